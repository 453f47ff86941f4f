import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from centroidal_mpc_b200 import synthetic
from centroidal_mpc_b200.device import BatchSolver
for name, B, sto, mode in (("solo12_trot", 4096, False, "B"), ("solo12_pace", 1024, False, "A"), ("solo12_bound", 4096, False, "B"), ("bolt", 8192, False, "B")):
    conf = synthetic.load_conf(name, N=100)
    solver = BatchSolver(synthetic.make_batch(conf, B, mode=mode, stochastic=sto))
    ref = None
    for a in (1.6, 1.72, 1.75, 1.78, 1.8):
        ov = dict(alpha=a)
        for _ in range(2): solver.solve(conf.scp_params, ov)
        torch.cuda.synchronize(); ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); solver.solve(conf.scp_params, ov); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        r, st = solver.results(), solver.stats()
        if ref is None: ref = r
        err = max(np.linalg.norm(r["U"][i] - ref["U"][i]) / np.linalg.norm(ref["U"][i]) for i in range(0, B, 61))
        print("%-12s B %5d sto %d alpha %.2f ms %6.2f failed %d nfac %.2f (max %d) pmm %.1f cert %.3f err %.1e" % (name, B, sto, a, np.median(ts), int((r["status"] != 0).sum()), st["n_factor"].mean(), st["n_factor"].max(), st["info"][:, 8].mean(), st["info"][:, 10].mean(), err), flush=True)
    solver.close()

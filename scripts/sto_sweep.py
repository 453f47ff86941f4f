import json, os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from centroidal_mpc_b200 import synthetic
from centroidal_mpc_b200.device import BatchSolver
for name in ("solo12_trot", "solo12_bound"):
    conf = synthetic.load_conf(name, N=100)
    solver = BatchSolver(synthetic.make_batch(conf, 4096, stochastic=True))
    ref = None
    for start in (20, 16, 12, 8):
        ov = dict(active_set_start=start, active_set_step=start)
        for _ in range(2): solver.solve(conf.scp_params, ov)
        torch.cuda.synchronize(); ts=[]
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); solver.solve(conf.scp_params, ov); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        r, st = solver.results(), solver.stats()
        if ref is None: ref = r
        err = max(np.linalg.norm(r["U"][i]-ref["U"][i])/np.linalg.norm(ref["U"][i]) for i in range(0,4096,64))
        print(name, "start", start, "ms %.2f" % np.median(ts), "failed", int((r["status"]!=0).sum()), "admm %.1f max %d nfac %.2f max %d pmm %.1f cert %.3f err %.1e" % (st["qp_iters"].mean(), st["qp_iters"].max(), st["n_factor"].mean(), st["n_factor"].max(), st["info"][:,8].mean(), st["info"][:,10].mean(), err), flush=True)
    solver.close()

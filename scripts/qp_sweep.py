"""Sweep of the QP settings that trade ADMM iterations against polish rounds, on the GPU (informational):
   python scripts/qp_sweep.py"""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from centroidal_mpc_b200 import synthetic
from centroidal_mpc_b200.device import BatchSolver

for name, B in (("solo12_trot", 4096), ("solo12_bound", 4096), ("bolt", 8192), ("solo12_pace", 4096)):
    conf = synthetic.load_conf(name, N=100)
    solver = BatchSolver(synthetic.make_batch(conf, B))
    ref = None
    grid = [dict()] + [dict(active_set_start=s, active_set_step=s) for s in (6, 7, 9, 10)] + \
           [dict(rho=r) for r in (1.0, 1.5, 3.0, 4.0)] + [dict(alpha=a) for a in (1.85,)] + \
           [dict(rho=3.0, active_set_start=6, active_set_step=6), dict(rho=1.5, active_set_start=10, active_set_step=10)]
    for ov in grid:
        for _ in range(2): solver.solve(conf.scp_params, ov or None)
        torch.cuda.synchronize(); ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); solver.solve(conf.scp_params, ov or None); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        r, st = solver.results(), solver.stats()
        if ref is None: ref = r
        idx = range(0, B, 97)
        err = max(np.linalg.norm(r["U"][i] - ref["U"][i]) / np.linalg.norm(ref["U"][i]) for i in idx)
        print("%-12s %-40s ms %6.2f failed %d admm %.1f nfac %.2f (max %d) pmm %.1f cert %.3f err %.1e" % (
            name, ov, np.median(ts), int((r["status"] != 0).sum()), st["qp_iters"].mean(), st["n_factor"].mean(), st["n_factor"].max(),
            st["info"][:, 8].mean(), st["info"][:, 10].mean(), err), flush=True)
    solver.close()

"""Timing sweep over cmpc_qp_settings on the headline workload (run on the GPU box): first polish
attempt (active_set_start), ADMM penalty (rho), relaxation (alpha).  Prints ms per batch, solver
statistics and the deviation from the default-settings solution."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from centroidal_mpc_b200 import synthetic
from centroidal_mpc_b200.device import BatchSolver

conf = synthetic.load_conf(os.environ.get("SWEEP_CONF", "solo12_trot"), N=100)
B = int(os.environ.get("SWEEP_BATCH", "4096"))
STOCH = os.environ.get("SWEEP_STOCH", "0") == "1"
solver = BatchSolver(synthetic.make_batch(conf, B, stochastic=STOCH))


def run(qp):
    for _ in range(2):
        solver.solve(conf.scp_params, qp)
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); solver.solve(conf.scp_params, qp); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    r, s = solver.results(), solver.stats()
    return min(ts), r, s


t0, r0, s0 = run(None)
print("default: %.2f ms  admm %.1f  nfac mean %.2f max %d  sweeps mean %.2f max %d" % (
    t0, s0["qp_iters"].mean(), s0["n_factor"].mean(), s0["n_factor"].max(), s0["info"][:, 8].mean(), s0["info"][:, 8].max()))
grid = []
for a in (4, 6, 8, 10, 12, 15, 30):
    grid.append(dict(active_set_start=a, active_set_step=a))
for rho in (0.5, 1.0, 4.0, 8.0):
    grid.append(dict(rho=rho))
for a, rho in ((10, 4.0), (10, 8.0), (8, 4.0), (12, 4.0), (6, 8.0)):
    grid.append(dict(active_set_start=a, active_set_step=a, rho=rho))
for al in (1.0, 1.8):
    grid.append(dict(alpha=al))
for rf in (1, 2, 5, 10):
    grid.append(dict(polish_refine_iter=rf))
for rf, a in ((5, 15), (10, 15), (10, 12), (10, 10)):
    grid.append(dict(polish_refine_iter=rf, active_set_start=a, active_set_step=a))
if os.environ.get("SWEEP_ONLY_REFINE", "0") == "1":
    grid = grid[-8:]
if STOCH:
    grid = [dict(polish_active_set_rounds=r) for r in (19, 39)]
    grid += [dict(active_set_start=a, active_set_step=a) for a in (40, 80)]
    grid += [dict(active_set_start=a, active_set_step=a, polish_active_set_rounds=29) for a in (20, 40)]
    grid += [dict(rho=r) for r in (0.5, 8.0)]
for qp in grid:
    t, r, s = run(qp)
    ex = np.linalg.norm(r["X"] - r0["X"]) / np.linalg.norm(r0["X"])
    eu = np.linalg.norm(r["U"] - r0["U"]) / np.linalg.norm(r0["U"])
    print("%-55s %.2f ms  failed %d  admm %.1f max %d  nfac mean %.2f max %d  sweeps mean %.2f max %d  attempts max %d  dX %.1e dU %.1e" % (
        qp, t, int((r["status"] != 0).sum()), s["qp_iters"].mean(), s["qp_iters"].max(), s["n_factor"].mean(), s["n_factor"].max(),
        s["info"][:, 8].mean(), s["info"][:, 8].max(), s["info"][:, 9].max(), ex, eu), flush=True)

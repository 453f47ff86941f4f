"""End-to-end host-entry timing: pinned (mapped) vs pageable result buffers (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from centroidal_mpc_b200 import synthetic
from centroidal_mpc_b200.device import BatchSolver
conf = synthetic.load_conf("solo12_trot", N=100)
B = 4096
batch = synthetic.make_batch(conf, B)
keep = []
for name in ("x_init", "x_final", "X_ref", "U_init", "contact_pos", "contact_active"):
    t = torch.from_numpy(getattr(batch, name)).pin_memory(); keep.append(t); setattr(batch, name, t.numpy())
solver = BatchSolver(batch)
def bufs(pin):
    mk = (lambda *a, **k: torch.zeros(*a, **k).pin_memory()) if pin else torch.zeros
    d = dict(X=mk((B, 101, 9), dtype=torch.float64), U=mk((B, 100, batch.nu), dtype=torch.float64), scp_iters=mk(B, dtype=torch.int32),
             status=mk(B, dtype=torch.int32), n_accepted=mk(B, dtype=torch.int32))
    keep.append(d)
    return {k: v.numpy() for k, v in d.items()}
for pin in (True, False):
    out = bufs(pin)
    for _ in range(2): solver.solve_host(conf.scp_params, out=out)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): solver.solve_host(conf.scp_params, out=out)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5 * 1e3
    print("pinned" if pin else "pageable", "%.2f ms per call" % dt, "status sum", int(out["status"].sum()))
for _ in range(2): solver.solve(conf.scp_params)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): solver.solve(conf.scp_params)
torch.cuda.synchronize(); print("device-resident %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))

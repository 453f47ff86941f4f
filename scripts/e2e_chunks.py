"""Host-entry (cmpc_solve_scp_host) timing for different chunk counts / result paths; one subprocess per setting
because the library reads CMPC_HOST_CHUNKS / CMPC_HOST_ZEROCOPY once.  python scripts/e2e_chunks.py"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time
sys.path.insert(0, %r)
import numpy as np, torch
from centroidal_mpc_b200 import synthetic
from centroidal_mpc_b200.device import BatchSolver
conf = synthetic.load_conf("solo12_trot", N=100)
B = int(os.environ.get("AB_BATCH", "4096"))
batch = synthetic.make_batch(conf, B)
keep = []
for name in ("x_init", "x_final", "X_ref", "U_init", "contact_pos", "contact_active"):
    t = torch.from_numpy(getattr(batch, name)).pin_memory(); keep.append(t); setattr(batch, name, t.numpy())
solver = BatchSolver(batch)
mk = lambda *a, **k: torch.zeros(*a, **k).pin_memory()
d = dict(X=mk((B, 101, 9), dtype=torch.float64), U=mk((B, 100, batch.nu), dtype=torch.float64), scp_iters=mk(B, dtype=torch.int32),
         status=mk(B, dtype=torch.int32), n_accepted=mk(B, dtype=torch.int32))
out = {k: v.numpy() for k, v in d.items()}
for _ in range(3): solver.solve_host(conf.scp_params, out=out)
torch.cuda.synchronize(); ts = []
for _ in range(7):
    t0 = time.perf_counter(); solver.solve_host(conf.scp_params, out=out); ts.append((time.perf_counter() - t0) * 1e3)
print("chunks=%%s zerocopy=%%s  median %%.2f ms  min %%.2f ms  (%%d k solves/s)  failed %%d" %% (os.environ.get("CMPC_HOST_CHUNKS"), os.environ.get("CMPC_HOST_ZEROCOPY"), float(np.median(ts)), min(ts), B / float(np.median(ts)), int((out["status"] != 0).sum())))
''' % ROOT
for zc in ("1", "0"):
    for ch in ("1", "2", "4", "8"):
        env = dict(os.environ, CMPC_HOST_CHUNKS=ch, CMPC_HOST_ZEROCOPY=zc)
        subprocess.run([sys.executable, "-c", CHILD], env=env)

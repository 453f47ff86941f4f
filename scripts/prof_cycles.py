"""Cycle breakdown per operation kind of one batched solve (profiling build of the library:
   scripts/variant.sh prof -DCMPC_PROFILE; CMPC_B200_LIB=.../variants/libcmpc_prof.so python scripts/prof_cycles.py [workload] [B])"""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from centroidal_mpc_b200 import synthetic, _lib as L
from centroidal_mpc_b200.device import BatchSolver

name = sys.argv[1] if len(sys.argv) > 1 else "solo12_trot"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
conf = synthetic.load_conf(name, N=100)
batch = synthetic.make_batch(conf, B)
if os.environ.get("CMPC_SAME"):      # every instance a copy of instance CMPC_SAME: all tiles run the same op sequence
    j = int(os.environ["CMPC_SAME"])
    for nm in ("x_init", "x_final", "X_ref", "U_init"):
        a = getattr(batch, nm)
        a[:] = a[j:j + 1]
solver = BatchSolver(batch)
lib = L.load()
out = (C.c_double * 32)()
for _ in range(2):
    solver.solve(conf.scp_params)
torch.cuda.synchronize()
lib.cmpc_debug_profile(out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); solver.solve(conf.scp_params); e1.record(); torch.cuda.synchronize()
rc = lib.cmpc_debug_profile(out)
v = np.array(list(out))
tiles = (B + 3) // 4
names = ["factor_admm", "sweep_admm", "build_as", "factor_pmm", "sweep_pmm", "rescale", "copy_sol", "eval", "write", "-",
         "  bwd of admm", "  bwd of pmm", "setup", "TILE", "copy wait", "n waits"]
print("rc", rc, "ms", e0.elapsed_time(e1), "tiles", tiles, "mean tile cycles %.0f (%.2f ms @1.965 GHz)" % (v[13] / tiles, v[13] / tiles / 1.965e6))
for n, x in zip(names, v):
    print("%-14s %12.0f cycles/tile  %5.1f%%" % (n, x / tiles, 100 * x / v[13]))
for n, x in zip(["bwd: acquire+peek", "bwd: kappa term", "bwd: phase 1 + sync", "bwd: phase 2 + sync", "bwd: release/issue/--k", "bwd: loop exit"], v[16:22]):
    print("%-24s %12.0f cycles/tile  %5.1f%%" % (n, x / tiles, 100 * x / v[13]))
for n, x in zip(["fac: phase 1 (Y rows)", "fac: phase 2 (build rows)", "fac: phase 3 (pivots)", "fac: output"], v[22:26]):
    print("%-24s %12.0f cycles/tile  %5.1f%%" % (n, x / tiles, 100 * x / v[13]))
print("wait per acquire %.0f cycles" % (v[14] / max(v[15], 1)))
st = solver.stats()
print("admm its mean", st["qp_iters"].mean(), "factor mean", st["n_factor"].mean(), "max", st["n_factor"].max(), "pmm sweeps mean", st["info"][:, 8].mean(), "max", st["info"][:, 8].max())

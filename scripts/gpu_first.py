"""First GPU contact: smoke, batched solve timing, statistics.  Run under gpurun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import ctypes as C
import __graft_entry__ as g
from centroidal_mpc_b200 import synthetic, _lib as L
from centroidal_mpc_b200.device import BatchSolver

g.smoke()
lib = L.load()
tf, ms = C.c_double(), C.c_double()
L.check(lib.cmpc_fp64_peak(C.byref(tf), C.byref(ms)))
print("fp64 peak TFLOP/s %.2f (%.2f ms)" % (tf.value, ms.value))
name = sys.argv[1] if len(sys.argv) > 1 else "solo12_trot"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
conf = synthetic.load_conf(name, N=100)
t0 = time.time(); batch = synthetic.make_batch(conf, B); print("make_batch %.1fs" % (time.time() - t0))
solver = BatchSolver(batch)
print("workspace MB", lib.cmpc_workspace_bytes(solver.handle) / 1e6)
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); solver.solve(conf.scp_params); e1.record(); torch.cuda.synchronize()
    dt = e0.elapsed_time(e1)
    print("rep", rep, "solve ms %.2f  solves/s %.0f" % (dt, B / dt * 1e3))
res = solver.results(); st = solver.stats()
print("status counts", np.bincount(res["status"]), "scp_iters", np.bincount(res["scp_iters"]), "accepted", np.bincount(res["n_accepted"]))
print("qp_iters mean %.1f min %d max %d; n_factor mean %.2f; polished %.3f" % (st["qp_iters"].mean(), st["qp_iters"].min(), st["qp_iters"].max(), st["n_factor"].mean(), st["info"][:, 7].mean()))
print("hist qp_iters", np.bincount(st["qp_iters"] // 25))
# parity of a few instances against the host build of the same source
import emu_binding as E
from centroidal_mpc_b200.batch import ProblemBatch
sub = synthetic.make_batch(conf, 4)
emu = E.solve_scp(sub, conf.scp_params)
for b in range(4):
    print(b, "gpu vs emu relerr X %.2e U %.2e; qp iters gpu %d emu %d" % (
        np.linalg.norm(res["X"][b] - emu["X"][b]) / np.linalg.norm(emu["X"][b]),
        np.linalg.norm(res["U"][b] - emu["U"][b]) / np.linalg.norm(emu["U"][b]), st["qp_iters"][b], emu["qp_iters"][b]))
t0 = time.time(); out = solver.solve_host(conf.scp_params); print("solve_host wall ms %.2f" % ((time.time() - t0) * 1e3))
print("host path equal:", np.array_equal(out["X"], res["X"]))

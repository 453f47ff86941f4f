"""GPU vs the host build of the same source on the WHOLE headline batch (run on the GPU box):
equal status / SCP iterations / ADMM iterations / factorisation counts per instance, and the
largest norm-wise difference of the trajectories."""
import os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import numpy as np
import emu_binding as E
from centroidal_mpc_b200 import synthetic
from centroidal_mpc_b200.src.scp_solver import solve_scp_batched

name, N, B = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ("solo12_trot", 100, 4096)
conf = synthetic.load_conf(name, N=N)
batch = synthetic.make_batch(conf, B)
out = solve_scp_batched(batch, conf.scp_params, return_stats=True)
emu = E.solve_scp(batch, conf.scp_params)
for k in ("status", "scp_iters", "n_accepted", "qp_iters", "n_factor"):
    print(k, "equal:", bool(np.array_equal(out[k], emu[k])))
ex = max(np.linalg.norm(out["X"][b] - emu["X"][b]) / np.linalg.norm(emu["X"][b]) for b in range(B))
eu = max(np.linalg.norm(out["U"][b] - emu["U"][b]) / np.linalg.norm(emu["U"][b]) for b in range(B))
print("bitwise equal X:", bool(np.array_equal(out["X"], emu["X"])), "U:", bool(np.array_equal(out["U"], emu["U"])))
print("max relerr X %.3e U %.3e over %d instances of %s N=%d" % (ex, eu, B, name, N))

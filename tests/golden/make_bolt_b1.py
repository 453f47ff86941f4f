"""tests/golden/bolt_b1_highs.npz: the QP of bolt, N = 40, instance 1 solved by scipy's HiGHS active-set QP solver
(independent of the oracle's OSQP restatement, which needs 96 275 iterations at eps 1e-9 for this QP and stops at
OSQP's max_iter = 4000 with the reference's settings, i.e. the reference returns False here).
Run from the repository root:  python tests/golden/make_bolt_b1.py"""
import os
import sys

import numpy as np
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from centroidal_mpc_b200 import synthetic                                   # noqa: E402
from centroidal_mpc_b200.src.centroidal_model import Centroidal_model      # noqa: E402
from oracle import qp_build, scp                                             # noqa: E402
from test_oracle import _highs_qp                                            # noqa: E402

conf = synthetic.load_conf("bolt", N=40)
m = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, 1))
prob = m.problem_arrays()
sp = conf.scp_params
P, q, A, l, u, blocks, td = scp.build_qp(prob, sp["trust_region_radius0"], sp["omega0"])
dyn = slice(blocks["dynamics"], blocks["final"])
lh, uh = l.copy(), u.copy()
lh[dyn] = uh[dyn] = 0.5 * (l[dyn] + u[dyn])          # the +-1e-12 band as an equality
n, N = P.shape[0], prob["N"]
reg = np.zeros(n)
reg[-(2 * N + 1):] = 1e-8                              # the slacks have no curvature
z, status = _highs_qp(sparse.csc_matrix(P + sparse.diags(reg)), q, A, lh, uh)
assert "Optimal" in status, status
X, U = qp_build.unpack(prob, z)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bolt_b1_highs.npz"), X=X, U=U, status=status)
print("saved", status, X.shape, U.shape)

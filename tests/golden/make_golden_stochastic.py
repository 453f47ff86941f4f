"""Generates tests/golden/stoch*.npz from the CPU oracle: LQR gains, covariances, friction back-offs
and the stochastic-mode SCP solution.  Like make_golden.py these
fixtures pin the ORACLE (the reference cannot run here: parity unpinned).
Run from the repository root:  python tests/golden/make_golden_stochastic.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from centroidal_mpc_b200 import synthetic                                   # noqa: E402
from centroidal_mpc_b200.src.centroidal_model import Centroidal_model      # noqa: E402
from oracle import dynamics, qp_build, scp                                   # noqa: E402

CASES = [("solo12_trot", 40), ("solo12_bound", 40), ("solo12_pace", 30), ("bolt", 40)]


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for idx, (name, N) in enumerate(CASES):
        conf = synthetic.load_conf(name, N=N)
        m = Centroidal_model(conf, STOCHASTIC_OCP=True, centroidal_traj=synthetic.reference_trajectory(conf, 0))
        prob = m.problem_arrays()
        gains, covs = dynamics.lqr_gains_covs(prob["X_ref"], prob["U_init"], prob, m._Q, m._R, m._Cov_w, m._Cov_eta)
        ub, xi = qp_build.friction_backoffs(prob, gains, covs, m._beta_u)
        sprob = dict(prob, friction_ub=ub)
        sol = scp.solve_scp(sprob, conf.scp_params)
        tight = scp.solve_scp(sprob, conf.scp_params, osqp_settings=dict(eps_abs=1e-9, eps_rel=1e-9, max_iter=40000,
                                                                        polish_refine_iter=30))
        assert sol is not False and tight is not False and sol["state"] and tight["state"]
        np.savez_compressed(os.path.join(here, "stoch%d_%s_N%d_b0.npz" % (idx, name, N)), name=name, N=N, b=0, xi=xi,
                            gains=gains, covs=covs, friction_ub=ub, iterations=sol["iterations"],
                            X=sol["state"][-1], U=sol["control"][-1], X_tight=tight["state"][-1],
                            U_tight=tight["control"][-1])
        print(idx, name, N, "iters", sol["iterations"], "ub min", ub.min())


if __name__ == "__main__":
    main()

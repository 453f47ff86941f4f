"""Generates tests/golden/*.npz from the CPU oracle (oracle/scp.py).

The reference itself cannot run in this environment and ships no golden vectors (SURVEY.md
section 8c: parity unpinned), so these fixtures pin the ORACLE, not the reference: they guard
the oracle against regressions and give the GPU tests a target that needs no oracle run.
Run from the repository root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from centroidal_mpc_b200 import synthetic                                   # noqa: E402
from centroidal_mpc_b200.src.centroidal_model import Centroidal_model      # noqa: E402
from oracle import scp                                                       # noqa: E402

CASES = [("solo12_trot", 40, {}), ("solo12_pace", 30, {}), ("solo12_bound", 40, {}), ("bolt", 40, {}),
         ("solo12_trot", 100, {}),
         # forced branches of the trust-region state machine
         ("solo12_trot", 40, dict(trust_region_radius0=1.0, max_iterations=4)),
         ("solo12_trot", 40, dict(rho1=1e-7, max_iterations=3)),
         ("solo12_trot", 40, dict(trust_region_radius0=0.05, max_iterations=2))]


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for idx, (name, N, upd) in enumerate(CASES):
        conf = synthetic.load_conf(name, N=N)
        sp = dict(conf.scp_params)
        sp.update(upd)
        for b in range(2):
            model = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, b))
            prob = model.problem_arrays()
            log = []
            sol = scp.solve_scp(prob, sp, log=log)
            ok = sol is not False
            X = sol["state"][-1] if ok and sol["state"] else np.zeros((9, N + 1))
            U = sol["control"][-1] if ok and sol["control"] else np.zeros((prob["U_init"].shape[0], N))
            # the same QP sequence solved tightly (eps 1e-9, 30 polish refinements): OSQP's own
            # answer at its default settings is only accurate to ~1e-7..4e-6 (bolt), this one to ~1e-9
            Xt, Ut = X, U
            if ok and sol["state"]:
                tight = scp.solve_scp(prob, sp, osqp_settings=dict(eps_abs=1e-9, eps_rel=1e-9, max_iter=40000,
                                                                  polish_refine_iter=30))
                if tight and tight["state"]:
                    Xt, Ut = tight["state"][-1], tight["control"][-1]
            np.savez_compressed(
                os.path.join(here, "case%d_%s_N%d_b%d.npz" % (idx, name, N, b)),
                name=name, N=N, b=b, scp_keys=np.array(sorted(upd)), scp_vals=np.array([upd[k] for k in sorted(upd)], dtype=float),
                returned_false=not ok, iterations=(sol["iterations"] if ok else len(log)),
                n_accepted=(len(sol["state"]) if ok else 0), X=X, U=U, X_tight=Xt, U_tight=Ut,
                verdicts=np.array([e.get("verdict", "qp_failed") for e in log]),
                qp_status=np.array([e["status"] for e in log]),
                snorm=np.array([e.get("snorm", np.nan) for e in log]))
            print(idx, name, N, b, "iters", len(log), [e.get("verdict") for e in log])


if __name__ == "__main__":
    main()

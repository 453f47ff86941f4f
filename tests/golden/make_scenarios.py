"""Generates tests/golden/scen_*.npz and tests/golden/n100_*.npz from the CPU oracle (oracle/scp.py, OSQP
restatement solved tightly: eps 1e-9, 30 polish refinements).  Like make_golden.py these fixtures pin the
ORACLE (the reference cannot run here: parity unpinned, oracle/__init__.py).

scen_*: synthetic cases for the branches of the trust-region loop (/root/reference/src/scp_solver.py:146-176)
that the shipped configurations never take.  The base problem is a trot window whose
warm start is (close to) its own solution, plus a one-knot spike of 0.1 on the three angular-momentum
references: the solution cannot follow the spike, so |kappa_k - kappa_bar_k|_1 = 0.247 at that knot while
sigma_max(X - X_bar) = 0.18 -- an L1 trust region with a radius in between BINDS on an ACCEPTED iterate.
  scen0  radius 0.22, weight 100: accepted at the first iteration, the kappa rows bind (slack active)
  scen1  radius 0.24, weight 1e4: accepted, the rows bind on the surface of the L1 ball
  scen2  rejected by the accuracy ratio (rho1 between the two ratios), radius updated, accepted with a
         DIFFERENT QP solution (beta_fail = 2 so that the second radius no longer binds)
  scen3  rejected by the trust test at weight 2500, weight updated (gamma_fail = 0.04), accepted at weight 100
  scen4  zero control warm start: the reference's convergence() is 0/0 = NaN, the loop runs to max_iterations
         and accepts every iterate
  scen5  free fall with a final state off the ballistic path: the QP is infeasible, the reference returns False
n100_*: the other BASELINE configurations at the benchmark horizon N = 100 (pace with perturbed initial
states, bound, bolt): tightly solved oracle trajectories of a sample of instances.

Run from the repository root:  python tests/golden/make_scenarios.py   (about 20 minutes)"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from centroidal_mpc_b200 import synthetic                                   # noqa: E402
from centroidal_mpc_b200.src.centroidal_model import Centroidal_model      # noqa: E402
from oracle import scp                                                       # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TIGHT = dict(eps_abs=1e-9, eps_rel=1e-9, max_iter=40000, polish_refine_iter=30)


def set_warm_start(model, X, U):
    model._init_trajectories = dict(state=np.array(X, dtype=np.float64), control=np.array(U, dtype=np.float64))
    model._x_init = model._init_trajectories["state"][:, 0].copy()
    model._x_final = model._init_trajectories["state"][:, -1].copy()


def spike_model(N=30, b=3, delta=0.1, k0=12):
    conf = synthetic.load_conf("solo12_trot", N=N)
    m = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, b))
    r = scp.solve_scp(m.problem_arrays(), conf.scp_params, osqp_settings=TIGHT)
    X, U = r["state"][-1].copy(), r["control"][-1].copy()
    for _ in range(3):      # towards the fixed point "the warm start is its own solution"
        set_warm_start(m, X, U)
        r = scp.solve_scp(m.problem_arrays(), conf.scp_params, osqp_settings=TIGHT)
        X, U = r["state"][-1].copy(), r["control"][-1].copy()
    Xs = X.copy()
    Xs[6:, k0] += delta
    set_warm_start(m, Xs, U)
    return conf, m


def save(name, conf_name, N, model, sp_upd, free_fall=False):
    conf = synthetic.load_conf(conf_name, N=N)
    sp = dict(conf.scp_params)
    sp.update(sp_upd)
    prob = model.problem_arrays()
    log = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sol = scp.solve_scp(prob, sp, osqp_settings=TIGHT, log=log)
    ok = sol is not False
    nacc = len(sol["state"]) if ok else 0
    X = sol["state"][-1] if nacc else np.zeros_like(prob["X_ref"])
    U = sol["control"][-1] if nacc else np.zeros_like(prob["U_init"])
    dk = np.abs(X[6:] - prob["X_ref"][6:]).sum(axis=0).max() if nacc else 0.0
    np.savez_compressed(os.path.join(HERE, name + ".npz"), name=conf_name, N=N, X_ref=prob["X_ref"], U_init=prob["U_init"],
                        free_fall=free_fall, scp_keys=np.array(sorted(sp_upd)),
                        scp_vals=np.array([sp_upd[k] for k in sorted(sp_upd)], dtype=float), returned_false=not ok,
                        iterations=(sol["iterations"] if ok else len(log)), n_accepted=nacc, X=X, U=U,
                        qp_status=np.array([e["status"] for e in log]), snorm=np.array([e.get("snorm", np.nan) for e in log]),
                        acc=np.array([e.get("rho", np.nan) if e.get("rho") is not None else np.nan for e in log]),
                        radius=np.array([e["radius"] for e in log]), weight=np.array([e["weight"] for e in log]),
                        max_dkappa_l1=dk)
    print(name, "iterations", len(log), "accepted", nacc, "returned_false", not ok,
          [(e["status"], round(float(e.get("snorm", -1)), 5), e.get("rho"), e["radius"], e["weight"]) for e in log], "max|dk|1", dk)


def scenarios():
    conf, m = spike_model()
    save("scen0_tr_binding_accepted", "solo12_trot", 30, m, dict(trust_region_radius0=0.22, max_iterations=4))
    save("scen1_tr_surface_accepted", "solo12_trot", 30, m, dict(trust_region_radius0=0.24, omega0=1e4, max_iterations=4))
    save("scen2_accuracy_reject_then_accept", "solo12_trot", 30, m,
         dict(trust_region_radius0=0.22, rho1=3.114e-11, beta_fail=2.0, max_iterations=4))
    save("scen3_weight_reject_then_accept", "solo12_trot", 30, m,
         dict(trust_region_radius0=0.188, omega0=2500.0, gamma_fail=0.04, max_iterations=4))
    conf = synthetic.load_conf("solo12_trot", N=20)
    m = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, 1))
    m._init_trajectories["control"][:] = 0.0
    save("scen4_nan_convergence", "solo12_trot", 20, m, dict(max_iterations=3))
    conf = synthetic.load_conf("solo12_trot", N=10)
    m = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, 1))
    m._contact_data["contacts_logic"][:] = 0
    m._init_trajectories["control"][:] = 0.0
    save("scen5_infeasible_free_fall", "solo12_trot", 10, m, {}, free_fall=True)


N100 = (("solo12_pace", "A", 1024, [int(i) for i in np.linspace(0, 1023, 32)], 100),
        ("solo12_bound", "B", 4096, [int(i) for i in np.linspace(0, 4095, 8)], 100),
        ("bolt", "B", 8192, [int(i) for i in np.linspace(0, 8191, 8)], 100))
# BASELINE configuration 5 (talos, CoP / wrench contact model): the smallest batch of its sweep at the benchmark
# horizon, and a short-horizon case for the CPU suite (host build of the wrench solver)
TALOS = (("talos", "B", 256, [int(i) for i in np.linspace(0, 255, 8)], 100),
         ("talos", "B", 8, [0, 3, 7], 30))


def n100(table=N100):
    for conf_name, mode, B, ids, N in table:
        conf = synthetic.load_conf(conf_name, N=N)
        Xs, Us, its, ok = [], [], [], []
        for b in ids:
            if mode == "A":
                m = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, 0, mode="A"))
                m._x_init = synthetic.perturbed_x_init(conf, b)
            else:
                m = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, b))
            sol = scp.solve_scp(m.problem_arrays(), conf.scp_params, osqp_settings=TIGHT)
            good = sol is not False and len(sol["state"]) > 0
            ok.append(good)
            its.append(sol["iterations"] if sol is not False else -1)
            Xs.append(sol["state"][-1] if good else np.zeros((9, N + 1)))
            Us.append(sol["control"][-1] if good else np.zeros((conf.n_u, N)))
            print(conf_name, b, "ok", good, "iterations", its[-1], flush=True)
        np.savez_compressed(os.path.join(HERE, "n%d_%s_mode%s.npz" % (N, conf_name, mode)), name=conf_name, mode=mode, batch=B,
                            ids=np.array(ids), ok=np.array(ok), iterations=np.array(its), X=np.array(Xs), U=np.array(Us))


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] == "scen":
        scenarios()
    if len(sys.argv) < 2 or sys.argv[1] == "n100":
        n100()
    if len(sys.argv) < 2 or sys.argv[1] == "talos":
        n100(TALOS)

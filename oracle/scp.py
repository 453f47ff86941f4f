"""SCP trust-region outer loop (oracle).  Follows /root/reference/src/scp_solver.py:

  solve_scp               :118-179   -> solve_scp()
  solve_subproblem        :59-68     -> osqp_restatement.solve(eps 1e-7, polish)
  compute_model_accuracy  :71-87     -> model_accuracy()
  convergence             :51-56     -> convergence()  (spectral norms)
  trust test              :151       -> np.linalg.norm(., 2) on a 9x(N+1) matrix

Reproduces the reference faithfully, including that the linearisation point is
never moved (:129-130).

PARITY UNPINNED: see oracle/__init__.py.
"""
import numpy as np

from . import dynamics, osqp_restatement, qp_build


def convergence(X_curr, U_curr, X_prev, U_prev):
    return (np.linalg.norm(U_curr - U_prev, 2) / np.linalg.norm(U_curr, 2)
            + np.linalg.norm(X_curr - X_prev, 2) / np.linalg.norm(X_curr, 2))


def model_accuracy(prob, X, U, traj_data):
    """rho = sum_k ||(f(x_k,u_k) - lin_k)[6:9]||^2 / sum_k ||lin_k||^2, scp_solver.py:71-87."""
    Fnl = dynamics.rollout(X, U, prob)
    F, Ax, Bu = traj_data["dynamics"], traj_data["f_x"], traj_data["f_u"]
    Xp, Up = prob["X_ref"], prob["U_init"]
    num = den = 0.0
    for k in range(U.shape[1]):
        lin = F[:, k] + Ax[k] @ (X[:, k] - Xp[:, k]) + Bu[k] @ (U[:, k] - Up[:, k])
        err = Fnl[6:, k] - lin[6:]
        num += err @ err
        den += lin @ lin
    return num / den


def build_qp(prob, radius, weight, emulate_jax_fp32=False):
    td = dynamics.trajectory_data(prob["X_ref"], prob["U_init"], prob, emulate_jax_fp32)
    P, q = qp_build.build_cost(prob)
    A, l, u, blocks = qp_build.build_constraints(prob, td, radius, weight, emulate_jax_fp32)
    return P, q, A, l, u, blocks, td


def solve_scp(prob, scp_params, emulate_jax_fp32=False, osqp_settings=None, log=None):
    """Returns dict(state=[X..], control=[U..], gains=[], covs=[]) like the reference,
    or False when a QP is not 'solved'.  ``log`` (a list) receives one dict per
    SCP iteration for the tests."""
    all_solution = dict(state=[], control=[], gains=[], covs=[])
    rho0, rho1 = scp_params["rho0"], scp_params["rho1"]
    omega_max = scp_params["omega_max"]
    beta_succ, beta_fail = scp_params["beta_succ"], scp_params["beta_fail"]
    max_iter = scp_params["max_iterations"]
    thresh = scp_params["convergence_threshold"]
    gamma_fail = scp_params["gamma_fail"]
    weight = float(scp_params["omega0"])
    radius = float(scp_params["trust_region_radius0"])
    settings = dict(eps_abs=1e-7, eps_rel=1e-7, polish=True)
    settings.update(osqp_settings or {})
    Xp, Up = prob["X_ref"], prob["U_init"]
    success = False
    it = 0
    # convergence(traj_tuple, prev_traj_dict) compares an object with itself (:129-134)
    while it < max_iter and weight < omega_max and \
            not (it != 0 and success and convergence(Xp, Up, Xp, Up) < thresh):
        success = False
        P, q, A, l, u, blocks, td = build_qp(prob, radius, weight, emulate_jax_fp32)
        res = osqp_restatement.solve(P, q, A, l, u, **settings)
        entry = dict(it=it, radius=radius, weight=weight, status=res.status, qp_iter=res.iter,
                     n_fact=res.n_fact, polished=res.polished)
        if res.status != "solved":
            if log is not None:
                log.append(entry)
            return False
        X, U = qp_build.unpack(prob, res.x)
        snorm = np.linalg.norm(X - Xp, 2)
        entry["snorm"] = snorm
        if snorm < radius:
            rho = model_accuracy(prob, X, U, td)
            entry["rho"] = rho
            if rho > rho1:
                radius *= beta_fail
                entry["verdict"] = "inaccurate"
            else:
                all_solution["state"].append(X)
                all_solution["control"].append(U)
                all_solution["gains"].append(None)
                all_solution["covs"].append(None)
                success = True
                entry["verdict"] = "accepted"
                if rho < rho0:
                    radius = min(beta_succ * radius, float(scp_params["trust_region_radius0"]))
        else:
            weight *= gamma_fail
            entry["verdict"] = "outside"
        if log is not None:
            log.append(entry)
        it += 1
    all_solution["iterations"] = it
    return all_solution

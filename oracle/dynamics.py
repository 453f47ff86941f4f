"""Centroidal dynamics, closed-form Jacobians and roll-outs (oracle, float64).

Follows /root/reference/src/centroidal_model.py:
  * integrate_model_one_step          :189-212  -> step()
  * jacfwd of that w.r.t. x, u, p     :229-232  -> jacobians()  (closed form)
  * integrate_dynamics_trajectory     :243-255  -> rollout()
  * compute_trajectory_data (f, A, B) :257-291  -> trajectory_data()

Problem data layout (all float64 unless noted), for one MPC instance:
  X        (9, N+1)   states  (com, lin_mom, ang_mom)
  U        (nu, N)    controls, nu = nc*npc; point contact npc=3 (fx,fy,fz),
                      TALOS npc=6 (cop_x,cop_y,fx,fy,fz,tau_z)
  contact_pos    (N, nc, 3)     zeros when inactive (:142-145)
  contact_R      (N, nc, 3, 3)  zeros when inactive
  contact_active (N, nc)        0/1
"""
import numpy as np


def skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def step(x, u, p, a, R, m, g, dt, robot="solo12"):
    """One explicit-Euler step, centroidal_model.py:189-212."""
    nc = len(a)
    npc = u.shape[0] // nc
    fdot = np.zeros(9)
    fdot[0:3] = x[3:6] / m
    fdot[5] = m * g
    for i in range(nc):
        ui = u[npc * i:npc * (i + 1)]
        p_com = p[i] - x[0:3]
        if robot == "TALOS":
            f = ui[2:5]
            lin = a[i] * f
            ang = a[i] * (np.cross(p_com, f) + np.cross(R[i][:, 0:2] @ ui[0:2], f)
                          + R[i][:, 2] * ui[5])
        else:
            lin = a[i] * ui
            ang = a[i] * np.cross(p_com, ui)
        fdot[3:6] += lin
        fdot[6:9] += ang
    return x + fdot * dt


def jacobians(x, u, p, a, R, m, g, dt, robot="solo12"):
    """A = df/dx (9x9), B = df/du (9xnu), C = df/dp (9x3nc); SURVEY.md A.3.

    The reference obtains these from jax.jacfwd (centroidal_model.py:230-232);
    the closed form below is checked against finite differences in the tests.
    """
    nc = len(a)
    npc = u.shape[0] // nc
    A = np.eye(9)
    A[0:3, 3:6] = dt / m * np.eye(3)
    B = np.zeros((9, u.shape[0]))
    C = np.zeros((9, 3 * nc))
    for i in range(nc):
        if not a[i]:
            continue
        ui = u[npc * i:npc * (i + 1)]
        if robot == "TALOS":
            f = ui[2:5]
            arm = p[i] - x[0:3] + R[i][:, 0:2] @ ui[0:2]
            B[6:9, npc * i:npc * i + 2] = -dt * skew(f) @ R[i][:, 0:2]
            B[3:6, npc * i + 2:npc * i + 5] = dt * np.eye(3)
            B[6:9, npc * i + 2:npc * i + 5] = dt * skew(arm)
            B[6:9, npc * i + 5] = dt * R[i][:, 2]
        else:
            f = ui
            B[3:6, 3 * i:3 * i + 3] = dt * np.eye(3)
            B[6:9, 3 * i:3 * i + 3] = dt * skew(p[i] - x[0:3])
        # d/dc of (p - c) x f = +[f]x ; d/dp = -[f]x
        A[6:9, 0:3] += dt * skew(f)
        C[6:9, 3 * i:3 * i + 3] = -dt * skew(f)
    return A, B, C


def rollout(X, U, prob):
    """One-step nonlinear predictions f(x_k,u_k), k<N (centroidal_model.py:243-255).

    The reference loops N+1 times with clamped indices; the extra column is
    never read (scp_solver.py:82-86), so only N columns are produced here.
    """
    N = U.shape[1]
    F = np.zeros((9, N))
    for k in range(N):
        F[:, k] = step(X[:, k], U[:, k], prob["contact_pos"][k], prob["contact_active"][k],
                       prob["contact_R"][k], prob["m"], prob["g"], prob["dt"], prob["robot"])
    return F


def trajectory_data(X, U, prob, emulate_jax_fp32=False):
    """f, A, B along (X,U) — the part of compute_trajectory_data (:257-291) the
    nominal SCP consumes.  ``emulate_jax_fp32`` rounds inputs and outputs to
    float32 the way default JAX would."""
    N = U.shape[1]
    nu = U.shape[0]
    if emulate_jax_fp32:
        X = X.astype(np.float32).astype(np.float64)
        U = U.astype(np.float32).astype(np.float64)
    F = np.zeros((9, N))
    Ax = np.zeros((N, 9, 9))
    Bu = np.zeros((N, 9, nu))
    Cw = np.zeros((N, 9, 3 * prob["contact_active"].shape[1]))
    for k in range(N):
        args = (X[:, k], U[:, k], prob["contact_pos"][k], prob["contact_active"][k],
                prob["contact_R"][k], prob["m"], prob["g"], prob["dt"], prob["robot"])
        F[:, k] = step(*args)
        Ax[k], Bu[k], Cw[k] = jacobians(*args)
    if emulate_jax_fp32:
        F, Ax, Bu, Cw = [v.astype(np.float32).astype(np.float64) for v in (F, Ax, Bu, Cw)]
    return dict(dynamics=F, f_x=Ax, f_u=Bu, f_w=Cw)


def lqr_feedback_gain(A, B, Q, R, niter=2):
    """compute_lqr_feedback_gains, centroidal_model.py:215-227: P = Q, `niter` Riccati steps,
    K = -(R + B'PB)^-1 B'PA."""
    P = Q
    for _ in range(niter):
        AtP = A.T @ P
        AtPB = AtP @ B
        P = (Q + AtP @ A) - AtPB @ np.linalg.solve(R + B.T @ P @ B, AtPB.T)
    return -np.linalg.solve(R + B.T @ P @ B, B.T @ P @ A)


def lqr_gains_covs(X, U, prob, Q, R, cov_w, cov_eta):
    """LQR_gains (N,nu,9) and Covs (N+1,9,9) of compute_trajectory_data
    (centroidal_model.py:233-238,284-285): Covs[0] = 0,
    Covs[k+1] = [A B] [[S, SK'],[KS, KSK']] [A B]' + C cov_w C' + cov_eta."""
    N = U.shape[1]
    nu = U.shape[0]
    gains = np.zeros((N, nu, 9))
    covs = np.zeros((N + 1, 9, 9))
    for k in range(N):
        A, B, C = jacobians(X[:, k], U[:, k], prob["contact_pos"][k], prob["contact_active"][k],
                            prob["contact_R"][k], prob["m"], prob["g"], prob["dt"], prob["robot"])
        K = lqr_feedback_gain(A, B, Q, R)
        S = covs[k]
        SKt = S @ K.T
        AB = np.hstack([A, B])
        Sxu = np.vstack([np.hstack([S, SKt]), np.hstack([SKt.T, K @ SKt])])
        gains[k] = K
        covs[k + 1] = AB @ Sxu @ AB.T + C @ cov_w @ C.T + cov_eta
    return gains, covs

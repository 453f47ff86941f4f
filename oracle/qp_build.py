"""QP assembly in the reference's own variable and row order (oracle, float64).

  z = [x_0..x_N | u_0..u_{N-1} | t_0..t_N | (N unused control slacks)]
      (/root/reference/src/optimizer.py:15-35,47-74,104-133)

  cost   : src/cost.py:9-39, summed as in src/scp_solver.py:10-26
  rows   : src/constraints.py:12-17 (initial), :20-50 (dynamics), :104-109 (final),
           :111-145 (CoP box, TALOS), :153-185,:215-217 (friction pyramid, nominal),
           :260-293 (L1 trust region on angular momentum), stacked in the order of
           src/scp_solver.py:28-48.
  pyramid: src/utils.py:9-16

Everything is built sparse (the reference builds dense and converts; the
matrices are identical).
"""
import numpy as np
from scipy import sparse


def dims(prob):
    N = int(prob["N"])
    nx = 9
    nu = int(prob["U_init"].shape[0])
    nc = int(prob["contact_active"].shape[1])
    nz = nx * (N + 1) + nu * N + (N + 1) + N       # centroidal_model.py:25-26
    return N, nx, nu, nc, nz


def idx_x(prob, k, i=0):
    return k * 9 + i


def idx_u(prob, k, j=0):
    N, nx, nu, nc, nz = dims(prob)
    return nx * (N + 1) + k * nu + j


def idx_t(prob, k):
    N, nx, nu, nc, nz = dims(prob)
    return nx * (N + 1) + nu * N + k


def penum_mat():
    """Slack_optimizer sign-enumeration matrix, optimizer.py:104-112."""
    S = np.zeros((8, 3))
    for i in range(3):
        S[:, i] = [(-1.0) ** (j // (2 ** i)) for j in range(8)]
    return S


def friction_pyramid(mu):
    """utils.py:9-16 (5x3; only the first four rows are ever written)."""
    ml = mu / np.sqrt(2.0)
    return np.array([[1.0, 0.0, -ml], [-1.0, 0.0, -ml], [0.0, 1.0, -ml],
                     [0.0, -1.0, -ml], [0.0, 0.0, -1.0]])


def build_cost(prob):
    """P (csc, nz x nz) and q; cost.py:9-39 + scp_solver.py:10-26."""
    N, nx, nu, nc, nz = dims(prob)
    Wx = np.asarray(prob["state_cost_weights"], dtype=np.float64)
    Wu = np.asarray(prob["control_cost_weights"], dtype=np.float64)
    P = sparse.block_diag([sparse.kron(sparse.eye(N + 1), Wx), sparse.kron(sparse.eye(N), Wu),
                           sparse.csc_matrix((N + 1, N + 1)), sparse.csc_matrix((N, N))], format="csc")
    q = np.zeros(nz)
    q[idx_t(prob, 0):idx_t(prob, 0) + N + 1] = 1.0                 # cost.py:34-39
    # tracking gradient only for solo12-type models with DYNAMICS_FIRST False
    # (scp_solver.py:13-20); TALOS gets none.
    if prob["robot"] != "TALOS" and not prob.get("DYNAMICS_FIRST", False):
        Xref = prob["X_ref"]
        for k in range(N + 1):
            q[9 * k:9 * k + 9] += -Wx @ Xref[:, k]                 # cost.py:21-29
    return P, q


def build_constraints(prob, traj_data, radius, weight, emulate_jax_fp32=False):
    """A (csc, m x nz), l, u in the reference row order; returns also a dict of
    row-block offsets."""
    N, nx, nu, nc, nz = dims(prob)
    npc = nu // nc
    X, U = prob["X_ref"], prob["U_init"]
    if emulate_jax_fp32:
        X32 = X.astype(np.float32).astype(np.float64)
        U32 = U.astype(np.float32).astype(np.float64)
    else:
        X32, U32 = X, U
    rows, cols, vals = [], [], []
    lo, up = [], []
    blocks = {}

    def add(r, c, v):
        if v != 0.0:
            rows.append(r); cols.append(c); vals.append(v)

    r0 = 0
    # 1. initial (constraints.py:12-17)
    blocks["initial"] = r0
    xi = np.asarray(prob["x_init"], dtype=np.float64)
    xf = np.asarray(prob["x_final"], dtype=np.float64)
    if emulate_jax_fp32:
        xi = xi.astype(np.float32).astype(np.float64)
        xf = xf.astype(np.float32).astype(np.float64)
    for i in range(9):
        add(r0 + i, i, 1.0)
    lo += list(xi); up += list(xi)
    r0 += 9
    # 2. dynamics (constraints.py:20-50)
    blocks["dynamics"] = r0
    F, Ax, Bu = traj_data["dynamics"], traj_data["f_x"], traj_data["f_u"]
    for k in range(N):
        Ak, Bk = Ax[k], Bu[k]
        for i in range(9):
            for j in range(9):
                if Ak[i, j] != 0.0:
                    add(r0 + i, idx_x(prob, k, j), Ak[i, j])
            for j in range(nu):
                if Bk[i, j] != 0.0:
                    add(r0 + i, idx_u(prob, k, j), Bk[i, j])
            add(r0 + i, idx_x(prob, k + 1, i), -1.0)
        lin = Ak @ X32[:, k] + Bk @ U32[:, k] - F[:, k]
        if emulate_jax_fp32:
            lo_k = (lin.astype(np.float32) - np.float32(1e-12)).astype(np.float64)
            up_k = (lin.astype(np.float32) + np.float32(1e-12)).astype(np.float64)
        else:
            lo_k, up_k = lin - 1e-12, lin + 1e-12
        lo += list(lo_k); up += list(up_k)
        r0 += 9
    # 3. final (constraints.py:104-109)
    blocks["final"] = r0
    for i in range(9):
        add(r0 + i, idx_x(prob, N, i), 1.0)
    lo += list(xf); up += list(xf)
    r0 += 9
    # 3b. CoP box, TALOS only (constraints.py:111-145); x rows then y rows per contact
    if prob["robot"] == "TALOS":
        blocks["cop"] = r0
        fr = prob["foot_range"]  # dict x:(lxp,lxn) y:(lyp,lyn)
        for c in range(nc):
            for ax, key in enumerate(("x", "y")):
                for k in range(N):
                    if prob["contact_active"][k, c]:
                        add(r0 + k, idx_u(prob, k, npc * c + ax), 1.0)
                        lo.append(-fr[key][1]); up.append(fr[key][0])
                    else:
                        lo.append(0.0); up.append(0.0)
                r0 += N
    # 4. friction pyramid (constraints.py:153-185,215-217): contact-major, 5 rows/knot
    blocks["friction"] = r0
    pyr = friction_pyramid(prob["mu"])
    f_off = 2 if prob["robot"] == "TALOS" else 0
    for c in range(nc):
        for k in range(N):
            if prob["contact_active"][k, c]:
                G = pyr @ prob["contact_R"][k, c].T
                for j in range(4):                      # range(4): 5th row never written
                    for a in range(3):
                        add(r0 + 5 * k + j, idx_u(prob, k, npc * c + f_off + a), G[j, a])
        ubc = [0.0] * (5 * N)
        if prob.get("friction_ub") is not None:       # stochastic mode: chance-constraint back-offs
            for k in range(N):
                ubc[5 * k:5 * k + 4] = list(prob["friction_ub"][k, c])
        lo += [-np.inf] * (5 * N); up += ubc
        r0 += 5 * N
    # 5. L1 trust region on angular momentum (constraints.py:260-293)
    blocks["trust"] = r0
    S = penum_mat()
    for k in range(N + 1):
        for j in range(8):
            for a in range(3):
                add(r0 + 8 * k + j, idx_x(prob, k, 6 + a), S[j, a])
            add(r0 + 8 * k + j, idx_t(prob, k), -1.0 / weight)
        lo += [-np.inf] * 8
        up += list(radius + S @ X32[6:9, k])
    r0 += 8 * (N + 1)
    blocks["slack_sign"] = r0
    for k in range(N + 1):
        add(r0 + k, idx_t(prob, k), -1.0)
    lo += [-np.inf] * (N + 1); up += [0.0] * (N + 1)
    r0 += N + 1
    blocks["m"] = r0
    A = sparse.csc_matrix((vals, (rows, cols)), shape=(r0, nz))
    return A, np.array(lo), np.array(up), blocks


def unpack(prob, z):
    """get_QP_solution, scp_solver.py:89-93 (order='F' reshape)."""
    N, nx, nu, nc, nz = dims(prob)
    X = np.reshape(z[:nx * (N + 1)], (nx, N + 1), order="F")
    U = np.reshape(z[nx * (N + 1):nx * (N + 1) + nu * N], (nu, N), order="F")
    return X, U


def friction_backoffs(prob, gains, covs, beta_u):
    """Upper bounds of the friction rows in stochastic mode, constraints.py:157-163,187-214:
    ub[k,c,j] = -sum_u xi 2 G_ju sqrt((K_c Sigma_k K_c')_uu) over the entries with G_ju > 1e-6 and
    sqrt(.) > 1e-6, for k > 0 and active contacts; xi = Phi^-1(1 - beta_u / 5 * 3).  The
    covariance-gradient terms of the reference are identically zero.
    Returns (ub [N, nc, 4], xi)."""
    from scipy.stats import norm
    N, nc = prob["N"], prob["contact_active"].shape[1]
    pyr = friction_pyramid(prob["mu"])
    xi = norm.ppf(1 - (beta_u / pyr.shape[0] * 3))
    ub = np.zeros((N, nc, 4))
    for c in range(nc):
        for k in range(1, N):
            if not prob["contact_active"][k, c]:
                continue
            G = pyr @ prob["contact_R"][k, c].T
            K = gains[k, 3 * c:3 * c + 3, :]
            KSK = K @ covs[k] @ K.T
            for j in range(4):
                for u in range(3):
                    with np.errstate(invalid="ignore"):
                        sq = np.sqrt(KSK[u, u])
                    if G[j, u] > 1e-6 and sq > 1e-6:
                        ub[k, c, j] -= xi * (2 * G[j, u] * sq)
    return ub, xi

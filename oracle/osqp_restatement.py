"""Restatement of the OSQP algorithm as the reference calls it (oracle, float64).

Call site: /root/reference/src/scp_solver.py:59-68
    prob.setup(P, q, A, l, u, warm_start=True, verbose=False,
               eps_abs=1e-7, eps_rel=1e-7, polish=True); res = prob.solve()
    status != 'solved'  ->  failure.

OSQP itself is a third-party dependency that is NOT under /root/reference and is
imported un-pinned (setup.py:1-7; the shipped egg targets Python 3.8, i.e. the
OSQP 0.6.x line).  What follows restates the published algorithm (Stellato,
Banjac, Goulart, Bemporad, Boyd: "OSQP: an operator splitting solver for
quadratic programs", Math. Prog. Comp. 2020) with the 0.6.x defaults:
rho=0.1, sigma=1e-6, alpha=1.6, max_iter=4000, scaling=10 (modified Ruiz on the
KKT matrix + cost normalisation), adaptive_rho with tolerance 5,
check_termination=25, eps_prim_inf=eps_dual_inf=1e-4, rho_eq = 1e3*rho for rows
with u-l < 1e-4, polish with delta=1e-6 and 3 refinement steps.

One knowing deviation: OSQP chooses its adaptive-rho interval from wall-clock
timing (a multiple of 25 that depends on setup time), so its inner iteration
counts are not reproducible.  Here the interval is the fixed setting
``adaptive_rho_interval`` (default 50).  Only the answer and the verdict
('solved' or not) are parity quantities, not the inner iteration count.

PARITY UNPINNED: see oracle/__init__.py.
"""
import numpy as np
from scipy import sparse
from scipy.sparse.linalg import splu

OSQP_INFTY = 1e30
MIN_SCALING, MAX_SCALING = 1e-4, 1e4
RHO_MIN, RHO_MAX = 1e-6, 1e6
RHO_TOL = 1e-4
RHO_EQ_OVER_RHO_INEQ = 1e3

DEFAULTS = dict(rho=0.1, sigma=1e-6, alpha=1.6, max_iter=4000, scaling=10,
                adaptive_rho=True, adaptive_rho_interval=50, adaptive_rho_tolerance=5.0,
                check_termination=25, eps_abs=1e-3, eps_rel=1e-3,
                eps_prim_inf=1e-4, eps_dual_inf=1e-4, polish=False, delta=1e-6,
                polish_refine_iter=3)


class Result:
    pass


def _limit_scaling(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.minimum(v, MAX_SCALING)


def _col_inf_norm(M):
    M = sparse.csc_matrix(M)
    out = np.zeros(M.shape[1])
    if M.nnz:
        absM = abs(M)
        out = np.asarray(absM.max(axis=0).todense()).ravel()
    return out


def _ruiz(P, q, A, l, u, iters):
    n, m = P.shape[0], A.shape[0]
    D, E, c = np.ones(n), np.ones(m), 1.0
    P, A, q = sparse.csc_matrix(P, dtype=np.float64), sparse.csc_matrix(A, dtype=np.float64), q.copy()
    for _ in range(iters):
        Dn = np.maximum(_col_inf_norm(P), _col_inf_norm(A))
        En = _col_inf_norm(A.T)
        Dt = 1.0 / np.sqrt(_limit_scaling(Dn))
        Et = 1.0 / np.sqrt(_limit_scaling(En))
        P = sparse.diags(Dt) @ P @ sparse.diags(Dt)
        A = sparse.diags(Et) @ A @ sparse.diags(Dt)
        q = Dt * q
        D *= Dt
        E *= Et
        ct = np.mean(_col_inf_norm(P))
        qn = _limit_scaling(np.array([np.max(np.abs(q))]))[0]
        ct = _limit_scaling(np.array([max(ct, qn)]))[0]
        ct = 1.0 / ct
        P = P * ct
        q = q * ct
        c *= ct
    return sparse.csc_matrix(P), q, sparse.csc_matrix(A), E * l, E * u, D, E, c


def _rho_vec(l, u, rho):
    v = np.full(l.shape, rho)
    eq = (u - l) < RHO_TOL
    v[eq] = RHO_EQ_OVER_RHO_INEQ * rho
    free = (l < -OSQP_INFTY * MIN_SCALING) & (u > OSQP_INFTY * MIN_SCALING)
    v[free] = RHO_MIN
    return v


def _factor(P, A, sigma, rho_vec):
    n, m = P.shape[0], A.shape[0]
    K = sparse.bmat([[P + sigma * sparse.eye(n), A.T],
                     [A, -sparse.diags(1.0 / rho_vec)]], format="csc")
    return splu(K)


def solve(P, q, A, l, u, **settings):
    """min 1/2 x'Px + q'x  s.t.  l <= Ax <= u.  Returns a Result with x, y, status, iter."""
    s = dict(DEFAULTS)
    s.update(settings)
    n, m = P.shape[0], A.shape[0]
    l = np.maximum(np.asarray(l, dtype=np.float64), -OSQP_INFTY)
    u = np.minimum(np.asarray(u, dtype=np.float64), OSQP_INFTY)
    Ps, qs, As, ls, us, D, E, c = _ruiz(P, np.asarray(q, dtype=np.float64), A, l, u, s["scaling"])
    Dinv, Einv, cinv = 1.0 / D, 1.0 / E, 1.0 / c
    rho = s["rho"]
    rv = _rho_vec(ls, us, rho)
    lu = _factor(Ps, As, s["sigma"], rv)
    n_fact = 1
    x, z, y = np.zeros(n), np.zeros(m), np.zeros(m)
    alpha, sigma = s["alpha"], s["sigma"]
    AsT = As.T.tocsc()
    status = "maximum iterations reached"
    it = 0

    def residuals(x, z, y):
        Ax = As @ x
        Px = Ps @ x
        Aty = AsT @ y
        pri = np.max(np.abs(Einv * (Ax - z))) if m else 0.0
        dua = cinv * np.max(np.abs(Dinv * (Px + qs + Aty)))
        eps_pri = s["eps_abs"] + s["eps_rel"] * max(np.max(np.abs(Einv * Ax)), np.max(np.abs(Einv * z)))
        eps_dua = s["eps_abs"] + s["eps_rel"] * cinv * max(np.max(np.abs(Dinv * Px)),
                                                         np.max(np.abs(Dinv * Aty)),
                                                         np.max(np.abs(Dinv * qs)))
        return pri, dua, eps_pri, eps_dua, Ax, Px, Aty

    pri = dua = np.inf
    for it in range(1, s["max_iter"] + 1):
        x_prev, z_prev = x, z
        rhs = np.concatenate([sigma * x_prev - qs, z_prev - y / rv])
        sol = lu.solve(rhs)
        xt, nu = sol[:n], sol[n:]
        zt = z_prev + (nu - y) / rv
        x = alpha * xt + (1 - alpha) * x_prev
        zr = alpha * zt + (1 - alpha) * z_prev
        z = np.clip(zr + y / rv, ls, us)
        dy = rv * (zr - z)
        y = y + dy
        dx = x - x_prev
        check = s["check_termination"] and it % s["check_termination"] == 0
        adapt = s["adaptive_rho"] and s["adaptive_rho_interval"] and it % s["adaptive_rho_interval"] == 0
        if check or adapt:
            pri, dua, eps_pri, eps_dua, Ax, Px, Aty = residuals(x, z, y)
            if pri <= eps_pri and dua <= eps_dua:
                status = "solved"
                break
            # primal infeasibility certificate: delta_y projected on the polar of the
            # recession cone of [l,u]
            big = OSQP_INFTY * MIN_SCALING
            dyp = dy.copy()
            up_inf, lo_inf = us > big, ls < -big
            dyp[up_inf & lo_inf] = 0.0
            dyp[up_inf & ~lo_inf] = np.minimum(dyp[up_inf & ~lo_inf], 0.0)
            dyp[lo_inf & ~up_inf] = np.maximum(dyp[lo_inf & ~up_inf], 0.0)
            ndy = np.max(np.abs(E * dyp)) if m else 0.0
            if ndy > s["eps_prim_inf"]:
                supp = np.where(up_inf, 0.0, us) @ np.maximum(dyp, 0) + np.where(lo_inf, 0.0, ls) @ np.minimum(dyp, 0)
                if supp < -s["eps_prim_inf"] * ndy:
                    if np.max(np.abs(Dinv * (AsT @ dyp))) <= s["eps_prim_inf"] * ndy:
                        status = "primal infeasible"
                        break
            # dual infeasibility certificate: delta_x
            ndx = np.max(np.abs(D * dx))
            if ndx > s["eps_dual_inf"] and (qs @ dx) < -c * s["eps_dual_inf"] * ndx:
                if np.max(np.abs(Dinv * (Ps @ dx))) < c * s["eps_dual_inf"] * ndx:
                    Adx = Einv * (As @ dx)
                    tol = s["eps_dual_inf"] * ndx
                    bad = ((~up_inf) & (Adx > tol)) | ((~lo_inf) & (Adx < -tol))
                    if not np.any(bad):
                        status = "dual infeasible"
                        break
        if adapt:
            p_n = np.max(np.abs(Ax - z)) / (max(np.max(np.abs(z)), np.max(np.abs(Ax))) + 1e-10)
            d_n = np.max(np.abs(Px + qs + Aty)) / (max(np.max(np.abs(qs)), np.max(np.abs(Aty)),
                                                       np.max(np.abs(Px))) + 1e-10)
            rho_new = float(np.clip(rho * np.sqrt(p_n / (d_n + 1e-10)), RHO_MIN, RHO_MAX))
            if rho_new > rho * s["adaptive_rho_tolerance"] or rho_new < rho / s["adaptive_rho_tolerance"]:
                rho = rho_new
                rv = _rho_vec(ls, us, rho)
                lu = _factor(Ps, As, sigma, rv)
                n_fact += 1
    else:
        # max_iter: OSQP reports 'solved inaccurate' if the 10x-looser test passes
        pri, dua, eps_pri, eps_dua, *_ = residuals(x, z, y)
        eps_pri10 = 10 * eps_pri
        eps_dua10 = 10 * eps_dua
        if pri <= eps_pri10 and dua <= eps_dua10:
            status = "solved inaccurate"

    res = Result()
    res.polished = False
    if status == "solved" and s["polish"]:
        pol = _polish(Ps, qs, As, ls, us, x, z, y, Dinv, Einv, cinv, s)
        if pol is not None:
            xp, zp, yp, ppri, pdua = pol
            if (ppri < pri and pdua < dua) or (ppri < pri and dua < 1e-10) or (pdua < dua and pri < 1e-10):
                x, z, y = xp, zp, yp
                pri, dua = ppri, pdua
                res.polished = True
    res.x = D * x
    res.y = cinv * (E * y)
    res.z = Einv * z
    res.status = status
    res.iter = it
    res.n_fact = n_fact
    res.rho = rho
    res.pri_res, res.dua_res = pri, dua
    return res


def _polish(Ps, qs, As, ls, us, x, z, y, Dinv, Einv, cinv, s):
    n, m = Ps.shape[0], As.shape[0]
    low = (z - ls) < -y
    upp = (us - z) < y
    act = np.concatenate([np.where(low)[0], np.where(upp)[0]])
    Ared = As.tocsr()[act, :].tocsc()
    na = len(act)
    delta = s["delta"]
    Kp = sparse.bmat([[Ps + delta * sparse.eye(n), Ared.T],
                      [Ared, -delta * sparse.eye(na)]], format="csc")
    K = sparse.bmat([[Ps, Ared.T], [Ared, None if na == 0 else sparse.csc_matrix((na, na))]], format="csc") \
        if na else Ps
    rhs = np.concatenate([-qs, ls[low], us[upp]])
    try:
        lu = splu(Kp)
    except RuntimeError:
        return None
    sol = lu.solve(rhs)
    for _ in range(s["polish_refine_iter"]):
        sol = sol + lu.solve(rhs - K @ sol)
    xp = sol[:n]
    yp = np.zeros(m)
    yp[act] = sol[n:]
    Ax = As @ xp
    zp = np.clip(Ax, ls, us)
    pri = np.max(np.abs(Einv * (Ax - zp)))
    dua = cinv * np.max(np.abs(Dinv * (Ps @ xp + qs + As.T @ yp)))
    return xp, zp, yp, pri, dua

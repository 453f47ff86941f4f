"""numpy model of the DEVICE algorithm (csrc/cmpc_tile.cuh) — test infrastructure only.

The CUDA solver does not restate OSQP's linear algebra; it solves the same QP with an ADMM
whose x-update treats the linearised dynamics exactly (a time-varying LQR solved by a
Riccati recursion) and splits only the friction / trust-region / slack / terminal rows.
This file is the executable specification of that algorithm: tests compare the CUDA
kernels (and their host build) against it step by step, and against the OSQP restatement
for the answer.  See DESIGN.md "device algorithm".
"""
import numpy as np

from . import dynamics

ALPHA = 1.8     # csrc/cmpc_params.h default_qp_settings
CHECK = 25           # termination test every CHECK iterations (OSQP default) (knot-local residuals, see iterate())
PER_ROW_FRICTION = True
RHO0 = 2.0
ADAPT_START = 200    # early residuals are transient; adapting on them hurts
AS_START = 8         # first certified-polish attempt after this many ADMM iterations
AS_STEP = 8          # interval to the next attempt (doubled after every failure)
AS_ROUNDS = 9        # active-set correction rounds per attempt
AS_TOL = 1e-9        # certificate tolerance (primal residual / row violation)
RHO_K_REL = 1.0       # kappa-copy penalty = rho * RHO_K_REL * min(W_kappa)
RHO_E_POL_REL = 1e4   # terminal-equality penalty while polishing
RHO_E_REL = 100.0     # terminal-equality penalty = RHO_E_REL * max(W_x), independent of rho


def skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


class Stage:
    """Linearised problem in stage order."""

    def __init__(self, prob):
        N = int(prob["N"])
        nc = prob["contact_active"].shape[1]
        self.N, self.nc, self.nu = N, nc, 3 * nc
        assert prob["robot"] != "TALOS"
        self.m, self.g, self.dt, self.mu = prob["m"], prob["g"], prob["dt"], prob["mu"]
        self.Wx = np.diag(prob["state_cost_weights"]).copy()
        self.Wu = np.diag(prob["control_cost_weights"]).copy()
        X, U = prob["X_ref"], prob["U_init"]
        self.Xbar, self.Ubar = X, U
        self.act = prob["contact_active"].astype(bool)
        self.R = prob["contact_R"]
        self.A = np.zeros((N, 9, 9)); self.B = np.zeros((N, 9, self.nu)); self.c = np.zeros((N, 9))
        self.fbar = np.zeros((N, 9))
        for k in range(N):
            args = (X[:, k], U[:, k], prob["contact_pos"][k], prob["contact_active"][k],
                    prob["contact_R"][k], self.m, self.g, self.dt, "solo12")
            f = dynamics.step(*args)
            A, B, _ = dynamics.jacobians(*args)
            self.A[k], self.B[k] = A, B
            self.fbar[k] = f
            self.c[k] = f - A @ X[:, k] - B @ U[:, k]
        self.q = np.stack([-self.Wx * X[:, k] for k in range(N + 1)])       # (N+1, 9)
        self.x_init = prob["x_init"].copy()
        self.x_final = prob["x_final"].copy()
        ml = self.mu / np.sqrt(2.0)
        self.pyr = np.array([[1, 0, -ml], [-1, 0, -ml], [0, 1, -ml], [0, -1, -ml]], dtype=float)
        self.S = np.array([[(-1.0) ** (j // 2 ** i) for i in range(3)] for j in range(8)])
        # friction row matrices per knot/contact (4x3), zero when inactive
        self.G = np.zeros((N, nc, 4, 3))
        for k in range(N):
            for i in range(nc):
                if self.act[k, i]:
                    self.G[k, i] = self.pyr @ self.R[k, i].T


def prox_trust(a, kbar, r, omega, rho):
    """argmin_v omega*max(0, ||v-kbar||_1 - r) + rho/2 ||v-a||^2  -> (v, branch).

    branch 0: inside the L1 ball (penalty inactive); 1: outside after soft-thresholding by
    omega/rho (penalty linear); 2: on the ball's surface (projection)."""
    b = a - kbar
    ab = np.abs(b)
    if ab.sum() <= r:
        return a.copy(), 0
    d = np.sign(b) * np.maximum(ab - omega / rho, 0.0)
    if np.abs(d).sum() >= r:
        return kbar + d, 1
    srt = np.sort(ab)[::-1]
    css = np.cumsum(srt)
    tau = 0.0
    for j in range(3):
        t = (css[j] - r) / (j + 1)
        if srt[j] - t > 0:
            tau = t
    return kbar + np.sign(b) * np.maximum(ab - tau, 0.0), 2


class RiccatiADMM:
    """ADMM on   min sum_k 1/2|x_k|^2_Wx + q_k'x_k + 1/2|u_k|^2_Wu
                       + sum_{k>=1} omega*max(0, |kappa_k - kappa_bar_k|_1 - radius)
                 s.t. x_{k+1} = A_k x_k + B_k u_k + c_k, x_0 = x_init   (exact, inside the x-update)
                      G f <= 0 (friction rows), x_N = x_final           (split)
    which is the reference QP with the slack t_k minimised out in closed form
    (t_k = omega*max(0, |kappa_k-kappa_bar_k|_1 - radius); constraints.py:277-289, cost.py:34-39).
    Split blocks: friction rows (w_f, y_f), a copy v_k of kappa_k with the prox of the exact
    penalty (w_k, y_k), the terminal equality (w_e, y_e).

    The x-update has no proximal term (OSQP's sigma): W_x, W_u > 0 make it strictly convex, so
    the iteration only carries (w, y) — stored on the device as v = w + y/rho per row."""

    def __init__(self, st, radius, weight, rho=RHO0, alpha=ALPHA):
        self.st = st
        self.radius, self.weight = radius, weight
        self.alpha = alpha
        N, nu, nc = st.N, st.nu, st.nc
        # start at the linearisation point (the answer of a convex QP does not depend on it;
        # the reference cold-starts OSQP at zero)
        self.x = st.Xbar.T.copy(); self.x[0] = st.x_init
        self.u = st.Ubar.T.copy() * np.repeat(st.act, 3, axis=1)
        cf0 = np.einsum("kiab,kib->kia", st.G, self.u.reshape(N, nc, 3))
        self.wf = np.minimum(cf0, 0.0); self.yf = np.zeros((N, nc, 4))
        self.wk = st.Xbar[6:9, :].T.copy(); self.yk = np.zeros((N + 1, 3))
        self.we = st.x_final.copy(); self.ye = np.zeros(9)
        self.kbar = st.Xbar[6:9, :].T.copy()
        self.branch = np.zeros(N + 1, dtype=int)
        self.n_fact = 0
        self.lin_k = np.zeros((N + 1, 3))      # polish only: linear penalty slope
        self.Mk = None                          # polish only: 3x3 kappa blocks
        self.dynrow = max(float(np.abs(st.c).max()), float(np.abs(st.x_init).max()))
        self.nq = float(np.abs(st.q).max())
        self.set_rho(rho)
        self.factor()

    def set_rho(self, rho):
        """Per-row penalties rho * e_row^2 with e_row the row's equilibration factor under the
        fixed variable scaling D = 1/sqrt(cost weight)."""
        st, N = self.st, self.st.N
        self.rho = rho
        Du = 1.0 / np.sqrt(st.Wu)
        rown = np.abs(st.G * Du.reshape(st.nc, 1, 3)[None]).max(axis=3)          # (N, nc, 4)
        with np.errstate(divide="ignore"):
            ef2 = np.where(st.act[:, :, None], 1.0 / np.maximum(rown, 1e-300) ** 2, 0.0)
        if not PER_ROW_FRICTION:
            ef2 = np.where(st.act[:, :, None], 1.0, 0.0) * np.min(st.Wu)
        self.rf = rho * ef2
        self.rk = np.full(N + 1, rho * RHO_K_REL * float(np.min(st.Wx[6:9])))
        self.rk[0] = 0.0
        self.re = RHO_E_REL * float(np.max(st.Wx))
        self.Mk = None

    # ---------------------------------------------------------------- factorisation
    def factor(self):
        st, N, nu = self.st, self.st.N, self.st.nu
        self.K = np.zeros((N, nu, 9)); self.Hinv = np.zeros((N, nu, nu)); self.Pc = np.zeros((N, 9))

        def Qk(k):
            Q = np.diag(st.Wx).copy()
            if self.Mk is not None:
                Q[6:9, 6:9] += self.Mk[k]
            else:
                Q[6:9, 6:9] += self.rk[k] * np.eye(3)
            return Q
        P = Qk(N) + self.re * np.eye(9)
        for k in range(N - 1, -1, -1):
            Rk = np.diag(st.Wu).copy()
            for i in range(st.nc):
                Rk[3 * i:3 * i + 3, 3 * i:3 * i + 3] += st.G[k, i].T @ (self.rf[k, i][:, None] * st.G[k, i])
            A, B = st.A[k], st.B[k]
            PB = P @ B
            Huu = Rk + B.T @ PB
            Hux = PB.T @ A
            Hinv = np.linalg.inv(Huu)
            Hinv = 0.5 * (Hinv + Hinv.T)
            K = -Hinv @ Hux
            self.K[k], self.Hinv[k] = K, Hinv
            self.Pc[k] = P @ st.c[k]
            P = Qk(k) + A.T @ P @ A + Hux.T @ K
            P = 0.5 * (P + P.T)
        self.n_fact += 1

    def rows(self, x, u):
        st = self.st
        cf = np.einsum("kiab,kib->kia", st.G, u.reshape(st.N, st.nc, 3))
        return cf, x[:, 6:9].copy(), x[st.N].copy()

    def lqr_solve(self, qx, ru):
        st, N = self.st, self.st.N
        p = qx[N].copy()
        d = np.zeros((N, st.nu))
        for k in range(N - 1, -1, -1):
            g = p + self.Pc[k]
            hu = ru[k] + st.B[k].T @ g
            d[k] = -self.Hinv[k] @ hu
            p = qx[k] + st.A[k].T @ g + self.K[k].T @ hu
        x = np.zeros((N + 1, 9)); u = np.zeros((N, st.nu))
        x[0] = st.x_init
        for k in range(N):
            u[k] = self.K[k] @ x[k] + d[k]
            x[k + 1] = st.A[k] @ x[k] + st.B[k] @ u[k] + st.c[k]
        return x, u

    def x_update(self):
        st, N = self.st, self.st.N
        vf = self.rf * self.wf - self.yf
        ve = self.re * self.we - self.ye
        qx = st.q.copy()
        if self.Mk is None:
            qx[:, 6:9] -= self.rk[:, None] * self.wk - self.yk
        else:   # polish: general rows on kappa, see polish()
            qx[:, 6:9] += self.lin_k - np.einsum("kab,kb->ka", self.Mk, self.wk) + self.yk
        qx[N] -= ve
        ru = -np.einsum("kiab,kia->kib", st.G, vf).reshape(N, st.nu)
        return self.lqr_solve(qx, ru)

    def iterate(self, check=False):
        """One ADMM iteration.  With check=True also returns (pri, dua, npri, ndua), OSQP's
        unscaled infinity-norm residuals of the point (x~, w_new, y_new, nu~), nu~ being the
        multipliers of the exactly-treated rows (dynamics, x_0) from the x-update itself.  With
        that choice the stationarity residual is  A_s'(y_new - y_old - rho (A_s x~ - w_old)),
        A_s the split rows, which is local to each knot: no costate recursion is needed."""
        st, al, N = self.st, self.alpha, self.st.N
        xt, ut = self.x_update()
        cf, ck, ce = self.rows(xt, ut)
        self.x, self.u = xt, ut
        wf0, yf0, wk0, yk0, ye0 = self.wf, self.yf, self.wk.copy(), self.yk.copy(), self.ye
        # friction rows: one-sided box
        zr = al * cf + (1 - al) * wf0
        with np.errstate(divide="ignore", invalid="ignore"):
            v = np.where(self.rf > 0, zr + yf0 / np.where(self.rf > 0, self.rf, 1.0), zr)
        wn = np.minimum(v, 0.0)
        self.yf = yf0 + self.rf * (zr - wn)
        self.wf = wn
        # kappa copies: prox of the exact trust-region penalty
        for k in range(1, N + 1):
            zk = al * ck[k] + (1 - al) * wk0[k]
            wkn, self.branch[k] = prox_trust(zk + yk0[k] / self.rk[k], self.kbar[k], self.radius,
                                             self.weight, self.rk[k])
            self.yk[k] = yk0[k] + self.rk[k] * (zk - wkn)
            self.wk[k] = wkn
        # terminal equality
        ze = al * ce + (1 - al) * self.we
        self.ye = ye0 + self.re * (ze - st.x_final)
        self.we = st.x_final.copy()
        if not check:
            return None
        df = self.yf - yf0 - self.rf * (cf - wf0)
        dk = self.yk - yk0 - self.rk[:, None] * (ck - wk0)
        dk[0] = 0.0
        # terminal rows are equalities: the certificate takes y_e + rho_e (x_N - x_f), the
        # multiplier of the x-update itself, so they contribute no stationarity residual
        rd_u = np.einsum("kiab,kia->kib", st.G, df).reshape(N, st.nu)
        rd_x = np.zeros((N + 1, 9))
        rd_x[:, 6:9] = dk
        pri = max(np.abs(cf - self.wf).max(), np.abs(ck[1:] - self.wk[1:]).max(), np.abs(ce - st.x_final).max())
        dua = max(np.abs(rd_u).max(), np.abs(rd_x).max())
        npri = max(np.abs(cf).max(), np.abs(ck[1:]).max(), np.abs(ce).max(), np.abs(self.wf).max(),
                   np.abs(self.wk[1:]).max(), np.abs(st.x_final).max(), self.dynrow)
        Px, Pu = st.Wx * xt, st.Wu * ut
        aty_x = rd_x - Px - st.q
        aty_u = rd_u - Pu
        ndua = max(np.abs(Px[1:]).max(), np.abs(Pu).max(), np.abs(aty_x[1:]).max(), np.abs(aty_u).max(), self.nq)
        return pri, dua, npri, ndua

    def solve(self, eps_abs=1e-7, eps_rel=1e-7, max_iter=4000, check=CHECK, adapt=True, adapt_tol=5.0,
              polish=True, verbose=False, adapt_start=ADAPT_START, as_start=AS_START, as_step=AS_STEP,
              as_rounds=AS_ROUNDS, as_tol=AS_TOL, refine=3):
        """ADMM (active-set predictor and globally convergent fallback) interleaved with
        certified active-set polishes: the first attempt after `as_start` iterations, later ones
        with a doubling interval; an attempt that certifies a KKT point ends the solve.  When the
        OSQP termination test passes first, the polish is accepted by OSQP's rule (it improves)."""
        status, it = "maximum iterations reached", max_iter
        self.polished = False
        self.pol_res = (np.nan, np.nan)
        next_as = as_start if (polish and as_start > 0) else -1
        for it in range(1, max_iter + 1):
            chk = (it % check == 0) or it == next_as
            res = self.iterate(check=chk)
            if res is None:
                continue
            pri, dua, npri, ndua = res
            if verbose:
                print(it, "pri %.3e dua %.3e rho %.3g" % (pri, dua, self.rho))
            term = pri <= eps_abs + eps_rel * npri and dua <= eps_abs + eps_rel * ndua
            if term or it == next_as:
                if polish:
                    cert, ppri, pnpri = self.polish(rounds=as_rounds, refine=refine, tol=as_tol, npri0=npri)
                    m0 = max(pri / (eps_abs + eps_rel * npri), dua / (eps_abs + eps_rel * ndua))
                    m1 = ppri / (eps_abs + eps_rel * pnpri)
                    if cert or (term and m1 < m0):
                        self.x, self.u = self.x_pol, self.u_pol
                        self.polished = True
                        self.pol_res = (ppri, 0.0)
                        status = "solved"
                        break
                if term:
                    status = "solved"
                    self.x, self.u = self.x_update()     # the device re-runs the LQR solve read-only
                    break
                next_as = it + as_step
                as_step *= 2
                self.factor()
            if adapt and it >= adapt_start and it % check == 0:
                est = self.rho * np.sqrt((pri / (npri + 1e-10)) / (dua / (ndua + 1e-10) + 1e-10))
                est = float(np.clip(est, 1e-6, 1e6))
                if est > self.rho * adapt_tol or est < self.rho / adapt_tol:
                    self.set_rho(est)       # w, y are kept (as OSQP does)
                    self.factor()
        return status, it

    # ---------------------------------------------------------------- polish
    def polish(self, delta=1e-6, refine=3, rounds=AS_ROUNDS, tol=AS_TOL, npri0=0.0):
        """OSQP-style polish (guess the active set, solve the equality-constrained QP) done as a
        method of multipliers with penalty 1/delta on the active rows — algebraically OSQP's
        regularised KKT solve + iterative refinement — reusing factor / x_update.

        Active structure: friction row active iff -w < y (OSQP's test u - z < y); terminal rows
        always; trust-region penalty per knot by the branch its prox took last:
          0 inside  -> nothing;  1 outside -> linear cost omega*sign on the non-zero
          components, zero components pinned to kappa_bar;  2 surface -> the equality
          sign'(kappa-kappa_bar) = radius, zero components pinned.
        After each round the friction active set is corrected (rows that came out violated by
        more than tol join, rows whose multiplier came out negative leave) and the round is
        repeated, at most 1 + rounds times.

        Returns (certified, pri, npri).  certified: no row wants to change and the primal
        residual is <= tol; stationarity is exact by construction, so the polished point
        (x_pol, u_pol) is a KKT point of the QP however far the ADMM iterate still was.  The ADMM
        state (w, y, rho) is left untouched."""
        st, N = self.st, self.st.N
        names = ("x", "u", "wf", "yf", "wk", "yk", "we", "ye", "rf", "rk", "re", "lin_k")
        keep = {n: np.copy(getattr(self, n)) for n in names}
        inv = 1.0 / delta
        af = ((0.0 - self.wf) < self.yf) & st.act[:, :, None]
        ypol = self.yf * af
        self.re = RHO_E_POL_REL * float(np.max(st.Wx))
        self.we = st.x_final.copy()
        # kappa rows: up to three pins e_i and one sign row per knot -> 3x3 block Mk
        self.Mk = np.zeros((N + 1, 3, 3))
        rows_k = [[] for _ in range(N + 1)]       # list of (row vector, rhs) per knot
        self.lin_k = np.zeros((N + 1, 3))
        ymul = [[] for _ in range(N + 1)]
        for k in range(1, N + 1):
            v = self.wk[k] + self.yk[k] / self.rk[k]
            wk, br = prox_trust(v, self.kbar[k], self.radius, self.weight, self.rk[k])
            if br == 0:
                continue
            d = wk - self.kbar[k]
            yk = self.rk[k] * (v - wk)
            sgn = np.sign(d)
            for i in range(3):
                if sgn[i] == 0.0:
                    e = np.zeros(3); e[i] = 1.0
                    rows_k[k].append((e, self.kbar[k][i])); ymul[k].append(yk[i])
            if br == 1:
                self.lin_k[k] = self.weight * sgn
            else:
                nz = sgn != 0
                mult = float(np.mean((yk / np.where(nz, sgn, 1.0))[nz])) if nz.any() else 0.0
                rows_k[k].append((sgn.copy(), self.radius + sgn @ self.kbar[k])); ymul[k].append(mult)
        for k in range(N + 1):
            for (a, b) in rows_k[k]:
                self.Mk[k] += inv * np.outer(a, a)
        lin_keep = self.lin_k.copy()
        certified, prev_chg = False, 1 << 30
        pri = np.inf
        npri = npri0      # thresholds are tol * (1 + npri of the previous sweep): absolute + relative

        def sweep():
            self.wk = np.zeros((N + 1, 3)); self.yk = np.zeros((N + 1, 3))
            extra = np.zeros((N + 1, 3))
            for k in range(N + 1):
                for (a, b), y in zip(rows_k[k], ymul[k]):
                    extra[k] += a * (inv * b - y)
            self.lin_k = lin_keep - extra
            xt, ut = self.x_update()
            self.lin_k = lin_keep
            cf, ck, ce = self.rows(xt, ut)
            self.x_pol, self.u_pol = xt, ut
            self.yf = self.yf + self.rf * cf
            self.ye = self.ye + self.re * (ce - self.we)
            for k in range(N + 1):
                ymul[k] = [y + inv * (a @ ck[k] - b) for (a, b), y in zip(rows_k[k], ymul[k])]
            p = max(np.abs(cf * af).max(), np.maximum(cf * ~af, 0).max(), np.abs(ce - st.x_final).max())
            n = max(np.abs(cf).max(), np.abs(ck[1:]).max(), np.abs(ce).max(), np.abs(st.x_final).max(), self.dynrow)
            return cf, p, n

        for rnd in range(1 + rounds):
            self.rf = af * inv
            self.yf = ypol * af
            self.wf = np.zeros_like(self.wf)
            self.factor()

            def correct(cf, npri_prev):
                nonlocal af
                keep_r = af & ~(self.yf < 0.0)
                join = (~af) & st.act[:, :, None] & (cf > tol * (1.0 + npri_prev))
                new_af = keep_r | join
                n = int((new_af != af).sum())
                self.yf = np.where(keep_r, self.yf, 0.0)
                af = new_af
                return n
            npri_prev = npri
            cf, pri, npri = sweep()
            chg = correct(cf, npri_prev)   # the first multiplier sweep already corrects the active set
            if chg == 0:
                for sw in range(1 + refine):
                    npri_prev = npri
                    cf, pri, npri = sweep()
                    chg = correct(cf, npri_prev)
                    if chg or pri <= tol * (1.0 + npri):
                        break
            ypol = self.yf.copy()
            if chg == 0:
                certified = pri <= tol * (1.0 + npri)
                break
            prev_chg = chg
        self.Mk = None
        for n in names:
            setattr(self, n, keep[n])
        return certified, pri, npri


def spectral_norm_9xn(D):
    """sigma_max of a 9 x n matrix through its 9x9 Gram matrix (what the device does)."""
    G = D @ D.T
    return float(np.sqrt(max(np.linalg.eigvalsh(G)[-1], 0.0)))


def solve_scp(prob, scp_params, log=None, warm_start=False, **solver_kw):
    """Device-side SCP state machine (mirrors oracle.scp.solve_scp / scp_solver.py:118-179)."""
    st = Stage(prob)
    N = st.N
    rho0, rho1 = scp_params["rho0"], scp_params["rho1"]
    radius0 = float(scp_params["trust_region_radius0"])
    radius, weight = radius0, float(scp_params["omega0"])
    out = dict(state=[], control=[], iterations=0, status="ok")
    it, success = 0, False
    prev = None
    while it < scp_params["max_iterations"] and weight < scp_params["omega_max"] and \
            not (it != 0 and success and 0.0 < scp_params["convergence_threshold"]):
        success = False
        s = RiccatiADMM(st, radius, weight, rho=solver_kw.get("rho", RHO0))
        if warm_start and prev is not None:
            for name in ("x", "u", "wf", "yf", "wk", "yk", "we", "ye"):
                setattr(s, name, getattr(prev, name).copy())
        status, qp_it = s.solve(**{k: v for k, v in solver_kw.items() if k != "rho"})
        prev = s
        entry = dict(it=it, radius=radius, weight=weight, status=status, qp_iter=qp_it, n_fact=s.n_fact)
        if status != "solved":
            out["status"] = status
            if log is not None:
                log.append(entry)
            return False
        X, U = s.x.T.copy(), s.u.T.copy()
        snorm = spectral_norm_9xn(X - st.Xbar)
        entry["snorm"] = snorm
        if snorm < radius:
            num = den = 0.0
            for k in range(N):
                lin = st.fbar[k] + st.A[k] @ (X[:, k] - st.Xbar[:, k]) + st.B[k] @ (U[:, k] - st.Ubar[:, k])
                nl = dynamics.step(X[:, k], U[:, k], prob["contact_pos"][k], prob["contact_active"][k],
                                   prob["contact_R"][k], st.m, st.g, st.dt, "solo12")
                e = nl[6:] - lin[6:]
                num += e @ e
                den += lin @ lin
            rho = num / den
            entry["rho"] = rho
            if rho > rho1:
                radius *= scp_params["beta_fail"]
                entry["verdict"] = "inaccurate"
            else:
                out["state"].append(X); out["control"].append(U)
                success = True
                entry["verdict"] = "accepted"
                if rho < rho0:
                    radius = min(scp_params["beta_succ"] * radius, radius0)
        else:
            weight *= scp_params["gamma_fail"]
            entry["verdict"] = "outside"
        if log is not None:
            log.append(entry)
        it += 1
    out["iterations"] = it
    return out

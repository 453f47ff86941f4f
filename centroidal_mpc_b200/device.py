"""Device side of the host package: torch tensors as the batch container, raw device pointers
into libcmpc_b200.so (include/cmpc.h).  No arithmetic of the hot path happens in Python/torch;
without the CUDA library or without a GPU every entry point raises."""
import ctypes as C

import numpy as np

from . import _lib as L
from .batch import ProblemBatch


def _torch_cuda():
    import torch
    if not torch.cuda.is_available():
        raise L.CmpcError("no CUDA device: the SCP hot path runs only on the GPU (there is no CPU fallback)")
    return torch


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


# QP settings used when friction upper bounds are bound (stochastic mode) unless the caller overrides
# them: the multiplier iteration of a polish round needs more sweeps to reach the 1e-9 certificate with
# back-offs (bound gait: 3 sweeps leave a quarter of the instances uncertified for several attempts),
# and the active set needs up to 13 correction rounds at N = 100.  The library
# applies the same values by itself when it is called with qp = NULL and upper bounds are bound
# (cmpc_api.cu: launch_tiles), so a C caller gets them without knowing this table.
STOCHASTIC_QP_DEFAULTS = dict(polish_refine_iter=10, polish_active_set_rounds=19)
# ... and for the wrench contact model (TALOS): multipliers of the order of the 900 N contact forces make the
# default certificate tolerance 1e-9 leave 6e-6 in X; 1e-11 reaches the tightly solved oracle to 3e-7
# (relaxation 1.6 instead of the 1.8 of the point-contact default: with 1.8 a receding-horizon tick of the talos loop
# fails to certify and runs into the ADMM iteration cap, tests/test_mpc.py)
WRENCH_QP_DEFAULTS = dict(polish_refine_iter=30, polish_active_set_rounds=19, active_set_tol=1e-11, delta=1e-9, alpha=1.6)


class BatchSolver:
    """Owns a cmpc handle (solver workspace) for a fixed (B, N, nc) and the device copies of one
    ProblemBatch.  ``solve`` runs solve_scp for all instances on the current stream."""

    def __init__(self, batch, device=None):
        torch = _torch_cuda()
        self.lib = L.load()
        self.batch = batch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dims = L.cmpc_dims(batch.B, batch.N, batch.nc, 1 if batch.shared_plan else 0,
                                L.contact_model_of(batch.proto["robot"]))
        self.model = L.make_model_struct(batch.proto)
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self.lib.cmpc_create(C.byref(self.dims), C.byref(self.handle)), self.lib)
        B, N, nu = batch.B, batch.N, batch.nu
        f64 = torch.float64
        self.X = torch.zeros((B, N + 1, 9), dtype=f64, device=self.device)   # zeros: nothing is written for an
        self.U = torch.zeros((B, N, nu), dtype=f64, device=self.device)      # instance whose QP fails
        self._stream = None
        self.ints = torch.zeros((5, B), dtype=torch.int32, device=self.device)
        self.info = torch.zeros((B, 12), dtype=f64, device=self.device)
        self.d = {}
        self.friction_ub = None
        self.upload(batch)

    def upload(self, batch):
        """Host -> device copy of the problem data and (re)binding of the pointers."""
        torch = _torch_cuda()
        if (batch.B, batch.N, batch.nc, 1 if batch.shared_plan else 0) != \
                (self.dims.batch, self.dims.N, self.dims.nc, self.dims.shared_plan):
            raise L.CmpcError("upload(): batch dims (B=%d, N=%d, nc=%d, shared_plan=%s) differ from the handle's "
                              "(B=%d, N=%d, nc=%d, shared_plan=%d)" % (batch.B, batch.N, batch.nc, batch.shared_plan,
                                                                       self.dims.batch, self.dims.N, self.dims.nc,
                                                                       self.dims.shared_plan))
        self.batch = batch
        self.model = L.make_model_struct(batch.proto)
        for name in ("x_init", "x_final", "X_ref", "U_init", "contact_pos", "contact_R", "contact_active"):
            a = getattr(batch, name)
            self.d[name] = None if a is None else torch.from_numpy(a).to(self.device, non_blocking=True)
        d = self.d
        L.check(self.lib.cmpc_set_problem(self.handle, C.byref(self.model), _ptr(d["x_init"]), _ptr(d["x_final"]),
                                          _ptr(d["X_ref"]), _ptr(d["U_init"]), _ptr(d["contact_pos"]),
                                          _ptr(d["contact_R"]), _ptr(d["contact_active"])), self.lib)
        # Centroidal_model(conf, STOCHASTIC_OCP=True): friction rows G f <= ub with the chance-constraint
        # back-offs along the warm start (constraints.py:157-163,187-214), computed on the device
        sto = batch.proto.get("stochastic")
        self.friction_ub = None
        if sto is not None:
            with torch.cuda.device(self.device):
                gains, covs = lqr_gains_covs_batched(batch, d["X_ref"], d["U_init"], sto["Q"], sto["R"],
                                                     sto["cov_w"], sto["cov_eta"])
                self.friction_ub = friction_backoffs_batched(batch, gains, covs, sto["beta_u"])
        L.check(self.lib.cmpc_set_friction_ub(self.handle, _ptr(self.friction_ub)), self.lib)

    def solve(self, scp_params, qp_overrides=None, stream=None, out=None):
        """``out``: raw device pointers dict(X, U, scp_iters, status, n_accepted) to write the results to instead of
        this solver's own tensors -- e.g. this rank's slice of a buffer on another GPU (parallel.PeerResults)."""
        torch = _torch_cuda()
        scp = L.make_scp_struct(scp_params)
        qp = L.make_qp_struct(self._qp(qp_overrides), self.lib)
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        self._stream = None if stream is None else torch.cuda.ExternalStream(stream, device=self.device)
        if out is None:
            ptrs = (_ptr(self.X), _ptr(self.U), _ptr(self.ints[0]), _ptr(self.ints[1]), _ptr(self.ints[2]))
        else:
            ptrs = tuple(C.c_void_p(int(out[k])) for k in ("X", "U", "scp_iters", "status", "n_accepted"))
        with torch.cuda.device(self.device):
            L.check(self.lib.cmpc_solve_scp(self.handle, C.byref(scp), C.byref(qp), *ptrs, C.c_void_p(st)), self.lib)
        return self

    def _qp(self, overrides):
        if getattr(self.batch, "wrench", False):
            return dict(WRENCH_QP_DEFAULTS, **(overrides or {}))
        if self.friction_ub is None:
            return overrides
        return dict(STOCHASTIC_QP_DEFAULTS, **(overrides or {}))

    def _wait_for_solve(self):
        """Orders torch's current stream after the stream the last solve() ran on (a raw stream handed in
        by the caller is not ordered with torch's streams by itself)."""
        if self._stream is not None:
            torch = _torch_cuda()
            torch.cuda.current_stream(self.device).wait_stream(self._stream)

    def stats(self):
        torch = _torch_cuda()
        self._wait_for_solve()
        st = torch.cuda.current_stream(self.device).cuda_stream
        L.check(self.lib.cmpc_get_stats(self.handle, _ptr(self.ints[3]), _ptr(self.ints[4]), _ptr(self.info),
                                        C.c_void_p(st)), self.lib)
        return dict(qp_iters=self.ints[3].cpu().numpy(), n_factor=self.ints[4].cpu().numpy(),
                    info=self.info.cpu().numpy())

    def results(self):
        """Device -> host: dict of numpy arrays."""
        self._wait_for_solve()
        ints = self.ints.cpu().numpy()
        return dict(X=self.X.cpu().numpy(), U=self.U.cpu().numpy(), scp_iters=ints[0], status=ints[1],
                    n_accepted=ints[2])

    def solve_host(self, scp_params, qp_overrides=None, out=None):
        """End-to-end call with HOST buffers (cmpc_solve_scp_host): H2D, solve, D2H, sync."""
        b = self.batch
        B, N, nu = b.B, b.N, b.nu
        if out is None:
            out = dict(X=np.empty((B, N + 1, 9)), U=np.empty((B, N, nu)), scp_iters=np.empty(B, np.int32),
                       status=np.empty(B, np.int32), n_accepted=np.empty(B, np.int32))
        scp = L.make_scp_struct(scp_params)
        qp = L.make_qp_struct(self._qp(qp_overrides), self.lib)

        def p(a):
            return None if a is None else a.ctypes.data_as(C.c_void_p)
        torch = _torch_cuda()
        with torch.cuda.device(self.device):
            L.check(self.lib.cmpc_solve_scp_host(self.handle, C.byref(self.model), C.byref(scp), C.byref(qp),
                                                 p(b.x_init), p(b.x_final), p(b.X_ref), p(b.U_init),
                                                 p(b.contact_pos), p(b.contact_R), p(b.contact_active),
                                                 p(out["X"]), p(out["U"]), p(out["scp_iters"]), p(out["status"]),
                                                 p(out["n_accepted"])), self.lib)
        return out

    def close(self):
        if self.handle:
            self.lib.cmpc_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# Centroidal_model method wrappers (src/centroidal_model.py:189-291 of the reference)
# ---------------------------------------------------------------------------------------------
def _lin_call(model, X, U, want_jac):
    torch = _torch_cuda()
    lib = L.load()
    prob = model.problem_arrays()
    N, nc = prob["N"], prob["contact_active"].shape[1]
    wrench = prob["robot"] == "TALOS"
    nu = (6 if wrench else 3) * nc
    dev = torch.device("cuda", torch.cuda.current_device())
    Xd = torch.from_numpy(np.ascontiguousarray(np.asarray(X, dtype=np.float64).T[None])).to(dev)
    Ud = torch.from_numpy(np.ascontiguousarray(np.asarray(U, dtype=np.float64).T[None])).to(dev)
    cp = torch.from_numpy(np.ascontiguousarray(prob["contact_pos"][None])).to(dev)
    ca = torch.from_numpy(np.ascontiguousarray(prob["contact_active"][None].astype(np.int32))).to(dev)
    dims = L.cmpc_dims(1, N, nc, 1, L.contact_model_of(prob["robot"]))
    mdl = L.make_model_struct(prob)
    f = torch.empty((1, N, 9), dtype=torch.float64, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if wrench:   # CoP / wrench contact model: the frames enter the dynamics
        cR = torch.from_numpy(np.ascontiguousarray(prob["contact_R"][None].astype(np.float64))).to(dev)
        fx = torch.empty((1, N, 9, 9), dtype=torch.float64, device=dev) if want_jac else None
        fu = torch.empty((1, N, 9, nu), dtype=torch.float64, device=dev) if want_jac else None
        L.check(lib.cmpc_linearize_wrench(C.byref(dims), C.byref(mdl), _ptr(Xd), _ptr(Ud), _ptr(cp), _ptr(cR), _ptr(ca),
                                          _ptr(f), _ptr(fx), _ptr(fu), st), lib)
        if want_jac:
            return f[0].cpu().numpy(), fx[0].cpu().numpy(), fu[0].cpu().numpy()
        return f[0].cpu().numpy()
    if want_jac:
        fx = torch.empty((1, N, 9, 9), dtype=torch.float64, device=dev)
        fu = torch.empty((1, N, 9, nu), dtype=torch.float64, device=dev)
        L.check(lib.cmpc_linearize(C.byref(dims), C.byref(mdl), _ptr(Xd), _ptr(Ud), _ptr(cp), _ptr(ca), _ptr(f),
                                   _ptr(fx), _ptr(fu), st), lib)
        return f[0].cpu().numpy(), fx[0].cpu().numpy(), fu[0].cpu().numpy()
    L.check(lib.cmpc_rollout(C.byref(dims), C.byref(mdl), _ptr(Xd), _ptr(Ud), _ptr(cp), _ptr(ca), _ptr(f), st), lib)
    return f[0].cpu().numpy()


def lqr_gains_covs_batched(batch, X, U, Q, R, cov_w, cov_eta, want_covs=True):
    """cmpc_lqr_covs for a ProblemBatch along the trajectories X [B,N+1,9], U [B,N,nu] (numpy or
    CUDA tensors): gains [B,N,nu,9] and covs [B,N+1,9,9] as CUDA tensors
    (centroidal_model.py:215-227,233-238,284-285 of the reference)."""
    torch = _torch_cuda()
    lib = L.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    B, N, nc, nu = batch.B, batch.N, batch.nc, batch.nu
    wrench = batch.proto["robot"] == "TALOS"

    def dv(a, dt):
        t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
        return t.to(device=dev, dtype=dt).contiguous()
    Xd, Ud = dv(X, torch.float64), dv(U, torch.float64)
    if tuple(Xd.shape) != (B, N + 1, 9) or tuple(Ud.shape) != (B, N, nu):
        raise L.CmpcError("X must be [B,N+1,9] and U [B,N,nu]")
    cp, ca = dv(batch.contact_pos, torch.float64), dv(batch.contact_active, torch.int32)
    dims = L.cmpc_dims(B, N, nc, 1 if batch.shared_plan else 0, L.contact_model_of(batch.proto["robot"]))
    mdl = L.make_model_struct(batch.proto)
    w = L.make_lqr_struct(Q, R, cov_w, cov_eta, nu)
    gains = torch.empty((B, N, nu, 9), dtype=torch.float64, device=dev)
    covs = torch.empty((B, N + 1, 9, 9), dtype=torch.float64, device=dev) if want_covs else None
    scratch = torch.empty(L.LQR_SCRATCH_BYTES, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    if wrench:
        cR = dv(batch.contact_R, torch.float64)
        L.check(lib.cmpc_lqr_covs_wrench(C.byref(dims), C.byref(mdl), C.byref(w), _ptr(Xd), _ptr(Ud), _ptr(cp), _ptr(cR), _ptr(ca),
                                         _ptr(gains), _ptr(covs), _ptr(scratch), C.c_void_p(st.cuda_stream)), lib)
    else:
        L.check(lib.cmpc_lqr_covs(C.byref(dims), C.byref(mdl), C.byref(w), _ptr(Xd), _ptr(Ud), _ptr(cp), _ptr(ca),
                                  _ptr(gains), _ptr(covs), _ptr(scratch), C.c_void_p(st.cuda_stream)), lib)
    st.synchronize()   # `w` (host) and `scratch` must outlive the asynchronous copy
    return gains, covs


def lqr_gains_covs(model, traj_tuple):
    """LQR_gains (N,nu,9) and Covs (N+1,9,9) of one model along traj_tuple, as numpy arrays."""
    X = np.asarray(traj_tuple["state"], dtype=np.float64).T[None]
    U = np.asarray(traj_tuple["control"], dtype=np.float64).T[None]
    g, c = lqr_gains_covs_batched(ProblemBatch([model]), X, U, model._Q, model._R, model._Cov_w, model._Cov_eta)
    return g[0].cpu().numpy(), c[0].cpu().numpy()


def chance_constraint_xi(beta_u):
    """xi of construct_friction_pyramid_constraints (constraints.py:157): the pyramid matrix has
    5 rows (utils.py:9-16), xi = Phi^-1(1 - beta_u / 5 * 3)."""
    from scipy.stats import norm
    return float(norm.ppf(1 - (beta_u / 5 * 3)))


def friction_backoffs_batched(batch, gains, covs, beta_u):
    """cmpc_friction_backoffs: friction-row upper bounds of the stochastic mode [B,N,nc,4] (CUDA
    tensor) from the CUDA tensors of lqr_gains_covs_batched (constraints.py:157-163,187-214)."""
    torch = _torch_cuda()
    lib = L.load()
    dev = gains.device
    B, N, nc = batch.B, batch.N, batch.nc
    ca = torch.from_numpy(batch.contact_active).to(dev)
    cR = None if batch.contact_R is None else torch.from_numpy(batch.contact_R).to(dev)
    dims = L.cmpc_dims(B, N, nc, 1 if batch.shared_plan else 0)
    mdl = L.make_model_struct(batch.proto)
    ub = torch.empty((B, N, nc, 4), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream()
        L.check(lib.cmpc_friction_backoffs(C.byref(dims), C.byref(mdl), chance_constraint_xi(beta_u),
                                           _ptr(gains.contiguous()), _ptr(covs.contiguous()), _ptr(cR), _ptr(ca),
                                           _ptr(ub), C.c_void_p(st.cuda_stream)), lib)
        st.synchronize()
    return ub


def friction_backoffs(model, traj_tuple=None):
    """Upper bounds (N, nc, 4) of one model's friction rows in stochastic mode, along traj_tuple
    (default: the warm start, as solve_scp linearises there)."""
    traj_tuple = model._init_trajectories if traj_tuple is None else traj_tuple
    batch = ProblemBatch([model])
    X = np.asarray(traj_tuple["state"], dtype=np.float64).T[None]
    U = np.asarray(traj_tuple["control"], dtype=np.float64).T[None]
    g, c = lqr_gains_covs_batched(batch, X, U, model._Q, model._R, model._Cov_w, model._Cov_eta)
    return friction_backoffs_batched(batch, g, c, model._beta_u)[0].cpu().numpy()


def compute_trajectory_data(model, traj_tuple):
    """dict(dynamics (9,N), gradients{f_x (N,9,9), f_u (N,9,nu), f_w (N,9,nw)}, LQR_gains (N,nu,9),
    Covs (N+1,9,9), Covs_gradients{Cov_dx (N+1,9,9,9,N+1), Cov_du (N+1,9,9,nu,N+1)}) like the reference's
    compute_trajectory_data (/root/reference/src/centroidal_model.py:257-291).
    f_w = df/dp, the contact-position Jacobian the covariance propagation uses (:233-238): rows 6..8 of
    the block of contact i are -dt [f_i]x for an active contact, zero otherwise.  The covariance-gradient
    tensors are identically zero in the reference (nothing ever writes them after
    the jnp.zeros of :266-267); they are returned as zeros of the reference's shapes unless
    ``model._SKIP_COV_GRADIENTS`` is set (they take 8 N^2 (9+nu) 81 bytes)."""
    f, fx, fu = _lin_call(model, traj_tuple["state"], traj_tuple["control"], True)
    gains, covs = lqr_gains_covs(model, traj_tuple)
    N, nu = model._N, model._n_u
    U = np.asarray(traj_tuple["control"], dtype=np.float64)
    logic = model._contact_data["contacts_logic"]
    nc = logic.shape[1]
    npc = model._n_u_per_contact
    f0 = 2 if model._robot == "TALOS" else 0
    fw = np.zeros((N, 9, 3 * nc))
    for k in range(N):
        for c in range(nc):
            if logic[k, c]:
                fx_, fy_, fz_ = U[npc * c + f0:npc * c + f0 + 3, k]
                fw[k, 6:9, 3 * c:3 * c + 3] = -model._dt * np.array([[0.0, -fz_, fy_], [fz_, 0.0, -fx_], [-fy_, fx_, 0.0]])
    out = dict(dynamics=f.T.copy(), gradients={"f_x": fx, "f_u": fu, "f_w": fw}, LQR_gains=gains, Covs=covs)
    if not getattr(model, "_SKIP_COV_GRADIENTS", False):
        out["Covs_gradients"] = {"Cov_dx": np.zeros((N + 1, 9, 9, 9, N + 1)), "Cov_du": np.zeros((N + 1, 9, 9, nu, N + 1))}
    return out


def integrate_dynamics_trajectory(model, traj_tuple):
    """(9, N+1) array of one-step predictions; the last column (never read by the reference,
    scp_solver.py:82-86) repeats column N-1."""
    f = _lin_call(model, traj_tuple["state"], traj_tuple["control"], False).T
    return np.concatenate([f, f[:, -1:]], axis=1)


def integrate_one_step(model, x, u, contacts_position_all, contacts_logic_all, contacts_orientation_all):
    torch = _torch_cuda()
    lib = L.load()
    prob = model.problem_arrays()
    nc = prob["contact_active"].shape[1]
    dev = torch.device("cuda", torch.cuda.current_device())
    X = np.zeros((1, 2, 9)); X[0, 0] = np.asarray(x, dtype=np.float64)
    Xd = torch.from_numpy(X).to(dev)
    Ud = torch.from_numpy(np.asarray(u, dtype=np.float64).reshape(1, 1, -1).copy()).to(dev)
    cp = torch.from_numpy(np.asarray(contacts_position_all, dtype=np.float64).reshape(1, 1, nc, 3).copy()).to(dev)
    ca = torch.from_numpy(np.asarray(contacts_logic_all).astype(np.int32).reshape(1, 1, nc).copy()).to(dev)
    dims = L.cmpc_dims(1, 1, nc, 1)
    mdl = L.make_model_struct(prob)
    f = torch.empty((1, 1, 9), dtype=torch.float64, device=dev)
    L.check(lib.cmpc_rollout(C.byref(dims), C.byref(mdl), _ptr(Xd), _ptr(Ud), _ptr(cp), _ptr(ca), _ptr(f),
                             C.c_void_p(torch.cuda.current_stream().cuda_stream)), lib)
    return f[0, 0].cpu().numpy()

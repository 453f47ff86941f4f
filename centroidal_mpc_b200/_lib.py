"""ctypes binding of libcmpc_b200.so (include/cmpc.h).  There is no fallback: if the shared
library is missing or no CUDA device is present, the entry points raise."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CMPC_B200_LIB: another build of the same library (kernel experiments); never a different backend
LIB_PATH = os.environ.get("CMPC_B200_LIB") or os.path.join(_HERE, "csrc", "libcmpc_b200.so")


class cmpc_dims(C.Structure):
    _fields_ = [("batch", C.c_int32), ("N", C.c_int32), ("nc", C.c_int32), ("shared_plan", C.c_int32),
                ("contact_model", C.c_int32)]   # 0 point contacts, 1 CoP / wrench (TALOS)


CONTACT_POINT, CONTACT_WRENCH = 0, 1


def contact_model_of(robot):
    return CONTACT_WRENCH if robot == "TALOS" else CONTACT_POINT


class cmpc_model(C.Structure):
    _fields_ = [("robot_mass", C.c_double), ("gravity_constant", C.c_double), ("dt", C.c_double),
                ("mu", C.c_double), ("state_cost_weights", C.c_double * 9),
                ("control_cost_weights", C.c_double * 12), ("foot_range", C.c_double * 4)]


class cmpc_scp_params(C.Structure):
    _fields_ = [("trust_region_radius0", C.c_double), ("omega0", C.c_double), ("omega_max", C.c_double),
                ("rho0", C.c_double), ("rho1", C.c_double), ("beta_succ", C.c_double),
                ("beta_fail", C.c_double), ("gamma_fail", C.c_double),
                ("convergence_threshold", C.c_double), ("max_iterations", C.c_int32)]


class cmpc_qp_settings(C.Structure):
    _fields_ = [("eps_abs", C.c_double), ("eps_rel", C.c_double), ("sigma", C.c_double),
                ("alpha", C.c_double), ("rho", C.c_double), ("delta", C.c_double),
                ("adaptive_rho_tolerance", C.c_double), ("max_iter", C.c_int32),
                ("check_termination", C.c_int32), ("polish", C.c_int32),
                ("polish_refine_iter", C.c_int32), ("adaptive_rho", C.c_int32),
                ("adaptive_rho_start", C.c_int32), ("polish_active_set_rounds", C.c_int32),
                ("active_set_start", C.c_int32), ("active_set_step", C.c_int32),
                ("active_set_tol", C.c_double), ("warm_start_tol", C.c_double), ("warm_start", C.c_int32)]


class cmpc_lqr_weights(C.Structure):
    _fields_ = [("Q", C.c_double * 81), ("R", C.c_double * 144), ("cov_w", C.c_double * 144),
                ("cov_eta", C.c_double * 81)]


LQR_SCRATCH_BYTES = 4096   # CMPC_LQR_SCRATCH_BYTES


EXPORTS = ["cmpc_debug_profile", "cmpc_default_qp_settings", "cmpc_create", "cmpc_destroy", "cmpc_workspace_bytes",
           "cmpc_set_problem", "cmpc_set_friction_ub", "cmpc_solve_scp", "cmpc_solve_scp_host", "cmpc_get_stats",
           "cmpc_linearize", "cmpc_rollout", "cmpc_linearize_wrench", "cmpc_lqr_covs", "cmpc_lqr_covs_wrench", "cmpc_friction_backoffs", "cmpc_fp64_peak", "cmpc_launch_count",
           "cmpc_last_error", "cmpc_version", "cmpc_build_id", "cmpc_peer_alloc", "cmpc_peer_open", "cmpc_peer_close",
           "cmpc_peer_free"]

_lib = None


class CmpcError(RuntimeError):
    pass


def make_model_struct(prob):
    m = cmpc_model()
    m.robot_mass, m.gravity_constant, m.dt, m.mu = prob["m"], prob["g"], prob["dt"], prob["mu"]
    import numpy as np
    wx = np.diag(np.asarray(prob["state_cost_weights"], dtype=float))
    wu = np.diag(np.asarray(prob["control_cost_weights"], dtype=float))
    Wx = np.asarray(prob["state_cost_weights"], dtype=float)
    Wu = np.asarray(prob["control_cost_weights"], dtype=float)
    if not (np.allclose(Wx, np.diag(wx)) and np.allclose(Wu, np.diag(wu))):
        raise CmpcError("only diagonal cost weights are supported (all reference configs are diagonal)")
    for i in range(9):
        m.state_cost_weights[i] = wx[i]
    for i in range(12):
        m.control_cost_weights[i] = wu[i] if i < len(wu) else 1.0
    if prob["robot"] == "TALOS":   # conf.robot_foot_range: -x[1] <= cop_x <= x[0], -y[1] <= cop_y <= y[0]
        fr = prob["foot_range"]
        m.foot_range[0], m.foot_range[1] = float(fr["x"][0]), float(fr["x"][1])
        m.foot_range[2], m.foot_range[3] = float(fr["y"][0]), float(fr["y"][1])
    return m


def make_lqr_struct(Q, R, cov_w, cov_eta, nu):
    """conf.Q, conf.R, conf.cov_w, conf.cov_white_noise -> cmpc_lqr_weights (dense, row-major)."""
    import numpy as np
    w = cmpc_lqr_weights()
    cov_w = np.asarray(cov_w, dtype=np.float64)
    if cov_w.shape[0] < nu and cov_w.shape == (cov_w.shape[0],) * 2:
        # wrench model: three position-noise components per foot (n_w = 3 nc < n_u = 6 nc): leading block of n_u x n_u
        full = np.zeros((nu, nu))
        full[:cov_w.shape[0], :cov_w.shape[0]] = cov_w
        cov_w = full
    for name, a, n in (("Q", Q, 9), ("R", R, nu), ("cov_w", cov_w, nu), ("cov_eta", cov_eta, 9)):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
        if a.shape != (n, n):
            raise CmpcError("%s must be %d x %d, got %r" % (name, n, n, a.shape))
        dst = getattr(w, name)
        flat = a.ravel()
        for i in range(n * n):
            dst[i] = flat[i]
    return w


def make_scp_struct(scp_params):
    s = cmpc_scp_params()
    s.trust_region_radius0 = float(scp_params["trust_region_radius0"])
    s.omega0 = float(scp_params["omega0"])
    s.omega_max = float(scp_params["omega_max"])
    s.rho0, s.rho1 = float(scp_params["rho0"]), float(scp_params["rho1"])
    s.beta_succ, s.beta_fail = float(scp_params["beta_succ"]), float(scp_params["beta_fail"])
    s.gamma_fail = float(scp_params["gamma_fail"])
    s.convergence_threshold = float(scp_params["convergence_threshold"])
    s.max_iterations = int(scp_params["max_iterations"])
    return s


def make_qp_struct(overrides=None, lib=None):
    q = cmpc_qp_settings()
    if lib is not None:
        lib.cmpc_default_qp_settings(C.byref(q))
    else:   # same numbers as csrc/cmpc_params.h default_qp_settings (checked by the tests)
        q.eps_abs = q.eps_rel = 1e-7
        q.sigma, q.alpha, q.rho, q.delta, q.adaptive_rho_tolerance = 1e-6, 1.8, 2.0, 1e-6, 5.0
        q.max_iter, q.check_termination, q.polish, q.polish_refine_iter, q.adaptive_rho = 4000, 25, 1, 3, 1
        q.adaptive_rho_start = 200
        q.polish_active_set_rounds = 9
        q.active_set_start, q.active_set_step, q.active_set_tol = 8, 8, 1e-9
        q.warm_start_tol, q.warm_start = 1e-7, 0
    for k, v in (overrides or {}).items():
        if not hasattr(q, k):
            raise CmpcError("unknown QP setting %r" % k)
        setattr(q, k, v)
    return q


def load():
    """Load libcmpc_b200.so.  Raises CmpcError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CmpcError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, ip, dp = C.c_void_p, C.c_void_p, C.c_void_p
    lib.cmpc_default_qp_settings.argtypes = [C.POINTER(cmpc_qp_settings)]
    lib.cmpc_default_qp_settings.restype = None
    lib.cmpc_create.argtypes = [C.POINTER(cmpc_dims), C.POINTER(C.c_void_p)]
    lib.cmpc_destroy.argtypes = [C.c_void_p]
    lib.cmpc_workspace_bytes.argtypes = [C.c_void_p]
    lib.cmpc_workspace_bytes.restype = C.c_int64
    lib.cmpc_set_problem.argtypes = [C.c_void_p, C.POINTER(cmpc_model)] + [dp] * 7
    lib.cmpc_set_friction_ub.argtypes = [C.c_void_p, dp]
    lib.cmpc_solve_scp.argtypes = [C.c_void_p, C.POINTER(cmpc_scp_params), C.POINTER(cmpc_qp_settings),
                                   dp, dp, ip, ip, ip, vp]
    lib.cmpc_solve_scp_host.argtypes = [C.c_void_p, C.POINTER(cmpc_model), C.POINTER(cmpc_scp_params),
                                        C.POINTER(cmpc_qp_settings)] + [dp] * 7 + [dp, dp, ip, ip, ip]
    lib.cmpc_get_stats.argtypes = [C.c_void_p, ip, ip, dp, vp]
    lib.cmpc_linearize.argtypes = [C.POINTER(cmpc_dims), C.POINTER(cmpc_model), dp, dp, dp, ip, dp, dp, dp, vp]
    lib.cmpc_rollout.argtypes = [C.POINTER(cmpc_dims), C.POINTER(cmpc_model), dp, dp, dp, ip, dp, vp]
    lib.cmpc_linearize_wrench.argtypes = [C.POINTER(cmpc_dims), C.POINTER(cmpc_model), dp, dp, dp, dp, ip, dp, dp, dp, vp]
    lib.cmpc_lqr_covs.argtypes = [C.POINTER(cmpc_dims), C.POINTER(cmpc_model), C.POINTER(cmpc_lqr_weights),
                                  dp, dp, dp, ip, dp, dp, vp, vp]
    lib.cmpc_lqr_covs_wrench.argtypes = [C.POINTER(cmpc_dims), C.POINTER(cmpc_model), C.POINTER(cmpc_lqr_weights),
                                         dp, dp, dp, dp, ip, dp, dp, vp, vp]
    lib.cmpc_friction_backoffs.argtypes = [C.POINTER(cmpc_dims), C.POINTER(cmpc_model), C.c_double,
                                           dp, dp, dp, ip, dp, vp]
    lib.cmpc_peer_alloc.argtypes = [C.c_int64, C.POINTER(C.c_void_p), C.c_char_p]
    lib.cmpc_peer_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.cmpc_peer_close.argtypes = [C.c_void_p]
    lib.cmpc_peer_free.argtypes = [C.c_void_p]
    lib.cmpc_fp64_peak.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.cmpc_launch_count.restype = C.c_int64
    lib.cmpc_last_error.restype = C.c_char_p
    lib.cmpc_version.restype = C.c_char_p
    lib.cmpc_build_id.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc, lib=None):
    if rc != 0:
        msg = (lib or load()).cmpc_last_error()
        raise CmpcError("libcmpc_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))

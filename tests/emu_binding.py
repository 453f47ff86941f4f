"""ctypes access to the TEST-ONLY host build of the device solver (tests/emu/cmpc_emu.cpp).
Used by the CPU test-suite to exercise the kernel logic without a GPU; never imported by the
product package."""
import ctypes as C
import os
import subprocess

import numpy as np

from centroidal_mpc_b200 import _lib as L

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "emu", "cmpc_emu.cpp")
_SO = os.path.join(_HERE, "emu", "libcmpc_emu.so")
_DEPS = [os.path.join(_HERE, "..", "centroidal_mpc_b200", "csrc", f)
         for f in ("cmpc_core.cuh", "cmpc_tile.cuh", "cmpc_lqr.cuh", "cmpc_params.h")] + [_SRC]
_libs = {}


def build(force=False, team_lanes=1, wrench=False):
    """team_lanes = 1: one host thread of control per instance.  team_lanes = 8: the lock-step build, in
    which the 8 lanes of an instance's team run as coroutines that switch at every team_sync, i.e. the
    work split, the shared-memory exchanges and the synchronisation points of the CUDA kernel."""
    so = _SO if team_lanes == 1 else _SO.replace(".so", "_nl%d.so" % team_lanes)
    if wrench:   # the CoP / wrench contact model: the same source compiled with CMPC_WRENCH=1 (csrc/cmpc_wrench.cu)
        so = so.replace(".so", "_wr.so")
    newest = max(os.path.getmtime(f) for f in _DEPS)
    if force or not os.path.exists(so) or os.path.getmtime(so) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas", "-shared", "-fPIC",
                               "-DCMPC_NL=%d" % team_lanes, "-DCMPC_WRENCH=%d" % (1 if wrench else 0), "-x", "c++", "-o",
                               so + ".tmp", _SRC])
        os.replace(so + ".tmp", so)
    return so


def load(team_lanes=1, wrench=False):
    key = (team_lanes, bool(wrench))
    if key not in _libs:
        _libs[key] = C.CDLL(build(team_lanes=team_lanes, wrench=wrench))
        assert _libs[key].cmpc_emu_team_lanes() == team_lanes
    return _libs[key]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def solve_scp(batch, scp_params, qp_overrides=None, friction_ub=None, team_lanes=1):
    """Run the host build of the device solver on a ProblemBatch; returns a dict of arrays.
    ``friction_ub`` [B,N,nc,4]: stochastic mode (upper bounds of the friction rows)."""
    lib = load(team_lanes, wrench=getattr(batch, "wrench", False))
    lib.cmpc_emu_set_friction_ub.argtypes = [C.c_void_p]
    lib.cmpc_emu_set_friction_ub.restype = None
    fub = None if friction_ub is None else np.ascontiguousarray(friction_ub, dtype=np.float64)
    lib.cmpc_emu_set_friction_ub(_p(fub))
    try:
        return _solve_scp(lib, batch, scp_params, qp_overrides)
    finally:
        lib.cmpc_emu_set_friction_ub(None)


def _solve_scp(lib, batch, scp_params, qp_overrides):
    B, N, nu = batch.B, batch.N, batch.nu
    dims = L.cmpc_dims(B, N, batch.nc, 1 if batch.shared_plan else 0, L.contact_model_of(batch.proto["robot"]))
    model = L.make_model_struct(batch.proto)
    scp = L.make_scp_struct(scp_params)
    qp = L.make_qp_struct(qp_overrides)
    out = dict(X=np.zeros((B, N + 1, 9)), U=np.zeros((B, N, nu)), scp_iters=np.zeros(B, np.int32),
               status=np.zeros(B, np.int32), n_accepted=np.zeros(B, np.int32),
               qp_iters=np.zeros(B, np.int32), n_factor=np.zeros(B, np.int32), info=np.zeros((B, 12)))
    rc = lib.cmpc_emu_solve_scp(C.byref(dims), C.byref(model), C.byref(scp), C.byref(qp), _p(batch.x_init),
                                _p(batch.x_final), _p(batch.X_ref), _p(batch.U_init), _p(batch.contact_pos),
                                _p(batch.contact_R), _p(batch.contact_active), _p(out["X"]), _p(out["U"]),
                                _p(out["scp_iters"]), _p(out["status"]), _p(out["n_accepted"]),
                                _p(out["qp_iters"]), _p(out["n_factor"]), _p(out["info"]))
    if rc != 0:
        raise RuntimeError("cmpc_emu_solve_scp returned %d" % rc)
    return out


def lqr_covs(batch, X, U, Q, R, cov_w, cov_eta):
    """Host build of csrc/cmpc_lqr.cuh: gains [B,N,nu,9], covs [B,N+1,9,9]."""
    wrench = getattr(batch, "wrench", False)
    lib = load(wrench=wrench)
    B, N, nu = batch.B, batch.N, batch.nu
    dims = L.cmpc_dims(B, N, batch.nc, 1 if batch.shared_plan else 0, L.contact_model_of(batch.proto["robot"]))
    model = L.make_model_struct(batch.proto)
    w = L.make_lqr_struct(Q, R, cov_w, cov_eta, nu)
    X = np.ascontiguousarray(X, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    gains, covs = np.zeros((B, N, nu, 9)), np.zeros((B, N + 1, 9, 9))
    rc = lib.cmpc_emu_lqr_covs(C.byref(dims), C.byref(model), C.byref(w), _p(X), _p(U), _p(batch.contact_pos),
                               _p(batch.contact_active), _p(gains), _p(covs), _p(batch.contact_R) if wrench else None)
    if rc != 0:
        raise RuntimeError("cmpc_emu_lqr_covs returned %d" % rc)
    return gains, covs


def friction_backoffs(batch, xi, gains, covs):
    """Host build of friction_backoff_knot: ub [B,N,nc,4]."""
    lib = load()
    lib.cmpc_emu_friction_backoffs.argtypes = [C.c_void_p, C.c_void_p, C.c_double] + [C.c_void_p] * 5
    B, N, nc = batch.B, batch.N, batch.nc
    dims = L.cmpc_dims(B, N, nc, 1 if batch.shared_plan else 0)
    model = L.make_model_struct(batch.proto)
    ub = np.zeros((B, N, nc, 4))
    rc = lib.cmpc_emu_friction_backoffs(C.byref(dims), C.byref(model), float(xi), _p(np.ascontiguousarray(gains)),
                                        _p(np.ascontiguousarray(covs)), _p(batch.contact_R),
                                        _p(batch.contact_active), _p(ub))
    if rc != 0:
        raise RuntimeError("cmpc_emu_friction_backoffs returned %d" % rc)
    return ub

"""Host-side mirror of the reference's problem objects (no GPU, no oracle arithmetic)."""
import ctypes
import os
import re

import numpy as np
import pytest

from centroidal_mpc_b200 import synthetic
from centroidal_mpc_b200.batch import ProblemBatch
from centroidal_mpc_b200.src import contact_plan, optimizer
from centroidal_mpc_b200.src.centroidal_model import Centroidal_model
from centroidal_mpc_b200.src.scp_solver import get_QP_solution, interpolate_SCP_solution

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,N", [("solo12_trot", 165), ("solo12_pace", 107), ("solo12_bound", 165), ("bolt", 122)])
def test_shipped_horizons(name, N):
    """N = 165 / 107 / 165 (solo12), 122 (bolt) from the gait tables."""
    assert synthetic.load_conf(name).N == N


def test_contact_sequence_trot():
    conf = synthetic.load_conf("solo12_trot")
    seq = conf.contact_sequence
    assert len(seq) == 4 * 4 + 1
    assert [d.CONTACT for d in seq[0]] == ["FR", "FL", "HR", "HL"] and all(d.ACTIVE for d in seq[0])
    assert [d.ACTIVE for d in seq[1]] == [False, True, True, False]      # rflhStep: FR and HL swing
    assert [d.ACTIVE for d in seq[3]] == [True, False, False, True]      # lfrhStep
    # swinging feet advance by stepLength after the step
    fr0 = seq[0][0].pose.translation[0]
    assert np.isclose(seq[2][0].pose.translation[0], fr0 + conf.gait["stepLength"])
    traj = contact_plan.create_contact_trajectory(conf)
    assert all(len(v) == 165 for v in traj.values())
    assert [d.idx for d in seq[0]] == [0, 1, 2, 3]
    np.testing.assert_array_equal(seq[0][0].pose.rotation, np.eye(3))


def test_debris_rotation_is_rodrigues():
    d = contact_plan.Debris("FR", x=0.1, y=0.2, z=0.0, axis=[1, 0], angle=0.3, ACTIVE=True)
    c, s = np.cos(0.3), np.sin(0.3)
    np.testing.assert_allclose(d.pose.rotation, [[1, 0, 0], [0, c, -s], [0, s, c]], atol=1e-15)


def test_model_attributes_and_layout():
    conf = synthetic.load_conf("solo12_trot", N=20)
    m = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, 0))
    N = 20
    assert m._total_nb_optimizers == 9 * (N + 1) + 12 * N + (N + 1) + N == 23 * N + 10
    cd = m._contact_data
    assert cd["contacts_logic"].shape == (N, 4) and cd["contacts_orient"].shape == (N, 4, 3, 3)
    assert cd["contacts_position"].shape == (N, 12)
    k = 7      # FR swings at knot 7 in a trot
    assert list(cd["contacts_logic"][k]) == [0, 1, 1, 0]
    assert np.all(cd["contacts_position"][k, 0:3] == 0) and np.all(cd["contacts_orient"][k, 0] == 0)
    U = m._init_trajectories["control"]
    assert np.allclose(U[3:6, k], [1e-3, 1e-3, 2.5 * 9.81 / 2]) and np.all(U[0:3, k] == 0)
    np.testing.assert_array_equal(m._x_init, m._init_trajectories["state"][:, 0])
    np.testing.assert_array_equal(m._x_final, m._init_trajectories["state"][:, N])
    # index objects
    so = m._state_optimizers_indices
    assert so["ang_moms"][1]._optimizer_idx_vector[3] == 3 * 9 + 6
    fz_hl = m._control_optimizers_indices["HL"]["forces"][2]
    assert fz_hl._optimizer_idx_vector[5] == 9 * (N + 1) + 5 * 12 + 3 * 3 + 2
    sl = m._state_slack_optimizers_indices
    assert sl._slack_optimizers_idx_vector[4] == 9 * (N + 1) + 12 * N + 4
    expect = np.array([[(-1.0) ** (j // 2 ** i) for i in range(3)] for j in range(8)])
    np.testing.assert_array_equal(sl._penum_mat, expect)


def test_model_requires_warm_start_file_like_the_reference(tmp_path, monkeypatch):
    conf = synthetic.load_conf("solo12_trot", N=10)
    monkeypatch.chdir(tmp_path)
    with pytest.raises(FileNotFoundError):
        Centroidal_model(conf)
    traj = synthetic.reference_trajectory(conf, 0)
    np.savez("wholeBody_to_centroidal_traj.npz", X=traj)
    m = Centroidal_model(conf)
    np.testing.assert_array_equal(m._x_final, traj[-1])


def test_interpolation_matches_naive_loops():
    rng = np.random.default_rng(3)
    X, U = rng.normal(size=(9, 6)), rng.normal(size=(12, 5))
    out = interpolate_SCP_solution(dict(state=[X], control=[U]))

    def naive(M):
        R = np.zeros((M.shape[0], (M.shape[1] - 1) * 10))
        for i in range(M.shape[1] - 1):
            d = (M[:, i + 1] - M[:, i]) / 10.0
            for j in range(10):
                R[:, i * 10 + j] = M[:, i] + j * d
        return R
    np.testing.assert_array_equal(out["X"], naive(X))     # same arithmetic as the reference's loops
    np.testing.assert_array_equal(out["U"], naive(U))


def test_npz_handoff_files_have_the_reference_names_keys_and_layouts(tmp_path):
    from centroidal_mpc_b200.src.scp_solver import load_scp_handoff, save_scp_handoff
    rng = np.random.default_rng(4)
    N = 7
    X, U = rng.normal(size=(9, N + 1)), rng.normal(size=(12, N))
    sol = dict(state=[X * 0, X], control=[U * 0, U], gains=[None, None], covs=[None, None])
    p1, p2 = save_scp_handoff(sol, tmp_path)
    assert p1.endswith("scp_sol_interpol_nom.npz") and p2.endswith("centroidal_to_wholeBody_traj.npz")
    f1, f2 = np.load(p1), np.load(p2)
    assert sorted(f1.files) == ["U", "X"] and sorted(f2.files) == ["U", "X"]
    assert f1["X"].shape == (9, 10 * N) and f1["U"].shape == (12, 10 * (N - 1))
    np.testing.assert_array_equal(f2["X"], X)             # the LAST accepted iterate
    np.testing.assert_array_equal(f2["U"], U)
    Xr, Ur = load_scp_handoff(tmp_path)
    np.testing.assert_array_equal(Xr, X)
    np.testing.assert_array_equal(Ur, U)
    with pytest.raises(ValueError):
        save_scp_handoff(dict(state=[], control=[]), tmp_path)
    with pytest.raises(ValueError):
        save_scp_handoff(False, tmp_path)


def test_get_qp_solution_layout():
    conf = synthetic.load_conf("solo12_trot", N=5)
    m = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, 0))
    z = np.arange(m._total_nb_optimizers, dtype=float)
    s = get_QP_solution(m, z)
    assert s["state"][3, 2] == 2 * 9 + 3 and s["control"][4, 1] == 9 * 6 + 12 + 4


def test_problem_batch_packing(cases):
    conf, models = cases["solo12_trot"]
    b = ProblemBatch(models)
    assert b.shared_plan and b.identity_R and b.contact_R is None
    assert b.X_ref.shape == (3, conf.N + 1, 9) and b.U_init.shape == (3, conf.N, 12)
    assert b.contact_pos.shape == (1, conf.N, 4, 3) and b.contact_active.dtype == np.int32
    np.testing.assert_array_equal(b.X_ref[1], models[1]._init_trajectories["state"].T)
    vb = synthetic.make_batch(conf, 3)
    np.testing.assert_array_equal(vb.X_ref, b.X_ref)
    np.testing.assert_array_equal(vb.U_init, b.U_init)
    np.testing.assert_array_equal(vb.x_final, b.x_final)


def test_cabi_library_exports_every_declared_symbol():
    """include/cmpc.h vs the built shared library (no compute calls: there is no GPU here)."""
    import __graft_entry__ as g
    g.build()
    hdr = open(os.path.join(ROOT, "include", "cmpc.h")).read()
    declared = set(re.findall(r"\b(cmpc_[a-z0-9_]+)\s*\(", hdr))
    from centroidal_mpc_b200 import _lib
    assert declared == set(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    lib.cmpc_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.cmpc_version()
    # ctypes mirror of the settings struct agrees with the library's defaults
    q = _lib.cmpc_qp_settings()
    lib.cmpc_default_qp_settings(ctypes.byref(q))
    d = _lib.make_qp_struct()
    for name, _ in _lib.cmpc_qp_settings._fields_:
        assert getattr(q, name) == getattr(d, name), name


def test_product_path_has_no_cpu_fallback(cases):
    """Without a CUDA device the drop-in entry point must raise, not compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from centroidal_mpc_b200._lib import CmpcError
    from centroidal_mpc_b200.src.scp_solver import solve_scp
    conf, models = cases["solo12_trot"]
    with pytest.raises(CmpcError):
        solve_scp(models[0], conf.scp_params)
    src = ""
    pkg = os.path.join(ROOT, "centroidal_mpc_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src += open(os.path.join(dp, f)).read()
    assert "import oracle" not in src and "from oracle" not in src and "emu_binding" not in src


def test_batch_rejects_mixed_models():
    """One model struct and one mode per batch: mixing weights or nominal / stochastic models raises
    instead of silently solving everything with instance 0's settings."""
    import copy
    conf = synthetic.load_conf("solo12_trot", N=10)
    a = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, 0))
    conf2 = copy.copy(conf)
    conf2.state_cost_weights = conf.state_cost_weights * 2.0
    b = Centroidal_model(conf2, centroidal_traj=synthetic.reference_trajectory(conf, 1))
    with pytest.raises(ValueError, match="state_cost_weights"):
        ProblemBatch([a, b])
    c = Centroidal_model(conf, STOCHASTIC_OCP=True, centroidal_traj=synthetic.reference_trajectory(conf, 1))
    with pytest.raises(ValueError, match="STOCHASTIC_OCP"):
        ProblemBatch([a, c])
    assert ProblemBatch([a, Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, 2))]).B == 2

"""The CoP / wrench contact model (TALOS; /root/reference/src/centroidal_model.py:204-208,
src/constraints.py:111-145, src/optimizer.py:48-64): host build of the wrench compilation of the solver
(csrc/cmpc_wrench.cu = the solver source with CMPC_WRENCH=1) against the tightly solved oracle, plus the
properties the QP promises (CoP box, friction pyramid in the foot frame, exact zeros on a swinging foot,
linearised dynamics)."""
import numpy as np
import pytest

import emu_binding as E
from conftest import relerr
from centroidal_mpc_b200 import synthetic
from centroidal_mpc_b200.batch import ProblemBatch
from centroidal_mpc_b200.device import WRENCH_QP_DEFAULTS
from centroidal_mpc_b200.src.centroidal_model import Centroidal_model
from oracle import dynamics
from scenarios import n100_samples

TOL = 1e-6


def test_wrench_host_build_matches_tight_oracle():
    conf, full, sub, g = n100_samples("talos", "B", N=30)
    assert bool(np.all(g["ok"]))
    out = E.solve_scp(sub, conf.scp_params, qp_overrides=WRENCH_QP_DEFAULTS)
    assert (out["status"] == 0).all() and (out["n_accepted"] == 1).all()
    assert (out["info"][:, 10] == 1).all()          # every QP ended with the KKT certificate
    for j in range(sub.B):
        assert out["scp_iters"][j] == int(g["iterations"][j])
        assert relerr(out["X"][j].T, g["X"][j]) < TOL and relerr(out["U"][j].T, g["U"][j]) < TOL


def test_wrench_solution_properties():
    conf = synthetic.load_conf("talos", N=40)
    models = [Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, b)) for b in range(3)]
    batch = ProblemBatch(models)
    assert batch.wrench and batch.nu == 12 and batch.contact_R is not None
    out = E.solve_scp(batch, conf.scp_params, qp_overrides=WRENCH_QP_DEFAULTS)
    assert (out["status"] == 0).all()
    kf = conf.mu / np.sqrt(2.0)
    binds = 0
    for b, m in enumerate(models):
        p = m.problem_arrays()
        X, U = out["X"][b].T, out["U"][b].T            # (9, N+1), (12, N)
        act = p["contact_active"]
        for c in range(2):
            u = U[6 * c:6 * c + 6]
            assert np.all(u[:, act[:, c] == 0] == 0.0)                 # swinging foot: exact zeros
            on = act[:, c] == 1
            assert np.all(u[0, on] <= conf.lxp + 1e-8) and np.all(-u[0, on] <= conf.lxn + 1e-8)   # CoP box
            assert np.all(u[1, on] <= conf.lyp + 1e-8) and np.all(-u[1, on] <= conf.lyn + 1e-8)
            binds += int(np.sum(np.abs(np.abs(u[1, on]) - conf.lyp) < 1e-7))
            for k in np.nonzero(on)[0]:                                # friction pyramid in the foot frame
                fl = p["contact_R"][k, c].T @ u[2:5, k]
                assert abs(fl[0]) <= kf * fl[2] + 1e-6 and abs(fl[1]) <= kf * fl[2] + 1e-6
        # the solution satisfies the dynamics linearised about the warm start, x_0 and x_N
        td = dynamics.trajectory_data(p["X_ref"], p["U_init"], p)
        for k in range(conf.N):
            lin = td["dynamics"][:, k] + td["f_x"][k] @ (X[:, k] - p["X_ref"][:, k]) + td["f_u"][k] @ (U[:, k] - p["U_init"][:, k])
            np.testing.assert_allclose(X[:, k + 1], lin, rtol=0, atol=1e-8)
        np.testing.assert_allclose(X[:, 0], p["x_init"], atol=1e-12)
        np.testing.assert_allclose(X[:, -1], p["x_final"], atol=1e-7)
    assert binds > 0      # single support with the CoM beside the foot: the lateral CoP sits on the edge of the sole


def test_wrench_team_of_8_lanes_is_bitwise_equal():
    """Lock-step host build (the eight lanes of a team as coroutines): the work split, shared-memory exchanges
    and synchronisation points of the wrench kernel, bit for bit the single-lane result."""
    conf = synthetic.load_conf("talos", N=22)
    models = [Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, b)) for b in range(2)]
    batch = ProblemBatch(models)
    a = E.solve_scp(batch, conf.scp_params, qp_overrides=WRENCH_QP_DEFAULTS)
    b8 = E.solve_scp(batch, conf.scp_params, qp_overrides=WRENCH_QP_DEFAULTS, team_lanes=8)
    for k in ("X", "U", "scp_iters", "status", "qp_iters", "n_factor", "info"):
        np.testing.assert_array_equal(a[k], b8[k], err_msg=k)


def test_wrench_batch_validation():
    conf = synthetic.load_conf("talos", N=10)
    m = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, 0))
    p = m.problem_arrays()
    p3 = dict(p, contact_active=np.ones((10, 3), np.int32))
    with pytest.raises(ValueError):
        ProblemBatch([p3])


@pytest.mark.gpu
def test_wrench_model_on_device_matches_tight_oracle(gpu):
    """BASELINE.json configuration 5 (talos) at the benchmark horizon N = 100, the smallest batch of its sweep
    (256): 8 sampled instances at 1e-6 against the tightly solved oracle, solved inside the full batch; the first
    tiles bit for bit equal to the host build of the same source."""
    from centroidal_mpc_b200.device import BatchSolver
    conf, full, sub, g = n100_samples("talos", "B", N=100)
    assert bool(np.all(g["ok"]))
    solver = BatchSolver(full)
    out = solver.solve(conf.scp_params).results()
    st = solver.stats()
    assert (out["status"] == 0).all() and (out["n_accepted"] == 1).all()
    worst = 0.0
    for j, b in enumerate(np.asarray(g["ids"])):
        assert out["scp_iters"][b] == int(g["iterations"][j])
        worst = max(worst, relerr(out["X"][b].T, g["X"][j]), relerr(out["U"][b].T, g["U"][j]))
    assert worst < TOL, worst
    head = ProblemBatch.from_arrays(full.proto, full.x_init[:8], full.x_final[:8], full.X_ref[:8], full.U_init[:8])
    emu = E.solve_scp(head, conf.scp_params, qp_overrides=WRENCH_QP_DEFAULTS)
    np.testing.assert_array_equal(out["X"][:8], emu["X"])
    np.testing.assert_array_equal(out["U"][:8], emu["U"])
    np.testing.assert_array_equal(st["qp_iters"][:8], emu["qp_iters"])
    # the host entry point (pinned / pageable host buffers) gives the same answer
    host = solver.solve_host(conf.scp_params)
    np.testing.assert_array_equal(host["X"], out["X"])
    np.testing.assert_array_equal(host["U"], out["U"])
    solver.close()


@pytest.mark.gpu
def test_wrench_linearize_kernel(gpu):
    conf = synthetic.load_conf("talos", N=40)
    m = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, 1))
    prob = m.problem_arrays()
    rng = np.random.default_rng(7)
    traj = dict(state=prob["X_ref"] + 0.01 * rng.normal(size=prob["X_ref"].shape),
                control=prob["U_init"] + np.array([0.01, 0.01, 5.0, 5.0, 20.0, 1.0] * 2)[:, None] * rng.normal(size=prob["U_init"].shape))
    from centroidal_mpc_b200 import device
    f, fx, fu = device._lin_call(m, traj["state"], traj["control"], True)
    ref = dynamics.trajectory_data(traj["state"], traj["control"], prob)
    np.testing.assert_allclose(f.T, ref["dynamics"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(fx, ref["f_x"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(fu, ref["f_u"], rtol=0, atol=1e-12)
    roll = m.integrate_dynamics_trajectory(traj)
    np.testing.assert_allclose(roll[:, :conf.N], dynamics.rollout(traj["state"], traj["control"], prob), atol=1e-11)


def _perturbed_talos(N=30):
    conf = synthetic.load_conf("talos", N=N)
    m = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, 1))
    p = m.problem_arrays()
    rng = np.random.default_rng(3)
    X = p["X_ref"] + 0.01 * rng.normal(size=p["X_ref"].shape)
    U = p["U_init"] + np.array([0.01, 0.01, 5.0, 5.0, 20.0, 1.0] * 2)[:, None] * rng.normal(size=p["U_init"].shape)
    return conf, m, p, X, U


def test_wrench_lqr_gains_and_covs_host_build():
    """LQR_gains / Covs of compute_trajectory_data for robot == 'TALOS' (/root/reference/src/centroidal_model.py:
    215-238,284-285 with the six-control Jacobians; position noise: three components per foot)."""
    conf, m, p, X, U = _perturbed_talos()
    g, c = E.lqr_covs(ProblemBatch([m]), X.T[None], U.T[None], conf.Q, conf.R, conf.cov_w, conf.cov_white_noise)
    go, co = dynamics.lqr_gains_covs(X, U, p, conf.Q, conf.R, conf.cov_w, conf.cov_white_noise)
    assert np.abs(g[0] - go).max() <= 1e-10 * np.abs(go).max()
    assert np.abs(c[0] - co).max() <= 1e-8 * np.abs(co).max()


@pytest.mark.gpu
def test_wrench_trajectory_data_and_drop_in_solve(gpu):
    """compute_trajectory_data (all keys) and the drop-in solve_scp for a TALOS model through the C ABI."""
    from centroidal_mpc_b200.src.scp_solver import solve_scp
    conf, m, p, X, U = _perturbed_talos()
    td = m.compute_trajectory_data(dict(state=X, control=U))
    ref = dynamics.trajectory_data(X, U, p)
    go, co = dynamics.lqr_gains_covs(X, U, p, conf.Q, conf.R, conf.cov_w, conf.cov_white_noise)
    np.testing.assert_allclose(td["dynamics"], ref["dynamics"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(td["gradients"]["f_u"], ref["f_u"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(td["gradients"]["f_w"], ref["f_w"], rtol=0, atol=1e-12)
    assert np.abs(td["LQR_gains"] - go).max() <= 1e-10 * np.abs(go).max()
    assert np.abs(td["Covs"] - co).max() <= 1e-8 * np.abs(co).max()
    sol = solve_scp(m, conf.scp_params)
    assert sol is not False and len(sol["state"]) == 1
    assert sol["state"][-1].shape == (9, conf.N + 1) and sol["control"][-1].shape == (12, conf.N)
    assert sol["gains"][-1].shape == (conf.N, 12, 9) and sol["covs"][-1].shape == (conf.N + 1, 9, 9)


@pytest.mark.gpu
def test_wrench_full_size_properties(gpu):
    """BASELINE configuration 5 at the size the benchmark runs it (4096 x N = 100): every instance accepted and
    certified, CoP box and friction pyramid hold, swinging feet carry exact zeros, boundary states are met."""
    from centroidal_mpc_b200.device import BatchSolver
    conf = synthetic.load_conf("talos", N=100)
    batch = synthetic.make_batch(conf, 4096)
    solver = BatchSolver(batch)
    out = solver.solve(conf.scp_params).results()
    st = solver.stats()
    solver.close()
    assert (out["status"] == 0).all() and (out["n_accepted"] == 1).all() and (out["scp_iters"] == 1).all()
    assert (st["info"][:, 10] == 1).all() and (st["qp_iters"] == 8).all()
    U, X = out["U"], out["X"]
    act = batch.contact_active[0]                     # shared plan [N, 2]
    kf = conf.mu / np.sqrt(2.0)
    for c in range(2):
        on = act[:, c] == 1
        u = U[:, :, 6 * c:6 * c + 6]
        assert np.all(u[:, ~on] == 0.0)
        assert np.all(np.abs(u[:, on, 0]) <= conf.lxp + 1e-7) and np.all(np.abs(u[:, on, 1]) <= conf.lyp + 1e-7)
        fz = u[:, on, 4]                              # identity foot frames in this gait
        assert np.all(np.abs(u[:, on, 2]) <= kf * fz + 1e-5) and np.all(np.abs(u[:, on, 3]) <= kf * fz + 1e-5)
    np.testing.assert_allclose(X[:, 0], batch.x_init, atol=1e-12)
    np.testing.assert_allclose(X[:, -1], batch.x_final, atol=1e-6)


def test_wrench_rotated_foot_frames_match_tight_oracle():
    """Yawed and pitched soles and a warm start with nonzero CoP / tau_z: every block of the wrench Jacobian
    (r1 x f, r2 x f, r3, lever arm p + R cop - c) and the friction pyramid in the foot frame are exercised
    (the synthetic gait has identity frames and a zero-CoP warm start)."""
    from oracle import scp as oscp
    conf = synthetic.load_conf("talos", N=24)
    m = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, 2))
    p = dict(m.problem_arrays())

    def Rz(a): return np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    def Ry(a): return np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    R = p["contact_R"].copy()
    for c, (yaw, pitch) in enumerate(((0.3, 0.05), (-0.2, -0.04))):
        R[:, c] = (Rz(yaw) @ Ry(pitch))[None] * p["contact_active"][:, c, None, None]
    p["contact_R"] = R
    U, rng = p["U_init"].copy(), np.random.default_rng(0)
    U[[0, 1, 6, 7]] += 0.01 * rng.normal(size=(4, conf.N)) * p["contact_active"].T[[0, 0, 1, 1]]
    U[[5, 11]] += 0.5 * rng.normal(size=(2, conf.N)) * p["contact_active"].T
    p["U_init"] = U
    out = E.solve_scp(ProblemBatch([p]), conf.scp_params, qp_overrides=WRENCH_QP_DEFAULTS)
    ref = oscp.solve_scp(p, conf.scp_params, osqp_settings=dict(eps_abs=1e-9, eps_rel=1e-9, max_iter=100000, polish_refine_iter=30))
    assert ref is not False and len(ref["state"]) == 1 and out["status"][0] == 0 and out["n_accepted"][0] == 1
    assert out["scp_iters"][0] == ref["iterations"]
    assert relerr(out["X"][0].T, ref["state"][-1]) < TOL and relerr(out["U"][0].T, ref["control"][-1]) < TOL

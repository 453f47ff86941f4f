"""GPU parity tests: the CUDA path through the C ABI (libcmpc_b200.so) against the oracle, the
golden fixtures, the host build of the same source, and size-independent properties at the
benchmark's full size.  Tolerance: 1e-6 norm-wise relative on (X, U) (BASELINE.json north_star),
equal SCP iteration counts."""
import os

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

TOL = 1e-6


def test_smoke(gpu):
    import __graft_entry__ as g
    g.smoke()


def test_golden_through_cabi(gpu):
    from centroidal_mpc_b200.src.scp_solver import solve_scp_batched
    from test_emu_parity import check_against_golden
    from test_oracle import _golden_files, load_golden
    for path in _golden_files():
        g, conf, sp, model = load_golden(path)
        out = solve_scp_batched([model], sp)
        if bool(g["returned_false"]):
            assert out["status"][0] == 0      # OSQP's iteration cap; the device solver converges
            continue
        assert out["status"][0] == 0, path
        assert out["scp_iters"][0] == int(g["iterations"]), path
        assert out["n_accepted"][0] == int(g["n_accepted"]), path
        if int(g["n_accepted"]):
            check_against_golden(out["X"][0].T, out["U"][0].T, g)


TIGHT = dict(eps_abs=1e-9, eps_rel=1e-9, max_iter=40000, polish_refine_iter=30)


@pytest.mark.parametrize("name", ["solo12_trot", "solo12_pace", "solo12_bound", "bolt"])
def test_oracle_parity(gpu, cases, name):
    """1e-6 against the oracle solved tightly (the oracle at OSQP's default tolerance, the reference's
    setting, is itself up to 3.7e-6 away from that answer on bolt), equal SCP iteration counts against
    the oracle at the reference's settings."""
    from centroidal_mpc_b200.src.scp_solver import solve_scp_batched
    from oracle import scp
    from test_emu_parity import check_against_tight_oracle
    conf, models = cases[name]
    out = solve_scp_batched(models, conf.scp_params, return_stats=True)
    check_against_tight_oracle(name, conf, models[:2], out)


@pytest.mark.parametrize("name,B", [("solo12_trot", 64), ("solo12_pace", 33), ("bolt", 7)])
def test_gpu_equals_host_build_of_the_same_source(gpu, name, B):
    """Same arithmetic in the same order (up to FMA contraction, which nvcc and g++ choose
    differently): results agree to amplified round-off, iteration counts exactly."""
    import emu_binding as E
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.src.scp_solver import solve_scp_batched
    conf = synthetic.load_conf(name, N=100)
    batch = synthetic.make_batch(conf, B)
    out = solve_scp_batched(batch, conf.scp_params, return_stats=True)
    emu = E.solve_scp(batch, conf.scp_params)
    np.testing.assert_array_equal(out["scp_iters"], emu["scp_iters"])
    np.testing.assert_array_equal(out["status"], emu["status"])
    np.testing.assert_array_equal(out["qp_iters"], emu["qp_iters"])
    for b in range(B):
        assert relerr(out["X"][b], emu["X"][b]) < 1e-7 and relerr(out["U"][b], emu["U"][b]) < 1e-7


def test_full_size_properties(gpu):
    """BASELINE.json headline size: solo12 trot, N=100, batch 4096."""
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.device import BatchSolver
    from test_emu_parity import check_properties
    conf = synthetic.load_conf("solo12_trot", N=100)
    batch = synthetic.make_batch(conf, 4096)
    solver = BatchSolver(batch)
    out = solver.solve(conf.scp_params).results()
    assert (out["status"] == 0).all() and (out["scp_iters"] == 1).all() and (out["n_accepted"] == 1).all()
    # properties on a sample (the check is O(N) numpy per instance)
    import copy
    idx = np.linspace(0, 4095, 24).astype(int)
    sub = copy.copy(batch)
    sub.B = len(idx)
    for k in ("x_init", "x_final", "X_ref", "U_init"):
        setattr(sub, k, getattr(batch, k)[idx])
    check_properties(sub, {k: v[idx] for k, v in out.items()})
    # idempotence: solving again gives the same answer; the host-buffer entry point too
    again = solver.solve(conf.scp_params).results()
    np.testing.assert_array_equal(again["X"], out["X"])
    host = solver.solve_host(conf.scp_params)
    np.testing.assert_array_equal(host["X"], out["X"])
    np.testing.assert_array_equal(host["U"], out["U"])
    # page-locked result buffers: the kernel writes them directly (mapped host memory, no device-to-host copy)
    import torch
    pinned = dict(X=torch.zeros((4096, 101, 9), dtype=torch.float64).pin_memory(),
                  U=torch.zeros((4096, 100, batch.nu), dtype=torch.float64).pin_memory(),
                  scp_iters=torch.zeros(4096, dtype=torch.int32).pin_memory(),
                  status=torch.full((4096,), -1, dtype=torch.int32).pin_memory(),
                  n_accepted=torch.zeros(4096, dtype=torch.int32).pin_memory())
    solver.solve_host(conf.scp_params, out={k: v.numpy() for k, v in pinned.items()})
    np.testing.assert_array_equal(pinned["X"].numpy(), out["X"])
    np.testing.assert_array_equal(pinned["U"].numpy(), out["U"])
    np.testing.assert_array_equal(pinned["status"].numpy(), out["status"])
    np.testing.assert_array_equal(pinned["scp_iters"].numpy(), out["scp_iters"])
    # the host entry works on private copies: the problem bound to the handle is untouched
    np.testing.assert_array_equal(solver.solve(conf.scp_params).results()["X"], out["X"])
    solver.close()


def test_two_handles_with_different_horizons(gpu):
    """The dynamic shared-memory limit is a per-function attribute: a second handle with a shorter horizon
    must not lower it under the first one."""
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.device import BatchSolver
    c100, c40 = synthetic.load_conf("solo12_trot", N=100), synthetic.load_conf("solo12_trot", N=40)
    s100 = BatchSolver(synthetic.make_batch(c100, 64))
    first = s100.solve(c100.scp_params).results()
    s40 = BatchSolver(synthetic.make_batch(c40, 64))
    assert (s40.solve(c40.scp_params).results()["status"] == 0).all()
    again = s100.solve(c100.scp_params).results()
    assert (again["status"] == 0).all()
    np.testing.assert_array_equal(again["X"], first["X"])
    s40.close()
    s100.close()


def test_forced_branches_in_one_batch(gpu, cases):
    """Mode A batch (shared reference, perturbed x_init) + a forced rejection sequence."""
    import emu_binding as E
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.src.scp_solver import solve_scp_batched
    conf = synthetic.load_conf("solo12_pace", N=60)
    batch = synthetic.make_batch(conf, 16, mode="A")
    sp = dict(conf.scp_params, trust_region_radius0=1.0, max_iterations=3)
    out = solve_scp_batched(batch, sp)
    emu = E.solve_scp(batch, sp)
    np.testing.assert_array_equal(out["scp_iters"], emu["scp_iters"])
    np.testing.assert_array_equal(out["n_accepted"], emu["n_accepted"])
    assert (out["scp_iters"] == 3).all() and (out["n_accepted"] == 0).all()


def test_drop_in_solve_scp(gpu, cases):
    from centroidal_mpc_b200.src.scp_solver import interpolate_SCP_solution, solve_scp
    conf, models = cases["solo12_trot"]
    sol = solve_scp(models[0], conf.scp_params)
    assert set(sol) == {"state", "control", "gains", "covs"} and len(sol["state"]) == 1
    assert sol["state"][-1].shape == (9, conf.N + 1) and sol["control"][-1].shape == (12, conf.N)
    # gains / covs along the warm start (scp_solver.py:165-166), against the oracle
    from oracle import dynamics
    m = models[0]
    g, c = dynamics.lqr_gains_covs(m._init_trajectories["state"], m._init_trajectories["control"],
                                   m.problem_arrays(), m._Q, m._R, m._Cov_w, m._Cov_eta)
    assert sol["gains"][-1].shape == (conf.N, 12, 9) and sol["covs"][-1].shape == (conf.N + 1, 9, 9)
    assert relerr(sol["gains"][-1], g) < 1e-12 and relerr(sol["covs"][-1], c) < 1e-12
    ip = interpolate_SCP_solution(sol)
    assert ip["X"].shape == (9, conf.N * 10) and ip["U"].shape == (12, (conf.N - 1) * 10)
    empty = solve_scp(models[0], dict(conf.scp_params, trust_region_radius0=1.0, max_iterations=2))
    assert empty["state"] == [] and empty["control"] == []
    # the QP failure path: an infeasible problem (free fall, final state off the ballistic path) makes the
    # reference return False (scp_solver.py:65-68,146-148); so does the drop-in
    from scenarios import load_scenario, scenario_files
    g, conf5, sp5, free_fall = load_scenario([p for p in scenario_files() if "scen5" in p][0])
    assert bool(g["returned_false"])
    with pytest.warns(UserWarning):
        assert solve_scp(free_fall, sp5) is False


def test_linearize_and_rollout_kernels(gpu, cases):
    from oracle import dynamics
    conf, models = cases["solo12_bound"]
    m = models[0]
    prob = m.problem_arrays()
    rng = np.random.default_rng(5)
    traj = dict(state=prob["X_ref"] + 0.01 * rng.normal(size=prob["X_ref"].shape),
                control=prob["U_init"] + 0.1 * rng.normal(size=prob["U_init"].shape))
    td = m.compute_trajectory_data(traj)
    ref = dynamics.trajectory_data(traj["state"], traj["control"], prob)
    np.testing.assert_allclose(td["dynamics"], ref["dynamics"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(td["gradients"]["f_x"], ref["f_x"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(td["gradients"]["f_u"], ref["f_u"], rtol=0, atol=1e-14)
    roll = m.integrate_dynamics_trajectory(traj)
    np.testing.assert_allclose(roll[:, :conf.N], dynamics.rollout(traj["state"], traj["control"], prob), atol=1e-13)
    one = m.integrate_model_one_step(traj["state"][:, 3], traj["control"][:, 3],
                                     m._contact_data["contacts_position"][3], m._contact_data["contacts_logic"][3],
                                     m._contact_data["contacts_orient"][3])
    np.testing.assert_allclose(one, ref["dynamics"][:, 3], atol=1e-13)


def test_general_friction_path_and_ragged_plans_on_device(gpu, cases):
    """Rotated contact frames (per-knot friction table instead of the constant pyramid) and a
    batch whose instances have different contact plans (padded slots): device == host build."""
    import emu_binding as E
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.batch import ProblemBatch
    from centroidal_mpc_b200.src.centroidal_model import Centroidal_model
    from centroidal_mpc_b200.src.scp_solver import solve_scp_batched
    conf = synthetic.load_conf("solo12_trot", N=20)
    models = []
    for b in range(3):
        m = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, b))
        c, s = np.cos(0.1 * (b + 1)), np.sin(0.1 * (b + 1))
        Rx = np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
        R = m._contact_data["contacts_orient"]
        for k in range(conf.N):
            for i in range(4):
                if m._contact_data["contacts_logic"][k, i]:
                    R[k, i] = Rx
        models.append(m)
    batch = ProblemBatch(models, shared_plan=False)
    assert not batch.identity_R
    out = solve_scp_batched(batch, conf.scp_params, return_stats=True)
    emu = E.solve_scp(batch, conf.scp_params)
    np.testing.assert_array_equal(out["status"], emu["status"])
    np.testing.assert_array_equal(out["qp_iters"], emu["qp_iters"])
    for b in range(3):
        assert relerr(out["X"][b], emu["X"][b]) < 1e-7 and relerr(out["U"][b], emu["U"][b]) < 1e-7
    # different gaits in one tile: trot and bound plans side by side
    conf_t, mt = cases["solo12_trot"]
    _, mb = cases["solo12_bound"]
    from conftest import with_weights_of
    mixed = ProblemBatch([mt[0], with_weights_of(mb[0], mt[0]), mt[1], with_weights_of(mb[1], mt[0])], shared_plan=False)
    out = solve_scp_batched(mixed, conf_t.scp_params, return_stats=True)
    emu = E.solve_scp(mixed, conf_t.scp_params)
    np.testing.assert_array_equal(out["qp_iters"], emu["qp_iters"])
    for b in range(4):
        assert relerr(out["X"][b], emu["X"][b]) < 1e-7 and relerr(out["U"][b], emu["U"][b]) < 1e-7


def test_shipped_horizon_on_device(gpu):
    """N = 165 / 107 are the horizons of the shipped trot and pace gait tables.
    (At N = 165 the oracle's OSQP restatement stops at max_iter = 4000 with 'solved inaccurate',
    i.e. the reference would return False; the device solver certifies a KKT point.)"""
    import emu_binding as E
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.src.scp_solver import solve_scp_batched
    for name, N, B in (("solo12_trot", 165, 40), ("solo12_pace", 107, 35)):
        conf = synthetic.load_conf(name, N=N)
        batch = synthetic.make_batch(conf, B)
        out = solve_scp_batched(batch, conf.scp_params, return_stats=True)
        emu = E.solve_scp(batch, conf.scp_params)
        assert (out["status"] == 0).all()
        np.testing.assert_array_equal(out["scp_iters"], emu["scp_iters"])
        np.testing.assert_array_equal(out["qp_iters"], emu["qp_iters"])
        for b in range(B):
            assert relerr(out["X"][b], emu["X"][b]) < 1e-7 and relerr(out["U"][b], emu["U"][b]) < 1e-7


def test_bitwise_equal_to_host_build(gpu):
    """The GPU and the host build of the same source agree BIT FOR BIT (FMA contraction is explicit
    in both, no lane depends on its neighbours): 512 instances of the headline workload.
    scripts/full_batch_check.py does the same for all 4096."""
    import emu_binding as E
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.src.scp_solver import solve_scp_batched
    conf = synthetic.load_conf("solo12_trot", N=100)
    batch = synthetic.make_batch(conf, 512, first=1024)
    out = solve_scp_batched(batch, conf.scp_params, return_stats=True)
    emu = E.solve_scp(batch, conf.scp_params)
    for k in ("status", "scp_iters", "n_accepted", "qp_iters", "n_factor"):
        np.testing.assert_array_equal(out[k], emu[k])
    np.testing.assert_array_equal(out["X"], emu["X"])
    np.testing.assert_array_equal(out["U"], emu["U"])


def test_headline_size_oracle_parity(gpu):
    """48 instances of the headline workload (solo12 trot, N = 100, instance ids spread over the
    4096 of the benchmark batch) against the oracle: equal SCP iteration counts, 1e-6 norm-wise."""
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.src.centroidal_model import Centroidal_model
    from centroidal_mpc_b200.src.scp_solver import solve_scp_batched
    from oracle import scp
    conf = synthetic.load_conf("solo12_trot", N=100)
    ids = [int(i) for i in np.linspace(0, 4095, 48)]
    models = [Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, b)) for b in ids]
    out = solve_scp_batched(models, conf.scp_params)
    worst = 0.0
    for j, m in enumerate(models):
        ref = scp.solve_scp(m.problem_arrays(), conf.scp_params)
        assert ref is not False and out["status"][j] == 0
        assert out["scp_iters"][j] == ref["iterations"] and out["n_accepted"][j] == len(ref["state"])
        worst = max(worst, relerr(out["X"][j].T, ref["state"][-1]), relerr(out["U"][j].T, ref["control"][-1]))
    assert worst < TOL, worst


@pytest.mark.parametrize("name", ["solo12_trot", "solo12_pace", "solo12_bound", "bolt"])
def test_lqr_gains_and_covs_through_cabi(gpu, cases, name):
    """cmpc_lqr_covs against the oracle (1e-12 relative) and bitwise against the host build of
    csrc/cmpc_lqr.cuh; unshared plans, perturbed trajectories."""
    import emu_binding as E
    from centroidal_mpc_b200.batch import ProblemBatch
    from centroidal_mpc_b200.device import lqr_gains_covs_batched
    from oracle import dynamics
    from test_emu_parity import lqr_case
    conf, models = cases[name]
    m0 = models[0]
    X, U = lqr_case(conf, models)
    batch = ProblemBatch(models)
    w = (m0._Q, m0._R, m0._Cov_w, m0._Cov_eta)
    gains, covs = lqr_gains_covs_batched(batch, X, U, *w)
    gains, covs = gains.cpu().numpy(), covs.cpu().numpy()
    ge, ce = E.lqr_covs(batch, X, U, *w)
    assert np.array_equal(gains, ge) and np.array_equal(covs, ce)
    for b, m in enumerate(models):
        g, c = dynamics.lqr_gains_covs(X[b].T, U[b].T, m.problem_arrays(), *w)
        assert relerr(gains[b], g) < 1e-12 and relerr(covs[b], c) < 1e-12
    td = m0.compute_trajectory_data(dict(state=X[0].T, control=U[0].T))
    assert np.array_equal(td["LQR_gains"], gains[0]) and np.array_equal(td["Covs"], covs[0])


def test_lqr_covs_full_batch_properties(gpu):
    """Headline shape (4096 x N=100): every instance's covariances are symmetric and grow from 0;
    gains of inactive contacts are exactly zero; a sample of instances against the oracle."""
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.device import lqr_gains_covs_batched
    from oracle import dynamics
    conf = synthetic.load_conf("solo12_trot", N=100)
    batch = synthetic.make_batch(conf, 4096)
    from centroidal_mpc_b200.src.centroidal_model import Centroidal_model
    m0 = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, 0))
    w = (m0._Q, m0._R, m0._Cov_w, m0._Cov_eta)
    gains, covs = lqr_gains_covs_batched(batch, batch.X_ref, batch.U_init, *w)
    assert bool(gains.isfinite().all()) and bool(covs.isfinite().all())
    assert float((covs - covs.transpose(2, 3)).abs().max()) < 1e-12
    assert float(covs[:, 0].abs().max()) == 0.0
    tr = covs.diagonal(dim1=2, dim2=3).sum(-1)
    assert bool((tr[:, 1:] > 0).all())
    act = np.repeat(batch.contact_active.astype(bool), 3, axis=-1)
    act = np.broadcast_to(act, (4096,) + act.shape[1:]) if act.shape[0] == 1 else act
    assert float(gains.cpu().numpy()[~act].__abs__().max()) == 0.0
    for b in (0, 4095):
        m = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, b))
        g, c = dynamics.lqr_gains_covs(batch.X_ref[b].T, batch.U_init[b].T, m.problem_arrays(), *w)
        assert relerr(gains[b].cpu().numpy(), g) < 1e-12 and relerr(covs[b].cpu().numpy(), c) < 1e-12


@pytest.mark.parametrize("name", ["solo12_trot", "solo12_bound", "bolt"])
def test_friction_backoffs_through_cabi(gpu, cases, name):
    """cmpc_friction_backoffs (stochastic-mode friction upper bounds) against the oracle and
    bitwise against the host build."""
    import emu_binding as E
    from centroidal_mpc_b200.batch import ProblemBatch
    from centroidal_mpc_b200.device import (chance_constraint_xi, friction_backoffs, friction_backoffs_batched,
                                            lqr_gains_covs_batched)
    from oracle import qp_build
    from test_emu_parity import lqr_case
    conf, models = cases[name]
    m0 = models[0]
    X, U = lqr_case(conf, models, seed=2)
    batch = ProblemBatch(models)
    gains, covs = lqr_gains_covs_batched(batch, X, U, m0._Q, m0._R, m0._Cov_w, m0._Cov_eta)
    ub = friction_backoffs_batched(batch, gains, covs, m0._beta_u).cpu().numpy()
    gh, ch = gains.cpu().numpy(), covs.cpu().numpy()
    assert np.array_equal(ub, E.friction_backoffs(batch, chance_constraint_xi(m0._beta_u), gh, ch))
    for b, m in enumerate(models):
        ref, xi = qp_build.friction_backoffs(m.problem_arrays(), gh[b], ch[b], m0._beta_u)
        assert xi == chance_constraint_xi(m0._beta_u)
        assert np.abs(ub[b] - ref).max() <= 1e-12 * np.abs(ref).max()
    one = friction_backoffs(m0)
    assert one.shape == (conf.N, batch.nc, 4) and np.all(one <= 0.0) and np.any(one < 0.0)


@pytest.mark.parametrize("name", ["solo12_trot", "solo12_bound"])
def test_stochastic_mode_on_device(gpu, name):
    """Centroidal_model(conf, STOCHASTIC_OCP=True) through solve_scp_batched: back-offs computed on
    the device, friction rows G f <= ub inside cmpc_scp_kernel; against the oracle's stochastic
    solve (1e-6) and bitwise against the host build fed with the same upper bounds."""
    import emu_binding as E
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.batch import ProblemBatch
    from centroidal_mpc_b200.device import BatchSolver
    from centroidal_mpc_b200.src.centroidal_model import Centroidal_model
    from centroidal_mpc_b200.src.scp_solver import solve_scp, solve_scp_batched
    from oracle import scp
    from test_emu_parity import oracle_backoffs
    conf = synthetic.load_conf(name, N=40)
    models = [Centroidal_model(conf, STOCHASTIC_OCP=True, centroidal_traj=synthetic.reference_trajectory(conf, b))
              for b in range(40)]
    batch = ProblemBatch(models)
    solver = BatchSolver(batch)
    out = solve_scp_batched(batch, conf.scp_params, solver=solver)
    ub = solver.friction_ub.cpu().numpy()
    solver.close()
    assert (out["status"] == 0).all() and ub.min() < -0.05
    from centroidal_mpc_b200.device import STOCHASTIC_QP_DEFAULTS
    host = E.solve_scp(batch, conf.scp_params, STOCHASTIC_QP_DEFAULTS, friction_ub=ub)
    assert np.array_equal(out["X"], host["X"]) and np.array_equal(out["U"], host["U"])
    assert np.array_equal(out["scp_iters"], host["scp_iters"])
    for b in (0, 39):
        ubo = oracle_backoffs(models[b])
        assert np.abs(ub[b] - ubo).max() <= 1e-12 * np.abs(ubo).max()
        ref = scp.solve_scp(dict(models[b].problem_arrays(), friction_ub=ubo), conf.scp_params)
        assert out["scp_iters"][b] == ref["iterations"]
        assert relerr(out["X"][b].T, ref["state"][-1]) < TOL and relerr(out["U"][b].T, ref["control"][-1]) < TOL
    # the nominal model of the same problem gives a different answer; the drop-in entry agrees with the batch
    nominal = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, 0))
    sol_n, sol_s = solve_scp(nominal, conf.scp_params), solve_scp(models[0], conf.scp_params)
    assert np.array_equal(sol_s["control"][-1], out["U"][0].T)
    assert relerr(sol_s["control"][-1], sol_n["control"][-1]) > 5e-4


@pytest.mark.parametrize("B", [1, 33, 300, 1100, 9001])
def test_host_entry_point_chunking(gpu, B):
    """cmpc_solve_scp_host cuts the batch into chunks of one resident set of tiles each (kernels on one stream,
    uploads of the next chunk on a second): ragged last tile, one chunk, three chunks (B = 9001); unshared plans
    at B=33; pageable result buffers (staged downloads) here, pinned ones in bench.py.
    Results must equal the device-pointer entry bit for bit, all outputs included."""
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.batch import ProblemBatch
    from centroidal_mpc_b200.device import BatchSolver
    from centroidal_mpc_b200.src.centroidal_model import Centroidal_model
    conf = synthetic.load_conf("solo12_trot", N=30)
    if B == 33:
        models = [Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, b)) for b in range(B)]
        batch = ProblemBatch(models, shared_plan=False)
    else:
        batch = synthetic.make_batch(conf, B)
    solver = BatchSolver(batch)
    dev = solver.solve(conf.scp_params).results()
    host = solver.solve_host(conf.scp_params)
    for key in ("X", "U", "scp_iters", "status", "n_accepted"):
        np.testing.assert_array_equal(host[key], dev[key], err_msg=key)
    assert (host["status"] == 0).all()
    solver.close()


def test_stochastic_golden_through_cabi(gpu):
    """tests/golden/stoch*.npz: gains, covariances, back-offs and the stochastic solve, all on the device."""
    from centroidal_mpc_b200.device import friction_backoffs
    from centroidal_mpc_b200.src.scp_solver import solve_scp
    from test_emu_parity import check_against_stochastic_golden
    from test_oracle import _stoch_golden_files, load_stoch_golden
    files = _stoch_golden_files()
    assert len(files) == 4
    for path in files:
        g, conf, m = load_stoch_golden(path)
        sol = solve_scp(m, conf.scp_params)
        assert sol is not False and len(sol["state"]) == int(g["iterations"]) == 1
        check_against_stochastic_golden(g, sol["gains"][-1], sol["covs"][-1], friction_backoffs(m),
                                        sol["state"][-1], sol["control"][-1], len(sol["state"]))


def test_stochastic_full_size_properties(gpu):
    """Stochastic mode at the headline shape (4096 x N=100, trot): every instance accepted after one SCP
    iteration and certified by the first polish attempt; the forces honour the backed-off rows; the forces
    of inactive contacts are exact zeros; the dynamics hold; the first tile equals the host build bit for bit."""
    import emu_binding as E
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.batch import ProblemBatch
    from centroidal_mpc_b200.device import STOCHASTIC_QP_DEFAULTS, BatchSolver
    conf = synthetic.load_conf("solo12_trot", N=100)
    batch = synthetic.make_batch(conf, 4096, stochastic=True)
    solver = BatchSolver(batch)
    out = solver.solve(conf.scp_params).results()
    st = solver.stats()
    ub = solver.friction_ub.cpu().numpy()
    solver.close()
    assert (out["status"] == 0).all() and (out["scp_iters"] == 1).all() and (out["n_accepted"] == 1).all()
    assert (st["qp_iters"] == 8).all() and (st["info"][:, 9] == 1).all()
    act = batch.contact_active[0].astype(bool)                       # [N, nc], shared plan
    U = out["U"].reshape(4096, 100, 4, 3)
    assert np.all(U[:, ~act] == 0.0)
    kf = batch.proto["mu"] / np.sqrt(2.0)
    rows = np.stack([U[..., 0], -U[..., 0], U[..., 1], -U[..., 1]], -1) - kf * U[..., 2:3] - ub
    assert rows[:, act].max() < 1e-7 and ub.min() < -1.0 and np.all(ub[:, ~act] == 0.0)
    # dynamics x_{k+1} = f(x_k, u_k) are linear in (x, u) about the warm start: check through the rollout
    X = out["X"]
    m, dt, g = batch.proto["m"], batch.proto["dt"], batch.proto["g"]
    F = U.sum(2)
    assert np.abs(X[:, 1:, 0:3] - (X[:, :-1, 0:3] + dt / m * X[:, :-1, 3:6])).max() < 1e-9
    lin = X[:, :-1, 3:6] + dt * F
    lin[..., 2] += dt * m * g
    assert np.abs(X[:, 1:, 3:6] - lin).max() < 1e-9
    np.testing.assert_allclose(X[:, 0], batch.x_init, atol=1e-9)
    np.testing.assert_allclose(X[:, -1], batch.x_final, atol=1e-7)
    # first tile against the host build of the kernel source
    import copy
    sub = copy.copy(batch)
    sub.B = 32
    for k in ("x_init", "x_final", "X_ref", "U_init"):
        setattr(sub, k, np.ascontiguousarray(getattr(batch, k)[:32]))
    host = E.solve_scp(sub, conf.scp_params, STOCHASTIC_QP_DEFAULTS, friction_ub=ub[:32])
    assert np.array_equal(host["X"], out["X"][:32]) and np.array_equal(host["U"], out["U"][:32])


# ---- branches of the trust-region loop the shipped configurations never take, and the other BASELINE
# ---- configurations at the benchmark horizon (tests/golden/scen_*.npz, n100_*.npz: tightly solved oracle)
def test_trust_region_scenarios_on_device(gpu):
    """Binding L1 trust region on an accepted iterate (slack active; on the surface of the ball), rejection
    followed by acceptance with a different QP solution (accuracy ratio; trust test), NaN convergence test
    of a zero warm start, infeasible QP: through the C ABI against the oracle fixtures (1e-6, equal counts
    and verdicts), and bit for bit against the host build."""
    import emu_binding as E
    from centroidal_mpc_b200.batch import ProblemBatch
    from centroidal_mpc_b200.src.scp_solver import solve_scp_batched
    from scenarios import check_scenario, load_scenario, scenario_files, scenario_qp
    files = scenario_files()
    assert len(files) >= 6
    for path in files:
        g, conf, sp, model = load_scenario(path)
        out = solve_scp_batched([model], sp, qp_settings=scenario_qp(path), return_stats=True)
        check_scenario(out, 0, g, relerr)
        emu = E.solve_scp(ProblemBatch([model]), sp, qp_overrides=scenario_qp(path))
        for k in ("status", "scp_iters", "n_accepted", "qp_iters", "n_factor"):
            np.testing.assert_array_equal(out[k], emu[k], err_msg=path)
        if int(g["n_accepted"]):
            np.testing.assert_array_equal(out["X"], emu["X"])
            np.testing.assert_array_equal(out["U"], emu["U"])
            if "binding" in path or "surface" in path:
                dk = np.abs(out["X"][0][:, 6:] - np.asarray(g["X_ref"]).T[:, 6:]).sum(axis=1).max()
                assert dk >= float(g["radius"][-1]) * (1 - 1e-9)


@pytest.mark.parametrize("name,mode", [("solo12_pace", "A"), ("solo12_bound", "B"), ("bolt", "B")])
def test_baseline_configs_at_n100_match_tight_oracle(gpu, name, mode):
    """BASELINE.json configurations 2-4 at the benchmark horizon N = 100: pace with 1024 perturbed initial
    states (32 sampled), bound 4096 (8 sampled), bolt 8192 (8 sampled), 1e-6 against the tightly solved
    oracle; the sample is solved inside the FULL batch (same tiles, same neighbours as the benchmark)."""
    from centroidal_mpc_b200.device import BatchSolver
    from scenarios import n100_samples
    conf, full, sub, g = n100_samples(name, mode)
    assert bool(np.all(g["ok"]))
    solver = BatchSolver(full)
    out = solver.solve(conf.scp_params).results()
    solver.close()
    assert (out["status"] == 0).all() and (out["n_accepted"] == 1).all()
    worst = 0.0
    for j, b in enumerate(np.asarray(g["ids"])):
        assert out["scp_iters"][b] == int(g["iterations"][j])
        worst = max(worst, relerr(out["X"][b].T, g["X"][j]), relerr(out["U"][b].T, g["U"][j]))
    assert worst < TOL, worst


def test_results_into_a_peer_buffer(gpu):
    """parallel.PeerResults / BatchSolver.solve(out=...): the kernel writes the solutions into a buffer obtained from
    cmpc_peer_alloc (on a multi-GPU box: rank 0's memory, mapped into the other ranks with cmpc_peer_open) at an
    instance offset; the zero-copy views of that buffer equal the solver's own result tensors.  One process here
    (the destination's half of the exchange); bench.py --gpus N runs the cross-process half."""
    from centroidal_mpc_b200 import parallel, synthetic
    from centroidal_mpc_b200.device import BatchSolver

    class OneRank:
        def get_rank(self): return 0
        def get_world_size(self): return 1
        def broadcast_object_list(self, box, src=0): return None
    conf = synthetic.load_conf("solo12_trot", N=30)
    batch = synthetic.make_batch(conf, 37)
    solver = BatchSolver(batch)
    own = solver.solve(conf.scp_params).results()
    first, total = 5, 50                                  # this "rank" owns instances 5..41 of a job of 50
    peer = parallel.PeerResults(total, conf.N, batch.nu, OneRank(), dst=0, slots=2)
    solver.solve(conf.scp_params, out=peer.pointers(1, first))
    gpu.cuda.synchronize()
    view = peer.tensors(1)
    for key in ("X", "U", "scp_iters", "status", "n_accepted"):
        np.testing.assert_array_equal(view[key][first:first + 37].cpu().numpy(), own[key], err_msg=key)
    peer.close()
    solver.close()

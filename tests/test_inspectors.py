"""The cost / constraint inspectors (centroidal_mpc_b200/src/cost.py, src/constraints.py: the reference's
``Cost(Q, p)`` / ``Constraint(mat, lb, ub)`` objects expanded from the device problem) against the
oracle's independent assembly (oracle/qp_build.py), entry by entry in the reference's row and column
order.  CPU: the linearisation comes from the oracle; GPU: from the device (cmpc_linearize, and in
stochastic mode cmpc_lqr_covs + cmpc_friction_backoffs), so the equality is a statement about the QP the
device solves."""
import numpy as np
import pytest

from centroidal_mpc_b200 import synthetic
from centroidal_mpc_b200.src import constraints as C
from centroidal_mpc_b200.src import cost as K
from centroidal_mpc_b200.src.centroidal_model import Centroidal_model
from centroidal_mpc_b200.src.scp_solver import stack_up_all_constraints, sum_up_all_costs
from oracle import dynamics, qp_build


def _model(name, N, b=0, stochastic=False, rotate=False):
    conf = synthetic.load_conf(name, N=N)
    m = Centroidal_model(conf, STOCHASTIC_OCP=stochastic, centroidal_traj=synthetic.reference_trajectory(conf, b))
    if rotate:
        import copy
        m._contact_trajectory = copy.deepcopy(m._contact_trajectory)   # the poses belong to the shared config module
        c, s = np.cos(0.2), np.sin(0.2)
        Rx = np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
        for contact in m._contact_trajectory:
            for d in m._contact_trajectory[contact]:
                if d.ACTIVE:
                    d.pose.rotation = Rx
        for k in range(N):
            for i in range(m._contact_data["contacts_logic"].shape[1]):
                if m._contact_data["contacts_logic"][k, i]:
                    m._contact_data["contacts_orient"][k, i] = Rx
    return conf, m


def _oracle_traj_data(prob):
    td = dynamics.trajectory_data(prob["X_ref"], prob["U_init"], prob)
    return dict(dynamics=td["dynamics"], gradients={"f_x": td["f_x"], "f_u": td["f_u"], "f_w": td["f_w"]})


def _check(model, traj_data, tr, friction_ub=None, atol=0.0):
    prob = model.problem_arrays()
    if friction_ub is not None:
        prob["friction_ub"] = friction_ub
    td_o = dynamics.trajectory_data(prob["X_ref"], prob["U_init"], prob)
    P, q = qp_build.build_cost(prob)
    A, lo, up, blocks = qp_build.build_constraints(prob, td_o, tr["radius"], tr["weight"])
    cost = sum_up_all_costs(model)
    assert isinstance(cost, K.Cost) and cost.Q.shape == P.shape
    assert abs(cost.Q - P).max() == 0.0
    np.testing.assert_array_equal(cost.p, q)
    cons = stack_up_all_constraints(model, model._init_trajectories, traj_data, tr, friction_ub=friction_ub)
    assert isinstance(cons, C.Constraint) and cons.mat.shape == A.shape
    d = abs(cons.mat - A)
    assert (d.max() if d.nnz else 0.0) <= atol
    # same sparsity pattern where the tolerance is zero (the reference order of rows and columns)
    if atol == 0.0:
        assert (cons.mat != A).nnz == 0
    np.testing.assert_allclose(cons.lb, lo, rtol=0, atol=atol)
    np.testing.assert_allclose(cons.ub, up, rtol=0, atol=atol)
    # block offsets of the reference order: initial, dynamics, final, friction, trust
    N, nc = model._N, prob["contact_active"].shape[1]
    cop = 2 * N * nc if prob["robot"] == "TALOS" else 0      # CoP box: x rows then y rows per foot (constraints.py:111-145)
    assert blocks["dynamics"] == 9 and blocks["final"] == 9 + 9 * N and blocks["friction"] == 18 + 9 * N + cop
    if cop:
        assert blocks["cop"] == 18 + 9 * N
    assert blocks["trust"] == 18 + 9 * N + cop + 5 * N * nc and blocks["m"] == cons.mat.shape[0]
    return cons


@pytest.mark.parametrize("name,N", [("solo12_trot", 3), ("solo12_trot", 40), ("solo12_bound", 40), ("bolt", 5), ("bolt", 40),
                                    ("talos", 5), ("talos", 30)])
def test_inspectors_match_the_oracle_assembly(name, N):
    conf, model = _model(name, N)
    tr = dict(radius=conf.scp_params["trust_region_radius0"], weight=conf.scp_params["omega0"])
    _check(model, _oracle_traj_data(model.problem_arrays()), tr)


def test_inspectors_rotated_contacts_and_upper_bounds():
    conf, model = _model("solo12_trot", 12, rotate=True)
    fub = -np.abs(np.random.default_rng(0).normal(size=(12, 4, 4))) * model._contact_data["contacts_logic"][:, :, None]
    _check(model, _oracle_traj_data(model.problem_arrays()), dict(radius=0.3, weight=50.0), friction_ub=fub)


def test_builders_keep_the_reference_shapes():
    conf, model = _model("solo12_pace", 7)
    N, n = 7, model._total_nb_optimizers
    assert C.construct_initial_constraints(model).mat.shape == (9, n)
    assert C.construct_final_constraints(model).mat.shape == (9, n)
    fr = C.construct_friction_pyramid_constraints(model)
    assert fr.mat.shape == (4 * 5 * N, n) and np.all(np.isneginf(fr.lb)) and np.all(fr.ub == 0.0)
    # the fifth pyramid row of every knot is never written (constraints.py:180, range(4))
    assert fr.mat[4::5].nnz == 0
    tr = C.construct_state_trust_region_constraints(model, model._init_trajectories, dict(radius=1.0, weight=10.0))
    assert tr.mat.shape == (8 * (N + 1) + (N + 1), n)
    assert K.construct_control_trust_region_cost(model).p[-N:].sum() == N


@pytest.mark.gpu
@pytest.mark.parametrize("name,N,stochastic", [("solo12_trot", 5, False), ("solo12_trot", 40, False), ("bolt", 40, False),
                                               ("solo12_bound", 40, True), ("talos", 30, False)])
def test_device_problem_expands_to_the_oracle_qp(gpu, name, N, stochastic):
    """traj_data from the device (cmpc_linearize [+ cmpc_lqr_covs, cmpc_friction_backoffs]): the QP the
    device solves, written out in the reference's matrices, is the oracle's."""
    from centroidal_mpc_b200 import device
    conf, model = _model(name, N, stochastic=stochastic)
    td = model.compute_trajectory_data(model._init_trajectories)
    assert set(td) >= {"dynamics", "gradients", "LQR_gains", "Covs", "Covs_gradients"}
    assert td["gradients"]["f_w"].shape == (N, 9, model._n_w)
    assert td["Covs_gradients"]["Cov_dx"].shape == (N + 1, 9, 9, 9, N + 1) and not td["Covs_gradients"]["Cov_dx"].any()
    prob = model.problem_arrays()
    ref = dynamics.trajectory_data(prob["X_ref"], prob["U_init"], prob)
    np.testing.assert_allclose(td["gradients"]["f_w"], ref["f_w"], rtol=0, atol=1e-14)
    fub = device.friction_backoffs(model) if stochastic else None
    if stochastic:
        g, c = dynamics.lqr_gains_covs(prob["X_ref"], prob["U_init"], prob, model._Q, model._R, model._Cov_w, model._Cov_eta)
        ref_ub, _ = qp_build.friction_backoffs(prob, g, c, model._beta_u)
        np.testing.assert_allclose(fub, ref_ub, rtol=1e-10, atol=1e-12)
    tr = dict(radius=conf.scp_params["trust_region_radius0"], weight=conf.scp_params["omega0"])
    _check(model, td, tr, friction_ub=fub, atol=1e-13)

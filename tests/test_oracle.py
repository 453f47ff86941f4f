"""The oracle against independent checks: finite differences, HiGHS, a KKT certificate, and the
committed golden fixtures."""
import glob
import os

import numpy as np
import pytest

from conftest import relerr
from oracle import dynamics, osqp_restatement, qp_build, scp


def _random_knot(rng, nc, robot):
    npc = 6 if robot == "TALOS" else 3
    x = rng.normal(size=9)
    u = rng.normal(size=nc * npc) * 3
    p = rng.normal(size=(nc, 3))
    a = (rng.random(nc) > 0.3).astype(int)
    R = np.stack([np.linalg.qr(rng.normal(size=(3, 3)))[0] for _ in range(nc)])
    return x, u, p, a, R


@pytest.mark.parametrize("robot,nc", [("solo12", 4), ("solo12", 2), ("TALOS", 2)])
def test_jacobians_match_finite_differences(robot, nc):
    rng = np.random.default_rng(0)
    for _ in range(5):
        x, u, p, a, R = _random_knot(rng, nc, robot)
        args = dict(m=2.5, g=-9.81, dt=0.01, robot=robot)
        A, B, C = dynamics.jacobians(x, u, p, a, R, **args)
        h = 1e-6
        for M, v, setter in ((A, x, 0), (B, u, 1), (C, p.reshape(-1), 2)):
            for j in range(v.size):
                vp, vm = v.copy(), v.copy()
                vp[j] += h
                vm[j] -= h
                if setter == 0:
                    fp, fm = dynamics.step(vp, u, p, a, R, **args), dynamics.step(vm, u, p, a, R, **args)
                elif setter == 1:
                    fp, fm = dynamics.step(x, vp, p, a, R, **args), dynamics.step(x, vm, p, a, R, **args)
                else:
                    fp = dynamics.step(x, u, vp.reshape(nc, 3), a, R, **args)
                    fm = dynamics.step(x, u, vm.reshape(nc, 3), a, R, **args)
                np.testing.assert_allclose(M[:, j], (fp - fm) / (2 * h), atol=1e-8)


def test_qp_dimensions_and_row_order(cases):
    conf, models = cases["solo12_trot"]
    prob = models[0].problem_arrays()
    N = prob["N"]
    P, q, A, l, u, blocks, td = scp.build_qp(prob, 100.0, 100.0)
    assert P.shape == (23 * N + 10, 23 * N + 10)
    assert A.shape == (38 * N + 27, 23 * N + 10)
    order = ["initial", "dynamics", "final", "friction", "trust", "slack_sign"]
    offs = [blocks[k] for k in order]
    assert offs == sorted(offs) and offs[0] == 0 and offs[1] == 9 and offs[2] == 9 + 9 * N
    assert offs[3] == 18 + 9 * N and offs[4] == offs[3] + 20 * N and offs[5] == offs[4] + 8 * (N + 1)
    # unilaterality is implied by rows 0+1 of each pyramid; the 5th row is never written
    fr = A[blocks["friction"]:blocks["trust"]].toarray()
    assert np.all(np.abs(fr[4::5]).sum(axis=1) == 0)
    # tracking gradient and slack cost
    np.testing.assert_allclose(q[:9], -np.diag(prob["state_cost_weights"]) * prob["X_ref"][:, 0])
    assert np.all(q[qp_build.idx_t(prob, 0):qp_build.idx_t(prob, 0) + N + 1] == 1.0)


def _highs_qp(P, q, A, l, u):
    from scipy.optimize._highspy import _core as hs
    from scipy import sparse
    n, m = P.shape[0], A.shape[0]
    h = hs._Highs()
    h.setOptionValue("output_flag", False)
    lp = hs.HighsLp()
    lp.num_col_, lp.num_row_ = n, m
    lp.col_cost_ = q
    lp.col_lower_ = np.full(n, -hs.kHighsInf)
    lp.col_upper_ = np.full(n, hs.kHighsInf)
    lp.row_lower_ = np.where(np.isinf(l), -hs.kHighsInf, l)
    lp.row_upper_ = np.where(np.isinf(u), hs.kHighsInf, u)
    Ac = sparse.csc_matrix(A)
    lp.a_matrix_.format_ = hs.MatrixFormat.kColwise
    lp.a_matrix_.start_, lp.a_matrix_.index_, lp.a_matrix_.value_ = Ac.indptr, Ac.indices, Ac.data
    hess = hs.HighsHessian()
    Pl = sparse.csc_matrix(sparse.tril(P))
    hess.dim_, hess.format_ = n, hs.HessianFormat.kTriangular
    hess.start_, hess.index_, hess.value_ = Pl.indptr, Pl.indices, Pl.data
    model = hs.HighsModel()
    model.lp_, model.hessian_ = lp, hess
    h.passModel(model)
    h.run()
    return np.array(h.getSolution().col_value), h.modelStatusToString(h.getModelStatus())


def test_osqp_restatement_against_highs_and_kkt():
    """Independent active-set QP solver on a small instance (equalities exact, unused control
    slacks regularised so that the Hessian is usable by HiGHS)."""
    from scipy import sparse
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.src.centroidal_model import Centroidal_model
    conf = synthetic.load_conf("solo12_trot", N=12)
    prob = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, 0)).problem_arrays()
    P, q, A, l, u, blocks, td = scp.build_qp(prob, 100.0, 100.0)
    res = osqp_restatement.solve(P, q, A, l, u, eps_abs=1e-7, eps_rel=1e-7, polish=True)
    assert res.status == "solved"
    # KKT certificate of the oracle's answer
    z, y = res.x, res.y
    Az = A @ z
    assert np.max(np.maximum(l - Az, 0) + np.maximum(Az - u, 0)) < 1e-7
    assert np.max(np.abs(P @ z + q + A.T @ y)) < 1e-5 * max(1.0, np.max(np.abs(q)))
    eq = (u - l) < 1e-4
    comp = np.where(~eq, np.minimum(np.abs(y), np.minimum(np.abs(Az - l), np.abs(Az - u))), 0.0)
    assert np.max(comp) < 1e-5
    # HiGHS
    mid = 0.5 * (l + u)
    l2, u2 = np.where(eq, mid, l), np.where(eq, mid, u)
    N = prob["N"]
    reg = np.zeros(P.shape[0])
    reg[-(2 * N + 1):] = 1e-8          # state and control slacks have no curvature
    xh, status = _highs_qp(P + sparse.diags(reg), q, A, l2, u2)
    assert "Optimal" in status, status
    Xo, Uo = qp_build.unpack(prob, z)
    Xh, Uh = qp_build.unpack(prob, xh)
    assert relerr(Xh, Xo) < 1e-6 and relerr(Uh, Uo) < 1e-6


def test_fp32_emulation_stays_within_budget(cases):
    """Rounding the JAX-touched quantities to float32 (what default JAX does in the reference,
    SURVEY.md Appendix C #3) moves the answer by far less than the 1e-6 parity budget...
    in norm-wise relative terms on (X, U)."""
    conf, models = cases["solo12_trot"]
    prob = models[0].problem_arrays()
    a = scp.solve_scp(prob, conf.scp_params)
    b = scp.solve_scp(prob, conf.scp_params, emulate_jax_fp32=True)
    assert a["iterations"] == b["iterations"] == 1
    assert relerr(b["state"][-1], a["state"][-1]) < 1e-6
    assert relerr(b["control"][-1], a["control"][-1]) < 1e-6


def _golden_files():
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return sorted(glob.glob(os.path.join(here, "case*.npz")))


def load_golden(path):
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.src.centroidal_model import Centroidal_model
    g = np.load(path)
    conf = synthetic.load_conf(str(g["name"]), N=int(g["N"]))
    sp = dict(conf.scp_params)
    for k, v in zip(g["scp_keys"], g["scp_vals"]):
        sp[str(k)] = int(v) if str(k) == "max_iterations" else float(v)
    model = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, int(g["b"])))
    return g, conf, sp, model


@pytest.mark.parametrize("path", [p for p in _golden_files() if "_b0" in p and "N100" not in p],
                         ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_reproduces_golden(path):
    g, conf, sp, model = load_golden(path)
    log = []
    sol = scp.solve_scp(model.problem_arrays(), sp, log=log)
    assert (sol is False) == bool(g["returned_false"])
    assert len(log) == int(g["iterations"])
    assert [e.get("verdict", "qp_failed") for e in log] == [str(v) for v in g["verdicts"]]
    if int(g["n_accepted"]):
        assert relerr(sol["state"][-1], g["X"]) < 1e-9 and relerr(sol["control"][-1], g["U"]) < 1e-9


def test_oracle_stochastic_mode_tightens_the_friction_rows(cases):
    """Stochastic mode of the oracle (constraints.py:157-163,187-214): the friction rows get the
    back-offs as upper bounds; the solve stays feasible, honours them, and moves the forces."""
    from oracle import dynamics, qp_build, scp
    conf, models = cases["solo12_trot"]
    m = models[0]
    prob = m.problem_arrays()
    g, c = dynamics.lqr_gains_covs(prob["X_ref"], prob["U_init"], prob, m._Q, m._R, m._Cov_w, m._Cov_eta)
    ub, xi = qp_build.friction_backoffs(prob, g, c, m._beta_u)
    assert abs(xi - 2.5121443279304616) < 1e-12          # Phi^-1(1 - 0.01/5*3)
    assert ub.shape == (conf.N, 4, 4) and ub.min() < -0.1 and ub.max() == 0.0
    nom = scp.solve_scp(prob, conf.scp_params)
    sto = scp.solve_scp(dict(prob, friction_ub=ub), conf.scp_params)
    assert sto is not False and len(sto["state"]) == 1
    U, Un = sto["control"][-1], nom["control"][-1]
    pyr = qp_build.friction_pyramid(prob["mu"])[:4]
    for k in range(conf.N):
        for cc in range(4):
            if prob["contact_active"][k, cc]:
                assert (pyr @ U[3 * cc:3 * cc + 3, k] - ub[k, cc]).max() < 1e-7
    assert np.linalg.norm(U - Un) / np.linalg.norm(Un) > 1e-3


def _stoch_golden_files():
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return sorted(glob.glob(os.path.join(here, "stoch*.npz")))


def load_stoch_golden(path):
    """(fixture, conf, stochastic model) of a tests/golden/stoch*.npz file."""
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.src.centroidal_model import Centroidal_model
    g = np.load(path)
    conf = synthetic.load_conf(str(g["name"]), N=int(g["N"]))
    model = Centroidal_model(conf, STOCHASTIC_OCP=True,
                             centroidal_traj=synthetic.reference_trajectory(conf, int(g["b"])))
    return g, conf, model


@pytest.mark.parametrize("path", _stoch_golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_reproduces_stochastic_golden(path):
    from oracle import dynamics, qp_build
    g, conf, m = load_stoch_golden(path)
    prob = m.problem_arrays()
    gains, covs = dynamics.lqr_gains_covs(prob["X_ref"], prob["U_init"], prob, m._Q, m._R, m._Cov_w, m._Cov_eta)
    ub, xi = qp_build.friction_backoffs(prob, gains, covs, m._beta_u)
    assert relerr(gains, g["gains"]) < 1e-12 and relerr(covs, g["covs"]) < 1e-12
    assert relerr(ub, g["friction_ub"]) < 1e-12 and xi == float(g["xi"])
    sol = scp.solve_scp(dict(prob, friction_ub=ub), conf.scp_params)
    assert sol["iterations"] == int(g["iterations"])
    assert relerr(sol["state"][-1], g["X"]) < 1e-9 and relerr(sol["control"][-1], g["U"]) < 1e-9
    # OSQP's answer at its default settings against the tightly solved one
    assert relerr(g["X"], g["X_tight"]) < 5e-6 and relerr(g["U"], g["U_tight"]) < 5e-6

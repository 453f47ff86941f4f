"""Receding-horizon driver and the warm-started active set
(cmpc_qp_settings.warm_start): a warm solve must return the same certified KKT point as a cold solve of the
same problem, without ADMM iterations."""
import numpy as np
import pytest

import emu_binding as E
from conftest import relerr
from centroidal_mpc_b200 import synthetic
from centroidal_mpc_b200.batch import ProblemBatch
from centroidal_mpc_b200.device import WRENCH_QP_DEFAULTS
from centroidal_mpc_b200.mpc import shift_window

TOL = 1e-6


def test_shift_window():
    rng = np.random.default_rng(0)
    X, U = rng.normal(size=(3, 6, 9)), rng.normal(size=(3, 5, 12))
    xt, ut = rng.normal(size=(3, 9)), rng.normal(size=(3, 12))
    xi, xf, Xr, Ui = shift_window(X, U, xt, ut)
    np.testing.assert_array_equal(Xr[:, :-1], X[:, 1:]); np.testing.assert_array_equal(Xr[:, -1], xt)
    np.testing.assert_array_equal(Ui[:, :-1], U[:, 1:]); np.testing.assert_array_equal(Ui[:, -1], ut)
    np.testing.assert_array_equal(xi, X[:, 1]); np.testing.assert_array_equal(xf, xt)
    x0 = rng.normal(size=(3, 9))
    xi, xf, Xr, Ui = shift_window(X, U, xt, ut, x_init=x0)
    np.testing.assert_array_equal(Xr[:, 0], x0); np.testing.assert_array_equal(X[:, 1], X[:, 1])


@pytest.mark.parametrize("name", ["solo12_trot", "solo12_bound", "bolt", "talos"])
def test_warm_start_equals_cold_solve_host_build(name):
    """Three ticks of a receding horizon in the host build: at every tick the warm-started solve equals the cold
    solve of the same problem to 1e-6 and runs no ADMM iteration."""
    H = 40
    full = synthetic.make_batch(synthetic.load_conf(name), 4)
    conf = synthetic.load_conf(name, N=H)
    ov = WRENCH_QP_DEFAULTS if name == "talos" else {}
    proto = dict(synthetic.make_batch(conf, 1).proto)
    b = synthetic.make_batch(conf, 4)
    sol = E.solve_scp(b, conf.scp_params, qp_overrides=ov or None)
    assert (sol["status"] == 0).all() and (sol["n_accepted"] > 0).all()
    for tick in range(1, 4):
        k1 = tick + H
        xi, xf, Xr, Ui = shift_window(sol["X"], sol["U"], full.X_ref[:, k1], full.U_init[:, k1 - 1])
        p = dict(proto, contact_pos=full.contact_pos[0, tick:k1], contact_active=full.contact_active[0, tick:k1],
                 contact_R=(full.contact_R[0, tick:k1] if full.contact_R is not None else
                            np.eye(3)[None, None] * full.contact_active[0, tick:k1, :, None, None]))
        nb = ProblemBatch.from_arrays(p, xi, xf, Xr, Ui)
        cold = E.solve_scp(nb, conf.scp_params, qp_overrides=ov or None)
        warm = E.solve_scp(nb, conf.scp_params, qp_overrides=dict(ov, warm_start=1))
        assert (cold["status"] == 0).all() and (warm["status"] == 0).all()
        np.testing.assert_array_equal(cold["n_accepted"], warm["n_accepted"])
        np.testing.assert_array_equal(cold["scp_iters"], warm["scp_iters"])
        # certified on the warm active set: no ADMM iteration (bolt: one instance is the known case whose polish cycles,
        # DESIGN.md section 5; it takes the ADMM route in the cold solve as well)
        if name in ("bolt", "talos"):   # (talos: an instance whose warm active set needs more than the 19 correction rounds)
            assert (warm["qp_iters"] == 0).mean() >= 0.5 and (warm["qp_iters"] <= cold["qp_iters"] + 16).all(), warm["qp_iters"]
        else:
            assert (warm["qp_iters"] == 0).all(), warm["qp_iters"]
        for i in range(4):
            assert relerr(warm["X"][i], cold["X"][i]) < TOL and relerr(warm["U"][i], cold["U"][i]) < TOL
        sol = warm


@pytest.mark.gpu
def test_receding_horizon_on_device(gpu):
    """The driver on the GPU: warm and cold loops apply the same controls (1e-6) over five ticks; warm ticks run
    without ADMM iterations; the first tile equals the host build bit for bit at every tick."""
    from centroidal_mpc_b200.mpc import RecedingHorizonMPC
    torch = gpu
    B, H = 64, 60
    warm = RecedingHorizonMPC("solo12_trot", B, H, warm=True)
    cold = RecedingHorizonMPC("solo12_trot", B, H, warm=False)
    for tick in range(5):
        x_w = [t.clone() for t in warm._keep]
        uw, sw = warm.step()
        uc, sc = cold.step()
        assert int(sw.sum()) == 0 and int(sc.sum()) == 0
        uw, uc = uw.cpu().numpy(), uc.cpu().numpy()
        assert np.linalg.norm(uw - uc) / np.linalg.norm(uc) < TOL
        st = warm.stats()
        if tick > 0:
            assert (st["qp_iters"] == 0).mean() > 0.9, st["qp_iters"]
        # host build on the problem the device just solved (first tile)
        k0 = tick
        full = synthetic.make_batch(synthetic.load_conf("solo12_trot"), 1)
        proto = dict(warm.solver.batch.proto, contact_pos=full.contact_pos[0, k0:k0 + H], contact_active=full.contact_active[0, k0:k0 + H],
                     contact_R=np.eye(3)[None, None] * full.contact_active[0, k0:k0 + H, :, None, None])
        nb = ProblemBatch.from_arrays(proto, *[t[:4].cpu().numpy() for t in x_w])
        emu = E.solve_scp(nb, warm.conf.scp_params, qp_overrides=dict(warm_start=1) if tick > 0 else None)
        np.testing.assert_array_equal(emu["U"][:, 0], uw[:4])
    assert not bool(warm.failed.any()) and not bool(cold.failed.any())
    warm.close(); cold.close()

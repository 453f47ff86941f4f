"""Loader of the synthetic trust-region scenarios and the N = 100 parity samples (tests/golden/scen_*.npz,
n100_*.npz; generator: tests/golden/make_scenarios.py)."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def scenario_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "scen*.npz")))


def scenario_qp(path):
    """QP settings of a scenario.  scen1 (penalty weight 1e4) is stiff: with multipliers of that size the
    default KKT certificate (primal residual <= 1e-9 (1 + norm)) leaves 5e-6 on the trajectories (the
    reference's own OSQP tolerance, 1e-7, leaves more); 1e-11 reaches the tightly solved oracle to 1e-7."""
    return dict(active_set_tol=1e-11) if "scen1" in os.path.basename(path) else None


def load_scenario(path):
    """-> (fixture, conf, scp_params, model) with the fixture's warm start / contact plan applied."""
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.src.centroidal_model import Centroidal_model
    g = np.load(path)
    conf = synthetic.load_conf(str(g["name"]), N=int(g["N"]))
    sp = dict(conf.scp_params)
    for k, v in zip(g["scp_keys"], g["scp_vals"]):
        sp[str(k)] = int(v) if str(k) == "max_iterations" else float(v)
    model = Centroidal_model(conf, centroidal_traj=np.asarray(g["X_ref"]).T)
    model._init_trajectories = dict(state=np.array(g["X_ref"]), control=np.array(g["U_init"]))
    model._x_init = model._init_trajectories["state"][:, 0].copy()
    model._x_final = model._init_trajectories["state"][:, -1].copy()
    if bool(g["free_fall"]):
        model._contact_data["contacts_logic"][:] = 0
    return g, conf, sp, model


def check_scenario(out, b, g, relerr, tol=1e-6):
    """Device / host-build result ``out`` (instance b) against the fixture: the reference's verdict, the SCP
    iteration and acceptance counts, the trust-region state and the trajectories."""
    if bool(g["returned_false"]):
        assert out["status"][b] != 0          # the reference returns False (scp_solver.py:146-148)
        return
    assert out["status"][b] == 0
    assert out["scp_iters"][b] == int(g["iterations"])
    assert out["n_accepted"][b] == int(g["n_accepted"])
    if int(g["n_accepted"]):
        assert relerr(out["X"][b].T, g["X"]) < tol and relerr(out["U"][b].T, g["U"]) < tol
    if "info" in out:
        assert abs(out["info"][b, 0] - float(g["snorm"][-1])) < 1e-6 * float(g["snorm"][-1])
        if not np.isnan(float(g["acc"][-1])):
            assert abs(out["info"][b, 1] - float(g["acc"][-1])) < 1e-4 * float(g["acc"][-1])


def n100_samples(name, mode, N=100):
    """-> (conf, full batch, batch of the sampled instances, fixture)."""
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.batch import ProblemBatch
    g = np.load(os.path.join(GOLDEN, "n%d_%s_mode%s.npz" % (N, name, mode)))
    conf = synthetic.load_conf(name, N=N)
    full = synthetic.make_batch(conf, int(g["batch"]), mode=mode)
    ids = np.asarray(g["ids"])
    sub = ProblemBatch.from_arrays(full.proto, full.x_init[ids], full.x_final[ids], full.X_ref[ids], full.U_init[ids])
    return conf, full, sub, g

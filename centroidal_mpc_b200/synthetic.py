"""Synthetic problem data for benchmarks and tests.

The reference's warm start comes from a crocoddyl whole-body DDP run that is not shipped
(``wholeBody_to_centroidal_traj.npz``; src/centroidal_model.py:87,174).  The stand-in is a
constant-velocity CoM reference with per-instance perturbations drawn from
``numpy.random.default_rng(1000 + b)``.
"""
import importlib
import types

import numpy as np

from .config import _robots

_COM_HEIGHT = {"solo12_trot": _robots.SOLO12_COM_HEIGHT, "solo12_pace": _robots.SOLO12_COM_HEIGHT,
               "solo12_bound": _robots.SOLO12_COM_HEIGHT, "bolt": _robots.BOLT_COM_HEIGHT,
               "talos": _robots.TALOS_COM_HEIGHT}


def load_conf(name, N=None):
    """Copy of ``config/conf_<name>`` as a namespace, optionally windowed to the first N knots
    (the contact plan is simply read for k < N; src/centroidal_model.py:132)."""
    mod = importlib.import_module(".config.conf_" + name, package=__package__)
    conf = types.SimpleNamespace(**{k: v for k, v in vars(mod).items() if not k.startswith("__")})
    conf.name = name
    if N is not None:
        if N > conf.N:
            raise ValueError("window N=%d exceeds the gait's horizon %d" % (N, conf.N))
        conf.N = int(N)
    return conf


def reference_trajectory(conf, b=0, mode="B", v=0.1):
    """(N+1, 9) warm-start / tracking trajectory for instance ``b``.

    mode 'B' (independent instances): CoM offset N(0,(1 cm)^2) per axis constant in k and
    angular-momentum noise N(0,(1e-3)^2) i.i.d. per knot.
    mode 'A' (perturbed initial state only): the shared nominal trajectory; the caller
    perturbs x_init (see ``perturbed_x_init``)."""
    N, m, dt = conf.N, conf.robot_mass, conf.dt
    if getattr(conf, "robot_name", "") == "TALOS":
        v = 0.0   # the feet step in place (stepLength 0): the CoM sways sideways, it does not travel
    k = np.arange(N + 1)
    X = np.zeros((N + 1, 9))
    X[:, 0] = v * dt * k
    X[:, 2] = _COM_HEIGHT.get(conf.name, 0.23)
    X[:, 3] = m * v
    if getattr(conf, "robot_name", "") == "TALOS":
        _lateral_sway(conf, X)
    if mode == "B":
        rng = np.random.default_rng(1000 + b)
        X[:, 0:3] += rng.normal(0.0, 0.01, size=3)[None, :]
        X[:, 6:9] += rng.normal(0.0, 1e-3, size=(N + 1, 3))
    return X


def _lateral_sway(conf, X, gain=0.85, half_window=6):
    """Flat-footed biped: the CoM reference sways over the stance foot (a biped cannot hold its CoM between the
    feet through a single-support phase: with the CoP confined to the sole the angular-momentum and lateral
    momentum rows of the terminal equality contradict each other).  y_ref = moving average of gain * mean y of
    the active feet, lateral momentum = m * dy/dt by finite differences."""
    from .src.contact_plan import create_contact_trajectory
    traj = create_contact_trajectory(conf)
    N = conf.N
    target = np.zeros(N + 1)
    for k in range(N + 1):
        kk = min(k, N - 1)
        ys = [traj[c][kk].pose.translation[1] for c in traj if traj[c][kk].ACTIVE]
        target[k] = gain * float(np.mean(ys)) if ys else 0.0
    pad = np.concatenate([np.full(half_window, target[0]), target, np.full(half_window, target[-1])])
    kernel = np.ones(2 * half_window + 1) / (2 * half_window + 1)
    y = np.convolve(np.convolve(pad, kernel, mode="same"), kernel, mode="same")[half_window:half_window + N + 1]
    X[:, 1] = y
    X[:, 4] = conf.robot_mass * np.gradient(y, conf.dt)


def perturbed_x_init(conf, b):
    """mode 'A' initial-state perturbation: N(0, diag(1 cm, 0.05 kg m/s, 0.01 kg m^2/s)^2)."""
    rng = np.random.default_rng(1000 + b)
    sig = np.array([0.01] * 3 + [0.05] * 3 + [0.01] * 3)
    return reference_trajectory(conf, 0, mode="A")[0] + rng.normal(0.0, 1.0, size=9) * sig


def make_batch(conf, B, mode="B", first=0, stochastic=False):
    """ProblemBatch of B synthetic instances sharing the contact plan of ``conf``.

    mode 'B': independent reference trajectories (every instance has its own linearisation);
    mode 'A': one shared reference, only x_init perturbed (BASELINE.json config 2).
    ``stochastic``: the instances are Centroidal_model(conf, STOCHASTIC_OCP=True)."""
    from .batch import ProblemBatch
    from .src.centroidal_model import Centroidal_model
    proto_model = Centroidal_model(conf, STOCHASTIC_OCP=stochastic,
                                   centroidal_traj=reference_trajectory(conf, first, mode=mode))
    proto = proto_model.problem_arrays()
    X_ref = np.stack([reference_trajectory(conf, first + b, mode=mode) for b in range(B)])   # [B,N+1,9]
    U_init = np.broadcast_to(proto["U_init"].T[None], (B,) + proto["U_init"].T.shape).copy()
    x_init = X_ref[:, 0].copy()
    x_final = X_ref[:, -1].copy()
    if mode == "A":
        x_init = np.stack([perturbed_x_init(conf, first + b) for b in range(B)])
    return ProblemBatch.from_arrays(proto, x_init, x_final, X_ref, U_init)

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def relerr(a, b):
    import numpy as np
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="session")
def cases():
    """Small synthetic problems shared by the tests: name -> (conf, [models])."""
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.src.centroidal_model import Centroidal_model
    out = {}
    for name, N in (("solo12_trot", 40), ("solo12_pace", 30), ("solo12_bound", 40), ("bolt", 40)):
        conf = synthetic.load_conf(name, N=N)
        out[name] = (conf, [Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, b))
                            for b in range(3)])
    return out


def with_weights_of(model, other):
    """Copy of ``model`` (its gait / contact plan / warm start) carrying the cost weights of ``other``: a batch
    shares one set of weights (ProblemBatch refuses anything else)."""
    import copy
    m = copy.copy(model)
    m._state_cost_weights = other._state_cost_weights
    m._control_cost_weights = other._control_cost_weights
    return m


@pytest.fixture(scope="session")
def gpu():
    """A CUDA device and the built library (the GPU tests skip without a device)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as g
    g.build()
    return torch

"""``Cost(Q, p)`` and the cost builders of the reference, as INSPECTORS of the device problem.

The device solver never materialises the QP (the knot records are the block-
banded KKT data).  These functions expand what the device works with -- the model's weights and the
linearisation point -- into the reference's objects, in the reference's variable order

    z = [x_0..x_N | u_0..u_{N-1} | t_0..t_N | N unused control slacks],

so that the implied (P, q) can be compared with an independent assembly (tests/test_host_api.py,
tests/test_gpu.py) or handed to any other QP solver.  Same names, arguments and return types as
/root/reference/src/cost.py:5-47; ``sum_up_all_costs`` lives in scp_solver.py as in the reference
(/root/reference/src/scp_solver.py:10-26).  Matrices are scipy.sparse (the reference builds dense
arrays of the same values)."""
from collections import namedtuple

import numpy as np
from scipy import sparse

Cost = namedtuple("Cost", "Q, p")


def construct_total_cost(model):
    """1/2 x'Wx x + 1/2 u'Wu u; slack blocks empty (/root/reference/src/cost.py:9-17)."""
    N = model._N
    Q = sparse.block_diag([sparse.kron(sparse.eye(N + 1), np.asarray(model._state_cost_weights, dtype=np.float64)),
                           sparse.kron(sparse.eye(N), np.asarray(model._control_cost_weights, dtype=np.float64)),
                           sparse.csc_matrix((N + 1, N + 1)), sparse.csc_matrix((N, N))], format="csc")
    return Cost(Q=Q, p=np.zeros(model._total_nb_optimizers))


def construct_state_tracking_cost(model):
    """q_x = -Wx xbar_k along the warm start (/root/reference/src/cost.py:21-29); the device keeps
    xbar_k per knot and forms this product on the fly (csrc/cmpc_tile.cuh: bwd_run)."""
    n_total = model._total_nb_optimizers
    gradient = np.zeros(n_total)
    com_x_indices = model._state_optimizers_indices["coms"][0]._optimizer_idx_vector
    tracking_traj = np.asarray(model._init_trajectories["state"], dtype=np.float64)
    Wx = np.asarray(model._state_cost_weights, dtype=np.float64)
    for time_idx in range(model._N + 1):
        i0 = com_x_indices[time_idx]
        gradient[i0:i0 + 9] = -Wx @ tracking_traj[:, time_idx]
    return Cost(Q=sparse.csc_matrix((n_total, n_total)), p=gradient)


def construct_state_trust_region_cost(model):
    """sum_k t_k: the exact L1 penalty of the trust region (/root/reference/src/cost.py:34-39).  On the
    device the slack is minimised out in closed form, t_k = w max(0, |kappa_k - kappa_bar_k|_1 - r)."""
    n_all, N = model._total_nb_optimizers, model._N
    s0 = model._n_x * (N + 1) + model._n_u * N
    p = np.zeros(n_all)
    p[s0:s0 + N + 1] = 1.0
    return Cost(Q=sparse.csc_matrix((n_all, n_all)), p=p)


def construct_control_trust_region_cost(model):
    """Defined by the reference but never added to the QP (/root/reference/src/cost.py:44-48)."""
    n_all, N = model._total_nb_optimizers, model._N
    p = np.zeros(n_all)
    p[n_all - N:] = 1.0
    return Cost(Q=sparse.csc_matrix((n_all, n_all)), p=p)

"""``Constraint(mat, lb, ub)`` and the constraint builders of the reference, as INSPECTORS of the
device problem (see cost.py).  Each builder expands the data the device solver works with into the
reference's rows, in the reference's row and column order:

  construct_initial_constraints            /root/reference/src/constraints.py:12-17
  construct_dynamics_constraints           :20-50     rows from traj_data = model.compute_trajectory_data(...)
                                                      (cmpc_linearize on the device)
  construct_final_constraints              :104-109
  construct_cop_constraints                :111-145   (TALOS)
  construct_friction_pyramid_constraints   :153-217   nominal rows; in stochastic mode the upper bounds are
                                                      the device's back-offs (cmpc_friction_backoffs)
  construct_state_trust_region_constraints :260-293

``stack_up_all_constraints`` lives in scp_solver.py as in the reference (:28-48).  Matrices are
scipy.sparse CSC; values, row order and bounds are the reference's."""
from collections import namedtuple

import numpy as np
from scipy import sparse

from .utils import construct_friction_pyramid_constraint_matrix

Constraint = namedtuple("Constraint", "mat, lb, ub")


def _coo(rows, cols, vals, shape):
    return sparse.csc_matrix((np.asarray(vals, dtype=np.float64), (np.asarray(rows, dtype=int), np.asarray(cols, dtype=int))),
                             shape=shape)


def construct_initial_constraints(model):
    nx = model._n_x
    A = _coo(range(nx), range(nx), np.ones(nx), (nx, model._total_nb_optimizers))
    x_init = np.asarray(model._x_init, dtype=np.float64)
    return Constraint(mat=A, lb=x_init, ub=x_init)


def construct_final_constraints(model):
    nx, N = model._n_x, model._N
    A = _coo(range(nx), nx * N + np.arange(nx), np.ones(nx), (nx, model._total_nb_optimizers))
    x_final = np.asarray(model._x_final, dtype=np.float64)
    return Constraint(mat=A, lb=x_final, ub=x_final)


def construct_dynamics_constraints(model, prev_traj_tuple, traj_data):
    """A_k x_k + B_k u_k - x_{k+1} = A_k xbar_k + B_k ubar_k - f(xbar_k, ubar_k) (+-1e-12).  The device
    eliminates these rows exactly through the Riccati recursion; its A_k, B_k, f are the ones of
    ``traj_data``."""
    nx, nu, N = model._n_x, model._n_u, model._N
    X = np.asarray(prev_traj_tuple["state"], dtype=np.float64)
    U = np.asarray(prev_traj_tuple["control"], dtype=np.float64)
    A_traj = np.asarray(traj_data["gradients"]["f_x"], dtype=np.float64)
    B_traj = np.asarray(traj_data["gradients"]["f_u"], dtype=np.float64)
    F_traj = np.asarray(traj_data["dynamics"], dtype=np.float64)
    x_idx = model._state_optimizers_indices["coms"][0]._optimizer_idx_vector
    first = next(iter(model._control_optimizers_indices))
    key = "cops" if model._robot == "TALOS" else "forces"
    u_idx = model._control_optimizers_indices[first][key][0]._optimizer_idx_vector
    rows, cols, vals = [], [], []
    lb, ub = np.zeros(nx * N), np.zeros(nx * N)
    for k in range(N):
        r0 = nx * k
        Ak, Bk = A_traj[k], B_traj[k]
        i, j = np.nonzero(Ak)
        rows += list(r0 + i); cols += list(x_idx[k] + j); vals += list(Ak[i, j])
        i, j = np.nonzero(Bk)
        rows += list(r0 + i); cols += list(u_idx[k] + j); vals += list(Bk[i, j])
        rows += list(r0 + np.arange(nx)); cols += list(x_idx[k] + nx + np.arange(nx)); vals += [-1.0] * nx
        lin = Ak @ X[:, k] + Bk @ U[:, k] - F_traj[:, k]
        lb[r0:r0 + nx] = lin - 1e-12
        ub[r0:r0 + nx] = lin + 1e-12
    return Constraint(mat=_coo(rows, cols, vals, (nx * N, model._total_nb_optimizers)), lb=lb, ub=ub)


def construct_cop_constraints(model):
    """CoP inside the foot rectangle, TALOS (/root/reference/src/constraints.py:111-145): per contact the
    x rows of all knots, then the y rows."""
    n_all, N = model._total_nb_optimizers, model._N
    logic = model._contact_data["contacts_logic"]
    rows, cols, vals, lb, ub = [], [], [], [], []
    r0 = 0
    for contact in model._contact_trajectory:
        c = list(model._contact_trajectory).index(contact)
        cops = model._control_optimizers_indices[contact]["cops"]
        for ax, key in enumerate(("x", "y")):
            rng = model._robot_foot_range[key]
            for k in range(N):
                if logic[k, c]:
                    rows.append(r0 + k); cols.append(cops[ax]._optimizer_idx_vector[k]); vals.append(1.0)
                    lb.append(-rng[1]); ub.append(rng[0])
                else:
                    lb.append(0.0); ub.append(0.0)
            r0 += N
    return Constraint(mat=_coo(rows, cols, vals, (r0, n_all)), lb=np.array(lb), ub=np.array(ub))


def construct_friction_pyramid_constraints(model, prev_traj_tuple=None, traj_data=None, friction_ub=None):
    """G_{k,c} f_{k,c} <= ub, contact-major, 5 rows per knot of which the reference writes 4
    (/root/reference/src/constraints.py:153-185,215-217).  Stochastic mode (:157-163,187-214): the upper
    bounds are the chance-constraint back-offs; ``friction_ub`` (N, nc, 4) are the device's
    (cmpc_friction_backoffs, computed when omitted).  The covariance-gradient terms of the reference are
    identically zero and add nothing to ``mat``."""
    n_all, N = model._total_nb_optimizers, model._N
    pyr = construct_friction_pyramid_constraint_matrix(model)
    nrow = pyr.shape[0]
    if model._STOCHASTIC_OCP and friction_ub is None:
        from .. import device
        friction_ub = device.friction_backoffs(model, prev_traj_tuple)
    rows, cols, vals = [], [], []
    ub_total = []
    r0 = 0
    for c, contact in enumerate(model._contact_trajectory):
        force_opt = model._control_optimizers_indices[contact]["forces"]
        ub = np.zeros(nrow * N)
        for k in range(N):
            d = model._contact_trajectory[contact][k]
            if d.ACTIVE:
                G = pyr @ np.asarray(d.pose.rotation, dtype=np.float64).T
                for j in range(4):
                    for a, opt in enumerate(force_opt):
                        if G[j, a] != 0.0:
                            rows.append(r0 + k * nrow + j); cols.append(opt._optimizer_idx_vector[k]); vals.append(G[j, a])
                if friction_ub is not None:
                    ub[k * nrow:k * nrow + 4] = np.asarray(friction_ub)[k, c]
        ub_total.append(ub)
        r0 += nrow * N
    return Constraint(mat=_coo(rows, cols, vals, (r0, n_all)), lb=-np.inf * np.ones(r0), ub=np.concatenate(ub_total))


def construct_state_trust_region_constraints(model, prev_traj_tuple, trust_region):
    """s_j'(kappa_k - kappa_bar_k) - t_k / w <= r for the 8 sign patterns, then -t_k <= 0
    (/root/reference/src/constraints.py:260-293).  The device handles these rows through the prox of the
    exact penalty (csrc/cmpc_tile.cuh: prox_trust) and, in the polish, as pinned / surface rows."""
    nx, n_all, N = model._n_x, model._total_nb_optimizers, model._N
    opt = model._state_slack_optimizers_indices
    S = opt._penum_mat
    nb = S.shape[0]
    X = np.asarray(prev_traj_tuple["state"], dtype=np.float64)
    w, r = float(trust_region["weight"]), float(trust_region["radius"])
    rows, cols, vals = [], [], []
    ub = np.zeros(nb * (N + 1))
    for k in range(N + 1):
        x6 = opt._x0_optimizer_idx_vector[k] + 6
        st = opt._slack_optimizers_idx_vector[k]
        for j in range(nb):
            for a in range(3):
                rows.append(k * nb + j); cols.append(x6 + a); vals.append(S[j, a])
            rows.append(k * nb + j); cols.append(st); vals.append(-1.0 / w)
        ub[k * nb:(k + 1) * nb] = r + S @ X[6:, k]
    nslack = model._n_t * (N + 1)
    s0 = nx * (N + 1) + model._n_u * N
    rows += list(nb * (N + 1) + np.arange(nslack)); cols += list(s0 + np.arange(nslack)); vals += [-1.0] * nslack
    m = nb * (N + 1) + nslack
    return Constraint(mat=_coo(rows, cols, vals, (m, n_all)),
                      ub=np.hstack([ub, np.zeros(nslack)]), lb=-np.inf * np.ones(m))

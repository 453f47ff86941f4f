"""``solve_scp`` — the reference's entry point (/root/reference/src/scp_solver.py:118-179) on
top of the CUDA library, plus the batched entry ``solve_scp_batched``.

The whole loop — linearisation (compute_trajectory_data), QP assembly (sum_up_all_costs,
stack_up_all_constraints), the QP solve (solve_subproblem), get_QP_solution, the spectral-norm
trust test, compute_model_accuracy and the radius/weight updates — runs inside one kernel per
batch (csrc/cmpc_api.cu: cmpc_scp_kernel); Python only packs inputs and unpacks outputs.
"""
from warnings import warn

import numpy as np

from ..batch import ProblemBatch


def solve_scp_batched(batch_or_models, scp_params, qp_settings=None, solver=None, return_stats=False):
    """Solve B independent SCP problems.  Returns dict(X [B,N+1,9], U [B,N,nu], scp_iters [B],
    status [B], n_accepted [B]) as numpy arrays (plus solver statistics when asked).
    ``status != 0`` marks instances for which the reference would return False (QP failure);
    ``n_accepted == 0`` marks instances for which it would return empty lists."""
    from ..device import BatchSolver
    batch = batch_or_models if isinstance(batch_or_models, ProblemBatch) else ProblemBatch(list(batch_or_models))
    own = solver is None
    if own:
        solver = BatchSolver(batch)
    try:
        solver.solve(scp_params, qp_settings)
        out = solver.results()
        if return_stats:
            out.update(solver.stats())
    finally:
        if own:
            solver.close()
    return out


def solve_scp(model, scp_params):
    """Drop-in for the reference's solve_scp(model, scp_params): returns
    dict(state=[X (9,N+1)], control=[U (n_u,N)], gains=[...], covs=[...]) with one entry per
    accepted SCP iteration, empty lists if nothing was accepted, or False when a QP subproblem
    failed (scp_solver.py:146-148)."""
    out = solve_scp_batched([model], scp_params)
    if int(out["status"][0]) != 0:
        warn("[solve_OSQP]: Problem unfeasible.")
        return False
    all_solution = dict(state=[], control=[], gains=[], covs=[])
    gains = covs = None
    if int(out["n_accepted"][0]) > 0:
        # traj_data is computed once, at the warm start (scp_solver.py:129-130), so every accepted
        # iterate carries the same LQR_gains / Covs (scp_solver.py:165-166)
        from ..device import lqr_gains_covs
        gains, covs = lqr_gains_covs(model, model._init_trajectories)
    # the linearisation point never moves (scp_solver.py:129-130), so every accepted iterate
    # solves the same QP; the last accepted one is what the kernel returns
    for _ in range(int(out["n_accepted"][0])):
        all_solution["state"].append(out["X"][0].T.copy())
        all_solution["control"].append(out["U"][0].T.copy())
        all_solution["gains"].append(gains)
        all_solution["covs"].append(covs)
    return all_solution


def sum_up_all_costs(model):
    """Cost(Q, p) of the QP in the reference's variable order (/root/reference/src/scp_solver.py:10-26): the
    inspector view of what the device minimises (src/cost.py)."""
    from scipy import sparse
    from .cost import Cost, construct_state_tracking_cost, construct_state_trust_region_cost, construct_total_cost
    total = [construct_total_cost(model), construct_state_trust_region_cost(model)]
    if model._robot != "TALOS" and not model._DYNAMICS_FIRST:
        total.append(construct_state_tracking_cost(model))
    n = model._total_nb_optimizers
    Q, p = sparse.csc_matrix((n, n)), np.zeros(n)
    for c in total:
        Q = Q + c.Q
        p = p + c.p
    return Cost(Q=sparse.csc_matrix(Q), p=p)


def stack_up_all_constraints(model, traj_tuple, traj_data, trust_region_updates, friction_ub=None):
    """Constraint(mat, lb, ub) of the QP in the reference's row order: initial, dynamics, final, (CoP,)
    friction pyramid, state trust region (/root/reference/src/scp_solver.py:28-48).  ``traj_data`` is
    ``model.compute_trajectory_data(traj_tuple)`` (the device's linearisation)."""
    from scipy import sparse
    from . import constraints as C
    parts = [C.construct_initial_constraints(model), C.construct_dynamics_constraints(model, traj_tuple, traj_data),
             C.construct_final_constraints(model)]
    if model._robot == "TALOS":
        parts.append(C.construct_cop_constraints(model))
    parts += [C.construct_friction_pyramid_constraints(model, traj_tuple, traj_data, friction_ub=friction_ub),
              C.construct_state_trust_region_constraints(model, traj_tuple, trust_region_updates)]
    return C.Constraint(mat=sparse.vstack([c.mat for c in parts], "csc"), lb=np.hstack([c.lb for c in parts]),
                        ub=np.hstack([c.ub for c in parts]))


def get_QP_solution(model, z):
    """Column-major unpack of a reference-ordered decision vector (scp_solver.py:89-93)."""
    n_x, n_u, N = model._n_x, model._n_u, model._N
    X_sol = np.reshape(z[:n_x * (N + 1)], (n_x, N + 1), order="F")
    U_sol = np.reshape(z[n_x * (N + 1):n_x * (N + 1) + n_u * N], (n_u, N), order="F")
    return dict(state=X_sol, control=U_sol)


def interpolate_SCP_solution(solution):
    """Linear x10 up-sampling of the last accepted (X, U) (scp_solver.py:95-111), vectorised with
    the reference's arithmetic (``M[:, i] + j * ((M[:, i+1] - M[:, i]) / 10)``) so the result is
    bit-identical to its double loop."""
    N_inner = 10
    X, U = solution["state"][-1], solution["control"][-1]
    j = np.arange(N_inner, dtype=float)

    def up(M):
        d = (M[:, 1:] - M[:, :-1]) / float(N_inner)
        seg = M[:, :-1, None] + j[None, None, :] * d[:, :, None]
        return seg.reshape(M.shape[0], -1)
    return dict(X=up(X), U=up(U))


# File names and keys of the hand-off between the reference's pipeline stages
# (build/lib/demos/run_motion.py:30,41-43; read back by src/whole_body_control.py:41-44 and
# src/centroidal_model.py:87,174).
WHOLEBODY_TO_CENTROIDAL = "wholeBody_to_centroidal_traj.npz"   # X [N+1, 9]: DDP -> SCP warm start
CENTROIDAL_TO_WHOLEBODY = "centroidal_to_wholeBody_traj.npz"   # X (9, N+1), U (n_u, N): SCP -> DDP
SCP_INTERPOLATED = "scp_sol_interpol_nom.npz"                  # X (9, 10 N), U (n_u, 10 (N-1))


def save_scp_handoff(solution, directory="."):
    """Write the two files the reference's whole-body stage consumes after ``solve_scp``
    (run_motion.py:41-43): the x10 interpolated solution and the last accepted (X, U), with the
    reference's file names, keys and array layouts.  Returns the two paths."""
    import os
    if solution is False or not solution["state"]:
        raise ValueError("no accepted SCP iterate to hand off")
    ip = interpolate_SCP_solution(solution)
    p1 = os.path.join(directory, SCP_INTERPOLATED)
    p2 = os.path.join(directory, CENTROIDAL_TO_WHOLEBODY)
    np.savez(p1, X=ip["X"], U=ip["U"])
    np.savez(p2, X=solution["state"][-1], U=solution["control"][-1])
    return p1, p2


def load_scp_handoff(directory="."):
    """Read ``centroidal_to_wholeBody_traj.npz`` the way whole_body_control.py:41-44 does."""
    import os
    f = np.load(os.path.join(directory, CENTROIDAL_TO_WHOLEBODY))
    return f["X"], f["U"]

"""Index bookkeeping: (variable name, knot) -> position in the reference's decision vector

    z = [x_0..x_N | u_0..u_{N-1} | t_0..t_N | N unused control slacks].

Same class and attribute names as /root/reference/src/optimizer.py (State_optimizer :5-35,
Control_optimizer :37-74, Slack_optimizer :94-133) so code that inspects
``model._state_optimizers_indices`` etc. keeps working.  The device solver itself uses a
stage-ordered layout; these objects define the drop-in view of it.
"""
import numpy as np

_STATE_NAMES = ("com_x", "com_y", "com_z", "lin_mom_x", "lin_mom_y", "lin_mom_z",
                "ang_mom_x", "ang_mom_y", "ang_mom_z")
_CONTROL_NAMES = {"solo12": ("fx", "fy", "fz"),
                  "TALOS": ("cop_x", "cop_y", "fx", "fy", "fz", "tau_z")}


class State_optimizer:
    def __init__(self, OPTIMIZER_IDENTIFIER, nb_x_optimizers, horizon_length):
        self._name = OPTIMIZER_IDENTIFIER
        self.nx = nb_x_optimizers
        self.N = horizon_length
        if self._name not in _STATE_NAMES:
            raise ValueError("not a state optimizer name: %r" % (self._name,))
        self._optimizer_idx = _STATE_NAMES.index(self._name)
        self._optimizer_idx_vector = np.arange(self.N + 1, dtype=int) * self.nx + self._optimizer_idx


class Control_optimizer:
    def __init__(self, OPTIMIZER_IDENTIFIER, contact_idx, robot_name, nb_x_optimizers,
                 nb_u_ptimizers, horizon_length):
        self._name = OPTIMIZER_IDENTIFIER
        self.nx = nb_x_optimizers
        self.nu = nb_u_ptimizers
        self.N = horizon_length
        names = _CONTROL_NAMES["TALOS" if robot_name == "TALOS" else "solo12"]
        if self._name not in names:
            raise ValueError("not a control optimizer name: %r" % (self._name,))
        self._optimizer_idx = names.index(self._name)
        per_contact = len(names)
        self._optimizer_idx_vector = (self.nx * (self.N + 1) + np.arange(self.N, dtype=int) * self.nu
                                      + per_contact * contact_idx + self._optimizer_idx)


class Slack_optimizer:
    """'state': L1 trust region on the 3 angular-momentum states (2^3 sign patterns);
    'control': the reference's unused control trust region (2^nu patterns)."""

    def __init__(self, OPTIMIZER_IDENTIFIER, nx_optimizers, nu_optimizers, nt_optimizers,
                 horizon_length):
        self._name = OPTIMIZER_IDENTIFIER
        self.N = horizon_length
        self.nx = nx_optimizers
        self.nu = nu_optimizers
        self.nt = nt_optimizers
        N, nx, nu, nt = self.N, self.nx, self.nu, self.nt
        if self._name == "state":
            width = nx - 6
            self._x0_optimizer_idx_vector = np.arange(N + 1, dtype=int) * nx
            self._slack_optimizers_idx_vector = nx * (N + 1) + nu * N + np.arange(N + 1, dtype=int)
        elif self._name == "control":
            width = nu
            self._u0_optimizer_idx_vector = nx * (N + 1) + np.arange(N, dtype=int) * nu
            self._slack_optimizers_idx_vector = (nx * (N + 1) + nu * N + nt * (N + 1)
                                                 + np.arange(N, dtype=int))
        else:
            raise ValueError("not a slack optimizer name: %r" % (self._name,))
        self._nb_slack_constraints = nt * 2 ** width
        j = np.arange(2 ** width)[:, None]
        i = np.arange(width)[None, :]
        self._penum_mat = (-1.0) ** (j // (2 ** i))       # s_j[i] = (-1)^(j // 2^i)

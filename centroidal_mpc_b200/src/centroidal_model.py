"""``Centroidal_model``: the problem object of the SCP hot path (host side, numpy only).

Mirrors /root/reference/src/centroidal_model.py:
  __init__ / attribute names           :15-47
  init and final states                :80-89
  optimizer indices                    :91-126
  contact data flattening              :127-156
  warm-start trajectories              :158-187
  integrate_model_one_step             :189-212   } thin wrappers over the CUDA kernels
  integrate_dynamics_trajectory        :243-255   } (csrc/cmpc_kernels.cu); there is no
  compute_trajectory_data              :257-291   } numpy/JAX fallback
JAX is not a dependency; arrays are float64 numpy (the reference's default-JAX float32
rounding is documented in DESIGN.md, "precision").
"""
from warnings import warn

import numpy as np

from .contact_plan import create_contact_trajectory
from .optimizer import Control_optimizer, Slack_optimizer, State_optimizer


class Centroidal_model:
    def __init__(self, conf, STOCHASTIC_OCP=False, centroidal_traj=None):
        """``centroidal_traj``: optional (N+1, 9) warm-start trajectory replacing the
        ``wholeBody_to_centroidal_traj.npz`` file the reference reads from the CWD."""
        self._DYNAMICS_FIRST = conf.DYNAMICS_FIRST
        self._robot = conf.robot_name
        self._n_x = conf.n_x
        self._n_u_per_contact = conf.n_u_per_contact
        self._n_u = conf.n_u
        self._n_w = conf.n_w
        self._n_t = conf.n_t
        self._N = conf.N
        self._total_nb_optimizers = (self._n_x * (self._N + 1) + self._n_u * self._N
                                     + self._n_t * (self._N + 1) + self._n_t * self._N)
        self._max_leg_length = conf.max_leg_length
        self._m = conf.robot_mass
        self._g = conf.gravity_constant
        self._dt = conf.dt
        self._state_cost_weights = conf.state_cost_weights
        self._control_cost_weights = conf.control_cost_weights
        self._linear_friction_coefficient = conf.mu
        self._Q = conf.Q
        self._R = conf.R
        if self._robot == "TALOS":
            self._robot_foot_range = {"x": np.array([conf.lxp, conf.lxn]),
                                      "y": np.array([conf.lyp, conf.lyn])}
        self._STOCHASTIC_OCP = STOCHASTIC_OCP
        self._beta_u = conf.beta_u
        self._Cov_w = conf.cov_w
        self._Cov_eta = conf.cov_white_noise
        self._warm_start = self._load_warm_start(conf, centroidal_traj)
        self._set_init_and_final_states(conf)
        self._fill_contact_data(conf)
        self._fill_optimizer_indices()
        self._fill_initial_trajectory()

    # ------------------------------------------------------------------ construction
    def _load_warm_start(self, conf, centroidal_traj):
        if conf.DYNAMICS_FIRST:
            return None
        if centroidal_traj is None:
            centroidal_traj = getattr(conf, "wholeBody_to_centroidal_traj", None)
        if centroidal_traj is None:
            # the reference's behaviour: file in the current directory (:87, :174)
            centroidal_traj = np.load("wholeBody_to_centroidal_traj.npz")["X"]
        traj = np.asarray(centroidal_traj, dtype=np.float64)
        if traj.shape != (self._N + 1, self._n_x):
            raise ValueError("warm start must have shape (N+1, n_x) = (%d, %d), got %r"
                             % (self._N + 1, self._n_x, traj.shape))
        return traj

    def _set_init_and_final_states(self, conf):
        if conf.DYNAMICS_FIRST:
            self._com_z = conf.com_z
            self._x_init = np.asarray(conf.x_init, dtype=np.float64)
            self._x_final = np.asarray(conf.x_final, dtype=np.float64)
        else:
            self._x_init = self._warm_start[0].copy()
            self._x_final = self._warm_start[-1].copy()

    def _fill_optimizer_indices(self):
        names = ["com_x", "com_y", "com_z", "lin_mom_x", "lin_mom_y", "lin_mom_z",
                 "ang_mom_x", "ang_mom_y", "ang_mom_z"]
        so = [State_optimizer(n, self._n_x, self._N) for n in names]
        # the reference's own (odd) grouping is kept: :101-103
        self._state_optimizers_indices = {"coms": so[:2], "lin_moms": so[2:5], "ang_moms": so[5:]}
        if self._robot == "TALOS":
            u_names = ["cop_x", "cop_y", "fx", "fy", "fz", "tau_z"]
        else:
            if self._robot != "solo12":
                warn("robot %r is handled with the point-contact ('solo12') force model" % self._robot)
            u_names = ["fx", "fy", "fz"]
        self._control_optimizers_indices = {}
        for contact_idx, contact in enumerate(self._contact_trajectory):
            co = [Control_optimizer(n, contact_idx, "TALOS" if self._robot == "TALOS" else "solo12",
                                    self._n_x, self._n_u, self._N) for n in u_names]
            if self._robot == "TALOS":
                entry = {"cops": co[:2], "forces": co[2:5], "moment": co[5:]}
            else:
                entry = {"forces": co}
            self._control_optimizers_indices[contact] = entry
        self._state_slack_optimizers_indices = Slack_optimizer("state", self._n_x, self._n_u,
                                                               self._n_t, self._N)

    def _fill_contact_data(self, conf):
        traj = create_contact_trajectory(conf)
        N, nc = self._N, len(traj)
        logic = np.zeros((N, nc), dtype=np.int32)
        orient = np.zeros((N, nc, 3, 3))
        position = np.zeros((N, nc * 3))
        for c, contact in enumerate(traj):
            for k in range(N):
                d = traj[contact][k]
                if d.ACTIVE:                       # inactive: logic 0, R = 0, p = 0 (:142-145)
                    logic[k, c] = 1
                    orient[k, c] = d.pose.rotation
                    position[k, 3 * c:3 * c + 3] = d.pose.translation
        self._contact_trajectory = traj
        self._contact_data = dict(contacts_logic=logic, contacts_orient=orient,
                                  contacts_position=position)

    def _fill_initial_trajectory(self):
        N = self._N
        init = {"state": np.zeros((self._n_x, N + 1)), "control": np.zeros((self._n_u, N))}
        if not self._DYNAMICS_FIRST:
            init["state"] = self._warm_start.T.copy()
            weight = -self._m * self._g
            logic = self._contact_data["contacts_logic"]
            npc = self._n_u_per_contact
            f0 = 2 if self._robot == "TALOS" else 0     # forces sit after the CoP slots for TALOS;
            # the reference indexes contact_idx*3 regardless - not copied.
            for k in range(N):
                n_active = int(np.sum(logic[k]))
                for c in range(logic.shape[1]):
                    if logic[k, c]:
                        init["control"][npc * c + f0:npc * c + f0 + 3, k] = [1e-3, 1e-3, weight / n_active]
        self._init_trajectories = init

    # ------------------------------------------------------------------ problem export
    def problem_arrays(self):
        """Plain-array view of the problem (what the C-ABI and the oracle consume)."""
        N, nc = self._N, self._contact_data["contacts_logic"].shape[1]
        prob = dict(
            N=N, robot="TALOS" if self._robot == "TALOS" else "solo12",
            m=float(self._m), g=float(self._g), dt=float(self._dt),
            mu=float(self._linear_friction_coefficient),
            state_cost_weights=np.asarray(self._state_cost_weights, dtype=np.float64),
            control_cost_weights=np.asarray(self._control_cost_weights, dtype=np.float64),
            x_init=np.asarray(self._x_init, dtype=np.float64),
            x_final=np.asarray(self._x_final, dtype=np.float64),
            X_ref=np.asarray(self._init_trajectories["state"], dtype=np.float64),
            U_init=np.asarray(self._init_trajectories["control"], dtype=np.float64),
            contact_pos=self._contact_data["contacts_position"].reshape(N, nc, 3).astype(np.float64),
            contact_R=self._contact_data["contacts_orient"].astype(np.float64),
            contact_active=self._contact_data["contacts_logic"].astype(np.int32),
            DYNAMICS_FIRST=bool(self._DYNAMICS_FIRST),
            # stochastic mode (constraints.py:157-163,187-214): what the friction back-offs need
            stochastic=dict(beta_u=float(self._beta_u), Q=np.asarray(self._Q, dtype=np.float64),
                            R=np.asarray(self._R, dtype=np.float64),
                            cov_w=np.asarray(self._Cov_w, dtype=np.float64),
                            cov_eta=np.asarray(self._Cov_eta, dtype=np.float64)) if self._STOCHASTIC_OCP else None,
        )
        if self._robot == "TALOS":
            prob["foot_range"] = {k: np.asarray(v, dtype=np.float64)
                                  for k, v in self._robot_foot_range.items()}
        return prob

    # ------------------------------------------------------------------ device wrappers
    def integrate_model_one_step(self, x, u, contacts_position_all, contacts_logic_all,
                                 contacts_orientation_all):
        from .. import device
        return device.integrate_one_step(self, x, u, contacts_position_all, contacts_logic_all,
                                         contacts_orientation_all)

    def integrate_dynamics_trajectory(self, traj_tuple):
        from .. import device
        return device.integrate_dynamics_trajectory(self, traj_tuple)

    def compute_trajectory_data(self, traj_tuple):
        from .. import device
        return device.compute_trajectory_data(self, traj_tuple)

"""solo12 pace.  Follows /root/reference/config/conf_solo12_pace.py (:9-98)."""
import numpy as np

from ..src.contact_plan import create_contact_sequence
from . import _robots

DYNAMICS_FIRST = False
dt = 0.01
dt_ctrl = 0.001
gait = {"type": "PACE", "stepLength": 0.0, "stepHeight": 0.05, "stepKnots": 10,
        "supportKnots": 3, "nbSteps": 4}
mu = 0.5

robot_name = "solo12"
ee_frame_names = ["FL_FOOT", "FR_FOOT", "HL_FOOT", "HR_FOOT"]
rmodel = _robots.solo12()
rdata = rmodel.createData()
robot_mass = _robots.SOLO12_MASS
gravity_constant = -9.81
max_leg_length = 0.34
foot_scaling = 1.0
lxp = lxn = lyp = lyn = 0.01

n_u_per_contact = 3
nb_contacts = 4
n_u = nb_contacts * n_u_per_contact
n_x = 9
n_t = 1

q0 = None
gait_templates, contact_sequence = create_contact_sequence(dt, gait, ee_frame_names, rmodel, rdata, q0)
N = int(round(contact_sequence[-1][0].t_end / dt, 2))
N_ctrl = int((N - 1) * (dt / dt_ctrl))

Q = np.diag([1e4] * 3 + [1e3] * 6)
R = np.diag([1e2, 5e2, 1e1] * 4)

n_w = nb_contacts * 3
cov_w, cov_white_noise = _robots.quadruped_noise(dt, 0.3, [0.7, 0.5, 0.01, 0.8, 0.6, 0.01, 0.7, 0.5, 0.01])
beta_u = 0.01

state_cost_weights = np.diag([1e4] * 3 + [1e3] * 3 + [1e5] * 3)
control_cost_weights = np.diag([1e2, 1e2, 1e1] * 4)

scp_params = {"trust_region_radius0": 50, "omega0": 100, "omega_max": 1.0e10, "epsilon": 1.0e-6,
              "rho0": 0.4, "rho1": 1.5, "beta_succ": 2.0, "beta_fail": 0.5, "gamma_fail": 5,
              "convergence_threshold": 1e-3, "max_iterations": 20}

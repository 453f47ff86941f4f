"""bolt biped, point-foot contacts — a SYNTHETIC centroidal config.

/root/reference/config/conf_bolt.py (:1-79) is whole-body (DDP) only: it has no n_x,
robot_mass, weights or scp_params, and the bolt branches of the reference contact planner
are commented out (src/contact_plan.py:132,173).  BASELINE.json config 4 therefore has to
be defined here: gait table from conf_bolt.py:59-64, two point feet under the reference's
point-contact ('solo12') force model, solo12-pace cost weights and scp_params
(SURVEY.md section 8d)."""
import numpy as np

from ..src.contact_plan import create_contact_sequence
from . import _robots

DYNAMICS_FIRST = False
dt = 0.01
dt_ctrl = 0.001
gait = {"type": "PACE", "stepLength": 0.0, "stepHeight": 0.05, "stepKnots": 10,
        "supportKnots": 2, "nbSteps": 5}
mu = 0.5

robot_name = "solo12"          # selects the 3-D point-force contact model (centroidal_model.py:104-109)
ee_frame_names = ["FL_ANKLE", "FR_ANKLE"]
rmodel = _robots.bolt()
rdata = rmodel.createData()
robot_mass = _robots.BOLT_MASS
gravity_constant = -9.81
max_leg_length = 0.4
foot_scaling = 1.0
lxp = lxn = lyp = lyn = 0.01

n_u_per_contact = 3
nb_contacts = 2
n_u = nb_contacts * n_u_per_contact
n_x = 9
n_t = 1

q0 = None
gait_templates, contact_sequence = create_contact_sequence(dt, gait, ee_frame_names, rmodel, rdata, q0)
N = int(round(contact_sequence[-1][0].t_end / dt, 2))
N_ctrl = int((N - 1) * (dt / dt_ctrl))

Q = np.diag([1e4] * 3 + [1e3] * 6)
R = np.diag([1e2, 5e2, 1e1] * nb_contacts)

n_w = nb_contacts * 3
cov_w = np.diag([0.4 ** 2, 0.4 ** 2, 0.3 ** 2] * nb_contacts)
cov_white_noise = dt * np.diag(np.array([0.7, 0.5, 0.01, 0.8, 0.6, 0.01, 0.7, 0.5, 0.01]) ** 2)
beta_u = 0.01

state_cost_weights = np.diag([1e4] * 3 + [1e3] * 3 + [1e5] * 3)
control_cost_weights = np.diag([1e2, 1e2, 1e1] * nb_contacts)

scp_params = {"trust_region_radius0": 50, "omega0": 100, "omega_max": 1.0e10, "epsilon": 1.0e-6,
              "rho0": 0.4, "rho1": 1.5, "beta_succ": 2.0, "beta_fail": 0.5, "gamma_fail": 5,
              "convergence_threshold": 1e-3, "max_iterations": 20}

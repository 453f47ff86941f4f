"""talos legs, flat feet with CoP / wrench contacts — a SYNTHETIC centroidal config.

/root/reference/config/conf_talos.py (:1-49) is whole-body (DDP) only: gait table, dt, mu and the frame
names, but no n_x, robot_mass, foot ranges, weights or scp_params.  BASELINE.json config 5 therefore has to
be defined here: gait table from conf_talos.py:18-23, two flat feet under the
reference's 'TALOS' contact model (six controls per foot: cop_x, cop_y, fx, fy, fz, tau_z;
src/centroidal_model.py:204-208, src/optimizer.py:48-64), the CoP box from the foot half-lengths
(src/centroidal_model.py:36-38, src/constraints.py:111-145), control weights scaled to the 90 kg robot
(forces of ~450-900 N against centimetres of CoP travel), the scp_params of the solo12 pace config."""
import numpy as np

from ..src.contact_plan import create_contact_sequence
from . import _robots

DYNAMICS_FIRST = False
dt = 0.03
dt_ctrl = 0.001
gait = {"type": "PACE", "stepLength": 0.0, "stepHeight": 0.1, "stepKnots": 15,
        "supportKnots": 5, "nbSteps": 4}
mu = 0.5

robot_name = "TALOS"           # selects the CoP / wrench contact model (centroidal_model.py:204-208)
ee_frame_names = ["left_sole_link", "right_sole_link"]
rmodel = _robots.talos()
rdata = rmodel.createData()
robot_mass = _robots.TALOS_MASS
gravity_constant = -9.81
max_leg_length = 1.1
foot_scaling = 1.0
lxp = 0.10   # foot length in positive x direction
lxn = 0.10   # ... negative x
lyp = 0.05   # ... positive y
lyn = 0.05   # ... negative y

n_u_per_contact = 6
nb_contacts = 2
n_u = nb_contacts * n_u_per_contact
n_x = 9
n_t = 1

q0 = None
gait_templates, contact_sequence = create_contact_sequence(dt, gait, ee_frame_names, rmodel, rdata, q0)
N = int(round(contact_sequence[-1][0].t_end / dt, 2))
N_ctrl = int((N - 1) * (dt / dt_ctrl))

Q = np.diag([1e4] * 3 + [1e3] * 6)
R = np.diag([1e4, 1e4, 1e-2, 1e-2, 1e-2, 1e2] * nb_contacts)

n_w = nb_contacts * 3
cov_w = np.diag([0.4 ** 2, 0.4 ** 2, 0.3 ** 2] * nb_contacts)
cov_white_noise = dt * np.diag(np.array([0.7, 0.5, 0.01, 0.8, 0.6, 0.01, 0.7, 0.5, 0.01]) ** 2)
beta_u = 0.01

# no tracking gradient for this robot (scp_solver.py:13-20): the state cost is 1/2 x' W x about the origin
# (about the nominal CoM: config/_robots.py puts the origin there)
state_cost_weights = np.diag([1e3] * 3 + [1e0] * 3 + [1e2] * 3)
control_cost_weights = np.diag([1e4, 1e4, 1e-2, 1e-2, 1e-2, 1e2] * nb_contacts)

# trust-region radius scaled to the momenta of a 90 kg robot (solo12: 2.5 kg, radius 50)
scp_params = {"trust_region_radius0": 500, "omega0": 100, "omega_max": 1.0e10, "epsilon": 1.0e-6,
              "rho0": 0.4, "rho1": 1.5, "beta_succ": 2.0, "beta_fail": 0.5, "gamma_fail": 5,
              "convergence_threshold": 1e-3, "max_iterations": 20}

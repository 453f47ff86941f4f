"""Robot constants that the reference obtains from pinocchio at import time
(config/conf_solo12_trot.py:25-28,46-47).  They cannot be verified offline (SURVEY.md
section 8d) and are recorded here as the synthetic stand-ins the benchmark is defined on.
"""
import numpy as np

from ..src.contact_plan import RobotStandIn

# solo12: pinocchio.computeTotalMass(example_robot_data 'solo12') ~ 2.5 kg; nominal stance
SOLO12_MASS = 2.5
SOLO12_COM_HEIGHT = 0.23
SOLO12_FEET = {"FL_FOOT": (0.195, 0.147, 0.0), "FR_FOOT": (0.195, -0.147, 0.0),
               "HL_FOOT": (-0.195, 0.147, 0.0), "HR_FOOT": (-0.195, -0.147, 0.0)}

# bolt biped: two point feet (placeholders: m = 1.3 kg, CoM height 0.35 m)
BOLT_MASS = 1.3
BOLT_COM_HEIGHT = 0.35
BOLT_FEET = {"FL_ANKLE": (0.0, 0.065, 0.0), "FR_ANKLE": (0.0, -0.065, 0.0)}

# talos legs: two flat feet.  The reference gives this robot NO tracking gradient (scp_solver.py:13-20), so
# its state cost is 1/2 x' W x about the world origin: the synthetic config puts the origin at the nominal
# CoM (soles 0.88 m below it), which makes that cost a regulariser about the nominal posture
TALOS_MASS = 90.0
TALOS_COM_HEIGHT = 0.0
TALOS_FEET = {"left_sole_link": (0.0, 0.085, -0.88), "right_sole_link": (0.0, -0.085, -0.88)}


def solo12():
    return RobotStandIn("solo", SOLO12_FEET, SOLO12_MASS)


def bolt():
    return RobotStandIn("bolt", BOLT_FEET, BOLT_MASS)


def talos():
    return RobotStandIn("talos", TALOS_FEET, TALOS_MASS)


def quadruped_noise(dt, z_pos_std, white):
    cov_w = np.diag([0.4 ** 2, 0.4 ** 2, z_pos_std ** 2] * 4)
    cov_eta = dt * np.diag(np.array(white) ** 2)
    return cov_w, cov_eta

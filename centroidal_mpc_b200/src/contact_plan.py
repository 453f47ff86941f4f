"""Contact plan objects for the SCP hot path, pinocchio-free.

Mirrors the live part of /root/reference/src/contact_plan.py:
  Debris                      :8-37
  create_contact_trajectory   :40-48
  create_contact_sequence     :112-264
The swing-foot interpolation / plotting helpers (:50-110, :266-328) are unused by the
SCP path and out of scope.

pinocchio is not a dependency: ``SE3`` below is the minimal stand-in for ``pin.SE3``
(attributes ``rotation`` and ``translation``), and ``create_contact_sequence`` takes the
initial foot placements from a ``RobotStandIn`` (or from pinocchio forward kinematics when
a real ``rmodel`` is handed in and pinocchio is importable).
"""
import numpy as np


class SE3:
    """rotation (3x3) + translation (3,), the two attributes the SCP path reads."""

    def __init__(self, rotation, translation):
        self.rotation = np.asarray(rotation, dtype=np.float64).reshape(3, 3)
        self.translation = np.asarray(translation, dtype=np.float64).reshape(3)

    @staticmethod
    def from_angle_axis(angle, axis3):
        """Rodrigues formula; replaces pin.AngleAxis(angle, axis).matrix()."""
        a = np.asarray(axis3, dtype=np.float64)
        n = np.linalg.norm(a)
        if n == 0.0 or angle == 0.0:
            return np.eye(3)
        a = a / n
        K = np.array([[0.0, -a[2], a[1]], [a[2], 0.0, -a[0]], [-a[1], a[0], 0.0]])
        return np.eye(3) + np.sin(angle) * K + (1.0 - np.cos(angle)) * (K @ K)


# contact name -> column of the control vector (contact_plan.py:29-37)
CONTACT_INDEX = {"RF": 0, "FR": 0, "LF": 1, "FL": 1, "HR": 2, "HL": 3}


class Debris:
    """One contact phase of one end-effector (contact_plan.py:8-37)."""

    def __init__(self, CONTACT, t_start=0.0, t_end=1.0, x=None, y=None, z=None, axis=None,
                 angle=None, ACTIVE=False):
        if ACTIVE:
            axis = np.array(axis, np.float64)
            axis /= np.linalg.norm(axis)
            self.axis = axis
            self.pose = SE3(SE3.from_angle_axis(angle, np.concatenate([axis, [0.0]])),
                            np.array([x, y, z], dtype=np.float64))
        self.t_start = t_start
        self.t_end = t_end
        self.CONTACT = CONTACT
        self.ACTIVE = ACTIVE
        self.idx = CONTACT_INDEX.get(CONTACT)


def create_contact_trajectory(conf):
    """Expand the per-phase plan to one Debris per knot and contact (contact_plan.py:40-48)."""
    plan = conf.contact_sequence
    traj = {foot.CONTACT: [] for foot in plan[0]}
    for phase in plan:
        for contact in phase:
            knots = int(round((contact.t_end - contact.t_start) / conf.dt))
            traj[contact.CONTACT].extend([contact] * knots)
    return traj


class RobotStandIn:
    """What create_contact_sequence needs from (rmodel, rdata, q0): a name that selects the
    gait family ('solo', 'talos', 'bolt') and the initial foot placements by frame name."""

    def __init__(self, name, foot_positions, mass=None):
        self.name = name
        self.foot_positions = {k: np.asarray(v, dtype=np.float64) for k, v in foot_positions.items()}
        self.mass = mass

    def createData(self):
        return None


# step phase -> feet in swing (every other foot of the robot is in stance)
_SWING = {
    "doubleSupport": (),
    "rflhStep": ("FR", "HL"), "lfrhStep": ("FL", "HR"),        # trot
    "rfrhStep": ("FR", "HR"), "lflhStep": ("FL", "HL"),        # pace (quadruped)
    "rflfStep": ("FR", "FL"), "rhlhStep": ("HR", "HL"),        # bound
    "rfStep": ("FR",), "lfStep": ("FL",),                      # biped pace
}
_CYCLE = {
    ("TROT", 4): ("rflhStep", "lfrhStep"),
    ("PACE", 4): ("rfrhStep", "lflhStep"),
    ("BOUND", 4): ("rflfStep", "rhlhStep"),
    ("PACE", 2): ("rfStep", "lfStep"),
}


def _initial_feet(ee_frame_names, rmodel, rdata, q0):
    if hasattr(rmodel, "foot_positions"):
        return [np.array(rmodel.foot_positions[n], dtype=np.float64) for n in ee_frame_names]
    import pinocchio as pin  # only reached with a real pinocchio model
    pin.forwardKinematics(rmodel, rdata, q0)
    pin.updateFramePlacements(rmodel, rdata)
    return [np.array(rdata.oMf[rmodel.getFrameId(n)].translation) for n in ee_frame_names]


def create_contact_sequence(dt, gait, ee_frame_names, rmodel, rdata, q0):
    """Gait table -> (gait_templates, contact_sequence)  (contact_plan.py:112-264).

    ``ee_frame_names`` is ordered [FL, FR, (HL, HR)] as in the reference configs
    (conf_solo12_trot.py:24; contact_plan.py:150-155).  Each phase lists its Debris in the
    fixed order FR, FL, HR, HL (contact_plan.py:165-171).  After a step phase the swinging
    feet advance by ``stepLength`` along x.
    """
    n_feet = 4 if rmodel.name == "solo" else 2
    first, second = _CYCLE[(gait["type"], n_feet)]
    steps = gait["nbSteps"]
    gait_templates = []
    for step in range(steps):
        cycle = ["doubleSupport", first, "doubleSupport", second]
        if step == steps - 1:
            cycle.append("doubleSupport")
        gait_templates.append(cycle)
    feet0 = _initial_feet(ee_frame_names, rmodel, rdata, q0)
    pos = {"FL": feet0[0], "FR": feet0[1]}
    if n_feet == 4:
        pos["HL"], pos["HR"] = feet0[2], feet0[3]
    order = ["FR", "FL", "HR", "HL"][:n_feet]
    t_start = 0.0
    contact_sequence = []
    for cycle in gait_templates:
        for phase in cycle:
            swing = _SWING[phase]
            knots = gait["supportKnots"] if phase == "doubleSupport" else gait["stepKnots"]
            t_end = t_start + knots * dt
            phase_k = []
            for name in order:
                if name in swing:
                    phase_k.append(Debris(CONTACT=name, t_start=t_start, t_end=t_end, ACTIVE=False))
                else:
                    p = pos[name]
                    phase_k.append(Debris(CONTACT=name, t_start=t_start, t_end=t_end, x=p[0], y=p[1],
                                          z=p[2], axis=[-1, 0], angle=0.0, ACTIVE=True))
            for name in swing:
                pos[name][0] += gait["stepLength"]
            t_start = t_end
            contact_sequence.append(phase_k)
    return gait_templates, contact_sequence

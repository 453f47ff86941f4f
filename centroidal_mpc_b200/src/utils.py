"""Friction-pyramid helper kept under the reference's name.

Mirrors /root/reference/src/utils.py:9-16 (``construct_friction_pyramid_constraint_matrix``).
The plotting / statistics helpers of that file are out of scope.
"""
import numpy as np


def construct_friction_pyramid_constraint_matrix(model):
    """5x3 inner linearisation of the friction cone: |f_x|,|f_y| <= (mu/sqrt2) f_z, f_z >= 0.

    Only the first four rows ever reach the QP (constraints.py:180 loops range(4)); the
    fifth is kept so the shape (and hence the chance-constraint quantile, f3) matches.
    """
    k = model._linear_friction_coefficient / np.sqrt(2.0)
    rows = [(1.0, 0.0), (-1.0, 0.0), (0.0, 1.0), (0.0, -1.0)]
    return np.array([[sx, sy, -k] for sx, sy in rows] + [[0.0, 0.0, -1.0]])

"""Free / fixed DOF partition (reference: fem/boundary.py:8-13)."""
from __future__ import annotations

import numpy as np


def free_and_fixed_dofs(ndof: int, fixed_dofs):
    """Sorted-unique fixed DOFs and the ascending complement, both int64."""
    fixed = np.unique(np.asarray(fixed_dofs, dtype=int).ravel())
    keep = np.ones(int(ndof), dtype=bool)
    keep[fixed] = False
    return np.flatnonzero(keep), fixed

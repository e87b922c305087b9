"""DOF index helpers (reference: fem/geometry.py:8-18).  Pure index arithmetic."""
from __future__ import annotations

import numpy as np


def element_dofs(node_i: int, node_j: int) -> np.ndarray:
    """Global DOFs ``[2i, 2i+1, 2j, 2j+1]`` of a 2-D bar (fem/geometry.py:8-9)."""
    i, j = int(node_i), int(node_j)
    return np.array((2 * i, 2 * i + 1, 2 * j, 2 * j + 1), dtype=int)


def split_element_data(nodes, disp, node_i, node_j):
    """``(x_i0, x_j0, u_i, u_j)`` of one element (fem/geometry.py:12-18)."""
    u = np.asarray(disp, dtype=float)
    return (nodes[node_i], nodes[node_j], u[2 * node_i:2 * node_i + 2].copy(), u[2 * node_j:2 * node_j + 2].copy())

"""Synthetic meshes of the benchmark configurations."""
from __future__ import annotations

import numpy as np


def lattice_truss(nx: int, ny: int | None = None):
    """C5 lattice truss: ``nx * ny`` nodes at integer coordinates (node id ``j*nx + i``);
    members in this order: all horizontals, all verticals, one diagonal
    ``(i,j)-(i+1,j+1)`` per cell, each group row-major; the left column is fully
    fixed.  ``nx = ny = 578`` gives 334,084 nodes and 999,941 elements.

    Returns ``(nodes[nnode,2] float64, elements[nelem,2] int64, fixed_dofs int64)``.
    """
    ny = nx if ny is None else ny
    ids = np.arange(nx * ny, dtype=np.int64).reshape(ny, nx)
    xs, ys = np.meshgrid(np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64))
    nodes = np.column_stack([xs.ravel(), ys.ravel()])
    groups = [
        (ids[:, :-1], ids[:, 1:]),      # horizontals
        (ids[:-1, :], ids[1:, :]),      # verticals
        (ids[:-1, :-1], ids[1:, 1:]),   # diagonals
    ]
    elements = np.concatenate([np.column_stack([a.ravel(), b.ravel()]) for a, b in groups])
    left = ids[:, 0]
    fixed = np.sort(np.concatenate([2 * left, 2 * left + 1]))
    return nodes, elements, fixed

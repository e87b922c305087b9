"""Single-element routines (reference: fem/element.py).  Each call evaluates a
one-element plan on the GPU -- they exist for API parity and tests; the solvers use
the batched assembly kernels directly."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from ..plan import AssemblyPlan
from ._device import default_device


@dataclass(frozen=True)
class ElementState:
    ke_total: np.ndarray
    fe_int: np.ndarray
    strain: float


def _one_element(nodes, u, young, area, dim, kind, strain_measure) -> ElementState:
    dev = default_device()
    plan = AssemblyPlan(np.asarray(nodes, dtype=float), [[0, 1]], [], dim=dim, device=dev)  # ValueError on l0 == 0
    ud = torch.as_tensor(np.asarray(u, dtype=float).reshape(-1)).to(dev)
    E = torch.full((1,), float(young), dtype=torch.float64, device=dev)
    A = torch.full((1,), float(area), dtype=torch.float64, device=dev)
    ke = plan.tangent_dense(E, A, ud, kind=kind)
    fe = plan.internal_force(ud, E, A, kind=kind)
    eps = plan.element_strain(ud, strain_measure)
    return ElementState(ke_total=ke.cpu().numpy(), fe_int=fe.cpu().numpy(), strain=float(eps[0]))


def truss1d_linear_element(x_i0, x_j0, u_i, u_j, young, area) -> ElementState:
    """fem/element.py:15-42."""
    return _one_element([float(x_i0), float(x_j0)], [u_i, u_j], young, area, 1, "linear", "linear")


def truss2d_linear_element(x_i0, x_j0, u_i, u_j, young, area) -> ElementState:
    """fem/element.py:45-102."""
    return _one_element(np.stack([x_i0, x_j0]), np.concatenate([u_i, u_j]), young, area, 2, "linear", "linear")


def truss2d_element_state(x_i0, x_j0, u_i, u_j, young, area) -> ElementState:
    """fem/element.py:105-133 (Green-Lagrange, verbatim)."""
    return _one_element(np.stack([x_i0, x_j0]), np.concatenate([u_i, u_j]), young, area, 2, "gl", "gl")

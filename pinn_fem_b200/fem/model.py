"""Data model of the drop-in layer: ``Material``, ``FEMModel`` and the legacy Newton ``SolverConfig`` /
``SolverResult``.  Field names, defaults and the ValueError texts are the reference's contract
(fem/model.py:12-107); the mesh itself ends up in an ``AssemblyPlan`` on the device (fem/_device.py)."""
from __future__ import annotations

import dataclasses
from typing import Dict, List

import numpy as np

from .properties import Property, to_property

_PROPERTY_SLOTS = ("young", "area", "density")  # Material order = parameter order of the device loops


@dataclasses.dataclass
class Material:
    """Three properties, each a plain number or a ``Property`` (scalar / network); plain numbers are
    wrapped on construction so that callers can always use ``material.young.value(...)``."""

    young: Property | float
    area: Property | float
    density: Property | float = 0.0

    def __post_init__(self):
        for slot in _PROPERTY_SLOTS:
            setattr(self, slot, to_property(getattr(self, slot)))

    def properties(self):
        return tuple(getattr(self, slot) for slot in _PROPERTY_SLOTS)

    def has_trainable_params(self) -> bool:
        return any(prop.is_trainable() for prop in self.properties())

    def get_all_torch_params(self) -> list:
        return [p for prop in self.properties() for p in prop.get_torch_params()]


def _mesh_errors(model):
    """(violated, message) pairs in the order the reference checks them."""
    dim, nodes, el = model.dimension, model.nodes, model.elements
    yield dim not in (1, 2), "dimension must be 1 or 2"
    yield dim == 1 and nodes.ndim != 1, "For 1D, nodes must be 1D array of positions"
    yield dim == 2 and (nodes.ndim != 2 or nodes.shape[1] != 2), "For 2D, nodes must have shape (nnode, 2)"
    yield el.ndim != 2 or el.shape[1] != 2, "elements must have shape (nelm, 2)"
    ndof = int(nodes.shape[0]) * dim
    yield model.loads.size != ndof, f"loads size must be {ndof}, got {model.loads.size}"
    yield bool(np.any((model.fixed_dofs < 0) | (model.fixed_dofs >= ndof))), "fixed_dofs contain out-of-range indices"


@dataclasses.dataclass
class FEMModel:
    nodes: np.ndarray
    elements: np.ndarray
    material: Material
    loads: np.ndarray
    fixed_dofs: np.ndarray
    dimension: int = 2

    def __post_init__(self) -> None:
        for name, kind, flat in (("nodes", float, False), ("elements", int, False), ("loads", float, True),
                                 ("fixed_dofs", int, True)):
            arr = np.asarray(getattr(self, name), dtype=kind)
            setattr(self, name, arr.reshape(-1) if flat else arr)
        for violated, message in _mesh_errors(self):
            if violated:
                raise ValueError(message)

    nnode = property(lambda self: int(self.nodes.shape[0]))
    nelm = property(lambda self: int(self.elements.shape[0]))
    ndof = property(lambda self: self.nnode * self.dimension)


# legacy configuration / result of fem.core.solve_incremental_newton (fem/model.py:94-107)
SolverConfig = dataclasses.make_dataclass(
    "SolverConfig",
    [("n_increments", int, 10), ("max_iterations", int, 80), ("tolerance", float, 1e-6), ("min_denominator", float, 1e-12)],
    frozen=True)
SolverConfig.__module__ = __name__

SolverResult = dataclasses.make_dataclass(
    "SolverResult",
    [("displacements", np.ndarray), ("reactions", np.ndarray), ("converged", bool),
     ("history", List[Dict[str, float]], dataclasses.field(default_factory=list))])
SolverResult.__module__ = __name__

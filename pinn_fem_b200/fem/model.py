"""Data model (reference: fem/model.py): Material, FEMModel, legacy SolverConfig / SolverResult."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np

from .properties import Property, to_property


@dataclass
class Material:
    """young / area / density, each a scalar or a Property (fem/model.py:12-43)."""

    young: Property | float
    area: Property | float
    density: Property | float = 0.0

    def __post_init__(self):
        for name in ("young", "area", "density"):
            setattr(self, name, to_property(getattr(self, name)))

    def _props(self):
        return (self.young, self.area, self.density)

    def has_trainable_params(self) -> bool:
        return any(p.is_trainable() for p in self._props())

    def get_all_torch_params(self) -> list:
        out = []
        for p in self._props():
            out.extend(p.get_torch_params())
        return out


@dataclass
class FEMModel:
    nodes: np.ndarray
    elements: np.ndarray
    material: Material
    loads: np.ndarray
    fixed_dofs: np.ndarray
    dimension: int = 2

    def __post_init__(self) -> None:
        self.nodes = np.asarray(self.nodes, dtype=float)
        self.elements = np.asarray(self.elements, dtype=int)
        self.loads = np.asarray(self.loads, dtype=float).reshape(-1)
        self.fixed_dofs = np.asarray(self.fixed_dofs, dtype=int).reshape(-1)
        if self.dimension not in (1, 2):
            raise ValueError("dimension must be 1 or 2")
        if self.dimension == 1 and self.nodes.ndim != 1:
            raise ValueError("For 1D, nodes must be 1D array of positions")
        if self.dimension == 2 and (self.nodes.ndim != 2 or self.nodes.shape[1] != 2):
            raise ValueError("For 2D, nodes must have shape (nnode, 2)")
        if self.elements.ndim != 2 or self.elements.shape[1] != 2:
            raise ValueError("elements must have shape (nelm, 2)")
        if self.loads.size != self.ndof:
            raise ValueError(f"loads size must be {self.ndof}, got {self.loads.size}")
        if np.any(self.fixed_dofs < 0) or np.any(self.fixed_dofs >= self.ndof):
            raise ValueError("fixed_dofs contain out-of-range indices")

    @property
    def nnode(self) -> int:
        return int(self.nodes.shape[0])

    @property
    def nelm(self) -> int:
        return int(self.elements.shape[0])

    @property
    def ndof(self) -> int:
        return self.nnode * self.dimension


@dataclass(frozen=True)
class SolverConfig:
    """Legacy NR configuration (fem/model.py:94-99)."""

    n_increments: int = 10
    max_iterations: int = 80
    tolerance: float = 1e-6
    min_denominator: float = 1e-12


@dataclass
class SolverResult:
    displacements: np.ndarray
    reactions: np.ndarray
    converged: bool
    history: List[Dict[str, float]] = field(default_factory=list)

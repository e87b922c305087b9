"""Legacy PINN gradient-descent solver (reference: fem/nn_solver_gd.py): mean-squared
losses, no load increments, convergence on the total loss only."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from .model import FEMModel


@dataclass
class PINNGradientDescentConfig:
    max_iterations: int = 1000
    tolerance: float = 1e-6
    learning_rate_u: float = 1e-7
    learning_rate_theta: float = 1e-4
    alpha_physics: float = 1.0
    alpha_data: float = 100.0
    print_every: int = 10


@dataclass
class PINNGradientDescentResult:
    displacements: np.ndarray
    nn_parameters: Dict[str, np.ndarray]
    converged: bool
    history: List[Dict[str, float]]


def solve_pinn_gradient_descent(model: FEMModel, f_ext: np.ndarray, measured_disp: Optional[np.ndarray] = None,
                                measured_dofs: Optional[List[int]] = None,
                                config: Optional[PINNGradientDescentConfig] = None) -> PINNGradientDescentResult:
    from .solver import _run_gd

    config = config or PINNGradientDescentConfig()
    if not model.material.has_trainable_params():
        raise ValueError("Model must have trainable NN parameters (use NNProperty)")
    if not model.material.get_all_torch_params():
        raise ValueError("No trainable parameters found in model.material")
    run = _run_gd(model, np.asarray(f_ext, dtype=float), measured_disp, measured_dofs, None,
                  max_iterations=config.max_iterations, tolerance=config.tolerance,
                  learning_rate_u=config.learning_rate_u, learning_rate_theta=config.learning_rate_theta,
                  alpha_physics=config.alpha_physics, alpha_data=config.alpha_data, load_factor=1.0, legacy_loss=True)
    keys = ("iteration", "loss_total", "loss_physics", "loss_data", "u_norm")
    history = [{k: h[k] for k in keys} for h in run["history"]]
    shape = (-1, 1) if model.dimension == 1 else (model.nnode, model.dimension)
    return PINNGradientDescentResult(displacements=run["u"].reshape(shape), nn_parameters=run["nn_parameters"] or {},
                                     converged=run["converged"], history=history)

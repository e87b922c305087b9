"""Drop-in mirror of the reference's ``fem`` package (FEM/python/fem/__init__.py): same
names, argument meaning and error behaviour; every numerical path runs on the GPU."""
from .core import solve_incremental_newton
from .model import FEMModel, Material, SolverConfig, SolverResult
from .nn_assembly import assemble_system_torch, compute_residual_and_jacobian
from .nn_solver import PINNSolverConfig, PINNSolverResult, solve_pinn_newton_raphson
from .nn_solver_gd import PINNGradientDescentConfig, PINNGradientDescentResult, solve_pinn_gradient_descent
from .properties import NNProperty, Property, ScalarProperty, to_property

__all__ = [
    "FEMModel", "Material", "SolverConfig", "SolverResult", "solve_incremental_newton", "Property",
    "ScalarProperty", "NNProperty", "to_property", "solve_pinn_newton_raphson", "PINNSolverConfig",
    "PINNSolverResult", "solve_pinn_gradient_descent", "PINNGradientDescentConfig", "PINNGradientDescentResult",
    "assemble_system_torch", "compute_residual_and_jacobian",
]

"""pinn_fem_b200 -- B200-native implementation of PINN-FEM's data-parallel hot path.

Element internal-force / tangent assembly, the residual that feeds
Newton-Raphson and gradient descent, the material MLPs and the Gauss-Newton
normal equations run as hand-written fp64 CUDA kernels (sm_100a) behind the C
ABI in ``include/pinnfem.h``; this package is the thin Python/PyTorch host side
that mirrors the reference's own entry points (``pinn_fem_b200.fem`` has the
same names as the reference's ``fem`` package).
"""
from ._lib import ELEM_GREEN_LAGRANGE, ELEM_LINEAR, PinnFemError, load as load_library
from .plan import AssemblyPlan

__all__ = ["AssemblyPlan", "PinnFemError", "load_library", "ELEM_LINEAR", "ELEM_GREEN_LAGRANGE"]

"""Dense assembly with NumPy in/out (reference: fem/assembly.py:16-75), computed on the GPU."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from ._device import get_plan, material_fields, to_dev
from .model import FEMModel


def assemble_system_device(model: FEMModel, disp, kind: str = "linear"):
    """Device-resident variant used by the solver loops: ``(plan, E, A, K dense, f_int, max|strain|)``."""
    plan = get_plan(model)
    u = to_dev(np.asarray(disp, dtype=float).reshape(-1), plan.device)
    E, A = material_fields(model, plan, load_factor=None)
    out = plan.residual(u, E, A, kind=kind, max_strain=True)
    K = plan.tangent_dense(E, A, u, kind=kind)
    return plan, E, A, K, out["f_int"], out["max_strain"]


def assemble_system(model: FEMModel, disp: np.ndarray) -> Tuple[np.ndarray, np.ndarray, float]:
    """``(K[ndof,ndof], f_int[ndof], max_abs_strain)`` exactly as the reference returns them."""
    _, _, _, K, f_int, eps = assemble_system_device(model, disp)
    return K.cpu().numpy(), f_int.cpu().numpy(), float(eps[0])

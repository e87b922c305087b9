"""Host-side glue shared by the solver mirrors: plan cache, material fields, theta packing.

Everything numerical below is a call into the CUDA library; there is no CPU path."""
from __future__ import annotations

import hashlib
from typing import List, Optional, Tuple

import numpy as np
import torch

from .. import ops
from ..plan import AssemblyPlan
from .properties import NNProperty, Property, ScalarProperty

PROPS = ("young", "area", "density")


def default_device() -> torch.device:
    if not torch.cuda.is_available():
        from .._lib import PF_ERR_NO_DEVICE, PinnFemError

        raise PinnFemError(PF_ERR_NO_DEVICE, "pinn_fem_b200 needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _nodes_np(model) -> np.ndarray:
    nodes = model.nodes
    if isinstance(nodes, torch.Tensor):  # solve_full_nr in the reference mutates nodes into tensors (A.16)
        nodes = nodes.detach().cpu().numpy()
    return np.asarray(nodes, dtype=np.float64)


def get_plan(model, device: Optional[torch.device] = None) -> AssemblyPlan:
    """AssemblyPlan of the model's mesh, cached on the model object (rebuilt if the mesh changes)."""
    device = device or default_device()
    nodes = np.ascontiguousarray(_nodes_np(model))
    el = np.ascontiguousarray(np.asarray(model.elements, dtype=np.int64))
    fixed = np.ascontiguousarray(np.asarray(model.fixed_dofs, dtype=np.int64))
    h = hashlib.blake2b(digest_size=16)
    for a in (nodes, el, fixed):
        h.update(a.tobytes())
    key = (h.hexdigest(), int(model.dimension), str(device))
    cache = model.__dict__.setdefault("_pf_plan_cache", {})
    if key not in cache:
        cache.clear()
        cache[key] = AssemblyPlan(nodes, el, fixed, dim=int(model.dimension), device=device)
    return cache[key]


def to_dev(x, device) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.detach().to(device=device, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to(device)


def nn_slots(model) -> List[Tuple[str, Optional[NNProperty]]]:
    """(name, NNProperty or None) for young, area, density -- Material.get_all_torch_params order."""
    return [(n, getattr(model.material, n) if isinstance(getattr(model.material, n), NNProperty) else None)
            for n in PROPS]


def scalar_or_scale(prop: Property) -> float:
    return prop.scale if isinstance(prop, NNProperty) else float(prop.value())


def pack_theta(model, device) -> torch.Tensor:
    parts = [p.flat_theta(device) for _, p in nn_slots(model) if p is not None]
    return torch.cat(parts) if parts else torch.empty(0, dtype=torch.float64, device=device)


def unpack_theta(model, theta: torch.Tensor) -> None:
    """Write a packed theta vector back into the nn.Modules (the solvers mutate the caller's model)."""
    off = 0
    for _, p in nn_slots(model):
        if p is not None:
            n = p.spec.n_params
            p.set_flat_theta(theta[off:off + n])
            off += n


def material_fields(model, plan: AssemblyPlan, load_factor: Optional[float]):
    """E and A per element on the device.

    ``load_factor`` given: the torch-path inputs ``{x, y, load_factor}`` whose sorted keys give
    ``[load_factor, x, y]`` (fem/nn_assembly.py:201-208).  ``None``: the NumPy-path call
    ``value(x_center)`` with zero padding up to ``input_dim`` (fem/assembly.py:59-61)."""
    out = []
    for name in ("young", "area"):
        prop = getattr(model.material, name)
        if isinstance(prop, NNProperty):
            if load_factor is not None:
                out.append(ops.mlp_forward(prop.spec, prop.flat_theta(plan.device), plan=plan,
                                           load_factor=load_factor, scale=prop.scale,
                                           enforce_positive=prop.enforce_positive))
            else:
                cen = plan.geometry("centroid")
                if cen.shape[1] > prop.input_dim:
                    raise ValueError("centroid has more coordinates than the network has inputs")
                X = np.zeros((plan.nelem, prop.input_dim))
                X[:, :cen.shape[1]] = cen
                out.append(prop.evaluate(X, plan.device))
        else:
            out.append(torch.full((plan.nelem,), float(prop.value()), dtype=torch.float64, device=plan.device))
    return out[0], out[1]

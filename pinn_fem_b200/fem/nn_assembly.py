"""Differentiable assembly (reference: fem/nn_assembly.py:105-261).

``assemble_system_torch`` returns dense ``K`` and ``f_int`` as torch tensors that are
differentiable with respect to ``disp`` and the material networks' parameters.  The
forward pass is the CUDA assembly + MLP kernels; the backward pass is the closed-form
VJP (``K g`` mat-vec, per-element material VJP, MLP backward) -- no autograd tape per
element.  Tensors are float64 on the CUDA device (the reference hard-codes float32 CPU,
SURVEY.md D2)."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from .. import ops
from ._device import default_device, get_plan, nn_slots, to_dev
from .model import FEMModel
from .properties import NNProperty


class _AssembleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, load_factor, model, plan, *params):
        dev = plan.device
        u = disp.detach().to(device=dev, dtype=torch.float64).contiguous()
        fields, thetas, specs = [], [], []
        off = 0
        for name in ("young", "area"):
            prop = getattr(model.material, name)
            if isinstance(prop, NNProperty):
                n = len(list(prop.net.parameters()))
                theta = torch.cat([p.detach().reshape(-1).to(device=dev, dtype=torch.float64)
                                   for p in params[off:off + n]]).contiguous()
                off += n
                fields.append(ops.mlp_forward(prop.spec, theta, plan=plan, load_factor=load_factor,
                                              scale=prop.scale, enforce_positive=prop.enforce_positive))
                thetas.append(theta)
                specs.append(prop)
            else:
                fields.append(torch.full((plan.nelem,), float(prop.value()), dtype=torch.float64, device=dev))
                thetas.append(None)
                specs.append(None)
        E, A = fields
        f_int = plan.internal_force(u, E, A)
        K = plan.tangent_dense(E, A)
        ctx.plan, ctx.load_factor, ctx.specs, ctx.thetas = plan, load_factor, specs, thetas
        ctx.save_for_backward(u, E, A)
        ctx.disp_meta = (disp.device, disp.dtype)
        ctx.param_meta = [(p.shape, p.device, p.dtype) for p in params]
        return K, f_int

    @staticmethod
    def backward(ctx, gK, gf):
        plan = ctx.plan
        u, E, A = ctx.saved_tensors
        dev = plan.device
        g_u = None
        gE = torch.zeros_like(E)
        gA = torch.zeros_like(A)
        if gf is not None:
            g = gf.to(device=dev, dtype=torch.float64).contiguous()
            g_u = plan.tangent_matvec(g, E, A)  # K is symmetric: dL/du = K^T g
            dE, dA = plan.material_vjp(u, E, A, g)
            gE += dE
            gA += dA
        if gK is not None and bool((gK != 0).any()):
            # dK/d(EA)_e = pattern_e / l0: contract the 2dim x 2dim block of gK of every element
            dofs = torch.as_tensor(plan.elem_dofs.copy(), device=dev)
            blk = gK.to(device=dev, dtype=torch.float64)[dofs[:, :, None], dofs[:, None, :]]
            d = plan.dim
            l0 = torch.as_tensor(plan.geometry("l0"), device=dev)
            if d == 2:
                c = torch.as_tensor(plan.geometry("cos"), device=dev)
                s = torch.as_tensor(plan.geometry("sin"), device=dev)
                dirv = torch.stack([c, s, -c, -s], dim=1)
            else:
                dirv = torch.tensor([1.0, -1.0], dtype=torch.float64, device=dev).expand(plan.nelem, 2)
            contr = torch.einsum("eab,ea,eb->e", blk, dirv, dirv) / l0
            gE += A * contr
            gA += E * contr
        grads = []
        for k, (prop, theta) in enumerate(zip(ctx.specs, ctx.thetas)):
            if prop is None:
                continue
            g_theta = ops.mlp_backward(prop.spec, theta, (gE if k == 0 else gA).contiguous(), plan=plan,
                                       load_factor=ctx.load_factor, scale=prop.scale,
                                       enforce_positive=prop.enforce_positive)
            grads.append(g_theta)
        flat = torch.cat(grads) if grads else None
        out, off = [], 0
        for shape, pdev, pdt in ctx.param_meta:
            n = int(np.prod(shape)) if len(shape) else 1
            out.append(flat[off:off + n].reshape(shape).to(device=pdev, dtype=pdt))
            off += n
        if g_u is not None:
            g_u = g_u.to(device=ctx.disp_meta[0], dtype=ctx.disp_meta[1])
        return (g_u, None, None, None, *out)


def assemble_system_torch(model: FEMModel, disp: torch.Tensor, load_factor: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(k_global[ndof,ndof], f_int[ndof])``, differentiable in ``disp`` and the NN parameters.
    Only young and area enter (density never does, fem/nn_assembly.py:207-208)."""
    plan = get_plan(model, default_device())
    params = []
    for name in ("young", "area"):
        prop = getattr(model.material, name)
        if isinstance(prop, NNProperty):
            params.extend(prop.net.parameters())
    if not isinstance(disp, torch.Tensor):
        disp = torch.as_tensor(np.asarray(disp, dtype=float))
    return _AssembleFn.apply(disp, float(load_factor), model, plan, *params)


def compute_residual_and_jacobian(model: FEMModel, disp: torch.Tensor, f_ext: torch.Tensor, free_dofs: np.ndarray):
    """``(R[free], K[free, free])`` with ``R = f_int - f_ext`` (fem/nn_assembly.py:234-261)."""
    k_global, f_int = assemble_system_torch(model, disp)
    f_ext = f_ext.to(device=f_int.device, dtype=f_int.dtype)
    idx = torch.as_tensor(np.asarray(free_dofs), device=f_int.device)
    return (f_int - f_ext)[idx], k_global[idx][:, idx]


def truss1d_linear_element_torch(x_i0, x_j0, u_i, u_j, young, area):
    """fem/nn_assembly.py:18-47 on a one-element plan: ``(ke[2,2], fe[2])``."""
    from ..plan import AssemblyPlan

    dev = default_device()
    plan = AssemblyPlan(np.array([float(x_i0), float(x_j0)]), [[0, 1]], [], dim=1, device=dev)
    u = torch.stack([torch.as_tensor(u_i).reshape(()), torch.as_tensor(u_j).reshape(())])
    return _single(plan, u, young, area)


def truss2d_linear_element_torch(x_i0, x_j0, u_i, u_j, young, area):
    """fem/nn_assembly.py:50-102 on a one-element plan: ``(ke[4,4], fe[4])``, differentiable in u."""
    from ..plan import AssemblyPlan

    dev = default_device()
    plan = AssemblyPlan(np.stack([np.asarray(x_i0, float), np.asarray(x_j0, float)]), [[0, 1]], [], dim=2, device=dev)
    u = torch.cat([torch.as_tensor(u_i).reshape(-1), torch.as_tensor(u_j).reshape(-1)])
    return _single(plan, u, young, area)


class _SingleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, plan, ea):
        dev = plan.device
        ud = u.detach().to(device=dev, dtype=torch.float64).contiguous()
        E = torch.full((1,), float(ea), dtype=torch.float64, device=dev)
        A = torch.ones(1, dtype=torch.float64, device=dev)
        ctx.plan, ctx.meta = plan, (u.device, u.dtype)
        ctx.save_for_backward(E, A)
        return plan.tangent_dense(E, A), plan.internal_force(ud, E, A)

    @staticmethod
    def backward(ctx, gK, gf):
        E, A = ctx.saved_tensors
        g = gf.to(device=ctx.plan.device, dtype=torch.float64).contiguous()
        return ctx.plan.tangent_matvec(g, E, A).to(device=ctx.meta[0], dtype=ctx.meta[1]), None, None


def _single(plan, u, young, area):
    ea = float(torch.as_tensor(young).reshape(-1)[0]) * float(torch.as_tensor(area).reshape(-1)[0])
    return _SingleFn.apply(u, plan, ea)

"""Gauss-Newton / Levenberg-Marquardt solver for the coupled (u, theta) inverse problem
(reference: fem/nn_solver.py).  The Jacobian is assembled in closed form on the device
(``pf_gn_jacobian``), J^T J runs on the fp64 tensor cores, the damped system is solved by
LU; the update / line-search logic -- including the reference's quirks -- is kept."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from .. import ops
from ._device import get_plan, nn_slots, pack_theta, to_dev, unpack_theta
from .boundary import free_and_fixed_dofs
from .model import FEMModel
from .properties import NNProperty


@dataclass
class PINNSolverConfig:
    max_iterations: int = 50
    tolerance: float = 1e-6
    alpha_physics: float = 1.0
    alpha_data: float = 1.0
    min_denominator: float = 1e-12
    max_step_u: float = 1e-3
    max_step_theta: float = 0.1
    line_search: bool = True


@dataclass
class PINNSolverResult:
    displacements: np.ndarray
    nn_parameters: Dict[str, np.ndarray]
    converged: bool
    history: List[Dict[str, float]]


class _State:
    """Device-side view of the model's networks for one Gauss-Newton run."""

    def __init__(self, model, load_factor):
        self.model, self.plan = model, get_plan(model)
        self.dev = self.plan.device
        self.lam = float(load_factor)
        self.slots = nn_slots(model)
        self.theta = pack_theta(model, self.dev)
        self.offsets, off = {}, 0
        for name, p in self.slots:
            if p is not None:
                self.offsets[name] = (off, p.spec.n_params)
                off += p.spec.n_params
        self.n_theta = off

    def _theta_of(self, name):
        o, n = self.offsets[name]
        return self.theta[o:o + n].contiguous()

    def field(self, name):
        prop = getattr(self.model.material, name)
        if isinstance(prop, NNProperty):
            return ops.mlp_forward(prop.spec, self._theta_of(name), plan=self.plan, load_factor=self.lam,
                                   scale=prop.scale, enforce_positive=prop.enforce_positive)
        return torch.full((self.plan.nelem,), float(prop.value()), dtype=torch.float64, device=self.dev)

    def param_jacobian(self, name):
        prop = getattr(self.model.material, name)
        if not isinstance(prop, NNProperty):
            return None
        return ops.mlp_param_jacobian(prop.spec, self._theta_of(name), plan=self.plan, load_factor=self.lam,
                                      scale=prop.scale, enforce_positive=prop.enforce_positive)


def compute_jacobian_blocks(model: FEMModel, u: torch.Tensor, f_ext_torch: torch.Tensor, theta_list, free_dofs,
                            measured_dofs=None):
    """``(j_uu, j_utheta, r_physics, j_data_u, None)`` as in fem/nn_solver.py:50-135 (float64 CUDA
    tensors).  ``theta_list`` is accepted for signature parity; the parameters are read from the model."""
    st = _State(model, 1.0)
    plan = st.plan
    ud = to_dev(u, st.dev)
    E, A = st.field("young"), st.field("area")
    jE, jA = st.param_jacobian("young"), st.param_jacobian("area")
    n_rest = st.n_theta - (0 if jE is None else jE.shape[1]) - (0 if jA is None else jA.shape[1])
    md = None if measured_dofs is None else np.asarray(measured_dofs, dtype=np.int64)
    J = ops.gn_jacobian(plan, ud, E, A, jE, jA, n_rest=n_rest, alpha_physics=1.0, alpha_data=1.0, meas_dofs=md)
    nf = plan.nfree
    free = torch.as_tensor(np.asarray(free_dofs), device=st.dev)
    r = plan.internal_force(ud, E, A) - to_dev(f_ext_torch, st.dev)
    j_data_u = J[nf:, :nf].clone() if md is not None else None
    return J[:nf, :nf].clone(), J[:nf, nf:].clone(), r[free], j_data_u, None


def _gauss_newton(model, f_ext, measured_disp, measured_dofs, config, load_factor=1.0):
    st = _State(model, load_factor)
    plan, dev = st.plan, st.dev
    f_ext_d = to_dev(f_ext, dev)
    free = torch.as_tensor(plan.free_dofs.copy(), device=dev)
    fixed = torch.as_tensor(plan.fixed_dofs.copy(), device=dev)
    nf = plan.nfree
    has_meas = measured_disp is not None and measured_dofs is not None
    md_np = np.asarray(measured_dofs, dtype=np.int64) if has_meas else None
    md = torch.as_tensor(md_np, device=dev) if has_meas else None
    mv = to_dev(measured_disp, dev) if has_meas else None
    u = torch.zeros(plan.ndof, dtype=torch.float64, device=dev)
    ap, ad = float(config.alpha_physics), float(config.alpha_data)
    history, converged = [], False

    def residual_parts(uu):
        E, A = st.field("young"), st.field("area")
        r_p = (plan.internal_force(uu, E, A) - f_ext_d)[free]
        r_d = (mv - uu[md]) if has_meas else torch.zeros(0, dtype=torch.float64, device=dev)
        return E, A, r_p, r_d

    for it in range(int(config.max_iterations)):
        E, A, r_p, r_d = residual_parts(u)
        jE, jA = st.param_jacobian("young"), st.param_jacobian("area")
        n_rest = st.n_theta - (0 if jE is None else jE.shape[1]) - (0 if jA is None else jA.shape[1])
        if has_meas:
            J = ops.gn_jacobian(plan, u, E, A, jE, jA, n_rest=n_rest, alpha_physics=ap, alpha_data=ad, meas_dofs=md_np)
            R = torch.cat([ap * r_p, ad * r_d])
        else:  # the reference leaves J unweighted here and scales only R (nn_solver.py:240-243)
            J = ops.gn_jacobian(plan, u, E, A, jE, jA, n_rest=n_rest, alpha_physics=1.0, alpha_data=1.0)
            R = ap * r_p
        rp_n, rd_n, rt_n = float(torch.linalg.vector_norm(r_p)), float(torch.linalg.vector_norm(r_d)), float(torch.linalg.vector_norm(R))
        jtj, jtr, _ = ops.gn_normal_equations(J, R.contiguous(), 1e-6)
        try:
            dx = ops.solve_spd(jtj, (-jtr).contiguous())  # J^T J + d I is SPD: Cholesky (the reference's LU agrees to rounding)
        except RuntimeError as exc:
            print(f"Solver failed at iteration {it + 1}: {exc}")
            break
        du_f, dth = dx[:nf], dx[nf:]
        step = 1.0
        if config.line_search:
            for _ in range(15):
                u_t = u.clone()
                u_t[free] += step * du_f
                u_t[fixed] = 0.0
                backup = st.theta.clone()
                st.theta += step * dth
                _, _, rp_t, rd_t = residual_parts(u_t)
                rt = torch.cat([ap * rp_t, ad * rd_t]) if has_meas else ap * rp_t
                if float(torch.linalg.vector_norm(rt)) < rt_n * (1.0 - 1e-4 * step):
                    break  # accepted: theta keeps the trial update (and is advanced again below, nn_solver.py:366-371)
                st.theta = backup
                step *= 0.7
                if step < 1e-10:
                    step = 0.0
                    break
            if 0.0 < step < 1e-8:
                step = 1e-6
        if step > 0:
            u[free] += step * du_f
            u[fixed] = 0.0
            st.theta += step * dth
        rel = rt_n / max(float(torch.linalg.vector_norm(u[free])), config.min_denominator)
        history.append({"iteration": float(it + 1), "r_physics": rp_n, "r_data": rd_n, "r_total": rt_n,
                        "relative_error": rel, "step_size": float(step)})
        print(f"{it + 1:5d} | {rp_n:12.3e} | {rd_n:12.3e} | {rt_n:12.3e} | {step:6.3f}")
        if rel < config.tolerance and step > 0:
            converged = True
            break
        if step == 0.0:
            break
    unpack_theta(model, st.theta)
    return {"u": u.cpu().numpy(), "converged": converged, "history": history}


def solve_pinn_newton_raphson(model: FEMModel, f_ext: np.ndarray, measured_disp: Optional[np.ndarray] = None,
                              measured_dofs: Optional[List[int]] = None,
                              config: Optional[PINNSolverConfig] = None) -> PINNSolverResult:
    config = config or PINNSolverConfig()
    if not model.material.has_trainable_params():
        raise ValueError("Model must have trainable NN parameters (use NNProperty)")
    theta_list = model.material.get_all_torch_params()
    if not theta_list:
        raise ValueError("No trainable parameters found in model.material")
    run = _gauss_newton(model, np.asarray(f_ext, dtype=float), measured_disp, measured_dofs, config, 1.0)
    shape = (-1, 1) if model.dimension == 1 else (model.nnode, model.dimension)
    params = {f"param_{i}": p.detach().cpu().numpy() for i, p in enumerate(model.material.get_all_torch_params())}
    return PINNSolverResult(displacements=run["u"].reshape(shape), nn_parameters=params, converged=run["converged"],
                            history=run["history"])

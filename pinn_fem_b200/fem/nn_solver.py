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


def _line_search_steps(n=15):
    """The step sizes the reference's backtracking loop visits (nn_solver.py:293-355): 1, then `step *= 0.7` per
    rejected trial, as the same fp64 products; entry n is the step it is left with when all n trials fail."""
    steps, step = [], 1.0
    for _ in range(n + 1):
        steps.append(step)
        step *= 0.7
    return steps


def _gauss_newton(model, f_ext, measured_disp, measured_dofs, config, load_factor=1.0, lm_path="auto"):
    """fem/nn_solver.py:223-371 with the iteration resident on the device: residual, closed-form Jacobian, the
    Levenberg-Marquardt step (dual m x m system when there are fewer residuals than unknowns), and the backtracking
    line search evaluated for ALL 15 trial steps at once (one batched network forward + one batched residual; the
    first step that passes the sufficient-decrease test is selected on the device, which is what the sequential loop
    would have stopped at).  The host reads one small record per iteration (norms, step, solver status)."""
    import os

    if lm_path == "auto":  # PF_GN_LM_PATH=primal|dual forces one formulation (tests compare the two)
        lm_path = os.environ.get("PF_GN_LM_PATH", "auto")
    st = _State(model, load_factor)
    plan, dev = st.plan, st.dev
    f_ext_d = to_dev(f_ext, dev)
    free = torch.as_tensor(plan.free_dofs.copy(), device=dev)
    nf = plan.nfree
    has_meas = measured_disp is not None and measured_dofs is not None
    md_np = np.asarray(measured_dofs, dtype=np.int64) if has_meas else None
    md = torch.as_tensor(md_np, device=dev) if has_meas else None
    mv = to_dev(measured_disp, dev) if has_meas else None
    u = torch.zeros(plan.ndof, dtype=torch.float64, device=dev)
    ap, ad = float(config.alpha_physics), float(config.alpha_data)
    history, converged = [], False
    n_trials = 15
    steps = torch.tensor(_line_search_steps(n_trials), dtype=torch.float64, device=dev)
    trial_steps = steps[:n_trials]
    nets = {name: getattr(model.material, name) for name in ("young", "area")}
    free_mask = torch.zeros(plan.ndof, dtype=torch.float64, device=dev)
    free_mask[free] = 1.0

    def trial_fields(theta_rows):
        """E, A [nelem, n_trials] for the rows of trial parameter vectors."""
        out = []
        for name in ("young", "area"):
            prop = nets[name]
            if isinstance(prop, NNProperty):
                o, n = st.offsets[name]
                if ops.mlp_acts_len(prop.spec, plan.nelem) > 0:
                    vals, _ = ops.mlp_forward_batched(prop.spec, theta_rows[:, o:], plan=plan, load_factor=st.lam,
                                                      scale=prop.scale, enforce_positive=prop.enforce_positive,
                                                      save_acts=False)
                else:  # shapes outside the batched kernels: one launch per trial
                    vals = torch.stack([ops.mlp_forward(prop.spec, theta_rows[k, o:o + n].contiguous(), plan=plan,
                                                        load_factor=st.lam, scale=prop.scale,
                                                        enforce_positive=prop.enforce_positive)
                                        for k in range(theta_rows.shape[0])], dim=1).contiguous()
            else:
                vals = torch.full((plan.nelem, theta_rows.shape[0]), float(prop.value()), dtype=torch.float64, device=dev)
            out.append(vals)
        return out

    for it in range(int(config.max_iterations)):
        E, A = st.field("young"), st.field("area")
        r_full = (plan.internal_force(u, E, A) - f_ext_d)
        r_p = r_full[free]
        r_d = (mv - u[md]) if has_meas else torch.zeros(0, dtype=torch.float64, device=dev)
        jE, jA = st.param_jacobian("young"), st.param_jacobian("area")
        n_rest = st.n_theta - (0 if jE is None else jE.shape[1]) - (0 if jA is None else jA.shape[1])
        if has_meas:
            J = ops.gn_jacobian(plan, u, E, A, jE, jA, n_rest=n_rest, alpha_physics=ap, alpha_data=ad, meas_dofs=md_np)
            R = torch.cat([ap * r_p, ad * r_d])
        else:  # the reference leaves J unweighted here and scales only R (nn_solver.py:240-243)
            J = ops.gn_jacobian(plan, u, E, A, jE, jA, n_rest=n_rest, alpha_physics=1.0, alpha_data=1.0)
            R = ap * r_p
        rp_n, rd_n, rt_n = torch.linalg.vector_norm(r_p), torch.linalg.vector_norm(r_d), torch.linalg.vector_norm(R)
        # (J^T J + d I) dx = -J^T R is SPD by construction (the reference's LU agrees to rounding)
        dx, _, info = ops.gn_lm_step(J, R.contiguous(), 1e-6, path=lm_path, check_info=False)
        du = torch.zeros(plan.ndof, dtype=torch.float64, device=dev)
        du[free] = dx[:nf]
        dth = dx[nf:]
        if config.line_search:
            # all trial states at once: u_k = u + s_k du on the free DOFs, theta_k = theta + s_k dtheta
            u_t = (u[:, None] + du[:, None] * trial_steps[None, :]) * free_mask[:, None]
            th_t = (st.theta[None, :] + trial_steps[:, None] * dth[None, :]).contiguous()
            E_t, A_t = trial_fields(th_t)
            sq_p = 2.0 * plan.residual(u_t.contiguous(), E_t, A_t, f_ext_d, 1.0, f_int=False, r=False, half_sq=True)["half_sq"]
            if has_meas:
                rd_t = mv[:, None] - u_t[md]
                rt_t = torch.sqrt(ap * ap * sq_p + ad * ad * (rd_t * rd_t).sum(0))
            else:
                rt_t = torch.sqrt(ap * ap * sq_p)
            ok = rt_t < rt_n * (1.0 - 1e-4 * trial_steps)
            any_ok = ok.any()
            k = torch.argmax(ok.to(torch.int32))  # first accepted trial
            step = torch.where(any_ok, trial_steps[k], steps[n_trials])
            # an accepted trial leaves its theta in place and the update below advances it again
            # (nn_solver.py:309-313 + :366-371); a failed search restores theta and takes the last step once
            th_mult = torch.where(any_ok, 2.0 * step, step)
        else:
            step = torch.ones((), dtype=torch.float64, device=dev)
            th_mult = step
        u = (u + step * du) * free_mask
        st.theta = st.theta + th_mult * dth
        un = torch.linalg.vector_norm(u[free])
        rec = torch.stack([rp_n, rd_n, rt_n, step, un, info[0].to(torch.float64)]).cpu()  # the iteration's one sync
        rp_v, rd_v, rt_v, step_v, un_v, info_v = (float(x) for x in rec)
        if info_v != 0.0 or not np.isfinite(step_v):
            # the state was advanced with an undefined step: undo is pointless, the reference stops here too
            print(f"Solver failed at iteration {it + 1}: Matrix is not positive definite (pivot {int(info_v)})")
            break
        rel = rt_v / max(un_v, config.min_denominator)
        history.append({"iteration": float(it + 1), "r_physics": rp_v, "r_data": rd_v, "r_total": rt_v,
                        "relative_error": rel, "step_size": step_v})
        print(f"{it + 1:5d} | {rp_v:12.3e} | {rd_v:12.3e} | {rt_v:12.3e} | {step_v:6.3f}")
        if rel < config.tolerance and step_v > 0:
            converged = True
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

"""Unified solver drivers (reference: fem/solver.py): gradient descent, Newton-Raphson,
hybrid, full-NR and the incremental ``solve`` driver.  Control flow, budgets, history
schema and convergence rules follow the reference; the numerical work of every
iteration runs on the GPU (device-resident GD loop, assembly kernels, LU/CG)."""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from .. import ops
from ._device import get_plan, material_fields, nn_slots, pack_theta, scalar_or_scale, to_dev, unpack_theta
from .boundary import free_and_fixed_dofs
from .core import newton_iterations
from .model import FEMModel


@dataclass
class SolverConfig:
    """Unified configuration (fem/solver.py:35-62)."""

    max_iterations: int = 1000
    tolerance: float = 1e-6
    print_every: int = 10
    n_increments: int = 10
    load_factor_initial: float = 0.0
    load_factor_final: float = 1.0
    min_denominator: float = 1e-10
    learning_rate_u: float = 1e-7
    learning_rate_theta: float = 1e-4
    alpha_physics: float = 1.0
    alpha_data: float = 100.0
    method: str = "auto"
    preconditioning: bool = False


@dataclass
class SolverResult:
    displacements: np.ndarray
    reactions: np.ndarray
    converged: bool
    history: List[Dict[str, float]] = field(default_factory=list)
    nn_parameters: Optional[Dict[str, np.ndarray]] = None


def _shape(model):
    return (-1, 1) if model.dimension == 1 else (model.nnode, model.dimension)


def _nn_parameter_dict(model):
    params = model.material.get_all_torch_params()
    if not params:
        return None
    return {f"param_{i}": p.detach().cpu().numpy() for i, p in enumerate(params)}


# ---------------------------------------------------------------------------------------
# gradient descent
# ---------------------------------------------------------------------------------------


# running totals over all phases / increments of a process (the result object only carries the last
# increment's history, like the reference); read by bench.py to report iterations per second
COUNTERS = {"gd_iterations": 0, "gd_calls": 0}


def _run_gd(model, loads, measured_disp, measured_dofs, u_initial, *, max_iterations, tolerance, learning_rate_u,
            learning_rate_theta, alpha_physics, alpha_data, load_factor, legacy_loss=False):
    """One ``solve_gd`` inner loop on the device; mutates the model's networks like the reference."""
    plan = get_plan(model)
    dev = plan.device
    slots = nn_slots(model)
    nets = [p.spec if p is not None else None for _, p in slots]
    scales = [scalar_or_scale(getattr(model.material, n)) for n, _ in slots]
    theta = pack_theta(model, dev).unsqueeze(0).contiguous()
    if u_initial is None:
        u = torch.zeros((1, plan.ndof), dtype=torch.float64, device=dev)
    else:
        u0 = u_initial.detach().cpu().numpy() if isinstance(u_initial, torch.Tensor) else np.asarray(u_initial)
        # the reference round-trips warm starts through float32 (solver.py:147-149, :1110)
        u = to_dev(np.asarray(u0, dtype=np.float32).astype(np.float64).reshape(1, -1), dev)
    has_meas = measured_disp is not None and measured_dofs is not None
    md = np.asarray(measured_dofs, dtype=np.int64) if has_meas else None
    mv = np.asarray(measured_disp, dtype=np.float64) if has_meas else None
    f_ext = to_dev(loads, dev)
    kw = dict(max_iterations=int(max_iterations), tolerance=float(tolerance), learning_rate_u=learning_rate_u,
              learning_rate_theta=learning_rate_theta, alpha_physics=alpha_physics, alpha_data=alpha_data,
              load_factor=float(load_factor), legacy_loss=legacy_loss)
    # one device-resident loop: a single CTA per problem when the mesh fits in shared memory, else the
    # multi-kernel loop of pf_gd_large.cu -- chosen inside pf_gd_solve
    res = ops.gd_solve(plan, nets, scales, theta if theta.numel() else None, u, f_ext, md, mv, **kw)
    n = int(res.n_iters[0])
    COUNTERS["gd_iterations"] += n
    COUNTERS["gd_calls"] += 1
    H = res.history[0, :n].cpu().numpy()
    converged = bool(res.converged[0])
    reactions = res.reactions[0]
    theta_out, u_out = res.theta[0], res.u[0]
    if theta_out.numel():
        unpack_theta(model, theta_out)
    has_nn = any(p is not None for _, p in slots)
    keys = ["iteration", "loss_total", "loss_physics", "loss_data", "u_norm", "residual_norm"] + (["theta_norm"] if has_nn else [])
    history = [{k: float(row[c]) for c, k in enumerate(keys)} for row in H]
    return {"u": u_out.cpu().numpy(), "reactions": reactions.cpu().numpy(), "converged": converged,
            "history": history, "nn_parameters": _nn_parameter_dict(model)}


def _print_history(history, print_every, has_nn):
    head = f"{'Iter':>6} | {'Loss Total':>12} | {'Loss Physics':>12} | {'||R||':>12} | {'Loss Data':>12} | {'||u||':>10}"
    print(head + (f" | {'NN Params':>10}" if has_nn else ""))
    for h in history:
        it = int(h["iteration"])
        if it == 1 or it % max(int(print_every), 1) == 0:
            msg = (f"{it:6d} | {h['loss_total']:12.3e} | {h['loss_physics']:12.3e} | {h['residual_norm']:12.3e} | "
                   f"{h['loss_data']:12.3e} | {h['u_norm']:10.3e}")
            print(msg + (f" | {h['theta_norm']:10.3e}" if has_nn else ""))


def solve_gd(model: FEMModel, config: Optional[SolverConfig] = None, measured_disp: Optional[np.ndarray] = None,
             measured_dofs: Optional[List[int]] = None, target_load_factor: float = 1.0,
             u_initial: Optional[torch.Tensor] = None, skip_preconditioning: bool = False) -> SolverResult:
    """Gradient-descent solver, optionally with the relaxed preconditioning phase (fem/solver.py:83-400)."""
    config = config or SolverConfig()
    if config.preconditioning and not skip_preconditioning:
        pre = copy.deepcopy(config)
        pre.max_iterations = min(300, config.max_iterations // 3)
        pre.tolerance = max(1e-4, config.tolerance * 10)
        pre.preconditioning = False
        try:
            first = solve_gd(model, pre, measured_disp, measured_dofs, target_load_factor, u_initial, True)
            if first.converged and first.history[-1].get("residual_norm", 1.0) < config.tolerance:
                return first
            main = copy.deepcopy(config)
            main.max_iterations = config.max_iterations - pre.max_iterations
            main.preconditioning = False
            warm = torch.tensor(first.displacements.flatten(), dtype=torch.float32)
            second = solve_gd(model, main, measured_disp, measured_dofs, target_load_factor, warm, True)
            shift = first.history[-1].get("iteration", 0) if first.history else 0
            second.history = list(first.history) + [dict(h, iteration=h.get("iteration", 0) + shift)
                                                    for h in second.history]
            return second
        except Exception as exc:  # same recovery as the reference (solver.py:197-198)
            print(f"  Preconditioning failed: {exc}, proceeding with standard GD")
    run = _run_gd(model, model.loads, measured_disp, measured_dofs, u_initial, max_iterations=config.max_iterations,
                  tolerance=config.tolerance, learning_rate_u=config.learning_rate_u,
                  learning_rate_theta=config.learning_rate_theta, alpha_physics=config.alpha_physics,
                  alpha_data=config.alpha_data, load_factor=target_load_factor)
    has_nn = model.material.has_trainable_params()
    _print_history(run["history"], config.print_every, has_nn)
    if run["converged"]:
        print(f"[CONVERGED] in {len(run['history'])} iterations")
    else:
        print(f"[WARNING] Did not converge in {config.max_iterations} iterations.")
    return SolverResult(displacements=run["u"].reshape(_shape(model)), reactions=run["reactions"].reshape(_shape(model)),
                        converged=run["converged"], history=run["history"], nn_parameters=run["nn_parameters"])


# ---------------------------------------------------------------------------------------
# Newton-Raphson
# ---------------------------------------------------------------------------------------


def solve_nr(model: FEMModel, config: Optional[SolverConfig] = None, target_load_factor: float = 1.0,
             u_initial: Optional[torch.Tensor] = None) -> SolverResult:
    """One load level of classical NR; always restarts from u = 0 (fem/solver.py:408-512)."""
    config = config or SolverConfig()
    if model.material.has_trainable_params():
        raise ValueError("Newton-Raphson solver with NN materials not fully supported yet. "
                         "Use solve_gd() for problems with NN parameters.")
    plan = get_plan(model)
    E, A = material_fields(model, plan, load_factor=None)
    loads = to_dev(model.loads, plan.device)
    u = torch.zeros(plan.ndof, dtype=torch.float64, device=plan.device)
    lam = float(target_load_factor)
    ok, res, eps, n_it = newton_iterations(plan, E, A, u, lam * loads, config.max_iterations, config.tolerance,
                                           config.min_denominator)
    history = [{"load_factor": lam, "iterations": float(n_it), "residual": float(res), "max_strain": float(eps),
                "converged": float(1.0 if ok else 0.0)}]
    reactions = plan.tangent_matvec(u, E, A) - lam * loads
    reactions[torch.as_tensor(plan.free_dofs.copy(), device=plan.device)] = 0.0
    return SolverResult(displacements=u.cpu().numpy().reshape(_shape(model)),
                        reactions=reactions.cpu().numpy().reshape(_shape(model)), converged=ok, history=history)


def solve_hybrid(model: FEMModel, config: Optional[SolverConfig] = None, measured_disp: Optional[np.ndarray] = None,
                 measured_dofs: Optional[List[int]] = None, target_load_factor: float = 1.0,
                 u_initial: Optional[torch.Tensor] = None) -> SolverResult:
    """GD warm-up, then NR for scalar materials or tight-tolerance GD for NN materials
    (fem/solver.py:520-692: with networks the second phase is GD again)."""
    config = config or SolverConfig()
    first, first_cfg = None, None
    if config.preconditioning:
        first_cfg = copy.deepcopy(config)
        first_cfg.max_iterations = min(300, config.max_iterations // 3)
        first_cfg.tolerance = max(1e-4, config.tolerance * 10)
        try:
            first = solve_gd(model, first_cfg, measured_disp, measured_dofs, target_load_factor, u_initial, True)
            if first.converged and first.history[-1].get("residual_norm", 1.0) < config.tolerance:
                return first
        except Exception as exc:
            print(f"  GD Phase failed: {exc}, proceeding with cold NR")
            first = None
    warm = torch.tensor(first.displacements.flatten(), dtype=torch.float32) if first else u_initial
    if model.material.has_trainable_params():
        final_cfg = copy.deepcopy(config)
        final_cfg.max_iterations = config.max_iterations - (first_cfg.max_iterations if first else 0)
        final = solve_gd(model, final_cfg, measured_disp, measured_dofs, target_load_factor, warm, True)
        if first:
            shift = first.history[-1].get("iteration", 0) if first.history else 0
            final.history = list(first.history) + [dict(h, iteration=h.get("iteration", 0) + shift)
                                                   for h in final.history]
        return final
    nr = solve_nr(model, config, target_load_factor, warm)
    if first:
        total = (first.history[-1].get("iteration", 0) if first.history else 0) + nr.history[-1].get("iterations", 1)
        last = dict(nr.history[-1], iteration=total)
        nr.history = list(first.history) + [last]
    return nr


def solve_full_nr(model: FEMModel, config: Optional[SolverConfig] = None, measured_disp: Optional[np.ndarray] = None,
                  measured_dofs: Optional[List[int]] = None, target_load_factor: float = 1.0) -> SolverResult:
    """Without networks: classical NR, like the reference (fem/solver.py:787-790).

    With networks the reference's implementation never builds the coupled system and crashes
    (SURVEY.md D4); here the coupled (u, theta) problem is solved with the Gauss-Newton /
    Levenberg-Marquardt iteration of fem/nn_solver.py at the requested load factor."""
    config = config or SolverConfig()
    if not model.material.has_trainable_params():
        return solve_nr(model, config, target_load_factor)
    from .nn_solver import PINNSolverConfig, _gauss_newton

    cfg = PINNSolverConfig(max_iterations=config.max_iterations, tolerance=config.tolerance,
                           alpha_physics=config.alpha_physics, alpha_data=config.alpha_data)
    run = _gauss_newton(model, np.asarray(model.loads, dtype=float) * target_load_factor, measured_disp,
                        measured_dofs, cfg, load_factor=target_load_factor)
    plan = get_plan(model)
    E, A = material_fields(model, plan, load_factor=target_load_factor)
    u = to_dev(run["u"], plan.device)
    reac = plan.internal_force(u, E, A) - target_load_factor * to_dev(model.loads, plan.device)
    reac[torch.as_tensor(plan.free_dofs.copy(), device=plan.device)] = 0.0
    return SolverResult(displacements=run["u"].reshape(_shape(model)), reactions=reac.cpu().numpy().reshape(_shape(model)),
                        converged=run["converged"], history=run["history"], nn_parameters=_nn_parameter_dict(model))


# ---------------------------------------------------------------------------------------
# incremental driver
# ---------------------------------------------------------------------------------------


def solve(model: FEMModel, config: Optional[SolverConfig] = None, measured_disp: Optional[np.ndarray] = None,
          measured_dofs: Optional[List[int]] = None) -> SolverResult:
    """Load-increment loop over the selected method; returns the LAST increment's result and
    stops at the first increment that does not converge (fem/solver.py:1045-1167)."""
    config = config or SolverConfig()
    if config.method != "auto":
        method = config.method.lower()
    else:
        has_meas = measured_disp is not None and measured_dofs is not None
        method = "nr" if (not model.material.has_trainable_params() and not has_meas) else "gd"
    result, u_cur = None, None
    for iinc in range(1, config.n_increments + 1):
        lam = config.load_factor_initial + (iinc / config.n_increments) * (config.load_factor_final - config.load_factor_initial)
        warm = None if u_cur is None else torch.tensor(u_cur, dtype=torch.float32, requires_grad=False)
        print(f"{iinc:>4} | {lam:>12.4f} | {'WARM_START' if warm is not None else 'COLD_START':>10}")
        if method == "gd":
            result = solve_gd(model, config, measured_disp, measured_dofs, target_load_factor=lam, u_initial=warm)
        elif method == "nr":
            result = solve_nr(model, config, target_load_factor=lam, u_initial=warm)
        elif method == "hybrid":
            result = solve_hybrid(model, config, measured_disp, measured_dofs, target_load_factor=lam, u_initial=warm)
        elif method == "full-nr":
            result = solve_full_nr(model, config, measured_disp, measured_dofs, target_load_factor=lam)
        else:
            raise ValueError(f"Unknown solver method: {method}")
        u_cur = result.displacements.flatten()
        print(f"{iinc:4d} | {lam:12.6f} | {'CONVERGED' if result.converged else 'FAILED':>10}")
        if not result.converged:
            print(f"[WARNING] Increment {iinc} did not converge, stopping incremental loading.")
            break
    return result

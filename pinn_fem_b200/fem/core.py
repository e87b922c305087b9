"""Classical incremental Newton-Raphson (reference: fem/core.py:10-79) on the device."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from ._device import get_plan, material_fields, to_dev
from .model import FEMModel, SolverConfig, SolverResult

DENSE_LIMIT = 4096  # free DOFs up to which K_ff is factorised densely; above: matrix-free CG
SPD_FROM = 121      # from here on K_ff goes through the blocked Cholesky first (below: LU in shared memory)


def newton_step(plan, E, A, u, rhs, kind="linear"):
    """Solve ``K_t(u)[free,free] du_f = rhs[free]``; returns ``du`` (zero on fixed DOFs).
    Raises the reference's RuntimeError when the tangent is singular (fem/core.py:36-37)."""
    free = torch.as_tensor(plan.free_dofs.copy(), device=plan.device)
    du = torch.zeros_like(u)
    if plan.nfree == 0:
        return du
    if plan.nfree <= DENSE_LIMIT:
        k_ff = plan.tangent_dense(E, A, u, kind=kind, free_only=True)
        rhs_f = rhs[free].contiguous()
        x = None
        if plan.nfree >= SPD_FROM:
            # K_ff of a constrained truss is symmetric positive definite: blocked Cholesky on all SMs (2.2 ms at
            # n = 1001, 11.5 ms at 4096; the pivoted LU the reference calls needs 10.9 / 161 ms).  A non-positive
            # pivot (mechanism, indefinite Green-Lagrange tangent) falls through to the LU, which decides
            # "singular" exactly like np.linalg.solve.
            try:
                x = ops.solve_spd(k_ff, rhs_f)
            except RuntimeError:
                x = None
        if x is None:
            try:
                x = ops.solve_dense(k_ff, rhs_f)
            except RuntimeError as exc:
                raise RuntimeError("Tangent stiffness became singular during solve") from exc
        du[free] = x
    else:
        if kind != "linear":
            raise NotImplementedError("CG path implements the linear element")
        x, _, resid = ops.cg_solve(plan, E, A, rhs.contiguous(), rel_tol=1e-12, max_iters=20 * plan.nfree)
        if not np.isfinite(resid):
            raise RuntimeError("Tangent stiffness became singular during solve")
        du = x
    return du


def newton_iterations(plan, E, A, u, f_ext, max_iterations, tolerance, min_denominator, kind="linear"):
    """NR iterations at one load level (fem/core.py:27-49 / fem/solver.py:456-481): step-norm test."""
    converged, res, eps, n_it = False, float("inf"), 0.0, 0
    for ite in range(int(max_iterations)):
        out = plan.residual(u, E, A, kind=kind, max_strain=True)
        eps = float(out["max_strain"][0])
        du = newton_step(plan, E, A, u, f_ext - out["f_int"], kind)
        u += du
        res = float(torch.linalg.vector_norm(du)) / max(float(torch.linalg.vector_norm(u)), min_denominator)
        n_it = ite + 1
        if res <= tolerance:
            converged = True
            break
    return converged, res, eps, n_it


def solve_incremental_newton(model: FEMModel, config: SolverConfig | None = None, kind: str = "linear") -> SolverResult:
    config = config or SolverConfig()
    plan = get_plan(model)
    E, A = material_fields(model, plan, load_factor=None)
    loads = to_dev(model.loads, plan.device)
    u = torch.zeros(plan.ndof, dtype=torch.float64, device=plan.device)
    history, ok_all = [], True
    for iinc in range(1, config.n_increments + 1):
        lam = iinc / config.n_increments
        ok, res, eps, n_it = newton_iterations(plan, E, A, u, lam * loads, config.max_iterations, config.tolerance,
                                               config.min_denominator, kind)
        history.append({"increment": float(iinc), "load_factor": float(lam), "iterations": float(n_it),
                        "residual": float(res), "max_strain": float(eps), "converged": float(1.0 if ok else 0.0)})
        ok_all = ok_all and ok
    # reactions = K u - loads, zero on free DOFs (core.py:61-63)
    reactions = plan.tangent_matvec(u, E, A, u, kind=kind) - loads
    reactions[torch.as_tensor(plan.free_dofs.copy(), device=plan.device)] = 0.0
    shape = (-1, 1) if model.dimension == 1 else (model.nnode, model.dimension)
    return SolverResult(displacements=u.cpu().numpy().reshape(shape), reactions=reactions.cpu().numpy().reshape(shape),
                        converged=ok_all, history=history)

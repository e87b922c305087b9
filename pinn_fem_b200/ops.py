"""Thin torch-facing wrappers over the C-ABI entry points that are not tied to
an :class:`AssemblyPlan` method: material networks, the device-resident
gradient-descent loop, dense solves and the Gauss-Newton normal equations.
All tensors are float64 CUDA tensors; nothing here computes on the CPU."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import check
from .plan import AssemblyPlan, _ptr, _stream_ptr


@dataclass(frozen=True)
class NetSpec:
    """SimpleNN architecture (examples/json/generic.py:118-142)."""

    input_dim: int
    hidden_layers: int
    width: int

    @property
    def n_params(self) -> int:
        n = int(_lib.load().pf_mlp_num_params(self.input_dim, self.hidden_layers, self.width))
        if n < 0:
            raise ValueError(f"unsupported network {self}")
        return n


def _dev_f64(t: torch.Tensor, name: str) -> torch.Tensor:
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise ValueError(f"{name} must be a contiguous float64 CUDA tensor")
    return t


def _mlp_args(spec: NetSpec, theta, X, plan):
    theta = _dev_f64(theta, "theta")
    if theta.numel() != spec.n_params:
        raise ValueError(f"theta has {theta.numel()} entries, network needs {spec.n_params}")
    if X is None:
        if plan is None:
            raise ValueError("either X or a plan (element centroids) is required")
        return theta, None, plan.nelem, plan._handle, plan.device
    X = _dev_f64(X, "X")
    if X.dim() != 2 or X.shape[1] != spec.input_dim:
        raise ValueError(f"X must be [n, {spec.input_dim}]")
    return theta, X, X.shape[0], (plan._handle if plan is not None else None), X.device


def mlp_forward(spec: NetSpec, theta, X=None, plan: Optional[AssemblyPlan] = None, load_factor=1.0, scale=1.0,
                enforce_positive=True) -> torch.Tensor:
    """``softplus(net(x)) * scale`` (fem/properties.py:150-156) at the rows of X, or at
    the plan's element centroids with inputs ``[load_factor, x_c(, y_c)]``."""
    theta, X, n, handle, dev = _mlp_args(spec, theta, X, plan)
    out = torch.empty(n, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().pf_mlp_forward(handle, spec.input_dim, spec.hidden_layers, spec.width, _ptr(theta), n,
                                         _ptr(X), float(load_factor), float(scale), int(enforce_positive),
                                         _ptr(out), _stream_ptr(dev)))
    return out


def mlp_backward(spec: NetSpec, theta, g_out, X=None, plan=None, load_factor=1.0, scale=1.0,
                 enforce_positive=True) -> torch.Tensor:
    """dL/dtheta (flat) given dL/dvalue per point."""
    theta, X, n, handle, dev = _mlp_args(spec, theta, X, plan)
    g_out = _dev_f64(g_out, "g_out")
    if g_out.numel() != n:
        raise ValueError("g_out must have one entry per point")
    g_theta = torch.empty(spec.n_params, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().pf_mlp_backward(handle, spec.input_dim, spec.hidden_layers, spec.width, _ptr(theta), n,
                                          _ptr(X), float(load_factor), float(scale), int(enforce_positive),
                                          _ptr(g_out), _ptr(g_theta), _stream_ptr(dev)))
    return g_theta


def mlp_acts_len(spec: NetSpec, n: int) -> int:
    """Doubles per problem of the activation record the batched forward leaves for the batched backward
    (0: the shape is not covered by the batched kernels)."""
    return int(_lib.load().pf_mlp_acts_len(spec.input_dim, spec.hidden_layers, spec.width, int(n)))


def _mlp_batched_args(spec: NetSpec, theta, X, plan):
    # rows may be slices of a wider [B, n_theta_total] array (unit stride inside a row)
    if not (isinstance(theta, torch.Tensor) and theta.is_cuda and theta.dtype == torch.float64 and theta.dim() == 2
            and theta.shape[1] >= spec.n_params and (theta.shape[1] == 1 or theta.stride(1) == 1)):
        raise ValueError(f"theta must be a float64 CUDA tensor [B, >= {spec.n_params}] with unit inner stride")
    if X is None:
        if plan is None:
            raise ValueError("either X or a plan (element centroids) is required")
        return theta, None, plan.nelem, plan._handle, plan.device
    X = _dev_f64(X, "X")
    if X.dim() != 2 or X.shape[1] != spec.input_dim:
        raise ValueError(f"X must be [n, {spec.input_dim}]")
    return theta, X, X.shape[0], (plan._handle if plan is not None else None), X.device


def mlp_forward_batched(spec: NetSpec, theta, X=None, plan=None, load_factor=1.0, scale=1.0, enforce_positive=True,
                        save_acts=True):
    """Values of B networks (rows of ``theta`` [B, n_params]) at the same points: ``out`` [n, B] (problem index last,
    the layout of the batched assembly kernels) and, with ``save_acts``, the activation record [B, acts_len] that
    :func:`mlp_backward_batched` consumes."""
    theta, X, n, handle, dev = _mlp_batched_args(spec, theta, X, plan)
    B = theta.shape[0]
    out = torch.empty((n, B), dtype=torch.float64, device=dev)
    acts = None
    if save_acts:
        alen = mlp_acts_len(spec, n)
        if alen <= 0:
            raise ValueError(f"network {spec} is not covered by the batched kernels")
        acts = torch.empty((B, alen), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().pf_mlp_forward_batched(handle, spec.input_dim, spec.hidden_layers, spec.width, _ptr(theta),
                                                 theta.stride(0), B, n, _ptr(X), float(load_factor), float(scale),
                                                 int(enforce_positive), _ptr(out), B, _ptr(acts), _stream_ptr(dev)))
    return out, acts


def mlp_backward_batched(spec: NetSpec, theta, g_out, acts, X=None, plan=None, load_factor=1.0) -> torch.Tensor:
    """dL/dtheta [B, n_params] given dL/dvalue [n, B] and the record of :func:`mlp_forward_batched`."""
    theta, X, n, handle, dev = _mlp_batched_args(spec, theta, X, plan)
    B = theta.shape[0]
    g_out = _dev_f64(g_out, "g_out")
    acts = _dev_f64(acts, "acts")
    if tuple(g_out.shape) != (n, B):
        raise ValueError(f"g_out must be [{n}, {B}]")
    if tuple(acts.shape) != (B, mlp_acts_len(spec, n)):
        raise ValueError("acts does not match the network / point count")
    g_theta = torch.empty((B, spec.n_params), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().pf_mlp_backward_batched(handle, spec.input_dim, spec.hidden_layers, spec.width, _ptr(theta),
                                                  theta.stride(0), B, n, _ptr(X), float(load_factor), _ptr(g_out), B,
                                                  _ptr(acts), _ptr(g_theta), spec.n_params, _stream_ptr(dev)))
    return g_theta


def mlp_param_jacobian(spec: NetSpec, theta, X=None, plan=None, load_factor=1.0, scale=1.0,
                       enforce_positive=True) -> torch.Tensor:
    """``jac[p, :] = d value_p / d theta`` for every point."""
    theta, X, n, handle, dev = _mlp_args(spec, theta, X, plan)
    jac = torch.empty((n, spec.n_params), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().pf_mlp_param_jacobian(handle, spec.input_dim, spec.hidden_layers, spec.width, _ptr(theta),
                                                n, _ptr(X), float(load_factor), float(scale),
                                                int(enforce_positive), _ptr(jac), _stream_ptr(dev)))
    return jac


@dataclass
class GDResult:
    u: torch.Tensor            # [nprob, ndof]
    theta: torch.Tensor        # [nprob, n_theta]
    reactions: torch.Tensor    # [nprob, ndof]
    history: torch.Tensor      # [nprob, max_iterations, 7] (rows beyond n_iters are undefined)
    n_iters: torch.Tensor      # int32 [nprob]
    converged: torch.Tensor    # int32 [nprob]


def gd_solve(plan: AssemblyPlan, nets: Sequence[Optional[NetSpec]], scales: Sequence[float], theta, u, f_ext,
             meas_dofs=None, meas_vals=None, *, max_iterations=1000, tolerance=1e-6, learning_rate_u=1e-7,
             learning_rate_theta=1e-4, alpha_physics=1.0, alpha_data=100.0, load_factor=1.0,
             record_history=True, legacy_loss=False) -> GDResult:
    """Run ``solve_gd``'s inner loop (fem/solver.py:252-355) on the device for ``nprob``
    independent problems sharing ``plan``.  ``nets[k]`` is the architecture of
    property k (young, area, density) or None for a scalar of value ``scales[k]``;
    ``theta`` is ``[nprob, sum(n_params of enabled nets)]`` and is updated in place,
    as is ``u`` ``[nprob, ndof]``."""
    plan._need_device()
    dev = plan.device
    u = _dev_f64(u, "u")
    if u.dim() == 1:
        u = u.unsqueeze(0)
    nprob = u.shape[0]
    if u.shape[1] != plan.ndof:
        raise ValueError(f"u must be [nprob, {plan.ndof}]")
    cfg = _lib.GDConfig()
    cfg.max_iterations = int(max_iterations)
    cfg.kind = _lib.ELEM_LINEAR
    cfg.tolerance = float(tolerance)
    cfg.learning_rate_u = float(learning_rate_u)
    cfg.learning_rate_theta = float(learning_rate_theta)
    cfg.alpha_physics = float(alpha_physics)
    cfg.alpha_data = float(alpha_data)
    cfg.load_factor = float(load_factor)
    cfg.loss_mode = 1 if legacy_loss else 0
    ntheta = 0
    for k in range(3):
        spec = nets[k] if k < len(nets) else None
        cfg.net_enabled[k] = 1 if spec is not None else 0
        cfg.net_scale[k] = float(scales[k]) if k < len(scales) else 0.0
        if spec is not None:
            cfg.net_input_dim[k], cfg.net_hidden_layers[k], cfg.net_width[k] = spec.input_dim, spec.hidden_layers, spec.width
            ntheta += spec.n_params
    if ntheta:
        theta = _dev_f64(theta, "theta")
        if theta.dim() == 1:
            theta = theta.unsqueeze(0)
        if tuple(theta.shape) != (nprob, ntheta):
            raise ValueError(f"theta must be [{nprob}, {ntheta}], got {tuple(theta.shape)}")
    else:
        theta = torch.empty((nprob, 0), dtype=torch.float64, device=dev)
    f_ext = _dev_f64(f_ext, "f_ext")
    n_meas = 0
    md = mv = None
    if meas_dofs is not None and meas_vals is not None and len(meas_dofs) > 0:
        md_host = torch.as_tensor(meas_dofs).cpu().to(torch.int64).reshape(-1)
        if int(md_host.min()) < 0 or int(md_host.max()) >= plan.ndof:  # the kernels index shared / global memory with them
            raise IndexError(f"measured DOF out of range [0, {plan.ndof}): min {int(md_host.min())}, max {int(md_host.max())}")
        md = md_host.to(torch.int32).to(dev).contiguous()
        mv = torch.as_tensor(meas_vals, dtype=torch.float64, device=dev)
        if mv.dim() == 1:
            mv = mv.unsqueeze(0).expand(nprob, -1)
        mv = mv.contiguous()
        n_meas = md.numel()
        if tuple(mv.shape) != (nprob, n_meas):
            raise ValueError("meas_vals must be [n_measured] or [nprob, n_measured]")
    cfg.n_measured = n_meas
    history = (torch.empty((nprob, max(int(max_iterations), 1), _lib.GD_HISTORY_COLS), dtype=torch.float64, device=dev)
               if record_history else None)
    n_iters = torch.zeros(nprob, dtype=torch.int32, device=dev)
    converged = torch.zeros(nprob, dtype=torch.int32, device=dev)
    reactions = torch.empty((nprob, plan.ndof), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().pf_gd_solve(plan._handle, C.byref(cfg), nprob, _ptr(theta) if ntheta else None, _ptr(u),
                                      _ptr(f_ext), _ptr(md), _ptr(mv), _ptr(history), _ptr(n_iters), _ptr(converged),
                                      _ptr(reactions), _stream_ptr(dev)))
    return GDResult(u=u, theta=theta, reactions=reactions, history=history, n_iters=n_iters, converged=converged)


def gd_solve_host(plan: AssemblyPlan, nets, scales, theta_host, u_host, f_ext, meas_dofs=None, meas_vals_host=None, *,
                  out_theta=None, out_u=None, out_history=None, **kw) -> GDResult:
    """:func:`gd_solve` for problems held in (pinned) host memory -- the end-to-end form of the batched inverse
    problem: ``theta_host`` [nprob, n_theta], ``u_host`` [nprob, ndof] and the targets go to the device, the solve
    runs there, and theta, u and the history rows come back into ``out_theta`` / ``out_u`` / ``out_history`` (host
    tensors, allocated pinned when not given).  The returned GDResult holds the host tensors (reactions, n_iters
    and converged stay on the device).  Synchronous."""
    plan._need_device()
    dev = plan.device
    th = theta_host.to(dev, non_blocking=True) if theta_host is not None else None
    uu = u_host.to(dev, non_blocking=True)
    mv = meas_vals_host.to(dev, non_blocking=True) if isinstance(meas_vals_host, torch.Tensor) else meas_vals_host
    res = gd_solve(plan, nets, scales, th, uu, f_ext, meas_dofs, mv, **kw)
    pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    out_u = pin(res.u) if out_u is None else out_u
    out_u.copy_(res.u, non_blocking=True)
    if res.theta.numel():
        out_theta = pin(res.theta) if out_theta is None else out_theta
        out_theta.copy_(res.theta, non_blocking=True)
    if res.history is not None:
        out_history = pin(res.history) if out_history is None else out_history
        out_history.copy_(res.history, non_blocking=True)
    torch.cuda.synchronize(dev)
    return GDResult(u=out_u, theta=out_theta if res.theta.numel() else res.theta, reactions=res.reactions,
                    history=out_history, n_iters=res.n_iters, converged=res.converged)


def solve_dense(A: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Solve ``A x = b`` by LU with partial pivoting (np.linalg.solve / torch.linalg.solve
    in the reference).  ``A`` is ``[n, n]`` or ``[batch, n, n]``; both inputs are left
    untouched.  Raises ``RuntimeError("Singular matrix")`` on an exactly zero pivot."""
    A = _dev_f64(A, "A")
    b = _dev_f64(b, "b")
    batched = A.dim() == 3
    A3 = (A if batched else A.unsqueeze(0)).clone()
    x = (b if batched else b.unsqueeze(0)).clone()
    nb, n = A3.shape[0], A3.shape[1]
    if A3.shape[2] != n or tuple(x.shape) != (nb, n):
        raise ValueError("solve_dense: shape mismatch")
    info = torch.zeros(nb, dtype=torch.int32, device=A.device)
    with torch.cuda.device(A.device):
        check(_lib.load().pf_solve_dense(nb, n, _ptr(A3), _ptr(x), _ptr(info), _stream_ptr(A.device)))
    if bool((info != 0).any()):
        raise RuntimeError("Singular matrix")
    return x if batched else x[0]


def solve_spd(A: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Solve ``A x = b`` for a symmetric positive definite ``A`` ``[n, n]`` by blocked Cholesky (only the lower
    triangle is read); inputs are left untouched.  Raises ``RuntimeError`` when a pivot is not positive."""
    A = _dev_f64(A, "A")
    b = _dev_f64(b, "b")
    n = A.shape[0]
    if A.dim() != 2 or A.shape[1] != n or tuple(b.shape) != (n,):
        raise ValueError("solve_spd: shape mismatch")
    L, x = A.clone(), b.clone()
    info = torch.zeros(1, dtype=torch.int32, device=A.device)
    with torch.cuda.device(A.device):
        check(_lib.load().pf_solve_spd(n, _ptr(L), _ptr(x), _ptr(info), _stream_ptr(A.device)))
    if int(info[0]) != 0:
        raise RuntimeError(f"Matrix is not positive definite (pivot {int(info[0])})")
    return x


def cg_solve(plan: AssemblyPlan, E, A, rhs, u=None, kind="linear", rel_tol=1e-10, max_iters=10000):
    """Jacobi-preconditioned CG on the free DOFs of ``K(E, A) x = rhs`` (matrix-free,
    batched ``[ndof, B]``).  Returns ``(x, iterations, worst relative residual)``."""
    from .plan import KINDS

    plan._need_device()
    B = plan._chk(rhs, plan.ndof, "rhs")
    mb = plan._materials(E, A, B)
    lib = _lib.load()
    work = torch.empty(int(lib.pf_cg_work_len(plan._handle, B)), dtype=torch.float64, device=plan.device)
    x = torch.empty_like(rhs)
    iters, resid = C.c_int32(0), C.c_double(0.0)
    with torch.cuda.device(plan.device):
        check(lib.pf_cg_solve(plan._handle, KINDS[kind], B, _ptr(u), _ptr(E), _ptr(A), mb, _ptr(rhs), _ptr(x),
                              float(rel_tol), int(max_iters), _ptr(work), work.numel(), C.byref(iters),
                              C.byref(resid), _stream_ptr(plan.device)))
    return x, int(iters.value), float(resid.value)


def gn_jacobian(plan: AssemblyPlan, u, E, A, jacE=None, jacA=None, n_rest=0, alpha_physics=1.0, alpha_data=1.0,
                meas_dofs=None) -> torch.Tensor:
    """Stacked Gauss-Newton Jacobian ``[(nfree + n_meas), nfree + nE + nA + n_rest]``
    (fem/nn_solver.py:50-135, :223-239) assembled in closed form on the device."""
    plan._need_device()
    dev = plan.device
    vals = plan.tangent_bsr(E, A)
    nE = 0 if jacE is None else jacE.shape[1]
    nA = 0 if jacA is None else jacA.shape[1]
    md = None if meas_dofs is None else torch.as_tensor(meas_dofs, device=dev).to(torch.int32).contiguous()
    n_meas = 0 if md is None else md.numel()
    J = torch.empty((plan.nfree + n_meas, plan.nfree + nE + nA + n_rest), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().pf_gn_jacobian(plan._handle, _lib.ELEM_LINEAR, _ptr(u), _ptr(E), _ptr(A), _ptr(vals),
                                         _ptr(jacE), nE, _ptr(jacA), nA, int(n_rest), float(alpha_physics),
                                         float(alpha_data), _ptr(md), n_meas, _ptr(J), _stream_ptr(dev)))
    return J


def gn_lm_step(J: torch.Tensor, R: torch.Tensor, damping_factor: float = 1e-6, path: str = "auto", check_info=True):
    """One Levenberg-Marquardt step ``dx = -(J^T J + d I)^-1 J^T R`` (fem/nn_solver.py:266-277).  ``path``:
    ``"auto"`` (dual m x m system when J has fewer rows than columns), ``"primal"`` (n x n), ``"dual"``.
    Returns ``(dx, damping, info)``; with ``check_info`` a non-positive pivot raises ``RuntimeError`` (one sync),
    otherwise ``info`` (device int32) is left for the caller to test."""
    J = _dev_f64(J, "J")
    R = _dev_f64(R, "R")
    m, n = J.shape
    if tuple(R.shape) != (m,):
        raise ValueError("gn_lm_step: R must have one entry per row of J")
    dx = torch.empty(n, dtype=torch.float64, device=J.device)
    damping = torch.empty(1, dtype=torch.float64, device=J.device)
    info = torch.zeros(1, dtype=torch.int32, device=J.device)
    with torch.cuda.device(J.device):
        check(_lib.load().pf_gn_lm_step(m, n, _ptr(J), _ptr(R), float(damping_factor),
                                        {"auto": 0, "primal": 1, "dual": 2}[path], _ptr(dx), _ptr(damping), _ptr(info),
                                        _stream_ptr(J.device)))
    if check_info and int(info[0]) != 0:
        raise RuntimeError(f"Matrix is not positive definite (pivot {int(info[0])})")
    return dx, damping, info


def gn_normal_equations(J: torch.Tensor, R: torch.Tensor, damping_factor: float = 1e-6):
    """``JtJ + d I`` with ``d = damping_factor * trace(JtJ) / n``, ``Jtr`` and ``d``
    (fem/nn_solver.py:268-274); J^T J runs on the fp64 tensor cores (DMMA)."""
    J = _dev_f64(J, "J")
    R = _dev_f64(R, "R")
    m, n = J.shape
    jtj = torch.empty((n, n), dtype=torch.float64, device=J.device)
    jtr = torch.empty(n, dtype=torch.float64, device=J.device)
    damping = torch.empty(1, dtype=torch.float64, device=J.device)
    with torch.cuda.device(J.device):
        check(_lib.load().pf_gn_normal_equations(m, n, _ptr(J), _ptr(R), float(damping_factor), _ptr(jtj), _ptr(jtr),
                                                 _ptr(damping), _stream_ptr(J.device)))
    return jtj, jtr, damping

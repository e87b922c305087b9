"""GPU parity: batched inverse problems on one large mesh (BASELINE config 5 in small: many independent PINN-GD
problems sharing a mesh that does not fit one CTA).  pf_gd_solve advances them in groups on the batched kernels
(patch-staged residual / K r / material VJP, fragment MLP kernels with per-problem parameters): every problem must
follow the oracle's solve_gd (fem/solver.py:252-355) iteration by iteration and agree with the one-problem loop."""
import numpy as np
import pytest
import torch

from oracle import pinnfem_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64).cuda()


HIST = ((1, "loss_total"), (2, "loss_physics"), (3, "loss_data"), (4, "u_norm"), (5, "residual_norm"), (6, "theta_norm"))


def _problem_set(nx, B, seed, with_meas=True):
    nodes, el, fixed = O.lattice_truss(nx)
    rng = np.random.default_rng(seed)
    loads = np.zeros(2 * len(nodes))
    loads[-2], loads[-1] = 0.05, -0.02
    specs = (O.NetSpec(3, 2, 20), O.NetSpec(3, 2, 15))
    theta = rng.normal(scale=0.3, size=(B, specs[0].n_params + specs[1].n_params))
    u0 = rng.uniform(-1e-3, 1e-3, size=(B, 2 * len(nodes)))
    u0[:, fixed] = 0.0
    nn = len(nodes)
    md = np.array([2 * (nn - 1), 2 * (nn - 1) + 1, 2 * (nn // 2) + 1, 2 * (nn // 2) + 1, 2 * (nn // 3)]) if with_meas else None
    mv = rng.normal(scale=0.01, size=(B, 5)) if with_meas else None
    return nodes, el, fixed, loads, specs, theta, u0, md, mv


@pytest.mark.parametrize("nx,B,n_it", [(24, 64, 30), (40, 70, 16), (40, 5, 20)])
def test_batched_large_mesh_gd_vs_oracle(nx, B, n_it):
    from pinn_fem_b200 import AssemblyPlan, ops

    nodes, el, fixed, loads, specs, theta, u0, md, mv = _problem_set(nx, B, 100 + nx + B)
    plan = AssemblyPlan(nodes, el, fixed, device="cuda")
    nets, scales = [ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15), None], [2.0, 1.5, 1.0]
    kw = dict(max_iterations=n_it, tolerance=1e-14, learning_rate_u=1e-4, learning_rate_theta=1e-3, alpha_data=10.0,
              load_factor=0.9)
    res = ops.gd_solve(plan, nets, scales, dev(theta).clone(), dev(u0).clone(), dev(loads), md, dev(mv), **kw)
    assert res.n_iters.cpu().tolist() == [n_it] * B and not res.converged.any()
    mesh = O.Mesh(nodes, el, loads, fixed)
    nE = specs[0].n_params
    for p in sorted({0, 1, B // 2, B - 1}):
        mat = O.MaterialNets((specs[0], theta[p, :nE].copy(), 2.0), (specs[1], theta[p, nE:].copy(), 1.5), 1.0)
        u_ref, reac_ref, ok, hist = O.solve_gd(mesh, mat, n_it, 1e-14, lr_u=1e-4, lr_theta=1e-3, alpha_d=10.0, meas_dofs=md,
                                               meas_vals=mv[p], lam=0.9, u_initial=u0[p])
        assert rel(res.u[p], u_ref) < 1e-8 and rel(res.reactions[p], reac_ref) < 1e-8
        assert rel(res.theta[p], np.concatenate([mat.young[1], mat.area[1]])) < 1e-8
        for col, key in HIST:
            assert rel(res.history[p, :n_it, col], np.array([h[key] for h in hist])) < 1e-8, (p, key)
    # the one-problem device loop on the same inputs (different kernels, different summation orders)
    for p in (1, B - 1):
        one = ops.gd_solve(plan, nets, scales, dev(theta[p:p + 1]).clone(), dev(u0[p:p + 1]).clone(), dev(loads), md,
                           dev(mv[p:p + 1]), **kw)
        assert rel(res.u[p], one.u[0].cpu().numpy()) < 1e-9 and rel(res.theta[p], one.theta[0].cpu().numpy()) < 1e-9
        assert rel(res.history[p, :n_it, 1], one.history[0, :n_it, 1].cpu().numpy()) < 1e-9
    # reproducible bit for bit
    res2 = ops.gd_solve(plan, nets, scales, dev(theta).clone(), dev(u0).clone(), dev(loads), md, dev(mv), **kw)
    assert torch.equal(res.u, res2.u) and torch.equal(res.theta, res2.theta) and torch.equal(res.history, res2.history)


def test_batched_problems_converge_independently():
    """Per-problem convergence flags: a problem that passes the test stops changing while its group runs on."""
    from pinn_fem_b200 import AssemblyPlan, ops

    B, n_it = 20, 30
    nodes, el, fixed, loads, specs, theta, u0, md, mv = _problem_set(24, B, 7)
    # problems 0..9 get a loose target they meet at iteration 12 (first check is at iteration index 11)
    plan = AssemblyPlan(nodes, el, fixed, device="cuda")
    nets, scales = [ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15), None], [2.0, 1.5, 1.0]
    mesh = O.Mesh(nodes, el, loads, fixed)
    nE = specs[0].n_params
    # pick the tolerance between the loss values of the two halves at iteration 12
    losses = []
    for p in range(B):
        mat = O.MaterialNets((specs[0], theta[p, :nE].copy(), 2.0), (specs[1], theta[p, nE:].copy(), 1.5), 1.0)
        _, _, _, hist = O.solve_gd(mesh, mat, 12, 0.0, lr_u=1e-4, lr_theta=1e-3, alpha_d=10.0, meas_dofs=md, meas_vals=mv[p],
                                   lam=0.9, u_initial=u0[p])
        losses.append(hist[-1]["loss_total"])
    tol = float(np.median(losses))
    res = ops.gd_solve(plan, nets, scales, dev(theta).clone(), dev(u0).clone(), dev(loads), md, dev(mv), max_iterations=n_it,
                       tolerance=tol, learning_rate_u=1e-4, learning_rate_theta=1e-3, alpha_data=10.0, load_factor=0.9)
    for p in range(B):
        mat = O.MaterialNets((specs[0], theta[p, :nE].copy(), 2.0), (specs[1], theta[p, nE:].copy(), 1.5), 1.0)
        u_ref, _, ok, hist = O.solve_gd(mesh, mat, n_it, tol, lr_u=1e-4, lr_theta=1e-3, alpha_d=10.0, meas_dofs=md,
                                        meas_vals=mv[p], lam=0.9, u_initial=u0[p])
        assert int(res.n_iters[p]) == len(hist) and bool(res.converged[p]) == bool(ok), p
        assert rel(res.u[p], u_ref) < 1e-8
    assert 0 < int(res.converged.sum()) < B  # the case exercises both outcomes


def test_batched_scalar_and_mixed_materials():
    """No network at all (only u is optimised) and one network + one scalar property, batched."""
    from pinn_fem_b200 import AssemblyPlan, ops

    B, n_it = 18, 6
    nodes, el, fixed, loads, specs, theta, u0, md, mv = _problem_set(24, B, 11)
    plan = AssemblyPlan(nodes, el, fixed, device="cuda")
    mesh = O.Mesh(nodes, el, loads, fixed)
    nE = specs[0].n_params
    kw = dict(max_iterations=n_it, tolerance=1e-14, learning_rate_u=1e-4, learning_rate_theta=1e-3, alpha_data=10.0,
              load_factor=0.9)
    res = ops.gd_solve(plan, [ops.NetSpec(3, 2, 20), None, None], [2.0, 1.5, 1.0], dev(theta[:, :nE]).clone(), dev(u0).clone(),
                       dev(loads), md, dev(mv), **kw)
    for p in (0, B - 1):
        mat = O.MaterialNets((specs[0], theta[p, :nE].copy(), 2.0), 1.5, 1.0)
        u_ref, reac_ref, ok, hist = O.solve_gd(mesh, mat, n_it, 1e-14, lr_u=1e-4, lr_theta=1e-3, alpha_d=10.0, meas_dofs=md,
                                               meas_vals=mv[p], lam=0.9, u_initial=u0[p])
        assert rel(res.u[p], u_ref) < 1e-8 and rel(res.theta[p], mat.young[1]) < 1e-8
        assert rel(res.history[p, :n_it, 1], np.array([h["loss_total"] for h in hist])) < 1e-8
    res = ops.gd_solve(plan, [None, None, None], [2.0, 1.5, 1.0], None, dev(u0).clone(), dev(loads), md, dev(mv), **kw)
    for p in (0, B - 1):
        u_ref, reac_ref, ok, hist = O.solve_gd(mesh, O.MaterialNets(2.0, 1.5, 1.0), n_it, 1e-14, lr_u=1e-4, lr_theta=1e-3,
                                               alpha_d=10.0, meas_dofs=md, meas_vals=mv[p], lam=0.9, u_initial=u0[p])
        assert rel(res.u[p], u_ref) < 1e-8 and rel(res.reactions[p], reac_ref) < 1e-8
        assert rel(res.history[p, :n_it, 2], np.array([h["loss_physics"] for h in hist])) < 1e-8

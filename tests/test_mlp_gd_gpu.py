"""GPU parity: material-network kernels and the device-resident GD loop against the oracle."""
import json

import numpy as np
import pytest
import torch

from oracle import pinnfem_oracle as O

pytestmark = pytest.mark.gpu

EX_NODES = np.array([[0.0, 0], [1, 0], [2, 0], [3, 0]])
EX_EL = np.array([[0, 1], [1, 2], [2, 3]])
EX_FIXED = np.array([0, 1, 3, 5, 7])
SPECS = {"young": O.NetSpec(3, 2, 20), "area": O.NetSpec(3, 2, 15), "density": O.NetSpec(3, 2, 10)}


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64).cuda()


@pytest.mark.parametrize("shape", [(3, 2, 20), (3, 2, 15), (2, 1, 7), (3, 4, 33), (5, 3, 64)])
@pytest.mark.parametrize("n", [1, 3, 130, 1000])
def test_mlp_forward_backward_jacobian(shape, n):
    from pinn_fem_b200 import ops

    spec_o = O.NetSpec(*shape)
    spec = ops.NetSpec(*shape)
    assert spec.n_params == spec_o.n_params
    rng = np.random.default_rng(n + shape[2])
    theta = rng.normal(scale=0.4, size=spec_o.n_params)
    X = rng.normal(size=(n, shape[0]))
    g = rng.normal(size=n)
    y = ops.mlp_forward(spec, dev(theta), dev(X), scale=2.5)
    assert rel(y, O.mlp_forward(spec_o, theta, X, 2.5)) < 1e-13
    gt = ops.mlp_backward(spec, dev(theta), dev(g), dev(X), scale=2.5)
    assert rel(gt, O.mlp_backward(spec_o, theta, X, g, 2.5)) < 1e-11
    if n <= 130:
        jac = ops.mlp_param_jacobian(spec, dev(theta), dev(X), scale=2.5)
        for p in (0, n - 1):
            e = np.zeros(n)
            e[p] = 1.0
            assert rel(jac[p], O.mlp_backward(spec_o, theta, X, e, 2.5)) < 1e-12
    # raw (not softplus) output path
    y0 = ops.mlp_forward(spec, dev(theta), dev(X), scale=1.0, enforce_positive=False)
    assert rel(y0, O.mlp_forward(spec_o, theta, X, 1.0, enforce_positive=False)) < 1e-13


@pytest.mark.parametrize("shape", [(3, 2, 20), (3, 2, 15), (3, 2, 16), (2, 1, 7), (3, 4, 33), (5, 3, 64), (3, 3, 8)])
@pytest.mark.parametrize("n", [2048, 4100, 70000])
def test_mlp_tensor_core_path(shape, n, monkeypatch):
    """Large point sets go through the DMMA (fp64 tensor-core) kernels: same parity bar vs the oracle,
    bitwise reproducible, and consistent with the generic kernels."""
    from pinn_fem_b200 import ops

    spec_o, spec = O.NetSpec(*shape), ops.NetSpec(*shape)
    rng = np.random.default_rng(n + shape[2])
    theta = rng.normal(scale=0.4, size=spec_o.n_params)
    X = rng.normal(size=(n, shape[0]))
    g = rng.normal(size=n)
    monkeypatch.setenv("PF_MLP_FWD_TC", "1")  # forward defaults to the FMA kernel (tanh bound); test the DMMA one
    y = ops.mlp_forward(spec, dev(theta), dev(X), scale=2.5)
    assert rel(y, O.mlp_forward(spec_o, theta, X, 2.5)) < 1e-13
    gt = ops.mlp_backward(spec, dev(theta), dev(g), dev(X), scale=2.5)
    assert rel(gt, O.mlp_backward(spec_o, theta, X, g, 2.5)) < 1e-11
    assert torch.equal(gt, ops.mlp_backward(spec, dev(theta), dev(g), dev(X), scale=2.5))
    y0 = ops.mlp_forward(spec, dev(theta), dev(X), scale=1.0, enforce_positive=False)
    assert rel(y0, O.mlp_forward(spec_o, theta, X, 1.0, enforce_positive=False)) < 1e-13
    # the same points through the FMA kernel
    monkeypatch.setenv("PF_MLP_FWD_TC", "0")
    y_fma = ops.mlp_forward(spec, dev(theta), dev(X), scale=2.5)
    assert rel(y, y_fma.cpu().numpy()) < 1e-14


def test_mlp_at_plan_centroids_matches_reference_values(golden_dir):
    """NNProperty.value at the example's centroids: inputs are [load_factor, x, y] (D6)."""
    from pinn_fem_b200 import AssemblyPlan, ops

    g = np.load(golden_dir / "assembly_torch_f32.npz")
    plan = AssemblyPlan(EX_NODES, EX_EL, EX_FIXED, device="cuda")
    for name, spec in SPECS.items():
        v = ops.mlp_forward(ops.NetSpec(3, 2, spec.width), dev(g[f"theta_{name}"]), plan=plan,
                            load_factor=float(g["lam"]))
        assert rel(v, g[f"value_{name}"]) < 2e-6  # the reference evaluated in fp32
    with pytest.raises(ValueError, match="input_dim"):
        ops.mlp_forward(ops.NetSpec(1, 2, 20), dev(np.zeros(O.NetSpec(1, 2, 20).n_params)), plan=plan)


def _ex4(golden_dir):
    with open(golden_dir / "solver_runs.json") as f:
        theta0 = json.load(f)["example4-P"]["theta0"]
    return theta0


def test_gd_loop_matches_oracle_iteration_by_iteration(golden_dir, example_inputs):
    """Same theta_0, same hyper-parameters as example 4-P, load factor 0.1: the
    device loop and the fp64 oracle loop must agree row by row in the history."""
    from pinn_fem_b200 import AssemblyPlan, ops

    theta0 = _ex4(golden_dir)
    d = example_inputs["example4-P"]
    pc = d["pinn_config"]
    mesh = O.Mesh(EX_NODES, EX_EL, np.array(d["loads"], dtype=float), EX_FIXED)
    mat = O.MaterialNets(*[(SPECS[n], np.array(theta0[n]), 1.0) for n in ("young", "area", "density")])
    md, mv = np.array([2, 3, 4, 5, 6, 7]), np.array([1.0, 0, 2, 0, 3, 0])
    n_it = 300
    u_ref, reac_ref, ok_ref, hist_ref = O.solve_gd(mesh, mat, n_it, 1e-4, lr_u=pc["learning_rate_u"],
                                                   lr_theta=pc["learning_rate_theta"], alpha_p=1.0, alpha_d=100.0,
                                                   meas_dofs=md, meas_vals=mv, lam=0.1)
    plan = AssemblyPlan(EX_NODES, EX_EL, EX_FIXED, device="cuda")
    nets = [ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15), ops.NetSpec(3, 2, 10)]
    theta = dev(np.concatenate([theta0[n] for n in ("young", "area", "density")]))[None, :].clone()
    u = torch.zeros((1, 8), dtype=torch.float64, device="cuda")
    res = ops.gd_solve(plan, nets, [1.0, 1.0, 1.0], theta, u, dev(d["loads"]), md, mv, max_iterations=n_it,
                       tolerance=1e-4, learning_rate_u=pc["learning_rate_u"],
                       learning_rate_theta=pc["learning_rate_theta"], alpha_physics=1.0, alpha_data=100.0,
                       load_factor=0.1)
    n = int(res.n_iters[0])
    assert n == len(hist_ref) and bool(res.converged[0]) == ok_ref
    H = res.history[0, :n].cpu().numpy()
    keys = ["iteration", "loss_total", "loss_physics", "loss_data", "u_norm", "residual_norm", "theta_norm"]
    Href = np.array([[h[k] for k in keys] for h in hist_ref])
    # fp64 on both sides: trajectories agree to ~1e-9 relative over hundreds of Adam steps
    scale = np.maximum(np.abs(Href), 1e-6)
    assert np.max(np.abs(H - Href) / scale) < 1e-7
    assert np.max(np.abs(H[:20] - Href[:20]) / scale[:20]) < 1e-11
    assert rel(res.u[0], u_ref) < 1e-8 and rel(res.reactions[0], reac_ref) < 1e-7
    th_ref = np.concatenate([mat.young[1], mat.area[1], mat.density[1]])
    assert rel(res.theta[0], th_ref) < 1e-8
    # density parameters never move (A.4)
    assert torch.equal(res.theta[0, 837:], dev(theta0["density"]))


def test_gd_loop_against_reference_fp32_run(golden_dir, example_inputs):
    """Against the REAL reference (fp32, gd_trace fixture): first iterations agree to
    fp32 round-off, the increment converges to the same displacements."""
    from pinn_fem_b200 import AssemblyPlan, ops

    theta0 = _ex4(golden_dir)
    with open(golden_dir / "gd_trace_example4P.json") as f:
        trace = json.load(f)[0]
    d = example_inputs["example4-P"]
    pc = d["pinn_config"]
    plan = AssemblyPlan(EX_NODES, EX_EL, EX_FIXED, device="cuda")
    nets = [ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15), ops.NetSpec(3, 2, 10)]
    theta = dev(np.concatenate([theta0[n] for n in ("young", "area", "density")]))[None, :].clone()
    u = torch.zeros((1, 8), dtype=torch.float64, device="cuda")
    # preconditioning phase of solve_gd (solver.py:116-121): min(300, 5000//3) iterations at tol 1e-4
    res = ops.gd_solve(plan, nets, [1.0, 1.0, 1.0], theta, u, dev(d["loads"]), [2, 3, 4, 5, 6, 7],
                       [1.0, 0, 2, 0, 3, 0], max_iterations=300, tolerance=1e-4,
                       learning_rate_u=pc["learning_rate_u"], learning_rate_theta=pc["learning_rate_theta"],
                       alpha_physics=1.0, alpha_data=100.0, load_factor=0.1)
    H = res.history[0].cpu().numpy()
    keys = ["iteration", "loss_total", "loss_physics", "loss_data", "u_norm", "residual_norm", "theta_norm"]
    for row, hr in zip(H[:12], trace["history_head"]):
        for c, k in enumerate(keys):
            assert abs(row[c] - hr[k]) <= 2e-4 * max(abs(hr[k]), 1e-3), (k, row[c], hr[k])


def test_gd_batched_problems_are_independent(golden_dir, example_inputs):
    """nprob problems in one launch == the same problems solved one by one (bitwise)."""
    from pinn_fem_b200 import AssemblyPlan, ops

    theta0 = _ex4(golden_dir)
    d = example_inputs["example4-P"]
    plan = AssemblyPlan(EX_NODES, EX_EL, EX_FIXED, device="cuda")
    nets = [ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15), ops.NetSpec(3, 2, 10)]
    base = np.concatenate([theta0[n] for n in ("young", "area", "density")])
    rng = np.random.default_rng(0)
    nprob = 5
    thetas = base[None, :] + 0.01 * rng.normal(size=(nprob, base.size))
    mvals = np.array([1.0, 0, 2, 0, 3, 0])[None, :] * rng.uniform(0.8, 1.2, size=(nprob, 1))
    kw = dict(max_iterations=60, tolerance=1e-9, learning_rate_u=0.01, learning_rate_theta=5e-4, load_factor=0.5)
    th_b, u_b = dev(thetas).clone(), torch.zeros((nprob, 8), dtype=torch.float64, device="cuda")
    rb = ops.gd_solve(plan, nets, [1.0, 1.0, 1.0], th_b, u_b, dev(d["loads"]), [2, 3, 4, 5, 6, 7], mvals, **kw)
    for p in range(nprob):
        th_1, u_1 = dev(thetas[p:p + 1]).clone(), torch.zeros((1, 8), dtype=torch.float64, device="cuda")
        r1 = ops.gd_solve(plan, nets, [1.0, 1.0, 1.0], th_1, u_1, dev(d["loads"]), [2, 3, 4, 5, 6, 7], mvals[p], **kw)
        assert torch.equal(r1.u[0], rb.u[p]) and torch.equal(r1.theta[0], rb.theta[p])
        assert torch.equal(r1.history[0], rb.history[p])


def test_gd_scalar_materials_and_mixed(example_inputs):
    """theta = {} (example 2: scalar E, A) and E = NN only (example 3), vs the oracle."""
    from pinn_fem_b200 import AssemblyPlan, ops

    d = example_inputs["example2-P"]
    mesh = O.Mesh(EX_NODES, EX_EL, np.array(d["loads"], dtype=float), EX_FIXED)
    plan = AssemblyPlan(EX_NODES, EX_EL, EX_FIXED, device="cuda")
    md, mv = np.array([2, 3, 4, 5, 6, 7]), np.array([1.0, 0, 2, 0, 3, 0])
    mat = O.MaterialNets(1.0, 1.0, 1.0)
    u_ref, reac_ref, ok, hist = O.solve_gd(mesh, mat, 200, 1e-6, lr_u=0.01, lr_theta=1e-4, meas_dofs=md,
                                           meas_vals=mv, lam=1.0)
    res = ops.gd_solve(plan, [None, None, None], [1.0, 1.0, 1.0], None,
                       torch.zeros((1, 8), dtype=torch.float64, device="cuda"), dev(d["loads"]), md, mv,
                       max_iterations=200, tolerance=1e-6, learning_rate_u=0.01, load_factor=1.0)
    assert int(res.n_iters[0]) == len(hist)
    assert rel(res.u[0], u_ref) < 1e-9
    assert rel(res.history[0, :len(hist), 1], np.array([h["loss_total"] for h in hist])) < 1e-9
    # young = NN, area scalar 2.0, no measurements
    rng = np.random.default_rng(2)
    th = rng.normal(scale=0.3, size=SPECS["young"].n_params)
    mat = O.MaterialNets((SPECS["young"], th.copy(), 3.0), 2.0, 1.0)
    u_ref, reac_ref, ok, hist = O.solve_gd(mesh, mat, 50, 1e-12, lr_u=0.02, lr_theta=1e-3, lam=0.7)
    res = ops.gd_solve(plan, [ops.NetSpec(3, 2, 20), None, None], [3.0, 2.0, 1.0], dev(th)[None].clone(),
                       torch.zeros((1, 8), dtype=torch.float64, device="cuda"), dev(d["loads"]),
                       max_iterations=50, tolerance=1e-12, learning_rate_u=0.02, learning_rate_theta=1e-3,
                       load_factor=0.7)
    assert rel(res.u[0], u_ref) < 1e-9 and rel(res.theta[0], mat.young[1]) < 1e-9
    assert rel(res.history[0, :50, 6], np.array([h["theta_norm"] for h in hist])) < 1e-12


def test_gd_on_a_larger_mesh_vs_oracle():
    """A 6x5 lattice (49 elements... fits one CTA): exercises multi-warp item loops."""
    from pinn_fem_b200 import AssemblyPlan, ops

    nodes, el, fixed = O.lattice_truss(6, 5)
    rng = np.random.default_rng(8)
    loads = np.zeros(2 * len(nodes))
    loads[-2] = 0.05
    mesh = O.Mesh(nodes, el, loads, fixed)
    specs = (O.NetSpec(3, 2, 12), O.NetSpec(3, 3, 9))
    th = [rng.normal(scale=0.3, size=s.n_params) for s in specs]
    mat = O.MaterialNets((specs[0], th[0].copy(), 2.0), (specs[1], th[1].copy(), 1.5), 1.0)
    md = np.array([2 * 29, 2 * 29 + 1, 11])
    mv = np.array([0.01, -0.02, 0.005])
    u_ref, reac_ref, ok, hist = O.solve_gd(mesh, mat, 25, 1e-14, lr_u=1e-3, lr_theta=1e-3, alpha_d=10.0,
                                           meas_dofs=md, meas_vals=mv, lam=0.9)
    plan = AssemblyPlan(nodes, el, fixed, device="cuda")
    theta = dev(np.concatenate(th))[None].clone()
    res = ops.gd_solve(plan, [ops.NetSpec(3, 2, 12), ops.NetSpec(3, 3, 9), None], [2.0, 1.5, 1.0], theta,
                       torch.zeros((1, plan.ndof), dtype=torch.float64, device="cuda"), dev(loads), md, mv,
                       max_iterations=25, tolerance=1e-14, learning_rate_u=1e-3, learning_rate_theta=1e-3,
                       alpha_data=10.0, load_factor=0.9)
    assert rel(res.u[0], u_ref) < 1e-9 and rel(res.reactions[0], reac_ref) < 1e-8
    assert rel(res.theta[0], np.concatenate([mat.young[1], mat.area[1]])) < 1e-9
    assert rel(res.history[0, :25, 1], np.array([h["loss_total"] for h in hist])) < 1e-9


@pytest.mark.parametrize("case", ["nets", "scalar_area", "converges", "all_scalar"])
def test_gd_large_mesh_device_loop_vs_oracle(case):
    """Meshes that do not fit one CTA's shared memory run the same iteration as a device-resident
    sequence of kernels (pf_gd_large.cu): 40x40 lattice (4641 elements, MLP backward on DMMA), duplicate
    measured DOFs, parity with the oracle's solve_gd iteration by iteration."""
    from pinn_fem_b200 import AssemblyPlan, ops

    nodes, el, fixed = O.lattice_truss(40)
    rng = np.random.default_rng(21)
    loads = np.zeros(2 * len(nodes))
    loads[-2], loads[-1] = 0.05, -0.02
    mesh = O.Mesh(nodes, el, loads, fixed)
    specs = (O.NetSpec(3, 2, 20), O.NetSpec(3, 2, 15))
    th = [rng.normal(scale=0.3, size=s.n_params) for s in specs]
    md = np.array([2 * 799, 2 * 799 + 1, 411, 411, 1202])
    mv = np.array([0.01, -0.02, 0.005, 0.004, -0.003])
    n_it, tol = 20, 1e-14
    if case == "nets":
        mat = O.MaterialNets((specs[0], th[0].copy(), 2.0), (specs[1], th[1].copy(), 1.5), 1.0)
        nets, scales, theta0 = [ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15), None], [2.0, 1.5, 1.0], np.concatenate(th)
    elif case == "scalar_area":
        mat = O.MaterialNets((specs[0], th[0].copy(), 2.0), 1.5, 1.0)
        nets, scales, theta0 = [ops.NetSpec(3, 2, 20), None, None], [2.0, 1.5, 1.0], th[0]
    elif case == "all_scalar":  # no network at all: only u is optimised (example 2 on a large mesh)
        mat = O.MaterialNets(2.0, 1.5, 1.0)
        nets, scales, theta0 = [None, None, None], [2.0, 1.5, 1.0], np.zeros(0)
        # uniform materials on a regular lattice: many DOFs have gradients that cancel to round-off, and Adam's
        # g / sqrt(v) normalisation turns a 1e-16 difference in summation order into +-lr steps within ~10
        # iterations (the losses still agree to 1e-3 after 20); compare while the trajectories are comparable
        n_it = 3
    else:  # loose tolerance: stops on the loss test right after iteration 11 (checked only for it > 10)
        mat = O.MaterialNets((specs[0], th[0].copy(), 2.0), (specs[1], th[1].copy(), 1.5), 1.0)
        nets, scales, theta0 = [ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15), None], [2.0, 1.5, 1.0], np.concatenate(th)
        n_it, tol = 40, 1e3
    u_ref, reac_ref, ok, hist = O.solve_gd(mesh, mat, n_it, tol, lr_u=1e-4, lr_theta=1e-3, alpha_d=10.0,
                                           meas_dofs=md, meas_vals=mv, lam=0.9)
    plan = AssemblyPlan(nodes, el, fixed, device="cuda")
    theta = dev(theta0)[None].clone() if theta0.size else None
    res = ops.gd_solve(plan, nets, scales, theta, torch.zeros((1, plan.ndof), dtype=torch.float64, device="cuda"),
                       dev(loads), md, mv, max_iterations=n_it, tolerance=tol, learning_rate_u=1e-4,
                       learning_rate_theta=1e-3, alpha_data=10.0, load_factor=0.9)
    n = int(res.n_iters[0])
    assert n == len(hist) and bool(res.converged[0]) == bool(ok)
    if case == "converges":
        assert n == 12 and ok
    assert rel(res.u[0], u_ref) < 1e-9 and rel(res.reactions[0], reac_ref) < 1e-8
    if case == "all_scalar":
        assert res.theta.numel() == 0
    else:
        th_ref = np.concatenate([mat.young[1], mat.area[1]]) if case != "scalar_area" else mat.young[1]
        assert rel(res.theta[0], th_ref) < 1e-9
    for col, key in ((1, "loss_total"), (2, "loss_physics"), (3, "loss_data"), (4, "u_norm"), (5, "residual_norm"),
                     (6, "theta_norm")):
        if key == "theta_norm" and case == "all_scalar":
            continue
        assert rel(res.history[0, :n, col], np.array([h[key] for h in hist])) < 1e-9, key
    # bitwise reproducible
    theta2 = dev(theta0)[None].clone() if theta0.size else None
    res2 = ops.gd_solve(plan, nets, scales, theta2, torch.zeros((1, plan.ndof), dtype=torch.float64, device="cuda"),
                        dev(loads), md, mv, max_iterations=n_it, tolerance=tol, learning_rate_u=1e-4,
                        learning_rate_theta=1e-3, alpha_data=10.0, load_factor=0.9)
    assert torch.equal(res.u, res2.u) and torch.equal(res.theta, res2.theta)


def test_gd_large_mesh_graph_replay_equals_eager(monkeypatch):
    """The large-mesh loop replays its iteration as a CUDA graph by default; PF_GD_GRAPH=0 launches every kernel
    eagerly.  Same kernels, same order: the results carry the same bits (also when the loop stops early)."""
    from pinn_fem_b200 import AssemblyPlan, ops

    nodes, el, fixed = O.lattice_truss(40)
    rng = np.random.default_rng(5)
    loads = np.zeros(2 * len(nodes))
    loads[-2], loads[-1] = 0.05, -0.02
    nets = [ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15), None]
    theta0 = rng.normal(scale=0.3, size=nets[0].n_params + nets[1].n_params)
    md, mv = np.array([2 * 799, 2 * 799 + 1, 411, 1202]), np.array([0.01, -0.02, 0.005, -0.003])
    plan = AssemblyPlan(nodes, el, fixed, device="cuda")
    for n_it, tol in ((30, 1e-14), (40, 1e3)):
        out = {}
        for mode in ("0", "1", None):
            if mode is None:
                monkeypatch.delenv("PF_GD_GRAPH", raising=False)
            else:
                monkeypatch.setenv("PF_GD_GRAPH", mode)
            out[mode] = ops.gd_solve(plan, nets, [2.0, 1.5, 1.0], dev(theta0)[None].clone(),
                                     torch.zeros((1, plan.ndof), dtype=torch.float64, device="cuda"), dev(loads), md, mv,
                                     max_iterations=n_it, tolerance=tol, learning_rate_u=1e-4, learning_rate_theta=1e-3,
                                     alpha_data=10.0, load_factor=0.9)
        for mode in ("1", None):
            a, b = out["0"], out[mode]
            assert int(a.n_iters[0]) == int(b.n_iters[0]) and bool(a.converged[0]) == bool(b.converged[0])
            assert torch.equal(a.u, b.u) and torch.equal(a.theta, b.theta) and torch.equal(a.reactions, b.reactions)
            n = int(a.n_iters[0])
            assert torch.equal(a.history[0, :n], b.history[0, :n])
        if tol > 1:
            assert int(out["0"].n_iters[0]) == 12


@pytest.mark.parametrize("entry", ["pf_debug_tanh", "pf_debug_tanh_table"])
def test_hidden_layer_tanh_accuracy(entry):
    """The tanh of the MLP kernels (branch-free Estrin version; table-based version of the fragment kernels):
    absolute error <= 4.5e-16 over the whole range, exact limits, odd symmetry, NaN propagation."""
    import ctypes as C

    from pinn_fem_b200 import _lib

    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-25, 25, 200000), rng.normal(scale=1.0, size=200000), rng.normal(scale=1e-3, size=50000),
                        np.linspace(-0.7, 0.7, 100001), [0.0, -0.0, 1e-300, -1e-300, 19.9999, 20.0, 20.0001, 700.0, -700.0,
                                                         np.inf, -np.inf, np.nan]])
    xd = dev(x)
    yd = torch.empty_like(xd)
    _lib.check(getattr(_lib.load(), entry)(x.size, C.c_void_p(xd.data_ptr()), C.c_void_p(yd.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    y = yd.cpu().numpy()
    ref = np.tanh(x)
    assert np.isnan(y[-1]) and y[-3] == 1.0 and y[-2] == -1.0
    ok = ~np.isnan(x)
    assert np.max(np.abs(y[ok] - ref[ok])) <= 4.5e-16  # measured 3.3e-16 (3 ulp at 1)
    assert np.all(np.abs(y[ok]) <= 1.0)
    xs = np.abs(x[ok])
    ys = dev(xs)
    ym = torch.empty_like(ys)
    _lib.check(getattr(_lib.load(), entry)(xs.size, C.c_void_p(ys.data_ptr()), C.c_void_p(ym.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    assert np.array_equal(np.abs(y[ok]), ym.cpu().numpy())  # odd

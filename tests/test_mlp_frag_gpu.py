"""GPU parity: the register-resident DMMA kernels of the material networks (pf_mlp_frag.cu) -- large point sets and
batches of problems with per-problem parameters -- against the oracle (fem/properties.py:150-156,
examples/json/generic.py:118-142)."""
import numpy as np
import pytest
import torch

from oracle import pinnfem_oracle as O

pytestmark = pytest.mark.gpu

# every (tile count, padding kind) combination of the kernels: width % 8 in {0}, {1..4}, {5..7}
SHAPES = [(3, 2, 20), (3, 2, 15), (3, 2, 10), (3, 2, 16), (3, 2, 24), (3, 2, 8), (2, 1, 7), (3, 3, 8), (3, 1, 20),
          (3, 3, 20), (1, 2, 4), (3, 2, 21), (3, 2, 5), (2, 3, 13), (3, 3, 24)]


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64).cuda()


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("n", [2048, 4101, 70001])
def test_single_problem_entry_points(shape, n, monkeypatch):
    """pf_mlp_forward / pf_mlp_backward route n >= 2048 points through the fragment kernels."""
    from pinn_fem_b200 import ops

    spec_o, spec = O.NetSpec(*shape), ops.NetSpec(*shape)
    assert ops.mlp_acts_len(spec, n) > 0
    rng = np.random.default_rng(n + 7 * shape[2] + shape[1])
    theta = rng.normal(scale=0.4, size=spec_o.n_params)
    X = rng.normal(size=(n, shape[0]))
    g = rng.normal(size=n)
    y = ops.mlp_forward(spec, dev(theta), dev(X), scale=2.5)
    assert rel(y, O.mlp_forward(spec_o, theta, X, 2.5)) < 1e-13
    gt = ops.mlp_backward(spec, dev(theta), dev(g), dev(X), scale=2.5)
    assert rel(gt, O.mlp_backward(spec_o, theta, X, g, 2.5)) < 1e-11
    assert torch.equal(gt, ops.mlp_backward(spec, dev(theta), dev(g), dev(X), scale=2.5))  # fixed reduction order
    y0 = ops.mlp_forward(spec, dev(theta), dev(X), scale=1.0, enforce_positive=False)
    assert rel(y0, O.mlp_forward(spec_o, theta, X, 1.0, enforce_positive=False)) < 1e-13
    # the same points through the older kernels (table-free tanh): agreement at rounding level
    monkeypatch.setenv("PF_MLP_NO_FRAG", "1")
    assert rel(y, ops.mlp_forward(spec, dev(theta), dev(X), scale=2.5).cpu().numpy()) < 1e-14
    assert rel(gt, ops.mlp_backward(spec, dev(theta), dev(g), dev(X), scale=2.5).cpu().numpy()) < 1e-11


@pytest.mark.parametrize("shape", [(3, 2, 20), (3, 2, 15), (3, 2, 10), (3, 3, 8), (2, 1, 7), (3, 2, 24)])
@pytest.mark.parametrize("n,B", [(1, 3), (33, 2), (3000, 5), (40000, 17), (700, 300)])
def test_batched_problems_with_saved_activations(shape, n, B):
    """B parameter sets at the same points: out[n][B], then dL/dtheta[B] from the saved activation record."""
    from pinn_fem_b200 import ops

    spec_o, spec = O.NetSpec(*shape), ops.NetSpec(*shape)
    rng = np.random.default_rng(n + B + shape[2])
    theta = rng.normal(scale=0.4, size=(B, spec_o.n_params))
    X = rng.normal(size=(n, shape[0]))
    g = rng.normal(size=(n, B))
    out, acts = ops.mlp_forward_batched(spec, dev(theta), dev(X), scale=1.7)
    gt = ops.mlp_backward_batched(spec, dev(theta), dev(g), acts, dev(X))
    check = range(B) if B <= 17 else (0, 1, B // 2, B - 1)
    for p in check:
        assert rel(out[:, p], O.mlp_forward(spec_o, theta[p], X, 1.7)) < 1e-13
        assert rel(gt[p], O.mlp_backward(spec_o, theta[p], X, g[:, p], 1.7)) < 1e-11
    # reproducible, and the forward does not depend on the batch it runs in
    out2, acts2 = ops.mlp_forward_batched(spec, dev(theta), dev(X), scale=1.7)
    assert torch.equal(out, out2) and torch.equal(acts, acts2)
    assert torch.equal(gt, ops.mlp_backward_batched(spec, dev(theta), dev(g), acts, dev(X)))
    one, _ = ops.mlp_forward_batched(spec, dev(theta[1:2]), dev(X), scale=1.7)
    assert torch.equal(one[:, 0], out[:, 1])
    # raw output (no softplus)
    raw, acts_raw = ops.mlp_forward_batched(spec, dev(theta), dev(X), scale=1.0, enforce_positive=False)
    assert rel(raw[:, B - 1], O.mlp_forward(spec_o, theta[B - 1], X, 1.0, enforce_positive=False)) < 1e-13
    gt_raw = ops.mlp_backward_batched(spec, dev(theta), dev(g), acts_raw, dev(X))
    assert rel(gt_raw[0], O.mlp_backward(spec_o, theta[0], X, g[:, 0], 1.0, enforce_positive=False)) < 1e-11


def test_batched_at_plan_centroids_and_strided_theta():
    """Centroid mode ([load_factor, x_c, y_c], D6) with theta rows embedded in a wider [B][n_theta_total] array, as the
    batched PINN loop holds them."""
    from pinn_fem_b200 import AssemblyPlan, ops

    nodes, el, fixed = O.lattice_truss(40)
    plan = AssemblyPlan(nodes, el, fixed, device="cuda")
    sE, sA = ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15)
    oE, oA = O.NetSpec(3, 2, 20), O.NetSpec(3, 2, 15)
    B = 6
    rng = np.random.default_rng(5)
    theta = rng.normal(scale=0.3, size=(B, oE.n_params + oA.n_params))
    th = dev(theta)
    cen = 0.5 * (nodes[el[:, 0]] + nodes[el[:, 1]])
    Xc = np.concatenate([np.full((len(el), 1), 0.7), cen], axis=1)
    E, actsE = ops.mlp_forward_batched(sE, th, plan=plan, load_factor=0.7, scale=2.0)
    A, actsA = ops.mlp_forward_batched(sA, th[:, oE.n_params:], plan=plan, load_factor=0.7, scale=0.5)
    g = rng.normal(size=(len(el), B))
    gE = ops.mlp_backward_batched(sE, th, dev(g), actsE, plan=plan, load_factor=0.7)
    gA = ops.mlp_backward_batched(sA, th[:, oE.n_params:], dev(g), actsA, plan=plan, load_factor=0.7)
    for p in range(B):
        assert rel(E[:, p], O.mlp_forward(oE, theta[p, :oE.n_params], Xc, 2.0)) < 1e-13
        assert rel(A[:, p], O.mlp_forward(oA, theta[p, oE.n_params:], Xc, 0.5)) < 1e-13
        assert rel(gE[p], O.mlp_backward(oE, theta[p, :oE.n_params], Xc, g[:, p], 2.0)) < 1e-11
        assert rel(gA[p], O.mlp_backward(oA, theta[p, oE.n_params:], Xc, g[:, p], 0.5)) < 1e-11


def test_uncovered_shapes_are_rejected_loudly():
    from pinn_fem_b200 import ops

    spec = ops.NetSpec(3, 4, 33)
    assert ops.mlp_acts_len(spec, 1000) == 0
    with pytest.raises(ValueError, match="not covered"):
        ops.mlp_forward_batched(spec, dev(np.zeros((2, spec.n_params))), dev(np.zeros((10, 3))))

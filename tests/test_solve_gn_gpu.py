"""GPU parity: dense LU, CG, the closed-form Gauss-Newton Jacobian and J^T J (DMMA)."""
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


@pytest.mark.parametrize("n,batch", [(1, 1), (3, 7), (17, 4), (120, 2), (121, 1), (300, 2), (1001, 1)])
def test_dense_lu_solve(n, batch):
    from pinn_fem_b200 import ops

    rng = np.random.default_rng(n)
    A = rng.normal(size=(batch, n, n)) + (0.0 if n < 50 else 0.05 * n) * np.eye(n)
    A[:, 0, 0] = 1e-14  # forces a row interchange in the first column
    b = rng.normal(size=(batch, n))
    x = ops.solve_dense(dev(A), dev(b))
    x_ref = np.linalg.solve(A, b[..., None])[..., 0]
    assert rel(x, x_ref) < 1e-9
    res = np.einsum("bij,bj->bi", A, x.cpu().numpy()) - b
    assert np.max(np.abs(res)) < 1e-9 * max(1.0, np.max(np.abs(x_ref))) * n
    if batch == 1:
        assert rel(ops.solve_dense(dev(A[0]), dev(b[0])), x_ref[0]) < 1e-9


@pytest.mark.parametrize("n", [121, 300, 1001, 1500])
def test_lu_cluster_panel_equals_single_cta_panel(n, monkeypatch):
    """The panel factorisation on a thread-block cluster (slabs in distributed shared memory) picks the same pivots
    and does the same arithmetic as the single-CTA panel: identical bits, also with ties in the pivot search."""
    from pinn_fem_b200 import ops

    rng = np.random.default_rng(100 + n)
    A = rng.normal(size=(n, n))
    A[:, 0] = np.where(rng.random(n) < 0.5, 2.0, -2.0)   # every row ties in the first column: lowest index wins
    A[5:40, 7] = 0.0
    A[n // 2] = A[n // 3] * 1.0                          # duplicate row -> exactly singular
    b = rng.normal(size=n)
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("PF_LU_CLUSTER", mode)
        try:
            outs[mode] = ops.solve_dense(dev(A), dev(b))
        except RuntimeError:
            outs[mode] = None
    assert (outs["0"] is None) == (outs["1"] is None)
    A[n // 2] = rng.normal(size=n)                       # regular again
    for mode in ("0", "1"):
        monkeypatch.setenv("PF_LU_CLUSTER", mode)
        outs[mode] = ops.solve_dense(dev(A), dev(b))
    assert torch.equal(outs["0"], outs["1"])
    assert rel(outs["1"], np.linalg.solve(A, b)) < 1e-8


def test_dense_solve_singular_raises():
    from pinn_fem_b200 import ops

    A = np.diag([1.0, 0.0, 2.0])
    with pytest.raises(RuntimeError, match="Singular"):
        ops.solve_dense(dev(A), dev(np.ones(3)))
    big = np.eye(200)
    big[57, 57] = 0.0
    with pytest.raises(RuntimeError, match="Singular"):
        ops.solve_dense(dev(big), dev(np.ones(200)))


@pytest.mark.parametrize("n", [3, 150, 700])
@pytest.mark.parametrize("mode", ["0", "1"])
def test_dense_solve_nan_matrix_raises_instead_of_faulting(n, mode, monkeypatch):
    """A diverged iterate hands the Newton / LM step a matrix of NaNs.  Every pivot search then finds no candidate
    (NaN compares false): the small-system kernel, the single-CTA panel and the cluster panel (which used to index
    distributed shared memory with the INT_MAX sentinel) must report the system as singular, and the device must
    stay usable."""
    from pinn_fem_b200 import ops

    monkeypatch.setenv("PF_LU_CLUSTER", mode)
    rng = np.random.default_rng(n)
    A = rng.normal(size=(n, n)) + n * np.eye(n)
    b = rng.normal(size=n)
    bad = A.copy()
    bad[:, n // 2:] = np.nan
    with pytest.raises(RuntimeError, match="Singular"):
        ops.solve_dense(dev(bad), dev(b))
    with pytest.raises(RuntimeError, match="Singular"):
        ops.solve_dense(dev(np.full((n, n), np.nan)), dev(b))
    assert rel(ops.solve_dense(dev(A), dev(b)), np.linalg.solve(A, b)) < 1e-9  # no sticky error


def test_cg_matches_dense_solution():
    from pinn_fem_b200 import AssemblyPlan, ops

    nodes, el, fixed = O.lattice_truss(14, 9)
    plan = AssemblyPlan(nodes, el, fixed, device="cuda")
    rng = np.random.default_rng(0)
    B = 40
    E = rng.uniform(0.5, 1.5, (plan.nelem, B))
    A = rng.uniform(0.5, 1.5, (plan.nelem, B))
    rhs = rng.normal(size=(plan.ndof, B))
    x, iters, resid = ops.cg_solve(plan, dev(E), dev(A), dev(rhs), rel_tol=1e-12, max_iters=5000)
    assert resid <= 1e-12 and 0 < iters < 5000
    free = plan.free_dofs
    for b in (0, B - 1):
        K = plan.tangent_dense(dev(E[:, b].copy()), dev(A[:, b].copy())).cpu().numpy()
        x_ref = np.zeros(plan.ndof)
        x_ref[free] = np.linalg.solve(K[np.ix_(free, free)], rhs[free, b])
        assert rel(x[:, b], x_ref) < 1e-9
    assert float(x[torch.as_tensor(plan.fixed_dofs).cuda()].abs().max()) == 0.0


def _ex10(golden_dir):
    g = np.load(golden_dir / "gauss_newton_ex10.npz")
    mat = O.MaterialNets(*[(SPECS[n], g[f"theta0_{n}"].copy(), 1.0) for n in ("young", "area", "density")])
    return g, mat


def test_gauss_newton_jacobian_and_normal_equations(golden_dir, example_inputs):
    """C4: example 10's model at u = [0,0,.5,0,1,0,1.5,0]: J (6 x 1001), JtJ, Jtr, damping, dx
    against the fp64 oracle (tight) and the reference's fp32 fixture (fp32 round-off)."""
    from pinn_fem_b200 import AssemblyPlan, ops

    g, mat = _ex10(golden_dir)
    d = example_inputs["example10"]
    mesh = O.Mesh(EX_NODES, EX_EL, np.array(d["loads"], dtype=float), EX_FIXED)
    md = np.array([2, 4, 6])
    j_uu, j_ut, r_p, j_du = O.jacobian_blocks(mesh, mat, g["u"], mesh.loads, md)
    J_ref, R_ref = O.gauss_newton_system(j_uu, j_ut, r_p, j_du, np.array([1.0, 2, 3]) - g["u"][md], 1.0, 1.0)

    plan = AssemblyPlan(EX_NODES, EX_EL, EX_FIXED, device="cuda")
    nets = {n: ops.NetSpec(3, 2, SPECS[n].width) for n in SPECS}
    th = {n: dev(g[f"theta0_{n}"]) for n in SPECS}
    E = ops.mlp_forward(nets["young"], th["young"], plan=plan, load_factor=1.0)
    A = ops.mlp_forward(nets["area"], th["area"], plan=plan, load_factor=1.0)
    jacE = ops.mlp_param_jacobian(nets["young"], th["young"], plan=plan, load_factor=1.0)
    jacA = ops.mlp_param_jacobian(nets["area"], th["area"], plan=plan, load_factor=1.0)
    u = dev(g["u"])
    J = ops.gn_jacobian(plan, u, E, A, jacE, jacA, n_rest=161, alpha_physics=1.0, alpha_data=1.0, meas_dofs=md)
    assert tuple(J.shape) == (6, 1001)
    assert rel(J, J_ref) < 1e-11
    assert rel(J, g["J"]) < 2e-5  # the reference's own fp32 Jacobian
    out = plan.residual(u, E, A, dev(mesh.loads), 1.0, f_int=False, r=True)
    free = torch.as_tensor(plan.free_dofs).cuda()
    R = torch.cat([out["r"][free], dev(np.array([1.0, 2, 3])) - u[torch.as_tensor(md).cuda()]])
    assert rel(R, R_ref) < 1e-12
    jtj, jtr, damping = ops.gn_normal_equations(J, R, 1e-6)
    dx_ref, jtj_ref, jtr_ref, damp_ref = O.lm_step(J_ref, R_ref)
    reg_ref = jtj_ref + damp_ref * np.eye(1001)
    assert rel(jtj, reg_ref) < 1e-11 and rel(jtr, jtr_ref) < 1e-11
    assert abs(float(damping[0]) - damp_ref) < 1e-12 * damp_ref
    assert rel(torch.diagonal(jtj) - damping, g["jtj_diag"]) < 5e-5  # reference fp32
    dx = ops.solve_dense(jtj, -jtr)
    # (JtJ + dI) is ill conditioned (rank 6 + damping): compare through the residual of the system
    # and on the well-determined displacement part
    resid = reg_ref @ dx.cpu().numpy() + jtr_ref
    assert np.linalg.norm(resid) < 1e-6 * np.linalg.norm(jtr_ref)
    assert np.linalg.norm(dx.cpu().numpy()[:3] - dx_ref[:3]) < 1e-5 * np.linalg.norm(dx_ref[:3])


@pytest.mark.parametrize("m,n", [(6, 1001), (70, 130), (512, 1100), (1, 5)])
def test_jtj_dmma_random(m, n):
    """J^T J on the fp64 tensor cores vs NumPy for ragged shapes."""
    from pinn_fem_b200 import ops

    rng = np.random.default_rng(m * n)
    J = rng.normal(size=(m, n))
    R = rng.normal(size=m)
    jtj, jtr, damping = ops.gn_normal_equations(dev(J), dev(R), 1e-6)
    ref = J.T @ J
    d = 1e-6 * np.trace(ref) / n
    assert rel(jtj, ref + d * np.eye(n)) < 1e-12
    assert rel(jtr, J.T @ R) < 1e-12
    assert torch.equal(jtj, jtj.T)


@pytest.mark.parametrize("n", [1, 5, 32, 33, 100, 257, 1001, 2100])
def test_cholesky_spd_solve(n):
    """Blocked Cholesky (panel solve + DMMA trailing update + one-CTA triangular solves) against numpy, incl. ragged
    last panels, and the not-positive-definite report."""
    from pinn_fem_b200 import ops

    rng = np.random.default_rng(n)
    M = rng.normal(size=(n, n + 3))
    A = M @ M.T + 1e-3 * np.eye(n)
    b = rng.normal(size=n)
    x_ref = np.linalg.solve(A, b)
    Au = A.copy()
    Au[np.triu_indices(n, 1)] = np.nan  # the upper triangle must not be read
    x = ops.solve_spd(dev(Au), dev(b)).cpu().numpy()
    assert np.linalg.norm(A @ x - b) <= 1e-10 * np.linalg.norm(A, 2) * np.linalg.norm(x_ref) * max(n, 1) ** 0.5 + 1e-12
    assert np.linalg.norm(x - x_ref) <= 1e-8 * np.linalg.cond(A) ** 0.5 * np.linalg.norm(x_ref) + 1e-12
    assert torch.equal(ops.solve_spd(dev(Au), dev(b)), ops.solve_spd(dev(Au), dev(b)))  # deterministic
    if n >= 5:
        bad = A.copy()
        bad[n // 2, n // 2] = -1.0
        with pytest.raises(RuntimeError, match="not positive definite"):
            ops.solve_spd(dev(bad), dev(b))


@pytest.mark.parametrize("m,n", [(6, 1001), (40, 300), (200, 64), (130, 130)])
def test_lm_step_dual_and_primal(m, n):
    """pf_gn_lm_step: dx = -(J^T J + d I)^-1 J^T R through the n x n system and through the m x m dual system
    (chosen when m < n) against numpy's evaluation of the same two formulas."""
    from pinn_fem_b200 import ops

    rng = np.random.default_rng(m + n)
    J = rng.normal(size=(m, n))
    R = rng.normal(size=m)
    d = 1e-6 * np.trace(J.T @ J) / n
    dual_ref = -(J.T @ np.linalg.solve(J @ J.T + d * np.eye(m), R))
    dx_dual, damp, _ = ops.gn_lm_step(dev(J), dev(R), 1e-6, path="dual")
    assert abs(float(damp[0]) - d) <= 1e-12 * d
    assert rel(dx_dual, dual_ref) < (1e-10 if m <= n else 1e-6)  # forced on m > n the dual is the ill-conditioned one
    dx_auto, _, _ = ops.gn_lm_step(dev(J), dev(R), 1e-6)
    dx_primal, _, _ = ops.gn_lm_step(dev(J), dev(R), 1e-6, path="primal")
    assert torch.equal(dx_auto, dx_dual if m < n else dx_primal)
    # the two formulations are the same vector; the n x n one is solved in a matrix of condition ~1e6 and worse when
    # m < n (rank-m matrix + ridge), so the agreement is bounded by that, not by the kernels
    primal_ref = np.linalg.solve(J.T @ J + d * np.eye(n), -(J.T @ R))
    bound = 1e-9 if m >= n else 1e-5
    assert rel(dx_primal, primal_ref) < bound and rel(dx_primal, dual_ref) < bound

"""GPU parity tests: the CUDA assembly kernels, called through the C ABI,
against the CPU oracle on the same seeded inputs.

Tolerances (north star): 1e-10 relative (max-norm) on f_int / K_t and the
derived products; index structures bit-exact (tests/test_cabi_and_plan.py)."""
import numpy as np
import pytest
import torch

from oracle import pinnfem_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64).cuda()


def case(assembly_golden, name):
    return {k.split(".", 1)[1]: assembly_golden[k] for k in assembly_golden.files if k.startswith(name + ".")}


@pytest.fixture(scope="module")
def plans():
    from pinn_fem_b200 import AssemblyPlan

    cache = {}

    def get(key, nodes, el, fixed, dim=None):
        if key not in cache:
            cache[key] = AssemblyPlan(nodes, el, fixed, dim=dim, device="cuda")
        return cache[key]

    return get


@pytest.mark.parametrize("name", ["ex1", "fem2d_like", "lattice8", "lattice5x3_perturbed", "bar1d"])
def test_against_reference_goldens(assembly_golden, plans, name):
    """Directly against what the reference's assemble_system returned."""
    c = case(assembly_golden, name)
    dim = int(c["dim"])
    p = plans(name, c["nodes"], c["elements"], c["fixed_in"], dim)
    u, E, A = dev(c["u"]), dev(c["E"]), dev(c["A"])
    out = p.residual(u, E, A, max_strain=True)
    assert rel(out["f_int"], c["f_int"]) < TOL
    assert abs(float(out["max_strain"][0]) - float(c["max_strain"])) < 1e-12 * max(1.0, float(c["max_strain"]))
    K = p.tangent_dense(E, A)
    assert rel(K, c["K"]) < TOL
    assert torch.equal(K, K.T)
    free = torch.as_tensor(c["free"]).cuda()
    Kff = p.tangent_dense(E, A, free_only=True)
    assert torch.equal(Kff, K[free][:, free])
    if dim == 2:
        ug = dev(c["u_gl"])
        assert rel(p.internal_force(ug, E, A, kind="gl"), c["f_gl"]) < TOL
        assert rel(p.tangent_dense(E, A, ug, kind="gl"), c["K_gl"]) < TOL


def test_known_answers(plans):
    """reference test_torch_element.py T2-A / T2-B on the device."""
    nodes = np.array([[0.0, 0], [1, 0], [2, 0], [3, 0]])
    el = np.array([[0, 1], [1, 2], [2, 3]])
    p = plans("kat", nodes, el, [0, 1, 3, 5, 7])
    one = dev(np.ones(3))
    f_ext = dev([0, 0, 0, 0, 0, 0, 1.0, 0])
    out = p.residual(dev(np.zeros(8)), one, one, f_ext, 1.0, r=True, half_sq=True)
    assert out["f_int"].tolist() == [0.0] * 8
    assert out["r"].tolist() == [0, 0, 0, 0, 0, 0, -1.0, 0]
    assert float(out["half_sq"][0]) == 0.5
    out = p.residual(dev([0, 0, 1.0, 0, 2, 0, 3, 0]), one, one, f_ext, 1.0, r=True)
    assert out["f_int"].tolist() == [-1.0, 0, 0, 0, 0, 0, 1.0, 0]
    assert out["r"].tolist() == [0.0] * 8
    g = dev([0, 0, 0, 0, 0, 0, -2.0 / 3.0, 0])
    assert np.allclose(p.tangent_matvec(g, one, one).cpu().numpy(), [0, 0, 0, 0, 2 / 3, 0, -2 / 3, 0], atol=1e-15)


def _random_graph(seed, nnode=300, nelem=1200):
    rng = np.random.default_rng(seed)
    nodes = rng.normal(size=(nnode, 2))
    el = rng.integers(0, nnode, size=(nelem, 2))
    el = el[el[:, 0] != el[:, 1]]
    return nodes, el, rng.integers(0, 2 * nnode, size=31)


@pytest.mark.parametrize("B", [1, 3, 32, 70, 257])
@pytest.mark.parametrize("kind", ["linear", "gl"])
def test_batched_random_graph(plans, B, kind):
    """Ragged batch sizes, duplicate edges, both element kinds, all kernels."""
    nodes, el, fixed = _random_graph(7)
    p = plans("rg", nodes, el, fixed)
    assert p.has_duplicate_edges
    okind = O.LINEAR if kind == "linear" else O.GREEN_LAGRANGE
    rng = np.random.default_rng(B)
    ndof, ne = 2 * len(nodes), len(el)
    scale = 1e-3 if kind == "linear" else 5e-2
    shp = (lambda n: (n,)) if B == 1 else (lambda n: (n, B))
    u = rng.uniform(-scale, scale, size=shp(ndof))
    E = rng.uniform(0.5, 1.5, size=shp(ne))
    A = rng.uniform(0.5, 1.5, size=shp(ne))
    fx = rng.normal(size=ndof)
    g = rng.normal(size=shp(ndof))
    lam = 0.37
    free, _ = O.free_and_fixed_dofs(ndof, fixed)
    f_ref, eps_ref = O.assemble_residual(nodes, el, E, A, u, 2, okind)
    r_ref = np.zeros_like(f_ref)
    r_ref[free] = f_ref[free] - lam * (fx[free] if B == 1 else fx[free][:, None])
    out = p.residual(dev(u), dev(E), dev(A), dev(fx), lam, kind=kind, r=True, half_sq=True, max_strain=True)
    assert rel(out["f_int"], f_ref) < TOL and rel(out["r"], r_ref) < TOL
    assert rel(out["half_sq"], np.atleast_1d(0.5 * np.sum(r_ref * r_ref, axis=0))) < TOL
    assert rel(out["max_strain"], np.atleast_1d(eps_ref)) < 1e-9
    kv_ref = O.tangent_matvec(nodes, el, E, A, g, u, 2, okind)
    assert rel(p.tangent_matvec(dev(g), dev(E), dev(A), dev(u), kind=kind), kv_ref) < TOL
    gE_ref, gA_ref = O.material_vjp(nodes, el, E, A, u, g, 2, okind)
    gE, gA = p.material_vjp(dev(u), dev(E), dev(A), dev(g), kind=kind)
    assert rel(gE, gE_ref) < 1e-9 and rel(gA, gA_ref) < 1e-9
    # shared (un-batched) materials against per-problem copies
    if B > 1:
        E1, A1 = E[:, 0].copy(), A[:, 0].copy()
        f_sh = p.internal_force(dev(u), dev(E1), dev(A1), kind=kind)
        f_pp = p.internal_force(dev(u), dev(np.repeat(E1[:, None], B, 1)), dev(np.repeat(A1[:, None], B, 1)), kind=kind)
        assert torch.equal(f_sh, f_pp)
    # batched tangent values against the un-batched oracle, two columns
    vals = p.tangent_bsr(dev(E), dev(A), dev(u), kind=kind)
    for b in sorted({0, B - 1}):
        Eb, Ab, ub = (E, A, u) if B == 1 else (E[:, b], A[:, b], u[:, b])
        _, _, v_ref = O.assemble_tangent_bsr(nodes, el, Eb, Ab, ub, 2, okind)
        vb = vals if B == 1 else vals[:, :, b].reshape(-1, 2, 2)
        assert rel(vb, v_ref) < TOL


def test_1d_batched_patch_path(plans):
    """1-D bars through the wide-batch (patch-staged) path: B = 160 -> 5 chunks."""
    rng = np.random.default_rng(5)
    x = np.sort(rng.uniform(0, 10, size=200))
    el = np.stack([np.arange(199), np.arange(1, 200)], axis=1)
    el = np.concatenate([el, [[0, 5], [7, 3], [100, 150]]])
    p = plans("bar1d_big", x, el, [0, 199], 1)
    B = 160
    u = rng.uniform(-1e-3, 1e-3, (200, B))
    E = rng.uniform(0.5, 1.5, (len(el), B))
    A = rng.uniform(0.5, 1.5, (len(el), B))
    fx = rng.normal(size=(200, B))
    out = p.residual(dev(u), dev(E), dev(A), dev(fx), 0.9, r=True, half_sq=True, max_strain=True)
    f_ref, eps_ref = O.assemble_residual(x, el, E, A, u, 1)
    r_ref = f_ref - 0.9 * fx
    r_ref[[0, 199]] = 0.0
    assert rel(out["f_int"], f_ref) < TOL and rel(out["r"], r_ref) < TOL
    assert rel(out["half_sq"], 0.5 * np.sum(r_ref * r_ref, axis=0)) < TOL
    assert rel(out["max_strain"], eps_ref) < 1e-9
    assert rel(p.tangent_matvec(dev(fx), dev(E), dev(A)), O.assemble_residual(x, el, E, A, fx, 1)[0]) < TOL


def test_patch_and_generic_paths_agree_bitwise(plans):
    """The same columns evaluated by the patch-staged kernel (wide batch) and by
    the generic gather (narrow batch) must give identical bits."""
    nodes, el, fixed = _random_graph(3)
    p = plans("rg3", nodes, el, fixed)
    rng = np.random.default_rng(9)
    B = 192
    u = rng.uniform(-1e-3, 1e-3, (2 * len(nodes), B))
    E = rng.uniform(0.5, 1.5, (len(el), B))
    A = rng.uniform(0.5, 1.5, (len(el), B))
    wide = p.internal_force(dev(u), dev(E), dev(A))
    for b0 in (0, 64, 160):
        narrow = p.internal_force(dev(u[:, b0:b0 + 32]), dev(E[:, b0:b0 + 32]), dev(A[:, b0:b0 + 32]))
        assert torch.equal(wide[:, b0:b0 + 32], narrow)


def test_determinism_and_symmetry(plans):
    nodes, el, fixed = _random_graph(11)
    p = plans("rg11", nodes, el, fixed)
    rng = np.random.default_rng(0)
    B = 64
    u, E, A = (dev(rng.uniform(-1e-3, 1e-3, (2 * len(nodes), B))), dev(rng.uniform(0.5, 1.5, (len(el), B))),
               dev(rng.uniform(0.5, 1.5, (len(el), B))))
    f1 = p.internal_force(u, E, A)
    f2 = p.internal_force(u, E, A)
    assert torch.equal(f1, f2)  # bitwise reproducible: no atomics in the sum
    # linearity in u (linear element): f(a u1 + u2) = a f(u1) + f(u2)
    u2 = dev(rng.uniform(-1e-3, 1e-3, (2 * len(nodes), B)))
    lhs = p.internal_force(2.5 * u + u2, E, A)
    assert rel(lhs, (2.5 * f1 + p.internal_force(u2, E, A)).cpu().numpy()) < 1e-12
    # <v, K w> == <w, K v>
    v, w = dev(rng.normal(size=(2 * len(nodes), B))), dev(rng.normal(size=(2 * len(nodes), B)))
    a = (v * p.tangent_matvec(w, E, A)).sum(0)
    b = (w * p.tangent_matvec(v, E, A)).sum(0)
    assert rel(a, b.cpu().numpy()) < 1e-11
    # rigid translation produces no force
    t = torch.ones_like(u)
    assert float(p.internal_force(t, E, A).abs().max()) < 1e-12


def test_c5s_lattices_and_scale_properties(plans):
    """C5s: lattices where the oracle still runs in seconds, then a
    larger lattice checked through size-independent properties."""
    for nx in (8, 24, 40):
        nodes, el, fixed = O.lattice_truss(nx)
        p = plans(f"lat{nx}", nodes, el, fixed)
        rng = np.random.default_rng(nx)
        B = 8
        u = rng.uniform(-1e-3, 1e-3, (2 * len(nodes), B))
        E = rng.uniform(0.5, 1.5, (len(el), B))
        A = rng.uniform(0.5, 1.5, (len(el), B))
        f_ref, _ = O.assemble_residual(nodes, el, E, A, u)
        assert rel(p.internal_force(dev(u), dev(E), dev(A)), f_ref) < TOL
    nx = 300
    nodes, el, fixed = O.lattice_truss(nx)
    p = plans("lat300", nodes, el, fixed)
    assert p.nelem == 2 * (nx - 1) * nx + (nx - 1) ** 2 and p.max_degree == 6
    rng = np.random.default_rng(1)
    B = 16
    u = dev(rng.uniform(-1e-3, 1e-3, (p.ndof, B)))
    E = dev(rng.uniform(0.5, 1.5, (p.nelem, B)))
    A = dev(rng.uniform(0.5, 1.5, (p.nelem, B)))
    f = p.internal_force(u, E, A)
    # a self-equilibrated internal force field: every column sums to zero per direction
    assert float(f.view(p.nnode, 2, B).sum(0).abs().max()) < 1e-9
    # energy identity  u . f = sum_e EA/l0 * axial^2  (linear element)
    l0 = torch.as_tensor(p.geometry("l0")).cuda()[:, None]
    c = torch.as_tensor(p.geometry("cos")).cuda()[:, None]
    s = torch.as_tensor(p.geometry("sin")).cuda()[:, None]
    elt = torch.as_tensor(el).cuda()
    ux, uy = u[0::2], u[1::2]
    axial = c * (ux[elt[:, 1]] - ux[elt[:, 0]]) + s * (uy[elt[:, 1]] - uy[elt[:, 0]])
    energy = (E * A / l0 * axial * axial).sum(0)
    assert rel((u * f).sum(0), energy.cpu().numpy()) < 1e-11


def test_residual_host_roundtrip(plans):
    nodes, el, fixed = O.lattice_truss(24)
    p = plans("lat24", nodes, el, fixed)
    rng = np.random.default_rng(2)
    B = 150  # not a multiple of the chunk
    u = torch.as_tensor(rng.uniform(-1e-3, 1e-3, (p.ndof, B))).pin_memory()
    E = torch.as_tensor(rng.uniform(0.5, 1.5, (p.nelem, B))).pin_memory()
    A = torch.as_tensor(rng.uniform(0.5, 1.5, (p.nelem, B))).pin_memory()
    fx = torch.as_tensor(rng.normal(size=p.ndof))
    r_host = p.residual_host(u, E, A, fx, 0.8, chunk=64)
    r_dev = p.residual(u.cuda(), E.cuda(), A.cuda(), fx.cuda(), 0.8, f_int=False, r=True)["r"]
    assert torch.equal(r_host, r_dev.cpu())


@pytest.mark.parametrize("mesh", ["lattice", "random_dups", "bar1d"])
def test_single_problem_kernel_matches_batched_bitwise(plans, mesh):
    """B = 1 goes through its own kernel (compact index table, geometry recomputed from the coordinates with
    correctly-rounded operations): f_int, r and K v must carry the same bits as column 0 of a batched call."""
    rng = np.random.default_rng(17)
    if mesh == "lattice":
        nodes, el, fixed = O.lattice_truss(9, 7)
        nodes = nodes + rng.uniform(-0.2, 0.2, nodes.shape)
        dim = 2
    elif mesh == "random_dups":
        nodes, el, fixed = _random_graph(5)
        dim = 2
    else:
        nodes = np.sort(rng.uniform(0, 10, size=60))
        el = np.concatenate([np.stack([np.arange(59), np.arange(1, 60)], axis=1), [[0, 7], [30, 10], [7, 0]]])
        fixed, dim = [0, 59], 1
    p = plans("b1_" + mesh, nodes, el, fixed, dim)
    B = 3
    u = rng.uniform(-1e-3, 1e-3, (p.ndof, B))
    E = rng.uniform(0.5, 1.5, (len(el), B))
    A = rng.uniform(0.5, 1.5, (len(el), B))
    fx = rng.normal(size=p.ndof)
    v = rng.normal(size=(p.ndof, B))
    ob = p.residual(dev(u), dev(E), dev(A), dev(fx), 0.7, r=True, half_sq=True, max_strain=True)
    kb = p.tangent_matvec(dev(v), dev(E), dev(A))
    for b in range(B):
        o1 = p.residual(dev(u[:, b].copy()), dev(E[:, b].copy()), dev(A[:, b].copy()), dev(fx), 0.7, r=True, half_sq=True,
                        max_strain=True)
        assert torch.equal(o1["f_int"], ob["f_int"][:, b]) and torch.equal(o1["r"], ob["r"][:, b])
        assert float(o1["max_strain"][0]) == float(ob["max_strain"][b])
        assert abs(float(o1["half_sq"][0]) / float(ob["half_sq"][b]) - 1.0) < 1e-13
        k1 = p.tangent_matvec(dev(v[:, b].copy()), dev(E[:, b].copy()), dev(A[:, b].copy()))
        assert torch.equal(k1, kb[:, b])
    f_ref, eps_ref = O.assemble_residual(nodes, el, E[:, 0], A[:, 0], u[:, 0], dim)
    o1 = p.residual(dev(u[:, 0].copy()), dev(E[:, 0].copy()), dev(A[:, 0].copy()), max_strain=True)
    assert rel(o1["f_int"], f_ref) < TOL and rel(o1["max_strain"], np.atleast_1d(eps_ref)) < 1e-9


@pytest.mark.parametrize("B", [1, 5, 160])
def test_isolated_nodes_and_empty_meshes(B):
    """Nodes without elements get zero force and a zero diagonal block; a mesh without elements is all zeros
    (single-problem, generic and patch-staged kernels)."""
    from pinn_fem_b200 import AssemblyPlan

    nodes = np.array([[0.0, 0], [1, 0], [5, 5], [2, 1], [9, 9]])
    el = np.array([[0, 1], [1, 3], [3, 0]])
    p = AssemblyPlan(nodes, el, [0, 1], device="cuda")
    rng = np.random.default_rng(2)
    shp = (lambda n: (n,)) if B == 1 else (lambda n: (n, B))
    u, E, A = rng.normal(size=shp(10)), rng.uniform(0.5, 1.5, shp(3)), rng.uniform(0.5, 1.5, shp(3))
    f = p.internal_force(dev(u), dev(E), dev(A)).cpu().numpy()
    f_ref, _ = O.assemble_residual(nodes, el, E, A, u)
    assert rel(f, f_ref) < TOL and np.all(f[[4, 5, 8, 9]] == 0.0)
    vals = p.tangent_bsr(dev(E), dev(A)).cpu().numpy()
    diag = p._array(10)  # PF_ARR_DIAG_SLOT
    assert np.all(vals[diag[[2, 4]]] == 0.0)
    empty = AssemblyPlan(nodes, np.zeros((0, 2), dtype=int), [], device="cuda")
    z = torch.zeros(shp(0), dtype=torch.float64, device="cuda")
    out = empty.residual(dev(u), z, z, dev(np.ones(10)), 2.0, r=True, half_sq=True)
    assert float(out["f_int"].abs().max()) == 0.0
    assert np.allclose(out["r"].cpu().numpy(), -2.0) and np.allclose(out["half_sq"].cpu().numpy(), 0.5 * 10 * 4.0)


def test_benchmarked_configuration_against_the_oracle():
    """The configuration bench.py quotes -- 578 x 578 nodes, 999,941 elements, 512 problems -- itself: columns of the
    residual against the C restatement of the reference's element loop (fem/assembly.py:52-73, fem/solver.py:267-269)
    on the same inputs, full mesh, 1e-10; first/last column of a chunk, chunk boundaries and the batch's last column."""
    from oracle import c_oracle
    from pinn_fem_b200 import AssemblyPlan
    from pinn_fem_b200.meshes import lattice_truss

    nodes, el, fixed = lattice_truss(578)
    p = AssemblyPlan(nodes, el, fixed, device="cuda")
    assert p.nelem == 999_941 and p.ndof == 668_168
    B = 512
    g = torch.Generator(device="cuda").manual_seed(11)
    u = (torch.rand((p.ndof, B), generator=g, device="cuda", dtype=torch.float64) - 0.5) * 2e-3
    E = torch.rand((p.nelem, B), generator=g, device="cuda", dtype=torch.float64) + 0.5
    A = torch.rand((p.nelem, B), generator=g, device="cuda", dtype=torch.float64) + 0.5
    fx = torch.randn(p.ndof, generator=g, device="cuda", dtype=torch.float64) * 1e-3
    out = p.residual(u, E, A, fx, 0.8, f_int=True, r=True, half_sq=True)
    cols = [0, 15, 16, 255, 256, 497, 511]
    uh, Eh, Ah = (np.ascontiguousarray(t[:, cols].cpu().numpy()) for t in (u, E, A))
    f_ref, r_ref = c_oracle.residual(nodes, el, Eh, Ah, uh, fx.cpu().numpy(), 0.8, fixed, want_f=True, want_r=True)
    assert rel(out["f_int"][:, cols], f_ref) < TOL
    assert rel(out["r"][:, cols], r_ref) < TOL
    assert rel(out["half_sq"][cols], 0.5 * np.sum(r_ref * r_ref, axis=0)) < 1e-11
    # and the C restatement itself against the numpy restatement on one column (oracle vs oracle, same algorithm)
    f_np, _ = O.assemble_residual(nodes, el, Eh[:, :1], Ah[:, :1], uh[:, :1])
    assert rel(f_ref[:, :1], f_np) < 1e-12

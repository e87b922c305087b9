"""Pins the CPU oracle (oracle/pinnfem_oracle.py) to the golden vectors that
tests/golden/make_golden.py captured from the real reference, and to the
known-answer cases of the reference's own test_torch_element.py."""
import json

import numpy as np
import pytest

from oracle import pinnfem_oracle as O


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


# --- element known-answer cases ------------------------------------------------


@pytest.fixture(scope="module")
def kat(golden_dir):
    with open(golden_dir / "elements_kat.json") as f:
        return json.load(f)


def test_index_helpers(kat):
    assert O.element_dofs(3, 7).tolist() == kat["element_dofs_3_7"] == [6, 7, 14, 15]
    free, fixed = O.free_and_fixed_dofs(12, kat["free_fixed_12"]["input"])
    assert free.tolist() == kat["free_fixed_12"]["free"]
    assert fixed.tolist() == kat["free_fixed_12"]["fixed"]
    assert free.dtype == np.int64 and fixed.dtype == np.int64


def test_t1_horizontal_bar(kat):
    # reference test_torch_element.py:21-59 (intent): unit bar, u_j = (1, 0)
    ke, fe, eps = O.truss2d_linear_element([0, 0], [1, 0], [0, 0], [1, 0], 1.0, 1.0)
    assert np.array_equal(ke, np.array(kat["T1_horizontal"]["ke"]))
    assert np.array_equal(fe, np.array(kat["T1_horizontal"]["fe"]))
    assert fe.tolist() == [-1.0, 0.0, 1.0, 0.0] and eps == 1.0


def test_t3_diagonal_bar(kat):
    # reference test_torch_element.py:196-244: axial force 100/sqrt2*0.1 = 7.0711
    d = 0.1 / np.sqrt(2)
    ke, fe, eps = O.truss2d_linear_element([0, 0], [1, 1], [0, 0], [d, d], 100.0, 1.0)
    assert rel(fe, kat["T3_diagonal"]["fe"]) < 1e-15
    assert rel(ke, kat["T3_diagonal"]["ke"]) < 1e-15
    assert np.allclose(fe, [-5, -5, 5, 5], rtol=0, atol=1e-14)
    assert abs(eps - 0.07071067811865474) < 1e-16
    ke32, fe32, _ = O.truss2d_linear_element([0, 0], [1, 1], [0, 0], np.float32([d, d]), 100.0, 1.0,
                                              dtype=np.float32)
    assert rel(fe32, kat["T3_diagonal"]["fe_torch32"]) < 5e-7
    assert rel(ke32, kat["T3_diagonal"]["ke_torch32"]) < 5e-7


def test_green_lagrange_element(kat):
    ke, fe, e = O.truss2d_element_state([0, 0], [1, 1], [0, 0], [0.1, 0.05], 100.0, 1.0)
    assert abs(e - 0.07812499999999999) < 1e-15 and e == kat["GL"]["strain"]
    assert rel(fe, kat["GL"]["fe"]) < 1e-15 and rel(ke, kat["GL"]["ke"]) < 1e-15
    assert abs(np.trace(ke) - 166.97111297940154) < 1e-11
    assert np.allclose(ke, ke.T, rtol=0, atol=1e-12)


def test_1d_element(kat):
    ke, fe, eps = O.truss1d_linear_element(0.5, 2.0, 0.01, -0.02, 3.0, 0.25)
    assert rel(ke, kat["T1D"]["ke"]) < 1e-16 and rel(fe, kat["T1D"]["fe"]) < 1e-16
    assert eps == kat["T1D"]["strain"]


def test_random_elements(kat):
    for c in kat["random"]:
        ke, fe, eps = O.truss2d_linear_element(c["xi"], c["xj"], c["ui"], c["uj"], c["E"], c["A"])
        assert rel(ke, c["lin"]["ke"]) < 1e-15 and rel(fe, c["lin"]["fe"]) < 1e-14
        assert abs(eps - c["lin"]["strain"]) < 1e-16
        ke, fe, eps = O.truss2d_element_state(c["xi"], c["xj"], c["ui"], c["uj"], c["E"], c["A"])
        assert rel(ke, c["gl"]["ke"]) < 1e-14 and rel(fe, c["gl"]["fe"]) < 1e-13


def test_t2_three_bars():
    # reference test_torch_element.py:79-187 (the one test that passes today)
    nodes = np.array([[0.0, 0], [1, 0], [2, 0], [3, 0]])
    el = np.array([[0, 1], [1, 2], [2, 3]])
    f_ext = np.array([0, 0, 0, 0, 0, 0, 1.0, 0])
    free = [2, 4, 6]
    f, _ = O.assemble_residual(nodes, el, 1.0, 1.0, np.zeros(8))
    assert np.array_equal(f, np.zeros(8))
    r = f[free] - f_ext[free]
    assert r.tolist() == [0, 0, -1]
    # d mean(R^2) / du = K^T (2/3 R) on the free rows
    g_f = np.zeros(8)
    g_f[free] = 2.0 / 3.0 * r
    g_u = O.tangent_matvec(nodes, el, 1.0, 1.0, g_f)
    assert np.allclose(g_u, [0, 0, 0, 0, 2 / 3, 0, -2 / 3, 0], rtol=0, atol=1e-15)
    f, _ = O.assemble_residual(nodes, el, 1.0, 1.0, np.array([0, 0, 1.0, 0, 2, 0, 3, 0]))
    assert f.tolist() == [-1, 0, 0, 0, 0, 0, 1, 0]


# --- assembly vs the reference's assemble_system (fp64) ---------------------------


def _case(g, name):
    return {k.split(".", 1)[1]: g[k] for k in g.files if k.startswith(name + ".")}


def test_assembly_cases_present(assembly_golden):
    assert set(assembly_golden["cases"].tolist()) == {"ex1", "fem2d_like", "lattice8", "lattice5x3_perturbed",
                                                      "bar1d"}


@pytest.mark.parametrize("name", ["ex1", "fem2d_like", "lattice8", "lattice5x3_perturbed", "bar1d"])
def test_assemble_system_matches_reference(assembly_golden, name):
    c = _case(assembly_golden, name)
    dim = int(c["dim"])
    K, f, eps = O.assemble_system_loop(c["nodes"], c["elements"], c["E"], c["A"], c["u"], dim)
    assert rel(K, c["K"]) < 1e-15 and rel(f, c["f_int"]) < 1e-13
    assert abs(eps - float(c["max_strain"])) <= 1e-16 * max(1.0, eps) + 1e-18
    Ks, fs, epss = O.assemble_system_loop(c["nodes"], c["elements"], 2.5, 0.4, c["u"], dim)
    assert rel(Ks, c["K_scalar"]) < 1e-15 and rel(fs, c["f_scalar"]) < 1e-13
    # vectorised forms used to check the CUDA kernels at size
    f2, eps2 = O.assemble_residual(c["nodes"], c["elements"], c["E"], c["A"], c["u"], dim)
    assert rel(f2, c["f_int"]) < 1e-12 and abs(eps2 - float(c["max_strain"])) < 1e-15
    rowptr, colind, vals = O.assemble_tangent_bsr(c["nodes"], c["elements"], c["E"], c["A"], c["u"], dim)
    assert rel(O.bsr_to_dense(rowptr, colind, vals, dim), c["K"]) < 1e-15
    free, fixed = O.free_and_fixed_dofs(len(c["u"]), c["fixed_in"])
    assert np.array_equal(free, c["free"]) and np.array_equal(fixed, c["fixed"])


@pytest.mark.parametrize("name", ["ex1", "fem2d_like", "lattice8", "lattice5x3_perturbed"])
def test_green_lagrange_assembly(assembly_golden, name):
    c = _case(assembly_golden, name)
    K, f, _ = O.assemble_system_loop(c["nodes"], c["elements"], c["E"], c["A"], c["u_gl"], 2, O.GREEN_LAGRANGE)
    assert rel(K, c["K_gl"]) < 1e-14 and rel(f, c["f_gl"]) < 1e-12
    f2, _ = O.assemble_residual(c["nodes"], c["elements"], c["E"], c["A"], c["u_gl"], 2, O.GREEN_LAGRANGE)
    assert rel(f2, c["f_gl"]) < 1e-12
    rowptr, colind, vals = O.assemble_tangent_bsr(c["nodes"], c["elements"], c["E"], c["A"], c["u_gl"], 2,
                                                  O.GREEN_LAGRANGE)
    assert rel(O.bsr_to_dense(rowptr, colind, vals), c["K_gl"]) < 1e-14
    v = np.cos(np.arange(len(c["u_gl"])))
    kv = O.tangent_matvec(c["nodes"], c["elements"], c["E"], c["A"], v, c["u_gl"], 2, O.GREEN_LAGRANGE)
    assert rel(kv, c["K_gl"] @ v) < 1e-13


@pytest.mark.parametrize("name", ["ex1", "fem2d_like", "lattice8", "bar1d"])
def test_structure_is_bit_exact(assembly_golden, name):
    c = _case(assembly_golden, name)
    dim = int(c["dim"])
    nnode = c["nodes"].shape[0]
    pat = O.structural_pattern(nnode, c["elements"], dim)
    assert np.all(pat[np.nonzero(c["K"])])  # numeric nonzeros are inside the structural pattern
    rowptr, colind, slots = O.bsr_pattern(nnode, c["elements"])
    blk = np.zeros((nnode, nnode), dtype=bool)
    for n in range(nnode):
        blk[n, colind[rowptr[n]:rowptr[n + 1]]] = True
    assert np.array_equal(np.kron(blk, np.ones((dim, dim), dtype=bool)), pat)
    ptr, inc_elem, inc_end = O.node_incidence(nnode, c["elements"])
    for n in range(nnode):
        es = inc_elem[ptr[n]:ptr[n + 1]]
        assert np.all(np.diff(es) >= 0)
        for e, end in zip(es, inc_end[ptr[n]:ptr[n + 1]]):
            assert c["elements"][e][end] == n
    if name == "fem2d_like":
        assert nnode == 82 and len(c["elements"]) == 162 and np.count_nonzero(c["K_scalar"]) == 884


def test_batched_layout_matches_unbatched(assembly_golden):
    c = _case(assembly_golden, "lattice8")
    rng = np.random.default_rng(0)
    B = 5
    u = rng.uniform(-1e-3, 1e-3, size=(len(c["u"]), B))
    E = rng.uniform(0.5, 1.5, size=(len(c["E"]), B))
    A = rng.uniform(0.5, 1.5, size=(len(c["E"]), B))
    for kind in (O.LINEAR, O.GREEN_LAGRANGE):
        f, _ = O.assemble_residual(c["nodes"], c["elements"], E, A, u, 2, kind)
        for b in range(B):
            fb, _ = O.assemble_residual(c["nodes"], c["elements"], E[:, b], A[:, b], u[:, b], 2, kind)
            assert np.array_equal(f[:, b], fb)


# --- material networks vs the reference's torch path (fp32) ----------------------


SPECS = {"young": O.NetSpec(3, 2, 20), "area": O.NetSpec(3, 2, 15), "density": O.NetSpec(3, 2, 10)}


@pytest.fixture(scope="module")
def torch_golden(golden_dir):
    return np.load(golden_dir / "assembly_torch_f32.npz")


def test_param_counts():
    # true counts 521 / 316 / 161
    assert [SPECS[n].n_params for n in ("young", "area", "density")] == [521, 316, 161]


def test_nnproperty_value(torch_golden):
    g = torch_golden
    X = np.array([[float(g["lam"]), x, 0.0] for x in (0.5, 1.5, 2.5)])
    for n in SPECS:
        v = O.mlp_forward(SPECS[n], g[f"theta_{n}"], X, 1.0)
        assert rel(v, g[f"value_{n}"]) < 2e-6  # reference evaluates in fp32


def test_assemble_system_torch_and_autograd(torch_golden, example_inputs):
    g = torch_golden
    nodes = np.array([[0.0, 0], [1, 0], [2, 0], [3, 0]])
    el = np.array([[0, 1], [1, 2], [2, 3]])
    lam = float(g["lam"])
    X = O.nn_inputs(nodes, el, lam)
    assert np.array_equal(X, [[lam, 0.5, 0], [lam, 1.5, 0], [lam, 2.5, 0]])  # [load_factor, x, y] (D6)
    E = O.mlp_forward(SPECS["young"], g["theta_young"], X)
    A = O.mlp_forward(SPECS["area"], g["theta_area"], X)
    K, f, _ = O.assemble_system_loop(nodes, el, E, A, g["u"])
    assert rel(K, g["K"]) < 2e-6 and rel(f, g["f_int"]) < 5e-6
    # L = sum(w * f_int): dL/du = K^T w, dL/dtheta through the material VJP + MLP backward
    g_u = O.tangent_matvec(nodes, el, E, A, g["w"])
    assert rel(g_u, g["g_u"]) < 5e-6
    gE, gA = O.material_vjp(nodes, el, E, A, g["u"], g["w"])
    gt_y = O.mlp_backward(SPECS["young"], g["theta_young"], X, gE)
    gt_a = O.mlp_backward(SPECS["area"], g["theta_area"], X, gA)
    assert rel(gt_y, g["g_theta_young"]) < 2e-5 and rel(gt_a, g["g_theta_area"]) < 2e-5
    assert not np.any(g["g_theta_density"])  # rho never enters the physics (A.4)


def test_mlp_backward_against_autograd_fp64():
    """Independent fp64 check of the closed-form reverse pass."""
    import torch

    spec = O.NetSpec(3, 3, 7)
    rng = np.random.default_rng(1)
    theta = rng.normal(scale=0.5, size=spec.n_params)
    X = rng.normal(size=(9, 3))
    g_out = rng.normal(size=9)
    t = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    a = torch.tensor(X)
    off = 0
    shapes = spec.layer_shapes
    for li, (o, i) in enumerate(shapes):
        W = t[off:off + o * i].reshape(o, i)
        off += o * i
        b = t[off:off + o]
        off += o
        a = a @ W.T + b
        if li < len(shapes) - 1:
            a = torch.tanh(a)
    y = torch.nn.functional.softplus(a[:, 0]) * 2.5
    (y * torch.tensor(g_out)).sum().backward()
    assert rel(O.mlp_forward(spec, theta, X, 2.5), y.detach().numpy()) < 1e-14
    assert rel(O.mlp_backward(spec, theta, X, g_out, 2.5), t.grad.numpy()) < 1e-12


# --- Newton-Raphson drivers -----------------------------------------------------


def test_example1_nr(golden_dir, example_inputs):
    with open(golden_dir / "solver_runs.json") as f:
        runs = json.load(f)
    mesh = O.Mesh(np.array([[0.0, 0], [1, 0], [2, 0], [3, 0]]), np.array([[0, 1], [1, 2], [2, 3]]),
                  np.array(example_inputs["example1"]["loads"], dtype=float), np.array([0, 1, 3, 5, 7]))
    u, reac, ok, hist = O.solve_nr(mesh, 1.0, 1.0, 1.0, 50, 1e-6)
    out = runs["example1"]["output"]
    assert ok and np.allclose(u, out["displacements"], rtol=0, atol=1e-15)
    assert np.allclose(reac, out["reactions"], rtol=0, atol=1e-15)
    assert hist[-1]["iterations"] == out["history"][-1]["iterations"] == 2.0
    assert out["iterations"] == 1  # len(history) of the last increment (A.6)


def test_fem2d_like_incremental_newton(golden_dir):
    g = np.load(golden_dir / "fem2d_like_nr.npz")
    mesh = O.Mesh(g["nodes"], g["elements"], g["loads"], g["fixed"])
    u, reac, ok, hist = O.solve_incremental_newton(mesh, float(g["young"]), float(g["area"]), 10, 120, 1e-5)
    assert ok == bool(g["converged"])
    assert rel(u, g["u"]) < 1e-11 and rel(reac, g["reactions"]) < 1e-9
    assert [h["iterations"] for h in hist] == g["iterations"].tolist()
    assert abs(np.max(np.linalg.norm(u.reshape(-1, 2), axis=1)) - 6.344508621253013e-4) < 1e-14


# --- gradient descent -----------------------------------------------------------


def _ex4_material(theta0, dtype=np.float64):
    return O.MaterialNets(*[(SPECS[n], np.array(theta0[n], dtype=dtype), 1.0) for n in ("young", "area", "density")])


def test_gd_first_increment_tracks_reference(golden_dir, example_inputs):
    """fp64 restatement vs the reference's fp32 run, same theta_0: the first
    iterations agree to fp32 round-off; the increment converges to the same u."""
    with open(golden_dir / "solver_runs.json") as f:
        theta0 = json.load(f)["example4-P"]["theta0"]
    with open(golden_dir / "gd_trace_example4P.json") as f:
        trace = json.load(f)
    d = example_inputs["example4-P"]
    mesh = O.Mesh(np.array([[0.0, 0], [1, 0], [2, 0], [3, 0]]), np.array([[0, 1], [1, 2], [2, 3]]),
                  np.array(d["loads"], dtype=float), np.array([0, 1, 3, 5, 7]))
    mat = _ex4_material(theta0)
    pc = d["pinn_config"]
    # gd_trace was produced by solve_gd with preconditioning=True -> two phases
    u, reac, ok, hist = O.solve_gd_preconditioned(
        mesh, mat, pc["max_iterations"], pc["tolerance"], lr_u=pc["learning_rate_u"],
        lr_theta=pc["learning_rate_theta"], alpha_p=pc["alpha_physics"], alpha_d=pc["alpha_data"],
        meas_dofs=np.array([2, 3, 4, 5, 6, 7]), meas_vals=np.array([1.0, 0, 2, 0, 3, 0]), lam=0.1)
    ref = trace[0]
    for h, hr in zip(hist[:12], ref["history_head"]):
        for key in ("loss_total", "loss_physics", "loss_data", "u_norm", "residual_norm", "theta_norm"):
            assert abs(h[key] - hr[key]) <= 2e-4 * max(abs(hr[key]), 1e-3), (key, h, hr)
    assert ok == ref["converged"]
    assert np.allclose(u, ref["u"], rtol=0, atol=5e-4)
    # iteration counts differ by a few percent between fp32 and fp64 trajectories
    assert abs(len(hist) - ref["n_history"]) <= 0.15 * ref["n_history"]


# --- Gauss-Newton / LM ----------------------------------------------------------


def test_jacobian_blocks_and_lm_step(golden_dir, example_inputs):
    g = np.load(golden_dir / "gauss_newton_ex10.npz")
    d = example_inputs["example10"]
    mesh = O.Mesh(np.array([[0.0, 0], [1, 0], [2, 0], [3, 0]]), np.array([[0, 1], [1, 2], [2, 3]]),
                  np.array(d["loads"], dtype=float), np.array([0, 1, 3, 5, 7]))
    mat = O.MaterialNets(*[(SPECS[n], g[f"theta0_{n}"].copy(), 1.0) for n in ("young", "area", "density")])
    md = np.array([2, 4, 6])
    j_uu, j_ut, r_p, j_du = O.jacobian_blocks(mesh, mat, g["u"], mesh.loads, md)
    assert j_ut.shape == (3, 998)
    assert rel(j_uu, g["j_uu"]) < 2e-6 and rel(r_p, g["r_physics"]) < 5e-6
    assert rel(j_ut, g["j_utheta"]) < 2e-5
    assert np.array_equal(j_du, g["j_data_u"]) and np.array_equal(j_du, -np.eye(3))
    J, R = O.gauss_newton_system(j_uu, j_ut, r_p, j_du, np.array([1.0, 2, 3]) - g["u"][md])
    assert J.shape == (6, 1001) and rel(J, g["J"]) < 2e-5 and rel(R, g["R"]) < 5e-6
    dx, jtj, jtr, damping = O.lm_step(J, R)
    assert rel(np.diag(jtj), g["jtj_diag"]) < 5e-5 and rel(jtj[:8], g["jtj_rows8"]) < 5e-5
    assert rel(jtr, g["jtr"]) < 5e-5
    assert abs(damping - float(g["damping"])) < 5e-5 * damping
    # the fp32 solve of a 1001x1001 system with damping 1e-6*tr/n is itself only
    # accurate to ~1e-2; compare the well-conditioned part (the u-step) loosely
    # and require the fp64 step to solve the reference's own system well.
    assert np.linalg.norm(dx[:3] - g["dx"][:3]) < 0.05 * np.linalg.norm(g["dx"][:3])


def test_gauss_newton_run_history(golden_dir, example_inputs):
    g = np.load(golden_dir / "gauss_newton_ex10.npz")
    d = example_inputs["example10"]
    mesh = O.Mesh(np.array([[0.0, 0], [1, 0], [2, 0], [3, 0]]), np.array([[0, 1], [1, 2], [2, 3]]),
                  np.array(d["loads"], dtype=float), np.array([0, 1, 3, 5, 7]))
    mat = O.MaterialNets(*[(SPECS[n], g[f"theta0_{n}"].copy(), 1.0) for n in ("young", "area", "density")])
    u, ok, hist = O.solve_pinn_newton_raphson(mesh, mat, mesh.loads, np.array([1.0, 2, 3]), [2, 4, 6],
                                              max_iterations=8)
    # first iteration starts from identical state: residual norm is exact
    assert abs(hist[0]["r_total"] - g["run_r_total"][0]) < 1e-5
    assert hist[0]["step_size"] == g["run_step"][0]


def test_lattice_generator_counts():
    nodes, el, fixed = O.lattice_truss(8)
    assert nodes.shape == (64, 2) and len(el) == 7 * 8 * 2 + 49 and len(fixed) == 16
    # C5 sizes from the closed form, without building the mesh
    nx = 578
    assert nx * nx == 334084 and 2 * (nx - 1) * nx + (nx - 1) ** 2 == 999941


def test_c_oracle_matches_numpy_oracle(assembly_golden):
    """oracle/pf_oracle.c (the multi-threaded CPU baseline) against the golden
    vectors and the NumPy restatement."""
    from oracle import c_oracle

    for name in ("fem2d_like", "lattice5x3_perturbed", "bar1d"):
        c = _case(assembly_golden, name)
        f, _ = c_oracle.residual(c["nodes"], c["elements"], c["E"], c["A"], c["u"])
        assert rel(f, c["f_int"]) < 1e-13
    c = _case(assembly_golden, "lattice8")
    rng = np.random.default_rng(3)
    B = 19
    u = rng.uniform(-1e-3, 1e-3, size=(len(c["u"]), B))
    E = rng.uniform(0.5, 1.5, size=(len(c["E"]), B))
    A = rng.uniform(0.5, 1.5, size=(len(c["E"]), B))
    fx = rng.normal(size=len(c["u"]))
    for kind in (O.LINEAR, O.GREEN_LAGRANGE):
        f, r = c_oracle.residual(c["nodes"], c["elements"], E, A, u, fx, 0.3, c["fixed_in"], kind, want_r=True)
        f_ref, _ = O.assemble_residual(c["nodes"], c["elements"], E, A, u, 2, kind)
        assert rel(f, f_ref) < 1e-13
        r_ref = np.zeros_like(f_ref)
        r_ref[c["free"]] = f_ref[c["free"]] - 0.3 * fx[c["free"]][:, None]
        assert rel(r, r_ref) < 1e-13


def test_green_lagrange_element_is_kept_verbatim_including_its_sign():
    """fem/element.py:105-133 returns fe = +EA/l0*e*d on node i (towards node j for a stretched bar) while its
    ke is positive definite: fe has the opposite sign of dU/du, so Newton on this element cannot converge --
    presumably why no reference solver calls it.  The port keeps it verbatim (element-level
    goldens above); this test pins the behaviour so that nobody "fixes" one side only."""
    xi, xj = np.array([0.0, 0.0]), np.array([1.0, 0.0])
    ui, uj = np.zeros(2), np.array([1e-3, 0.0])  # stretched
    ke_l, fe_l, _ = O.truss2d_linear_element(xi, xj, ui, uj, 1.0, 1.0)
    ke_g, fe_g, e_g = O.truss2d_element_state(xi, xj, ui, uj, 1.0, 1.0)
    assert fe_l[0] < 0 < fe_l[2]          # linear: internal force resists the stretch
    assert fe_g[0] > 0 > fe_g[2]          # Green-Lagrange (verbatim): opposite sign
    assert abs(abs(fe_g[0]) - abs(fe_l[0])) < 2e-6 and e_g > 0
    assert np.all(np.linalg.eigvalsh(ke_g) > -1e-12)
    # finite-difference tangent of the verbatim fe is -ke (to first order in the strain)
    h = 1e-7
    col = (O.truss2d_element_state(xi, xj, ui + np.array([h, 0]), uj, 1.0, 1.0)[1] - fe_g) / h
    assert np.allclose(col, -ke_g[:, 0], atol=5e-3)

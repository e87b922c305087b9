"""GPU end-to-end tests through the reference's own entry points (the drop-in boundary):
generic.py, api_fem_solver.py, fem.solver.solve*, fem.core, fem.nn_solver, assemble_system*."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import pinnfem_oracle as O

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
SPECS = {"young": O.NetSpec(3, 2, 20), "area": O.NetSpec(3, 2, 15), "density": O.NetSpec(3, 2, 10)}


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def _write(tmp_path, data, name="problem.json"):
    p = tmp_path / name
    p.write_text(json.dumps(data))
    return p


@pytest.fixture(scope="module")
def runs(golden_dir):
    with open(golden_dir / "solver_runs.json") as f:
        return json.load(f)


def _run_generic(tmp_path, data, seed=0, name="problem.json"):
    """generic.main() in-process (so the seed applies), output read back from the default file."""
    from pinn_fem_b200.examples.json import generic

    p = _write(tmp_path, data, name)
    argv = sys.argv
    sys.argv = ["generic.py", str(p)]
    torch.manual_seed(seed)
    try:
        generic.main()
    finally:
        sys.argv = argv
    out = p.with_name(p.stem + ".res.json")
    assert out.exists() and p.with_name(p.stem + ".log").exists()
    with open(out) as f:
        return json.load(f)


@pytest.mark.parametrize("name", ["example1", "example1-1", "example8", "example5"])
def test_generic_newton_examples(tmp_path, example_inputs, runs, name):
    out = _run_generic(tmp_path, example_inputs[name])
    ref = runs[name]["output"]
    assert out["success"] is True and out["converged"] is True and out["iterations"] == ref["iterations"] == 1
    assert np.allclose(out["displacements"], ref["displacements"], rtol=0, atol=1e-14)
    assert np.allclose(out["reactions"], ref["reactions"], rtol=0, atol=1e-14)
    h, hr = out["history"][-1], ref["history"][-1]
    assert set(h) == set(hr) and h["iterations"] == hr["iterations"] == 2.0 and h["converged"] == 1.0
    assert abs(h["max_strain"] - hr["max_strain"]) < 1e-14 and h["load_factor"] == 1.0


def test_generic_cli_subprocess_and_error_contract(tmp_path, example_inputs):
    """The script form: explicit output path; on failure: exit code 1 and NO output file."""
    script = ROOT / "pinn_fem_b200" / "examples" / "json" / "generic.py"
    p = _write(tmp_path, example_inputs["example1"])
    out = tmp_path / "custom_out.json"
    r = subprocess.run([sys.executable, str(script), str(p), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert json.loads(out.read_text())["displacements"] == [0.0, 0.0, 1.0, 0.0, 2.0, 0.0, 3.0, 0.0]
    bad = json.loads(json.dumps(example_inputs["example1"]))
    bad["loads"] = [0.0, 1.0]  # wrong size -> ValueError from FEMModel
    pb = _write(tmp_path, bad, "bad.json")
    r = subprocess.run([sys.executable, str(script), str(pb)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and not (tmp_path / "bad.res.json").exists()
    assert "loads size must be 8" in (tmp_path / "bad.log").read_text()


def test_api_fem_solver_contract(tmp_path, golden_dir):
    with open(golden_dir / "api_fem_solver.json") as f:
        g = json.load(f)
    script = ROOT / "pinn_fem_b200" / "api_fem_solver.py"
    pin, pout = _write(tmp_path, g["input"], "in.json"), tmp_path / "out.json"
    r = subprocess.run([sys.executable, str(script), str(pin), str(pout)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = json.loads(pout.read_text())
    ref = g["output"]
    assert set(out) == set(ref)
    assert np.allclose(out["displacements"], ref["displacements"], atol=1e-13)
    assert np.allclose(out["stresses"], ref["stresses"], atol=1e-12) and np.allclose(out["strains"], ref["strains"], atol=1e-12)
    assert out["converged"] == ref["converged"] and len(out["convergence_history"]) == 10
    for h, hr in zip(out["convergence_history"], ref["convergence_history"]):
        assert h["iterations"] == hr["iterations"] and h["load_factor"] == hr["load_factor"]
    # if/elif boundary parsing -> singular tangent -> error JSON + exit 1 (appendix A.18)
    pbad, pbo = _write(tmp_path, g["bad_input"], "bad.json"), tmp_path / "bad_out.json"
    r = subprocess.run([sys.executable, str(script), str(pbad), str(pbo)], capture_output=True, text=True, timeout=300)
    assert r.returncode == g["bad_exit"] == 1
    assert json.loads(pbo.read_text()) == g["bad_output"]


def test_fem2d_like_incremental_newton(golden_dir):
    from pinn_fem_b200.fem import FEMModel, Material, SolverConfig, solve_incremental_newton

    g = np.load(golden_dir / "fem2d_like_nr.npz")
    model = FEMModel(nodes=g["nodes"], elements=g["elements"], material=Material(young=float(g["young"]), area=float(g["area"]), density=7800.0),
                     loads=g["loads"], fixed_dofs=g["fixed"])
    res = solve_incremental_newton(model, SolverConfig(n_increments=10, max_iterations=120, tolerance=1e-5))
    assert res.converged == bool(g["converged"]) and res.displacements.shape == (82, 2)
    assert rel(res.displacements.reshape(-1), g["u"]) < 1e-8   # north star: 1e-8 on final u
    assert rel(res.reactions.reshape(-1), g["reactions"]) < 1e-8
    assert [h["iterations"] for h in res.history] == g["iterations"].tolist()
    assert abs(np.max(np.linalg.norm(res.displacements, axis=1)) - 6.344508621253013e-4) < 1e-11


def test_newton_step_cholesky_first_then_lu_decides_singular():
    """From 121 free DOFs on the Newton step tries the blocked Cholesky (K_ff is SPD on a constrained truss) and
    agrees with the pivoted LU; a zero pivot (free node without elements) falls through to the LU and raises the
    reference's RuntimeError (fem/core.py:36-37)."""
    from oracle import pinnfem_oracle as O
    from pinn_fem_b200 import AssemblyPlan, ops
    from pinn_fem_b200.fem import core

    nodes, el, fixed = O.lattice_truss(12)
    rng = np.random.default_rng(2)
    plan = AssemblyPlan(nodes, el, fixed, device="cuda")
    assert plan.nfree >= core.SPD_FROM
    E = torch.as_tensor(rng.uniform(0.5, 1.5, len(el))).cuda()
    A = torch.as_tensor(rng.uniform(0.5, 1.5, len(el))).cuda()
    rhs = torch.as_tensor(rng.normal(size=plan.ndof)).cuda()
    u = torch.zeros(plan.ndof, dtype=torch.float64, device="cuda")
    du = core.newton_step(plan, E, A, u, rhs)
    free = torch.as_tensor(plan.free_dofs.copy(), device="cuda")
    k_ff = plan.tangent_dense(E, A, u, free_only=True)
    ref = ops.solve_dense(k_ff, rhs[free].contiguous())
    assert rel(du[free], ref.cpu().numpy()) < 1e-11 and float(du.abs().sum() - du[free].abs().sum()) == 0.0
    # one more node that no element touches: zero row and column in K_ff
    nodes2 = np.vstack([nodes, [[99.0, 99.0]]])
    plan2 = AssemblyPlan(nodes2, el, fixed, device="cuda")
    with pytest.raises(RuntimeError, match="Tangent stiffness became singular"):
        core.newton_step(plan2, E, A, torch.zeros(plan2.ndof, dtype=torch.float64, device="cuda"),
                         torch.ones(plan2.ndof, dtype=torch.float64, device="cuda"))


def test_incremental_newton_through_matrix_free_cg(golden_dir, monkeypatch):
    """Large meshes solve the Newton step with the matrix-free batched CG instead of dense LU; forcing that
    path on fem2d_like must reproduce the reference's displacements to the same 1e-8."""
    from pinn_fem_b200.fem import FEMModel, Material, SolverConfig, core, solve_incremental_newton

    monkeypatch.setattr(core, "DENSE_LIMIT", 0)
    g = np.load(golden_dir / "fem2d_like_nr.npz")
    model = FEMModel(nodes=g["nodes"], elements=g["elements"], material=Material(young=float(g["young"]), area=float(g["area"]), density=7800.0),
                     loads=g["loads"], fixed_dofs=g["fixed"])
    res = solve_incremental_newton(model, SolverConfig(n_increments=10, max_iterations=120, tolerance=1e-5))
    assert res.converged == bool(g["converged"])
    assert rel(res.displacements.reshape(-1), g["u"]) < 1e-8
    assert rel(res.reactions.reshape(-1), g["reactions"]) < 1e-8
    assert [h["iterations"] for h in res.history] == g["iterations"].tolist()


def test_assemble_system_numpy_api(assembly_golden):
    from pinn_fem_b200.fem import FEMModel, Material
    from pinn_fem_b200.fem.assembly import assemble_system

    c = {k.split(".", 1)[1]: assembly_golden[k] for k in assembly_golden.files if k.startswith("fem2d_like.")}
    model = FEMModel(nodes=c["nodes"], elements=c["elements"], material=Material(young=2.5, area=0.4, density=1.0),
                     loads=np.zeros(len(c["u"])), fixed_dofs=c["fixed_in"])
    K, f, eps = assemble_system(model, c["u"])
    assert isinstance(K, np.ndarray) and K.dtype == np.float64 and K.shape == (164, 164)
    assert rel(K, c["K_scalar"]) < 1e-10 and rel(f, c["f_scalar"]) < 1e-10
    assert abs(eps - float(c["max_strain_scalar"])) < 1e-15


def test_element_functions_known_answers():
    from pinn_fem_b200.fem.element import truss1d_linear_element, truss2d_element_state, truss2d_linear_element
    from pinn_fem_b200.fem.nn_assembly import truss2d_linear_element_torch

    s = truss2d_linear_element(np.array([0.0, 0]), np.array([1.0, 0]), np.array([0.0, 0]), np.array([1.0, 0]), 1.0, 1.0)
    assert s.fe_int.tolist() == [-1.0, 0.0, 1.0, 0.0] and s.strain == 1.0
    d = 0.1 / np.sqrt(2)
    s = truss2d_linear_element(np.array([0.0, 0]), np.array([1.0, 1]), np.array([0.0, 0]), np.array([d, d]), 100.0, 1.0)
    assert np.allclose(s.fe_int, [-5, -5, 5, 5], atol=1e-13) and abs(s.strain - 0.07071067811865474) < 1e-15
    s = truss2d_element_state(np.array([0.0, 0]), np.array([1.0, 1]), np.array([0.0, 0]), np.array([0.1, 0.05]), 100.0, 1.0)
    assert abs(s.strain - 0.07812499999999999) < 1e-14 and abs(np.trace(s.ke_total) - 166.97111297940154) < 1e-9
    assert np.allclose(s.fe_int[:2], [6.076698900821891, 5.8004853144208965], rtol=1e-12)
    s = truss1d_linear_element(0.5, 2.0, 0.01, -0.02, 3.0, 0.25)
    assert np.allclose(s.ke_total, [[0.5, -0.5], [-0.5, 0.5]]) and np.allclose(s.fe_int, [0.015, -0.015]) and abs(s.strain + 0.02) < 1e-17
    with pytest.raises(ValueError, match="zero initial length"):
        truss2d_linear_element(np.array([1.0, 1]), np.array([1.0, 1]), np.zeros(2), np.zeros(2), 1.0, 1.0)
    # torch variant: differentiable in u (T1 of the reference's test script)
    u_j = torch.tensor([1.0, 0.0], dtype=torch.float64, requires_grad=True)
    ke, fe = truss2d_linear_element_torch(np.array([0.0, 0]), np.array([1.0, 0]), torch.zeros(2, dtype=torch.float64), u_j,
                                          torch.tensor(1.0), torch.tensor(1.0))
    fe.sum().backward()
    assert fe.tolist() == [-1.0, 0.0, 1.0, 0.0] and u_j.grad is not None


def _ex_model(example_inputs, golden_theta, name="example4-P", tmp=None):
    from pinn_fem_b200.examples.json import generic

    p = tmp / f"{name}.json"
    p.write_text(json.dumps(example_inputs[name]))
    torch.manual_seed(0)
    parsed = generic.parse_problem(str(p))
    return parsed


def test_assemble_system_torch_autograd(tmp_path, golden_dir, example_inputs):
    """K, f_int and autograd gradients w.r.t. u and every network parameter vs the reference (fp32
    fixture) and vs the fp64 oracle (tight)."""
    from pinn_fem_b200.fem import assemble_system_torch

    g = np.load(golden_dir / "assembly_torch_f32.npz")
    model = _ex_model(example_inputs, None, tmp=tmp_path)["model"]
    u = torch.tensor(g["u"], dtype=torch.float64, requires_grad=True)
    K, f = assemble_system_torch(model, u, load_factor=float(g["lam"]))
    assert K.shape == (8, 8) and f.shape == (8,) and f.dtype == torch.float64 and f.is_cuda
    assert rel(K, g["K"]) < 2e-6 and rel(f, g["f_int"]) < 5e-6
    w = torch.tensor(g["w"], dtype=torch.float64, device=f.device)
    (w * f).sum().backward()
    assert rel(u.grad, g["g_u"]) < 5e-6
    for name in ("young", "area", "density"):
        prop = getattr(model.material, name)
        grads = [p.grad for p in prop.net.parameters()]
        if name == "density":
            assert all(gr is None for gr in grads)  # never enters the physics
            continue
        flat = torch.cat([gr.reshape(-1) for gr in grads])
        assert rel(flat, g[f"g_theta_{name}"]) < 2e-5
    # fp64 oracle
    nodes, el = np.array([[0.0, 0], [1, 0], [2, 0], [3, 0]]), np.array([[0, 1], [1, 2], [2, 3]])
    X = O.nn_inputs(nodes, el, float(g["lam"]))
    E = O.mlp_forward(SPECS["young"], g["theta_young"], X)
    A = O.mlp_forward(SPECS["area"], g["theta_area"], X)
    gE, gA = O.material_vjp(nodes, el, E, A, g["u"], g["w"])
    gy = torch.cat([p.grad.reshape(-1) for p in model.material.young.net.parameters()])
    assert rel(gy, O.mlp_backward(SPECS["young"], g["theta_young"], X, gE)) < 1e-10
    assert rel(u.grad, O.tangent_matvec(nodes, el, E, A, g["w"])) < 1e-12
    # gradient through K as well: d sum(K * M) / d theta
    for p in model.material.get_all_torch_params():
        p.grad = None
    K2, _ = assemble_system_torch(model, u.detach(), load_factor=float(g["lam"]))
    M = torch.arange(64, dtype=torch.float64, device=K2.device).reshape(8, 8)
    (K2 * M).sum().backward()
    ke_unit = O.element_stiffness(nodes, el, 1.0, 1.0)
    dofs = O.all_element_dofs(el)
    contr = np.array([np.sum(M.cpu().numpy()[np.ix_(dofs[e], dofs[e])] * ke_unit[e]) for e in range(3)])
    gy = torch.cat([p.grad.reshape(-1) for p in model.material.young.net.parameters()])
    assert rel(gy, O.mlp_backward(SPECS["young"], g["theta_young"], X, A * contr)) < 1e-10


def test_example4p_whole_run_against_reference(tmp_path, example_inputs, runs):
    """C2: whole generic.py run of example 4-P, seed 0 on both sides."""
    out = _run_generic(tmp_path, example_inputs["example4-P"])
    ref = runs["example4-P"]["output"]
    assert out["converged"] is True and ref["converged"] is True
    # fp64 here vs fp32 in the reference: displacement targets are met to the GD tolerance on both sides
    assert np.allclose(out["displacements"], ref["displacements"], rtol=0, atol=2e-3)
    assert np.allclose(out["displacements"], [0, 0, 1, 0, 2, 0, 3, 0], rtol=0, atol=2e-3)
    assert abs(out["iterations"] - ref["n_history"]) <= max(30, 0.5 * ref["n_history"])
    assert set(out["history"][0]) == set(ref["history"][0])
    assert set(out) == {"success", "converged", "iterations", "displacements", "reactions", "history",
                        "nn_parameters", "identified_properties"}
    assert sorted(out["nn_parameters"]) == sorted(ref["nn_parameters"]) and len(out["nn_parameters"]) == 18
    ip = out["identified_properties"]
    for name in ("young", "area", "density"):
        assert ip[name]["type"] == "nn_load_dependent" and ip[name]["input_dim"] == 3
        assert set(ip[name]["load_factor_variations"]) == {"load_factor_0.2", "load_factor_0.5", "load_factor_1.0"}
    # only the product E*A is identifiable: E*A ~ 1 at full load on every element
    lv = "load_factor_1.0"
    EA = np.array(ip["young"]["load_factor_variations"][lv]["at_elements"]["values"]) * \
        np.array(ip["area"]["load_factor_variations"][lv]["at_elements"]["values"])
    assert np.allclose(EA, 1.0, atol=5e-3)


@pytest.mark.parametrize("name", ["example4-P", "example3-P"])
def test_extract_nn_properties_matches_reference_output(tmp_path, example_inputs, runs, name):
    """Post-processing (examples/json/generic.py:498-799): load the reference run's final nn_parameters into
    our networks and compare the identified property fields at nodes and centroids for every load factor
    with what the reference wrote (fp32 on its side: 1e-6)."""
    from pinn_fem_b200.examples.json import generic

    ref = runs[name]["output"]
    p = _write(tmp_path, example_inputs[name], "problem.json")
    torch.manual_seed(0)
    model = generic.parse_problem(str(p))["model"]
    params = model.material.get_all_torch_params()
    assert len(params) == len(ref["nn_parameters"])
    with torch.no_grad():
        for i, prm in enumerate(params):
            prm.copy_(torch.as_tensor(np.asarray(ref["nn_parameters"][f"param_{i}"], dtype=np.float64)).reshape(prm.shape))
    got = generic.extract_nn_properties(model)
    assert set(got) == set(ref["identified_properties"])

    def cmp(a, b, path):
        assert type(a) is type(b) or isinstance(a, (int, float)) and isinstance(b, (int, float)), path
        if isinstance(a, dict):
            assert set(a) == set(b), path
            for k in a:
                cmp(a[k], b[k], path + "/" + k)
        elif isinstance(a, list):
            assert np.allclose(np.asarray(a, dtype=float), np.asarray(b, dtype=float), rtol=2e-6, atol=1e-9), path
        elif isinstance(a, (int, float)) and not isinstance(a, bool):
            assert abs(a - b) <= 2e-6 * max(1.0, abs(b)), path
        else:
            assert a == b, path

    cmp(got, ref["identified_properties"], name)


def test_solve_gd_matches_oracle_end_state(tmp_path, example_inputs):
    """fem.solver.solve (method gd, 3 increments, preconditioning) vs the fp64 oracle driver from the
    same theta_0: final u and identified E, A agree to 1e-8 (north star)."""
    from pinn_fem_b200.fem.solver import solve

    d = json.loads(json.dumps(example_inputs["example4-P"]))
    d["solver_config"] = {"n_increments": 3}
    d["pinn_config"]["max_iterations"] = 900
    parsed = _ex_model({"example4-P": d}, None, tmp=tmp_path)
    model, cfg, md = parsed["model"], parsed["solver_config"], parsed["measured_data"]
    theta0 = {n: torch.cat([p.detach().reshape(-1) for p in getattr(model.material, n).net.parameters()]).numpy().copy()
              for n in SPECS}
    res = solve(model, cfg, md["values"], md["dofs"])
    mesh = O.Mesh(np.array([[0.0, 0], [1, 0], [2, 0], [3, 0]]), np.array([[0, 1], [1, 2], [2, 3]]),
                  np.array(d["loads"], dtype=float), np.array([0, 1, 3, 5, 7]))
    mat = O.MaterialNets(*[(SPECS[n], theta0[n].copy(), 1.0) for n in ("young", "area", "density")])
    u_ref, reac_ref, ok_ref, hist_ref = O.solve_incremental_gd(
        mesh, mat, n_increments=3, preconditioning=True, max_iterations=900, tolerance=cfg.tolerance,
        lr_u=cfg.learning_rate_u, lr_theta=cfg.learning_rate_theta, alpha_p=1.0, alpha_d=100.0,
        meas_dofs=md["dofs"], meas_vals=md["values"])
    assert res.converged == ok_ref and len(res.history) == len(hist_ref)
    assert rel(res.displacements.reshape(-1), u_ref) < 1e-8
    X = O.nn_inputs(mesh.nodes, mesh.elements, 1.0)
    from pinn_fem_b200.fem._device import get_plan, material_fields

    E, A = material_fields(model, get_plan(model), load_factor=1.0)
    assert rel(E, O.mlp_forward(SPECS["young"], mat.young[1], X)) < 1e-8
    assert rel(A, O.mlp_forward(SPECS["area"], mat.area[1], X)) < 1e-8


def test_hybrid_and_scalar_paths(tmp_path, example_inputs, runs):
    for name in ("example5-P", "example2-P"):
        out = _run_generic(tmp_path, example_inputs[name], name=f"{name}.json")
        ref = runs[name]["output"]
        assert out["converged"] == ref["converged"]
        assert np.allclose(out["displacements"], ref["displacements"], rtol=0, atol=5e-3)
        assert "nn_parameters" not in out
    out5, ref5 = _run_generic(tmp_path, example_inputs["example5-P"], name="e5.json"), runs["example5-P"]["output"]
    assert set(out5["history"][-1]) == set(ref5["history"][-1])  # GD rows + one NR row tagged with "iteration"


def test_gauss_newton_solver_vs_oracle(golden_dir, example_inputs, tmp_path, monkeypatch):
    """C4: compute_jacobian_blocks and a full solve_pinn_newton_raphson run on example 10's model."""
    from pinn_fem_b200.fem import PINNSolverConfig, solve_pinn_newton_raphson
    from pinn_fem_b200.fem.nn_solver import compute_jacobian_blocks

    g = np.load(golden_dir / "gauss_newton_ex10.npz")
    parsed = _ex_model(example_inputs, None, name="example10", tmp=tmp_path)
    model = parsed["model"]
    j_uu, j_ut, r_p, j_du, _ = compute_jacobian_blocks(model, torch.tensor(g["u"]), torch.tensor(model.loads),
                                                       model.material.get_all_torch_params(), np.array([2, 4, 6]),
                                                       np.array([2, 4, 6]))
    assert tuple(j_ut.shape) == (3, 998)
    assert rel(j_uu, g["j_uu"]) < 2e-6 and rel(j_ut, g["j_utheta"]) < 2e-5 and rel(r_p, g["r_physics"]) < 5e-6
    assert torch.equal(j_du.cpu(), -torch.eye(3, dtype=torch.float64))
    # full run vs the oracle's restatement of the same loop (same quirks), fp64 both sides
    mesh = O.Mesh(np.array([[0.0, 0], [1, 0], [2, 0], [3, 0]]), np.array([[0, 1], [1, 2], [2, 3]]),
                  np.array(example_inputs["example10"]["loads"], dtype=float), np.array([0, 1, 3, 5, 7]))
    mat = O.MaterialNets(*[(SPECS[n], g[f"theta0_{n}"].copy(), 1.0) for n in ("young", "area", "density")])
    # full runs vs the oracle's restatement of the same loop (same quirks), fp64 both sides: every history row, the
    # final u and theta.  The LM step goes through the m x m dual system here (6 residuals, 1001 unknowns); the
    # n x n formulation of the reference is run too.  Its matrix is rank 6 plus a 1e-6-relative ridge, so n x n
    # solves (ours, numpy's) agree with each other and with the dual only to what that conditioning leaves.
    theta0 = {n: g[f"theta0_{n}"].copy() for n in ("young", "area", "density")}
    keys = ("iteration", "r_physics", "r_data", "r_total", "relative_error", "step_size")
    # targets: the example's own (every step accepted at 1), one where iteration 2 is accepted after four rejected
    # trials (step 0.7^4), one where every trial of iterations 3+ fails (the loop is left with step 0.7^15)
    cases = (("base", [1.0, 2.0, 3.0], 1.0), ("backtrack", [0.1, 0.2, 0.3], 0.7 ** 4), ("failed", [1.0, -2.0, 3.0], None))
    # Measured agreement with the oracle over six iterations (scripts/gn_diff_vs_oracle.py): n x n formulation 2e-16 on
    # the first two rows, <= 1e-8 later; dual formulation 1e-9 on the first row, <= 5e-7 later (the 6 x 6 Gram matrix
    # has condition ~1e7, so the last-bit differences of J J^T between numpy and the DMMA kernel show at 1e-9 in dx).
    for path, dual, tol in (("dual", True, 2e-6), ("primal", False, 1e-7)):
        for case, targets, expect_step in cases:
            mv = np.array(targets)
            mat = O.MaterialNets(*[(SPECS[n], theta0[n].copy(), 1.0) for n in ("young", "area", "density")])
            u_ref, ok_ref, hist_ref = O.solve_pinn_newton_raphson(mesh, mat, mesh.loads, mv, [2, 4, 6], max_iterations=6,
                                                                  dual=dual)
            model = _ex_model(example_inputs, None, name="example10", tmp=tmp_path)["model"]
            monkeypatch.setenv("PF_GN_LM_PATH", path)
            res = solve_pinn_newton_raphson(model, model.loads, mv, [2, 4, 6], PINNSolverConfig(max_iterations=6))
            assert len(res.history) == len(hist_ref) == 6 and res.converged == ok_ref
            for row, ref in zip(res.history, hist_ref):
                assert set(row) == set(keys)
                assert row["step_size"] == ref["step_size"] and row["iteration"] == ref["iteration"], (path, case, row, ref)
                for k in ("r_physics", "r_data", "r_total", "relative_error"):
                    assert abs(row[k] - ref[k]) <= tol * max(abs(ref[k]), 1e-3), (path, case, row["iteration"], k, row[k], ref[k])
            assert rel(res.displacements.reshape(-1), u_ref) < tol
            th = np.concatenate([p.detach().cpu().numpy().reshape(-1) for p in model.material.get_all_torch_params()])
            assert rel(th, np.concatenate([mat.young[1], mat.area[1], mat.density[1]])) < tol
            steps_seen = [r["step_size"] for r in hist_ref]
            if case == "backtrack":
                assert abs(steps_seen[1] - expect_step) < 1e-15
            if case == "failed":
                assert abs(steps_seen[2] - 0.7 ** 15) < 1e-12
            if case == "base" and path == "dual":
                assert abs(res.history[0]["r_total"] - float(g["run_r_total"][0])) < 1e-5  # the reference's own first row
                assert res.history[0]["step_size"] == float(g["run_step"][0])


def _api_input(max_iterations):
    return {"nodes": [{"x": 0.0, "y": 0.0, "fixed": True}, {"x": 1.0, "y": 0.0, "fixed_y": True},
                      {"x": 2.0, "y": 0.0, "fixed_y": True}, {"x": 3.0, "y": 0.0, "fixed_y": True}],
            "elements": [{"nodes": [0, 1]}, {"nodes": [1, 2]}, {"nodes": [2, 3]}],
            "material": {"young": 1.0, "area": 1.0}, "loads": [0, 0, 0, 0, 0, 0, 1.0, 0],
            "measured_disp": [1.0, 2.0, 3.0], "measured_dofs": [2, 4, 6],
            "solver_config": {"max_iterations": max_iterations, "learning_rate": 0.01, "tolerance": 1e-10, "seed": 0}}


@pytest.mark.parametrize("module,iters", [("api_pinn_gradient_descent", 3000), ("api_pinn_newton_raphson", 6)])
def test_api_pinn_wrappers_run_and_keep_the_error_contract(tmp_path, module, iters):
    """The reference's two PINN API scripts fail at import; ours run the solvers that exist and keep
    the documented schema and the `{"error","type"}` + exit 1 contract of api_fem_solver.py."""
    pin, pout = tmp_path / "in.json", tmp_path / "out.json"
    pin.write_text(json.dumps(_api_input(iters)))
    r = subprocess.run([sys.executable, "-m", f"pinn_fem_b200.{module}", str(pin), str(pout)], capture_output=True,
                       text=True, timeout=600, cwd=str(Path(__file__).resolve().parent.parent))
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    out = json.loads(pout.read_text())
    assert set(out) == {"displacements", "stresses", "strains", "identified_params", "converged", "convergence_history",
                        "final_loss"}
    assert len(out["displacements"]) == 8 and len(out["stresses"]) == 3 and len(out["strains"]) == 3
    assert set(out["identified_params"]) >= {"young", "area"} and np.isfinite(out["final_loss"])
    u = np.asarray(out["displacements"])
    assert np.all(u[[0, 1, 3, 5, 7]] == 0.0) and np.all(np.isfinite(u))
    if module.endswith("gradient_descent"):
        hist = out["convergence_history"]
        assert hist[-1]["loss_total"] < hist[0]["loss_total"]      # the data misfit is being reduced
        assert np.allclose(u[[2, 4, 6]], [1.0, 2.0, 3.0], atol=0.2)
    bad = _api_input(5)
    del bad["measured_disp"]
    pin.write_text(json.dumps(bad))
    r = subprocess.run([sys.executable, "-m", f"pinn_fem_b200.{module}", str(pin), str(pout)], capture_output=True,
                       text=True, timeout=600, cwd=str(Path(__file__).resolve().parent.parent))
    assert r.returncode == 1
    err = json.loads(pout.read_text())
    assert set(err) == {"error", "type"} and err["type"] == "ValueError" and "measured_disp" in err["error"]


@pytest.mark.parametrize("name", ["example9", "example10"])
def test_examples_9_and_10_run_where_the_reference_crashes(tmp_path, example_inputs, name):
    """`full-nr` with networks: the reference crashes (example 10 at iteration 0, example 9 after 1000 fallback
    steps) and never parses these files' `measured_data` (landmine 13).  Here the run completes: with no
    measurements it is the forward problem for the networks' initial fields, and the returned displacements must
    satisfy equilibrium K(E(theta), A(theta)) u = f for the returned parameters (checked with the oracle)."""
    out = _run_generic(tmp_path, example_inputs[name])
    assert out["converged"] is True and out["success"] is True
    assert set(out) == {"success", "converged", "iterations", "displacements", "reactions", "history",
                        "nn_parameters", "identified_properties"}
    d = example_inputs[name]
    nodes = np.array([[n["x"], n["y"]] for n in d["nodes"]], dtype=float)
    el = np.array([e["nodes"] if isinstance(e, dict) else e for e in d["elements"]])
    u = np.asarray(out["displacements"])
    ip = out["identified_properties"]

    def field(prop):
        if ip[prop]["type"] == "scalar":
            return np.full(len(el), ip[prop]["value"])
        return np.asarray(ip[prop]["load_factor_variations"]["load_factor_1.0"]["at_elements"]["values"])

    E, A = field("young"), field("area")
    f, _ = O.assemble_residual(nodes, el, E, A, u)
    free = np.setdiff1d(np.arange(8), [0, 1, 3, 5, 7])
    loads = np.asarray(d["loads"], dtype=float)
    tol = 10 * float(d.get("solver_config", {}).get("tolerance", 1e-6))  # the solve stops at its own tolerance
    assert np.max(np.abs(f[free] - loads[free])) < tol * max(1.0, np.abs(loads).max())
    assert np.allclose(np.asarray(out["reactions"])[0], -loads.sum(), atol=tol)


@pytest.mark.parametrize("name", ["example3-P", "example6-P", "example7-P"])
def test_remaining_pinn_examples_against_reference_runs(tmp_path, example_inputs, runs, name):
    """Whole generic.py runs of the other PINN examples (E = NN / hybrid with 1 and 3 networks), seed 0 on both
    sides: same convergence flag, displacements within the GD tolerance of the reference's fp32 run, same schema."""
    out = _run_generic(tmp_path, example_inputs[name], name=f"{name}.json")
    ref = runs[name]["output"]
    assert out["converged"] == ref["converged"] is True
    assert np.allclose(out["displacements"], ref["displacements"], rtol=0, atol=5e-3)
    assert np.allclose(out["displacements"], [0, 0, 1, 0, 2, 0, 3, 0], rtol=0, atol=5e-3)
    assert sorted(out["nn_parameters"]) == sorted(ref["nn_parameters"])
    assert set(out["history"][0]) == set(ref["history"][0])
    for k, v in ref["nn_parameters"].items():
        assert np.asarray(out["nn_parameters"][k]).shape == np.asarray(v).shape

#!/usr/bin/env python3
"""Generate the golden fixtures in this directory FROM THE REAL REFERENCE.

Run in the build container only (``/root/reference`` does not exist on the
GPU box): ``python tests/golden/make_golden.py``.  It imports the reference's
pure-Python ``fem`` package read-only, runs it on seeded inputs and stores
inputs + outputs as small ``.npz`` / ``.json`` files.  Nothing here is imported
by the product; the tests only read the files it writes.

What is pinned:
  * element known-answer cases (reference test_torch_element.py T1/T2/T3, GL)
  * ``assemble_system`` (NumPy fp64) on example-1, fem2d_like and lattice meshes
  * ``assemble_system_torch`` (torch fp32) + autograd gradients with 3 MLPs
  * ``NNProperty.value`` on seeded SimpleNN weights
  * NR drivers: ``solve`` (method nr), ``solve_incremental_newton``, api_fem_solver
  * GD driver: whole ``generic.py`` runs of examples 2-P, 3-P, 4-P, 7-P (seed 0)
  * GN/LM: ``compute_jacobian_blocks`` + the JtJ/solve block, and a full
    ``solve_pinn_newton_raphson`` run, on example 10's model
"""

from __future__ import annotations

import contextlib
import io
import json
import os
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/FEM/python")
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(REF))
sys.path.insert(0, str(REF / "examples" / "json"))
sys.path.insert(0, str(HERE.parent.parent))

import fem  # noqa: E402  (the reference)
from fem.assembly import assemble_system  # noqa: E402
from fem.boundary import free_and_fixed_dofs  # noqa: E402
from fem.element import truss1d_linear_element, truss2d_element_state, truss2d_linear_element  # noqa: E402
from fem.geometry import element_dofs  # noqa: E402
from fem.model import FEMModel, Material  # noqa: E402
from fem.nn_assembly import assemble_system_torch, truss2d_linear_element_torch  # noqa: E402
from fem.nn_solver import PINNSolverConfig, compute_jacobian_blocks, solve_pinn_newton_raphson  # noqa: E402
from fem.properties import NNProperty, Property  # noqa: E402
import generic as ref_generic  # noqa: E402  (examples/json/generic.py)

from oracle.pinnfem_oracle import lattice_truss  # noqa: E402  (mesh generator only)


@contextlib.contextmanager
def quiet():
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield buf


class TableProperty(Property):
    """Harness-only: per-element constants handed out in call order, so the
    reference's own assembly loop (one ``value()`` call per element, in element
    order) can be driven with element-wise E, A.  (Centroids are ignored:
    fem2d_like has distinct elements with identical centroids.)"""

    def __init__(self, centroids, values):
        self.values = [float(v) for v in values]
        self.calls = 0

    def value(self, inputs=None):
        v = self.values[self.calls % len(self.values)]
        self.calls += 1
        return v


def flat_params(net):
    return np.concatenate([p.detach().numpy().reshape(-1) for p in net.parameters()]).astype(np.float64)


def example_inputs():
    """Copy the example JSONs we use to a scratch dir (generic.py writes next
    to its input) and keep their content as fixtures."""
    names = ["example1", "example1-1", "example2", "example2-P", "example3-P", "example4-P", "example5",
             "example5-P", "example6-P", "example7-P", "example8", "example9", "example10"]
    out = {}
    for n in names:
        with open(REF / "examples" / "json" / f"{n}.json") as f:
            out[n] = json.load(f)
    return out


# ----------------------------------------------------------------------------


def gen_elements():
    kat = {}
    ke, fe, eps = (lambda s: (s.ke_total, s.fe_int, s.strain))(
        truss2d_linear_element(np.array([0.0, 0.0]), np.array([1.0, 0.0]), np.array([0.0, 0.0]),
                               np.array([1.0, 0.0]), 1.0, 1.0))
    kat["T1_horizontal"] = {"ke": ke.tolist(), "fe": fe.tolist(), "strain": eps}
    d = 0.1 / np.sqrt(2)
    s = truss2d_linear_element(np.array([0.0, 0.0]), np.array([1.0, 1.0]), np.array([0.0, 0.0]),
                               np.array([d, d]), 100.0, 1.0)
    ke32, fe32 = truss2d_linear_element_torch(np.array([0.0, 0.0]), np.array([1.0, 1.0]),
                                              torch.tensor([0.0, 0.0]), torch.tensor([d, d], dtype=torch.float32),
                                              torch.tensor(100.0), torch.tensor(1.0))
    kat["T3_diagonal"] = {"ke": s.ke_total.tolist(), "fe": s.fe_int.tolist(), "strain": s.strain,
                          "fe_torch32": fe32.numpy().astype(float).tolist(),
                          "ke_torch32": ke32.numpy().astype(float).tolist()}
    s = truss2d_element_state(np.array([0.0, 0.0]), np.array([1.0, 1.0]), np.array([0.0, 0.0]),
                              np.array([0.1, 0.05]), 100.0, 1.0)
    kat["GL"] = {"ke": s.ke_total.tolist(), "fe": s.fe_int.tolist(), "strain": s.strain}
    s = truss1d_linear_element(0.5, 2.0, 0.01, -0.02, 3.0, 0.25)
    kat["T1D"] = {"ke": s.ke_total.tolist(), "fe": s.fe_int.tolist(), "strain": s.strain}
    rng = np.random.default_rng(11)
    rnd = []
    for _ in range(16):
        xi, xj = rng.normal(size=2), rng.normal(size=2)
        ui, uj = 0.05 * rng.normal(size=2), 0.05 * rng.normal(size=2)
        E, A = rng.uniform(0.5, 200.0), rng.uniform(0.1, 2.0)
        sl = truss2d_linear_element(xi, xj, ui, uj, E, A)
        sg = truss2d_element_state(xi, xj, ui, uj, E, A)
        rnd.append({"xi": xi.tolist(), "xj": xj.tolist(), "ui": ui.tolist(), "uj": uj.tolist(), "E": E, "A": A,
                    "lin": {"ke": sl.ke_total.tolist(), "fe": sl.fe_int.tolist(), "strain": sl.strain},
                    "gl": {"ke": sg.ke_total.tolist(), "fe": sg.fe_int.tolist(), "strain": sg.strain}})
    kat["random"] = rnd
    kat["element_dofs_3_7"] = element_dofs(3, 7).tolist()
    free, fixed = free_and_fixed_dofs(12, np.array([7, 0, 1, 7, 3]))
    kat["free_fixed_12"] = {"input": [7, 0, 1, 7, 3], "free": free.tolist(), "fixed": fixed.tolist()}
    with open(HERE / "elements_kat.json", "w") as f:
        json.dump(kat, f, indent=1)


def fem2d_like_mesh():
    sys.path.insert(0, str(REF / "examples"))
    import fem2d_like

    m = fem2d_like.build_model()
    return m


def gen_assembly_f64():
    """assemble_system on several meshes, scalar and per-element materials."""
    out = {}
    cases = {}
    ex1_nodes = np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0], [3.0, 0.0]])
    ex1_el = np.array([[0, 1], [1, 2], [2, 3]])
    cases["ex1"] = (ex1_nodes, ex1_el, np.array([0, 1, 3, 5, 7]), 2)
    m = fem2d_like_mesh()
    cases["fem2d_like"] = (m.nodes, m.elements, m.fixed_dofs, 2)
    n, e, fx = lattice_truss(8)
    cases["lattice8"] = (n, e, fx, 2)
    n, e, fx = lattice_truss(5, 3)
    # perturb coordinates so that no direction cosine is trivial
    n = n + 0.1 * np.random.default_rng(5).normal(size=n.shape)
    cases["lattice5x3_perturbed"] = (n, e, fx, 2)
    cases["bar1d"] = (np.array([0.0, 0.7, 1.5, 3.0, 3.2]), np.array([[0, 1], [1, 2], [2, 3], [3, 4], [0, 2]]),
                      np.array([0]), 1)
    for name, (nodes, el, fixed, dim) in cases.items():
        rng = np.random.default_rng(abs(hash(name)) % (2**31) if False else len(name) * 7919)
        nnode = nodes.shape[0]
        ndof = nnode * dim
        u = rng.uniform(-1e-3, 1e-3, size=ndof)
        E = rng.uniform(0.5, 1.5, size=len(el))
        A = rng.uniform(0.5, 1.5, size=len(el))
        cen = (nodes[el[:, 0]] + nodes[el[:, 1]]) / 2.0
        mat = Material(young=TableProperty(cen.reshape(len(el), -1), E),
                       area=TableProperty(cen.reshape(len(el), -1), A), density=1.0)
        model = FEMModel(nodes=nodes, elements=el, material=mat, loads=np.zeros(ndof), fixed_dofs=fixed,
                         dimension=dim)
        K, f, eps = assemble_system(model, u)
        mat_s = Material(young=2.5, area=0.4, density=1.0)
        model_s = FEMModel(nodes=nodes, elements=el, material=mat_s, loads=np.zeros(ndof), fixed_dofs=fixed,
                           dimension=dim)
        Ks, fs, epss = assemble_system(model_s, u)
        free, fixed_u = free_and_fixed_dofs(ndof, fixed)
        out.update({f"{name}.nodes": nodes, f"{name}.elements": el, f"{name}.fixed_in": np.asarray(fixed),
                    f"{name}.dim": np.array(dim), f"{name}.u": u, f"{name}.E": E, f"{name}.A": A,
                    f"{name}.K": K, f"{name}.f_int": f, f"{name}.max_strain": np.array(eps),
                    f"{name}.K_scalar": Ks, f"{name}.f_scalar": fs, f"{name}.max_strain_scalar": np.array(epss),
                    f"{name}.free": free, f"{name}.fixed": fixed_u})
        if dim == 2:
            # Green-Lagrange: the reference never wires it into assembly (D1);
            # drive its element routine with the same assembly loop here.
            Kg = np.zeros((ndof, ndof))
            fg = np.zeros(ndof)
            ug = 50.0 * u  # finite strains
            for k, (i, j) in enumerate(el):
                s = truss2d_element_state(nodes[i], nodes[j], ug[[2 * i, 2 * i + 1]], ug[[2 * j, 2 * j + 1]],
                                          E[k], A[k])
                dofs = element_dofs(i, j)
                Kg[np.ix_(dofs, dofs)] += s.ke_total
                fg[dofs] += s.fe_int
            out.update({f"{name}.u_gl": ug, f"{name}.K_gl": Kg, f"{name}.f_gl": fg})
    out["cases"] = np.array(list(cases.keys()))
    np.savez_compressed(HERE / "assembly_f64.npz", **out)


def build_nn_model(data, seed):
    """parse_problem needs a file; seed first so hidden layers are reproducible."""
    tmp = Path(tempfile.mkdtemp())
    p = tmp / "problem.json"
    with open(p, "w") as f:
        json.dump(data, f)
    torch.manual_seed(seed)
    with quiet():
        parsed = ref_generic.parse_problem(str(p))
    shutil.rmtree(tmp)
    return parsed


def gen_assembly_torch(inputs):
    parsed = build_nn_model(inputs["example4-P"], 0)
    model = parsed["model"]
    mat = model.material
    theta = {n: flat_params(getattr(mat, n).net) for n in ("young", "area", "density")}
    rng = np.random.default_rng(3)
    u_np = rng.uniform(-0.5, 0.5, size=model.ndof).astype(np.float32)
    u = torch.tensor(u_np, requires_grad=True)
    lam = 0.7
    K, f = assemble_system_torch(model, u, load_factor=lam)
    w = torch.tensor(rng.normal(size=model.ndof).astype(np.float32))
    loss = torch.sum(w * f)
    loss.backward()
    g_theta = {}
    for n in ("young", "area", "density"):
        gs = []
        for p in getattr(mat, n).net.parameters():
            gs.append(np.zeros(p.numel()) if p.grad is None else p.grad.numpy().reshape(-1).astype(np.float64))
        g_theta[n] = np.concatenate(gs)
    vals = {}
    with torch.no_grad():
        for n in ("young", "area", "density"):
            vals[n] = np.array([getattr(mat, n).value({"x": x, "y": 0.0, "load_factor": lam})
                                for x in (0.5, 1.5, 2.5)], dtype=np.float64).reshape(-1)
    np.savez_compressed(
        HERE / "assembly_torch_f32.npz", u=u_np.astype(np.float64), lam=np.array(lam), w=w.numpy().astype(np.float64),
        K=K.detach().numpy().astype(np.float64), f_int=f.detach().numpy().astype(np.float64),
        g_u=u.grad.numpy().astype(np.float64),
        **{f"theta_{n}": theta[n] for n in theta}, **{f"g_theta_{n}": g_theta[n] for n in g_theta},
        **{f"value_{n}": vals[n] for n in vals})


def run_generic(data, seed):
    """Whole generic.py run (parse -> solve -> output dict) with a fixed seed;
    also returns theta_0 so the build can start from identical weights."""
    parsed = build_nn_model(data, seed)
    mat = parsed["model"].material
    theta0 = {}
    for n in ("young", "area", "density"):
        prop = getattr(mat, n)
        if hasattr(prop, "net"):
            theta0[n] = flat_params(prop.net).tolist()
    with quiet():
        out = ref_generic.solve_problem(parsed)
    return theta0, out


def trim_history(hist, head=15, tail=3):
    if len(hist) <= head + tail:
        return hist
    return hist[:head] + hist[-tail:]


def gen_solver_runs(inputs):
    runs = {}
    for name in ("example1", "example1-1", "example8", "example5", "example2-P", "example3-P", "example4-P",
                 "example7-P", "example6-P", "example5-P"):
        theta0, out = run_generic(inputs[name], 0)
        out = dict(out)
        out["n_history"] = len(out["history"])
        out["history"] = trim_history(out["history"])
        out.pop("identified_properties", None) if name not in ("example4-P", "example3-P") else None
        runs[name] = {"theta0": theta0, "output": out}
        print(name, "converged", out["converged"], "iterations", out["iterations"], file=sys.stderr)
    with open(HERE / "solver_runs.json", "w") as f:
        json.dump(runs, f)

    # per-increment trace for example4-P: iterations used in each increment
    # (needed to compare control flow, not just the end state)
    parsed = build_nn_model(inputs["example4-P"], 0)
    from fem.solver import solve_gd

    model, cfg, md = parsed["model"], parsed["solver_config"], parsed["measured_data"]
    trace = []
    u_cur = None
    for iinc in range(1, 4):
        lam = iinc / cfg.n_increments
        with quiet():
            res = solve_gd(model, cfg, md["values"], md["dofs"], target_load_factor=lam,
                           u_initial=None if u_cur is None else torch.tensor(u_cur, dtype=torch.float32))
        u_cur = res.displacements.flatten()
        trace.append({"load_factor": lam, "n_history": len(res.history), "converged": bool(res.converged),
                      "u": u_cur.astype(float).tolist(), "history_head": res.history[:12],
                      "history_tail": res.history[-2:]})
    with open(HERE / "gd_trace_example4P.json", "w") as f:
        json.dump(trace, f)


def gen_api_fem_solver(inputs):
    sys.path.insert(0, str(REF))
    import api_fem_solver

    data = json.loads(json.dumps(inputs["example1"]))
    data["nodes"][0] = {"x": 0.0, "y": 0.0, "fixed": True}
    data["elements"] = [{"nodes": e} for e in data["elements"]]
    tmp = Path(tempfile.mkdtemp())
    with open(tmp / "in.json", "w") as f:
        json.dump(data, f)
    argv = sys.argv
    sys.argv = ["api_fem_solver.py", str(tmp / "in.json"), str(tmp / "out.json")]
    with quiet():
        api_fem_solver.main()
    sys.argv = argv
    with open(tmp / "out.json") as f:
        out = json.load(f)
    # the singular variant (if/elif BC parsing) -> error JSON + exit 1
    data_bad = json.loads(json.dumps(inputs["example1"]))
    data_bad["elements"] = [{"nodes": e} for e in data_bad["elements"]]
    with open(tmp / "bad.json", "w") as f:
        json.dump(data_bad, f)
    sys.argv = ["api_fem_solver.py", str(tmp / "bad.json"), str(tmp / "bad_out.json")]
    code = 0
    try:
        with quiet(), contextlib.redirect_stderr(io.StringIO()):
            api_fem_solver.main()
    except SystemExit as e:
        code = e.code
    sys.argv = argv
    with open(tmp / "bad_out.json") as f:
        bad = json.load(f)
    shutil.rmtree(tmp)
    with open(HERE / "api_fem_solver.json", "w") as f:
        json.dump({"input": data, "output": out, "bad_input": data_bad, "bad_output": bad, "bad_exit": code}, f)


def gen_fem2d_like():
    from fem import SolverConfig, solve_incremental_newton

    m = fem2d_like_mesh()
    res = solve_incremental_newton(m, SolverConfig(n_increments=10, max_iterations=120, tolerance=1e-5))
    np.savez_compressed(HERE / "fem2d_like_nr.npz", nodes=m.nodes, elements=m.elements, loads=m.loads,
                        fixed=m.fixed_dofs, young=np.array(2.1e11), area=np.array(1.0),
                        u=res.displacements.reshape(-1), reactions=res.reactions.reshape(-1),
                        converged=np.array(res.converged),
                        iterations=np.array([h["iterations"] for h in res.history]),
                        residual=np.array([h["residual"] for h in res.history]),
                        max_strain=np.array([h["max_strain"] for h in res.history]))


def gen_gauss_newton(inputs):
    """C4: example 10's model (3 NNs, seed 0); compute_jacobian_blocks and the
    JtJ/solve block at u = [0,0,.5,0,1,0,1.5,0]; then a full GN/LM run."""
    data = json.loads(json.dumps(inputs["example10"]))
    parsed = build_nn_model(data, 0)
    model = parsed["model"]
    mat = model.material
    theta0 = {n: flat_params(getattr(mat, n).net) for n in ("young", "area", "density")}
    theta_list = mat.get_all_torch_params()
    free, fixed = free_and_fixed_dofs(model.ndof, model.fixed_dofs)
    u = torch.tensor([0, 0, 0.5, 0, 1.0, 0, 1.5, 0], dtype=torch.float32)
    f_ext = torch.tensor(model.loads, dtype=torch.float32)
    meas_dofs = np.array([2, 4, 6])
    meas = torch.tensor([1.0, 2.0, 3.0])
    j_uu, j_ut, r_p, j_du, _ = compute_jacobian_blocks(model, u, f_ext, theta_list, free, meas_dofs)
    r_d = meas - u[meas_dofs]
    J = torch.cat([torch.cat([j_uu, j_ut], 1), torch.cat([j_du, torch.zeros(3, j_ut.shape[1])], 1)], 0)
    R = torch.cat([r_p, r_d])
    jtj = J.T @ J
    jtr = J.T @ R
    damping = 1e-6 * torch.trace(jtj) / jtj.shape[0]
    dx = torch.linalg.solve(jtj + damping * torch.eye(jtj.shape[0]), -jtr)
    out = {"u": u.numpy().astype(np.float64), "j_uu": j_uu.numpy().astype(np.float64),
           "j_utheta": j_ut.numpy().astype(np.float64), "r_physics": r_p.numpy().astype(np.float64),
           "j_data_u": j_du.numpy().astype(np.float64), "J": J.numpy().astype(np.float64),
           "R": R.numpy().astype(np.float64), "jtj_diag": torch.diagonal(jtj).numpy().astype(np.float64),
           "jtj_rows8": jtj[:8].numpy().astype(np.float64), "jtj_fro": np.array(float(torch.linalg.norm(jtj))),
           "jtr": jtr.numpy().astype(np.float64), "damping": np.array(float(damping)),
           "dx": dx.numpy().astype(np.float64)}
    for n in theta0:
        out[f"theta0_{n}"] = theta0[n]
    # full run from the same theta_0 (re-seed to rebuild identical nets)
    parsed = build_nn_model(data, 0)
    model = parsed["model"]
    with quiet():
        res = solve_pinn_newton_raphson(model, model.loads, np.array([1.0, 2.0, 3.0]), [2, 4, 6],
                                        PINNSolverConfig(max_iterations=8))
    out["run_u"] = res.displacements.reshape(-1).astype(np.float64)
    out["run_converged"] = np.array(res.converged)
    out["run_r_total"] = np.array([h["r_total"] for h in res.history])
    out["run_step"] = np.array([h["step_size"] for h in res.history])
    np.savez_compressed(HERE / "gauss_newton_ex10.npz", **out)


def main():
    inputs = example_inputs()
    (HERE / "inputs").mkdir(exist_ok=True)
    for n, d in inputs.items():
        with open(HERE / "inputs" / f"{n}.json", "w") as f:
            json.dump(d, f, indent=1)
    gen_elements()
    gen_assembly_f64()
    gen_assembly_torch(inputs)
    gen_api_fem_solver(inputs)
    gen_fem2d_like()
    gen_gauss_newton(inputs)
    gen_solver_runs(inputs)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

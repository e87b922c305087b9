"""Shared implementation of the two PINN API wrappers.

The reference's api_pinn_gradient_descent.py / api_pinn_newton_raphson.py import functions
that exist nowhere in its tree (ImportError), so there is no behaviour to match;
these wrappers keep the documented input/output schema and run the solvers that do exist
(fem.nn_solver_gd.solve_pinn_gradient_descent / fem.nn_solver.solve_pinn_newton_raphson) with
E and A identified as [load_factor, x, y] networks scaled by the initial guesses."""
from __future__ import annotations

import json
import sys
import traceback

import numpy as np

from .examples.json.generic import SimpleNN
from .fem._device import get_plan, material_fields, to_dev
from .fem.model import FEMModel, Material
from .fem.nn_solver import PINNSolverConfig, solve_pinn_newton_raphson
from .fem.nn_solver_gd import PINNGradientDescentConfig, solve_pinn_gradient_descent
from .fem.properties import NNProperty


def parse_input(input_data):
    nodes = np.array([[n["x"], n["y"]] for n in input_data["nodes"]])
    elements = np.array([[e["nodes"][0], e["nodes"][1]] for e in input_data["elements"]])
    material = input_data.get("material", {})
    fixed = []
    for i, node in enumerate(input_data["nodes"]):
        if node.get("fixed", False):
            fixed += [2 * i, 2 * i + 1]
        elif node.get("fixed_x", False):
            fixed.append(2 * i)
        elif node.get("fixed_y", False):
            fixed.append(2 * i + 1)
    measured_disp, measured_dofs = input_data.get("measured_disp", []), input_data.get("measured_dofs", [])
    if not measured_disp or not measured_dofs:
        raise ValueError("PINN requires measured_disp and measured_dofs for inverse problem")
    sc = input_data.get("solver_config", {})
    return {"nodes": nodes, "elements": elements, "f_ext": np.array(input_data.get("loads", [0.0] * (2 * len(nodes)))),
            "fixed_dofs": fixed, "young_init": material.get("young", 210e9), "area_init": material.get("area", 0.01),
            "u_measured": np.array(measured_disp, dtype=float), "measured_dofs": np.array(measured_dofs, dtype=int),
            "n_iterations": sc.get("max_iterations", 500), "learning_rate": sc.get("learning_rate", 0.001),
            "alpha": sc.get("alpha", 1.0), "beta": sc.get("beta", 100.0), "tolerance": sc.get("tolerance", 1e-6),
            "seed": sc.get("seed", 0)}


def solve(problem, method: str):
    import torch

    torch.manual_seed(int(problem["seed"]))
    mat = Material(young=NNProperty(SimpleNN(2, 20, 3), input_dim=3, scale=problem["young_init"]),
                   area=NNProperty(SimpleNN(2, 15, 3), input_dim=3, scale=problem["area_init"]), density=0.0)
    model = FEMModel(nodes=problem["nodes"], elements=problem["elements"], material=mat, loads=problem["f_ext"],
                     fixed_dofs=np.array(problem["fixed_dofs"], dtype=int), dimension=2)
    if method == "gd":
        cfg = PINNGradientDescentConfig(max_iterations=problem["n_iterations"], tolerance=problem["tolerance"],
                                        learning_rate_u=problem["learning_rate"], learning_rate_theta=problem["learning_rate"],
                                        alpha_physics=problem["alpha"], alpha_data=problem["beta"])
        res = solve_pinn_gradient_descent(model, problem["f_ext"], problem["u_measured"], problem["measured_dofs"], cfg)
        hist = [{k: h[k] for k in ("iteration", "loss_total", "loss_physics", "loss_data")} for h in res.history[::10]]
        final = res.history[-1]["loss_total"] if res.history else None
    else:
        cfg = PINNSolverConfig(max_iterations=problem["n_iterations"], tolerance=problem["tolerance"],
                               alpha_physics=problem["alpha"], alpha_data=problem["beta"])
        res = solve_pinn_newton_raphson(model, problem["f_ext"], problem["u_measured"], problem["measured_dofs"], cfg)
        hist = res.history
        final = res.history[-1]["r_total"] if res.history else None
    u = res.displacements.reshape(-1)
    plan = get_plan(model)
    E, A = material_fields(model, plan, load_factor=1.0)
    young, area = float(E.mean()), float(A.mean())
    eps = plan.element_strain(to_dev(u, plan.device), "engineering").cpu().numpy()
    return {"displacements": u.tolist(), "stresses": [float(young * e) for e in eps], "strains": [float(e) for e in eps],
            "identified_params": {"young": young, "area": area, "young_per_element": E.cpu().tolist(),
                                  "area_per_element": A.cpu().tolist()},
            "converged": bool(res.converged), "convergence_history": hist, "final_loss": final}


def main(method: str, name: str):
    if len(sys.argv) != 3:
        print(f"Usage: python {name} input.json output.json")
        sys.exit(1)
    input_file, output_file = sys.argv[1], sys.argv[2]
    try:
        with open(input_file, "r") as f:
            input_data = json.load(f)
        result = solve(parse_input(input_data), method)
        with open(output_file, "w") as f:
            json.dump(result, f, indent=2)
        print(f"[OK] Results written to {output_file}")
    except Exception as exc:  # noqa: BLE001
        with open(output_file, "w") as f:
            json.dump({"error": str(exc), "type": type(exc).__name__}, f, indent=2)
        print(f"[ERROR] {exc}")
        traceback.print_exc()
        sys.exit(1)

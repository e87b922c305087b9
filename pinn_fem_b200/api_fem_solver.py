#!/usr/bin/env python3
"""API wrapper of the classical FEM solver (drop-in for the reference's api_fem_solver.py).

    python api_fem_solver.py input.json output.json

Object-schema input only; boundary flags are read with the reference's ``if/elif`` chain
(a node with both fixed_x and fixed_y only gets x fixed, api_fem_solver.py:47-53).  On any
exception ``{"error", "type"}`` is written to the output file and the exit code is 1."""
from __future__ import annotations

import json
import sys
import traceback
from pathlib import Path

import numpy as np

_ROOT = Path(__file__).resolve().parents[1]
if str(_ROOT) not in sys.path:
    sys.path.insert(0, str(_ROOT))

from pinn_fem_b200.fem._device import get_plan, to_dev  # noqa: E402
from pinn_fem_b200.fem.core import solve_incremental_newton  # noqa: E402
from pinn_fem_b200.fem.model import FEMModel, Material, SolverConfig  # noqa: E402


def parse_input(input_data):
    nodes = np.array([[n["x"], n["y"]] for n in input_data["nodes"]])
    elements = np.array([[e["nodes"][0], e["nodes"][1]] for e in input_data["elements"]])
    mat = input_data.get("material", {})
    material = Material(young=mat.get("young", 210e9), area=mat.get("area", 0.01), density=mat.get("density", 7850))
    loads = np.array(input_data.get("loads", [0.0] * (2 * len(nodes))))
    fixed = []
    for i, node in enumerate(input_data["nodes"]):
        if node.get("fixed", False):
            fixed += [2 * i, 2 * i + 1]
        elif node.get("fixed_x", False):
            fixed.append(2 * i)
        elif node.get("fixed_y", False):
            fixed.append(2 * i + 1)
    sc = input_data.get("solver_config", {})
    config = SolverConfig(tolerance=sc.get("tolerance", 1e-6), max_iterations=sc.get("max_iterations", 50),
                          n_increments=sc.get("n_increments", 10))
    model = FEMModel(nodes=nodes, elements=elements, material=material, loads=loads,
                     fixed_dofs=np.array(fixed, dtype=int), dimension=2)
    return model, config


def compute_element_stresses(model: FEMModel, u: np.ndarray):
    """Engineering strain (L - L0)/L0 from the deformed length and stress E*strain per element
    (api_fem_solver.py:78-118), evaluated by the strain kernel."""
    plan = get_plan(model)
    eps = plan.element_strain(to_dev(np.asarray(u, dtype=float).reshape(-1), plan.device), "engineering").cpu().numpy()
    young = model.material.young.value()
    return [float(young * e) for e in eps], [float(e) for e in eps]


def main():
    if len(sys.argv) != 3:
        print("Usage: python api_fem_solver.py input.json output.json")
        sys.exit(1)
    input_file, output_file = sys.argv[1], sys.argv[2]
    try:
        with open(input_file, "r") as f:
            input_data = json.load(f)
        model, config = parse_input(input_data)
        print(f"Solving FEM problem: nodes {model.nnode}, elements {model.nelm}, DOFs {model.ndof}, "
              f"fixed DOFs {len(model.fixed_dofs)}, increments {config.n_increments}")
        result = solve_incremental_newton(model, config)
        u_flat = result.displacements.reshape(-1)
        stresses, strains = compute_element_stresses(model, u_flat)
        output = {"displacements": u_flat.tolist(), "stresses": stresses, "strains": strains,
                  "converged": result.converged, "convergence_history": result.history}
        with open(output_file, "w") as f:
            json.dump(output, f, indent=2)
        print(f"[OK] Results written to {output_file}  converged: {result.converged}")
    except Exception as exc:  # noqa: BLE001
        with open(output_file, "w") as f:
            json.dump({"error": str(exc), "type": type(exc).__name__}, f, indent=2)
        print(f"[ERROR] {exc}")
        traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()

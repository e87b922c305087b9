#!/usr/bin/env python3
"""Generic JSON driver (drop-in for the reference's examples/json/generic.py).

    python generic.py problem.json [output.json]

Same input schema (nodes / elements / fixed_dofs / loads / material / nn_config /
measured_displacements / solver_type / solver_config / pinn_config), same result schema
(success, converged, iterations, displacements, reactions, history, nn_parameters,
identified_properties), same side effects (``<stem>.log`` next to the input, default
output ``<stem>.res.json``, traceback + exit 1 without an output file on failure).
The numerical work runs in the CUDA library."""
from __future__ import annotations

import json
import logging
import sys
import traceback
from datetime import datetime
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

_PKG_ROOT = Path(__file__).resolve().parents[3]
if str(_PKG_ROOT) not in sys.path:
    sys.path.insert(0, str(_PKG_ROOT))

from pinn_fem_b200.fem.model import FEMModel, Material  # noqa: E402
from pinn_fem_b200.fem.properties import NNProperty  # noqa: E402
from pinn_fem_b200.fem.solver import SolverConfig, SolverResult, solve, solve_gd, solve_nr  # noqa: E402,F401

logger = None
PROPERTY_DEFAULTS = {"young": 210e9, "area": 0.01, "density": 7850}


def setup_logging(problem_file):
    """File + console logging; the log file sits next to the problem file and is overwritten."""
    global logger
    log_file = Path(problem_file).parent / f"{Path(problem_file).stem}.log"
    logging.basicConfig(level=logging.DEBUG, format="%(asctime)s [%(levelname)s] %(message)s", force=True,
                        handlers=[logging.FileHandler(log_file, mode="w", encoding="utf-8"),
                                  logging.StreamHandler(sys.stdout)])
    logger = logging.getLogger(__name__)
    for line in ("=" * 60, "PINN-FEM Generic Solver Log (pinn_fem_b200, CUDA)",
                 f"Timestamp: {datetime.now():%Y-%m-%d %H:%M:%S}", f"Problem file: {problem_file}",
                 f"Log file: {log_file}", "=" * 60):
        logger.info(line)
    return log_file


def log_print(msg="", level="info"):
    if logger is None:
        print(msg)
    else:
        getattr(logger, level if level in ("debug", "warning", "error") else "info")(msg)


class SimpleNN(nn.Module):
    """Linear(in,h) Tanh [Linear(h,h) Tanh]*(L-1) Linear(h,1); last layer weight 0.1, bias 1.0.

    Built in float32 exactly like the reference (so that a given ``torch.manual_seed`` yields the
    same initial weights), then promoted to float64, the dtype the kernels compute in."""

    def __init__(self, hidden_layers=2, neurons_per_layer=20, input_dim=1):
        super().__init__()
        mods = [nn.Linear(input_dim, neurons_per_layer), nn.Tanh()]
        for _ in range(hidden_layers - 1):
            mods += [nn.Linear(neurons_per_layer, neurons_per_layer), nn.Tanh()]
        mods.append(nn.Linear(neurons_per_layer, 1))
        self.net = nn.Sequential(*mods)
        with torch.no_grad():
            self.net[-1].bias.fill_(1.0)
            self.net[-1].weight.fill_(0.1)
        self.double()

    def forward(self, x):
        return self.net(x)


def _nodes(data):
    raw = data.get("nodes", [])
    if raw and isinstance(raw[0], list):
        arr = np.array(raw, dtype=float)
        dim = arr.shape[1]
        return (arr.flatten() if dim == 1 else arr), dim, raw
    return np.array([[n["x"], n["y"]] for n in raw]), 2, raw


def _elements(data):
    raw = data.get("elements", [])
    if raw and isinstance(raw[0], list):
        return np.array(raw)
    return np.array([[e["nodes"][0], e["nodes"][1]] for e in raw])


def _fixed_dofs(data, raw_nodes):
    listed = data.get("fixed_dofs", [])
    if listed:
        return np.array(listed, dtype=int)
    out = []
    if raw_nodes and isinstance(raw_nodes[0], dict):
        for i, node in enumerate(raw_nodes):
            if node.get("fixed", False):
                out += [2 * i, 2 * i + 1]
                continue
            if node.get("fixed_x", False):
                out.append(2 * i)
            if node.get("fixed_y", False):
                out.append(2 * i + 1)
    return np.array(out, dtype=int)


def _material(data):
    base = data.get("material", {})
    nn_config = data.get("nn_config", {})
    props = {}
    for name, default in PROPERTY_DEFAULTS.items():
        value = base.get(name, default)
        arch = nn_config.get(name, {})
        if arch.get("enabled", False):
            in_dim = arch.get("input_dim", 1)
            net = SimpleNN(hidden_layers=arch.get("hidden_layers", arch.get("hiddenLayers", 2)),
                           neurons_per_layer=arch.get("neurons_per_layer", arch.get("neuronsPerLayer", 20)),
                           input_dim=in_dim)
            props[name] = NNProperty(net=net, input_dim=in_dim, enforce_positive=True, scale=value)
            log_print(f"[DEBUG] {name}: NNProperty (scale={value}, input_dim={in_dim})", level="debug")
        else:
            props[name] = value
            log_print(f"[DEBUG] {name}: Scalar ({value})", level="debug")
    return Material(**props)


def _measurements(data, raw_nodes):
    """Only parsed for solver_type starting with "pinn" (examples 9/10 keep theirs under another key)."""
    if not data.get("solver_type", "fem").startswith("pinn"):
        return {}
    dofs, vals = [], []
    block = data.get("measured_displacements", None)
    if block:
        if "global_dof" in block and "measured_u" in block:
            dofs, vals = block["global_dof"], block["measured_u"]
        else:
            ux, uy = block.get("ux", []), block.get("uy", [])
            for k, node in enumerate(block.get("nodes", [])):
                if k < len(ux):
                    dofs.append(2 * node)
                    vals.append(ux[k])
                if k < len(uy):
                    dofs.append(2 * node + 1)
                    vals.append(uy[k])
    else:
        for i, node in enumerate(raw_nodes):
            for comp, key in enumerate(("measured_ux", "measured_uy")):
                v = node.get(key, 0) if isinstance(node, dict) else 0
                if v != 0:
                    dofs.append(2 * i + comp)
                    vals.append(v)
    return {"dofs": np.array(dofs, dtype=int), "values": np.array(vals)}


def _solver_config(data):
    sc, pc = data.get("solver_config", {}), data.get("pinn_config", {})
    solver_type = data.get("solver_type", "auto")
    method = sc.get("method") or {"fem": "nr", "pinn-gd": "gd", "pinn": "gd", "pinn-hybrid": "hybrid"}.get(solver_type, "auto")
    return SolverConfig(
        max_iterations=pc.get("max_iterations", sc.get("max_iterations", 1000)),
        tolerance=pc.get("tolerance", sc.get("tolerance", 1e-6)),
        print_every=pc.get("print_every", 10),
        n_increments=sc.get("n_increments", 10),
        min_denominator=sc.get("min_denominator", 1e-10),
        learning_rate_u=sc.get("learning_rate_u", pc.get("learning_rate_u", 1e-7)),
        learning_rate_theta=sc.get("learning_rate_theta", pc.get("learning_rate_theta", 1e-4)),
        alpha_physics=pc.get("alpha_physics", 1.0),
        alpha_data=pc.get("alpha_data", 100.0),
        preconditioning=pc.get("preconditioning", sc.get("preconditioning", False)),
        method=method,
    )


def parse_problem(problem_file):
    """problem.json -> {"model", "solver_config", "measured_data"}."""
    with open(problem_file, "r") as f:
        data = json.load(f)
    nodes, dim, raw_nodes = _nodes(data)
    n_dofs = len(raw_nodes) * dim
    elements = _elements(data)
    fixed = _fixed_dofs(data, raw_nodes)
    loads = np.array(data.get("loads", [0.0] * n_dofs), dtype=float)
    log_print(f"[DEBUG] Nodes: {len(raw_nodes)}, DOFs: {n_dofs}, Elements: {len(elements)}, Fixed DOFs: {fixed}",
              level="debug")
    material = _material(data)
    measured = _measurements(data, raw_nodes)
    model = FEMModel(nodes=nodes, elements=elements, material=material, loads=loads, fixed_dofs=fixed, dimension=dim)
    cfg = _solver_config(data)
    log_print(f"[DEBUG] Solver config: method={cfg.method}, tol={cfg.tolerance}, max_iter={cfg.max_iterations}",
              level="debug")
    return {"model": model, "solver_config": cfg, "measured_data": measured}


def extract_nn_properties(model, load_factors=None):
    """Identified property fields at the nodes and element centroids (batched MLP evaluations);
    load-dependent networks (input_dim > dimension) are sampled at load factors 0.2, 0.5, 1.0."""
    load_factors = [0.2, 0.5, 1.0] if load_factors is None else load_factors
    coords = np.asarray(model.nodes, dtype=float)
    pts_nodes = coords.reshape(model.nnode, -1)
    el = np.asarray(model.elements)
    centroids = ((coords[el[:, 0]] + coords[el[:, 1]]) / 2.0)
    pts_cen = centroids.reshape(len(el), -1)
    out = {}
    for name in ("young", "area", "density"):
        prop = getattr(model.material, name)
        if not hasattr(prop, "net"):
            out[name] = {"value": float(prop.value()), "type": "scalar"}
            continue

        def sample(points, lam=None):
            if lam is None:
                X = np.zeros((len(points), prop.input_dim))
                X[:, :points.shape[1]] = points
            else:  # sorted dict keys: [load_factor, x(, y)]
                X = np.column_stack([np.full(len(points), lam), points])
            return [float(v) for v in prop.evaluate(X).cpu().numpy()]

        def block(lam=None):
            return {"at_nodes": {"coords": coords.tolist(), "values": sample(pts_nodes, lam)},
                    "at_elements": {"centroids": [c.tolist() for c in centroids], "values": sample(pts_cen, lam)}}

        if prop.input_dim > model.dimension:
            out[name] = {"load_factor_variations": {f"load_factor_{lf:.1f}": block(lf) for lf in load_factors},
                         "type": "nn_load_dependent", "input_dim": prop.input_dim}
        else:
            out[name] = dict(block(), type="nn", input_dim=prop.input_dim)
    return out


def solve_problem(parsed_data):
    model, cfg = parsed_data["model"], parsed_data["solver_config"]
    measured = parsed_data.get("measured_data", {})
    log_print(f"Nodes: {len(model.nodes)}  Elements: {len(model.elements)}  Fixed DOFs: {len(model.fixed_dofs)}  "
              f"Has NN: {model.material.has_trainable_params()}  Method: {cfg.method}")
    result = solve(model=model, config=cfg, measured_disp=measured.get("values", None),
                   measured_dofs=measured.get("dofs", None))
    output = {
        "success": result.converged,
        "converged": result.converged,
        "iterations": len(result.history),
        "displacements": result.displacements.flatten().tolist(),
        "reactions": result.reactions.flatten().tolist() if result.reactions is not None else [],
        "history": result.history,
    }
    if result.nn_parameters:
        output["nn_parameters"] = {k: v.tolist() for k, v in result.nn_parameters.items()}
        output["identified_properties"] = extract_nn_properties(model)
    return output


def main():
    if len(sys.argv) < 2:
        print("Usage: python generic.py problem.json [output.json]")
        sys.exit(1)
    problem_file = sys.argv[1]
    log_file = setup_logging(problem_file)
    output_file = sys.argv[2] if len(sys.argv) > 2 else str(Path(problem_file).parent / f"{Path(problem_file).stem}.res.json")
    log_print(f"Output file will be: {output_file}")
    try:
        parsed = parse_problem(problem_file)
        log_print("[OK] Problem parsed successfully")
        result = solve_problem(parsed)
        log_print("[OK] Problem solved")
        with open(output_file, "w") as f:
            json.dump(result, f, indent=2)
        log_print(f"[OK] Results written to {output_file}")
        status = "SUCCESS" if result.get("success") else "FAILED"
        log_print(f"SOLUTION SUMMARY: status {status}, iterations {result.get('iterations')}, "
                  f"max |u| {max(abs(d) for d in result['displacements']):.6e}")
        log_print(f"Log file saved: {log_file}")
    except Exception as exc:  # noqa: BLE001 -- same contract as the reference: log, exit 1, no output file
        log_print(f"\n[ERROR] {exc}", level="error")
        log_print(traceback.format_exc(), level="error")
        sys.exit(1)


if __name__ == "__main__":
    main()

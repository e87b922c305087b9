"""CPU-side checks of the host mirror: JSON parsing, config precedence, network
construction (same theta_0 as the reference for a given seed), model validation."""
import json

import numpy as np
import pytest
import torch


def _write(tmp_path, data, name="problem.json"):
    p = tmp_path / name
    p.write_text(json.dumps(data))
    return str(p)


def test_parse_example4p(tmp_path, example_inputs, golden_dir):
    from pinn_fem_b200.examples.json import generic

    torch.manual_seed(0)
    parsed = generic.parse_problem(_write(tmp_path, example_inputs["example4-P"]))
    model, cfg, meas = parsed["model"], parsed["solver_config"], parsed["measured_data"]
    assert model.nnode == 4 and model.nelm == 3 and model.ndof == 8 and model.dimension == 2
    assert model.fixed_dofs.tolist() == [0, 1, 3, 5, 7]
    # legacy {nodes, ux, uy}: BOTH components become targets (appendix A.12)
    assert meas["dofs"].tolist() == [2, 3, 4, 5, 6, 7] and meas["values"].tolist() == [1.0, 0.0, 2.0, 0.0, 3.0, 0.0]
    assert cfg.method == "gd" and cfg.preconditioning is True and cfg.max_iterations == 5000
    assert cfg.learning_rate_u == 0.01 and cfg.learning_rate_theta == 0.0005 and cfg.n_increments == 10
    assert cfg.alpha_data == 100.0 and cfg.print_every == 100
    # same seed -> same initial weights as the reference (fixture captured from the reference itself)
    with open(golden_dir / "solver_runs.json") as f:
        theta0 = json.load(f)["example4-P"]["theta0"]
    for name, n in (("young", 521), ("area", 316), ("density", 161)):
        prop = getattr(model.material, name)
        flat = torch.cat([p.detach().reshape(-1) for p in prop.net.parameters()]).numpy()
        assert flat.size == n and flat.dtype == np.float64
        assert np.array_equal(flat, np.asarray(theta0[name]))  # float32 values, promoted exactly
        assert prop.spec.input_dim == 3 and prop.scale == 1.0


def test_parse_methods_and_precedence(tmp_path, example_inputs):
    from pinn_fem_b200.examples.json import generic

    cases = {"example1": "nr", "example7-P": "hybrid", "example10": "full-nr", "example2-P": "gd"}
    for name, method in cases.items():
        cfg = generic.parse_problem(_write(tmp_path, example_inputs[name]))["solver_config"]
        assert cfg.method == method, name
    # measurements are only parsed for solver_type "pinn*" (A.13): example 10 keeps none
    assert generic.parse_problem(_write(tmp_path, example_inputs["example10"]))["measured_data"] == {}
    # pinn_config wins for max_iterations/tolerance, solver_config wins for learning rates
    d = json.loads(json.dumps(example_inputs["example2-P"]))
    d["solver_config"] = {"max_iterations": 7, "tolerance": 1e-3, "learning_rate_u": 0.5, "n_increments": 3}
    d["pinn_config"].update({"max_iterations": 9, "learning_rate_u": 0.25})
    cfg = generic.parse_problem(_write(tmp_path, d))["solver_config"]
    assert cfg.max_iterations == 9 and cfg.learning_rate_u == 0.5 and cfg.n_increments == 3
    # global_dof measurement format and list-style nodes / 1-D problems
    d = {"nodes": [[0.0], [1.0], [2.5]], "elements": [{"nodes": [0, 1]}, {"nodes": [1, 2]}], "fixed_dofs": [0],
         "loads": [0, 0, 1.0], "material": {"young": 2.0, "area": 3.0},
         "measured_displacements": {"global_dof": [1, 2], "measured_u": [0.1, 0.2]}, "solver_type": "pinn-gd"}
    parsed = generic.parse_problem(_write(tmp_path, d))
    assert parsed["model"].dimension == 1 and parsed["model"].nodes.tolist() == [0.0, 1.0, 2.5]
    assert parsed["measured_data"]["dofs"].tolist() == [1, 2]


def test_model_validation_messages():
    from pinn_fem_b200.fem import FEMModel, Material, ScalarProperty, to_property

    mat = Material(young=1.0, area=2.0)
    assert isinstance(mat.young, ScalarProperty) and mat.density.value() == 0.0 and not mat.has_trainable_params()
    ok = dict(nodes=[[0, 0], [1, 0]], elements=[[0, 1]], material=mat, loads=[0, 0, 0, 0], fixed_dofs=[0, 1])
    FEMModel(**ok)
    with pytest.raises(ValueError, match="loads size must be 4"):
        FEMModel(**{**ok, "loads": [0, 0]})
    with pytest.raises(ValueError, match="out-of-range"):
        FEMModel(**{**ok, "fixed_dofs": [9]})
    with pytest.raises(ValueError, match="dimension must be 1 or 2"):
        FEMModel(**{**ok, "dimension": 3})
    with pytest.raises(ValueError, match=r"shape \(nnode, 2\)"):
        FEMModel(**{**ok, "nodes": [0.0, 1.0]})
    with pytest.raises(TypeError):
        to_property("steel")


def test_index_helpers_match_oracle():
    from oracle import pinnfem_oracle as O
    from pinn_fem_b200.fem.boundary import free_and_fixed_dofs
    from pinn_fem_b200.fem.geometry import element_dofs

    assert element_dofs(3, 7).tolist() == O.element_dofs(3, 7).tolist()
    f1, x1 = free_and_fixed_dofs(12, [7, 0, 1, 7, 3])
    f2, x2 = O.free_and_fixed_dofs(12, [7, 0, 1, 7, 3])
    assert np.array_equal(f1, f2) and np.array_equal(x1, x2)


def test_nnproperty_rejects_unsupported_modules():
    from pinn_fem_b200.fem import NNProperty

    bad = torch.nn.Sequential(torch.nn.Linear(2, 4), torch.nn.ReLU(), torch.nn.Linear(4, 1))
    with pytest.raises(NotImplementedError):
        NNProperty(bad, input_dim=2).spec
    good = torch.nn.Sequential(torch.nn.Linear(2, 4), torch.nn.Tanh(), torch.nn.Linear(4, 4), torch.nn.Tanh(),
                               torch.nn.Linear(4, 1))
    spec = NNProperty(good, input_dim=2).spec
    assert (spec.input_dim, spec.hidden_layers, spec.width) == (2, 2, 4)

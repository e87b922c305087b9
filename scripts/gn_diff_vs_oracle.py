import json, sys, numpy as np, torch, os
sys.path.insert(0, '.')
from oracle import pinnfem_oracle as O
from pinn_fem_b200.fem import PINNSolverConfig, solve_pinn_newton_raphson
from pinn_fem_b200.examples.json import generic
import io, contextlib
g = np.load('tests/golden/gauss_newton_ex10.npz')
SPECS = {"young": O.NetSpec(3, 2, 20), "area": O.NetSpec(3, 2, 15), "density": O.NetSpec(3, 2, 10)}
mesh = O.Mesh(np.array([[0.0, 0], [1, 0], [2, 0], [3, 0]]), np.array([[0, 1], [1, 2], [2, 3]]), np.array([0, 0, 0, 0, 0, 0, 1.0, 0]), np.array([0, 1, 3, 5, 7]))
for path, dual in (("dual", True), ("primal", False)):
    for case, targets in (("base", [1.0, 2, 3]), ("backtrack", [0.1, 0.2, 0.3]), ("failed", [1.0, -2, 3])):
        mv = np.array(targets)
        mat = O.MaterialNets(*[(SPECS[n], g[f"theta0_{n}"].copy(), 1.0) for n in ("young", "area", "density")])
        u_ref, ok, hist = O.solve_pinn_newton_raphson(mesh, mat, mesh.loads, mv, [2, 4, 6], max_iterations=6, dual=dual)
        torch.manual_seed(0)
        model = generic.parse_problem('tests/golden/inputs/example10.json')["model"]
        os.environ["PF_GN_LM_PATH"] = path
        with contextlib.redirect_stdout(io.StringIO()):
            res = solve_pinn_newton_raphson(model, model.loads, mv, [2, 4, 6], PINNSolverConfig(max_iterations=6))
        d = [max(abs(r[k] - h[k]) / max(abs(h[k]), 1e-3) for k in ("r_physics", "r_data", "r_total", "relative_error")) for r, h in zip(res.history, hist)]
        print(path, case, ["%.1e" % x for x in d], [r["step_size"] == h["step_size"] for r, h in zip(res.history, hist)],
              "u %.1e" % (np.abs(res.displacements.reshape(-1) - u_ref).max() / np.abs(u_ref).max()))

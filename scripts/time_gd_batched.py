"""PINN-GD on the C5 lattice (999,941 elements, E and A as 3-20-20-1 / 3-15-15-1 networks): ms per iteration of the
one-problem loop and of the batched loop (groups of B problems), device-timed around whole pf_gd_solve calls."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pinn_fem_b200 import AssemblyPlan, ops  # noqa: E402
from pinn_fem_b200.meshes import lattice_truss  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=578)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--batches", type=str, default="1,16,64")
a = ap.parse_args()
dev = torch.device("cuda", 0)
nodes, el, fixed = lattice_truss(a.nx)
plan = AssemblyPlan(nodes, el, fixed, device=dev)
nets = [ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15), None]
ntheta = nets[0].n_params + nets[1].n_params
fx = torch.zeros(plan.ndof, dtype=torch.float64, device=dev)
fx[-2] = 0.05
md = np.array([plan.ndof - 2, plan.ndof - 1, plan.ndof // 2])
for B in [int(x) for x in a.batches.split(",")]:
    g = torch.Generator(device=dev).manual_seed(B)
    theta = 0.3 * torch.randn((B, ntheta), generator=g, device=dev, dtype=torch.float64)
    u = (torch.rand((B, plan.ndof), generator=g, device=dev, dtype=torch.float64) - 0.5) * 2e-3
    mv = 0.01 * torch.randn((B, 3), generator=g, device=dev, dtype=torch.float64)

    def run(iters):
        return ops.gd_solve(plan, nets, [1.0, 1.0, 1.0], theta.clone(), u.clone(), fx, md, mv, max_iterations=iters,
                            tolerance=1e-30, learning_rate_u=1e-5, learning_rate_theta=1e-4, alpha_data=10.0, load_factor=1.0,
                            record_history=True)

    run(3)
    torch.cuda.synchronize()
    t = {}
    for iters in (a.iters, 2 * a.iters):  # the difference removes set-up, transposes and the reactions pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = run(iters)
        torch.cuda.synchronize()
        t[iters] = time.perf_counter() - t0
    ms_it = (t[2 * a.iters] - t[a.iters]) / a.iters * 1e3
    print(json.dumps({"nx": a.nx, "nelem": plan.nelem, "B": B, "ms_per_iteration_of_the_group": round(ms_it, 4),
                      "ms_per_problem_iteration": round(ms_it / B, 4), "problem_iterations_per_s": round(B / ms_it * 1e3, 1),
                      "whole_call_ms": round(t[2 * a.iters] * 1e3, 1), "loss_last": float(res.history[0, 2 * a.iters - 1, 1])}),
          flush=True)

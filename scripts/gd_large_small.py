"""The large-mesh PINN-GD loop (pf_gd_large.cu) on a lattice of --nx nodes per side, one GPU: ms per iteration.
Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel launch list."""
import argparse, json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pinn_fem_b200 import AssemblyPlan
from bench_gd import gd_large_mesh_iterations_per_second
from pinn_fem_b200.meshes import lattice_truss

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=205)
ap.add_argument("--iters", type=int, default=100)
a = ap.parse_args()
dev = torch.device("cuda", 0)
plan = AssemblyPlan(*lattice_truss(a.nx), device=dev)
res = gd_large_mesh_iterations_per_second(dev, plan, iters=a.iters)
print(json.dumps({"nelem": plan.nelem, "ms_per_iteration": res["ms_per_iteration"]}))

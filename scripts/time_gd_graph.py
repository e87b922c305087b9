import sys, torch
sys.path.insert(0, ".")
from pinn_fem_b200 import AssemblyPlan
from pinn_fem_b200.meshes import lattice_truss
from bench_gd import gd_large_mesh_iterations_per_second
dev = torch.device("cuda", 0)
plan = AssemblyPlan(*lattice_truss(int(sys.argv[1]) if len(sys.argv) > 1 else 578), device=dev)
for rep in range(3):
    print(gd_large_mesh_iterations_per_second(dev, plan, iters=200)["ms_per_iteration"])

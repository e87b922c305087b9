import sys, torch
sys.path.insert(0, ".")
from pinn_fem_b200 import AssemblyPlan
from pinn_fem_b200.meshes import lattice_truss
dev = torch.device("cuda", 0)
plan = AssemblyPlan(*lattice_truss(578), device=dev)
g = torch.Generator(device=dev).manual_seed(0)
for B in (1, 4, 16, 32, 64, 128):
    shp = (plan.nelem,) if B == 1 else (plan.nelem, B)
    E = torch.rand(shp, generator=g, device=dev, dtype=torch.float64) + 0.5
    A = torch.rand(shp, generator=g, device=dev, dtype=torch.float64) + 0.5
    vals = plan.tangent_bsr(E, A, B=B)
    for _ in range(2): plan.tangent_bsr(E, A, B=B, out=vals)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): plan.tangent_bsr(E, A, B=B, out=vals)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ab = B * (16 * plan.nelem + 32 * plan.nnzb) + 24 * plan.nelem + 16 * plan.nnode
    print(B, f"{ms:.3f} ms  {ab/ms/1e6:.0f} GB/s  frac {ab/ms/1e6/6543.1:.3f}")
    del vals, E, A

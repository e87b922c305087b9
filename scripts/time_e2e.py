import sys, time, torch
sys.path.insert(0, ".")
from pinn_fem_b200 import AssemblyPlan
from pinn_fem_b200.meshes import lattice_truss
dev = torch.device("cuda", 0)
plan = AssemblyPlan(*lattice_truss(578), device=dev)
B = 512
uh = torch.rand((plan.ndof, B), dtype=torch.float64).pin_memory()
Eh = (torch.rand((plan.nelem, B), dtype=torch.float64) + 0.5).pin_memory()
Ah = (torch.rand((plan.nelem, B), dtype=torch.float64) + 0.5).pin_memory()
rh = torch.empty((plan.ndof, B), dtype=torch.float64).pin_memory()
fx = torch.randn(plan.ndof, dtype=torch.float64)
for chunk in (64, 128, 256, 512):
    plan.residual_host(uh, Eh, Ah, fx, 1.0, rh, chunk=chunk)
    t0 = time.perf_counter()
    for _ in range(2):
        plan.residual_host(uh, Eh, Ah, fx, 1.0, rh, chunk=chunk)
    dt = (time.perf_counter() - t0) / 2
    print(chunk, f"{dt*1e3:.1f} ms  {B*plan.nelem/dt/1e9:.3f} G evals/s  H2D {(plan.ndof+2*plan.nelem)*B*8/dt/1e9:.1f} GB/s")

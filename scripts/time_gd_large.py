"""Device-timed components of one PINN-GD iteration on the C5 lattice (single problem, 3 MLPs)."""
import argparse, json, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pinn_fem_b200 import AssemblyPlan, ops
from pinn_fem_b200.meshes import lattice_truss

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=578)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda", 0)
nodes, el, fixed = lattice_truss(a.nx)
plan = AssemblyPlan(nodes, el, fixed, device=dev)
specs = [ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15)]
g = torch.Generator(device=dev).manual_seed(0)
thetas = [0.3 * torch.randn(s.n_params, generator=g, device=dev, dtype=torch.float64) for s in specs]
u = (torch.rand(plan.ndof, generator=g, device=dev, dtype=torch.float64) - 0.5) * 2e-3
fx = torch.randn(plan.ndof, generator=g, device=dev, dtype=torch.float64) * 1e-3
gv = torch.randn(plan.nelem, generator=g, device=dev, dtype=torch.float64)


def timeit(fn, name, flops=None, bytes_=None):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.iters):
        out = fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / a.iters
    rec = {"op": name, "ms": round(ms, 4)}
    if flops:
        rec["TFLOPs"] = round(flops / (ms * 1e-3) / 1e12, 3)
    if bytes_:
        rec["GBs"] = round(bytes_ / (ms * 1e-3) / 1e9, 1)
    print(json.dumps(rec))
    return out


n = plan.nelem
for s, th, nm in zip(specs, thetas, ("young", "area")):
    fl = 2 * (3 * s.width + s.width * s.width + s.width) * n
    timeit(lambda: ops.mlp_forward(s, th, plan=plan, load_factor=1.0), f"mlp_forward[{nm}]", flops=fl)
    timeit(lambda: ops.mlp_backward(s, th, gv, plan=plan, load_factor=1.0), f"mlp_backward[{nm}]", flops=3 * fl)
E = ops.mlp_forward(specs[0], thetas[0], plan=plan, load_factor=1.0)
A = ops.mlp_forward(specs[1], thetas[1], plan=plan, load_factor=1.0)
ab = 8 * n + 16 * n + 48 * plan.nnode
timeit(lambda: plan.residual(u, E, A, fx, 1.0, f_int=False, r=True, half_sq=True), "residual B=1", bytes_=ab)
r = plan.residual(u, E, A, fx, 1.0, f_int=False, r=True)["r"]
timeit(lambda: plan.tangent_matvec(r, E, A), "matvec B=1", bytes_=ab)
timeit(lambda: plan.material_vjp(u, E, A, r), "material_vjp B=1", bytes_=8 * n + 32 * n + 32 * plan.nnode)
timeit(lambda: plan.tangent_bsr(E, A), "tangent_bsr B=1", bytes_=ab + 16 * n + 32 * plan.nnzb)

"""Quick device-timed sweep of the residual kernel on the C5 lattice."""
import argparse, json, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import pinnfem_oracle as O
from pinn_fem_b200 import AssemblyPlan

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=578)
ap.add_argument("--B", type=int, nargs="+", default=[1, 32, 128, 512])
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
t0 = time.time()
nodes, el, fixed = O.lattice_truss(a.nx)
t1 = time.time()
plan = AssemblyPlan(nodes, el, fixed, device="cuda:0")
t2 = time.time()
print(f"mesh {t1-t0:.2f}s plan {t2-t1:.2f}s nnode {plan.nnode} nelem {plan.nelem} nnzb {plan.nnzb}")
peak = 6543.1
for B in a.B:
    g = torch.Generator(device="cuda").manual_seed(B)
    shp = (lambda n: (n,)) if B == 1 else (lambda n: (n, B))
    u = (torch.rand(shp(plan.ndof), generator=g, device="cuda", dtype=torch.float64) - 0.5) * 2e-3
    E = torch.rand(shp(plan.nelem), generator=g, device="cuda", dtype=torch.float64) + 0.5
    A = torch.rand(shp(plan.nelem), generator=g, device="cuda", dtype=torch.float64) + 0.5
    fx = torch.randn(plan.ndof, generator=g, device="cuda", dtype=torch.float64)
    r = torch.empty_like(u)
    for _ in range(3):
        plan.residual_into(u, E, A, fx, 1.0, r)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
    ev[0].record()
    for i in range(a.iters):
        plan.residual_into(u, E, A, fx, 1.0, r)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(a.iters)])
    ab = (16 * plan.nelem + 32 * plan.nnode) * B + (8 * plan.nelem + 16 * plan.nnode if B == 1 else 0)
    gbs = ab / (ms.min() * 1e-3) / 1e9
    print(json.dumps({"B": B, "ms_min": float(ms.min()), "ms_med": float(np.median(ms)), "Gevals_s": B * plan.nelem / (ms.min() * 1e-3) / 1e9,
                      "alg_GBs": gbs, "frac": gbs / peak}))
    # tangent-vector product and material VJP for reference
    v = torch.randn_like(u)
    for fn, name in ((lambda: plan.tangent_matvec(v, E, A), "matvec"), (lambda: plan.material_vjp(u, E, A, v), "vjp")):
        fn(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        print("   ", name, f"{s.elapsed_time(e):.3f} ms")
    del u, E, A, r, v
if True:
    E1 = torch.rand(plan.nelem, device="cuda", dtype=torch.float64) + 0.5
    for _ in range(2):
        vals = plan.tangent_bsr(E1, E1)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); vals = plan.tangent_bsr(E1, E1); e.record(); torch.cuda.synchronize()
    ab = 40.04e6 / 999941 * plan.nelem + 16 * plan.nelem + 32 * plan.nnzb
    print("tangent_bsr B=1", f"{s.elapsed_time(e):.3f} ms", f"{ab / (s.elapsed_time(e) * 1e-3) / 1e9:.0f} GB/s alg")

import sys, time, torch
sys.path.insert(0, ".")
from pinn_fem_b200 import ops
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
for n in (1001, 2048, 4096):
    J = torch.randn((n, n), generator=g, device=dev, dtype=torch.float64)
    A = J.t() @ J + 1e-3 * torch.eye(n, device=dev, dtype=torch.float64)
    b = torch.randn(n, generator=g, device=dev, dtype=torch.float64)
    for name, fn in (("lu", ops.solve_dense), ("cholesky", ops.solve_spd)):
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); x = fn(A, b); e1.record(); torch.cuda.synchronize()
            print(n, name, rep, f"event {e0.elapsed_time(e1):.2f} ms wall {1e3*(time.perf_counter()-t0):.2f} ms resid {float((A@x-b).norm()/b.norm()):.2e}")

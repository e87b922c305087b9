import sys, torch
sys.path.insert(0, '.')
from pinn_fem_b200 import ops
n = 4096
g = torch.Generator(device="cuda").manual_seed(n)
M = torch.randn((n, n + 8), generator=g, device="cuda", dtype=torch.float64)
A = M @ M.T + 1e-3 * torch.eye(n, device="cuda", dtype=torch.float64)
b = torch.randn(n, generator=g, device="cuda", dtype=torch.float64)
ops.solve_spd(A, b); ops.solve_spd(A, b)
torch.cuda.synchronize()
print("ok")

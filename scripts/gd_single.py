import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pinn_fem_b200 import ops, AssemblyPlan
import bench_gd as B
dev = torch.device("cuda", 0)
plan = AssemblyPlan(B.NODES, B.ELEMENTS, B.FIXED, device=dev)
nets = [ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15), ops.NetSpec(3, 2, 10)]
for _ in range(2):
    theta = B._theta0(1, dev); u = torch.zeros((1, 8), dtype=torch.float64, device=dev)
    ops.gd_solve(plan, nets, [1.0, 1.0, 1.0], theta, u, torch.as_tensor(B.LOADS).to(dev), B.MEAS_DOFS, B.MEAS_VALS,
                 max_iterations=500, tolerance=0.0, learning_rate_u=0.01, learning_rate_theta=5e-4)
torch.cuda.synchronize()
print("ok")

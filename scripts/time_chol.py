"""Blocked Cholesky solve (pf_solve_spd): device time for n = 1001 / 2048 / 4096 and residual check."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pinn_fem_b200 import ops  # noqa: E402

for n in (257, 1001, 2048, 4096, 8192):
    g = torch.Generator(device="cuda").manual_seed(n)
    M = torch.randn((n, n + 8), generator=g, device="cuda", dtype=torch.float64)
    A = M @ M.T + 1e-3 * torch.eye(n, device="cuda", dtype=torch.float64)
    b = torch.randn(n, generator=g, device="cuda", dtype=torch.float64)
    x = ops.solve_spd(A, b)
    res = float(torch.linalg.vector_norm(A @ x - b) / torch.linalg.vector_norm(b))
    best = 1e9
    for rep in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.solve_spd(A, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"n={n}: {best:.3f} ms (incl. the clone of A), {n ** 3 / 3 / best / 1e9:.2f} TFLOP/s, relative residual {res:.2e}", flush=True)

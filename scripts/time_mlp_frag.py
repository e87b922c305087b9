"""Time the material-network kernels on the 999,941-element lattice's centroids (3-20-20-1 and 3-15-15-1):
fragment kernels (pf_mlp_frag.cu) single problem and batched, against the older kernels."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pinn_fem_b200 import ops  # noqa: E402


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(os.environ.get("N", 999941))
    rng = np.random.default_rng(0)
    X = torch.as_tensor(rng.uniform(0, 577, size=(n, 3))).cuda()
    X[:, 0] = 1.0
    for shape in ((3, 2, 20), (3, 2, 15)):
        spec = ops.NetSpec(*shape)
        theta = torch.as_tensor(rng.normal(scale=0.3, size=spec.n_params)).cuda()
        g = torch.as_tensor(rng.normal(size=n)).cuda()
        os.environ["PF_MLP_NO_FRAG"] = "1"
        t_f_old = timed(lambda: ops.mlp_forward(spec, theta, X))
        t_b_old = timed(lambda: ops.mlp_backward(spec, theta, g, X))
        os.environ["PF_MLP_NO_FRAG"] = "0"
        t_f = timed(lambda: ops.mlp_forward(spec, theta, X))
        t_b = timed(lambda: ops.mlp_backward(spec, theta, g, X))
        print(f"{shape} n={n}: forward old {t_f_old:.3f} ms, frag {t_f:.3f} ms; backward old {t_b_old:.3f} ms, "
              f"frag (forward+save, backward) {t_b:.3f} ms", flush=True)
        for B in (1, 4, 16, 64):
            th = torch.as_tensor(rng.normal(scale=0.3, size=(B, spec.n_params))).cuda()
            gb = torch.as_tensor(rng.normal(size=(n, B))).cuda()
            alen = ops.mlp_acts_len(spec, n)
            if B * alen * 8 > 60e9:
                continue
            out, acts = ops.mlp_forward_batched(spec, th, X)
            from pinn_fem_b200 import _lib
            from pinn_fem_b200.plan import _ptr, _stream_ptr
            lib = _lib.load()
            gt = torch.empty((B, spec.n_params), dtype=torch.float64, device="cuda")
            st = _stream_ptr(out.device)

            def fwd(save=True):
                _lib.check(lib.pf_mlp_forward_batched(None, *shape, _ptr(th), th.stride(0), B, n, _ptr(X), 1.0, 1.0, 1,
                                                      _ptr(out), B, _ptr(acts) if save else None, st))

            def bwd():
                _lib.check(lib.pf_mlp_backward_batched(None, *shape, _ptr(th), th.stride(0), B, n, _ptr(X), 1.0, _ptr(gb), B,
                                                       _ptr(acts), _ptr(gt), spec.n_params, st))

            tf, tfn, tb = timed(fwd, 5, 2), timed(lambda: fwd(False), 5, 2), timed(bwd, 5, 2)
            print(f"   B={B:3d}: forward+save {tf / B:.3f} ms/problem (no save {tfn / B:.3f}), backward {tb / B:.3f} ms/problem",
                  flush=True)
            del out, acts, gb


if __name__ == "__main__":
    main()

"""One batched forward (saving activations) and backward of the two material networks on the C5 centroids:
the command profiled with ncu for the fragment kernels (pf_mlp_frag.cu)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pinn_fem_b200 import ops  # noqa: E402

n, B = int(os.environ.get("N", 999941)), int(os.environ.get("B", 16))
rng = np.random.default_rng(0)
X = torch.as_tensor(rng.uniform(0, 577, size=(n, 3))).cuda()
X[:, 0] = 1.0
for rep in range(2):
    for shape in ((3, 2, 20), (3, 2, 15)):
        spec = ops.NetSpec(*shape)
        th = torch.as_tensor(rng.normal(scale=0.3, size=(B, spec.n_params))).cuda()
        g = torch.as_tensor(rng.normal(size=(n, B))).cuda()
        out, acts = ops.mlp_forward_batched(spec, th, X)
        gt = ops.mlp_backward_batched(spec, th, g, acts, X)
        torch.cuda.synchronize()
        del out, acts, g
print("ok", float(gt.abs().max()))

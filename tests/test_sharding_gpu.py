"""GPU: batch-sharded solves.  Single process (world 1) always; a 2-rank NCCL run under torchrun when
the box has two GPUs.  Sharded results must equal the unsharded batch bit for bit (problems are
independent and every kernel is deterministic)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent

WORKER = r"""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["PF_ROOT"])
from pinn_fem_b200 import AssemblyPlan, ops, sharding as S
from pinn_fem_b200 import bench_gd as B
rank, local, world = S.env_rank_world()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
plan = AssemblyPlan(B.NODES, B.ELEMENTS, B.FIXED, device=dev)
nets = [ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15), ops.NetSpec(3, 2, 10)]
nprob = 7
theta_all = B._theta0(nprob, torch.device("cuda", local)).cpu()
u_all = torch.zeros((nprob, 8), dtype=torch.float64)
rng = np.random.default_rng(3)
mv = torch.as_tensor(np.array(B.MEAS_VALS)[None, :] * rng.uniform(0.9, 1.1, size=(nprob, 1)))
kw = dict(max_iterations=80, tolerance=1e-7, learning_rate_u=0.01, learning_rate_theta=5e-4)
res, shard, summ = S.gd_solve_sharded(plan, nets, [1.0, 1.0, 1.0], theta_all, u_all, torch.as_tensor(B.LOADS).to(dev),
                                      B.MEAS_DOFS, mv, **kw)
# residual of this rank's shard of a batched lattice problem
from pinn_fem_b200.meshes import lattice_truss
nodes, el, fixed = lattice_truss(12)
lp = AssemblyPlan(nodes, el, fixed, device=dev)
g = torch.Generator().manual_seed(5)
Bt = 130
u = (torch.rand((lp.ndof, Bt), generator=g, dtype=torch.float64) - 0.5) * 2e-3
E = torch.rand((lp.nelem, Bt), generator=g, dtype=torch.float64) + 0.5
A = torch.rand((lp.nelem, Bt), generator=g, dtype=torch.float64) + 0.5
sh = S.shard_range(Bt, rank, world)
f = lp.internal_force(S.shard_problem_minor(u, sh).to(dev), S.shard_problem_minor(E, sh).to(dev),
                      S.shard_problem_minor(A, sh).to(dev))
f_all = S.gather_problem_rows(f.t().contiguous(), sh)          # [Bt, ndof] on every rank
u_rows = S.gather_problem_rows(res.u, shard)
th_rows = S.gather_problem_rows(res.theta, shard)
if rank == 0:
    torch.save({"u": u_rows.cpu(), "theta": th_rows.cpu(), "n_iters": summ.n_iters.cpu(), "converged": summ.converged.cpu(),
                "final": summ.final.cpu(), "f": f_all.cpu(), "world": world}, os.environ["PF_OUT"])
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
"""


def _run(tmp_path, world):
    out = tmp_path / f"w{world}.pt"
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, PF_ROOT=str(ROOT), PF_OUT=str(out))
    if world == 1:
        cmd = [sys.executable, str(script)]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)]
    proc = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    return torch.load(out)


def test_sharded_gd_world1_matches_direct_call(tmp_path):
    from pinn_fem_b200 import AssemblyPlan, ops
    from pinn_fem_b200 import bench_gd as B

    got = _run(tmp_path, 1)
    dev = torch.device("cuda", 0)
    plan = AssemblyPlan(B.NODES, B.ELEMENTS, B.FIXED, device=dev)
    nets = [ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15), ops.NetSpec(3, 2, 10)]
    theta = B._theta0(7, dev)
    u = torch.zeros((7, 8), dtype=torch.float64, device=dev)
    rng = np.random.default_rng(3)
    mv = np.array(B.MEAS_VALS)[None, :] * rng.uniform(0.9, 1.1, size=(7, 1))
    res = ops.gd_solve(plan, nets, [1.0, 1.0, 1.0], theta, u, torch.as_tensor(B.LOADS).to(dev), B.MEAS_DOFS, mv,
                       max_iterations=80, tolerance=1e-7, learning_rate_u=0.01, learning_rate_theta=5e-4)
    assert torch.equal(got["u"], res.u.cpu()) and torch.equal(got["theta"], res.theta.cpu())
    assert got["n_iters"].tolist() == res.n_iters.cpu().tolist()
    assert got["converged"].tolist() == res.converged.cpu().tolist()
    last = res.history[torch.arange(7, device=dev), res.n_iters.long() - 1].cpu()
    assert torch.equal(got["final"], last)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_nccl_shards_equal_single_gpu(tmp_path):
    one = _run(tmp_path, 1)
    two = _run(tmp_path, 2)
    assert two["world"] == 2
    for k in ("u", "theta", "n_iters", "converged", "final", "f"):
        assert torch.equal(one[k], two[k]), k

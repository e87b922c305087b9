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
import bench_gd as B
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
    import bench_gd as B

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


ELEM_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["PF_ROOT"])
from pinn_fem_b200 import AssemblyPlan, ops, sharding as S
from pinn_fem_b200.element_sharding import Communicator, ShardedMesh, gd_solve_element_sharded
from pinn_fem_b200.meshes import lattice_truss
rank, local, world = S.env_rank_world()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nodes, el, fixed = lattice_truss(40)
ndof = 2 * len(nodes)
rng = np.random.default_rng(21)
comm = Communicator(dev)
mesh = ShardedMesh(nodes, el, fixed, comm)
lm = mesh.local
# ---- internal force of the owned rows, batched (B = 3), vs the single-GPU plan on the same data
u = rng.uniform(-1e-3, 1e-3, (ndof, 3)); E = rng.uniform(0.5, 1.5, (len(el), 3)); A = rng.uniform(0.5, 1.5, (len(el), 3))
ul = torch.as_tensor(lm.to_local_vector(u)).to(dev)
ul[mesh.n_owned_dofs:] = 7.0   # stale halo rows: the exchange must overwrite them
f_own = mesh.internal_force_owned(ul, torch.as_tensor(E[lm.elements_global]).to(dev).contiguous(),
                                  torch.as_tensor(A[lm.elements_global]).to(dev).contiguous())
full = AssemblyPlan(nodes, el, fixed, device=dev)
f_full = full.internal_force(torch.as_tensor(u).to(dev), torch.as_tensor(E).to(dev), torch.as_tensor(A).to(dev))
own_dofs = torch.as_tensor(lm.local_dofs_global()[: mesh.n_owned_dofs]).to(dev)
force_bitwise = bool(torch.equal(f_own, f_full[own_dofs]))
t = torch.arange(4, dtype=torch.float64, device=dev) + rank
comm.allreduce_sum_(t)
# ---- element-sharded PINN-GD vs the single-GPU large-mesh loop
loads = np.zeros(ndof); loads[-2], loads[-1] = 0.05, -0.02
nets = [ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15), None]
theta0 = rng.normal(scale=0.3, size=nets[0].n_params + nets[1].n_params)
md = np.array([2 * 799, 2 * 799 + 1, 411, 411, 1202, 3100]); mv = np.array([0.01, -0.02, 0.005, 0.004, -0.003, 0.002])
if os.environ.get("PF_MEAS_RANK0"):   # every measurement on rank 0's rows: the other rank holds none
    md = np.array([2 * 799, 2 * 799 + 1, 411, 411, 1202, 1500])
kw = dict(max_iterations=int(os.environ.get("PF_ITERS", "20")), tolerance=float(os.environ.get("PF_TOL", "1e-14")),
          learning_rate_u=1e-4, learning_rate_theta=1e-3, alpha_data=10.0, load_factor=0.9)
out = gd_solve_element_sharded(mesh, nets, [2.0, 1.5, 1.0], theta0, np.zeros(ndof), loads, md, mv, **kw)
ref = ops.gd_solve(full, nets, [2.0, 1.5, 1.0], torch.as_tensor(theta0).to(dev)[None].clone(),
                   torch.zeros((1, ndof), dtype=torch.float64, device=dev), torch.as_tensor(loads).to(dev), md, mv, **kw)
# gather owned rows of every rank in rank order -> global vectors
def gather_owned(x_owned):
    counts = [m * 2 for m in S.shard_counts(len(nodes), world)]
    pad = torch.zeros(max(counts), dtype=torch.float64, device=dev); pad[: x_owned.numel()] = x_owned
    parts = [torch.empty_like(pad) for _ in range(world)]
    if world > 1: dist.all_gather(parts, pad)
    else: parts = [pad]
    return torch.cat([p[:c] for p, c in zip(parts, counts)])
u_all, reac_all = gather_owned(out["u_owned"]), gather_owned(out["reactions_owned"])
thetas = [torch.empty_like(out["theta"]) for _ in range(world)]
if world > 1: dist.all_gather(thetas, out["theta"])
else: thetas = [out["theta"]]
if rank == 0:
    torch.save({"world": world, "transport": comm.transport, "force_bitwise": force_bitwise, "allreduce": t.cpu(), "u": u_all.cpu(), "reac": reac_all.cpu(),
                "theta": out["theta"].cpu(), "theta_same": all(torch.equal(thetas[0], x) for x in thetas),
                "hist": out["history"].cpu(), "n": out["n_iters"], "conv": out["converged"],
                "u_ref": ref.u[0].cpu(), "reac_ref": ref.reactions[0].cpu(), "theta_ref": ref.theta[0].cpu(),
                "hist_ref": ref.history[0, : int(ref.n_iters[0])].cpu(), "n_ref": int(ref.n_iters[0]),
                "conv_ref": bool(ref.converged[0])}, os.environ["PF_OUT"])
force_ok = torch.tensor([1.0 if force_bitwise else 0.0], device=dev)
if world > 1:
    dist.all_reduce(force_ok, op=dist.ReduceOp.MIN); dist.barrier()
assert force_ok.item() == 1.0
mesh.close(); comm.close()
if world > 1: dist.destroy_process_group()
"""


def _run_elem(tmp_path, world, env_extra=None):
    out = tmp_path / f"e{world}.pt"
    script = tmp_path / "elem_worker.py"
    script.write_text(ELEM_WORKER)
    env = dict(os.environ, PF_ROOT=str(ROOT), PF_OUT=str(out), **(env_extra or {}))
    if world == 1:
        cmd = [sys.executable, str(script)]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)]
    proc = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    return torch.load(out)


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def _check_elem(g):
    assert g["force_bitwise"]
    assert g["n"] == g["n_ref"] and g["conv"] == g["conv_ref"] and g["theta_same"]
    assert _rel(g["u"], g["u_ref"]) < 1e-10 and _rel(g["theta"], g["theta_ref"]) < 1e-10
    assert _rel(g["reac"], g["reac_ref"]) < 1e-9
    for col in range(1, 7):
        assert _rel(g["hist"][:, col], g["hist_ref"][:, col]) < 1e-10, col


def test_element_sharded_world1_is_the_large_mesh_loop(tmp_path):
    """World size 1: the sharded entry point with an empty halo must reproduce pf_gd_solve's large-mesh loop."""
    g = _run_elem(tmp_path, 1)
    assert g["allreduce"].tolist() == [0.0, 1.0, 2.0, 3.0]
    _check_elem(g)
    assert torch.equal(g["u"], g["u_ref"]) and torch.equal(g["theta"], g["theta_ref"])


def _peer_capable():
    return torch.cuda.device_count() >= 2 and torch.cuda.can_device_access_peer(0, 1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_element_sharded_two_gpus_match_single_gpu(tmp_path, transport):
    """Two ranks, halo exchange + gradient all-reduce over peer memory (stores into the neighbour's mailbox,
    all-reduce fused into the Adam kernel) and over NCCL: owned rows of f_int bitwise equal to the single-GPU
    result, GD end state and history equal to 1e-10 (the all-reduce changes the summation order of dL/dtheta)."""
    if transport == "peer" and not _peer_capable():
        pytest.skip("GPUs 0 and 1 cannot map each other's memory")
    env = {"PF_COMM_TRANSPORT": transport}
    g = _run_elem(tmp_path, 2, env)
    assert g["world"] == 2 and g["allreduce"].tolist() == [1.0, 3.0, 5.0, 7.0]
    assert g["transport"] == transport
    _check_elem(g)
    # early stop: every rank sees the same flag, at the iteration the single-GPU loop stops
    g2 = _run_elem(tmp_path, 2, dict(env, PF_ITERS="40", PF_TOL="1e3"))
    assert g2["n"] == g2["n_ref"] == 12 and g2["conv"] and g2["conv_ref"]
    _check_elem(g2)
    # all measurements on rank 0's rows: rank 1 contributes an empty data loss every iteration
    g3 = _run_elem(tmp_path, 2, dict(env, PF_MEAS_RANK0="1"))
    _check_elem(g3)


@pytest.mark.skipif(torch.cuda.device_count() < 3, reason="needs more than two GPUs")
def test_element_sharded_all_gpus_match_single_gpu(tmp_path):
    """Every GPU of the box (interior ranks have two neighbours, the all-reduce sums world slots)."""
    world = torch.cuda.device_count()
    g = _run_elem(tmp_path, world, {"PF_COMM_TRANSPORT": "auto"})
    assert g["world"] == world and g["allreduce"].tolist() == [sum(range(world)) + world * k for k in range(4)]
    if all(torch.cuda.can_device_access_peer(0, d) for d in range(1, world)):
        assert g["transport"] == "peer"
    _check_elem(g)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_and_nccl_transports_agree(tmp_path):
    """Same run on both transports: u and f_int rows are bitwise equal (the halo rows carry the same values);
    theta agrees to rounding (different all-reduce order)."""
    if not _peer_capable():
        pytest.skip("GPUs 0 and 1 cannot map each other's memory")
    a = _run_elem(tmp_path, 2, {"PF_COMM_TRANSPORT": "peer", "PF_ITERS": "1"})
    b = _run_elem(tmp_path, 2, {"PF_COMM_TRANSPORT": "nccl", "PF_ITERS": "1"})
    assert a["transport"] == "peer" and b["transport"] == "nccl"
    assert torch.equal(a["u"], b["u"])  # one iteration: u does not depend on the reduced gradient yet
    assert _rel(a["theta"], b["theta"]) < 1e-13

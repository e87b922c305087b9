"""Batch sharding of independent problems over the GPUs of one box.

The hot path shards by *problem*: inverse problems / load cases that share a mesh are
independent, so every rank owns a contiguous slice of the batch, the mesh plan is
replicated, and there is **no collective on the data path**.  The only traffic is what
the caller wants to know about the whole batch afterwards -- per-problem iteration counts,
convergence flags, final losses, the timing of the slowest rank -- a few bytes per problem,
gathered once per solve.

One process per GPU (``torchrun``); the same code runs under the ``gloo`` backend on CPU
tensors, which is how the tests exercise it without a GPU.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Shard:
    """Contiguous slice ``[start, stop)`` of ``total`` problems owned by ``rank`` of ``world``."""
    rank: int
    world: int
    total: int
    start: int
    stop: int

    @property
    def count(self) -> int:
        return self.stop - self.start

    def slice(self) -> slice:
        return slice(self.start, self.stop)


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world) from the torchrun environment; (0, 0, 1) outside it."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(total: int, rank: int, world: int) -> Shard:
    """Balanced contiguous partition: the first ``total % world`` ranks get one extra problem.
    Shards are ordered by rank, so concatenating them in rank order restores problem order."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world: {rank}/{world}")
    if total < 0:
        raise ValueError("total must be >= 0")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return Shard(rank, world, total, start, start + base + (1 if rank < extra else 0))


def shard_counts(total: int, world: int) -> Sequence[int]:
    return [shard_range(total, r, world).count for r in range(world)]


def shard_problem_major(t: torch.Tensor, shard: Shard) -> torch.Tensor:
    """Slice of a problem-major array ``[nprob, ...]`` (theta, u of the GD loop)."""
    if t.shape[0] != shard.total:
        raise ValueError(f"leading dimension {t.shape[0]} != total problems {shard.total}")
    return t[shard.slice()].contiguous()


def shard_problem_minor(t: torch.Tensor, shard: Shard) -> torch.Tensor:
    """Slice of a batched array ``[rows, B]`` with the problem index last (u, E, A of the
    assembly kernels)."""
    if t.shape[-1] != shard.total:
        raise ValueError(f"trailing dimension {t.shape[-1]} != total problems {shard.total}")
    return t[..., shard.slice()].contiguous()


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def gather_problem_rows(local: torch.Tensor, shard: Shard, group=None) -> torch.Tensor:
    """All-gather per-problem rows ``[shard.count, k]`` (or ``[shard.count]``) into
    ``[total, k]`` on every rank, in problem order.  Ragged shards are padded to the widest
    one for the collective and trimmed afterwards."""
    if local.shape[0] != shard.count:
        raise ValueError(f"local rows {local.shape[0]} != shard size {shard.count}")
    if _world(group) == 1:
        return local.clone()
    counts = shard_counts(shard.total, shard.world)
    width = max(counts)
    tail = tuple(local.shape[1:])
    padded = torch.zeros((width,) + tail, dtype=local.dtype, device=local.device)
    padded[: shard.count] = local
    parts = [torch.empty_like(padded) for _ in range(shard.world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def max_over_ranks(value: float, device: Optional[torch.device] = None, group=None) -> float:
    """The slowest rank's timing: multi-GPU numbers are the max over ranks, never an average."""
    if _world(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value: float, device: Optional[torch.device] = None, group=None) -> float:
    if _world(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


@dataclass
class BatchSummary:
    """What every rank knows about the whole batch after a sharded solve."""
    n_iters: torch.Tensor       # int64 [total]
    converged: torch.Tensor     # int64 [total]
    final: torch.Tensor         # float64 [total, k] last history row (or any per-problem metrics)

    @property
    def all_converged(self) -> bool:
        return bool(self.converged.all().item()) if self.converged.numel() else True


def summarize_batch(n_iters: torch.Tensor, converged: torch.Tensor, final: torch.Tensor, shard: Shard,
                    group=None) -> BatchSummary:
    """One small all-gather per solve: iteration counts, convergence flags and final metrics of
    every problem, identical on all ranks."""
    packed = torch.cat([n_iters.to(torch.float64).reshape(shard.count, 1),
                        converged.to(torch.float64).reshape(shard.count, 1),
                        final.to(torch.float64).reshape(shard.count, math.prod(final.shape[1:]))], dim=1)
    full = gather_problem_rows(packed, shard, group)
    return BatchSummary(n_iters=full[:, 0].round().to(torch.int64), converged=full[:, 1].round().to(torch.int64),
                        final=full[:, 2:].contiguous())


def gd_solve_sharded(plan, nets, scales, theta_all, u_all, f_ext, meas_dofs=None, meas_vals_all=None, group=None,
                     **kw):
    """Batch-sharded ``ops.gd_solve``: every rank runs the device-resident GD loop on its slice of
    ``theta_all`` ``[nprob, n_theta]`` / ``u_all`` ``[nprob, ndof]`` (host or device tensors holding
    the whole batch) and learns the convergence summary of all problems.  Returns
    ``(local GDResult, Shard, BatchSummary)``."""
    from . import ops

    rank, _, world = env_rank_world()
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    shard = shard_range(u_all.shape[0], rank, world)
    dev = plan.device
    theta = shard_problem_major(theta_all, shard).to(dev)
    u = shard_problem_major(u_all, shard).to(dev)
    mv = None
    if meas_vals_all is not None:
        mv = torch.as_tensor(meas_vals_all)
        if mv.dim() == 2:
            mv = shard_problem_major(mv, shard)
    if shard.count == 0:
        raise ValueError("fewer problems than ranks: every rank needs at least one problem")
    res = ops.gd_solve(plan, nets, scales, theta, u, f_ext, meas_dofs, mv, **kw)
    if res.history is not None:
        last = (res.n_iters.to(torch.int64) - 1).clamp(min=0)
        final = res.history[torch.arange(shard.count, device=dev), last]
    else:
        final = torch.zeros((shard.count, 0), dtype=torch.float64, device=dev)
    backend = dist.get_backend(group) if world > 1 else None
    to = (lambda t: t) if backend == "nccl" else (lambda t: t.cpu())
    summary = summarize_batch(to(res.n_iters), to(res.converged), to(final), shard, group)
    return res, shard, summary

"""Batch sharding on CPU: world_size-2 gloo processes exercise the N>1 host logic
of pinn_fem_b200/sharding.py -- partition, ragged all-gather of per-problem metrics, max-over-ranks
timing -- and the shard-seeded synthetic inputs of bench.py."""
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pinn_fem_b200 import sharding as S


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_q):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        shard = S.shard_range(total, rank, world)
        # per-problem metrics every rank computes from the global problem id
        ids = torch.arange(shard.start, shard.stop, dtype=torch.float64)
        n_it = (10 + ids).to(torch.int32)
        conv = (ids.to(torch.int64) % 2 == 0).to(torch.int32)
        final = torch.stack([ids * 0.5, -ids], dim=1) if shard.count else torch.zeros((0, 2), dtype=torch.float64)
        summ = S.summarize_batch(n_it, conv, final, shard)
        rows = S.gather_problem_rows(ids, shard)
        slow = S.max_over_ranks(1.0 + rank)
        tot = S.sum_over_ranks(float(shard.count))
        # batched arrays: problem index last / first
        full_minor = torch.arange(3 * total, dtype=torch.float64).reshape(3, total)
        full_major = torch.arange(total * 4, dtype=torch.float64).reshape(total, 4)
        mine_minor = S.shard_problem_minor(full_minor, shard)
        mine_major = S.shard_problem_major(full_major, shard)
        back = S.gather_problem_rows(mine_major, shard)
        out_q.put((rank, shard.start, shard.stop, summ.n_iters.tolist(), summ.converged.tolist(),
                   summ.final.tolist(), rows.tolist(), slow, tot, mine_minor.tolist(),
                   bool(torch.equal(back, full_major)), summ.all_converged))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 8, 2, 1])
def test_two_rank_gloo_gather(total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # shards tile [0, total) in rank order
    assert results[0][1] == 0 and results[0][2] == results[1][1] and results[1][2] == total
    exp_ids = list(range(total))
    for rank, start, stop, n_it, conv, final, rows, slow, tot, minor, roundtrip, allc in results:
        assert n_it == [10 + i for i in exp_ids]
        assert conv == [1 if i % 2 == 0 else 0 for i in exp_ids]
        assert final == [[0.5 * i, -float(i)] for i in exp_ids]
        assert rows == [float(i) for i in exp_ids]
        assert slow == 2.0 and tot == float(total)
        assert minor == [[float(r * total + c) for c in range(start, stop)] for r in range(3)]
        assert roundtrip
        assert allc == (total == 1)


def test_shard_range_properties():
    for total in (0, 1, 7, 512, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            shards = [S.shard_range(total, r, world) for r in range(world)]
            assert shards[0].start == 0 and shards[-1].stop == total
            assert all(a.stop == b.start for a, b in zip(shards, shards[1:]))
            counts = [s.count for s in shards]
            assert max(counts) - min(counts) <= 1 and counts == sorted(counts, reverse=True)
            assert list(S.shard_counts(total, world)) == counts
    assert S.shard_range(4096, 3, 8).count == 512
    with pytest.raises(ValueError):
        S.shard_range(4, 2, 2)


def test_single_process_paths_need_no_process_group():
    shard = S.shard_range(3, 0, 1)
    x = torch.arange(6, dtype=torch.float64).reshape(3, 2)
    assert torch.equal(S.gather_problem_rows(x, shard), x)
    assert S.max_over_ranks(3.5) == 3.5 and S.sum_over_ranks(2.0) == 2.0
    with pytest.raises(ValueError):
        S.gather_problem_rows(x[:2], shard)


def test_bench_inputs_are_seeded_per_problem():
    """bench.py seeds problem p with np.random.default_rng(p): rank k holds problems [k B, (k+1) B), so
    shards differ, reruns agree, and a shard does not depend on how many ranks there are."""
    import bench

    a = bench.synthetic_inputs(ndof=10, nelem=7, B=4, first_problem=0, device="cpu", group=3)
    b = bench.synthetic_inputs(ndof=10, nelem=7, B=4, first_problem=4, device="cpu")
    a2 = bench.synthetic_inputs(ndof=10, nelem=7, B=4, first_problem=0, device="cpu")
    assert all(torch.equal(x, y) for x, y in zip(a, a2))
    assert not torch.equal(a[0], b[0])
    wide = bench.synthetic_inputs(ndof=10, nelem=7, B=8, first_problem=0, device="cpu")
    assert torch.equal(wide[0][:, 4:], b[0]) and torch.equal(wide[2][:, :4], a[2])
    pu, pe, pa = bench.synthetic_problem(5, 10, 7)
    assert np.array_equal(b[0][:, 1].numpy(), pu) and np.array_equal(b[1][:, 1].numpy(), pe)
    rng = np.random.default_rng(5)
    assert np.array_equal(pu, rng.uniform(-1e-3, 1e-3, 10)) and np.array_equal(pe, rng.uniform(0.5, 1.5, 7))
    u, E, A, fx = a
    assert u.shape == (10, 4) and E.shape == (7, 4) and fx.shape == (10,)
    assert float(u.abs().max()) <= 1e-3 and 0.5 <= float(E.min()) and float(A.max()) <= 1.5
    assert bench.algorithmic_bytes(999941, 334084, 512) == 13678493800
    assert np.isclose(bench.algorithmic_bytes(999941, 334084, 1) / 999941, 40.04, atol=0.02)

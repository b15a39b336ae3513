"""N > 1 host logic on the CPU: two gloo ranks shard a batch, evaluate (with the oracle standing in for the
device — this file tests sharding and gathering, not kernels) and all_gather; the result must equal the unsharded run."""
import os
import socket

import numpy as np
import pytest

from pde_controller_b200 import sharding, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, seed, q):
    import torch
    import torch.distributed as dist
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (a) every rank regenerates only ITS shard of the global synthetic set
        lo, hi = sharding.shard_bounds(n, rank, world)
        mine = synth.make_batch(hi - lo, "mixed", seed=seed, start=lo, total=n)
        rob, st = oracle.eval_batch(mine, nthreads=1)
        sizes = [sharding.shard_bounds(n, r, world)[1] - sharding.shard_bounds(n, r, world)[0] for r in range(world)]
        full = sharding.gather_results(torch.from_numpy(rob), sizes)
        # (b) zero-copy shard views of one global batch give the same numbers
        whole = synth.make_batch(n, "mixed", seed=seed, total=n)
        rob2, _ = oracle.eval_batch(whole.shard(lo, hi), nthreads=1)
        assert np.array_equal(rob2, rob)
        dist.barrier()
        if rank == 0:
            q.put(full.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_gather_matches_single_process(oracle_lib):
    import torch.multiprocessing as mp
    n, seed, world = 9001, 4, 2          # odd size, spans several generator blocks and the heat/wave boundary
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref, _ = oracle_lib.eval_batch(synth.make_batch(n, "mixed", seed=seed, total=n))
    assert np.array_equal(got, ref)


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            b = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_cost_balanced_bounds():
    b = synth.make_batch(4096, "mixed", seed=1)
    cost = sharding.batch_cost(b)
    cuts = sharding.cost_balanced_bounds(cost, 8)
    assert cuts[0][0] == 0 and cuts[-1][1] == b.n
    loads = np.array([cost[l:h].sum() for l, h in cuts])
    assert loads.max() / loads.mean() < 1.05
    by_count = np.array([cost[l:h].sum() for l, h in (sharding.shard_bounds(b.n, r, 8) for r in range(8))])
    assert loads.max() <= by_count.max()


def test_kind_slices_partition_the_mixed_set_and_balance_cost():
    """make_batch(kind_slice=(r, R)): the R shards are a partition of the global mixed set, each with both kinds."""
    R, n = 4, 2048
    whole = synth.make_batch(R * n, "mixed", seed=5)
    key = lambda x: set(map(tuple, np.stack([x.kind, x.L, x.T, x.g0, x.E_in, x.pred_b[0]]).T.tolist()))
    parts = [synth.make_batch(n, "mixed", seed=5, kind_slice=(r, R)) for r in range(R)]
    union = set()
    for p in parts:
        assert (p.kind == 0).sum() == n // 2 and (p.kind == 1).sum() == n - n // 2
        k = key(p)
        assert len(k) == n and not (k & union)
        union |= k
    assert union == key(whole)
    loads = np.array([sharding.batch_cost(p).sum() for p in parts])
    assert loads.max() / loads.mean() < 1.05
    contiguous = np.array([sharding.batch_cost(synth.make_batch(n, "mixed", seed=5, start=r * n, total=R * n)).sum() for r in range(R)])
    assert contiguous.max() / contiguous.mean() > 1.3

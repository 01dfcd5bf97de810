"""CPU, world_size 2, gloo: the N>1 host logic — contiguous candidate shards, no data-path collective, rank-ordered
gather of per-candidate rows."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dl4vc_b200.shard import shard_range, gather_in_rank_order


@pytest.mark.parametrize("n,world", [(0, 2), (1, 2), (7, 2), (8, 8), (1_000_003, 8), (5, 8)])
def test_shard_range_partitions_in_order(n, world):
    edges = [shard_range(n, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == n
    for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in edges]
    assert max(sizes) - min(sizes) <= 1


def test_shard_range_rejects_bad_rank():
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n, rank, world)
        # stand-in for the per-rank forward: row i carries its global candidate index in every head column
        rows = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 27)
        full = gather_in_rank_order(rows, n)
        if rank == 0:
            q.put(full.numpy())
        else:
            assert full is None
        # the timing reduction bench.py performs: max over ranks
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gather_preserves_candidate_order():
    n, world = 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    full = q.get(timeout=90)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert full.shape == (n, 27)
    assert (full[:, 0] == torch.arange(n).numpy()).all()

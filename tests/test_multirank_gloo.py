"""world_size-2 CPU test (gloo) of the multi-GPU host logic: QP sharding with no
data-path collective, and the max/sum scalar reductions used for timing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fenics_constitutive_b200.partition import max_over_ranks, shard_range, sum_over_ranks


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n: int, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n, rank, world)
        # each rank "processes" its shard; the only communication is scalar
        total = sum_over_ranks(float(hi - lo))
        slowest = max_over_ranks(float(rank + 1))
        # shards tile [0, n) exactly: gather the bounds and check on rank 0
        bounds = [None] * world
        dist.all_gather_object(bounds, (lo, hi))
        if rank == 0:
            out.put((total, slowest, bounds))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_sharding_and_reductions():
    world, n = 2, 1_000_003
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    total, slowest, bounds = out.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert total == float(n)
    assert slowest == 2.0
    assert bounds[0][0] == 0 and bounds[0][1] == bounds[1][0] and bounds[1][1] == n

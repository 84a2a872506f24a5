"""N>1 host logic on CPU: world_size-2 gloo group, round-robin world sharding with no data-path
collective, and the whole-job metric = sum(units) / max(time)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from box2d_optimized_b200.sharding import aggregate, shard_worlds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_worlds, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_worlds(num_worlds, rank, world)
    # every world "simulates" 211 bodies for 10 steps; rank 1 is slower
    units = len(mine) * 211 * 10
    ms = 5.0 if rank == 0 else 8.0
    total, tmax, rate = aggregate(units, ms)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out.put((gathered, total, tmax, rate))
    dist.destroy_process_group()


def test_two_rank_sharding_and_metric():
    num_worlds = 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_worlds, q)) for r in range(2)]
    [p.start() for p in procs]
    gathered, total, tmax, rate = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert sorted(gathered[0] + gathered[1]) == list(range(num_worlds))   # every world exactly once
    assert set(gathered[0]).isdisjoint(gathered[1])
    assert gathered[0] == [0, 2, 4, 6] and gathered[1] == [1, 3, 5]
    assert total == num_worlds * 211 * 10
    assert tmax == 8.0                                                     # max over ranks, not mean
    assert rate == pytest.approx(total / 0.008)


def test_single_process_passthrough():
    assert shard_worlds(5, 0, 1) == [0, 1, 2, 3, 4]
    u, t, r = aggregate(1000, 2.0)
    assert (u, t, r) == (1000.0, 2.0, 500000.0)
    with pytest.raises(ValueError):
        shard_worlds(4, 2, 2)

"""world_size-2 gloo test of the multi-GPU host logic (env sharding has no data-path collective; only
timings and end-of-run scalars are reduced)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import pkg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg()
    from snake_b200 import shard
    lo, hi = shard.shard_range(n_total, rank, world)
    ranges = [None] * world
    dist.all_gather_object(ranges, (lo, hi))
    t = shard.max_over_ranks(10.0 + rank)            # rank 1 is the slow one
    s = shard.sum_over_ranks(hi - lo)
    wl, wh = shard.weak_range(1000, rank)
    q.put((rank, ranges, t, s, (wl, wh), shard.rank_seed(42, rank)))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    world, n_total = 2, 1_048_577
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ranges, t, s, weak, seed in res:
        assert ranges[0][0] == 0 and ranges[-1][1] == n_total
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))      # disjoint, complete
        assert max(h - l for l, h in ranges) - min(h - l for l, h in ranges) <= 1
        assert t == 11.0 and s == n_total
        assert weak == (rank * 1000, (rank + 1) * 1000)
    assert res[0][5] != res[1][5]


def test_shard_range_properties():
    pkg()
    from snake_b200 import shard
    for n in (1, 7, 4096, 1 << 20, 1_000_003):
        for w in (1, 2, 4, 8):
            rs = [shard.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
    assert shard.max_over_ranks(3.5) == 3.5

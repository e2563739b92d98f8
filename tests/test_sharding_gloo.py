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


def _gram_worker(rank, world, port, K, P, q):
    """The sharded-Gram schedule (ring order, column offsets, transposed-peer symmetrise) under gloo, with
    numpy standing in for the device kernels: planes are exchanged with all_gather, block products follow
    ring_schedule, and the result must be the full Gram's row block."""
    import numpy as np
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg()
    from snake_b200 import shard
    from snake_b200.gram_sharded import ring_schedule
    rng = np.random.default_rng(0)
    A = rng.normal(0, 1, (K, P))
    rows_all = [shard.shard_range(K, r, world)[1] - shard.shard_range(K, r, world)[0] for r in range(world)]
    col0 = [sum(rows_all[:r]) for r in range(world)]
    mine = A[col0[rank]:col0[rank] + rows_all[rank]]
    hi = mine.astype(np.float32).astype(np.float64)            # stand-in split: hi + lo == mine
    lo2 = 2 * (mine - hi)
    planes = [None] * world
    dist.all_gather_object(planes, (hi, lo2))
    Y = np.zeros((rows_all[rank], K))
    order = ring_schedule(rank, world)
    assert order[0] == rank and sorted(order) == list(range(world))
    for p in order:
        ph, pl = planes[p]
        Y[:, col0[p]:col0[p] + rows_all[p]] = hi @ ph.T + hi @ pl.T
    Ys = [None] * world
    dist.all_gather_object(Ys, Y)
    G = np.zeros_like(Y)
    for p in range(world):
        yt = Ys[p][:, col0[rank]:col0[rank] + rows_all[rank]]
        G[:, col0[p]:col0[p] + rows_all[p]] = 0.5 * (Y[:, col0[p]:col0[p] + rows_all[p]] + yt.T)
    ref = (A @ A.T)[col0[rank]:col0[rank] + rows_all[rank]]
    q.put((rank, float(np.abs(G - ref).max() / np.abs(ref).max())))
    dist.destroy_process_group()


def test_sharded_gram_schedule_two_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_gram_worker, args=(r, world, port, 37, 50, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    # only lo*lo^T is dropped: relative error ~ (2^-24)^2
    assert all(err < 1e-12 for _, err in res), res


def test_cpulist_parsing_and_affinity_helper_never_raises():
    pkg()
    from snake_b200 import shard
    assert shard.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert shard.parse_cpulist("5") == [5]
    assert shard.parse_cpulist("") == []
    info = shard.bind_host_near_gpu(0)          # no GPU here: reports why nothing was bound
    assert info["bound"] is False and "reason" in info

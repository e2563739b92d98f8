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
    """The sharded-Gram schedule (half ring, column offsets, the shared last step, transposed-peer mirror) under gloo, with
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
    lo = mine - hi
    planes = [None] * world
    dist.all_gather_object(planes, (hi, lo))
    Y = np.full((rows_all[rank], K), np.nan)
    sched = ring_schedule(rows_all, rank)
    assert sched[0][0] == rank and len(sched) == world // 2 + 1
    for p, a0, a1, b0, b1 in sched:
        ph, pl = planes[p]
        Y[a0:a1, col0[p] + b0:col0[p] + b1] = hi[a0:a1] @ ph[b0:b1].T + hi[a0:a1] @ pl[b0:b1].T + lo[a0:a1] @ ph[b0:b1].T
    Ys = [None] * world
    dist.all_gather_object(Ys, Y)
    G = Y.copy()
    for i in range(1, len(sched)):
        src = (rank - i) % world                               # rank src computed (its rows a0:a1) x (my rows b0:b1) at ITS step i
        _, a0, a1, b0, b1 = ring_schedule(rows_all, src)[i]
        assert np.isnan(G[b0:b1, col0[src] + a0:col0[src] + a1]).all()          # nobody computes a block twice
        G[b0:b1, col0[src] + a0:col0[src] + a1] = Ys[src][a0:a1, col0[rank] + b0:col0[rank] + b1].T
    assert not np.isnan(G).any()
    ref = (A @ A.T)[col0[rank]:col0[rank] + rows_all[rank]]
    q.put((rank, float(np.abs(G - ref).max() / np.abs(ref).max())))
    dist.destroy_process_group()


def test_ring_schedule_covers_every_block_exactly_once():
    """pure arithmetic (snk_gram_shard_schedule needs no GPU): for every world size the computed blocks and their mirror images
    tile the K x K matrix exactly once, and the shared last step of an even world splits the lower rank's rows"""
    import numpy as np
    pkg()
    from snake_b200.gram_sharded import ring_schedule
    for world, rows_all in [(1, [5]), (2, [3, 4]), (2, [1250, 1250]), (3, [7, 7, 6]), (4, [300, 300, 299, 299]), (5, [2] * 5),
                            (8, [6250] * 8), (8, [1] * 8), (16, [40] * 16)]:
        K = sum(rows_all)
        col0 = [sum(rows_all[:r]) for r in range(world)]
        cover = np.zeros((K, K), dtype=np.int32) if K <= 5000 else None
        work = []
        for g in range(world):
            sched = ring_schedule(rows_all, g)
            assert len(sched) == world // 2 + 1 and sched[0] == (g, 0, rows_all[g], 0, rows_all[g])
            w = 0
            for i, (p, a0, a1, b0, b1) in enumerate(sched):
                assert p == (g + i) % world and 0 <= a0 <= a1 <= rows_all[g] and 0 <= b0 <= b1 <= rows_all[p]
                w += (a1 - a0) * (b1 - b0) * (0.5 if i == 0 else 1.0)
                if cover is not None:
                    cover[col0[g] + a0:col0[g] + a1, col0[p] + b0:col0[p] + b1] += 1
                    if i > 0:
                        cover[col0[p] + b0:col0[p] + b1, col0[g] + a0:col0[g] + a1] += 1
            work.append(w)
        if cover is not None:
            assert (cover == 1).all(), (world, rows_all)
        if min(rows_all) >= 40:
            assert max(work) / min(work) < 1.02, (world, work)          # the shared step keeps the ranks balanced
    p8 = ring_schedule([6250] * 8, 1)
    assert p8[4] == (5, 0, 3125, 0, 6250) and ring_schedule([6250] * 8, 5)[4] == (1, 0, 6250, 3125, 6250)


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

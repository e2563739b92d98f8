#!/usr/bin/env python3
"""Phase timing of the row-sharded Gram with 2 virtual ranks on one GPU (same kernels as the multi-GPU run): pack, the two block
Grams of a rank (against its own planes and against the peer's), the mirror pass — through the snk_gram_shard_* phase calls."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

S = graft.load_package()
from snake_b200 import gram_sharded as GS  # noqa: E402

R, P = 6250, 181395
dev = torch.device("cuda", 0)
A = torch.randn(2 * R, P, device=dev, dtype=torch.float32)
peers = GS.LocalPeers(2 * R, P, 2, dev)
rows = [A[s.col0[s.rank]:s.col0[s.rank] + s.rows].contiguous() for s in peers.shards]


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


sh = peers.shards[0]
print("pack (6,250 x 181,395 fp32 -> bf16 planes)  ms", timed(lambda: sh.pack(rows[0])))
peers.shards[1].pack(rows[1])
torch.cuda.synchronize()
print("ring of rank 0 (2 block Grams + staged copy)  ms", timed(lambda: sh.ring(3)))
peers.shards[1].ring(3)
torch.cuda.synchronize()
print("mirror of rank 0 (copy + peer-read transpose)  ms", timed(lambda: sh.mirror()))
for cg in (1, 2):
    S.lib().snk_gram_config(cg)
    print("cta_group", cg, ": ring of rank 0 ms", timed(lambda: sh.ring(3)))
peers.free()

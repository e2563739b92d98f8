#!/usr/bin/env python3
"""Phase timing of the sharded Gram with 2 virtual ranks on one GPU (same kernels as the multi-GPU run)."""
import ctypes as C
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
L = S.lib()
for s in peers.shards:
    s.pack(A[s.col0[s.rank]:s.col0[s.rank] + s.rows].contiguous())
torch.cuda.synchronize()
sh = peers.shards[0]
other = peers.shards[1]
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def block(bplanes, col):
    S._check(L.snk_gram_block(sh.planes.ptr, sh.rows, C.c_void_p(bplanes), C.c_void_p(bplanes + sh.plane_bytes), R, P, 3, 0, 0,
                              sh.scratch.ptr, C.c_void_p(sh.Y.ptr.value + 4 * col), sh.K, st))


print("block vs own planes   ms", timed(lambda: block(sh.planes.ptr.value, 0)))
print("block vs other planes ms", timed(lambda: block(other.planes.ptr.value, R)))
S._check(L.snk_copy_async(sh.stage[1].ptr, other.planes.ptr, 2 * sh.plane_bytes, st)); torch.cuda.synchronize()
print("block vs staged copy  ms", timed(lambda: block(sh.stage[1].ptr.value, R)))
print("copy of one peer's planes (local) ms", timed(lambda: S._check(L.snk_copy_async(sh.stage[1].ptr, other.planes.ptr, 2 * sh.plane_bytes, st))))
print("pack ms", timed(lambda: sh.pack(A[:R].contiguous())))
for cg in (1, 2):
    L.snk_gram_config(cg)
    print("cta_group", cg, "block ms", timed(lambda: block(sh.planes.ptr.value, 0)))

"""Row-sharded Gram over the GPUs of one box (BASELINE config 5b).

Rank g owns rows_g of A (K_total x P).  G[rows_g, :] needs every other rank's rows, so this is the one
place on the path with a real exchange step:

  1. every rank packs its rows into bf16 planes (hi, lo) in IPC-exportable device memory;
  2. planes ring: at step i rank g multiplies its rows against the planes of rank (g+i) % R.  While the
     tensor cores work on step i, the planes of rank (g+i+1) % R are copied from that peer's memory over
     NVLink into the other half of a double buffer (cudaMemcpyAsync on a peer-mapped pointer, own copy
     stream) — the all-gather is never materialised as a separate phase;
        Y[rows_g, rows_p] = hi_g hi_p^T + hi_g lo_p^T + lo_g hi_p^T       (a finished block of G)
     G is symmetric, so the ring stops half way (i = 0 .. R/2, the own block on its upper-triangle tiles only,
     the step R/2 of an even R shared by the two ranks of the pair: `ring_schedule`);
  3. the blocks a rank did not compute are transposes of blocks its peers did: the transpose kernel reads
     Y[rows_p, rows_g] straight out of rank p's memory (peer loads over NVLink) — the transpose
     "all-to-all" is fused into that kernel.

The orchestration lives in the C library (csrc/gram_shard.cu: snk_gram_shard_*, one call per rank and Gram, device-side
barriers over peer memory); this module is the thin Python caller.  torch.distributed only carries the 192-byte IPC
handles at set-up.  `LocalPeers` emulates R ranks inside one process on one GPU with the same kernels (tests on a
single B200); `AllGatherGram` is the library-collective (NCCL) baseline the planes ring is measured against.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _check, _ptr, lib, DTYPE_F32, DTYPE_F64
from .shard import shard_range


def ring_schedule(rows_all, rank):
    """The blocks rank `rank` computes, in ring order (snk_gram_shard_schedule — the library's own arithmetic, no GPU):
    a list of (peer, a0, a1, b0, b1): its own rows [a0, a1) against rows [b0, b1) of `peer`.  Every other block of its row
    slab is the transpose of a block in a peer's list."""
    world = len(rows_all)
    n = world // 2 + 1
    arr = (C.c_int64 * world)(*[int(r) for r in rows_all])
    steps = C.c_int(0)
    a0, a1, b0, b1 = [(C.c_int64 * n)() for _ in range(4)]
    _check(lib().snk_gram_shard_schedule(arr, world, int(rank), C.byref(steps), a0, a1, b0, b1))
    return [((rank + i) % world, a0[i], a1[i], b0[i], b1[i]) for i in range(steps.value)]


HANDLE_BYTES = 192            # SNK_GRAM_SHARD_HANDLE_BYTES


class GramShard:
    """One rank's snk_gram_shard: planes of its own rows, a double buffer for peers' planes, its Y row block — all owned
    by the C library (csrc/gram_shard.cu).  Python only carries the IPC handles between the processes."""

    def __init__(self, rows_all, rank, P, device, splits=0):
        self.rows_all = [int(r) for r in rows_all]
        self.rank, self.world, self.P = int(rank), len(rows_all), int(P)
        self.rows, self.K = self.rows_all[rank], sum(self.rows_all)
        self.col0 = [sum(self.rows_all[:r]) for r in range(self.world)]
        self.device = torch.device(device)
        arr = (C.c_int64 * self.world)(*self.rows_all)
        self._g = C.c_void_p()
        _check(lib().snk_gram_shard_create(C.byref(self._g), arr, self.world, self.rank, self.P, int(splits),
                                           self.device.index or 0))
        self.G = torch.empty(self.rows, self.K, dtype=torch.float32, device=self.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def handle(self):
        h = (C.c_uint8 * HANDLE_BYTES)()
        _check(lib().snk_gram_shard_export_host(self._g, h))
        return bytes(h)

    def connect(self, handles):
        """handles: list of every rank's 192-byte handle (own entry ignored)"""
        blob = b"".join(handles)
        _check(lib().snk_gram_shard_connect_host(self._g, (C.c_uint8 * len(blob)).from_buffer_copy(blob)))

    def planes(self):
        """(hi pointer, lo pointer, pitch in elements) of this rank's bf16 planes, for a producer that writes them directly"""
        hi, lo, pitch = C.c_void_p(), C.c_void_p(), C.c_int64()
        _check(lib().snk_gram_shard_planes(self._g, C.byref(hi), C.byref(lo), C.byref(pitch)))
        return hi.value, lo.value, pitch.value

    def _a(self, A_rows):
        if A_rows is None:
            return None, 0
        assert tuple(A_rows.shape) == (self.rows, self.P)
        return _ptr(A_rows, device=self.device), {torch.float64: DTYPE_F64, torch.float32: DTYPE_F32}[A_rows.dtype]

    def run(self, A_rows, terms=3, block_k=0):
        """the whole sharded Gram for this rank, enqueued on the current stream: G[rows_rank, :] (an internal buffer, valid
        until the next run).  A_rows None = the planes were written by a producer."""
        a, dt = self._a(A_rows)
        _check(lib().snk_gram_shard_run(self._g, a, dt, int(terms), int(block_k), _ptr(self.G), self.K, self._stream()))
        return self.G

    # the phases on their own (virtual ranks in one process order them themselves)
    def pack(self, A_rows):
        a, dt = self._a(A_rows)
        _check(lib().snk_gram_shard_pack(self._g, a, dt, self._stream()))

    def ring(self, terms=3, block_k=0):
        _check(lib().snk_gram_shard_ring(self._g, int(terms), int(block_k), self._stream()))

    def mirror(self):
        _check(lib().snk_gram_shard_mirror(self._g, _ptr(self.G), self.K, self._stream()))
        return self.G

    def check(self):
        t = C.c_int(0)
        _check(lib().snk_gram_shard_status_host(self._g, C.byref(t)))

    def free(self):
        if getattr(self, "_g", None) and self._g.value:
            lib().snk_gram_shard_destroy(self._g)
            self._g = C.c_void_p()


class LocalPeers:
    """R virtual ranks in ONE process on ONE GPU — same kernels, local pointers (snk_gram_shard_connect_local); the phases
    are ordered by the host here because the virtual ranks share one stream.  For tests on a single B200."""

    def __init__(self, K_total, P, world, device, splits=0):
        rows = [shard_range(K_total, r, world)[1] - shard_range(K_total, r, world)[0] for r in range(world)]
        self.shards = [GramShard(rows, r, P, device, splits) for r in range(world)]
        arr = (C.c_void_p * world)(*[s._g.value for s in self.shards])
        for s in self.shards:
            _check(lib().snk_gram_shard_connect_local(s._g, arr))

    def gram(self, A, terms=3, block_k=0):
        for s in self.shards:
            lo = s.col0[s.rank]
            s.pack(A[lo:lo + s.rows].contiguous())
        torch.cuda.synchronize()                  # "barrier": every virtual rank's planes are packed
        for s in self.shards:
            s.ring(terms, block_k)
        torch.cuda.synchronize()
        out = [s.mirror() for s in self.shards]
        torch.cuda.synchronize()
        return torch.cat(out, 0)

    def free(self):
        for s in self.shards:
            s.free()


class DistributedGram:
    """Persistent multi-process sharded Gram (one process per GPU): buffers and peer mappings are set up once (the 192-byte
    IPC handles travel through torch.distributed — the only thing it is used for), run() is then ONE library call per rank:
    no host synchronisation, the phases are separated by device-side barriers over peer memory."""

    def __init__(self, rows_all, P, device, splits=0, group=None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device)
        self.shard = GramShard(rows_all, self.rank, P, self.device, splits)
        handles = [None] * self.world
        dist.all_gather_object(handles, self.shard.handle(), group=group)
        self.shard.connect(handles)

    def run(self, A_rows, terms=3, block_k=0):
        """G[rows_rank, :] (a view of an internal buffer, valid until the next run)."""
        return self.shard.run(A_rows, terms, block_k)

    def close(self):
        torch.cuda.synchronize(self.device)
        self.shard.check()
        dist.barrier(group=self.group)
        self.shard.free()


class AllGatherGram:
    """The same row-sharded Gram through library collectives — the baseline the planes ring is measured against
    (BASELINE config 5b names "NCCL all-gather"): NCCL all-gather of every rank's packed planes (materialised:
    world x 2 planes per GPU), the same blocks of the same schedule against the gathered planes, an NCCL all-to-all of
    the computed blocks to the ranks that need their transposes, and a local transpose.  Same kernels, same result
    bits as DistributedGram; the exchange is a separate phase here instead of running under the MMA main loop."""

    def __init__(self, rows_all, P, device, splits=0, group=None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.rows_all, self.P, self.splits = list(rows_all), int(P), splits
        self.rows, self.K = self.rows_all[self.rank], sum(self.rows_all)
        self.col0 = [sum(self.rows_all[:r]) for r in range(self.world)]
        self.device = torch.device(device)
        max_rows = max(self.rows_all)
        pb, pitch = C.c_size_t(0), C.c_int64(0)
        _check(lib().snk_gram_planes_layout(max_rows, self.P, C.byref(pb), C.byref(pitch)))
        self.plane_bytes, self.pitch = pb.value, pitch.value
        self.sched = [ring_schedule(self.rows_all, r) for r in range(self.world)]
        need = 256
        for (_, a0, a1, b0, b1) in self.sched[self.rank]:
            if a1 > a0 and b1 > b0:
                sb = C.c_size_t(0)
                _check(lib().snk_gram_block_scratch_bytes(a1 - a0, b1 - b0, self.P, splits, C.byref(sb)))
                need = max(need, sb.value)
        self.all_planes = torch.empty(self.world, 2 * self.plane_bytes, dtype=torch.uint8, device=self.device)
        self.scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        self.G = torch.empty(self.rows, self.K, dtype=torch.float32, device=self.device)

    def run(self, A_rows, terms=3, block_k=0):
        L = lib()
        assert tuple(A_rows.shape) == (self.rows, self.P)
        dt = {torch.float64: DTYPE_F64, torch.float32: DTYPE_F32}[A_rows.dtype]
        mine = self.all_planes[self.rank]
        row_bytes = 2 * self.pitch
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _check(L.snk_gram_pack_planes(_ptr(A_rows, device=self.device), dt, self.P, self.rows, C.c_void_p(mine.data_ptr()),
                                          C.c_void_p(mine.data_ptr() + self.plane_bytes), st))
            dist.all_gather_into_tensor(self.all_planes.view(-1), mine, group=self.group)
            send = [torch.empty(0, dtype=torch.float32, device=self.device) for _ in range(self.world)]
            for i, (p, a0, a1, b0, b1) in enumerate(self.sched[self.rank]):
                if a1 == a0 or b1 == b0:
                    continue
                a_hi = mine.data_ptr() + a0 * row_bytes
                b_hi = self.all_planes[p].data_ptr() + b0 * row_bytes
                blk = self.G[a0:a1, self.col0[p] + b0:self.col0[p] + b1]
                _check(L.snk_gram_block(C.c_void_p(a_hi), C.c_void_p(a_hi + self.plane_bytes), a1 - a0, C.c_void_p(b_hi),
                                        C.c_void_p(b_hi + self.plane_bytes), b1 - b0, self.P, terms, 1 if i == 0 else 0, block_k,
                                        self.splits, C.c_void_p(self.scratch.data_ptr()), C.c_void_p(blk.data_ptr()), self.K, st))
                if i > 0:
                    send[p] = blk.contiguous()
            # what rank q = rank - i computed against my rows comes back as its (a1-a0) x (b1-b0) block
            recv = [torch.empty(0, dtype=torch.float32, device=self.device) for _ in range(self.world)]
            for i in range(1, len(self.sched[self.rank])):
                q = (self.rank - i) % self.world
                _, a0, a1, b0, b1 = self.sched[q][i]
                recv[q] = torch.empty((a1 - a0) * (b1 - b0), dtype=torch.float32, device=self.device)
            dist.all_to_all(recv, [t.reshape(-1) for t in send], group=self.group)
            for i in range(1, len(self.sched[self.rank])):
                q = (self.rank - i) % self.world
                _, a0, a1, b0, b1 = self.sched[q][i]
                if a1 == a0 or b1 == b0:
                    continue
                dst = self.G[b0:b1, self.col0[q] + a0:self.col0[q] + a1]
                _check(L.snk_gram_transpose_block(C.c_void_p(recv[q].data_ptr()), b1 - b0, b1 - b0, a1 - a0,
                                                  C.c_void_p(dst.data_ptr()), self.K, st))
        return self.G

    def close(self):
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        self.all_planes = self.scratch = self.G = None


def gram_distributed(A_rows, rows_all, terms=3, block_k=0, splits=0, group=None):
    """One-shot G[rows_rank, :] for this rank.  A_rows: this rank's (rows, P) CUDA tensor."""
    dg = DistributedGram(rows_all, A_rows.shape[1], A_rows.device, splits, group)
    try:
        return dg.run(A_rows, terms, block_k).clone()
    finally:
        dg.close()

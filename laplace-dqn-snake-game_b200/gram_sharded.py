"""Row-sharded Gram over the GPUs of one box (BASELINE config 5b).

Rank g owns rows_g of A (K_total x P).  G[rows_g, :] needs every other rank's rows, so this is the one
place on the path with a real exchange step:

  1. every rank packs its rows into bf16 planes (hi, 2*lo) in IPC-exportable device memory;
  2. planes ring: at step i rank g multiplies its rows against the planes of rank (g+i) % R.  While the
     tensor cores work on step i, the planes of rank (g+i+1) % R are copied from that peer's memory over
     NVLink into the other half of a double buffer (cudaMemcpyAsync on a peer-mapped pointer, own copy
     stream) — the all-gather is never materialised as a separate phase;
        Y[rows_g, rows_p] = hi_g hi_p^T + hi_g (2 lo_p)^T
  3. the hi/lo cross terms are transposes of each other, so G = (Y + Y^T)/2: the symmetrise kernel reads
     the transposed block Y[rows_p, rows_g] straight out of rank p's memory (peer loads over NVLink) —
     the transpose "all-to-all" is fused into that kernel.

torch.distributed is only plumbing here (IPC-handle exchange, barriers); it works with gloo as well, which
is how the CPU test drives the schedule.  `LocalPeers` emulates R ranks inside one process on one GPU with
the same kernels (tests on a single B200).
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _check, _ptr, lib, DTYPE_F32, DTYPE_F64
from .shard import shard_range


def ring_schedule(rank, world):
    """Peers in the order rank multiplies against them: itself first, then around the ring."""
    return [(rank + i) % world for i in range(world)]


class _Buf:
    """cudaMalloc'ed (IPC-exportable) device buffer."""

    def __init__(self, nbytes):
        self.ptr = C.c_void_p()
        self.nbytes = int(nbytes)
        _check(lib().snk_ipc_alloc(C.byref(self.ptr), max(self.nbytes, 256)))

    def handle(self):
        h = (C.c_uint8 * 64)()
        _check(lib().snk_ipc_export(self.ptr, h))
        return bytes(h)

    def free(self):
        if self.ptr:
            lib().snk_ipc_free(self.ptr)
            self.ptr = C.c_void_p()


class GramShard:
    """One rank's state: planes of its own rows, a double buffer for peers' planes, its Y and G row blocks."""

    def __init__(self, rows_all, rank, P, device, splits=0):
        self.rows_all = list(rows_all)                     # rows owned by every rank
        self.rank, self.world, self.P = rank, len(rows_all), int(P)
        self.rows = self.rows_all[rank]
        self.K = sum(self.rows_all)
        self.col0 = [sum(self.rows_all[:r]) for r in range(self.world)]
        self.device = torch.device(device)
        self.splits = splits
        max_rows = max(self.rows_all)
        pb, pitch = C.c_size_t(0), C.c_int64(0)
        _check(lib().snk_gram_planes_layout(max_rows, self.P, C.byref(pb), C.byref(pitch)))
        self.plane_bytes, self.pitch = pb.value, pitch.value
        with torch.cuda.device(self.device):
            self.planes = _Buf(2 * self.plane_bytes)                       # [hi | lo2] of my rows (exported)
            self.stage = [_Buf(2 * self.plane_bytes) for _ in range(2)]    # peers' planes, double buffered
            self.Y = _Buf(self.rows * self.K * 4)                          # my row block of Y (exported)
            sb = C.c_size_t(0)
            _check(lib().snk_gram_block_scratch_bytes(self.rows, max_rows, self.P, splits, C.byref(sb)))
            self.scratch = _Buf(sb.value)
        self.G = torch.empty(self.rows, self.K, dtype=torch.float32, device=self.device)
        self.compute = torch.cuda.current_stream(self.device)
        self.copy = torch.cuda.Stream(self.device)

    def _p(self, buf, off=0):
        return C.c_void_p(buf.ptr.value + off)

    def pack(self, A_rows):
        assert tuple(A_rows.shape) == (self.rows, self.P)
        dt = {torch.float64: DTYPE_F64, torch.float32: DTYPE_F32}[A_rows.dtype]
        with torch.cuda.device(self.device):
            _check(lib().snk_gram_pack_planes(_ptr(A_rows, device=self.device), dt, self.P, self.rows, self.planes.ptr,
                                              self._p(self.planes, self.plane_bytes),
                                              C.c_void_p(self.compute.cuda_stream)))

    def free(self):
        for b in [self.planes, self.Y, self.scratch] + self.stage:
            b.free()


def run_ring(shard, peer_planes, terms=3, block_k=0):
    """Steps 2 of the module docstring for one rank.  peer_planes[r] = device pointer (int) to rank r's
    [hi | lo2] planes as visible from this process (own pointer for r == rank)."""
    L = lib()
    sched = ring_schedule(shard.rank, shard.world)
    cs, ks = shard.copy, shard.compute
    ev_copied = [torch.cuda.Event() for _ in sched]
    ev_used = [torch.cuda.Event() for _ in sched]
    with torch.cuda.device(shard.device):
        for i, p in enumerate(sched):
            # prefetch the NEXT peer's planes into the other staging buffer while this step computes
            if i + 1 < len(sched):
                nxt = sched[i + 1]
                if i >= 1:
                    cs.wait_event(ev_used[i - 1])          # that buffer was the B operand of step i-1
                nbytes = shard.plane_bytes + shard.rows_all[nxt] * shard.pitch * 2
                _check(L.snk_copy_async(shard.stage[(i + 1) % 2].ptr, C.c_void_p(peer_planes[nxt]), nbytes,
                                        C.c_void_p(cs.cuda_stream)))
                ev_copied[i + 1].record(cs)
            if i == 0:
                b_hi = shard.planes.ptr.value
            else:
                ks.wait_event(ev_copied[i])
                b_hi = shard.stage[i % 2].ptr.value
            b_lo = b_hi + shard.plane_bytes
            ycol = C.c_void_p(shard.Y.ptr.value + 4 * shard.col0[p])
            _check(L.snk_gram_block(shard.planes.ptr, shard.rows, C.c_void_p(b_hi), C.c_void_p(b_lo), shard.rows_all[p],
                                    shard.P, terms, block_k, shard.splits, shard.scratch.ptr, ycol, shard.K,
                                    C.c_void_p(ks.cuda_stream)))
            ev_used[i].record(ks)


def run_symmetrize(shard, peer_Y, terms=3):
    """Step 3: G[rows_g, rows_p] = (Y[rows_g, rows_p] + Y_p[rows_p, rows_g]^T)/2, Y_p read from peer memory."""
    L = lib()
    g = shard
    with torch.cuda.device(g.device):
        st = C.c_void_p(g.compute.cuda_stream)
        for p in range(g.world):
            y = C.c_void_p(g.Y.ptr.value + 4 * g.col0[p])
            out = C.c_void_p(g.G.data_ptr() + 4 * g.col0[p])
            if terms == 1:
                _check(L.snk_copy_async(C.c_void_p(g.G.data_ptr()), g.Y.ptr, g.rows * g.K * 4, st))
                break
            yt = C.c_void_p(peer_Y[p] + 4 * g.col0[g.rank])       # rank p's block (rows_p x rows_g), ld K
            _check(L.snk_gram_symmetrize_block(y, g.K, yt, g.K, g.rows, g.rows_all[p], out, g.K, st))
    return g.G


class LocalPeers:
    """R virtual ranks in ONE process on ONE GPU — same kernels, local pointers.  For tests on a single B200."""

    def __init__(self, K_total, P, world, device, splits=0):
        rows = [shard_range(K_total, r, world)[1] - shard_range(K_total, r, world)[0] for r in range(world)]
        self.shards = [GramShard(rows, r, P, device, splits) for r in range(world)]

    def gram(self, A, terms=3, block_k=0):
        for s in self.shards:
            lo = s.col0[s.rank]
            s.pack(A[lo:lo + s.rows].contiguous())
        planes = [s.planes.ptr.value for s in self.shards]
        Ys = [s.Y.ptr.value for s in self.shards]
        torch.cuda.synchronize()                  # "barrier": every virtual rank's planes are packed
        for s in self.shards:
            run_ring(s, planes, terms, block_k)
        torch.cuda.synchronize()
        out = [run_symmetrize(s, Ys, terms) for s in self.shards]
        torch.cuda.synchronize()
        return torch.cat(out, 0)

    def free(self):
        for s in self.shards:
            s.free()


class DistributedGram:
    """Persistent multi-process sharded Gram (one process per GPU): buffers and peer mappings are set up once
    (cudaMalloc + cudaIpc handle exchange through torch.distributed), run() can then be called repeatedly."""

    def __init__(self, rows_all, P, device, splits=0, group=None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device)
        self.shard = GramShard(rows_all, self.rank, P, self.device, splits)
        handles = [None] * self.world
        dist.all_gather_object(handles, (self.shard.planes.handle(), self.shard.Y.handle()), group=group)
        self.planes, self.Ys, self._opened = [], [], []
        for r, (hp, hy) in enumerate(handles):
            if r == self.rank:
                self.planes.append(self.shard.planes.ptr.value)
                self.Ys.append(self.shard.Y.ptr.value)
                continue
            pp, py = C.c_void_p(), C.c_void_p()
            with torch.cuda.device(self.device):
                _check(lib().snk_ipc_import((C.c_uint8 * 64).from_buffer_copy(hp), C.byref(pp)))
                _check(lib().snk_ipc_import((C.c_uint8 * 64).from_buffer_copy(hy), C.byref(py)))
            self._opened += [pp, py]
            self.planes.append(pp.value)
            self.Ys.append(py.value)

    def _barrier(self):
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)

    def run(self, A_rows, terms=3, block_k=0):
        """G[rows_rank, :] (a view of an internal buffer, valid until the next run)."""
        self.shard.pack(A_rows)
        self._barrier()                           # every rank's planes are packed
        run_ring(self.shard, self.planes, terms, block_k)
        self._barrier()                           # every rank's Y row block is complete
        G = run_symmetrize(self.shard, self.Ys, terms)
        self._barrier()                           # nobody still reads my Y / planes
        return G

    def close(self):
        self._barrier()
        for p in self._opened:
            lib().snk_ipc_close(p)
        self._opened = []
        self.shard.free()


class AllGatherGram:
    """The same row-sharded Gram through library collectives — the baseline the planes ring is measured against
    (BASELINE config 5b names "NCCL all-gather"): NCCL all-gather of every rank's packed planes (materialised:
    world x 2 planes per GPU), the block Grams against the gathered planes, an NCCL all-to-all of the transposed Y
    blocks, and the local symmetrise kernel.  Same kernels, same result bits as DistributedGram; the exchange is
    a separate phase here instead of running under the MMA main loop."""

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
        self.plane_bytes = pb.value
        sb = C.c_size_t(0)
        _check(lib().snk_gram_block_scratch_bytes(self.rows, max_rows, self.P, splits, C.byref(sb)))
        self.all_planes = torch.empty(self.world, 2 * self.plane_bytes, dtype=torch.uint8, device=self.device)
        self.scratch = torch.empty(max(sb.value, 256), dtype=torch.uint8, device=self.device)
        self.Y = torch.empty(self.rows, self.K, dtype=torch.float32, device=self.device)
        self.G = torch.empty(self.rows, self.K, dtype=torch.float32, device=self.device)

    def run(self, A_rows, terms=3, block_k=0):
        L = lib()
        assert tuple(A_rows.shape) == (self.rows, self.P)
        dt = {torch.float64: DTYPE_F64, torch.float32: DTYPE_F32}[A_rows.dtype]
        mine = self.all_planes[self.rank]
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _check(L.snk_gram_pack_planes(_ptr(A_rows, device=self.device), dt, self.P, self.rows, C.c_void_p(mine.data_ptr()),
                                          C.c_void_p(mine.data_ptr() + self.plane_bytes), st))
            dist.all_gather_into_tensor(self.all_planes.view(-1), mine, group=self.group)
            for p in ring_schedule(self.rank, self.world):
                b_hi = self.all_planes[p].data_ptr()
                _check(L.snk_gram_block(C.c_void_p(mine.data_ptr()), self.rows, C.c_void_p(b_hi), C.c_void_p(b_hi + self.plane_bytes),
                                        self.rows_all[p], self.P, terms, block_k, self.splits, C.c_void_p(self.scratch.data_ptr()),
                                        C.c_void_p(self.Y.data_ptr() + 4 * self.col0[p]), self.K, st))
            if terms == 1:
                self.G.copy_(self.Y)
                return self.G
            send = [self.Y[:, self.col0[p]:self.col0[p] + self.rows_all[p]].contiguous() for p in range(self.world)]
            recv = [torch.empty(self.rows_all[p], self.rows, dtype=torch.float32, device=self.device) for p in range(self.world)]
            dist.all_to_all(recv, send, group=self.group)
            for p in range(self.world):
                _check(L.snk_gram_symmetrize_block(C.c_void_p(self.Y.data_ptr() + 4 * self.col0[p]), self.K,
                                                   C.c_void_p(recv[p].data_ptr()), self.rows, self.rows, self.rows_all[p],
                                                   C.c_void_p(self.G.data_ptr() + 4 * self.col0[p]), self.K, st))
        return self.G

    def close(self):
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        self.all_planes = self.scratch = self.Y = self.G = None


def gram_distributed(A_rows, rows_all, terms=3, block_k=0, splits=0, group=None):
    """One-shot G[rows_rank, :] for this rank.  A_rows: this rank's (rows, P) CUDA tensor."""
    dg = DistributedGram(rows_all, A_rows.shape[1], A_rows.device, splits, group)
    try:
        return dg.run(A_rows, terms, block_k).clone()
    finally:
        dg.close()

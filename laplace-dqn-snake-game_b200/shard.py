"""Env sharding across the GPUs of one box: envs are independent, so rank g of G owns a
contiguous range of global env ids and nothing is exchanged on the step path.  Only timings / end-of-run
scalars are reduced, with torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """Contiguous [lo, hi) of global env ids owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def weak_range(envs_per_rank, rank):
    """Weak scaling: every rank owns `envs_per_rank` envs; global ids are rank-major."""
    return rank * envs_per_rank, (rank + 1) * envs_per_rank


def rank_seed(seed, rank):
    """Per-rank seed of the synthetic draw streams (distinct shards must not replay each other)."""
    return (int(seed) * 0x9E3779B1 + 0x85EBCA6B * (rank + 1)) & 0x7FFFFFFFFFFFFFFF


def max_over_ranks(x, device=None):
    """max of a host scalar over all ranks (1 rank: identity)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())

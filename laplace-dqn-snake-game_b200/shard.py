"""Env sharding across the GPUs of one box: envs are independent, so rank g of G owns a
contiguous range of global env ids and nothing is exchanged on the step path.  Only timings / end-of-run
scalars are reduced, with torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
import os

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


def parse_cpulist(text):
    """'0-3,8,10-11' (sysfs cpulist syntax) -> sorted list of CPU ids."""
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return sorted(cpus)


def bind_host_near_gpu(device_index, sysfs="/sys/bus/pci/devices"):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned host buffers of the
    host-buffer entry points (snk_host_alloc) are allocated: with one process per GPU the device<->host copies of
    all ranks then stay on their own socket's memory controllers and PCIe root instead of crossing the socket
    interconnect.  Returns a dict describing what was done (also when nothing could be done: single NUMA node,
    virtualised topology without numa_node, restricted cpuset)."""
    info = {"bound": False}
    try:
        p = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        info["pci"] = bus
        with open(os.path.join(sysfs, bus, "numa_node")) as f:
            info["numa_node"] = int(f.read().strip())
        with open(os.path.join(sysfs, bus, "local_cpulist")) as f:
            local = parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus = sorted(set(local) & allowed)
        if info["numa_node"] >= 0 and cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, cpus="%d-%d (%d)" % (cpus[0], cpus[-1], len(cpus)))
        else:
            info["reason"] = "single NUMA node or no narrower CPU set for this GPU"
    except Exception as e:  # noqa: BLE001 - affinity is an optimisation, never a failure
        info["reason"] = "%s: %s" % (type(e).__name__, e)
    return info

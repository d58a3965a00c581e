"""Index-range sharding of a batch across GPUs (one process per GPU).

The path has no exchange step: lanes are independent (the reference's only
"parallelism" is 4 independent SIMD lanes, include/ecsimd/bignum.h:101-102), so
a batch of n lanes is cut into `world` contiguous ranges, each rank runs the
same kernels on its own range, and nothing crosses NVLink.  torch.distributed
is used for the rendezvous, the start barrier and the reduction of a few host
scalars (max of the device time, sum of lanes, min of the parity verdicts)
only -- over the `gloo` backend on CPU tensors: no NCCL communicator is created
and no collective kernel runs on any GPU.
"""
import os

ALIGN = 128  # shard boundaries fall on multiples of 128 lanes (whole 4-lane packs, whole warps)


def shard_range(n, rank, world, align=ALIGN):
    """[lo, hi) of rank `rank`: contiguous, aligned to `align` lanes, sizes differ by less than 2*`align`."""
    assert 0 <= rank < world
    units = (n + align - 1) // align
    base, extra = divmod(units, world)
    lo_u = rank * base + min(rank, extra)
    hi_u = lo_u + base + (1 if rank < extra else 0)
    return min(lo_u * align, n), min(hi_u * align, n)


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_distributed(backend=None):
    """Rendezvous from the torchrun environment (MASTER_ADDR/MASTER_PORT/RANK/WORLD_SIZE); gloo unless told otherwise."""
    import torch.distributed as dist
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend or "gloo", rank=rank, world_size=world)
    return rank, world, local


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (the job's step time is its slowest rank's)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def min_over_ranks(value, device=None):
    """min of a python number over all ranks (a verdict that must hold on every rank)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def xor_over_ranks(words, device=None):
    """XOR-combine per-rank checksums (uint32 numpy array) -- order independent"""
    import numpy as np
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return words
    world = dist.get_world_size()
    t = torch.from_numpy(words.astype(np.int64)).to(device or "cpu")
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    acc = np.zeros_like(words)
    for o in out:
        acc ^= o.cpu().numpy().astype(np.uint32)
    return acc

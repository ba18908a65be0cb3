"""Multi-GPU plumbing for the forward path: one process per GPU, scenes sharded across ranks, no collective on the
data path (scenes are independent units).  The only communication is bookkeeping: a barrier around
timed regions and a MAX-reduction of per-rank device times / a SUM of per-rank scene counts.  Works with the `nccl`
backend on GPUs and with `gloo` on CPU (tests/test_dist_gloo.py)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None, device: torch.device | None = None):
    """(rank, world, local_rank) from the torchrun environment; initialises the process group when world > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kwargs = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kwargs)
    return rank, world, local_rank


def shard_scenes(n_scenes: int, rank: int, world: int):
    """Indices of the scenes rank `rank` owns: contiguous blocks, sizes differ by at most one, every scene exactly once."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_scenes, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def balance_by_voxels(voxel_counts, world: int):
    """Greedy longest-processing-time assignment of scenes to ranks so that per-rank voxel sums are even (ragged
    scene sizes are the weak-scaling risk named in SURVEY.md §8(e)).  Returns a list of index lists."""
    order = sorted(range(len(voxel_counts)), key=lambda i: -voxel_counts[i])
    loads, owned = [0] * world, [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        owned[r].append(i)
        loads[r] += voxel_counts[i]
    return [sorted(o) for o in owned]


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device=None) -> float:
    """Slowest rank's time (the job's time)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def job_throughput(units_this_rank: float, seconds_this_rank: float, device=None) -> float:
    """Whole-job throughput = units processed by all ranks / time of the slowest rank."""
    return sum_over_ranks(units_this_rank, device) / max_over_ranks(seconds_this_rank, device)

"""Data-parallel plumbing for the replica benchmark (inference shards by images; no data-path collective).

One process per GPU (``torch.distributed``: NCCL on the GPU box, gloo in the CPU tests).  The only communication is the
timing protocol of bench.py: a barrier on both sides of the timed region and a MAX all-reduce of the per-rank device time,
so the reported throughput is (images of all ranks) / (slowest rank's time).
"""
import os

import torch


def init_from_env(backend=None):
    """Initialise the default process group from RANK / WORLD_SIZE / MASTER_* (torchrun); returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            kw = {}
            if backend == "nccl":
                kw["device_id"] = torch.device("cuda", local)
            dist.init_process_group(backend, **kw)
    return rank, world, local


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(value, device=None):
    """MAX all-reduce of a python float (per-rank elapsed milliseconds)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    if device is None:
        device = "cuda" if (torch.cuda.is_available() and dist.get_backend() == "nccl") else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def replica_throughput(images_per_rank_per_step, steps, elapsed_ms_this_rank):
    """Whole-job images/s of N independent replicas: all ranks' images over the slowest rank's time."""
    import torch.distributed as dist
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    ms = max_over_ranks(elapsed_ms_this_rank)
    return world * images_per_rank_per_step * steps / (ms * 1e-3), ms

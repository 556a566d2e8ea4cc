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


class SyncExchange:
    """Symmetric-memory buffer for the fused SyncBatchNorm statistics exchange (csrc/syncbn.cu): every rank allocates the same
    layout -- ``max_slots`` data slots of [world][2][cmax] fp32 followed by ``max_slots`` flag slots of [world][32] u32 -- through
    ``torch.distributed._symmetric_memory`` (CUDA VMM / IPC handles exchanged at rendezvous), which maps all peers' buffers into this
    process; ``peers_dev`` is the device array of their base pointers.  A plan draws one slot per exchange at build time (the draw
    order is the launch order, identical on all ranks); the trainer bumps ``epoch`` once per step."""

    def __init__(self, dist, device, cmax, max_slots=1024):
        import torch.distributed._symmetric_memory as symm
        from . import lib as L
        lib = L.load()
        self.world, self.rank, self.cmax, self.max_slots = dist.get_world_size(), dist.get_rank(), int(cmax), int(max_slots)
        self.slot_floats = lib.segb200_syncbn_slot_floats(self.world, self.cmax)
        self.slot_flags = lib.segb200_syncbn_slot_flags(self.world)
        self.flags_base = self.max_slots * self.slot_floats                       # in 4-byte words from the buffer base
        nwords = self.max_slots * (self.slot_floats + self.slot_flags)
        group = dist.group.WORLD
        if hasattr(symm, "enable_symm_mem_for_group"):                             # needed by torch < 2.8, a deprecated no-op since
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                try:
                    symm.enable_symm_mem_for_group(group.group_name)
                except Exception:
                    pass
        self.buf = symm.empty(nwords, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group)
        self.peers_dev = int(self.hdl.buffer_ptrs_dev)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.next_slot = 0
        torch.cuda.synchronize(device)
        dist.barrier()                                                              # every buffer is zeroed before any peer writes

    def new_slot(self):
        s = self.next_slot
        if s >= self.max_slots:
            raise RuntimeError("segb200: SyncBatchNorm exchange slots exhausted (raise max_slots)")
        self.next_slot += 1
        return s * self.slot_floats, self.flags_base + s * self.slot_flags


def replica_throughput(images_per_rank_per_step, steps, elapsed_ms_this_rank):
    """Whole-job images/s of N independent replicas: all ranks' images over the slowest rank's time."""
    import torch.distributed as dist
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    ms = max_over_ranks(elapsed_ms_this_rank)
    return world * images_per_rank_per_step * steps / (ms * 1e-3), ms

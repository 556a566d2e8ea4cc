"""world_size-2 gloo test of the replica-benchmark plumbing (barrier + max-over-ranks + aggregate throughput)."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from segmentron_b200 import parallel
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    parallel.barrier()
    # rank 1 is the slow replica: the job's time is ITS time, the images are everybody's
    value, ms = parallel.replica_throughput(images_per_rank_per_step=8, steps=10, elapsed_ms_this_rank=100.0 * (rank + 1))
    parallel.barrier()
    out[rank] = (value, ms)
    import torch.distributed as dist
    dist.destroy_process_group()


def test_replica_throughput_two_ranks():
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    for rank in (0, 1):
        value, ms = out[rank]
        assert ms == 200.0
        assert abs(value - 2 * 8 * 10 / 0.2) < 1e-6


def test_single_process_is_identity():
    from segmentron_b200 import parallel
    v, ms = parallel.replica_throughput(8, 10, 50.0)
    assert ms == 50.0 and abs(v - 8 * 10 / 0.05) < 1e-9

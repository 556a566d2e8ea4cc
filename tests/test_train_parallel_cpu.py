"""world_size-2 gloo test of the data-parallel TRAINING path (gradient all-reduce + SyncBatchNorm):
each rank interprets its own launch list on its own half batch (tests/emulate_plan.py, fp64), runs the bucketed gradient
all-reduces and the SyncBN statistic all-reduces exactly where the GPU path runs them, and takes an SGD step.  Expected result
= the oracle on the CONCATENATED batch (batch statistics over all samples == SyncBatchNorm with equal per-rank counts) with
the DDP loss, i.e. the mean over ranks of each rank's own mean cross-entropy."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

MODEL, SEED, SHAPE = "deeplabv3plus_resnet101", 31, (2, 3, 65, 97)      # per-rank batch 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data(rank):
    g = torch.Generator().manual_seed(4000 + rank)
    x = torch.randn(*SHAPE, generator=g)
    t = torch.randint(-1, 19, (SHAPE[0], SHAPE[2], SHAPE[3]), generator=g)
    if rank == 1:
        t[:, :20] = -1                                    # unequal valid-pixel counts: per-rank mean losses differ from the global mean
    m = (torch.rand(SHAPE[0], 256, 1, 1, generator=g) > 0.1).double() / 0.9
    return x, t, m


class _FakeExchange:
    """stand-in for parallel.SyncExchange (needs CUDA symmetric memory): lets the plan builder emit the FUSED exchange steps
    (bn_finalize_sync / bn_bwd_finalize_sync), which the CPU interpreter executes as local sums + all_reduce(SUM) + finalize"""

    def __init__(self, world, rank):
        self.world, self.rank, self.cmax, self.peers_dev, self.n = world, rank, 2048, 0, 0
        self.epoch = torch.zeros(1, dtype=torch.int32)

    def new_slot(self):
        self.n += 1
        return (self.n - 1) * self.world * 2 * self.cmax, 10 ** 9 + (self.n - 1) * self.world * 32


def _worker(rank, world, port, out, fused=False):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(4)
    sys.path.insert(0, os.path.dirname(__file__))
    import emulate_plan as E
    from oracle import segref as R
    from segmentron_b200 import ops, parallel
    from segmentron_b200.train import DeepLabV3PlusTrainerB200, TrainPlan
    import torch.distributed as dist
    parallel.init_from_env("gloo")
    ops._PLAN_DRY_RUN = True
    TrainPlan.STAT_DTYPE = torch.float64
    P = R.build_params(MODEL, SEED)
    tr = DeepLabV3PlusTrainerB200(P.state_dict(), dtype=torch.float64, device="cpu", lr=0.02, bucket_mb=20)
    assert tr.world == 2 and tr.dist is not None
    if fused:
        tr.xchg = _FakeExchange(world, rank)
    x, t, m = _data(rank)
    st = tr.plan_for(x.shape)
    pl, S = st["plan"], tr.store
    kinds = [s.kind for s in pl.fwd + pl.bwd]
    if fused:                                                                       # no collective left on the compute path
        assert kinds.count("allreduce") == 0 and kinds.count("bn_finalize_sync") > 100 and kinds.count("bn_bwd_finalize_sync") > 100
        assert tr.xchg.n == kinds.count("bn_finalize_sync") + kinds.count("bn_bwd_finalize_sync")
    else:
        assert kinds.count("allreduce") > 200                                      # SyncBN collectives are in the list
    pl.x_in.copy_(x); pl.target.copy_(t)
    pl.masks["head.aspp.dropout"].copy_(m.reshape(SHAPE[0], 256))
    S.grad.zero_()
    E.gather_cast(S.master, S.idx16, S.w16); E.gather_cast(S.master, S.idx32, S.w32)
    for s in pl.fwd:
        E.run_step(s)
    pos0, works = 0, []                                   # mirrors DeepLabV3PlusTrainerB200.forward_backward
    for pos, lo, hi in st["buckets"]:
        for s in pl.bwd[pos0:pos]:
            E.run_step(s)
        pos0 = pos
        works.append(dist.all_reduce(S.grad[lo:hi], async_op=True))
    for s in pl.bwd[pos0:]:
        E.run_step(s)
    for w in works:
        w.wait()
    grads = {k: v.clone() for k, v in tr.store.grads().items()}       # SUM over ranks
    E.sgd(tr)
    out[rank] = dict(loss=float(pl.out3[0]), grads=grads if rank == 0 else None,
                     w=tr.state_dict()["head.block.2.weight"], rm=tr.state_dict()["encoder.bn1.running_mean"], nb=len(st["buckets"]))
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("fused", [False, True], ids=["nccl_form", "fused_exchange_form"])
def test_two_rank_training_step_matches_oracle_on_the_joint_batch(fused):
    import torch.nn.functional as F
    from oracle import segref as R
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out, fused), nprocs=2, join=True)
    assert out[0]["nb"] >= 3
    # oracle: joint batch, DDP loss
    P = R.build_params(MODEL, SEED).to(dtype=torch.float64)
    P.frozen = True
    (x0, t0, m0), (x1, t1, m1) = _data(0), _data(1)
    P.dropout_masks["head.aspp.dropout"] = torch.cat([m0, m1])
    names = R.trainable(P)
    for k in names:
        P.t[k] = P.t[k].detach().clone().requires_grad_(True)
    before = P.t["head.block.2.weight"].detach().clone()
    P.training = True
    o = R.deeplabv3plus(P, torch.cat([x0, x1]).double(), nclass=19, **R.MODELS[MODEL])
    l0 = F.cross_entropy(o[:2], t0, ignore_index=-1)
    l1 = F.cross_entropy(o[2:], t1, ignore_index=-1)
    (0.5 * (l0 + l1)).backward()
    assert abs(out[0]["loss"] - float(l0)) < 1e-6 and abs(out[1]["loss"] - float(l1)) < 1e-6
    worst = ("", 0.0)
    for k in names:
        e = float((0.5 * out[0]["grads"][k].double() - P.t[k].grad).norm() / (P.t[k].grad.norm() + 1e-12))
        if e > worst[1]:
            worst = (k, e)
    assert worst[1] < 1e-5, worst
    assert torch.equal(out[0]["w"], out[1]["w"]), "replicas diverged after the step"
    ref_w = before - 0.2 * (P.t["head.block.2.weight"].grad + 1e-4 * before)
    assert float((out[0]["w"].double() - ref_w).norm() / ref_w.norm()) < 1e-6
    assert torch.allclose(out[0]["rm"].double(), P.t["encoder.bn1.running_mean"].detach(), atol=1e-6)

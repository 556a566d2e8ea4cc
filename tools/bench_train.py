"""BASELINE.json configs[2]: DeepLabv3+ / ResNet101 bf16 TRAINING at 1025x2049, per-GPU batch 4 -- images/sec of a full
iteration (forward, CrossEntropy(ignore -1), backward, SGD step; gradient all-reduce when launched under torchrun) through the
segb200 training engine, with the reference's own recipe timed beside it on the same GPU: the oracle port (the same torch ops
the reference executes) with fp32 master weights, torch.autocast(bf16), torch.optim.SGD, cudnn.benchmark -- NCHW and
channels_last, the faster one counts.

    python tools/bench_train.py [--steps K] [--warmup W] [--batch B] [--height H] [--width W] [--no-ref] [--kernels out.tsv]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py     (weak scaling)

One JSON line on stdout (rank 0)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
MODEL = "deeplabv3plus_resnet101"          # --backbone xception65 switches to deeplabv3plus_xception65


def ref_train_leg(P, x, target, fmt, iters):
    """the reference recipe: fp32 params, autocast(bf16) forward+loss, backward, SGD(momentum 0.9, wd 1e-4, head lr x10)"""
    import torch.nn.functional as F
    from oracle import segref as R
    Pg = P.to("cuda", torch.float32)
    Pg.frozen, Pg.training = True, True
    names = R.trainable(Pg)
    for k in names:
        v = Pg.t[k]
        if fmt == "channels_last" and v.dim() == 4:
            v = v.contiguous(memory_format=torch.channels_last)
        Pg.t[k] = v.detach().clone().requires_grad_(True)
    enc = [Pg.t[k] for k in names if k.startswith("encoder.")]
    head = [Pg.t[k] for k in names if not k.startswith("encoder.")]
    hr = MODEL.startswith("hrnet")             # the HRNet YAML: lr 0.01, no decoder factor, BatchNorm momentum 0.01
    Pg.bn_momentum = 0.01 if hr else 0.1
    opt = torch.optim.SGD([{"params": enc, "lr": 0.01 if hr else 0.02}, {"params": head, "lr": 0.01 if hr else 0.2}], lr=0.02,
                          momentum=0.9, weight_decay=1e-4)
    xb = x.contiguous(memory_format=torch.channels_last) if fmt == "channels_last" else x

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if MODEL == "danet_resnet101":            # three outputs, summed cross-entropy (solver/loss.py:31-36)
                loss = sum(F.cross_entropy(o.float(), target, ignore_index=-1) for o in R.danet(Pg, xb, nclass=19))
            else:
                out = R.hrnet_seg(Pg, xb, nclass=19) if hr else (R.ccnet(Pg, xb, nclass=19) if MODEL == "ccnet_resnet101" else
                                                                R.deeplabv3plus(Pg, xb, nclass=19, **R.MODELS[MODEL]))
                loss = F.cross_entropy(out.float(), target, ignore_index=-1)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--height", type=int, default=1025)
    ap.add_argument("--width", type=int, default=2049)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--kernels", default=None)
    ap.add_argument("--backbone", default="resnet101", choices=["resnet101", "xception65", "mobilenet_v2", "hrnet", "danet", "ccnet"])
    ap.add_argument("--same-data", action="store_true", help="every rank trains on rank 0's batch (N-GPU result must equal the 1-GPU one)")
    ap.add_argument("--no-dropout", action="store_true")
    ap.add_argument("--graph", action="store_true", help="single GPU: replay the step as one CUDA graph")
    ap.add_argument("--bf16-grads", action="store_true", help="all-reduce the gradient buckets in bf16 (opt-in; the reference's DDP moves fp32)")
    ap.add_argument("--cpu-baseline", action="store_true", help="also time ONE training iteration of the oracle port on the host cores (batch 1)")
    args = ap.parse_args()
    global MODEL
    MODEL = {"hrnet": "hrnet_w18_small_v1", "danet": "danet_resnet101", "ccnet": "ccnet_resnet101"}.get(args.backbone, "deeplabv3plus_" + args.backbone)
    if args.backbone == "hrnet" and (args.height, args.width) == (1025, 2049):
        args.height, args.width = 1024, 2048         # HRNet needs multiples of 32 (nearest up-sampling + add in the fuse layers)
    if args.backbone == "danet" and (args.height, args.width) == (1025, 2049):
        args.height, args.width, args.batch = 768, 768, min(args.batch, 2)   # the DANet YAML's training crop; attention is O(N^2)
    import __graft_entry__ as ge
    ge.build()
    from oracle import segref as R          # parameter generator + the reference leg only
    from segmentron_b200 import parallel
    from segmentron_b200.train import CCNetTrainerB200, DANetTrainerB200, DeepLabV3PlusTrainerB200, HRNetTrainerB200
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    rank, world, local = parallel.init_from_env("nccl")
    torch.backends.cudnn.benchmark = True
    P = R.build_params(MODEL, 0)
    shape = (args.batch, 3, args.height, args.width)
    g = torch.Generator().manual_seed(1024 + (0 if args.same_data else rank))
    x = torch.randn(*shape, generator=g).cuda()
    target = torch.randint(-1, 19, (shape[0], shape[2], shape[3]), generator=g).cuda()
    if args.backbone == "hrnet":
        tr = HRNetTrainerB200(P.state_dict(), dtype=torch.bfloat16, lr=0.01)
    elif args.backbone in ("danet", "ccnet"):
        tr = (DANetTrainerB200 if args.backbone == "danet" else CCNetTrainerB200)(P.state_dict(), dtype=torch.bfloat16, lr=0.02,
                                                                                 dropout=not args.no_dropout)
    else:
        tr = DeepLabV3PlusTrainerB200(P.state_dict(), backbone=args.backbone, dtype=torch.bfloat16, lr=0.02, dropout=not args.no_dropout,
                                      grad_comm_dtype=torch.bfloat16 if args.bf16_grads else torch.float32, cuda_graph=args.graph)
    losses = [float(tr.step(x, target)) for _ in range(max(args.warmup, 3))]
    parallel.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        tr.step(x, target)
    e1.record()
    torch.cuda.synchronize()
    ms = parallel.max_over_ranks(e0.elapsed_time(e1)) / args.steps
    parallel.barrier()
    out = None
    if rank == 0:
        pl = tr.plan_for(shape)["plan"]
        # per-launch timing replays the launch list on THIS rank only: not possible when the list contains collectives
        rows = []
        if world == 1:
            pl.run_timed()
            rows = pl.run_timed()
        agg = {}
        for st, t in rows:
            a = agg.setdefault(st.kind, dict(ms=0.0, n=0, flops=0.0))
            a["ms"] += t; a["n"] += 1; a["flops"] += st.info.get("flops", 0.0)
        if args.kernels and rows:
            with open(args.kernels, "w") as f:
                f.write("phase\tkind\tms\tGFLOP\tTFLOP/s\tdesc\n")
                nf = len(pl.fwd)
                for i, (st, t) in enumerate(rows):
                    fl = st.info.get("flops", 0.0)
                    desc = {k: v for k, v in st.info.items() if isinstance(v, (int, float, str, bool)) and k != "flops"}
                    for key in ("x", "y", "dy", "z"):
                        if torch.is_tensor(st.info.get(key)):
                            desc[key] = tuple(st.info[key].shape)
                    f.write(f"{'fwd' if i < nf else 'bwd'}\t{st.kind}\t{t:.4f}\t{fl / 1e9:.2f}\t{fl / max(t, 1e-6) / 1e9:.1f}\t{desc}\n")
        tot = sum(a["ms"] for a in agg.values())
        mm = {k: agg[k] for k in ("conv", "wgrad") if k in agg and agg[k]["ms"] > 0}
        out = {"config": "c3_train" if args.backbone == "resnet101" else "train_" + args.backbone,
               "what": f"{MODEL} bf16 training step (fwd + CE loss + bwd + SGD) at "
               f"{args.height}x{args.width}, per-GPU batch {args.batch}", "n_gpus": world, "segb200_img_s": world * args.batch / (ms * 1e-3),
               "segb200_ms_per_step": ms, "launches_per_step": tr.n_launches(shape), "loss_first_steps": losses,
               "per_kind_ms": {k: round(v["ms"], 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])},
               "sum_kernel_ms": tot,
               "tensor_core_tflops": {k: v["flops"] / (v["ms"] * 1e-3) / 1e12 for k, v in mm.items()},
               "activation_GB": pl.act_bytes / 1e9, "grad_pool_GB": pl.pool_bytes / 1e9,
               "sync_bn": bool(world > 1), "allreduce_buckets": len(tr.plan_for(shape)["buckets"]) if world > 1 else 0,
               "fused_syncbn_exchange": bool(getattr(tr, "xchg", None) is not None),
               "nccl_collectives_per_step": tr.collectives_per_step(shape),
               "param_digest": float(tr.store.master.double().abs().sum())}
    if not args.no_ref and world == 1:
        del tr
        torch.cuda.empty_cache()
        ref = {}
        for fmt in ("nchw", "channels_last"):
            try:
                ref[fmt] = args.batch / (ref_train_leg(P, x, target, fmt, 3) * 1e-3)
            except Exception as ex:                                   # noqa: BLE001
                ref[fmt] = f"failed: {type(ex).__name__}: {str(ex)[:120]}"
            torch.cuda.empty_cache()
        nums = [v for v in ref.values() if isinstance(v, float)]
        out["ref_cudnn_img_s"] = max(nums) if nums else None
        out["ref_by_layout"] = ref
        out["speedup"] = out["segb200_img_s"] / max(nums) if nums else None
        out["ref_what"] = "oracle port (the reference's torch ops) fp32 params + autocast(bf16) + torch SGD, cudnn.benchmark"
    if args.cpu_baseline and rank == 0:
        import time
        try:                                       # batch 2: train-mode BatchNorm of the image-pooling branch needs > 1 sample
            Pc = R.build_params(MODEL, 0)
            xc, tc = x[:2].cpu(), target[:2].cpu()
            t0 = time.perf_counter()
            R.loss_and_grads(MODEL, Pc, xc, tc)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": 2.0 / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": f"1 training iteration (fwd + CE + bwd, fp32) of the oracle port on 2x3x{args.height}x{args.width}"}
        except Exception as ex:                    # noqa: BLE001
            out["cpu_baseline"] = {"error": f"{type(ex).__name__}: {str(ex)[:160]}"}
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

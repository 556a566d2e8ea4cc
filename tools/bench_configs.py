"""Throughput of the other BASELINE.json configs (inference) through the segb200 engines, with the oracle port (the same
torch ops the reference runs, cuDNN/cuBLAS bf16|fp16 eager, faster of NCHW / channels_last) timed beside them.

    python tools/bench_configs.py [c3 c4 c5 c1]

Prints one JSON line per config.  (The headline config C2 is bench.py.)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import segref as R  # noqa: E402  (weights generator + the cuDNN reference leg)
from segmentron_b200 import engine as E  # noqa: E402

CONFIGS = {
    # name: (oracle model, engine factory, input shape, dtype, what)
    "c3": ("deeplabv3plus_resnet101", lambda sd, dt: E.DeepLabV3PlusB200(sd, backbone="resnet101", dtype=dt), (4, 3, 1025, 2049),
           torch.bfloat16, "DeepLabv3+/ResNet101 bf16 INFERENCE forward at 1025x2049, batch 4 (config 3 is training; forward only here)"),
    "c4": ("danet_resnet101", lambda sd, dt: E.DANetB200(sd, dtype=dt), (2, 3, 1024, 2048), torch.bfloat16,
           "DANet/ResNet101 OS8 multi-grid, PAM N=32768 + CAM, 1024x2048, batch 2"),
    "c5": ("hrnet_w18_small_v1", lambda sd, dt: E.HRNetB200(sd, dtype=dt), (16, 3, 1024, 2048), torch.float16,
           "HRNet-w18-small-v1 fp16 inference at 1024x2048, batch 16"),
    "c1": ("deeplabv3plus_mobilenet_v2", lambda sd, dt: E.DeepLabV3PlusB200(sd, backbone="mobilenet_v2", use_aspp=False,
                                                                              use_decoder=False, dtype=dt), (1, 3, 512, 1024),
           torch.bfloat16, "DeepLabv3+/MobileNetV2 512x1024 batch 1 (config 1 shape) on the GPU engine"),
}


def time_it(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    no_ref = "--no-ref" in sys.argv
    kinds = "--kinds" in sys.argv                    # also print the per-kernel-kind time table of each plan
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c3", "c4", "c5", "c1"]
    for a in sys.argv[1:]:                          # e.g. --opt:gemm_kgroup=9 --opt:gemm_kgroup_kb=48  (segb200_set_option knobs)
        if a.startswith("--opt:"):
            k, v = a[6:].split("=")
            from segmentron_b200 import lib as L
            L.check(L.load().segb200_set_option(k.encode(), int(v)), "set_option")
    torch.backends.cudnn.benchmark = True
    for name in names:
        model, factory, shape, dt, what = CONFIGS[name]
        P = R.build_params(model, 0)
        x = torch.randn(*shape, generator=torch.Generator().manual_seed(1024)).cuda()
        eng = factory(P.state_dict(), dt)
        eng(x)
        ms = time_it(lambda: eng(x, copy_input=False), 10)
        ours = shape[0] / (ms * 1e-3)
        per_kind = None
        if kinds:
            pl = eng.plan_for(x)["plan"]
            pl.run_timed()
            agg = {}
            for m, t in pl.run_timed():
                a = agg.setdefault(m["kind"], [0.0, 0])
                a[0] += t; a[1] += 1
            per_kind = {k: [round(v[0], 3), v[1]] for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}
        del eng
        torch.cuda.empty_cache()
        ref = {}
        for fmt in (() if no_ref else ("nchw", "channels_last")):
            try:
                Pg = P.to("cuda", dt)
                xb = x.to(dt)
                if fmt == "channels_last":
                    xb = xb.contiguous(memory_format=torch.channels_last)
                    for k, v in Pg.t.items():
                        if v.dim() == 4:
                            Pg.t[k] = v.contiguous(memory_format=torch.channels_last)
                ref[fmt] = shape[0] / (time_it(lambda: R.forward(model, Pg, xb), 3) * 1e-3)
                del Pg, xb
            except Exception as ex:                              # e.g. the materialised N x N attention does not fit
                ref[fmt] = f"failed: {type(ex).__name__}: {str(ex)[:80]}"
            torch.cuda.empty_cache()
        nums = [v for v in ref.values() if isinstance(v, float)]
        best = max(nums) if nums else None
        print(json.dumps({"config": name, "what": what, "shape": shape, "dtype": str(dt).split(".")[-1], "segb200_img_s": ours,
                          "segb200_ms": ms, "ref_cudnn_img_s": best, "ref_by_layout": ref,
                          "speedup": (ours / best) if best else None, "per_kind_ms": per_kind}), flush=True)


if __name__ == "__main__":
    main()

"""A/B of the single-CTA and the CTA-pair GEMM kernels on the hot shapes (CUDA events, L2 flushed): correctness vs each other
first (max |diff|), then TFLOP/s.  GPU box only; every launch is bounded by the kernels' mbarrier watchdog."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmentron_b200 import fold, lib as L, ops  # noqa: E402

dt = torch.bfloat16
lib = L.load()
SHAPES = [  # n, h, w, cin, cout, k, dil, residual
    (8, 65, 129, 728, 728, 1, 1, True), (8, 65, 129, 728, 728, 1, 1, False), (8, 65, 129, 1024, 1024, 1, 1, False), (8, 65, 129, 2048, 256, 1, 1, False),
    (8, 65, 129, 728, 1024, 1, 1, False), (8, 65, 129, 1536, 2048, 1, 1, False), (4, 65, 129, 256, 256, 3, 1, False),
    (4, 65, 129, 512, 512, 3, 2, False), (4, 65, 129, 1024, 256, 1, 1, False), (4, 65, 129, 256, 1024, 1, 1, True),
    (8, 129, 257, 256, 728, 1, 1, False), (2, 513, 1025, 128, 128, 1, 1, False),
]
flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device="cuda")
for (n, h, w, cin, cout, k, dil, use_res) in SHAPES:
    pad = dil * (k - 1) // 2
    x = torch.randn(n, h, w, cin, device="cuda").to(dt)
    wt = (torch.randn(cout, cin, k, k, device="cuda") / math.sqrt(cin * k * k)).to(dt)
    wpk = fold.pack_conv_weight(wt, dt)
    sc, sh = torch.rand(cout, device="cuda") + 0.5, torch.randn(cout, device="cuda") * 0.1
    res = torch.randn(n, h, w, cout, device="cuda").to(dt) if use_res else None
    ys, rec = [], {"shape": f"{k}x{k}d{dil} {cin}->{cout} @{n}x{h}x{w}" + (" +res" if use_res else "")}
    for mode in (0, 1, 2):                                  # 0: single CTA; 1: CTA pair, one k-block per observation; 2: CTA pair, paired k-blocks
        L.check(lib.segb200_set_option(b"gemm_2cta", 1 if mode else 0))
        L.check(lib.segb200_set_option(b"gemm_mma_pairs", 1 if mode == 2 else 0))
        y = torch.zeros(n, h, w, cout, device="cuda", dtype=dt)
        fn = lambda: ops.conv_gemm(x, wpk, y, cin=cin, cout=cout, kh=k, kw=k, dilation=dil, pad_t=pad, pad_l=pad, scale=sc, shift=sh,  # noqa: E731
                                   act="relu", residual=res)
        try:
            fn(); fn()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            rec[f"mode{mode}"] = f"FAILED {e}"
            print(json.dumps(rec), flush=True)
            sys.exit(0)
        ms = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = sorted(ms)[len(ms) // 2]
        rec[f"mode{mode}_ms"] = round(t, 4)
        rec[f"mode{mode}_TFLOPs"] = round(2.0 * n * h * w * cin * cout * k * k / t / 1e9, 1)
        ys.append(y.float())
    rec["max_abs_diff"] = max(float((ys[0] - ys[1]).abs().max()), float((ys[0] - ys[2]).abs().max()))
    rec["speedup_pair"] = round(rec["mode0_ms"] / rec["mode1_ms"], 3)
    rec["speedup_pair_paired_kblocks"] = round(rec["mode0_ms"] / rec["mode2_ms"], 3)
    print(json.dumps(rec), flush=True)
L.check(lib.segb200_set_option(b"gemm_2cta", 0))
L.check(lib.segb200_set_option(b"gemm_mma_pairs", 1))

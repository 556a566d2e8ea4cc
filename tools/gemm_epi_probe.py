"""A/B of the epilogue-group threshold (segb200_set_option("gemm_epi2_maxk", K)) on the mid-K shapes; CUDA events, L2 flushed."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmentron_b200 import fold, lib as L, ops  # noqa: E402

dt = torch.bfloat16
lib = L.load()
SHAPES = [(8, 65, 129, 728, 728, 1, 1, True), (8, 65, 129, 728, 728, 1, 1, False), (8, 65, 129, 1024, 1024, 1, 1, False),
          (4, 65, 129, 1024, 256, 1, 1, False), (4, 65, 129, 512, 2048, 1, 1, True), (8, 65, 129, 2048, 256, 1, 1, False),
          (4, 65, 129, 256, 256, 3, 1, False), (8, 129, 257, 256, 728, 1, 1, False)]
flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device="cuda")
for (n, h, w, cin, cout, k, dil, use_res) in SHAPES:
    pad = dil * (k - 1) // 2
    x = torch.randn(n, h, w, cin, device="cuda").to(dt)
    wpk = fold.pack_conv_weight((torch.randn(cout, cin, k, k, device="cuda") / math.sqrt(cin * k * k)).to(dt), dt)
    sc, sh = torch.rand(cout, device="cuda") + 0.5, torch.randn(cout, device="cuda") * 0.1
    res = torch.randn(n, h, w, cout, device="cuda").to(dt) if use_res else None
    rec = {"shape": f"{k}x{k}d{dil} {cin}->{cout} @{n}x{h}x{w}" + (" +res" if use_res else "")}
    for maxk in (512, 1024, 4096):
        L.check(lib.segb200_set_option(b"gemm_epi2_maxk", maxk))
        y = torch.zeros(n, h, w, cout, device="cuda", dtype=dt)
        fn = lambda: ops.conv_gemm(x, wpk, y, cin=cin, cout=cout, kh=k, kw=k, dilation=dil, pad_t=pad, pad_l=pad, scale=sc, shift=sh,  # noqa: E731
                                   act="relu", residual=res)
        fn(); fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(9):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = sorted(ms)[len(ms) // 2]
        rec[f"maxk{maxk}_ms"] = round(t, 4)
        rec[f"maxk{maxk}_TFLOPs"] = round(2.0 * n * h * w * cin * cout * k * k / t / 1e9, 1)
    print(json.dumps(rec), flush=True)
L.check(lib.segb200_set_option(b"gemm_epi2_maxk", 512))

"""A/B of the two depthwise weight-gradient kernels (default segb200_dw_wgrad vs opt-in segb200_dw_wgrad_v2) on the shapes of the
Xception65 / DeepLabv3+ training step: error of each against autograd and CUDA-event time (L2 flushed between launches).
(GPU box only.)   python tools/dw_wgrad_probe.py > gpurun_out/dw_wgrad_probe.jsonl"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmentron_b200 import train_ops as T  # noqa: E402

# (n, h, w, c, dilation, pre_relu): entry flow, middle flow (48 of the 68 instances), exit flow, ASPP, decoder
SHAPES = [(4, 513, 1025, 128, 1, True), (4, 257, 513, 256, 1, True), (4, 65, 129, 728, 1, True), (4, 65, 129, 1536, 2, True),
          (4, 65, 129, 2048, 6, False), (4, 65, 129, 2048, 18, False), (4, 257, 513, 304, 1, False)]
flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=5):
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for (n, h, w, c, d, pre) in SHAPES:
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, h, w, c, generator=g).bfloat16().cuda()
    dy = torch.randn(n, h, w, c, generator=g).bfloat16().cuda()
    row = {"shape": (n, h, w, c), "dilation": d, "pre_relu": pre}
    ref = None
    if n * h * w * c < 3e8:
        wr = torch.zeros(c, 1, 3, 3, device="cuda", requires_grad=True)
        xin = x.float().permute(0, 3, 1, 2)
        F.conv2d(F.relu(xin) if pre else xin, wr, None, 1, d, d, groups=c).backward(dy.float().permute(0, 3, 1, 2))
        ref = wr.grad.reshape(c, 9)
    for variant in (1, 2):
        dw = torch.zeros(c, 9, device="cuda")
        try:
            T.dw_wgrad(x, dy, dw, dilation=d, pre_relu=pre, variant=variant)
            torch.cuda.synchronize()
            if ref is not None:
                row[f"v{variant}_rel_l2"] = float((dw - ref).norm() / ref.norm())
            ms = timed(lambda: T.dw_wgrad(x, dy, dw, dilation=d, pre_relu=pre, variant=variant))
            row[f"v{variant}_ms"] = ms
            row[f"v{variant}_GBps"] = 2 * x.numel() * 2 / ms / 1e6            # algorithmic bytes: x and dy read once
        except Exception as e:  # noqa: BLE001
            row[f"v{variant}_error"] = str(e)[:200]
    print(json.dumps(row), flush=True)

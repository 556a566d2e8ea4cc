"""Depthwise 3x3 micro-benchmark over the kernel's tuning knobs (segb200_set_option: dw_v8, dw_persistent, dw_ring_slots) on the
headline model's layer shapes.  Buffers rotate through > 126 MB so that no launch finds its input in L2.  One JSON line per
(shape, configuration): achieved GB/s of algorithmic bytes (in + out) and the fraction of the measured HBM peak.
    python tools/dw_sweep.py [--quick]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from segmentron_b200 import fold, lib, ops  # noqa: E402

L = lib.load()
dt = torch.bfloat16
peak = 6570.9
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

SHAPES = [(8, 65, 129, 728, 1, 1), (8, 513, 1025, 128, 1, 1), (8, 257, 513, 256, 1, 1), (8, 257, 513, 304, 1, 1), (8, 129, 257, 728, 1, 1),
          (8, 513, 1025, 64, 1, 1), (8, 513, 1025, 128, 2, 1), (8, 65, 129, 2048, 1, 12), (8, 65, 129, 1536, 1, 2)]
if "--quick" in sys.argv:
    SHAPES = SHAPES[:3]
CONFIGS = [dict(dw_v8=1, dw_cols2=0), dict(dw_cols2=0), dict(), dict(dw_ring_slots=8), dict(dw_ring_slots=6), dict(dw_persistent=1, dw_cols2=0)]
if "--cols2" in sys.argv:                      # round-2 A/B of the two-column kernel only
    CONFIGS = [dict(dw_cols2=0), dict(), dict(dw_cw5=1), dict(dw_ring_slots=8)]
    SHAPES = SHAPES[:6]
KNOBS = {"dw_v8": 0, "dw_persistent": 0, "dw_ring_slots": 0, "dw_cols2": 1, "dw_cw5": 0}


def bench(n, h, w, c, stride, dil, reps=30):
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    nbuf = max(2, int(400e6 // (n * h * w * c * 2)) + 1)
    xs = [torch.randn(n, h, w, c, device="cuda").to(dt) for _ in range(nbuf)]
    ys = [torch.empty(n, ho, wo, c, device="cuda", dtype=dt) for _ in range(nbuf)]
    wt = fold.pack_dw_weight(torch.randn(c, 1, 3, 3, device="cuda"), torch.ones(c, device="cuda"))
    sh = torch.zeros(c, device="cuda")
    for i in range(3):
        ops.dwconv3x3(xs[i % nbuf], wt, ys[i % nbuf], stride=stride, dilation=dil, shift=sh, pre_relu=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        ops.dwconv3x3(xs[i % nbuf], wt, ys[i % nbuf], stride=stride, dilation=dil, shift=sh, pre_relu=True)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    byts = 2.0 * (n * h * w * c + n * ho * wo * c)
    return us, byts / us / 1e3


for shp in SHAPES:
    for cfg in CONFIGS:
        for k, dflt in KNOBS.items():
            L.segb200_set_option(k.encode(), int(cfg.get(k, dflt)))
        us, gbs = bench(*shp)
        print(json.dumps(dict(shape="dw3x3 s%dd%d c%d @%dx%dx%d" % (shp[4], shp[5], shp[3], shp[0], shp[1], shp[2]), cfg=cfg or "default",
                              us=round(us, 1), GBps=round(gbs), frac_of_measured_hbm=round(gbs / peak, 3))), flush=True)
for k, dflt in KNOBS.items():
    L.segb200_set_option(k.encode(), dflt)

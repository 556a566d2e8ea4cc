"""Launch a few representative hot-path kernels once each (after a warm-up launch) so that ncu can capture them:

    ncu --set full --clock-control none --import-source on -k regex:'dwconv|conv_gemm|bilinear' -o gpurun_out/prof python tools/prof_kernels.py

Shapes are the BASELINE config-2 layer shapes (batch reduced to 2 for the big early layers to keep replays short)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmentron_b200 import fold, ops  # noqa: E402

dt = torch.bfloat16
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def dw(n, h, w, c, stride=1, dil=1):
    x = torch.randn(n, h, w, c, device="cuda").to(dt)
    wt = fold.pack_dw_weight(torch.randn(c, 1, 3, 3, device="cuda"), torch.ones(c, device="cuda"))
    sh = torch.zeros(c, device="cuda")
    y = torch.empty(n, (h - 1) // stride + 1, (w - 1) // stride + 1, c, device="cuda", dtype=dt)
    for _ in range(2):
        ops.dwconv3x3(x, wt, y, stride=stride, dilation=dil, shift=sh, pre_relu=True)
    torch.cuda.synchronize()


def pw(n, h, w, cin, cout, res=False):
    x = torch.randn(n, h, w, cin, device="cuda").to(dt)
    wt = fold.pack_conv_weight((torch.randn(cout, cin, 1, 1, device="cuda") / math.sqrt(cin)).to(dt), dt)
    sc, sh = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
    r = torch.randn(n, h, w, cout, device="cuda").to(dt) if res else None
    y = torch.empty(n, h, w, cout, device="cuda", dtype=dt)
    for _ in range(2):
        ops.conv_gemm(x, wt, y, cin=cin, cout=cout, scale=sc, shift=sh, residual=r)
    torch.cuda.synchronize()


if which in ("all", "dw"):
    dw(8, 65, 129, 728)
    dw(2, 513, 1025, 128)
    dw(2, 513, 1025, 128, stride=2)
def up(n, hi, wi, c, k):
    lg = torch.randn(n, hi, wi, 24, device="cuda").to(dt)
    o = torch.empty(n, c, (hi - 1) * k + 1, (wi - 1) * k + 1, device="cuda", dtype=dt)
    am = torch.empty(n, (hi - 1) * k + 1, (wi - 1) * k + 1, device="cuda", dtype=torch.uint8)
    for _ in range(2):
        ops.bilinear_nchw_out(lg, o, c, True, am)
    x = torch.randn(n, 65, 129, 256, device="cuda").to(dt)
    y = torch.empty(n, 257, 513, 256, device="cuda", dtype=dt)
    for _ in range(2):
        ops.bilinear_nhwc(x, y, align_corners=True)
    torch.cuda.synchronize()


if which in ("all", "up"):
    up(8, 257, 513, 19, 4)
if which in ("all", "gemm"):
    pw(8, 65, 129, 728, 728, res=True)
    pw(8, 65, 129, 1536, 2048)
    pw(2, 513, 1025, 128, 128)

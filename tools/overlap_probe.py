"""Feasibility probe: does a depthwise kernel (FMA/issue-bound) overlap with a tensor-bound GEMM when both co-reside on
the SMs (GEMM ring shrunk)?  Times GEMM alone, dw alone, and both concurrently on two streams (half-batch shapes)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmentron_b200 import fold, lib, ops  # noqa: E402

L = lib.load()
dt = torch.bfloat16
n, h, w, c = 4, 65, 129, 728
x = torch.randn(n, h, w, c, device="cuda").to(dt)
wt = fold.pack_conv_weight((torch.randn(c, c, 1, 1, device="cuda") / math.sqrt(c)).to(dt), dt)
sc, sh = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
y = torch.empty(n, h, w, c, device="cuda", dtype=dt)
x2 = torch.randn(n, h, w, c, device="cuda").to(dt)
wd = fold.pack_dw_weight(torch.randn(c, 1, 3, 3, device="cuda"), torch.ones(c, device="cuda"))
y2 = torch.empty(n, h, w, c, device="cuda", dtype=dt)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
REP = 30


def gemm():
    ops.conv_gemm(x, wt, y, cin=c, cout=c, scale=sc, shift=sh)


def dw():
    ops.dwconv3x3(x2, wd, y2, shift=sh, pre_relu=True)


def timed(fn_a, fn_b):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        for _ in range(REP):
            if fn_a: fn_a()
    with torch.cuda.stream(s2):
        for _ in range(REP):
            if fn_b: fn_b()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REP * 1e3


for ring, slots in [(0, 0), (144, 5), (144, 3), (96, 5)]:
    L.segb200_set_option(b"gemm_ring_kb", ring)
    L.segb200_set_option(b"dw_ring_slots", slots)
    for _ in range(3):
        gemm(); dw()
    g, d, both = timed(gemm, None), timed(None, dw), timed(gemm, dw)
    print(f"ring {ring or 192} KB, dw slots {slots or 11}: gemm {g:.1f} us, dw {d:.1f} us, serial {g + d:.1f} us, concurrent {both:.1f} us "
          f"(x{(g + d) / both:.2f})")

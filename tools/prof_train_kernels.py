"""Representative TRAINING-path kernels at the BASELINE config-3 layer shapes (DeepLabv3+/ResNet101, batch 4, 1025x2049):
CUDA-event timing (printed as achieved TFLOP/s or GB/s of ALGORITHMIC work) and, under ncu, the launches to capture:

    python tools/prof_train_kernels.py                      # timing table (JSON lines)
    ncu --set full --clock-control none --import-source on -k regex:'conv_wgrad|bn_' -c 12 -o gpurun_out/train_prof \\
        python tools/prof_train_kernels.py once             # one launch per kernel after a warm-up
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmentron_b200 import ops, train_ops as T  # noqa: E402

dt = torch.bfloat16
once = len(sys.argv) > 1 and sys.argv[1] == "once"
ITERS = 1 if once else 10


def timed(name, fn, flops=0.0, nbytes=0.0):
    fn(); fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device="cuda")
    ms = []
    for _ in range(ITERS):
        flush.zero_()                                   # L2 flush between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = sorted(ms)[len(ms) // 2]
    print(json.dumps({"kernel": name, "ms": round(t, 4), "TFLOP/s": round(flops / t / 1e9, 1) if flops else None,
                      "GB/s": round(nbytes / t / 1e6, 1) if nbytes else None}), flush=True)


def rnd(*s):
    return torch.randn(*s, device="cuda").to(dt)


def wgrad(n, h, w, cin, cout, k=1, stride=1, dil=1):
    pad = dil * (k - 1) // 2
    ho, wo = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1, (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
    x, dy = rnd(n, h, w, cin), rnd(n, ho, wo, cout)
    dw = torch.zeros(cout, k * k, cin, device="cuda")
    fl = 2.0 * n * ho * wo * cin * cout * k * k
    by = 2.0 * (x.numel() + dy.numel()) + 4.0 * dw.numel()
    timed(f"conv_wgrad {k}x{k}s{stride}d{dil} {cin}->{cout} @{n}x{ho}x{wo}",
          lambda: T.conv_wgrad(x, dy, dw, cin=cin, cout=cout, kh=k, kw=k, stride=stride, dilation=dil, pad_t=pad, pad_l=pad), fl, by)


def dgrad(n, h, w, cin, cout, k=1, dil=1):
    pad = dil * (k - 1) // 2
    dy = rnd(n, h, w, cout)
    wpk = T.pack_dgrad_weight(torch.randn(cout, cin, k, k, device="cuda").to(dt) * (cin * k * k) ** -0.5, dt)
    dx = torch.empty(n, h, w, cin, device="cuda", dtype=dt)
    timed(f"dgrad(conv_gemm) {k}x{k}d{dil} {cout}->{cin} @{n}x{h}x{w}",
          lambda: ops.conv_gemm(dy, wpk, dx, cin=cout, cout=cin, kh=k, kw=k, dilation=dil, pad_t=pad, pad_l=pad),
          2.0 * n * h * w * cin * cout * k * k, 2.0 * (dy.numel() + dx.numel()))


def bn(n, h, w, c):
    y, z, dz, dy, res = rnd(n, h, w, c), torch.empty(n, h, w, c, device="cuda", dtype=dt), rnd(n, h, w, c), \
        torch.empty(n, h, w, c, device="cuda", dtype=dt), rnd(n, h, w, c)
    g, b = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    e = y.numel() * 2.0
    st = T.bn_forward(y, z, g, b, rm, rv, 0.1, 1e-5, act="relu", residual=res)
    from segmentron_b200 import lib as L
    lib = L.load()
    rows, c_, ld = n * h * w, c, c
    slabs = T.reduce_slabs(rows, c)
    part = torch.empty(slabs * 2 * c, device="cuda")
    s = lambda: ops._stream()  # noqa: E731
    P = ops._ptr
    timed(f"bn_stats c{c} @{n}x{h}x{w}", lambda: L.check(lib.segb200_bn_stats(P(y), rows, c, ld, 0, P(part), 0, s())), 0, e)
    timed(f"bn_apply(+res,relu) c{c}", lambda: L.check(lib.segb200_bn_apply(P(y), P(st.scale), P(st.shift), P(res), None, P(z), rows, h * w, c, ld, ld, ld, 1, 0, s())), 0, 3 * e)
    timed(f"bn_bwd_reduce c{c}", lambda: L.check(lib.segb200_bn_bwd_reduce(P(dz), P(z), P(y), P(st.scale), P(st.shift), None, P(part), rows, h * w, c, ld, ld, ld, 1, 0, 0, s())), 0, 3 * e)
    timed(f"bn_bwd_reduce(mask from y) c{c}", lambda: L.check(lib.segb200_bn_bwd_reduce(P(dz), None, P(y), P(st.scale), P(st.shift), None, P(part), rows, h * w, c, ld, 0, ld, 1, 0, 0, s())), 0, 2 * e)
    timed(f"bn_bwd_apply(mask from y) c{c}", lambda: L.check(lib.segb200_bn_bwd_apply(P(dz), None, P(y), P(st.mean), P(st.invstd), P(st.scale), P(st.shift), P(st.sums), float(rows), None, P(dy), None, 0, rows, h * w, c, ld, 0, ld, ld, 0, 1, 0, s())), 0, 3 * e)
    timed(f"bn_bwd_apply(+dres) c{c}", lambda: L.check(lib.segb200_bn_bwd_apply(P(dz), P(z), P(y), P(st.mean), P(st.invstd), P(st.scale), P(st.shift), P(st.sums), float(rows), None, P(dy), P(res), 0, rows, h * w, c, ld, ld, ld, ld, ld, 1, 0, s())), 0, 5 * e)


wgrad(4, 65, 129, 512, 512, 3, 1, 2)
wgrad(4, 65, 129, 256, 1024)
wgrad(4, 65, 129, 1024, 256)
wgrad(4, 65, 129, 256, 256, 3)
wgrad(4, 257, 513, 64, 256)
wgrad(4, 257, 513, 64, 64, 3)
wgrad(4, 129, 257, 128, 512)
dgrad(4, 65, 129, 512, 512, 3, 2)
dgrad(4, 65, 129, 256, 1024)
dgrad(4, 257, 513, 64, 256)
bn(4, 257, 513, 256)
bn(4, 65, 129, 1024)

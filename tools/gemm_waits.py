"""Per-role wait-cycle breakdown of the conv_gemm kernel for a few layer shapes (uses segb200_debug_set_counters)."""
import ctypes as C
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmentron_b200 import fold, lib, ops  # noqa: E402

dt = torch.bfloat16
L = lib.load()
HAS_DBG = hasattr(L, "segb200_debug_set_counters") and "segb200_debug_set_counters" in dir(L) or False
try:
    L.segb200_debug_set_counters
    HAS_DBG = True
except AttributeError:
    HAS_DBG = False
cnt = torch.zeros(16, dtype=torch.int64, device="cuda")
NAMES = ["prod:slot_free", "mma:acc_free", "mma:operands", "epi:acc_ready", "epi:busy", "-", "tiles", "total"]


def run(name, n, h, w, cin, cout, k=1, res=False, pad=0):
    x = torch.randn(n, h, w, cin, device="cuda").to(dt)
    wt = fold.pack_conv_weight((torch.randn(cout, cin, k, k, device="cuda") / math.sqrt(cin * k * k)).to(dt), dt)
    sc, sh = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
    r = torch.randn(n, h, w, cout, device="cuda").to(dt) if res else None
    y = torch.empty(n, h, w, cout, device="cuda", dtype=dt)
    kw = dict(cin=cin, cout=cout, kh=k, kw=k, pad_t=pad, pad_l=pad, scale=sc, shift=sh, residual=r)
    for _ in range(3):
        ops.conv_gemm(x, wt, y, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.conv_gemm(x, wt, y, **kw)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    cnt.zero_()
    if HAS_DBG and L.segb200_debug_set_counters(C.c_void_p(cnt.data_ptr())) == 0:
        ops.conv_gemm(x, wt, y, **kw)
        torch.cuda.synchronize()
        L.segb200_debug_set_counters(None)
    c = cnt.tolist()
    ctas = min(148, 10 ** 9)
    tot = c[7] / 148.0
    flops = 2.0 * n * h * w * cin * cout * k * k
    byts = 2.0 * (n * h * w * (cin + cout * (2 if res else 1)))
    print(f"{name}: {us:.1f} us  {flops / us / 1e6:.0f} TFLOP/s  {byts / us / 1e3:.0f} GB/s   cycles/CTA {tot:.0f}")
    print("    " + "  ".join(f"{nm}={100.0 * v / 148.0 / max(tot, 1):.0f}%" for nm, v in zip(NAMES[:5], c[:5])) +
          f"   tiles/CTA {c[6] / 148.0:.2f}  cycles/tile {tot / max(c[6] / 148.0, 1e-9):.0f}  epi busy cycles/tile {c[4] / max(c[6], 1):.0f}")


modes = [int(m) for m in sys.argv[1:]] or [0]
for mode in modes:
    print(f"##### debug mode {mode}  (0 normal, 1 no MMA, 2 no TMA loads, 3 no epilogue work)")
    L.segb200_set_option(b"gemm_dbg_mode", mode)
    run("pw 128->128 @8x513x1025", 8, 513, 1025, 128, 128)
    run("pw 728->728 @8x65x129", 8, 65, 129, 728, 728)
    run("pw 1024->1024 @8x65x129", 8, 65, 129, 1024, 1024)
    run("pw 1536->2048", 8, 65, 129, 1536, 2048)
    if mode == 0:
        run("pw 64->128 @8x513x1025", 8, 513, 1025, 64, 128)
        run("pw 728->728 +res", 8, 65, 129, 728, 728, res=True)
        run("c3 32->64 @8x513x1025", 8, 513, 1025, 32, 64, k=3, pad=1)

"""Diagnostic: run the tcgen05 weight-gradient kernel on a few shapes with both LBO/SBO descriptor conventions and print
the error against torch autograd.  (GPU box only.)"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmentron_b200 import lib as L, train_ops as T  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
CASES = [(2, 33, 47, 64, 128, 1, 1, 1, 0), (1, 16, 64, 128, 256, 1, 1, 1, 0), (2, 33, 65, 128, 128, 3, 1, 1, 1),
         (1, 33, 65, 64, 128, 3, 2, 1, 1)]
for swap in (0, 1):
    L.load().segb200_wgrad_debug_swap(swap)
    for (n, h, w, cin, cout, k, s, d, p) in CASES:
        g = torch.Generator().manual_seed(1)
        x = torch.randn(n, h, w, cin, generator=g).bfloat16().cuda()
        ho, wo = (h + 2 * p - d * (k - 1) - 1) // s + 1, (w + 2 * p - d * (k - 1) - 1) // s + 1
        dy = torch.randn(n, ho, wo, cout, generator=g).bfloat16().cuda()
        wr = torch.zeros(cout, cin, k, k, device="cuda", requires_grad=True)
        F.conv2d(x.float().permute(0, 3, 1, 2), wr, None, s, p, d).backward(dy.float().permute(0, 3, 1, 2))
        ref = wr.grad.permute(0, 2, 3, 1).reshape(cout, k * k, cin)
        dw = torch.zeros(cout, k * k, cin, device="cuda")
        try:
            T.conv_wgrad(x, dy, dw, cin=cin, cout=cout, kh=k, kw=k, stride=s, dilation=d, pad_t=p, pad_l=p)
            torch.cuda.synchronize()
            rel = float((dw - ref).norm() / ref.norm())
            print(f"swap={swap} case={(n, h, w, cin, cout, k, s, d, p)} rel-L2={rel:.3e} max|dw|={float(dw.abs().max()):.3f} "
                  f"max|ref|={float(ref.abs().max()):.3f}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"swap={swap} case={(n, h, w, cin, cout, k, s, d, p)} FAILED: {e}", flush=True)
            sys.exit(0)
L.load().segb200_wgrad_debug_swap(0)

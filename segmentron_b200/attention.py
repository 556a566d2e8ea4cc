"""Attention blocks of DANet / CCNet on the C-ABI kernels (NHWC tensors in, NHWC tensors out).

CAM_Module (modules/module.py:142-162):  E = X^T X -> softmax(rowmax - E) -> gamma * (A X) + x
  = transpose kernel + tcgen05 GEMM (fp32 energies) + cam_softmax + tcgen05 GEMM (scale = gamma, residual = x).
CrissCrossAttention (modules/cc_attention.py:62-72): q/k/v 1x1 convs (GEMMs, bias as shift) ->
  cca_weight_softmax (ca_forward + softmax fused) -> cca_map (ca_map_forward + gamma*out + x fused).
"""
import ctypes as C

import torch

from . import fold, lib as L, ops


def cam_nhwc(x, gamma, out=None):
    """x [n,h,w,c] (16-bit, c % 8 == 0), gamma: 1-element fp32 CUDA tensor -> gamma * CAM(x) + x."""
    n, h, w, c, x_ld = ops._nhwc(x, "x")
    dt = x.dtype
    hw = h * w
    pitch = fold.round_up(hw, 64)
    lib = L.load()
    st = ops._stream
    if out is None:
        out = torch.empty(n, h, w, c, dtype=dt, device=x.device)
    gvec = gamma.detach().float().reshape(1).expand(c).contiguous()
    for b in range(n):
        xb = x[b:b + 1]
        xt = torch.zeros(c, pitch, dtype=dt, device=x.device)                   # [C][N] K-major copy (zero tail)
        L.check(lib.segb200_nhwc_to_cn(ops._ptr(xb), ops._ptr(xt), 1, c, hw, x_ld, pitch, ops.dt_code(dt), st()), "nhwc_to_cn")
        energy = torch.empty(1, 1, c, c, dtype=torch.float32, device=x.device)
        ops.conv_gemm(xt.view(1, 1, c, pitch), xt, energy, cin=pitch, cout=c)   # E[c1][c2] = sum_p x[p,c1] x[p,c2]
        cpad = fold.round_up(c, fold.conv_kblock(c))                               # K pitch the second GEMM expects
        att = torch.zeros(c, cpad, dtype=dt, device=x.device)
        L.check(lib.segb200_cam_softmax(ops._ptr(energy), ops._ptr(att), c, c, c, cpad, ops.dt_code(dt), st()), "cam_softmax")
        ops.conv_gemm(xb, att, out[b:b + 1], cin=c, cout=c, scale=gvec, residual=xb)   # out[p,c1] = g*sum_c2 att[c1,c2] x[p,c2] + x
    return out


def cca_nhwc(x, wq, bq, wk, bk, wv, bv, gamma, out=None):
    """One criss-cross attention step.  wq/wk/wv: packed 1x1 weights (fold.pack_conv_weight), b*: fp32 bias vectors."""
    n, h, w, c, x_ld = ops._nhwc(x, "x")
    dt = x.dtype
    cq = wq.shape[0]
    lib = L.load()
    q = torch.empty(n, h, w, cq, dtype=dt, device=x.device)
    k = torch.empty(n, h, w, cq, dtype=dt, device=x.device)
    v = torch.empty(n, h, w, c, dtype=dt, device=x.device)
    ops.conv_gemm(x, wq, q, cin=c, cout=cq, shift=bq)
    ops.conv_gemm(x, wk, k, cin=c, cout=cq, shift=bk)
    ops.conv_gemm(x, wv, v, cin=c, cout=c, shift=bv)
    att_ld = fold.round_up(h + w - 1, 4)
    att = torch.empty(n, h, w, att_ld, dtype=torch.float32, device=x.device)
    L.check(lib.segb200_cca_weight_softmax(ops._ptr(q), ops._ptr(k), ops._ptr(att), n, h, w, cq, cq, cq, att_ld,
                                           ops.dt_code(dt), ops._stream()), "cca_weight_softmax")
    y = out if out is not None else torch.empty(n, h, w, c, dtype=dt, device=x.device)
    g = gamma.detach().float().reshape(1).contiguous()
    L.check(lib.segb200_cca_map(ops._ptr(att), ops._ptr(v), ops._ptr(x), ops._ptr(y), ops._ptr(g), n, h, w, c, att_ld, c, x_ld,
                                c, ops.dt_code(dt), ops._stream()), "cca_map")
    return y


def pam_nhwc(x, wq, bq, wk, bk, wv, bv, gamma, out=None):
    """PAM_Module on NHWC x [n,h,w,c] (c % 64 == 0, contiguous).  wq/wk: packed [c/8 = 64][1][c], wv: packed [c][1][c]
    (fold.pack_conv_weight); bq/bk/bv fp32 bias vectors; gamma 1-element tensor."""
    n, h, w, c, x_ld = ops._nhwc(x, "x")
    if x_ld != c or c % 64 != 0:
        raise RuntimeError("segb200 pam: x must be a contiguous NHWC tensor with channels % 64 == 0")
    dt = x.dtype
    ntok = h * w
    dq = wq.shape[0]
    dv = wv.shape[0]
    if dq != 64 or dv % 256 != 0:
        raise RuntimeError(f"segb200 pam: kernel is specialised for query depth 64 and d_v % 256 == 0 (got {dq}, {dv})")
    lib = L.load()
    q = torch.empty(n, h, w, dq, dtype=dt, device=x.device)
    k = torch.empty(n, h, w, dq, dtype=dt, device=x.device)
    ops.conv_gemm(x, wq, q, cin=c, cout=dq, shift=bq)
    ops.conv_gemm(x, wk, k, cin=c, cout=dq, shift=bk)
    # V^T[b] = W_v . X_b^T : the same GEMM with the roles swapped (A = weight rows, B = the image's pixels)
    pitch = fold.round_up(ntok, 8)
    vt = torch.empty(n, dv, pitch, dtype=dt, device=x.device)
    wv_as_x = wv.view(1, 1, dv, wv.shape[-1])
    for b in range(n):
        ops.conv_gemm(wv_as_x, x[b], vt[b].view(1, 1, dv, pitch)[..., :ntok], cin=c, cout=ntok)
    y = out if out is not None else torch.empty(n, h, w, c, dtype=dt, device=x.device)
    sm = torch.empty(n * ntok, dtype=torch.float32, device=x.device)
    sl = torch.empty(n * ntok, dtype=torch.float32, device=x.device)
    g = gamma.detach().float().reshape(1).contiguous()
    L.check(lib.segb200_pam_attention(ops._ptr(q), ops._ptr(k), ops._ptr(vt), ops._ptr(bv), ops._ptr(g), ops._ptr(x), ops._ptr(y),
                                      ops._ptr(sm), ops._ptr(sl), n, ntok, dv, dq, dq, pitch, x_ld, c, ops.dt_code(dt),
                                      ops._stream()), "pam_attention")
    return y

"""Tensor-level wrappers of the training-path C-ABI kernels (include/segb200.h, "TRAINING PATH").

Same conventions as ``ops.py``: NHWC CUDA tensors / channel-slice views, current stream, RuntimeError on any
failure, no PyTorch fallback.  These are the single-call forms used by the tests and by small harnesses; the
whole-model training plan (``train.py``) pre-marshals the same C calls into a static launch list.
"""
import ctypes as C

import torch

from . import fold, lib as L
from .ops import _nhwc, _ptr, _stream, dt_code


def _rows(t):
    n, h, w, c, ld = _nhwc(t, "tensor")
    return n * h * w, h * w, c, ld


def reduce_slabs(rows, c, max_slabs=0):
    return L.load().segb200_reduce_slabs(rows, c, max_slabs)


def conv_wgrad(x, dy, dw, *, cin, cout, kh=1, kw=1, stride=1, dilation=1, pad_t=0, pad_l=0, splits=0, max_ctas=0):
    """dw [cout][kh*kw][cin] fp32 += conv weight gradient (x: conv input, dy: gradient of the conv output)."""
    n, h, w, cx, x_ld = _nhwc(x, "x")
    _, ho, wo, cy, dy_ld = _nhwc(dy, "dy")
    if cx < cin or cy < cout or dw.dtype != torch.float32 or not dw.is_contiguous() or dw.numel() != cout * kh * kw * cin:
        raise RuntimeError("segb200 conv_wgrad: shape mismatch")
    a = L.WgradArgs()
    a.x, a.dy, a.dw = _ptr(x), _ptr(dy), _ptr(dw)
    a.n, a.h, a.w, a.cin, a.x_ld = n, h, w, cin, x_ld
    a.ho, a.wo, a.cout, a.dy_ld = ho, wo, cout, dy_ld
    a.kh, a.kw, a.stride, a.dilation, a.pad_t, a.pad_l = kh, kw, stride, dilation, pad_t, pad_l
    a.dtype, a.max_ctas, a.splits = dt_code(x.dtype), max_ctas, splits
    L.check(L.load().segb200_conv_wgrad(C.byref(a), _stream()), "conv_wgrad")
    return dw


def pack_dgrad_weight(w, dtype):
    """[Cout, Cin, kh, kw] (OIHW) -> the forward-kernel operand of the DATA gradient: [Cin_pad8][kh*kw reversed][Cout_padK]
    (transposed and tap-flipped; stride-1 'same' convs: dX = conv(dY, this) with the same dilation and padding)."""
    co, ci, kh, kw = w.shape
    bk = fold.conv_kblock(co)
    cop = fold.round_up(co, bk)
    cip = fold.round_up(ci, 8)
    out = torch.zeros(cip, kh * kw, cop, dtype=dtype, device=w.device)
    out[:ci, :, :co] = w.permute(1, 2, 3, 0).reshape(ci, kh * kw, co).flip(1).to(dtype)
    return out.contiguous()


class BNState:
    """Per-layer fp32 vectors of one train-mode BatchNorm application."""

    def __init__(self, c, device):
        self.mean, self.invstd, self.scale, self.shift = (torch.empty(c, dtype=torch.float32, device=device) for _ in range(4))
        self.sums = torch.empty(2, c, dtype=torch.float32, device=device)


def bn_forward(y, z, gamma, beta, running_mean, running_var, momentum, eps, act=None, residual=None, nc_scale=None,
               state=None):
    """z = act(BN_train(y) + residual) * nc_scale; returns the BNState (mean / invstd / scale / shift) for backward."""
    rows, hw, c, y_ld = _rows(y)
    _, _, cz, z_ld = _rows(z)
    lib = L.load()
    st = state or BNState(c, y.device)
    slabs = reduce_slabs(rows, c)
    partial = torch.empty(slabs * 2 * c, dtype=torch.float32, device=y.device)
    s = _stream()
    L.check(lib.segb200_bn_stats(_ptr(y), rows, c, y_ld, dt_code(y.dtype), _ptr(partial), 0, s), "bn_stats")
    L.check(lib.segb200_bn_finalize(_ptr(partial), slabs, c, float(rows), _ptr(gamma), _ptr(beta), _ptr(running_mean),
                                    _ptr(running_var), momentum, eps, _ptr(st.mean), _ptr(st.invstd), _ptr(st.scale),
                                    _ptr(st.shift), s), "bn_finalize")
    res_ld = _rows(residual)[3] if residual is not None else 0
    L.check(lib.segb200_bn_apply(_ptr(y), _ptr(st.scale), _ptr(st.shift), _ptr(residual), _ptr(nc_scale), _ptr(z), rows, hw, c,
                                 y_ld, res_ld, z_ld, L.ACT[act], dt_code(y.dtype), s), "bn_apply")
    return st


def bn_backward(dz, z, y, st, dy, dgamma, dbeta, act=None, dres=None, dres_accumulate=False, nc_scale=None):
    """dy = dBN(dz * act'(.)); dgamma/dbeta (fp32) are ACCUMULATED; dres (+)= dz * act'(.).  z=None: the activation mask is
    recomputed from y (only valid when no residual entered the activation)."""
    rows, hw, c, dz_ld = _rows(dz)
    lib = L.load()
    s = _stream()
    slabs = reduce_slabs(rows, c)
    partial = torch.empty(slabs * 2 * c, dtype=torch.float32, device=dz.device)
    z_ld = _rows(z)[3] if z is not None else 0
    y_ld = _rows(y)[3]
    L.check(lib.segb200_bn_bwd_reduce(_ptr(dz), _ptr(z), _ptr(y), _ptr(st.scale), _ptr(st.shift), _ptr(nc_scale), _ptr(partial),
                                      rows, hw, c, dz_ld, z_ld, y_ld, L.ACT[act], dt_code(dz.dtype), 0, s), "bn_bwd_reduce")
    L.check(lib.segb200_bn_bwd_finalize(_ptr(partial), slabs, c, _ptr(st.mean), _ptr(st.invstd), _ptr(st.sums), _ptr(dgamma),
                                        _ptr(dbeta), s), "bn_bwd_finalize")
    L.check(lib.segb200_bn_bwd_apply(_ptr(dz), _ptr(z), _ptr(y), _ptr(st.mean), _ptr(st.invstd), _ptr(st.scale), _ptr(st.shift),
                                     _ptr(st.sums), float(rows), _ptr(nc_scale), _ptr(dy), _ptr(dres), int(bool(dres_accumulate)), rows, hw, c,
                                     dz_ld, z_ld, y_ld, _rows(dy)[3], _rows(dres)[3] if dres is not None else 0, L.ACT[act],
                                     dt_code(dz.dtype), s), "bn_bwd_apply")
    return dy


def relu_mask(g, x, out, accumulate=False):
    """out (+)= g * [x > 0]  -- the backward of a leading ReLU (SeparableConv2d relu_first, modules/basic.py:45-46), through
    the residual path of the BatchNorm-backward kernel"""
    rows, hw, c, g_ld = _rows(g)
    L.check(L.load().segb200_bn_bwd_apply(_ptr(g), _ptr(x), None, None, None, None, None, None, float(rows), None, None, _ptr(out),
                                          int(bool(accumulate)), rows, hw, c, g_ld, _rows(x)[3], 0, 0, _rows(out)[3], L.ACT["relu"],
                                          dt_code(g.dtype), _stream()), "bn_bwd_apply")
    return out


def maxpool3x3s2_bwd(x, dy, dx):
    n, h, w, c, x_ld = _nhwc(x, "x")
    L.check(L.load().segb200_maxpool3x3s2_bwd(_ptr(x), _ptr(dy), _ptr(dx), n, h, w, c, x_ld, _nhwc(dy, "dy")[4], _nhwc(dx, "dx")[4],
                                              dt_code(x.dtype), _stream()), "maxpool3x3s2_bwd")
    return dx


def maxpool3x3s2_idx(x, y, idx):
    n, h, w, c, x_ld = _nhwc(x, "x")
    assert idx.dtype == torch.uint8 and idx.is_contiguous() and idx.numel() == y.shape[0] * y.shape[1] * y.shape[2] * c
    L.check(L.load().segb200_maxpool3x3s2_idx(_ptr(x), _ptr(y), _ptr(idx), n, h, w, c, x_ld, _nhwc(y, "y")[4], dt_code(x.dtype),
                                              _stream()), "maxpool3x3s2_idx")
    return y


def maxpool3x3s2_bwd_idx(idx, dy, dx):
    n, h, w, c, dx_ld = _nhwc(dx, "dx")
    L.check(L.load().segb200_maxpool3x3s2_bwd_idx(_ptr(idx), _ptr(dy), _ptr(dx), n, h, w, c, _nhwc(dy, "dy")[4], dx_ld,
                                                  dt_code(dx.dtype), _stream()), "maxpool3x3s2_bwd_idx")
    return dx


def bilinear_nhwc_bwd(dy, dx, align_corners=True, accumulate=False, gscale=None):
    n, ho, wo, c, dy_ld = _nhwc(dy, "dy")
    _, hi, wi, cx, dx_ld = _nhwc(dx, "dx")
    assert cx == c
    L.check(L.load().segb200_bilinear_nhwc_bwd(_ptr(dy), _ptr(dx), n, hi, wi, c, dx_ld, ho, wo, dy_ld, int(align_corners),
                                               int(bool(accumulate)), _ptr(gscale), dt_code(dy.dtype), _stream()),
            "bilinear_nhwc_bwd")
    return dx


def upsample_ce(logits, target, dfull, nclass, align_corners=True, ignore_index=-1):
    """-> out3 (device fp32: mean loss, 1/valid, valid); dfull [n][H][W][d_ld] = softmax - onehot."""
    n, hi, wi, _, x_ld = _nhwc(logits, "logits")
    _, ho, wo, _, d_ld = _nhwc(dfull, "dfull")
    assert target.dtype == torch.int64 and target.is_contiguous() and tuple(target.shape) == (n, ho, wo)
    lib = L.load()
    nb = lib.segb200_upsample_ce_blocks(n, ho, wo)
    partial = torch.empty(2 * nb, dtype=torch.float32, device=logits.device)
    out3 = torch.empty(3, dtype=torch.float32, device=logits.device)
    L.check(lib.segb200_upsample_ce(_ptr(logits), _ptr(target), _ptr(dfull), _ptr(partial), _ptr(out3), n, hi, wi, nclass, x_ld,
                                    ho, wo, d_ld, int(align_corners), ignore_index, dt_code(logits.dtype), _stream()),
            "upsample_ce")
    return out3


def dw_wgrad(x, dy, dw, dilation=1, pre_relu=False, accumulate=True, variant=1):
    """dw: fp32 [c][9] (torch depthwise weight [C,1,3,3] flattened), accumulated.  variant=2: the opt-in sliding-window kernel
    (csrc/dw_wgrad2.cu)."""
    n, h, w, c, x_ld = _nhwc(x, "x")
    lib = L.load()
    fn, slabs = (lib.segb200_dw_wgrad, reduce_slabs(n * h * w, c)) if variant == 1 else \
        (lib.segb200_dw_wgrad_v2, lib.segb200_dw_wgrad_v2_slabs(n * h * w, c, 0))
    partial = torch.empty(slabs * 9 * c, dtype=torch.float32, device=x.device)
    s = _stream()
    L.check(fn(_ptr(x), _ptr(dy), _ptr(partial), n, h, w, c, x_ld, _nhwc(dy, "dy")[4], dilation,
               int(bool(pre_relu)), dt_code(x.dtype), 0, s), "dw_wgrad")
    L.check(lib.segb200_reduce_partials(_ptr(partial), slabs, 9, c, _ptr(dw), 1, 9, int(bool(accumulate)), 1.0, s),
            "reduce_partials")
    return dw


def nc_broadcast(v, y, scale=1.0, accumulate=False):
    n, h, w, c, y_ld = _nhwc(y, "y")
    v_ld = _nhwc(v, "v")[4]
    L.check(L.load().segb200_nc_broadcast(_ptr(v), _ptr(y), n, h * w, c, v_ld, y_ld, scale, int(bool(accumulate)),
                                          dt_code(y.dtype), _stream()), "nc_broadcast")
    return y


def stride2_place(t, z, mode):
    n, h, w, c, z_ld = _nhwc(z, "z")
    _, ht, wt, ct, t_ld = _nhwc(t, "t")
    assert ct == c and ht == (h - 1) // 2 + 1 and wt == (w - 1) // 2 + 1
    L.check(L.load().segb200_stride2_place(_ptr(t), _ptr(z), n, h, w, c, t_ld, z_ld, mode, dt_code(z.dtype), _stream()),
            "stride2_place")
    return z


def gather_cast(src, index, dst):
    assert src.dtype == torch.float32 and index.dtype == torch.int32 and index.numel() == dst.numel()
    L.check(L.load().segb200_gather_cast(_ptr(src), _ptr(index), _ptr(dst), dst.numel(), dt_code(dst.dtype), _stream()),
            "gather_cast")
    return dst


def scatter_add(src, index, dst):
    assert src.dtype == dst.dtype == torch.float32 and index.dtype == torch.int32 and index.numel() == src.numel()
    L.check(L.load().segb200_scatter_add(_ptr(src), _ptr(index), _ptr(dst), src.numel(), _stream()), "scatter_add")
    return dst


def sgd_step(p, g, m, lr, momentum=0.9, weight_decay=0.0, grad_scale=1.0):
    assert p.dtype == g.dtype == m.dtype == torch.float32 and p.numel() == g.numel() == m.numel()
    L.check(L.load().segb200_sgd_step(_ptr(p), _ptr(g), _ptr(m), p.numel(), lr, momentum, weight_decay, grad_scale, _stream()),
            "sgd_step")
    return p

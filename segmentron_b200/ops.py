"""Tensor-level wrappers of the C-ABI kernels.

Activations are torch CUDA tensors of shape [N, H, W, C] ("NHWC"); a tensor may be a channel slice
of a wider buffer (``buf[..., a:b]``) -- the channel pitch is taken from the strides.  Every wrapper
enqueues on torch's current CUDA stream and raises RuntimeError on any failure; nothing here falls
back to PyTorch ops.
"""
import ctypes as C

import torch

from . import lib as L

_DT = {torch.bfloat16: L.BF16, torch.float16: L.F16, torch.float32: L.F32}


def dt_code(dtype):
    try:
        return _DT[dtype]
    except KeyError:
        if _PLAN_DRY_RUN:                 # fp64 buffers of the CPU plan interpreter (tests); never launched
            return L.F32
        raise RuntimeError(f"segb200: unsupported dtype {dtype}")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# Plan DRY RUN (tests only): lets a launch list be BUILT over CPU tensors so that its host logic (buffer routing,
# accumulate flags, index tables) can be checked without a GPU.  Nothing can be launched in that mode -- the kernels
# would fault on host pointers -- and no compute path ever sets it.
_PLAN_DRY_RUN = False


def _nhwc(t, name):
    """-> (n, h, w, c, ld) of an NHWC tensor / channel-slice view; validates the layout."""
    if not t.is_cuda and not _PLAN_DRY_RUN:
        raise RuntimeError(f"segb200: {name} must be a CUDA tensor (no CPU implementation)")
    if t.dim() != 4:
        raise RuntimeError(f"segb200: {name} must be [N,H,W,C]")
    n, h, w, c = t.shape
    # the pixel pitch; a size-1 dimension carries an arbitrary stride (torch keeps whatever a permute left there)
    ld = t.stride(2) if w > 1 else (t.stride(1) if h > 1 else (t.stride(0) if n > 1 else max(c, t.stride(2))))
    if t.stride(3) != 1 or ld < c or (h > 1 and t.stride(1) != w * ld) or (n > 1 and t.stride(0) != h * w * ld):
        raise RuntimeError(f"segb200: {name} is not an NHWC tensor / channel slice (shape {tuple(t.shape)}, "
                           f"strides {t.stride()})")
    return n, h, w, c, ld


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def make_conv_args(x, wgt, y, *, cin, cout, kh=1, kw=1, stride=1, dilation=1, pad_t=0, pad_l=0,
                   scale=None, shift=None, act=None, residual=None, max_ctas=0):
    n, h, w, cx, x_ld = _nhwc(x, "x")
    ny, ho, wo, cy, y_ld = _nhwc(y, "y")
    if cx < cin or cy < cout:
        raise RuntimeError("segb200 conv: shape mismatch")
    a = L.ConvArgs()
    a.x, a.wgt, a.y = _ptr(x), _ptr(wgt), _ptr(y)
    a.scale, a.shift = _ptr(scale), _ptr(shift)
    a.residual = _ptr(residual)
    a.res_ld = _nhwc(residual, "residual")[4] if residual is not None else 0
    a.n, a.h, a.w, a.cin, a.x_ld = n, h, w, cin, x_ld
    a.ho, a.wo, a.cout, a.y_ld = ho, wo, cout, y_ld
    a.kh, a.kw, a.stride, a.dilation, a.pad_t, a.pad_l = kh, kw, stride, dilation, pad_t, pad_l
    a.act = L.ACT[act]
    a.dtype = dt_code(x.dtype)
    a.max_ctas = max_ctas
    a.y_f32 = 1 if y.dtype == torch.float32 else 0
    return a


def conv_gemm(x, wgt, y, **kw):
    a = make_conv_args(x, wgt, y, **kw)
    L.check(L.load().segb200_conv_gemm(C.byref(a), _stream()), "conv_gemm")
    return y


def make_dw_args(x, wgt, y, *, stride=1, dilation=1, shift=None, pre_relu=False, act=None):
    n, h, w, c, x_ld = _nhwc(x, "x")
    ny, ho, wo, cy, y_ld = _nhwc(y, "y")
    if cy != c or ny != n:
        raise RuntimeError("segb200 dwconv: shape mismatch")
    a = L.DwArgs()
    a.x, a.wgt, a.shift, a.y = _ptr(x), _ptr(wgt), _ptr(shift), _ptr(y)
    a.n, a.h, a.w, a.c, a.x_ld, a.y_ld = n, h, w, c, x_ld, y_ld
    a.ho, a.wo, a.stride, a.dilation = ho, wo, stride, dilation
    a.pre_relu, a.act, a.dtype = int(bool(pre_relu)), L.ACT[act], dt_code(x.dtype)
    return a


def dwconv3x3(x, wgt, y, **kw):
    a = make_dw_args(x, wgt, y, **kw)
    L.check(L.load().segb200_dwconv3x3(C.byref(a), _stream()), "dwconv3x3")
    return y


def pack_s2d(x_nchw, out):
    """NCHW image -> space-to-depth NHWC [n, ceil(h/2), ceil(w/2), ld]."""
    if not x_nchw.is_cuda:
        raise RuntimeError("segb200: input must be a CUDA tensor (no CPU implementation)")
    x_nchw = x_nchw.contiguous()
    n, c, h, w = x_nchw.shape
    _, hs, ws, co, ld = _nhwc(out, "out")
    assert hs == (h + 1) // 2 and ws == (w + 1) // 2
    L.check(L.load().segb200_pack_s2d(_ptr(x_nchw), dt_code(x_nchw.dtype), _ptr(out), dt_code(out.dtype),
                                      n, c, h, w, ld, _stream()), "pack_s2d")
    return out


def global_avgpool(x, out):
    n, h, w, c, ld = _nhwc(x, "x")
    L.check(L.load().segb200_global_avgpool(_ptr(x), _ptr(out), n, h, w, c, ld, dt_code(x.dtype), _stream()),
            "global_avgpool")
    return out


def adaptive_avgpool(x, out, s):
    n, h, w, c, ld = _nhwc(x, "x")
    _, _, _, _, old = _nhwc(out, "out")
    L.check(L.load().segb200_adaptive_avgpool(_ptr(x), _ptr(out), n, h, w, c, ld, s, old, dt_code(x.dtype), _stream()),
            "adaptive_avgpool")
    return out


def maxpool3x3s2(x, y):
    n, h, w, c, x_ld = _nhwc(x, "x")
    _, ho, wo, cy, y_ld = _nhwc(y, "y")
    assert cy == c and ho == (h - 1) // 2 + 1 and wo == (w - 1) // 2 + 1
    L.check(L.load().segb200_maxpool3x3s2(_ptr(x), _ptr(y), n, h, w, c, x_ld, y_ld, dt_code(x.dtype), _stream()),
            "maxpool3x3s2")
    return y


def bilinear_nhwc(x, y, align_corners=True):
    n, hi, wi, c, x_ld = _nhwc(x, "x")
    _, ho, wo, cy, y_ld = _nhwc(y, "y")
    assert cy == c
    L.check(L.load().segb200_bilinear_nhwc(_ptr(x), _ptr(y), n, hi, wi, c, x_ld, ho, wo, y_ld, int(align_corners),
                                           dt_code(x.dtype), _stream()), "bilinear_nhwc")
    return y


def bilinear_nchw_out(x, y_nchw, c, align_corners=True, argmax_out=None):
    n, hi, wi, _, x_ld = _nhwc(x, "x")
    ho, wo = y_nchw.shape[2], y_nchw.shape[3]
    assert y_nchw.is_contiguous() and y_nchw.shape[1] == c
    L.check(L.load().segb200_bilinear_nchw_out(_ptr(x), _ptr(y_nchw), _ptr(argmax_out), n, hi, wi, c, x_ld, ho, wo,
                                               int(align_corners), dt_code(x.dtype), dt_code(y_nchw.dtype), _stream()),
            "bilinear_nchw_out")
    return y_nchw


def nchw_to_nhwc(x_nchw, out):
    if not x_nchw.is_cuda:
        raise RuntimeError("segb200: input must be a CUDA tensor (no CPU implementation)")
    x_nchw = x_nchw.contiguous()
    n, c, h, w = x_nchw.shape
    _, _, _, _, ld = _nhwc(out, "out")
    L.check(L.load().segb200_nchw_to_nhwc(_ptr(x_nchw), dt_code(x_nchw.dtype), _ptr(out), dt_code(out.dtype), n, c, h, w,
                                          ld, _stream()), "nchw_to_nhwc")
    return out


def nhwc_to_nchw(x, out_nchw):
    n, h, w, c, ld = _nhwc(x, "x")
    assert out_nchw.is_contiguous()
    L.check(L.load().segb200_nhwc_to_nchw(_ptr(x), dt_code(x.dtype), _ptr(out_nchw), dt_code(out_nchw.dtype), n, c, h, w,
                                          ld, _stream()), "nhwc_to_nchw")
    return out_nchw

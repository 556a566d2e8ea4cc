"""Host emulation of a SUBSET of the C ABI (include/segb200.h) for CPU tests of the Python layers above it.  TEST INFRASTRUCTURE.

Each method receives exactly what the C entry point receives (ctypes pointers, sizes, pitches, dtype code) and performs the
documented semantics with numpy on the host memory behind the pointers.  For kernels that have run on the B200 this is an
emulation of verified behaviour; for kernels written after round 1's GPU budget was spent (cam_softmax_bwd, cam_bwd_pack) it is a
line-by-line transcription of the .cu file, so the test pins their algorithm.  Only fp32 storage (dtype code F32) is emulated --
the tests run the "16-bit" paths with fp32 tensors under ops._PLAN_DRY_RUN.
"""
import ctypes as C

import numpy as np

F32 = 2


def view(ptr, shape, dtype=np.float32):
    addr = ptr.value if isinstance(ptr, C.c_void_p) else ptr
    if addr is None:
        return None
    n = int(np.prod(shape))
    ct = {np.float32: C.c_float, np.int64: C.c_longlong, np.int32: C.c_int32}[dtype]
    return np.frombuffer((ct * n).from_address(addr), dtype=dtype).reshape(shape)


def kblock(cin):
    return 64 if cin >= 64 else (32 if cin >= 32 else 16)


def act(v, a):
    return np.maximum(v, 0) if a == 1 else (np.clip(v, 0, 6) if a == 2 else v)


class FakeLib:
    def segb200_last_error(self):
        return b"(fake lib)"

    def segb200_conv_kblock(self, cin):
        return kblock(cin)

    def segb200_conv_gemm(self, args_ref, stream):
        a = args_ref._obj
        assert a.dtype == F32 and a.kh == 1 and a.kw == 1 and a.stride == 1, "the emulation covers 1x1 GEMMs on fp32 storage"
        m = a.n * a.h * a.w
        assert (a.ho, a.wo) == (a.h, a.w)
        x = view(C.c_void_p(a.x), (m, a.x_ld))[:, :a.cin]
        kp = -(-a.cin // kblock(a.cin)) * kblock(a.cin)
        wgt = view(C.c_void_p(a.wgt), (a.cout, kp))[:, :a.cin]
        y = view(C.c_void_p(a.y), (m, a.y_ld))
        o = x.astype(np.float64) @ wgt.astype(np.float64).T
        if a.scale:
            o = o * view(C.c_void_p(a.scale), (a.cout,)).astype(np.float64)
        if a.shift:
            o = o + view(C.c_void_p(a.shift), (a.cout,)).astype(np.float64)
        if a.residual:
            o = o + view(C.c_void_p(a.residual), (m, a.res_ld))[:, :a.cout].astype(np.float64)
        y[:, :a.cout] = act(o, a.act).astype(np.float32)
        return 0

    def segb200_nhwc_to_cn(self, x, y, n, c, hw, x_ld, pitch, dtype, stream):
        assert dtype == F32
        xs = view(x, (n, hw, x_ld))[:, :, :c]
        ys = view(y, (n, c, pitch))
        ys[:, :, :hw] = xs.transpose(0, 2, 1)
        return 0

    def segb200_cam_softmax(self, energy, att, rows, c, e_ld, att_ld, dtype, stream):
        e = view(energy, (rows, e_ld))[:, :c].astype(np.float64)
        a = view(att, (rows, att_ld))
        z = e.min(1, keepdims=True) - e                      # == rowmax - E up to the softmax's shift invariance
        p = np.exp(z)
        a[:, :c] = (p / p.sum(1, keepdims=True)).astype(np.float32)
        return 0

    def segb200_cam_softmax_bwd(self, att, g, gamma, de, dgamma_partial, c, att_ld, g_ld, de_ld, dtype, stream):
        a = view(att, (c, att_ld))[:, :c]
        gm = float(view(gamma, (1,))[0])
        gg = view(g, (c, g_ld))[:, :c]
        r = (a * gg).sum(1, keepdims=True, dtype=np.float32)
        view(de, (c, de_ld))[:, :c] = -gm * a * (gg - r)
        view(dgamma_partial, (c,))[:] = r[:, 0]
        return 0

    def segb200_cam_bwd_pack(self, att, de, gamma, w1, w2, c, att_ld, de_ld, w_ld, dtype, stream):
        a = view(att, (c, att_ld))[:, :c]
        d = view(de, (c, de_ld))[:, :c]
        gm = float(view(gamma, (1,))[0])
        o1, o2 = view(w1, (c, w_ld)), view(w2, (c, w_ld))
        o1[...] = 0
        o2[...] = 0
        o1[:, :c] = gm * a.T
        o2[:, :c] = d + d.T
        return 0

    def segb200_reduce_partials(self, partial, slabs, k, c, out, sk, sc, accumulate, scale, stream):
        p = view(partial, (slabs, k, c)).astype(np.float64).sum(0) * scale
        o = view(out, (max(1, (k - 1) * sk + (c - 1) * sc + 1),))
        for kk in range(k):
            for ch in range(c):
                o[kk * sk + ch * sc] = (o[kk * sk + ch * sc] if accumulate else 0.0) + p[kk, ch]
        return 0


def _more(cls):
    def segb200_row_softmax(self, s, p, rows, n, s_ld, p_ld, dtype, stream):
        e = view(s, (rows, s_ld))[:, :n].astype(np.float64)
        o = view(p, (rows, p_ld))
        z = np.exp(e - e.max(1, keepdims=True))
        o[...] = 0
        o[:, :n] = (z / z.sum(1, keepdims=True)).astype(np.float32)
        return 0

    def segb200_row_softmax_bwd(self, p, d, gamma, ds, partial, rows, n, p_ld, d_ld, ds_ld, dtype, stream):
        pp = view(p, (rows, p_ld))[:, :n]
        dd = view(d, (rows, d_ld))[:, :n]
        gm = float(view(gamma, (1,))[0])
        r = (pp * dd).sum(1, keepdims=True, dtype=np.float32)
        o = view(ds, (rows, ds_ld))
        o[...] = 0
        o[:, :n] = gm * pp * (dd - r)
        view(partial, (rows,))[:] = r[:, 0]
        return 0

    def segb200_conv_wgrad(self, args_ref, stream):
        a = args_ref._obj
        assert a.dtype == F32 and a.kh == 1 and a.kw == 1 and a.stride == 1
        m = a.n * a.h * a.w
        x = view(C.c_void_p(a.x), (m, a.x_ld))[:, :a.cin].astype(np.float64)
        dy = view(C.c_void_p(a.dy), (m, a.dy_ld))[:, :a.cout].astype(np.float64)
        dw = view(C.c_void_p(a.dw), (a.cout, a.cin))
        dw += (dy.T @ x).astype(np.float32)
        return 0

    def segb200_reduce_slabs(self, rows, c, max_slabs):
        return 1

    def segb200_bn_stats(self, x, rows, c, x_ld, dtype, partial, max_slabs, stream):
        xs = view(x, (rows, x_ld))[:, :c].astype(np.float64)
        p = view(partial, (2, c))
        p[0] = xs.sum(0)
        p[1] = (xs * xs).sum(0)
        return 0
    for f in (segb200_row_softmax, segb200_row_softmax_bwd, segb200_conv_wgrad, segb200_reduce_slabs, segb200_bn_stats):
        setattr(cls, f.__name__, f)


_more(FakeLib)


def _pooling(cls):
    import torch
    import torch.nn.functional as F

    def segb200_adaptive_avgpool(self, x, out, n, h, w, c, x_ld, s, out_ld, dtype, stream):
        xs = view(x, (n, h, w, x_ld))[..., :c]
        o = view(out, (n, s, s, out_ld))
        for bi in range(s):
            h0, h1 = (bi * h) // s, ((bi + 1) * h + s - 1) // s
            for bj in range(s):
                w0, w1 = (bj * w) // s, ((bj + 1) * w + s - 1) // s
                o[:, bi, bj, :c] = xs[:, h0:h1, w0:w1].mean((1, 2), dtype=np.float64)
        return 0

    def segb200_adaptive_avgpool_bwd(self, dy, dx, n, h, w, c, dy_ld, dx_ld, s, accumulate, dtype, stream):
        """transcription of adaptive_pool_bwd_kernel: per input pixel, gather over the bins that contain it"""
        g = view(dy, (n, s, s, dy_ld))[..., :c]
        o = view(dx, (n, h, w, dx_ld))
        for y in range(h):
            for x in range(w):
                acc = np.zeros((n, c), dtype=np.float32)
                for bi in range(s):
                    h0, h1 = (bi * h) // s, ((bi + 1) * h + s - 1) // s
                    if y < h0 or y >= h1:
                        continue
                    for bj in range(s):
                        w0, w1 = (bj * w) // s, ((bj + 1) * w + s - 1) // s
                        if x < w0 or x >= w1:
                            continue
                        acc += g[:, bi, bj] * np.float32(1.0 / ((h1 - h0) * (w1 - w0)))
                o[:, y, x, :c] = acc + (o[:, y, x, :c] if accumulate else 0)
        return 0

    def segb200_bilinear_nhwc(self, x, y, n, hi, wi, c, x_ld, ho, wo, y_ld, align, dtype, stream):
        xs = torch.from_numpy(view(x, (n, hi, wi, x_ld))[..., :c].copy()).permute(0, 3, 1, 2)
        view(y, (n, ho, wo, y_ld))[..., :c] = F.interpolate(xs, (ho, wo), mode="bilinear", align_corners=bool(align)).permute(0, 2, 3, 1).numpy()
        return 0

    def segb200_bilinear_nhwc_bwd(self, dy, dx, n, hi, wi, c, dx_ld, ho, wo, dy_ld, align, accumulate, gscale, dtype, stream):
        g = torch.from_numpy(view(dy, (n, ho, wo, dy_ld))[..., :c].copy()).permute(0, 3, 1, 2)
        with torch.enable_grad():                                  # called from inside an autograd backward
            xr = torch.zeros(n, c, hi, wi, requires_grad=True)
            F.interpolate(xr, (ho, wo), mode="bilinear", align_corners=bool(align)).backward(g)
        o = view(dx, (n, hi, wi, dx_ld))
        sc = float(view(gscale, (1,))[0]) if (gscale is not None and getattr(gscale, "value", gscale)) else 1.0
        o[..., :c] = xr.grad.permute(0, 2, 3, 1).numpy() * sc + (o[..., :c] if accumulate else 0)
        return 0
    for f in (segb200_adaptive_avgpool, segb200_adaptive_avgpool_bwd, segb200_bilinear_nhwc, segb200_bilinear_nhwc_bwd):
        setattr(cls, f.__name__, f)


_pooling(FakeLib)

"""Per-kernel parity on the B200: every C-ABI kernel against the same op in plain PyTorch fp32
(inputs/weights rounded to the kernel's 16-bit dtype first, so the only differences are the output
rounding and summation order).  Tolerance: |err| <= tol * |ref| + tol * rms(ref) with tol = 2^-7 for bf16 outputs (8 bits of
mantissa) and 2^-10 for fp16 outputs (11 bits) -- one output rounding plus slack for fp32 summation order."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def _close(got, ref, what, tol=None):
    if tol is None:
        tol = 2.0 ** -10 if getattr(got, "src_dtype", got.dtype) == torch.float16 else 2.0 ** -7
    got, ref = got.float(), ref.float()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    rms = float(ref.pow(2).mean().sqrt()) + 1e-12
    err = (got - ref).abs()
    bound = tol * ref.abs() + tol * rms
    bad = err > bound
    if bad.any():
        idx = bad.nonzero()[0].tolist()
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} elements out of tolerance; first at {idx}: "
                             f"got {float(got[tuple(idx)])} ref {float(ref[tuple(idx)])}; max err {float(err.max())}, "
                             f"rms {rms}, rel-L2 {float((got - ref).norm() / ref.norm())}")


def _rand(*shape, dtype, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


def _to_nchw(x_nhwc):
    t = x_nhwc.float().permute(0, 3, 1, 2).contiguous()
    t.src_dtype = x_nhwc.dtype                      # _close picks the output-rounding tolerance from the kernel's dtype
    return t


CONV_CASES = [
    # name, n,h,w, cin,cout, k, stride, dil, pad, act, residual, out slice (ld, off)
    ("pw_flat_64_128", 1, 25, 40, 64, 128, 1, 1, 1, 0, "relu", False, None),
    ("pw_flat_728_728_res", 2, 33, 65, 728, 728, 1, 1, 1, 0, None, True, None),
    ("pw_flat_304_256", 1, 33, 65, 304, 256, 1, 1, 1, 0, "relu", False, None),
    ("pw_flat_2048_256_slice", 1, 17, 33, 2048, 256, 1, 1, 1, 0, "relu", False, (1280, 512)),
    ("pw_flat_256_48_slice", 1, 33, 65, 256, 48, 1, 1, 1, 0, "relu", False, (304, 256)),
    ("pw_flat_256_24_cls", 1, 33, 65, 256, 24, 1, 1, 1, 0, None, False, (32, 0)),
    ("pw_flat_tinyM", 8, 1, 1, 2048, 256, 1, 1, 1, 0, "relu", False, None),
    ("pw_flat_1536_2048", 1, 17, 33, 1536, 2048, 1, 1, 1, 0, "relu", False, None),
    ("pw_s2_64_128", 2, 33, 65, 64, 128, 1, 2, 1, 0, None, False, None),
    ("pw_s2_256_728_even", 1, 34, 66, 256, 728, 1, 2, 1, 0, None, False, None),
    ("c3_32_64", 2, 37, 53, 32, 64, 3, 1, 1, 1, "relu", False, None),
    ("c3_64_64_d2", 1, 33, 65, 64, 64, 3, 1, 2, 2, "relu", True, None),
    ("c3_128_128_d4", 1, 33, 65, 128, 128, 3, 1, 4, 4, "relu", False, None),
    ("c3_64_128_s2", 2, 33, 65, 64, 128, 3, 2, 1, 1, "relu", False, None),
    ("c3_16_32", 1, 20, 24, 16, 32, 3, 1, 1, 1, "relu6", False, None),
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_gemm(case, dtype, max_ctas=0):
    from segmentron_b200 import fold, ops
    name, n, h, w, cin, cout, k, stride, dil, pad, act, use_res, sl = case
    x = _rand(n, h, w, cin, dtype=dtype, seed=1)
    wt = _rand(cout, cin, k, k, dtype=dtype, seed=2, scale=1.0 / math.sqrt(cin * k * k))
    scale = (0.5 + torch.rand(cout, generator=torch.Generator().manual_seed(3))).cuda()
    shift = (0.2 * torch.randn(cout, generator=torch.Generator().manual_seed(4))).cuda()
    ho = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1
    wo = (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
    res = _rand(n, ho, wo, cout, dtype=dtype, seed=5) if use_res else None
    if sl is None:
        buf = torch.full((n, ho, wo, cout), float("nan"), dtype=dtype, device="cuda")
        y = buf
    else:
        ld, off = sl
        buf = torch.full((n, ho, wo, ld), 7.0, dtype=dtype, device="cuda")
        y = buf[..., off:off + cout]
    wpk = fold.pack_conv_weight(wt, dtype)
    ops.conv_gemm(x, wpk, y, cin=cin, cout=cout, kh=k, kw=k, stride=stride, dilation=dil, pad_t=pad, pad_l=pad,
                  scale=scale, shift=shift, act=act, residual=res, max_ctas=max_ctas)
    torch.cuda.synchronize()
    ref = F.conv2d(_to_nchw(x), wt.float(), None, stride, pad, dil)
    ref = ref * scale[None, :, None, None] + shift[None, :, None, None]
    if use_res:
        ref = ref + _to_nchw(res)
    ref = F.relu(ref) if act == "relu" else (F.relu6(ref) if act == "relu6" else ref)
    _close(_to_nchw(y), ref, name)
    if sl is not None:                      # bytes outside the slice untouched
        ld, off = sl
        mask = torch.ones(ld, dtype=torch.bool); mask[off:off + cout] = False
        assert (buf[..., mask.cuda()] == 7.0).all(), f"{name}: wrote outside its channel slice"
    return y.clone()  # (used by the A/B tests below)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("max_ctas,subk", [(1, 0), (3, 0), (3, 1)])
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_gemm_dual_streams(case, max_ctas, subk, dtype):
    """The opt-in two-tile-streams-per-CTA kernel (two producer warps + two MMA-issuing warps, conv_gemm_kernel<..., kDual = true>;
    segb200_set_option("gemm_dual", 2), optionally walking half k-blocks, "gemm_dual_subk") with a 1- or 3-CTA persistent grid so
    that every CTA walks several tiles, odd and even counts; against the torch reference and BIT-identical to the single-stream
    kernel (a tile's MMA order does not change).  Slower than the default on B200, kept as a measured experiment."""
    from segmentron_b200 import lib as L
    lib = L.load()
    outs = []
    for dual in (2, 0):
        L.check(lib.segb200_set_option(b"gemm_dual", dual))
        L.check(lib.segb200_set_option(b"gemm_dual_subk", subk))
        try:
            outs.append(test_conv_gemm(case, dtype, max_ctas=max_ctas))
        finally:
            L.check(lib.segb200_set_option(b"gemm_dual", 0))
            L.check(lib.segb200_set_option(b"gemm_dual_subk", 0))
    assert torch.equal(outs[0], outs[1]), "dual-stream result differs from the single-stream kernel"


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_gemm_cta_pair(case):
    """the opt-in CTA-pair kernel (tcgen05 cta_group::2, 256 x BN tiles; segb200_set_option("gemm_2cta", 1)) on every conv case;
    shapes it does not take (N tile not splittable, single M tile) fall through to the single-CTA kernel"""
    from segmentron_b200 import lib as L
    lib = L.load()
    L.check(lib.segb200_set_option(b"gemm_2cta", 1))
    try:
        test_conv_gemm(case, torch.bfloat16)
    finally:
        L.check(lib.segb200_set_option(b"gemm_2cta", 0))


def test_conv_gemm_fp32_out():
    """fp32 epilogue (attention energies): K = 4096 accumulation, y_f32 store path."""
    from segmentron_b200 import fold, ops
    dtype = torch.bfloat16
    x = _rand(1, 1, 200, 4096, dtype=dtype, seed=1)
    wt = _rand(136, 4096, 1, 1, dtype=dtype, seed=2)
    y = torch.full((1, 1, 200, 136), float("nan"), dtype=torch.float32, device="cuda")
    ops.conv_gemm(x, fold.pack_conv_weight(wt, dtype), y, cin=4096, cout=136)
    torch.cuda.synchronize()
    ref = x.float().view(200, 4096) @ wt.float().view(136, 4096).t()
    assert torch.allclose(y.view(200, 136), ref, rtol=1e-4, atol=1e-2), float((y.view(200, 136) - ref).abs().max())


@pytest.mark.parametrize("k,pad,cout", [(3, 1, 32), (7, 3, 64)])
def test_stem_s2d(k, pad, cout):
    from segmentron_b200 import fold, ops
    dtype = torch.bfloat16
    n, h, w = 2, 65, 129
    x = _rand(n, 3, h, w, dtype=torch.float32, seed=1)
    wt = _rand(cout, 3, k, k, dtype=dtype, seed=2, scale=0.3)
    wpk, T, pad2, ld = fold.pack_stem_s2d(wt, pad, dtype)
    hs, ws = (h + 1) // 2, (w + 1) // 2
    s2d = torch.empty(n, hs, ws, ld, dtype=dtype, device="cuda")
    ops.pack_s2d(x, s2d)
    ho, wo = (h + 2 * pad - k) // 2 + 1, (w + 2 * pad - k) // 2 + 1
    y = torch.empty(n, ho, wo, cout, dtype=dtype, device="cuda")
    ops.conv_gemm(s2d, wpk, y, cin=ld, cout=cout, kh=T, kw=T, pad_t=pad2, pad_l=pad2, act="relu")
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(x.to(dtype).float(), wt.float(), None, 2, pad))
    _close(_to_nchw(y), ref, f"stem{k}")


DW_CASES = [(2, 33, 65, 728, 1, 1, True, None), (1, 33, 65, 128, 2, 1, True, None), (1, 34, 66, 256, 2, 1, True, None),
            (1, 33, 65, 2048, 1, 6, False, "relu"), (1, 33, 65, 256, 1, 18, False, "relu"),
            (2, 17, 19, 304, 1, 1, False, "relu"), (1, 16, 32, 96, 1, 2, False, "relu6"), (1, 9, 9, 8, 1, 12, False, None)]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("case", DW_CASES, ids=[f"dw{i}" for i in range(len(DW_CASES))])
def test_dwconv(case, dtype):
    from segmentron_b200 import fold, ops
    n, h, w, c, stride, dil, pre_relu, act = case
    x = _rand(n, h, w, c, dtype=dtype, seed=1)
    wt = _rand(c, 1, 3, 3, dtype=torch.float32, seed=2, scale=0.4)
    scale = (0.5 + torch.rand(c, generator=torch.Generator().manual_seed(3))).cuda()
    shift = (0.2 * torch.randn(c, generator=torch.Generator().manual_seed(4))).cuda()
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    y = torch.full((n, ho, wo, c), float("nan"), dtype=dtype, device="cuda")
    ops.dwconv3x3(x, fold.pack_dw_weight(wt, scale), y, stride=stride, dilation=dil, shift=shift, pre_relu=pre_relu, act=act)
    torch.cuda.synchronize()
    xin = _to_nchw(x)
    if pre_relu:
        xin = F.relu(xin)
    ref = F.conv2d(xin, wt, None, stride, dil, dil, groups=c) * scale[None, :, None, None] + shift[None, :, None, None]
    ref = F.relu(ref) if act == "relu" else (F.relu6(ref) if act == "relu6" else ref)
    _close(_to_nchw(y), ref, "dwconv")


DW2_CASES = [(2, 33, 65, 728, 1, True, None), (1, 40, 29, 64, 1, False, "relu"), (1, 7, 28, 304, 1, True, "relu6"), (2, 70, 57, 128, 1, False, None),
             (1, 1, 31, 72, 1, False, None), (1, 65, 129, 1536, 1, True, None),
             (1, 65, 129, 1536, 2, True, None), (1, 33, 65, 2048, 6, False, "relu"), (2, 33, 65, 256, 18, False, "relu"), (1, 16, 32, 96, 2, False, "relu6"),
             (1, 9, 29, 64, 12, True, None), (1, 65, 129, 128, 36, False, "relu"), (1, 70, 30, 64, 64, False, None)]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("case", DW2_CASES, ids=[f"dw2_{i}" for i in range(len(DW2_CASES))])
def test_dwconv_two_column_kernel(case, dtype):
    """stride 1 (any dilation) runs the two-columns-per-thread ring kernel (csrc/dwconv.cu dwconv3x3_ring4x2_kernel): against the fp32
    torch reference, BIT-identical to the one-column kernel (segb200_set_option("dw_cols2", 0); same fp32 operation order per output),
    into a channel slice of a wider buffer (y_ld > c) with the neighbouring channels untouched, from a channel slice (x_ld > c)."""
    from segmentron_b200 import fold, lib as L, ops
    lib = L.load()
    n, h, w, c, dil, pre_relu, act = case
    xbuf = _rand(n, h, w, c + 24, dtype=dtype, seed=11)
    x = xbuf[..., 8:8 + c]
    wt = _rand(c, 1, 3, 3, dtype=torch.float32, seed=12, scale=0.4)
    scale = (0.5 + torch.rand(c, generator=torch.Generator().manual_seed(13))).cuda()
    shift = (0.2 * torch.randn(c, generator=torch.Generator().manual_seed(14))).cuda()
    outs = []
    # (dw_cols2, dw_cw5): two-column kernel with 5 consumer warps / 3 CTAs per SM (opt-in, dilation 1), with 7 warps / 2 CTAs (default),
    # and the one-column kernel
    for cols2, cw5 in ((1, 1), (1, 0), (0, 1)):
        L.check(lib.segb200_set_option(b"dw_cols2", cols2))
        L.check(lib.segb200_set_option(b"dw_cw5", cw5))
        try:
            ybuf = torch.full((n, h, w, c + 16), 7.0, dtype=dtype, device="cuda")
            ybuf[..., 8:8 + c] = float("nan")
            ops.dwconv3x3(x, fold.pack_dw_weight(wt, scale), ybuf[..., 8:8 + c], stride=1, dilation=dil, shift=shift, pre_relu=pre_relu, act=act)
            torch.cuda.synchronize()
        finally:
            L.check(lib.segb200_set_option(b"dw_cols2", 1))
            L.check(lib.segb200_set_option(b"dw_cw5", 0))
        assert (ybuf[..., :8] == 7.0).all() and (ybuf[..., 8 + c:] == 7.0).all(), "wrote outside its channel slice"
        outs.append(ybuf[..., 8:8 + c].clone())
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[2]) and torch.equal(outs[1], outs[2]), "two-column kernel differs from the one-column kernel"
    xin = _to_nchw(x)
    if pre_relu:
        xin = F.relu(xin)
    ref = F.conv2d(xin, wt, None, 1, dil, dil, groups=c) * scale[None, :, None, None] + shift[None, :, None, None]
    ref = F.relu(ref) if act == "relu" else (F.relu6(ref) if act == "relu6" else ref)
    _close(_to_nchw(outs[0]), ref, "dwconv two-column")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_pool_and_resize(dtype):
    from segmentron_b200 import ops
    n, h, w, c = 2, 33, 65, 256
    x = _rand(n, h, w, c, dtype=dtype, seed=1)
    xn = _to_nchw(x)
    # global average pool
    out = torch.empty(n, 1, 1, c, dtype=dtype, device="cuda")
    ops.global_avgpool(x, out)
    _close(_to_nchw(out), F.adaptive_avg_pool2d(xn, 1), "gap")
    # adaptive pools (PSP sizes)
    for s in (1, 2, 3, 6):
        o = torch.empty(n, s, s, c, dtype=dtype, device="cuda")
        ops.adaptive_avgpool(x, o, s)
        _close(_to_nchw(o), F.adaptive_avg_pool2d(xn, s), f"adaptive{s}")
    # bilinear, both corner modes, up and from 1x1, into a channel slice
    for align in (True, False):
        buf = torch.zeros(n, 129, 257, c + 48, dtype=dtype, device="cuda")
        ops.bilinear_nhwc(x, buf[..., :c], align_corners=align)
        _close(_to_nchw(buf[..., :c]), F.interpolate(xn, (129, 257), mode="bilinear", align_corners=align), f"bilinear{align}")
        assert (buf[..., c:] == 0).all()
    y1 = torch.empty(n, 17, 33, c, dtype=dtype, device="cuda")
    ops.bilinear_nhwc(out, y1, align_corners=True)
    _close(_to_nchw(y1), F.interpolate(_to_nchw(out), (17, 33), mode="bilinear", align_corners=True), "broadcast")
    # final logits upsample to NCHW + fused argmax
    lg = _rand(n, h, w, 32, dtype=dtype, seed=9)
    for od in (dtype, torch.float32):
        yo = torch.empty(n, 19, 4 * h - 3, 4 * w - 3, dtype=od, device="cuda")
        am = torch.empty(n, 4 * h - 3, 4 * w - 3, dtype=torch.uint8, device="cuda")
        ops.bilinear_nchw_out(lg, yo, 19, True, am)
        ref = F.interpolate(_to_nchw(lg)[:, :19], (4 * h - 3, 4 * w - 3), mode="bilinear", align_corners=True)
        _close(yo, ref, "logits_up")
        assert (am.long() == yo.float().argmax(1)).all(), "fused argmax != torch.argmax of the same output"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("case", [(2, 17, 33, 19, 24, 65, 129, True), (1, 33, 65, 19, 32, 129, 257, True), (2, 9, 13, 21, 24, 70, 59, False),
                                  (1, 16, 16, 2, 8, 64, 64, False), (1, 5, 7, 32, 32, 5, 7, True), (1, 12, 20, 11, 16, 45, 131, True)],
                         ids=lambda c: "x".join(str(int(v)) for v in c))
def test_logits_upsample_strip_kernel(case, dtype):
    """segb200_bilinear_nchw_out for c <= 32 runs the strip kernel (misc.cu bilinear_nchw_strip_kernel: one column x 4 rows per
    thread, compile-time dtypes): logits and fused argmax BIT-identical to the one-pixel-per-thread kernel
    (segb200_set_option("bilinear_out_v1", 1)) for every output dtype, both align_corners modes, integer and non-integer scales;
    and against torch's bilinear interpolation."""
    from segmentron_b200 import lib as L, ops
    lib = L.load()
    n, hi, wi, c, ld, ho, wo, align = case
    lg = _rand(n, hi, wi, ld, dtype=dtype, seed=21)
    for od in (dtype, torch.float32):
        res = []
        for v1 in (0, 1):
            L.check(lib.segb200_set_option(b"bilinear_out_v1", v1))
            try:
                yo = torch.full((n, c, ho, wo), float("nan"), dtype=od, device="cuda")
                am = torch.full((n, ho, wo), 255, dtype=torch.uint8, device="cuda")
                ops.bilinear_nchw_out(lg, yo, c, align, am)
                torch.cuda.synchronize()
            finally:
                L.check(lib.segb200_set_option(b"bilinear_out_v1", 0))
            res.append((yo, am))
        assert torch.isfinite(res[0][0]).all()
        assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
        assert (res[0][1].long() == res[0][0].float().argmax(1)).all(), "fused argmax != torch.argmax of the same output"
        ref = F.interpolate(_to_nchw(lg)[:, :c], (ho, wo), mode="bilinear", align_corners=align)
        _close(res[0][0], ref, "logits_up(strip)")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("case", [(2, 17, 33, 256, 65, 129, True), (1, 9, 13, 48, 70, 59, False), (2, 1, 1, 256, 33, 65, True),
                                  (1, 33, 65, 512, 17, 33, True), (1, 6, 6, 64, 33, 33, False)],
                         ids=lambda c: "x".join(str(int(v)) for v in c))
def test_bilinear_nhwc_strip_kernel(case, dtype):
    """segb200_bilinear_nhwc runs the strip kernel (one column x 8 channels x 4 rows per thread): BIT-identical to the one-output-
    per-thread kernel (segb200_set_option("bilinear_out_v1", 1)) for up- and down-sampling, broadcast from 1x1, both align_corners
    modes, into a channel slice; and against torch."""
    from segmentron_b200 import lib as L, ops
    lib = L.load()
    n, hi, wi, c, ho, wo, align = case
    x = _rand(n, hi, wi, c, dtype=dtype, seed=31)
    res = []
    for v1 in (0, 1):
        L.check(lib.segb200_set_option(b"bilinear_out_v1", v1))
        try:
            buf = torch.full((n, ho, wo, c + 16), 3.0, dtype=dtype, device="cuda")
            ops.bilinear_nhwc(x, buf[..., 8:8 + c], align_corners=align)
            torch.cuda.synchronize()
        finally:
            L.check(lib.segb200_set_option(b"bilinear_out_v1", 0))
        assert (buf[..., :8] == 3.0).all() and (buf[..., 8 + c:] == 3.0).all()
        res.append(buf[..., 8:8 + c].clone())
    assert torch.equal(res[0], res[1])
    _close(_to_nchw(res[0]), F.interpolate(_to_nchw(x), (ho, wo), mode="bilinear", align_corners=align), "bilinear_nhwc(strip)")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_maxpool(dtype):
    from segmentron_b200 import ops
    for (n, h, w, c) in [(2, 33, 65, 64), (1, 34, 66, 128)]:
        x = _rand(n, h, w, c, dtype=dtype, seed=1)
        y = torch.empty(n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c, dtype=dtype, device="cuda")
        ops.maxpool3x3s2(x, y)
        assert torch.equal(_to_nchw(y), F.max_pool2d(_to_nchw(x), 3, 2, 1))


def test_layout_converters():
    from segmentron_b200 import ops
    x = _rand(2, 19, 37, 41, dtype=torch.float32, seed=3)
    y = torch.zeros(2, 37, 41, 24, dtype=torch.bfloat16, device="cuda")
    ops.nchw_to_nhwc(x, y[..., :19])
    assert torch.equal(y[..., :19].permute(0, 3, 1, 2), x.to(torch.bfloat16))
    z = torch.empty(2, 19, 37, 41, dtype=torch.float32, device="cuda")
    ops.nhwc_to_nchw(y[..., :19], z)
    assert torch.equal(z, x.to(torch.bfloat16).float())


def test_errors_are_loud():
    from segmentron_b200 import ops
    x = torch.zeros(1, 4, 4, 64, dtype=torch.bfloat16)          # CPU tensor
    with pytest.raises(RuntimeError):
        ops.dwconv3x3(x, x, x)
    xc = torch.zeros(1, 4, 4, 60, dtype=torch.bfloat16, device="cuda")   # 60 channels: not a multiple of 8
    with pytest.raises(RuntimeError):
        ops.dwconv3x3(xc, torch.zeros(9, 60, device="cuda"), torch.empty_like(xc))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("case", [(1, 300, 256, 256, False), (2, 1000, 256, 512, True), (1, 64, 256, 256, False), (2, 129, 64, 256, True),
                                  (1, 8192, 256, 256, False)],
                         ids=lambda c: "x".join(str(int(v)) for v in c))
def test_nonlocal_attention(case, dtype):
    """segb200_nonlocal_attention (csrc/pam.cu, query/key depth 64 or 256): y = gamma * (softmax(q k^T) v + b_v) + x against the
    fp32 torch evaluation of the same formula on the same 16-bit inputs, with and without the residual / gamma, ragged query and key
    tiles (n_tok not a multiple of 128 / 64), q and k as channel slices of one tensor (OCNet: shared key/query transform), and
    N = 8192 = a 1024x2048 image at output stride 16."""
    import ctypes as C
    from segmentron_b200 import lib as L, ops
    lib = L.load()
    b, n, dk, dv, with_x = case
    g = torch.Generator().manual_seed(n + dk)
    qk = (torch.randn(b, n, 2 * dk, generator=g) * (2.0 / dk ** 0.5)).to(dtype).cuda()          # energies O(4)
    q, k = qk[..., :dk], qk[..., dk:]
    v = torch.randn(b, n, dv, generator=g).to(dtype).cuda()
    pitch = (n + 7) // 8 * 8
    vt = torch.zeros(b, dv, pitch, dtype=dtype, device="cuda")
    vt[..., :n] = v.transpose(1, 2)
    bv = (0.1 * torch.randn(dv, generator=g)).cuda()
    x = torch.randn(b, n, dv, generator=g).to(dtype).cuda() if with_x else None
    gamma = torch.tensor([0.7], device="cuda") if with_x else None
    y = torch.full((b, n, dv), float("nan"), dtype=dtype, device="cuda")
    sm = torch.empty(b * n, device="cuda"); sl = torch.empty(b * n, device="cuda")
    L.check(lib.segb200_nonlocal_attention(ops._ptr(q), ops._ptr(k), ops._ptr(vt), ops._ptr(bv), ops._ptr(gamma), ops._ptr(x), ops._ptr(y),
                                           ops._ptr(sm), ops._ptr(sl), b, n, dk, dv, 2 * dk, 2 * dk, pitch, dv if with_x else 0, dv,
                                           ops.dt_code(dtype), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "nonlocal_attention")
    torch.cuda.synchronize()
    att = torch.softmax(torch.bmm(q.float(), k.float().transpose(1, 2)), dim=-1)
    ref = torch.bmm(att, v.float()) + bv
    if with_x:
        ref = 0.7 * ref + x.float()
    got = y.float()
    assert torch.isfinite(got).all()
    # P is rounded to 16 bits before the P V product (as the reference's 16-bit bmm would): 2^-7 / 2^-9 of the output scale
    tol = 2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -9
    err = float((got - ref).abs().max() / ref.abs().max())
    assert err < tol, (err, tol)

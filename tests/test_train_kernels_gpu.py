"""Per-kernel parity of the TRAINING path on the B200: every backward / train-mode kernel against torch autograd of the
same op in fp32 on inputs rounded to the kernel's 16-bit dtype (SURVEY.md 8a rows a15/a16, appendix D semantics).
Tolerance as in test_kernels_gpu.py: |err| <= tol*|ref| + tol*rms(ref), tol = 2^-7 unless stated."""
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


def _close(got, ref, what, tol=2.0 ** -7, max_bad_frac=0.0):
    """max_bad_frac: fraction of elements allowed outside the bound (ReLU-mask flips at |pre-activation| ~ 1e-7)."""
    got, ref = got.float(), ref.float()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    rms = float(ref.pow(2).mean().sqrt()) + 1e-12
    err = (got - ref).abs()
    bad = err > tol * ref.abs() + tol * rms
    if int(bad.sum()) > max_bad_frac * bad.numel():
        idx = bad.nonzero()[0].tolist()
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} out of tolerance; first at {idx}: got "
                             f"{float(got[tuple(idx)])} ref {float(ref[tuple(idx)])}; max err {float(err.max())}, rms {rms}, "
                             f"rel-L2 {float((got - ref).norm() / (ref.norm() + 1e-30))}")


def _rand(*shape, dtype, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


def _nchw(x_nhwc):
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous()


def _nhwc(x_nchw, dtype):
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(dtype)


def _out_hw(h, w, k, s, p, d):
    return (h + 2 * p - d * (k - 1) - 1) // s + 1, (w + 2 * p - d * (k - 1) - 1) // s + 1


WGRAD_CASES = [
    # name, n,h,w, cin,cout, k, stride, dil, pad
    ("pw_flat_64_256", 2, 33, 47, 64, 256, 1, 1, 1, 0),
    ("pw_flat_256_64", 2, 33, 47, 256, 64, 1, 1, 1, 0),
    ("pw_flat_304_48", 1, 33, 65, 304, 48, 1, 1, 1, 0),
    ("pw_flat_256_19", 1, 33, 65, 256, 19, 1, 1, 1, 0),
    ("pw_flat_1280_256", 1, 17, 33, 1280, 256, 1, 1, 1, 0),
    ("pw_tinyM", 4, 1, 1, 2048, 256, 1, 1, 1, 0),
    ("c3_128_128", 2, 33, 65, 128, 128, 3, 1, 1, 1),
    ("c3_64_64_d2", 1, 33, 65, 64, 64, 3, 1, 2, 2),
    ("c3_512_512_d2", 1, 17, 33, 512, 512, 3, 1, 2, 2),
    ("c3_64_128_s2", 2, 33, 65, 64, 128, 3, 2, 1, 1),
    ("pw_s2_256_512", 2, 33, 65, 256, 512, 1, 2, 1, 0),
    ("c4_stem_64_64", 2, 33, 65, 64, 64, 4, 1, 1, 2),
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("case", WGRAD_CASES, ids=[c[0] for c in WGRAD_CASES])
def test_conv_wgrad(case, dtype):
    from segmentron_b200 import train_ops as T
    name, n, h, w, cin, cout, k, s, d, p = case
    ho, wo = _out_hw(h, w, k, s, p, d)
    cop = (cout + 7) // 8 * 8
    x = _rand(n, h, w, cin, dtype=dtype, seed=1)
    dyb = torch.zeros(n, ho, wo, max(cop, 32), dtype=dtype, device="cuda")
    dyb[..., :cout] = _rand(n, ho, wo, cout, dtype=dtype, seed=2)
    dy = dyb[..., :cop]
    xr = _nchw(x).requires_grad_(False)
    wr = torch.zeros(cout, cin, k, k, device="cuda", requires_grad=True)
    yr = F.conv2d(xr, wr, None, s, p, d)
    yr.backward(_nchw(dy[..., :cout]))
    ref = wr.grad.permute(0, 2, 3, 1).reshape(cout, k * k, cin)
    for splits in (0, 1, 3):
        dw = torch.zeros(cout, k * k, cin, dtype=torch.float32, device="cuda")
        T.conv_wgrad(x, dy, dw, cin=cin, cout=cout, kh=k, kw=k, stride=s, dilation=d, pad_t=p, pad_l=p, splits=splits)
        torch.cuda.synchronize()
        _close(dw, ref, f"wgrad {name} splits={splits}", tol=2.0 ** -9)
    # accumulation semantics: a second call doubles the result
    T.conv_wgrad(x, dy, dw, cin=cin, cout=cout, kh=k, kw=k, stride=s, dilation=d, pad_t=p, pad_l=p, splits=3)
    _close(dw, 2 * ref, f"wgrad {name} accumulate", tol=2.0 ** -9)


DGRAD_CASES = [
    ("pw_256_64", 2, 33, 47, 256, 64, 1, 1, 1, 0),
    ("pw_304_48", 1, 33, 65, 304, 48, 1, 1, 1, 0),
    ("c3_128_128", 2, 33, 65, 128, 128, 3, 1, 1, 1),
    ("c3_512_512_d2", 1, 17, 33, 512, 512, 3, 1, 2, 2),
    ("c3_64_128_s2", 2, 33, 65, 64, 128, 3, 2, 1, 1),
    ("c3_64_128_s2_even", 1, 34, 66, 64, 128, 3, 2, 1, 1),
    ("pw_s2_256_512", 2, 33, 65, 256, 512, 1, 2, 1, 0),
]


@pytest.mark.parametrize("case", DGRAD_CASES, ids=[c[0] for c in DGRAD_CASES])
def test_conv_dgrad(case):
    """Data gradient = the forward GEMM kernel on dY with transposed, tap-flipped weights; stride 2 through stride2_place."""
    from segmentron_b200 import ops, train_ops as T
    dtype = torch.bfloat16
    name, n, h, w, cin, cout, k, s, d, p = case
    ho, wo = _out_hw(h, w, k, s, p, d)
    wt = _rand(cout, cin, k, k, dtype=dtype, seed=3, scale=(cin * k * k) ** -0.5)
    dy = _rand(n, ho, wo, cout, dtype=dtype, seed=4)
    xr = torch.zeros(n, cin, h, w, device="cuda", requires_grad=True)
    F.conv2d(xr, wt.float(), None, s, p, d).backward(_nchw(dy))
    ref = xr.grad.permute(0, 2, 3, 1)
    wpk = T.pack_dgrad_weight(wt, dtype)
    dx = torch.zeros(n, h, w, cin, dtype=dtype, device="cuda")
    if s == 1:
        ops.conv_gemm(dy, wpk, dx, cin=cout, cout=cin, kh=k, kw=k, dilation=d, pad_t=p, pad_l=p)
    elif k == 1:
        t = torch.empty(n, ho, wo, cin, dtype=dtype, device="cuda")
        ops.conv_gemm(dy, wpk, t, cin=cout, cout=cin)
        base = _rand(n, h, w, cin, dtype=dtype, seed=5)
        dx.copy_(base)
        T.stride2_place(t, dx, 1)
        ref = ref + base.float()
    else:
        z = torch.empty(n, h, w, cout, dtype=dtype, device="cuda")
        T.stride2_place(dy, z, 0)
        ops.conv_gemm(z, wpk, dx, cin=cout, cout=cin, kh=k, kw=k, dilation=1, pad_t=p, pad_l=p)
    torch.cuda.synchronize()
    _close(dx, ref, f"dgrad {name}")


BN_CASES = [("c64", 2, 33, 47, 64), ("c256_res", 2, 17, 33, 256), ("c48", 1, 33, 65, 48), ("c2048_tinyM", 4, 1, 1, 2048),
            ("c24", 2, 33, 65, 24)]


@pytest.mark.parametrize("act", [None, "relu"])
@pytest.mark.parametrize("case", BN_CASES, ids=[c[0] for c in BN_CASES])
def test_batchnorm_train(case, act):
    from segmentron_b200 import train_ops as T
    dtype = torch.bfloat16
    name, n, h, w, c = case
    use_res = "res" in name
    y = _rand(n, h, w, c, dtype=dtype, seed=6, scale=2.0) + 0.5
    res = _rand(n, h, w, c, dtype=dtype, seed=7) if use_res else None
    gamma = (0.75 + 0.5 * torch.rand(c, generator=torch.Generator().manual_seed(8))).cuda()
    beta = (0.2 * torch.randn(c, generator=torch.Generator().manual_seed(9))).cuda()
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    rm_ref, rv_ref = rm.clone(), rv.clone()
    nc = (torch.rand(n, c, generator=torch.Generator().manual_seed(10)) > 0.2).float().cuda() / 0.8 if act == "relu" else None
    z = torch.empty(n, h, w, c, dtype=dtype, device="cuda")
    st = T.bn_forward(y, z, gamma, beta, rm, rv, 0.1, 1e-5, act=act, residual=res, nc_scale=nc)
    # reference
    yr = _nchw(y).requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rr = _nchw(res).requires_grad_(True) if use_res else None
    o = F.batch_norm(yr, rm_ref, rv_ref, gr, br, True, 0.1, 1e-5)
    if use_res:
        o = o + rr
    if act == "relu":
        o = F.relu(o)
    if nc is not None:
        o = o * nc[:, :, None, None]
    _close(z, o.permute(0, 2, 3, 1), f"bn fwd {name}")
    _close(rm, rm_ref, f"running_mean {name}", tol=1e-4)
    _close(rv, rv_ref, f"running_var {name}", tol=1e-4)
    # backward: use OUR z for the activation mask in the reference too (mask = z > 0), by feeding dz that is zero where the
    # two masks could differ is unnecessary: compare with the reference grads, masks agree except at |pre-act| ~ 0
    dz = _rand(n, h, w, c, dtype=dtype, seed=11)
    o.backward(_nchw(dz))
    dy = torch.empty_like(y)
    dres = _rand(n, h, w, c, dtype=dtype, seed=12) if use_res else None
    dres0 = dres.clone() if use_res else None
    dgamma, dbeta = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    # the activation mask comes from z when a residual entered the activation, else it is recomputed from y
    T.bn_backward(dz, z if (act and use_res) else None, y, st, dy, dgamma, dbeta, act=act, dres=dres, dres_accumulate=True,
                  nc_scale=nc)
    torch.cuda.synchronize()
    _close(dy, yr.grad.permute(0, 2, 3, 1), f"bn dy {name}", tol=2.0 ** -6, max_bad_frac=1e-4)
    _close(dgamma, gr.grad, f"bn dgamma {name}", tol=2.0 ** -7)
    _close(dbeta, br.grad, f"bn dbeta {name}", tol=2.0 ** -7)
    if use_res:
        _close(dres, dres0.float() + rr.grad.permute(0, 2, 3, 1), f"bn dres {name}", max_bad_frac=1e-4)


def test_maxpool_bwd_first_max_rule():
    from segmentron_b200 import ops, train_ops as T
    dtype = torch.bfloat16
    n, h, w, c = 2, 33, 47, 64
    x = torch.relu(_rand(n, h, w, c, dtype=dtype, seed=13))          # many exact ties at 0 (post-ReLU input, as in ResNet)
    x = (x * 4).round() / 4                                             # and coarse values -> ties between positives too
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    dy = _rand(n, ho, wo, c, dtype=dtype, seed=14)
    xr = _nchw(x).requires_grad_(True)
    F.max_pool2d(xr, 3, 2, 1).backward(_nchw(dy))
    dx = torch.empty_like(x)
    T.maxpool3x3s2_bwd(x, dy, dx)
    torch.cuda.synchronize()
    _close(dx, xr.grad.permute(0, 2, 3, 1), "maxpool bwd")


def test_maxpool_index_pair():
    """forward with argmax taps + index-based backward (the pair the training plan uses) == F.max_pool2d autograd, ties included"""
    from segmentron_b200 import train_ops as T
    dtype = torch.bfloat16
    for (n, h, w, c) in ((2, 33, 47, 64), (1, 34, 66, 64), (1, 5, 7, 8)):
        x = torch.relu(_rand(n, h, w, c, dtype=dtype, seed=113))
        x = (x * 4).round() / 4
        ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        dy = _rand(n, ho, wo, c, dtype=dtype, seed=114)
        xr = _nchw(x).requires_grad_(True)
        yr = F.max_pool2d(xr, 3, 2, 1)
        yr.backward(_nchw(dy))
        y = torch.empty(n, ho, wo, c, dtype=dtype, device="cuda")
        idx = torch.empty(n, ho, wo, c, dtype=torch.uint8, device="cuda")
        T.maxpool3x3s2_idx(x, y, idx)
        dx = torch.empty_like(x)
        T.maxpool3x3s2_bwd_idx(idx, dy, dx)
        torch.cuda.synchronize()
        assert torch.equal(y.float(), yr.detach().permute(0, 2, 3, 1)), "maxpool forward (idx variant)"
        assert int(idx.max()) <= 8
        _close(dx, xr.grad.permute(0, 2, 3, 1), f"maxpool idx bwd {(n, h, w, c)}")


BIL_CASES = [("x4_ac", 2, 17, 33, 65, 129, 64, True), ("x4_ac_small", 1, 9, 9, 33, 33, 24, True),
             ("noalign_x2", 2, 16, 24, 32, 48, 32, False), ("bcast", 2, 1, 1, 17, 33, 64, True),
             ("same", 1, 17, 33, 17, 33, 16, True), ("down", 1, 33, 65, 17, 33, 16, True)]


@pytest.mark.parametrize("case", BIL_CASES, ids=[c[0] for c in BIL_CASES])
def test_bilinear_bwd(case):
    from segmentron_b200 import train_ops as T
    dtype = torch.bfloat16
    name, n, hi, wi, ho, wo, c, ac = case
    dy = _rand(n, ho, wo, c, dtype=dtype, seed=15)
    xr = torch.zeros(n, c, hi, wi, device="cuda", requires_grad=True)
    F.interpolate(xr, (ho, wo), mode="bilinear", align_corners=ac).backward(_nchw(dy))
    base = _rand(n, hi, wi, c, dtype=dtype, seed=16)
    dx = base.clone()
    gs = torch.tensor([0.5], device="cuda")
    T.bilinear_nhwc_bwd(dy, dx, align_corners=ac, accumulate=True, gscale=gs)
    torch.cuda.synchronize()
    _close(dx, base.float() + 0.5 * xr.grad.permute(0, 2, 3, 1), f"bilinear bwd {name}")
    dx2 = torch.empty_like(dx)
    T.bilinear_nhwc_bwd(dy, dx2, align_corners=ac)
    _close(dx2, xr.grad.permute(0, 2, 3, 1), f"bilinear bwd {name} overwrite")


def test_upsample_cross_entropy():
    """Fused final upsample + CrossEntropyLoss(ignore_index=-1) (deeplabv3_plus.py:39 + loss.py:16-46): loss value and the
    gradient w.r.t. the low-resolution logits."""
    from segmentron_b200 import train_ops as T
    dtype = torch.bfloat16
    n, hi, wi, H, W, nclass = 2, 17, 33, 65, 129, 19
    lg = torch.zeros(n, hi, wi, 32, dtype=dtype, device="cuda")
    lg[..., :nclass] = _rand(n, hi, wi, nclass, dtype=dtype, seed=17, scale=2.0)
    target = torch.randint(-1, nclass, (n, H, W), generator=torch.Generator().manual_seed(18)).cuda()
    lr = _nchw(lg[..., :nclass]).requires_grad_(True)
    loss_ref = F.cross_entropy(F.interpolate(lr, (H, W), mode="bilinear", align_corners=True), target, ignore_index=-1)
    loss_ref.backward()
    dfull = torch.empty(n, H, W, 24, dtype=dtype, device="cuda")
    out3 = T.upsample_ce(lg, target, dfull, nclass)
    dlg = torch.zeros(n, hi, wi, 32, dtype=dtype, device="cuda")
    T.bilinear_nhwc_bwd(dfull, dlg[..., :24], align_corners=True, gscale=out3[1:2])
    torch.cuda.synchronize()
    assert abs(float(out3[0]) - float(loss_ref)) <= 2e-3 * abs(float(loss_ref)), (float(out3[0]), float(loss_ref))
    assert float(out3[2]) == float((target >= 0).sum())
    _close(dlg[..., :nclass], lr.grad.permute(0, 2, 3, 1), "dlogits", tol=2.0 ** -6)
    assert float(dlg[..., nclass:24].abs().max()) == 0.0


@pytest.mark.parametrize("dil", [1, 6])
def test_depthwise_backward(dil):
    from segmentron_b200 import ops, train_ops as T
    dtype = torch.bfloat16
    n, h, w, c = 2, 17, 33, 128
    x = _rand(n, h, w, c, dtype=dtype, seed=19)
    wt = _rand(c, 1, 3, 3, dtype=torch.float32, seed=20, scale=0.3)
    dy = _rand(n, h, w, c, dtype=dtype, seed=21)
    xr = _nchw(x).requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    F.conv2d(xr, wr, None, 1, dil, dil, groups=c).backward(_nchw(dy))
    dw = torch.zeros(c, 9, device="cuda")
    T.dw_wgrad(x, dy, dw, dilation=dil)
    _close(dw, wr.grad.reshape(c, 9), f"dw wgrad d{dil}", tol=2.0 ** -9)
    wflip = wt.reshape(c, 9).flip(1).t().contiguous()            # [9][c], taps reversed
    dx = torch.empty_like(x)
    ops.dwconv3x3(dy, wflip, dx, stride=1, dilation=dil)
    torch.cuda.synchronize()
    _close(dx, xr.grad.permute(0, 2, 3, 1), f"dw dgrad d{dil}")


def test_small_helpers():
    from segmentron_b200 import train_ops as T
    dtype = torch.bfloat16
    n, h, w, c = 2, 9, 13, 64
    v = _rand(n, 1, 1, c, dtype=dtype, seed=22)
    base = _rand(n, h, w, c, dtype=dtype, seed=23)
    y = base.clone()
    T.nc_broadcast(v, y, scale=0.25, accumulate=True)
    _close(y, base.float() + 0.25 * v.float(), "nc_broadcast acc")
    T.nc_broadcast(v, y, scale=2.0)
    _close(y, (2.0 * v.float()).expand(n, h, w, c), "nc_broadcast")
    # gather / scatter
    src = torch.randn(1000, generator=torch.Generator().manual_seed(24)).cuda()
    idx = torch.randperm(1000, generator=torch.Generator().manual_seed(25))[:600].int()
    idx[::7] = -1
    idx = idx.cuda()
    dst = torch.empty(600, dtype=dtype, device="cuda")
    T.gather_cast(src, idx, dst)
    ref = torch.where(idx >= 0, src[idx.clamp(min=0).long()], torch.zeros((), device="cuda"))
    assert torch.equal(dst, ref.to(dtype))
    g = torch.randn(600, generator=torch.Generator().manual_seed(26)).cuda()
    acc = torch.ones(1000, device="cuda")
    T.scatter_add(g, idx, acc)
    ref2 = torch.ones(1000, device="cuda")
    ref2.index_add_(0, idx[idx >= 0].long(), g[idx >= 0])
    assert torch.allclose(acc, ref2)


def test_sgd_matches_torch():
    from segmentron_b200 import train_ops as T
    nel = 100003
    p0 = torch.randn(nel, generator=torch.Generator().manual_seed(27)).cuda()
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([p], lr=0.02, momentum=0.9, weight_decay=1e-4)
    mine, m = p0.clone(), torch.zeros(nel, device="cuda")
    for it in range(3):
        g = torch.randn(nel, generator=torch.Generator().manual_seed(28 + it)).cuda()
        p.grad = g.clone()
        opt.step()
        T.sgd_step(mine, g, m, 0.02, 0.9, 1e-4)
    torch.cuda.synchronize()
    assert torch.allclose(mine, p.detach(), rtol=1e-5, atol=1e-6), float((mine - p.detach()).abs().max())


@pytest.mark.parametrize("geo", [(2, 17, 33, 128, 1, False), (2, 17, 33, 128, 6, False), (1, 65, 129, 728, 1, True), (4, 33, 41, 24, 2, True),
                                 (1, 5, 3, 8, 1, False)])
def test_depthwise_wgrad_v2(geo):
    """csrc/dw_wgrad2.cu (sliding window, four channels per thread, packed FMA) against autograd AND against the default kernel
    (same inputs: the two differ only in summation order)."""
    from segmentron_b200 import train_ops as T
    n, h, w, c, dil, pre_relu = geo
    dtype = torch.bfloat16
    x = _rand(n, h, w, c, dtype=dtype, seed=31)
    dy = _rand(n, h, w, c, dtype=dtype, seed=32)
    xr = _nchw(x)
    wr = torch.zeros(c, 1, 3, 3, device="cuda", requires_grad=True)
    F.conv2d(F.relu(xr) if pre_relu else xr, wr, None, 1, dil, dil, groups=c).backward(_nchw(dy))
    dw1 = torch.zeros(c, 9, device="cuda")
    dw2 = torch.zeros(c, 9, device="cuda")
    T.dw_wgrad(x, dy, dw1, dilation=dil, pre_relu=pre_relu)
    T.dw_wgrad(x, dy, dw2, dilation=dil, pre_relu=pre_relu, variant=2)
    T.dw_wgrad(x, dy, dw2, dilation=dil, pre_relu=pre_relu, variant=2)          # accumulates
    torch.cuda.synchronize()
    _close(dw2 * 0.5, wr.grad.reshape(c, 9), f"dw wgrad v2 {geo}", tol=2.0 ** -9)
    _close(dw2 * 0.5, dw1, f"dw wgrad v2 vs v1 {geo}", tol=2.0 ** -12)


@pytest.mark.parametrize("k,act", [(1, "relu"), (2, None), (3, "relu")])
def test_upsample_add_backward(k, act):
    """segb200_upsample_add_bwd (HRNet fuse sum) against autograd through relu(a + nearest_up(z)); overwrite and accumulate modes."""
    from segmentron_b200 import lib as L
    from segmentron_b200.ops import _ptr, _stream, dt_code
    dtype = torch.bfloat16
    n, c, hz, wz = 2, 32, 5, 7
    h, w = hz << k, wz << k
    a = _rand(n, h, w, c, dtype=dtype, seed=51)
    z = _rand(n, hz, wz, c, dtype=dtype, seed=52)
    dy = _rand(n, h, w, c, dtype=dtype, seed=53)
    lib = L.load()
    y = torch.empty_like(a)
    L.check(lib.segb200_upsample_add(_ptr(a), _ptr(z), _ptr(y), n, h, w, c, c, c, c, k, L.ACT[act], dt_code(dtype), _stream()))
    ar, zr = a.float().requires_grad_(True), z.float().requires_grad_(True)
    s = ar + zr.repeat_interleave(1 << k, 1).repeat_interleave(1 << k, 2)
    (F.relu(s) if act == "relu" else s).backward(dy.float())
    da0 = _rand(n, h, w, c, dtype=dtype, seed=54)
    dz0 = _rand(n, hz, wz, c, dtype=dtype, seed=55)
    for acc in (0, 1):
        da, dz = da0.clone(), dz0.clone()
        L.check(lib.segb200_upsample_add_bwd(_ptr(dy), _ptr(y), _ptr(da), _ptr(dz), n, h, w, c, c, c, c, c, k, L.ACT[act], acc, acc,
                                             dt_code(dtype), _stream()))
        torch.cuda.synchronize()
        _close(da, ar.grad + (da0.float() if acc else 0), f"upsample_add_bwd da k{k} acc{acc}")
        _close(dz, zr.grad + (dz0.float() if acc else 0), f"upsample_add_bwd dz k{k} acc{acc}")

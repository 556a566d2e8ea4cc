"""Drop-in nn.Modules (segmentron_b200/modules.py) on the B200 against the oracle's functional restatement of the same
reference modules, with the oracle's reference-named parameters loaded through load_state_dict(strict=True)."""
import os

import pytest
import torch

from oracle import segref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def _load(mod, P, prefix):
    sd = {k[len(prefix) + 1:]: v for k, v in P.state_dict().items() if k.startswith(prefix + ".")}
    mod.load_state_dict(sd, strict=True)
    return mod.cuda().eval()


def _cmp(y, ref, tol):
    rel = float((y.float().cpu() - ref).norm() / ref.norm())
    assert rel < tol, rel
    return rel


def _x(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 1.5e-3), (torch.bfloat16, 1e-2)], ids=["f16", "bf16"])
def test_dropin_modules(dtype, tol):
    from segmentron_b200 import modules as M
    with torch.no_grad():
        # SeparableConv2d, both orderings, stride/dilation
        for relu_first, stride, dil in [(True, 1, 1), (True, 2, 1), (False, 1, 2), (False, 1, 12)]:
            P = R.Params(1)
            x = _x(2, 64, 33, 41, seed=3)
            ref = R.separable_conv2d(P, x, "m", 128, stride, dil, relu_first, 1e-3)
            m = M.SeparableConv2d(64, 128, stride=stride, dilation=dil, relu_first=relu_first)
            for bn in (m.block.bn_depth, m.block.bn_point):
                bn.eps = 1e-3                                  # mutated after construction, like tools/eval.py:50-53
            m = _load(m, P, "m")
            y = m(x.cuda().to(dtype))
            assert y.shape == ref.shape and y.dtype == dtype
            _cmp(y, ref, tol)
            y32 = m(x.cuda())                                  # fp32 in -> bf16 compute -> fp32 out
            assert y32.dtype == torch.float32
        # _ConvBNReLU 3x3 dilated + relu6 1x1
        P = R.Params(2)
        x = _x(1, 64, 19, 23, seed=4)
        ref = R.conv_bn_act(P, x, "c", 96, 3, 1, 2, 2)
        _cmp(_load(M._ConvBNReLU(64, 96, 3, 1, 2, 2), P, "c")(x.cuda().to(dtype)), ref, tol)
        # InvertedResidual with and without the skip
        for cin, cout, stride, t in [(32, 32, 1, 6), (32, 64, 2, 6), (16, 24, 1, 1)]:
            P = R.Params(5)
            x = _x(1, cin, 20, 28, seed=6)
            ref = R.inverted_residual(P, x, "ir", cout, stride, t)
            _cmp(_load(M.InvertedResidual(cin, cout, stride, t), P, "ir")(x.cuda().to(dtype)), ref, tol)
        # _ASPP (output stride 16) and PyramidPooling
        P = R.Params(7)
        x = _x(2, 256, 17, 33, seed=8)
        ref = R.aspp(P, x, "aspp", 256, 16)
        _cmp(_load(M._ASPP(256, 256, output_stride=16), P, "aspp")(x.cuda().to(dtype)), ref, 2 * tol)
        P = R.Params(9)
        x = _x(1, 64, 17, 33, seed=10)
        ref = R.pyramid_pooling(P, x, "psp")
        _cmp(_load(M.PyramidPooling(64), P, "psp")(x.cuda().to(dtype)), ref, tol)


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-3), (torch.bfloat16, 1.2e-2)], ids=["f16", "bf16"])
def test_attention_modules(dtype, tol):
    """CAM_Module and CrissCrossAttention against the oracle (which is pinned to the reference classes / to a scalar
    transcription of ca_cuda.cu in tests/golden/make_golden.py)."""
    from segmentron_b200 import modules as M
    with torch.no_grad():
        for (n, c, h, w, s) in [(2, 64, 12, 20, 0.2), (1, 512, 17, 33, 0.05)]:
            P = R.Params(21)
            x = (_x(n, c, h, w, seed=22) * s).to(dtype).float()     # 16-bit-representable input: isolates the kernels' own error
            ref = R.cam(P, x, "cam", gamma=0.7)
            _cmp(_load(M.CAM_Module(c), P, "cam")(x.cuda().to(dtype)), ref, tol)
        for (n, c, h, w) in [(2, 512, 12, 20), (1, 512, 32, 40)]:       # N = 240 (ragged key tile) and 1280
            P = R.Params(25)
            x = (_x(n, c, h, w, seed=26) * 0.5).to(dtype).float()
            ref = R.pam(P, x, "pam", gamma=0.8)
            _cmp(_load(M.PAM_Module(c), P, "pam")(x.cuda().to(dtype)), ref, tol)
        for (n, c, h, w) in [(2, 64, 9, 13), (1, 512, 16, 24)]:
            P = R.Params(23)
            x = _x(n, c, h, w, seed=24)
            ref = R.criss_cross_attention(P, x, "cca", gamma=0.6)
            _cmp(_load(M.CrissCrossAttention(c), P, "cca")(x.cuda().to(dtype)), ref, tol)


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 4e-3), (torch.bfloat16, 2.5e-2)], ids=["f16", "bf16"])
def test_criss_cross_attention_backward(dtype, tol):
    """Training-mode CrissCrossAttention drop-in: forward and EVERY gradient (input, q/k/v conv weights and biases, gamma)
    against autograd through the oracle (ca_weight / ca_map are pinned to a scalar transcription of ca_cuda.cu) -- i.e. the
    replacement of _C.ca_backward / _C.ca_map_backward (csrc/criss_cross_attention/ca_cuda.cu:38-92,122-177)."""
    from segmentron_b200 import modules as M
    for (n, c, h, w) in [(2, 64, 9, 13), (1, 512, 16, 24)]:
        P = R.Params(23)
        x = _x(n, c, h, w, seed=24).to(dtype).float()
        R.criss_cross_attention(P, x, "cca", gamma=0.6)                 # creates the parameters
        names = [k for k in P.t if k.startswith("cca.")]
        for k in names:
            P.t[k] = P.t[k].to(dtype).float().detach().requires_grad_(True)     # 16-bit-representable parameters
        xr = x.clone().requires_grad_(True)
        ref = R.criss_cross_attention(P, xr, "cca", gamma=0.6)
        dy = _x(n, c, h, w, seed=25).to(dtype).float()
        ref.backward(dy)
        m = M.CrissCrossAttention(c)
        m.load_state_dict({k[4:]: v.detach() for k, v in P.t.items() if k.startswith("cca.")}, strict=True)
        m = m.cuda().train()
        xg = x.cuda().to(dtype).requires_grad_(True)
        y = m(xg)
        _cmp(y.detach(), ref.detach(), tol)
        y.backward(dy.cuda().to(dtype))
        _cmp(xg.grad, xr.grad, tol)
        # key_conv.bias has an analytically ZERO gradient (a per-query constant added to every energy cancels in the softmax):
        # errors are measured against |g_ref| + 5 % of the largest parameter-gradient norm
        floor = 0.05 * max(float(P.t[k].grad.norm()) for k in names)
        errs = {}
        for k in names:
            g = dict(m.named_parameters())[k[4:]].grad
            assert g is not None, k
            errs[k] = float((g.float().cpu().reshape(-1) - P.t[k].grad.reshape(-1)).norm() / (float(P.t[k].grad.norm()) + floor))
        print(f"[cca bwd {dtype} {(n, c, h, w)}] " + ", ".join(f"{k[4:]}: {v:.2e}" for k, v in errs.items()))
        assert max(errs.values()) < 2 * tol, errs


def _train_reference(oracle_fn, prefix, x, dtype, autocast):
    """fp32 (or bf16-autocast) forward + backward of the oracle's functional restatement in train mode"""
    P = R.Params(41)
    with torch.no_grad():
        oracle_fn(P, x)                                           # creates the parameters
    names = [k for k in P.t if k.startswith(prefix + ".") and not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
    for k in names:
        P.t[k] = P.t[k].to(dtype).float().detach().requires_grad_(True)
    sd0 = {k[len(prefix) + 1:]: v.detach().clone() for k, v in P.state_dict().items() if k.startswith(prefix + ".")}
    P.training = True
    xr = x.clone().requires_grad_(True)
    if autocast:
        with torch.autocast("cpu", dtype=torch.bfloat16):
            ref = oracle_fn(P, xr)
    else:
        ref = oracle_fn(P, xr)
    dy = _x(*ref.shape, seed=43).to(dtype).float()
    ref.float().backward(dy)
    return P, names, sd0, ref.detach().float(), xr.grad, dy


TRAIN_DROPIN_CASES = {
    # all verified on the B200 in round 1 (gpurun_out/final5_tests.log -> profiles/r1_final_dropin_train_tests.log: 11 passed)
    "sep_relu_first": (lambda M: M.SeparableConv2d(64, 128, 3, 1, 1, True), lambda P, t: R.separable_conv2d(P, t, "m", 128, 1, 1, True, 1e-5)),
    "sep_relu_first_s2": (lambda M: M.SeparableConv2d(64, 128, 3, 2, 1, True), lambda P, t: R.separable_conv2d(P, t, "m", 128, 2, 1, True, 1e-5)),
    "sep_d2": (lambda M: M.SeparableConv2d(64, 128, 3, 1, 2, False), lambda P, t: R.separable_conv2d(P, t, "m", 128, 1, 2, False, 1e-5)),
    # same Functions, other geometry
    "cbr_3x3_d2": (lambda M: M._ConvBNReLU(64, 128, 3, 1, 2, 2), lambda P, t: R.conv_bn_act(P, t, "m", 128, 3, 1, 2, 2)),
    "cbr6_3x3_s2": (lambda M: M._ConvBNReLU(64, 64, 3, 2, 1, 1, relu6=True), lambda P, t: R.conv_bn_act(P, t, "m", 64, 3, 2, 1, 1, act="relu6")),
    "cb_1x1_s2": (lambda M: M._ConvBN(64, 256, 1, 2), lambda P, t: R.conv_bn_act(P, t, "m", 256, 1, 2, act=None)),
    "dw_cbr": (lambda M: M._ConvBNReLU(64, 64, 3, 1, 1, 1, groups=64), lambda P, t: R.conv_bn_act(P, t, "m", 64, 3, 1, 1, 1, groups=64)),
    # composite classes, unit by unit through the same Functions (wiring also checked on the CPU: test_host_cpu.py)
    "inverted_residual_skip": (lambda M: M.InvertedResidual(64, 64, 1, 6), lambda P, t: R.inverted_residual(P, t, "m", 64, 1, 6)),
    "inverted_residual_s2": (lambda M: M.InvertedResidual(64, 96, 2, 6), lambda P, t: R.inverted_residual(P, t, "m", 96, 2, 6)),
    "inverted_residual_t1_d2": (lambda M: M.InvertedResidual(64, 32, 1, 1, dilation=2), lambda P, t: R.inverted_residual(P, t, "m", 32, 1, 1, 2)),
    "aspp": (lambda M: _no_dropout(M._ASPP(64, 64, output_stride=16)), lambda P, t: _aspp_no_dropout(P, t)),
}
TRAIN_DROPIN_CASES["pyramid_pooling"] = (lambda M: M.PyramidPooling(64), lambda P, t: R.pyramid_pooling(P, t, "m"))


def _no_dropout(m):
    m.dropout.p = 0.0
    return m


def _aspp_no_dropout(P, t):
    P.dropout_masks["m.dropout"] = torch.ones(1)
    return R.aspp(P, t, "m", 64, 16)


_ALL_TRAIN_CASES = TRAIN_DROPIN_CASES


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 4e-3), (torch.bfloat16, 2.5e-2)], ids=["f16", "bf16"])
def test_pam_module_backward(dtype, tol):
    """Training-mode PAM_Module drop-in (attention.PamFunction, csrc/softmax_rows.cu): output and EVERY gradient (input, q/k/v conv
    weights and biases, gamma) against autograd through the oracle's pam() (modules/module.py:112-131)."""
    from segmentron_b200 import modules as M
    for (n, c, h, w) in [(2, 512, 12, 20), (1, 512, 16, 24), (2, 64, 8, 12)]:
        P = R.Params(25)
        x = (_x(n, c, h, w, seed=26) * 0.5).to(dtype).float()
        with torch.no_grad():
            R.pam(P, x, "pam", gamma=0.8)
        names = [k for k in P.t if k.startswith("pam.")]
        for k in names:
            P.t[k] = P.t[k].to(dtype).float().detach().requires_grad_(True)
        xr = x.clone().requires_grad_(True)
        ref = R.pam(P, xr, "pam", gamma=0.8)
        dy = _x(n, c, h, w, seed=27).to(dtype).float()
        ref.backward(dy)
        m = M.PAM_Module(c)
        m.load_state_dict({k[4:]: v.detach() for k, v in P.t.items() if k.startswith("pam.")}, strict=True)
        m = m.cuda().train()
        xg = x.cuda().to(dtype).requires_grad_(True)
        y = m(xg)
        _cmp(y.detach(), ref.detach(), tol)
        y.backward(dy.cuda().to(dtype))
        _cmp(xg.grad, xr.grad, tol)
        floor = 0.05 * max(float(P.t[k].grad.norm()) for k in names)       # key_conv.bias: analytically zero gradient
        errs = {}
        for k in names:
            g = dict(m.named_parameters())[k[4:]].grad
            assert g is not None, k
            errs[k] = float((g.float().cpu().reshape(-1) - P.t[k].grad.reshape(-1)).norm() / (float(P.t[k].grad.norm()) + floor))
        print(f"[pam bwd {dtype} {(n, c, h, w)}] " + ", ".join(f"{k[4:]}: {v:.2e}" for k, v in errs.items()))
        assert max(errs.values()) < 2 * tol, errs


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 4e-3), (torch.bfloat16, 2.5e-2)], ids=["f16", "bf16"])
def test_cam_module_backward(dtype, tol):
    """Training-mode CAM_Module drop-in (attention.CamFunction, csrc/cam_bwd.cu): output, input gradient and gamma gradient against
    autograd through the oracle's cam() (modules/module.py:142-162)."""
    from segmentron_b200 import modules as M
    for (n, c, h, w, s) in [(2, 64, 12, 20, 0.2), (1, 512, 16, 24, 0.05)]:       # scales as in test_attention_modules (soft attention)
        P = R.Params(27)
        x = (_x(n, c, h, w, seed=28) * s).to(dtype).float()
        with torch.no_grad():
            R.cam(P, x, "cam", gamma=0.8)
        P.t["cam.gamma"] = P.t["cam.gamma"].to(dtype).float().detach().requires_grad_(True)
        xr = x.clone().requires_grad_(True)
        ref = R.cam(P, xr, "cam", gamma=0.8)
        dy = _x(n, c, h, w, seed=29).to(dtype).float()
        ref.backward(dy)
        m = M.CAM_Module(c)
        m.load_state_dict({"gamma": P.t["cam.gamma"].detach()}, strict=True)
        m = m.cuda().train()
        xg = x.cuda().to(dtype).requires_grad_(True)
        y = m(xg)
        _cmp(y.detach(), ref.detach(), tol)
        y.backward(dy.cuda().to(dtype))
        _cmp(xg.grad, xr.grad, tol)
        g_ref = float(P.t["cam.gamma"].grad)
        assert abs(float(m.gamma.grad) - g_ref) <= 2 * tol * abs(g_ref) + 1e-3, (float(m.gamma.grad), g_ref)


@pytest.mark.parametrize("case", list(_ALL_TRAIN_CASES))
def test_dropin_modules_train_backward(case):
    """Training-mode drop-ins (train-mode BatchNorm + backward kernels behind torch.autograd.Function) against autograd through
    the oracle in train mode: output, input gradient, every parameter gradient, BatchNorm running statistics.
    Yardstick: a ReLU whose pre-activation was rounded to bf16 flips its mask where |pre-activation| is below the rounding
    noise, so the REFERENCE'S OWN bf16-autocast gradients differ from fp32 by up to 4.2 % rel-L2 (sep_d2); each quantity must
    be within max(2.5e-2, 1.5x the reference's bf16-vs-fp32 error + 1e-2) (measured on the B200 for sep_d2's input gradient:
    4.184e-2, identical to the autocast error)."""
    from segmentron_b200 import modules as M
    dtype = torch.bfloat16
    make_mod, oracle_fn = _ALL_TRAIN_CASES[case]
    x = (_x(2, 64, 33, 41, seed=42)).to(dtype).float()
    P, names, sd0, ref, gx, dy = _train_reference(oracle_fn, "m", x, dtype, False)
    Pa, _, _, ref_a, gx_a, _ = _train_reference(oracle_fn, "m", x, dtype, True)

    def rel(a, b):
        return float((a.float().cpu() - b).norm() / (b.norm() + 1e-30))

    m = make_mod(M)
    m.load_state_dict(sd0, strict=True)
    m = m.cuda().train()
    xg = x.cuda().to(dtype).requires_grad_(True)
    y = m(xg)
    assert rel(y.detach(), ref) <= max(2.5e-2, 1.5 * rel(ref_a, ref) + 1e-2)
    y.backward(dy.cuda().to(dtype))
    assert rel(xg.grad, gx) <= max(2.5e-2, 1.5 * rel(gx_a, gx) + 1e-2), (rel(xg.grad, gx), rel(gx_a, gx))
    floor = 0.05 * max(float(P.t[k].grad.norm()) for k in names)          # analytically-zero gradients are rounding noise
    for k in names:
        g = dict(m.named_parameters())[k[2:]].grad
        assert g is not None, k
        ours = float((g.float().cpu() - P.t[k].grad).norm() / (float(P.t[k].grad.norm()) + floor))
        yard = float((Pa.t[k].grad - P.t[k].grad).norm() / (float(P.t[k].grad.norm()) + floor))
        assert ours <= max(5e-2, 1.5 * yard + 2e-2), (k, ours, yard)
    for k, v in m.state_dict().items():
        if k.endswith(("running_mean", "running_var")):
            assert torch.allclose(v.cpu(), P.t["m." + k], atol=2e-2, rtol=2e-2), k


def test_dropin_errors_and_cache_invalidation():
    from segmentron_b200 import modules as M
    m = M.SeparableConv2d(64, 64).cuda().eval()
    x = _x(1, 64, 9, 9).cuda().to(torch.bfloat16)
    with torch.no_grad():
        y0 = m(x).clone()
        m.block.bn_point.eps = 0.5                             # eps change must invalidate the folded-BN cache
        y1 = m(x).clone()
        assert not torch.equal(y0, y1)
        m.block.pointwise.weight.mul_(2.0)                     # in-place parameter update bumps the version
        y2 = m(x)
        assert not torch.equal(y1, y2)
    with pytest.raises(RuntimeError):
        m(x.cpu())
    with pytest.raises(RuntimeError):                             # training mode has no CPU path either
        m.train()(x.cpu())
    b = M._ASPP(64, 64, output_stride=16).cuda().train()          # batch statistics over ONE value (1x1 image pooling, batch 1):
    with pytest.raises(ValueError):                                # the same error torch's batch_norm raises
        b(x)

"""Training-mode CAM_Module (attention.CamFunction) without a GPU: the Function runs on CPU fp32 tensors against a host emulation of
the C ABI (tests/fake_lib.py: verified kernels by their documented semantics, the two new CAM backward kernels by transcription) and
must reproduce autograd through the oracle's cam() (modules/module.py:142-162): output, input gradient, gamma gradient.
Pins the backward's math (G = dy^T x, dE = -gamma A (G - sum A G), W1 = gamma A^T, W2 = dE + dE^T, dx = dy + dy.W1 + x.W2), the
pitches / K paddings handed to the GEMMs and the pointer plumbing; the CUDA build is covered by the gated GPU test."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from fake_lib import FakeLib  # noqa: E402
from oracle import segref as R  # noqa: E402


@pytest.mark.parametrize("shape", [(2, 16, 5, 7), (1, 72, 9, 13), (3, 8, 4, 4)])
def test_cam_function_matches_oracle_autograd(shape, monkeypatch):
    from segmentron_b200 import attention as A, lib as L, ops
    monkeypatch.setattr(ops, "_PLAN_DRY_RUN", True)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    monkeypatch.setattr(L, "load", lambda: FakeLib())
    n, c, h, w = shape
    g = torch.Generator().manual_seed(c)
    x = torch.randn(n, c, h, w, generator=g) * 0.5
    dy = torch.randn(n, c, h, w, generator=g)
    P = R.Params(3)
    xr = x.clone().requires_grad_(True)
    with torch.no_grad():
        R.cam(P, x, "m")
    P.t["m.gamma"] = P.t["m.gamma"].detach().clone().requires_grad_(True)
    ref = R.cam(P, xr, "m")
    ref.backward(dy)
    xh = x.permute(0, 2, 3, 1).contiguous().requires_grad_(True)
    gamma = P.t["m.gamma"].detach().clone().requires_grad_(True)
    y = A.CamFunction.apply(xh, gamma)
    y.backward(dy.permute(0, 2, 3, 1).contiguous())
    scale = float(ref.abs().max())
    assert float((y.permute(0, 3, 1, 2) - ref).abs().max()) <= 1e-5 * scale
    assert float((xh.grad.permute(0, 3, 1, 2) - xr.grad).abs().max()) <= 2e-5 * float(xr.grad.abs().max())
    assert abs(float(gamma.grad) - float(P.t["m.gamma"].grad)) <= 1e-4 * abs(float(P.t["m.gamma"].grad)) + 1e-6


@pytest.mark.parametrize("shape", [(2, 64, 4, 6), (1, 128, 8, 5)])
def test_pam_function_matches_oracle_autograd(shape, monkeypatch):
    """Training-mode PAM_Module (attention.PamFunction: materialised attention, csrc/softmax_rows.cu) against autograd through the
    oracle's pam() (modules/module.py:112-131): output, input gradient, the three conv weights and biases, gamma.  h*w % 8 == 0,
    query depth 8 / 16 (< the GEMM K block: exercises the padded q / k pitches)."""
    from segmentron_b200 import attention as A, lib as L, ops
    monkeypatch.setattr(ops, "_PLAN_DRY_RUN", True)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    monkeypatch.setattr(L, "load", lambda: FakeLib())
    n, c, h, w = shape
    g = torch.Generator().manual_seed(c + 1)
    x = torch.randn(n, c, h, w, generator=g) * 0.5
    dy = torch.randn(n, c, h, w, generator=g)
    P = R.Params(5)
    with torch.no_grad():
        R.pam(P, x, "m", gamma=0.8)
    names = [k for k in P.t if k.startswith("m.")]
    for k in names:
        P.t[k] = P.t[k].detach().clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = R.pam(P, xr, "m", gamma=0.8)
    ref.backward(dy)
    xh = x.permute(0, 2, 3, 1).contiguous().requires_grad_(True)
    prm = {k: P.t[k].detach().clone().requires_grad_(True) for k in names}
    y = A.PamFunction.apply(xh, prm["m.query_conv.weight"], prm["m.query_conv.bias"], prm["m.key_conv.weight"], prm["m.key_conv.bias"],
                            prm["m.value_conv.weight"], prm["m.value_conv.bias"], prm["m.gamma"])
    y.backward(dy.permute(0, 2, 3, 1).contiguous())
    assert float((y.permute(0, 3, 1, 2) - ref).abs().max()) <= 1e-5 * float(ref.detach().abs().max())
    assert float((xh.grad.permute(0, 3, 1, 2) - xr.grad).abs().max()) <= 5e-5 * float(xr.grad.abs().max())
    # key_conv.bias has an analytically ZERO gradient (a per-query constant added to every energy cancels in the softmax): errors
    # are measured against |g_ref| + 1e-4 of the largest parameter-gradient magnitude
    floor = 1e-4 * max(float(P.t[k].grad.abs().max()) for k in names)
    for k in names:
        gr, go = P.t[k].grad, prm[k].grad
        assert go is not None, k
        assert float((go.reshape(-1) - gr.reshape(-1)).abs().max()) <= 5e-5 * float(gr.abs().max()) + 1e-2 * floor, k


@pytest.mark.parametrize("hw", [(12, 18), (7, 11)])
def test_pyramid_pooling_training_matches_oracle_autograd(hw, monkeypatch):
    """Training-mode PyramidPooling (modules/module.py:82-97): adaptive pool (forward kernel by its semantics, the new backward
    kernel by transcription -- overlapping bins when the size is not a multiple of s), 1x1 conv + BN + ReLU units (substituted
    by their torch definition here; they are GPU kernels tested with -m gpu), bilinear up-sampling and its gather backward, concat
    order [x, 1, 2, 3, 6] -- against autograd through the oracle."""
    import torch.nn.functional as F
    from segmentron_b200 import lib as L, modules as M, ops

    def unit(xh, conv, bn, act, pre_relu=False):
        t = F.conv2d(xh.permute(0, 3, 1, 2), conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
        t = F.batch_norm(t, bn.running_mean, bn.running_var, bn.weight, bn.bias, True, bn.momentum, bn.eps)
        return (F.relu(t) if act == "relu" else t).permute(0, 2, 3, 1).contiguous()
    monkeypatch.setattr(ops, "_PLAN_DRY_RUN", True)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    monkeypatch.setattr(L, "load", lambda: FakeLib())
    monkeypatch.setattr(M, "_train_conv_bn_act", unit)
    monkeypatch.setattr(M, "_train_enter", lambda x, m: x.permute(0, 2, 3, 1).contiguous())
    h, w = hw
    g = torch.Generator().manual_seed(h)
    x = torch.randn(2, 32, h, w, generator=g)
    P = R.Params(9)
    with torch.no_grad():
        R.pyramid_pooling(P, x, "m")
    keys = [k for k in P.t if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
    for k in keys:
        P.t[k] = P.t[k].detach().requires_grad_(True)
    sd0 = {k[2:]: v.detach().clone() for k, v in P.state_dict().items()}
    P.training = True
    xr = x.clone().requires_grad_(True)
    ref = R.pyramid_pooling(P, xr, "m")
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy)
    m = M.PyramidPooling(32)
    m.load_state_dict(sd0, strict=True)
    m.train()
    xg = x.clone().requires_grad_(True)
    y = m(xg)
    y.backward(dy)
    assert torch.allclose(y, ref, atol=1e-5, rtol=1e-5)
    assert torch.allclose(xg.grad, xr.grad, atol=1e-5, rtol=1e-4)
    for k in keys:
        gm = dict(m.named_parameters())[k[2:]].grad
        assert gm is not None and torch.allclose(gm, P.t[k].grad, atol=1e-4, rtol=1e-4), k

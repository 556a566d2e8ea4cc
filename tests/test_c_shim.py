"""`segmentron._C` drop-in (segmentron_b200/c_shim.py over csrc/ca_nchw.cu): the four functions of the reference's pybind module
(segmentron/modules/csrc/vision.cpp:6-11, ca.h:25-72).

CPU part: the module object exposes exactly the four names, CPU tensors raise the reference's "Not implemented on the CPU".
GPU part (through the C ABI):
  * each function against the oracle's restatement (oracle/segref.py ca_weight / ca_map, pinned to a scalar transcription of
    ca_cuda.cu by tests/golden) and, for the two backward functions, against autograd through that restatement -- fp32 to 1e-5 of
    the output's max, fp16 / bf16 on 16-bit-representable inputs to 2^-10 / 2^-7;
  * the REFERENCE'S OWN code on top of it: `segmentron.modules.cc_attention` (its autograd Functions _CAWeight / _CAMap and its
    CrissCrossAttention module, cc_attention.py:11-72) imported from the staged, unmodified reference (baseline/_ref) with the shim
    installed as `segmentron._C` -- forward and every gradient against the oracle;
  * the reference's CCNet built through its own registry (`get_segmentation_model()`, models/ccnet.py) on the shim, against the
    committed fixture tests/golden/ccnet_resnet101_65x97.pt.
"""
import os
import sys

import pytest
import torch

from oracle import segref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
G = os.path.join(os.path.dirname(__file__), "golden")


def test_shim_surface_cpu():
    from segmentron_b200 import c_shim
    m = c_shim.make_module()
    assert sorted(n for n in dir(m) if n.startswith("ca_")) == ["ca_backward", "ca_forward", "ca_map_backward", "ca_map_forward"]
    t = torch.zeros(1, 8, 4, 5)
    for call in (lambda: m.ca_forward(t, t), lambda: m.ca_backward(torch.zeros(1, 8, 4, 5), t, t),
                 lambda: m.ca_map_forward(torch.zeros(1, 8, 4, 5), t), lambda: m.ca_map_backward(t, torch.zeros(1, 8, 4, 5), t)):
        with pytest.raises(RuntimeError, match="Not implemented on the CPU"):       # ca.h:34,46,58,70
            call()


def _rt(shape, seed, dtype):
    t = torch.randn(*shape, generator=torch.Generator().manual_seed(seed))
    return t.to(dtype).float()                          # representable in the kernel dtype


def _close(a, b, tol, what, floor=0.0):
    err = float((a.float().cpu() - b).abs().max() / max(float(b.abs().max()), floor, 1e-30))
    assert err <= tol, (what, err, tol)
    return err


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.float16, 2.0 ** -10), (torch.bfloat16, 2.0 ** -7)],
                         ids=["f32", "f16", "bf16"])
@pytest.mark.parametrize("shape", [(2, 8, 5, 7), (1, 64, 33, 49), (2, 16, 40, 129), (1, 3, 1, 37), (1, 5, 33, 1)],
                         ids=lambda s: "x".join(map(str, s)))
def test_c_functions_vs_oracle(shape, dtype, tol):
    from segmentron_b200 import c_shim as S
    n, c, h, w = shape
    t, f = _rt(shape, 1, dtype), _rt(shape, 2, dtype)
    g = _rt((n, 2 * c, h, w), 3, dtype)
    a = torch.softmax(_rt((n, h + w - 1, h, w), 4, torch.float32), 1).to(dtype).float()
    dwt = _rt((n, h + w - 1, h, w), 5, dtype)
    dout = _rt((n, 2 * c, h, w), 6, dtype)
    dev = lambda v: v.to("cuda", dtype)                     # noqa: E731
    # ca_forward / ca_map_forward
    _close(S.ca_forward(dev(t), dev(f)), R.ca_weight(t, f), tol, "ca_forward")
    _close(S.ca_map_forward(dev(a), dev(g)), R.ca_map(a, g), tol, "ca_map_forward")
    # ca_backward == autograd of ca_weight;  ca_map_backward == autograd of ca_map
    tt, ff = t.clone().requires_grad_(), f.clone().requires_grad_()
    R.ca_weight(tt, ff).backward(dwt)
    dt_, df_ = S.ca_backward(dev(dwt), dev(t), dev(f))
    _close(dt_, tt.grad, tol, "ca_backward dt")
    _close(df_, ff.grad, tol, "ca_backward df")
    aa, gg = a.clone().requires_grad_(), g.clone().requires_grad_()
    R.ca_map(aa, gg).backward(dout)
    dw_, dg_ = S.ca_map_backward(dev(dout), dev(a), dev(g))
    _close(dw_, aa.grad, tol, "ca_map_backward dw")
    _close(dg_, gg.grad, tol, "ca_map_backward dg")
    # non-contiguous inputs are accepted like the reference's (.contiguous() inside, ca_cuda.cu:205-207)
    tn = dev(t).permute(0, 1, 3, 2).contiguous().permute(0, 1, 3, 2)
    assert not tn.is_contiguous() or h == 1 or w == 1
    _close(S.ca_forward(tn, dev(f)), R.ca_weight(t, f), tol, "ca_forward (strided input)")


def _reference_cc():
    import ref_harness as H
    if not H.available():
        pytest.fail("baseline/_ref is not staged: run `python tools/make_baseline_ref.py` in the build container")
    H.enter()
    from segmentron_b200 import c_shim
    c_shim.install()
    import segmentron.modules.cc_attention as cc
    assert cc._C.__doc__.startswith("segb200")
    return cc


@pytest.mark.gpu
def test_reference_autograd_functions_on_the_shim():
    """The reference's own _CAWeight / _CAMap / CrissCrossAttention (cc_attention.py:11-72), unmodified, on `segmentron._C` = shim."""
    cc = _reference_cc()
    assert cc.__file__.startswith(os.path.join(ROOT, "baseline", "_ref"))
    n, c, h, w = 2, 64, 17, 23
    P = R.Params(7)
    x = torch.randn(n, c, h, w, generator=torch.Generator().manual_seed(8))
    dy = torch.randn(n, c, h, w, generator=torch.Generator().manual_seed(9))
    xo = x.clone().requires_grad_()
    yo = R.criss_cross_attention(P, xo, "cca", gamma=0.7)                  # fp32 oracle on the CPU (creates the parameters)
    for k in P.t:
        P.t[k].requires_grad_()
    xo.grad = None
    yo = R.criss_cross_attention(P, xo, "cca", gamma=0.7)
    yo.backward(dy)
    m = cc.CrissCrossAttention(c).cuda()
    sd = {k[len("cca."):]: v.detach() for k, v in P.t.items() if k.startswith("cca.")}
    m.load_state_dict(sd, strict=True)
    xr = x.cuda().requires_grad_()
    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False      # the module's own 1x1 convs run in cuDNN
    try:
        yr = m(xr)
        yr.backward(dy.cuda())
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf
    # fp32 on both sides, but cuDNN / cuBLAS and the CPU sum in different orders and the softmax sits on energies of O(10):
    # forward to 2e-5, gradients to 1e-3 of their max
    errs = {"y": _close(yr.detach(), yo.detach(), 2e-5, "CrissCrossAttention forward"), "dx": _close(xr.grad, xo.grad, 1e-3, "dx")}
    wscale = float(P.t["cca.value_conv.weight"].grad.abs().max())
    for name, p in m.named_parameters():
        # key_conv.bias: adding a constant to every key shifts all energies of a query equally -> the softmax, hence the loss, does not
        # depend on it; its gradient is pure rounding noise (1e-5) on both sides, so it is measured against the weight-gradient scale
        errs[name] = _close(p.grad, P.t["cca." + name].grad.reshape(p.shape), 1e-3, "d" + name,
                            floor=1e-3 * wscale if name == "key_conv.bias" else 0.0)
    print("\n[reference CrissCrossAttention on the segb200 _C shim] max-rel errors vs the CPU oracle:", {k: f"{v:.1e}" for k, v in errs.items()})


@pytest.mark.gpu
def test_reference_ccnet_builds_and_runs_on_the_shim():
    """`get_segmentation_model()` for configs/cityscapes_ccnet_resnet.yaml (models/ccnet.py, commented out of models/__init__.py:11
    because the extension cannot be built) works once the shim is installed; output == the committed reference fixture."""
    import ref_harness as H
    _reference_cc()
    fx = torch.load(os.path.join(G, "ccnet_resnet101_65x97.pt"))
    model = H.build_model("cityscapes_ccnet_resnet.yaml", install_c=True)
    P = R.build_params(fx["model"], fx["seed"])
    model.load_state_dict(P.state_dict(), strict=True)
    x = torch.randn(*fx["shape"], generator=torch.Generator().manual_seed(fx["input_seed"]))
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            y = model.cuda()(x.cuda())[0].float().cpu()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b
    err = float((y - fx["y_ref"].float()).norm() / fx["y_ref"].float().norm())
    print(f"\n[reference CCNet on the segb200 _C shim] rel-L2 vs committed reference fixture: {err:.2e}")
    assert err < 1e-4, err

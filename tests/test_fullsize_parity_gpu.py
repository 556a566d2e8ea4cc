"""Numeric parity AT THE BASELINE.json SIZES (one image per config; the engines are batch-invariant, tests/test_model_gpu.py):

  C2  DeepLabv3+/Xception65   bf16  1x3x1025x2049
  C3  DeepLabv3+/ResNet101    bf16  1x3x1025x2049   (forward of the training config)
  C4  DANet/ResNet101         bf16  1x3x1024x2048   (PAM over N = 128*256 = 32 768 tokens; the fp32 oracle materialises N x N on the GPU)
  C5  HRNet-w18-small-v1      fp16  1x3x1024x2048

For each: this engine's logits, and the reference's own 16-bit forward (the same torch ops through cuDNN/cuBLAS, `.to(dtype)` --
the real reference model from baseline/_ref for C2, the oracle port, which is pinned to it, for the others), both against the fp32
oracle run on the same GPU with TF32 off.  Criterion (the north-star's "match the reference's own forward"): our rel-L2 error must not
exceed the reference's own 16-bit error by more than 5 %, and no argmax mismatch may sit at a pixel whose fp32 top-2 margin exceeds
4x the 16-bit resolution of the logits.  The numbers are printed (pytest -s) and recorded in BASELINE.md.
"""
import os
import sys

import pytest
import torch

from oracle import segref as R

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def _engine(model, P, dtype):
    from segmentron_b200.engine import DANetB200, DeepLabV3PlusB200, HRNetB200
    if model == "hrnet_w18_small_v1":
        return HRNetB200(P.state_dict(), dtype=dtype, cuda_graph=False, want_argmax=True)
    if model == "danet_resnet101":
        return DANetB200(P.state_dict(), dtype=dtype, cuda_graph=False, want_argmax=True)
    cfg = R.MODELS[model]
    return DeepLabV3PlusB200(P.state_dict(), backbone=cfg["backbone"], eps_encoder=cfg["eps_encoder"], use_aspp=cfg["use_aspp"],
                             use_decoder=cfg["use_decoder"], dtype=dtype, cuda_graph=False, want_argmax=True)


def _reference_16bit(model, P, x, dtype):
    """the reference's own 16-bit forward on this GPU: the staged reference package for the headline model, else the oracle port"""
    import ref_harness as H
    if model == "deeplabv3plus_xception65" and H.available():
        m = H.build_model("cityscapes_deeplabv3_plus.yaml")
        m.load_state_dict(P.state_dict(), strict=True)
        m = m.cuda().to(dtype)
        with torch.no_grad():
            return m(x.to(dtype))[0].float(), "reference package (baseline/_ref)"
    return R.forward(model, P.to("cuda", dtype), x.to(dtype)).float(), "oracle port"


CASES = [("C2", "deeplabv3plus_xception65", torch.bfloat16, (1, 3, 1025, 2049), 31),
         ("C3", "deeplabv3plus_resnet101", torch.bfloat16, (1, 3, 1025, 2049), 32),
         ("C4", "danet_resnet101", torch.bfloat16, (1, 3, 1024, 2048), 33),
         ("C5", "hrnet_w18_small_v1", torch.float16, (1, 3, 1024, 2048), 34)]


@pytest.mark.parametrize("tag,model,dtype,shape,seed", CASES, ids=[c[0] for c in CASES])
def test_parity_at_baseline_size(tag, model, dtype, shape, seed):
    P = R.build_params(model, seed)
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(100 + seed)).cuda()
    with torch.no_grad():
        y32 = R.forward(model, P.to("cuda"), x).float()
        torch.cuda.empty_cache()
        y16, src = _reference_16bit(model, P, x, dtype)
        torch.cuda.empty_cache()
    eng = _engine(model, P, dtype)
    y = eng(x).float()
    am = eng.argmax(x).long()
    assert torch.isfinite(y).all()
    e_ours, e_ref, e_x = _rel(y, y32), _rel(y16, y32), _rel(y, y16)
    a32 = y32.argmax(1)
    top2 = y32.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    mism, mism_ref = am != a32, y16.argmax(1) != a32
    res = float(y32.abs().max()) * (2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10)
    hard = int((mism & (margin > 4 * res)).sum())
    hard_ref = int((mism_ref & (margin > 4 * res)).sum())
    print(f"\n[{tag} {model} {str(dtype).replace('torch.', '')} {shape[2]}x{shape[3]}] rel-L2 vs fp32: ours {e_ours:.3e}, reference-16bit "
          f"({src}) {e_ref:.3e}; ours vs reference-16bit {e_x:.3e}; argmax mismatches vs fp32: ours {int(mism.sum())} / ref16 "
          f"{int(mism_ref.sum())} of {mism.numel()} (beyond 16-bit resolution: ours {hard} / ref16 {hard_ref})")
    assert torch.equal(am, y.argmax(1)), "fused argmax disagrees with argmax of the engine's own logits"
    if e_ref == e_ref:                                        # the reference's own 16-bit forward is finite
        assert e_ours <= 1.05 * e_ref + 1e-4, (e_ours, e_ref)
    else:
        assert e_ours < (1e-2 if dtype == torch.float16 else 6e-2)
    assert hard <= max(hard_ref, 0), (hard, hard_ref)

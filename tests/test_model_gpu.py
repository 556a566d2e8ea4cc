"""Whole-model parity on the B200: the engine (CUDA kernels through the C ABI) against
(a) the committed outputs of the real reference (tests/golden) and (b) the oracle run in fp32.

Criteria (SURVEY.md section 7 "parity at bf16"): rel-L2 of the logits against
the fp32 reference must not exceed the reference's OWN 16-bit error (same torch ops, `.to(dtype)` model,
measured in the same test) by more than 5 %, with absolute sanity caps 1e-2 (fp16) / 6e-2 (bf16) -- the reference's own
fp16 forward is at 1.9e-3..4.2e-3 and its bf16 forward at 1.7e-2 on these fixtures, so the north-star "1e-3" is a
per-kernel bound (tests/test_kernels_gpu.py), not a whole-model one; argmax
maps must agree with the fp32 reference wherever the fp32 top-2 margin exceeds the 16-bit resolution,
and the total mismatch count is reported.
"""
import os

import pytest
import torch

from oracle import segref as R

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def _engine(model, P, dtype, **kw):
    from segmentron_b200.engine import CCNetB200, DANetB200, DeepLabV3PlusB200, HRNetB200, OCNetB200, PSPNetB200
    if model == "hrnet_w18_small_v1":
        return HRNetB200(P.state_dict(), dtype=dtype, **kw)
    if model == "danet_resnet101":
        return DANetB200(P.state_dict(), dtype=dtype, **kw)
    if model == "ccnet_resnet101":
        return CCNetB200(P.state_dict(), dtype=dtype, **kw)
    if model == "pspnet_resnet101":
        return PSPNetB200(P.state_dict(), dtype=dtype, **kw)
    if model == "ocnet_resnet50":
        return OCNetB200(P.state_dict(), dtype=dtype, **kw)
    cfg = R.MODELS[model]
    return DeepLabV3PlusB200(P.state_dict(), backbone=cfg["backbone"], eps_encoder=cfg["eps_encoder"],
                             use_aspp=cfg["use_aspp"], use_decoder=cfg["use_decoder"], dtype=dtype, **kw)


def _check(model, P, x, y32, dtype, tol):
    eng = _engine(model, P, dtype, cuda_graph=False, want_argmax=True)
    y = eng(x.cuda()).float().cpu()
    am = eng.argmax(x.cuda()).long().cpu()
    # the reference's own 16-bit forward on the same device (same torch ops as the oracle)
    y16 = R.forward(model, P.to("cuda", dtype), x.cuda().to(dtype)).float().cpu()
    e_ours, e_ref = _rel(y, y32), _rel(y16, y32)
    top2 = y32.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    mism = am != y32.argmax(1)
    res = float(y32.abs().max()) * (2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10)
    hard = int((mism & (margin > 4 * res)).sum())
    mism16 = y16.argmax(1) != y32.argmax(1)
    hard16 = int((mism16 & (margin > 4 * res)).sum())         # the reference's own 16-bit forward, same criterion
    print(f"\n[{model} {dtype}] rel-L2 ours={e_ours:.3e} ref16={e_ref:.3e}; argmax mismatches {int(mism.sum())}/"
          f"{mism.numel()} (beyond-resolution: {hard}); ref16 mismatches {int(mism16.sum())} (beyond-resolution: {hard16})")
    assert torch.equal(am, y.argmax(1)), "fused argmax disagrees with argmax of the engine's own logits"
    # DANet: CAM's softmax(rowmax(E) - E) runs on UNNORMALISED Gram energies (|E| ~ 1e2..1e3 with these synthetic weights), which
    # amplifies any 16-bit rounding upstream of it by |E|: on ONE fixture the engine and the reference's own 16-bit forward are two
    # noisy samples of the same error distribution (tools/danet_diag.py: per-seed ratios 0.5 .. 1.5).  The comparison against the
    # reference's 16-bit forward is therefore made on the MEAN over seeds (test_danet_error_vs_reference_16bit_over_seeds below);
    # here DANet only has to stay under the absolute cap (fp16: 2e-2 -- the reference's own fp16 forward is NaN on this fixture).
    ill = model == "danet_resnet101"
    assert e_ours < (2e-2 if (ill and dtype == torch.float16) else tol), (e_ours, tol)            # absolute sanity cap
    if e_ref == e_ref and not ill:                                                                  # reference 16-bit forward finite
        assert e_ours < 1.05 * e_ref + 1e-4, (e_ours, e_ref)                                       # vs the reference's own 16-bit forward
    if e_ref != e_ref:
        # The reference's OWN 16-bit forward overflows to NaN on this fixture (CCNet / DANet in fp16: activations beyond 65504), i.e.
        # the network is outside the dtype's range for the reference itself.  The engine stays finite (fp32 epilogues), but in that
        # regime its output moves by ~5e-3 when a handful of folded-BN scales change by ONE fp32 ulp (host-side vs device-side
        # folding, tools/fold_diff.py): there is no 16-bit reference to be faithful to, so only the absolute cap above applies.
        return
    # beyond-resolution flips: none, unless the reference's own 16-bit forward has them too (then: no more than it has)
    assert hard <= max(hard16, mism.numel() // 500 if ill else 0), (hard, hard16)


def test_danet_error_vs_reference_16bit_over_seeds():
    """DANet (PAM + CAM head): mean rel-L2 error against the fp32 oracle over 6 seeded (weights, input) pairs, bf16, 64x96 and
    128x192 -- the engine's mean error must not exceed the mean error of the reference's own bf16 forward (same torch ops, oracle
    port in bf16) by more than 5 %.  Measured on B200: 2.3e-2 vs 2.8e-2 (64x96), 2.4e-2 vs 2.7e-2 (128x192)."""
    from segmentron_b200.engine import DANetB200
    model, dtype = "danet_resnet101", torch.bfloat16
    for shape in ((1, 3, 64, 96), (1, 3, 128, 192)):
        ours, ref = [], []
        for seed in range(6):
            P = R.build_params(model, 100 + seed)
            x = torch.randn(*shape, generator=torch.Generator().manual_seed(200 + seed)).cuda()
            with torch.no_grad():
                y32 = R.forward(model, P.to("cuda"), x).float()
                y16 = R.forward(model, P.to("cuda", dtype), x.to(dtype)).float()
            y = DANetB200(P.state_dict(), dtype=dtype)(x).float()
            ours.append(_rel(y, y32)); ref.append(_rel(y16, y32))
        mo, mr = sum(ours) / len(ours), sum(ref) / len(ref)
        print(f"\n[danet bf16 {shape[2]}x{shape[3]}] mean rel-L2 over 6 seeds: ours {mo:.3e}, reference-bf16 {mr:.3e}  (per seed ours "
              f"{[round(v, 4) for v in ours]}, ref {[round(v, 4) for v in ref]})")
        assert mo <= 1.05 * mr, (mo, mr)


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 1e-2), (torch.bfloat16, 6e-2)], ids=["f16", "bf16"])
@pytest.mark.parametrize("case", ["dlv3p_xception65_65x129", "dlv3p_xception65_97x161_b2", "dlv3p_mobilenetv2_64x128",
                                  "dlv3p_resnet101_65x129", "danet_resnet101_64x96", "ccnet_resnet101_65x97",
                                  "hrnet_w18s_128x192", "pspnet_resnet101_65x97", "ocnet_resnet50_65x97"])
def test_engine_vs_reference_fixture(case, dtype, tol):
    fx = torch.load(os.path.join(G, case + ".pt"))
    P = R.build_params(fx["model"], fx["seed"])
    x = torch.randn(*fx["shape"], generator=torch.Generator().manual_seed(fx["input_seed"]))
    _check(fx["model"], P, x, fx["y_ref"], dtype, tol)


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 1e-2), (torch.bfloat16, 6e-2)], ids=["f16", "bf16"])
def test_engine_vs_oracle_larger(dtype, tol):
    """Odd Cityscapes-like aspect (257x513, batch 2), CUDA-graph replay path, checked against the fp32 oracle."""
    model = "deeplabv3plus_xception65"
    P = R.build_params(model, 3)
    x = torch.randn(2, 3, 257, 513, generator=torch.Generator().manual_seed(5))
    y32 = R.forward(model, P.to("cuda"), x.cuda()).cpu()
    _check(model, P, x, y32, dtype, tol)
    eng = _engine(model, P, dtype, cuda_graph=True)
    y_a = eng(x.cuda()).clone()
    y_b = eng(x.cuda()).clone()
    assert torch.equal(y_a, y_b), "graph replay is not deterministic"
    eng2 = _engine(model, P, dtype, cuda_graph=False)
    assert torch.equal(eng2(x.cuda()), y_a), "graph replay differs from direct launch"


def test_full_size_batch_invariance_and_determinism():
    """Size-independent properties at the BASELINE resolution (1025x2049): every image of a batch is computed exactly as
    it is alone (tiles of the flattened-pixel GEMMs cross image boundaries, the per-element arithmetic must not), replays
    are bit-reproducible, outputs are finite, and the fused argmax equals torch.argmax of the logits."""
    model = "deeplabv3plus_xception65"
    P = R.build_params(model, 11)
    x = torch.randn(2, 3, 1025, 2049, generator=torch.Generator().manual_seed(12)).cuda()
    eng = _engine(model, P, torch.bfloat16, cuda_graph=True, want_argmax=True)
    y2 = eng(x).clone()
    am2 = eng.argmax(x).clone()
    assert torch.isfinite(y2.float()).all()
    assert torch.equal(eng(x), y2)
    assert torch.equal(am2.long(), y2.float().argmax(1))
    y1 = eng(x[1:2].contiguous()).clone()
    assert torch.equal(y1[0], y2[1]), float((y1[0].float() - y2[1].float()).abs().max())


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 1e-2), (torch.bfloat16, 6e-2)], ids=["f16", "bf16"])
def test_ocnet_larger_map(dtype, tol):
    """OCNet (base object-context head, models/ocnet.py) on a 257x385 input: 17 x 25 = 425 tokens, i.e. several 128-query tiles with
    a ragged last one and a ragged last 64-key tile in the depth-256 attention kernel; against the fp32 oracle and the reference's
    own 16-bit forward."""
    model = "ocnet_resnet50"
    P = R.build_params(model, 41)
    x = torch.randn(2, 3, 257, 385, generator=torch.Generator().manual_seed(42))
    y32 = R.forward(model, P.to("cuda"), x.cuda()).cpu()
    _check(model, P, x, y32, dtype, tol)

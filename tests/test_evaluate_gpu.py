"""Multi-scale + flip evaluation driver (segmentron_b200/evaluate.py, csrc/evaluate.cu) on the B200 against the oracle
(oracle/evalref.py == the reference's SegBaseModel.evaluate, tests/golden/evaluate_cases.pt) with the same seeded stub model as
`forward`, and against the committed reference scores themselves.  fp32: |err| <= 2e-5 * max|ref| (the stub's conv runs in
cuDNN fp32 vs the CPU's, TF32 off); bf16 logits: 2^-6 (three roundings to bf16 on both sides)."""
import importlib.util
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu]

G = os.path.join(os.path.dirname(__file__), "golden")


def _gen():
    spec = importlib.util.spec_from_file_location("make_eval_golden", os.path.join(G, "make_eval_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    return gen


@pytest.fixture(autouse=True)
def _no_tf32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


@pytest.mark.parametrize("case", range(6))
def test_scores_match_reference_fixture(case):
    from segmentron_b200.evaluate import evaluate
    gen = _gen()
    fx = torch.load(os.path.join(G, "evaluate_cases.pt"))[case]
    seed, b, h, w, scales, flip, crop = gen.CASES[case]
    stub = _cuda_stub(gen, seed)
    calls = []

    def forward(x):                                              # the stub's weights on the GPU; returns the models' tuple
        calls.append(tuple(x.shape))
        return (stub(x),)
    with torch.no_grad():
        got = evaluate(forward, gen.make_image(seed, b, h, w).cuda(), scales, flip, crop)
    ref = fx["scores"]
    assert got.shape == ref.shape and got.dtype == torch.float32
    assert float((got.cpu() - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    assert len(calls) == len(scales) and all(s[0] == (2 if flip else 1) * b for s in calls)      # ONE model call per scale


def _cuda_stub(gen, seed, dtype=torch.float32):
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(1000 + seed)
    w1 = (torch.randn(8, 3, 3, 3, generator=g) * 0.3).cuda()
    w2 = (torch.randn(5, 8, 1, 1, generator=g) * 0.5).cuda()
    b2 = torch.randn(5, generator=g).cuda()

    def forward(x):
        y = F.relu(F.conv2d(x, w1, None, stride=4, padding=1))
        y = F.conv2d(y, w2, b2)
        return F.interpolate(y, x.shape[2:], mode="bilinear", align_corners=True).to(dtype)
    return forward


def test_bf16_logits_round_like_the_reference():
    """model dtype bf16: `outputs +=`, the resized score and `scores +=` each round to bf16 in the reference; same here."""
    from oracle import evalref as E
    from segmentron_b200.evaluate import evaluate
    gen = _gen()
    seed, b, h, w, scales, flip, crop = 2, 1, 40, 72, [0.5, 1.0, 1.75], True, None
    img = gen.make_image(seed, b, h, w)
    cpu = gen.stub_forward(seed)
    with torch.no_grad():
        ref = E.evaluate(lambda x: cpu(x).to(torch.bfloat16), img, scales, flip, crop).float()
        got = evaluate(_cuda_stub(gen, seed, torch.bfloat16), img.cuda(), scales, flip, crop)
    assert got.dtype == torch.bfloat16
    assert float((got.float().cpu() - ref).abs().max()) <= 2.0 ** -6 * float(ref.abs().max())


def test_prepare_kernel_is_exact_layout():
    """eval_prepare against torch ops on the GPU: resized block, zero padding, and the mirrored copy == flip of the padded image."""
    import torch.nn.functional as F
    from segmentron_b200 import lib as L
    from segmentron_b200.ops import _ptr, _stream
    img = torch.randn(2, 3, 37, 53, generator=torch.Generator().manual_seed(9)).cuda()
    height, width, hp, wp = 28, 40, 36, 48
    out = torch.full((4, 3, hp, wp), float("nan"), device="cuda")
    L.check(L.load().segb200_eval_prepare(_ptr(img), _ptr(out), 2, 3, 37, 53, height, width, hp, wp, 1, _stream()))
    ref = F.pad(F.interpolate(img, size=[height, width], mode="bilinear", align_corners=True), (0, wp - width, 0, hp - height))
    assert float((out[:2] - ref).abs().max()) <= 1e-5
    assert torch.equal(out[2:], out[:2].flip(3)) and torch.equal(out[:2, :, height:, :], torch.zeros_like(out[:2, :, height:, :]))

"""The evaluation driver's host logic AND the algorithm of its two kernels, without a GPU.

csrc/evaluate.cu cannot run here, so the test substitutes a Python transcription of the two kernels that works on the same raw
pointers and integer arguments the C ABI receives (index rules, padding, mirrored reads, rounding points follow the .cu file line
by line, in float32) and runs `segmentron_b200.evaluate.evaluate` end to end on CPU tensors (ops._PLAN_DRY_RUN, a tests-only
switch).  Result vs the scores recorded from the reference's own SegBaseModel.evaluate (tests/golden/evaluate_cases.pt).
What this does NOT cover -- launch geometry, coalescing, the CUDA build of the same formulas -- is tests/test_evaluate_gpu.py."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest
import torch

G = os.path.join(os.path.dirname(__file__), "golden")


def _view(ptr, shape, ctype=C.c_float, np_dtype=np.float32):
    n = int(np.prod(shape))
    addr = ptr.value if isinstance(ptr, C.c_void_p) else ptr
    return np.frombuffer((ctype * n).from_address(addr), dtype=np_dtype).reshape(shape)


def _lerp(dst, n_in, n_out):
    """vec.cuh lerp_coord, align_corners=True, float32"""
    scale = np.float32(n_in - 1) / np.float32(n_out - 1) if n_out > 1 else np.float32(0)
    src = (scale * dst.astype(np.float32)).astype(np.float32)
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    i1 = i0 + (i0 < n_in - 1)
    l1 = (src - i0.astype(np.float32)).astype(np.float32)
    return i0, i1, (np.float32(1) - l1).astype(np.float32), l1


class FakeLib:
    """eval_prepare_kernel / eval_accumulate_kernel transcribed (vectorised over the grid-stride index)."""

    def segb200_eval_prepare(self, image, out, b, c, h, w, height, width, hp, wp, flip, stream):
        img = _view(image, (b * c, h, w))
        o = _view(out, ((2 if flip else 1) * b * c, hp, wp))
        y0, y1, ly0, ly1 = _lerp(np.arange(height), h, height)
        x0, x1, lx0, lx1 = _lerp(np.arange(width), w, width)
        v = ly0[None, :, None] * (lx0 * img[:, y0][:, :, x0] + lx1 * img[:, y0][:, :, x1]) + \
            ly1[None, :, None] * (lx0 * img[:, y1][:, :, x0] + lx1 * img[:, y1][:, :, x1])
        first = np.zeros((b * c, hp, wp), dtype=np.float32)
        first[:, :height, :width] = v
        o[:b * c] = first
        if flip:
            o[b * c:] = first[:, :, ::-1]                       # out[half + .. + (wp-1-x)] = v
        return 0

    def segb200_eval_accumulate(self, logits, scores, dtype, b, k, hp, wp, height, width, h, w, flip, accumulate, stream):
        assert dtype == 2                                        # this transcription covers fp32 logits
        lg = _view(logits, ((2 if flip else 1) * b * k, hp, wp))
        sc = _view(scores, (b * k, h, w))
        y0, y1, ly0, ly1 = _lerp(np.arange(h), height, h)
        x0, x1, lx0, lx1 = _lerp(np.arange(w), width, w)

        def tap(yy, xx):
            v = lg[:b * k][:, yy][:, :, xx]
            if flip:
                v = v + lg[b * k:][:, yy][:, :, wp - 1 - xx]
            return v
        o = ly0[None, :, None] * (lx0 * tap(y0, x0) + lx1 * tap(y0, x1)) + ly1[None, :, None] * (lx0 * tap(y1, x0) + lx1 * tap(y1, x1))
        sc[...] = (sc + o) if accumulate else o
        return 0

    def segb200_last_error(self):
        return b""


@pytest.mark.parametrize("case", range(6))
def test_driver_and_kernel_algorithm_reproduce_the_reference_scores(case, monkeypatch):
    from segmentron_b200 import evaluate as V, ops
    spec = importlib.util.spec_from_file_location("make_eval_golden", os.path.join(G, "make_eval_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    fx = torch.load(os.path.join(G, "evaluate_cases.pt"))[case]
    seed, b, h, w, scales, flip, crop = gen.CASES[case]
    monkeypatch.setattr(ops, "_PLAN_DRY_RUN", True)
    monkeypatch.setattr(V.L, "load", lambda: FakeLib())
    fwd = gen.stub_forward(seed)
    calls = []

    def forward(x):
        calls.append(tuple(x.shape))
        return (fwd(x),)
    with torch.no_grad():
        got = V.evaluate(forward, gen.make_image(seed, b, h, w), scales, flip, crop)
    ref = fx["scores"]
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 2e-6 * float(ref.abs().max()), float((got - ref).abs().max())
    assert len(calls) == len(scales) and all(s[0] == (2 if flip else 1) * b for s in calls)
    hp_wp = [V.padded_size(*V.scaled_size(h, w, s), V._to_tuple(crop) if crop else None, s) for s in scales]
    assert [s[2:] for s in calls] == [tuple(p) for p in hp_wp]

"""Device-resident evaluation metric (csrc/metric.cu, segmentron_b200/metric.py) on the B200: integer counts bit-exact against the
oracle (oracle/scoreref.py, itself pinned to the reference's segmentron/utils/score.py by tests/golden/score_cases.pt), the
accumulated pixAcc / mIoU bit-identical to the reference's recorded values, the fused-up-sampling source identical to the metric of
the engine's own full-resolution output, and size-independent properties at the full 8 x 1025 x 2049 size."""
import importlib.util
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu]

G = os.path.join(os.path.dirname(__file__), "golden")


def _gen():
    spec = importlib.util.spec_from_file_location("make_score_golden", os.path.join(G, "make_score_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    return gen


def _counts(m):
    return m._counts.cpu().numpy().copy()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_counts_match_oracle_and_reference_fixture(dtype):
    from oracle import scoreref as S
    from segmentron_b200 import lib as L
    from segmentron_b200.metric import SegmentationMetric
    from segmentron_b200.ops import _ptr, _stream, dt_code
    gen = _gen()
    fx = torch.load(os.path.join(G, "score_cases.pt"))
    lib = L.load()
    m19 = SegmentationMetric(19, False)
    ref19 = S.SegmentationMetric(19)
    for i, args in enumerate(gen.CASES):
        x, t = gen.make_inputs(*args)
        x = x.to(dtype)
        n, c, h, w = x.shape
        want = S.counts(x.float().numpy(), t.numpy(), c)
        cnt = torch.zeros(2 + 3 * c, dtype=torch.int64, device="cuda")
        xg, tg = x.cuda(), t.cuda()
        L.check(lib.segb200_seg_metric(_ptr(xg), dt_code(dtype), _ptr(tg), n, c, h, w, _ptr(cnt), _stream()))
        assert np.array_equal(cnt.cpu().numpy(), want), (args, cnt.cpu().numpy(), want)
        L.check(lib.segb200_seg_metric(_ptr(xg), dt_code(dtype), _ptr(tg), n, c, h, w, _ptr(cnt), _stream()))   # counts accumulate
        assert np.array_equal(cnt.cpu().numpy(), 2 * want)
        if dtype == torch.float32:
            assert (int(want[0]), int(want[1])) == (fx["cases"][i]["correct"], fx["cases"][i]["labeled"])
        if i < 6:
            m19.update(xg, tg)
            ref19.update(x.float().numpy(), t.numpy())
    pix, miou, iou = m19.get(return_category_iou=True)
    rp, rm, ri = ref19.get()
    assert (m19.total_correct, m19.total_label) == (ref19.total_correct, ref19.total_label)
    assert np.array_equal(m19.total_inter.cpu().numpy(), ref19.total_inter) and np.array_equal(m19.total_union.cpu().numpy(), ref19.total_union)
    assert pix == rp and np.array_equal(iou, ri) and abs(miou - rm) <= 1e-7
    if dtype == torch.float32:                                   # the reference's own recorded result
        acc = fx["accumulated"]
        assert pix == acc["pixAcc"] and abs(miou - acc["mIoU"]) <= 1e-7 and np.array_equal(iou, acc["IoU"].numpy())
    m19.reset()
    assert m19.get() == (0.0, 0.0)
    m19.update([torch.zeros(1, 19, 4, 4, device="cuda")], [torch.zeros(1, 4, 4, dtype=torch.long, device="cuda")])
    assert m19.get()[0] == 1.0                                   # list / tuple form (score.py:57-59)


@pytest.mark.parametrize("dtype,out_dtype", [(torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32), (torch.float16, torch.float16)])
def test_fused_upsampling_source_equals_metric_of_engine_output(dtype, out_dtype):
    from segmentron_b200 import ops
    from segmentron_b200.metric import SegmentationMetric
    g = torch.Generator().manual_seed(3)
    n, hi, wi, c, ld, ho, wo = 2, 17, 33, 19, 24, 65, 129
    low = torch.zeros(n, hi, wi, ld, dtype=dtype)
    low[..., :c] = (torch.randn(n, hi, wi, c, generator=g) * 2.0).to(dtype)
    low = low.cuda()
    t = torch.randint(-1, c, (n, ho, wo), generator=g).cuda()
    full = torch.empty(n, c, ho, wo, dtype=out_dtype, device="cuda")
    ops.bilinear_nchw_out(low, full, c, align_corners=True)
    a, b = SegmentationMetric(c), SegmentationMetric(c)
    a.update(full, t)
    b.update_lowres(low[..., :c], t, align_corners=True, out_dtype=out_dtype)
    assert a.total_label == b.total_label == int((t >= 0).sum()) and a.total_correct == b.total_correct
    assert torch.equal(a.total_inter, b.total_inter) and torch.equal(a.total_union, b.total_union)
    assert a.get() == b.get()


def test_full_size_properties():
    """8 x 19 x 1025 x 2049 logits (2.5 GB fp32 would not be what the engine emits: bf16): counts against torch ops on the GPU and
    the identities  sum(pred) == labeled,  sum(lab) == #{0 <= label < nclass},  inter <= min(pred, lab),  correct <= labeled."""
    from segmentron_b200.metric import SegmentationMetric
    n, c, h, w = 8, 19, 1025, 2049
    g = torch.Generator(device="cuda").manual_seed(11)
    x = (torch.randn(n, c, h, w, generator=g, device="cuda") * 2.0).to(torch.bfloat16)
    t = torch.randint(-1, c, (n, h, w), generator=g, device="cuda")
    m = SegmentationMetric(c)
    m.update(x, t)
    m2 = SegmentationMetric(c)
    m2._check(x, t, "update")
    from segmentron_b200 import lib as L
    from segmentron_b200.ops import _ptr, _stream, dt_code
    L.check(L.load().segb200_seg_metric(_ptr(x), dt_code(x.dtype), _ptr(t), n, c, h, w, _ptr(m2._counts), _stream()))
    cnt = m2._counts.cpu().numpy()
    valid = t >= 0
    pred = x.float().argmax(1)
    pred_t = x.float().trunc().long().argmax(1)
    assert cnt[1] == int(valid.sum()) and cnt[0] == int(((pred_t == t) & valid).sum()) and cnt[0] <= cnt[1]
    inter, parea, larea = cnt[2:2 + c], cnt[2 + c:2 + 2 * c], cnt[2 + 2 * c:]
    assert np.array_equal(parea, torch.bincount(pred[valid], minlength=c).cpu().numpy())
    assert np.array_equal(larea, torch.bincount(t[valid], minlength=c).cpu().numpy())
    assert np.array_equal(inter, torch.bincount(pred[valid & (pred == t)], minlength=c).cpu().numpy())
    assert parea.sum() == cnt[1] == larea.sum() and (inter <= np.minimum(parea, larea)).all()
    assert m.total_label == int(cnt[1])


def test_input_normalize_kernel_is_torchvision_bit_for_bit():
    """segb200_image_normalize against transforms.ToTensor() + transforms.Normalize on the same uint8 images: exact."""
    from PIL import Image
    from torchvision import transforms
    from segmentron_b200 import data as D
    g = torch.Generator().manual_seed(0)
    imgs = torch.randint(0, 256, (3, 65, 129, 3), generator=g, dtype=torch.uint8)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean, std)])
    ref = torch.stack([tf(Image.fromarray(im.numpy())) for im in imgs])
    got = D.normalize(imgs.cuda(), mean, std)
    assert got.dtype == torch.float32 and torch.equal(got.cpu(), ref)

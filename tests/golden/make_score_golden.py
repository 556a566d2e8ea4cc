"""Golden fixtures for the evaluation metric (oracle/scoreref.py): runs the REFERENCE's own ``segmentron/utils/score.py`` in the
build container (needs /root/reference, read-only) on seeded logits / labels and stores inputs + outputs in
``tests/golden/score_cases.pt``.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_score_golden.py

Cases cover: float ties and integer-valued logits (the truncated argmax of score.py:86 differs from the float argmax of :102),
ignored labels (-1), labels >= nclass, an all-ignored batch, several updates accumulated in the float32 totals.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def make_inputs(seed, n, c, h, w, kind):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, c, h, w, generator=g) * 3.0
    if kind == "integers":
        x = torch.round(x)                                        # many exact ties
    elif kind == "small":
        x = x * 0.2                                               # |x| < 1 mostly: the truncated logits are all 0 -> class 0
    elif kind == "bf16":
        x = x.to(torch.bfloat16).float()
    t = torch.randint(-1, c, (n, h, w), generator=g)
    if kind == "over":
        t = torch.randint(-1, c + 3, (n, h, w), generator=g)      # labels beyond the last class
    if kind == "ignored":
        t = torch.full((n, h, w), -1, dtype=torch.long)
    return x, t


CASES = [(0, 2, 19, 33, 47, "plain"), (1, 1, 19, 64, 64, "integers"), (2, 3, 19, 17, 29, "small"), (3, 2, 19, 40, 24, "bf16"),
         (4, 2, 19, 31, 31, "over"), (5, 1, 19, 16, 16, "ignored"), (6, 2, 7, 25, 35, "plain"), (7, 1, 32, 20, 20, "integers")]


def main():
    sys.path.insert(0, "/root/reference")
    from segmentron.utils.score import SegmentationMetric, batch_intersection_union, batch_pix_accuracy
    out = {"cases": []}
    for seed, n, c, h, w, kind in CASES:
        x, t = make_inputs(seed, n, c, h, w, kind)
        correct, labeled = batch_pix_accuracy(x, t)
        inter, union = batch_intersection_union(x, t, c)
        out["cases"].append({"args": (seed, n, c, h, w, kind), "correct": int(correct), "labeled": int(labeled),
                             "inter": inter.clone(), "union": union.clone()})
    m = SegmentationMetric(19, False)
    torch.cuda.synchronize = lambda: None                          # score.py:49 synchronises unconditionally; CPU container
    for seed, n, c, h, w, kind in CASES[:6]:
        x, t = make_inputs(seed, n, c, h, w, kind)
        m.update(x, t)
    pix_acc, miou, iou = m.get(return_category_iou=True)
    out["accumulated"] = {"pixAcc": float(pix_acc), "mIoU": float(miou), "IoU": torch.from_numpy(iou.copy()),
                          "total_inter": m.total_inter.clone(), "total_union": m.total_union.clone(),
                          "total_correct": int(m.total_correct), "total_label": int(m.total_label)}
    torch.save(out, os.path.join(HERE, "score_cases.pt"))
    print("wrote score_cases.pt:", {k: (v["correct"], v["labeled"]) for k, v in zip(range(len(CASES)), out["cases"])})


if __name__ == "__main__":
    main()

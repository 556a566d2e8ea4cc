"""Golden fixtures for the evaluation driver (oracle/evalref.py): calls the REFERENCE's own ``SegBaseModel.evaluate``
(segmentron/models/segbase.py:44-79, unbound, on a stub object whose ``forward`` is a small seeded conv net) in the build
container and stores the scores in ``tests/golden/evaluate_cases.pt``.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_eval_golden.py
"""
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))

# (seed, batch, h, w, scales, flip, crop_size)
CASES = [
    (0, 1, 33, 65, [1.0], False, None),                       # the default config: one plain forward
    (1, 2, 33, 65, [1.0], True, None),
    (2, 1, 40, 72, [0.5, 1.0, 1.75], True, None),
    (3, 2, 37, 53, [0.75, 1.25], True, (48, 64)),             # zero padding up to ceil(crop * scale); flip moves the padding left
    (4, 1, 57, 31, [0.5, 1.5], False, (64, 64)),              # portrait: h > w branch of segbase.py:54-56
    (5, 1, 24, 24, [2.0], True, 24),                          # scalar crop size (_to_tuple :119-127)
]


def stub_forward(seed, nclass=5):
    """A tiny 'segmentation model': stride-4 conv features -> 1x1 classifier -> bilinear up-sampling to the input size."""
    g = torch.Generator().manual_seed(1000 + seed)
    w1 = torch.randn(8, 3, 3, 3, generator=g) * 0.3
    w2 = torch.randn(nclass, 8, 1, 1, generator=g) * 0.5
    b2 = torch.randn(nclass, generator=g)

    def forward(x):
        y = F.relu(F.conv2d(x, w1, None, stride=4, padding=1))
        y = F.conv2d(y, w2, b2)
        return F.interpolate(y, x.shape[2:], mode="bilinear", align_corners=True)
    return forward


def make_image(seed, b, h, w):
    return torch.randn(b, 3, h, w, generator=torch.Generator().manual_seed(seed))


def main():
    sys.path.insert(0, "/root/reference")
    from segmentron.config import cfg
    from segmentron.models.segbase import SegBaseModel
    out = []
    for seed, b, h, w, scales, flip, crop in CASES:
        cfg.TEST.SCALES, cfg.TEST.FLIP, cfg.TEST.CROP_SIZE = scales, flip, crop
        fwd = stub_forward(seed)

        class Stub:
            def forward(self, x):
                return (fwd(x),)                                  # models return a tuple; evaluate takes [0] (segbase.py:69)
        with torch.no_grad():
            scores = SegBaseModel.evaluate(Stub(), make_image(seed, b, h, w))
        out.append({"args": (seed, b, h, w, scales, flip, crop), "scores": scores.clone()})
        print(seed, tuple(scores.shape), float(scores.abs().mean()))
    torch.save(out, os.path.join(HERE, "evaluate_cases.pt"))


if __name__ == "__main__":
    main()

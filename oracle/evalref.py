"""CPU ORACLE for the multi-scale + flip evaluation driver.  TEST INFRASTRUCTURE ONLY.

Restatement of ``SegBaseModel.evaluate`` (segmentron/models/segbase.py:44-79 with its helpers ``_resize_image`` :82-83,
``_pad_image`` :86-107, ``_flip_image`` :114-116) over an arbitrary ``forward(image) -> logits`` callable.  Only ``tests/`` may
import it.  Pinned to the real method by ``tests/golden/evaluate_cases.pt`` (``tests/golden/make_eval_golden.py`` calls the
reference's unbound ``SegBaseModel.evaluate`` on a stub object in the build container).
"""
import math

import torch.nn.functional as F


def scaled_size(h, w, scale):
    """segbase.py:53-60: the long side becomes ceil(max(h, w) * scale), the short side keeps the aspect ratio (round half up)."""
    long_size = int(math.ceil(max(h, w) * scale))
    if h > w:
        return long_size, int(1.0 * w * long_size / h + 0.5)
    return int(1.0 * h * long_size / w + 0.5), long_size


def padded_size(height, width, crop_size, scale):
    """segbase.py:64-68 + _pad_image (:86-107).  The reference computes padh = crop_h - height, padw = crop_w - width and then
    calls ``F.pad(img, (0, padh, 0, padw))`` (:93) -- F.pad's tuple is (left, right, top, bottom), so the HEIGHT deficit is
    appended on the RIGHT and the WIDTH deficit at the BOTTOM.  Kept as is (the fixture pins it): the padded image is
    [height + padw, width + padh], which equals the intended crop only when both deficits agree."""
    if crop_size is None:
        return height, width
    ch, cw = int(math.ceil(crop_size[0] * scale)), int(math.ceil(crop_size[1] * scale))
    padh, padw = max(0, ch - height), max(0, cw - width)
    return height + padw, width + padh


def evaluate(forward, image, scales=(1.0,), flip=False, crop_size=None):
    """-> scores [B, nclass, h, w]: sum over scales of the (flip-averaged, un-normalised) logits resized back to the input size."""
    if isinstance(crop_size, (int, float)):
        crop_size = (crop_size, crop_size)
    _, _, h, w = image.shape
    scores = None
    for scale in scales:
        height, width = scaled_size(h, w, scale)
        cur = F.interpolate(image, size=[height, width], mode="bilinear", align_corners=True)            # :63
        if crop_size is not None:
            assert crop_size[0] >= h and crop_size[1] >= w                                                # :65
            hp, wp = padded_size(height, width, crop_size, scale)
            cur = F.pad(cur, (0, wp - width, 0, hp - height))                                             # :68, _pad_image
        out = forward(cur)[..., :height, :width]                                                          # :69
        if flip:
            out = out + forward(cur.flip(3)).flip(3)[..., :height, :width]                                # :70-71
        score = F.interpolate(out, size=[h, w], mode="bilinear", align_corners=True)                      # :73
        scores = score if scores is None else scores + score                                              # :75-78
    return scores

"""Multi-scale + flip evaluation driver: the replacement for ``SegBaseModel.evaluate`` (segmentron/models/segbase.py:44-79).

    scores = evaluate(forward, image, scales=cfg.TEST.SCALES, flip=cfg.TEST.FLIP, crop_size=cfg.TEST.CROP_SIZE)

``forward`` maps a CUDA fp32 NCHW batch to full-resolution logits [B', nclass, H', W'] (a tensor, or the reference models' tuple
whose first entry is that tensor) -- e.g. a ``patch.accelerate``d reference model or ``engine.DeepLabV3PlusB200``.  Per scale the
reference issues F.interpolate -> F.pad -> forward -> flip -> forward -> flip -> += -> crop -> F.interpolate -> += as separate
ops; here one kernel builds the resized + padded batch together with its mirrored copy (``segb200_eval_prepare``), the model runs
ONCE on the stacked 2B images, and one kernel folds flip-add, crop, resize and accumulation (``segb200_eval_accumulate``).
The default configuration (scales [1.0], no flip, no crop) degenerates to a single plain forward, with which the reference's two
identity resizes agree exactly (align_corners=True at equal size is the identity).

Size rules are the reference's, including its transposed padding amounts (``F.pad(img, (0, padh, 0, padw))``, segbase.py:93:
the height deficit lands on the right, the width deficit at the bottom) -- kept so that scores stay comparable.
No CPU implementation: non-CUDA images raise RuntimeError.
"""
import math

import torch

from . import lib as L
from . import ops
from .ops import _ptr, dt_code


def _to_tuple(size):
    """segbase.py:119-127"""
    if isinstance(size, (list, tuple)):
        if len(size) != 2:
            raise RuntimeError(f"segb200: eval crop size must have two elements, got {len(size)}")
        return tuple(size)
    return (size, size)


def scaled_size(h, w, scale):
    """segbase.py:53-60 -> (height, width) of the resized image for one test scale."""
    long_size = int(math.ceil(max(h, w) * scale))
    if h > w:
        return long_size, int(1.0 * w * long_size / h + 0.5)
    return int(1.0 * h * long_size / w + 0.5), long_size


def padded_size(height, width, crop_size, scale):
    """segbase.py:64-68 with _pad_image's argument order (:93): -> (hp, wp) = (height + padw, width + padh)."""
    if crop_size is None:
        return height, width
    ch, cw = int(math.ceil(crop_size[0] * scale)), int(math.ceil(crop_size[1] * scale))
    return height + max(0, cw - width), width + max(0, ch - height)


def _stream(t):
    return ops._stream() if t.is_cuda else None


def _logits(out):
    return out[0] if isinstance(out, (tuple, list)) else out


def evaluate(forward, image, scales=(1.0,), flip=False, crop_size=None):
    """-> scores [B, nclass, h, w] in the dtype of the model's logits (segbase.py:44-79)."""
    if not image.is_cuda and not ops._PLAN_DRY_RUN:       # dry run: tests drive the host logic against an emulated library
        raise RuntimeError("segb200: evaluate is not implemented on the CPU (the image must be a CUDA tensor)")
    if image.dim() != 4 or image.dtype != torch.float32:
        raise RuntimeError("segb200: evaluate expects an fp32 NCHW image batch")
    image = image.contiguous()
    b, c, h, w = image.shape
    crop = _to_tuple(crop_size) if crop_size else None
    if crop is not None and not (crop[0] >= h and crop[1] >= w):
        raise RuntimeError(f"segb200: TEST.CROP_SIZE {crop} is smaller than the image {(h, w)} (segbase.py:65)")
    scales = list(scales)
    lib = L.load()
    scores = None
    for scale in scales:
        height, width = scaled_size(h, w, scale)
        hp, wp = padded_size(height, width, crop, scale)
        if not flip and (height, width, hp, wp) == (h, w, h, w):
            logits = _logits(forward(image))                       # both resizes are identities at this scale
            if len(scales) == 1:
                # a fresh tensor like every other path of this function: an engine forward hands out its static (CUDA-graph)
                # output buffer, which the next call overwrites
                return logits.clone()
        else:
            batch = torch.empty((2 if flip else 1) * b, c, hp, wp, dtype=torch.float32, device=image.device)
            L.check(lib.segb200_eval_prepare(_ptr(image), _ptr(batch), b, c, h, w, height, width, hp, wp, int(flip), _stream(image)),
                    "eval_prepare")
            logits = _logits(forward(batch))
        if logits.dim() != 4 or logits.shape[0] != (2 if flip else 1) * b or tuple(logits.shape[2:]) != (hp, wp):
            raise RuntimeError(f"segb200: forward returned {tuple(logits.shape)}, expected [{(2 if flip else 1) * b}, nclass, {hp}, {wp}]")
        logits = logits.contiguous()
        k = logits.shape[1]
        if scores is None:
            scores = torch.empty(b, k, h, w, dtype=logits.dtype, device=image.device)
            acc = 0
        else:
            acc = 1
        L.check(lib.segb200_eval_accumulate(_ptr(logits), _ptr(scores), dt_code(logits.dtype), b, k, hp, wp, height, width, h, w,
                                            int(flip), acc, _stream(image)), "eval_accumulate")
    return scores

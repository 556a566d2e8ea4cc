"""GPU-side input transform: ``transforms.ToTensor()`` + ``transforms.Normalize(mean, std)`` of the reference's data pipeline
(tools/train.py:36-39, tools/eval.py:33-36) for decoded uint8 images that are already on the device, so the loader ships 1 byte
per sample instead of 4 and the CPU workers stop doing float arithmetic.  Bit-identical to torchvision (same IEEE operations in
the same order).  No CPU implementation."""
import torch

from . import lib as L
from .ops import _ptr, _stream


def normalize(images_u8_nhwc, mean, std):
    """images_u8_nhwc: uint8 CUDA tensor [N, H, W, C] (C <= 4); mean / std: C floats (cfg.DATASET.MEAN / STD) -> fp32 [N, C, H, W]."""
    x = images_u8_nhwc
    if not x.is_cuda:
        raise RuntimeError("segb200: normalize is not implemented on the CPU (the image batch must be a CUDA tensor)")
    if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[3] > 4:
        raise RuntimeError("segb200: normalize expects a uint8 [N,H,W,C] batch with C <= 4")
    x = x.contiguous()
    n, h, w, c = x.shape
    ms = torch.tensor(list(mean) + list(std), dtype=torch.float32, device=x.device)
    if ms.numel() != 2 * c:
        raise RuntimeError("segb200: mean / std must have one entry per channel")
    out = torch.empty(n, c, h, w, dtype=torch.float32, device=x.device)
    L.check(L.load().segb200_image_normalize(_ptr(x), _ptr(out), n, h, w, c, _ptr(ms[:c]), _ptr(ms[c:]), _stream()), "image_normalize")
    return out

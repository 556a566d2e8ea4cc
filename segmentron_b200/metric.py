"""Device-resident evaluation metric: the drop-in for ``segmentron.utils.score.SegmentationMetric`` (score.py:11-81).

Same constructor, ``update(preds, labels)``, ``get(return_category_iou=False)``, ``reset()`` and the same numbers: the integer
areas that the reference obtains with two argmax passes, three host transfers and three ``torch.histc`` calls per batch behind a
``torch.cuda.synchronize()`` (score.py:49,86,102,108-110) come from ONE kernel over the logits (``segb200_seg_metric``); the totals
live on the device (two int64 pixel counters, float32 per-class totals advanced once per update exactly like the reference's
``total_inter += inter``) and nothing synchronises before ``get()``.  Distributed: the batch counts are all-reduced (SUM) as one
int64 tensor before they are accumulated (the reference all-reduces its four tensors one by one, score.py:32-44).

``update_lowres`` feeds the classifier's low-resolution NHWC logits instead and fuses the final bilinear up-sampling, so an
evaluation loop never materialises the [N, nclass, H, W] logits.

Reference behaviours kept on purpose (see oracle/scoreref.py for the list): the pixel accuracy uses the argmax of the logits
truncated to integers; labels >= nclass count as labeled but fall out of the label histogram.  Not reproduced: ``torch.histc``
accumulating a single batch's area in float32 beyond 2**24 pixels of one class (the counts here are exact integers).

No CPU implementation: non-CUDA inputs raise RuntimeError.
"""
import torch

from . import lib as L
from .ops import _ptr, _stream, dt_code, _nhwc


class SegmentationMetric:
    """Computes pixAcc and mIoU (score.py:11-81) without leaving the device."""

    def __init__(self, nclass, distributed=False, device=None):
        if not 1 <= nclass <= 64:
            raise RuntimeError("segb200: SegmentationMetric supports 1..64 classes")
        self.nclass = nclass
        self.distributed = distributed
        self.device = torch.device(device) if device is not None else None
        self._counts = None
        self.reset()

    # -- state ------------------------------------------------------------------------------------------------------
    def _alloc(self, device):
        self.device = device
        k = self.nclass
        self._counts = torch.zeros(2 + 3 * k, dtype=torch.int64, device=device)        # written as unsigned 64-bit
        self._pixels = torch.zeros(2, dtype=torch.int64, device=device)                # total_correct, total_label
        self.total_inter = torch.zeros(k, dtype=torch.float32, device=device)
        self.total_union = torch.zeros(k, dtype=torch.float32, device=device)

    def reset(self):
        """Resets the internal evaluation result to initial state (score.py:76-81)."""
        if self._counts is not None:
            for t in (self._counts, self._pixels, self.total_inter, self.total_union):
                t.zero_()
        else:
            self.total_inter = torch.zeros(self.nclass)
            self.total_union = torch.zeros(self.nclass)

    @property
    def total_correct(self):
        return int(self._pixels[0]) if self._counts is not None else 0

    @property
    def total_label(self):
        return int(self._pixels[1]) if self._counts is not None else 0

    # -- updates ----------------------------------------------------------------------------------------------------
    def _check(self, t, labels, what):
        if not t.is_cuda or not labels.is_cuda:
            raise RuntimeError(f"segb200: SegmentationMetric.{what} is not implemented on the CPU (CUDA tensors expected)")
        if self._counts is None or self.device != t.device:
            if self._counts is not None and (self.total_label or float(self.total_union.sum())):
                raise RuntimeError("segb200: SegmentationMetric was already updated on another device")
            self._alloc(t.device)
        if labels.dtype != torch.int64:
            labels = labels.long()
        return labels.contiguous()

    def _finish(self):
        lib = L.load()
        if self.distributed:
            torch.distributed.all_reduce(self._counts, op=torch.distributed.ReduceOp.SUM)
        L.check(lib.segb200_seg_metric_accumulate(_ptr(self._counts), self.nclass, _ptr(self._pixels), _ptr(self.total_inter),
                                                  _ptr(self.total_union), _stream()), "seg_metric_accumulate")

    def _update_one(self, pred, label):
        label = self._check(pred, label, "update")
        if pred.dim() != 4 or pred.shape[1] != self.nclass or tuple(label.shape) != (pred.shape[0],) + tuple(pred.shape[2:]):
            raise RuntimeError(f"segb200: expected logits [N,{self.nclass},H,W] and labels [N,H,W], got {tuple(pred.shape)} / "
                               f"{tuple(label.shape)}")
        pred = pred.contiguous()
        n, _, h, w = pred.shape
        L.check(L.load().segb200_seg_metric(_ptr(pred), dt_code(pred.dtype), _ptr(label), n, self.nclass, h, w, _ptr(self._counts),
                                            _stream()), "seg_metric")
        self._finish()

    def update(self, preds, labels):
        """preds: logits [N, nclass, H, W] (fp32 / bf16 / fp16, CUDA) or a list/tuple of them; labels likewise [N, H, W]
        (negative = ignored).  One kernel + one tiny accumulate kernel per tensor; asynchronous."""
        if isinstance(preds, torch.Tensor):
            self._update_one(preds, labels)
        elif isinstance(preds, (list, tuple)):
            for pred, label in zip(preds, labels):
                self._update_one(pred, label)
        else:
            raise RuntimeError("segb200: preds must be a tensor or a list / tuple of tensors")

    def update_lowres(self, logits_nhwc, labels, align_corners=True, out_dtype=None):
        """logits_nhwc: the classifier's NHWC 16-bit output [N, hi, wi, >= nclass] (a channel slice of a padded buffer is fine);
        labels [N, H, W].  Equivalent to update(engine_output, labels) where engine_output is what segb200_bilinear_nchw_out
        would have written in `out_dtype` (default: the logits' dtype)."""
        labels = self._check(logits_nhwc, labels, "update_lowres")
        n, hi, wi, c, ld = _nhwc(logits_nhwc, "logits_nhwc")
        if c < self.nclass or labels.dim() != 3 or labels.shape[0] != n:
            raise RuntimeError("segb200: update_lowres shape mismatch")
        L.check(L.load().segb200_seg_metric_lowres(_ptr(logits_nhwc), dt_code(logits_nhwc.dtype), ld, hi, wi, int(align_corners),
                                                   dt_code(out_dtype or logits_nhwc.dtype), _ptr(labels), n, self.nclass,
                                                   labels.shape[1], labels.shape[2], _ptr(self._counts), _stream()),
                "seg_metric_lowres")
        self._finish()

    # -- result -----------------------------------------------------------------------------------------------------
    def get(self, return_category_iou=False):
        """-> (pixAcc, mIoU[, per-class IoU as numpy]) with the reference's arithmetic (score.py:69-74), evaluated on the host like
        the reference's single-process path (its totals are CPU tensors there); the only host synchronisation."""
        pixAcc = 1.0 * self.total_correct / (2.220446049250313e-16 + self.total_label)
        IoU = 1.0 * self.total_inter.cpu() / (2.220446049250313e-16 + self.total_union.cpu())
        mIoU = IoU.mean().item()
        if return_category_iou:
            return pixAcc, mIoU, IoU.numpy()
        return pixAcc, mIoU

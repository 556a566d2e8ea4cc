"""CPU ORACLE for the evaluation metric of the hot path.  TEST INFRASTRUCTURE ONLY.

Integer restatement (numpy ``bincount`` over class ids) of ``segmentron/utils/score.py``: ``batch_pix_accuracy`` (:83-93),
``batch_intersection_union`` (:96-113) and the accumulator ``SegmentationMetric`` (:11-81).  Only ``tests/`` may import it.
Pinned to the real reference by ``tests/golden/score_cases.pt`` (written by ``tests/golden/make_score_golden.py``, which runs the
reference's own functions in the build container).

Reference behaviours that are kept on purpose:
  * pixel accuracy takes the argmax of the logits TRUNCATED to integers (``output.long()``, score.py:86), so it can
    disagree with the argmax of the float logits used for the IoU (:102); ties go to the lowest class index;
  * a label is "labeled" when ``label >= 0`` (the ``+1`` / ``> 0`` dance, :87,:89,:103); labels >= nclass still count as
    labeled and in the predicted areas, but fall outside ``histc``'s range for the label areas (:111);
  * the per-class totals are float32 tensors and are accumulated per batch in float32 (:54-55,:77-78), the two pixel
    totals are Python ints (:50-51).
"""
import numpy as np


def counts(logits, target, nclass):
    """-> int64 vector [2 + 3*nclass]: correct, labeled, inter[c], pred_area[c], label_area[c].

    logits: float array [N, C, H, W]; target: integer array [N, H, W]."""
    logits = np.asarray(logits, dtype=np.float32)
    target = np.asarray(target).astype(np.int64)
    n, c, h, w = logits.shape
    assert c == nclass
    valid = target >= 0                                                   # target + 1 > 0
    pred_trunc = np.argmax(np.trunc(logits).astype(np.int64), axis=1)     # score.py:86 (first maximum)
    pred = np.argmax(logits, axis=1)                                      # score.py:102
    out = np.zeros(2 + 3 * nclass, dtype=np.int64)
    out[0] = np.count_nonzero((pred_trunc == target) & valid)             # :90
    out[1] = np.count_nonzero(valid)                                      # :89
    out[2:2 + nclass] = np.bincount(pred[valid & (pred == target)], minlength=nclass)[:nclass]             # :106,:109
    out[2 + nclass:2 + 2 * nclass] = np.bincount(pred[valid], minlength=nclass)[:nclass]                   # :105,:110
    in_range = valid & (target < nclass)
    out[2 + 2 * nclass:] = np.bincount(target[in_range], minlength=nclass)[:nclass]                       # :111
    return out


class SegmentationMetric:
    """score.py:11-81 on top of ``counts`` (single process; the distributed variant sums the counts over ranks first)."""

    def __init__(self, nclass):
        self.nclass = nclass
        self.reset()

    def reset(self):
        self.total_inter = np.zeros(self.nclass, dtype=np.float32)
        self.total_union = np.zeros(self.nclass, dtype=np.float32)
        self.total_correct = 0
        self.total_label = 0

    def update_counts(self, cnt):
        k = self.nclass
        inter = cnt[2:2 + k].astype(np.float32)
        union = cnt[2 + k:2 + 2 * k].astype(np.float32) + cnt[2 + 2 * k:].astype(np.float32) - inter      # :112
        self.total_correct += int(cnt[0])
        self.total_label += int(cnt[1])
        self.total_inter = (self.total_inter + inter).astype(np.float32)
        self.total_union = (self.total_union + union).astype(np.float32)

    def update(self, logits, target):
        self.update_counts(counts(logits, target, self.nclass))

    def get(self):
        eps = np.float64(2.220446049250313e-16)
        pix_acc = 1.0 * self.total_correct / (eps + self.total_label)                                     # :69
        iou = (1.0 * self.total_inter / (np.float32(2.220446049250313e-16) + self.total_union)).astype(np.float32)   # :70
        return pix_acc, float(iou.mean(dtype=np.float32)), iou

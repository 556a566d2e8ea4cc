"""The reference in the loop, on the B200: its UNMODIFIED scripts (staged under baseline/_ref by tools/make_baseline_ref.py) run
(a) stock -- plain reference, torch/cuDNN fp32 -- and (b) through ``python -m segmentron_b200.launch`` -- the same script, YAML,
registry and checkpoint format, with the L1 modules rebound to the segb200 drop-ins -- on the same synthetic Cityscapes tree,
same seed, same GPU; the numbers the scripts log are compared.

  tools/train.py (tools/train.py:128-147 + validation :163-196):  per-iteration losses, first iteration within 3 % (bf16 compute
      vs fp32), later ones within 20 % (train-mode BatchNorm makes the net chaotic), all finite; the
      validation pass runs the reference's SegmentationMetric on the drop-ins' outputs and writes best_model.pth;
  tools/eval.py  (tools/eval.py:78-85 -> SegBaseModel.evaluate, segbase.py:44-79): pixAcc / mIoU of the checkpoint the STOCK
      training run saved, stock vs drop-in modules vs the whole-model plan (``--accelerate``): within 1.5 points.
"""
import glob
import math
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
pytestmark = pytest.mark.gpu

YAML = "configs/cityscapes_deeplabv3_plus.yaml"          # DeepLabv3+ / Xception65: the headline model
COMMON = ["TRAIN.BACKBONE_PRETRAINED", "False", "DATASET.WORKERS", "0", "TEST.BATCH_SIZE", "1"]
TRAIN = ["--config-file", YAML, "--log-iter", "1", *COMMON, "TRAIN.EPOCHS", "1", "TRAIN.BATCH_SIZE", "2", "TRAIN.CROP_SIZE", "129",
         "TRAIN.BASE_SIZE", "160"]


@pytest.fixture(scope="module")
def runs(tmp_path_factory):
    import ref_harness as H
    if not H.available():
        pytest.fail("baseline/_ref is not staged: run `python tools/make_baseline_ref.py` in the build container")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_baseline_ref
    assert make_baseline_ref.verify(), "baseline/_ref differs from its manifest: the reference copy must stay unmodified"
    out = {}
    for tag, through in (("stock", False), ("ours", True)):
        d = str(tmp_path_factory.mktemp("run_" + tag))
        H.make_run_dir(d, n_train=8, n_val=4, h=128, w=256, seed=3)
        rc, log = H.run_script(d, "train.py", TRAIN, through_launch=through)
        assert rc == 0, f"tools/train.py ({tag}) failed:\n{log[-3000:]}"
        out[tag] = dict(dir=d, log=log, losses=H.parse_train_losses(log))
    return out


def test_train_py_through_launch_matches_stock(runs):
    s, o = runs["stock"]["losses"], runs["ours"]["losses"]
    print(f"\n[tools/train.py] stock losses {s}\n[tools/train.py] segb200 drop-in losses {o}")
    assert "rebound" in runs["ours"]["log"] and "rebound" not in runs["stock"]["log"]
    assert len(s) == len(o) == 4
    assert all(math.isfinite(v) for v in s + o)
    assert abs(o[0] - s[0]) <= 0.03 * s[0], (o[0], s[0])
    for a, b in zip(o[1:], s[1:]):
        assert abs(a - b) <= 0.20 * b, (o, s)
    for tag in ("stock", "ours"):                         # validation ran through SegmentationMetric and a checkpoint was written
        assert "[EVAL END]" in runs[tag]["log"], runs[tag]["log"][-2000:]
        assert glob.glob(os.path.join(runs[tag]["dir"], "runs", "checkpoints", "*", "1.pth"))


def test_eval_py_through_launch_matches_stock(runs):
    import ref_harness as H
    import torch
    ck = glob.glob(os.path.join(runs["stock"]["dir"], "runs", "checkpoints", "*", "best_model.pth"))
    if not ck:                                            # mIoU 0 on the first validation: fall back to the epoch checkpoint
        ep = glob.glob(os.path.join(runs["stock"]["dir"], "runs", "checkpoints", "*", "1.pth"))[0]
        ck = [os.path.join(os.path.dirname(ep), "plain.pth")]
        torch.save(torch.load(ep, map_location="cpu", weights_only=False)["state_dict"], ck[0])
    argv = ["--config-file", YAML, *COMMON, "TEST.TEST_MODEL_PATH", ck[0]]
    res = {}
    for tag, through, flags in (("stock", False, ()), ("drop-ins", True, ()), ("plan", True, ("--accelerate",))):
        rc, log = H.run_script(runs["stock"]["dir"], "eval.py", argv, through_launch=through, launch_flags=flags)
        assert rc == 0, f"tools/eval.py ({tag}) failed:\n{log[-3000:]}"
        res[tag] = H.parse_eval_result(log)
        assert res[tag] is not None, log[-2000:]
    print(f"\n[tools/eval.py] (pixAcc %, mIoU %): {res}")
    for tag in ("drop-ins", "plan"):
        assert abs(res[tag][0] - res["stock"][0]) <= 1.5 and abs(res[tag][1] - res["stock"][1]) <= 1.5, res

"""Drop-in mechanics against the REAL reference (build container only: needs /root/reference; skipped elsewhere):
class rebinding covers every consumer namespace, state_dict keys/shapes are unchanged, convert_to_b200 shares
parameters, and the swapped modules refuse to run on the CPU (no fallback)."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "segmentron")), reason="reference tree not present")

SCRIPT = r'''
import sys, numpy as np
np.int = int
sys.path.insert(0, "%(root)s"); sys.path.insert(0, "%(ref)s")
import torch
import segmentron
from segmentron.config import cfg
from segmentron.models.model_zoo import get_segmentation_model
from segmentron_b200 import patch, modules as M
from oracle import segref as R
cfg.update_from_file("%(ref)s/configs/%(yaml)s")
cfg.PHASE = "test"; cfg.check_and_freeze()
ref_model = get_segmentation_model().eval()                 # built from reference classes
ref_keys = {k: tuple(v.shape) for k, v in ref_model.state_dict().items()}
n = patch.install()
assert n >= 7, n
model = get_segmentation_model().eval()                     # same registry / YAML, now built from the drop-ins
cnt = sum(isinstance(m, M.SeparableConv2d) for m in model.modules())
assert cnt == %(nsep)d, cnt
assert all(type(m).__module__.startswith("segmentron_b200") for m in model.modules() if type(m).__name__ in M.REPLACEMENTS)
assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == ref_keys
P = R.build_params("%(oracle)s", 0)
model.load_state_dict(P.state_dict(), strict=True)          # the oracle's dict (= a reference checkpoint) loads
conv = patch.convert_to_b200(ref_model)                     # in-place conversion of a reference-built model
assert sum(isinstance(m, M.SeparableConv2d) for m in conv.modules()) == %(nsep)d
assert {k: tuple(v.shape) for k, v in conv.state_dict().items()} == ref_keys
try:
    model(torch.zeros(1, 3, 33, 33))
    raise SystemExit("CPU forward did not raise")
except RuntimeError as e:
    assert "CPU" in str(e) or "CUDA" in str(e), e
print("OK")
'''


@pytest.mark.parametrize("yaml_file,oracle,nsep", [("cityscapes_deeplabv3_plus.yaml", "deeplabv3plus_xception65", 68),
                                                   ("cityscapes_deeplabv3_plus_mobilenet.yaml", "deeplabv3plus_mobilenet_v2", 2)])
def test_install_and_convert(yaml_file, oracle, nsep):
    code = SCRIPT % dict(root=ROOT, ref=REF, yaml=yaml_file, oracle=oracle, nsep=nsep)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_install_metric_and_evaluate_rebind_the_reference_names():
    """patch.install_metric / install_evaluate against the real reference tree: the names tools/eval.py uses resolve to the drop-ins
    (constructor signature of score.py:14 accepted; SegBaseModel.evaluate reads cfg.TEST and refuses CPU images loudly)."""
    script = r"""
import sys, numpy as np
np.int = int
sys.path.insert(0, "%(root)s"); sys.path.insert(0, "%(ref)s")
import torch
import segmentron.utils.score as score
from segmentron.models.segbase import SegBaseModel
from segmentron_b200 import patch, metric
assert patch.install_metric() >= 1
from segmentron.utils.score import SegmentationMetric
assert SegmentationMetric is metric.SegmentationMetric
m = SegmentationMetric(19, False)
assert m.get() == (0.0, 0.0)
cls = patch.install_evaluate()
assert cls is SegBaseModel and SegBaseModel.evaluate.__name__ == "_evaluate"
class Stub:
    forward = staticmethod(lambda x: (x,))
try:
    SegBaseModel.evaluate(Stub(), torch.zeros(1, 3, 8, 8))
    raise SystemExit("CPU evaluate did not raise")
except RuntimeError as e:
    assert "CPU" in str(e), e
print("OK")
""" % dict(root=ROOT, ref=REF)
    out = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr

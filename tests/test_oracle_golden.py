"""The oracle (oracle/segref.py) against the committed outputs of the REAL reference
(tests/golden/*.pt, written by tests/golden/make_golden.py in the build container)."""
import os

import pytest
import torch

from oracle import segref as R

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("case", ["dlv3p_xception65_65x129", "dlv3p_xception65_97x161_b2",
                                  "dlv3p_mobilenetv2_64x128", "dlv3p_resnet101_65x129",
                                  "danet_resnet101_64x96", "ccnet_resnet101_65x97", "hrnet_w18s_128x192"])
def test_model_matches_reference_fixture(case):
    fx = torch.load(os.path.join(G, case + ".pt"))
    P = R.build_params(fx["model"], fx["seed"])
    assert len(P.t) == fx["n_params"]
    x = torch.randn(*fx["shape"], generator=torch.Generator().manual_seed(fx["input_seed"]))
    y = R.forward(fx["model"], P, x)
    ref = fx["y_ref"]
    rel = float((y - ref).abs().max() / ref.abs().max())
    assert rel < 2e-5, rel                       # fp32 CPU kernels may differ by ISA
    mism = int((y.argmax(1) != fx["argmax"].long()).sum())
    assert mism <= y[:, 0].numel() // 5000, mism  # ties only


def test_modules_match_reference_fixture():
    fx = torch.load(os.path.join(G, "modules.pt"))
    with torch.no_grad():
        y = R.pam(R.Params(fx["pam"]["seed"]), fx["pam"]["x"], "pam")
        assert torch.allclose(y, fx["pam"]["y"], atol=1e-4)
        y = R.cam(R.Params(fx["cam"]["seed"]), fx["cam"]["x"], "cam")
        assert torch.allclose(y, fx["cam"]["y"], atol=1e-4)
        y = R.pyramid_pooling(R.Params(fx["psp"]["seed"]), fx["psp"]["x"], "psp")
        assert torch.allclose(y, fx["psp"]["y"], atol=1e-4)
        c = fx["cca"]
        assert torch.allclose(R.ca_weight(c["t"], c["f"]), c["weight"], atol=1e-5)
        assert torch.allclose(R.ca_map(c["att"], c["g"]), c["out"], atol=1e-5)


def test_param_names_are_reference_names():
    P = R.build_params("deeplabv3plus_xception65", 0)
    for k in ["encoder.conv1.weight", "encoder.bn1.running_var",
              "encoder.block1.sep_conv1.block.depthwise.weight",
              "encoder.block1.sep_conv1.block.bn_depth.num_batches_tracked",
              "encoder.block2.conv.weight", "encoder.block21.sep_conv3.block.bn_point.bias",
              "head.aspp.image_pooling.conv.weight", "head.aspp.aspp3.block.pointwise.weight",
              "head.aspp.conv.weight", "head.c1_block.bn.weight", "head.block.1.block.pointwise.weight",
              "head.block.2.bias"]:
        assert k in P.t, k
    assert tuple(P.t["head.aspp.conv.weight"].shape) == (256, 1280, 1, 1)
    assert tuple(P.t["encoder.block3.conv.weight"].shape) == (728, 256, 1, 1)

"""The oracle (oracle/segref.py) against the committed outputs of the REAL reference
(tests/golden/*.pt, written by tests/golden/make_golden.py in the build container)."""
import os

import pytest
import torch

from oracle import segref as R

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("case", ["dlv3p_xception65_65x129", "dlv3p_xception65_97x161_b2",
                                  "dlv3p_mobilenetv2_64x128", "dlv3p_resnet101_65x129",
                                  "danet_resnet101_64x96", "ccnet_resnet101_65x97", "hrnet_w18s_128x192", "pspnet_resnet101_65x97",
                                  "ocnet_resnet50_65x97"])
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


def _digest(grads, seed=12345):
    out = {}
    for i, k in enumerate(sorted(grads)):
        g = grads[k].detach().double().flatten()
        r = torch.randn(g.numel(), generator=torch.Generator().manual_seed(seed + i), dtype=torch.float64)
        out[k] = (float(g.norm()), float((g * r).sum()))
    return out


@pytest.mark.parametrize("case", ["train_dlv3p_resnet101_65x97_b4", "train_dlv3p_xception65_65x97_b4",
                                  "train_dlv3p_mobilenetv2_64x96_b4", "train_ccnet_resnet101_65x97_b2",
                                  "train_hrnet_w18s_64x96_b2", "train_danet_resnet101_64x96_b2"])
def test_training_step_matches_reference_fixture(case):
    """Oracle train step (train-mode BN, Dropout2d mask, CE(ignore -1), backward, SGD groups) against the real reference's
    tools/train.py iteration recorded in tests/golden/train_dlv3p_resnet101_65x97_b4.pt: loss, a (norm, random projection)
    digest of all 356 parameter gradients, selected full gradients, BN running statistics, parameters after optimizer.step()."""
    fx = torch.load(os.path.join(G, case + ".pt"))
    P = R.build_params(fx["model"], fx["seed"])
    n, _, h, w = fx["shape"]
    g = torch.Generator().manual_seed(fx["input_seed"])
    x = torch.randn(*fx["shape"], generator=g)
    target = torch.randint(-1, 19, (n, h, w), generator=g)
    P.dropout_masks[fx.get("mask_key", "head.aspp.dropout")] = fx["mask"]
    P.dropout_masks.update(fx.get("more_masks", {}))      # DANet: three Dropout2d layers
    P.bn_momentum = fx.get("bn_momentum", 0.1)          # MODEL.BN_MOMENTUM of the YAML (HRNet: 0.01), solver/optimizer.py:37-39
    before = {k: v.clone() for k, v in P.t.items()}
    loss, grads, out, low = R.loss_and_grads(fx["model"], P, x, target)
    assert abs(float(loss) - fx["loss"]) < 1e-4 * abs(fx["loss"])
    assert float((out[:, :, ::8, ::8] - fx["low"]).abs().max() / fx["low"].abs().max()) < 1e-4
    dg = _digest({k: v for k, v in grads.items() if k in fx["digest"]})
    assert set(dg) == set(fx["digest"])
    for k, (nrm, proj) in fx["digest"].items():
        assert abs(dg[k][0] - nrm) <= 2e-3 * nrm + 1e-9, (k, dg[k], (nrm, proj))
        assert abs(dg[k][1] - proj) <= 1e-2 * nrm + 1e-9, (k, dg[k], (nrm, proj))     # |<e, r>| ~ |e| for a unit-variance r
    for k, gr in fx["grads_small"].items():
        assert float((grads[k] - gr).norm() / gr.norm()) < 2e-3, k
    for k, v in fx["running"].items():
        assert torch.allclose(P.t[k], v, atol=1e-4), k
    # torch.optim.SGD first step with the reference's param groups (solver/optimizer.py:14-34,50-51): p -= lr (g + wd p)
    stepped = {}
    for k, (lr, wd, mom) in fx["hyper"].items():
        # parameters without a gradient (the unused ImageNet fc) are skipped by torch.optim.SGD
        stepped[k] = before[k] - lr * (grads[k] + wd * before[k]) if k in grads else before[k]
    sd = _digest(stepped, 999)
    for k, (nrm, proj) in fx["stepped_digest"].items():
        assert abs(sd[k][0] - nrm) <= 1e-4 * nrm + 1e-9, k
    enc0 = "encoder.conv1.weight" if "encoder.conv1.weight" in fx["hyper"] else "encoder.conv1.conv.weight"
    head0 = next(k for k in ("head.block.2.weight", "head.out.weight", "hrnet_head.last_layer.3.weight", "head.conv8.1.weight") if k in fx["hyper"])
    factor = 1.0 if "hrnet" in case else 10.0            # SOLVER.DECODER_LR_FACTOR: 10 in the DeepLab / CCNet YAMLs, default 1
    assert fx["hyper"][head0][0] == pytest.approx(factor * fx["hyper"][enc0][0])


def test_score_oracle_matches_reference_fixture():
    """oracle/scoreref.py (integer bincount restatement) against the outputs of the reference's own segmentron/utils/score.py
    (batch_pix_accuracy, batch_intersection_union, SegmentationMetric accumulated over six updates), recorded by
    tests/golden/make_score_golden.py: every count and the float32 totals are EXACT; pixAcc / mIoU bit-identical."""
    import importlib.util
    import numpy as np
    from oracle import scoreref as S
    spec = importlib.util.spec_from_file_location("make_score_golden", os.path.join(G, "make_score_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    fx = torch.load(os.path.join(G, "score_cases.pt"))
    assert len(fx["cases"]) == len(gen.CASES)
    m = S.SegmentationMetric(19)
    for i, (case, args) in enumerate(zip(fx["cases"], gen.CASES)):
        assert tuple(case["args"]) == tuple(args)
        x, t = gen.make_inputs(*args)
        c = args[2]
        cnt = S.counts(x.numpy(), t.numpy(), c)
        assert (int(cnt[0]), int(cnt[1])) == (case["correct"], case["labeled"]), args
        inter = cnt[2:2 + c].astype(np.float32)
        union = cnt[2 + c:2 + 2 * c].astype(np.float32) + cnt[2 + 2 * c:].astype(np.float32) - inter
        assert np.array_equal(inter, case["inter"].numpy()) and np.array_equal(union, case["union"].numpy()), args
        if i < 6:
            m.update_counts(cnt)
    acc = fx["accumulated"]
    assert (m.total_correct, m.total_label) == (acc["total_correct"], acc["total_label"])
    assert np.array_equal(m.total_inter, acc["total_inter"].numpy()) and np.array_equal(m.total_union, acc["total_union"].numpy())
    pix_acc, miou, iou = m.get()
    assert pix_acc == acc["pixAcc"] and np.array_equal(iou, acc["IoU"].numpy())
    assert abs(miou - acc["mIoU"]) <= 1e-7


def test_evaluate_oracle_matches_reference_fixture():
    """oracle/evalref.py against the scores the reference's own SegBaseModel.evaluate (segbase.py:44-79) produced for six
    (scales, flip, crop) configurations of a seeded stub model (tests/golden/make_eval_golden.py): same torch ops in the same
    order, so bit-identical on the machine that wrote the fixture (1e-6 elsewhere: CPU conv kernels differ by ISA)."""
    import importlib.util
    from oracle import evalref as E
    spec = importlib.util.spec_from_file_location("make_eval_golden", os.path.join(G, "make_eval_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    fx = torch.load(os.path.join(G, "evaluate_cases.pt"))
    assert len(fx) == len(gen.CASES)
    for case, args in zip(fx, gen.CASES):
        seed, b, h, w, scales, flip, crop = args
        with torch.no_grad():
            got = E.evaluate(gen.stub_forward(seed), gen.make_image(seed, b, h, w), scales, flip, crop)
        ref = case["scores"]
        assert got.shape == ref.shape
        assert float((got - ref).abs().max()) <= 1e-6 * float(ref.abs().max()), args
    assert E.scaled_size(1024, 2048, 0.75) == (768, 1536) and E.scaled_size(57, 31, 1.5) == (86, 47)
    assert E.padded_size(28, 40, (48, 64), 0.75) == (36, 48)
    assert E.padded_size(29, 16, (64, 64), 0.5) == (45, 19)            # the reference's swapped F.pad amounts (segbase.py:93)

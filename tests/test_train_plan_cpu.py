"""Host logic of the training launch list, checked WITHOUT a GPU: the plan is built over CPU fp32 buffers (dry run) and
interpreted by tests/emulate_plan.py (torch restatement of each kernel's documented semantics); loss, every parameter
gradient, BatchNorm running statistics and the SGD update must match the oracle's training step (which is pinned to the real
reference by tests/golden/train_dlv3p_resnet101_65x97_b4.pt)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
import emulate_plan as E  # noqa: E402
from oracle import segref as R  # noqa: E402


@pytest.fixture()
def dry_run():
    """plan built over CPU buffers; fp64 activations AND statistics so that rounding noise cannot flip ReLU masks (train-mode
    BatchNorm through 100 layers amplifies a 1e-7 perturbation ~1e4-fold with these synthetic weights)"""
    from segmentron_b200 import ops
    from segmentron_b200.train import TrainPlan
    ops._PLAN_DRY_RUN = True
    TrainPlan.STAT_DTYPE = torch.float64
    yield
    ops._PLAN_DRY_RUN = False
    TrainPlan.STAT_DTYPE = torch.float32


@pytest.mark.parametrize("model,backbone", [("deeplabv3plus_resnet101", "resnet101"), ("deeplabv3plus_xception65", "xception65"),
                                            ("deeplabv3plus_mobilenet_v2", "mobilenet_v2")])
def test_train_plan_matches_oracle_on_cpu(dry_run, model, backbone):
    from segmentron_b200.train import DeepLabV3PlusTrainerB200
    seed, shape = 21, (4, 3, 65, 97)
    P = R.build_params(model, seed)
    g = torch.Generator().manual_seed(2000 + seed)
    x = torch.randn(*shape, generator=g)
    target = torch.randint(-1, 19, (shape[0], shape[2], shape[3]), generator=g)
    torch.manual_seed(777)
    mask = torch.empty(shape[0], 256, 1, 1).bernoulli_(0.9) / 0.9
    tr = DeepLabV3PlusTrainerB200(P.state_dict(), backbone=backbone, dtype=torch.float64, device="cpu", lr=0.02)
    loss = E.forward_backward(tr, x, target, {"head.aspp.dropout": mask})
    grads = tr.store.grads()
    sd_mid = tr.state_dict()
    P.dropout_masks["head.aspp.dropout"] = mask
    before = {k: v.clone() for k, v in P.t.items()}
    P64 = P.to(dtype=torch.float64)
    P64.frozen, P64.dropout_masks = True, P.dropout_masks
    o_loss, o_grads, _, _ = R.loss_and_grads(model, P64, x.double(), target)
    assert abs(float(loss) - float(o_loss)) < 1e-6 * abs(float(o_loss)), (float(loss), float(o_loss))
    worst = ("", 0.0)
    # some gradients are analytically ZERO (a BatchNorm bias followed by conv + train-mode BatchNorm: Xception's bn_depth.bias)
    # and come out as ~1e-14 rounding noise on both sides: errors are measured against |g_ref| + 1e-6 * (largest gradient norm)
    floor = 1e-6 * max(float(v.norm()) for v in o_grads.values())
    for k, gr in o_grads.items():
        e = float((grads[k] - gr).norm() / (gr.norm() + floor))
        if e > worst[1]:
            worst = (k, e)
    assert worst[1] < 1e-6, worst                  # fp32 gradient buffers: ~3e-8 observed
    for k in sd_mid:
        if k.endswith(("running_mean", "running_var")):
            assert torch.allclose(sd_mid[k].double(), P64.t[k], atol=1e-6, rtol=1e-6), k
    # SGD: encoder lr, decoder lr x10, weight decay on everything (solver/optimizer.py:14-34,50-51)
    E.sgd(tr)
    sd = tr.state_dict()
    probes = {"resnet101": ("encoder.conv1.weight", "encoder.layer3.5.conv2.weight", "head.aspp.bn.bias"),
              "xception65": ("encoder.conv1.weight", "encoder.block7.sep_conv2.block.depthwise.weight", "head.aspp.bn.bias"),
              "mobilenet_v2": ("encoder.conv1.conv.weight", "encoder.block4.2.conv.1.conv.weight", "head.block.1.block.bn_point.bias")}
    for k in probes[backbone] + ("head.block.2.weight",):
        lr = 0.02 * (10.0 if k.startswith("head.") else 1.0)
        ref = before[k].double() - lr * (o_grads[k] + 1e-4 * before[k].double())
        assert float((sd[k].double() - ref).norm() / ref.norm()) < 1e-6, k
    # bucket plan: contiguous, covers the whole gradient, launch positions non-decreasing
    bk = tr.plan_for(shape)["buckets"]
    assert bk[0][2] == tr.store.total and bk[-1][1] == 0
    assert all(a[1] == b[2] for a, b in zip(bk, bk[1:])) and all(a[0] <= b[0] for a, b in zip(bk, bk[1:]))
    pl = tr.plan_for(shape)["plan"]
    for pos, lo, hi in bk:                              # nothing after `pos` may touch the gradient range of the bucket
        for st in pl.bwd[pos:]:
            for key in ("dw", "dgamma", "dbeta", "out"):
                t = st.info.get(key)
                if t is not None and t.untyped_storage().data_ptr() == tr.store.grad.untyped_storage().data_ptr():
                    off = t.storage_offset()
                    assert not (lo <= off < hi), (st.kind, off, lo, hi)
            if st.kind == "scatter_add":               # the stem's s2d-space gradient is un-packed into encoder.conv1.weight
                assert not (lo <= tr.store.meta[tr.store.stem]["off"] < hi), "stem gradient written after its bucket"


def test_ccnet_train_plan_matches_oracle_on_cpu(dry_run):
    """CCNet / ResNet101 (models/ccnet.py): criss-cross attention forward + backward steps in the launch list, recurrence 2 with
    SHARED weights (gradients of both applications meet in the same slots), c4 produced into -- and its gradient accumulated
    from -- a channel slice of the concat buffer, Dropout2d after a ReLU-less BatchNorm."""
    from segmentron_b200.train import CCNetTrainerB200
    seed, shape = 31, (2, 3, 65, 97)
    P = R.build_params("ccnet_resnet101", seed)
    g = torch.Generator().manual_seed(3000 + seed)
    x = torch.randn(*shape, generator=g)
    target = torch.randint(-1, 19, (shape[0], shape[2], shape[3]), generator=g)
    mask = (torch.rand(shape[0], 512, 1, 1, generator=g) > 0.1).double() / 0.9
    tr = CCNetTrainerB200(P.state_dict(), dtype=torch.float64, device="cpu", lr=0.02)
    loss = E.forward_backward(tr, x, target, {"head.rcca.bottleneck.dropout": mask})
    grads = tr.store.grads()
    P64 = P.to(dtype=torch.float64)
    P64.frozen, P64.dropout_masks = True, {"head.rcca.bottleneck.dropout": mask}
    o_loss, o_grads, _, _ = R.loss_and_grads("ccnet_resnet101", P64, x.double(), target)
    assert abs(float(loss) - float(o_loss)) < 1e-6 * abs(float(o_loss))
    assert float(o_grads["head.rcca.cca.gamma"].abs()) > 1e-3 and float(P.t["head.rcca.cca.gamma"]) != 0.0      # non-vacuous
    floor = 1e-6 * max(float(v.norm()) for v in o_grads.values())
    worst = max(((float((grads[k] - gr).norm() / (gr.norm() + floor)), k) for k, gr in o_grads.items()))
    assert worst[0] < 1e-6, worst
    kinds = {s.kind for s in tr.plan_for(shape)["plan"].bwd}
    assert {"cca_weight_bwd", "cca_gather", "cca_scatter"} <= kinds


def test_state_dict_roundtrip(dry_run):
    from segmentron_b200.train import DeepLabV3PlusTrainerB200
    P = R.build_params("deeplabv3plus_resnet101", 3)
    sd0 = P.state_dict()
    tr = DeepLabV3PlusTrainerB200(sd0, dtype=torch.float64, device="cpu")
    sd = tr.state_dict()
    assert set(sd) == set(sd0)
    for k in sd0:
        assert tuple(sd[k].shape) == tuple(sd0[k].shape), k
        assert torch.equal(sd[k].float(), sd0[k].float()), k


def test_hrnet_train_plan_matches_oracle_on_cpu(dry_run):
    """HRNet-w18-small-v1 (backbones/hrnet.py + models/hrnet_seg.py): stride-4 stem of two 3x3/2 convs, Bottleneck + BasicBlock
    branches, transition layers, the fuse sums (identity + stride-2 conv chains through the BatchNorm residual operand + 1x1 /
    BatchNorm / nearest-up terms through upsample_add, one ReLU at the end), the head's align_corners=False resizes into concat
    slices, a conv bias in front of BatchNorm, BatchNorm momentum 0.01, no decoder LR factor (the YAML's recipe)."""
    from segmentron_b200.train import HRNetTrainerB200
    seed, shape = 41, (2, 3, 64, 96)
    P = R.build_params("hrnet_w18_small_v1", seed)
    g = torch.Generator().manual_seed(4000 + seed)
    x = torch.randn(*shape, generator=g)
    target = torch.randint(-1, 19, (shape[0], shape[2], shape[3]), generator=g)
    tr = HRNetTrainerB200(P.state_dict(), dtype=torch.float64, device="cpu", lr=0.01)
    loss = E.forward_backward(tr, x, target, {})
    grads = tr.store.grads()
    sd_mid = tr.state_dict()
    before = {k: v.clone() for k, v in P.t.items()}
    P64 = P.to(dtype=torch.float64)
    P64.frozen, P64.bn_momentum = True, 0.01
    o_loss, o_grads, _, _ = R.loss_and_grads("hrnet_w18_small_v1", P64, x.double(), target)
    assert abs(float(loss) - float(o_loss)) < 1e-6 * abs(float(o_loss)), (float(loss), float(o_loss))
    floor = 1e-6 * max(float(v.norm()) for v in o_grads.values())
    assert set(grads) == set(o_grads)
    worst = max(((float((grads[k] - gr).norm() / (gr.norm() + floor)), k) for k, gr in o_grads.items()))
    assert worst[0] < 1e-6, worst
    for k in sd_mid:
        if k.endswith(("running_mean", "running_var")):
            assert torch.allclose(sd_mid[k].double(), P64.t[k], atol=1e-6, rtol=1e-6), k
    E.sgd(tr)
    sd = tr.state_dict()
    for k in ("encoder.conv1.weight", "encoder.stage3.0.fuse_layers.2.0.1.0.weight", "hrnet_head.last_layer.0.bias",
              "hrnet_head.last_layer.3.weight"):
        ref = before[k].double() - 0.01 * (o_grads[k] + 1e-4 * before[k].double())          # one LR for encoder and head
        assert float((sd[k].double() - ref).norm() / (ref.norm() + 1e-12)) < 1e-6, k
    kinds = {s.kind for s in tr.plan_for(shape)["plan"].fwd + tr.plan_for(shape)["plan"].bwd}
    assert {"upsample_add", "upsample_add_bwd"} <= kinds


def test_danet_train_plan_matches_oracle_on_cpu(dry_run):
    """DANet / ResNet101 (models/danet.py; OS8 with the multi-grid dilations 4/8/16 in layer4): position attention with the
    attention matrix materialised per image (Q K^T, row softmax, P V and the four backward GEMMs over transposed copies), channel
    attention (Gram matrix, softmax(rowmax - E), its backward through dE + dE^T), the gamma residuals and their gradients, three
    Dropout2d + classifier heads on sa, sc and sa + sc, and the summed cross-entropy of the three outputs."""
    from segmentron_b200.train import DANetTrainerB200
    seed, shape = 51, (2, 3, 64, 96)                         # 8 x 12 = 96 positions per image
    P = R.build_params("danet_resnet101", seed)
    g = torch.Generator().manual_seed(5000 + seed)
    x = torch.randn(*shape, generator=g)
    target = torch.randint(-1, 19, (shape[0], shape[2], shape[3]), generator=g)
    masks = {f"head.conv{j}.0": (torch.rand(shape[0], 512, 1, 1, generator=g) > 0.1).double() / 0.9 for j in (6, 7, 8)}
    tr = DANetTrainerB200(P.state_dict(), dtype=torch.float64, device="cpu", lr=0.02)
    loss = E.forward_backward(tr, x, target, masks)
    grads = tr.store.grads()
    sd_mid = tr.state_dict()
    P64 = P.to(dtype=torch.float64)
    P64.frozen, P64.dropout_masks = True, dict(masks)
    o_loss, o_grads, _, _ = R.loss_and_grads("danet_resnet101", P64, x.double(), target)
    assert abs(float(loss) - float(o_loss)) < 1e-6 * abs(float(o_loss)), (float(loss), float(o_loss))
    assert float(P.t["head.sa.gamma"]) != 0.0 and float(P.t["head.sc.gamma"]) != 0.0                  # non-vacuous attention
    floor = 1e-6 * max(float(v.norm()) for v in o_grads.values())
    assert set(grads) == set(o_grads)
    worst = max(((float((grads[k] - gr).norm() / (gr.norm() + floor)), k) for k, gr in o_grads.items()))
    assert worst[0] < 1e-6, worst
    for k in sd_mid:
        if k.endswith(("running_mean", "running_var")):
            assert torch.allclose(sd_mid[k].double(), P64.t[k], atol=1e-6, rtol=1e-6), k
    kinds = {s.kind for s in tr.plan_for(shape)["plan"].fwd + tr.plan_for(shape)["plan"].bwd}
    assert {"row_softmax", "row_softmax_bwd", "cam_softmax", "cam_softmax_bwd", "cam_bwd_pack", "transpose", "upsample_add_bwd"} <= kinds


def test_every_plan_step_matches_its_c_signature(dry_run, monkeypatch):
    """Arity / type-position check of EVERY kernel call the plans record (all six model families): the argument tuple handed to
    TrainPlan.add must line up with the ctypes signature declared from include/segb200.h (pointers where pointers go, ints where
    ints go, floats where floats go) -- a mistake here would otherwise only surface as a TypeError or a wild pointer on the GPU."""
    import ctypes as C
    from segmentron_b200 import train as T
    seen = {}
    orig_add = T.TrainPlan.add

    def checked_add(self, kind, fn, args, **info):
        sig = fn.argtypes
        assert sig is not None and len(args) + 1 == len(sig), (kind, fn.__name__, len(args), len(sig) if sig else None)
        for pos, (a, t) in enumerate(zip(args, sig)):
            if t is C.c_void_p:
                assert a is None or isinstance(a, C.c_void_p), (kind, fn.__name__, pos, type(a))
            elif t in (C.c_int, C.c_longlong):
                assert isinstance(a, int) and not isinstance(a, bool), (kind, fn.__name__, pos, a)
            elif t in (C.c_float, C.c_double):
                assert isinstance(a, (int, float)) and not isinstance(a, bool), (kind, fn.__name__, pos, a)
        seen[fn.__name__] = seen.get(fn.__name__, 0) + 1
        return orig_add(self, kind, fn, args, **info)
    monkeypatch.setattr(T.TrainPlan, "add", checked_add)
    cases = [(T.DeepLabV3PlusTrainerB200, "deeplabv3plus_resnet101", dict(backbone="resnet101"), (2, 3, 65, 97)),
             (T.DeepLabV3PlusTrainerB200, "deeplabv3plus_xception65", dict(backbone="xception65"), (2, 3, 65, 97)),
             (T.DeepLabV3PlusTrainerB200, "deeplabv3plus_mobilenet_v2", dict(backbone="mobilenet_v2"), (2, 3, 64, 96)),
             (T.CCNetTrainerB200, "ccnet_resnet101", {}, (2, 3, 65, 97)), (T.HRNetTrainerB200, "hrnet_w18_small_v1", {}, (2, 3, 64, 96)),
             (T.DANetTrainerB200, "danet_resnet101", {}, (2, 3, 64, 96))]
    for cls, model, kw, shape in cases:
        tr = cls(R.build_params(model, 1).state_dict(), dtype=torch.float64, device="cpu", **kw)
        tr.plan_for(shape)
    assert {"segb200_upsample_add_bwd", "segb200_row_softmax_bwd", "segb200_cam_bwd_pack", "segb200_cca_weight_bwd",
            "segb200_upsample_ce", "segb200_dw_wgrad"} <= set(seen), sorted(seen)

"""CPU/torch ORACLE for the SegmenTron dense hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (plain ``torch.nn.functional`` calls over a flat
parameter dict) of the reference's forward pass for the hot path of SURVEY.md section 8.
It is NOT the product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the checker or the
timed CPU baseline.  Nothing under ``segmentron_b200/`` imports it.

Pinning: the reference ships no golden vectors.  The oracle is pinned by
``tests/golden/make_golden.py``, which imports the real reference from /root/reference in the
build container, loads the SAME parameter dict into the reference ``nn.Module`` tree
(``load_state_dict(strict=True)`` => names and shapes are the reference's) and checks that
reference and oracle agree, then commits the reference's outputs as fixtures under
``tests/golden/``.  ``tests/test_oracle_golden.py`` re-checks the oracle against those
fixtures everywhere (no /root/reference needed).

Parameter dicts use the reference's ``state_dict`` key names.  A ``Params`` object creates
missing entries lazily from a seeded CPU generator, so running a forward once on a tiny input
yields a complete, reproducible state dict.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------
# parameter store
# ----------------------------------------------------------------------------------------
class Params:
    """Flat ``name -> tensor`` store with lazy, seeded creation.

    BatchNorm buffers are randomised (NOT the identity default of nn.BatchNorm2d) so that a
    folded-BN bug cannot hide (SURVEY.md 8c "non-vacuous oracle").  Gains are chosen so the
    activation scale neither collapses nor explodes through ~70 layers.
    """

    def __init__(self, seed: int = 0, tensors: Optional[Dict[str, torch.Tensor]] = None,
                 device: str = "cpu", dtype: torch.dtype = torch.float32):
        self.t: Dict[str, torch.Tensor] = dict(tensors) if tensors else {}
        self.g = torch.Generator(device="cpu")
        self.g.manual_seed(seed)
        self.device, self.dtype = device, dtype
        self.frozen = tensors is not None
        # training-mode switches (tools/train.py:133 `self.model.train()`): batch-statistics BatchNorm with running-stat
        # updates, Dropout2d active.  ``dropout_masks[name]`` = explicit [N,C,1,1] keep-mask already scaled by 1/(1-p)
        # (so that the engine under test and the oracle can share one mask); absent -> F.dropout2d with the torch RNG.
        self.training = False
        self.bn_momentum = 0.1
        self.dropout_masks: Dict[str, torch.Tensor] = {}

    # -- creation helpers (always drawn on CPU in fp32, then moved) ----------------------
    def _new(self, name, make):
        if name not in self.t:
            if self.frozen:
                raise KeyError(f"oracle parameter '{name}' missing from the supplied state dict")
            self.t[name] = make().to(torch.float32)
        v = self.t[name]
        if name.endswith("num_batches_tracked"):
            return v
        if v.device != torch.device(self.device) or v.dtype != self.dtype:
            v = v.to(device=self.device, dtype=self.dtype)
        return v

    def conv_w(self, name, cout, cin_g, k, gain=1.0):
        fan_in = cin_g * k * k
        std = gain / math.sqrt(fan_in)
        return self._new(name, lambda: torch.randn(cout, cin_g, k, k, generator=self.g) * std)

    def vec(self, name, n, kind):
        def make():
            if kind == "gamma":
                return 0.75 + 0.5 * torch.rand(n, generator=self.g)
            if kind == "var":
                return 0.6 + 0.8 * torch.rand(n, generator=self.g)
            if kind in ("beta", "mean", "bias"):
                return 0.1 * torch.randn(n, generator=self.g)
            raise ValueError(kind)
        return self._new(name, make)

    def scalar(self, name, value):
        return self._new(name, lambda: torch.full((1,), float(value)))

    def count(self, name):
        return self._new(name, lambda: torch.zeros(()))

    def state_dict(self) -> Dict[str, torch.Tensor]:
        out = {}
        for k, v in self.t.items():
            out[k] = v.long() if k.endswith("num_batches_tracked") else v
        return out

    def to(self, device=None, dtype=None) -> "Params":
        """Materialised copy on ``device`` / in ``dtype`` (like ``model.to(...)``: converted once, not per use)."""
        device, dtype = device or self.device, dtype or self.dtype
        conv = {k: (v.to(device=device, dtype=dtype) if not k.endswith("num_batches_tracked") else v.to(device))
                for k, v in self.t.items()}
        p = Params(tensors=conv)
        p.device, p.dtype = device, dtype
        return p


# ----------------------------------------------------------------------------------------
# leaf ops  (Appendix D semantics)
# ----------------------------------------------------------------------------------------
def conv2d(P: Params, x, name, cout, k=1, stride=1, padding=0, dilation=1, groups=1,
           bias=False, gain=1.0):
    """nn.Conv2d, OIHW cross-correlation, zero padding (torch semantics)."""
    cin = x.shape[1]
    w = P.conv_w(name + ".weight", cout, cin // groups, k, gain)
    b = P.vec(name + ".bias", cout, "bias") if bias else None
    return F.conv2d(x, w, b, stride, padding, dilation, groups)


def batchnorm(P: Params, x, name, eps=1e-5):
    """nn.BatchNorm2d in eval mode: (x-mean)/sqrt(var+eps)*weight+bias  (batch_norm.py:126)."""
    c = x.shape[1]
    g = P.vec(name + ".weight", c, "gamma")
    b = P.vec(name + ".bias", c, "beta")
    m = P.vec(name + ".running_mean", c, "mean")
    v = P.vec(name + ".running_var", c, "var")
    P.count(name + ".num_batches_tracked")
    if P.training:
        # F.batch_norm(training=True): biased batch variance for the output, running stats updated in place with
        # momentum (running_var with the unbiased variance) -- SURVEY.md appendix D, torch semantics
        if name + ".num_batches_tracked" in P.t:
            P.t[name + ".num_batches_tracked"] = P.t[name + ".num_batches_tracked"] + 1
        return F.batch_norm(x, m, v, g, b, True, P.bn_momentum, eps)
    return F.batch_norm(x, m, v, g, b, False, 0.0, eps)


def dropout2d(P: Params, x, name, p=0.1):
    """nn.Dropout2d(p): identity in eval; in training zeroes whole (sample, channel) planes and scales by 1/(1-p)
    (modules/module.py:60,75)."""
    if not P.training:
        return x
    if name in P.dropout_masks:
        return x * P.dropout_masks[name].to(device=x.device, dtype=x.dtype)
    return F.dropout2d(x, p, True)


def conv_bn_act(P, x, prefix, cout, k, stride=1, padding=0, dilation=1, groups=1, act="relu",
                eps=1e-5, conv="conv", bn="bn"):
    """_ConvBNReLU / _ConvBN  (modules/basic.py:65-77, :95-105)."""
    gain = math.sqrt(2.0)
    x = conv2d(P, x, f"{prefix}.{conv}" if conv else prefix, cout, k, stride, padding, dilation,
               groups, gain=gain)
    x = batchnorm(P, x, f"{prefix}.{bn}", eps)
    if act == "relu":
        x = F.relu(x)
    elif act == "relu6":
        x = F.relu6(x)
    return x


def separable_conv2d(P, x, prefix, planes, stride=1, dilation=1, relu_first=True, eps=1e-5,
                     pw_gain=1.0):
    """SeparableConv2d (modules/basic.py:34-62).

    relu_first=True : ReLU -> dw3x3(pad=dil) -> BN -> pw1x1 -> BN
    relu_first=False: dw3x3 -> BN -> ReLU -> pw1x1 -> BN -> ReLU
    """
    c = x.shape[1]
    p = prefix + ".block"
    if relu_first:
        x = F.relu(x)                                   # not in place: input re-used by the skip
    x = conv2d(P, x, p + ".depthwise", c, 3, stride, dilation, dilation, groups=c,
               gain=math.sqrt(2.0) if relu_first else 1.0)
    x = batchnorm(P, x, p + ".bn_depth", eps)
    if not relu_first:
        x = F.relu(x)
    x = conv2d(P, x, p + ".pointwise", planes, 1,
               gain=pw_gain * (math.sqrt(2.0) if not relu_first else 1.0))
    x = batchnorm(P, x, p + ".bn_point", eps)
    if not relu_first:
        x = F.relu(x)
    return x


# ----------------------------------------------------------------------------------------
# Xception65  (models/backbones/xception.py)
# ----------------------------------------------------------------------------------------
def xception_block(P, x, prefix, chans, stride=1, dilation=1, skip="conv", relu_first=True,
                   low_feat=False, eps=1e-5):
    """XceptionBlock (xception.py:10-51): 3 separable convs + {conv|sum|none} skip, no ReLU
    after the add."""
    sc1 = separable_conv2d(P, x, prefix + ".sep_conv1", chans[1], 1, dilation, relu_first, eps)
    sc2 = separable_conv2d(P, sc1, prefix + ".sep_conv2", chans[2], 1, dilation, relu_first, eps)
    # (synthetic-weight gain only: keeps the 16 'sum' blocks from doubling the variance each)
    res = separable_conv2d(P, sc2, prefix + ".sep_conv3", chans[3], stride, dilation, relu_first, eps,
                           pw_gain=0.35 if skip == "sum" else 1.0)
    if skip == "conv":
        s = conv2d(P, x, prefix + ".conv", chans[3], 1, stride)       # xception.py:21,38
        s = batchnorm(P, s, prefix + ".bn", eps)                      # :22,39
        out = res + s                                                 # :40
    elif skip == "sum":
        out = res + x                                                 # :42
    else:
        out = res                                                     # :44
    return (out, sc2) if low_feat else out


def xception65(P, x, prefix="encoder", output_stride=16, eps=1e-5):
    """Xception65.forward (xception.py:129-165); stride/dilation table :57-74."""
    if output_stride == 32:
        b3s, mid_d, exit_d, exit_s = 2, 1, (1, 1), 2
    elif output_stride == 16:
        b3s, mid_d, exit_d, exit_s = 2, 1, (1, 2), 1
    elif output_stride == 8:
        b3s, mid_d, exit_d, exit_s = 1, 2, (2, 4), 1
    else:
        raise NotImplementedError(output_stride)
    p = prefix
    x = conv2d(P, x, p + ".conv1", 32, 3, 2, 1, gain=math.sqrt(2.0))  # :131
    x = F.relu(batchnorm(P, x, p + ".bn1", eps))
    x = conv2d(P, x, p + ".conv2", 64, 3, 1, 1, gain=math.sqrt(2.0))  # :135
    x = F.relu(batchnorm(P, x, p + ".bn2", eps))
    x = xception_block(P, x, p + ".block1", [64, 128, 128, 128], 2, eps=eps)
    x, c1 = xception_block(P, x, p + ".block2", [128, 256, 256, 256], 2, low_feat=True, eps=eps)
    x, c2 = xception_block(P, x, p + ".block3", [256, 728, 728, 728], b3s, low_feat=True, eps=eps)
    for i in range(4, 20):                                            # middle flow :144-159
        x = xception_block(P, x, f"{p}.block{i}", [728] * 4, 1, mid_d, skip="sum", eps=eps)
    c3 = x
    x = xception_block(P, c3, p + ".block20", [728, 728, 1024, 1024], exit_s, exit_d[0], eps=eps)
    c4 = xception_block(P, x, p + ".block21", [1024, 1536, 1536, 2048], 1, exit_d[1],
                        skip="none", relu_first=False, eps=eps)
    return c1, c2, c3, c4


# ----------------------------------------------------------------------------------------
# MobileNetV2  (models/backbones/mobilenet.py:55-143, modules/basic.py:139-163)
# ----------------------------------------------------------------------------------------
def inverted_residual(P, x, prefix, cout, stride, expand, dilation=1, eps=1e-5):
    """InvertedResidual (basic.py:139-163): [pw+BN+ReLU6] -> dw(stride,dil)+BN+ReLU6 -> pw+BN."""
    cin = x.shape[1]
    inter = int(round(cin * expand))
    y, i = x, 0
    if expand != 1:
        y = conv_bn_act(P, y, f"{prefix}.conv.{i}", inter, 1, act="relu6", eps=eps)
        i += 1
    y = conv_bn_act(P, y, f"{prefix}.conv.{i}", inter, 3, stride, dilation, dilation, groups=inter,
                    act="relu6", eps=eps)
    i += 1
    y = conv2d(P, y, f"{prefix}.conv.{i}", cout, 1)
    y = batchnorm(P, y, f"{prefix}.conv.{i + 1}", eps)
    return x + y if (stride == 1 and cin == cout) else y


def mobilenet_v2(P, x, prefix="encoder", output_stride=16, eps=1e-5):
    """MobileNetV2.forward (mobilenet.py:131-143).  Quirk kept: inside a dilated stage only the
    FIRST block of each (t,c,n,s) group gets the dilation (mobilenet.py:125 vs :128)."""
    dil = {32: (1, 1), 16: (1, 2), 8: (2, 4)}[output_stride]
    setting = [[1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1],
               [6, 160, 3, 2], [6, 320, 1, 1]]
    x = conv_bn_act(P, x, prefix + ".conv1", 32, 3, 2, 1, act="relu6", eps=eps)

    def layer(x, name, rows, dilation=1):
        j = 0
        for t, c, n, s in rows:
            stride = s if dilation == 1 else 1
            x = inverted_residual(P, x, f"{prefix}.{name}.{j}", c, stride, t, dilation, eps); j += 1
            for _ in range(n - 1):
                x = inverted_residual(P, x, f"{prefix}.{name}.{j}", c, 1, t, 1, eps); j += 1
        return x

    x = layer(x, "block1", setting[0:1])
    c1 = layer(x, "block2", setting[1:2])
    c2 = layer(c1, "block3", setting[2:3])
    c3 = layer(c2, "block4", setting[3:5], dil[0])
    c4 = layer(c3, "block5", setting[5:], dil[1])
    return c1, c2, c3, c4



# ----------------------------------------------------------------------------------------
# ResNetV1 / BottleneckV1b  (models/backbones/resnet.py:44-199)
# ----------------------------------------------------------------------------------------
def bottleneck_v1b(P, x, prefix, planes, stride=1, dilation=1, downsample=False, eps=1e-5, last_gain=0.5):
    """BottleneckV1b.forward (resnet.py:60-81): 1x1 -> 3x3(stride, pad=dil, dil) -> 1x1, + identity, ReLU."""
    g = math.sqrt(2.0)
    out = F.relu(batchnorm(P, conv2d(P, x, prefix + ".conv1", planes, 1, gain=g), prefix + ".bn1", eps))
    out = F.relu(batchnorm(P, conv2d(P, out, prefix + ".conv2", planes, 3, stride, dilation, dilation, gain=g),
                           prefix + ".bn2", eps))
    out = batchnorm(P, conv2d(P, out, prefix + ".conv3", planes * 4, 1, gain=g * last_gain), prefix + ".bn3", eps)
    identity = x
    if downsample:                                                    # resnet.py:142-147
        identity = batchnorm(P, conv2d(P, x, prefix + ".downsample.0", planes * 4, 1, stride, gain=g),
                             prefix + ".downsample.1", eps)
    return F.relu(out + identity)                                     # :78-79


def resnet_v1(P, x, prefix="encoder", layers=(3, 4, 23, 3), output_stride=16, multi_grid=False,
              multi_dilation=None, eps=1e-5):
    """ResNetV1.forward (resnet.py:183-199), non-deep-stem (resnet50/101/152).  Dilation/stride table :90-100,
    first-block dilation rule and DANet multi-grid :149-179."""
    dil, strides = {32: ((1, 1), (2, 2)), 16: ((1, 2), (2, 1)), 8: ((2, 4), (1, 1))}[output_stride]
    x = conv2d(P, x, prefix + ".conv1", 64, 7, 2, 3, gain=math.sqrt(2.0))        # :116
    x = F.relu(batchnorm(P, x, prefix + ".bn1", eps))
    x = F.max_pool2d(x, 3, 2, 1)                                                  # :119
    inplanes = 64

    def make_layer(x, name, planes, blocks, stride=1, dilation=1, mg=False):
        nonlocal inplanes
        ds = stride != 1 or inplanes != planes * 4
        if not mg:
            first_d = 1 if dilation in (1, 2) else 2                               # :151-159
        else:
            first_d = multi_dilation[0]                                            # :161
        x = bottleneck_v1b(P, x, f"{prefix}.{name}.0", planes, stride, first_d, ds, eps)
        inplanes = planes * 4
        for i in range(1, blocks):
            d = multi_dilation[i % len(multi_dilation)] if mg else dilation        # :166-175
            x = bottleneck_v1b(P, x, f"{prefix}.{name}.{i}", planes, 1, d, False, eps)
        return x

    c1 = make_layer(x, "layer1", 64, layers[0])
    c2 = make_layer(c1, "layer2", 128, layers[1], 2)
    c3 = make_layer(c2, "layer3", 256, layers[2], strides[0], dil[0])
    c4 = make_layer(c3, "layer4", 512, layers[3], strides[1], dil[1], multi_grid)
    # the (unused) classifier exists in the reference state_dict: resnet.py:129-131
    P._new(prefix + ".fc.weight", lambda: torch.zeros(1000, 2048))
    P._new(prefix + ".fc.bias", lambda: torch.zeros(1000))
    return c1, c2, c3, c4


# ----------------------------------------------------------------------------------------
# HRNet  (models/backbones/hrnet.py, models/hrnet_seg.py)
# ----------------------------------------------------------------------------------------
HRNET_W18_SMALL_V1 = dict(                     # configs/cityscapes_hrnet_w18_small_v1.yaml:23-64
    stage1=dict(block="BOTTLENECK", blocks=[1], channels=[32]),
    stage2=dict(modules=1, block="BASIC", blocks=[2, 2], channels=[16, 32]),
    stage3=dict(modules=1, block="BASIC", blocks=[2, 2, 2], channels=[16, 32, 64]),
    stage4=dict(modules=1, block="BASIC", blocks=[2, 2, 2, 2], channels=[16, 32, 64, 128]),
    final_conv_kernel=1)


def hr_basic_block(P, x, prefix, planes):
    """BasicBlock.forward (hrnet.py:38-55): conv3x3-BN-ReLU-conv3x3-BN, + x, ReLU (no downsample inside HR modules)."""
    g = math.sqrt(2.0)
    out = F.relu(batchnorm(P, conv2d(P, x, prefix + ".conv1", planes, 3, 1, 1, gain=g), prefix + ".bn1"))
    out = batchnorm(P, conv2d(P, out, prefix + ".conv2", planes, 3, 1, 1, gain=g * 0.5), prefix + ".bn2")
    return F.relu(out + x)


def hr_module(P, xs, prefix, blocks, channels):
    """HighResolutionModule.forward (hrnet.py:215-232) with the fuse layers of _make_fuse_layers (:165-209)."""
    nb = len(xs)
    xs = list(xs)
    for i in range(nb):
        for b in range(blocks[i]):
            xs[i] = hr_basic_block(P, xs[i], f"{prefix}.branches.{i}.{b}", channels[i])
    if nb == 1:
        return xs
    outs = []
    for i in range(nb):
        y = None
        for j in range(nb):
            if j == i:
                t = xs[j]
            elif j > i:                                                               # 1x1 conv + BN + nearest up 2^(j-i)
                t = batchnorm(P, conv2d(P, xs[j], f"{prefix}.fuse_layers.{i}.{j}.0", channels[i], 1),
                              f"{prefix}.fuse_layers.{i}.{j}.1")
                t = F.interpolate(t, scale_factor=2 ** (j - i), mode="nearest")
            else:                                                                     # chain of 3x3 stride-2 convs
                t = xs[j]
                for k in range(i - j):
                    last = k == i - j - 1
                    co = channels[i] if last else channels[j]
                    t = batchnorm(P, conv2d(P, t, f"{prefix}.fuse_layers.{i}.{j}.{k}.0", co, 3, 2, 1, gain=math.sqrt(2.0)),
                                  f"{prefix}.fuse_layers.{i}.{j}.{k}.1")
                    if not last:
                        t = F.relu(t)
            y = t if y is None else y + t
        outs.append(F.relu(y))
    return outs


def hrnet_backbone(P, x, prefix="encoder", hcfg=None):
    """HighResolutionNet.forward (hrnet.py:429-479)."""
    hcfg = hcfg or HRNET_W18_SMALL_V1
    g = math.sqrt(2.0)
    x = F.relu(batchnorm(P, conv2d(P, x, prefix + ".conv1", 64, 3, 2, 1, gain=g), prefix + ".bn1"))
    x = F.relu(batchnorm(P, conv2d(P, x, prefix + ".conv2", 64, 3, 2, 1, gain=g), prefix + ".bn2"))
    # layer1: Bottleneck blocks (hrnet.py:58-94, _make_layer :383-399)
    planes = hcfg["stage1"]["channels"][0]
    inpl = 64
    for b in range(hcfg["stage1"]["blocks"][0]):
        ds = inpl != planes * 4
        x = bottleneck_v1b(P, x, f"{prefix}.layer1.{b}", planes, 1, 1, ds)    # same structure/keys as ResNet's bottleneck
        inpl = planes * 4
    pre = [inpl]
    ys = [x]
    for si, sname in enumerate(("stage2", "stage3", "stage4")):
        sc = hcfg[sname]
        cur = sc["channels"]
        tname = f"{prefix}.transition{si + 1}"
        xs = []
        for i in range(len(cur)):                                             # _make_transition_layer :347-381
            if i < len(pre):
                if cur[i] != pre[i]:
                    xs.append(F.relu(batchnorm(P, conv2d(P, ys[i], f"{tname}.{i}.0", cur[i], 3, 1, 1, gain=g), f"{tname}.{i}.1")))
                else:
                    xs.append(ys[i])
            else:
                t = ys[-1]
                for j in range(i + 1 - len(pre)):
                    co = cur[i] if j == i - len(pre) else pre[-1]
                    t = F.relu(batchnorm(P, conv2d(P, t, f"{tname}.{i}.{j}.0", co, 3, 2, 1, gain=g), f"{tname}.{i}.{j}.1"))
                xs.append(t)
        for m in range(sc["modules"]):
            xs = hr_module(P, xs, f"{prefix}.{sname}.{m}", sc["blocks"], cur)
        ys, pre = xs, cur
    return ys


def hrnet_seg(P, x, nclass=19, hcfg=None):
    """HighResolutionNet(SegBaseModel).forward (models/hrnet_seg.py:23-29) + _HRNetHead (:54-63): bilinear
    (align_corners=False) to the /4 branch, cat, 1x1 conv(+bias)+BN+ReLU, 1x1 conv(+bias), bilinear(False) to the input."""
    hcfg = hcfg or HRNET_W18_SMALL_V1
    size = x.shape[2:]
    ys = hrnet_backbone(P, x, "encoder", hcfg)
    s0 = ys[0].shape[2:]
    cat = torch.cat([ys[0]] + [F.interpolate(t, size=s0, mode="bilinear", align_corners=False) for t in ys[1:]], 1)
    c = cat.shape[1]
    y = conv2d(P, cat, "hrnet_head.last_layer.0", c, 1, bias=True, gain=math.sqrt(2.0))
    y = F.relu(batchnorm(P, y, "hrnet_head.last_layer.1"))
    k = hcfg["final_conv_kernel"]
    y = conv2d(P, y, "hrnet_head.last_layer.3", nclass, k, 1, 1 if k == 3 else 0, bias=True, gain=4.0)
    return F.interpolate(y, size=size, mode="bilinear", align_corners=False)

# ----------------------------------------------------------------------------------------
# heads
# ----------------------------------------------------------------------------------------
def aspp(P, x, prefix="head.aspp", out_ch=256, output_stride=16):
    """_ASPP.forward (modules/module.py:62-77).  Concat order [pool, aspp0..3] (:70).  Always
    nn.BatchNorm2d with default eps (:46,54,58).  Dropout2d is identity in eval."""
    d = {16: (6, 12, 18), 8: (12, 24, 36), 32: (6, 12, 18)}[output_stride]
    size = x.shape[2:]
    pool = F.adaptive_avg_pool2d(x, 1)                                                # :52
    pool = conv_bn_act(P, pool, prefix + ".image_pooling", out_ch, 1)                  # :53-55
    pool = F.interpolate(pool, size=size, mode="bilinear", align_corners=True)        # :64
    x0 = conv_bn_act(P, x, prefix + ".aspp0", out_ch, 1)
    x1 = separable_conv2d(P, x, prefix + ".aspp1", out_ch, 1, d[0], relu_first=False)
    x2 = separable_conv2d(P, x, prefix + ".aspp2", out_ch, 1, d[1], relu_first=False)
    x3 = separable_conv2d(P, x, prefix + ".aspp3", out_ch, 1, d[2], relu_first=False)
    y = torch.cat((pool, x0, x1, x2, x3), dim=1)
    y = conv_bn_act(P, y, prefix, out_ch, 1)                                          # :72-74
    return dropout2d(P, y, prefix + ".dropout", 0.1)                                  # :75 (identity in eval)


def deeplab_head(P, c4, c1, nclass, prefix="head", use_aspp=True, use_decoder=True,
                 output_stride=16):
    """_DeepLabHead.forward (models/deeplabv3_plus.py:66-75)."""
    x = c4
    if use_aspp:
        x = aspp(P, x, prefix + ".aspp", 256, output_stride)
    if use_decoder:
        x = F.interpolate(x, c1.shape[2:], mode="bilinear", align_corners=True)       # :71
        c1 = conv_bn_act(P, c1, prefix + ".c1_block", 48, 1)                          # :72
        x = torch.cat([x, c1], dim=1)                                                 # :73
    x = separable_conv2d(P, x, prefix + ".block.0", 256, relu_first=False)            # :62
    x = separable_conv2d(P, x, prefix + ".block.1", 256, relu_first=False)            # :63
    return conv2d(P, x, prefix + ".block.2", nclass, 1, bias=True, gain=4.0)          # :64


def deeplabv3plus(P, x, backbone="xception65", nclass=19, output_stride=16, eps_encoder=1e-5,
                  use_aspp=True, use_decoder=True, return_lowres=False):
    """DeepLabV3Plus.forward (models/deeplabv3_plus.py:33-46), aux head off (SOLVER.AUX False).

    ``eps_encoder`` mirrors cfg.MODEL.BN_EPS_FOR_ENCODER applied by tools/eval.py:50-53
    (1e-3 for the Xception65 YAML)."""
    size = x.shape[2:]
    if backbone == "xception65":
        c1, _, _, c4 = xception65(P, x, "encoder", output_stride, eps_encoder)
    elif backbone == "mobilenet_v2":
        c1, _, _, c4 = mobilenet_v2(P, x, "encoder", output_stride, eps_encoder)
    elif backbone == "resnet101":
        c1, _, _, c4 = resnet_v1(P, x, "encoder", (3, 4, 23, 3), output_stride, eps=eps_encoder)
    else:
        raise NotImplementedError(backbone)
    y = deeplab_head(P, c4, c1, nclass, "head", use_aspp, use_decoder, output_stride)
    out = F.interpolate(y, size, mode="bilinear", align_corners=True)                 # :39
    return (out, y) if return_lowres else out


def pyramid_pooling(P, x, prefix, sizes=(1, 2, 3, 6)):
    """PyramidPooling.forward (modules/module.py:92-97): cat[x, up(conv(pool_s(x)))...]."""
    c = x.shape[1]
    size = x.shape[2:]
    feats = [x]
    for i, s in enumerate(sizes):
        y = F.adaptive_avg_pool2d(x, s)
        y = conv_bn_act(P, y, f"{prefix}.convs.{i}", c // 4, 1)
        feats.append(F.interpolate(y, size, mode="bilinear", align_corners=True))
    return torch.cat(feats, dim=1)


def pam(P, x, prefix, gamma=0.5):
    """PAM_Module.forward (modules/module.py:112-131): softmax(Q^T K) (no 1/sqrt(d)), V A^T,
    gamma*out + x.  q/k/v 1x1 convs carry bias (:106-108).  ``gamma`` non-zero on creation so
    the branch is not vacuous (reference default 0, :109)."""
    b, c, h, w = x.shape
    q = conv2d(P, x, prefix + ".query_conv", c // 8, 1, bias=True).view(b, -1, h * w).permute(0, 2, 1)
    k = conv2d(P, x, prefix + ".key_conv", c // 8, 1, bias=True).view(b, -1, h * w)
    v = conv2d(P, x, prefix + ".value_conv", c, 1, bias=True).view(b, -1, h * w)
    g = P.scalar(prefix + ".gamma", gamma)
    att = torch.softmax(torch.bmm(q, k), dim=-1)
    out = torch.bmm(v, att.permute(0, 2, 1)).view(b, c, h, w)
    return g * out + x


def cam(P, x, prefix, gamma=0.5):
    """CAM_Module.forward (modules/module.py:142-162): E = X X^T, softmax(rowmax(E) - E), A X."""
    b, c, h, w = x.shape
    q = x.view(b, c, -1)
    e = torch.bmm(q, q.permute(0, 2, 1))
    e = torch.max(e, -1, keepdim=True)[0].expand_as(e) - e
    att = torch.softmax(e, dim=-1)
    out = torch.bmm(att, q).view(b, c, h, w)
    return P.scalar(prefix + ".gamma", gamma) * out + x


def ca_weight(t, f):
    """_C.ca_forward (csrc/criss_cross_attention/ca_cuda.cu:8-36): energies over the criss-cross
    neighbourhood.  Output channel z<W: same row, column z (self included); z>=W: same column,
    row j = i<y ? i : i+1 with i=z-W (self excluded)."""
    n, c, h, w = t.shape
    row = torch.einsum("nchw,nchv->nvhw", t, f)          # [n, W(key col v), h, w]
    col = torch.einsum("nchw,ncjw->njhw", t, f)          # [n, H(key row j), h, w]
    out = t.new_zeros(n, h + w - 1, h, w)
    out[:, :w] = row
    for y in range(h):
        js = [j for j in range(h) if j != y]
        out[:, w:, y, :] = col[:, js, y, :]
    return out


def ca_map(wgt, g):
    """_C.ca_map_forward (ca_cuda.cu:94-120): aggregate values with the same index map."""
    n, c, h, w = g.shape
    out = torch.einsum("nvhw,nchv->nchw", wgt[:, :w], g)
    for y in range(h):
        js = [j for j in range(h) if j != y]
        out[:, :, y, :] += torch.einsum("njw,ncjw->ncw", wgt[:, w:, y, :], g[:, :, js, :])
    return out


def criss_cross_attention(P, x, prefix, gamma=0.5):
    """CrissCrossAttention.forward (modules/cc_attention.py:62-72): softmax over H+W-1 (dim 1)."""
    c = x.shape[1]
    q = conv2d(P, x, prefix + ".query_conv", c // 8, 1, bias=True)
    k = conv2d(P, x, prefix + ".key_conv", c // 8, 1, bias=True)
    v = conv2d(P, x, prefix + ".value_conv", c, 1, bias=True)
    att = torch.softmax(ca_weight(q, k), dim=1)
    return P.scalar(prefix + ".gamma", gamma) * ca_map(att, v) + x


def ccnet(P, x, nclass=19, output_stride=16, recurrence=2):
    """CCNet.forward (models/ccnet.py:27-40) with _CCHead / _RCCAModule (:43-82): conva -> CCA x recurrence (shared
    weights, cfg.MODEL.CCNET.RECURRENCE = 2, config/settings.py:171) -> convb -> cat[x, out] -> bottleneck (3x3 + BN, no
    ReLU; Dropout2d identity in eval) -> 1x1 classifier -> bilinear(align_corners=True)."""
    size = x.shape[2:]
    _, _, _, c4 = resnet_v1(P, x, "encoder", (3, 4, 23, 3), output_stride)
    out = conv_bn_act(P, c4, "head.rcca.conva", 512, 3, 1, 1, conv="0", bn="1")
    for _ in range(recurrence):
        out = criss_cross_attention(P, out, "head.rcca.cca")
    out = conv_bn_act(P, out, "head.rcca.convb", 512, 3, 1, 1, conv="0", bn="1")
    out = torch.cat([c4, out], dim=1)                                               # ccnet.py:79
    out = conv_bn_act(P, out, "head.rcca.bottleneck", 512, 3, 1, 1, act=None, conv="0", bn="1")
    out = dropout2d(P, out, "head.rcca.bottleneck.dropout", 0.1)                    # ccnet.py:70 (identity in eval)
    out = conv2d(P, out, "head.out", nclass, 1, bias=True, gain=4.0)
    return F.interpolate(out, size, mode="bilinear", align_corners=True)


def fcn_head(P, x, prefix, nclass):
    """_FCNHead.forward (modules/module.py:13-27): conv3x3(C -> C/4, no bias) + BN + ReLU + Dropout(0.1) (identity in eval) +
    conv1x1(C/4 -> nclass, bias)."""
    y = conv_bn_act(P, x, prefix + ".block", x.shape[1] // 4, 3, 1, 1, conv="0", bn="1")
    return conv2d(P, y, prefix + ".block.4", nclass, 1, bias=True, gain=4.0)


def pspnet(P, x, nclass=19, output_stride=8, aux=True, all=False):
    """PSPNet.forward (models/pspnet.py:30-43) with _PSPHead (:46-61): ResNet101 (OS8, cityscapes_pspnet_resnet.yaml:25) ->
    PyramidPooling(2048) -> conv3x3(4096 -> 512, no bias) + BN + ReLU + Dropout(0.1) (identity in eval) -> conv1x1(512 -> nclass)
    -> bilinear(align_corners=True); the auxiliary _FCNHead(1024) on c3 (SOLVER.AUX True in the YAML) is the second output.
    Inference only (nn.Dropout, not Dropout2d, sits in both heads)."""
    assert not P.training, "the PSPNet oracle is inference-only"
    size = x.shape[2:]
    _, _, c3, c4 = resnet_v1(P, x, "encoder", (3, 4, 23, 3), output_stride)
    y = pyramid_pooling(P, c4, "head.psp")
    y = conv_bn_act(P, y, "head.block", 512, 3, 1, 1, conv="0", bn="1")
    y = conv2d(P, y, "head.block.4", nclass, 1, bias=True, gain=4.0)
    out = F.interpolate(y, size, mode="bilinear", align_corners=True)
    if not aux:
        return out
    auxout = F.interpolate(fcn_head(P, c3, "auxlayer", nclass), size, mode="bilinear", align_corners=True)
    return (out, auxout) if all else out


def danet_head(P, x, nclass, prefix="head"):
    """DANetHead.forward (models/danet.py:70-88).  Each classifier is nn.Sequential(Dropout2d(0.1), Conv2d) (:64-68): the three
    dropouts are identity in eval and, in training, draw their masks in the order conv6, conv7, conv8."""
    def cbr(x, name):                                  # nn.Sequential(conv3x3(pad 1, no bias), BN, ReLU)  danet.py:48-61
        return conv_bn_act(P, x, f"{prefix}.{name}", 512, 3, 1, 1, conv="0", bn="1")
    feat1 = cbr(x, "conv5a")
    sa_conv = cbr(pam(P, feat1, prefix + ".sa"), "conv51")
    sa_output = conv2d(P, dropout2d(P, sa_conv, prefix + ".conv6.0"), prefix + ".conv6.1", nclass, 1, bias=True, gain=4.0)
    feat2 = cbr(x, "conv5c")
    sc_conv = cbr(cam(P, feat2, prefix + ".sc"), "conv52")
    sc_output = conv2d(P, dropout2d(P, sc_conv, prefix + ".conv7.0"), prefix + ".conv7.1", nclass, 1, bias=True, gain=4.0)
    sasc_output = conv2d(P, dropout2d(P, sa_conv + sc_conv, prefix + ".conv8.0"), prefix + ".conv8.1", nclass, 1, bias=True, gain=4.0)
    return sasc_output, sa_output, sc_output


def danet(P, x, nclass=19, output_stride=8, multi_grid=True, multi_dilation=(4, 8, 16)):
    """DANet.forward (models/danet.py:26-41): ResNet101 (OS8, multi-grid 4/8/16) + DANetHead, three logit maps, each
    bilinearly upsampled (align_corners=True) to the input size."""
    size = x.shape[2:]
    _, _, _, c4 = resnet_v1(P, x, "encoder", (3, 4, 23, 3), output_stride, multi_grid, list(multi_dilation))
    outs = danet_head(P, c4, nclass)
    return tuple(F.interpolate(o, size, mode="bilinear", align_corners=True) for o in outs)


def base_attention_block(P, x, prefix, key_ch, value_ch, out_ch):
    """BaseAttentionBlock.forward (models/ocnet.py:95-113) at scale 1: value = conv1x1(bias); key = query = ReLU(BN(conv1x1(bias)))
    -- ONE module registered under both names (`self.f_query = self.f_key`, :88), so the state dict carries each of its tensors
    twice --; sim = softmax(Q K^T * key_ch^-0.5) over the keys; context = sim V; W = conv1x1(bias).  ``W`` is non-zero on creation
    so that the branch is not vacuous (the reference initialises it to 0, :90-91)."""
    b, c, h, w = x.shape
    value = conv2d(P, x, prefix + ".f_value", value_ch, 1, bias=True).view(b, value_ch, -1).permute(0, 2, 1)
    kq = F.relu(batchnorm(P, conv2d(P, x, prefix + ".f_key.0", key_ch, 1, bias=True), prefix + ".f_key.1"))
    for suffix in ("0.weight", "0.bias", "1.weight", "1.bias", "1.running_mean", "1.running_var", "1.num_batches_tracked"):
        if f"{prefix}.f_query.{suffix}" not in P.t and not P.frozen:
            P.t[f"{prefix}.f_query.{suffix}"] = P.t[f"{prefix}.f_key.{suffix}"]          # shared module: same tensors
    query = kq.view(b, key_ch, -1).permute(0, 2, 1)
    key = kq.view(b, key_ch, -1)
    sim = torch.softmax(torch.bmm(query, key) * (key_ch ** -0.5), dim=-1)
    ctx = torch.bmm(sim, value).permute(0, 2, 1).contiguous().view(b, value_ch, h, w)
    return conv2d(P, ctx, prefix + ".W", out_ch, 1, bias=True)


def ocnet(P, x, nclass=19, output_stride=16, layers=(3, 4, 6, 3)):
    """OCNet.forward (models/ocnet.py:30-43) with _OCHead 'base' (:47-54, cfg.MODEL.OCNet.OC_ARCH = 'base', settings.py:162):
    ResNet50 OS16 (configs/cityscapes_ocnet.yaml) -> conv3x3(2048 -> 512, no bias) + BN + ReLU -> BaseOCModule(512, 512, 256, 256,
    scales [1]) (:116-141: one BaseAttentionBlock, cat[context, x], project = conv1x1(1024 -> 512, bias) + BN + ReLU + Dropout2d(0.05),
    identity in eval) -> conv1x1(512 -> nclass, bias) -> bilinear(align_corners=True).  SOLVER.AUX is off in the YAML.
    Inference only."""
    assert not P.training, "the OCNet oracle is inference-only"
    size = x.shape[2:]
    _, _, _, c4 = resnet_v1(P, x, "encoder", layers, output_stride)
    f = conv_bn_act(P, c4, "head.context", 512, 3, 1, 1, conv="0", bn="1")
    ctx = base_attention_block(P, f, "head.context.3.stages.0", 256, 256, 512)
    y = torch.cat([ctx, f], dim=1)                                                  # ocnet.py:138-139
    y = F.relu(batchnorm(P, conv2d(P, y, "head.context.3.project.0", 512, 1, bias=True), "head.context.3.project.1"))
    y = conv2d(P, y, "head.out", nclass, 1, bias=True, gain=4.0)
    return F.interpolate(y, size, mode="bilinear", align_corners=True)


# ----------------------------------------------------------------------------------------
# convenience
# ----------------------------------------------------------------------------------------
MODELS = {
    # name: (forward kwargs)  -- BASELINE.json configs
    "deeplabv3plus_xception65": dict(backbone="xception65", eps_encoder=1e-3, use_aspp=True,
                                     use_decoder=True),
    "deeplabv3plus_mobilenet_v2": dict(backbone="mobilenet_v2", eps_encoder=1e-5, use_aspp=False,
                                       use_decoder=False),
    "deeplabv3plus_resnet101": dict(backbone="resnet101", eps_encoder=1e-5, use_aspp=True, use_decoder=True),
}


def build_params(model: str, seed: int = 0, nclass: int = 19) -> Params:
    """Create the full reference-named parameter dict for ``model`` by tracing one tiny forward."""
    P = Params(seed)
    with torch.no_grad():
        if model == "danet_resnet101":
            danet(P, torch.zeros(1, 3, 33, 33), nclass=nclass)
        elif model == "ccnet_resnet101":
            ccnet(P, torch.zeros(1, 3, 33, 33), nclass=nclass)
        elif model == "pspnet_resnet101":
            pspnet(P, torch.zeros(1, 3, 49, 49), nclass=nclass)
        elif model == "ocnet_resnet50":
            ocnet(P, torch.zeros(1, 3, 33, 33), nclass=nclass)
        elif model == "hrnet_w18_small_v1":
            hrnet_seg(P, torch.zeros(1, 3, 64, 64), nclass=nclass)
        else:
            deeplabv3plus(P, torch.zeros(1, 3, 33, 33), nclass=nclass, **MODELS[model])
    P.frozen = True
    return P


def trainable(P: Params):
    """Names of the tensors an optimizer would see (everything except BatchNorm running statistics and the unused
    ImageNet classifier)."""
    return [k for k in P.t if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))
            and ".fc." not in k]


def loss_and_grads(model: str, P: Params, x, target, nclass: int = 19, ignore_index: int = -1, **kw):
    """One training forward/backward of tools/train.py:135-147 for a DeepLabv3+ model: train-mode forward,
    nn.CrossEntropyLoss(ignore_index=-1) on the full-resolution logits (solver/loss.py:16-46, aux off), backward.
    -> (loss, {name: grad}, full-res logits, low-res logits).  P must hold fp32 CPU (or same-device) tensors; BatchNorm
    running statistics in P are updated in place like the reference's buffers."""
    assert P.frozen, "build the parameters first"
    names = trainable(P)
    leaves = {}
    for k in names:
        P.t[k] = P.t[k].detach().clone().requires_grad_(True)
        leaves[k] = P.t[k]
    was = P.training
    P.training = True
    try:
        if model == "ccnet_resnet101":
            out, low = ccnet(P, x, nclass=nclass, **kw), None
        elif model == "hrnet_w18_small_v1":
            out, low = hrnet_seg(P, x, nclass=nclass, **kw), None
        elif model == "danet_resnet101":
            outs = danet(P, x, nclass=nclass, **kw)
            out, low = outs[0], None
        else:
            out, low = deeplabv3plus(P, x, nclass=nclass, return_lowres=True, **MODELS[model], **kw)
        if model == "danet_resnet101":                 # MixSoftmaxCrossEntropyLoss._multiple_forward (solver/loss.py:31-36)
            loss = sum(F.cross_entropy(o.float(), target, ignore_index=ignore_index) for o in outs)
        else:
            loss = F.cross_entropy(out.float(), target, ignore_index=ignore_index)
        loss.backward()
    finally:
        P.training = was
    grads = {k: (v.grad.detach() if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    for k in names:
        P.t[k] = P.t[k].detach()
    return loss.detach(), grads, out.detach(), (low.detach() if low is not None else None)


def forward(model: str, P: Params, x, nclass: int = 19, **kw):
    """-> logits [N, nclass, H, W] (for DANet: the first of its three outputs unless all=True)."""
    with torch.no_grad():
        if model == "danet_resnet101":
            outs = danet(P, x, nclass=nclass)
            return outs if kw.get("all") else outs[0]
        if model == "ccnet_resnet101":
            return ccnet(P, x, nclass=nclass)
        if model == "pspnet_resnet101":
            return pspnet(P, x, nclass=nclass, all=bool(kw.get("all")))
        if model == "ocnet_resnet50":
            return ocnet(P, x, nclass=nclass)
        if model == "hrnet_w18_small_v1":
            return hrnet_seg(P, x, nclass=nclass)
        return deeplabv3plus(P, x, nclass=nclass, **MODELS[model], **kw)

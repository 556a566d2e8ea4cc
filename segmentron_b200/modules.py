"""Drop-in ``nn.Module`` replacements for the reference's L1 building blocks.

Same constructor signatures, same sub-module names, hence the same ``state_dict`` keys as
``segmentron/modules/basic.py`` and ``segmentron/modules/module.py``; parameters stay ordinary
``nn.Parameter``s owned by the module (optimizer param groups, ``load_state_dict``, ``convert_sync_batchnorm``
and the post-construction ``eps`` mutation of tools/eval.py:50-53 keep working).  ``forward`` takes and returns
logical-NCHW tensors; physically the engine works in NHWC, so a ``channels_last`` 16-bit input crosses the
boundary with zero copies and the output is returned as a ``channels_last`` view.

Folded BatchNorm / packed weights are caches keyed on parameter versions and ``eps``; they are rebuilt
lazily at the next forward after any change.

No CPU implementation and no PyTorch fallback: a non-CUDA input raises RuntimeError, like the reference's
native op does (csrc/criss_cross_attention/ca.h:34 "Not implemented on the CPU").  Training mode: ``SeparableConv2d``,
``_ConvBNReLU``, ``_ConvBN``, ``InvertedResidual``, ``_ASPP``, ``PyramidPooling``, ``PAM_Module``, ``CAM_Module`` and
``CrissCrossAttention`` -- every class -- are differentiable (train_modules.py / attention.py: train-mode BatchNorm + the backward
kernels, composed unit by unit) -- whole models train fastest through ``train.DeepLabV3PlusTrainerB200``.
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from . import fold, ops

_COMPUTE_DTYPE = torch.bfloat16


def set_compute_dtype(dtype):
    """16-bit dtype used when a module receives fp32 activations (default bf16)."""
    global _COMPUTE_DTYPE
    assert dtype in (torch.bfloat16, torch.float16)
    _COMPUTE_DTYPE = dtype


# ------------------------------------------------------------------------------------------------
# boundary helpers
# ------------------------------------------------------------------------------------------------
def _enter(x, module):
    """logical NCHW tensor -> (NHWC 16-bit tensor, original dtype)."""
    if not x.is_cuda:
        raise RuntimeError(f"segb200: {type(module).__name__} is not implemented on the CPU (input must be a CUDA tensor)")
    if module.training:
        raise RuntimeError(f"segb200: {type(module).__name__}.forward_nhwc is the inference path (folded BatchNorm); in training "
                           "mode call the module itself (forward), which runs the differentiable train-mode kernels")
    if x.dim() != 4:
        raise RuntimeError("segb200: expected a 4-D NCHW tensor")
    dt = x.dtype if x.dtype in (torch.bfloat16, torch.float16) else _COMPUTE_DTYPE
    n, c, h, w = x.shape
    if x.dtype == dt and c % 8 == 0 and x.permute(0, 2, 3, 1).is_contiguous():
        return x.permute(0, 2, 3, 1), x.dtype                      # zero-copy: already channels_last
    buf = torch.empty(n, h, w, fold.round_up(c, 8), dtype=dt, device=x.device)
    if buf.shape[3] != c:
        buf.zero_()
    ops.nchw_to_nhwc(x, buf[..., :c])
    return (buf[..., :c] if buf.shape[3] != c else buf), x.dtype


def _train_enter(x, module):
    """training mode: logical NCHW -> NHWC 16-bit through autograd-visible torch ops (boundary plumbing)"""
    if not x.is_cuda:
        raise RuntimeError(f"segb200: {type(module).__name__} is not implemented on the CPU (input must be a CUDA tensor)")
    dt = x.dtype if x.dtype in (torch.bfloat16, torch.float16) else _COMPUTE_DTYPE
    return x.permute(0, 2, 3, 1).contiguous().to(dt)


def _train_conv_bn_act(xh, conv, bn, act, pre_relu=False):
    """one fused training unit (conv or depthwise 3x3 + train-mode BatchNorm + activation) on an NHWC tensor"""
    from . import train_modules as TM
    if conv.bias is not None:
        raise RuntimeError("segb200: conv bias in front of BatchNorm is not supported in training mode")
    if conv.weight.dtype != torch.float32 or bn.running_mean.dtype != torch.float32:
        raise RuntimeError("segb200: training mode expects fp32 parameters / BatchNorm buffers (compute is 16-bit)")
    g, b, rm, rv, mom, eps = TM._bn_args(bn)
    k, s, d, p = conv.kernel_size[0], conv.stride[0], conv.dilation[0], conv.padding[0]
    ho, wo = _out_hw(xh.shape[1], xh.shape[2], k, s, p, d)
    if xh.shape[0] * ho * wo < 2:                # torch.nn.functional.batch_norm raises the same way
        raise ValueError(f"Expected more than 1 value per channel when training, got input size {[xh.shape[0], conv.out_channels, ho, wo]}")
    if conv.groups == 1:
        if pre_relu:
            raise RuntimeError("segb200: leading ReLU is only fused into the depthwise unit")
        weight = conv.weight
        cpad = (-xh.shape[3]) % 8
        if cpad:
            # small-Cin stems (RGB: MobileNetV2 backbones/mobilenet.py:80, FastSCNN, ICNet): zero channels on both operands, like
            # the inference path does; autograd slices the padding off the weight gradient again
            xh = torch.nn.functional.pad(xh, (0, cpad))
            weight = torch.nn.functional.pad(weight, (0, 0, 0, 0, 0, cpad))
        y = TM.ConvBNActFunction.apply(xh, weight, g, b, rm, rv, mom, eps, s, d, p, act)
    elif conv.groups == conv.in_channels == conv.out_channels and k == 3 and p == d:
        y = TM.DwBNActFunction.apply(xh, conv.weight, g, b, rm, rv, mom, eps, s, d, pre_relu, act)
    else:
        raise RuntimeError(f"segb200: no training kernel for Conv2d(groups={conv.groups}, k={k})")
    if bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    return y


def _leave(y_nhwc, out_dtype):
    """NHWC tensor -> logical NCHW (channels_last view); converts back if the caller's dtype was fp32."""
    y = y_nhwc.permute(0, 3, 1, 2)
    return y if y.dtype == out_dtype else y.to(out_dtype)


def _versions(*tensors):
    return tuple((t.data_ptr(), t._version) for t in tensors if t is not None)


class _Cache:
    """Lazily rebuilt derived tensors (folded BN, packed weights)."""

    def __init__(self):
        self.key, self.val = None, None

    def get(self, key, build):
        if self.key != key:
            self.val, self.key = build(), key
        return self.val


def _bn_tensors(bn):
    if not isinstance(bn, nn.modules.batchnorm._BatchNorm) and not hasattr(bn, "running_var"):
        raise RuntimeError(f"segb200: unsupported norm layer {type(bn).__name__}")
    return bn.weight, bn.bias, bn.running_mean, bn.running_var


def _fold(bn):
    w, b, m, v = _bn_tensors(bn)
    return fold.bn_fold(w.detach(), b.detach(), m, v, bn.eps)


def _out_hw(h, w, k, stride, pad, dil):
    return (h + 2 * pad - dil * (k - 1) - 1) // stride + 1, (w + 2 * pad - dil * (k - 1) - 1) // stride + 1


def _run_conv_bn_act(x, conv, bn, act, cache, dt, out=None, residual=None):
    """conv (dense, or depthwise 3x3) + optional BN + activation on an NHWC tensor."""
    n, h, w, cin = x.shape
    k, s, d, p = conv.kernel_size[0], conv.stride[0], conv.dilation[0], conv.padding[0]
    ho, wo = _out_hw(h, w, k, s, p, d)
    cout = conv.out_channels
    bn_t = _bn_tensors(bn) if bn is not None else ()
    key = (dt, bn.eps if bn is not None else None) + _versions(conv.weight, conv.bias, *bn_t)
    if conv.groups == 1:
        def build():
            cop = fold.round_up(cout, 8)
            if bn is not None:
                sc, sh = _fold(bn)
                if conv.bias is not None:
                    sh = sh + conv.bias.detach().float() * sc
                sc, sh = fold.pad_vec(sc, cop, 1.0), fold.pad_vec(sh, cop)
            else:
                sc, sh = None, (fold.pad_vec(conv.bias.detach(), cop) if conv.bias is not None else None)
            return fold.pack_conv_weight(conv.weight.detach(), dt, cop), sc, sh, cop
        wpk, sc, sh, cop = cache.get(key, build)
        if out is None:
            buf = torch.empty(n, ho, wo, cop, dtype=dt, device=x.device)
            out = buf
        ops.conv_gemm(x, wpk, out, cin=cin, cout=cop, kh=k, kw=k, stride=s, dilation=d, pad_t=p, pad_l=p, scale=sc, shift=sh,
                      act=act, residual=residual)
        return out[..., :cout] if out.shape[3] != cout else out
    if conv.groups == cin == cout and k == 3 and p == d:
        def build():
            if bn is not None:
                sc, sh = _fold(bn)
            else:
                sc, sh = torch.ones(cout, device=x.device), torch.zeros(cout, device=x.device)
            if conv.bias is not None:
                sh = sh + conv.bias.detach().float() * sc
            return fold.pack_dw_weight(conv.weight.detach(), sc), sh.contiguous()
        wdw, sh = cache.get(key, build)
        if out is None:
            out = torch.empty(n, ho, wo, cout, dtype=dt, device=x.device)
        return ops.dwconv3x3(x, wdw, out, stride=s, dilation=d, shift=sh, pre_relu=False, act=act)
    raise RuntimeError(f"segb200: no kernel for Conv2d(groups={conv.groups}, k={k}) yet")


# ------------------------------------------------------------------------------------------------
# modules/basic.py replacements
# ------------------------------------------------------------------------------------------------
class SeparableConv2d(nn.Module):
    """Replaces segmentron.modules.basic.SeparableConv2d (basic.py:34-62): one depthwise kernel (pre-ReLU, BN_depth,
    ReLU fused) + one tcgen05 GEMM (BN_point, ReLU fused)."""

    def __init__(self, inplanes, planes, kernel_size=3, stride=1, dilation=1, relu_first=True, bias=False,
                 norm_layer=nn.BatchNorm2d):
        super().__init__()
        depthwise = nn.Conv2d(inplanes, inplanes, kernel_size, stride=stride, padding=dilation, dilation=dilation,
                              groups=inplanes, bias=bias)
        bn_depth = norm_layer(inplanes)
        pointwise = nn.Conv2d(inplanes, planes, 1, bias=bias)
        bn_point = norm_layer(planes)
        self.relu_first = relu_first
        if relu_first:
            layers = [("relu", nn.ReLU()), ("depthwise", depthwise), ("bn_depth", bn_depth), ("pointwise", pointwise),
                      ("bn_point", bn_point)]
        else:
            layers = [("depthwise", depthwise), ("bn_depth", bn_depth), ("relu1", nn.ReLU(inplace=True)),
                      ("pointwise", pointwise), ("bn_point", bn_point), ("relu2", nn.ReLU(inplace=True))]
        self.block = nn.Sequential(OrderedDict(layers))
        self._c_dw, self._c_pw = _Cache(), _Cache()

    def forward_nhwc(self, x, out=None, residual=None):
        b = self.block
        dt = x.dtype
        n, h, w, c = x.shape
        dwc = b.depthwise
        s, d = dwc.stride[0], dwc.dilation[0]
        key = (b.bn_depth.eps,) + _versions(dwc.weight, dwc.bias, *_bn_tensors(b.bn_depth))

        def build():
            sc, sh = _fold(b.bn_depth)
            if dwc.bias is not None:
                sh = sh + dwc.bias.detach().float() * sc
            return fold.pack_dw_weight(dwc.weight.detach(), sc), sh.contiguous()
        wdw, sh = self._c_dw.get(key, build)
        ho, wo = (h - 1) // s + 1, (w - 1) // s + 1
        tmp = torch.empty(n, ho, wo, c, dtype=dt, device=x.device)
        ops.dwconv3x3(x, wdw, tmp, stride=s, dilation=d, shift=sh, pre_relu=self.relu_first,
                      act=None if self.relu_first else "relu")
        return _run_conv_bn_act(tmp, b.pointwise, b.bn_point, None if self.relu_first else "relu", self._c_pw, dt, out=out,
                                residual=residual)

    def forward(self, x):
        if self.training:                          # differentiable path: train-mode BatchNorm + backward kernels
            b = self.block
            act = None if self.relu_first else "relu"
            y = _train_conv_bn_act(_train_enter(x, self), b.depthwise, b.bn_depth, act, pre_relu=self.relu_first)
            y = _train_conv_bn_act(y, b.pointwise, b.bn_point, act)
            return y.permute(0, 3, 1, 2).to(x.dtype)
        xh, odt = _enter(x, self)
        return _leave(self.forward_nhwc(xh), odt)


class _ConvBNReLU(nn.Module):
    """Replaces segmentron.modules.basic._ConvBNReLU (basic.py:65-77): conv + BN + ReLU/ReLU6 in one kernel."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, relu6=False,
                 norm_layer=nn.BatchNorm2d):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias=False)
        self.bn = norm_layer(out_channels)
        self.relu = nn.ReLU6(True) if relu6 else nn.ReLU(True)
        self._act = "relu6" if relu6 else "relu"
        self._cache = _Cache()

    def forward_nhwc(self, x, out=None):
        return _run_conv_bn_act(x, self.conv, self.bn, self._act, self._cache, x.dtype, out=out)

    def forward(self, x):
        if self.training:
            return _train_conv_bn_act(_train_enter(x, self), self.conv, self.bn, self._act).permute(0, 3, 1, 2).to(x.dtype)
        xh, odt = _enter(x, self)
        return _leave(self.forward_nhwc(xh), odt)


class _ConvBN(nn.Module):
    """Replaces segmentron.modules.basic._ConvBN (basic.py:95-105)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 norm_layer=nn.BatchNorm2d, **kwargs):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias=False)
        self.bn = norm_layer(out_channels)
        self._cache = _Cache()

    def forward_nhwc(self, x, out=None):
        return _run_conv_bn_act(x, self.conv, self.bn, None, self._cache, x.dtype, out=out)

    def forward(self, x):
        if self.training:
            return _train_conv_bn_act(_train_enter(x, self), self.conv, self.bn, None).permute(0, 3, 1, 2).to(x.dtype)
        xh, odt = _enter(x, self)
        return _leave(self.forward_nhwc(xh), odt)


class InvertedResidual(nn.Module):
    """Replaces segmentron.modules.basic.InvertedResidual (basic.py:139-163); the residual add is fused into the
    pw-linear GEMM epilogue."""

    def __init__(self, in_channels, out_channels, stride, expand_ratio, dilation=1, norm_layer=nn.BatchNorm2d):
        super().__init__()
        assert stride in [1, 2]
        self.use_res_connect = stride == 1 and in_channels == out_channels
        layers = []
        inter = int(round(in_channels * expand_ratio))
        if expand_ratio != 1:
            layers.append(_ConvBNReLU(in_channels, inter, 1, relu6=True, norm_layer=norm_layer))
        layers.extend([_ConvBNReLU(inter, inter, 3, stride, dilation, dilation, groups=inter, relu6=True,
                                   norm_layer=norm_layer),
                       nn.Conv2d(inter, out_channels, 1, bias=False), norm_layer(out_channels)])
        self.conv = nn.Sequential(*layers)
        self._cache = _Cache()

    def forward_nhwc(self, x):
        y = x
        mods = list(self.conv)
        for m in mods[:-2]:
            y = m.forward_nhwc(y)
        return _run_conv_bn_act(y, mods[-2], mods[-1], None, self._cache, x.dtype,
                                residual=x if self.use_res_connect else None)

    def forward(self, x):
        if self.training:                          # unit by unit through the training kernels; the skip is a torch add
            xh = _train_enter(x, self)
            mods = list(self.conv)
            y = xh
            for m in mods[:-2]:
                y = _train_conv_bn_act(y, m.conv, m.bn, m._act)
            y = _train_conv_bn_act(y, mods[-2], mods[-1], None)
            if self.use_res_connect:
                y = y + xh
            return y.permute(0, 3, 1, 2).to(x.dtype)
        xh, odt = _enter(x, self)
        return _leave(self.forward_nhwc(xh), odt)


# ------------------------------------------------------------------------------------------------
# modules/module.py replacements
# ------------------------------------------------------------------------------------------------
class _ASPP(nn.Module):
    """Replaces segmentron.modules.module._ASPP (module.py:32-77).  ``output_stride`` comes from the reference's
    global cfg when it is importable (module.py:35), else from the keyword.  Every branch writes its 256-channel
    slice of one 1280-channel NHWC buffer (no torch.cat); the image-pooling branch is GAP -> GEMM -> broadcast."""

    def __init__(self, in_channels=2048, out_channels=256, output_stride=None):
        super().__init__()
        if output_stride is None:
            try:
                from segmentron.config import cfg          # the reference's global config (module.py:8,35)
                output_stride = cfg.MODEL.OUTPUT_STRIDE
            except Exception:
                output_stride = 16
        dilations = {16: [6, 12, 18], 8: [12, 24, 36], 32: [6, 12, 18]}.get(output_stride)
        if dilations is None:
            raise NotImplementedError
        self.aspp0 = nn.Sequential(OrderedDict([("conv", nn.Conv2d(in_channels, out_channels, 1, bias=False)),
                                                ("bn", nn.BatchNorm2d(out_channels)), ("relu", nn.ReLU(inplace=True))]))
        self.aspp1 = SeparableConv2d(in_channels, out_channels, dilation=dilations[0], relu_first=False)
        self.aspp2 = SeparableConv2d(in_channels, out_channels, dilation=dilations[1], relu_first=False)
        self.aspp3 = SeparableConv2d(in_channels, out_channels, dilation=dilations[2], relu_first=False)
        self.image_pooling = nn.Sequential(OrderedDict([("gap", nn.AdaptiveAvgPool2d((1, 1))),
                                                        ("conv", nn.Conv2d(in_channels, out_channels, 1, bias=False)),
                                                        ("bn", nn.BatchNorm2d(out_channels)),
                                                        ("relu", nn.ReLU(inplace=True))]))
        self.conv = nn.Conv2d(out_channels * 5, out_channels, 1, bias=False)
        self.bn = nn.BatchNorm2d(out_channels)
        self.relu = nn.ReLU(inplace=True)
        self.dropout = nn.Dropout2d(p=0.1)
        self._c0, self._cp, self._cproj = _Cache(), _Cache(), _Cache()

    def forward_nhwc(self, x):
        n, h, w, c = x.shape
        dt = x.dtype
        oc = self.conv.out_channels
        cat = torch.empty(n, h, w, 5 * oc, dtype=dt, device=x.device)
        pooled = torch.empty(n, 1, 1, c, dtype=dt, device=x.device)
        ops.global_avgpool(x, pooled)
        pf = _run_conv_bn_act(pooled, self.image_pooling.conv, self.image_pooling.bn, "relu", self._cp, dt)
        ops.bilinear_nhwc(pf, cat[..., 0:oc], align_corners=True)
        _run_conv_bn_act(x, self.aspp0.conv, self.aspp0.bn, "relu", self._c0, dt, out=cat[..., oc:2 * oc])
        for i, m in enumerate((self.aspp1, self.aspp2, self.aspp3)):
            m.forward_nhwc(x, out=cat[..., (2 + i) * oc:(3 + i) * oc])
        return _run_conv_bn_act(cat, self.conv, self.bn, "relu", self._cproj, dt)     # Dropout2d: identity in eval

    def forward_train(self, x):
        """training mode (module.py:62-77 with batch statistics): every conv+BN(+ReLU) is one differentiable unit of the training
        kernels; GAP / broadcast are their own Functions; the concat is a torch.cat of NHWC tensors (boundary plumbing) and
        Dropout2d is torch's own on the logical-NCHW result, so it draws the same random planes as the reference would."""
        from . import train_modules as TM
        xh = _train_enter(x, self)
        _, h, w, _ = xh.shape
        ip = self.image_pooling
        pool = TM.BroadcastFunction.apply(_train_conv_bn_act(TM.GlobalAvgPoolFunction.apply(xh), ip.conv, ip.bn, "relu"), h, w)
        branches = [pool, _train_conv_bn_act(xh, self.aspp0.conv, self.aspp0.bn, "relu")]
        for m in (self.aspp1, self.aspp2, self.aspp3):
            b = m.block
            act = None if m.relu_first else "relu"
            y = _train_conv_bn_act(xh, b.depthwise, b.bn_depth, act, pre_relu=m.relu_first)
            branches.append(_train_conv_bn_act(y, b.pointwise, b.bn_point, act))
        y = _train_conv_bn_act(torch.cat(branches, dim=3), self.conv, self.bn, "relu")
        return self.dropout(y.permute(0, 3, 1, 2).to(x.dtype))

    def forward(self, x):
        if self.training:
            return self.forward_train(x)
        xh, odt = _enter(x, self)
        return _leave(self.forward_nhwc(xh), odt)


class PyramidPooling(nn.Module):
    """Replaces segmentron.modules.module.PyramidPooling (module.py:82-97): adaptive pools, 1x1 conv+BN+ReLU GEMMs and
    bilinear up-samples all write channel slices of the 2C output buffer (no torch.cat)."""

    def __init__(self, in_channels, sizes=(1, 2, 3, 6), norm_layer=nn.BatchNorm2d, **kwargs):
        super().__init__()
        out_channels = int(in_channels / 4)
        self.sizes = tuple(sizes)
        self.avgpools = nn.ModuleList()
        self.convs = nn.ModuleList()
        for size in sizes:
            self.avgpools.append(nn.AdaptiveAvgPool2d(size))
            self.convs.append(_ConvBNReLU(in_channels, out_channels, 1, norm_layer=norm_layer, **kwargs))

    def forward_nhwc(self, x):
        n, h, w, c = x.shape
        oc = c // 4
        out = torch.empty(n, h, w, c + oc * len(self.sizes), dtype=x.dtype, device=x.device)
        out[..., :c].copy_(x)
        for i, (s, conv) in enumerate(zip(self.sizes, self.convs)):
            p = torch.empty(n, s, s, c, dtype=x.dtype, device=x.device)
            ops.adaptive_avgpool(x, p, s)
            f = conv.forward_nhwc(p)
            ops.bilinear_nhwc(f, out[..., c + i * oc:c + (i + 1) * oc], align_corners=True)
        return out

    def forward(self, x):
        if self.training:                          # unit by unit: adaptive pool / conv+BN+ReLU / bilinear Functions, torch.cat of NHWC tensors
            from . import train_modules as TM
            xh = _train_enter(x, self)
            _, h, w, _ = xh.shape
            feats = [xh]
            for s, conv in zip(self.sizes, self.convs):
                f = _train_conv_bn_act(TM.AdaptiveAvgPoolFunction.apply(xh, s), conv.conv, conv.bn, conv._act)
                feats.append(TM.BilinearFunction.apply(f, h, w, True))
            return torch.cat(feats, dim=3).permute(0, 3, 1, 2).to(x.dtype)
        xh, odt = _enter(x, self)
        return _leave(self.forward_nhwc(xh), odt)


class PAM_Module(nn.Module):
    """Replaces segmentron.modules.module.PAM_Module (module.py:100-131): q/k 1x1 convs and V^T as tcgen05 GEMMs, then the
    tiled softmax(Q K^T) V kernel (csrc/pam.cu) with gamma*out + x in its epilogue."""

    def __init__(self, in_dim):
        super().__init__()
        self.chanel_in = in_dim
        self.query_conv = nn.Conv2d(in_channels=in_dim, out_channels=in_dim // 8, kernel_size=1)
        self.key_conv = nn.Conv2d(in_channels=in_dim, out_channels=in_dim // 8, kernel_size=1)
        self.value_conv = nn.Conv2d(in_channels=in_dim, out_channels=in_dim, kernel_size=1)
        self.gamma = nn.Parameter(torch.zeros(1))
        self.softmax = nn.Softmax(dim=-1)
        self._cache = _Cache()

    def forward_nhwc(self, x):
        from .attention import pam_nhwc
        dt = x.dtype
        convs = (self.query_conv, self.key_conv, self.value_conv)
        key = (dt,) + _versions(*[t for cv in convs for t in (cv.weight, cv.bias)])

        def build():
            return tuple(t for cv in convs for t in (fold.pack_conv_weight(cv.weight.detach(), dt),
                                                     cv.bias.detach().float().contiguous()))
        wq, bq, wk, bk, wv, bv = self._cache.get(key, build)
        if not x.is_contiguous():
            x = x.contiguous()
        return pam_nhwc(x, wq, bq, wk, bk, wv, bv, self.gamma)

    def forward(self, x):
        if self.training:                          # differentiable path (attention.PamFunction: materialised attention, like the reference)
            from .attention import PamFunction
            y = PamFunction.apply(_train_enter(x, self), self.query_conv.weight, self.query_conv.bias, self.key_conv.weight,
                                  self.key_conv.bias, self.value_conv.weight, self.value_conv.bias, self.gamma)
            return y.permute(0, 3, 1, 2).to(x.dtype)
        xh, odt = _enter(x, self)
        return _leave(self.forward_nhwc(xh), odt)


class CAM_Module(nn.Module):
    """Replaces segmentron.modules.module.CAM_Module (module.py:134-162): Gram matrix and A.X on the tensor cores (fp32
    energies), softmax(rowmax - E) in between; gamma*out + x fused into the second GEMM's epilogue."""

    def __init__(self, in_dim):
        super().__init__()
        self.chanel_in = in_dim
        self.gamma = nn.Parameter(torch.zeros(1))
        self.softmax = nn.Softmax(dim=-1)

    def forward_nhwc(self, x):
        from .attention import cam_nhwc
        return cam_nhwc(x, self.gamma)

    def forward(self, x):
        if self.training:                          # differentiable path (attention.CamFunction / csrc/cam_bwd.cu)
            from .attention import CamFunction
            y = CamFunction.apply(_train_enter(x, self), self.gamma)
            return y.permute(0, 3, 1, 2).to(x.dtype)
        xh, odt = _enter(x, self)
        return _leave(self.forward_nhwc(xh), odt)


class CrissCrossAttention(nn.Module):
    """Replaces segmentron.modules.cc_attention.CrissCrossAttention (cc_attention.py:52-72) and the four
    segmentron._C.ca_* functions it calls (csrc/vision.cpp:6-11)."""

    def __init__(self, in_channels):
        super().__init__()
        self.query_conv = nn.Conv2d(in_channels, in_channels // 8, 1)
        self.key_conv = nn.Conv2d(in_channels, in_channels // 8, 1)
        self.value_conv = nn.Conv2d(in_channels, in_channels, 1)
        self.gamma = nn.Parameter(torch.zeros(1))
        self._cache = _Cache()

    def forward_train(self, x):
        """training mode: differentiable through the CUDA backward kernels (attention.CrissCrossFunction).  x: logical NCHW."""
        from .attention import CrissCrossFunction
        if not x.is_cuda:
            raise RuntimeError("segb200: CrissCrossAttention is not implemented on the CPU (input must be a CUDA tensor)")
        dt = x.dtype if x.dtype in (torch.bfloat16, torch.float16) else _COMPUTE_DTYPE
        xh = x.permute(0, 2, 3, 1).contiguous().to(dt)                     # boundary plumbing (autograd-visible)
        y = CrissCrossFunction.apply(xh, self.query_conv.weight, self.query_conv.bias, self.key_conv.weight, self.key_conv.bias,
                                     self.value_conv.weight, self.value_conv.bias, self.gamma)
        return y.permute(0, 3, 1, 2).to(x.dtype)

    def forward_nhwc(self, x):
        from .attention import cca_nhwc
        dt = x.dtype
        convs = (self.query_conv, self.key_conv, self.value_conv)
        key = (dt,) + _versions(*[t for cv in convs for t in (cv.weight, cv.bias)])

        def build():
            return tuple(t for cv in convs for t in (fold.pack_conv_weight(cv.weight.detach(), dt),
                                                     cv.bias.detach().float().contiguous()))
        wq, bq, wk, bk, wv, bv = self._cache.get(key, build)
        return cca_nhwc(x, wq, bq, wk, bk, wv, bv, self.gamma)

    def forward(self, x):
        if self.training:
            return self.forward_train(x)
        xh, odt = _enter(x, self)
        return _leave(self.forward_nhwc(xh), odt)


REPLACEMENTS = {
    "SeparableConv2d": SeparableConv2d,
    "_ConvBNReLU": _ConvBNReLU,
    "_ConvBN": _ConvBN,
    "InvertedResidual": InvertedResidual,
    "_ASPP": _ASPP,
    "PyramidPooling": PyramidPooling,
    "PAM_Module": PAM_Module,
    "CAM_Module": CAM_Module,
    "CrissCrossAttention": CrissCrossAttention,
}

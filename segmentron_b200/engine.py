"""Whole-model execution plans for the DeepLabv3+ family on the C-ABI kernels.

A ``Plan`` is a flat list of pre-marshalled kernel launches over statically allocated NHWC buffers
(180 GB of HBM3e: no buffer re-use needed), replayed either directly or as one CUDA graph.  Fusion
structure (what the reference runs as 5-6 eager kernels per SeparableConv2d):

  SeparableConv2d   = dwconv3x3[pre-ReLU, BN_depth folded, ReLU]  ->  tcgen05 GEMM[BN_point, ReLU, +skip]
  XceptionBlock     = 3 x SeparableConv2d; the 1x1 stride-s shortcut conv+BN is one GEMM whose output is
                      consumed as the residual operand of the third GEMM's epilogue (no add kernel)
  _ASPP             = every branch GEMM writes its 256-channel slice of the 1280-channel buffer (no cat)
  decoder           = bilinear x4 and c1_block write their slices of the 304-channel buffer (no cat)
  logits            = classifier GEMM (bias as shift) -> one bilinear kernel writing NCHW (+ optional argmax)

Parameters are taken from a reference ``state_dict`` (same key names), so a checkpoint of the reference
model drives this engine unchanged.
"""
import ctypes as C

import torch

from . import fold, lib as L, ops


_DEVICE_FOLD = bool(__import__("os").environ.get("SEGB200_DEVICE_FOLD"))     # A/B switch: fold / pack on the device (round-1 behaviour)


class Plan:
    def __init__(self, params, dtype, device):
        self.P = params            # name -> tensor (reference state_dict keys)
        self.dtype = dtype
        self.device = device
        self.steps = []            # (fn, args struct | tuple)
        self.keep = []             # keep packed weights / buffers alive
        self.lib = L.load()
        self.n_launch = 0
        self._host = {}

    # ---- buffers -----------------------------------------------------------------------
    def new(self, n, h, w, c, ld=None):
        ld = ld or fold.round_up(c, 8)
        buf = torch.empty(n, h, w, ld, dtype=self.dtype, device=self.device)
        self.keep.append(buf)
        return buf[..., :c] if ld != c else buf

    def w(self, name):
        """parameter on the HOST: BatchNorm folding and weight packing run on the CPU at plan-build time (a few hundred tiny
        elementwise ops per model; on the device they would be ~1000 torch launches in front of the first engine kernel), only the
        folded / packed results are uploaded (``dev``)"""
        t = self._host.get(name)
        if t is None:
            t = self._host[name] = self.P[name].detach().to(self.device if _DEVICE_FOLD else "cpu")
        return t

    def dev(self, t):
        """upload a host-side derived tensor (no-op for device tensors / None)"""
        if t is None or t.device.type != "cpu":
            return t
        return t.contiguous().to(self.device)

    def bn(self, prefix, eps):
        return fold.bn_fold(self.w(prefix + ".weight"), self.w(prefix + ".bias"), self.w(prefix + ".running_mean"),
                            self.w(prefix + ".running_var"), eps)

    # ---- op recorders ------------------------------------------------------------------
    def conv(self, x, wpk, y, **kw):
        wpk = self.dev(wpk)
        for k in ("scale", "shift"):
            if kw.get(k) is not None:
                kw[k] = self.dev(kw[k])
                self.keep.append(kw[k])
        self.keep.append(wpk)
        a = ops.make_conv_args(x, wpk, y, **kw)
        fn = self.lib.segb200_conv_gemm
        pix = a.n * a.ho * a.wo
        taps = a.kh * a.kw
        meta = dict(kind="conv_gemm", flops=2.0 * pix * a.cout * a.cin * taps,
                    bytes=2.0 * (a.n * a.h * a.w * a.cin + pix * a.cout * (2 if kw.get("residual") is not None else 1)
                                 + a.cout * a.cin * taps),
                    desc=f"{a.kh}x{a.kw}s{a.stride}d{a.dilation} {a.cin}->{a.cout} @{a.n}x{a.ho}x{a.wo}")
        self.steps.append((lambda s, a=a, fn=fn: L.check(fn(C.byref(a), s), "conv_gemm"), meta))
        self.n_launch += 1
        return y

    def dw(self, x, wdw, y, **kw):
        wdw = self.dev(wdw)
        self.keep.append(wdw)
        if kw.get("shift") is not None:
            kw["shift"] = self.dev(kw["shift"])
            self.keep.append(kw["shift"])
        a = ops.make_dw_args(x, wdw, y, **kw)
        fn = self.lib.segb200_dwconv3x3
        meta = dict(kind="dwconv3x3", flops=2.0 * 9 * a.n * a.ho * a.wo * a.c,
                    bytes=2.0 * (a.n * a.h * a.w * a.c + a.n * a.ho * a.wo * a.c),
                    desc=f"dw3x3 s{a.stride}d{a.dilation} c{a.c} @{a.n}x{a.ho}x{a.wo}")
        self.steps.append((lambda s, a=a, fn=fn: L.check(fn(C.byref(a), s), "dwconv3x3"), meta))
        self.n_launch += 1
        return y

    def call(self, name, *args, launches=1, nbytes=0.0):
        fn = getattr(self.lib, name)
        meta = dict(kind=name.replace("segb200_", ""), flops=0.0, bytes=float(nbytes), desc=name)
        self.steps.append((lambda s, fn=fn, args=args, name=name: L.check(fn(*args, s), name), meta))
        self.n_launch += launches

    def run(self):
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for f, _ in self.steps:
            f(s)

    def run_timed(self):
        """Direct (non-graph) replay with a CUDA event pair around every step on the launching stream.
        Returns [(meta, milliseconds)] -- used by bench.py for the per-kernel roofline numbers."""
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(self.steps) + 1)]
        evs[0].record()
        for i, (f, _) in enumerate(self.steps):
            f(s)
            evs[i + 1].record()
        torch.cuda.synchronize()
        return [(m, evs[i].elapsed_time(evs[i + 1])) for i, (_, m) in enumerate(self.steps)]

    # ---- composite layers -----------------------------------------------------------------
    def conv_bn_act(self, x, prefix, cout, k=1, stride=1, dilation=1, pad=0, act="relu", eps=1e-5, out=None,
                    conv="conv", bn="bn", residual=None, bias=False):
        """_ConvBNReLU / _ConvBN / bare conv (modules/basic.py:65-105)."""
        n, h, w_, cin = x.shape
        wname = f"{prefix}.{conv}" if conv else prefix
        wt = self.w(wname + ".weight")
        ho = (h + 2 * pad - dilation * (k - 1) - 1) // stride + 1
        wo = (w_ + 2 * pad - dilation * (k - 1) - 1) // stride + 1
        cop = fold.round_up(cout, 8)
        if bn is not None:
            scale, shift = self.bn(f"{prefix}.{bn}", eps)
            scale, shift = fold.pad_vec(scale, cop, 1.0), fold.pad_vec(shift, cop)
        else:
            scale = None
            shift = fold.pad_vec(self.w(wname + ".bias"), cop) if bias else None
        wpk = fold.pack_conv_weight(wt, self.dtype, cop)
        if out is None:
            out = self.new(n, ho, wo, cop)
        return self.conv(x, wpk, out, cin=cin, cout=cop, kh=k, kw=k, stride=stride, dilation=dilation, pad_t=pad,
                         pad_l=pad, scale=scale, shift=shift, act=act, residual=residual)

    def sepconv(self, x, prefix, planes, stride=1, dilation=1, relu_first=True, eps=1e-5, out=None, residual=None,
                x_is_relu=False, out_relu=False):
        """SeparableConv2d (modules/basic.py:34-62) = dw kernel + GEMM.  ``x_is_relu``: the producer already applied the leading ReLU
        of a relu_first block (its only consumer is this block: xception.py:37-42), so the depthwise kernel skips it; ``out_relu``:
        apply the NEXT block's leading ReLU in this GEMM's epilogue (free there; round(max(v,0)) == max(round(v),0), bit-identical)."""
        n, h, w_, c = x.shape
        p = prefix + ".block"
        s1, t1 = self.bn(p + ".bn_depth", eps)
        wdw = fold.pack_dw_weight(self.w(p + ".depthwise.weight"), s1)
        ho, wo = (h - 1) // stride + 1, (w_ - 1) // stride + 1
        tmp = self.new(n, ho, wo, c)
        self.dw(x, wdw, tmp, stride=stride, dilation=dilation, shift=t1, pre_relu=relu_first and not x_is_relu,
                act=None if relu_first else "relu")
        s2, t2 = self.bn(p + ".bn_point", eps)
        wpk = fold.pack_conv_weight(self.w(p + ".pointwise.weight"), self.dtype)
        if out is None:
            out = self.new(n, ho, wo, planes)
        assert not (out_relu and (residual is not None or not relu_first))
        return self.conv(tmp, wpk, out, cin=c, cout=planes, scale=s2, shift=t2, act="relu" if (out_relu or not relu_first) else None,
                         residual=residual)

    def stem_s2d(self, x_nchw_shape, x_holder, prefix_conv, prefix_bn, cout, k, pad, act, eps):
        """Stride-2 kxk stem conv + BN + act via space-to-depth packing + tensor-core GEMM."""
        n, cin, h, w_ = x_nchw_shape
        wpk, T, pad2, ld = fold.pack_stem_s2d(self.w(prefix_conv + ".weight"), pad, self.dtype)
        hs, ws = (h + 1) // 2, (w_ + 1) // 2
        s2d = self.new(n, hs, ws, ld)
        meta = dict(kind="pack_s2d", flops=0.0, bytes=float(n * cin * h * w_ * 4 + n * hs * ws * ld * 2), desc="pack_s2d")
        self.steps.append((lambda s, holder=x_holder, s2d=s2d: ops.pack_s2d(holder["x"], s2d), meta))
        self.n_launch += 1
        ho = (h + 2 * pad - k) // 2 + 1
        wo = (w_ + 2 * pad - k) // 2 + 1
        scale, shift = self.bn(prefix_bn, eps)
        out = self.new(n, ho, wo, cout)
        return self.conv(s2d, wpk, out, cin=ld, cout=cout, kh=T, kw=T, stride=1, dilation=1, pad_t=pad2, pad_l=pad2,
                         scale=scale, shift=shift, act=act)


# ----------------------------------------------------------------------------------------------
# model builders (mirror the reference forward graphs; see oracle/segref.py for the line citations)
# ----------------------------------------------------------------------------------------------
def _xception_block(pl, x, prefix, chans, stride=1, dilation=1, skip="conv", relu_first=True, eps=1e-5, sc2_used=True):
    """XceptionBlock.forward (backbones/xception.py:32-51).  In a relu_first block sep_conv1's output feeds only sep_conv2's leading
    ReLU, and sep_conv2's only sep_conv3's unless the caller also takes it (``sc2_used``: the decoder's low-level feature): those
    ReLUs move into the producing GEMM's epilogue."""
    n, h, w_, cin = x.shape
    f1 = relu_first
    f2 = relu_first and not sc2_used
    sc1 = pl.sepconv(x, prefix + ".sep_conv1", chans[1], 1, dilation, relu_first, eps, out_relu=f1)
    sc2 = pl.sepconv(sc1, prefix + ".sep_conv2", chans[2], 1, dilation, relu_first, eps, x_is_relu=f1, out_relu=f2)
    if skip == "conv":
        res = pl.conv_bn_act(x, prefix, chans[3], 1, stride=stride, act=None, eps=eps, conv="conv", bn="bn")
    elif skip == "sum":
        res = x
    else:
        res = None
    out = pl.sepconv(sc2, prefix + ".sep_conv3", chans[3], stride, dilation, relu_first, eps, residual=res, x_is_relu=f2)
    return out, (sc2 if sc2_used else None)


def _xception65(pl, x_shape, holder, output_stride, eps):
    """Xception65.forward (backbones/xception.py:129-165)."""
    b3s, mid_d, exit_d, exit_s = {32: (2, 1, (1, 1), 2), 16: (2, 1, (1, 2), 1), 8: (1, 2, (2, 4), 1)}[output_stride]
    p = "encoder"
    x = pl.stem_s2d(x_shape, holder, p + ".conv1", p + ".bn1", 32, 3, 1, "relu", eps)
    x = pl.conv_bn_act(x, p, 64, 3, pad=1, act="relu", eps=eps, conv="conv2", bn="bn2")
    x, _ = _xception_block(pl, x, p + ".block1", [64, 128, 128, 128], 2, eps=eps, sc2_used=False)
    x, c1 = _xception_block(pl, x, p + ".block2", [128, 256, 256, 256], 2, eps=eps)
    x, c2 = _xception_block(pl, x, p + ".block3", [256, 728, 728, 728], b3s, eps=eps)
    for i in range(4, 20):
        x, _ = _xception_block(pl, x, f"{p}.block{i}", [728] * 4, 1, mid_d, "sum", eps=eps, sc2_used=False)
    c3 = x
    x, _ = _xception_block(pl, c3, p + ".block20", [728, 728, 1024, 1024], exit_s, exit_d[0], eps=eps, sc2_used=False)
    c4, _ = _xception_block(pl, x, p + ".block21", [1024, 1536, 1536, 2048], 1, exit_d[1], "none", False, eps)
    return c1, c2, c3, c4


def _inverted_residual(pl, x, prefix, cout, stride, expand, dilation, eps):
    """InvertedResidual (modules/basic.py:139-163)."""
    n, h, w_, cin = x.shape
    inter = int(round(cin * expand))
    y, i = x, 0
    if expand != 1:
        y = pl.conv_bn_act(y, f"{prefix}.conv.{i}", inter, 1, act="relu6", eps=eps)
        i += 1
    s, t = pl.bn(f"{prefix}.conv.{i}.bn", eps)
    wdw = fold.pack_dw_weight(pl.w(f"{prefix}.conv.{i}.conv.weight"), s)
    ho, wo = (h - 1) // stride + 1, (w_ - 1) // stride + 1
    z = pl.new(n, ho, wo, inter)
    pl.dw(y, wdw, z, stride=stride, dilation=dilation, shift=t, pre_relu=False, act="relu6")
    i += 1
    res = x if (stride == 1 and cin == cout) else None
    # pw-linear: conv at index i, its BN at index i+1 of the nn.Sequential
    wt = pl.w(f"{prefix}.conv.{i}.weight")
    s2, t2 = pl.bn(f"{prefix}.conv.{i + 1}", eps)
    out = pl.new(n, ho, wo, cout)
    return pl.conv(z, fold.pack_conv_weight(wt, pl.dtype), out, cin=inter, cout=cout, scale=s2, shift=t2, act=None,
                   residual=res)


def _mobilenet_v2(pl, x_shape, holder, output_stride, eps):
    """MobileNetV2.forward (backbones/mobilenet.py:131-143), incl. the first-block-only dilation quirk."""
    dil = {32: (1, 1), 16: (1, 2), 8: (2, 4)}[output_stride]
    setting = [[1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1], [6, 160, 3, 2], [6, 320, 1, 1]]
    p = "encoder"
    x = pl.stem_s2d(x_shape, holder, p + ".conv1.conv", p + ".conv1.bn", 32, 3, 1, "relu6", eps)

    def layer(x, name, rows, dilation=1):
        j = 0
        for t, c, nrep, s in rows:
            stride = s if dilation == 1 else 1
            x = _inverted_residual(pl, x, f"{p}.{name}.{j}", c, stride, t, dilation, eps); j += 1
            for _ in range(nrep - 1):
                x = _inverted_residual(pl, x, f"{p}.{name}.{j}", c, 1, t, 1, eps); j += 1
        return x

    x = layer(x, "block1", setting[0:1])
    c1 = layer(x, "block2", setting[1:2])
    c2 = layer(c1, "block3", setting[2:3])
    c3 = layer(c2, "block4", setting[3:5], dil[0])
    c4 = layer(c3, "block5", setting[5:], dil[1])
    return c1, c2, c3, c4


def _bottleneck(pl, x, prefix, planes, stride, dilation, downsample, eps, out=None):
    """BottleneckV1b.forward (backbones/resnet.py:60-81): three GEMMs; the identity / downsample branch is the
    residual operand of the third GEMM's epilogue, followed by the fused ReLU (out += identity; relu)."""
    y = pl.conv_bn_act(x, prefix, planes, 1, act="relu", eps=eps, conv="conv1", bn="bn1")
    y = pl.conv_bn_act(y, prefix, planes, 3, stride=stride, dilation=dilation, pad=dilation, act="relu", eps=eps,
                       conv="conv2", bn="bn2")
    if downsample:
        idn = pl.conv_bn_act(x, prefix + ".downsample", planes * 4, 1, stride=stride, act=None, eps=eps, conv="0", bn="1")
    else:
        idn = x
    return pl.conv_bn_act(y, prefix, planes * 4, 1, act="relu", eps=eps, conv="conv3", bn="bn3", residual=idn, out=out)


def _resnet(pl, x_shape, holder, layers, output_stride, eps, multi_grid=False, multi_dilation=None, c4_out=None):
    """ResNetV1.forward (backbones/resnet.py:183-199): 7x7/2 stem (space-to-depth -> 4x4 tensor-core conv) + BN + ReLU,
    MaxPool(3,2,1), four bottleneck stages with the reference's stride/dilation table (:90-100, :149-179)."""
    dil, strides = {32: ((1, 1), (2, 2)), 16: ((1, 2), (2, 1)), 8: ((2, 4), (1, 1))}[output_stride]
    p = "encoder"
    x = pl.stem_s2d(x_shape, holder, p + ".conv1", p + ".bn1", 64, 7, 3, "relu", eps)
    n, h, w_, c = x.shape
    y = pl.new(n, (h - 1) // 2 + 1, (w_ - 1) // 2 + 1, c)
    pl.call("segb200_maxpool3x3s2", ops._ptr(x), ops._ptr(y), n, h, w_, c, x.stride(2), y.stride(2), ops.dt_code(pl.dtype),
            nbytes=2.0 * (x.numel() + y.numel()))
    x = y
    inplanes = [64]

    def make_layer(x, name, planes, blocks, stride=1, dilation=1, mg=False, last_out=None):
        ds = stride != 1 or inplanes[0] != planes * 4
        first_d = (1 if dilation in (1, 2) else 2) if not mg else multi_dilation[0]
        x = _bottleneck(pl, x, f"{p}.{name}.0", planes, stride, first_d, ds, eps)
        inplanes[0] = planes * 4
        for i in range(1, blocks):
            d = multi_dilation[i % len(multi_dilation)] if mg else dilation
            x = _bottleneck(pl, x, f"{p}.{name}.{i}", planes, 1, d, False, eps, out=last_out if i == blocks - 1 else None)
        return x

    c1 = make_layer(x, "layer1", 64, layers[0])
    c2 = make_layer(c1, "layer2", 128, layers[1], 2)
    c3 = make_layer(c2, "layer3", 256, layers[2], strides[0], dil[0])
    c4 = make_layer(c3, "layer4", 512, layers[3], strides[1], dil[1], multi_grid, last_out=c4_out)
    return c1, c2, c3, c4


def _aspp(pl, x, prefix, output_stride):
    """_ASPP.forward (modules/module.py:62-77); concat order [pool, aspp0, aspp1, aspp2, aspp3]."""
    d = {16: (6, 12, 18), 8: (12, 24, 36), 32: (6, 12, 18)}[output_stride]
    n, h, w_, c = x.shape
    cat = pl.new(n, h, w_, 1280)
    # image pooling branch: GAP -> 1x1 conv+BN+ReLU on n "pixels" -> broadcast (bilinear from 1x1)
    pooled = pl.new(n, 1, 1, c)
    pl.call("segb200_global_avgpool", ops._ptr(x), ops._ptr(pooled), n, h, w_, c, x.stride(2), ops.dt_code(pl.dtype))
    pf = pl.conv_bn_act(pooled, prefix + ".image_pooling", 256, 1, act="relu")
    pl.call("segb200_bilinear_nhwc", ops._ptr(pf), ops._ptr(cat[..., 0:256]), n, 1, 1, 256, pf.stride(2), h, w_,
            cat.stride(2), 1, ops.dt_code(pl.dtype))
    pl.conv_bn_act(x, prefix + ".aspp0", 256, 1, act="relu", out=cat[..., 256:512])
    for i in range(3):
        pl.sepconv(x, f"{prefix}.aspp{i + 1}", 256, 1, d[i], relu_first=False, out=cat[..., 512 + 256 * i:768 + 256 * i])
    return pl.conv_bn_act(cat, prefix, 256, 1, act="relu")          # Dropout2d: identity in eval (:75)


def build_deeplabv3plus(pl, x_shape, holder, backbone, nclass, output_stride, eps_encoder, use_aspp, use_decoder,
                        out_dtype, want_argmax):
    """DeepLabV3Plus.forward (models/deeplabv3_plus.py:33-46) + _DeepLabHead (:66-75)."""
    n, _, H, W = x_shape
    if backbone == "xception65":
        c1, _, _, c4 = _xception65(pl, x_shape, holder, output_stride, eps_encoder)
    elif backbone == "mobilenet_v2":
        c1, _, _, c4 = _mobilenet_v2(pl, x_shape, holder, output_stride, eps_encoder)
    elif backbone in ("resnet50", "resnet101", "resnet152"):
        layers = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3), "resnet152": (3, 8, 36, 3)}[backbone]
        c1, _, _, c4 = _resnet(pl, x_shape, holder, layers, output_stride, eps_encoder)
    else:
        raise RuntimeError(f"segb200: backbone '{backbone}' has no B200 plan yet")
    x = c4
    if use_aspp:
        x = _aspp(pl, x, "head.aspp", output_stride)
    if use_decoder:
        _, h1, w1, _ = c1.shape
        cat = pl.new(n, h1, w1, 256 + 48)
        pl.call("segb200_bilinear_nhwc", ops._ptr(x), ops._ptr(cat[..., 0:256]), n, x.shape[1], x.shape[2], 256,
                x.stride(2), h1, w1, cat.stride(2), 1, ops.dt_code(pl.dtype))
        pl.conv_bn_act(c1, "head.c1_block", 48, 1, act="relu", out=cat[..., 256:304])
        x = cat
    x = pl.sepconv(x, "head.block.0", 256, relu_first=False)
    x = pl.sepconv(x, "head.block.1", 256, relu_first=False)
    logits = pl.new(n, x.shape[1], x.shape[2], fold.round_up(nclass, 8), ld=32)
    pl.conv_bn_act(x, "head.block.2", nclass, 1, act=None, conv=None, bn=None, bias=True, out=logits)
    out = torch.empty(n, nclass, H, W, dtype=out_dtype, device=pl.device)
    amax = torch.empty(n, H, W, dtype=torch.uint8, device=pl.device) if want_argmax else None
    pl.call("segb200_bilinear_nchw_out", ops._ptr(logits), ops._ptr(out), ops._ptr(amax), n, logits.shape[1],
            logits.shape[2], nclass, logits.stride(2), H, W, 1, ops.dt_code(pl.dtype), ops.dt_code(out_dtype))
    pl.keep += [out] + ([amax] if amax is not None else [])
    return out, amax, logits


def build_danet(pl, x_shape, holder, nclass, output_stride, multi_grid, multi_dilation, out_dtype, want_argmax):
    """DANet.forward (models/danet.py:26-41) + DANetHead.forward (:70-88).  conv8(sa_conv + sc_conv) is evaluated as
    conv8(sa_conv) + conv8(sc_conv) (linear), the second GEMM taking the first as its residual operand."""
    from . import attention as A
    n, _, H, W = x_shape
    _, _, _, c4 = _resnet(pl, x_shape, holder, (3, 4, 23, 3), output_stride, 1e-5, multi_grid, multi_dilation)
    hp = "head"

    def cbr(x, name):
        return pl.conv_bn_act(x, f"{hp}.{name}", 512, 3, pad=1, act="relu", conv="0", bn="1")

    def qkv(prefix):
        ws = []
        for nm in ("query_conv", "key_conv", "value_conv"):
            ws += [pl.dev(fold.pack_conv_weight(pl.w(f"{prefix}.{nm}.weight"), pl.dtype)), pl.dev(pl.w(f"{prefix}.{nm}.bias").float())]
        return ws

    feat1 = cbr(c4, "conv5a")
    sa_feat = pl.new(*feat1.shape)
    wq, bq, wk, bk, wv, bv = qkv(hp + ".sa")
    g_sa = pl.dev(pl.w(hp + ".sa.gamma").float())
    pl.keep += [wq, bq, wk, bk, wv, bv, g_sa]
    ntok = feat1.shape[1] * feat1.shape[2]
    pl.steps.append((lambda s: A.pam_nhwc(feat1, wq, bq, wk, bk, wv, bv, g_sa, out=sa_feat),
                     dict(kind="pam", flops=2.0 * n * ntok * ntok * (64 + 512), bytes=2.0 * n * ntok * (512 * 3 + 128),
                          desc=f"PAM N={ntok} d=64 dv=512 b={n}")))
    pl.n_launch += 4 + n
    sa_conv = cbr(sa_feat, "conv51")
    feat2 = cbr(c4, "conv5c")
    sc_feat = pl.new(*feat2.shape)
    g_sc = pl.dev(pl.w(hp + ".sc.gamma").float())
    pl.keep.append(g_sc)
    pl.steps.append((lambda s: A.cam_nhwc(feat2, g_sc, out=sc_feat),
                     dict(kind="cam", flops=2.0 * 2 * n * ntok * 512 * 512, bytes=2.0 * n * ntok * 512 * 3, desc=f"CAM C=512 N={ntok} b={n}")))
    pl.n_launch += 4 * n
    sc_conv = cbr(sc_feat, "conv52")

    def classifier(x, name, residual=None, bias=True):
        lg = pl.new(n, x.shape[1], x.shape[2], fold.round_up(nclass, 8), ld=32)
        cop = fold.round_up(nclass, 8)
        wpk = fold.pack_conv_weight(pl.w(f"{hp}.{name}.1.weight"), pl.dtype, cop)
        shift = fold.pad_vec(pl.w(f"{hp}.{name}.1.bias"), cop) if bias else None
        return pl.conv(x, wpk, lg, cin=512, cout=cop, shift=shift, residual=residual)

    sa_out = classifier(sa_conv, "conv6")
    sc_out = classifier(sc_conv, "conv7")
    t = classifier(sa_conv, "conv8", bias=False)
    sasc_out = classifier(sc_conv, "conv8", residual=t)
    outs = []
    for lg in (sasc_out, sa_out, sc_out):
        o = torch.empty(n, nclass, H, W, dtype=out_dtype, device=pl.device)
        am = torch.empty(n, H, W, dtype=torch.uint8, device=pl.device) if (want_argmax and lg is sasc_out) else None
        pl.call("segb200_bilinear_nchw_out", ops._ptr(lg), ops._ptr(o), ops._ptr(am), n, lg.shape[1], lg.shape[2], nclass,
                lg.stride(2), H, W, 1, ops.dt_code(pl.dtype), ops.dt_code(out_dtype))
        pl.keep += [o] + ([am] if am is not None else [])
        outs.append((o, am))
    return outs


HRNET_W18_SMALL_V1 = dict(                     # configs/cityscapes_hrnet_w18_small_v1.yaml:23-64
    stage1=dict(block="BOTTLENECK", blocks=[1], channels=[32]),
    stage2=dict(modules=1, block="BASIC", blocks=[2, 2], channels=[16, 32]),
    stage3=dict(modules=1, block="BASIC", blocks=[2, 2, 2], channels=[16, 32, 64]),
    stage4=dict(modules=1, block="BASIC", blocks=[2, 2, 2, 2], channels=[16, 32, 64, 128]),
    final_conv_kernel=1)


def _hr_module(pl, xs, prefix, blocks, channels):
    """HighResolutionModule.forward (backbones/hrnet.py:215-232).  BasicBlocks are two GEMM convs (residual + ReLU fused);
    every fuse sum is built incrementally: down paths add the running sum as the residual of their last 3x3/2 GEMM, up paths
    run the 1x1 GEMM at low resolution and one nearest-up + add (+ final ReLU) kernel."""
    nb = len(xs)
    xs = list(xs)
    for i in range(nb):
        for b in range(blocks[i]):
            p = f"{prefix}.branches.{i}.{b}"
            y = pl.conv_bn_act(xs[i], p, channels[i], 3, pad=1, act="relu", conv="conv1", bn="bn1")
            xs[i] = pl.conv_bn_act(y, p, channels[i], 3, pad=1, act="relu", conv="conv2", bn="bn2", residual=xs[i])
    if nb == 1:
        return xs
    dtc = ops.dt_code(pl.dtype)
    outs = []
    for i in range(nb):
        n, h, w_, c = xs[i].shape
        acc = None
        for j in range(nb):
            last = j == nb - 1
            act = "relu" if last else None
            if j == i:
                if acc is None:
                    acc = xs[j]                      # first term (i == 0): nothing to add yet; nb > 1 so it is never last
                else:
                    out = pl.new(n, h, w_, c)
                    pl.call("segb200_upsample_add", ops._ptr(acc), ops._ptr(xs[j]), ops._ptr(out), n, h, w_, c, acc.stride(2),
                            xs[j].stride(2), out.stride(2), 0, L.ACT[act], dtc)
                    acc = out
            elif j > i:
                z = pl.conv_bn_act(xs[j], f"{prefix}.fuse_layers.{i}.{j}", channels[i], 1, act=None, conv="0", bn="1")
                out = pl.new(n, h, w_, c)
                pl.call("segb200_upsample_add", ops._ptr(acc), ops._ptr(z), ops._ptr(out), n, h, w_, c, acc.stride(2), z.stride(2),
                        out.stride(2), j - i, L.ACT[act], dtc)
                acc = out
            else:
                t = xs[j]
                for k in range(i - j):
                    lastk = k == i - j - 1
                    co = channels[i] if lastk else channels[j]
                    t = pl.conv_bn_act(t, f"{prefix}.fuse_layers.{i}.{j}.{k}", co, 3, stride=2, pad=1,
                                       act=(act if lastk else "relu"), conv="0", bn="1", residual=acc if lastk else None)
                acc = t
        outs.append(acc)
    return outs


def build_hrnet(pl, x_shape, holder, nclass, hcfg, out_dtype, want_argmax):
    """HighResolutionNet.forward (backbones/hrnet.py:429-479) + _HRNetHead (models/hrnet_seg.py:54-63) + final bilinear
    (align_corners=False, hrnet_seg.py:28).  The three head up-samplings write their channel slices of the cat buffer."""
    n, _, H, W = x_shape
    p = "encoder"
    x = pl.stem_s2d(x_shape, holder, p + ".conv1", p + ".bn1", 64, 3, 1, "relu", 1e-5)
    x = pl.conv_bn_act(x, p, 64, 3, stride=2, pad=1, act="relu", conv="conv2", bn="bn2")
    planes = hcfg["stage1"]["channels"][0]
    inpl = 64
    for b in range(hcfg["stage1"]["blocks"][0]):
        x = _bottleneck(pl, x, f"{p}.layer1.{b}", planes, 1, 1, inpl != planes * 4, 1e-5)
        inpl = planes * 4
    pre, ys = [inpl], [x]
    for si, sname in enumerate(("stage2", "stage3", "stage4")):
        sc = hcfg[sname]
        cur = sc["channels"]
        tname = f"{p}.transition{si + 1}"
        xs = []
        for i in range(len(cur)):
            if i < len(pre):
                xs.append(ys[i] if cur[i] == pre[i] else
                          pl.conv_bn_act(ys[i], f"{tname}.{i}", cur[i], 3, pad=1, act="relu", conv="0", bn="1"))
            else:
                t = ys[-1]
                for j in range(i + 1 - len(pre)):
                    co = cur[i] if j == i - len(pre) else pre[-1]
                    t = pl.conv_bn_act(t, f"{tname}.{i}.{j}", co, 3, stride=2, pad=1, act="relu", conv="0", bn="1")
                xs.append(t)
        for m in range(sc["modules"]):
            xs = _hr_module(pl, xs, f"{p}.{sname}.{m}", sc["blocks"], cur)
        ys, pre = xs, cur
    _, h0, w0, c0 = ys[0].shape
    ctot = sum(t.shape[3] for t in ys)
    cat = pl.new(n, h0, w0, ctot)
    # branch 0 is copied by a k=0 "add" against itself? no: bilinear to the same size is the identity -> use the resize kernel
    off = 0
    for t in ys:
        c = t.shape[3]
        pl.call("segb200_bilinear_nhwc", ops._ptr(t), ops._ptr(cat[..., off:off + c]), n, t.shape[1], t.shape[2], c, t.stride(2),
                h0, w0, cat.stride(2), 0, ops.dt_code(pl.dtype))
        off += c
    hp = "hrnet_head.last_layer"
    # conv(+bias) -> BN -> ReLU : fold the conv bias into the BN shift
    sc_, sh_ = pl.bn(hp + ".1", 1e-5)
    sh_ = sh_ + pl.w(hp + ".0.bias").float() * sc_
    y = pl.new(n, h0, w0, ctot)
    pl.conv(cat, fold.pack_conv_weight(pl.w(hp + ".0.weight"), pl.dtype), y, cin=ctot, cout=ctot, scale=sc_, shift=sh_.contiguous(),
            act="relu")
    k = hcfg["final_conv_kernel"]
    cop = fold.round_up(nclass, 8)
    logits = pl.new(n, h0, w0, cop, ld=32)
    pl.conv(y, fold.pack_conv_weight(pl.w(hp + ".3.weight"), pl.dtype, cop), logits, cin=ctot, cout=cop, kh=k, kw=k,
            pad_t=1 if k == 3 else 0, pad_l=1 if k == 3 else 0, shift=fold.pad_vec(pl.w(hp + ".3.bias"), cop))
    o = torch.empty(n, nclass, H, W, dtype=out_dtype, device=pl.device)
    am = torch.empty(n, H, W, dtype=torch.uint8, device=pl.device) if want_argmax else None
    pl.call("segb200_bilinear_nchw_out", ops._ptr(logits), ops._ptr(o), ops._ptr(am), n, h0, w0, nclass, logits.stride(2), H, W, 0,
            ops.dt_code(pl.dtype), ops.dt_code(out_dtype))
    pl.keep += [o] + ([am] if am is not None else [])
    return [(o, am)]


def build_ccnet(pl, x_shape, holder, nclass, output_stride, recurrence, out_dtype, want_argmax):
    """CCNet.forward + _RCCAModule.forward (models/ccnet.py:27-40, :73-82).  The backbone's last block and convb write
    their channel slices of one 2560-channel buffer (no torch.cat); the recurrent criss-cross attention shares weights."""
    from . import attention as A
    n, _, H, W = x_shape
    # the concat buffer must exist before the backbone is planned: c4 is produced straight into cat[..., :2048]
    ho = wo = None
    hh, ww = H, W
    for s in ([2, 2, 2] + ([2] if output_stride >= 16 else []) + ([2] if output_stride == 32 else [])):
        hh, ww = (hh - 1) // s + 1, (ww - 1) // s + 1
    cat = pl.new(n, hh, ww, 2048 + 512)
    _, _, _, c4 = _resnet(pl, x_shape, holder, (3, 4, 23, 3), output_stride, 1e-5, c4_out=cat[..., :2048])
    assert tuple(c4.shape[1:3]) == (hh, ww), (c4.shape, hh, ww)
    hp = "head.rcca"
    out = pl.conv_bn_act(c4, hp + ".conva", 512, 3, pad=1, act="relu", conv="0", bn="1")
    ws = []
    for nm in ("query_conv", "key_conv", "value_conv"):
        ws += [pl.dev(fold.pack_conv_weight(pl.w(f"{hp}.cca.{nm}.weight"), pl.dtype)), pl.dev(pl.w(f"{hp}.cca.{nm}.bias").float())]
    g = pl.dev(pl.w(hp + ".cca.gamma").float())
    pl.keep += ws + [g]
    bufs = [pl.new(*out.shape) for _ in range(recurrence)]
    cur = out
    for i in range(recurrence):
        pl.steps.append((lambda s, src=cur, dst=bufs[i]: A.cca_nhwc(src, *ws, g, out=dst),
                         dict(kind="cca", flops=2.0 * n * hh * ww * (hh + ww - 1) * (64 + 512), bytes=2.0 * n * hh * ww * 512 * 4,
                              desc=f"CCA {hh}x{ww} c512 b={n}")))
        pl.n_launch += 5
        cur = bufs[i]
    pl.conv_bn_act(cur, hp + ".convb", 512, 3, pad=1, act="relu", conv="0", bn="1", out=cat[..., 2048:2560])
    y = pl.conv_bn_act(cat, hp + ".bottleneck", 512, 3, pad=1, act=None, conv="0", bn="1")
    logits = pl.new(n, hh, ww, fold.round_up(nclass, 8), ld=32)
    pl.conv_bn_act(y, "head.out", nclass, 1, act=None, conv=None, bn=None, bias=True, out=logits)
    o = torch.empty(n, nclass, H, W, dtype=out_dtype, device=pl.device)
    am = torch.empty(n, H, W, dtype=torch.uint8, device=pl.device) if want_argmax else None
    pl.call("segb200_bilinear_nchw_out", ops._ptr(logits), ops._ptr(o), ops._ptr(am), n, hh, ww, nclass, logits.stride(2), H, W, 1,
            ops.dt_code(pl.dtype), ops.dt_code(out_dtype))
    pl.keep += [o] + ([am] if am is not None else [])
    return [(o, am)]


def build_pspnet(pl, x_shape, holder, nclass=19, output_stride=8, out_dtype=None, want_argmax=False):
    """PSPNet.forward + _PSPHead.forward (models/pspnet.py:30-61; `forward()[0]`, the output SegBaseModel.evaluate uses).  The
    backbone's last block writes channels [0, 2048) of the 4096-channel pyramid buffer and every pooled branch (adaptive average pool
    -> 1x1 conv + BN + ReLU GEMM -> bilinear, PyramidPooling, modules/module.py:82-97) writes its 512-channel slice: no torch.cat.
    Then the 3x3 4096 -> 512 conv + BN + ReLU (the head's dominant GEMM, K = 36 864), Dropout(0.1) = identity in eval, the 1x1
    classifier and the fused bilinear NCHW output (+ argmax)."""
    n, _, H, W = x_shape
    hh, ww = H, W
    for s in ([2, 2, 2] + ([2] if output_stride >= 16 else []) + ([2] if output_stride == 32 else [])):
        hh, ww = (hh - 1) // s + 1, (ww - 1) // s + 1
    sizes = (1, 2, 3, 6)
    cat = pl.new(n, hh, ww, 2048 + 512 * len(sizes))
    _, _, _, c4 = _resnet(pl, x_shape, holder, (3, 4, 23, 3), output_stride, 1e-5, c4_out=cat[..., :2048])
    assert tuple(c4.shape[1:3]) == (hh, ww), (c4.shape, hh, ww)
    for i, s in enumerate(sizes):
        pooled = pl.new(n, s, s, 2048)
        pl.call("segb200_adaptive_avgpool", ops._ptr(c4), ops._ptr(pooled), n, hh, ww, 2048, c4.stride(2), s, pooled.stride(2),
                ops.dt_code(pl.dtype), nbytes=2.0 * c4.numel())
        f = pl.conv_bn_act(pooled, f"head.psp.convs.{i}", 512, 1, act="relu")
        pl.call("segb200_bilinear_nhwc", ops._ptr(f), ops._ptr(cat[..., 2048 + 512 * i:2560 + 512 * i]), n, s, s, 512, f.stride(2), hh, ww,
                cat.stride(2), 1, ops.dt_code(pl.dtype), nbytes=2.0 * n * hh * ww * 512)
    y = pl.conv_bn_act(cat, "head.block", 512, 3, pad=1, act="relu", conv="0", bn="1")
    logits = pl.new(n, hh, ww, fold.round_up(nclass, 8), ld=32)
    pl.conv_bn_act(y, "head.block.4", nclass, 1, act=None, conv=None, bn=None, bias=True, out=logits)
    out_dtype = out_dtype or pl.dtype
    o = torch.empty(n, nclass, H, W, dtype=out_dtype, device=pl.device)
    am = torch.empty(n, H, W, dtype=torch.uint8, device=pl.device) if want_argmax else None
    pl.call("segb200_bilinear_nchw_out", ops._ptr(logits), ops._ptr(o), ops._ptr(am), n, hh, ww, nclass, logits.stride(2), H, W, 1,
            ops.dt_code(pl.dtype), ops.dt_code(out_dtype))
    pl.keep += [o] + ([am] if am is not None else [])
    return [(o, am)]


def build_ocnet(pl, x_shape, holder, nclass=19, output_stride=16, layers=(3, 4, 6, 3), out_dtype=None, want_argmax=False):
    """OCNet.forward + _OCHead 'base' (models/ocnet.py:30-54; cfg.MODEL.OCNet.OC_ARCH = 'base', configs/cityscapes_ocnet.yaml:
    ResNet50, OS16): conv3x3(2048 -> 512) + BN + ReLU -> BaseOCModule (:116-141) = BaseAttentionBlock (:72-113) on the tensor-core
    attention kernel + its output conv W, then project(cat[context, feats]) -- evaluated as two GEMMs, the feats half (pre-multiplied by
    the folded BN scale) being the residual operand of the context half (the 1x1 conv is linear; no concat buffer, and feats stays
    contiguous for the V^T GEMM) -- + BN + ReLU (Dropout2d(0.05) = identity in eval) -> 1x1 classifier -> fused bilinear NCHW output."""
    from . import attention as A
    n, _, H, W = x_shape
    _, _, _, c4 = _resnet(pl, x_shape, holder, layers, output_stride, 1e-5)
    hh, ww = c4.shape[1], c4.shape[2]
    feats = pl.conv_bn_act(c4, "head.context", 512, 3, pad=1, act="relu", conv="0", bn="1")
    bp = "head.context.3.stages.0"
    key_ch, val_ch = pl.w(bp + ".f_key.0.weight").shape[0], pl.w(bp + ".f_value.weight").shape[0]
    s = float(key_ch) ** -0.5
    ksc, ksh = pl.bn(bp + ".f_key.1", 1e-5)
    ksh = ksh + ksc * pl.w(bp + ".f_key.0.bias").float()                            # conv bias through the folded BN
    wk = pl.w(bp + ".f_key.0.weight")
    wqk = pl.dev(fold.pack_conv_weight(torch.cat([wk, wk], 0), pl.dtype))
    scale_qk = pl.dev(torch.cat([ksc * s, ksc]).float())                            # ReLU(s z) = s ReLU(z), s > 0
    shift_qk = pl.dev(torch.cat([ksh * s, ksh]).float())
    wv = pl.dev(fold.pack_conv_weight(pl.w(bp + ".f_value.weight"), pl.dtype))
    bv = pl.dev(pl.w(bp + ".f_value.bias").float())
    pl.keep += [wqk, scale_qk, shift_qk, wv, bv]
    ctx = pl.new(n, hh, ww, val_ch)
    ntok = hh * ww
    pl.steps.append((lambda s_: A.nonlocal_nhwc(feats, wqk, scale_qk, shift_qk, wv, bv, key_ch, ctx),
                     dict(kind="nonlocal", flops=2.0 * n * ntok * ntok * (key_ch + val_ch), bytes=2.0 * n * ntok * (512 + 2 * key_ch + 2 * val_ch),
                          desc=f"non-local attention N={ntok} d={key_ch} dv={val_ch} b={n}")))
    pl.n_launch += 3 + n
    cw = pl.conv_bn_act(ctx, bp + ".W", 512, 1, act=None, conv=None, bn=None, bias=True)
    # project: BN(W_p [context | feats] + b_p) = scale (W_a context) + [scale (W_b feats)] + (scale b_p + shift)
    pp = "head.context.3.project"
    psc, psh = pl.bn(pp + ".1", 1e-5)
    psh = psh + psc * pl.w(pp + ".0.bias").float()
    wp = pl.w(pp + ".0.weight")
    t = pl.new(n, hh, ww, 512)
    pl.conv(feats, fold.pack_conv_weight(wp[:, 512:].contiguous(), pl.dtype), t, cin=512, cout=512, scale=psc.float())
    y = pl.new(n, hh, ww, 512)
    pl.conv(cw, fold.pack_conv_weight(wp[:, :512].contiguous(), pl.dtype), y, cin=512, cout=512, scale=psc.float(), shift=psh.float(),
            act="relu", residual=t)
    logits = pl.new(n, hh, ww, fold.round_up(nclass, 8), ld=32)
    pl.conv_bn_act(y, "head.out", nclass, 1, act=None, conv=None, bn=None, bias=True, out=logits)
    out_dtype = out_dtype or pl.dtype
    o = torch.empty(n, nclass, H, W, dtype=out_dtype, device=pl.device)
    am = torch.empty(n, H, W, dtype=torch.uint8, device=pl.device) if want_argmax else None
    pl.call("segb200_bilinear_nchw_out", ops._ptr(logits), ops._ptr(o), ops._ptr(am), n, hh, ww, nclass, logits.stride(2), H, W, 1,
            ops.dt_code(pl.dtype), ops.dt_code(out_dtype))
    pl.keep += [o] + ([am] if am is not None else [])
    return [(o, am)]


class DeepLabV3PlusB200:
    """Inference engine: ``engine(x_nchw) -> logits [N, nclass, H, W]`` (same contract as
    ``DeepLabV3Plus.forward(x)[0]``, models/deeplabv3_plus.py:33-46)."""

    def __init__(self, state_dict, backbone="xception65", nclass=19, output_stride=16, eps_encoder=1e-5,
                 use_aspp=True, use_decoder=True, dtype=torch.bfloat16, out_dtype=None, device="cuda",
                 cuda_graph=True, want_argmax=False):
        if not torch.cuda.is_available():
            raise RuntimeError("segb200: a CUDA device (sm_100a) is required; there is no CPU fallback")
        L.load()
        self.sd = {k: v.detach() for k, v in state_dict.items()}
        self.cfg = dict(backbone=backbone, nclass=nclass, output_stride=output_stride, eps_encoder=eps_encoder,
                        use_aspp=use_aspp, use_decoder=use_decoder)
        self.dtype, self.out_dtype = dtype, out_dtype or dtype
        self.device = torch.device(device)
        self.cuda_graph = cuda_graph
        self.want_argmax = want_argmax
        self.plans = {}

    def _build(self, shape, in_dtype):
        holder = {"x": torch.empty(shape, dtype=in_dtype, device=self.device)}
        pl = Plan(self.sd, self.dtype, self.device)
        out, amax, logits = build_deeplabv3plus(pl, shape, holder, out_dtype=self.out_dtype,
                                                want_argmax=self.want_argmax, **self.cfg)
        graph = None
        if self.cuda_graph:
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                pl.run()                              # warm-up (module load, attribute set) outside capture
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                pl.run()
        return dict(plan=pl, holder=holder, out=out, amax=amax, logits=logits, graph=graph)

    def plan_for(self, x):
        key = (tuple(x.shape), x.dtype)
        if key not in self.plans:
            self.plans[key] = self._build(tuple(x.shape), x.dtype)
        return self.plans[key]

    def __call__(self, x, copy_input=True):
        if not x.is_cuda:
            raise RuntimeError("segb200: input must be a CUDA tensor (no CPU implementation)")
        st = self.plan_for(x)
        if copy_input:
            st["holder"]["x"].copy_(x)
        if st["graph"] is not None:
            st["graph"].replay()
        else:
            st["plan"].run()
        return st["out"]

    def argmax(self, x):
        st = self.plan_for(x)
        if st["amax"] is None:
            raise RuntimeError("segb200: engine was built without want_argmax=True")
        self(x)
        return st["amax"]


class DANetB200(DeepLabV3PlusB200):
    """``engine(x) -> sasc logits`` (``DANet.forward(x)[0]``, models/danet.py:26-41); ``engine.outputs(x)`` returns all three
    maps (sasc, sa, sc).  ResNet101 backbone at output stride 8 with multi-grid dilations, PAM/CAM on the tensor cores."""

    def __init__(self, state_dict, nclass=19, output_stride=8, multi_grid=True, multi_dilation=(4, 8, 16), dtype=torch.bfloat16,
                 out_dtype=None, device="cuda", cuda_graph=False, want_argmax=False):
        super().__init__(state_dict, backbone="resnet101", nclass=nclass, output_stride=output_stride, dtype=dtype,
                         out_dtype=out_dtype, device=device, cuda_graph=cuda_graph, want_argmax=want_argmax)
        self.mg = (multi_grid, list(multi_dilation))

    def _build(self, shape, in_dtype):
        holder = {"x": torch.empty(shape, dtype=in_dtype, device=self.device)}
        pl = Plan(self.sd, self.dtype, self.device)
        outs = build_danet(pl, shape, holder, self.cfg["nclass"], self.cfg["output_stride"], self.mg[0], self.mg[1],
                           self.out_dtype, self.want_argmax)
        graph = None
        if self.cuda_graph:
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                pl.run()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                pl.run()
        return dict(plan=pl, holder=holder, out=outs[0][0], amax=outs[0][1], all=[o for o, _ in outs], graph=graph, logits=None)

    def outputs(self, x):
        self(x)
        return tuple(self.plan_for(x)["all"])


class CCNetB200(DANetB200):
    """``engine(x) -> logits`` (``CCNet.forward(x)[0]``, models/ccnet.py:27-40): ResNet101 OS16 + recurrent criss-cross
    attention (RECURRENCE = 2)."""

    def __init__(self, state_dict, nclass=19, output_stride=16, recurrence=2, dtype=torch.bfloat16, out_dtype=None,
                 device="cuda", cuda_graph=False, want_argmax=False):
        DeepLabV3PlusB200.__init__(self, state_dict, backbone="resnet101", nclass=nclass, output_stride=output_stride,
                                   dtype=dtype, out_dtype=out_dtype, device=device, cuda_graph=cuda_graph, want_argmax=want_argmax)
        self.recurrence = recurrence

    def _build(self, shape, in_dtype):
        holder = {"x": torch.empty(shape, dtype=in_dtype, device=self.device)}
        pl = Plan(self.sd, self.dtype, self.device)
        outs = build_ccnet(pl, shape, holder, self.cfg["nclass"], self.cfg["output_stride"], self.recurrence, self.out_dtype,
                           self.want_argmax)
        return dict(plan=pl, holder=holder, out=outs[0][0], amax=outs[0][1], all=[o for o, _ in outs], graph=None, logits=None)


class PSPNetB200(DANetB200):
    """``engine(x) -> logits`` (``PSPNet.forward(x)[0]``, models/pspnet.py:30-43): ResNet101 OS8 + PyramidPooling + _PSPHead.  The
    auxiliary _FCNHead is a training-time output (``forward()[1]``) and is not computed."""

    def __init__(self, state_dict, nclass=19, output_stride=8, dtype=torch.bfloat16, out_dtype=None, device="cuda", cuda_graph=False,
                 want_argmax=False):
        DeepLabV3PlusB200.__init__(self, state_dict, backbone="resnet101", nclass=nclass, output_stride=output_stride,
                                   dtype=dtype, out_dtype=out_dtype, device=device, cuda_graph=cuda_graph, want_argmax=want_argmax)

    def _build(self, shape, in_dtype):
        holder = {"x": torch.empty(shape, dtype=in_dtype, device=self.device)}
        pl = Plan(self.sd, self.dtype, self.device)
        outs = build_pspnet(pl, shape, holder, self.cfg["nclass"], self.cfg["output_stride"], self.out_dtype, self.want_argmax)
        return dict(plan=pl, holder=holder, out=outs[0][0], amax=outs[0][1], all=[o for o, _ in outs], graph=None, logits=None)


class OCNetB200(DANetB200):
    """``engine(x) -> logits`` (``OCNet.forward(x)[0]``, models/ocnet.py:30-43) for the 'base' object-context head on ResNet50 at
    output stride 16 (configs/cityscapes_ocnet.yaml); the 'pyramid' and 'asp' variants are not built."""

    def __init__(self, state_dict, nclass=19, output_stride=16, layers=(3, 4, 6, 3), dtype=torch.bfloat16, out_dtype=None, device="cuda",
                 cuda_graph=False, want_argmax=False):
        DeepLabV3PlusB200.__init__(self, state_dict, backbone="resnet50", nclass=nclass, output_stride=output_stride,
                                   dtype=dtype, out_dtype=out_dtype, device=device, cuda_graph=cuda_graph, want_argmax=want_argmax)
        self.layers = tuple(layers)

    def _build(self, shape, in_dtype):
        holder = {"x": torch.empty(shape, dtype=in_dtype, device=self.device)}
        pl = Plan(self.sd, self.dtype, self.device)
        outs = build_ocnet(pl, shape, holder, self.cfg["nclass"], self.cfg["output_stride"], self.layers, self.out_dtype, self.want_argmax)
        return dict(plan=pl, holder=holder, out=outs[0][0], amax=outs[0][1], all=[o for o, _ in outs], graph=None, logits=None)


class HRNetB200(DANetB200):
    """``engine(x) -> logits`` (``HighResolutionNet.forward(x)[0]``, models/hrnet_seg.py:23-29); default = hrnet_w18_small_v1."""

    def __init__(self, state_dict, nclass=19, hcfg=None, dtype=torch.float16, out_dtype=None, device="cuda", cuda_graph=True,
                 want_argmax=False):
        DeepLabV3PlusB200.__init__(self, state_dict, backbone="hrnet", nclass=nclass, dtype=dtype, out_dtype=out_dtype,
                                   device=device, cuda_graph=cuda_graph, want_argmax=want_argmax)
        self.hcfg = hcfg or HRNET_W18_SMALL_V1

    def _build(self, shape, in_dtype):
        holder = {"x": torch.empty(shape, dtype=in_dtype, device=self.device)}
        pl = Plan(self.sd, self.dtype, self.device)
        outs = build_hrnet(pl, shape, holder, self.cfg["nclass"], self.hcfg, self.out_dtype, self.want_argmax)
        graph = None
        if self.cuda_graph:
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                pl.run()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                pl.run()
        return dict(plan=pl, holder=holder, out=outs[0][0], amax=outs[0][1], all=[o for o, _ in outs], graph=graph, logits=None)

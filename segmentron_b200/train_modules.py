"""Training-mode forward/backward of the drop-in building blocks (``modules.SeparableConv2d``, ``_ConvBNReLU``, ``_ConvBN``):
``torch.autograd.Function``s over the C-ABI training kernels, so that a reference model whose L1 classes were swapped
(``patch.install()`` / ``convert_to_b200``) stays differentiable under the unmodified ``tools/train.py`` loop.

One Function = one fused unit of the training plan (``train.py``), in eager form:
  conv (tcgen05 implicit GEMM, raw 16-bit output) -> batch statistics -> normalise + activation;
  backward: BN/activation backward (mask recomputed from y) -> tcgen05 weight gradient -> data gradient on the forward GEMM
  kernel with transposed, tap-reversed weights (zero insertion / strided placement for stride 2).
The depthwise unit is the first half of SeparableConv2d (optional leading ReLU, modules/basic.py:45-46).
``GlobalAvgPoolFunction`` / ``BroadcastFunction`` are the two ends of _ASPP's image-pooling branch (module.py:52,64: a
bilinear up-sampling from 1x1 is a broadcast); each one's backward is the other's forward kernel with a scale.

BatchNorm semantics are torch's (batch statistics, running-stat update with ``momentum``, ``num_batches_tracked``); statistics
are per process -- for SyncBatchNorm across ranks use the whole-model engine (``train.DeepLabV3PlusTrainerB200``).
No CPU / PyTorch fallback: non-CUDA input raises.
"""
import torch

from . import fold, ops, train_ops as T


def _bn_args(bn):
    if bn.momentum is None:
        raise RuntimeError("segb200: BatchNorm momentum=None (cumulative average) is not supported by the training kernels")
    if not bn.track_running_stats or bn.running_mean is None:
        raise RuntimeError("segb200: BatchNorm without running statistics is not supported by the training kernels")
    if isinstance(bn, torch.nn.SyncBatchNorm) and torch.distributed.is_available() and torch.distributed.is_initialized() \
            and torch.distributed.get_world_size() > 1:
        raise RuntimeError("segb200: SyncBatchNorm across ranks is only available in the whole-model training engine "
                           "(segmentron_b200.train.DeepLabV3PlusTrainerB200)")
    return bn.weight, bn.bias, bn.running_mean, bn.running_var, float(bn.momentum), float(bn.eps)


class ConvBNActFunction(torch.autograd.Function):
    """z = act(BN_train(conv(x, W)))  on NHWC 16-bit x;  W [Co,Ci,k,k] fp32 (groups = 1, no conv bias)."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, running_mean, running_var, momentum, eps, stride, dilation, pad, act):
        n, h, w, ci, _ = ops._nhwc(x, "x")
        co, ci_w, k, _ = weight.shape
        if ci != ci_w or ci % 8:
            raise RuntimeError(f"segb200: conv expects {ci_w} input channels (multiple of 8), got {ci}")
        if co % 8:
            raise RuntimeError("segb200: training-mode conv units need Cout % 8 == 0")
        dt = x.dtype
        ho = (h + 2 * pad - dilation * (k - 1) - 1) // stride + 1
        wo = (w + 2 * pad - dilation * (k - 1) - 1) // stride + 1
        y = torch.empty(n, ho, wo, co, dtype=dt, device=x.device)
        ops.conv_gemm(x, fold.pack_conv_weight(weight.detach(), dt), y, cin=ci, cout=co, kh=k, kw=k, stride=stride,
                      dilation=dilation, pad_t=pad, pad_l=pad)
        z = torch.empty_like(y)
        st = T.bn_forward(y, z, gamma.detach().float().contiguous(), beta.detach().float().contiguous(), running_mean, running_var,
                          momentum, eps, act=act)
        ctx.save_for_backward(x, y, weight)
        ctx.geo = (stride, dilation, pad, act, co, k)
        ctx.st = st
        return z

    @staticmethod
    def backward(ctx, dz):
        x, y, weight = ctx.saved_tensors
        stride, dilation, pad, act, co, k = ctx.geo
        st = ctx.st
        n, h, w, ci, _ = ops._nhwc(x, "x")
        _, ho, wo, _ = y.shape
        dt = x.dtype
        dz = dz.contiguous()
        dy = torch.empty_like(y)
        dgamma = torch.zeros(co, dtype=torch.float32, device=x.device)
        dbeta = torch.zeros(co, dtype=torch.float32, device=x.device)
        T.bn_backward(dz, None, y, st, dy, dgamma, dbeta, act=act)
        dwk = torch.zeros(co, k * k, ci, dtype=torch.float32, device=x.device)
        T.conv_wgrad(x, dy, dwk, cin=ci, cout=co, kh=k, kw=k, stride=stride, dilation=dilation, pad_t=pad, pad_l=pad)
        dweight = dwk.view(co, k, k, ci).permute(0, 3, 1, 2)
        dx = None
        if ctx.needs_input_grad[0]:
            if 2 * pad != dilation * (k - 1):
                raise RuntimeError("segb200: the data gradient needs 'same' padding (2*pad == dilation*(k-1))")
            wd = T.pack_dgrad_weight(weight.detach().to(dt), dt)
            dx = torch.empty(n, h, w, ci, dtype=dt, device=x.device)
            if stride == 1:
                ops.conv_gemm(dy, wd, dx, cin=co, cout=ci, kh=k, kw=k, dilation=dilation, pad_t=pad, pad_l=pad)
            elif k == 1:
                t = torch.empty(n, ho, wo, ci, dtype=dt, device=x.device)
                ops.conv_gemm(dy, wd, t, cin=co, cout=ci)
                T.stride2_place(t, dx, 0)
            else:
                if dilation != 1:
                    raise RuntimeError("segb200: stride-2 dilated conv has no data-gradient kernel")
                zb = torch.empty(n, h, w, co, dtype=dt, device=x.device)
                T.stride2_place(dy, zb, 0)
                ops.conv_gemm(zb, wd, dx, cin=co, cout=ci, kh=k, kw=k, pad_t=pad, pad_l=pad)
        return dx, dweight, dgamma, dbeta, None, None, None, None, None, None, None, None


class DwBNActFunction(torch.autograd.Function):
    """z = act(BN_train(dw3x3(pre_relu ? relu(x) : x)))  (groups = C, padding = dilation)."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, running_mean, running_var, momentum, eps, stride, dilation, pre_relu, act):
        n, h, w, c, _ = ops._nhwc(x, "x")
        if c % 8 or tuple(weight.shape) != (c, 1, 3, 3):
            raise RuntimeError("segb200: depthwise unit expects a [C,1,3,3] weight and C % 8 == 0")
        dt = x.dtype
        ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
        wf = weight.detach().float().reshape(c, 9).t().contiguous()
        y = torch.empty(n, ho, wo, c, dtype=dt, device=x.device)
        ops.dwconv3x3(x, wf, y, stride=stride, dilation=dilation, pre_relu=pre_relu)
        z = torch.empty_like(y)
        st = T.bn_forward(y, z, gamma.detach().float().contiguous(), beta.detach().float().contiguous(), running_mean, running_var,
                          momentum, eps, act=act)
        ctx.save_for_backward(x, y, weight)
        ctx.geo = (stride, dilation, pre_relu, act)
        ctx.st = st
        return z

    @staticmethod
    def backward(ctx, dz):
        x, y, weight = ctx.saved_tensors
        stride, dilation, pre_relu, act = ctx.geo
        st = ctx.st
        n, h, w, c, _ = ops._nhwc(x, "x")
        dt = x.dtype
        dz = dz.contiguous()
        dy = torch.empty_like(y)
        dgamma = torch.zeros(c, dtype=torch.float32, device=x.device)
        dbeta = torch.zeros(c, dtype=torch.float32, device=x.device)
        T.bn_backward(dz, None, y, st, dy, dgamma, dbeta, act=act)
        g_full = dy
        if stride == 2:
            if dilation != 1:
                raise RuntimeError("segb200: stride-2 dilated depthwise conv has no gradient kernel")
            g_full = torch.empty(n, h, w, c, dtype=dt, device=x.device)
            T.stride2_place(dy, g_full, 0)
        dwk = torch.zeros(c, 9, dtype=torch.float32, device=x.device)
        T.dw_wgrad(x, g_full, dwk, dilation=dilation, pre_relu=pre_relu)
        dx = None
        if ctx.needs_input_grad[0]:
            wflip = weight.detach().float().reshape(c, 9).flip(1).t().contiguous()
            dx = torch.empty(n, h, w, c, dtype=dt, device=x.device)
            if pre_relu:
                t = torch.empty(n, h, w, c, dtype=dt, device=x.device)
                ops.dwconv3x3(g_full, wflip, t, stride=1, dilation=dilation)
                T.relu_mask(t, x, dx)
            else:
                ops.dwconv3x3(g_full, wflip, dx, stride=1, dilation=dilation)
        return dx, dwk.view(c, 1, 3, 3), dgamma, dbeta, None, None, None, None, None, None, None, None


class GlobalAvgPoolFunction(torch.autograd.Function):
    """nn.AdaptiveAvgPool2d(1) on NHWC: [n,h,w,c] -> [n,1,1,c];  backward: dx = dy / (h w) at every pixel."""

    @staticmethod
    def forward(ctx, x):
        n, h, w, c, _ = ops._nhwc(x, "x")
        y = torch.empty(n, 1, 1, c, dtype=x.dtype, device=x.device)
        ops.global_avgpool(x, y)
        ctx.geo = (n, h, w, c)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, h, w, c = ctx.geo
        dx = torch.empty(n, h, w, c, dtype=dy.dtype, device=dy.device)
        T.nc_broadcast(dy.contiguous(), dx, scale=1.0 / (h * w))
        return dx


class BroadcastFunction(torch.autograd.Function):
    """F.interpolate(v[n,c,1,1], (h, w), 'bilinear', align_corners=True) == broadcast;  backward: dv = sum over pixels."""

    @staticmethod
    def forward(ctx, v, h, w):
        n, _, _, c, _ = ops._nhwc(v, "v")
        y = torch.empty(n, h, w, c, dtype=v.dtype, device=v.device)
        T.nc_broadcast(v, y)
        ctx.geo = (n, h, w, c)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, h, w, c = ctx.geo
        mean = torch.empty(n, 1, 1, c, dtype=dy.dtype, device=dy.device)
        ops.global_avgpool(dy.contiguous(), mean)
        dv = torch.empty_like(mean)
        T.nc_broadcast(mean, dv, scale=float(h * w))
        return dv, None, None


class AdaptiveAvgPoolFunction(torch.autograd.Function):
    """nn.AdaptiveAvgPool2d(s) on NHWC: [n,h,w,c] -> [n,s,s,c]  (PyramidPooling, modules/module.py:82-97)"""

    @staticmethod
    def forward(ctx, x, s):
        n, h, w, c, _ = ops._nhwc(x, "x")
        y = torch.empty(n, s, s, c, dtype=x.dtype, device=x.device)
        ops.adaptive_avgpool(x, y, s)
        ctx.geo = (n, h, w, c, s)
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import lib as L
        n, h, w, c, s = ctx.geo
        dy = dy.contiguous()
        dx = torch.empty(n, h, w, c, dtype=dy.dtype, device=dy.device)
        L.check(L.load().segb200_adaptive_avgpool_bwd(ops._ptr(dy), ops._ptr(dx), n, h, w, c, c, c, s, 0, ops.dt_code(dy.dtype),
                                                      ops._stream()), "adaptive_avgpool_bwd")
        return dx, None


class BilinearFunction(torch.autograd.Function):
    """F.interpolate(x, (ho, wo), mode='bilinear', align_corners=align) on NHWC; backward = the gather kernel of the training plan"""

    @staticmethod
    def forward(ctx, x, ho, wo, align):
        n, hi, wi, c, _ = ops._nhwc(x, "x")
        y = torch.empty(n, ho, wo, c, dtype=x.dtype, device=x.device)
        ops.bilinear_nhwc(x, y, align_corners=align)
        ctx.geo = (n, hi, wi, c, align)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, hi, wi, c, align = ctx.geo
        dx = torch.empty(n, hi, wi, c, dtype=dy.dtype, device=dy.device)
        T.bilinear_nhwc_bwd(dy.contiguous(), dx, align_corners=align)
        return dx, None, None, None

"""Training engine for DeepLabv3+ (ResNet backbones): one iteration of the reference's training loop
(tools/train.py:135-147 -- forward, CrossEntropy(ignore_index=-1), backward, SGD step; DDP gradient all-reduce) as a
STATIC launch list over the C-ABI kernels, SURVEY.md 8a rows a15/a16.

Design (B200-first, not an autograd port):
  * the whole step is planned once per input shape: every buffer (activations, gradients, BatchNorm statistics) is
    allocated up front, every kernel call is pre-marshalled; a step = replaying ~1.5k launches (optionally one CUDA graph);
  * parameters live in ONE flat fp32 master buffer (+ flat gradient and momentum buffers); conv weights are stored
    [Cout][kh*kw][Cin] (the channels_last physical layout of an OIHW tensor) so that the tcgen05 weight-gradient kernel
    accumulates straight into the gradient buffer, SGD is two launches, and the DDP all-reduce works on contiguous
    buckets of the same buffer, launched asynchronously while backward is still running;
  * per step the 16-bit GEMM operands (forward packing and the transposed/tap-flipped data-gradient packing) are
    regenerated from the masters by one index-table gather;
  * conv -> raw 16-bit output; BatchNorm batch statistics by a fixed-order two-level reduction; normalise + residual +
    ReLU (+ Dropout2d mask) in one pass.  Backward per unit: BN/ReLU backward (2 passes), weight gradient on the tensor
    cores straight from the NHWC activations (MN-major UMMA operands), data gradient = the forward implicit-GEMM kernel
    with transposed weights, accumulating into the consumer-shared gradient through its residual operand;
  * torch.cat never exists: branch outputs are channel slices of one buffer, and so are their gradients;
  * the [N,19,H,W] logits are never materialised: up-sampling + softmax cross-entropy + its gradient are one kernel.

``state_dict`` in / out uses the reference's parameter names (a reference checkpoint trains unchanged).
There is no CPU or PyTorch fallback: everything below calls the C ABI (lib.py) and raises RuntimeError otherwise.
"""
import ctypes as C
import os

import torch

from . import fold, lib as L, ops
from .ops import _ptr


class Step:
    __slots__ = ("kind", "call", "info")

    def __init__(self, kind, call, info):
        self.kind, self.call, self.info = kind, call, info


# ------------------------------------------------------------------------------------------------------------
# parameters: flat fp32 master / gradient / momentum buffers + packed 16-bit operands
# ------------------------------------------------------------------------------------------------------------
def _kind(name, t):
    if t.dim() == 4:
        # depthwise 3x3: [C, 1, 3, 3] (SeparableConv2d.depthwise, the groups=C _ConvBNReLU of InvertedResidual); no dense conv of
        # these models has a single input channel
        return "dw" if (t.shape[1] == 1 and t.shape[0] > 1 and tuple(t.shape[2:]) == (3, 3)) else "conv"
    return "vec"


class ParamStore:
    SKIP = ("running_mean", "running_var", "num_batches_tracked")

    def __init__(self, state_dict, device, dtype, stem=None):
        self.device, self.dtype = device, dtype
        names = [k for k in state_dict if not k.endswith(self.SKIP) and ".fc." not in k]
        names.sort(key=lambda k: 0 if k.startswith("encoder.") else 1)       # stable: encoder group first
        self.meta, off = {}, 0
        for k in names:
            t = state_dict[k]
            self.meta[k] = dict(off=off, shape=tuple(t.shape), kind=_kind(k, t), numel=t.numel())
            off += fold.round_up(t.numel(), 64)
            if k.startswith("encoder."):
                self.n_encoder = off
        self.total = off
        self.master = torch.zeros(off, dtype=torch.float32, device=device)
        self.grad = torch.zeros(off, dtype=torch.float32, device=device)
        self.mom = torch.zeros(off, dtype=torch.float32, device=device)
        for k in names:
            self.view(self.master, k).copy_(self.to_internal(k, state_dict[k].detach().float()))
        # BatchNorm running statistics (not trained): one flat buffer
        self.stat_meta, soff = {}, 0
        for k in state_dict:
            if k.endswith(("running_mean", "running_var")):
                self.stat_meta[k] = (soff, state_dict[k].numel())
                soff += fold.round_up(state_dict[k].numel(), 64)
        self.stats = torch.zeros(max(soff, 64), dtype=torch.float32, device=device)
        for k, (o, n) in self.stat_meta.items():
            self.stats[o:o + n].copy_(state_dict[k].detach().float())
        self.extra = {k: v for k, v in state_dict.items() if k.endswith("num_batches_tracked") or ".fc." in k}
        self.steps = 0                                 # training-mode forwards since construction (-> num_batches_tracked)
        # ---- packed operands: index tables (built on the CPU, int64 -> int32) ----
        self.stem = stem
        idx16, idx32, self.pk = [], [], {}
        n16 = n32 = 0
        for k in names:
            m = self.meta[k]
            if m["kind"] == "conv":
                co, ci, kh, kw = m["shape"]
                T = kh * kw
                base = torch.arange(co * T * ci, dtype=torch.int64).reshape(co, T, ci) + m["off"]
                if k == stem:
                    fwd, T2, pad2, scat = _stem_indices(base.reshape(co, kh, kw, ci), pad=(kh - 1) // 2)
                    self.pk[k] = dict(fwd=(n16, tuple(fwd.shape)), T=T2, pad=pad2, scatter=scat.to(torch.int32).to(device))
                    idx16.append(fwd.flatten()); n16 += fold.round_up(fwd.numel(), 64)
                    idx16.append(torch.full((fold.round_up(fwd.numel(), 64) - fwd.numel(),), -1, dtype=torch.int64))
                    continue
                cip = fold.round_up(ci, fold.conv_kblock(ci))
                cop = fold.round_up(co, 8)
                fwd = torch.full((cop, T, cip), -1, dtype=torch.int64)
                fwd[:co, :, :ci] = base
                ci8 = fold.round_up(ci, 8)
                cok = fold.round_up(cop, fold.conv_kblock(cop))
                dg = torch.full((ci8, T, cok), -1, dtype=torch.int64)
                dg[:ci, :, :co] = base.permute(2, 1, 0).flip(1)
                ent = {}
                for nm, tab in (("fwd", fwd), ("dgrad", dg)):
                    ent[nm] = (n16, tuple(tab.shape))
                    pad = fold.round_up(tab.numel(), 64) - tab.numel()
                    idx16 += [tab.flatten(), torch.full((pad,), -1, dtype=torch.int64)]
                    n16 += tab.numel() + pad
                self.pk[k] = ent
            elif m["kind"] == "dw":
                c = m["shape"][0]
                base = torch.arange(c * 9, dtype=torch.int64).reshape(c, 9) + m["off"]
                ent = {}
                for nm, tab in (("fwd", base.t().contiguous()), ("flip", base.flip(1).t().contiguous())):
                    ent[nm] = (n32, (9, c))
                    pad = fold.round_up(tab.numel(), 64) - tab.numel()
                    idx32 += [tab.flatten(), torch.full((pad,), -1, dtype=torch.int64)]
                    n32 += tab.numel() + pad
                self.pk[k] = ent
        assert self.total < 2 ** 31 and n16 < 2 ** 31
        self.idx16 = torch.cat(idx16).to(torch.int32).to(device) if idx16 else None
        self.idx32 = torch.cat(idx32).to(torch.int32).to(device) if idx32 else None
        self.w16 = torch.zeros(max(n16, 64), dtype=dtype, device=device)
        self.w32 = torch.zeros(max(n32, 64), dtype=torch.float32, device=device)

    # ---- layout ----
    def to_internal(self, k, t):
        return t.permute(0, 2, 3, 1).reshape(-1) if self.meta[k]["kind"] == "conv" else t.reshape(-1)

    def from_internal(self, k, flat):
        m = self.meta[k]
        if m["kind"] == "conv":
            co, ci, kh, kw = m["shape"]
            return flat.reshape(co, kh, kw, ci).permute(0, 3, 1, 2).contiguous()
        return flat.reshape(m["shape"]).clone()

    def view(self, buf, k):
        m = self.meta[k]
        return buf[m["off"]:m["off"] + m["numel"]]

    def packed(self, k, which):
        off, shape = self.pk[k][which]
        buf = self.w32 if self.meta[k]["kind"] == "dw" else self.w16
        n = 1
        for s in shape:
            n *= s
        return buf[off:off + n].view(shape)

    def stat(self, k):
        o, n = self.stat_meta[k]
        return self.stats[o:o + n]

    def state_dict(self):
        out = {k: self.from_internal(k, self.view(self.master, k)) for k in self.meta}
        out.update({k: self.stat(k).clone() for k in self.stat_meta})
        out.update(self.extra)
        for k, v in self.extra.items():               # nn.BatchNorm2d counts its training-mode forwards (batch_norm.py:126)
            if k.endswith("num_batches_tracked"):
                out[k] = v.detach().clone() + self.steps
        return out

    def grads(self):
        """{reference parameter name: gradient in the reference (OIHW) layout} -- a copy, for inspection / tests."""
        return {k: self.from_internal(k, self.view(self.grad, k)) for k in self.meta}


def _stem_indices(base, pad):
    """Index form of fold.pack_stem_s2d for a stride-2 kxk stem on Cin<=4 channels.  base: int64 [co][k][k][ci] of master
    offsets.  -> (forward table [co][T*T][16], T, pad2, scatter table [co][T*T][64] for the s2d-space weight gradient)."""
    co, k, _, ci = base.shape
    offs = [ky - pad for ky in range(k)]
    a = [o // 2 for o in offs]
    par = [o % 2 for o in offs]
    a_min, a_max = min(a), max(a)
    T = a_max - a_min + 1
    ld = fold.round_up(4 * ci, 16)
    assert ld == 16
    fwd = torch.full((co, T, T, ld), -1, dtype=torch.int64)
    scat = torch.full((co, T, T, 64), -1, dtype=torch.int64)
    for ky in range(k):
        for kx in range(k):
            ch0 = (par[ky] * 2 + par[kx]) * ci
            fwd[:, a[ky] - a_min, a[kx] - a_min, ch0:ch0 + ci] = base[:, ky, kx, :]
            scat[:, a[ky] - a_min, a[kx] - a_min, ch0:ch0 + ci] = base[:, ky, kx, :]
    return fwd.reshape(co, T * T, ld), T, -a_min, scat.reshape(-1)


# ------------------------------------------------------------------------------------------------------------
# activations with gradient routing
# ------------------------------------------------------------------------------------------------------------
class Act:
    """An NHWC activation (or a channel slice of one) plus its gradient buffer and the overwrite/accumulate state."""

    def __init__(self, plan, t, parent=None, lo=0, needs_grad=True):
        self.plan, self.t, self.parent, self.lo = plan, t, parent, lo
        self.g = None
        self.written = False
        self.needs_grad = needs_grad

    def slice(self, lo, hi):
        assert self.parent is None
        return Act(self.plan, self.t[..., lo:hi], parent=self, lo=lo)

    def _root_grad(self):
        r = self.parent or self
        if r.g is None:
            n, h, w, c = r.t.shape
            r.g = self.plan.pool_get(n, h, w, c, r.t.stride(2))
        return r.g

    def grad(self):
        """the gradient as a tensor view (must have been written by every consumer before it is read)"""
        r = self.parent or self
        if not r.written:
            raise RuntimeError("segb200 train plan: gradient read before any consumer wrote it")
        g = self._root_grad()
        return g[..., self.lo:self.lo + self.t.shape[3]] if self.parent is not None else g

    def take(self):
        """-> (gradient tensor to write into, accumulate?) and mark it written"""
        if self.parent is not None:
            # a slice can only ACCUMULATE: the consumer of the whole buffer (e.g. the conv over a concat) must have written the
            # root gradient first -- true whenever the slice's other consumers were recorded before the concat consumer
            if not self.parent.written:
                raise RuntimeError("segb200 train plan: a channel slice's gradient was written before its concat buffer's")
            return self.parent._root_grad()[..., self.lo:self.lo + self.t.shape[3]], True
        g = self._root_grad()
        acc = self.written
        self.written = True
        return g, acc


# ------------------------------------------------------------------------------------------------------------
# the plan
# ------------------------------------------------------------------------------------------------------------
class TrainPlan:
    def __init__(self, store, shape, nclass, dtype, device, bn_momentum=0.1, ignore_index=-1, dist=None, sync_bn=False, xchg=None):
        self.S, self.dtype, self.device = store, dtype, device
        # SyncBatchNorm (the reference's default for distributed training, tools/train.py:73-79): batch statistics and the two
        # backward sums are all-reduced over the ranks ([2][C] fp32 per layer, count-weighted == equal per-rank counts here)
        self.dist = dist if (dist is not None and sync_bn) else None
        self.world = dist.get_world_size() if self.dist is not None else 1
        # xchg (parallel.SyncExchange): the two exchanges become part of the finalize kernels (P2P stores over NVLink, csrc/syncbn.cu)
        # instead of reduce_partials + NCCL all_reduce + finalize; None -> the all_reduce form (gloo CPU tests, no symmetric memory)
        self.xchg = xchg if self.dist is not None else None
        self.lib = L.load()
        self.dt = ops.dt_code(dtype)
        self.n, _, self.H, self.W = shape
        self.nclass, self.ignore_index, self.bn_momentum = nclass, ignore_index, bn_momentum
        self.fwd, self.bwd, self.tape = [], [], []
        self.cur = self.fwd
        self.keep, self.pool, self.pool_bytes, self.act_bytes = [], {}, 0, 0
        self.x_in = torch.zeros(shape, dtype=torch.float32, device=device)
        self.target = torch.zeros(self.n, self.H, self.W, dtype=torch.int64, device=device)
        self.out3 = torch.zeros(3, dtype=self.STAT_DTYPE, device=device)
        self.masks = {}
        self.done_at = {}                 # parameter name -> number of backward steps after which its gradient is final

    # ---- memory ----
    def new(self, n, h, w, c, ld=None):
        ld = ld or fold.round_up(c, 8)
        buf = torch.zeros(n, h, w, ld, dtype=self.dtype, device=self.device)
        self.keep.append(buf)
        self.act_bytes += buf.numel() * 2
        return buf[..., :c] if ld != c else buf

    def pool_get(self, n, h, w, c, ld=None):
        ld = ld or fold.round_up(c, 8)
        key = (n, h, w, ld)
        lst = self.pool.setdefault(key, [])
        if lst:
            buf = lst.pop()
        else:
            buf = torch.zeros(n, h, w, ld, dtype=self.dtype, device=self.device)
            self.keep.append(buf)
            self.pool_bytes += buf.numel() * 2
        return buf[..., :c] if ld != c else buf

    def pool_put(self, t):
        base = t._base if t._base is not None else t
        self.pool.setdefault((base.shape[0], base.shape[1], base.shape[2], base.shape[3]), []).append(base)

    STAT_DTYPE = torch.float32            # (the CPU plan interpreter of the tests raises it to fp64 for an exact comparison)

    def f32(self, *shape):
        t = torch.zeros(*shape, dtype=self.STAT_DTYPE, device=self.device)
        self.keep.append(t)
        return t

    # ---- step recording ----
    def add(self, kind, fn, args, **info):
        self.cur.append(Step(kind, (lambda s, fn=fn, args=args, kind=kind: L.check(fn(*args, s), kind)), info))

    def allreduce(self, t):
        d = self.dist
        self.cur.append(Step("allreduce", (lambda s, t=t, d=d: d.all_reduce(t)), dict(t=t)))

    def conv(self, x, w, y, *, cin, cout, k=1, stride=1, dilation=1, pad=0, shift=None, residual=None):
        a = ops.make_conv_args(x, w, y, cin=cin, cout=cout, kh=k, kw=k, stride=stride, dilation=dilation, pad_t=pad,
                               pad_l=pad, shift=shift, residual=residual)
        fn = self.lib.segb200_conv_gemm
        self.cur.append(Step("conv", (lambda s, a=a, fn=fn: L.check(fn(C.byref(a), s), "conv_gemm")),
                             dict(x=x, w=w, y=y, cin=cin, cout=cout, k=k, stride=stride, dilation=dilation, pad=pad, shift=shift,
                                  residual=residual, flops=2.0 * a.n * a.ho * a.wo * cout * cin * k * k)))

    def wgrad(self, x, dy, dw, *, cin, cout, k=1, stride=1, dilation=1, pad=0):
        a = L.WgradArgs()
        n, h, w_, _, x_ld = ops._nhwc(x, "x")
        _, ho, wo, _, dy_ld = ops._nhwc(dy, "dy")
        a.x, a.dy, a.dw = _ptr(x), _ptr(dy), _ptr(dw)
        a.n, a.h, a.w, a.cin, a.x_ld = n, h, w_, cin, x_ld
        a.ho, a.wo, a.cout, a.dy_ld = ho, wo, cout, dy_ld
        a.kh, a.kw, a.stride, a.dilation, a.pad_t, a.pad_l = k, k, stride, dilation, pad, pad
        a.dtype, a.max_ctas, a.splits = self.dt, 0, 0
        fn = self.lib.segb200_conv_wgrad
        self.cur.append(Step("wgrad", (lambda s, a=a, fn=fn: L.check(fn(C.byref(a), s), "conv_wgrad")),
                             dict(x=x, dy=dy, dw=dw, cin=cin, cout=cout, k=k, stride=stride, dilation=dilation, pad=pad,
                                  flops=2.0 * n * ho * wo * cout * cin * k * k)))

    def dwconv(self, x, w, y, dilation, stride=1, pre_relu=False):
        a = ops.make_dw_args(x, w, y, stride=stride, dilation=dilation, pre_relu=pre_relu)
        fn = self.lib.segb200_dwconv3x3
        self.cur.append(Step("dw", (lambda s, a=a, fn=fn: L.check(fn(C.byref(a), s), "dwconv3x3")),
                             dict(x=x, w=w, y=y, dilation=dilation, stride=stride, pre_relu=pre_relu)))

    @staticmethod
    def _rows(t):
        n, h, w, c, ld = ops._nhwc(t, "t")
        return n * h * w, h * w, c, ld

    # ---- BatchNorm(+residual, act, channel mask) unit: forward steps + backward builder ----
    def bn_act_fwd(self, y, z, bn, act, eps, residual=None, nc_scale=None):
        rows, hw, c, y_ld = self._rows(y)
        S = self.S
        st = dict(mean=self.f32(c), invstd=self.f32(c), scale=self.f32(c), shift=self.f32(c), sums=self.f32(2, c))
        slabs = self.lib.segb200_reduce_slabs(rows, c, 0)
        partial = self.f32(slabs * 2 * c)
        st.update(slabs=slabs, partial=partial)
        gamma, beta = S.view(S.master, bn + ".weight"), S.view(S.master, bn + ".bias")
        rm, rv = S.stat(bn + ".running_mean"), S.stat(bn + ".running_var")
        self.add("bn_stats", self.lib.segb200_bn_stats, (_ptr(y), rows, c, y_ld, self.dt, _ptr(partial), 0), x=y, partial=partial, c=c)
        fin_partial, fin_slabs, count = partial, slabs, float(rows)
        if self.xchg is not None:
            X = self.xchg
            data_off, flag_off = X.new_slot()
            count = float(rows * self.world)
            st["count"] = count
            self.add("bn_finalize_sync", self.lib.segb200_bn_finalize_sync,
                     (_ptr(partial), slabs, c, count, _ptr(gamma), _ptr(beta), _ptr(rm), _ptr(rv), self.bn_momentum, eps,
                      _ptr(st["mean"]), _ptr(st["invstd"]), _ptr(st["scale"]), _ptr(st["shift"]), C.c_void_p(X.peers_dev), X.world, X.rank,
                      X.cmax, data_off, flag_off, _ptr(X.epoch)),
                     partial=partial, slabs=slabs, c=c, count=count, gamma=gamma, beta=beta, rm=rm, rv=rv, momentum=self.bn_momentum,
                     eps=eps, st=st, dist=self.dist)
        elif self.dist is not None:
            gsum = self.f32(2 * c)
            self.add("reduce_partials", self.lib.segb200_reduce_partials, (_ptr(partial), slabs, 2, c, _ptr(gsum), c, 1, 0, 1.0),
                     partial=partial, slabs=slabs, K=2, c=c, out=gsum, sk=c, sc=1, accumulate=0, scale=1.0)
            self.allreduce(gsum)
            fin_partial, fin_slabs, count = gsum, 1, float(rows * self.world)
        st["count"] = count
        if self.xchg is None:
            self.add("bn_finalize", self.lib.segb200_bn_finalize,
                     (_ptr(fin_partial), fin_slabs, c, count, _ptr(gamma), _ptr(beta), _ptr(rm), _ptr(rv), self.bn_momentum, eps,
                      _ptr(st["mean"]), _ptr(st["invstd"]), _ptr(st["scale"]), _ptr(st["shift"])),
                     partial=fin_partial, slabs=fin_slabs, c=c, count=count, gamma=gamma, beta=beta, rm=rm, rv=rv,
                     momentum=self.bn_momentum, eps=eps, st=st)
        res_ld = self._rows(residual)[3] if residual is not None else 0
        self.add("bn_apply", self.lib.segb200_bn_apply,
                 (_ptr(y), _ptr(st["scale"]), _ptr(st["shift"]), _ptr(residual), _ptr(nc_scale), _ptr(z), rows, hw, c, y_ld, res_ld,
                  self._rows(z)[3], L.ACT[act], self.dt),
                 y=y, scale=st["scale"], shift=st["shift"], residual=residual, nc_scale=nc_scale, z=z, act=act)
        return st

    def bn_act_bwd(self, dz, z, y, st, bn, act, dy, residual=None, nc_scale=None):
        """emit: BN/act backward -> dy; dgamma/dbeta accumulated; residual.grad (+)= g"""
        rows, hw, c, dz_ld = self._rows(dz)
        S = self.S
        # the ReLU mask needs the stored output only when a residual entered the activation; otherwise both passes recompute
        # it from y (one activation read less in each)
        zz = z if (act is not None and residual is not None) else None
        z_ld = self._rows(z)[3] if zz is not None else 0
        y_ld = self._rows(y)[3]
        self.add("bn_bwd_reduce", self.lib.segb200_bn_bwd_reduce,
                 (_ptr(dz), _ptr(zz), _ptr(y), _ptr(st["scale"]), _ptr(st["shift"]), _ptr(nc_scale), _ptr(st["partial"]), rows, hw, c,
                  dz_ld, z_ld, y_ld, L.ACT[act], self.dt, 0),
                 dz=dz, z=zz, y=y, st=st, nc_scale=nc_scale, act=act, c=c)
        dgamma, dbeta = S.view(S.grad, bn + ".weight"), S.view(S.grad, bn + ".bias")
        if self.xchg is not None:
            X = self.xchg
            data_off, flag_off = X.new_slot()
            self.add("bn_bwd_finalize_sync", self.lib.segb200_bn_bwd_finalize_sync,
                     (_ptr(st["partial"]), st["slabs"], c, _ptr(st["mean"]), _ptr(st["invstd"]), _ptr(st["sums"]), _ptr(dgamma), _ptr(dbeta),
                      C.c_void_p(X.peers_dev), X.world, X.rank, X.cmax, data_off, flag_off, _ptr(X.epoch)),
                     st=st, c=c, dgamma=dgamma, dbeta=dbeta, dist=self.dist)
        else:
            self.add("bn_bwd_finalize", self.lib.segb200_bn_bwd_finalize,
                     (_ptr(st["partial"]), st["slabs"], c, _ptr(st["mean"]), _ptr(st["invstd"]), _ptr(st["sums"]), _ptr(dgamma), _ptr(dbeta)),
                     st=st, c=c, dgamma=dgamma, dbeta=dbeta)
            if self.dist is not None:
                self.allreduce(st["sums"])           # dgamma / dbeta above are the LOCAL sums (DDP averages them with the rest)
        count = st.get("count", float(rows))
        dres, dres_acc, dres_ld = None, False, 0
        if residual is not None:
            dres, dres_acc = residual.take()
            dres_ld = self._rows(dres)[3]
        self.add("bn_bwd_apply", self.lib.segb200_bn_bwd_apply,
                 (_ptr(dz), _ptr(zz), _ptr(y), _ptr(st["mean"]), _ptr(st["invstd"]), _ptr(st["scale"]), _ptr(st["shift"]),
                  _ptr(st["sums"]), count, _ptr(nc_scale), _ptr(dy), _ptr(dres), int(dres_acc), rows, hw, c, dz_ld, z_ld, y_ld,
                  self._rows(dy)[3], dres_ld, L.ACT[act], self.dt),
                 dz=dz, z=zz, y=y, st=st, count=count, nc_scale=nc_scale, dy=dy, dres=dres, dres_acc=dres_acc, act=act)

    # ---- conv (+BN +act) unit ----
    def conv_unit(self, x, wname, bn=None, act=None, *, k=1, stride=1, dilation=1, pad=0, residual=None, out=None, bias=None,
                  nc_scale=None, eps=1e-5, stem=False):
        S = self.S
        n, h, w_, cx = x.t.shape
        co, ci = S.meta[wname]["shape"][:2]
        cop = fold.round_up(co, 8)
        if stem:
            T, pad2 = S.pk[wname]["T"], S.pk[wname]["pad"]
            ho, wo = (self.H + 2 * pad - k) // 2 + 1, (self.W + 2 * pad - k) // 2 + 1
            geo = dict(cin=16, k=T, stride=1, dilation=1, pad=pad2)
        else:
            ho = (h + 2 * pad - dilation * (k - 1) - 1) // stride + 1
            wo = (w_ + 2 * pad - dilation * (k - 1) - 1) // stride + 1
            assert cx == ci, (wname, cx, ci)
            geo = dict(cin=ci, k=k, stride=stride, dilation=dilation, pad=pad)
        wf = S.packed(wname, "fwd")
        if bn is not None:
            y = self.new(n, ho, wo, cop)
            z = out if out is not None else Act(self, self.new(n, ho, wo, cop))
            # a conv bias in front of BatchNorm (hrnet_seg.py:42-47) is the GEMM's shift; its gradient is the column sum of dy
            self.conv(x.t, wf, y, cout=cop, shift=S.view(S.master, bias) if bias else None, **geo)
            st = self.bn_act_fwd(y, z.t, bn, act, eps, residual.t if residual is not None else None, nc_scale)
        else:
            assert act is None and residual is None and nc_scale is None
            z = out if out is not None else Act(self, self.new(n, ho, wo, cop))
            y, st = z.t, None
            self.conv(x.t, wf, y, cout=cop, shift=S.view(S.master, bias) if bias else None, **geo)

        def backward():
            dz = z.grad()
            if bn is not None:
                dy = self.pool_get(n, ho, wo, cop)
                self.bn_act_bwd(dz, z.t, y, st, bn, act, dy, residual, nc_scale)
                if z.parent is None:
                    self.pool_put(dz)
            else:
                dy = dz
            if bias:
                rows, hw, c, ld = self._rows(dy)
                slabs = self.lib.segb200_reduce_slabs(rows, c, 0)
                partial, sums = self.f32(slabs * 2 * c), self.f32(2, c)
                self.add("bn_bwd_reduce", self.lib.segb200_bn_bwd_reduce,
                         (_ptr(dy), None, _ptr(dy), None, None, None, _ptr(partial), rows, hw, c, ld, 0, ld, 0, self.dt, 0),
                         dz=dy, z=None, y=dy, st=dict(mean=None, invstd=None, partial=partial), nc_scale=None, act=None, c=c)
                self.add("bn_bwd_finalize", self.lib.segb200_bn_bwd_finalize, (_ptr(partial), slabs, c, None, None, _ptr(sums), None, None),
                         st=dict(partial=partial, slabs=slabs, sums=sums), c=c, dgamma=None, dbeta=None)
                gb = S.view(S.grad, bias)
                self.add("reduce_partials", self.lib.segb200_reduce_partials, (_ptr(sums), 1, 1, co, _ptr(gb), 0, 1, 1, 1.0),
                         partial=sums, slabs=1, K=1, c=co, out=gb, sk=0, sc=1, accumulate=1, scale=1.0)
            # weight gradient
            if stem:
                dws = self.f32(co * geo["k"] ** 2 * 64)
                self.cur.append(Step("zero", (lambda s, t=dws: t.zero_()), dict(t=dws)))
                self.wgrad(x.t.as_strided((n, x.t.shape[1], x.t.shape[2], 64), x.t.stride()) if x.t.shape[3] != 64 else x.t, dy, dws,
                           cin=64, cout=co, k=geo["k"], stride=1, dilation=1, pad=geo["pad"])
                sc = S.pk[wname]["scatter"]
                self.add("scatter_add", self.lib.segb200_scatter_add, (_ptr(dws), _ptr(sc), _ptr(S.grad), dws.numel()), src=dws, index=sc,
                         dst=S.grad)
            else:
                self.wgrad(x.t, dy, S.view(S.grad, wname), cin=ci, cout=co, **{kk: geo[kk] for kk in ("k", "stride", "dilation", "pad")})
            # data gradient
            if x.needs_grad:
                wd = S.packed(wname, "dgrad")
                gx, acc = x.take()
                if stride == 1:
                    assert 2 * pad == dilation * (k - 1), "data gradient via the forward kernel needs 'same' padding"
                    self.conv(dy, wd, gx, cin=cop, cout=ci, k=k, dilation=dilation, pad=pad, residual=gx if acc else None)
                elif k == 1:
                    t = self.pool_get(n, ho, wo, ci)
                    self.conv(dy, wd, t, cin=cop, cout=ci)
                    self.add("stride2_place", self.lib.segb200_stride2_place,
                             (_ptr(t), _ptr(gx), n, h, w_, ci, self._rows(t)[3], self._rows(gx)[3], 1 if acc else 0, self.dt), t=t, z=gx,
                             mode=1 if acc else 0)
                    self.pool_put(t)
                else:
                    assert dilation == 1 and 2 * pad == k - 1
                    zb = self.pool_get(n, h, w_, cop)
                    self.add("stride2_place", self.lib.segb200_stride2_place,
                             (_ptr(dy), _ptr(zb), n, h, w_, cop, self._rows(dy)[3], self._rows(zb)[3], 0, self.dt), t=dy, z=zb, mode=0)
                    self.conv(zb, wd, gx, cin=cop, cout=ci, k=k, dilation=1, pad=pad, residual=gx if acc else None)
                    self.pool_put(zb)
            if bn is not None:
                self.pool_put(dy)
            elif z.parent is None:
                self.pool_put(dz)
            self.mark_done(wname, bias, bn + ".weight" if bn else None, bn + ".bias" if bn else None)

        self.tape.append(backward)
        return z

    # ---- depthwise 3x3 + BN + act (the first half of SeparableConv2d, relu_first=False: modules/basic.py:52-59) ----
    def dw_unit(self, x, wname, bn, act, dilation, eps=1e-5, stride=1, pre_relu=False):
        """depthwise 3x3 (+ optional leading ReLU, modules/basic.py:45-46) + BatchNorm (+ ReLU): the first half of SeparableConv2d"""
        S = self.S
        n, h, w_, c = x.t.shape
        ho, wo = (h - 1) // stride + 1, (w_ - 1) // stride + 1
        y = self.new(n, ho, wo, c)
        z = Act(self, self.new(n, ho, wo, c))
        self.dwconv(x.t, S.packed(wname, "fwd"), y, dilation, stride, pre_relu)
        st = self.bn_act_fwd(y, z.t, bn, act, eps)

        def backward():
            dz = z.grad()
            dy = self.pool_get(n, ho, wo, c)
            self.bn_act_bwd(dz, z.t, y, st, bn, act, dy)
            self.pool_put(dz)
            g_full = dy
            if stride == 2:                              # zero insertion turns both stride-2 gradients into stride-1 ones
                assert dilation == 1
                g_full = self.pool_get(n, h, w_, c)
                self.add("stride2_place", self.lib.segb200_stride2_place,
                         (_ptr(dy), _ptr(g_full), n, h, w_, c, self._rows(dy)[3], self._rows(g_full)[3], 0, self.dt), t=dy, z=g_full, mode=0)
            rows = n * h * w_
            slabs = self.lib.segb200_reduce_slabs(rows, c, 0)
            partial = self.f32(slabs * 9 * c)
            self.add("dw_wgrad", self.lib.segb200_dw_wgrad,
                     (_ptr(x.t), _ptr(g_full), _ptr(partial), n, h, w_, c, self._rows(x.t)[3], self._rows(g_full)[3], dilation,
                      int(pre_relu), self.dt, 0),
                     x=x.t, dy=g_full, partial=partial, dilation=dilation, c=c, pre_relu=pre_relu)
            gw = S.view(S.grad, wname)
            self.add("reduce_partials", self.lib.segb200_reduce_partials, (_ptr(partial), slabs, 9, c, _ptr(gw), 1, 9, 1, 1.0),
                     partial=partial, slabs=slabs, K=9, c=c, out=gw, sk=1, sc=9, accumulate=1, scale=1.0)
            if x.needs_grad:
                gx, acc = x.take()
                wfl = S.packed(wname, "flip")
                if pre_relu:
                    # dx (+)= dw_dgrad * [x > 0]: the mask of the leading ReLU, applied by the BN-backward kernel's residual path
                    t = self.pool_get(n, h, w_, c)
                    self.dwconv(g_full, wfl, t, dilation)
                    rws, hw, cc, ld = self._rows(gx)
                    self.add("bn_bwd_apply", self.lib.segb200_bn_bwd_apply,
                             (_ptr(t), _ptr(x.t), None, None, None, None, None, None, float(rws), None, None, _ptr(gx), int(acc), rws, hw,
                              cc, self._rows(t)[3], self._rows(x.t)[3], 0, 0, ld, L.ACT["relu"], self.dt),
                             dz=t, z=x.t, y=None, st=dict(mean=None, invstd=None, scale=None, shift=None, sums=None), count=float(rws),
                             nc_scale=None, dy=None, dres=gx, dres_acc=acc, act="relu")
                    self.pool_put(t)
                elif acc:
                    t = self.pool_get(n, h, w_, c)
                    self.dwconv(g_full, wfl, t, dilation)
                    rws, hw, cc, ld = self._rows(gx)
                    self.add("bn_apply", self.lib.segb200_bn_apply,
                             (_ptr(t), None, None, _ptr(gx), None, _ptr(gx), rws, hw, cc, self._rows(t)[3], ld, ld, 0, self.dt),
                             y=t, scale=None, shift=None, residual=gx, nc_scale=None, z=gx, act=None)
                    self.pool_put(t)
                else:
                    self.dwconv(g_full, wfl, gx, dilation)
            if stride == 2:
                self.pool_put(g_full)
            self.pool_put(dy)
            self.mark_done(wname, bn + ".weight", bn + ".bias")

        self.tape.append(backward)
        return z

    def mark_done(self, *names):
        for k in names:
            if k:
                self.done_at[k] = len(self.bwd)

    # ---- criss-cross attention (CrissCrossAttention.forward, modules/cc_attention.py:62-72), weights shared between calls ----
    def cca_unit(self, x, prefix):
        S = self.S
        n, h, w_, c = x.t.shape
        cq = S.meta[prefix + ".query_conv.weight"]["shape"][0]
        q = self.conv_unit(x, prefix + ".query_conv.weight", bias=prefix + ".query_conv.bias")
        k = self.conv_unit(x, prefix + ".key_conv.weight", bias=prefix + ".key_conv.bias")
        v = self.conv_unit(x, prefix + ".value_conv.weight", bias=prefix + ".value_conv.bias")
        att_ld = fold.round_up(h + w_ - 1, 4)
        att = self.f32(n, h, w_, att_ld)
        gamma = S.view(S.master, prefix + ".gamma")
        y = Act(self, self.new(n, h, w_, c))
        self.add("cca_weight_softmax", self.lib.segb200_cca_weight_softmax,
                 (_ptr(q.t), _ptr(k.t), _ptr(att), n, h, w_, cq, q.t.stride(2), k.t.stride(2), att_ld, self.dt), q=q.t, k=k.t, att=att)
        self.add("cca_map", self.lib.segb200_cca_map,
                 (_ptr(att), _ptr(v.t), _ptr(x.t), _ptr(y.t), _ptr(gamma), n, h, w_, c, att_ld, v.t.stride(2), x.t.stride(2),
                  y.t.stride(2), self.dt), att=att, v=v.t, x=x.t, y=y.t, gamma=gamma)

        def backward():
            dy = y.grad()
            de = self.f32(n, h, w_, att_ld)
            nb = self.lib.segb200_cca_weight_bwd_blocks(n, h, w_)
            part = self.f32(nb)
            self.add("cca_weight_bwd", self.lib.segb200_cca_weight_bwd,
                     (_ptr(dy), _ptr(v.t), _ptr(att), _ptr(de), _ptr(part), _ptr(gamma), n, h, w_, c, dy.stride(2), v.t.stride(2), att_ld,
                      self.dt), dy=dy, v=v.t, att=att, de=de, part=part, gamma=gamma)
            gg = S.view(S.grad, prefix + ".gamma")
            self.add("reduce_partials", self.lib.segb200_reduce_partials, (_ptr(part), nb, 1, 1, _ptr(gg), 0, 1, 1, 1.0),
                     partial=part, slabs=nb, K=1, c=1, out=gg, sk=0, sc=1, accumulate=1, scale=1.0)
            gq, aq = q.take()
            self.add("cca_gather", self.lib.segb200_cca_gather,
                     (_ptr(de), _ptr(k.t), _ptr(gq), n, h, w_, cq, att_ld, k.t.stride(2), gq.stride(2), 1.0, int(aq), self.dt),
                     a=de, src=k.t, out=gq, scale=1.0, scale_dev=None, accumulate=aq)
            gk, ak = k.take()
            self.add("cca_scatter", self.lib.segb200_cca_scatter,
                     (_ptr(de), _ptr(q.t), _ptr(gk), n, h, w_, cq, att_ld, q.t.stride(2), gk.stride(2), 1.0, None, int(ak), self.dt),
                     a=de, src=q.t, out=gk, scale=1.0, scale_dev=None, accumulate=ak)
            gv, av = v.take()
            self.add("cca_scatter", self.lib.segb200_cca_scatter,
                     (_ptr(att), _ptr(dy), _ptr(gv), n, h, w_, c, att_ld, dy.stride(2), gv.stride(2), 1.0, _ptr(gamma), int(av), self.dt),
                     a=att, src=dy, out=gv, scale=1.0, scale_dev=gamma, accumulate=av)
            gx, ax = x.take()                                   # the residual: dx (+)= dy
            rws, hw, cc, ld = self._rows(gx)
            self.add("bn_apply", self.lib.segb200_bn_apply,
                     (_ptr(dy), None, None, _ptr(gx) if ax else None, None, _ptr(gx), rws, hw, cc, dy.stride(2), ld if ax else 0, ld, 0,
                      self.dt), y=dy, scale=None, shift=None, residual=gx if ax else None, nc_scale=None, z=gx, act=None)
            self.pool_put(dy)
            self.mark_done(prefix + ".gamma")

        self.tape.append(backward)
        return y

    # ---- glue ops ----
    def maxpool(self, x):
        n, h, w_, c = x.t.shape
        ho, wo = (h - 1) // 2 + 1, (w_ - 1) // 2 + 1
        y = Act(self, self.new(n, ho, wo, c))
        idx = torch.zeros(n, ho, wo, c, dtype=torch.uint8, device=self.device)
        self.keep.append(idx)
        self.add("maxpool", self.lib.segb200_maxpool3x3s2_idx, (_ptr(x.t), _ptr(y.t), _ptr(idx), n, h, w_, c, x.t.stride(2),
                                                                y.t.stride(2), self.dt), x=x.t, y=y.t, idx=idx)

        def backward():
            dy = y.grad()
            gx, acc = x.take()
            assert not acc
            self.add("maxpool_bwd", self.lib.segb200_maxpool3x3s2_bwd_idx,
                     (_ptr(idx), _ptr(dy), _ptr(gx), n, h, w_, c, dy.stride(2), gx.stride(2), self.dt), x=x.t, dy=dy, dx=gx, idx=idx)
            self.pool_put(dy)
        self.tape.append(backward)
        return y

    def bilinear(self, x, out, align=True):
        """F.interpolate(x, out.shape, 'bilinear', align_corners=align) into the channel slice `out` (deeplabv3_plus.py:71;
        align_corners=False in _HRNetHead, hrnet_seg.py:57-59)."""
        n, hi, wi, c = x.t.shape
        _, ho, wo, _ = out.t.shape
        self.add("bilinear", self.lib.segb200_bilinear_nhwc,
                 (_ptr(x.t), _ptr(out.t), n, hi, wi, c, x.t.stride(2), ho, wo, out.t.stride(2), int(align), self.dt), x=x.t, y=out.t,
                 align=align)

        def backward():
            dy = out.grad()
            gx, acc = x.take()
            self.add("bilinear_bwd", self.lib.segb200_bilinear_nhwc_bwd,
                     (_ptr(dy), _ptr(gx), n, hi, wi, c, gx.stride(2), ho, wo, dy.stride(2), int(align), int(acc), None, self.dt), dy=dy,
                     dx=gx, accumulate=acc, gscale=None, align=align)
        self.tape.append(backward)

    def upsample_add(self, a, t, k, act=None):
        """y = act(a + nearest_up_{2^k}(t)): one term of the HRNet fuse sum (backbones/hrnet.py:178-186,215-232)"""
        n, h, w_, c = a.t.shape
        assert tuple(t.t.shape) == (n, h >> k, w_ >> k, c) and h % (1 << k) == 0 and w_ % (1 << k) == 0, (a.t.shape, t.t.shape, k)
        y = Act(self, self.new(n, h, w_, c))
        self.add("upsample_add", self.lib.segb200_upsample_add,
                 (_ptr(a.t), _ptr(t.t), _ptr(y.t), n, h, w_, c, a.t.stride(2), t.t.stride(2), y.t.stride(2), k, L.ACT[act], self.dt),
                 a=a.t, z=t.t, y=y.t, k=k, act=act)

        def backward():
            dy = y.grad()
            ga, acc_a = a.take()
            gt, acc_t = t.take()
            self.add("upsample_add_bwd", self.lib.segb200_upsample_add_bwd,
                     (_ptr(dy), _ptr(y.t), _ptr(ga), _ptr(gt), n, h, w_, c, dy.stride(2), y.t.stride(2), ga.stride(2), gt.stride(2), k,
                      L.ACT[act], int(acc_a), int(acc_t), self.dt),
                     dy=dy, y=y.t, da=ga, dz=gt, k=k, act=act, acc_a=acc_a, acc_z=acc_t)
            if y.parent is None:
                self.pool_put(dy)
        self.tape.append(backward)
        return y

    def sum_losses(self, triples):
        """out3 = sum of the per-output (loss, 1/valid, valid) triples (MixSoftmaxCrossEntropyLoss._multiple_forward,
        solver/loss.py:31-36); `triples` is one contiguous [k][3] buffer."""
        k = triples.numel() // 3
        self.add("reduce_partials", self.lib.segb200_reduce_partials, (_ptr(triples), k, 1, 3, _ptr(self.out3), 0, 1, 0, 1.0),
                 partial=triples, slabs=k, K=1, c=3, out=self.out3, sk=0, sc=1, accumulate=0, scale=1.0)

    def dropout(self, x, name):
        """nn.Dropout2d on an activation that is not followed by BatchNorm (DANetHead.conv6/7/8, models/danet.py:64-68): the
        [n][c] mask (already scaled by 1/(1-p)) is an input of the plan, refreshed per step by the trainer."""
        n, h, w_, c = x.t.shape
        mask = self.f32(n, c)
        mask.fill_(1.0)
        self.masks[name] = mask
        z = Act(self, self.new(n, h, w_, c))
        rows, hw = n * h * w_, h * w_
        self.add("bn_apply", self.lib.segb200_bn_apply,
                 (_ptr(x.t), None, None, None, _ptr(mask), _ptr(z.t), rows, hw, c, x.t.stride(2), 0, z.t.stride(2), 0, self.dt),
                 y=x.t, scale=None, shift=None, residual=None, nc_scale=mask, z=z.t, act=None)

        def backward():
            dz = z.grad()
            gx, acc = x.take()
            self.add("bn_bwd_apply", self.lib.segb200_bn_bwd_apply,
                     (_ptr(dz), None, None, None, None, None, None, None, float(rows), _ptr(mask), None, _ptr(gx), int(acc), rows, hw, c,
                      dz.stride(2), 0, 0, 0, gx.stride(2), 0, self.dt),
                     dz=dz, z=None, y=None, st={}, count=float(rows), nc_scale=mask, dy=None, dres=gx, dres_acc=acc, act=None)
            if z.parent is None:
                self.pool_put(dz)
        self.tape.append(backward)
        return z

    def _gamma_vec(self, gname, c):
        """[c] fp32 copies of a scalar parameter (gather_cast with a constant index table): refreshed inside the step because the
        optimizer moves gamma"""
        S = self.S
        idx = torch.full((c,), S.meta[gname]["off"], dtype=torch.int32, device=self.device)
        vec = self.f32(c)
        self.keep.append(idx)
        self.add("gather_cast", self.lib.segb200_gather_cast, (_ptr(S.master), _ptr(idx), _ptr(vec), c, L.F32), src=S.master, index=idx,
                 dst=vec)
        return vec

    def _transpose(self, x2d, rows, cols, x_ld, out, pitch):
        """out[cols][pitch] (16-bit) <- x2d[rows][x_ld] transposed (segb200_nhwc_to_cn; columns rows..pitch-1 of `out` stay zero)"""
        self.add("transpose", self.lib.segb200_nhwc_to_cn, (_ptr(x2d), _ptr(out), 1, cols, rows, x_ld, pitch, self.dt), x=x2d, y=out,
                 rows=rows, cols=cols, x_ld=x_ld, pitch=pitch)

    def _gamma_residual_bwd(self, y, o, x, gname, gvec, c):
        """shared tail of the two attention units, y = gamma * o + x:  dx (+)= dy;  dgamma += sum(dy * o);  -> do = gamma * dy"""
        S = self.S
        dy = y.grad()
        n, h, w_, _ = dy.shape
        rows, hw = n * h * w_, h * w_
        gx, acc = x.take()
        self.add("bn_apply", self.lib.segb200_bn_apply,
                 (_ptr(dy), None, None, _ptr(gx) if acc else None, None, _ptr(gx), rows, hw, c, dy.stride(2), gx.stride(2) if acc else 0,
                  gx.stride(2), 0, self.dt), y=dy, scale=None, shift=None, residual=gx if acc else None, nc_scale=None, z=gx, act=None)
        slabs = self.lib.segb200_reduce_slabs(rows, c, 0)
        partial, sums = self.f32(slabs * 2 * c), self.f32(2, c)
        self.add("bn_bwd_reduce", self.lib.segb200_bn_bwd_reduce,
                 (_ptr(dy), None, _ptr(o), None, None, None, _ptr(partial), rows, hw, c, dy.stride(2), 0, o.stride(2), 0, self.dt, 0),
                 dz=dy, z=None, y=o, st=dict(mean=None, invstd=None, partial=partial), nc_scale=None, act=None, c=c)
        self.add("bn_bwd_finalize", self.lib.segb200_bn_bwd_finalize, (_ptr(partial), slabs, c, None, None, _ptr(sums), None, None),
                 st=dict(partial=partial, slabs=slabs, sums=sums), c=c, dgamma=None, dbeta=None)
        gg = S.view(S.grad, gname)
        self.add("reduce_partials", self.lib.segb200_reduce_partials, (_ptr(sums[1]), c, 1, 1, _ptr(gg), 0, 1, 1, 1.0),
                 partial=sums[1], slabs=c, K=1, c=1, out=gg, sk=0, sc=1, accumulate=1, scale=1.0)
        do = self.pool_get(n, h, w_, c)
        self.add("bn_apply", self.lib.segb200_bn_apply,
                 (_ptr(dy), _ptr(gvec), None, None, None, _ptr(do), rows, hw, c, dy.stride(2), 0, do.stride(2), 0, self.dt),
                 y=dy, scale=gvec, shift=None, residual=None, nc_scale=None, z=do, act=None)
        if y.parent is None:
            self.pool_put(dy)
        return do, gx

    def pam_unit(self, x, prefix):
        """PAM_Module in training (modules/module.py:100-131), attention materialised per image like the reference's
        bmm / softmax / bmm: S = Q K^T -> row_softmax -> P (kept, 16-bit) -> o = P V; y = gamma o + x.  Backward: D = do V^T ->
        row_softmax_bwd -> dV = P^T do, dQ = dS K, dK = dS^T Q (GEMMs over transposed copies); the three 1x1 convs are conv_units."""
        S = self.S
        n, h, w_, c = x.t.shape
        ntok = h * w_
        if ntok % 8 or c % 64:
            raise RuntimeError("segb200: the PAM training unit needs h*w % 8 == 0 and channels % 64 == 0")
        q = self.conv_unit(x, prefix + ".query_conv.weight", bias=prefix + ".query_conv.bias")
        k = self.conv_unit(x, prefix + ".key_conv.weight", bias=prefix + ".key_conv.bias")
        v = self.conv_unit(x, prefix + ".value_conv.weight", bias=prefix + ".value_conv.bias")
        cq = q.t.shape[3]
        if cq % fold.conv_kblock(cq):
            raise RuntimeError("segb200: the PAM training unit needs a query depth that is a multiple of the GEMM K block")
        pitch = fold.round_up(ntok, 64)
        gname = prefix + ".gamma"
        gvec = self._gamma_vec(gname, c)
        one = self.f32(1)
        one.fill_(1.0)
        probs = [self.new(1, 1, ntok, pitch) for _ in range(n)]
        energy = self.f32(1, 1, ntok, ntok)
        vt = self.new(1, 1, c, pitch)[0, 0]
        o = self.new(n, h, w_, c)
        y = Act(self, self.new(n, h, w_, c))
        for b in range(n):
            self.conv(q.t[b].view(1, 1, ntok, cq), k.t[b].view(ntok, 1, cq), energy, cin=cq, cout=ntok)
            self.add("row_softmax", self.lib.segb200_row_softmax, (_ptr(energy), _ptr(probs[b]), ntok, ntok, ntok, pitch, self.dt),
                     s=energy, p=probs[b], n=ntok)
            self._transpose(v.t[b], ntok, c, c, vt, pitch)
            self.conv(probs[b], vt.view(c, 1, pitch), o[b].view(1, 1, ntok, c), cin=pitch, cout=c)
        rows, hw = n * ntok, ntok
        self.add("bn_apply", self.lib.segb200_bn_apply,
                 (_ptr(o), _ptr(gvec), None, _ptr(x.t), None, _ptr(y.t), rows, hw, c, o.stride(2), x.t.stride(2), y.t.stride(2), 0, self.dt),
                 y=o, scale=gvec, shift=None, residual=x.t, nc_scale=None, z=y.t, act=None)

        def backward():
            do, _ = self._gamma_residual_bwd(y, o, x, gname, gvec, c)
            gq, aq = q.take()
            gk, ak = k.take()
            gv, av = v.take()
            assert not (aq or ak or av)
            dmat = self.f32(1, 1, ntok, ntok)
            ds = self.pool_get(1, 1, ntok, pitch)
            tr = self.pool_get(1, 1, ntok, pitch)
            dot = self.pool_get(1, 1, c, pitch)
            kt = self.pool_get(1, 1, cq, pitch)
            part = self.f32(ntok)
            for t in (tr, dot, kt):                           # K tails of the transposed operands must be zero
                self.cur.append(Step("zero", (lambda s_, t=t: t.zero_()), dict(t=t)))
            for b in range(n):
                self.conv(do[b].view(1, 1, ntok, c), v.t[b].view(ntok, 1, c), dmat, cin=c, cout=ntok)
                self.add("row_softmax_bwd", self.lib.segb200_row_softmax_bwd,
                         (_ptr(probs[b]), _ptr(dmat), _ptr(one), _ptr(ds), _ptr(part), ntok, ntok, pitch, ntok, pitch, self.dt),
                         p=probs[b], d=dmat, gamma=one, ds=ds, part=part, n=ntok)
                self._transpose(probs[b][0, 0], ntok, ntok, pitch, tr[0, 0], pitch)
                self._transpose(do[b], ntok, c, do.stride(2), dot[0, 0], pitch)
                self.conv(tr, dot[0, 0].view(c, 1, pitch), gv[b].view(1, 1, ntok, c), cin=pitch, cout=c)
                self._transpose(k.t[b], ntok, cq, cq, kt[0, 0], pitch)
                self.conv(ds, kt[0, 0].view(cq, 1, pitch), gq[b].view(1, 1, ntok, cq), cin=pitch, cout=cq)
                self._transpose(ds[0, 0], ntok, ntok, pitch, tr[0, 0], pitch)
                self._transpose(q.t[b], ntok, cq, cq, kt[0, 0], pitch)
                self.conv(tr, kt[0, 0].view(cq, 1, pitch), gk[b].view(1, 1, ntok, cq), cin=pitch, cout=cq)
            for t in (ds, tr, dot, kt, do):
                self.pool_put(t)
            self.mark_done(gname)
        self.tape.append(backward)
        return y

    def cam_unit(self, x, prefix):
        """CAM_Module in training (modules/module.py:134-162): E = X^T X -> cam_softmax -> A (kept); o = A X; y = gamma o + x.
        Backward (csrc/cam_bwd.cu): G = do^T x -> dE = -A (G - sum A G) -> dx += do A + x (dE + dE^T)."""
        n, h, w_, c = x.t.shape
        ntok = h * w_
        pitch = fold.round_up(ntok, 64)
        cpad = fold.round_up(c, fold.conv_kblock(c))
        gname = prefix + ".gamma"
        gvec = self._gamma_vec(gname, c)
        one = self.f32(1)
        one.fill_(1.0)
        xts = [self.new(1, 1, c, pitch) for _ in range(n)]
        atts = [self.new(1, 1, c, cpad) for _ in range(n)]
        energy = self.f32(1, 1, c, c)
        o = self.new(n, h, w_, c)
        y = Act(self, self.new(n, h, w_, c))
        for b in range(n):
            self._transpose(x.t[b], ntok, c, x.t.stride(2), xts[b][0, 0], pitch)
            self.conv(xts[b], xts[b][0, 0].view(c, 1, pitch), energy, cin=pitch, cout=c)
            self.add("cam_softmax", self.lib.segb200_cam_softmax, (_ptr(energy), _ptr(atts[b]), c, c, c, cpad, self.dt), e=energy,
                     att=atts[b], c=c)
            self.conv(x.t[b:b + 1], atts[b][0, 0].view(c, 1, cpad), o[b:b + 1], cin=c, cout=c)
        rows, hw = n * ntok, ntok
        self.add("bn_apply", self.lib.segb200_bn_apply,
                 (_ptr(o), _ptr(gvec), None, _ptr(x.t), None, _ptr(y.t), rows, hw, c, o.stride(2), x.t.stride(2), y.t.stride(2), 0, self.dt),
                 y=o, scale=gvec, shift=None, residual=x.t, nc_scale=None, z=y.t, act=None)

        def backward():
            do, gx = self._gamma_residual_bwd(y, o, x, gname, gvec, c)
            dot = self.pool_get(1, 1, c, pitch)
            self.cur.append(Step("zero", (lambda s_, t=dot: t.zero_()), dict(t=dot)))
            gmat, de, part = self.f32(1, 1, c, c), self.f32(c, c), self.f32(c)
            w1, w2 = self.pool_get(1, 1, c, cpad), self.pool_get(1, 1, c, cpad)
            for b in range(n):
                self._transpose(do[b], ntok, c, do.stride(2), dot[0, 0], pitch)
                self.conv(dot, xts[b][0, 0].view(c, 1, pitch), gmat, cin=pitch, cout=c)
                self.add("cam_softmax_bwd", self.lib.segb200_cam_softmax_bwd,
                         (_ptr(atts[b]), _ptr(gmat), _ptr(one), _ptr(de), _ptr(part), c, cpad, c, c, self.dt),
                         att=atts[b], g=gmat, gamma=one, de=de, part=part, c=c)
                self.add("cam_bwd_pack", self.lib.segb200_cam_bwd_pack,
                         (_ptr(atts[b]), _ptr(de), _ptr(one), _ptr(w1), _ptr(w2), c, cpad, c, cpad, self.dt),
                         att=atts[b], de=de, gamma=one, w1=w1, w2=w2, c=c)
                self.conv(do[b:b + 1], w1[0, 0].view(c, 1, cpad), gx[b:b + 1], cin=c, cout=c, residual=gx[b:b + 1])
                self.conv(x.t[b:b + 1], w2[0, 0].view(c, 1, cpad), gx[b:b + 1], cin=c, cout=c, residual=gx[b:b + 1])
            for t in (dot, w1, w2, do):
                self.pool_put(t)
            self.mark_done(gname)
        self.tape.append(backward)
        return y

    def image_pool(self, x):
        """nn.AdaptiveAvgPool2d(1) (module.py:52) -> [n,1,1,c]"""
        n, h, w_, c = x.t.shape
        y = Act(self, self.new(n, 1, 1, c))
        self.add("gap", self.lib.segb200_global_avgpool, (_ptr(x.t), _ptr(y.t), n, h, w_, c, x.t.stride(2), self.dt), x=x.t, y=y.t)

        def backward():
            dy = y.grad()
            gx, acc = x.take()
            self.add("nc_broadcast", self.lib.segb200_nc_broadcast,
                     (_ptr(dy), _ptr(gx), n, h * w_, c, dy.stride(2), gx.stride(2), 1.0 / (h * w_), int(acc), self.dt), v=dy, y=gx,
                     scale=1.0 / (h * w_), accumulate=acc)
        self.tape.append(backward)
        return y

    def broadcast(self, v, out):
        """bilinear up-sampling from 1x1 == broadcast (module.py:64) into the channel slice `out`"""
        n, h, w_, c = out.t.shape
        self.add("nc_broadcast", self.lib.segb200_nc_broadcast,
                 (_ptr(v.t), _ptr(out.t), n, h * w_, c, v.t.stride(2), out.t.stride(2), 1.0, 0, self.dt), v=v.t, y=out.t, scale=1.0,
                 accumulate=False)

        def backward():
            dy = out.grad()                                        # [n,h,w,c] slice -> sum over pixels
            mean = self.pool_get(n, 1, 1, c)
            self.add("gap", self.lib.segb200_global_avgpool, (_ptr(dy), _ptr(mean), n, h, w_, c, dy.stride(2), self.dt), x=dy, y=mean)
            gv, acc = v.take()
            assert not acc
            hwv = self.f32(c).fill_(float(h * w_))
            self.add("bn_apply", self.lib.segb200_bn_apply,
                     (_ptr(mean), _ptr(hwv), None, None, None, _ptr(gv), n, 1, c, mean.stride(2), 0, gv.stride(2), 0, self.dt),
                     y=mean, scale=hwv, shift=None, residual=None, nc_scale=None, z=gv, act=None)
            self.pool_put(mean)
        self.tape.append(backward)

    def loss(self, logits, align=True, out3=None):
        """fused F.interpolate(logits, (H, W), align_corners=align) + CrossEntropyLoss(ignore_index) + its gradient.  out3: where
        (mean loss, 1/valid, valid) go -- the plan's own triple by default; models with several outputs pass one triple per output
        and sum them with sum_losses()."""
        out3 = self.out3 if out3 is None else out3
        n, hi, wi, _ = logits.t.shape
        c8 = fold.round_up(self.nclass, 8)
        dfull = self.new(n, self.H, self.W, c8)
        nb = self.lib.segb200_upsample_ce_blocks(n, self.H, self.W)
        partial = self.f32(2 * nb)
        self.add("upsample_ce", self.lib.segb200_upsample_ce,
                 (_ptr(logits.t), _ptr(self.target), _ptr(dfull), _ptr(partial), _ptr(out3), n, hi, wi, self.nclass,
                  logits.t.stride(2), self.H, self.W, dfull.stride(2), int(align), self.ignore_index, self.dt),
                 logits=logits.t, target=self.target, dfull=dfull, out3=out3, nclass=self.nclass, ignore_index=self.ignore_index,
                 align=align)

        def backward():
            gl, acc = logits.take()
            assert not acc
            inv = out3[1:2]
            self.add("bilinear_bwd", self.lib.segb200_bilinear_nhwc_bwd,
                     (_ptr(dfull), _ptr(gl), n, hi, wi, c8, gl.stride(2), self.H, self.W, dfull.stride(2), int(align), 0, _ptr(inv),
                      self.dt), dy=dfull, dx=gl, accumulate=False, gscale=inv, align=align)
        self.tape.append(backward)

    def build_backward(self):
        self.cur = self.bwd
        for fn in reversed(self.tape):
            fn()
        self.cur = self.fwd

    # ---- execution ----
    def run(self, steps=None):
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for st in (steps if steps is not None else self.fwd + self.bwd):
            st.call(s)

    def run_timed(self):
        """Replay with a CUDA event pair around every launch (events on the launching stream) -> [(Step, ms)]."""
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        steps = self.fwd + self.bwd
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(steps) + 1)]
        evs[0].record()
        for i, st in enumerate(steps):
            st.call(s)
            evs[i + 1].record()
        torch.cuda.synchronize()
        return [(st, evs[i].elapsed_time(evs[i + 1])) for i, st in enumerate(steps)]


# ------------------------------------------------------------------------------------------------------------
# model builders (mirror oracle/segref.py <-> the reference forward graphs)
# ------------------------------------------------------------------------------------------------------------
def _bottleneck(pl, x, prefix, planes, stride, dilation, downsample, out=None):
    """BottleneckV1b (backbones/resnet.py:44-81).  The downsample branch is recorded FIRST so that its (possibly strided)
    data gradient is the last contribution to the block input's gradient."""
    idn = x
    if downsample:
        idn = pl.conv_unit(x, prefix + ".downsample.0.weight", prefix + ".downsample.1", None, k=1, stride=stride)
    y = pl.conv_unit(x, prefix + ".conv1.weight", prefix + ".bn1", "relu")
    y = pl.conv_unit(y, prefix + ".conv2.weight", prefix + ".bn2", "relu", k=3, stride=stride, dilation=dilation, pad=dilation)
    return pl.conv_unit(y, prefix + ".conv3.weight", prefix + ".bn3", "relu", residual=idn, out=out)


def _resnet(pl, layers, output_stride, c4_out=None, multi_dilation=None):
    """ResNetV1.forward (backbones/resnet.py:183-199), stride/dilation table :90-100,:149-179."""
    dil, strides = {32: ((1, 1), (2, 2)), 16: ((1, 2), (2, 1)), 8: ((2, 4), (1, 1))}[output_stride]
    p = "encoder"
    n, H, W = pl.n, pl.H, pl.W
    s2d = Act(pl, pl.new(n, (H + 1) // 2, (W + 1) // 2, 16, ld=64), needs_grad=False)
    pl.cur.append(Step("pack_s2d", (lambda s: ops.pack_s2d(pl.x_in, s2d.t._base if s2d.t._base is not None else s2d.t)),
                       dict(x=pl.x_in, out=s2d.t)))
    x = pl.conv_unit(s2d, p + ".conv1.weight", p + ".bn1", "relu", k=7, stride=2, pad=3, stem=True)
    x = pl.maxpool(x)
    inpl = [64]

    def make_layer(x, name, planes, blocks, stride=1, dilation=1, last_out=None, mg=None):
        ds = stride != 1 or inpl[0] != planes * 4
        first_d = mg[0] if mg else (1 if dilation in (1, 2) else 2)            # resnet.py:151-161 (multi-grid: DANet)
        x = _bottleneck(pl, x, f"{p}.{name}.0", planes, stride, first_d, ds)
        inpl[0] = planes * 4
        for i in range(1, blocks):
            d = mg[i % len(mg)] if mg else dilation                              # resnet.py:166-175
            x = _bottleneck(pl, x, f"{p}.{name}.{i}", planes, 1, d, False, out=last_out if i == blocks - 1 else None)
        return x

    c1 = make_layer(x, "layer1", 64, layers[0])
    c2 = make_layer(c1, "layer2", 128, layers[1], 2)
    c3 = make_layer(c2, "layer3", 256, layers[2], strides[0], dil[0])
    c4 = make_layer(c3, "layer4", 512, layers[3], strides[1], dil[1], last_out=c4_out, mg=multi_dilation)
    return c1, c2, c3, c4


def _sepconv(pl, x, prefix, dilation, out=None, planes=None, stride=1, relu_first=False, eps=1e-5, residual=None):
    """SeparableConv2d (modules/basic.py:34-62).  relu_first=False: dw -> BN -> ReLU -> pw -> BN -> ReLU;
    relu_first=True: ReLU -> dw -> BN -> pw -> BN (the leading ReLU is NOT in place: the caller's x stays un-rectified)."""
    b = prefix + ".block"
    z = pl.dw_unit(x, b + ".depthwise.weight", b + ".bn_depth", None if relu_first else "relu", dilation, eps=eps, stride=stride,
                   pre_relu=relu_first)
    return pl.conv_unit(z, b + ".pointwise.weight", b + ".bn_point", None if relu_first else "relu", out=out, eps=eps,
                        residual=residual)


def _xception_block(pl, x, prefix, stride=1, dilation=1, skip="conv", relu_first=True, eps=1e-5):
    """XceptionBlock.forward (backbones/xception.py:32-51): three separable convs + shortcut (1x1 stride-s conv + BN / identity /
    none); NO ReLU after the add (:40-42).  The shortcut is recorded first: its gradient is the last to reach the block input."""
    res = None
    if skip == "conv":
        res = pl.conv_unit(x, prefix + ".conv.weight", prefix + ".bn", None, k=1, stride=stride, eps=eps)
    elif skip == "sum":
        res = x
    sc1 = _sepconv(pl, x, prefix + ".sep_conv1", dilation, relu_first=relu_first, eps=eps)
    sc2 = _sepconv(pl, sc1, prefix + ".sep_conv2", dilation, relu_first=relu_first, eps=eps)
    out = _sepconv(pl, sc2, prefix + ".sep_conv3", dilation, stride=stride, relu_first=relu_first, eps=eps, residual=res)
    return out, sc2


def _xception65(pl, output_stride, eps):
    """Xception65.forward (backbones/xception.py:129-165)."""
    b3s, mid_d, exit_d, exit_s = {32: (2, 1, (1, 1), 2), 16: (2, 1, (1, 2), 1), 8: (1, 2, (2, 4), 1)}[output_stride]
    p = "encoder"
    n, H, W = pl.n, pl.H, pl.W
    s2d = Act(pl, pl.new(n, (H + 1) // 2, (W + 1) // 2, 16, ld=64), needs_grad=False)
    pl.cur.append(Step("pack_s2d", (lambda s: ops.pack_s2d(pl.x_in, s2d.t._base if s2d.t._base is not None else s2d.t)),
                       dict(x=pl.x_in, out=s2d.t)))
    x = pl.conv_unit(s2d, p + ".conv1.weight", p + ".bn1", "relu", k=3, stride=2, pad=1, stem=True, eps=eps)
    x = pl.conv_unit(x, p + ".conv2.weight", p + ".bn2", "relu", k=3, pad=1, eps=eps)
    x, _ = _xception_block(pl, x, p + ".block1", 2, eps=eps)
    x, c1 = _xception_block(pl, x, p + ".block2", 2, eps=eps)
    x, _ = _xception_block(pl, x, p + ".block3", b3s, eps=eps)
    for i in range(4, 20):
        x, _ = _xception_block(pl, x, f"{p}.block{i}", 1, mid_d, "sum", eps=eps)
    x, _ = _xception_block(pl, x, p + ".block20", exit_s, exit_d[0], eps=eps)
    c4, _ = _xception_block(pl, x, p + ".block21", 1, exit_d[1], "none", False, eps)
    return c1, c4


def _inverted_residual(pl, x, prefix, cout, stride, expand, dilation, eps):
    """InvertedResidual (modules/basic.py:139-163): [pw + BN + ReLU6] -> dw(stride, dil) + BN + ReLU6 -> pw + BN (+ x)."""
    cin = x.t.shape[3]
    y, i = x, 0
    if expand != 1:
        y = pl.conv_unit(y, f"{prefix}.conv.{i}.conv.weight", f"{prefix}.conv.{i}.bn", "relu6", eps=eps)
        i += 1
    y = pl.dw_unit(y, f"{prefix}.conv.{i}.conv.weight", f"{prefix}.conv.{i}.bn", "relu6", dilation, eps=eps, stride=stride)
    i += 1
    res = x if (stride == 1 and cin == cout) else None
    return pl.conv_unit(y, f"{prefix}.conv.{i}.weight", f"{prefix}.conv.{i + 1}", None, eps=eps, residual=res)


def _mobilenet_v2(pl, output_stride, eps):
    """MobileNetV2.forward (backbones/mobilenet.py:131-143), incl. the first-block-only dilation quirk (:125 vs :128)."""
    dil = {32: (1, 1), 16: (1, 2), 8: (2, 4)}[output_stride]
    setting = [[1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1], [6, 160, 3, 2], [6, 320, 1, 1]]
    p = "encoder"
    n, H, W = pl.n, pl.H, pl.W
    s2d = Act(pl, pl.new(n, (H + 1) // 2, (W + 1) // 2, 16, ld=64), needs_grad=False)
    pl.cur.append(Step("pack_s2d", (lambda s: ops.pack_s2d(pl.x_in, s2d.t._base if s2d.t._base is not None else s2d.t)),
                       dict(x=pl.x_in, out=s2d.t)))
    x = pl.conv_unit(s2d, p + ".conv1.conv.weight", p + ".conv1.bn", "relu6", k=3, stride=2, pad=1, stem=True, eps=eps)

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
    return c1, c4


def _aspp(pl, c4, prefix, output_stride):
    """_ASPP.forward (modules/module.py:62-77) in training mode (Dropout2d active, :75)."""
    d = {16: (6, 12, 18), 8: (12, 24, 36), 32: (6, 12, 18)}[output_stride]
    n, h, w_, c = c4.t.shape
    cat = Act(pl, pl.new(n, h, w_, 1280))
    pooled = pl.image_pool(c4)
    pf = pl.conv_unit(pooled, prefix + ".image_pooling.conv.weight", prefix + ".image_pooling.bn", "relu")
    pl.broadcast(pf, cat.slice(0, 256))
    pl.conv_unit(c4, prefix + ".aspp0.conv.weight", prefix + ".aspp0.bn", "relu", out=cat.slice(256, 512))
    for i in range(3):
        _sepconv(pl, c4, f"{prefix}.aspp{i + 1}", d[i], out=cat.slice(512 + 256 * i, 768 + 256 * i))
    mask = pl.f32(n, 256).fill_(1.0)
    pl.masks[prefix + ".dropout"] = mask
    return pl.conv_unit(cat, prefix + ".conv.weight", prefix + ".bn", "relu", nc_scale=mask)


def build_deeplabv3plus_train(pl, backbone="resnet101", output_stride=16, eps_encoder=1e-5, use_aspp=True, use_decoder=True):
    """DeepLabV3Plus.forward + _DeepLabHead (models/deeplabv3_plus.py:33-75) + the loss of solver/loss.py:16-46 (aux off).
    use_aspp / use_decoder False = the MobileNetV2 YAML (configs/cityscapes_deeplabv3_plus_mobilenet.yaml:21-23)."""
    if backbone == "xception65":
        c1, c4 = _xception65(pl, output_stride, eps_encoder)
    elif backbone == "mobilenet_v2":
        c1, c4 = _mobilenet_v2(pl, output_stride, eps_encoder)
    else:
        layers = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3), "resnet152": (3, 8, 36, 3)}[backbone]
        c1, _, _, c4 = _resnet(pl, layers, output_stride)
    x = c4
    if use_aspp:
        x = _aspp(pl, x, "head.aspp", output_stride)
    if use_decoder:
        n, h1, w1, _ = c1.t.shape
        cat = Act(pl, pl.new(n, h1, w1, 304))
        pl.bilinear(x, cat.slice(0, 256))
        pl.conv_unit(c1, "head.c1_block.conv.weight", "head.c1_block.bn", "relu", out=cat.slice(256, 304))
        x = cat
    x = _sepconv(pl, x, "head.block.0", 1)
    x = _sepconv(pl, x, "head.block.1", 1)
    n, hl, wl, _ = x.t.shape
    logits = Act(pl, pl.new(n, hl, wl, fold.round_up(pl.nclass, 8), ld=32))
    pl.conv_unit(x, "head.block.2.weight", bias="head.block.2.bias", out=logits)
    pl.logits = logits
    pl.loss(logits)
    pl.build_backward()
    return pl


def build_ccnet_train(pl, output_stride=16, recurrence=2):
    """CCNet.forward + _CCHead / _RCCAModule (models/ccnet.py:27-82) + the loss of solver/loss.py:16-46 (aux off): ResNet101,
    conva -> criss-cross attention x RECURRENCE (shared weights) -> convb -> cat[c4, out] -> bottleneck (3x3 + BN + Dropout2d)
    -> 1x1 classifier.  c4 is produced straight into its channel slice of the concat buffer."""
    n, H, W = pl.n, pl.H, pl.W
    hh, ww = H, W
    for s_ in ([2, 2, 2] + ([2] if output_stride >= 16 else []) + ([2] if output_stride == 32 else [])):
        hh, ww = (hh - 1) // s_ + 1, (ww - 1) // s_ + 1
    cat = Act(pl, pl.new(n, hh, ww, 2048 + 512))
    _, _, _, c4 = _resnet(pl, (3, 4, 23, 3), output_stride, c4_out=cat.slice(0, 2048))
    hp = "head.rcca"
    out = pl.conv_unit(c4, hp + ".conva.0.weight", hp + ".conva.1", "relu", k=3, pad=1)
    for _ in range(recurrence):
        out = pl.cca_unit(out, hp + ".cca")
    pl.conv_unit(out, hp + ".convb.0.weight", hp + ".convb.1", "relu", k=3, pad=1, out=cat.slice(2048, 2560))
    mask = pl.f32(n, 512).fill_(1.0)
    pl.masks[hp + ".bottleneck.dropout"] = mask
    y = pl.conv_unit(cat, hp + ".bottleneck.0.weight", hp + ".bottleneck.1", None, k=3, pad=1, nc_scale=mask)
    logits = Act(pl, pl.new(n, hh, ww, fold.round_up(pl.nclass, 8), ld=32))
    pl.conv_unit(y, "head.out.weight", bias="head.out.bias", out=logits)
    pl.logits = logits
    pl.loss(logits)
    pl.build_backward()
    return pl


def _hr_basic_block(pl, x, prefix):
    """BasicBlock (backbones/hrnet.py:25-55): conv3x3-BN-ReLU, conv3x3-BN, + x, ReLU"""
    y = pl.conv_unit(x, prefix + ".conv1.weight", prefix + ".bn1", "relu", k=3, pad=1)
    return pl.conv_unit(y, prefix + ".conv2.weight", prefix + ".bn2", "relu", k=3, pad=1, residual=x)


def _hr_module(pl, xs, prefix, blocks):
    """HighResolutionModule.forward (backbones/hrnet.py:215-232) with the fuse layers of _make_fuse_layers (:165-209).  The sum of
    output i starts from the identity term x_i, the stride-2 chains (j < i) are added through their last BatchNorm's residual
    operand, the 1x1 + BN + nearest-up terms (j > i) through upsample_add; the last term carries the ReLU."""
    nb = len(xs)
    xs = list(xs)
    for i in range(nb):
        for b in range(blocks[i]):
            xs[i] = _hr_basic_block(pl, xs[i], f"{prefix}.branches.{i}.{b}")
    if nb == 1:
        return xs
    outs = []
    for i in range(nb):
        terms = [j for j in range(nb) if j != i]
        cur = xs[i]
        for q, j in enumerate(terms):
            act = "relu" if q == len(terms) - 1 else None
            f = f"{prefix}.fuse_layers.{i}.{j}"
            if j > i:
                t = pl.conv_unit(xs[j], f + ".0.weight", f + ".1", None)
                cur = pl.upsample_add(cur, t, j - i, act)
            else:
                t = xs[j]
                for k in range(i - j):
                    if k == i - j - 1:
                        cur = pl.conv_unit(t, f"{f}.{k}.0.weight", f"{f}.{k}.1", act, k=3, stride=2, pad=1, residual=cur)
                    else:
                        t = pl.conv_unit(t, f"{f}.{k}.0.weight", f"{f}.{k}.1", "relu", k=3, stride=2, pad=1)
        outs.append(cur)
    return outs


def build_hrnet_train(pl, hcfg):
    """HighResolutionNet.forward (backbones/hrnet.py:429-479) + _HRNetHead (models/hrnet_seg.py:32-63) + the final bilinear
    (align_corners=False, hrnet_seg.py:28) fused into the loss."""
    p = "encoder"
    n, H, W = pl.n, pl.H, pl.W
    if H % 32 or W % 32:
        raise RuntimeError("segb200: the HRNet training plan needs input sizes that are multiples of 32 (the reference's nearest "
                           "up-sampling + add has the same requirement, backbones/hrnet.py:178-186)")
    s2d = Act(pl, pl.new(n, H // 2, W // 2, 16, ld=64), needs_grad=False)
    pl.cur.append(Step("pack_s2d", (lambda s: ops.pack_s2d(pl.x_in, s2d.t._base if s2d.t._base is not None else s2d.t)),
                       dict(x=pl.x_in, out=s2d.t)))
    x = pl.conv_unit(s2d, p + ".conv1.weight", p + ".bn1", "relu", k=3, stride=2, pad=1, stem=True)
    x = pl.conv_unit(x, p + ".conv2.weight", p + ".bn2", "relu", k=3, stride=2, pad=1)
    planes, inpl = hcfg["stage1"]["channels"][0], 64
    for b in range(hcfg["stage1"]["blocks"][0]):
        x = _bottleneck(pl, x, f"{p}.layer1.{b}", planes, 1, 1, inpl != planes * 4)
        inpl = planes * 4
    pre, ys = [inpl], [x]
    for si, sname in enumerate(("stage2", "stage3", "stage4")):
        sc = hcfg[sname]
        cur = sc["channels"]
        tname = f"{p}.transition{si + 1}"
        xs = []
        for i in range(len(cur)):                                            # _make_transition_layer, hrnet.py:347-381
            if i < len(pre):
                xs.append(ys[i] if cur[i] == pre[i] else
                          pl.conv_unit(ys[i], f"{tname}.{i}.0.weight", f"{tname}.{i}.1", "relu", k=3, pad=1))
            else:
                t = ys[-1]
                for j in range(i + 1 - len(pre)):
                    t = pl.conv_unit(t, f"{tname}.{i}.{j}.0.weight", f"{tname}.{i}.{j}.1", "relu", k=3, stride=2, pad=1)
                xs.append(t)
        for m in range(sc["modules"]):
            xs = _hr_module(pl, xs, f"{p}.{sname}.{m}", sc["blocks"])
        ys, pre = xs, cur
    _, h0, w0, _ = ys[0].t.shape
    ctot = sum(t.t.shape[3] for t in ys)
    cat = Act(pl, pl.new(n, h0, w0, ctot))
    off = 0
    for t in ys:                                                             # hrnet_seg.py:55-61 (same-size resize == identity)
        c = t.t.shape[3]
        pl.bilinear(t, cat.slice(off, off + c), align=False)
        off += c
    hp = "hrnet_head.last_layer"
    y = pl.conv_unit(cat, hp + ".0.weight", hp + ".1", "relu", bias=hp + ".0.bias")
    k = hcfg["final_conv_kernel"]
    logits = Act(pl, pl.new(n, h0, w0, fold.round_up(pl.nclass, 8), ld=32))
    pl.conv_unit(y, hp + ".3.weight", bias=hp + ".3.bias", k=k, pad=1 if k == 3 else 0, out=logits)
    pl.logits = logits
    pl.loss(logits, align=False)
    pl.build_backward()
    return pl


def build_danet_train(pl, output_stride=8, multi_dilation=(4, 8, 16)):
    """DANet.forward (models/danet.py:26-41) + DANetHead (:44-88): ResNet101 (OS8, multi-grid) -> conv5a -> PAM -> conv51 and
    conv5c -> CAM -> conv52; three classifiers, each behind its own Dropout2d, on sa_conv, sc_conv and their sum; the loss is
    the sum of the three cross-entropies (solver/loss.py:31-36)."""
    _, _, _, c4 = _resnet(pl, (3, 4, 23, 3), output_stride, multi_dilation=list(multi_dilation) if multi_dilation else None)
    n, hh, ww, _ = c4.t.shape
    hd = "head"
    feat1 = pl.conv_unit(c4, hd + ".conv5a.0.weight", hd + ".conv5a.1", "relu", k=3, pad=1)
    sa_conv = pl.conv_unit(pl.pam_unit(feat1, hd + ".sa"), hd + ".conv51.0.weight", hd + ".conv51.1", "relu", k=3, pad=1)
    feat2 = pl.conv_unit(c4, hd + ".conv5c.0.weight", hd + ".conv5c.1", "relu", k=3, pad=1)
    sc_conv = pl.conv_unit(pl.cam_unit(feat2, hd + ".sc"), hd + ".conv52.0.weight", hd + ".conv52.1", "relu", k=3, pad=1)
    feat_sum = pl.upsample_add(sa_conv, sc_conv, 0)
    triples = pl.f32(3, 3)
    c8 = fold.round_up(pl.nclass, 8)
    # output order of DANetHead.forward: (sasc, sa, sc); the dropout masks are drawn in the order conv6, conv7, conv8
    for i, (src, name) in enumerate(((feat_sum, "conv8"), (sa_conv, "conv6"), (sc_conv, "conv7"))):
        logits = Act(pl, pl.new(n, hh, ww, c8, ld=32))
        pl.conv_unit(pl.dropout(src, f"{hd}.{name}.0"), f"{hd}.{name}.1.weight", bias=f"{hd}.{name}.1.bias", out=logits)
        if i == 0:
            pl.logits = logits
        pl.loss(logits, out3=triples[i])
    pl.sum_losses(triples)
    pl.build_backward()
    return pl


class DeepLabV3PlusTrainerB200:
    """``trainer.step(images_nchw_fp32, targets_int64) -> loss`` : one iteration of tools/train.py:135-147 (forward, criterion,
    zero_grad, backward, optimizer.step) for DeepLabV3_Plus / ResNet on the CUDA engine.  ``state_dict()`` returns reference-named
    tensors.  Hyper-parameters default to the reference's (config/settings.py:65-75: momentum 0.9, weight decay 1e-4, decoder
    LR x10; LR from the YAML, cityscapes_deeplabv3_plus_resnet.yaml:15).  Multi-GPU: construct under an initialised
    torch.distributed NCCL group; gradients are averaged over ranks by bucketed all-reduces overlapped with backward."""

    def __init__(self, state_dict, backbone="resnet101", nclass=19, output_stride=16, eps_encoder=None, use_aspp=None,
                 use_decoder=None, dtype=torch.bfloat16, device="cuda",
                 lr=0.02, momentum=0.9, weight_decay=1e-4, decoder_lr_factor=10.0, bn_momentum=0.1, dropout=True,
                 bucket_mb=25, cuda_graph=False, sync_bn=True, fused_sync_bn=True, grad_comm_dtype=torch.float32):
        if not ops._PLAN_DRY_RUN and not torch.cuda.is_available():
            raise RuntimeError("segb200: a CUDA device (sm_100a) is required; there is no CPU fallback")
        self.device = torch.device(device)
        self.dtype = dtype
        # cfg.MODEL.BN_EPS_FOR_ENCODER (1e-3 in cityscapes_deeplabv3_plus.yaml:20, applied by solver/optimizer.py:18-20)
        lite = backbone == "mobilenet_v2"         # cfg.MODEL.DEEPLABV3_PLUS.USE_ASPP / ENABLE_DECODER are False in the MobileNet YAML
        self.cfg = dict(backbone=backbone, output_stride=output_stride,
                        eps_encoder=eps_encoder if eps_encoder is not None else (1e-3 if backbone == "xception65" else 1e-5),
                        use_aspp=(not lite) if use_aspp is None else use_aspp,
                        use_decoder=(not lite) if use_decoder is None else use_decoder)
        self.nclass, self.bn_momentum = nclass, bn_momentum
        self.lr, self.momentum, self.weight_decay, self.decoder_lr_factor = lr, momentum, weight_decay, decoder_lr_factor
        self.dropout = dropout
        self.bucket_bytes = int(bucket_mb * 2 ** 20)
        self.cuda_graph = cuda_graph
        self.sync_bn = sync_bn                  # cfg.TRAIN.SYNC_BATCH_NORM (config/settings.py:59): only matters when world > 1
        # dtype of the gradient all-reduce payload: fp32 = what the reference's DDP moves (tools/train.py:108-111, 191 MB per step);
        # bf16 halves the NVLink bytes at the price of one rounding of every summand (opt-in: the sums differ from DDP's in the last bits)
        if grad_comm_dtype not in (torch.float32, torch.bfloat16):
            raise RuntimeError("segb200: grad_comm_dtype must be torch.float32 or torch.bfloat16")
        self.grad_comm_dtype = grad_comm_dtype
        stem = "encoder.conv1.conv.weight" if backbone == "mobilenet_v2" else "encoder.conv1.weight"
        self.store = ParamStore({k: v.detach() for k, v in state_dict.items()}, self.device, dtype, stem=stem)
        self.plans = {}
        self.world = 1
        self.dist = None
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                self.dist, self.world = dist, dist.get_world_size()
        except Exception:                                            # pragma: no cover
            self.dist = None
        # SyncBatchNorm exchange over NVLink peer memory (csrc/syncbn.cu); falls back -- loudly -- to per-layer NCCL all-reduces when
        # symmetric memory cannot be set up (and always under gloo: the CPU tests interpret the all_reduce form)
        self.xchg = None
        if os.environ.get("SEGB200_NO_FUSED_SYNCBN"):                # A/B switch: per-layer NCCL all-reduces (round-1 form)
            fused_sync_bn = False
        if self.dist is not None and sync_bn and fused_sync_bn and self.device.type == "cuda" and self.dist.get_backend() == "nccl":
            try:
                from .parallel import SyncExchange
                cmax = max([v.numel() for k, v in state_dict.items() if k.endswith("running_mean")] + [64])
                self.xchg = SyncExchange(self.dist, self.device, fold.round_up(cmax, 128))
            except Exception as e:                                   # noqa: BLE001
                import sys
                print(f"[segb200] fused SyncBatchNorm exchange unavailable ({type(e).__name__}: {e}); using NCCL all-reduces",
                      file=sys.stderr)
                self.xchg = None

    def plan_for(self, shape):
        shape = tuple(shape)
        if shape not in self.plans:
            pl = TrainPlan(self.store, shape, self.nclass, self.dtype, self.device, self.bn_momentum, dist=self.dist,
                           sync_bn=self.sync_bn, xchg=self.xchg)
            self._build(pl)
            self.plans[shape] = dict(plan=pl, graph=None, buckets=self._buckets(pl))
        return self.plans[shape]

    def _build(self, pl):
        build_deeplabv3plus_train(pl, **self.cfg)

    def _buckets(self, pl):
        """contiguous ranges of the flat gradient, cut where backward has finished everything above an offset"""
        S = self.store
        names = sorted(S.meta, key=lambda k: -S.meta[k]["off"])          # from the end of the flat buffer (= start of backward)
        out, hi, pos, prev = [], S.total, 0, 0
        for k in names:
            lo = S.meta[k]["off"]
            pos = max(pos, pl.done_at.get(k, len(pl.bwd)))
            if (hi - lo) * 4 >= self.bucket_bytes or lo == 0:
                pos = max(pos, prev)                                      # launch order must follow the step order
                out.append((pos, lo, hi))
                hi, prev, pos = lo, pos, 0
        return out

    # ---- pieces of a step (public so that tests can check gradients before the update) ----
    def pack_weights(self):
        S, lib = self.store, L.load()
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if S.idx16 is not None:
            L.check(lib.segb200_gather_cast(_ptr(S.master), _ptr(S.idx16), _ptr(S.w16), S.idx16.numel(), ops.dt_code(self.dtype), s),
                    "gather_cast")
        if S.idx32 is not None:
            L.check(lib.segb200_gather_cast(_ptr(S.master), _ptr(S.idx32), _ptr(S.w32), S.idx32.numel(), L.F32, s), "gather_cast")

    def forward_backward(self, x, target, dropout_masks=None):
        """-> loss (device scalar tensor); gradients (SUMMED over ranks; optimizer_step divides by the world size) are left in
        ``self.store.grad``."""
        if not x.is_cuda or not target.is_cuda:
            raise RuntimeError("segb200: inputs must be CUDA tensors (no CPU implementation)")
        st = self.plan_for(x.shape)
        pl = st["plan"]
        pl.x_in.copy_(x)
        pl.target.copy_(target)
        for name, m in pl.masks.items():
            if dropout_masks is not None and name in dropout_masks:
                m.copy_(dropout_masks[name].reshape(m.shape))
            elif self.dropout:
                m.bernoulli_(0.9).div_(0.9)                  # nn.Dropout2d(0.1): [N,C] keep mask / (1-p)  (module.py:60)
            else:
                m.fill_(1.0)
        self.store.grad.zero_()
        self.store.steps += 1
        if self.xchg is not None:                        # new epoch for this step's SyncBatchNorm exchanges
            L.check(L.load().segb200_counter_add(_ptr(self.xchg.epoch), 1, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                    "counter_add")
        if self.dist is None and self.cuda_graph:
            # single-GPU step as ONE CUDA graph (operand packing + the ~1.1k launches): same kernels, same order, bit-identical
            # results; removes the inter-launch gaps of the ~350 tiny BatchNorm finalize / reduce kernels.  The first call of a
            # shape runs eagerly (module loading, attribute setting), the second is captured, later ones replay.
            calls = st.setdefault("graph_calls", 0)
            st["graph_calls"] = calls + 1
            if st["graph"] is not None:
                st["graph"].replay()
                return pl.out3[0]
            if calls == 1:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.pack_weights()
                    pl.run()
                st["graph"] = g
                g.replay()
                return pl.out3[0]
        self.pack_weights()
        if self.dist is None:
            pl.run()
        else:
            pl.run(pl.fwd)
            works, pos0 = [], 0
            for pos, lo, hi in st["buckets"]:
                pl.run(pl.bwd[pos0:pos])
                pos0 = pos
                g = self.store.grad[lo:hi]
                if self.grad_comm_dtype == torch.float32:
                    works.append((self.dist.all_reduce(g, async_op=True), None, None))
                else:                                # bf16 payload: cast -> all-reduce -> copy back (after the wait)
                    t = g.to(self.grad_comm_dtype)
                    works.append((self.dist.all_reduce(t, async_op=True), g, t))
            pl.run(pl.bwd[pos0:])
            for w, g, t in works:
                w.wait()
                if t is not None:
                    g.copy_(t)
        return pl.out3[0]

    def optimizer_step(self, lr=None):
        S, lib = self.store, L.load()
        lr = self.lr if lr is None else lr
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        ne = S.n_encoder
        gs = 1.0 / self.world
        L.check(lib.segb200_sgd_step(_ptr(S.master), _ptr(S.grad), _ptr(S.mom), ne, lr, self.momentum, self.weight_decay, gs, s), "sgd")
        if S.total > ne:
            L.check(lib.segb200_sgd_step(_ptr(S.master[ne:]), _ptr(S.grad[ne:]), _ptr(S.mom[ne:]), S.total - ne,
                                         lr * self.decoder_lr_factor, self.momentum, self.weight_decay, gs, s), "sgd")

    def step(self, x, target, lr=None, dropout_masks=None):
        loss = self.forward_backward(x, target, dropout_masks)
        self.optimizer_step(lr)
        return loss

    def state_dict(self):
        return self.store.state_dict()

    def collectives_per_step(self, shape):
        """NCCL collectives issued per step: gradient buckets + (only without the fused exchange) 2 per BatchNorm layer"""
        st = self.plan_for(shape)
        pl = st["plan"]
        return (len(st["buckets"]) if self.dist is not None else 0) + sum(1 for s in pl.fwd + pl.bwd if s.kind == "allreduce")

    def n_launches(self, shape):
        pl = self.plan_for(shape)["plan"]
        return len(pl.fwd) + len(pl.bwd) + 5


class CCNetTrainerB200(DeepLabV3PlusTrainerB200):
    """``trainer.step(images, targets)`` for CCNet / ResNet101 (models/ccnet.py): the same engine, with the criss-cross attention
    forward + backward kernels in the launch list (recurrence 2, shared weights).  Plan verified against the oracle in fp64
    (tests/test_train_plan_cpu.py); every launch of a real bf16 step replayed on the B200 (tests/test_train_model_gpu.py)."""

    def __init__(self, state_dict, nclass=19, output_stride=16, recurrence=2, lr=0.003, **kw):
        kw["lr"] = lr                                  # configs/cityscapes_ccnet_resnet.yaml: SOLVER.LR 0.003
        super().__init__(state_dict, backbone="resnet101", nclass=nclass, output_stride=output_stride, use_aspp=False,
                         use_decoder=False, **kw)
        self.recurrence = recurrence

    def _build(self, pl):
        build_ccnet_train(pl, self.cfg["output_stride"], self.recurrence)


class HRNetTrainerB200(DeepLabV3PlusTrainerB200):
    """``trainer.step(images, targets)`` for HRNet (models/hrnet_seg.py + backbones/hrnet.py; default: hrnet_w18_small_v1).
    Defaults follow configs/cityscapes_hrnet_w18_small_v1.yaml: BatchNorm momentum 0.01 (MODEL.BN_MOMENTUM, applied by
    solver/optimizer.py:37-39) and no decoder LR factor.  Plan verified against the oracle in fp64 on the CPU
    (tests/test_train_plan_cpu.py); every launch of a real step replayed on the B200: tests/test_train_model_gpu.py."""

    def __init__(self, state_dict, nclass=19, hcfg=None, bn_momentum=0.01, decoder_lr_factor=1.0, lr=0.01, **kw):
        from .engine import HRNET_W18_SMALL_V1
        self.hcfg = hcfg or HRNET_W18_SMALL_V1
        super().__init__(state_dict, backbone="hrnet", nclass=nclass, use_aspp=False, use_decoder=False, bn_momentum=bn_momentum,
                         decoder_lr_factor=decoder_lr_factor, lr=lr, dropout=False, **kw)

    def _build(self, pl):
        build_hrnet_train(pl, self.hcfg)



class DANetTrainerB200(DeepLabV3PlusTrainerB200):
    """``trainer.step(images, targets)`` for DANet / ResNet101 (models/danet.py; OS8, multi-grid 4/8/16): position and channel
    attention run with the attention matrices materialised per image (as the reference does), three classifiers behind three
    Dropout2d layers, loss = sum of the three cross-entropies.  Plan verified against the oracle in fp64 on the CPU
    (tests/test_train_plan_cpu.py); every launch of a real step replayed on the B200 (tests/test_train_model_gpu.py).  Needs (H/8)*(W/8) % 8 == 0."""

    def __init__(self, state_dict, nclass=19, output_stride=8, multi_dilation=(4, 8, 16), lr=0.003, **kw):
        kw["lr"] = lr                                  # configs/cityscapes_danet_resnet.yaml: SOLVER.LR 0.003
        super().__init__(state_dict, backbone="resnet101", nclass=nclass, output_stride=output_stride, use_aspp=False,
                         use_decoder=False, **kw)
        self.multi_dilation = multi_dilation

    def _build(self, pl):
        build_danet_train(pl, self.cfg["output_stride"], self.multi_dilation)

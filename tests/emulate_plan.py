"""TEST INFRASTRUCTURE: a torch/CPU interpreter for the training launch list (segmentron_b200.train.TrainPlan).

The product only ever executes ``Step.call`` (the C-ABI kernel launch).  This interpreter executes ``Step.info`` instead --
the same buffers, the documented semantics of each kernel (include/segb200.h) restated with torch ops -- so that the HOST
logic of the plan (buffer routing, overwrite/accumulate decisions, channel-slice gradients, index tables, parameter layout,
bucket ordering) can be checked against the oracle on a machine without a GPU.  It is never imported by the package.
"""
import torch
import torch.nn.functional as F

CD = torch.float64          # compute dtype of the interpreter (fp64: no ReLU-mask flips from rounding noise)


def _nchw(t):
    return t.permute(0, 3, 1, 2).to(CD)


def _act(v, act):
    return F.relu(v) if act == "relu" else (F.relu6(v) if act == "relu6" else v)


def _conv_core(x_nchw, W, info, ho, wo):
    k, s, d, p = info["k"], info["stride"], info["dilation"], info["pad"]
    big = d * (k - 1) + s * max(ho, wo) + 8
    xp = F.pad(x_nchw, (p, big, p, big))
    return F.conv2d(xp, W, None, s, 0, d)[:, :, :ho, :wo]


def _unpack_w(w, cout, cin, k):
    return w[:cout, :, :cin].reshape(cout, k, k, cin).permute(0, 3, 1, 2).to(CD)


def run_step(st):
    i, kind = st.info, st.kind
    if kind == "zero":
        i["t"].zero_()
    elif kind == "allreduce":
        st.call(None)                       # a torch.distributed collective, not a kernel: executed as is (gloo in the CPU tests)
    elif kind == "pack_s2d":
        x, out = i["x"], i["out"]
        n, c, h, w = x.shape
        base = out._base if out._base is not None else out
        base.zero_()
        hs, ws = base.shape[1], base.shape[2]
        xp = F.pad(x, (0, 2 * ws - w, 0, 2 * hs - h))
        for dy in range(2):
            for dx in range(2):
                base[..., (dy * 2 + dx) * c:(dy * 2 + dx + 1) * c] = xp[:, :, dy::2, dx::2].permute(0, 2, 3, 1).to(base.dtype)
    elif kind == "conv":
        x, w, y = i["x"], i["w"], i["y"]
        cin, cout, k = i["cin"], i["cout"], i["k"]
        xin = x if x.shape[3] == cin else x.as_strided((x.shape[0], x.shape[1], x.shape[2], cin), x.stride())
        o = _conv_core(_nchw(xin), _unpack_w(w, cout, cin, k), i, y.shape[1], y.shape[2]).permute(0, 2, 3, 1)
        if i["shift"] is not None:
            sh = torch.zeros(cout, dtype=CD, device=x.device)
            m = min(cout, i["shift"].numel())
            sh[:m] = i["shift"][:m]
            o = o + sh
        if i["residual"] is not None:
            o = o + i["residual"][..., :cout].to(CD)
        y[..., :cout] = o.to(y.dtype)
    elif kind == "wgrad":
        x, dy, dw = i["x"], i["dy"], i["dw"]
        cin, cout, k = i["cin"], i["cout"], i["k"]
        W0 = torch.zeros(cout, cin, k, k, dtype=CD, device=x.device, requires_grad=True)
        o = _conv_core(_nchw(x[..., :cin]), W0, i, dy.shape[1], dy.shape[2])
        o.backward(_nchw(dy[..., :cout]))
        dw.view(cout, k * k, cin).add_(W0.grad.permute(0, 2, 3, 1).reshape(cout, k * k, cin).to(dw.dtype))
    elif kind == "dw":
        x, w, y, d = i["x"], i["w"], i["y"], i["dilation"]
        c = x.shape[3]
        W = w.t().reshape(c, 1, 3, 3).to(CD)
        xin = F.relu(_nchw(x)) if i.get("pre_relu") else _nchw(x)
        y.copy_(F.conv2d(xin, W, None, i.get("stride", 1), d, d, groups=c).permute(0, 2, 3, 1).to(y.dtype))
    elif kind == "bn_stats":
        x, partial, c = i["x"], i["partial"], i["c"]
        xf = x.to(CD).reshape(-1, c)
        partial.zero_()
        partial[:c] = xf.sum(0)
        partial[c:2 * c] = (xf * xf).sum(0)
    elif kind in ("bn_finalize", "bn_finalize_sync"):
        c, cnt, st_ = i["c"], i["count"], i["st"]
        p = i["partial"].view(i["slabs"], 2, c).double().sum(0)
        if kind == "bn_finalize_sync":              # the kernel's peer-memory exchange == a SUM over ranks of the local [2][c] sums
            p = p.contiguous()
            i["dist"].all_reduce(p)
        m = p[0] / cnt
        var = (p[1] / cnt - m * m).clamp(min=0)
        inv = 1.0 / torch.sqrt(var + i["eps"])
        st_["mean"].copy_(m.to(CD)); st_["invstd"].copy_(inv.to(CD))
        sc = i["gamma"].double() * inv
        st_["scale"].copy_(sc.to(CD)); st_["shift"].copy_((i["beta"].double() - m * sc).to(CD))
        mom = i["momentum"]
        i["rm"].mul_(1 - mom).add_((mom * m).to(i["rm"].dtype))
        i["rv"].mul_(1 - mom).add_((mom * (var * cnt / (cnt - 1) if cnt > 1 else var)).to(i["rv"].dtype))
    elif kind == "bn_apply":
        y, z = i["y"], i["z"]
        v = y.to(CD)
        if i["scale"] is not None:
            v = v * i["scale"]
        if i["shift"] is not None:
            v = v + i["shift"]
        if i["residual"] is not None:
            v = v + i["residual"].to(CD)
        v = _act(v, i["act"])
        if i["nc_scale"] is not None:
            v = v * i["nc_scale"][:, None, None, :]
        z.copy_(v.to(z.dtype))
    elif kind in ("bn_bwd_reduce", "bn_bwd_apply"):
        dz, z, y, st_, act = i["dz"], i["z"], i["y"], i["st"], i["act"]
        g = dz.to(CD)
        if i["nc_scale"] is not None:
            g = g * i["nc_scale"][:, None, None, :]
        if act is not None:
            # mask source: the stored output when given, else the pre-activation recomputed from y (kernel contract)
            pre = z.to(CD) if z is not None else y.to(CD) * st_["scale"] + st_["shift"]
            g = g * ((pre > 0) if act == "relu" else ((pre > 0) & (pre < 6)))
        mean = st_["mean"] if st_.get("mean") is not None else 0.0
        inv = st_["invstd"] if st_.get("invstd") is not None else 1.0
        xh = (y.to(CD) - mean) * inv if y is not None else None
        c = g.shape[3]
        if kind == "bn_bwd_reduce":
            p = st_["partial"]
            p.zero_()
            p[:c] = g.reshape(-1, c).sum(0)
            p[c:2 * c] = (g * y.to(CD)).reshape(-1, c).sum(0)            # sum(g*y); the finalize turns it into sum(g*xhat)
        else:
            if i["dres"] is not None:
                dres = i["dres"]
                dres.copy_(((dres.to(CD) if i["dres_acc"] else 0) + g).to(dres.dtype))
            if i["dy"] is not None:
                s = st_.get("sums")
                if s is not None:
                    o = st_["scale"] * (g - s[0] / i["count"] - xh * s[1] / i["count"])
                else:
                    o = g * st_["scale"] if st_.get("scale") is not None else g
                i["dy"].copy_(o.to(i["dy"].dtype))
    elif kind in ("bn_bwd_finalize", "bn_bwd_finalize_sync"):
        st_, c = i["st"], i["c"]
        p = st_["partial"].view(st_["slabs"], 2, c).double().sum(0)
        mean = st_["mean"].double() if st_.get("mean") is not None else 0.0
        inv = st_["invstd"].double() if st_.get("invstd") is not None else 1.0
        p = torch.stack([p[0], inv * (p[1] - mean * p[0])])
        if i["dgamma"] is not None:                 # LOCAL sums (DDP averages the parameter gradients)
            i["dgamma"].add_(p[1].to(i["dgamma"].dtype))
        if i["dbeta"] is not None:
            i["dbeta"].add_(p[0].to(i["dbeta"].dtype))
        if kind == "bn_bwd_finalize_sync":
            p = p.contiguous()
            i["dist"].all_reduce(p)
        st_["sums"][0] = p[0].to(CD); st_["sums"][1] = p[1].to(CD)
    elif kind == "reduce_partials":
        K, c, slabs = i["K"], i["c"], i["slabs"]
        # partial[(slab*K + k)*cc + ch] with cc = the channel count the producer used
        cc = i["partial"].numel() // (slabs * K)
        p = i["partial"].view(slabs, K, cc).double().sum(0)[:, :c]
        out = i["out"]
        for k in range(K):
            idx = k * i["sk"] + torch.arange(c, device=out.device) * i["sc"]
            v = (p[k] * i["scale"]).to(CD)
            out[idx] = (out[idx].to(CD) + v if i["accumulate"] else v).to(out.dtype)
    elif kind == "maxpool":
        i["y"].copy_(F.max_pool2d(_nchw(i["x"]), 3, 2, 1).permute(0, 2, 3, 1).to(i["y"].dtype))
    elif kind == "maxpool_bwd":
        xr = _nchw(i["x"]).requires_grad_(True)
        F.max_pool2d(xr, 3, 2, 1).backward(_nchw(i["dy"]))
        i["dx"].copy_(xr.grad.permute(0, 2, 3, 1).to(i["dx"].dtype))
    elif kind == "bilinear":
        y = i["y"]
        y.copy_(F.interpolate(_nchw(i["x"]), y.shape[1:3], mode="bilinear", align_corners=i.get("align", True)).permute(0, 2, 3, 1).to(y.dtype))
    elif kind == "bilinear_bwd":
        dy, dx = i["dy"], i["dx"]
        xr = torch.zeros(dx.shape[0], dy.shape[3], dx.shape[1], dx.shape[2], dtype=CD, device=dy.device, requires_grad=True)
        F.interpolate(xr, dy.shape[1:3], mode="bilinear", align_corners=i.get("align", True)).backward(_nchw(dy))
        g = xr.grad.permute(0, 2, 3, 1) * (float(i["gscale"][0]) if i["gscale"] is not None else 1.0)
        c = dy.shape[3]
        dx[..., :c] = ((dx[..., :c].to(CD) if i["accumulate"] else 0) + g).to(dx.dtype)
    elif kind == "upsample_add":                # y = act(a + nearest_up_{2^k}(z))   (segb200_upsample_add)
        a, z, y, k = i["a"], i["z"], i["y"], i["k"]
        up = z.to(CD).repeat_interleave(1 << k, dim=1).repeat_interleave(1 << k, dim=2)
        y.copy_(_act(a.to(CD) + up, i["act"]).to(y.dtype))
    elif kind == "upsample_add_bwd":            # g = dy * act'(y); da (+)= g; dz (+)= block sums of g
        dy, y, da, dz, k = i["dy"], i["y"], i["da"], i["dz"], i["k"]
        g = dy.to(CD)
        if i["act"] == "relu":
            g = g * (y.to(CD) > 0)
        elif i["act"] == "relu6":
            g = g * ((y.to(CD) > 0) & (y.to(CD) < 6))
        n, h, w, c = g.shape
        s_ = 1 << k
        gz = g.reshape(n, h // s_, s_, w // s_, s_, c).sum((2, 4))
        da.copy_(((da.to(CD) if i["acc_a"] else 0) + g).to(da.dtype))
        dz.copy_(((dz.to(CD) if i["acc_z"] else 0) + gz).to(dz.dtype))
    elif kind == "gather_cast":
        gather_cast(i["src"], i["index"], i["dst"])
    elif kind == "transpose":                   # segb200_nhwc_to_cn: y[col][row] = x[row][col]; y's columns >= rows are left alone
        x, y, rows, cols = i["x"], i["y"], i["rows"], i["cols"]
        x2 = x.reshape(-1, x.shape[-1]) if x.dim() > 2 else x
        src = torch.as_strided(x2, (rows, cols), (i["x_ld"], 1), x2.storage_offset())
        y2 = y.reshape(-1, y.shape[-1])
        y2[:cols, :rows] = src.t().to(y.dtype)
    elif kind == "row_softmax":
        n = i["n"]
        s2, p2 = i["s"].reshape(-1, i["s"].shape[-1]), i["p"].reshape(-1, i["p"].shape[-1])
        p2.zero_()
        p2[:n, :n] = torch.softmax(s2[:n, :n].to(CD), 1).to(p2.dtype)
    elif kind == "row_softmax_bwd":             # r = sum_j P D; dS = gamma P (D - r)
        n = i["n"]
        p2 = i["p"].reshape(-1, i["p"].shape[-1])[:n, :n].to(CD)
        d2 = i["d"].reshape(-1, i["d"].shape[-1])[:n, :n].to(CD)
        r = (p2 * d2).sum(1, keepdim=True)
        ds2 = i["ds"].reshape(-1, i["ds"].shape[-1])
        ds2.zero_()
        ds2[:n, :n] = (float(i["gamma"][0]) * p2 * (d2 - r)).to(ds2.dtype)
        i["part"][:n] = r[:, 0].to(i["part"].dtype)
    elif kind == "cam_softmax":                 # A = softmax(rowmax(E) - E)
        c = i["c"]
        e = i["e"].reshape(-1, i["e"].shape[-1])[:c, :c].to(CD)
        a2 = i["att"].reshape(-1, i["att"].shape[-1])
        a2.zero_()
        a2[:c, :c] = torch.softmax(e.max(1, keepdim=True)[0] - e, 1).to(a2.dtype)
    elif kind == "cam_softmax_bwd":             # r = sum A G; dE = -gamma A (G - r)
        c = i["c"]
        a2 = i["att"].reshape(-1, i["att"].shape[-1])[:c, :c].to(CD)
        g2 = i["g"].reshape(-1, i["g"].shape[-1])[:c, :c].to(CD)
        r = (a2 * g2).sum(1, keepdim=True)
        i["de"][:c, :c] = (-float(i["gamma"][0]) * a2 * (g2 - r)).to(i["de"].dtype)
        i["part"][:c] = r[:, 0].to(i["part"].dtype)
    elif kind == "cam_bwd_pack":                # w1 = gamma A^T, w2 = dE + dE^T (zero K padding)
        c = i["c"]
        a2 = i["att"].reshape(-1, i["att"].shape[-1])[:c, :c].to(CD)
        de = i["de"][:c, :c].to(CD)
        w1, w2 = i["w1"].reshape(-1, i["w1"].shape[-1]), i["w2"].reshape(-1, i["w2"].shape[-1])
        w1.zero_(); w2.zero_()
        w1[:c, :c] = (float(i["gamma"][0]) * a2.t()).to(w1.dtype)
        w2[:c, :c] = (de + de.t()).to(w2.dtype)
    elif kind == "gap":
        i["y"].copy_(i["x"].to(CD).mean((1, 2), keepdim=True).to(i["y"].dtype))
    elif kind == "nc_broadcast":
        y = i["y"]
        v = i["v"].to(CD) * i["scale"]
        y.copy_(((y.to(CD) if i["accumulate"] else 0) + v.expand_as(y)).to(y.dtype))
    elif kind == "stride2_place":
        t, z = i["t"], i["z"]
        if i["mode"] == 0:
            z.zero_()
            z[:, ::2, ::2, :] = t
        else:
            z[:, ::2, ::2, :] = (z[:, ::2, ::2, :].to(CD) + t.to(CD)).to(z.dtype)
    elif kind == "dw_wgrad":
        x, dy, c, d = i["x"], i["dy"], i["c"], i["dilation"]
        W0 = torch.zeros(c, 1, 3, 3, dtype=CD, device=x.device, requires_grad=True)
        xin = F.relu(_nchw(x)) if i.get("pre_relu") else _nchw(x)
        F.conv2d(xin, W0, None, 1, d, d, groups=c).backward(_nchw(dy))
        p = i["partial"]
        p.zero_()
        p[:9 * c] = W0.grad.reshape(c, 9).t().reshape(-1)
    elif kind == "upsample_ce":
        lg, tgt, dfull, out3, nc = i["logits"], i["target"], i["dfull"], i["out3"], i["nclass"]
        up = F.interpolate(_nchw(lg[..., :nc]), tgt.shape[1:3], mode="bilinear", align_corners=i.get("align", True))
        valid = (tgt != i["ignore_index"]) & (tgt >= 0) & (tgt < nc)
        lsm = F.log_softmax(up, 1)
        oh = F.one_hot(tgt.clamp(0, nc - 1), nc).permute(0, 3, 1, 2).to(CD)
        cnt = float(valid.sum())
        out3[0] = float(-(lsm * oh).sum(1)[valid].sum() / max(cnt, 1.0))
        out3[1] = 1.0 / cnt if cnt > 0 else 0.0
        out3[2] = cnt
        g = (lsm.exp() - oh) * valid[:, None].to(CD)
        dfull.zero_()
        dfull[..., :nc] = g.permute(0, 2, 3, 1).to(dfull.dtype)
    elif kind in ("cca_weight_softmax", "cca_map", "cca_weight_bwd", "cca_gather", "cca_scatter"):
        from oracle import segref as R          # criss-cross index map: the oracle's ca_weight / ca_map (pinned to ca_cuda.cu)

        def att_nchw(a, L):
            return a[..., :L].permute(0, 3, 1, 2).to(CD)          # [n,h,w,L] -> [n,L,h,w]
        if kind == "cca_weight_softmax":
            q, k, att = i["q"], i["k"], i["att"]
            L = q.shape[1] + q.shape[2] - 1
            att.zero_()
            att[..., :L] = torch.softmax(R.ca_weight(_nchw(q), _nchw(k)), 1).permute(0, 2, 3, 1).to(att.dtype)
        elif kind == "cca_map":
            att, v, x, y = i["att"], i["v"], i["x"], i["y"]
            L = v.shape[1] + v.shape[2] - 1
            o = i["gamma"].to(CD)[0] * R.ca_map(att_nchw(att, L), _nchw(v)) + _nchw(x)
            y.copy_(o.permute(0, 2, 3, 1).to(y.dtype))
        elif kind == "cca_weight_bwd":
            dy, v, att, de, part = i["dy"], i["v"], i["att"], i["de"], i["part"]
            L = v.shape[1] + v.shape[2] - 1
            D = R.ca_weight(_nchw(dy), _nchw(v))                    # D[p][z] = dy[p] . v[key(p,z)]
            A = att_nchw(att, L)
            ssum = (A * D).sum(1, keepdim=True)
            de.zero_()
            de[..., :L] = (i["gamma"].to(CD)[0] * A * (D - ssum)).permute(0, 2, 3, 1).to(de.dtype)
            part.zero_()
            part[0] = float(ssum.sum())
        elif kind == "cca_gather":
            a, src, out = i["a"], i["src"], i["out"]
            L = src.shape[1] + src.shape[2] - 1
            o = i["scale"] * R.ca_map(att_nchw(a, L), _nchw(src)).permute(0, 2, 3, 1)
            out.copy_(((out.to(CD) if i["accumulate"] else 0) + o).to(out.dtype))
        else:
            a, src, out = i["a"], i["src"], i["out"]
            L = src.shape[1] + src.shape[2] - 1
            g0 = torch.zeros_like(_nchw(src), requires_grad=True)
            (R.ca_map(att_nchw(a, L), g0) * _nchw(src)).sum().backward()
            sc = i["scale"] * (float(i["scale_dev"][0]) if i["scale_dev"] is not None else 1.0)
            out.copy_(((out.to(CD) if i["accumulate"] else 0) + sc * g0.grad.permute(0, 2, 3, 1)).to(out.dtype))
    elif kind == "scatter_add":
        src, idx, dst = i["src"], i["index"].long(), i["dst"]
        m = idx >= 0
        dst.index_add_(0, idx[m], src[m].to(dst.dtype))
    else:
        raise NotImplementedError(kind)


def gather_cast(src, index, dst):
    idx = index.long()
    v = torch.where(idx >= 0, src[idx.clamp(min=0)], torch.zeros((), dtype=src.dtype, device=src.device))
    dst.copy_(v.to(dst.dtype))


def forward_backward(trainer, x, target, dropout_masks=None):
    """CPU interpretation of DeepLabV3PlusTrainerB200.forward_backward (single rank)."""
    st = trainer.plan_for(x.shape)
    pl, S = st["plan"], trainer.store
    pl.x_in.copy_(x)
    pl.target.copy_(target)
    for name, m in pl.masks.items():
        if dropout_masks is not None and name in dropout_masks:
            m.copy_(dropout_masks[name].reshape(m.shape))
        else:
            m.fill_(1.0)
    S.grad.zero_()
    gather_cast(S.master, S.idx16, S.w16)
    if S.idx32 is not None:
        gather_cast(S.master, S.idx32, S.w32)
    for s in pl.fwd + pl.bwd:
        run_step(s)
    return pl.out3[0].clone()


def sgd(trainer, lr=None):
    S = trainer.store
    lr = trainer.lr if lr is None else lr
    for lo, hi, l in ((0, S.n_encoder, lr), (S.n_encoder, S.total, lr * trainer.decoder_lr_factor)):
        g = S.grad[lo:hi] / trainer.world + trainer.weight_decay * S.master[lo:hi]
        S.mom[lo:hi] = trainer.momentum * S.mom[lo:hi] + g
        S.master[lo:hi] -= l * S.mom[lo:hi]

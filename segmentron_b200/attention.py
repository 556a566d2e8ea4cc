"""Attention blocks of DANet / CCNet on the C-ABI kernels (NHWC tensors in, NHWC tensors out).

CAM_Module (modules/module.py:142-162):  E = X^T X -> softmax(rowmax - E) -> gamma * (A X) + x
  = transpose kernel + tcgen05 GEMM (fp32 energies) + cam_softmax + tcgen05 GEMM (scale = gamma, residual = x).
CrissCrossAttention (modules/cc_attention.py:62-72): q/k/v 1x1 convs (GEMMs, bias as shift) ->
  cca_weight_softmax (ca_forward + softmax fused) -> cca_map (ca_map_forward + gamma*out + x fused).
"""
import torch

from . import fold, lib as L, ops


def cam_nhwc(x, gamma, out=None):
    """x [n,h,w,c] (16-bit, c % 8 == 0), gamma: 1-element fp32 CUDA tensor -> gamma * CAM(x) + x."""
    n, h, w, c, x_ld = ops._nhwc(x, "x")
    dt = x.dtype
    hw = h * w
    pitch = fold.round_up(hw, 64)
    lib = L.load()
    st = ops._stream
    if out is None:
        out = torch.empty(n, h, w, c, dtype=dt, device=x.device)
    gvec = gamma.detach().float().reshape(1).expand(c).contiguous()
    for b in range(n):
        xb = x[b:b + 1]
        xt = torch.zeros(c, pitch, dtype=dt, device=x.device)                   # [C][N] K-major copy (zero tail)
        L.check(lib.segb200_nhwc_to_cn(ops._ptr(xb), ops._ptr(xt), 1, c, hw, x_ld, pitch, ops.dt_code(dt), st()), "nhwc_to_cn")
        energy = torch.empty(1, 1, c, c, dtype=torch.float32, device=x.device)
        ops.conv_gemm(xt.view(1, 1, c, pitch), xt, energy, cin=pitch, cout=c)   # E[c1][c2] = sum_p x[p,c1] x[p,c2]
        cpad = fold.round_up(c, fold.conv_kblock(c))                               # K pitch the second GEMM expects
        att = torch.zeros(c, cpad, dtype=dt, device=x.device)
        L.check(lib.segb200_cam_softmax(ops._ptr(energy), ops._ptr(att), c, c, c, cpad, ops.dt_code(dt), st()), "cam_softmax")
        ops.conv_gemm(xb, att, out[b:b + 1], cin=c, cout=c, scale=gvec, residual=xb)   # out[p,c1] = g*sum_c2 att[c1,c2] x[p,c2] + x
    return out


class PamFunction(torch.autograd.Function):
    """y = gamma * PAM(x) + x on an NHWC 16-bit tensor, differentiable (training-mode PAM_Module, modules/module.py:100-131).

    Training materialises the attention matrix per image, like the reference's bmm / softmax / bmm (the inference kernel,
    csrc/pam.cu, never forms it): S = Q K^T (tcgen05 GEMM, fp32) -> row_softmax -> P (16-bit, [N][N]) -> y = gamma P V + x (GEMM
    with V^T as the weight operand).  Backward: D = dy V^T (GEMM, fp32) -> row_softmax_bwd: dS = gamma P (D - sum P D) and the gamma
    gradient -> dV = gamma P^T dy, dQ = dS K, dK = dS^T Q (GEMMs over transposed copies), then the weight / bias / data gradients of
    the three 1x1 convs.  Needs h*w % 8 == 0 and channels % 64 == 0 (GEMM operand pitches); 2 * N^2 bytes per image for P."""

    @staticmethod
    def forward(ctx, x, wq, bq, wk, bk, wv, bv, gamma):
        from . import train_ops as T
        n, h, w, c, x_ld = ops._nhwc(x, "x")
        ntok = h * w
        if ntok % 8 or c % 64 or x_ld != c:
            raise RuntimeError("segb200: training-mode PAM needs a contiguous NHWC input with h*w % 8 == 0 and channels % 64 == 0")
        dt = x.dtype
        dtc = ops.dt_code(dt)
        lib = L.load()
        st = ops._stream
        cq = wq.shape[0]
        cqp = fold.round_up(cq, fold.conv_kblock(cq))              # q / k also serve as GEMM weight operands: K pitch = padded depth
        pitch = fold.round_up(ntok, 64)
        pk = [fold.pack_conv_weight(t.detach(), dt) for t in (wq, wk, wv)]
        q = torch.zeros(n, h, w, cqp, dtype=dt, device=x.device)
        k = torch.zeros(n, h, w, cqp, dtype=dt, device=x.device)
        v = torch.empty(n, h, w, c, dtype=dt, device=x.device)
        ops.conv_gemm(x, pk[0], q[..., :cq], cin=c, cout=cq, shift=bq.detach().float().contiguous())
        ops.conv_gemm(x, pk[1], k[..., :cq], cin=c, cout=cq, shift=bk.detach().float().contiguous())
        ops.conv_gemm(x, pk[2], v, cin=c, cout=c, shift=bv.detach().float().contiguous())
        g1 = gamma.detach().float().reshape(1).contiguous()
        gvec = g1.expand(c).contiguous()
        y = torch.empty(n, h, w, c, dtype=dt, device=x.device)
        probs = torch.empty(n, ntok, pitch, dtype=dt, device=x.device)
        energy = torch.empty(1, 1, ntok, ntok, dtype=torch.float32, device=x.device)
        vt = torch.zeros(c, pitch, dtype=dt, device=x.device)
        for b in range(n):
            ops.conv_gemm(q[b].view(1, 1, ntok, cqp), k[b].view(ntok, cqp), energy, cin=cqp, cout=ntok)        # S = Q K^T
            L.check(lib.segb200_row_softmax(ops._ptr(energy), ops._ptr(probs[b]), ntok, ntok, ntok, pitch, dtc, st()), "row_softmax")
            L.check(lib.segb200_nhwc_to_cn(ops._ptr(v[b]), ops._ptr(vt), 1, c, ntok, c, pitch, dtc, st()), "nhwc_to_cn")
            ops.conv_gemm(probs[b].view(1, 1, ntok, pitch), vt, y[b].view(1, 1, ntok, c), cin=pitch, cout=c, scale=gvec,
                          residual=x[b].view(1, 1, ntok, c))                                                    # gamma P V + x
        ctx.save_for_backward(x, q, k, v, probs, g1, wq, wk, wv)
        ctx.T = T
        return y

    @staticmethod
    def backward(ctx, dy):
        T = ctx.T
        x, q, k, v, probs, g1, wq, wk, wv = ctx.saved_tensors
        n, h, w, c, _ = ops._nhwc(x, "x")
        ntok, pitch, cqp, cq = h * w, probs.shape[2], q.shape[3], wq.shape[0]
        dt = x.dtype
        dtc = ops.dt_code(dt)
        lib = L.load()
        st = ops._stream
        dy = dy.contiguous()
        gvec = g1.expand(c).contiguous()
        dq = torch.zeros(n, h, w, cqp, dtype=dt, device=x.device)
        dk = torch.zeros(n, h, w, cqp, dtype=dt, device=x.device)
        dv = torch.empty(n, h, w, c, dtype=dt, device=x.device)
        dmat = torch.empty(1, 1, ntok, ntok, dtype=torch.float32, device=x.device)
        ds = torch.empty(ntok, pitch, dtype=dt, device=x.device)
        tr = torch.zeros(ntok, pitch, dtype=dt, device=x.device)                   # P^T, then dS^T
        dyt = torch.zeros(c, pitch, dtype=dt, device=x.device)
        kt = torch.zeros(cq, pitch, dtype=dt, device=x.device)
        qt = torch.zeros(cq, pitch, dtype=dt, device=x.device)
        part = torch.empty(n, ntok, dtype=torch.float32, device=x.device)
        for b in range(n):
            dyb = dy[b].view(1, 1, ntok, c)
            ops.conv_gemm(dyb, v[b].view(ntok, c), dmat, cin=c, cout=ntok)                                      # D = dy V^T
            L.check(lib.segb200_row_softmax_bwd(ops._ptr(probs[b]), ops._ptr(dmat), ops._ptr(g1), ops._ptr(ds), ops._ptr(part[b]), ntok,
                                                ntok, pitch, ntok, pitch, dtc, st()), "row_softmax_bwd")
            L.check(lib.segb200_nhwc_to_cn(ops._ptr(probs[b]), ops._ptr(tr), 1, ntok, ntok, pitch, pitch, dtc, st()), "nhwc_to_cn")
            L.check(lib.segb200_nhwc_to_cn(ops._ptr(dy[b]), ops._ptr(dyt), 1, c, ntok, c, pitch, dtc, st()), "nhwc_to_cn")
            ops.conv_gemm(tr.view(1, 1, ntok, pitch), dyt, dv[b].view(1, 1, ntok, c), cin=pitch, cout=c, scale=gvec)   # gamma P^T dy
            L.check(lib.segb200_nhwc_to_cn(ops._ptr(k[b]), ops._ptr(kt), 1, cq, ntok, cqp, pitch, dtc, st()), "nhwc_to_cn")
            ops.conv_gemm(ds.view(1, 1, ntok, pitch), kt, dq[b].view(1, 1, ntok, cqp)[..., :cq], cin=pitch, cout=cq)    # dS K
            L.check(lib.segb200_nhwc_to_cn(ops._ptr(ds), ops._ptr(tr), 1, ntok, ntok, pitch, pitch, dtc, st()), "nhwc_to_cn")
            L.check(lib.segb200_nhwc_to_cn(ops._ptr(q[b]), ops._ptr(qt), 1, cq, ntok, cqp, pitch, dtc, st()), "nhwc_to_cn")
            ops.conv_gemm(tr.view(1, 1, ntok, pitch), qt, dk[b].view(1, 1, ntok, cqp)[..., :cq], cin=pitch, cout=cq)    # dS^T Q
        dgamma = torch.zeros(1, dtype=torch.float32, device=x.device)
        L.check(lib.segb200_reduce_partials(ops._ptr(part), n * ntok, 1, 1, ops._ptr(dgamma), 0, 1, 0, 1.0, st()), "reduce_partials")
        grads_w, grads_b = [], []
        dx = torch.empty(n, h, w, c, dtype=dt, device=x.device)
        res = dy
        for wt, d, co in ((wq, dq[..., :cq], cq), (wk, dk[..., :cq], cq), (wv, dv, c)):
            dw = torch.zeros(co, 1, c, dtype=torch.float32, device=x.device)
            T.conv_wgrad(x, d, dw, cin=c, cout=co)
            grads_w.append(dw.view(co, c, 1, 1))
            grads_b.append(_colsum(d))
            ops.conv_gemm(d, T.pack_dgrad_weight(wt.detach().to(dt), dt), dx, cin=co, cout=c, residual=res)
            res = dx
        return dx, grads_w[0], grads_b[0], grads_w[1], grads_b[1], grads_w[2], grads_b[2], dgamma.to(g1.dtype)


class CamFunction(torch.autograd.Function):
    """y = gamma * CAM(x) + x on NHWC 16-bit x, differentiable (training-mode CAM_Module, modules/module.py:134-162).

    Forward = cam_nhwc keeping, per image, the K-major copy of x and the attention matrix.  Backward (csrc/cam_bwd.cu):
      G = dy^T x (tcgen05 GEMM, fp32)  ->  cam_softmax_bwd: dE = -gamma A (G - sum A G), dgamma  ->  cam_bwd_pack: W1 = gamma A^T,
      W2 = dE + dE^T  ->  dx = dy + dy.W1 + x.W2 (two tcgen05 GEMMs chained through the residual operand)."""

    @staticmethod
    def forward(ctx, x, gamma):
        n, h, w, c, x_ld = ops._nhwc(x, "x")
        if c % 8:
            raise RuntimeError("segb200: CAM needs a channel count that is a multiple of 8")
        dt = x.dtype
        hw = h * w
        pitch = fold.round_up(hw, 64)
        cpad = fold.round_up(c, fold.conv_kblock(c))
        lib = L.load()
        st = ops._stream
        g1 = gamma.detach().float().reshape(1).contiguous()
        gvec = g1.expand(c).contiguous()
        out = torch.empty(n, h, w, c, dtype=dt, device=x.device)
        xts = torch.zeros(n, c, pitch, dtype=dt, device=x.device)                 # [C][N] K-major copies (zero tails)
        atts = torch.zeros(n, c, cpad, dtype=dt, device=x.device)
        energy = torch.empty(1, 1, c, c, dtype=torch.float32, device=x.device)
        for b in range(n):
            xb = x[b:b + 1]
            L.check(lib.segb200_nhwc_to_cn(ops._ptr(xb), ops._ptr(xts[b]), 1, c, hw, x_ld, pitch, ops.dt_code(dt), st()), "nhwc_to_cn")
            ops.conv_gemm(xts[b].view(1, 1, c, pitch), xts[b], energy, cin=pitch, cout=c)
            L.check(lib.segb200_cam_softmax(ops._ptr(energy), ops._ptr(atts[b]), c, c, c, cpad, ops.dt_code(dt), st()), "cam_softmax")
            ops.conv_gemm(xb, atts[b], out[b:b + 1], cin=c, cout=c, scale=gvec, residual=xb)
        ctx.save_for_backward(x, xts, atts, g1)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, xts, atts, g1 = ctx.saved_tensors
        n, h, w, c, x_ld = ops._nhwc(x, "x")
        dt = x.dtype
        hw = h * w
        pitch, cpad = xts.shape[2], atts.shape[2]
        lib = L.load()
        st = ops._stream
        dy = dy.contiguous()
        dx = torch.empty(n, h, w, c, dtype=dt, device=x.device)
        tmp = torch.empty(1, h, w, c, dtype=dt, device=x.device)
        dyt = torch.zeros(c, pitch, dtype=dt, device=x.device)
        gmat = torch.empty(1, 1, c, c, dtype=torch.float32, device=x.device)
        de = torch.empty(c, c, dtype=torch.float32, device=x.device)
        part = torch.empty(n, c, dtype=torch.float32, device=x.device)
        w1 = torch.empty(c, cpad, dtype=dt, device=x.device)
        w2 = torch.empty(c, cpad, dtype=dt, device=x.device)
        for b in range(n):
            xb, dyb = x[b:b + 1], dy[b:b + 1]
            L.check(lib.segb200_nhwc_to_cn(ops._ptr(dyb), ops._ptr(dyt), 1, c, hw, c, pitch, ops.dt_code(dt), st()), "nhwc_to_cn")
            ops.conv_gemm(dyt.view(1, 1, c, pitch), xts[b], gmat, cin=pitch, cout=c)          # G[c1][c2] = sum_p dy[p,c1] x[p,c2]
            L.check(lib.segb200_cam_softmax_bwd(ops._ptr(atts[b]), ops._ptr(gmat), ops._ptr(g1), ops._ptr(de), ops._ptr(part[b]), c, cpad,
                                                c, c, ops.dt_code(dt), st()), "cam_softmax_bwd")
            L.check(lib.segb200_cam_bwd_pack(ops._ptr(atts[b]), ops._ptr(de), ops._ptr(g1), ops._ptr(w1), ops._ptr(w2), c, cpad, c, cpad,
                                             ops.dt_code(dt), st()), "cam_bwd_pack")
            ops.conv_gemm(dyb, w1, tmp, cin=c, cout=c, residual=dyb)                          # dy + dy.(gamma A)
            ops.conv_gemm(xb, w2, dx[b:b + 1], cin=c, cout=c, residual=tmp)                   # ... + x.(dE + dE^T)
        dgamma = torch.zeros(1, dtype=torch.float32, device=x.device)
        L.check(lib.segb200_reduce_partials(ops._ptr(part), n * c, 1, 1, ops._ptr(dgamma), 0, 1, 0, 1.0, st()), "reduce_partials")
        return dx, dgamma


def cca_nhwc(x, wq, bq, wk, bk, wv, bv, gamma, out=None):
    """One criss-cross attention step.  wq/wk/wv: packed 1x1 weights (fold.pack_conv_weight), b*: fp32 bias vectors."""
    n, h, w, c, x_ld = ops._nhwc(x, "x")
    dt = x.dtype
    cq = wq.shape[0]
    lib = L.load()
    q = torch.empty(n, h, w, cq, dtype=dt, device=x.device)
    k = torch.empty(n, h, w, cq, dtype=dt, device=x.device)
    v = torch.empty(n, h, w, c, dtype=dt, device=x.device)
    ops.conv_gemm(x, wq, q, cin=c, cout=cq, shift=bq)
    ops.conv_gemm(x, wk, k, cin=c, cout=cq, shift=bk)
    ops.conv_gemm(x, wv, v, cin=c, cout=c, shift=bv)
    att_ld = fold.round_up(h + w - 1, 4)
    att = torch.empty(n, h, w, att_ld, dtype=torch.float32, device=x.device)
    L.check(lib.segb200_cca_weight_softmax(ops._ptr(q), ops._ptr(k), ops._ptr(att), n, h, w, cq, cq, cq, att_ld,
                                           ops.dt_code(dt), ops._stream()), "cca_weight_softmax")
    y = out if out is not None else torch.empty(n, h, w, c, dtype=dt, device=x.device)
    g = gamma.detach().float().reshape(1).contiguous()
    L.check(lib.segb200_cca_map(ops._ptr(att), ops._ptr(v), ops._ptr(x), ops._ptr(y), ops._ptr(g), n, h, w, c, att_ld, c, x_ld,
                                c, ops.dt_code(dt), ops._stream()), "cca_map")
    return y


def pam_nhwc(x, wq, bq, wk, bk, wv, bv, gamma, out=None):
    """PAM_Module on NHWC x [n,h,w,c] (c % 64 == 0, contiguous).  wq/wk: packed [c/8 = 64][1][c], wv: packed [c][1][c]
    (fold.pack_conv_weight); bq/bk/bv fp32 bias vectors; gamma 1-element tensor."""
    n, h, w, c, x_ld = ops._nhwc(x, "x")
    if x_ld != c or c % 64 != 0:
        raise RuntimeError("segb200 pam: x must be a contiguous NHWC tensor with channels % 64 == 0")
    dt = x.dtype
    ntok = h * w
    dq = wq.shape[0]
    dv = wv.shape[0]
    if dq != 64 or dv % 256 != 0:
        raise RuntimeError(f"segb200 pam: kernel is specialised for query depth 64 and d_v % 256 == 0 (got {dq}, {dv})")
    lib = L.load()
    q = torch.empty(n, h, w, dq, dtype=dt, device=x.device)
    k = torch.empty(n, h, w, dq, dtype=dt, device=x.device)
    ops.conv_gemm(x, wq, q, cin=c, cout=dq, shift=bq)
    ops.conv_gemm(x, wk, k, cin=c, cout=dq, shift=bk)
    # V^T[b] = W_v . X_b^T : the same GEMM with the roles swapped (A = weight rows, B = the image's pixels)
    pitch = fold.round_up(ntok, 8)
    vt = torch.empty(n, dv, pitch, dtype=dt, device=x.device)
    wv_as_x = wv.view(1, 1, dv, wv.shape[-1])
    for b in range(n):
        ops.conv_gemm(wv_as_x, x[b], vt[b].view(1, 1, dv, pitch)[..., :ntok], cin=c, cout=ntok)
    y = out if out is not None else torch.empty(n, h, w, c, dtype=dt, device=x.device)
    sm = torch.empty(n * ntok, dtype=torch.float32, device=x.device)
    sl = torch.empty(n * ntok, dtype=torch.float32, device=x.device)
    g = gamma.detach().float().reshape(1).contiguous()
    L.check(lib.segb200_pam_attention(ops._ptr(q), ops._ptr(k), ops._ptr(vt), ops._ptr(bv), ops._ptr(g), ops._ptr(x), ops._ptr(y),
                                      ops._ptr(sm), ops._ptr(sl), n, ntok, dv, dq, dq, pitch, x_ld, c, ops.dt_code(dt),
                                      ops._stream()), "pam_attention")
    return y


def nonlocal_nhwc(x, wqk, scale_qk, shift_qk, wv, bv, key_ch, out):
    """OCNet BaseAttentionBlock (models/ocnet.py:95-113, scale 1) up to, not including, its output conv W, on a contiguous NHWC
    x [n,h,w,c]:  out[n,h,w,dv] = softmax(Q K^T * key_ch^-0.5) V + b_v.
      wqk / scale_qk / shift_qk : the shared f_key = f_query 1x1 conv (+ bias + folded BN + ReLU) packed TWICE into one
                                  2*key_ch-channel GEMM -- channels [0, key_ch) pre-multiplied by key_ch^-0.5 (the query copy; exact
                                  for key_ch = 256 = 2^8), channels [key_ch, 2 key_ch) plain (the keys)
      wv, bv                    : f_value weights packed [dv][1][c] and bias
    Q K^T, the softmax and P V run in segb200_nonlocal_attention (csrc/pam.cu, query/key depth 256): the N x N similarity map of the
    reference's two torch.bmm is never materialised."""
    n, h, w, c, x_ld = ops._nhwc(x, "x")
    if x_ld != c or c % 64 != 0:
        raise RuntimeError("segb200 nonlocal: x must be a contiguous NHWC tensor with channels % 64 == 0")
    dt = x.dtype
    ntok = h * w
    dv = wv.shape[0]
    lib = L.load()
    qk = torch.empty(n, h, w, 2 * key_ch, dtype=dt, device=x.device)
    ops.conv_gemm(x, wqk, qk, cin=c, cout=2 * key_ch, scale=scale_qk, shift=shift_qk, act="relu")
    pitch = fold.round_up(ntok, 8)
    vt = torch.empty(n, dv, pitch, dtype=dt, device=x.device)
    wv_as_x = wv.view(1, 1, dv, wv.shape[-1])
    for b in range(n):                                   # V^T[b] = W_v . X_b^T (the conv GEMM with the operand roles swapped)
        ops.conv_gemm(wv_as_x, x[b], vt[b].view(1, 1, dv, pitch)[..., :ntok], cin=c, cout=ntok)
    sm = torch.empty(n * ntok, dtype=torch.float32, device=x.device)
    sl = torch.empty(n * ntok, dtype=torch.float32, device=x.device)
    q, k = qk[..., :key_ch], qk[..., key_ch:]
    L.check(lib.segb200_nonlocal_attention(ops._ptr(q), ops._ptr(k), ops._ptr(vt), ops._ptr(bv), None, None, ops._ptr(out),
                                           ops._ptr(sm), ops._ptr(sl), n, ntok, key_ch, dv, 2 * key_ch, 2 * key_ch, pitch, 0,
                                           out.stride(2), ops.dt_code(dt), ops._stream()), "nonlocal_attention")
    return out


# ------------------------------------------------------------------------------------------------------------
# criss-cross attention with a backward pass (training): torch.autograd.Function over the C-ABI kernels
# ------------------------------------------------------------------------------------------------------------
def _colsum(t):
    """fp32 per-channel sum of an NHWC 16-bit tensor through the fixed-order reduction kernels (conv bias gradients)."""
    n, h, w, c, ld = ops._nhwc(t, "t")
    lib = L.load()
    rows = n * h * w
    slabs = lib.segb200_reduce_slabs(rows, c, 0)
    partial = torch.empty(slabs * 2 * c, dtype=torch.float32, device=t.device)
    out = torch.zeros(2, c, dtype=torch.float32, device=t.device)
    s = ops._stream()
    L.check(lib.segb200_bn_stats(ops._ptr(t), rows, c, ld, ops.dt_code(t.dtype), ops._ptr(partial), 0, s), "bn_stats")
    L.check(lib.segb200_reduce_partials(ops._ptr(partial), slabs, 2, c, ops._ptr(out), c, 1, 0, 1.0, s), "reduce_partials")
    return out[0]


class CrissCrossFunction(torch.autograd.Function):
    """y = gamma * CCA(q(x), k(x), v(x)) + x  on an NHWC 16-bit tensor; parameters are the reference module's
    (query_conv / key_conv / value_conv weight [Co,C,1,1] + bias, gamma [1]; modules/cc_attention.py:52-72).
    Backward = segb200_cca_weight_bwd / cca_gather / cca_scatter (the _C.ca_backward / ca_map_backward pair, softmax backward
    and gamma gradient fused) + the tcgen05 weight-/data-gradient kernels of the three 1x1 convs."""

    @staticmethod
    def forward(ctx, x, wq, bq, wk, bk, wv, bv, gamma):
        from . import train_ops as T
        n, h, w, c, x_ld = ops._nhwc(x, "x")
        dt = x.dtype
        lib = L.load()
        cq = wq.shape[0]
        pk = [fold.pack_conv_weight(t.detach(), dt) for t in (wq, wk, wv)]
        q = torch.empty(n, h, w, cq, dtype=dt, device=x.device)
        k = torch.empty(n, h, w, cq, dtype=dt, device=x.device)
        v = torch.empty(n, h, w, c, dtype=dt, device=x.device)
        ops.conv_gemm(x, pk[0], q, cin=c, cout=cq, shift=bq.detach().float().contiguous())
        ops.conv_gemm(x, pk[1], k, cin=c, cout=cq, shift=bk.detach().float().contiguous())
        ops.conv_gemm(x, pk[2], v, cin=c, cout=c, shift=bv.detach().float().contiguous())
        att_ld = fold.round_up(h + w - 1, 4)
        att = torch.empty(n, h, w, att_ld, dtype=torch.float32, device=x.device)
        L.check(lib.segb200_cca_weight_softmax(ops._ptr(q), ops._ptr(k), ops._ptr(att), n, h, w, cq, cq, cq, att_ld,
                                               ops.dt_code(dt), ops._stream()), "cca_weight_softmax")
        y = torch.empty(n, h, w, c, dtype=dt, device=x.device)
        g = gamma.detach().float().reshape(1).contiguous()
        L.check(lib.segb200_cca_map(ops._ptr(att), ops._ptr(v), ops._ptr(x), ops._ptr(y), ops._ptr(g), n, h, w, c, att_ld, c, x_ld,
                                    c, ops.dt_code(dt), ops._stream()), "cca_map")
        ctx.save_for_backward(x, q, k, v, att, g, wq, wk, wv)
        ctx.T = T
        return y

    @staticmethod
    def backward(ctx, dy):
        T = ctx.T
        x, q, k, v, att, g, wq, wk, wv = ctx.saved_tensors
        lib = L.load()
        s = ops._stream()
        dy = dy.contiguous()
        n, h, w, c, _ = ops._nhwc(x, "x")
        cq = q.shape[3]
        dt = x.dtype
        dtc = ops.dt_code(dt)
        att_ld = att.shape[3]
        de = torch.empty_like(att)
        nb = lib.segb200_cca_weight_bwd_blocks(n, h, w)
        part = torch.empty(nb, dtype=torch.float32, device=x.device)
        L.check(lib.segb200_cca_weight_bwd(ops._ptr(dy), ops._ptr(v), ops._ptr(att), ops._ptr(de), ops._ptr(part), ops._ptr(g), n, h,
                                           w, c, dy.stride(2), c, att_ld, dtc, s), "cca_weight_bwd")
        dgamma = torch.zeros(1, dtype=torch.float32, device=x.device)
        L.check(lib.segb200_reduce_partials(ops._ptr(part), nb, 1, 1, ops._ptr(dgamma), 0, 1, 0, 1.0, s), "reduce_partials")
        dq = torch.empty_like(q)
        dk = torch.empty_like(k)
        dv = torch.empty_like(v)
        L.check(lib.segb200_cca_gather(ops._ptr(de), ops._ptr(k), ops._ptr(dq), n, h, w, cq, att_ld, cq, cq, 1.0, 0, dtc, s), "cca_gather")
        L.check(lib.segb200_cca_scatter(ops._ptr(de), ops._ptr(q), ops._ptr(dk), n, h, w, cq, att_ld, cq, cq, 1.0, None, 0, dtc, s),
                "cca_scatter")
        L.check(lib.segb200_cca_scatter(ops._ptr(att), ops._ptr(dy), ops._ptr(dv), n, h, w, c, att_ld, dy.stride(2), c, 1.0, ops._ptr(g),
                                        0, dtc, s), "cca_scatter")
        grads_w, grads_b = [], []
        dx = torch.empty(n, h, w, c, dtype=dt, device=x.device)
        res = dy
        for wt, d, co in ((wq, dq, cq), (wk, dk, cq), (wv, dv, c)):
            dw = torch.zeros(co, 1, c, dtype=torch.float32, device=x.device)
            T.conv_wgrad(x, d, dw, cin=c, cout=co)
            grads_w.append(dw.view(co, c, 1, 1))
            grads_b.append(_colsum(d))
            ops.conv_gemm(d, T.pack_dgrad_weight(wt.detach().to(dt), dt), dx, cin=co, cout=c, residual=res)
            res = dx
        return dx, grads_w[0], grads_b[0], grads_w[1], grads_b[1], grads_w[2], grads_b[2], dgamma.to(g.dtype)

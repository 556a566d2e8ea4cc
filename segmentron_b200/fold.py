"""Host-side parameter preparation: BatchNorm folding and weight packing for the C-ABI kernels.

BatchNorm (eval): y = (x - mean) / sqrt(var + eps) * weight + bias  (reference: modules/batch_norm.py:126,
torch F.batch_norm) == x * scale + shift with scale = weight / sqrt(var + eps), shift = bias - mean * scale.
``eps`` is read at fold time, never at construction, because the reference mutates it after
construction (tools/eval.py:50-53, solver/optimizer.py:18-20).
"""
import torch


def bn_fold(weight, bias, mean, var, eps):
    scale = weight.float() / torch.sqrt(var.float() + eps)
    shift = bias.float() - mean.float() * scale
    return scale.contiguous(), shift.contiguous()


def round_up(v, m):
    return (v + m - 1) // m * m


def conv_kblock(cin):
    return 64 if cin >= 64 else (32 if cin >= 32 else 16)


def pack_conv_weight(w, dtype, cout_pad=None):
    """[Cout, Cin, kh, kw] (OIHW, torch) -> [Cout_pad, kh*kw, Cin_pad] K-major, ``dtype``."""
    co, ci, kh, kw = w.shape
    bk = conv_kblock(ci)
    cin_pad = round_up(ci, bk)
    cop = cout_pad or round_up(co, 8)
    out = torch.zeros(cop, kh * kw, cin_pad, dtype=dtype, device=w.device)
    out[:co, :, :ci] = w.permute(0, 2, 3, 1).reshape(co, kh * kw, ci).to(dtype)
    return out.contiguous()


def pad_vec(v, n, fill=0.0):
    if v.numel() == n:
        return v.float().contiguous()
    out = torch.full((n,), fill, dtype=torch.float32, device=v.device)
    out[:v.numel()] = v.float()
    return out


def pack_dw_weight(w, scale):
    """[C, 1, 3, 3] -> fp32 [9, C] with the following BN's scale folded in."""
    c = w.shape[0]
    return (w.float().reshape(c, 9) * scale.float()[:, None]).t().contiguous()


def pack_stem_s2d(w, pad, dtype):
    """Stride-2 kxk conv on Cin (<=4) channels -> stride-1 TxT conv on the space-to-depth input.

    out(ho,wo) = sum_{ky,kx,ch} w[o,ch,ky,kx] x[2ho+ky-pad, 2wo+kx-pad, ch].  With 2ho+ky-pad = 2(ho+a)+p the
    tap index a runs over [a_min, a_max]; the packed weight is [Cout, T*T, ld] with channel (py*2+px)*Cin+ch.
    Returns (packed, T, pad2, ld): conv kernel T, top/left padding pad2 = -a_min, channel pitch ld.
    """
    co, ci, k, _ = w.shape
    offs = [ky - pad for ky in range(k)]
    a = [o // 2 for o in offs]                 # python floor division
    par = [o % 2 for o in offs]
    a_min, a_max = min(a), max(a)
    T = a_max - a_min + 1
    ld = round_up(4 * ci, 16)
    out = torch.zeros(co, T, T, ld, dtype=torch.float32, device=w.device)
    for ky in range(k):
        for kx in range(k):
            ch0 = (par[ky] * 2 + par[kx]) * ci
            out[:, a[ky] - a_min, a[kx] - a_min, ch0:ch0 + ci] = w[:, :, ky, kx].float()
    return out.reshape(co, T * T, ld).to(dtype).contiguous(), T, -a_min, ld

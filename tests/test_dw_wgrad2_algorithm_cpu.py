"""Algorithm of the opt-in depthwise weight-gradient kernel (csrc/dw_wgrad2.cu) without a GPU.

A thread-by-thread Python transcription of ``dw_wgrad_rows_kernel`` (same geometry arguments, same segment walk, the same
three-way rotation of the window columns with its 1- and 2-pixel tails, zero window entries outside the image, the dilated
gather path, the per-slab partial layout) is executed for every (block, thread) of small launches and compared with autograd's
depthwise weight gradient.  It pins the INDEXING of the kernel; the CUDA build itself is covered by the gated GPU test
(tests/test_train_kernels_gpu.py::test_depthwise_wgrad_v2)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def rows_geom(rows, c, max_slabs):
    """csrc/dw_wgrad2.cu rows_geom"""
    cqn = c // 4
    cls = 0
    while (1 << cls) < cqn and cls < 5:
        cls += 1
    cl, pl = 1 << cls, 256 >> cls
    gx = (cqn + cl - 1) // cl
    want = (148 * 6 + gx - 1) // gx
    by_rows = (rows + pl * 64 - 1) // (pl * 64)
    want = max(1, min(want, by_rows, max_slabs))
    run = ((rows + want - 1) // want + pl - 1) // pl
    pps = run * pl
    slabs = max(1, (rows + pps - 1) // pps)
    return cls, gx, slabs, pps, run


def simulate(x, dy, dilation, pre_relu, cls, gx, slabs, pps, run):
    """x, dy: float32 [n,h,w,c] (already rounded to the 16-bit dtype).  -> partial [slabs][9][c]"""
    n, h, w, c = x.shape
    xf = np.maximum(x, 0) if pre_relu else x
    xf = xf.reshape(-1, c)
    gf = dy.reshape(-1, c)
    rows = n * h * w
    cl, pl = 1 << cls, 256 >> cls
    partial = np.zeros((slabs, 9, c), dtype=np.float32)
    for by in range(slabs):
        for bx in range(gx):
            red = np.zeros((256, 9, 4), dtype=np.float32)
            for tid in range(256):
                lc, lp = tid & (cl - 1), tid >> cls
                cq = bx * cl + lc
                if cq >= c // 4:
                    continue
                ch = slice(cq * 4, cq * 4 + 4)
                s0 = by * pps
                s1 = min(s0 + pps, rows)
                p = s0 + lp * run
                pe = min(p + run, s1)
                acc = np.zeros((9, 4), dtype=np.float32)
                if p < pe:
                    xw, yh = p % w, (p // w) % h

                    def col(pix, dcol, col_ok, up_ok, down_ok):          # col_load: rows y-d, y, y+d at column offset dcol
                        out = np.zeros((3, 4), dtype=np.float32)
                        if col_ok:
                            if up_ok:
                                out[0] = xf[pix - dilation * w + dcol, ch]
                            out[1] = xf[pix + dcol, ch]
                            if down_ok:
                                out[2] = xf[pix + dilation * w + dcol, ch]
                        return out

                    def fma(g, l, m, r):                                  # window_fma
                        for ky in range(3):
                            acc[ky * 3 + 0] += g * l[ky]
                            acc[ky * 3 + 1] += g * m[ky]
                            acc[ky * 3 + 2] += g * r[ky]
                    while p < pe:
                        ln = min(pe - p, w - xw)
                        up_ok, down_ok = yh - dilation >= 0, yh + dilation < h
                        if dilation == 1:
                            a = col(p, -1, xw - 1 >= 0, up_ok, down_ok)
                            b = col(p, 0, True, up_ok, down_ok)
                            i = 0
                            while i + 3 <= ln:
                                cc = col(p + i, 1, xw + i + 1 < w, up_ok, down_ok)
                                fma(gf[p + i, ch], a, b, cc)
                                a = col(p + i + 1, 1, xw + i + 2 < w, up_ok, down_ok)
                                fma(gf[p + i + 1, ch], b, cc, a)
                                b = col(p + i + 2, 1, xw + i + 3 < w, up_ok, down_ok)
                                fma(gf[p + i + 2, ch], cc, a, b)
                                i += 3
                            if i < ln:
                                cc = col(p + i, 1, xw + i + 1 < w, up_ok, down_ok)
                                fma(gf[p + i, ch], a, b, cc)
                                if i + 1 < ln:
                                    a = col(p + i + 1, 1, xw + i + 2 < w, up_ok, down_ok)
                                    fma(gf[p + i + 1, ch], b, cc, a)
                        else:
                            for i in range(ln):
                                xc = xw + i
                                fma(gf[p + i, ch], col(p + i, -dilation, xc - dilation >= 0, up_ok, down_ok),
                                    col(p + i, 0, True, up_ok, down_ok), col(p + i, dilation, xc + dilation < w, up_ok, down_ok))
                        p += ln
                        xw = 0
                        yh = yh + 1 if yh + 1 < h else 0
                red[lp * cl + lc] = acc
            for lc in range(cl):                                          # lp == 0 threads: fixed-order sum over the pixel lanes
                cq = bx * cl + lc
                if cq < c // 4:
                    partial[by, :, cq * 4:cq * 4 + 4] = red[np.arange(pl) * cl + lc].sum(0)
    return partial


CASES = [
    # n, h, w, c, dilation, pre_relu, forced (run, pixels_per_slab) or None for the kernel's own geometry
    (2, 5, 7, 16, 1, False, None),
    (1, 6, 9, 8, 1, True, None),            # c = 8: two quads, cls 1
    (2, 4, 5, 24, 1, False, (3, 3 * 32)),    # runs of 3 pixels: every rotation phase starts mid-row, several slabs
    (1, 7, 8, 16, 1, False, (5, 5 * 64)),    # runs of 5: tails of 2 and row wraps inside a run
    (1, 3, 4, 16, 1, False, (1, 64)),        # single-pixel runs: tail of 1 only
    (2, 9, 11, 16, 2, False, None),          # dilated gather path
    (1, 13, 8, 16, 6, True, (7, 7 * 64)),    # dilation larger than half the width, leading ReLU
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_rows_kernel_transcription_matches_autograd(case):
    n, h, w, c, dil, pre_relu, forced = CASES[case]
    g = torch.Generator().manual_seed(100 + case)
    x = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).float()
    dy = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).float()
    wt = torch.zeros(c, 1, 3, 3, requires_grad=True)
    xin = x.permute(0, 3, 1, 2)
    F.conv2d(F.relu(xin) if pre_relu else xin, wt, None, 1, dil, dil, groups=c).backward(dy.permute(0, 3, 1, 2))
    cls, gx, slabs, pps, run = rows_geom(n * h * w, c, 1 << 20)
    if forced is not None:
        run, pps = forced
        assert pps == run * (256 >> cls)
        slabs = (n * h * w + pps - 1) // pps
    partial = simulate(x.numpy(), dy.numpy(), dil, pre_relu, cls, gx, slabs, pps, run)
    dw = partial.sum(0).T                                               # reduce_partials(stride_k=1, stride_c=9) -> [c][9]
    ref = wt.grad.reshape(c, 9).numpy()
    assert np.abs(dw - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max()), np.abs(dw - ref).max()


def test_geometry_covers_every_pixel_once():
    for rows, c in [(8 * 65 * 129, 728), (4 * 513 * 1025, 128), (4 * 257 * 513, 256), (2 * 17 * 33, 128), (70, 16), (1, 8)]:
        cls, gx, slabs, pps, run = rows_geom(rows, c, 1 << 20)
        pl = 256 >> cls
        assert pps == run * pl and slabs * pps >= rows and (slabs - 1) * pps < rows
        assert gx * (1 << cls) * 4 >= c

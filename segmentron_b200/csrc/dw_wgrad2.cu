// segb200 -- depthwise 3x3 weight gradient, sliding-window variant (OPT-IN: segb200_dw_wgrad_v2; the default is
// segb200_dw_wgrad, train.cu).  Same contract and the same partial[(slab*9 + tap)*c + ch] layout.
//
// Why: the default kernel re-loads and re-converts all nine taps for every pixel (9 x 16 B cached loads, 72 conversions and 72
// scalar FMAs per 8 channels, 128 registers) and is issue-bound (17.5 of 75 ms of the Xception65 training step, round 1).
// Here a thread owns FOUR channels and walks CONTIGUOUS pixels of an image row, keeping the 3x3 window of its channels as fp32
// registers: per pixel it loads and converts only the window's new column (3 x 8 B) and the gradient (8 B) and issues 18 packed
// fma.rn.f32x2 -- about half the issue slots per channel-pixel (128 registers, 2 blocks / SM like the default kernel).
// Dilation d > 1: consecutive pixels share no taps, so the nine taps are gathered directly (the packed FMAs still apply).
// Out-of-image taps are ZERO window entries, so the inner loop has no branches.  Bound: FP32 issue, then L2 (each x element is
// read 3 times from L1/L2, once from HBM).
#include "vec.cuh"
#include "../../include/segb200.h"

namespace segb200 {

__device__ __forceinline__ float2 ffma2_(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

struct Quad { float2 lo, hi; };                        // four consecutive channels as two fp32 pairs

__device__ __forceinline__ Quad quad_zero() { Quad q; q.lo = make_float2(0.f, 0.f); q.hi = make_float2(0.f, 0.f); return q; }

// 8-byte load of 4 consecutive 16-bit channels -> fp32 (optionally through ReLU)
__device__ __forceinline__ Quad quad_load(const char* p, int dtype, int relu) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  Quad q;
  q.lo = unpack_any(v.x, dtype);
  q.hi = unpack_any(v.y, dtype);
  if (relu) {
    q.lo.x = fmaxf(q.lo.x, 0.f); q.lo.y = fmaxf(q.lo.y, 0.f);
    q.hi.x = fmaxf(q.hi.x, 0.f); q.hi.y = fmaxf(q.hi.y, 0.f);
  }
  return q;
}

__device__ __forceinline__ void quad_fma(Quad& acc, const Quad& g, const Quad& x) {
  acc.lo = ffma2_(g.lo, x.lo, acc.lo);
  acc.hi = ffma2_(g.hi, x.hi, acc.hi);
}

// one 3-row column of the window (rows y-d, y, y+d at one image column); rows / columns outside the image are zeros
struct Col { Quad r[3]; };

__device__ __forceinline__ Col col_load(const char* xpix, long long row_off, long long col_byte_off, bool col_ok, bool up_ok,
                                        bool down_ok, int dtype, int relu) {
  Col c;
  c.r[0] = (col_ok && up_ok) ? quad_load(xpix - row_off + col_byte_off, dtype, relu) : quad_zero();
  c.r[1] = col_ok ? quad_load(xpix + col_byte_off, dtype, relu) : quad_zero();
  c.r[2] = (col_ok && down_ok) ? quad_load(xpix + row_off + col_byte_off, dtype, relu) : quad_zero();
  return c;
}

// acc[ky*3 + kx] += g * x'[y + (ky-1) d][x + (kx-1) d]   for the three columns (left, mid, right)
__device__ __forceinline__ void window_fma(Quad (&acc)[9], const Quad& g, const Col& l, const Col& m, const Col& r) {
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    quad_fma(acc[ky * 3 + 0], g, l.r[ky]);
    quad_fma(acc[ky * 3 + 1], g, m.r[ky]);
    quad_fma(acc[ky * 3 + 2], g, r.r[ky]);
  }
}

// block = (256 >> cls) pixel lanes x (1 << cls) channel-quad lanes; grid.x covers the quads, grid.y the pixel slabs; inside a slab
// every pixel lane owns ONE contiguous run of `run` pixels (flattened n*h*w order).
__global__ void __launch_bounds__(256, 2)
dw_wgrad_rows_kernel(const void* __restrict__ x, const void* __restrict__ dy, int n, int h, int w, int c, int x_ld, int dy_ld,
                     int dilation, int pre_relu, int dtype, int cls, long long pixels_per_slab, long long run,
                     float* __restrict__ partial) {
  __shared__ float red[256][4];
  const int cl = 1 << cls, pl = 256 >> cls;
  const int lc = threadIdx.x & (cl - 1), lp = threadIdx.x >> cls;
  const int cq = blockIdx.x * cl + lc;
  const bool valid = cq < c / 4;
  const long long rows = (long long)n * h * w;
  const long long s0 = (long long)blockIdx.y * pixels_per_slab;
  long long s1 = s0 + pixels_per_slab; if (s1 > rows) s1 = rows;
  long long p = s0 + (long long)lp * run;
  long long pe = p + run; if (pe > s1) pe = s1;
  Quad acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = quad_zero();
  if (valid && p < pe) {
    const char* xb = reinterpret_cast<const char*>(x) + (long long)cq * 8;
    const char* gb = reinterpret_cast<const char*>(dy) + (long long)cq * 8;
    const long long xs = (long long)x_ld * 2, gs = (long long)dy_ld * 2;
    const long long row_off = (long long)dilation * w * xs, col_off = (long long)dilation * xs;
    int xw = (int)(p % w), yh = (int)((p / w) % h);
    while (p < pe) {
      // one segment: pixels p .. p+len-1 of image row yh, starting at column xw
      long long len = pe - p;
      if (len > w - xw) len = w - xw;
      const bool up_ok = yh - dilation >= 0, down_ok = yh + dilation < h;
      const char* xp = xb + p * xs;
      const char* gp = gb + p * gs;
      if (dilation == 1) {
        Col a = col_load(xp, row_off, -xs, xw - 1 >= 0, up_ok, down_ok, dtype, pre_relu);
        Col b = col_load(xp, row_off, 0, true, up_ok, down_ok, dtype, pre_relu);
        Col cc;
        long long i = 0;
        // roles rotate (left, mid, right) = (a, b, cc) -> (b, cc, a) -> (cc, a, b): no register moves
        for (; i + 3 <= len; i += 3) {
          cc = col_load(xp + i * xs, row_off, xs, xw + (int)i + 1 < w, up_ok, down_ok, dtype, pre_relu);
          window_fma(acc, quad_load(gp + i * gs, dtype, 0), a, b, cc);
          a = col_load(xp + (i + 1) * xs, row_off, xs, xw + (int)i + 2 < w, up_ok, down_ok, dtype, pre_relu);
          window_fma(acc, quad_load(gp + (i + 1) * gs, dtype, 0), b, cc, a);
          b = col_load(xp + (i + 2) * xs, row_off, xs, xw + (int)i + 3 < w, up_ok, down_ok, dtype, pre_relu);
          window_fma(acc, quad_load(gp + (i + 2) * gs, dtype, 0), cc, a, b);
        }
        if (i < len) {                                   // 1 or 2 pixels left; (a, b) are (left, mid) of pixel i
          cc = col_load(xp + i * xs, row_off, xs, xw + (int)i + 1 < w, up_ok, down_ok, dtype, pre_relu);
          window_fma(acc, quad_load(gp + i * gs, dtype, 0), a, b, cc);
          if (i + 1 < len) {
            a = col_load(xp + (i + 1) * xs, row_off, xs, xw + (int)i + 2 < w, up_ok, down_ok, dtype, pre_relu);
            window_fma(acc, quad_load(gp + (i + 1) * gs, dtype, 0), b, cc, a);
          }
        }
      } else {
        for (long long i = 0; i < len; ++i) {
          const int xc = xw + (int)i;
          const Col l = col_load(xp + i * xs, row_off, -col_off, xc - dilation >= 0, up_ok, down_ok, dtype, pre_relu);
          const Col m = col_load(xp + i * xs, row_off, 0, true, up_ok, down_ok, dtype, pre_relu);
          const Col r = col_load(xp + i * xs, row_off, col_off, xc + dilation < w, up_ok, down_ok, dtype, pre_relu);
          window_fma(acc, quad_load(gp + i * gs, dtype, 0), l, m, r);
        }
      }
      p += len;
      xw = 0;
      if (++yh == h) yh = 0;
    }
  }
  // fixed-order reduction over the pixel lanes, one tap at a time
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    __syncthreads();
    red[lp * cl + lc][0] = acc[t].lo.x; red[lp * cl + lc][1] = acc[t].lo.y;
    red[lp * cl + lc][2] = acc[t].hi.x; red[lp * cl + lc][3] = acc[t].hi.y;
    __syncthreads();
    if (lp == 0 && valid) {
      float s[4] = {0.f, 0.f, 0.f, 0.f};
      for (int q = 0; q < pl; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] += red[q * cl + lc][j];
      *reinterpret_cast<float4*>(partial + ((long long)blockIdx.y * 9 + t) * c + cq * 4) = make_float4(s[0], s[1], s[2], s[3]);
    }
  }
}

struct RowsGeom { int cls, gx, slabs; long long pixels_per_slab, run; };

static RowsGeom rows_geom(long long rows, int c, int max_slabs) {
  RowsGeom g;
  const int cqn = c / 4;
  g.cls = 0;
  while ((1 << g.cls) < cqn && g.cls < 5) ++g.cls;
  const int cl = 1 << g.cls, pl = 256 >> g.cls;
  g.gx = (cqn + cl - 1) / cl;
  long long want = (148LL * 6 + g.gx - 1) / g.gx;                 // ~6 blocks per SM in total (3 resident)
  const long long by_rows = (rows + (long long)pl * 64 - 1) / ((long long)pl * 64);   // runs of >= 64 pixels amortise the window set-up
  if (want > by_rows) want = by_rows;
  if (want > max_slabs) want = max_slabs;
  if (want < 1) want = 1;
  g.run = ((rows + want - 1) / want + pl - 1) / pl;
  g.pixels_per_slab = g.run * pl;
  g.slabs = (int)((rows + g.pixels_per_slab - 1) / g.pixels_per_slab);
  if (g.slabs < 1) g.slabs = 1;
  return g;
}

}  // namespace segb200

using namespace segb200;

// number of slabs segb200_dw_wgrad_v2 will write for this problem (partial must hold slabs * 9 * c floats)
extern "C" int segb200_dw_wgrad_v2_slabs(long long rows, int c, int max_slabs) {
  if (rows < 1 || c < 8) return 1;
  return rows_geom(rows, c, max_slabs > 0 ? max_slabs : 1 << 20).slabs;
}

extern "C" int segb200_dw_wgrad_v2(const void* x, const void* dy, float* partial, int n, int h, int w, int c, int x_ld, int dy_ld,
                                   int dilation, int pre_relu, int dtype, int max_slabs, void* stream) {
  if (!x || !dy || !partial) return set_error(-1, "dw_wgrad_v2: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "dw_wgrad_v2: bad dtype");
  if (c < 8 || (c & 7) || (x_ld & 7) || (dy_ld & 7) || x_ld < c || dy_ld < c || dilation < 1 || n < 1 || h < 1 || w < 1)
    return set_error(-4, "dw_wgrad_v2: bad c/pitches");
  const long long rows = (long long)n * h * w;
  const RowsGeom g = rows_geom(rows, c, max_slabs > 0 ? max_slabs : 1 << 20);
  dw_wgrad_rows_kernel<<<dim3(g.gx, g.slabs), 256, 0, STREAM(stream)>>>(x, dy, n, h, w, c, x_ld, dy_ld, dilation, pre_relu ? 1 : 0,
                                                                        dtype, g.cls, g.pixels_per_slab, g.run, partial);
  return check_launch("dw_wgrad_v2");
}

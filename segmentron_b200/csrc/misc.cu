// segb200 -- memory-bound glue kernels (NHWC, 128-bit vectors, fp32 math): stem packing, pooling,
// bilinear resize, layout converters.  All are HBM-roofline kernels (no data reuse beyond L1/L2).
#include "common.cuh"
#include "vec.cuh"
#include "../../include/segb200.h"

namespace segb200 {

// -------------------------------------------------------------------------------------------
// space-to-depth stem packing
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_s2d_kernel(const void* __restrict__ x, int x_dtype, void* __restrict__ out, int out_dtype, int n, int c, int h,
                int w, int hs, int ws, int out_ld) {
  const int groups = out_ld / 8;
  const long long total = (long long)n * hs * ws * groups;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % groups);
    long long r = idx / groups;
    const int j = (int)(r % ws); r /= ws;
    const int i = (int)(r % hs);
    const int b = (int)(r / hs);
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch4 = g * 8 + k;
      float v = 0.f;
      if (ch4 < 4 * c) {
        const int par = ch4 / c, ch = ch4 - par * c;
        const int yy = 2 * i + (par >> 1), xx = 2 * j + (par & 1);
        if (yy < h && xx < w) v = load_any(x, (((long long)b * c + ch) * h + yy) * w + xx, x_dtype);
      }
      f[k] = v;
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<char*>(out) + (((long long)b * hs + i) * ws + j) * out_ld * 2 + g * 16) =
        pack8(f, out_dtype);
  }
}

// One thread per (output pixel, 8-channel group) on a 2-D grid (x: pixel-of-row x group, y: image x row): same loads and stores as
// pack_s2d_kernel without its 64-bit index divisions and grid-stride loop (196 -> measured in profiles/ for the 8x3x1025x2049 input).
__global__ void __launch_bounds__(256)
pack_s2d_rows_kernel(const void* __restrict__ x, int x_dtype, void* __restrict__ out, int out_dtype, int c, int h, int w, int hs,
                     int ws, int out_ld) {
  const int groups = out_ld / 8;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= ws * groups) return;
  const int j = t / groups, g = t - j * groups;
  const int b = blockIdx.y / hs, i = blockIdx.y - b * hs;
  float f[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch4 = g * 8 + k;
    float v = 0.f;
    if (ch4 < 4 * c) {
      const int par = ch4 / c, ch = ch4 - par * c;
      const int yy = 2 * i + (par >> 1), xx = 2 * j + (par & 1);
      if (yy < h && xx < w) v = load_any(x, (((long long)b * c + ch) * h + yy) * w + xx, x_dtype);
    }
    f[k] = v;
  }
  *reinterpret_cast<uint4*>(reinterpret_cast<char*>(out) + (((long long)b * hs + i) * ws + j) * out_ld * 2 + g * 16) = pack8(f, out_dtype);
}

// -------------------------------------------------------------------------------------------
// global average pool: one block per (image, 16 channel vectors); 16 pixel lanes x 16 channel lanes,
// fixed-order accumulation and a fixed-order shared-memory reduction => bit-reproducible (no atomics)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gap_kernel(const void* __restrict__ x, void* __restrict__ out, int hw, int c, int x_ld, int dtype) {
  __shared__ float red[16][16][8];
  const int n = blockIdx.y;
  const int lc = threadIdx.x & 15, lp = threadIdx.x >> 4;
  const int cv = blockIdx.x * 16 + lc;
  const int cvn = c / 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (cv < cvn) {
    const char* base = reinterpret_cast<const char*>(x) + ((long long)n * hw * x_ld + cv * 8) * 2;
    int p = lp;
    for (; p + 48 < hw; p += 64) {                   // 4 independent 128-bit loads in flight per thread
      float f0[8], f1[8], f2[8], f3[8];
      unpack8(ldg_nc_v4(base + (long long)p * x_ld * 2), dtype, f0);
      unpack8(ldg_nc_v4(base + (long long)(p + 16) * x_ld * 2), dtype, f1);
      unpack8(ldg_nc_v4(base + (long long)(p + 32) * x_ld * 2), dtype, f2);
      unpack8(ldg_nc_v4(base + (long long)(p + 48) * x_ld * 2), dtype, f3);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += (f0[j] + f1[j]) + (f2[j] + f3[j]);
    }
    for (; p < hw; p += 16) {
      float f[8];
      unpack8(ldg_nc_v4(base + (long long)p * x_ld * 2), dtype, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[lp][lc][j] = acc[j];
  __syncthreads();
  if (lp == 0 && cv < cvn) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = 0.f;
    for (int q = 0; q < 16; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] += red[q][lc][j];
    const float inv = 1.f / (float)hw;
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] *= inv;
    *reinterpret_cast<uint4*>(reinterpret_cast<char*>(out) + ((long long)n * c + cv * 8) * 2) = pack8(t, dtype);
  }
}

// Wide variant (default): 1024 threads = 8 channel-vector lanes x 128 pixel lanes per block, one block per (image, 8 channel
// vectors): twice the blocks and 8x the loads in flight per SM of gap_kernel (which leaves 20 of 148 SMs idle at c = 2048, batch 8).
// Fixed-order accumulation per lane, fixed-order shared-memory tree => bit-reproducible.
__global__ void __launch_bounds__(1024)
gap_wide_kernel(const void* __restrict__ x, void* __restrict__ out, int hw, int c, int x_ld, int dtype) {
  __shared__ float red[128][8][8];                 // 32 KB
  const int n = blockIdx.y;
  const int lc = threadIdx.x & 7, lp = threadIdx.x >> 3;
  const int cv = blockIdx.x * 8 + lc;
  const int cvn = c / 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (cv < cvn) {
    const char* base = reinterpret_cast<const char*>(x) + ((long long)n * hw * x_ld + cv * 8) * 2;
    int p = lp;
    for (; p + 384 < hw; p += 512) {                 // 4 independent 128-bit loads in flight per thread
      float f0[8], f1[8], f2[8], f3[8];
      unpack8(ldg_nc_v4(base + (long long)p * x_ld * 2), dtype, f0);
      unpack8(ldg_nc_v4(base + (long long)(p + 128) * x_ld * 2), dtype, f1);
      unpack8(ldg_nc_v4(base + (long long)(p + 256) * x_ld * 2), dtype, f2);
      unpack8(ldg_nc_v4(base + (long long)(p + 384) * x_ld * 2), dtype, f3);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += (f0[j] + f1[j]) + (f2[j] + f3[j]);
    }
    for (; p < hw; p += 128) {
      float f[8];
      unpack8(ldg_nc_v4(base + (long long)p * x_ld * 2), dtype, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[lp][lc][j] = acc[j];
  __syncthreads();
  for (int half = 64; half >= 1; half >>= 1) {       // fixed-order tree over the 128 pixel lanes
    if (lp < half) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[lp][lc][j] += red[lp + half][lc][j];
    }
    __syncthreads();
  }
  if (lp == 0 && cv < cvn) {
    float t[8];
    const float inv = 1.f / (float)hw;
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = red[0][lc][j] * inv;
    *reinterpret_cast<uint4*>(reinterpret_cast<char*>(out) + ((long long)n * c + cv * 8) * 2) = pack8(t, dtype);
  }
}

// -------------------------------------------------------------------------------------------
// adaptive average pool to s x s
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
adaptive_pool_kernel(const void* __restrict__ x, void* __restrict__ out, int n, int h, int w, int c, int x_ld, int s,
                     int out_ld, int dtype) {
  const int cvn = c / 8;
  const long long total = (long long)n * s * s * cvn;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvn);
    long long r = idx / cvn;
    const int bj = (int)(r % s); r /= s;
    const int bi = (int)(r % s);
    const int b = (int)(r / s);
    const int h0 = (bi * h) / s, h1 = ((bi + 1) * h + s - 1) / s;
    const int w0 = (bj * w) / s, w1 = ((bj + 1) * w + s - 1) / s;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int yy = h0; yy < h1; ++yy)
      for (int xx = w0; xx < w1; ++xx) {
        float f[8];
        unpack8(ldg_nc_v4(reinterpret_cast<const char*>(x) + ((((long long)b * h + yy) * w + xx) * x_ld + cv * 8) * 2),
                dtype, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    *reinterpret_cast<uint4*>(reinterpret_cast<char*>(out) + ((((long long)b * s + bi) * s + bj) * out_ld + cv * 8) * 2) =
        pack8(acc, dtype);
  }
}


// -------------------------------------------------------------------------------------------
// max pool 3x3, stride 2, padding 1 (-inf padding), NHWC
// -------------------------------------------------------------------------------------------
template <bool kBF16>
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const void* __restrict__ x, void* __restrict__ y, int n, int h, int w, int c, int x_ld, int ho, int wo,
                    int y_ld) {
  using H = Half2<kBF16>;
  const int cvn = c / 8;
  const long long total = (long long)n * ho * wo * cvn;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvn);
    long long r = idx / cvn;
    const int ox = (int)(r % wo); r /= wo;
    const int oy = (int)(r % ho);
    const int b = (int)(r / ho);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - 1 + ky;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - 1 + kx;
        if (ix < 0 || ix >= w) continue;
        const uint4 v = ldg_v4(reinterpret_cast<const char*>(x) + ((((long long)b * h + iy) * w + ix) * x_ld + cv * 8) * 2);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = H::unpack(u[j]);
          m[2 * j] = fmaxf(m[2 * j], f.x); m[2 * j + 1] = fmaxf(m[2 * j + 1], f.y);
        }
      }
    }
    uint4 o;
    o.x = H::pack(m[0], m[1]); o.y = H::pack(m[2], m[3]); o.z = H::pack(m[4], m[5]); o.w = H::pack(m[6], m[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<char*>(y) + ((((long long)b * ho + oy) * wo + ox) * y_ld + cv * 8) * 2) = o;
  }
}

// -------------------------------------------------------------------------------------------
// HRNet fuse: y = act(a + nearest_up(z, 2^k))   (k = 0: plain add).  hrnet.py:178-186,215-232
// -------------------------------------------------------------------------------------------
template <bool kBF16>
__global__ void __launch_bounds__(256)
upsample_add_kernel(const void* __restrict__ a, const void* __restrict__ z, void* __restrict__ y, int n, int h, int w, int c,
                    int a_ld, int z_ld, int y_ld, int k, int act) {
  using H = Half2<kBF16>;
  const int cvn = c / 8;
  const int hz = h >> k, wz = w >> k;
  const long long total = (long long)n * h * w * cvn;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvn);
    long long r = idx / cvn;
    const int x = (int)(r % w); r /= w;
    const int yy = (int)(r % h);
    const int b = (int)(r / h);
    const uint4 va = ldg_nc_v4(reinterpret_cast<const char*>(a) + ((((long long)b * h + yy) * w + x) * a_ld + cv * 8) * 2);
    const uint4 vz = ldg_v4(reinterpret_cast<const char*>(z) + ((((long long)b * hz + (yy >> k)) * wz + (x >> k)) * z_ld + cv * 8) * 2);
    const uint32_t ua[4] = {va.x, va.y, va.z, va.w}, uz[4] = {vz.x, vz.y, vz.z, vz.w};
    uint4 o; uint32_t* po = &o.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = H::unpack(ua[j]), fz = H::unpack(uz[j]);
      po[j] = H::pack(apply_act(fa.x + fz.x, act), apply_act(fa.y + fz.y, act));
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<char*>(y) + ((((long long)b * h + yy) * w + x) * y_ld + cv * 8) * 2) = o;
  }
}

// -------------------------------------------------------------------------------------------
// bilinear resize (torch upsample_bilinear2d index rules, fp32 math)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bilinear_nhwc_kernel(const void* __restrict__ x, void* __restrict__ y, int n, int hi, int wi, int c, int x_ld, int ho,
                     int wo, int y_ld, int align, int dtype) {
  const int cvn = c / 8;
  const long long total = (long long)n * ho * wo * cvn;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvn);
    long long r = idx / cvn;
    const int ox = (int)(r % wo); r /= wo;
    const int oy = (int)(r % ho);
    const int b = (int)(r / ho);
    const Lerp ly = lerp_coord(oy, hi, ho, align), lx = lerp_coord(ox, wi, wo, align);
    const char* base = reinterpret_cast<const char*>(x) + ((long long)b * hi * wi * x_ld + cv * 8) * 2;
    float f00[8], f01[8], f10[8], f11[8], o[8];
    unpack8(ldg_v4(base + ((long long)ly.i0 * wi + lx.i0) * x_ld * 2), dtype, f00);
    unpack8(ldg_v4(base + ((long long)ly.i0 * wi + lx.i1) * x_ld * 2), dtype, f01);
    unpack8(ldg_v4(base + ((long long)ly.i1 * wi + lx.i0) * x_ld * 2), dtype, f10);
    unpack8(ldg_v4(base + ((long long)ly.i1 * wi + lx.i1) * x_ld * 2), dtype, f11);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] = ly.l0 * (lx.l0 * f00[j] + lx.l1 * f01[j]) + ly.l1 * (lx.l0 * f10[j] + lx.l1 * f11[j]);
    *reinterpret_cast<uint4*>(reinterpret_cast<char*>(y) + ((((long long)b * ho + oy) * wo + ox) * y_ld + cv * 8) * 2) =
        pack8(o, dtype);
  }
}

// Strip variant (default): a thread owns one (output column, 8-channel vector) and kRows consecutive output rows; rows that share
// their two source rows (4x up-sampling: all four) reuse the loaded neighbours and the horizontal interpolation.  2-D grid, 32-bit
// index math, compile-time dtype.  Explicit mul / fma in the contraction nvcc applies to bilinear_nhwc_kernel's expression:
// bit-identical results.
template <bool kBF16, int kRows>
__global__ void __launch_bounds__(256)
bilinear_nhwc_strip_kernel(const void* __restrict__ x, void* __restrict__ y, int hi, int wi, int c, int x_ld, int ho, int wo, int y_ld,
                           int align, int groups) {
  using H = Half2<kBF16>;
  const int cvn = c / 8;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= wo * cvn) return;
  const int ox = t / cvn, cv = t - ox * cvn;
  const int b = blockIdx.y / groups, g = blockIdx.y - b * groups;
  const Lerp lx = lerp_coord(ox, wi, wo, align);
  const char* base = reinterpret_cast<const char*>(x) + ((long long)b * hi * wi * x_ld + cv * 8) * 2;
  float A[8], B[8];
  int ci0 = -1, ci1 = -1;
#pragma unroll 1
  for (int r = 0; r < kRows; ++r) {
    const int oy = g * kRows + r;
    if (oy >= ho) break;
    const Lerp ly = lerp_coord(oy, hi, ho, align);
    if (ly.i0 != ci0 || ly.i1 != ci1) {
      ci0 = ly.i0; ci1 = ly.i1;
      const uint4 v00 = ldg_v4(base + ((long long)ly.i0 * wi + lx.i0) * x_ld * 2), v01 = ldg_v4(base + ((long long)ly.i0 * wi + lx.i1) * x_ld * 2);
      const uint4 v10 = ldg_v4(base + ((long long)ly.i1 * wi + lx.i0) * x_ld * 2), v11 = ldg_v4(base + ((long long)ly.i1 * wi + lx.i1) * x_ld * 2);
      const uint32_t u00[4] = {v00.x, v00.y, v00.z, v00.w}, u01[4] = {v01.x, v01.y, v01.z, v01.w};
      const uint32_t u10[4] = {v10.x, v10.y, v10.z, v10.w}, u11[4] = {v11.x, v11.y, v11.z, v11.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f00 = H::unpack(u00[j]), f01 = H::unpack(u01[j]), f10 = H::unpack(u10[j]), f11 = H::unpack(u11[j]);
        A[2 * j] = __fmaf_rn(lx.l0, f00.x, __fmul_rn(lx.l1, f01.x));
        A[2 * j + 1] = __fmaf_rn(lx.l0, f00.y, __fmul_rn(lx.l1, f01.y));
        B[2 * j] = __fmaf_rn(lx.l0, f10.x, __fmul_rn(lx.l1, f11.x));
        B[2 * j + 1] = __fmaf_rn(lx.l0, f10.y, __fmul_rn(lx.l1, f11.y));
      }
    }
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      pk[j] = H::pack(__fmaf_rn(ly.l0, A[2 * j], __fmul_rn(ly.l1, B[2 * j])), __fmaf_rn(ly.l0, A[2 * j + 1], __fmul_rn(ly.l1, B[2 * j + 1])));
    *reinterpret_cast<uint4*>(reinterpret_cast<char*>(y) + ((((long long)b * ho + oy) * wo + ox) * y_ld + cv * 8) * 2) =
        make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// NHWC low-res logits -> NCHW full-res (+ optional fused argmax). One thread per output pixel.
template <int kMaxC>
__global__ void __launch_bounds__(256)
bilinear_nchw_out_kernel(const void* __restrict__ x, void* __restrict__ y, uint8_t* __restrict__ amax, int n, int hi,
                         int wi, int c, int x_ld, int ho, int wo, int align, int dtype, int out_dtype) {
  const long long total = (long long)n * ho * wo;
  const long long plane = (long long)ho * wo;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % wo);
    long long r = idx / wo;
    const int oy = (int)(r % ho);
    const int b = (int)(r / ho);
    const Lerp ly = lerp_coord(oy, hi, ho, align), lx = lerp_coord(ox, wi, wo, align);
    const char* base = reinterpret_cast<const char*>(x) + (long long)b * hi * wi * x_ld * 2;
    const char* p00 = base + ((long long)ly.i0 * wi + lx.i0) * x_ld * 2;
    const char* p01 = base + ((long long)ly.i0 * wi + lx.i1) * x_ld * 2;
    const char* p10 = base + ((long long)ly.i1 * wi + lx.i0) * x_ld * 2;
    const char* p11 = base + ((long long)ly.i1 * wi + lx.i1) * x_ld * 2;
    float best = -INFINITY; int besti = 0;
    char* yb = reinterpret_cast<char*>(y);
#pragma unroll
    for (int cv = 0; cv < kMaxC / 8; ++cv) {
      if (cv * 8 >= c) break;
      float f00[8], f01[8], f10[8], f11[8];
      unpack8(ldg_v4(p00 + cv * 16), dtype, f00);
      unpack8(ldg_v4(p01 + cv * 16), dtype, f01);
      unpack8(ldg_v4(p10 + cv * 16), dtype, f10);
      unpack8(ldg_v4(p11 + cv * 16), dtype, f11);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ch = cv * 8 + j;
        if (ch < c) {
          float o = ly.l0 * (lx.l0 * f00[j] + lx.l1 * f01[j]) + ly.l1 * (lx.l0 * f10[j] + lx.l1 * f11[j]);
          o = round_any(o, out_dtype);
          const long long oi = ((long long)b * c + ch) * plane + (long long)oy * wo + ox;
          if (out_dtype == DT_F32) reinterpret_cast<float*>(yb)[oi] = o;
          else if (out_dtype == DT_BF16) reinterpret_cast<__nv_bfloat16*>(yb)[oi] = __float2bfloat16_rn(o);
          else reinterpret_cast<__half*>(yb)[oi] = __float2half_rn(o);
          if (o > best) { best = o; besti = ch; }
        }
      }
    }
    if (amax != nullptr) amax[idx] = (uint8_t)besti;
  }
}

// Strip variant (default for c <= 32): a thread owns ONE output column and kRows consecutive output rows.  With the 4x up-sampling
// of the segmentation heads the rows of a strip share their two source rows, so the four neighbours are loaded and unpacked, and
// the horizontal interpolation is done, once per strip instead of once per pixel; data types are template parameters (the
// one-pixel kernel above dispatches on them per element).  The arithmetic is spelled with explicit mul / fma in exactly the
// contraction nvcc applies to `ly.l0 * (lx.l0 * f00 + lx.l1 * f01) + ly.l1 * (lx.l0 * f10 + lx.l1 * f11)` -- bit-identical
// to bilinear_nchw_out_kernel and to the fused consumers (metric.cu, evaluate.cu, train.cu upsample_ce).
template <int kCV, bool kInBF16, int kOut, int kRows>
__global__ void __launch_bounds__(128)
bilinear_nchw_strip_kernel(const void* __restrict__ x, void* __restrict__ y, uint8_t* __restrict__ amax, int n, int hi, int wi, int c,
                           int x_ld, int ho, int wo, int align, int groups) {
  using H = Half2<kInBF16>;
  const int ox = blockIdx.x * 128 + threadIdx.x;
  if (ox >= wo) return;
  const int b = blockIdx.y / groups, g = blockIdx.y - b * groups;
  const long long plane = (long long)ho * wo;
  const Lerp lx = lerp_coord(ox, wi, wo, align);
  const char* base = reinterpret_cast<const char*>(x) + (long long)b * hi * wi * x_ld * 2;
  float A[kCV * 8], B[kCV * 8];
  int ci0 = -1, ci1 = -1;
  char* yb = reinterpret_cast<char*>(y);
#pragma unroll 1
  for (int r = 0; r < kRows; ++r) {
    const int oy = g * kRows + r;
    if (oy >= ho) break;
    const Lerp ly = lerp_coord(oy, hi, ho, align);
    if (ly.i0 != ci0 || ly.i1 != ci1) {                // new pair of source rows: horizontal interpolation of both
      ci0 = ly.i0; ci1 = ly.i1;
      const char* p00 = base + ((long long)ly.i0 * wi + lx.i0) * x_ld * 2;
      const char* p01 = base + ((long long)ly.i0 * wi + lx.i1) * x_ld * 2;
      const char* p10 = base + ((long long)ly.i1 * wi + lx.i0) * x_ld * 2;
      const char* p11 = base + ((long long)ly.i1 * wi + lx.i1) * x_ld * 2;
#pragma unroll
      for (int cv = 0; cv < kCV; ++cv) {
        const uint4 v00 = ldg_v4(p00 + cv * 16), v01 = ldg_v4(p01 + cv * 16), v10 = ldg_v4(p10 + cv * 16), v11 = ldg_v4(p11 + cv * 16);
        const uint32_t u00[4] = {v00.x, v00.y, v00.z, v00.w}, u01[4] = {v01.x, v01.y, v01.z, v01.w};
        const uint32_t u10[4] = {v10.x, v10.y, v10.z, v10.w}, u11[4] = {v11.x, v11.y, v11.z, v11.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f00 = H::unpack(u00[j]), f01 = H::unpack(u01[j]), f10 = H::unpack(u10[j]), f11 = H::unpack(u11[j]);
          A[cv * 8 + 2 * j] = __fmaf_rn(lx.l0, f00.x, __fmul_rn(lx.l1, f01.x));
          A[cv * 8 + 2 * j + 1] = __fmaf_rn(lx.l0, f00.y, __fmul_rn(lx.l1, f01.y));
          B[cv * 8 + 2 * j] = __fmaf_rn(lx.l0, f10.x, __fmul_rn(lx.l1, f11.x));
          B[cv * 8 + 2 * j + 1] = __fmaf_rn(lx.l0, f10.y, __fmul_rn(lx.l1, f11.y));
        }
      }
    }
    float best = -INFINITY; int besti = 0;
    const long long pix = (long long)oy * wo + ox;
#pragma unroll
    for (int ch = 0; ch < kCV * 8; ++ch) {
      if (ch < c) {
        float o = __fmaf_rn(ly.l0, A[ch], __fmul_rn(ly.l1, B[ch]));
        const long long oi = ((long long)b * c + ch) * plane + pix;
        if (kOut == DT_F32) {
          reinterpret_cast<float*>(yb)[oi] = o;
        } else if (kOut == DT_BF16) {
          const __nv_bfloat16 q = __float2bfloat16_rn(o);
          reinterpret_cast<__nv_bfloat16*>(yb)[oi] = q;
          o = __bfloat162float(q);
        } else {
          const __half q = __float2half_rn(o);
          reinterpret_cast<__half*>(yb)[oi] = q;
          o = __half2float(q);
        }
        if (o > best) { best = o; besti = ch; }
      }
    }
    if (amax != nullptr) amax[(long long)b * plane + pix] = (uint8_t)besti;
  }
}

static int g_bilinear_out_v1 = 0;   // 1: the one-pixel-per-thread kernel for every c (A/B knob "bilinear_out_v1")
int set_bilinear_out_v1(int v) { g_bilinear_out_v1 = v ? 1 : 0; return 0; }

// -------------------------------------------------------------------------------------------
// layout converters: [n][c][hw] <-> [n][hw][ld]   (32x32 smem tile transpose)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const void* __restrict__ x, int x_dtype, void* __restrict__ y, int y_dtype, int c, long long hw,
                    int y_ld) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int ch = c0 + k; const long long p = p0 + tx;
    tile[k][tx] = (ch < c && p < hw) ? load_any(x, ((long long)n * c + ch) * hw + p, x_dtype) : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const long long p = p0 + k; const int ch = c0 + tx;
    if (p < hw && ch < c) store_any(y, ((long long)n * hw + p) * y_ld + ch, tile[tx][k], y_dtype);
  }
}
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const void* __restrict__ x, int x_dtype, void* __restrict__ y, int y_dtype, int c, long long hw,
                    int x_ld, long long pitch) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int k = ty; k < 32; k += 8) {
    const long long p = p0 + k; const int ch = c0 + tx;
    tile[k][tx] = (p < hw && ch < c) ? load_any(x, ((long long)n * hw + p) * x_ld + ch, x_dtype) : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int ch = c0 + k; const long long p = p0 + tx;
    if (ch < c && p < hw) store_any(y, ((long long)n * c + ch) * pitch + p, tile[tx][k], y_dtype);
  }
}

}  // namespace segb200

using namespace segb200;

extern "C" int segb200_pack_s2d(const void* x, int x_dtype, void* out, int out_dtype, int n, int c, int h, int w,
                                int out_ld, void* stream) {
  if (!x || !out) return set_error(-1, "pack_s2d: null pointer");
  if (!half_dt(out_dtype) || x_dtype < 0 || x_dtype > 2) return set_error(-2, "pack_s2d: bad dtype");
  if ((out_ld & 7) || out_ld < 4 * c) return set_error(-4, "pack_s2d: out_ld must be a multiple of 8 and >= 4*c");
  const int hs = (h + 1) / 2, ws = (w + 1) / 2;
  const long long total = (long long)n * hs * ws * (out_ld / 8);
  if ((long long)n * hs <= 65535) {
    pack_s2d_rows_kernel<<<dim3((unsigned)((ws * (out_ld / 8) + 255) / 256), (unsigned)(n * hs)), 256, 0, STREAM(stream)>>>(
        x, x_dtype, out, out_dtype, c, h, w, hs, ws, out_ld);
    return check_launch("pack_s2d(rows)");
  }
  pack_s2d_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(x, x_dtype, out, out_dtype, n, c, h, w, hs, ws, out_ld);
  return check_launch("pack_s2d");
}

extern "C" int segb200_global_avgpool(const void* x, void* out, int n, int h, int w, int c, int x_ld, int dtype,
                                      void* stream) {
  if (!x || !out) return set_error(-1, "global_avgpool: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "global_avgpool: bad dtype");
  if ((c & 7) || (x_ld & 7)) return set_error(-4, "global_avgpool: c and x_ld must be multiples of 8");
  if (((uintptr_t)x & 15) || ((uintptr_t)out & 15)) return set_error(-7, "global_avgpool: pointers must be 16-byte aligned");
  if (h * w >= 2048) {
    gap_wide_kernel<<<dim3((c / 8 + 7) / 8, n), 1024, 0, STREAM(stream)>>>(x, out, h * w, c, x_ld, dtype);
    return check_launch("global_avgpool(wide)");
  }
  gap_kernel<<<dim3((c / 8 + 15) / 16, n), 256, 0, STREAM(stream)>>>(x, out, h * w, c, x_ld, dtype);
  return check_launch("global_avgpool");
}

extern "C" int segb200_adaptive_avgpool(const void* x, void* out, int n, int h, int w, int c, int x_ld, int s, int out_ld,
                                        int dtype, void* stream) {
  if (!x || !out) return set_error(-1, "adaptive_avgpool: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "adaptive_avgpool: bad dtype");
  if ((c & 7) || (x_ld & 7) || (out_ld & 7) || s < 1) return set_error(-4, "adaptive_avgpool: bad sizes");
  const long long total = (long long)n * s * s * (c / 8);
  adaptive_pool_kernel<<<grid_for(total, 128), 128, 0, STREAM(stream)>>>(x, out, n, h, w, c, x_ld, s, out_ld, dtype);
  return check_launch("adaptive_avgpool");
}

extern "C" int segb200_maxpool3x3s2(const void* x, void* y, int n, int h, int w, int c, int x_ld, int y_ld, int dtype,
                                   void* stream) {
  if (!x || !y) return set_error(-1, "maxpool3x3s2: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "maxpool3x3s2: bad dtype");
  if ((c & 7) || (x_ld & 7) || (y_ld & 7)) return set_error(-4, "maxpool3x3s2: c/pitches must be multiples of 8");
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const long long total = (long long)n * ho * wo * (c / 8);
  if (dtype == DT_BF16)
    maxpool3x3s2_kernel<true><<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(x, y, n, h, w, c, x_ld, ho, wo, y_ld);
  else
    maxpool3x3s2_kernel<false><<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(x, y, n, h, w, c, x_ld, ho, wo, y_ld);
  return check_launch("maxpool3x3s2");
}

extern "C" int segb200_upsample_add(const void* a, const void* z, void* y, int n, int h, int w, int c, int a_ld, int z_ld,
                                    int y_ld, int k, int act, int dtype, void* stream) {
  if (!a || !z || !y) return set_error(-1, "upsample_add: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "upsample_add: bad dtype");
  if ((c & 7) || (a_ld & 7) || (z_ld & 7) || (y_ld & 7) || k < 0 || k > 8 || (h & ((1 << k) - 1)) || (w & ((1 << k) - 1)))
    return set_error(-4, "upsample_add: sizes must be multiples of 8 channels and of the 2^k scale");
  const long long total = (long long)n * h * w * (c / 8);
  if (dtype == DT_BF16)
    upsample_add_kernel<true><<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(a, z, y, n, h, w, c, a_ld, z_ld, y_ld, k, act);
  else
    upsample_add_kernel<false><<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(a, z, y, n, h, w, c, a_ld, z_ld, y_ld, k, act);
  return check_launch("upsample_add");
}

extern "C" int segb200_bilinear_nhwc(const void* x, void* y, int n, int hi, int wi, int c, int x_ld, int ho, int wo,
                                     int y_ld, int align_corners, int dtype, void* stream) {
  if (!x || !y) return set_error(-1, "bilinear_nhwc: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "bilinear_nhwc: bad dtype");
  if ((c & 7) || (x_ld & 7) || (y_ld & 7)) return set_error(-4, "bilinear_nhwc: c/pitches must be multiples of 8");
  const long long total = (long long)n * ho * wo * (c / 8);
  constexpr int kRows = 4;
  const int groups = (ho + kRows - 1) / kRows;
  if (!g_bilinear_out_v1 && (long long)n * groups <= 65535 && (long long)wo * (c / 8) < 0x7fffffffLL) {
    const dim3 grid((unsigned)((wo * (c / 8) + 255) / 256), (unsigned)(n * groups));
    if (dtype == DT_BF16) bilinear_nhwc_strip_kernel<true, kRows><<<grid, 256, 0, STREAM(stream)>>>(x, y, hi, wi, c, x_ld, ho, wo, y_ld, align_corners, groups);
    else bilinear_nhwc_strip_kernel<false, kRows><<<grid, 256, 0, STREAM(stream)>>>(x, y, hi, wi, c, x_ld, ho, wo, y_ld, align_corners, groups);
    return check_launch("bilinear_nhwc(strip)");
  }
  bilinear_nhwc_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(x, y, n, hi, wi, c, x_ld, ho, wo, y_ld,
                                                                         align_corners, dtype);
  return check_launch("bilinear_nhwc");
}

extern "C" int segb200_bilinear_nchw_out(const void* x, void* y, uint8_t* argmax_out, int n, int hi, int wi, int c,
                                         int x_ld, int ho, int wo, int align_corners, int dtype, int out_dtype,
                                         void* stream) {
  if (!x || !y) return set_error(-1, "bilinear_nchw_out: null pointer");
  if (!half_dt(dtype) || out_dtype < 0 || out_dtype > 2) return set_error(-2, "bilinear_nchw_out: bad dtype");
  if ((x_ld & 7) || c < 1 || c > 256 || ((c + 7) & ~7) > x_ld) return set_error(-4, "bilinear_nchw_out: bad c/x_ld");
  const long long total = (long long)n * ho * wo;
  if (c <= 32 && !g_bilinear_out_v1) {
    constexpr int kRows = 4;
    const int groups = (ho + kRows - 1) / kRows;
    if ((long long)n * groups > 65535) return set_error(-6, "bilinear_nchw_out: too many row groups");
    const dim3 grid((unsigned)((wo + 127) / 128), (unsigned)(n * groups));
    const int cv = (c + 7) / 8;
    typedef void (*Fn)(const void*, void*, uint8_t*, int, int, int, int, int, int, int, int, int);
#define BL_OUT(CV, BF) {bilinear_nchw_strip_kernel<CV, BF, DT_BF16, kRows>, bilinear_nchw_strip_kernel<CV, BF, DT_F16, kRows>, bilinear_nchw_strip_kernel<CV, BF, DT_F32, kRows>}
    static const Fn fns[4][2][3] = {{BL_OUT(1, false), BL_OUT(1, true)}, {BL_OUT(2, false), BL_OUT(2, true)}, {BL_OUT(3, false), BL_OUT(3, true)},
                                    {BL_OUT(4, false), BL_OUT(4, true)}};
#undef BL_OUT
    fns[cv - 1][dtype == DT_BF16 ? 1 : 0][out_dtype]<<<grid, 128, 0, STREAM(stream)>>>(x, y, argmax_out, n, hi, wi, c, x_ld, ho, wo, align_corners, groups);
    return check_launch("bilinear_nchw_out(strip)");
  }
  const int g = grid_for(total, 256);
  if (c <= 32)
    bilinear_nchw_out_kernel<32><<<g, 256, 0, STREAM(stream)>>>(x, y, argmax_out, n, hi, wi, c, x_ld, ho, wo,
                                                                align_corners, dtype, out_dtype);
  else
    bilinear_nchw_out_kernel<256><<<g, 256, 0, STREAM(stream)>>>(x, y, argmax_out, n, hi, wi, c, x_ld, ho, wo,
                                                                 align_corners, dtype, out_dtype);
  return check_launch("bilinear_nchw_out");
}

extern "C" int segb200_nchw_to_nhwc(const void* x, int x_dtype, void* y, int y_dtype, int n, int c, int h, int w,
                                    int y_ld, void* stream) {
  if (!x || !y) return set_error(-1, "nchw_to_nhwc: null pointer");
  const long long hw = (long long)h * w;
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((c + 31) / 32), (unsigned)n);
  nchw_to_nhwc_kernel<<<grid, 256, 0, STREAM(stream)>>>(x, x_dtype, y, y_dtype, c, hw, y_ld);
  return check_launch("nchw_to_nhwc");
}
extern "C" int segb200_nhwc_to_nchw(const void* x, int x_dtype, void* y, int y_dtype, int n, int c, int h, int w,
                                    int x_ld, void* stream) {
  if (!x || !y) return set_error(-1, "nhwc_to_nchw: null pointer");
  const long long hw = (long long)h * w;
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((c + 31) / 32), (unsigned)n);
  nhwc_to_nchw_kernel<<<grid, 256, 0, STREAM(stream)>>>(x, x_dtype, y, y_dtype, c, hw, x_ld, hw);
  return check_launch("nhwc_to_nchw");
}
extern "C" int segb200_nhwc_to_cn(const void* x, void* y, int n, int c, int hw, int x_ld, int pitch, int dtype, void* stream) {
  if (!x || !y) return set_error(-1, "nhwc_to_cn: null pointer");
  if (!half_dt(dtype) || pitch < hw) return set_error(-4, "nhwc_to_cn: bad dtype/pitch");
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((c + 31) / 32), (unsigned)n);
  nhwc_to_nchw_kernel<<<grid, 256, 0, STREAM(stream)>>>(x, dtype, y, dtype, c, hw, x_ld, pitch);
  return check_launch("nhwc_to_cn");
}

// segb200 -- training-path memory-bound kernels (NHWC, 128-bit vectors, fp32 math):
//   train-mode BatchNorm (batch statistics, normalise + residual + activation, backward), column reductions,
//   max-pool / bilinear / depthwise-weight backward, the fused logits-upsample + cross-entropy loss,
//   weight packing (index-table gather), SGD with momentum.
// Every reduction is two-level and fixed-order (per-slab partial sums, then a per-channel finalize): bit-reproducible,
// no floating-point atomics in this file.
#include "common.cuh"
#include "vec.cuh"
#include "../../include/segb200.h"

#include <math.h>

namespace segb200 {

// ---------------------------------------------------------------------------------------------
// column-reduction geometry: a block is 256 threads = (256 >> cls) row lanes x (1 << cls) channel-vector lanes;
// grid.x covers the channel vectors, grid.y the row slabs.  partial[(slab * K + k) * c + ch].
// ---------------------------------------------------------------------------------------------
struct RedGeom { int cls, gx, slabs; long long rows_per_slab; };

static RedGeom red_geom(long long rows, int c, int max_slabs, int blocks_target = 148 * 4) {
  RedGeom g;
  const int cvn = c / 8;
  g.cls = 0;
  while ((1 << g.cls) < cvn && g.cls < 4) ++g.cls;
  const int cl = 1 << g.cls, pl = 256 >> g.cls;
  g.gx = (cvn + cl - 1) / cl;
  long long want = ((long long)blocks_target + g.gx - 1) / g.gx;  // ~4 (reductions) .. 8 (elementwise) blocks per SM in total
  const long long by_rows = (rows + (long long)pl * 4 - 1) / ((long long)pl * 4);   // at least 4 rows per thread
  if (want > by_rows) want = by_rows;
  if (want > max_slabs) want = max_slabs;
  if (want < 1) want = 1;
  g.rows_per_slab = ((rows + want - 1) / want + pl - 1) / pl * pl;
  g.slabs = (int)((rows + g.rows_per_slab - 1) / g.rows_per_slab);
  if (g.slabs < 1) g.slabs = 1;
  return g;
}

// reduce acc[8] over the row lanes of the block (fixed order) and store to dst[ch .. ch+8)
__device__ __forceinline__ void block_colsum_store(float (&acc)[8], float (*red)[8], int cls, int lc, int lp, bool valid,
                                                   float* dst) {
  const int cl = 1 << cls, pl = 256 >> cls;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) red[lp * cl + lc][j] = acc[j];
  __syncthreads();
  if (lp == 0 && valid) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = 0.f;
    for (int q = 0; q < pl; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] += red[q * cl + lc][j];
    *reinterpret_cast<float4*>(dst) = make_float4(t[0], t[1], t[2], t[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(t[4], t[5], t[6], t[7]);
  }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm forward statistics: per-slab sum and sum of squares per channel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_stats_kernel(const void* __restrict__ x, long long rows, int c, int ld, int dtype, int cls, long long rows_per_slab,
                float* __restrict__ partial) {
  __shared__ float red[256][8];
  const int cl = 1 << cls, pl = 256 >> cls;
  const int lc = threadIdx.x & (cl - 1), lp = threadIdx.x >> cls;
  const int cv = blockIdx.x * cl + lc;
  const bool valid = cv < c / 8;
  const long long r0 = (long long)blockIdx.y * rows_per_slab;
  long long r1 = r0 + rows_per_slab; if (r1 > rows) r1 = rows;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
  if (valid) {
    const char* base = reinterpret_cast<const char*>(x) + (long long)cv * 16;
    long long p = r0 + lp;
    for (; p + 3LL * pl < r1; p += 4LL * pl) {           // 4 independent 128-bit loads in flight
      float f0[8], f1[8], f2[8], f3[8];
      unpack8(ldg_nc_v4(base + p * ld * 2), dtype, f0);
      unpack8(ldg_nc_v4(base + (p + pl) * ld * 2), dtype, f1);
      unpack8(ldg_nc_v4(base + (p + 2LL * pl) * ld * 2), dtype, f2);
      unpack8(ldg_nc_v4(base + (p + 3LL * pl) * ld * 2), dtype, f3);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += (f0[j] + f1[j]) + (f2[j] + f3[j]);
        q[j] += (f0[j] * f0[j] + f1[j] * f1[j]) + (f2[j] * f2[j] + f3[j] * f3[j]);
      }
    }
    for (; p < r1; p += pl) {
      float f[8];
      unpack8(ldg_nc_v4(base + p * ld * 2), dtype, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] = fmaf(f[j], f[j], q[j]); }
    }
  }
  float* dst = partial + ((long long)blockIdx.y * 2) * c + cv * 8;
  block_colsum_store(s, red, cls, lc, lp, valid, dst);
  block_colsum_store(q, red, cls, lc, lp, valid, dst + c);
}

// out[k * sk + ch * sc] (+)= sum over slabs of partial[(slab * K + k) * c + ch]   (double accumulation, fixed order)
__global__ void __launch_bounds__(128)
reduce_partials_kernel(const float* __restrict__ partial, int slabs, int K, int c, float* __restrict__ out, long long sk,
                       long long sc, int accumulate, float scale) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * c) return;
  const int k = idx / c, ch = idx - k * c;
  double t = 0.0;
  for (int s = 0; s < slabs; ++s) t += (double)partial[((long long)s * K + k) * c + ch];
  float* o = out + k * sk + ch * sc;
  const float v = (float)t * scale;
  *o = accumulate ? *o + v : v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per channel: lanes stride over the slabs (fixed order), shuffle tree => deterministic
__global__ void __launch_bounds__(128)
bn_finalize_kernel(const float* __restrict__ partial, int slabs, int c, double count, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ running_mean, float* __restrict__ running_var,
                   float momentum, float eps, float* __restrict__ mean, float* __restrict__ invstd,
                   float* __restrict__ scale, float* __restrict__ shift) {
  const int ch = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (ch >= c) return;
  double s = 0.0, q = 0.0;
  for (int sl = lane; sl < slabs; sl += 32) {
    s += (double)partial[((long long)sl * 2) * c + ch];
    q += (double)partial[((long long)sl * 2 + 1) * c + ch];
  }
  s = warp_sum(s); q = warp_sum(q);
  if (lane != 0) return;
  const double m = s / count;
  double var = q / count - m * m;
  if (var < 0.0) var = 0.0;
  const float is = (float)(1.0 / sqrt(var + (double)eps));
  mean[ch] = (float)m;
  invstd[ch] = is;
  const float g = gamma != nullptr ? gamma[ch] : 1.f;
  const float sc = g * is;
  scale[ch] = sc;
  shift[ch] = (beta != nullptr ? beta[ch] : 0.f) - (float)m * sc;
  if (running_mean != nullptr) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
  if (running_var != nullptr) {
    const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unb;
  }
}

// ---------------------------------------------------------------------------------------------
// z = act(y * scale[c] + shift[c] + residual) * nc_scale[n][c]      (every optional operand may be NULL)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_apply_kernel(const void* __restrict__ y, int y_ld, const float* __restrict__ scale, const float* __restrict__ shift,
                const void* __restrict__ res, int res_ld, const float* __restrict__ nc_scale, long long rows_per_img,
                void* __restrict__ z, int z_ld, long long rows, int c, int act, int dtype, int cls, long long rows_per_slab) {
  const int cl = 1 << cls, pl = 256 >> cls;
  const int lc = threadIdx.x & (cl - 1), lp = threadIdx.x >> cls;
  const int cv = blockIdx.x * cl + lc;
  if (cv >= c / 8) return;
  const long long r0 = (long long)blockIdx.y * rows_per_slab;
  long long r1 = r0 + rows_per_slab; if (r1 > rows) r1 = rows;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = scale != nullptr ? scale[cv * 8 + j] : 1.f; sh[j] = shift != nullptr ? shift[cv * 8 + j] : 0.f; }
  const char* yb = reinterpret_cast<const char*>(y) + (long long)cv * 16;
  const char* rb = reinterpret_cast<const char*>(res) + (long long)cv * 16;
  char* zb = reinterpret_cast<char*>(z) + (long long)cv * 16;
  auto one = [&](long long p, const uint4& vy, const uint4& vr) {
    float f[8];
    unpack8(vy, dtype, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sh[j]);
    if (res != nullptr) {
      float r[8];
      unpack8(vr, dtype, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += r[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = apply_act(f[j], act);
    if (nc_scale != nullptr) {
      const float* m = nc_scale + (p / rows_per_img) * c + cv * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= __ldg(m + j);
    }
    *reinterpret_cast<uint4*>(zb + p * z_ld * 2) = pack8(f, dtype);
  };
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  long long p = r0 + lp;
  for (; p + pl < r1; p += 2LL * pl) {                 // two rows in flight per thread
    const uint4 a0 = ldg_nc_v4(yb + p * y_ld * 2), a1 = ldg_nc_v4(yb + (p + pl) * y_ld * 2);
    uint4 b0 = zero4, b1 = zero4;
    if (res != nullptr) { b0 = ldg_nc_v4(rb + p * res_ld * 2); b1 = ldg_nc_v4(rb + (p + pl) * res_ld * 2); }
    one(p, a0, b0); one(p + pl, a1, b1);
  }
  if (p < r1) {
    const uint4 a0 = ldg_nc_v4(yb + p * y_ld * 2);
    const uint4 b0 = res != nullptr ? ldg_nc_v4(rb + p * res_ld * 2) : zero4;
    one(p, a0, b0);
  }
}

// gradient through (optional channel mask, activation): g = dz * nc_scale * [act'(pre-activation)].
// The activation mask comes from the stored output z when it is given (needed when a residual entered the activation),
// else it is recomputed from the raw conv output: pre = y * scale + shift  (saves one activation read per pass).
__device__ __forceinline__ void act_grad8(float (&g)[8], const uint4& vz, bool have_z, const float (&yy)[8], const float (&sc)[8],
                                          const float (&sh)[8], int act, const float* nc_scale, long long nc_off, int dtype) {
  if (nc_scale != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= __ldg(nc_scale + nc_off + j);
  }
  if (act != ACT_NONE) {
    float zz[8];
    if (have_z) unpack8(vz, dtype, zz);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) zz[j] = fmaf(yy[j], sc[j], sh[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool on = act == ACT_RELU6 ? (zz[j] > 0.f && zz[j] < 6.f) : (zz[j] > 0.f);
      g[j] = on ? g[j] : 0.f;
    }
  }
}

// BatchNorm backward, pass 1: per-slab sum(g) and sum(g * y) per channel (the finalize turns the second into
// sum(g * xhat) = invstd * (sum(g*y) - mean * sum(g)); keeping mean/invstd out of the loop saves 16 registers)
__global__ void __launch_bounds__(256, 3)
bn_bwd_reduce_kernel(const void* __restrict__ dz, int dz_ld, const void* __restrict__ z, int z_ld,
                     const void* __restrict__ y, int y_ld, const float* __restrict__ scale, const float* __restrict__ shift,
                     const float* __restrict__ nc_scale, long long rows_per_img, long long rows, int c, int act, int dtype,
                     int cls, long long rows_per_slab, float* __restrict__ partial) {
  __shared__ float red[256][8];
  const int cl = 1 << cls, pl = 256 >> cls;
  const int lc = threadIdx.x & (cl - 1), lp = threadIdx.x >> cls;
  const int cv = blockIdx.x * cl + lc;
  const bool valid = cv < c / 8;
  const long long r0 = (long long)blockIdx.y * rows_per_slab;
  long long r1 = r0 + rows_per_slab; if (r1 > rows) r1 = rows;
  float s1[8], s2[8], sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; sc[j] = 1.f; sh[j] = 0.f; }
  const bool have_z = z != nullptr;
  if (valid) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = scale != nullptr ? scale[cv * 8 + j] : 1.f; sh[j] = shift != nullptr ? shift[cv * 8 + j] : 0.f; }
    const char* db = reinterpret_cast<const char*>(dz) + (long long)cv * 16;
    const char* yb = reinterpret_cast<const char*>(y) + (long long)cv * 16;
    const char* zb = reinterpret_cast<const char*>(z) + (long long)cv * 16;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    auto one = [&](long long p, const uint4& vd, const uint4& vy, const uint4& vz) {
      float g[8], yy[8];
      unpack8(vd, dtype, g);
      unpack8(vy, dtype, yy);
      act_grad8(g, vz, have_z, yy, sc, sh, act, nc_scale, nc_scale != nullptr ? (p / rows_per_img) * c + cv * 8 : 0, dtype);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s1[j] += g[j]; s2[j] = fmaf(g[j], yy[j], s2[j]); }
    };
    long long p = r0 + lp;
    for (; p + pl < r1; p += 2LL * pl) {
      const uint4 d0 = ldg_nc_v4(db + p * dz_ld * 2), d1 = ldg_nc_v4(db + (p + pl) * dz_ld * 2);
      const uint4 y0 = ldg_nc_v4(yb + p * y_ld * 2), y1 = ldg_nc_v4(yb + (p + pl) * y_ld * 2);
      uint4 z0 = zero4, z1 = zero4;
      if (have_z && act != ACT_NONE) { z0 = ldg_nc_v4(zb + p * z_ld * 2); z1 = ldg_nc_v4(zb + (p + pl) * z_ld * 2); }
      one(p, d0, y0, z0); one(p + pl, d1, y1, z1);
    }
    if (p < r1) {
      const uint4 d0 = ldg_nc_v4(db + p * dz_ld * 2), y0 = ldg_nc_v4(yb + p * y_ld * 2);
      const uint4 z0 = (have_z && act != ACT_NONE) ? ldg_nc_v4(zb + p * z_ld * 2) : zero4;
      one(p, d0, y0, z0);
    }
  }
  float* dst = partial + ((long long)blockIdx.y * 2) * c + cv * 8;
  block_colsum_store(s1, red, cls, lc, lp, valid, dst);
  block_colsum_store(s2, red, cls, lc, lp, valid, dst + c);
}

// sums[0][c] = sum g, sums[1][c] = sum g*xhat = invstd * (sum g*y - mean * sum g); dgamma += sums[1], dbeta += sums[0]
// (one warp per channel)
__global__ void __launch_bounds__(128)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int slabs, int c, const float* __restrict__ mean,
                       const float* __restrict__ invstd, float* __restrict__ sums, float* __restrict__ dgamma,
                       float* __restrict__ dbeta) {
  const int ch = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (ch >= c) return;
  double a = 0.0, b = 0.0;
  for (int sl = lane; sl < slabs; sl += 32) {
    a += (double)partial[((long long)sl * 2) * c + ch];
    b += (double)partial[((long long)sl * 2 + 1) * c + ch];
  }
  a = warp_sum(a); b = warp_sum(b);
  if (lane != 0) return;
  b = (invstd != nullptr ? (double)invstd[ch] : 1.0) * (b - (mean != nullptr ? (double)mean[ch] : 0.0) * a);
  sums[ch] = (float)a;
  sums[c + ch] = (float)b;
  if (dgamma != nullptr) dgamma[ch] += (float)b;
  if (dbeta != nullptr) dbeta[ch] += (float)a;
}

// BatchNorm backward, pass 2: dy = scale * (g - sum_g/count - xhat * sum_gx/count) = A*g + B*y + C per channel; dres (+)= g
__global__ void __launch_bounds__(256, 3)
bn_bwd_apply_kernel(const void* __restrict__ dz, int dz_ld, const void* __restrict__ z, int z_ld,
                    const void* __restrict__ y, int y_ld, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ scale, const float* __restrict__ shift,
                    const float* __restrict__ sums, float inv_count, const float* __restrict__ nc_scale,
                    long long rows_per_img, void* __restrict__ dy, int dy_ld, void* __restrict__ dres, int dres_ld,
                    int dres_accumulate, long long rows, int c, int act, int dtype, int cls, long long rows_per_slab) {
  const int cl = 1 << cls, pl = 256 >> cls;
  const int lc = threadIdx.x & (cl - 1), lp = threadIdx.x >> cls;
  const int cv = blockIdx.x * cl + lc;
  if (cv >= c / 8) return;
  const long long r0 = (long long)blockIdx.y * rows_per_slab;
  long long r1 = r0 + rows_per_slab; if (r1 > rows) r1 = rows;
  float A[8], B[8], Cc[8], sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = cv * 8 + j;
    sc[j] = scale != nullptr ? scale[ch] : 1.f;
    sh[j] = shift != nullptr ? shift[ch] : 0.f;
    if (sums != nullptr) {
      const float is = invstd[ch], mu = mean[ch], k2 = sc[j] * sums[c + ch] * inv_count * is;
      A[j] = sc[j]; B[j] = -k2; Cc[j] = fmaf(k2, mu, -sc[j] * sums[ch] * inv_count);
    } else {
      A[j] = sc[j]; B[j] = 0.f; Cc[j] = 0.f;
    }
  }
  const bool have_z = z != nullptr;
  const bool need_y = sums != nullptr || (!have_z && act != ACT_NONE);
  const char* db = reinterpret_cast<const char*>(dz) + (long long)cv * 16;
  const char* yb = reinterpret_cast<const char*>(y) + (long long)cv * 16;
  const char* zb = reinterpret_cast<const char*>(z) + (long long)cv * 16;
  char* ob = reinterpret_cast<char*>(dy) + (long long)cv * 16;
  char* rb = reinterpret_cast<char*>(dres) + (long long)cv * 16;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  auto one = [&](long long p, const uint4& vd, const uint4& vy, const uint4& vz) {
    float g[8], yy[8];
    unpack8(vd, dtype, g);
    unpack8(vy, dtype, yy);
    act_grad8(g, vz, have_z, yy, sc, sh, act, nc_scale, nc_scale != nullptr ? (p / rows_per_img) * c + cv * 8 : 0, dtype);
    if (dres != nullptr) {
      char* rp = rb + p * dres_ld * 2;
      float o[8];
      if (dres_accumulate) {
        unpack8(*reinterpret_cast<const uint4*>(rp), dtype, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += g[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = g[j];
      }
      *reinterpret_cast<uint4*>(rp) = pack8(o, dtype);
    }
    if (dy != nullptr) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(A[j], g[j], fmaf(B[j], yy[j], Cc[j]));
      *reinterpret_cast<uint4*>(ob + p * dy_ld * 2) = pack8(o, dtype);
    }
  };
  long long p = r0 + lp;
  for (; p + pl < r1; p += 2LL * pl) {
    const uint4 d0 = ldg_nc_v4(db + p * dz_ld * 2), d1 = ldg_nc_v4(db + (p + pl) * dz_ld * 2);
    uint4 y0 = zero4, y1 = zero4, z0 = zero4, z1 = zero4;
    if (need_y) { y0 = ldg_nc_v4(yb + p * y_ld * 2); y1 = ldg_nc_v4(yb + (p + pl) * y_ld * 2); }
    if (have_z && act != ACT_NONE) { z0 = ldg_nc_v4(zb + p * z_ld * 2); z1 = ldg_nc_v4(zb + (p + pl) * z_ld * 2); }
    one(p, d0, y0, z0); one(p + pl, d1, y1, z1);
  }
  if (p < r1) {
    const uint4 d0 = ldg_nc_v4(db + p * dz_ld * 2);
    const uint4 y0 = need_y ? ldg_nc_v4(yb + p * y_ld * 2) : zero4;
    const uint4 z0 = (have_z && act != ACT_NONE) ? ldg_nc_v4(zb + p * z_ld * 2) : zero4;
    one(p, d0, y0, z0);
  }
}

// ---------------------------------------------------------------------------------------------
// MaxPool2d(3, 2, 1) backward (gather): dx[p] = sum over the windows whose FIRST maximum (row-major scan, strict >,
// torch's rule) is p of dy[window]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
maxpool3x3s2_bwd_kernel(const void* __restrict__ x, const void* __restrict__ dy, void* __restrict__ dx, int n, int h, int w,
                        int c, int x_ld, int ho, int wo, int dy_ld, int dx_ld, int dtype) {
  const int cvn = c / 8;
  const long long total = (long long)n * h * w * cvn;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvn);
    long long r = idx / cvn;
    const int ix = (int)(r % w); r /= w;
    const int iy = (int)(r % h);
    const int b = (int)(r / h);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const char* xb = reinterpret_cast<const char*>(x) + ((long long)b * h * w * x_ld + cv * 8) * 2;
    const int oy0 = iy / 2, oy1 = (iy + 1) / 2;          // windows with 2*oy-1 <= iy <= 2*oy+1
    const int ox0 = ix / 2, ox1 = (ix + 1) / 2;
    for (int oy = oy0; oy <= oy1; ++oy) {
      if (oy >= ho) continue;
      for (int ox = ox0; ox <= ox1; ++ox) {
        if (ox >= wo) continue;
        float m[8]; int pos[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { m[j] = -INFINITY; pos[j] = -1; }
        for (int ky = 0; ky < 3; ++ky) {
          const int yy = oy * 2 - 1 + ky;
          if (yy < 0 || yy >= h) continue;
          for (int kx = 0; kx < 3; ++kx) {
            const int xx = ox * 2 - 1 + kx;
            if (xx < 0 || xx >= w) continue;
            float f[8];
            unpack8(ldg_v4(xb + ((long long)yy * w + xx) * x_ld * 2), dtype, f);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (f[j] > m[j] || pos[j] < 0) { m[j] = f[j]; pos[j] = yy * w + xx; }
          }
        }
        float g[8];
        unpack8(ldg_v4(reinterpret_cast<const char*>(dy) + ((((long long)b * ho + oy) * wo + ox) * dy_ld + cv * 8) * 2), dtype, g);
        const int me = iy * w + ix;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += pos[j] == me ? g[j] : 0.f;
      }
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<char*>(dx) + ((((long long)b * h + iy) * w + ix) * dx_ld + cv * 8) * 2) = pack8(acc, dtype);
  }
}

// ---------------------------------------------------------------------------------------------
// MaxPool2d(3, 2, 1) forward that also records the argmax tap (0..8, first maximum in row-major scan order = torch's rule),
// and the backward that uses it: 4 x (8 B index + 16 B gradient) per input vector instead of re-scanning 4 x 9 inputs.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
maxpool3x3s2_idx_kernel(const void* __restrict__ x, void* __restrict__ y, uint8_t* __restrict__ idx, int n, int h, int w, int c,
                        int x_ld, int ho, int wo, int y_ld, int dtype) {
  const int cvn = c / 8;
  const long long total = (long long)n * ho * wo * cvn;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvn);
    long long r = i / cvn;
    const int ox = (int)(r % wo); r /= wo;
    const int oy = (int)(r % ho);
    const long long b = r / ho;
    float m[8]; int pos[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { m[j] = -INFINITY; pos[j] = -1; }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - 1 + ky;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - 1 + kx;
        if (ix < 0 || ix >= w) continue;
        float f[8];
        unpack8(ldg_v4(reinterpret_cast<const char*>(x) + (((b * h + iy) * w + ix) * x_ld + cv * 8) * 2), dtype, f);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (f[j] > m[j] || pos[j] < 0) { m[j] = f[j]; pos[j] = ky * 3 + kx; }
      }
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<char*>(y) + (((b * ho + oy) * wo + ox) * y_ld + cv * 8) * 2) = pack8(m, dtype);
    uint2 pk;
    pk.x = (uint32_t)pos[0] | ((uint32_t)pos[1] << 8) | ((uint32_t)pos[2] << 16) | ((uint32_t)pos[3] << 24);
    pk.y = (uint32_t)pos[4] | ((uint32_t)pos[5] << 8) | ((uint32_t)pos[6] << 16) | ((uint32_t)pos[7] << 24);
    *reinterpret_cast<uint2*>(idx + ((b * ho + oy) * wo + ox) * c + cv * 8) = pk;
  }
}

__global__ void __launch_bounds__(256)
maxpool3x3s2_bwd_idx_kernel(const uint8_t* __restrict__ idx, const void* __restrict__ dy, void* __restrict__ dx, int n, int h,
                            int w, int c, int ho, int wo, int dy_ld, int dx_ld, int dtype) {
  const int cvn = c / 8;
  const long long total = (long long)n * h * w * cvn;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvn);
    long long r = i / cvn;
    const int ix = (int)(r % w); r /= w;
    const int iy = (int)(r % h);
    const long long b = r / h;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int oy0 = iy / 2, oy1 = (iy + 1) / 2, ox0 = ix / 2, ox1 = (ix + 1) / 2;
    for (int oy = oy0; oy <= oy1; ++oy) {
      if (oy >= ho) continue;
      for (int ox = ox0; ox <= ox1; ++ox) {
        if (ox >= wo) continue;
        const uint32_t me = (uint32_t)((iy - (oy * 2 - 1)) * 3 + (ix - (ox * 2 - 1)));
        const long long o = (b * ho + oy) * wo + ox;
        const uint2 pk = __ldg(reinterpret_cast<const uint2*>(idx + o * c + cv * 8));
        float g[8];
        unpack8(ldg_v4(reinterpret_cast<const char*>(dy) + (o * dy_ld + cv * 8) * 2), dtype, g);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[j] += ((pk.x >> (8 * j)) & 0xffu) == me ? g[j] : 0.f;
          acc[4 + j] += ((pk.y >> (8 * j)) & 0xffu) == me ? g[4 + j] : 0.f;
        }
      }
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<char*>(dx) + (((b * h + iy) * w + ix) * dx_ld + cv * 8) * 2) = pack8(acc, dtype);
  }
}

// ---------------------------------------------------------------------------------------------
// bilinear resize backward (gather): dx[b][i][j] (+)= gscale * sum over destination pixels of weight * dy
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dst_range(int i, int in, int out, int align, int* lo, int* hi) {
  // destination indices whose source coordinate can fall within (i-1, i+1); conservative, the exact test follows
  float center, half;
  if (align) {
    const float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    if (scale <= 0.f) { *lo = 0; *hi = out - 1; return; }
    center = (float)i / scale; half = 1.f / scale;
  } else {
    const float scale = (float)in / (float)out;
    center = ((float)i + 0.5f) / scale - 0.5f; half = 1.f / scale;
  }
  int l = (int)floorf(center - half) - 1, u = (int)ceilf(center + half) + 1;
  if (l < 0) l = 0;
  if (u > out - 1) u = out - 1;
  *lo = l; *hi = u;
}

__global__ void __launch_bounds__(256)
bilinear_bwd_kernel(const void* __restrict__ dy, void* __restrict__ dx, int n, int hi, int wi, int c, int dx_ld, int ho,
                    int wo, int dy_ld, int align, int accumulate, const float* __restrict__ gscale, int dtype) {
  const int cvn = c / 8;
  const long long total = (long long)n * hi * wi * cvn;
  const float gs = gscale != nullptr ? *gscale : 1.f;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvn);
    long long r = idx / cvn;
    const int j = (int)(r % wi); r /= wi;
    const int i = (int)(r % hi);
    const int b = (int)(r / hi);
    int rlo, rhi, clo, chi;
    dst_range(i, hi, ho, align, &rlo, &rhi);
    dst_range(j, wi, wo, align, &clo, &chi);
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    const char* base = reinterpret_cast<const char*>(dy) + ((long long)b * ho * wo * dy_ld + cv * 8) * 2;
    for (int rr = rlo; rr <= rhi; ++rr) {
      const Lerp ly = lerp_coord(rr, hi, ho, align);
      const float wy = (ly.i0 == i ? ly.l0 : 0.f) + (ly.i1 == i ? ly.l1 : 0.f);
      if (wy == 0.f) continue;
      for (int cc = clo; cc <= chi; ++cc) {
        const Lerp lx = lerp_coord(cc, wi, wo, align);
        const float wx = (lx.i0 == j ? lx.l0 : 0.f) + (lx.i1 == j ? lx.l1 : 0.f);
        if (wx == 0.f) continue;
        float f[8];
        unpack8(ldg_v4(base + ((long long)rr * wo + cc) * dy_ld * 2), dtype, f);
        const float wgt = wy * wx;
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = fmaf(wgt, f[q], acc[q]);
      }
    }
    char* op = reinterpret_cast<char*>(dx) + ((((long long)b * hi + i) * wi + j) * dx_ld + cv * 8) * 2;
    float o[8];
    if (accumulate) {
      unpack8(*reinterpret_cast<const uint4*>(op), dtype, o);
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = fmaf(gs, acc[q], o[q]);
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = gs * acc[q];
    }
    *reinterpret_cast<uint4*>(op) = pack8(o, dtype);
  }
}

// ---------------------------------------------------------------------------------------------
// fused logits upsample (bilinear) + softmax cross-entropy (ignore_index), forward AND gradient w.r.t. the up-sampled
// logits in one pass: dfull[b][oy][ox][k] = softmax_k - [k == target]  (zero at ignored pixels; NOT yet divided by the
// number of valid pixels -- that factor is applied by the bilinear backward through `gscale`).
// per-block partial (loss sum, valid count) -> ce_finalize.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxClasses = 32;

__global__ void __launch_bounds__(256)
upsample_ce_kernel(const void* __restrict__ logits, const long long* __restrict__ target, void* __restrict__ dfull,
                   float* __restrict__ partial, int n, int hi, int wi, int nclass, int x_ld, int ho, int wo, int d_ld,
                   int align, int ignore_index, int dtype) {
  __shared__ float sred[2][8];
  const long long total = (long long)n * ho * wo;
  float lsum = 0.f, lcnt = 0.f;
  const int nv = (nclass + 7) / 8;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % wo);
    long long r = idx / wo;
    const int oy = (int)(r % ho);
    const int b = (int)(r / ho);
    const Lerp ly = lerp_coord(oy, hi, ho, align), lx = lerp_coord(ox, wi, wo, align);
    const char* base = reinterpret_cast<const char*>(logits) + (long long)b * hi * wi * x_ld * 2;
    const char* p00 = base + ((long long)ly.i0 * wi + lx.i0) * x_ld * 2;
    const char* p01 = base + ((long long)ly.i0 * wi + lx.i1) * x_ld * 2;
    const char* p10 = base + ((long long)ly.i1 * wi + lx.i0) * x_ld * 2;
    const char* p11 = base + ((long long)ly.i1 * wi + lx.i1) * x_ld * 2;
    float v[kMaxClasses];
    float mx = -INFINITY;
#pragma unroll
    for (int cv = 0; cv < kMaxClasses / 8; ++cv) {
      if (cv < nv) {
        float f00[8], f01[8], f10[8], f11[8];
        unpack8(ldg_v4(p00 + cv * 16), dtype, f00);
        unpack8(ldg_v4(p01 + cv * 16), dtype, f01);
        unpack8(ldg_v4(p10 + cv * 16), dtype, f10);
        unpack8(ldg_v4(p11 + cv * 16), dtype, f11);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = cv * 8 + j;
          const float o = ly.l0 * (lx.l0 * f00[j] + lx.l1 * f01[j]) + ly.l1 * (lx.l0 * f10[j] + lx.l1 * f11[j]);
          v[k] = k < nclass ? o : -INFINITY;
          mx = fmaxf(mx, v[k]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[cv * 8 + j] = -INFINITY;
      }
    }
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k) { v[k] = k < nclass ? __expf(v[k] - mx) : 0.f; se += v[k]; }
    const long long t = target[idx];
    const bool ok = t != (long long)ignore_index && t >= 0 && t < nclass;
    const float inv = 1.f / se;
    float pt = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k) {
      const float p = v[k] * inv;
      if ((long long)k == t) pt = p;
      v[k] = ok ? (p - ((long long)k == t ? 1.f : 0.f)) : 0.f;
    }
    if (ok) { lsum += -__logf(fmaxf(pt, 1e-37f)); lcnt += 1.f; }
    char* dp = reinterpret_cast<char*>(dfull) + idx * d_ld * 2;
#pragma unroll
    for (int cv = 0; cv < kMaxClasses / 8; ++cv) {
      if (cv * 8 < d_ld && cv < nv) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = v[cv * 8 + j];
        *reinterpret_cast<uint4*>(dp + cv * 16) = pack8(o, dtype);
      }
    }
  }
  // block reduction (fixed order)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { lsum += __shfl_xor_sync(0xffffffffu, lsum, o); lcnt += __shfl_xor_sync(0xffffffffu, lcnt, o); }
  if ((threadIdx.x & 31) == 0) { sred[0][threadIdx.x >> 5] = lsum; sred[1][threadIdx.x >> 5] = lcnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c2 = 0.f;
    for (int i = 0; i < 8; ++i) { a += sred[0][i]; c2 += sred[1][i]; }
    partial[blockIdx.x * 2] = a; partial[blockIdx.x * 2 + 1] = c2;
  }
}

// out[0] = mean loss over valid pixels, out[1] = 1 / valid count (0 if none), out[2] = valid count
__global__ void ce_finalize_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double a = 0.0, c = 0.0;
  for (int i = 0; i < nblocks; ++i) { a += (double)partial[2 * i]; c += (double)partial[2 * i + 1]; }
  out[0] = c > 0.0 ? (float)(a / c) : 0.f;
  out[1] = c > 0.0 ? (float)(1.0 / c) : 0.f;
  out[2] = (float)c;
}

// ---------------------------------------------------------------------------------------------
// depthwise 3x3 weight gradient: partial[(slab*9 + tap)*c + ch] = sum_p dy[p][ch] * x'[p + off(tap)][ch]
// (x' = relu(x) when the forward had a leading ReLU); stride 1, padding = dilation
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dw_wgrad_kernel(const void* __restrict__ x, const void* __restrict__ dy, int n, int h, int w, int c, int x_ld, int dy_ld,
                int dilation, int pre_relu, int dtype, int cls, long long rows_per_slab, float* __restrict__ partial) {
  __shared__ float red[256][8];
  const int cl = 1 << cls, pl = 256 >> cls;
  const int lc = threadIdx.x & (cl - 1), lp = threadIdx.x >> cls;
  const int cv = blockIdx.x * cl + lc;
  const bool valid = cv < c / 8;
  const long long rows = (long long)n * h * w;
  const long long r0 = (long long)blockIdx.y * rows_per_slab;
  long long r1 = r0 + rows_per_slab; if (r1 > rows) r1 = rows;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  if (valid) {
    const char* xb = reinterpret_cast<const char*>(x) + (long long)cv * 16;
    const char* gb = reinterpret_cast<const char*>(dy) + (long long)cv * 16;
    // no per-pixel divisions / 64-bit multiplies: the (column, row) of the running pixel is tracked incrementally and every
    // tap address is the pixel's own address plus a precomputed offset
    const long long row_off = (long long)dilation * w * x_ld * 2, col_off = (long long)dilation * x_ld * 2;
    auto pixel = [&](long long p, int xw, int yh, const uint4& vg) {
      float g[8];
      unpack8(vg, dtype, g);
      const char* xp = xb + p * x_ld * 2;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = yh + (ky - 1) * dilation;
        if (iy < 0 || iy >= h) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = xw + (kx - 1) * dilation;
          if (ix < 0 || ix >= w) continue;
          float f[8];
          unpack8(ldg_v4(xp + (ky - 1) * row_off + (kx - 1) * col_off), dtype, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xv = pre_relu ? fmaxf(f[j], 0.f) : f[j];
            acc[ky * 3 + kx][j] = fmaf(g[j], xv, acc[ky * 3 + kx][j]);
          }
        }
      }
    };
    auto advance = [&](int& xw, int& yh, int step) {          // (xw, yh) of pixel p -> of pixel p + step (the image index is not needed)
      xw += step;
      while (xw >= w) { xw -= w; if (++yh == h) yh = 0; }
    };
    long long p = r0 + lp;
    int xw = (int)(p % w), yh = (int)((p / w) % h);
    for (; p + pl < r1; p += 2LL * pl) {               // two pixels in flight
      const uint4 g0 = ldg_nc_v4(gb + p * dy_ld * 2), g1 = ldg_nc_v4(gb + (p + pl) * dy_ld * 2);
      int xw1 = xw, yh1 = yh;
      advance(xw1, yh1, pl);
      pixel(p, xw, yh, g0); pixel(p + pl, xw1, yh1, g1);
      xw = xw1; yh = yh1;
      advance(xw, yh, pl);
    }
    if (p < r1) pixel(p, xw, yh, ldg_nc_v4(gb + p * dy_ld * 2));
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
    block_colsum_store(acc[t], red, cls, lc, lp, valid, partial + ((long long)blockIdx.y * 9 + t) * c + cv * 8);
}

// ---------------------------------------------------------------------------------------------
// small elementwise helpers
// ---------------------------------------------------------------------------------------------
// y[b][p][:] (+)= v[b][:] * scale      (gradient of a global average pool; broadcast of a pooled feature)
__global__ void __launch_bounds__(256)
nc_broadcast_kernel(const void* __restrict__ v, void* __restrict__ y, int n, long long hw, int c, int v_ld, int y_ld,
                    float scale, int accumulate, int dtype) {
  const int cvn = c / 8;
  const long long total = (long long)n * hw * cvn;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvn);
    const long long row = idx / cvn;
    const long long b = row / hw;
    float f[8], o[8];
    unpack8(ldg_v4(reinterpret_cast<const char*>(v) + (b * v_ld + cv * 8) * 2), dtype, f);
    char* op = reinterpret_cast<char*>(y) + (row * y_ld + cv * 8) * 2;
    if (accumulate) {
      unpack8(*reinterpret_cast<const uint4*>(op), dtype, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(scale, f[j], o[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = scale * f[j];
    }
    *reinterpret_cast<uint4*>(op) = pack8(o, dtype);
  }
}

// stride-2 placement: mode 0: z[b][y][x] = (y, x even and inside) ? t[b][y/2][x/2] : 0   (zero insertion, writes all of z)
//                     mode 1: z[b][2i][2j] += t[b][i][j]                                   (strided accumulate)
__global__ void __launch_bounds__(256)
stride2_place_kernel(const void* __restrict__ t, void* __restrict__ z, int n, int h, int w, int c, int ht, int wt, int t_ld,
                     int z_ld, int mode, int dtype) {
  const int cvn = c / 8;
  if (mode == 0) {
    const long long total = (long long)n * h * w * cvn;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
      const int cv = (int)(idx % cvn);
      long long r = idx / cvn;
      const int x = (int)(r % w); r /= w;
      const int y = (int)(r % h);
      const long long b = r / h;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (!(y & 1) && !(x & 1) && (y >> 1) < ht && (x >> 1) < wt)
        v = ldg_nc_v4(reinterpret_cast<const char*>(t) + (((b * ht + (y >> 1)) * wt + (x >> 1)) * t_ld + cv * 8) * 2);
      *reinterpret_cast<uint4*>(reinterpret_cast<char*>(z) + (((b * h + y) * w + x) * z_ld + cv * 8) * 2) = v;
    }
  } else {
    const long long total = (long long)n * ht * wt * cvn;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
      const int cv = (int)(idx % cvn);
      long long r = idx / cvn;
      const int j = (int)(r % wt); r /= wt;
      const int i = (int)(r % ht);
      const long long b = r / ht;
      if (2 * i >= h || 2 * j >= w) continue;
      float f[8], o[8];
      unpack8(ldg_nc_v4(reinterpret_cast<const char*>(t) + (((b * ht + i) * wt + j) * t_ld + cv * 8) * 2), dtype, f);
      char* op = reinterpret_cast<char*>(z) + (((b * h + 2 * i) * w + 2 * j) * z_ld + cv * 8) * 2;
      unpack8(*reinterpret_cast<const uint4*>(op), dtype, o);
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] += f[q];
      *reinterpret_cast<uint4*>(op) = pack8(o, dtype);
    }
  }
}

// dst[i] = idx[i] >= 0 ? cast(src[idx[i]]) : 0     (weight packing: fp32 master -> packed 16-bit / fp32 operands)
__global__ void __launch_bounds__(256)
gather_cast_kernel(const float* __restrict__ src, const int* __restrict__ index, void* __restrict__ dst, long long n,
                   int dst_dtype) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = __ldg(index + i);
    store_any(dst, i, k >= 0 ? __ldg(src + k) : 0.f, dst_dtype);
  }
}

// dst[idx[i]] += src[i] for idx[i] >= 0 (index table is injective): un-packs a packed-layout gradient into the master layout
__global__ void __launch_bounds__(256)
scatter_add_kernel(const float* __restrict__ src, const int* __restrict__ index, float* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = __ldg(index + i);
    if (k >= 0) dst[k] += src[i];
  }
}

// torch.optim.SGD(momentum, dampening 0, weight_decay): g' = g + wd p; m = mu m + g'; p -= lr m   (m starts at zero)
__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, long long n, float lr, float mu,
           float wd, float gscale) {
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    mm.x = fmaf(mu, mm.x, fmaf(wd, pp.x, gg.x * gscale)); pp.x = fmaf(-lr, mm.x, pp.x);
    mm.y = fmaf(mu, mm.y, fmaf(wd, pp.y, gg.y * gscale)); pp.y = fmaf(-lr, mm.y, pp.y);
    mm.z = fmaf(mu, mm.z, fmaf(wd, pp.z, gg.z * gscale)); pp.z = fmaf(-lr, mm.z, pp.z);
    mm.w = fmaf(mu, mm.w, fmaf(wd, pp.w, gg.w * gscale)); pp.w = fmaf(-lr, mm.w, pp.w);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4 * 4)) {
    const long long i = n4 * 4 + threadIdx.x;
    const float mm = fmaf(mu, m[i], fmaf(wd, p[i], g[i] * gscale));
    m[i] = mm; p[i] = fmaf(-lr, mm, p[i]);
  }
}

}  // namespace segb200

using namespace segb200;

static inline bool vec_ok(int c, int ld) { return c > 0 && !(c & 7) && !(ld & 7) && ld >= c; }

extern "C" int segb200_reduce_slabs(long long rows, int c, int max_slabs) {
  if (rows < 1 || c < 8) return 1;
  return red_geom(rows, c, max_slabs > 0 ? max_slabs : 1 << 20).slabs;
}

extern "C" int segb200_bn_stats(const void* x, long long rows, int c, int x_ld, int dtype, float* partial, int max_slabs,
                                void* stream) {
  if (!x || !partial) return set_error(-1, "bn_stats: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "bn_stats: bad dtype");
  if (!vec_ok(c, x_ld) || rows < 1) return set_error(-4, "bn_stats: c/x_ld must be multiples of 8 and rows >= 1");
  const RedGeom g = red_geom(rows, c, max_slabs > 0 ? max_slabs : 1 << 20);
  bn_stats_kernel<<<dim3(g.gx, g.slabs), 256, 0, STREAM(stream)>>>(x, rows, c, x_ld, dtype, g.cls, g.rows_per_slab, partial);
  return check_launch("bn_stats");
}

extern "C" int segb200_reduce_partials(const float* partial, int slabs, int k, int c, float* out, long long stride_k,
                                       long long stride_c, int accumulate, float scale, void* stream) {
  if (!partial || !out) return set_error(-1, "reduce_partials: null pointer");
  if (slabs < 1 || k < 1 || c < 1) return set_error(-4, "reduce_partials: bad sizes");
  reduce_partials_kernel<<<(k * c + 127) / 128, 128, 0, STREAM(stream)>>>(partial, slabs, k, c, out, stride_k, stride_c,
                                                                        accumulate, scale);
  return check_launch("reduce_partials");
}

extern "C" int segb200_bn_finalize(const float* partial, int slabs, int c, double count, const float* gamma,
                                   const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                                   float* mean, float* invstd, float* scale, float* shift, void* stream) {
  if (!partial || !mean || !invstd || !scale || !shift) return set_error(-1, "bn_finalize: null pointer");
  if (slabs < 1 || c < 1 || count < 1.0) return set_error(-4, "bn_finalize: bad sizes");
  bn_finalize_kernel<<<(c + 3) / 4, 128, 0, STREAM(stream)>>>(partial, slabs, c, count, gamma, beta, running_mean,
                                                                 running_var, momentum, eps, mean, invstd, scale, shift);
  return check_launch("bn_finalize");
}

extern "C" int segb200_bn_apply(const void* y, const float* scale, const float* shift, const void* residual,
                                const float* nc_scale, void* z, long long rows, long long rows_per_img, int c, int y_ld,
                                int res_ld, int z_ld, int act, int dtype, void* stream) {
  if (!y || !z) return set_error(-1, "bn_apply: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "bn_apply: bad dtype");
  if (!vec_ok(c, y_ld) || !vec_ok(c, z_ld) || (residual && !vec_ok(c, res_ld)) || rows < 1 || rows_per_img < 1)
    return set_error(-4, "bn_apply: c and pitches must be multiples of 8");
  const RedGeom g = red_geom(rows, c, 1 << 20, 148 * 8);
  bn_apply_kernel<<<dim3(g.gx, g.slabs), 256, 0, STREAM(stream)>>>(y, y_ld, scale, shift, residual, res_ld, nc_scale, rows_per_img,
                                                                   z, z_ld, rows, c, act, dtype, g.cls, g.rows_per_slab);
  return check_launch("bn_apply");
}

extern "C" int segb200_bn_bwd_reduce(const void* dz, const void* z, const void* y, const float* scale, const float* shift,
                                     const float* nc_scale, float* partial, long long rows, long long rows_per_img, int c,
                                     int dz_ld, int z_ld, int y_ld, int act, int dtype, int max_slabs, void* stream) {
  if (!dz || !y || !partial) return set_error(-1, "bn_bwd_reduce: null pointer");
  if (act != ACT_NONE && !z && (!scale || !shift))
    return set_error(-1, "bn_bwd_reduce: the activation mask needs z, or scale and shift to recompute it from y");
  if (!half_dt(dtype)) return set_error(-2, "bn_bwd_reduce: bad dtype");
  if (!vec_ok(c, dz_ld) || !vec_ok(c, y_ld) || (z && !vec_ok(c, z_ld)) || rows < 1 || rows_per_img < 1)
    return set_error(-4, "bn_bwd_reduce: c and pitches must be multiples of 8");
  const RedGeom g = red_geom(rows, c, max_slabs > 0 ? max_slabs : 1 << 20);
  bn_bwd_reduce_kernel<<<dim3(g.gx, g.slabs), 256, 0, STREAM(stream)>>>(dz, dz_ld, z, z_ld, y, y_ld, scale, shift, nc_scale,
                                                                        rows_per_img, rows, c, act, dtype, g.cls,
                                                                        g.rows_per_slab, partial);
  return check_launch("bn_bwd_reduce");
}

extern "C" int segb200_bn_bwd_finalize(const float* partial, int slabs, int c, const float* mean, const float* invstd,
                                       float* sums, float* dgamma, float* dbeta, void* stream) {
  if (!partial || !sums) return set_error(-1, "bn_bwd_finalize: null pointer");
  if (slabs < 1 || c < 1) return set_error(-4, "bn_bwd_finalize: bad sizes");
  bn_bwd_finalize_kernel<<<(c + 3) / 4, 128, 0, STREAM(stream)>>>(partial, slabs, c, mean, invstd, sums, dgamma, dbeta);
  return check_launch("bn_bwd_finalize");
}

extern "C" int segb200_bn_bwd_apply(const void* dz, const void* z, const void* y, const float* mean, const float* invstd,
                                    const float* scale, const float* shift, const float* sums, double count,
                                    const float* nc_scale, void* dy, void* dres, int dres_accumulate, long long rows,
                                    long long rows_per_img, int c, int dz_ld, int z_ld, int y_ld, int dy_ld, int dres_ld, int act,
                                    int dtype, void* stream) {
  if (!dz || (!dy && !dres)) return set_error(-1, "bn_bwd_apply: null pointer");
  if (act != ACT_NONE && !z && (!y || !scale || !shift))
    return set_error(-1, "bn_bwd_apply: the activation mask needs z, or y with scale and shift to recompute it");
  if (sums && (!y || !mean || !invstd || !scale)) return set_error(-1, "bn_bwd_apply: normalised unit needs y/mean/invstd/scale");
  if (!half_dt(dtype)) return set_error(-2, "bn_bwd_apply: bad dtype");
  if (!vec_ok(c, dz_ld) || (y && !vec_ok(c, y_ld)) || (z && !vec_ok(c, z_ld)) || (dy && !vec_ok(c, dy_ld)) ||
      (dres && !vec_ok(c, dres_ld)) || rows < 1 || rows_per_img < 1 || count < 1.0)
    return set_error(-4, "bn_bwd_apply: c and pitches must be multiples of 8");
  const RedGeom g = red_geom(rows, c, 1 << 20, 148 * 8);
  bn_bwd_apply_kernel<<<dim3(g.gx, g.slabs), 256, 0, STREAM(stream)>>>(
      dz, dz_ld, z, z_ld, y, y_ld, mean, invstd, scale, shift, sums, (float)(1.0 / count), nc_scale, rows_per_img, dy, dy_ld, dres,
      dres_ld, dres_accumulate, rows, c, act, dtype, g.cls, g.rows_per_slab);
  return check_launch("bn_bwd_apply");
}

extern "C" int segb200_maxpool3x3s2_bwd(const void* x, const void* dy, void* dx, int n, int h, int w, int c, int x_ld,
                                        int dy_ld, int dx_ld, int dtype, void* stream) {
  if (!x || !dy || !dx) return set_error(-1, "maxpool3x3s2_bwd: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "maxpool3x3s2_bwd: bad dtype");
  if (!vec_ok(c, x_ld) || !vec_ok(c, dy_ld) || !vec_ok(c, dx_ld)) return set_error(-4, "maxpool3x3s2_bwd: bad c/pitches");
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  maxpool3x3s2_bwd_kernel<<<grid_for((long long)n * h * w * (c / 8), 256), 256, 0, STREAM(stream)>>>(x, dy, dx, n, h, w, c, x_ld,
                                                                                                  ho, wo, dy_ld, dx_ld, dtype);
  return check_launch("maxpool3x3s2_bwd");
}

extern "C" int segb200_maxpool3x3s2_idx(const void* x, void* y, uint8_t* idx, int n, int h, int w, int c, int x_ld, int y_ld,
                                        int dtype, void* stream) {
  if (!x || !y || !idx) return set_error(-1, "maxpool3x3s2_idx: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "maxpool3x3s2_idx: bad dtype");
  if (!vec_ok(c, x_ld) || !vec_ok(c, y_ld)) return set_error(-4, "maxpool3x3s2_idx: bad c/pitches");
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  maxpool3x3s2_idx_kernel<<<grid_for((long long)n * ho * wo * (c / 8), 256), 256, 0, STREAM(stream)>>>(x, y, idx, n, h, w, c, x_ld, ho,
                                                                                                    wo, y_ld, dtype);
  return check_launch("maxpool3x3s2_idx");
}

extern "C" int segb200_maxpool3x3s2_bwd_idx(const uint8_t* idx, const void* dy, void* dx, int n, int h, int w, int c, int dy_ld,
                                            int dx_ld, int dtype, void* stream) {
  if (!idx || !dy || !dx) return set_error(-1, "maxpool3x3s2_bwd_idx: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "maxpool3x3s2_bwd_idx: bad dtype");
  if (!vec_ok(c, dy_ld) || !vec_ok(c, dx_ld)) return set_error(-4, "maxpool3x3s2_bwd_idx: bad c/pitches");
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  maxpool3x3s2_bwd_idx_kernel<<<grid_for((long long)n * h * w * (c / 8), 256), 256, 0, STREAM(stream)>>>(idx, dy, dx, n, h, w, c, ho,
                                                                                                      wo, dy_ld, dx_ld, dtype);
  return check_launch("maxpool3x3s2_bwd_idx");
}

extern "C" int segb200_bilinear_nhwc_bwd(const void* dy, void* dx, int n, int hi, int wi, int c, int dx_ld, int ho, int wo,
                                         int dy_ld, int align_corners, int accumulate, const float* gscale, int dtype,
                                         void* stream) {
  if (!dy || !dx) return set_error(-1, "bilinear_nhwc_bwd: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "bilinear_nhwc_bwd: bad dtype");
  if (!vec_ok(c, dx_ld) || !vec_ok(c, dy_ld)) return set_error(-4, "bilinear_nhwc_bwd: bad c/pitches");
  bilinear_bwd_kernel<<<grid_for((long long)n * hi * wi * (c / 8), 256), 256, 0, STREAM(stream)>>>(
      dy, dx, n, hi, wi, c, dx_ld, ho, wo, dy_ld, align_corners, accumulate, gscale, dtype);
  return check_launch("bilinear_nhwc_bwd");
}

extern "C" int segb200_upsample_ce_blocks(int n, int ho, int wo) { return grid_for((long long)n * ho * wo, 256); }

extern "C" int segb200_upsample_ce(const void* logits, const long long* target, void* dfull, float* partial, float* out3,
                                   int n, int hi, int wi, int nclass, int x_ld, int ho, int wo, int d_ld, int align_corners,
                                   int ignore_index, int dtype, void* stream) {
  if (!logits || !target || !dfull || !partial || !out3) return set_error(-1, "upsample_ce: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "upsample_ce: bad dtype");
  if (nclass < 1 || nclass > kMaxClasses || (x_ld & 7) || (d_ld & 7) || x_ld < ((nclass + 7) & ~7) || d_ld < ((nclass + 7) & ~7))
    return set_error(-4, "upsample_ce: nclass <= 32, pitches multiples of 8 and >= round_up(nclass, 8)");
  const int g = grid_for((long long)n * ho * wo, 256);
  upsample_ce_kernel<<<g, 256, 0, STREAM(stream)>>>(logits, target, dfull, partial, n, hi, wi, nclass, x_ld, ho, wo, d_ld,
                                                    align_corners, ignore_index, dtype);
  int rc = check_launch("upsample_ce");
  if (rc) return rc;
  ce_finalize_kernel<<<1, 32, 0, STREAM(stream)>>>(partial, g, out3);
  return check_launch("ce_finalize");
}

extern "C" int segb200_dw_wgrad(const void* x, const void* dy, float* partial, int n, int h, int w, int c, int x_ld,
                                int dy_ld, int dilation, int pre_relu, int dtype, int max_slabs, void* stream) {
  if (!x || !dy || !partial) return set_error(-1, "dw_wgrad: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "dw_wgrad: bad dtype");
  if (!vec_ok(c, x_ld) || !vec_ok(c, dy_ld) || dilation < 1) return set_error(-4, "dw_wgrad: bad c/pitches");
  const long long rows = (long long)n * h * w;
  const RedGeom g = red_geom(rows, c, max_slabs > 0 ? max_slabs : 1 << 20);
  dw_wgrad_kernel<<<dim3(g.gx, g.slabs), 256, 0, STREAM(stream)>>>(x, dy, n, h, w, c, x_ld, dy_ld, dilation, pre_relu, dtype,
                                                                   g.cls, g.rows_per_slab, partial);
  return check_launch("dw_wgrad");
}

extern "C" int segb200_nc_broadcast(const void* v, void* y, int n, long long hw, int c, int v_ld, int y_ld, float scale,
                                    int accumulate, int dtype, void* stream) {
  if (!v || !y) return set_error(-1, "nc_broadcast: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "nc_broadcast: bad dtype");
  if (!vec_ok(c, v_ld) || !vec_ok(c, y_ld)) return set_error(-4, "nc_broadcast: bad c/pitches");
  nc_broadcast_kernel<<<grid_for((long long)n * hw * (c / 8), 256), 256, 0, STREAM(stream)>>>(v, y, n, hw, c, v_ld, y_ld, scale,
                                                                                           accumulate, dtype);
  return check_launch("nc_broadcast");
}

extern "C" int segb200_stride2_place(const void* t, void* z, int n, int h, int w, int c, int t_ld, int z_ld, int mode,
                                     int dtype, void* stream) {
  if (!t || !z) return set_error(-1, "stride2_place: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "stride2_place: bad dtype");
  if (!vec_ok(c, t_ld) || !vec_ok(c, z_ld) || (mode != 0 && mode != 1)) return set_error(-4, "stride2_place: bad arguments");
  const int ht = (h - 1) / 2 + 1, wt = (w - 1) / 2 + 1;
  const long long total = mode == 0 ? (long long)n * h * w * (c / 8) : (long long)n * ht * wt * (c / 8);
  stride2_place_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(t, z, n, h, w, c, ht, wt, t_ld, z_ld, mode, dtype);
  return check_launch("stride2_place");
}

extern "C" int segb200_gather_cast(const float* src, const int* index, void* dst, long long n, int dst_dtype, void* stream) {
  if (!src || !index || !dst) return set_error(-1, "gather_cast: null pointer");
  if (dst_dtype < 0 || dst_dtype > 2 || n < 1) return set_error(-2, "gather_cast: bad arguments");
  gather_cast_kernel<<<grid_for(n, 256), 256, 0, STREAM(stream)>>>(src, index, dst, n, dst_dtype);
  return check_launch("gather_cast");
}

extern "C" int segb200_scatter_add(const float* src, const int* index, float* dst, long long n, void* stream) {
  if (!src || !index || !dst) return set_error(-1, "scatter_add: null pointer");
  if (n < 1) return set_error(-4, "scatter_add: empty");
  scatter_add_kernel<<<grid_for(n, 256), 256, 0, STREAM(stream)>>>(src, index, dst, n);
  return check_launch("scatter_add");
}

extern "C" int segb200_sgd_step(float* p, const float* g, float* m, long long n, float lr, float momentum,
                                float weight_decay, float grad_scale, void* stream) {
  if (!p || !g || !m) return set_error(-1, "sgd_step: null pointer");
  if (n < 1 || (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m) & 15)) return set_error(-4, "sgd_step: buffers must be 16-byte aligned");
  sgd_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, STREAM(stream)>>>(p, g, m, n, lr, momentum, weight_decay, grad_scale);
  return check_launch("sgd_step");
}

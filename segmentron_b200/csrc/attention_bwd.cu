// segb200 -- criss-cross attention BACKWARD (CCNet; replaces _C.ca_backward / _C.ca_map_backward of
// segmentron/modules/csrc/criss_cross_attention/ca_cuda.cu:38-92,122-177 together with the softmax backward and the
// gamma gradient that autograd runs around them in CrissCrossAttention.forward, modules/cc_attention.py:62-72).
//
// forward (attention.cu):  E[p][z] = q[p].k[key(p,z)],  A = softmax_z(E),  out[p] = sum_z A[p][z] v[key(p,z)],  y = gamma*out + x
// backward, given dy:
//   D[p][z]  = dy[p] . v[key(p,z)]                                   (ca_map_backward's dw  ==  ca_forward(dy, v))
//   dgamma   = sum_p sum_z A[p][z] D[p][z]                           (= <dy, out> without storing out)
//   dE[p][z] = gamma * A[p][z] * (D[p][z] - sum_z' A[p][z'] D[p][z'])      <- cca_weight_bwd_kernel (one warp per pixel)
//   dq[p]    = sum_z dE[p][z] k[key(p,z)]                            (ca_backward's dt  ==  ca_map_forward(dE, k))   <- cca_gather_kernel
//   dk[r]    = sum_{(p,z): key(p,z)=r} dE[p][z] q[p]                 (ca_backward's df)                              <- cca_scatter_kernel
//   dv[r]    = gamma * sum_{(p,z): key(p,z)=r} A[p][z] dy[p]         (ca_map_backward's dg)                          <- cca_scatter_kernel
// key(p,z) for p = (y, x):  z < W : (y, z);  z >= W : (j, x) with i = z - W, j = i < y ? i : i + 1.
// Its inverse, used by the "scatter" written as a gather: the queries that see r = (y', x') are (y', x) for every x (weight
// index z = x') and (y, x') for every y != y' (weight index z = W + (y' < y ? y' : y' - 1)).
#include "common.cuh"
#include "vec.cuh"
#include "../../include/segb200.h"

namespace segb200 {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per pixel: D, the softmax backward, and the per-pixel gamma-gradient term
__global__ void __launch_bounds__(256)
cca_weight_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ v, const float* __restrict__ att,
                      float* __restrict__ de, float* __restrict__ dgamma_partial, const float* __restrict__ gamma_p, int n, int h,
                      int w, int c, int dy_ld, int v_ld, int att_ld, int dtype) {
  __shared__ float wsum[8];
  const long long pix = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long npix = (long long)n * h * w;
  float s = 0.f;
  if (pix < npix) {
    const float gamma = __ldg(gamma_p);
    const int x = (int)(pix % w);
    const int y = (int)((pix / w) % h);
    const long long b = pix / ((long long)w * h);
    const int L = h + w - 1;
    const char* dp = reinterpret_cast<const char*>(dy) + pix * dy_ld * 2;
    const char* vb = reinterpret_cast<const char*>(v) + b * h * w * v_ld * 2;
    const float* ap = att + pix * att_ld;
    float* ep = de + pix * att_ld;
    for (int z = lane; z < L; z += 32) {
      int ky, kx;
      if (z < w) { ky = y; kx = z; } else { const int i = z - w; ky = i < y ? i : i + 1; kx = x; }
      const char* kp = vb + ((long long)ky * w + kx) * v_ld * 2;
      float dot = 0.f;
      for (int c0 = 0; c0 < c; c0 += 8) {
        float fa[8], fb[8];
        unpack8(ldg_v4(dp + c0 * 2), dtype, fa);
        unpack8(ldg_v4(kp + c0 * 2), dtype, fb);
#pragma unroll
        for (int j = 0; j < 8; ++j) dot = fmaf(fa[j], fb[j], dot);
      }
      ep[z] = dot;                                   // D staged in the output row
      s = fmaf(__ldg(ap + z), dot, s);
    }
    s = warp_sum_f(s);
    __syncwarp();
    for (int z = lane; z < L; z += 32) ep[z] = gamma * __ldg(ap + z) * (ep[z] - s);
  }
  if (lane == 0) wsum[threadIdx.x >> 5] = pix < npix ? s : 0.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += wsum[i];
    dgamma_partial[blockIdx.x] = t;                  // fixed-order per-block partial; summed by segb200_reduce_partials
  }
}

// out[p] (+)= scale * sum_z a[p][z] * src[key(p,z)]      (thread = 8 channels of one pixel)
__global__ void __launch_bounds__(256)
cca_gather_kernel(const float* __restrict__ a, const void* __restrict__ src, void* __restrict__ out, int n, int h, int w, int c,
                  int a_ld, int src_ld, int out_ld, float scale, int accumulate, int dtype) {
  const int cvn = c / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * h * w * cvn) return;
  const int cv = (int)(idx % cvn);
  const long long pix = idx / cvn;
  const int x = (int)(pix % w);
  const int y = (int)((pix / w) % h);
  const long long b = pix / ((long long)w * h);
  const int L = h + w - 1;
  const float* ap = a + pix * a_ld;
  const char* sb = reinterpret_cast<const char*>(src) + (b * h * w * src_ld + cv * 8) * 2;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int z = 0; z < L; ++z) {
    int ky, kx;
    if (z < w) { ky = y; kx = z; } else { const int i = z - w; ky = i < y ? i : i + 1; kx = x; }
    const float wt = __ldg(ap + z);
    float f[8];
    unpack8(ldg_v4(sb + ((long long)ky * w + kx) * src_ld * 2), dtype, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(wt, f[j], acc[j]);
  }
  char* op = reinterpret_cast<char*>(out) + (pix * out_ld + cv * 8) * 2;
  float o[8];
  if (accumulate) {
    unpack8(*reinterpret_cast<const uint4*>(op), dtype, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(scale, acc[j], o[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = scale * acc[j];
  }
  *reinterpret_cast<uint4*>(op) = pack8(o, dtype);
}

// out[r] (+)= scale * sum over the queries p that see r of a[p][z(p,r)] * src[p]      (thread = 8 channels of one pixel r)
__global__ void __launch_bounds__(256)
cca_scatter_kernel(const float* __restrict__ a, const void* __restrict__ src, void* __restrict__ out, int n, int h, int w, int c,
                   int a_ld, int src_ld, int out_ld, float scale, const float* __restrict__ scale_p, int accumulate, int dtype) {
  const int cvn = c / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * h * w * cvn) return;
  if (scale_p != nullptr) scale *= __ldg(scale_p);
  const int cv = (int)(idx % cvn);
  const long long pix = idx / cvn;
  const int xr = (int)(pix % w);
  const int yr = (int)((pix / w) % h);
  const long long b = pix / ((long long)w * h);
  const float* ab = a + b * h * w * a_ld;
  const char* sb = reinterpret_cast<const char*>(src) + (b * h * w * src_ld + cv * 8) * 2;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int x = 0; x < w; ++x) {                       // queries of the same row: weight index z = xr
    const long long p = (long long)yr * w + x;
    const float wt = __ldg(ab + p * a_ld + xr);
    float f[8];
    unpack8(ldg_v4(sb + p * src_ld * 2), dtype, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(wt, f[j], acc[j]);
  }
  for (int y = 0; y < h; ++y) {                       // queries of the same column (not the pixel itself)
    if (y == yr) continue;
    const long long p = (long long)y * w + xr;
    const int z = w + (yr < y ? yr : yr - 1);
    const float wt = __ldg(ab + p * a_ld + z);
    float f[8];
    unpack8(ldg_v4(sb + p * src_ld * 2), dtype, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(wt, f[j], acc[j]);
  }
  char* op = reinterpret_cast<char*>(out) + (pix * out_ld + cv * 8) * 2;
  float o[8];
  if (accumulate) {
    unpack8(*reinterpret_cast<const uint4*>(op), dtype, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(scale, acc[j], o[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = scale * acc[j];
  }
  *reinterpret_cast<uint4*>(op) = pack8(o, dtype);
}

}  // namespace segb200

using namespace segb200;

extern "C" int segb200_cca_weight_bwd_blocks(int n, int h, int w) { return (int)(((long long)n * h * w + 7) / 8); }

extern "C" int segb200_cca_weight_bwd(const void* dy, const void* v, const float* att, float* de, float* dgamma_partial,
                                      const float* gamma, int n, int h, int w, int c, int dy_ld, int v_ld, int att_ld, int dtype,
                                      void* stream) {
  if (!dy || !v || !att || !de || !dgamma_partial || !gamma) return set_error(-1, "cca_weight_bwd: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "cca_weight_bwd: bad dtype");
  if ((c & 7) || (dy_ld & 7) || (v_ld & 7) || att_ld < h + w - 1) return set_error(-4, "cca_weight_bwd: bad sizes");
  const long long blocks = ((long long)n * h * w + 7) / 8;
  if (blocks > 0x7fffffffLL) return set_error(-8, "cca_weight_bwd: too many pixels");
  cca_weight_bwd_kernel<<<(int)blocks, 256, 0, STREAM(stream)>>>(dy, v, att, de, dgamma_partial, gamma, n, h, w, c, dy_ld, v_ld,
                                                                att_ld, dtype);
  return check_launch("cca_weight_bwd");
}

extern "C" int segb200_cca_gather(const float* a, const void* src, void* out, int n, int h, int w, int c, int a_ld, int src_ld,
                                  int out_ld, float scale, int accumulate, int dtype, void* stream) {
  if (!a || !src || !out) return set_error(-1, "cca_gather: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "cca_gather: bad dtype");
  if ((c & 7) || (src_ld & 7) || (out_ld & 7) || a_ld < h + w - 1) return set_error(-4, "cca_gather: bad sizes");
  const long long blocks = ((long long)n * h * w * (c / 8) + 255) / 256;
  if (blocks > 0x7fffffffLL) return set_error(-8, "cca_gather: too large");
  cca_gather_kernel<<<(int)blocks, 256, 0, STREAM(stream)>>>(a, src, out, n, h, w, c, a_ld, src_ld, out_ld, scale, accumulate, dtype);
  return check_launch("cca_gather");
}

extern "C" int segb200_cca_scatter(const float* a, const void* src, void* out, int n, int h, int w, int c, int a_ld, int src_ld,
                                   int out_ld, float scale, const float* scale_dev, int accumulate, int dtype, void* stream) {
  if (!a || !src || !out) return set_error(-1, "cca_scatter: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "cca_scatter: bad dtype");
  if ((c & 7) || (src_ld & 7) || (out_ld & 7) || a_ld < h + w - 1) return set_error(-4, "cca_scatter: bad sizes");
  const long long blocks = ((long long)n * h * w * (c / 8) + 255) / 256;
  if (blocks > 0x7fffffffLL) return set_error(-8, "cca_scatter: too large");
  cca_scatter_kernel<<<(int)blocks, 256, 0, STREAM(stream)>>>(a, src, out, n, h, w, c, a_ld, src_ld, out_ld, scale, scale_dev,
                                                             accumulate, dtype);
  return check_launch("cca_scatter");
}

// segb200 -- row softmax forward / backward on a materialised attention matrix: the TRAINING-mode PAM_Module
// (modules/module.py:100-131: energy = bmm(Q, K), attention = softmax(energy, -1), out = bmm(V, attention^T)), which the
// reference also materialises (a [N, N] fp32 matrix per image).  The inference path never forms it (csrc/pam.cu); training keeps
// P in 16 bits because the backward needs it four times (dV, dS, and through dS: dQ, dK).
//   row_softmax     : P[r][j] = softmax_j(S[r][j]), j < n;  columns n .. p_ld-1 are written as zeros (K padding of the next GEMM)
//   row_softmax_bwd : r = sum_j P D (-> partial[r], the gamma gradient term);  dS = (*gamma) * P * (D - r), zero padding as above
// One warp per row; S / D are fp32 GEMM outputs.  Bound: HBM (N^2 elements read once, written once).
#include "vec.cuh"
#include "../../include/segb200.h"

namespace segb200 {

__device__ __forceinline__ float wsum_(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax_(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(256)
row_softmax_kernel(const float* __restrict__ s, void* __restrict__ p, int rows, int n, int s_ld, int p_ld, int dtype) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* sr = s + (long long)row * s_ld;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32) mx = fmaxf(mx, sr[j]);
  mx = wmax_(mx);
  float sum = 0.f;
  for (int j = lane; j < n; j += 32) sum += __expf(sr[j] - mx);
  sum = wsum_(sum);
  const float inv = 1.f / sum;
  for (int j = lane; j < p_ld; j += 32)
    store_any(p, (long long)row * p_ld + j, j < n ? __expf(sr[j] - mx) * inv : 0.f, dtype);
}

__global__ void __launch_bounds__(256)
row_softmax_bwd_kernel(const void* __restrict__ p, const float* __restrict__ d, const float* __restrict__ gamma,
                       void* __restrict__ ds, float* __restrict__ partial, int rows, int n, int p_ld, int d_ld, int ds_ld, int dtype) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float gm = __ldg(gamma);
  const float* dr = d + (long long)row * d_ld;
  float r = 0.f;
  for (int j = lane; j < n; j += 32) r += load_any(p, (long long)row * p_ld + j, dtype) * dr[j];
  r = wsum_(r);
  for (int j = lane; j < ds_ld; j += 32)
    store_any(ds, (long long)row * ds_ld + j, j < n ? gm * load_any(p, (long long)row * p_ld + j, dtype) * (dr[j] - r) : 0.f, dtype);
  if (lane == 0) partial[row] = r;
}

}  // namespace segb200

using namespace segb200;

extern "C" int segb200_row_softmax(const float* s, void* p, int rows, int n, int s_ld, int p_ld, int dtype, void* stream) {
  if (!s || !p) return set_error(-1, "row_softmax: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "row_softmax: bad dtype");
  if (rows < 1 || n < 1 || s_ld < n || p_ld < n) return set_error(-4, "row_softmax: bad sizes");
  row_softmax_kernel<<<(rows + 7) / 8, 256, 0, STREAM(stream)>>>(s, p, rows, n, s_ld, p_ld, dtype);
  return check_launch("row_softmax");
}

extern "C" int segb200_row_softmax_bwd(const void* p, const float* d, const float* gamma, void* ds, float* partial, int rows, int n,
                                       int p_ld, int d_ld, int ds_ld, int dtype, void* stream) {
  if (!p || !d || !gamma || !ds || !partial) return set_error(-1, "row_softmax_bwd: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "row_softmax_bwd: bad dtype");
  if (rows < 1 || n < 1 || p_ld < n || d_ld < n || ds_ld < n) return set_error(-4, "row_softmax_bwd: bad sizes");
  row_softmax_bwd_kernel<<<(rows + 7) / 8, 256, 0, STREAM(stream)>>>(p, d, gamma, ds, partial, rows, n, p_ld, d_ld, ds_ld, dtype);
  return check_launch("row_softmax_bwd");
}

// segb200 -- CAM_Module backward glue (modules/module.py:134-162; autograd through bmm / softmax / bmm in the reference).
//
//   forward :  E = X^T X,  A = softmax(rowmax(E) - E),  out = A X,  y = gamma * out + x          (attention.cu / conv_gemm.cu)
//   backward:  G[c1][c2] = sum_p dy[p,c1] x[p,c2]                       (tcgen05 GEMM, fp32 output, like E)
//              r[c1]     = sum_c2 A[c1][c2] G[c1][c2]                   -> dgamma = sum_c1 r[c1]
//              dE        = -gamma * A * (G - r)                         (softmax backward; the sign is the "rowmax - E" of the forward)
//              dx        = dy + dy . (gamma A)  +  x . (dE + dE^T)      (two tcgen05 GEMMs whose 16-bit weights W1 = gamma A^T and
//                                                                        W2 = dE + dE^T are packed here)
// Both kernels work on C x C matrices (C = 512 in DANet): latency-bound, one warp per row.
#include "vec.cuh"
#include "../../include/segb200.h"

namespace segb200 {

__device__ __forceinline__ float warp_sum_(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
cam_softmax_bwd_kernel(const void* __restrict__ att, const float* __restrict__ g, const float* __restrict__ gamma,
                       float* __restrict__ de, float* __restrict__ dgamma_partial, int c, int att_ld, int g_ld, int de_ld, int dtype) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= c) return;
  const float gm = __ldg(gamma);
  float r = 0.f;
  for (int j = lane; j < c; j += 32) r += load_any(att, (long long)row * att_ld + j, dtype) * g[(long long)row * g_ld + j];
  r = warp_sum_(r);
  for (int j = lane; j < c; j += 32)
    de[(long long)row * de_ld + j] = -gm * load_any(att, (long long)row * att_ld + j, dtype) * (g[(long long)row * g_ld + j] - r);
  if (lane == 0) dgamma_partial[row] = r;
}

// w1[c2][c1] = gamma * A[c1][c2];  w2[c1][c2] = dE[c1][c2] + dE[c2][c1];  columns c .. w_ld-1 are zero (K padding of the GEMMs)
__global__ void __launch_bounds__(256)
cam_bwd_pack_kernel(const void* __restrict__ att, const float* __restrict__ de, const float* __restrict__ gamma, void* __restrict__ w1,
                    void* __restrict__ w2, int c, int att_ld, int de_ld, int w_ld, int dtype) {
  const long long total = (long long)c * w_ld;
  const float gm = __ldg(gamma);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(idx % w_ld);
    const int row = (int)(idx / w_ld);
    float a = 0.f, s = 0.f;
    if (col < c) {
      a = gm * load_any(att, (long long)col * att_ld + row, dtype);
      s = de[(long long)row * de_ld + col] + de[(long long)col * de_ld + row];
    }
    store_any(w1, idx, a, dtype);
    store_any(w2, idx, s, dtype);
  }
}

}  // namespace segb200

using namespace segb200;

extern "C" int segb200_cam_softmax_bwd(const void* att, const float* g, const float* gamma, float* de, float* dgamma_partial, int c,
                                       int att_ld, int g_ld, int de_ld, int dtype, void* stream) {
  if (!att || !g || !gamma || !de || !dgamma_partial) return set_error(-1, "cam_softmax_bwd: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "cam_softmax_bwd: bad dtype");
  if (c < 1 || att_ld < c || g_ld < c || de_ld < c) return set_error(-4, "cam_softmax_bwd: bad sizes");
  cam_softmax_bwd_kernel<<<(c + 7) / 8, 256, 0, STREAM(stream)>>>(att, g, gamma, de, dgamma_partial, c, att_ld, g_ld, de_ld, dtype);
  return check_launch("cam_softmax_bwd");
}

extern "C" int segb200_cam_bwd_pack(const void* att, const float* de, const float* gamma, void* w1, void* w2, int c, int att_ld,
                                    int de_ld, int w_ld, int dtype, void* stream) {
  if (!att || !de || !gamma || !w1 || !w2) return set_error(-1, "cam_bwd_pack: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "cam_bwd_pack: bad dtype");
  if (c < 1 || att_ld < c || de_ld < c || w_ld < c) return set_error(-4, "cam_bwd_pack: bad sizes");
  cam_bwd_pack_kernel<<<grid_for((long long)c * w_ld, 256), 256, 0, STREAM(stream)>>>(att, de, gamma, w1, w2, c, att_ld, de_ld, w_ld,
                                                                                    dtype);
  return check_launch("cam_bwd_pack");
}

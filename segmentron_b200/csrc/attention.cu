// segb200 -- attention glue: CAM softmax, criss-cross attention (CCNet).  NHWC, fp32 statistics.
//
// CAM (modules/module.py:142-162):   E = X^T X  (tcgen05 GEMM with fp32 epilogue, see conv_gemm.cu)
//                                    A = softmax(rowmax(E) - E)          <- cam_softmax_kernel (this file)
//                                    y = gamma * (A X) + x               (tcgen05 GEMM with scale = gamma, residual = x)
// CCA (modules/cc_attention.py:62-72, csrc/criss_cross_attention/ca_cuda.cu):
//   energy[p][z] = q[p] . k[key(p,z)],  z <  W : key = (y, z)                         (same row, self included)
//                                       z >= W : key = (j, x), j = i<y ? i : i+1, i=z-W (same column, self excluded)
//   A = softmax_z(energy);   out[p] = sum_z A[p][z] v[key(p,z)];   y = gamma*out + x
//   cca_weight_softmax_kernel fuses ca_forward + softmax, cca_map_kernel fuses ca_map_forward + gamma*out + x.
#include "common.cuh"
#include "../../include/segb200.h"

namespace segb200 {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per row of E [rows][c]: att[j] = exp(min_j(E) - E[j]) / sum  (== softmax(rowmax - E))
template <bool kBF16>
__global__ void __launch_bounds__(256)
cam_softmax_kernel(const float* __restrict__ e, void* __restrict__ att, int rows, int c, int e_ld, int att_ld) {
  using H = Half2<kBF16>;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* er = e + (long long)row * e_ld;
  float mn = INFINITY;
  for (int j = lane; j < c; j += 32) mn = fminf(mn, er[j]);
  mn = -warp_max(-mn);
  float sum = 0.f;
  for (int j = lane; j < c; j += 32) sum += __expf(mn - er[j]);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  typename H::T* ar = reinterpret_cast<typename H::T*>(att) + (long long)row * att_ld;
  for (int j = lane; j < c; j += 32) ar[j] = H::from_f(__expf(mn - er[j]) * inv);
}

// one warp per pixel: energies over the criss-cross neighbourhood + softmax -> att [n][h][w][att_ld] fp32
template <bool kBF16>
__global__ void __launch_bounds__(256)
cca_weight_softmax_kernel(const void* __restrict__ q, const void* __restrict__ k, float* __restrict__ att, int n, int h,
                          int w, int c, int q_ld, int k_ld, int att_ld) {
  using H = Half2<kBF16>;
  using T = typename H::T;
  const long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long npix = (long long)n * h * w;
  if (pix >= npix) return;
  const int x = (int)(pix % w);
  const int y = (int)((pix / w) % h);
  const int b = (int)(pix / ((long long)w * h));
  const int L = h + w - 1;
  const T* qp = reinterpret_cast<const T*>(q) + pix * q_ld;
  const T* kb = reinterpret_cast<const T*>(k) + (long long)b * h * w * k_ld;
  float* ap = att + pix * att_ld;
  float mx = -INFINITY;
  for (int z = lane; z < L; z += 32) {
    int ky, kx;
    if (z < w) { ky = y; kx = z; } else { const int i = z - w; ky = i < y ? i : i + 1; kx = x; }
    const T* kp = kb + ((long long)ky * w + kx) * k_ld;
    float dot = 0.f;
    for (int c0 = 0; c0 < c; c0 += 8) {
      const uint4 a = ldg_v4(qp + c0), bb = ldg_v4(kp + c0);
      const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fa = H::unpack(ua[j]), fb = H::unpack(ub[j]);
        dot = fmaf(fa.x, fb.x, dot); dot = fmaf(fa.y, fb.y, dot);
      }
    }
    ap[z] = dot;                       // staged in the output row, rewritten below
    mx = fmaxf(mx, dot);
  }
  mx = warp_max(mx);
  __syncwarp();
  float sum = 0.f;
  for (int z = lane; z < L; z += 32) { const float e = __expf(ap[z] - mx); ap[z] = e; sum += e; }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int z = lane; z < L; z += 32) ap[z] *= inv;
}

// aggregate: thread = 8 channels of one pixel; block = 256 threads = 256 / (c/8) consecutive pixels
template <bool kBF16>
__global__ void __launch_bounds__(256)
cca_map_kernel(const float* __restrict__ att, const void* __restrict__ v, const void* __restrict__ xres, void* __restrict__ y,
               const float* __restrict__ gamma_p, int n, int h, int w, int c, int att_ld, int v_ld, int x_ld, int y_ld) {
  using H = Half2<kBF16>;
  const float gamma = __ldg(gamma_p);
  using T = typename H::T;
  const int cvn = c / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n * h * w * cvn;
  if (idx >= total) return;
  const int cv = (int)(idx % cvn);
  const long long pix = idx / cvn;
  const int x = (int)(pix % w);
  const int yy = (int)((pix / w) % h);
  const int b = (int)(pix / ((long long)w * h));
  const int L = h + w - 1;
  const float* ap = att + pix * att_ld;
  const T* vb = reinterpret_cast<const T*>(v) + (long long)b * h * w * v_ld + cv * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int z = 0; z < L; ++z) {
    int ky, kx;
    if (z < w) { ky = yy; kx = z; } else { const int i = z - w; ky = i < yy ? i : i + 1; kx = x; }
    const float a = __ldg(ap + z);
    const uint4 vv = ldg_v4(vb + ((long long)ky * w + kx) * v_ld);
    const uint32_t u[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = H::unpack(u[j]);
      acc[2 * j] = fmaf(a, f.x, acc[2 * j]); acc[2 * j + 1] = fmaf(a, f.y, acc[2 * j + 1]);
    }
  }
  const uint4 xr = ldg_v4(reinterpret_cast<const T*>(xres) + pix * x_ld + cv * 8);
  const uint32_t ux[4] = {xr.x, xr.y, xr.z, xr.w};
  uint4 o;
  uint32_t* po = &o.x;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = H::unpack(ux[j]);
    po[j] = H::pack(fmaf(gamma, acc[2 * j], f.x), fmaf(gamma, acc[2 * j + 1], f.y));
  }
  *reinterpret_cast<uint4*>(reinterpret_cast<T*>(y) + pix * y_ld + cv * 8) = o;
}

}  // namespace segb200

using namespace segb200;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)
static inline bool half_dt(int d) { return d == DT_BF16 || d == DT_F16; }

extern "C" int segb200_cam_softmax(const float* energy, void* att, int rows, int c, int e_ld, int att_ld, int dtype,
                                   void* stream) {
  if (!energy || !att) return set_error(-1, "cam_softmax: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "cam_softmax: bad dtype");
  if (rows < 1 || c < 1) return set_error(-6, "cam_softmax: empty");
  const int blocks = (rows + 7) / 8;
  if (dtype == DT_BF16) cam_softmax_kernel<true><<<blocks, 256, 0, STREAM(stream)>>>(energy, att, rows, c, e_ld, att_ld);
  else cam_softmax_kernel<false><<<blocks, 256, 0, STREAM(stream)>>>(energy, att, rows, c, e_ld, att_ld);
  return check_launch("cam_softmax");
}

extern "C" int segb200_cca_weight_softmax(const void* q, const void* k, float* att, int n, int h, int w, int c, int q_ld,
                                          int k_ld, int att_ld, int dtype, void* stream) {
  if (!q || !k || !att) return set_error(-1, "cca_weight_softmax: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "cca_weight_softmax: bad dtype");
  if ((c & 7) || (q_ld & 7) || (k_ld & 7) || att_ld < h + w - 1) return set_error(-4, "cca_weight_softmax: bad sizes");
  const long long npix = (long long)n * h * w;
  const long long blocks = (npix + 7) / 8;
  if (blocks > 0x7fffffffLL) return set_error(-8, "cca_weight_softmax: too many pixels");
  if (dtype == DT_BF16)
    cca_weight_softmax_kernel<true><<<(int)blocks, 256, 0, STREAM(stream)>>>(q, k, att, n, h, w, c, q_ld, k_ld, att_ld);
  else
    cca_weight_softmax_kernel<false><<<(int)blocks, 256, 0, STREAM(stream)>>>(q, k, att, n, h, w, c, q_ld, k_ld, att_ld);
  return check_launch("cca_weight_softmax");
}

extern "C" int segb200_cca_map(const float* att, const void* v, const void* x, void* y, const float* gamma, int n, int h, int w,
                               int c, int att_ld, int v_ld, int x_ld, int y_ld, int dtype, void* stream) {
  if (!att || !v || !x || !y || !gamma) return set_error(-1, "cca_map: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "cca_map: bad dtype");
  if ((c & 7) || (v_ld & 7) || (x_ld & 7) || (y_ld & 7) || att_ld < h + w - 1) return set_error(-4, "cca_map: bad sizes");
  const long long total = (long long)n * h * w * (c / 8);
  const long long blocks = (total + 255) / 256;
  if (blocks > 0x7fffffffLL) return set_error(-8, "cca_map: too large");
  if (dtype == DT_BF16)
    cca_map_kernel<true><<<(int)blocks, 256, 0, STREAM(stream)>>>(att, v, x, y, gamma, n, h, w, c, att_ld, v_ld, x_ld, y_ld);
  else
    cca_map_kernel<false><<<(int)blocks, 256, 0, STREAM(stream)>>>(att, v, x, y, gamma, n, h, w, c, att_ld, v_ld, x_ld, y_ld);
  return check_launch("cca_map");
}

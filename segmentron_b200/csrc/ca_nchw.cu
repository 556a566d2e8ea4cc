// segb200 -- the four functions of the reference's only native module, `segmentron._C`
// (segmentron/modules/csrc/vision.cpp:6-11; signatures csrc/criss_cross_attention/ca.h:25-72), as C-ABI entry points on the
// REFERENCE's own tensor layout: contiguous NCHW, element type fp32 / fp16 / bf16, fp32 accumulation.
//
//   ca_forward(t[N,C,H,W], f[N,C,H,W])            -> w[N,H+W-1,H,W]      energies of a pixel against its row and its column
//   ca_backward(dw, t, f)                         -> dt, df
//   ca_map_forward(w[N,H+W-1,H,W], g[N,C,H,W])    -> out[N,C,H,W]        aggregation along the same criss-cross
//   ca_map_backward(dout, w, g)                   -> dw, dg
//
// Index map (the algorithm, ca_cuda.cu:8-36): energy channel i < W pairs pixel (y,x) with (y,i) (the pixel itself included);
// channel W+j, j < H-1, pairs it with (yy,x), yy = j < y ? j : j+1 (the pixel itself skipped).
//
// The four functions are three kernels:
//   cc_energy  (a, b)  -> e[p][i] = sum_c a[c][p] * b[c][key(p,i)]            ca_forward(t,f), ca_map_backward's dw (dout, g)
//   cc_gather  (e, s)  -> o[c][p] = sum_i e[p][i] * s[c][key(p,i)]            ca_map_forward(w,g), ca_backward's dt (dw, f)
//   cc_scatter (e, s)  -> o[c][r] = sum_{(p,i): key(p,i)=r} e[p][i] * s[c][p] ca_backward's df (dw, t), ca_map_backward's dg (w, dout)
// Unlike the reference kernels (one thread per output element, partial sums added into global memory), every output is produced
// by one thread from registers: no atomics, no zero-initialised outputs, bit-reproducible.  A block owns 32 consecutive x of one
// image row; the criss-cross coefficients of those 32 pixels are staged once in shared memory and reused for all channels.
// The fused NHWC kernels of csrc/attention.cu are what the drop-in CrissCrossAttention module runs; these exist so that the
// reference's own `cc_attention.py` (its autograd Functions _CAWeight/_CAMap, :11-45) runs unchanged on a current stack.
#include "common.cuh"
#include "../../include/segb200.h"

namespace segb200 {

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(__ldg(p)); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(__ldg(p)); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__half>(__half* p, float v) { *p = __float2half_rn(v); }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

constexpr int kTX = 32;      // x lanes per block
constexpr int kTY = 8;       // second block dimension: energy channels (cc_energy) or feature channels (gather / scatter)

// e[n][i][y][x] = sum_c a[n][c][y][x] * b[n][c][ky][kx],  (ky,kx) = i < W ? (y,i) : (yy(i-W, y), x)
template <typename T>
__global__ void __launch_bounds__(kTX * kTY) cc_energy_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ e,
                                                               int C, int H, int W) {
  extern __shared__ float sm[];                      // a[c][x-lane] of this row segment: C x 32
  const int lx = threadIdx.x, ly = threadIdx.y;
  const int x = blockIdx.x * kTX + lx, y = blockIdx.y, n = blockIdx.z;
  const size_t plane = (size_t)H * W;
  const T* an = a + (size_t)n * C * plane;
  const T* bn = b + (size_t)n * C * plane;
  for (int c = ly; c < C; c += kTY) sm[c * kTX + lx] = x < W ? ldf(an + c * plane + (size_t)y * W + x) : 0.f;
  __syncthreads();
  if (x >= W) return;
  const int L = H + W - 1;
  T* en = e + (size_t)n * L * plane + (size_t)y * W + x;
  for (int i = ly; i < L; i += kTY) {
    size_t off;
    if (i < W) off = (size_t)y * W + i;
    else { const int j = i - W; off = (size_t)(j < y ? j : j + 1) * W + x; }
    const T* bp = bn + off;
    float acc = 0.f;
#pragma unroll 4
    for (int c = 0; c < C; ++c) acc = fmaf(sm[c * kTX + lx], ldf(bp + c * plane), acc);
    stf(en + i * plane, acc);
  }
}

// o[n][c][y][x] = sum_i e[n][i][y][x] * s[n][c][ky][kx]
template <typename T>
__global__ void __launch_bounds__(kTX * kTY) cc_gather_kernel(const T* __restrict__ e, const T* __restrict__ s, T* __restrict__ o,
                                                               int C, int H, int W) {
  extern __shared__ float sm[];                      // e[i][x-lane]: (H+W-1) x 32
  const int lx = threadIdx.x, ly = threadIdx.y;
  const int x = blockIdx.x * kTX + lx, y = blockIdx.y, n = blockIdx.z;
  const size_t plane = (size_t)H * W;
  const int L = H + W - 1;
  const T* en = e + (size_t)n * L * plane + (size_t)y * W;
  for (int i = ly; i < L; i += kTY) sm[i * kTX + lx] = x < W ? ldf(en + i * plane + x) : 0.f;
  __syncthreads();
  if (x >= W) return;
  const T* sn = s + (size_t)n * C * plane;
  T* on = o + (size_t)n * C * plane + (size_t)y * W + x;
  for (int c = ly; c < C; c += kTY) {
    const T* sc = sn + c * plane;
    const T* row = sc + (size_t)y * W;               // (y, i): the same address for the whole warp
    float acc = 0.f;
#pragma unroll 4
    for (int i = 0; i < W; ++i) acc = fmaf(sm[i * kTX + lx], ldf(row + i), acc);
    const T* col = sc + x;                           // (yy, x): coalesced over the warp
    for (int j = 0; j < y; ++j) acc = fmaf(sm[(W + j) * kTX + lx], ldf(col + (size_t)j * W), acc);
    for (int j = y; j < H - 1; ++j) acc = fmaf(sm[(W + j) * kTX + lx], ldf(col + (size_t)(j + 1) * W), acc);
    stf(on + c * plane, acc);
  }
}

// o[n][c][y][x] = sum_{i<W} e[n][x][y][i] * s[n][c][y][i]  +  sum_{y' != y} e[n][W + jj][y'][x] * s[n][c][y'][x],
//                 jj = y < y' ? y : y - 1   (the energy channel through which row y' sees row y)
template <typename T>
__global__ void __launch_bounds__(kTX * kTY) cc_scatter_kernel(const T* __restrict__ e, const T* __restrict__ s, T* __restrict__ o,
                                                                int C, int H, int W) {
  extern __shared__ float sm[];
  const int lx = threadIdx.x, ly = threadIdx.y;
  const int x0 = blockIdx.x * kTX, x = x0 + lx, y = blockIdx.y, n = blockIdx.z;
  const size_t plane = (size_t)H * W;
  const int L = H + W - 1;
  const int pitch = W | 1;                           // odd pitch: the transposed read below is bank-conflict free
  float* srow = sm;                                  // srow[lane][i] = e[n][x0+lane][y][i]      (32 x W, rows read contiguously)
  float* scol = sm + kTX * pitch;                    // scol[y'][lane] = e[n][W+jj][y'][x]        (H x 32; row y itself = 0)
  const T* en = e + (size_t)n * L * plane;
  for (int r = ly; r < kTX; r += kTY) {
    const int xc = x0 + r;
    for (int i = lx; i < W; i += kTX) srow[r * pitch + i] = xc < W ? ldf(en + (size_t)xc * plane + (size_t)y * W + i) : 0.f;
  }
  for (int yp = ly; yp < H; yp += kTY) {
    float v = 0.f;
    if (yp != y && x < W) v = ldf(en + (size_t)(W + (y < yp ? y : y - 1)) * plane + (size_t)yp * W + x);
    scol[yp * kTX + lx] = v;
  }
  __syncthreads();
  if (x >= W) return;
  const T* sn = s + (size_t)n * C * plane;
  T* on = o + (size_t)n * C * plane + (size_t)y * W + x;
  for (int c = ly; c < C; c += kTY) {
    const T* sc = sn + c * plane;
    const T* row = sc + (size_t)y * W;
    float acc = 0.f;
#pragma unroll 4
    for (int i = 0; i < W; ++i) acc = fmaf(srow[lx * pitch + i], ldf(row + i), acc);
    const T* col = sc + x;
#pragma unroll 4
    for (int yp = 0; yp < H; ++yp) acc = fmaf(scol[yp * kTX + lx], ldf(col + (size_t)yp * W), acc);
    stf(on + c * plane, acc);
  }
}

static int check_geo(const char* what, int n, int c, int h, int w, int dtype, size_t smem_floats) {
  if (n < 1 || c < 1 || h < 1 || w < 1) return set_error(-6, "%s: empty tensor", what);
  if (dtype != DT_BF16 && dtype != DT_F16 && dtype != DT_F32) return set_error(-2, "%s: dtype must be f32, f16 or bf16", what);
  if (h > 65535 || n > 65535) return set_error(-6, "%s: H and N must be < 65536", what);
  if (smem_floats * 4 > 200 * 1024) return set_error(-8, "%s: H + W too large for the shared-memory staging (%zu bytes)", what, smem_floats * 4);
  return 0;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return set_error((int)e, "ca: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  return 0;
}

#define CC_DISPATCH(KERNEL, SMEM, ...)                                                                                    \
  do {                                                                                                                    \
    const dim3 grid((unsigned)((w + kTX - 1) / kTX), (unsigned)h, (unsigned)n), block(kTX, kTY);                          \
    int rc_;                                                                                                              \
    if (dtype == DT_F32) {                                                                                                \
      if ((rc_ = set_smem(KERNEL<float>, SMEM))) return rc_;                                                              \
      KERNEL<float><<<grid, block, SMEM, stream>>>(__VA_ARGS__);                                                          \
    } else if (dtype == DT_F16) {                                                                                         \
      if ((rc_ = set_smem(KERNEL<__half>, SMEM))) return rc_;                                                             \
      KERNEL<__half><<<grid, block, SMEM, stream>>>(__VA_ARGS__);                                                         \
    } else {                                                                                                              \
      if ((rc_ = set_smem(KERNEL<__nv_bfloat16>, SMEM))) return rc_;                                                      \
      KERNEL<__nv_bfloat16><<<grid, block, SMEM, stream>>>(__VA_ARGS__);                                                  \
    }                                                                                                                     \
  } while (0)

template <typename T> static const T* cp(const void* p) { return reinterpret_cast<const T*>(p); }

static int energy(const void* a, const void* b, void* e, int n, int c, int h, int w, int dtype, cudaStream_t stream, const char* what) {
  const size_t smem = (size_t)c * kTX * 4;
  int rc = check_geo(what, n, c, h, w, dtype, (size_t)c * kTX);
  if (rc) return rc;
#define ARGS(T) cp<T>(a), cp<T>(b), reinterpret_cast<T*>(e), c, h, w
  { const dim3 grid((unsigned)((w + kTX - 1) / kTX), (unsigned)h, (unsigned)n), block(kTX, kTY);
    if (dtype == DT_F32) { if ((rc = set_smem(cc_energy_kernel<float>, smem))) return rc; cc_energy_kernel<float><<<grid, block, smem, stream>>>(ARGS(float)); }
    else if (dtype == DT_F16) { if ((rc = set_smem(cc_energy_kernel<__half>, smem))) return rc; cc_energy_kernel<__half><<<grid, block, smem, stream>>>(ARGS(__half)); }
    else { if ((rc = set_smem(cc_energy_kernel<__nv_bfloat16>, smem))) return rc; cc_energy_kernel<__nv_bfloat16><<<grid, block, smem, stream>>>(ARGS(__nv_bfloat16)); } }
#undef ARGS
  return check_launch(what);
}

static int gather(const void* e, const void* s, void* o, int n, int c, int h, int w, int dtype, cudaStream_t stream, const char* what) {
  const size_t fl = (size_t)(h + w - 1) * kTX, smem = fl * 4;
  int rc = check_geo(what, n, c, h, w, dtype, fl);
  if (rc) return rc;
#define ARGS(T) cp<T>(e), cp<T>(s), reinterpret_cast<T*>(o), c, h, w
  { const dim3 grid((unsigned)((w + kTX - 1) / kTX), (unsigned)h, (unsigned)n), block(kTX, kTY);
    if (dtype == DT_F32) { if ((rc = set_smem(cc_gather_kernel<float>, smem))) return rc; cc_gather_kernel<float><<<grid, block, smem, stream>>>(ARGS(float)); }
    else if (dtype == DT_F16) { if ((rc = set_smem(cc_gather_kernel<__half>, smem))) return rc; cc_gather_kernel<__half><<<grid, block, smem, stream>>>(ARGS(__half)); }
    else { if ((rc = set_smem(cc_gather_kernel<__nv_bfloat16>, smem))) return rc; cc_gather_kernel<__nv_bfloat16><<<grid, block, smem, stream>>>(ARGS(__nv_bfloat16)); } }
#undef ARGS
  return check_launch(what);
}

static int scatter(const void* e, const void* s, void* o, int n, int c, int h, int w, int dtype, cudaStream_t stream, const char* what) {
  const size_t fl = (size_t)kTX * (w | 1) + (size_t)h * kTX, smem = fl * 4;
  int rc = check_geo(what, n, c, h, w, dtype, fl);
  if (rc) return rc;
#define ARGS(T) cp<T>(e), cp<T>(s), reinterpret_cast<T*>(o), c, h, w
  { const dim3 grid((unsigned)((w + kTX - 1) / kTX), (unsigned)h, (unsigned)n), block(kTX, kTY);
    if (dtype == DT_F32) { if ((rc = set_smem(cc_scatter_kernel<float>, smem))) return rc; cc_scatter_kernel<float><<<grid, block, smem, stream>>>(ARGS(float)); }
    else if (dtype == DT_F16) { if ((rc = set_smem(cc_scatter_kernel<__half>, smem))) return rc; cc_scatter_kernel<__half><<<grid, block, smem, stream>>>(ARGS(__half)); }
    else { if ((rc = set_smem(cc_scatter_kernel<__nv_bfloat16>, smem))) return rc; cc_scatter_kernel<__nv_bfloat16><<<grid, block, smem, stream>>>(ARGS(__nv_bfloat16)); } }
#undef ARGS
  return check_launch(what);
}

}  // namespace segb200

using namespace segb200;

#define NULLCHK(what, ...)                                                              \
  do {                                                                                  \
    const void* ps_[] = {__VA_ARGS__};                                                  \
    for (const void* p_ : ps_) if (!p_) return set_error(-1, what ": null pointer argument"); \
  } while (0)

extern "C" int segb200_ca_forward(const void* t, const void* f, void* weight, int n, int c, int h, int w, int dtype, void* stream) {
  NULLCHK("ca_forward", t, f, weight);
  return energy(t, f, weight, n, c, h, w, dtype, reinterpret_cast<cudaStream_t>(stream), "ca_forward");
}

extern "C" int segb200_ca_backward(const void* dw, const void* t, const void* f, void* dt, void* df, int n, int c, int h, int w, int dtype,
                                   void* stream) {
  NULLCHK("ca_backward", dw, t, f, dt, df);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int rc = gather(dw, f, dt, n, c, h, w, dtype, s, "ca_backward(dt)");
  if (rc) return rc;
  return scatter(dw, t, df, n, c, h, w, dtype, s, "ca_backward(df)");
}

extern "C" int segb200_ca_map_forward(const void* weight, const void* g, void* out, int n, int c, int h, int w, int dtype, void* stream) {
  NULLCHK("ca_map_forward", weight, g, out);
  return gather(weight, g, out, n, c, h, w, dtype, reinterpret_cast<cudaStream_t>(stream), "ca_map_forward");
}

extern "C" int segb200_ca_map_backward(const void* dout, const void* weight, const void* g, void* dw, void* dg, int n, int c, int h, int w,
                                       int dtype, void* stream) {
  NULLCHK("ca_map_backward", dout, weight, g, dw, dg);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int rc = energy(dout, g, dw, n, c, h, w, dtype, s, "ca_map_backward(dw)");
  if (rc) return rc;
  return scatter(weight, dout, dg, n, c, h, w, dtype, s, "ca_map_backward(dg)");
}

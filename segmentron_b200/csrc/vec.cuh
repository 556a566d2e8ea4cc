// segb200 -- shared helpers of the memory-bound kernels: dtype-generic loads/stores, 8-wide (128-bit) pack/unpack,
// grid sizing, the torch bilinear index rule.
#pragma once
#include "common.cuh"

namespace segb200 {

__device__ __forceinline__ float load_any(const void* p, long long i, int dtype) {
  if (dtype == DT_F32) return reinterpret_cast<const float*>(p)[i];
  if (dtype == DT_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(reinterpret_cast<const __half*>(p)[i]);
}
__device__ __forceinline__ void store_any(void* p, long long i, float v, int dtype) {
  if (dtype == DT_F32) reinterpret_cast<float*>(p)[i] = v;
  else if (dtype == DT_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
}
__device__ __forceinline__ uint32_t pack_any(float a, float b, int dtype) {
  return dtype == DT_BF16 ? Half2<true>::pack(a, b) : Half2<false>::pack(a, b);
}
__device__ __forceinline__ float2 unpack_any(uint32_t u, int dtype) {
  return dtype == DT_BF16 ? Half2<true>::unpack(u) : Half2<false>::unpack(u);
}
__device__ __forceinline__ float round_any(float v, int dtype) {
  if (dtype == DT_BF16) return __bfloat162float(__float2bfloat16_rn(v));
  if (dtype == DT_F16) return __half2float(__float2half_rn(v));
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& v, int dtype, float (&f)[8]) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float2 t = unpack_any(u[j], dtype); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8], int dtype) {
  return make_uint4(pack_any(f[0], f[1], dtype), pack_any(f[2], f[3], dtype), pack_any(f[4], f[5], dtype),
                    pack_any(f[6], f[7], dtype));
}

static inline int grid_for(long long total, int block) {
  long long b = (total + block - 1) / block;
  const long long cap = 148LL * 32;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

// torch upsample_bilinear2d source index rule (fp32 math, both align_corners modes)
struct Lerp { int i0, i1; float l0, l1; };
__device__ __forceinline__ Lerp lerp_coord(int dst, int in, int out, int align) {
  float src;
  if (align) {
    const float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    src = scale * (float)dst;
  } else {
    const float scale = (float)in / (float)out;
    src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
  }
  Lerp r;
  r.i0 = (int)src;
  if (r.i0 > in - 1) r.i0 = in - 1;
  r.i1 = r.i0 + (r.i0 < in - 1 ? 1 : 0);
  r.l1 = src - (float)r.i0;
  r.l0 = 1.f - r.l1;
  return r;
}

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)
static inline bool half_dt(int d) { return d == DT_BF16 || d == DT_F16; }

}  // namespace segb200

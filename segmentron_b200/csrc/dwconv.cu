// segb200 -- depthwise 3x3 convolution, NHWC, HBM/L2-bound (sm_100a).
//
//   y[n,ho,wo,c] = act( sum_{ky,kx} wgt[ky*3+kx][c] * pre(x[n, ho*s + (ky-1)*d, wo*s + (kx-1)*d, c]) + shift[c] )
//
// One thread owns 8 consecutive channels (one 16-byte vector) of kPix horizontally adjacent output
// pixels; the lanes of a warp cover consecutive channel vectors first (fully coalesced 128-bit
// loads/stores), then pixels.  Pixels are enumerated in 8-row x 16-column tiles so the 3x3 halo is
// re-used out of L1/L2.  BN scale is pre-folded into the fp32 weights; fp32 accumulation.
#include "common.cuh"
#include "../../include/segb200.h"

namespace segb200 {

struct DwParams {
  const void* x; const float* wgt; const float* shift; void* y;
  int n, h, w, c, x_ld, y_ld, ho, wo, stride, dil, pre_relu, act;
  int cv;             // channel vectors (c / 8)
  int tiles_w, tiles_h;
  long long total;    // threads of work: n * tiles_h * tiles_w * (8*16/kPix) * cv
};

constexpr int kTileH = 8, kTileW = 16, kPix = 2;

template <bool kBF16>
__global__ void __launch_bounds__(256)
dwconv3x3_kernel(const DwParams p) {
  using H = Half2<kBF16>;
  using T = typename H::T;
  const T* __restrict__ x = reinterpret_cast<const T*>(p.x);
  T* __restrict__ y = reinterpret_cast<T*>(p.y);
  constexpr int kSlots = kTileH * kTileW / kPix;     // work items per tile per channel vector
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % p.cv);
    long long r = idx / p.cv;
    const int slot = (int)(r % kSlots); r /= kSlots;
    const int tw = (int)(r % p.tiles_w); r /= p.tiles_w;
    const int th = (int)(r % p.tiles_h);
    const int n = (int)(r / p.tiles_h);
    const int ho = th * kTileH + slot / (kTileW / kPix);
    const int wo0 = tw * kTileW + (slot % (kTileW / kPix)) * kPix;
    if (ho >= p.ho || wo0 >= p.wo) continue;
    const int c0 = cv * 8;

    float acc[kPix][8];
#pragma unroll
    for (int q = 0; q < kPix; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;

#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int hi = ho * p.stride + (ky - 1) * p.dil;
      if (hi < 0 || hi >= p.h) continue;
      const T* xrow = x + ((long long)n * p.h + hi) * p.w * p.x_ld + c0;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.wgt + (ky * 3 + kx) * p.c + c0));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.wgt + (ky * 3 + kx) * p.c + c0 + 4));
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int q = 0; q < kPix; ++q) {
          const int wi = (wo0 + q) * p.stride + (kx - 1) * p.dil;
          if (wi < 0 || wi >= p.w || wo0 + q >= p.wo) continue;
          const uint4 v = ldg_v4(xrow + (long long)wi * p.x_ld);
          const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = H::unpack(vv[j]);
            if (p.pre_relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); }
            acc[q][2 * j] = fmaf(f.x, wv[2 * j], acc[q][2 * j]);
            acc[q][2 * j + 1] = fmaf(f.y, wv[2 * j + 1], acc[q][2 * j + 1]);
          }
        }
      }
    }
    float sh[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (p.shift != nullptr) {
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.shift + c0));
      const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + 4));
      sh[0] = s0.x; sh[1] = s0.y; sh[2] = s0.z; sh[3] = s0.w; sh[4] = s1.x; sh[5] = s1.y; sh[6] = s1.z; sh[7] = s1.w;
    }
#pragma unroll
    for (int q = 0; q < kPix; ++q) {
      if (wo0 + q >= p.wo) continue;
      uint4 o;
      o.x = H::pack(apply_act(acc[q][0] + sh[0], p.act), apply_act(acc[q][1] + sh[1], p.act));
      o.y = H::pack(apply_act(acc[q][2] + sh[2], p.act), apply_act(acc[q][3] + sh[3], p.act));
      o.z = H::pack(apply_act(acc[q][4] + sh[4], p.act), apply_act(acc[q][5] + sh[5], p.act));
      o.w = H::pack(apply_act(acc[q][6] + sh[6], p.act), apply_act(acc[q][7] + sh[7], p.act));
      *reinterpret_cast<uint4*>(y + (((long long)n * p.ho + ho) * p.wo + wo0 + q) * p.y_ld + c0) = o;
    }
  }
}

}  // namespace segb200

using namespace segb200;

extern "C" int segb200_dwconv3x3(const segb200_dwconv_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->x || !a->wgt || !a->y) return set_error(-1, "dwconv3x3: null pointer argument");
  if (a->dtype != DT_BF16 && a->dtype != DT_F16) return set_error(-2, "dwconv3x3: dtype must be bf16 or f16");
  if ((a->c & 7) || (a->x_ld & 7) || (a->y_ld & 7) || a->c > a->x_ld || a->c > a->y_ld)
    return set_error(-4, "dwconv3x3: channels/pitches must be multiples of 8");
  if (a->stride < 1 || a->stride > 2 || a->dilation < 1) return set_error(-3, "dwconv3x3: bad stride/dilation");
  if (((uintptr_t)a->x & 15) || ((uintptr_t)a->y & 15) || ((uintptr_t)a->wgt & 15) || ((uintptr_t)a->shift & 15))
    return set_error(-7, "dwconv3x3: pointers must be 16-byte aligned");
  DwParams p;
  p.x = a->x; p.wgt = a->wgt; p.shift = a->shift; p.y = a->y;
  p.n = a->n; p.h = a->h; p.w = a->w; p.c = a->c; p.x_ld = a->x_ld; p.y_ld = a->y_ld;
  p.ho = a->ho; p.wo = a->wo; p.stride = a->stride; p.dil = a->dilation; p.pre_relu = a->pre_relu; p.act = a->act;
  p.cv = a->c / 8;
  p.tiles_w = (a->wo + kTileW - 1) / kTileW;
  p.tiles_h = (a->ho + kTileH - 1) / kTileH;
  p.total = (long long)a->n * p.tiles_h * p.tiles_w * (kTileH * kTileW / kPix) * p.cv;
  if (p.total <= 0) return set_error(-6, "dwconv3x3: empty");
  long long blocks = (p.total + 255) / 256;
  const long long cap = 148LL * 8 * 4;
  if (blocks > cap) blocks = cap;
  if (a->dtype == DT_BF16) dwconv3x3_kernel<true><<<(int)blocks, 256, 0, stream>>>(p);
  else dwconv3x3_kernel<false><<<(int)blocks, 256, 0, stream>>>(p);
  return check_launch("dwconv3x3");
}

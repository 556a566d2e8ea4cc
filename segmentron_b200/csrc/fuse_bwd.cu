// segb200 -- backward of the HRNet fuse step  y = act(a + nearest_up_{2^k}(z))   (forward: segb200_upsample_add, misc.cu;
// reference: the `y = y + fuse_layers[i][j](x[j])` sum of HighResolutionModule.forward, backbones/hrnet.py:215-232, with the
// nn.Upsample(scale_factor=2^(j-i), mode='nearest') of :178-186 and the final ReLU :231).
//   g  = dy * [y > 0]           (act == relu; the mask comes from the stored output)
//   da (+)= g                   at full resolution
//   dz (+)= sum of g over each 2^k x 2^k block (fp32, fixed order)
// One thread per low-resolution pixel and 8-channel vector.  Bound: HBM (reads dy and y once, writes da once).
#include "vec.cuh"
#include "../../include/segb200.h"

namespace segb200 {

__global__ void __launch_bounds__(256)
upsample_add_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ y, void* __restrict__ da, void* __restrict__ dz, int n,
                        int h, int w, int c, int dy_ld, int y_ld, int da_ld, int dz_ld, int k, int act, int acc_a, int acc_z,
                        int dtype) {
  const int cvn = c / 8;
  const int hz = h >> k, wz = w >> k, s = 1 << k;
  const long long total = (long long)n * hz * wz * cvn;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvn);
    long long r = idx / cvn;
    const int xz = (int)(r % wz); r /= wz;
    const int yz = (int)(r % hz);
    const long long b = r / hz;
    float sum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sum[j] = 0.f;
    for (int dyy = 0; dyy < s; ++dyy) {
      for (int dxx = 0; dxx < s; ++dxx) {
        const long long pix = (b * h + (long long)yz * s + dyy) * w + (long long)xz * s + dxx;
        float g[8];
        unpack8(ldg_nc_v4(reinterpret_cast<const char*>(dy) + (pix * dy_ld + cv * 8) * 2), dtype, g);
        if (act != ACT_NONE) {
          float o[8];
          unpack8(ldg_nc_v4(reinterpret_cast<const char*>(y) + (pix * y_ld + cv * 8) * 2), dtype, o);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const bool on = act == ACT_RELU6 ? (o[j] > 0.f && o[j] < 6.f) : (o[j] > 0.f);
            g[j] = on ? g[j] : 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sum[j] += g[j];
        if (da != nullptr) {
          char* ap = reinterpret_cast<char*>(da) + (pix * da_ld + cv * 8) * 2;
          if (acc_a) {
            float a0[8];
            unpack8(*reinterpret_cast<const uint4*>(ap), dtype, a0);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] += a0[j];
          }
          *reinterpret_cast<uint4*>(ap) = pack8(g, dtype);
        }
      }
    }
    if (dz != nullptr) {
      char* zp = reinterpret_cast<char*>(dz) + ((((long long)b * hz + yz) * wz + xz) * dz_ld + cv * 8) * 2;
      if (acc_z) {
        float z0[8];
        unpack8(*reinterpret_cast<const uint4*>(zp), dtype, z0);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum[j] += z0[j];
      }
      *reinterpret_cast<uint4*>(zp) = pack8(sum, dtype);
    }
  }
}

}  // namespace segb200

using namespace segb200;

extern "C" int segb200_upsample_add_bwd(const void* dy, const void* y, void* da, void* dz, int n, int h, int w, int c, int dy_ld,
                                        int y_ld, int da_ld, int dz_ld, int k, int act, int accumulate_a, int accumulate_z,
                                        int dtype, void* stream) {
  if (!dy || (!da && !dz) || (act != ACT_NONE && !y)) return set_error(-1, "upsample_add_bwd: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "upsample_add_bwd: bad dtype");
  if (c < 8 || (c & 7) || (dy_ld & 7) || dy_ld < c || (y && ((y_ld & 7) || y_ld < c)) || (da && ((da_ld & 7) || da_ld < c)) ||
      (dz && ((dz_ld & 7) || dz_ld < c)) || k < 0 || k > 8 || n < 1 || h < 1 || w < 1 || (h & ((1 << k) - 1)) || (w & ((1 << k) - 1)))
    return set_error(-4, "upsample_add_bwd: sizes must be multiples of 8 channels and of the 2^k scale");
  const long long total = (long long)n * (h >> k) * (w >> k) * (c / 8);
  upsample_add_bwd_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(dy, y, da, dz, n, h, w, c, dy_ld, y_ld, da_ld, dz_ld, k, act,
                                                                            accumulate_a ? 1 : 0, accumulate_z ? 1 : 0, dtype);
  return check_launch("upsample_add_bwd");
}

// -------------------------------------------------------------------------------------------
// backward of nn.AdaptiveAvgPool2d(s) (PyramidPooling, modules/module.py:82-97; forward: adaptive_pool_kernel, misc.cu):
// bin (bi, bj) covers rows [floor(bi h / s), ceil((bi+1) h / s)) x the same rule in w (bins overlap when s does not divide the size);
// dx[p] (+)= sum over the bins that contain p of dy[bin] / area(bin).  One thread per input pixel and 8-channel vector (gather).
// -------------------------------------------------------------------------------------------
namespace segb200 {

__global__ void __launch_bounds__(256)
adaptive_pool_bwd_kernel(const void* __restrict__ dy, void* __restrict__ dx, int n, int h, int w, int c, int dy_ld, int dx_ld, int s,
                         int accumulate, int dtype) {
  const int cvn = c / 8;
  const long long total = (long long)n * h * w * cvn;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvn);
    long long r = idx / cvn;
    const int x = (int)(r % w); r /= w;
    const int y = (int)(r % h);
    const long long b = r / h;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int bi = 0; bi < s; ++bi) {
      const int h0 = (bi * h) / s, h1 = ((bi + 1) * h + s - 1) / s;
      if (y < h0 || y >= h1) continue;
      for (int bj = 0; bj < s; ++bj) {
        const int w0 = (bj * w) / s, w1 = ((bj + 1) * w + s - 1) / s;
        if (x < w0 || x >= w1) continue;
        float g[8];
        unpack8(ldg_v4(reinterpret_cast<const char*>(dy) + (((b * s + bi) * s + bj) * dy_ld + cv * 8) * 2), dtype, g);
        const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(g[j], inv, acc[j]);
      }
    }
    char* op = reinterpret_cast<char*>(dx) + (((b * h + y) * w + x) * dx_ld + cv * 8) * 2;
    if (accumulate) {
      float o[8];
      unpack8(*reinterpret_cast<const uint4*>(op), dtype, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += o[j];
    }
    *reinterpret_cast<uint4*>(op) = pack8(acc, dtype);
  }
}

}  // namespace segb200

extern "C" int segb200_adaptive_avgpool_bwd(const void* dy, void* dx, int n, int h, int w, int c, int dy_ld, int dx_ld, int s,
                                            int accumulate, int dtype, void* stream) {
  if (!dy || !dx) return set_error(-1, "adaptive_avgpool_bwd: null pointer");
  if (!half_dt(dtype)) return set_error(-2, "adaptive_avgpool_bwd: bad dtype");
  if (c < 8 || (c & 7) || (dy_ld & 7) || (dx_ld & 7) || dy_ld < c || dx_ld < c || s < 1 || n < 1 || h < 1 || w < 1)
    return set_error(-4, "adaptive_avgpool_bwd: bad sizes");
  adaptive_pool_bwd_kernel<<<grid_for((long long)n * h * w * (c / 8), 256), 256, 0, STREAM(stream)>>>(dy, dx, n, h, w, c, dy_ld, dx_ld, s,
                                                                                                  accumulate ? 1 : 0, dtype);
  return check_launch("adaptive_avgpool_bwd");
}

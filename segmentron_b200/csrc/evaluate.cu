// segb200 -- the two data-movement steps of the multi-scale + flip evaluation driver.
//
// SegBaseModel.evaluate (segmentron/models/segbase.py:44-79) runs, per scale, F.interpolate(image) -> F.pad -> forward ->
// [flip -> forward -> flip -> +=] -> crop -> F.interpolate(outputs) -> scores +=   as separate torch ops (9 passes over the
// image / logits).  Here:
//   eval_prepare    : image -> [resized + zero padded ; horizontally flipped copy of the same]   (one pass, both forward
//                     inputs stacked in one batch so that the model runs ONCE on 2B images);
//   eval_accumulate : scores (+)= resize( crop( logits[0:B] + flip(logits[B:2B]) ) )            (one pass; the flipped half is
//                     read through mirrored indices, the sum is formed at the four bilinear taps on the fly).
// Both are HBM-bound element-wise kernels (coalesced along x).  Bilinear rule = torch's align_corners=True rule (vec.cuh
// lerp_coord), the arithmetic order of upsample_bilinear2d; every place where the reference materialises a tensor in the model
// dtype (outputs +=, the resized score, scores +=) rounds to that dtype here too.
#include "vec.cuh"
#include "../../include/segb200.h"

namespace segb200 {

__global__ void __launch_bounds__(256)
eval_prepare_kernel(const float* __restrict__ image, float* __restrict__ out, int b, int c, int h, int w, int height, int width,
                    int hp, int wp, int flip) {
  const long long total = (long long)b * c * hp * wp;
  const long long half = total;                                   // element offset of the flipped half
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % wp);
    long long r = idx / wp;
    const int y = (int)(r % hp);
    const long long bc = r / hp;                                  // image * c + channel
    float v = 0.f;
    if (y < height && x < width) {
      const Lerp ly = lerp_coord(y, h, height, 1), lx = lerp_coord(x, w, width, 1);
      const float* p = image + bc * h * w;
      const float v00 = __ldg(p + (long long)ly.i0 * w + lx.i0), v01 = __ldg(p + (long long)ly.i0 * w + lx.i1);
      const float v10 = __ldg(p + (long long)ly.i1 * w + lx.i0), v11 = __ldg(p + (long long)ly.i1 * w + lx.i1);
      v = ly.l0 * (lx.l0 * v00 + lx.l1 * v01) + ly.l1 * (lx.l0 * v10 + lx.l1 * v11);
    }
    out[idx] = v;
    if (flip) out[half + (bc * hp + y) * wp + (wp - 1 - x)] = v;
  }
}

__global__ void __launch_bounds__(256)
eval_accumulate_kernel(const void* __restrict__ logits, void* __restrict__ scores, int dtype, int b, int k, int hp, int wp,
                       int height, int width, int h, int w, int flip, int accumulate) {
  const long long total = (long long)b * k * h * w;
  const long long plane = (long long)hp * wp;
  const long long half = (long long)b * k * plane;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % w);
    long long r = idx / w;
    const int y = (int)(r % h);
    const long long bk = r / h;
    const Lerp ly = lerp_coord(y, height, h, 1), lx = lerp_coord(x, width, w, 1);
    const long long base = bk * plane;
    float s[4];
    const int yy[2] = {ly.i0, ly.i1}, xx[2] = {lx.i0, lx.i1};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int sy = yy[q >> 1], sx = xx[q & 1];
      float v = load_any(logits, base + (long long)sy * wp + sx, dtype);
      if (flip) v = round_any(v + load_any(logits, half + base + (long long)sy * wp + (wp - 1 - sx), dtype), dtype);
      s[q] = v;
    }
    float o = ly.l0 * (lx.l0 * s[0] + lx.l1 * s[1]) + ly.l1 * (lx.l0 * s[2] + lx.l1 * s[3]);
    o = round_any(o, dtype);
    if (accumulate) o = round_any(load_any(scores, idx, dtype) + o, dtype);
    store_any(scores, idx, o, dtype);
  }
}

}  // namespace segb200

using namespace segb200;

extern "C" int segb200_eval_prepare(const float* image, float* out, int b, int c, int h, int w, int height, int width, int hp,
                                    int wp, int flip, void* stream) {
  if (!image || !out) return set_error(-1, "eval_prepare: null pointer");
  if (b < 1 || c < 1 || h < 1 || w < 1 || height < 1 || width < 1 || hp < height || wp < width)
    return set_error(-4, "eval_prepare: sizes must be positive and the padded size must cover the resized image");
  eval_prepare_kernel<<<grid_for((long long)b * c * hp * wp, 256), 256, 0, STREAM(stream)>>>(image, out, b, c, h, w, height, width,
                                                                                          hp, wp, flip ? 1 : 0);
  return check_launch("eval_prepare");
}

extern "C" int segb200_eval_accumulate(const void* logits, void* scores, int dtype, int b, int k, int hp, int wp, int height,
                                       int width, int h, int w, int flip, int accumulate, void* stream) {
  if (!logits || !scores) return set_error(-1, "eval_accumulate: null pointer");
  if (dtype < 0 || dtype > 2) return set_error(-2, "eval_accumulate: bad dtype");
  if (b < 1 || k < 1 || h < 1 || w < 1 || height < 1 || width < 1 || hp < height || wp < width)
    return set_error(-4, "eval_accumulate: sizes must be positive and the crop must lie inside the logits");
  eval_accumulate_kernel<<<grid_for((long long)b * k * h * w, 256), 256, 0, STREAM(stream)>>>(
      logits, scores, dtype, b, k, hp, wp, height, width, h, w, flip ? 1 : 0, accumulate ? 1 : 0);
  return check_launch("eval_accumulate");
}

// segb200 -- evaluation metric counts on the device.
//
// Replaces segmentron/utils/score.py:83-113 (batch_pix_accuracy, batch_intersection_union), which per batch runs two argmax
// passes over the full-resolution logits, moves three float maps to the host and calls torch.histc three times behind a
// torch.cuda.synchronize() (score.py:49,108-110).  Here one pass over the logits produces all integer counts; the totals stay
// on the device and nothing synchronises until get().
//
//   counts (unsigned 64-bit, length 2 + 3*nclass, ACCUMULATED):
//     [0] correct   = #{label >= 0 and argmax_c trunc(logit_c) == label}          (score.py:86-90: the logits are truncated
//                                                                                  to integers before that argmax)
//     [1] labeled   = #{label >= 0}
//     [2 + c]            inter[c] = #{label >= 0, argmax_c logit_c == c == label}  (score.py:102-109)
//     [2 + nclass + c]   pred[c]  = #{label >= 0, argmax_c logit_c == c}           (:105,:110)
//     [2 + 2 nclass + c] lab[c]   = #{label == c}                                   (:111; labels >= nclass are outside histc's range)
//   ties -> lowest class index (torch.argmax).  Integer sums: the result does not depend on the order of the atomics.
//
// Two sources: full-resolution NCHW logits (what the reference's metric is given), or the low-resolution NHWC logits of the
// classifier with the final bilinear up-sampling fused (same arithmetic and output rounding as bilinear_nchw_out_kernel, misc.cu),
// so that the [N,19,H,W] tensor is never written when only the metric is wanted.
// Bound: HBM (4..76 bytes per pixel read once; the counting is warp-aggregated shared-memory atomics).
#include <climits>

#include "vec.cuh"
#include "../../include/segb200.h"

namespace segb200 {

constexpr int kMetricMaxClasses = 64;

// every lane of the warp calls this; lanes with key < 0 do not count
__device__ __forceinline__ void warp_bin_add(unsigned int* bins, int key, int lane) {
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  if (key >= 0 && (__ffs(peers) - 1) == lane) atomicAdd(bins + key, (unsigned int)__popc(peers));
}

template <bool kFused>
__global__ void __launch_bounds__(256)
seg_metric_kernel(const void* __restrict__ logits, const long long* __restrict__ target, unsigned long long* __restrict__ counts,
                  int n, int nclass, int ho, int wo, int dtype,
                  // fused source only: low-res NHWC geometry
                  int hi, int wi, int x_ld, int align, int out_dtype) {
  __shared__ unsigned int sh[2 + 3 * kMetricMaxClasses];
  const int nbins = 2 + 3 * nclass;
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long plane = (long long)ho * wo;
  const long long total = (long long)n * plane;
  unsigned int* inter = sh + 2;
  unsigned int* pred = inter + nclass;
  unsigned int* lab = pred + nclass;
  // uniform trip count per block: every lane reaches the warp-synchronous calls
  for (long long base = (long long)blockIdx.x * blockDim.x; base < total; base += (long long)gridDim.x * blockDim.x) {
    const long long idx = base + threadIdx.x;
    const bool in = idx < total;
    float best = -INFINITY; int besti = 0;
    long long bestt = LLONG_MIN; int bestti = 0;
    long long tg = -1;
    if (in) {
      tg = __ldg(target + idx);
      const long long pix = idx % plane;
      const int b = (int)(idx / plane);
      if (!kFused) {
        for (int ch = 0; ch < nclass; ++ch) {
          const float o = load_any(logits, ((long long)b * nclass + ch) * plane + pix, dtype);
          if (o > best) { best = o; besti = ch; }
          const long long t = (long long)o;                                  // truncation toward zero == Tensor.long()
          if (t > bestt) { bestt = t; bestti = ch; }
        }
      } else {
        const int ox = (int)(pix % wo), oy = (int)(pix / wo);
        const Lerp ly = lerp_coord(oy, hi, ho, align), lx = lerp_coord(ox, wi, wo, align);
        const char* img = reinterpret_cast<const char*>(logits) + (long long)b * hi * wi * x_ld * 2;
        const char* p00 = img + ((long long)ly.i0 * wi + lx.i0) * x_ld * 2;
        const char* p01 = img + ((long long)ly.i0 * wi + lx.i1) * x_ld * 2;
        const char* p10 = img + ((long long)ly.i1 * wi + lx.i0) * x_ld * 2;
        const char* p11 = img + ((long long)ly.i1 * wi + lx.i1) * x_ld * 2;
        for (int cv = 0; cv * 8 < nclass; ++cv) {
          float f00[8], f01[8], f10[8], f11[8];
          unpack8(ldg_v4(p00 + cv * 16), dtype, f00);
          unpack8(ldg_v4(p01 + cv * 16), dtype, f01);
          unpack8(ldg_v4(p10 + cv * 16), dtype, f10);
          unpack8(ldg_v4(p11 + cv * 16), dtype, f11);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int ch = cv * 8 + j;
            if (ch < nclass) {
              float o = ly.l0 * (lx.l0 * f00[j] + lx.l1 * f01[j]) + ly.l1 * (lx.l0 * f10[j] + lx.l1 * f11[j]);
              o = round_any(o, out_dtype);
              if (o > best) { best = o; besti = ch; }
              const long long t = (long long)o;
              if (t > bestt) { bestt = t; bestti = ch; }
            }
          }
        }
      }
    }
    const bool valid = in && tg >= 0;
    const unsigned m_valid = __ballot_sync(0xffffffffu, valid);
    const unsigned m_correct = __ballot_sync(0xffffffffu, valid && (long long)bestti == tg);
    if (lane == 0) {
      if (m_valid) atomicAdd(sh + 1, (unsigned int)__popc(m_valid));
      if (m_correct) atomicAdd(sh + 0, (unsigned int)__popc(m_correct));
    }
    if (m_valid) {                                                            // warp-uniform
      warp_bin_add(pred, valid ? besti : -1, lane);
      warp_bin_add(inter, (valid && (long long)besti == tg) ? besti : -1, lane);
      warp_bin_add(lab, (valid && tg < nclass) ? (int)tg : -1, lane);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += blockDim.x)
    if (sh[i]) atomicAdd(counts + i, (unsigned long long)sh[i]);
}

// SegmentationMetric.update's accumulation (score.py:50-55) on the device: the two pixel totals are exact integers, the per-class
// totals are float32 and are advanced by float32(inter), float32(pred) + float32(lab) - float32(inter) (score.py:112) once per
// update, like the reference's `total_inter += inter`.  The batch counts are cleared for the next update.
__global__ void __launch_bounds__(128)
seg_metric_accumulate_kernel(unsigned long long* __restrict__ counts, int nclass, long long* __restrict__ total_pixels,
                             float* __restrict__ total_inter, float* __restrict__ total_union) {
  const int i = threadIdx.x;
  if (i < 2) total_pixels[i] += (long long)counts[i];
  if (i < nclass) {
    const float a_inter = (float)counts[2 + i];
    const float a_pred = (float)counts[2 + nclass + i];
    const float a_lab = (float)counts[2 + 2 * nclass + i];
    total_inter[i] += a_inter;
    total_union[i] += (a_pred + a_lab) - a_inter;
  }
  __syncthreads();
  for (int j = i; j < 2 + 3 * nclass; j += blockDim.x) counts[j] = 0ull;
}

}  // namespace segb200

using namespace segb200;

extern "C" int segb200_seg_metric(const void* logits, int dtype, const long long* target, int n, int nclass, int h, int w,
                                  unsigned long long* counts, void* stream) {
  if (!logits || !target || !counts) return set_error(-1, "seg_metric: null pointer");
  if (dtype < 0 || dtype > 2) return set_error(-2, "seg_metric: bad dtype");
  if (nclass < 1 || nclass > kMetricMaxClasses || n < 1 || h < 1 || w < 1)
    return set_error(-4, "seg_metric: 1 <= nclass <= 64 and a non-empty batch");
  seg_metric_kernel<false><<<grid_for((long long)n * h * w, 256), 256, 0, STREAM(stream)>>>(logits, target, counts, n, nclass, h, w,
                                                                                         dtype, 0, 0, 0, 0, 0);
  return check_launch("seg_metric");
}

extern "C" int segb200_seg_metric_lowres(const void* logits_nhwc, int dtype, int x_ld, int hi, int wi, int align_corners,
                                         int out_dtype, const long long* target, int n, int nclass, int ho, int wo,
                                         unsigned long long* counts, void* stream) {
  if (!logits_nhwc || !target || !counts) return set_error(-1, "seg_metric_lowres: null pointer");
  if (!half_dt(dtype) || out_dtype < 0 || out_dtype > 2) return set_error(-2, "seg_metric_lowres: bad dtype");
  if (nclass < 1 || nclass > kMetricMaxClasses || (x_ld & 7) || ((nclass + 7) & ~7) > x_ld || n < 1 || hi < 1 || wi < 1 || ho < 1 ||
      wo < 1)
    return set_error(-4, "seg_metric_lowres: 1 <= nclass <= 64, x_ld a multiple of 8 and >= round_up(nclass, 8)");
  seg_metric_kernel<true><<<grid_for((long long)n * ho * wo, 256), 256, 0, STREAM(stream)>>>(
      logits_nhwc, target, counts, n, nclass, ho, wo, dtype, hi, wi, x_ld, align_corners, out_dtype);
  return check_launch("seg_metric_lowres");
}

extern "C" int segb200_seg_metric_accumulate(unsigned long long* counts, int nclass, long long* total_pixels, float* total_inter,
                                             float* total_union, void* stream) {
  if (!counts || !total_pixels || !total_inter || !total_union) return set_error(-1, "seg_metric_accumulate: null pointer");
  if (nclass < 1 || nclass > kMetricMaxClasses) return set_error(-4, "seg_metric_accumulate: 1 <= nclass <= 64");
  seg_metric_accumulate_kernel<<<1, 128, 0, STREAM(stream)>>>(counts, nclass, total_pixels, total_inter, total_union);
  return check_launch("seg_metric_accumulate");
}

// segb200 -- GPU-side input transform: transforms.ToTensor() + transforms.Normalize(mean, std) of the reference's
// data pipeline (tools/train.py:36-39, tools/eval.py:33-36) for a batch of decoded uint8 HWC images already on the device:
//   out[n][c][y][x] = (img[n][y][x][c] / 255 - mean[c]) / std[c]        (fp32, the same two IEEE divisions and one subtraction
//                                                                          torchvision performs, so the result is bit-identical)
// One thread per output element (coalesced NCHW writes; the 3-byte-strided reads are served by L1).  Bound: HBM (1 B in, 4 B out).
#include "vec.cuh"
#include "../../include/segb200.h"

namespace segb200 {

__global__ void __launch_bounds__(256)
image_normalize_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, int n, int h, int w, int c,
                       const float* __restrict__ mean, const float* __restrict__ stdv) {
  const long long plane = (long long)h * w;
  const long long total = (long long)n * c * plane;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long pix = idx % plane;
    const long long r = idx / plane;
    const int ch = (int)(r % c);
    const long long b = r / c;
    const float u = (float)img[(b * plane + pix) * c + ch];
    out[idx] = __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.f), __ldg(mean + ch)), __ldg(stdv + ch));
  }
}

}  // namespace segb200

using namespace segb200;

extern "C" int segb200_image_normalize(const unsigned char* img, float* out, int n, int h, int w, int c, const float* mean,
                                       const float* stdv, void* stream) {
  if (!img || !out || !mean || !stdv) return set_error(-1, "image_normalize: null pointer");
  if (n < 1 || h < 1 || w < 1 || c < 1 || c > 4) return set_error(-4, "image_normalize: 1..4 channels, non-empty batch");
  image_normalize_kernel<<<grid_for((long long)n * c * h * w, 256), 256, 0, STREAM(stream)>>>(img, out, n, h, w, c, mean, stdv);
  return check_launch("image_normalize");
}

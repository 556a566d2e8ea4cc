// segb200 -- convolution weight gradient as a pixel-reduction GEMM on tcgen05 tensor cores (sm_100a).
//
//   dW[cout, tap, cin] += sum_{pixel} dY[pixel, cout] * X[pixel shifted by tap, cin]
//
// Both operands are consumed straight from their NHWC activations: a TMA box {64 channels, BW, BH, 1} (BW*BH = 64
// pixels) lands in 128B-swizzled shared memory as 64 rows (pixels = the GEMM K dimension) of 128 bytes (64 channels =
// the GEMM M / N dimension), which is exactly the canonical *MN-major* UMMA operand layout
//   Swizzle<3,4,3> o ((8,n),(8,k)) : ((1,LBO),(8,SBO))   [16-byte units; LBO = one box = 8 KB, SBO = 8 rows = 1 KB]
// so no transposed copy of the activations is ever made.  The conv zero padding (and the image borders of the pixel
// tile) are TMA out-of-bounds zero fill on both operands; stride-2 convs read parity views of X exactly like the forward
// kernel does.
//
// Work item = (128-cout tile, <=256-cin tile, tap, pixel split); persistent CTAs, warp-specialised like conv_gemm:
//   warp 0 TMA producer (ring of stages {A: 2 boxes, B: up to 4 boxes}), warp 1 TMEM allocator + MMA issuer (M = 128,
//   N = 64..256, K = 16 pixels per instruction, fp32 accumulate, two accumulator stages), warps 2..5 epilogue:
//   tcgen05.ld -> red.global.add.v4.f32 into the fp32 gradient (split-K partial sums meet in L2).
// Items are ordered so that CTAs running concurrently share a pixel split: every dY / X tile is fetched from HBM once
// and re-read from L2 by the other (cout tile, cin tile, tap) items.
#include "common.cuh"
#include "../../include/segb200.h"

#include <mutex>
#include <stddef.h>

namespace segb200 {

constexpr int kWgRing = 196608;
constexpr int kWgBox = 8192;              // 64 pixels x 128 B
constexpr int kWgMaxStages = 12;
constexpr int kWgCtl = 512;
constexpr int kWgSmem = kWgRing + kWgCtl;

struct WgControl {
  uint64_t full[kWgMaxStages];
  uint64_t empty[kWgMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(WgControl) <= kWgCtl, "control block too large");

struct WgradParams {
  int bw, bh, wtiles, htiles;
  int bi;                      // images per pixel tile (the TMA box spans bi images)
  long long pix_tiles;
  int cout, cin, ntaps;
  int co_tiles, ci_tiles, bn;
  int splits, total_items;
  int num_stages, stage_bytes;
  int lbo_sbo_swap;            // diagnostics: exchange the LBO / SBO fields of the operand descriptors
  float* dw;
  uint32_t taps[64];           // map id (bits 0..1) | (off_w + 128) << 8 | (off_h + 128) << 16
};

// MN-major shared-memory matrix descriptor, 128B swizzle (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::f16: fp32 accumulate, MN-major A and B (bits 15, 16), M = 128
__device__ __forceinline__ uint32_t make_idesc_mn(bool bf16, uint32_t n) {
  const uint32_t fmt = bf16 ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct WgItem { int co0, ci0, tap, na, nb; long long pt0, pt1; };

__device__ __forceinline__ WgItem wg_decode(const WgradParams& p, int item) {
  WgItem it;
  const int co = item % p.co_tiles; item /= p.co_tiles;
  const int ci = item % p.ci_tiles; item /= p.ci_tiles;
  it.tap = item % p.ntaps;
  const int split = item / p.ntaps;
  it.co0 = co * 128; it.ci0 = ci * p.bn;
  it.na = (p.cout - it.co0 > 64) ? 2 : 1;
  int nb = (p.cin - it.ci0 + 63) >> 6;
  if (nb > (p.bn >> 6)) nb = p.bn >> 6;
  it.nb = nb;
  it.pt0 = p.pix_tiles * split / p.splits;
  it.pt1 = p.pix_tiles * (split + 1) / p.splits;
  return it;
}

template <bool kBF16>
__global__ void __launch_bounds__(192, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
                  const __grid_constant__ CUtensorMap tmX2, const __grid_constant__ CUtensorMap tmX3,
                  const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  WgControl* ctl = reinterpret_cast<WgControl*>(smem + kWgRing);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("segb200: dynamic smem base not 1024B aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmX0); prefetch_tmap(&tmDY); }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < p.num_stages; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&ctl->tmem_full[i], 1); mbar_init(&ctl->tmem_empty[i], 128); }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&ctl->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const long long tiles_per_img = (long long)p.wtiles * p.htiles;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      const CUtensorMap* xmaps[4] = {&tmX0, &tmX1, &tmX2, &tmX3};
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const WgItem it = wg_decode(p, item);
        const uint32_t t = p.taps[it.tap];
        const int ow = (int)((t >> 8) & 0xff) - 128, oh = (int)((t >> 16) & 0xff) - 128;
        const CUtensorMap* xm = xmaps[t & 3];
        const uint32_t tx = (uint32_t)((it.na + it.nb) * kWgBox);
        for (long long pt = it.pt0; pt < it.pt1; ++pt) {
          const int img = (int)(pt / tiles_per_img);
          const int rem = (int)(pt - (long long)img * tiles_per_img);
          const int hb = rem / p.wtiles, wb = rem - hb * p.wtiles;
          const int w0 = wb * p.bw, h0 = hb * p.bh;
          mbar_wait(&ctl->empty[stage], phase ^ 1);
          mbar_expect_tx(&ctl->full[stage], tx);
          uint8_t* sa = smem + stage * p.stage_bytes;
          for (int i = 0; i < it.na; ++i)
            tma_load_4d(&tmDY, &ctl->full[stage], sa + i * kWgBox, it.co0 + i * 64, w0, h0, img * p.bi);
          for (int j = 0; j < it.nb; ++j)
            tma_load_4d(xm, &ctl->full[stage], sa + 2 * kWgBox + j * kWgBox, it.ci0 + j * 64, w0 + ow, h0 + oh, img * p.bi);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      const uint32_t lbo = p.lbo_sbo_swap ? 1024u : (uint32_t)kWgBox;
      const uint32_t sbo = p.lbo_sbo_swap ? (uint32_t)kWgBox : 1024u;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const WgItem it = wg_decode(p, item);
        const uint32_t idesc = make_idesc_mn(kBF16, (uint32_t)(it.nb * 64));
        mbar_wait(&ctl->tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
        uint32_t first = 1;
        for (long long pt = it.pt0; pt < it.pt1; ++pt) {
          mbar_wait(&ctl->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * p.stage_bytes);
          const uint32_t sb = sa + 2 * kWgBox;
#pragma unroll
          for (int k = 0; k < 4; ++k) {              // 64 pixels per stage = 4 x (K = 16); 16 rows x 128 B = 2 KB per step
            const uint64_t adesc = make_mnmajor_desc(sa + (uint32_t)(k * 2048), lbo, sbo);
            const uint64_t bdesc = make_mnmajor_desc(sb + (uint32_t)(k * 2048), lbo, sbo);
            umma_f16(d_tmem, adesc, bdesc, idesc, first ? 0u : 1u);
            first = 0;
          }
          umma_commit(&ctl->empty[stage]);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&ctl->tmem_full[acc]);
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue (warps 2..5) ------------------------------
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const WgItem it = wg_decode(p, item);
      mbar_wait(&ctl->tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + (uint32_t)(acc * 256) + ((uint32_t)(q * 32) << 16);
      const int co = it.co0 + row;
      const bool row_ok = co < p.cout && it.pt1 > it.pt0;
      float* drow = p.dw + ((long long)co * p.ntaps + it.tap) * p.cin;
      for (int ch = 0; ch < it.nb * 2; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32(t_acc + (uint32_t)(ch * 32), v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int ci = it.ci0 + ch * 32 + g * 4;
            if (ci < p.cin)
              red_add_v4(drow + ci, __uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]), __uint_as_float(v[g * 4 + 2]),
                         __uint_as_float(v[g * 4 + 3]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&ctl->tmem_empty[acc]);
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static inline int wg_floordiv2(int o, int* parity) {
  const int p = ((o % 2) + 2) % 2;
  *parity = p;
  return (o - p) / 2;
}

}  // namespace segb200

using namespace segb200;

static int g_wg_swap = 0;
extern "C" int segb200_wgrad_debug_swap(int v) { g_wg_swap = v ? 1 : 0; return 0; }

extern "C" int segb200_conv_wgrad(const segb200_wgrad_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->x || !a->dy || !a->dw) return set_error(-1, "conv_wgrad: null pointer argument");
  if (a->dtype != DT_BF16 && a->dtype != DT_F16) return set_error(-2, "conv_wgrad: dtype must be bf16 or f16");
  if (a->stride != 1 && a->stride != 2) return set_error(-3, "conv_wgrad: stride must be 1 or 2");
  if ((a->x_ld & 7) || (a->dy_ld & 7) || (a->cin & 7) || a->cin < 8 || a->cout < 1 || a->cin > a->x_ld || a->cout > a->dy_ld)
    return set_error(-4, "conv_wgrad: cin and the pitches must be multiples of 8 elements (cin %d x_ld %d cout %d dy_ld %d)",
                     a->cin, a->x_ld, a->cout, a->dy_ld);
  const int ntaps = a->kh * a->kw;
  if (ntaps < 1 || ntaps > 64) return set_error(-5, "conv_wgrad: unsupported kernel %dx%d", a->kh, a->kw);
  if (a->n < 1 || a->ho < 1 || a->wo < 1) return set_error(-6, "conv_wgrad: empty output");
  if (((uintptr_t)a->x & 15) || ((uintptr_t)a->dy & 15) || ((uintptr_t)a->dw & 15))
    return set_error(-7, "conv_wgrad: pointers must be 16-byte aligned");

  WgradParams p;
  memset(&p, 0, sizeof(p));
  const bool flat = (ntaps == 1 && a->stride == 1 && a->pad_t == 0 && a->pad_l == 0 && a->ho == a->h && a->wo == a->w);
  long long Wv, Hv, Nv;
  if (flat) {
    Wv = (long long)a->n * a->h * a->w; Hv = 1; Nv = 1;
    p.bw = 64; p.bh = 1; p.bi = 1;
  } else {
    Wv = a->wo; Hv = a->ho; Nv = a->n;
    // 64 pixels per K step as a BW x BH patch of BI images (BW*BH*BI = 64): spanning images cuts the padding of small maps
    long long best = -1;
    for (int bi = 1; bi <= 8 && bi <= a->n * 2 - 1; bi *= 2)
      for (int bw = 64 / bi; bw >= 1; bw /= 2) {
        const int bh = 64 / bi / bw;
        const long long t = (long long)((a->wo + bw - 1) / bw) * ((a->ho + bh - 1) / bh) * ((a->n + bi - 1) / bi);
        if (best < 0 || t < best) { best = t; p.bw = bw; p.bh = bh; p.bi = bi; }
      }
  }
  if (Wv > 0x7fffffffLL) return set_error(-8, "conv_wgrad: too many pixels");
  p.wtiles = (int)((Wv + p.bw - 1) / p.bw);
  p.htiles = (int)((Hv + p.bh - 1) / p.bh);
  p.pix_tiles = (long long)p.wtiles * p.htiles * ((Nv + p.bi - 1) / p.bi);
  p.cout = a->cout; p.cin = a->cin; p.ntaps = ntaps;
  p.co_tiles = (a->cout + 127) / 128;
  p.bn = a->cin >= 256 ? 256 : ((a->cin + 63) & ~63);
  p.ci_tiles = (a->cin + p.bn - 1) / p.bn;
  const long long base_items = (long long)p.co_tiles * p.ci_tiles * ntaps;
  // pixel splits: fill the machine in whole waves
  const int sms = a->max_ctas > 0 ? a->max_ctas : num_sms();
  long long max_splits = p.pix_tiles / 4; if (max_splits < 1) max_splits = 1;
  int splits = a->splits;
  if (splits <= 0) {
    double best_eff = -1.0; splits = 1;
    for (int k = 1; k <= 4; ++k) {
      long long s = (long long)sms * k / base_items;
      if (s < 1) s = 1;
      if (s > max_splits) s = max_splits;
      const long long items = base_items * s;
      const double eff = (double)items / (double)(((items + sms - 1) / sms) * sms);
      if (eff > best_eff + 1e-9) { best_eff = eff; splits = (int)s; }
    }
  }
  if (splits > max_splits) splits = (int)max_splits;
  if (splits > p.pix_tiles) splits = (int)p.pix_tiles;
  p.splits = splits;
  const long long total = base_items * splits;
  if (total > 0x7fffffffLL) return set_error(-8, "conv_wgrad: too many work items");
  p.total_items = (int)total;
  p.stage_bytes = (2 + p.bn / 64) * kWgBox;
  p.num_stages = kWgRing / p.stage_bytes;
  if (p.num_stages > kWgMaxStages) p.num_stages = kWgMaxStages;
  p.lbo_sbo_swap = g_wg_swap;
  p.dw = a->dw;

  CUtensorMap tmX[4], tmDY;
  memset(tmX, 0, sizeof(tmX));
  bool used[4] = {false, false, false, false};
  for (int ky = 0; ky < a->kh; ++ky)
    for (int kx = 0; kx < a->kw; ++kx) {
      const int oh = ky * a->dilation - a->pad_t, ow = kx * a->dilation - a->pad_l;
      int ph = 0, pw = 0, ah = oh, aw = ow;
      if (a->stride == 2) { ah = wg_floordiv2(oh, &ph); aw = wg_floordiv2(ow, &pw); }
      if (ah < -128 || ah > 127 || aw < -128 || aw > 127) return set_error(-9, "conv_wgrad: tap offset out of range");
      const int mid = ph * 2 + pw;
      used[mid] = true;
      p.taps[ky * a->kw + kx] = (uint32_t)mid | ((uint32_t)(aw + 128) << 8) | ((uint32_t)(ah + 128) << 16);
    }
  const char* xb = reinterpret_cast<const char*>(a->x);
  const uint32_t box[4] = {64u, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bi};
  for (int mid = 0; mid < 4; ++mid) {
    if (!used[mid]) continue;
    int rc;
    if (flat) {
      const uint64_t dims[4] = {(uint64_t)a->cin, (uint64_t)Wv, 1, 1};
      const uint64_t str[3] = {(uint64_t)a->x_ld * 2, (uint64_t)a->x_ld * 2 * (uint64_t)Wv, (uint64_t)a->x_ld * 2 * (uint64_t)Wv};
      rc = encode_map(&tmX[mid], a->dtype, 4, xb, dims, str, box, 128, "wgrad X/flat");
    } else {
      const int s = a->stride, ph = mid >> 1, pw = mid & 1;
      if (ph >= a->h || pw >= a->w) return set_error(-9, "conv_wgrad: input too small for stride-2 parity view");
      const uint64_t dims[4] = {(uint64_t)a->cin, (uint64_t)((a->w - pw + s - 1) / s), (uint64_t)((a->h - ph + s - 1) / s),
                                (uint64_t)a->n};
      const uint64_t str[3] = {(uint64_t)a->x_ld * 2 * s, (uint64_t)a->x_ld * 2 * a->w * s, (uint64_t)a->x_ld * 2 * a->w * a->h};
      rc = encode_map(&tmX[mid], a->dtype, 4, xb + ((long long)ph * a->w + pw) * a->x_ld * 2, dims, str, box, 128,
                      "wgrad X/patch");
    }
    if (rc) return rc;
  }
  for (int mid = 0; mid < 4; ++mid)
    if (!used[mid]) for (int j = 0; j < 4; ++j) if (used[j]) { tmX[mid] = tmX[j]; break; }
  {
    const uint64_t dims[4] = {(uint64_t)a->cout, (uint64_t)Wv, (uint64_t)Hv, (uint64_t)Nv};
    const uint64_t str[3] = {(uint64_t)a->dy_ld * 2, (uint64_t)a->dy_ld * 2 * (uint64_t)Wv,
                             (uint64_t)a->dy_ld * 2 * (uint64_t)Wv * (uint64_t)Hv};
    int rc = encode_map(&tmDY, a->dtype, 4, a->dy, dims, str, box, 128, "wgrad dY");
    if (rc) return rc;
  }

  int grid = sms;
  if (grid > p.total_items) grid = p.total_items;
  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    cudaFuncSetAttribute(conv_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem);
    cudaFuncSetAttribute(conv_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem);
  });
  if (a->dtype == DT_BF16)
    conv_wgrad_kernel<true><<<grid, 192, kWgSmem, stream>>>(tmX[0], tmX[1], tmX[2], tmX[3], tmDY, p);
  else
    conv_wgrad_kernel<false><<<grid, 192, kWgSmem, stream>>>(tmX[0], tmX[1], tmX[2], tmX[3], tmDY, p);
  return check_launch("conv_wgrad");
}

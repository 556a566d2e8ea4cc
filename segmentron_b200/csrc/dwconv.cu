// segb200 -- depthwise 3x3 convolution, NHWC, HBM-bound (sm_100a).
//
//   y[n,ho,wo,c] = act( sum_{ky,kx} wgt[ky*3+kx][c] * pre(x[n, ho*s + (ky-1)*d, wo*s + (kx-1)*d, c]) + shift[c] )
//
// Row-streaming design: a thread owns one output column `wo` and 8 consecutive channels (one 16-byte vector)
// and walks DOWN the image.  For every input row it touches it loads the three horizontal taps once
// (3 x 128-bit loads), converts them to fp32 once and forms the three per-kernel-row partial sums
//        s_ky(r) = sum_kx w[ky][kx] * x[r][wo*s + (kx-1)*d]
// so that  y[h] = s_0(h*s - d) + s_1(h*s) + s_2(h*s + d)  is assembled from two rolling accumulators:
// each input element is loaded 3x (from L1: the 16x8 thread block shares its horizontal halo) instead of
// 9x, and converted once.  The 72 folded weights of the thread's 8 channels live in registers for the whole
// walk.  Dilation d>1 (stride 1) is handled as d interleaved row chains (rows h0, h0+d, h0+2d, ...).
// Math is packed fp32x2 FMA (fma.rn.f32x2), fp32 accumulation, BN scale pre-folded into the weights.
#include "common.cuh"
#include "../../include/segb200.h"

#include <mutex>

namespace segb200 {

struct DwParams {
  const void* x; const float* wgt; const float* shift; void* y;
  int n, h, w, c, x_ld, y_ld, ho, wo, stride, dil, pre_relu, act;
  int cv;              // channel vectors (c / 8)
  int lc;              // channel-vector lanes per block (8 or 16); 128 / lc column lanes
  int cblocks;         // ceil(cv / lc)
  int rows_per_block;  // output rows per blockIdx.y segment
};

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

template <bool kBF16>
struct RowSums {
  float2 s[3][4];   // [ky][channel pair]
};

// partial sums of one input row for the thread's column; xrow = &x[n][r][0][c0] or nullptr for a padded row
template <bool kBF16>
__device__ __forceinline__ void row_sums(const typename Half2<kBF16>::T* xrow, int wi0, int wstep, int w, int x_ld,
                                         bool pre_relu, const float2 (&wt)[9][4], float2 (&s)[3][4]) {
  using H = Half2<kBF16>;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[ky][j] = make_float2(0.f, 0.f);
  if (xrow == nullptr) return;
  uint4 v[3];
#pragma unroll
  for (int kx = 0; kx < 3; ++kx) {
    const int wi = wi0 + kx * wstep;
    v[kx] = (wi >= 0 && wi < w) ? ldg_v4(xrow + (long long)wi * x_ld) : make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int kx = 0; kx < 3; ++kx) {
    const uint32_t u[4] = {v[kx].x, v[kx].y, v[kx].z, v[kx].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = H::unpack(u[j]);
      if (pre_relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) s[ky][j] = ffma2(f, wt[ky * 3 + kx][j], s[ky][j]);
    }
  }
}

template <bool kBF16>
__device__ __forceinline__ void store_out(typename Half2<kBF16>::T* dst, const float2 (&o)[4], int act) {
  using H = Half2<kBF16>;
  uint4 r;
  r.x = H::pack(apply_act(o[0].x, act), apply_act(o[0].y, act));
  r.y = H::pack(apply_act(o[1].x, act), apply_act(o[1].y, act));
  r.z = H::pack(apply_act(o[2].x, act), apply_act(o[2].y, act));
  r.w = H::pack(apply_act(o[3].x, act), apply_act(o[3].y, act));
  *reinterpret_cast<uint4*>(dst) = r;
}

template <bool kBF16, int kStride>
__global__ void __launch_bounds__(128, 4)
dwconv3x3_kernel(const DwParams p) {
  using H = Half2<kBF16>;
  using T = typename H::T;
  const int lc = threadIdx.x % p.lc, lw = threadIdx.x / p.lc;
  const int cblk = blockIdx.x % p.cblocks, wblk = blockIdx.x / p.cblocks;
  const int cv = cblk * p.lc + lc;
  const int wo = wblk * (128 / p.lc) + lw;
  if (cv >= p.cv || wo >= p.wo) return;
  const int c0 = cv * 8;
  const int n = blockIdx.z;
  const int h_begin = blockIdx.y * p.rows_per_block;
  int h_end = h_begin + p.rows_per_block; if (h_end > p.ho) h_end = p.ho;

  float2 wt[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.wgt + t * p.c + c0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.wgt + t * p.c + c0 + 4));
    wt[t][0] = make_float2(a.x, a.y); wt[t][1] = make_float2(a.z, a.w);
    wt[t][2] = make_float2(b.x, b.y); wt[t][3] = make_float2(b.z, b.w);
  }
  float2 sh[4];
  if (p.shift != nullptr) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.shift + c0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + 4));
    sh[0] = make_float2(a.x, a.y); sh[1] = make_float2(a.z, a.w); sh[2] = make_float2(b.x, b.y); sh[3] = make_float2(b.z, b.w);
  } else {
    sh[0] = sh[1] = sh[2] = sh[3] = make_float2(0.f, 0.f);
  }
  const T* xn = reinterpret_cast<const T*>(p.x) + (long long)n * p.h * p.w * p.x_ld + c0;
  T* yn = reinterpret_cast<T*>(p.y) + ((long long)n * p.ho * p.wo + wo) * p.y_ld + c0;
  const int d = p.dil;
  const int wi0 = wo * kStride - d;
  const bool relu = p.pre_relu != 0;
  const long long xrow_stride = (long long)p.w * p.x_ld;
  const long long yrow_stride = (long long)p.wo * p.y_ld;

  if (kStride == 1) {
    // d interleaved chains: chain `ch` produces output rows h0, h0+d, ... from input rows r_k = h0 + (k-1)*d.
    // Before step k:  acc1 = shift + s0(r-2d) + s1(r-d)  (pending out(r-d)),  acc0 = shift + s0(r-d)  (pending out(r)).
    const int nchains = d < (h_end - h_begin) ? d : (h_end - h_begin);
    for (int ch = 0; ch < nchains; ++ch) {
      const int h0 = h_begin + ch;
      const int n_out = (h_end - h0 + d - 1) / d;
      float2 acc0[4], acc1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc0[j] = sh[j]; acc1[j] = sh[j]; }
      for (int k = 0; k < n_out + 2; ++k) {
        const int r = h0 + (k - 1) * d;
        float2 s[3][4];
        row_sums<kBF16>((r >= 0 && r < p.h) ? xn + r * xrow_stride : nullptr, wi0, d, p.w, p.x_ld, relu, wt, s);
        if (k >= 2) {
          float2 o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = fadd2(acc1[j], s[2][j]);
          store_out<kBF16>(yn + (r - d) * yrow_stride, o, p.act);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc1[j] = fadd2(acc0[j], s[1][j]); acc0[j] = fadd2(sh[j], s[0][j]); }
      }
    }
  } else if (d == 1) {
    // stride 2, dilation 1: out(ho) = s0(2ho-1) + s1(2ho) + s2(2ho+1); row 2ho+1 is shared with out(ho+1)
    float2 acc0[4];
    {
      const int r = 2 * h_begin - 1;
      float2 s[3][4];
      row_sums<kBF16>((r >= 0 && r < p.h) ? xn + r * xrow_stride : nullptr, wi0, 1, p.w, p.x_ld, relu, wt, s);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc0[j] = fadd2(sh[j], s[0][j]);
    }
    for (int ho = h_begin; ho < h_end; ++ho) {
      float2 sa[3][4], sb[3][4];
      const int r0 = 2 * ho, r1 = 2 * ho + 1;
      row_sums<kBF16>((r0 < p.h) ? xn + r0 * xrow_stride : nullptr, wi0, 1, p.w, p.x_ld, relu, wt, sa);
      row_sums<kBF16>((r1 < p.h) ? xn + r1 * xrow_stride : nullptr, wi0, 1, p.w, p.x_ld, relu, wt, sb);
      float2 o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[j] = fadd2(fadd2(acc0[j], sa[1][j]), sb[2][j]);
        acc0[j] = fadd2(sh[j], sb[0][j]);
      }
      store_out<kBF16>(yn + ho * yrow_stride, o, p.act);
    }
  } else {
    // stride 2 with dilation > 1 (not used by the zoo's hot path): no row sharing
    for (int ho = h_begin; ho < h_end; ++ho) {
      float2 o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = sh[j];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int r = ho * 2 + (ky - 1) * d;
        if (r < 0 || r >= p.h) continue;
        float2 s[3][4];
        row_sums<kBF16>(xn + r * xrow_stride, wi0, d, p.w, p.x_ld, relu, wt, s);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = fadd2(o[j], s[ky][j]);
      }
      store_out<kBF16>(yn + ho * yrow_stride, o, p.act);
    }
  }
}


// ---------------------------------------------------------------------------------------------
// TMA-ring variant (C >= 64): one producer warp streams input rows (box = 64 channels x (tile width + halo)
// columns x 1 row; conv zero padding = TMA out-of-bounds zero fill) into a shared-memory ring guarded by
// full/empty mbarriers; 8 consumer warps read their three horizontal taps with conflict-free LDS.128 and
// run the same rolling partial-sum recurrence as above.  Many rows are in flight per SM without holding
// registers, which is what an HBM-bound stencil needs (Little's law: ~26 KB/SM at 6.5 TB/s).
// ---------------------------------------------------------------------------------------------
struct DwRingParams {
  DwParams b;
  int twin;          // input columns per ring slot
  int slot_bytes;    // twin * 128
  int nslots;
  int wblocks, segs; // persistent (4-channel) kernel: column blocks and row segments per image
};

static int g_dw_ring_slots = 0;
static int g_dw_v8 = 0;           // 1: the round-1 8-channel ring kernel (A/B knob "dw_v8")
static int g_dw_persistent = 0;   // 1: grid = 3 CTAs / SM walking the tiles through one continuous ring (A/B knob "dw_persistent")
int set_dw_ring_slots(int n) { g_dw_ring_slots = n; return 0; }
int set_dw_v8(int v) { g_dw_v8 = v ? 1 : 0; return 0; }
int set_dw_persistent(int v) { g_dw_persistent = v ? 1 : 0; return 0; }
// 1: two-column kernel with 5 consumer warps (20 columns, 192 threads, 3 CTAs / SM) for dilation 1 (A/B knob "dw_cw5").  Measured on
// B200 (profiles/r2_dw_sweep_cw5.jsonl): within +-2 % of the 7-warp / 2-CTA form on five of six shapes (c728 @129x257 +4.5 %, c128
// @513x1025 -3 %), step 16.34 vs 16.37 ms: no gain, off by default.
static int g_dw_cw5 = 0;
int set_dw_cw5(int v) { g_dw_cw5 = v ? 1 : 0; return 0; }
static int g_dw_cols2 = 1;        // 1: two output columns per thread for stride 1 / dilation 1 (A/B knob "dw_cols2")
int set_dw_cols2(int v) { g_dw_cols2 = v ? 1 : 0; return 0; }

constexpr int kDwConsumers = 224;   // 7 warps: thread -> (c8 = t & 7, column = t >> 3); +1 producer warp = 256 threads
constexpr int kDwTW = kDwConsumers / 8;   // 28 output columns per CTA (128 regs x 256 threads -> 2 CTAs / SM)

template <bool kBF16>
__device__ __forceinline__ void ring_row_sums(uint32_t row_addr, uint32_t tap_step, bool pre_relu, const float2 (&wt)[9][4],
                                              float2 (&s)[3][4], uint32_t empty_bar) {
  // On entry s[ky] holds the SEED of kernel row ky's FMA chain (the rolling accumulator it is added to), on exit
  // seed + sum_kx w[ky][kx] * x[kx]: the accumulator updates cost no extra add instructions.
  using H = Half2<kBF16>;
  uint4 v[3];
#pragma unroll
  for (int kx = 0; kx < 3; ++kx) v[kx] = lds_v4(row_addr + kx * tap_step);
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive_addr(empty_bar);     // slot may be refilled once all consumer warps have read it
#pragma unroll
  for (int kx = 0; kx < 3; ++kx) {
    uint32_t u[4] = {v[kx].x, v[kx].y, v[kx].z, v[kx].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (pre_relu) u[j] = H::relu2(u[j]);
      const float2 f = H::unpack(u[j]);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) s[ky][j] = ffma2(f, wt[ky * 3 + kx][j], s[ky][j]);
    }
  }
}

template <bool kBF16, int kStride, bool kPreRelu>
__global__ void __launch_bounds__(kDwConsumers + 32, 2)
dwconv3x3_ring_kernel(const __grid_constant__ CUtensorMap tmX, const DwRingParams rp) {
  using H = Half2<kBF16>;
  using T = typename H::T;
  extern __shared__ __align__(128) uint8_t dsm[];
  const DwParams& p = rp.b;
  uint64_t* full = reinterpret_cast<uint64_t*>(dsm + rp.nslots * rp.slot_bytes);
  uint64_t* empty = full + rp.nslots;
  const int warp = threadIdx.x >> 5;
  const int cblk = blockIdx.x % p.cblocks, wblk = blockIdx.x / p.cblocks;
  const int n = blockIdx.z;
  const int h_begin = blockIdx.y * p.rows_per_block;
  int h_end = h_begin + p.rows_per_block; if (h_end > p.ho) h_end = p.ho;
  const int d = p.dil;
  const int w0 = wblk * kDwTW;
  if (threadIdx.x == 0) {
    for (int i = 0; i < rp.nslots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kDwConsumers / 32); }
    fence_mbar_init();
  }
  __syncthreads();

  // the row sequence is identical for producer and consumers:
  //   stride 1: for each chain ch: rows h_begin + ch + (k-1)*d, k = 0 .. n_out+1
  //   stride 2 (d == 1): rows 2*h_begin-1 .. 2*(h_end-1)+1
  if (warp == kDwConsumers / 32) {
    if ((threadIdx.x & 31) == 0) {
      int slot = 0; uint32_t phase = 0;
      auto issue = [&](int r) {
        mbar_wait(&empty[slot], phase ^ 1);
        mbar_expect_tx(&full[slot], (uint32_t)rp.slot_bytes);
        tma_load_4d(&tmX, &full[slot], dsm + slot * rp.slot_bytes, cblk * 64, w0 * kStride - d, r, n);
        if (++slot == rp.nslots) { slot = 0; phase ^= 1; }
      };
      if (kStride == 1) {
        const int nchains = d < (h_end - h_begin) ? d : (h_end - h_begin);
        for (int ch = 0; ch < nchains; ++ch) {
          const int h0 = h_begin + ch;
          const int n_out = (h_end - h0 + d - 1) / d;
          for (int k = 0; k < n_out + 2; ++k) issue(h0 + (k - 1) * d);
        }
      } else {
        for (int r = 2 * h_begin - 1; r <= 2 * (h_end - 1) + 1; ++r) issue(r);
      }
    }
    return;
  }

  const int c8 = threadIdx.x & 7, wl = threadIdx.x >> 3;
  const int cv = cblk * 8 + c8;
  const int wo = w0 + wl;
  const bool active = cv < p.cv && wo < p.wo;
  const int c0 = (cv < p.cv ? cv : 0) * 8;
  float2 wt[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.wgt + t * p.c + c0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.wgt + t * p.c + c0 + 4));
    wt[t][0] = make_float2(a.x, a.y); wt[t][1] = make_float2(a.z, a.w);
    wt[t][2] = make_float2(b.x, b.y); wt[t][3] = make_float2(b.z, b.w);
  }
  float2 sh[4];
  if (p.shift != nullptr) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.shift + c0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + 4));
    sh[0] = make_float2(a.x, a.y); sh[1] = make_float2(a.z, a.w); sh[2] = make_float2(b.x, b.y); sh[3] = make_float2(b.z, b.w);
  } else {
    sh[0] = sh[1] = sh[2] = sh[3] = make_float2(0.f, 0.f);
  }
  T* yn = reinterpret_cast<T*>(p.y) + ((long long)n * p.ho * p.wo + (wo < p.wo ? wo : 0)) * p.y_ld + c0;
  const long long yrow_stride = (long long)p.wo * p.y_ld;
  constexpr bool relu = kPreRelu;
  // slot column of the left tap is wl*s (a slot starts at input column w0*s - d); taps are d columns (d*128 B) apart
  const uint32_t ring0 = smem_u32(dsm) + (uint32_t)(wl * kStride * 128 + c8 * 16);
  const uint32_t tap_step = (uint32_t)(d * 128);
  const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
  int slot = 0; uint32_t phase = 0;
#define DW_ROW(S) do { \
    mbar_wait_lean(full0 + slot * 8, phase); \
    ring_row_sums<kBF16>(ring0 + slot * rp.slot_bytes, tap_step, relu, wt, S, empty0 + slot * 8); \
    if (++slot == rp.nslots) { slot = 0; phase ^= 1; } } while (0)

  if (kStride == 1) {
    const int nchains = d < (h_end - h_begin) ? d : (h_end - h_begin);
    for (int ch = 0; ch < nchains; ++ch) {
      const int h0 = h_begin + ch;
      const int n_out = (h_end - h0 + d - 1) / d;
      float2 acc0[4], acc1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc0[j] = sh[j]; acc1[j] = sh[j]; }
      for (int k = 0; k < n_out + 2; ++k) {
        const int r = h0 + (k - 1) * d;
        float2 s[3][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { s[2][j] = acc1[j]; s[1][j] = acc0[j]; s[0][j] = sh[j]; }   // chain seeds
        DW_ROW(s);
        if (k >= 2 && active) store_out<kBF16>(yn + (r - d) * yrow_stride, s[2], p.act);         // acc1 + s2(r)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc1[j] = s[1][j]; acc0[j] = s[0][j]; }
      }
    }
  } else {
    float2 acc0[4];
    {
      float2 s[3][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[0][j] = sh[j]; s[1][j] = sh[j]; s[2][j] = sh[j]; }
      DW_ROW(s);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc0[j] = s[0][j];
    }
    for (int ho = h_begin; ho < h_end; ++ho) {
      float2 sa[3][4], sb[3][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { sa[1][j] = acc0[j]; sa[0][j] = sh[j]; sa[2][j] = sh[j]; }   // only sa[1] is used
      DW_ROW(sa);
#pragma unroll
      for (int j = 0; j < 4; ++j) { sb[2][j] = sa[1][j]; sb[0][j] = sh[j]; sb[1][j] = sh[j]; }   // sb[2] = out, sb[0] = next acc0
      DW_ROW(sb);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc0[j] = sb[0][j];
      if (active) store_out<kBF16>(yn + ho * yrow_stride, sb[2], p.act);
    }
  }
}

#undef DW_ROW

// ---------------------------------------------------------------------------------------------
// Lean ring variant (round 2; default for C >= 64): FOUR channels per thread (16 lanes cover the 128-byte channel block of a
// column, a warp two adjacent columns), 14 output columns per CTA.  Against the 8-channel kernel above:
//   * 36 weight registers instead of 72 -> <= 80 registers -> 3 CTAs (21 consumer warps) per SM instead of 2 (14): the 8-channel
//     kernel sat at 54-57 % issue-slot use with 21 % occupancy -- latency-bound (profiles/r1_ncu_full_summary.md);
//   * the per-row instruction stream is cut to the arithmetic: activation and leading ReLU are template parameters, the output
//     pointer and the ring addresses advance incrementally (the 8-channel loop spent ~60 of its ~190 instructions per row on 64-bit
//     address arithmetic and run-time activation dispatch, cuobjdump).
// Same producer protocol, same rolling partial-sum recurrence, bit-identical results (same fp32 operation order per channel).
// ---------------------------------------------------------------------------------------------
constexpr int kDw4Cols = kDwConsumers / 16;   // 14 output columns per CTA

template <bool kBF16, bool kPreRelu>
__device__ __forceinline__ void ring_row_sums4(uint32_t row_addr, uint32_t tap_step, const float2 (&wt)[9][2], float2 (&s)[3][2],
                                               uint32_t empty_bar) {
  using H = Half2<kBF16>;
  uint2 v[3];
#pragma unroll
  for (int kx = 0; kx < 3; ++kx)
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v[kx].x), "=r"(v[kx].y) : "r"(row_addr + kx * tap_step));
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive_addr(empty_bar);
#pragma unroll
  for (int kx = 0; kx < 3; ++kx) {
    uint32_t u[2] = {v[kx].x, v[kx].y};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (kPreRelu) u[j] = H::relu2(u[j]);
      const float2 f = H::unpack(u[j]);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) s[ky][j] = ffma2(f, wt[ky * 3 + kx][j], s[ky][j]);
    }
  }
}

template <bool kBF16, int kAct>
__device__ __forceinline__ void store_out4(void* dst, const float2 (&o)[2]) {
  using H = Half2<kBF16>;
  uint2 r;
  r.x = H::pack(o[0].x, o[0].y);
  r.y = H::pack(o[1].x, o[1].y);
  if (kAct != ACT_NONE) { r.x = H::relu2(r.x); r.y = H::relu2(r.y); }      // round(max(v,0)) == max(round(v),0)
  if (kAct == ACT_RELU6) { r.x = H::min2(r.x, 6.f); r.y = H::min2(r.y, 6.f); }
  *reinterpret_cast<uint2*>(dst) = r;
}

template <bool kBF16, int kStride, bool kPreRelu, int kAct, bool kD1>
__global__ void __launch_bounds__(kDwConsumers + 32, 3)
dwconv3x3_ring4_kernel(const __grid_constant__ CUtensorMap tmX, const DwRingParams rp) {
  // Work items (tiles):  item -> (image | row segment | column block | channel block), channel block fastest.  Default grid: one
  // item per CTA.  Optional persistent mode (segb200_set_option("dw_persistent", 1)): the grid is 3 CTAs per SM, a CTA walks
  // item = blockIdx.x + k * gridDim.x and streams the rows of successive items through ONE continuous ring (the producer runs
  // ahead across item boundaries).  Measured on B200 (profiles/r2_dw_sweep.jsonl): not faster -- the hardware scheduler balances
  // one-tile CTAs better than a static round-robin, and 3 resident CTAs already overlap one tile's fill with another's drain.
  using H = Half2<kBF16>;
  using T = typename H::T;
  extern __shared__ __align__(128) uint8_t dsm[];
  const DwParams& p = rp.b;
  uint64_t* full = reinterpret_cast<uint64_t*>(dsm + rp.nslots * rp.slot_bytes);
  uint64_t* empty = full + rp.nslots;
  const int warp = threadIdx.x >> 5;
  const int d = kD1 ? 1 : p.dil;                        // kD1: dilation 1 is a compile-time constant (tap offsets become immediates)
  const int wblocks = rp.wblocks, segs = rp.segs;
  const int per_cblk = wblocks * segs * p.n;
  const int total = per_cblk * p.cblocks;
  if (threadIdx.x == 0) {
    for (int i = 0; i < rp.nslots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kDwConsumers / 32); }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();
  auto decode = [&](int item, int& cblk, int& n, int& h_begin, int& h_end, int& w0) {
    // channel block fastest: CTAs that run together read neighbouring 128-byte chunks of the SAME pixels (DRAM page / L2 sector
    // locality: a pixel's 2*C bytes are touched once, not cblocks times at different moments)
    cblk = item % p.cblocks;
    int r = item / p.cblocks;
    n = r / (wblocks * segs); r -= n * (wblocks * segs);
    const int seg = r / wblocks;
    w0 = (r - seg * wblocks) * kDw4Cols;
    h_begin = seg * p.rows_per_block;
    h_end = h_begin + p.rows_per_block; if (h_end > p.ho) h_end = p.ho;
  };

  if (warp == kDwConsumers / 32) {                     // ---- producer warp: identical item / row sequence to the consumers ----
    if ((threadIdx.x & 31) == 0) {
      int slot = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        int cblk, n, h_begin, h_end, w0;
        decode(item, cblk, n, h_begin, h_end, w0);
        auto issue = [&](int r) {
          mbar_wait(&empty[slot], phase ^ 1);
          mbar_expect_tx(&full[slot], (uint32_t)rp.slot_bytes);
          tma_load_4d(&tmX, &full[slot], dsm + slot * rp.slot_bytes, cblk * 64, w0 * kStride - d, r, n);
          if (++slot == rp.nslots) { slot = 0; phase ^= 1; }
        };
        if (kStride == 1) {
          const int nchains = d < (h_end - h_begin) ? d : (h_end - h_begin);
          for (int ch = 0; ch < nchains; ++ch) {
            const int h0 = h_begin + ch;
            const int n_out = (h_end - h0 + d - 1) / d;
            for (int k = 0; k < n_out + 2; ++k) issue(h0 + (k - 1) * d);
          }
        } else {
          for (int r = 2 * h_begin - 1; r <= 2 * (h_end - 1) + 1; ++r) issue(r);
        }
      }
    }
    return;
  }

  const int c4 = threadIdx.x & 15, wl = threadIdx.x >> 4;
  const long long yrow_bytes = (long long)p.wo * p.y_ld * (long long)sizeof(T);
  const uint32_t ring0 = smem_u32(dsm) + (uint32_t)(wl * kStride * 128 + c4 * 8);
  const uint32_t tap_step = (uint32_t)(d * 128);
  const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
  const uint32_t slot_bytes = (uint32_t)rp.slot_bytes;
  const uint32_t bar_end = full0 + (uint32_t)rp.nslots * 8;
  uint32_t ring = ring0, fbar = full0, ebar = empty0, phase = 0;
#define DW4_ROW(S) do { \
    mbar_wait_lean(fbar, phase); \
    ring_row_sums4<kBF16, kPreRelu>(ring, tap_step, wt, S, ebar); \
    ring += slot_bytes; fbar += 8; ebar += 8; \
    if (fbar == bar_end) { ring = ring0; fbar = full0; ebar = empty0; phase ^= 1; } } while (0)

  float2 wt[9][2];
  float2 sh[2];
  int cur_cblk = -1;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    int cblk, n, h_begin, h_end, w0;
    decode(item, cblk, n, h_begin, h_end, w0);
    const int cq = cblk * 16 + c4;                     // channel quad
    const int wo = w0 + wl;
    const bool active = cq * 4 < p.c && wo < p.wo;
    const int c0 = (cq * 4 < p.c ? cq : 0) * 4;
    if (cblk != cur_cblk) {                            // (items are channel-block major: this happens once or twice per CTA)
      cur_cblk = cblk;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p.wgt + t * p.c + c0));
        wt[t][0] = make_float2(a.x, a.y); wt[t][1] = make_float2(a.z, a.w);
      }
      if (p.shift != nullptr) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p.shift + c0));
        sh[0] = make_float2(a.x, a.y); sh[1] = make_float2(a.z, a.w);
      } else {
        sh[0] = sh[1] = make_float2(0.f, 0.f);
      }
    }
    uint8_t* ybase = reinterpret_cast<uint8_t*>(reinterpret_cast<T*>(p.y) + ((long long)n * p.ho * p.wo + (wo < p.wo ? wo : 0)) * p.y_ld + c0);

    if (kStride == 1) {
      const int nchains = d < (h_end - h_begin) ? d : (h_end - h_begin);
      const long long ystep = yrow_bytes * d;
      for (int ch = 0; ch < nchains; ++ch) {
        const int h0 = h_begin + ch;
        const int n_out = (h_end - h0 + d - 1) / d;
        uint8_t* yp = ybase + (long long)h0 * yrow_bytes;
        float2 acc0[2] = {sh[0], sh[1]}, acc1[2] = {sh[0], sh[1]};
        // rows k = 0, 1 only feed the accumulators; rows k >= 2 also complete output row h0 + (k-2)*d
        for (int k = 0; k < n_out + 2; ++k) {
          float2 s[3][2] = {{sh[0], sh[1]}, {acc0[0], acc0[1]}, {acc1[0], acc1[1]}};      // chain seeds
          DW4_ROW(s);
          if (k >= 2) {
            if (active) store_out4<kBF16, kAct>(yp, s[2]);
            yp += ystep;
          }
          acc1[0] = s[1][0]; acc1[1] = s[1][1]; acc0[0] = s[0][0]; acc0[1] = s[0][1];
        }
      }
    } else {
      float2 acc0[2];
      {
        float2 s[3][2] = {{sh[0], sh[1]}, {sh[0], sh[1]}, {sh[0], sh[1]}};
        DW4_ROW(s);
        acc0[0] = s[0][0]; acc0[1] = s[0][1];
      }
      uint8_t* yp = ybase + (long long)h_begin * yrow_bytes;
      for (int ho = h_begin; ho < h_end; ++ho) {
        float2 sa[3][2] = {{sh[0], sh[1]}, {acc0[0], acc0[1]}, {sh[0], sh[1]}};          // only sa[1] is used
        DW4_ROW(sa);
        float2 sb[3][2] = {{sh[0], sh[1]}, {sh[0], sh[1]}, {sa[1][0], sa[1][1]}};        // sb[2] = out, sb[0] = next acc0
        DW4_ROW(sb);
        acc0[0] = sb[0][0]; acc0[1] = sb[0][1];
        if (active) store_out4<kBF16, kAct>(yp, sb[2]);
        yp += yrow_bytes;
      }
    }
  }
#undef DW4_ROW
}

// ---------------------------------------------------------------------------------------------
// Two-column variant of the lean ring kernel for stride 1, dilation 1 (59 of the 68 depthwise launches of the headline step):
// a thread owns FOUR channels of TWO adjacent output columns.  The two outputs share two of their three taps, so a ring row costs
// 4 shared-memory loads + 8 unpacks per 8 outputs instead of 3 + 6 per 4, and the per-row ring hand-shake (wait, arrive, slot
// advance) is paid once per 8 outputs: ~10.5 issued instructions per output element against ~15.5 (the one-column kernel is bound by
// instruction issue, profiles/r2_ncu_full_summary.md).  28 output columns per CTA, slot = 30 input columns x 128 B.  The first two
// rows of a segment (accumulate only) are peeled, so the steady-state loop has no row-index test.  Same producer protocol, same
// fp32 operation order per output -> bit-identical to the one-column kernel.
// ---------------------------------------------------------------------------------------------
constexpr int kDw2Cols = 2 * kDw4Cols;   // 28 output columns per CTA

// One ring row -> the three kernel-row partial sums of both columns.  kD1: the columns share two of their three taps (4 loads);
// dilated: taps are d columns apart and nothing is shared (6 loads), the gain is the per-row hand-shake paid once per 8 outputs.
template <bool kBF16, bool kPreRelu, bool kD1>
__device__ __forceinline__ void ring_row_sums4x2(uint32_t row_addr, uint32_t tap_step, const float2 (&wt)[9][2], float2 (&sa)[3][2],
                                                 float2 (&sb)[3][2], uint32_t empty_bar) {
  using H = Half2<kBF16>;
  constexpr int kLoads = kD1 ? 4 : 6;
  uint2 v[kLoads];
#pragma unroll
  for (int k = 0; k < kLoads; ++k) {
    // kD1: slot columns 2wl + k.  dilated: k = 2 * kx + e -> slot column 2wl + e + kx * d
    const uint32_t off = kD1 ? (uint32_t)(k * 128) : (uint32_t)((k & 1) * 128) + (uint32_t)(k >> 1) * tap_step;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v[k].x), "=r"(v[k].y) : "r"(row_addr + off));
  }
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive_addr(empty_bar);
  float2 f[kLoads][2];
#pragma unroll
  for (int k = 0; k < kLoads; ++k) {
    uint32_t u[2] = {v[k].x, v[k].y};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (kPreRelu) u[j] = H::relu2(u[j]);
      f[k][j] = H::unpack(u[j]);
    }
  }
#pragma unroll
  for (int kx = 0; kx < 3; ++kx)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        sa[ky][j] = ffma2(f[kD1 ? kx : 2 * kx][j], wt[ky * 3 + kx][j], sa[ky][j]);
        sb[ky][j] = ffma2(f[kD1 ? kx + 1 : 2 * kx + 1][j], wt[ky * 3 + kx][j], sb[ky][j]);
      }
}

// kCW = consumer warps: 7 (28 output columns, 256 threads, 2 CTAs / SM) or 5 (20 columns, 192 threads, 3 CTAs / SM when the kernel
// stays within 113 registers: 18 warps per SM instead of 16, and while one CTA fills its ring two others compute).
template <bool kBF16, bool kPreRelu, int kAct, bool kD1, int kCW>
__global__ void __launch_bounds__(kCW * 32 + 32, kCW == 7 ? 2 : 3)
dwconv3x3_ring4x2_kernel(const __grid_constant__ CUtensorMap tmX, const DwRingParams rp) {
  using H = Half2<kBF16>;
  using T = typename H::T;
  extern __shared__ __align__(128) uint8_t dsm[];
  const DwParams& p = rp.b;
  uint64_t* full = reinterpret_cast<uint64_t*>(dsm + rp.nslots * rp.slot_bytes);
  uint64_t* empty = full + rp.nslots;
  const int warp = threadIdx.x >> 5;
  const int d = kD1 ? 1 : p.dil;
  if (threadIdx.x == 0) {
    for (int i = 0; i < rp.nslots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kCW); }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_launch_dependents();                             // programmatic dependent launch (common.cuh): the barrier set-up above overlapped
  pdl_wait();                                          // the previous kernel's tail; its results are visible from here on
  // one tile per CTA; channel block fastest (see dwconv3x3_ring4_kernel)
  const int item = blockIdx.x;
  const int cblk = item % p.cblocks;
  int r = item / p.cblocks;
  const int n = r / (rp.wblocks * rp.segs); r -= n * (rp.wblocks * rp.segs);
  const int seg = r / rp.wblocks;
  const int w0 = (r - seg * rp.wblocks) * (4 * kCW);
  const int h_begin = seg * p.rows_per_block;
  const int h_end = h_begin + p.rows_per_block < p.ho ? h_begin + p.rows_per_block : p.ho;
  // d interleaved row chains (output rows h0, h0 + d, ...): a chain streams input rows h0 - d, h0, ..., one ring slot each
  const int nchains = d < (h_end - h_begin) ? d : (h_end - h_begin);

  if (warp == kCW) {                     // ---- producer warp ----
    if ((threadIdx.x & 31) == 0) {
      int slot = 0; uint32_t phase = 0;
      for (int ch = 0; ch < nchains; ++ch) {
        const int h0 = h_begin + ch;
        const int n_out = (h_end - h0 + d - 1) / d;
        for (int k = 0; k < n_out + 2; ++k) {
          mbar_wait(&empty[slot], phase ^ 1);
          mbar_expect_tx(&full[slot], (uint32_t)rp.slot_bytes);
          tma_load_4d(&tmX, &full[slot], dsm + slot * rp.slot_bytes, cblk * 64, w0 - d, h0 + (k - 1) * d, n);
          if (++slot == rp.nslots) { slot = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  const int c4 = threadIdx.x & 15, wl = threadIdx.x >> 4;
  const long long yrow_bytes = (long long)p.wo * p.y_ld * (long long)sizeof(T);
  const uint32_t ring0 = smem_u32(dsm) + (uint32_t)(wl * 256 + c4 * 8);
  const uint32_t tap_step = (uint32_t)(d * 128);
  const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
  const uint32_t slot_bytes = (uint32_t)rp.slot_bytes;
  const uint32_t bar_end = full0 + (uint32_t)rp.nslots * 8;
  uint32_t ring = ring0, fbar = full0, ebar = empty0, phase = 0;
#define DW2_ROW(SA, SB) do { \
    mbar_wait_lean(fbar, phase); \
    ring_row_sums4x2<kBF16, kPreRelu, kD1>(ring, tap_step, wt, SA, SB, ebar); \
    ring += slot_bytes; fbar += 8; ebar += 8; \
    if (fbar == bar_end) { ring = ring0; fbar = full0; ebar = empty0; phase ^= 1; } } while (0)

  const int cq = cblk * 16 + c4;                       // channel quad
  const int wo = w0 + 2 * wl;
  const bool act_a = cq * 4 < p.c && wo < p.wo, act_b = cq * 4 < p.c && wo + 1 < p.wo;
  const int c0 = (cq * 4 < p.c ? cq : 0) * 4;
  float2 wt[9][2];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.wgt + t * p.c + c0));
    wt[t][0] = make_float2(a.x, a.y); wt[t][1] = make_float2(a.z, a.w);
  }
  float2 sh[2];
  if (p.shift != nullptr) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.shift + c0));
    sh[0] = make_float2(a.x, a.y); sh[1] = make_float2(a.z, a.w);
  } else {
    sh[0] = sh[1] = make_float2(0.f, 0.f);
  }
  const long long ypix = (long long)p.y_ld * (long long)sizeof(T);
  const long long ystep = yrow_bytes * d;
  uint8_t* ybase = reinterpret_cast<uint8_t*>(reinterpret_cast<T*>(p.y) + ((long long)n * p.ho * p.wo + (wo < p.wo ? wo : 0)) * p.y_ld + c0);

  for (int ch = 0; ch < nchains; ++ch) {
    const int h0 = h_begin + ch;
    const int n_out = (h_end - h0 + d - 1) / d;
    uint8_t* yp = ybase + (long long)h0 * yrow_bytes;
    // rolling partial sums per column: a0 = kernel row 0 of the newest input row, a1 = rows 0..1 of the two newest
    float2 a0a[2], a1a[2], a0b[2], a1b[2];
    {                                                  // input row h0 - d: only its ky = 0 sum survives
      float2 sa[3][2] = {{sh[0], sh[1]}, {sh[0], sh[1]}, {sh[0], sh[1]}}, sb[3][2] = {{sh[0], sh[1]}, {sh[0], sh[1]}, {sh[0], sh[1]}};
      DW2_ROW(sa, sb);
      a0a[0] = sa[0][0]; a0a[1] = sa[0][1]; a0b[0] = sb[0][0]; a0b[1] = sb[0][1];
    }
    {                                                  // input row h0
      float2 sa[3][2] = {{sh[0], sh[1]}, {a0a[0], a0a[1]}, {sh[0], sh[1]}}, sb[3][2] = {{sh[0], sh[1]}, {a0b[0], a0b[1]}, {sh[0], sh[1]}};
      DW2_ROW(sa, sb);
      a1a[0] = sa[1][0]; a1a[1] = sa[1][1]; a0a[0] = sa[0][0]; a0a[1] = sa[0][1];
      a1b[0] = sb[1][0]; a1b[1] = sb[1][1]; a0b[0] = sb[0][0]; a0b[1] = sb[0][1];
    }
    for (int k = 0; k < n_out; ++k) {                  // input row h0 + (k + 1) d completes output row h0 + k d
      float2 sa[3][2] = {{sh[0], sh[1]}, {a0a[0], a0a[1]}, {a1a[0], a1a[1]}}, sb[3][2] = {{sh[0], sh[1]}, {a0b[0], a0b[1]}, {a1b[0], a1b[1]}};
      DW2_ROW(sa, sb);
      if (act_a) store_out4<kBF16, kAct>(yp, sa[2]);
      if (act_b) store_out4<kBF16, kAct>(yp + ypix, sb[2]);
      yp += ystep;
      a1a[0] = sa[1][0]; a1a[1] = sa[1][1]; a0a[0] = sa[0][0]; a0a[1] = sa[0][1];
      a1b[0] = sb[1][0]; a1b[1] = sb[1][1]; a0b[0] = sb[0][0]; a0b[1] = sb[0][1];
    }
  }
#undef DW2_ROW
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
  if (a->n < 1 || a->ho < 1 || a->wo < 1) return set_error(-6, "dwconv3x3: empty");
  if (a->n > 65535) return set_error(-6, "dwconv3x3: batch too large");
  DwParams p;
  p.x = a->x; p.wgt = a->wgt; p.shift = a->shift; p.y = a->y;
  p.n = a->n; p.h = a->h; p.w = a->w; p.c = a->c; p.x_ld = a->x_ld; p.y_ld = a->y_ld;
  p.ho = a->ho; p.wo = a->wo; p.stride = a->stride; p.dil = a->dilation; p.pre_relu = a->pre_relu; p.act = a->act;
  p.cv = a->c / 8;
  const bool ring = a->c >= 64 && (a->stride == 1 || a->dilation == 1) && a->dilation <= 64;
  int lw;
  const bool cols2 = ring && !g_dw_v8 && !g_dw_persistent && g_dw_cols2 && a->stride == 1;
  const int cw2 = (cols2 && g_dw_cw5 && a->dilation == 1) ? 5 : 7;          // consumer warps of the two-column kernel
  if (ring) { p.lc = 8; p.cblocks = (a->c + 63) / 64; lw = g_dw_v8 ? kDwTW : cols2 ? 4 * cw2 : kDw4Cols; }
  else { p.lc = p.cv <= 8 ? 8 : 16; p.cblocks = (p.cv + p.lc - 1) / p.lc; lw = 128 / p.lc; }
  const int wblocks = (a->wo + lw - 1) / lw;
  // rows per block: long enough to amortise the 2-row halo of each chain, short enough to fill the GPU
  int rows = a->stride == 1 ? (ring ? 32 : 16) * a->dilation : (ring ? 32 : 16);
  if (rows > a->ho) rows = a->ho;
  long long blocks_xy = (long long)p.cblocks * wblocks * a->n;
  while (rows > 4 * a->dilation && blocks_xy * ((a->ho + rows - 1) / rows) < 148LL * 8) rows = (rows + 1) / 2;
  if (ring && !g_dw_v8) {
    // persistent kernel: no need to over-decompose for occupancy; balanced segments (65 rows -> 3 x 22, not 32 + 32 + 1)
    rows = a->stride == 1 ? 32 * a->dilation : 32;
    if (rows > a->ho) rows = a->ho;
    while (rows > 8 * a->dilation && blocks_xy * ((a->ho + rows - 1) / rows) < 148LL * ((cols2 && cw2 == 7) ? 2 : 3) * (g_dw_persistent ? 2 : 4)) rows = (rows + 1) / 2;
    const int nseg = (a->ho + rows - 1) / rows;
    rows = (a->ho + nseg - 1) / nseg;
  }
  if (a->stride == 1 && rows < a->ho) rows = ((rows + a->dilation - 1) / a->dilation) * a->dilation;   // whole chains
  p.rows_per_block = rows;
  dim3 grid((unsigned)(p.cblocks * wblocks), (unsigned)((a->ho + rows - 1) / rows), (unsigned)a->n);
  if (grid.y > 65535) return set_error(-6, "dwconv3x3: too many row segments");
  if (ring) {
    DwRingParams rp;
    rp.b = p;
    const int tw = g_dw_v8 ? kDwTW : cols2 ? 4 * cw2 : kDw4Cols;
    rp.twin = (tw - 1) * a->stride + 2 * a->dilation + 1;
    rp.slot_bytes = rp.twin * 128;
    rp.nslots = (cols2 && a->dilation > 1 ? 65536 : 49152) / rp.slot_bytes;
    if (rp.nslots > (g_dw_v8 ? 12 : 24)) rp.nslots = g_dw_v8 ? 12 : 24;
    if (g_dw_ring_slots > 0 && rp.nslots > g_dw_ring_slots) rp.nslots = g_dw_ring_slots;
    if (rp.nslots < 3) rp.nslots = 3;
    const int smem = rp.nslots * rp.slot_bytes + 2 * rp.nslots * 8;
    CUtensorMap tmX;
    const uint64_t dims[4] = {(uint64_t)a->c, (uint64_t)a->w, (uint64_t)a->h, (uint64_t)a->n};
    const uint64_t str[3] = {(uint64_t)a->x_ld * 2, (uint64_t)a->x_ld * 2 * a->w, (uint64_t)a->x_ld * 2 * a->w * a->h};
    const uint32_t box[4] = {64u, (uint32_t)rp.twin, 1u, 1u};
    int rc = encode_map(&tmX, a->dtype, 4, a->x, dims, str, box, 0, "dw/X");
    if (rc) return rc;
    typedef void (*RingFn)(const CUtensorMap, const DwRingParams);
    const int threads = kDwConsumers + 32;
    if (cols2) {
      // [dilation == 1][dtype][pre_relu][act]
#define DW2_ACTS(BF, PR, D1) {dwconv3x3_ring4x2_kernel<BF, PR, 0, D1, 7>, dwconv3x3_ring4x2_kernel<BF, PR, 1, D1, 7>, dwconv3x3_ring4x2_kernel<BF, PR, 2, D1, 7>}
      static const RingFn fns2[2][2][2][3] = {{{DW2_ACTS(false, false, false), DW2_ACTS(false, true, false)}, {DW2_ACTS(true, false, false), DW2_ACTS(true, true, false)}},
                                              {{DW2_ACTS(false, false, true), DW2_ACTS(false, true, true)}, {DW2_ACTS(true, false, true), DW2_ACTS(true, true, true)}}};
#undef DW2_ACTS
      // 5 consumer warps, dilation 1 only: [dtype][pre_relu][act]
#define DW2_ACTS5(BF, PR) {dwconv3x3_ring4x2_kernel<BF, PR, 0, true, 5>, dwconv3x3_ring4x2_kernel<BF, PR, 1, true, 5>, dwconv3x3_ring4x2_kernel<BF, PR, 2, true, 5>}
      static const RingFn fns25[2][2][3] = {{DW2_ACTS5(false, false), DW2_ACTS5(false, true)}, {DW2_ACTS5(true, false), DW2_ACTS5(true, true)}};
#undef DW2_ACTS5
      static std::once_flag once2;
      std::call_once(once2, [] {
        for (int i = 0; i < 24; ++i) cudaFuncSetAttribute(fns2[i / 12][(i / 6) & 1][(i / 3) & 1][i % 3], cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        for (int i = 0; i < 12; ++i) cudaFuncSetAttribute(fns25[i / 6][(i / 3) & 1][i % 3], cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      });
      if (a->act < 0 || a->act > 2) return set_error(-3, "dwconv3x3: bad activation code");
      rp.wblocks = wblocks; rp.segs = (int)grid.y;
      const long long total = (long long)p.cblocks * wblocks * grid.y * a->n;
      if (total > 0x7fffffffLL) return set_error(-6, "dwconv3x3: too many tiles");
      const RingFn fn2 = cw2 == 5 ? fns25[a->dtype == DT_BF16 ? 1 : 0][a->pre_relu ? 1 : 0][a->act]
                                  : fns2[a->dilation == 1 ? 1 : 0][a->dtype == DT_BF16 ? 1 : 0][a->pre_relu ? 1 : 0][a->act];
      cudaError_t le = launch_kernel(fn2, dim3((unsigned)total), dim3((unsigned)(cw2 * 32 + 32)), (size_t)smem, stream, pdl_enabled() != 0, tmX, rp);
      if (le != cudaSuccess) return set_error((int)le, "dwconv3x3: launch failed: %s", cudaGetErrorString(le));
      return check_launch("dwconv3x3(ring4x2)");
    }
    if (!g_dw_v8) {
      // [dtype][stride-1][pre_relu][act]
      // [dilation == 1][dtype][stride-1][pre_relu][act]
#define DW4_ROWS(BF, S, D1) {{dwconv3x3_ring4_kernel<BF, S, false, 0, D1>, dwconv3x3_ring4_kernel<BF, S, false, 1, D1>, dwconv3x3_ring4_kernel<BF, S, false, 2, D1>}, \
                             {dwconv3x3_ring4_kernel<BF, S, true, 0, D1>, dwconv3x3_ring4_kernel<BF, S, true, 1, D1>, dwconv3x3_ring4_kernel<BF, S, true, 2, D1>}}
      static const RingFn fns4x[2][2][2][2][3] = {{{DW4_ROWS(false, 1, false), DW4_ROWS(false, 2, false)}, {DW4_ROWS(true, 1, false), DW4_ROWS(true, 2, false)}},
                                                  {{DW4_ROWS(false, 1, true), DW4_ROWS(false, 2, true)}, {DW4_ROWS(true, 1, true), DW4_ROWS(true, 2, true)}}};
#undef DW4_ROWS
      const auto& fns4 = fns4x[a->dilation == 1 ? 1 : 0];
      static std::once_flag once4;
      std::call_once(once4, [] {
        for (int i = 0; i < 48; ++i)
          cudaFuncSetAttribute(fns4x[i / 24][(i / 12) & 1][(i / 6) & 1][(i / 3) & 1][i % 3], cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      });
      if (a->act < 0 || a->act > 2) return set_error(-3, "dwconv3x3: bad activation code");
      rp.wblocks = wblocks; rp.segs = (int)grid.y;
      const long long total = (long long)p.cblocks * wblocks * grid.y * a->n;
      if (total > 0x7fffffffLL) return set_error(-6, "dwconv3x3: too many tiles");
      long long pgrid = g_dw_persistent ? (long long)num_sms() * 3 : total;     // non-persistent: one tile per CTA, same kernel
      if (pgrid > total) pgrid = total;
      cudaError_t le = launch_kernel(fns4[a->dtype == DT_BF16 ? 1 : 0][a->stride - 1][a->pre_relu ? 1 : 0][a->act], dim3((unsigned)pgrid),
                                     dim3((unsigned)threads), (size_t)smem, stream, pdl_enabled() != 0, tmX, rp);
      if (le != cudaSuccess) return set_error((int)le, "dwconv3x3: launch failed: %s", cudaGetErrorString(le));
      return check_launch("dwconv3x3(ring4)");
    }
    static const RingFn fns[2][2][2] = {
        {{dwconv3x3_ring_kernel<false, 1, false>, dwconv3x3_ring_kernel<false, 1, true>},
         {dwconv3x3_ring_kernel<false, 2, false>, dwconv3x3_ring_kernel<false, 2, true>}},
        {{dwconv3x3_ring_kernel<true, 1, false>, dwconv3x3_ring_kernel<true, 1, true>},
         {dwconv3x3_ring_kernel<true, 2, false>, dwconv3x3_ring_kernel<true, 2, true>}}};
    static std::once_flag once;
    std::call_once(once, [] {
      for (int i = 0; i < 8; ++i)
        cudaFuncSetAttribute(fns[i >> 2][(i >> 1) & 1][i & 1], cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    });
    fns[a->dtype == DT_BF16 ? 1 : 0][a->stride - 1][a->pre_relu ? 1 : 0]<<<grid, threads, smem, stream>>>(tmX, rp);
    return check_launch("dwconv3x3(ring)");
  }
  if (a->dtype == DT_BF16) {
    if (a->stride == 1) dwconv3x3_kernel<true, 1><<<grid, 128, 0, stream>>>(p);
    else dwconv3x3_kernel<true, 2><<<grid, 128, 0, stream>>>(p);
  } else {
    if (a->stride == 1) dwconv3x3_kernel<false, 1><<<grid, 128, 0, stream>>>(p);
    else dwconv3x3_kernel<false, 2><<<grid, 128, 0, stream>>>(p);
  }
  return check_launch("dwconv3x3");
}

// segb200 -- dense convolution as an implicit GEMM on tcgen05 tensor cores (sm_100a).
//
//   D[pixel, cout] = sum_{tap, cin} X[pixel shifted by tap, cin] * W[cout, tap, cin]
//   y = act(D * scale + shift + residual)
//
// One persistent CTA per SM, 6 warps, warp-specialised:
//   warp 0      : TMA producer. A tiles are 4-D boxes {K-block channels, BW, BH, 1} of the NHWC
//                 activation (BW*BH = 128 output pixels); the tap shift is a coordinate offset and the
//                 conv zero padding is TMA out-of-bounds zero fill.  Stride-2 convs read one of up
//                 to four parity views (doubled strides) of the same tensor.  B tiles are 2-D boxes
//                 {K-block, BN} of the packed K-major weights.  Both land in 128B/64B/32B-swizzled smem.
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (M=128, N<=256, K=16 per instr),
//                 fp32 accumulators in TMEM, two accumulator stages (2 x 256 columns) so the epilogue
//                 of tile i overlaps the main loop of tile i+1.
//   warps 2..5  : epilogue. tcgen05.ld 32 columns at a time -> scale/shift (+residual) -> activation
//                 -> bf16/fp16 -> 128B-swizzled smem staging -> TMA store (which clips the M and N tails
//                 and writes straight into a channel slice of the consumer's buffer).
// smem ring: num_stages x {A 128 x bk_bytes, B bn x bk_bytes} in the first 192 KB, 2 x 16 KB store
// staging, then the control block (mbarriers, TMEM base, scale/shift of the current N tile).
#include "common.cuh"
#include "../../include/segb200.h"

#include <mutex>
#include <stddef.h>
#include <stdio.h>

namespace segb200 {

constexpr int kStageRegion = 196608;          // bytes for the A/B ring
constexpr int kEpiBufBytes = 16384;           // 128 rows x 64 ch x 2 B
constexpr int kCtlOffset = kStageRegion + 2 * kEpiBufBytes;
constexpr int kResRegion = 2 * kEpiBufBytes;  // residual tiles live at the top of the ring region when used
constexpr int kSmemBytes = kCtlOffset + 2432;  // 231680 <= 232448
constexpr int kMaxStages = 16;
// kEpiGroups (template parameter): number of 4-warp epilogue groups.  Two groups put two epilogue warps on every SM
// sub-partition, which hides the TMEM-load / smem latencies of the epilogue: HBM-bound layers (K <= 512) go from 4.2 to
// 6.1 TB/s; tensor-bound layers lose 2-6 % to the extra warps, so the host picks per launch.

struct Control {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t res_full[2];
  uint64_t b_full;      // weights resident in smem (B-stationary mode)
  uint64_t pad64;       // keeps scale/shift 16-byte aligned (float4 loads)
  uint32_t tmem_base;
  uint32_t pad[3];
  float scale[256];
  float shift[256];
};
static_assert(sizeof(Control) <= 2432, "control block too large");
static_assert(offsetof(Control, scale) % 16 == 0 && offsetof(Control, shift) % 16 == 0, "scale/shift must be 16 B aligned");

struct ConvGemmParams {
  int n_img, ho, wo;
  int bw, bh, wtiles, htiles;
  int bi;               // images per M tile (the TMA box spans bi images); n_img counts image GROUPS
  int n_tiles, total_tiles;
  int cout, bn;
  int cblocks, ntaps, bk_bytes, num_stages;
  int group;            // single-CTA kernel: k-blocks per ring stage (one full/empty hand-shake per `group` k-blocks)
  int mma_pairs;        // CTA-pair kernel: consume k-blocks two at a time (needs a deep ring)
  int kmma_tail;        // K=16 MMA steps of the LAST channel block of a tap (the zero-filled K tail beyond it is skipped)
  int a_stage_bytes, b_stage_bytes;
  int b_resident;       // 1: all K blocks of the (single) N tile stay in smem for the whole kernel; the ring holds A only
  int ring_bytes;       // A/B ring region (its top 32 KB hold the residual tiles when a residual is fused); default 192 KB
  int act;
  int out_f32;          // 1: y is fp32 (32-column TMA store units), no residual
  const float* scale;
  const float* shift;
  const void* residual;
  long long res_ld;
  unsigned long long* dbg;   // only used when compiled with -DSEGB200_DBG
  int dbg_mode;              // DBG build only: 1 = skip the MMAs, 2 = skip the TMA loads, 3 = skip the epilogue work (resource decomposition)
  uint32_t taps[64];   // map id (bits 0..1) | (off_w + 128) << 8 | (off_h + 128) << 16
};

// Diagnostics build (-DSEGB200_DBG): per-role wait-cycle counters, see segb200_debug_set_counters in the header.
#ifdef SEGB200_DBG
#define TIMED_WAIT(bar, parity, slot)                                             \
  do {                                                                            \
    if (p.dbg != nullptr) {                                                       \
      const long long t0_ = clock64();                                            \
      mbar_wait(bar, parity);                                                     \
      atomicAdd(p.dbg + (slot), (unsigned long long)(clock64() - t0_));           \
    } else {                                                                      \
      mbar_wait(bar, parity);                                                     \
    }                                                                             \
  } while (0)
#define DBG_T0() const long long dbg_t0_ = (p.dbg != nullptr) ? clock64() : 0
#define DBG_ADD(slot) do { if (p.dbg != nullptr) atomicAdd(p.dbg + (slot), (unsigned long long)(clock64() - dbg_t0_)); } while (0)
#define DBG_COUNT(slot, v) do { if (p.dbg != nullptr) atomicAdd(p.dbg + (slot), (unsigned long long)(v)); } while (0)
#else
#define TIMED_WAIT(bar, parity, slot) mbar_wait(bar, parity)
#define DBG_T0() do {} while (0)
#define DBG_ADD(slot) do {} while (0)
#define DBG_COUNT(slot, v) do {} while (0)
#endif

// kDual: TWO tile streams per CTA.  Stream s owns the CTA's tiles j = s (mod 2) -- hence accumulator stage s --, half of the ring
// stages, and its own producer warp and MMA-issuing warp (two extra warps after the epilogue warps).  The tensor pipe interleaves
// the two instruction streams, so the several hundred cycles one issuing thread spends on a full/empty hand-shake are covered by
// the other stream's MMAs (tools/mma_probe.py mode 8: 128 cycles per 128x256x16 MMA with two issuers against 239 with one, both
// running this kernel's ring protocol).  The epilogue is unchanged: it drains tile j from accumulator j & 1 in tile order.
template <bool kBF16, int kEpiGroups, bool kDual>
__global__ void __launch_bounds__(64 + 128 * kEpiGroups + (kDual ? 64 : 0), 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmA3,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                 const __grid_constant__ CUtensorMap tmR, const __grid_constant__ ConvGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Control* ctl = reinterpret_cast<Control*>(smem + p.ring_bytes + 2 * kEpiBufBytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int sub_bytes = p.a_stage_bytes + (p.b_resident ? 0 : p.b_stage_bytes);    // one k-block: A 128 x bk (+ B bn x bk)
  const int stage_bytes = p.group * sub_bytes;                                      // a ring stage holds `group` k-blocks
  const int num_kb = p.ntaps * p.cblocks;
  const int bk_elems = p.bk_bytes >> 1;
  constexpr int kStreams = kDual ? 2 : 1;
  constexpr int kWarp2 = 2 + 4 * kEpiGroups;                  // dual: producer of stream 1 (kWarp2) and its MMA issuer (kWarp2 + 1)
  const int strm = (kDual && warp >= kWarp2) ? 1 : 0;
  const int st_n = kDual ? (p.num_stages >> 1) : p.num_stages;   // ring stages of one stream
  const int st_base = strm * st_n;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("segb200: dynamic smem base not 1024B aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0); prefetch_tmap(&tmB); prefetch_tmap(&tmC);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < p.num_stages; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&ctl->tmem_full[i], 1); mbar_init(&ctl->tmem_empty[i], 128 * kEpiGroups); mbar_init(&ctl->res_full[i], 1);
      }
      mbar_init(&ctl->b_full, 1);
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
  // set-up done (barriers, TMEM, tensor-map prefetch): let the next kernel of the stream be scheduled onto SMs as they free up, and
  // only now wait for the previous kernel's results (programmatic dependent launch; no-ops without the launch attribute)
  pdl_launch_dependents();
  pdl_wait();
#ifdef SEGB200_DBG
  const long long dbg_kernel_t0 = clock64();
#endif

  if (warp == 0 || (kDual && warp == kWarp2)) {
    // ------------------------------ TMA producer (of stream `strm`) ------------------------------
    if (lane == 0) {
      const CUtensorMap* amaps[4] = {&tmA0, &tmA1, &tmA2, &tmA3};
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx = (uint32_t)(128 * p.bk_bytes + (p.b_resident ? 0 : p.bn * p.bk_bytes));
      if (p.b_resident && strm == 0) {          // weights: loaded once, [num_kb] tiles of {bk, bn} behind the A ring
        mbar_expect_tx(&ctl->b_full, (uint32_t)(num_kb * p.bn * p.bk_bytes));
        for (int kb = 0; kb < num_kb; ++kb)
          tma_load_2d(&tmB, &ctl->b_full, smem + p.num_stages * stage_bytes + kb * p.b_stage_bytes, kb * bk_elems, 0);
      }
      for (int tile = blockIdx.x + strm * gridDim.x; tile < p.total_tiles; tile += kStreams * gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        int m_tile = tile / p.n_tiles;
        const int wb = m_tile % p.wtiles; m_tile /= p.wtiles;
        const int hb = m_tile % p.htiles;
        const int img = m_tile / p.htiles;
        const int w0 = wb * p.bw, h0 = hb * p.bh, n0 = n_tile * p.bn;
        int tap = 0, cb = 0;
        for (int kb0 = 0; kb0 < num_kb; kb0 += p.group) {
          const int gn = num_kb - kb0 < p.group ? num_kb - kb0 : p.group;
          TIMED_WAIT(&ctl->empty[st_base + stage], phase ^ 1, 0);
#ifdef SEGB200_DBG
          if (p.dbg_mode == 2) {
            tap += (cb + gn) / p.cblocks; cb = (cb + gn) % p.cblocks;
            mbar_arrive(&ctl->full[st_base + stage]); if (++stage == st_n) { stage = 0; phase ^= 1; } continue;
          }
#endif
          uint64_t* fullb = &ctl->full[st_base + stage];
          mbar_expect_tx(fullb, tx * (uint32_t)gn);
          uint8_t* sa = smem + (st_base + stage) * stage_bytes;
          for (int g = 0; g < gn; ++g, sa += sub_bytes) {
            const uint32_t t = p.taps[tap];
            const int ow = (int)((t >> 8) & 0xff) - 128, oh = (int)((t >> 16) & 0xff) - 128;
            tma_load_4d(amaps[t & 3], fullb, sa, cb * bk_elems, w0 + ow, h0 + oh, img * p.bi);
            if (!p.b_resident) tma_load_2d(&tmB, fullb, sa + p.a_stage_bytes, (kb0 + g) * bk_elems, n0);
            if (++cb == p.cblocks) { cb = 0; ++tap; }
          }
          if (++stage == st_n) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 || (kDual && warp == kWarp2 + 1)) {
    // ------------------------------ MMA issuer ------------------------------
    // tcgen05.mma issue is nearly synchronous: the tensor pipe queues about one instruction ahead of the one it executes
    // (tools/mma_probe.py: every instruction this thread executes between two MMAs beyond ~128 cycles of slack is a cycle the pipe
    // idles).  A single thread runs at one dependent instruction per ~4-6 cycles, so the k-block boundary (commit, barrier wait,
    // fence, descriptors, loop control) is written for minimum instruction count: every kernel parameter is hoisted into registers,
    // the descriptors' low words advance by one multiply-add per k-block (the high words are constants), the channel-block counter
    // replaces a modulo, and barrier addresses are 32-bit shared addresses computed once.
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const int kmma = p.bk_bytes >> 5;          // K=16 elements (32 B) per instruction
      const int kmma_tail = p.kmma_tail, cblocks = p.cblocks;
      const uint32_t nstages = (uint32_t)st_n, sbase = (uint32_t)st_base;
      const bool b_res_mode = p.b_resident != 0;
      const uint32_t full0 = smem_u32(&ctl->full[0]), empty0 = smem_u32(&ctl->empty[0]);
      // descriptor = {lo: start>>4 | LBO(1)<<16, hi: SBO | version<<14 | layout<<29}  (make_kmajor_desc)
      const uint64_t d0 = make_kmajor_desc(smem_u32(smem), (uint32_t)p.bk_bytes);
      const uint32_t desc_hi = (uint32_t)(d0 >> 32), a_lo0 = (uint32_t)d0;
      const uint32_t stage_step = (uint32_t)stage_bytes >> 4, b_off = (uint32_t)p.a_stage_bytes >> 4;
      const uint32_t bres_lo0 = a_lo0 + (uint32_t)p.num_stages * stage_step, bres_step = (uint32_t)p.b_stage_bytes >> 4;
      if (p.b_resident) mbar_wait(&ctl->b_full, 0);
      // tile j of this CTA uses accumulator stage j & 1 in its (j >> 1)-th use; a dual stream walks j = strm, strm + 2, ...
      for (int j = strm, tile = blockIdx.x + strm * gridDim.x; tile < p.total_tiles; j += kStreams, tile += kStreams * gridDim.x) {
        const int acc = j & 1; const uint32_t acc_phase = (uint32_t)(j >> 1) & 1u;
        const int n_tile = tile % p.n_tiles;
        const int n0 = n_tile * p.bn;
        int nvalid = p.cout - n0; if (nvalid > p.bn) nvalid = p.bn;
        const uint32_t umma_n = (uint32_t)((nvalid + 15) & ~15);
        const uint32_t idesc = make_idesc(kBF16, umma_n);
        TIMED_WAIT(&ctl->tmem_empty[acc], acc_phase ^ 1, 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
        int cb = 0;
        // (Consuming k-blocks in pairs -- two barrier observations, then 8 MMAs back to back -- runs at 162 instead of 235 cycles per MMA
        // in the stand-alone probe (tools/mma_probe.py mode 6 vs 3) but LOSES in this kernel when a k-block is 48 KB: only 3-4 fit, and a
        // pair-granular ring halves the TMA prefetch distance (operand starvation 21 % -> 29 %, +res shapes 29 % -> 46 %;
        // profiles/r2_gemm_decomposition.md).  Small k-blocks are different: the host packs `group` of them into one ring stage
        // (<= 32 KB, so the ring stays >= 5 stages deep) and the whole group rides on ONE full/empty hand-shake.)
        const uint32_t sub_step = (uint32_t)sub_bytes >> 4;
        const int group = p.group;
        for (int kb0 = 0; kb0 < num_kb; kb0 += group) {
          const int gn = num_kb - kb0 < group ? num_kb - kb0 : group;
#ifdef SEGB200_DBG
          TIMED_WAIT(&ctl->full[sbase + stage], phase, 2);
#else
          mbar_wait_guarded(full0 + (sbase + stage) * 8, phase);
#endif
          tc_fence_after();
          uint32_t a_lo = a_lo0 + (sbase + stage) * stage_step;
          for (int g = 0; g < gn; ++g, a_lo += sub_step) {
            const int kb = kb0 + g;
            const uint32_t b_lo = b_res_mode ? bres_lo0 + (uint32_t)kb * bres_step : a_lo + b_off;
            int kcnt = kmma;
            if (++cb == cblocks) { cb = 0; kcnt = kmma_tail; }
#ifdef SEGB200_DBG
            if (p.dbg_mode != 1)
#endif
            for (int k = 0; k < kcnt; ++k)
              umma_f16(d_tmem, ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2u * (uint32_t)k), ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2u * (uint32_t)k),
                       idesc, (uint32_t)(kb | k));
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty0 + (sbase + stage) * 8) : "memory");
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&ctl->tmem_full[acc]);            // accumulator ready for the epilogue
      }
    }
    __syncwarp();
  } else if (!kDual || warp < kWarp2) {
    // ------------------------------ epilogue (warps 2..5 [, 6..9]) ------------------------------
    // kEpiGroups groups of 4 warps walk the same (tile, 64-column chunk) sequence; with two groups, group g owns the chunks
    // whose running index is g (mod 2), plus its own staging buffer, residual buffer, named barrier and bulk-store group.
    using H = Half2<kBF16>;
    const int grp = (warp - 2) >> 2;
    const int et = (threadIdx.x - 64) & 127;         // 0..127 within the group
    const int bar_id = 1 + grp;
    const int q = warp & 3;                          // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;                   // tile row == TMEM lane
    uint8_t* epi = smem + p.ring_bytes;              // 2 store-staging buffers
    uint8_t* resb = smem + p.ring_bytes - kResRegion;  // 2 residual buffers (ring is shortened by the host)
    const bool has_res = p.residual != nullptr;
    int acc = 0; uint32_t acc_phase = 0; uint32_t chunk_ctr = 0, my_ctr = 0; int staged_n_tile = -1;

    // residual prefetch cursor (group leader only) over the global (tile, chunk) sequence
    int pf_tile = blockIdx.x, pf_ch = 0;
    auto pf_step = [&](bool load, uint32_t bi) {
      // optionally issue the TMA load of the residual chunk under the cursor into resb[bi], then advance the cursor
      if (pf_tile >= p.total_tiles) return;
      const int n_tile = pf_tile % p.n_tiles;
      int m_tile = pf_tile / p.n_tiles;
      const int wb = m_tile % p.wtiles; m_tile /= p.wtiles;
      const int hb = m_tile % p.htiles;
      const int img = m_tile / p.htiles;
      const int n0 = n_tile * p.bn;
      int nvalid = p.cout - n0; if (nvalid > p.bn) nvalid = p.bn;
      if (load) {
        mbar_expect_tx(&ctl->res_full[bi], (uint32_t)kEpiBufBytes);
        tma_load_4d(&tmR, &ctl->res_full[bi], resb + bi * kEpiBufBytes, n0 + pf_ch * 64, wb * p.bw, hb * p.bh, img * p.bi);
      }
      if (++pf_ch >= ((nvalid + 63) >> 6)) { pf_ch = 0; pf_tile += gridDim.x; }
    };
    if (has_res && et == 0) {
      if (kEpiGroups == 1) pf_step(true, 0);
      else { if (grp == 1) pf_step(false, 0); pf_step(true, (uint32_t)grp); }
    }

    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      int m_tile = tile / p.n_tiles;
      const int wb = m_tile % p.wtiles; m_tile /= p.wtiles;
      const int hb = m_tile % p.htiles;
      const int img = m_tile / p.htiles;
      const int w0 = wb * p.bw, h0 = hb * p.bh, n0 = n_tile * p.bn;
      int nvalid = p.cout - n0; if (nvalid > p.bn) nvalid = p.bn;
      const int nchunks = (nvalid + 63) >> 6;

#ifndef SEGB200_EXP
#define SEGB200_EXP 0
#endif
      if (n_tile != staged_n_tile) {                   // (re)stage the folded-BN scale/shift when the N tile changes
        if (kEpiGroups == 1) named_bar_sync(1, 128); else named_bar_sync(3, 256);   // previous readers are done
        if (grp == 0) {
          for (int i = et; i < p.bn; i += 128) {
            const int c = n0 + i;
            ctl->scale[i] = (p.scale != nullptr && c < p.cout) ? p.scale[c] : 1.f;
            ctl->shift[i] = (p.shift != nullptr && c < p.cout) ? p.shift[c] : 0.f;
          }
        }
        if (kEpiGroups == 2) named_bar_sync(3, 256);   // publish to the other group (one group: the chunk barrier does)
        staged_n_tile = n_tile;
      }

      if (et == 0) { TIMED_WAIT(&ctl->tmem_full[acc], acc_phase, 3); } else { mbar_wait(&ctl->tmem_full[acc], acc_phase); }
      tc_fence_after();
#ifdef SEGB200_DBG
      const long long dbg_epi_t0 = clock64();
      if (p.dbg_mode == 3) {
        chunk_ctr += (uint32_t)nchunks;
        tc_fence_before();
        mbar_arrive(&ctl->tmem_empty[acc]);
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
        continue;
      }
#endif
      const uint32_t t_acc = tmem_base + (uint32_t)(acc * 256) + ((uint32_t)(q * 32) << 16);
      if (p.out_f32) {
        if (grp == 0) {
        // fp32 output (attention energies etc.): one 32-column x 128-row unit (128 B rows) per staging buffer / TMA store
        const int nunits = (nvalid + 31) >> 5;
        for (int un = 0; un < nunits; ++un, ++chunk_ctr) {
          uint8_t* buf = epi + (chunk_ctr & 1) * kEpiBufBytes;
          if (et == 0) tma_store_wait_read<1>();
          named_bar_sync(1, 128);
          const int col0 = un * 32;
          uint32_t v[32];
          tmem_ld_32x32(t_acc + (uint32_t)col0, v);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 8; ++g) {                // 8 groups of 4 floats = one 16 B vector each
            float f[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = col0 + g * 4 + j;
              f[j] = apply_act(fmaf(__uint_as_float(v[g * 4 + j]), ctl->scale[c], ctl->shift[c]), p.act);
            }
            *reinterpret_cast<float4*>(buf + row * 128 + ((g ^ (row & 7)) * 16)) = make_float4(f[0], f[1], f[2], f[3]);
          }
          fence_proxy_async();
          named_bar_sync(1, 128);
          if (et == 0) {
            tma_store_4d(&tmC, buf, n0 + col0, w0, h0, img * p.bi);
            tma_store_commit();
          }
        }
        }
      } else
      for (int ch = 0; ch < nchunks; ++ch, ++chunk_ctr) {
        if (kEpiGroups == 2 && (int)(chunk_ctr & 1) != grp) continue;
        const uint32_t bi = kEpiGroups == 1 ? (my_ctr & 1) : (uint32_t)grp;   // staging / residual buffer of this chunk
        uint8_t* buf = epi + bi * kEpiBufBytes;
        const uint8_t* rbuf = resb + bi * kEpiBufBytes;
        if (et == 0) {
          if (kEpiGroups == 1) {
            tma_store_wait_read<1>();                // the store that last used `buf` has drained it
            if (has_res) pf_step(true, (my_ctr + 1) & 1);   // resb[other] was last read before the previous chunk's barrier
          } else {
            tma_store_wait_read<0>();
          }
        }
        named_bar_sync(bar_id, 128);                 // (one group: also publishes scale/shift on the first chunk)
        if (has_res) mbar_wait(&ctl->res_full[bi], kEpiGroups == 1 ? ((my_ctr >> 1) & 1) : (my_ctr & 1));
        ++my_ctr;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int col0 = ch * 64 + half * 32;
          uint32_t v[32];
          tmem_ld_32x32(t_acc + (uint32_t)col0, v);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {              // 4 groups of 8 channels = one 16 B vector each
            const int chunk16 = (half * 4 + g) ^ (row & 7);   // 128B swizzle: 16 B chunk c lives at c ^ (row & 7)
            const int cl = col0 + g * 8;
            const float4 s0 = *reinterpret_cast<const float4*>(&ctl->scale[cl]);
            const float4 s1 = *reinterpret_cast<const float4*>(&ctl->scale[cl + 4]);
            const float4 h0 = *reinterpret_cast<const float4*>(&ctl->shift[cl]);
            const float4 h1 = *reinterpret_cast<const float4*>(&ctl->shift[cl + 4]);
            float f[8];
            f[0] = fmaf(__uint_as_float(v[g * 8 + 0]), s0.x, h0.x); f[1] = fmaf(__uint_as_float(v[g * 8 + 1]), s0.y, h0.y);
            f[2] = fmaf(__uint_as_float(v[g * 8 + 2]), s0.z, h0.z); f[3] = fmaf(__uint_as_float(v[g * 8 + 3]), s0.w, h0.w);
            f[4] = fmaf(__uint_as_float(v[g * 8 + 4]), s1.x, h1.x); f[5] = fmaf(__uint_as_float(v[g * 8 + 5]), s1.y, h1.y);
            f[6] = fmaf(__uint_as_float(v[g * 8 + 6]), s1.z, h1.z); f[7] = fmaf(__uint_as_float(v[g * 8 + 7]), s1.w, h1.w);
            if (has_res) {
              const uint4 r = *reinterpret_cast<const uint4*>(rbuf + row * 128 + chunk16 * 16);
              const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 t2 = H::unpack(rr[j]);
                f[2 * j] += t2.x; f[2 * j + 1] += t2.y;
              }
            }
            // activation on the PACKED 16-bit pairs: round(max(v,0)) == max(round(v),0) and 6.0 is representable
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              pk[j] = H::pack(f[2 * j], f[2 * j + 1]);
              if (p.act != ACT_NONE) pk[j] = H::relu2(pk[j]);
              if (p.act == ACT_RELU6) pk[j] = H::min2(pk[j], 6.f);
            }
            *reinterpret_cast<uint4*>(buf + row * 128 + chunk16 * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        fence_proxy_async();
        named_bar_sync(bar_id, 128);
        if (et == 0) {
          tma_store_4d(&tmC, buf, n0 + ch * 64, w0, h0, img * p.bi);
          tma_store_commit();
          if (kEpiGroups == 2 && has_res) { pf_step(false, 0); pf_step(true, (uint32_t)grp); }   // rbuf is free again
        }
      }
      tc_fence_before();
      mbar_arrive(&ctl->tmem_empty[acc]);             // 128 * kEpiGroups arrivals release the accumulator stage
#ifdef SEGB200_DBG
      if (p.dbg != nullptr && et == 0 && grp == 0) { atomicAdd(p.dbg + 4, (unsigned long long)(clock64() - dbg_epi_t0)); atomicAdd(p.dbg + 6, 1ull); }
#endif
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
    if (et == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
#ifdef SEGB200_DBG
  if (p.dbg != nullptr && threadIdx.x == 0) atomicAdd(p.dbg + 7, (unsigned long long)(clock64() - dbg_kernel_t0));
#endif
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ---------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) variant: two CTAs of a cluster compute one 256 x BN tile.  Each CTA stages ITS 128 rows of A and
// ITS half of the BN weight rows, so a stage is 16 KB + 16 KB instead of 16 KB + 32 KB: a third fewer operand bytes per FLOP
// per SM and 6 ring stages instead of 4.  (The single-CTA kernel is bandwidth-delay bound on its ring: profiles/.)
//   * both CTAs run the TMA producer; the peer's loads signal the LEADER's `full` barrier (cta_group::2 TMA form);
//   * only the leader issues tcgen05.mma.cta_group::2 (M = 256); its commits are multicast to both CTAs' barriers;
//   * each CTA's epilogue drains its own 128 TMEM lanes; the peer's epilogue threads arrive on the leader's tmem_empty.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (once all previously issued MMAs retire) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// TMA loads whose completion bytes are credited to the barrier of the pair's leader CTA (bit 24 of the cluster address = rank)
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* map, uint32_t leader_bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(const CUtensorMap* map, uint32_t leader_bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ uint32_t make_idesc_m256(bool bf16, uint32_t n) {
  const uint32_t fmt = bf16 ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}

template <bool kBF16, int kEpiGroups>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 128 * kEpiGroups, 1)
conv_gemm2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmA3,
                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                  const __grid_constant__ CUtensorMap tmR, const __grid_constant__ ConvGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Control* ctl = reinterpret_cast<Control*>(smem + p.ring_bytes + 2 * kEpiBufBytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int stage_bytes = p.a_stage_bytes + p.b_stage_bytes;           // b_stage_bytes = this CTA's HALF of the N tile
  const int num_kb = p.ntaps * p.cblocks;
  const int bk_elems = p.bk_bytes >> 1;
  const int half_n = p.bn >> 1;
  const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  const int m_tiles = p.wtiles * p.htiles * p.n_img;
  const int pair_tiles = ((m_tiles + 1) >> 1) * p.n_tiles;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("segb200: dynamic smem base not 1024B aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA0); prefetch_tmap(&tmB); prefetch_tmap(&tmC); }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < p.num_stages; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&ctl->tmem_full[i], 1);
        mbar_init(&ctl->tmem_empty[i], 2 * 128 * kEpiGroups);          // both CTAs' epilogue threads (used on the leader only)
        mbar_init(&ctl->res_full[i], 1);
      }
      mbar_init(&ctl->b_full, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc2(&ctl->tmem_base, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                                  // barriers of both CTAs exist before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  auto decode = [&](int t, int& n_tile, int& wb, int& hb, int& img) {
    n_tile = t % p.n_tiles;
    int m_tile = (t / p.n_tiles) * 2 + (int)rank;                       // a phantom tile (m_tile >= m_tiles) decodes to img >= n_img:
    wb = m_tile % p.wtiles; m_tile /= p.wtiles;                         // its loads are zero-filled, its stores clipped
    hb = m_tile % p.htiles;
    img = m_tile / p.htiles;
  };

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    if (lane == 0) {
      const CUtensorMap* amaps[4] = {&tmA0, &tmA1, &tmA2, &tmA3};
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx_pair = (uint32_t)(2 * (128 * p.bk_bytes + half_n * p.bk_bytes));
      for (int tile = cluster_id; tile < pair_tiles; tile += n_clusters) {
        int n_tile, wb, hb, img;
        decode(tile, n_tile, wb, hb, img);
        const int w0 = wb * p.bw, h0 = hb * p.bh, n0 = n_tile * p.bn + (int)rank * half_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / p.cblocks;
          const int cb = kb - tap * p.cblocks;
          const uint32_t t = p.taps[tap];
          const int ow = (int)((t >> 8) & 0xff) - 128, oh = (int)((t >> 16) & 0xff) - 128;
          mbar_wait(&ctl->empty[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&ctl->full[stage], tx_pair);
          const uint32_t lbar = mapa_u32(smem_u32(&ctl->full[stage]), 0);      // the LEADER's barrier (shared::cluster address)
          uint8_t* sa = smem + stage * stage_bytes;
          tma2_load_4d(amaps[t & 3], lbar, sa, cb * bk_elems, w0 + ow, h0 + oh, img * p.bi);
          tma2_load_2d(&tmB, lbar, sa + p.a_stage_bytes, kb * bk_elems, n0);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) ------------------------------
    // Same lean loop as the single-CTA kernel.  Here one barrier observation feeds 4 MMAs of M = 256 (twice the FLOPs), and the
    // 32 KB stages leave 5-6 ring slots, so k-blocks can be consumed in PAIRS (two observations, then 8 MMAs back to back:
    // tools/mma_probe.py mode 6) without starving the TMA prefetch -- p.mma_pairs, set by the host when num_stages >= 5.
    if (lane == 0 && leader) {
      uint32_t stage = 0, phase = 0; int acc = 0; uint32_t acc_phase = 0;
      const int kmma = p.bk_bytes >> 5;
      const int kmma_tail = p.kmma_tail, cblocks = p.cblocks;
      const uint32_t nstages = (uint32_t)p.num_stages;
      const uint32_t full0 = smem_u32(&ctl->full[0]), empty0 = smem_u32(&ctl->empty[0]);
      const uint64_t d0 = make_kmajor_desc(smem_u32(smem), (uint32_t)p.bk_bytes);
      const uint32_t desc_hi = (uint32_t)(d0 >> 32), a_lo0 = (uint32_t)d0;
      const uint32_t stage_step = (uint32_t)stage_bytes >> 4, b_off = (uint32_t)p.a_stage_bytes >> 4;
      const int group = p.mma_pairs ? 2 : 1;
      // the instruction's N spans both halves: CTA r supplies columns [r*half_n, r*half_n + N/2); keep the split at half_n
      const uint32_t idesc = make_idesc_m256(kBF16, (uint32_t)p.bn);
      for (int tile = cluster_id; tile < pair_tiles; tile += n_clusters) {
        mbar_wait(&ctl->tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
        int cb = 0;
        for (int kb = 0; kb < num_kb; kb += group) {
          const bool two = group == 2 && kb + 1 < num_kb;
          uint32_t s1 = stage + 1, ph1 = phase;
          if (s1 == nstages) { s1 = 0; ph1 ^= 1; }
          mbar_wait_guarded(full0 + stage * 8, phase);
          if (two) mbar_wait_guarded(full0 + s1 * 8, ph1);
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 1 && !two) break;
            const uint32_t st = h == 0 ? stage : s1;
            const int kbb = kb + h;
            const uint32_t a_lo = a_lo0 + st * stage_step, b_lo = a_lo + b_off;
            int kcnt = kmma;
            if (++cb == cblocks) { cb = 0; kcnt = kmma_tail; }
            for (int k = 0; k < kcnt; ++k)
              umma2_f16(d_tmem, ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2u * (uint32_t)k), ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2u * (uint32_t)k),
                        idesc, (uint32_t)(kbb | k));
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(empty0 + st * 8), "h"((uint16_t)3) : "memory");
          }
          if (two) { stage = s1 + 1; phase = ph1; if (stage == nstages) { stage = 0; phase ^= 1; } }
          else { stage = s1; phase = ph1; }
        }
        umma2_commit_both(&ctl->tmem_full[acc]);
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue (warps 2..5 [, 6..9]) of BOTH CTAs: own 128 rows ------------------------------
    using H = Half2<kBF16>;
    const int grp = (warp - 2) >> 2;
    const int et = (threadIdx.x - 64) & 127;
    const int bar_id = 1 + grp;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint8_t* epi = smem + p.ring_bytes;
    uint8_t* resb = smem + p.ring_bytes - kResRegion;
    const bool has_res = p.residual != nullptr;
    int acc = 0; uint32_t acc_phase = 0; uint32_t chunk_ctr = 0, my_ctr = 0; int staged_n_tile = -1;
    const uint32_t leader_empty0 = mapa_u32(smem_u32(&ctl->tmem_empty[0]), 0), leader_empty1 = mapa_u32(smem_u32(&ctl->tmem_empty[1]), 0);

    int pf_tile = cluster_id, pf_ch = 0;
    auto pf_step = [&](bool load, uint32_t bi) {
      if (pf_tile >= pair_tiles) return;
      int n_tile, wb, hb, img;
      decode(pf_tile, n_tile, wb, hb, img);
      const int n0 = n_tile * p.bn;
      int nvalid = p.cout - n0; if (nvalid > p.bn) nvalid = p.bn;
      if (load) {
        mbar_expect_tx(&ctl->res_full[bi], (uint32_t)kEpiBufBytes);
        tma_load_4d(&tmR, &ctl->res_full[bi], resb + bi * kEpiBufBytes, n0 + pf_ch * 64, wb * p.bw, hb * p.bh, img * p.bi);
      }
      if (++pf_ch >= ((nvalid + 63) >> 6)) { pf_ch = 0; pf_tile += n_clusters; }
    };
    if (has_res && et == 0) {
      if (kEpiGroups == 1) pf_step(true, 0);
      else { if (grp == 1) pf_step(false, 0); pf_step(true, (uint32_t)grp); }
    }

    for (int tile = cluster_id; tile < pair_tiles; tile += n_clusters) {
      int n_tile, wb, hb, img;
      decode(tile, n_tile, wb, hb, img);
      const int w0 = wb * p.bw, h0 = hb * p.bh, n0 = n_tile * p.bn;
      int nvalid = p.cout - n0; if (nvalid > p.bn) nvalid = p.bn;
      const int nchunks = (nvalid + 63) >> 6;

      if (n_tile != staged_n_tile) {
        if (kEpiGroups == 1) named_bar_sync(1, 128); else named_bar_sync(3, 256);
        if (grp == 0) {
          for (int i = et; i < p.bn; i += 128) {
            const int c = n0 + i;
            ctl->scale[i] = (p.scale != nullptr && c < p.cout) ? p.scale[c] : 1.f;
            ctl->shift[i] = (p.shift != nullptr && c < p.cout) ? p.shift[c] : 0.f;
          }
        }
        if (kEpiGroups == 2) named_bar_sync(3, 256);
        staged_n_tile = n_tile;
      }

      mbar_wait(&ctl->tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + (uint32_t)(acc * 256) + ((uint32_t)(q * 32) << 16);
      for (int ch = 0; ch < nchunks; ++ch, ++chunk_ctr) {
        if (kEpiGroups == 2 && (int)(chunk_ctr & 1) != grp) continue;
        const uint32_t bi = kEpiGroups == 1 ? (my_ctr & 1) : (uint32_t)grp;
        uint8_t* buf = epi + bi * kEpiBufBytes;
        const uint8_t* rbuf = resb + bi * kEpiBufBytes;
        if (et == 0) {
          if (kEpiGroups == 1) {
            tma_store_wait_read<1>();
            if (has_res) pf_step(true, (my_ctr + 1) & 1);
          } else {
            tma_store_wait_read<0>();
          }
        }
        named_bar_sync(bar_id, 128);
        if (has_res) mbar_wait(&ctl->res_full[bi], kEpiGroups == 1 ? ((my_ctr >> 1) & 1) : (my_ctr & 1));
        ++my_ctr;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int col0 = ch * 64 + half * 32;
          uint32_t v[32];
          tmem_ld_32x32(t_acc + (uint32_t)col0, v);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int chunk16 = (half * 4 + g) ^ (row & 7);
            const int cl = col0 + g * 8;
            const float4 s0 = *reinterpret_cast<const float4*>(&ctl->scale[cl]);
            const float4 s1 = *reinterpret_cast<const float4*>(&ctl->scale[cl + 4]);
            const float4 h0v = *reinterpret_cast<const float4*>(&ctl->shift[cl]);
            const float4 h1v = *reinterpret_cast<const float4*>(&ctl->shift[cl + 4]);
            float f[8];
            f[0] = fmaf(__uint_as_float(v[g * 8 + 0]), s0.x, h0v.x); f[1] = fmaf(__uint_as_float(v[g * 8 + 1]), s0.y, h0v.y);
            f[2] = fmaf(__uint_as_float(v[g * 8 + 2]), s0.z, h0v.z); f[3] = fmaf(__uint_as_float(v[g * 8 + 3]), s0.w, h0v.w);
            f[4] = fmaf(__uint_as_float(v[g * 8 + 4]), s1.x, h1v.x); f[5] = fmaf(__uint_as_float(v[g * 8 + 5]), s1.y, h1v.y);
            f[6] = fmaf(__uint_as_float(v[g * 8 + 6]), s1.z, h1v.z); f[7] = fmaf(__uint_as_float(v[g * 8 + 7]), s1.w, h1v.w);
            if (has_res) {
              const uint4 r = *reinterpret_cast<const uint4*>(rbuf + row * 128 + chunk16 * 16);
              const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 t2 = H::unpack(rr[j]);
                f[2 * j] += t2.x; f[2 * j + 1] += t2.y;
              }
            }
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              pk[j] = H::pack(f[2 * j], f[2 * j + 1]);
              if (p.act != ACT_NONE) pk[j] = H::relu2(pk[j]);
              if (p.act == ACT_RELU6) pk[j] = H::min2(pk[j], 6.f);
            }
            *reinterpret_cast<uint4*>(buf + row * 128 + chunk16 * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        fence_proxy_async();
        named_bar_sync(bar_id, 128);
        if (et == 0) {
          tma_store_4d(&tmC, buf, n0 + ch * 64, w0, h0, img * p.bi);
          tma_store_commit();
          if (kEpiGroups == 2 && has_res) { pf_step(false, 0); pf_step(true, (uint32_t)grp); }
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(acc == 0 ? leader_empty0 : leader_empty1);   // 2 x 128 x groups arrivals release the pair's accumulator
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
    if (et == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                                  // neither CTA may exit while the peer still uses its smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static inline int floordiv2(int o, int* parity) {   // o = 2*a + p, p in {0,1}
  const int p = ((o % 2) + 2) % 2;
  *parity = p;
  return (o - p) / 2;
}

}  // namespace segb200

using namespace segb200;

static int g_dbg_mode = 0;
static int g_kgroup_kb = 40;   // largest ring stage a group of k-blocks may fill, in KB ("gemm_kgroup_kb"; measured: 40 KB / 9 best)
static int g_kgroup = 0;       // k-blocks per ring stage: 0 = auto (<= 32 KB, <= 4), 1 = off, n = cap ("gemm_kgroup")
static int g_mma_pairs = 1;    // CTA-pair kernel: pair-wise k-block consumption when the ring has >= 5 stages ("gemm_mma_pairs")
static int g_ring_kb = 0;
static int g_no_b_resident = 0;
static int g_no_img_tiles = 0;
static int g_epi2_maxk = 512;  // K (taps x padded cin) up to which the two-epilogue-group variant is used
static int g_2cta = 0;        // opt-in: CTA-pair kernel for the tensor-bound shapes
// Two tile streams per CTA (opt-in, "gemm_dual": 0 off (default), 1 auto, 2 wherever the ring allows).  Measured on B200
// (profiles/r2_gemm_decomposition.md section 5): the bare issue loops reach 128 cycles per MMA with two issuers against 239 with one,
// but in this kernel two streams halve each stream's ring -- 2 x 48 KB slots: 728->728 85 -> 96 us, 1536->2048 318 -> 355 us -- and
// half k-blocks (4 x 24 KB slots per stream, 64-byte rows) are slower still (conv_gemm 9.4 -> 11.3 ms per step).
static int g_dual = 0;
static int g_dual_subk = 0;   // dual streams walk half k-blocks (32 channels) so that each keeps >= 3 ring slots ("gemm_dual_subk")
static int g_dual_min_kb = 3; // auto: at least this many ring hand-shakes per tile ("gemm_dual_min_kb")
static int g_no_bn128 = 1;   // measured: 128-wide tiles lose 45 % on the 3x3 256->256 layers (operand traffic per FLOP up 33 %)
extern "C" int segb200_set_option(const char* name, int value) {
  if (name && !strcmp(name, "gemm_ring_kb")) { g_ring_kb = value; return 0; }
  if (name && !strcmp(name, "gemm_b_resident")) { g_no_b_resident = value ? 0 : 1; return 0; }
  if (name && !strcmp(name, "gemm_bn128")) { g_no_bn128 = value ? 0 : 1; return 0; }
  if (name && !strcmp(name, "gemm_2cta")) { g_2cta = value; return 0; }
  if (name && !strcmp(name, "gemm_img_tiles")) { g_no_img_tiles = value ? 0 : 1; return 0; }
  if (name && !strcmp(name, "gemm_epi2_maxk")) { g_epi2_maxk = value > 0 ? value : 512; return 0; }
  if (name && !strcmp(name, "dw_ring_slots")) return segb200::set_dw_ring_slots(value);
  if (name && !strcmp(name, "dw_v8")) return segb200::set_dw_v8(value);
  if (name && !strcmp(name, "dw_persistent")) return segb200::set_dw_persistent(value);
  if (name && !strcmp(name, "dw_cols2")) return segb200::set_dw_cols2(value);
  if (name && !strcmp(name, "dw_cw5")) return segb200::set_dw_cw5(value);
  if (name && !strcmp(name, "pdl")) return segb200::set_pdl(value);
  if (name && !strcmp(name, "gemm_dual")) { g_dual = value; return 0; }
  if (name && !strcmp(name, "gemm_dual_min_kb")) { g_dual_min_kb = value; return 0; }
  if (name && !strcmp(name, "gemm_dual_subk")) { g_dual_subk = value; return 0; }
  if (name && !strcmp(name, "bilinear_out_v1")) return segb200::set_bilinear_out_v1(value);
  if (name && !strcmp(name, "gemm_mma_pairs")) { g_mma_pairs = value; return 0; }
  if (name && !strcmp(name, "gemm_kgroup")) { g_kgroup = value; return 0; }
  if (name && !strcmp(name, "gemm_kgroup_kb")) { g_kgroup_kb = value > 0 ? value : 32; return 0; }
  if (name && !strcmp(name, "gemm_dbg_mode")) { g_dbg_mode = value; return 0; }     // effective in -DSEGB200_DBG builds only
  return set_error(-30, "segb200_set_option: unknown option '%s'", name ? name : "(null)");
}

static unsigned long long* g_dbg_counters = nullptr;
extern "C" int segb200_debug_set_counters(void* dev_ptr_16_u64) {
#ifdef SEGB200_DBG
  g_dbg_counters = reinterpret_cast<unsigned long long*>(dev_ptr_16_u64);
  return 0;
#else
  (void)dev_ptr_16_u64;
  return set_error(-20, "segb200_debug_set_counters: library was built without -DSEGB200_DBG");
#endif
}

extern "C" int segb200_conv_kblock(int cin) { return cin >= 64 ? 64 : (cin >= 32 ? 32 : 16); }

extern "C" int segb200_conv_gemm(const segb200_conv_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->x || !a->wgt || !a->y) return set_error(-1, "conv_gemm: null pointer argument");
  if (a->dtype != DT_BF16 && a->dtype != DT_F16) return set_error(-2, "conv_gemm: dtype must be bf16 or f16");
  if (a->stride != 1 && a->stride != 2) return set_error(-3, "conv_gemm: stride must be 1 or 2");
  if ((a->x_ld & 7) || (a->y_ld & (a->y_f32 ? 3 : 7)) || (a->cin & 7) || a->cout < 1 || a->cin > a->x_ld || a->cout > a->y_ld)
    return set_error(-4, "conv_gemm: cin and the pitches must be multiples of 8 elements (cin %d x_ld %d cout %d y_ld %d)",
                     a->cin, a->x_ld, a->cout, a->y_ld);
  if (a->residual && ((a->res_ld & 7) || a->res_ld < a->cout)) return set_error(-4, "conv_gemm: bad res_ld");
  if (a->y_f32 && a->residual) return set_error(-4, "conv_gemm: fp32 output does not take a residual");
  const int ntaps = a->kh * a->kw;
  if (ntaps < 1 || ntaps > 64) return set_error(-5, "conv_gemm: unsupported kernel %dx%d", a->kh, a->kw);
  if (a->n < 1 || a->ho < 1 || a->wo < 1) return set_error(-6, "conv_gemm: empty output");
  if (((uintptr_t)a->x & 15) || ((uintptr_t)a->y & 15) || ((uintptr_t)a->wgt & 15) || ((uintptr_t)a->residual & 15))
    return set_error(-7, "conv_gemm: pointers must be 16-byte aligned");

  int bk = segb200_conv_kblock(a->cin);            // k-block of the weight PACKING (cin is padded to a multiple of it per tap)
  int bk_bytes = bk * 2;
  int cblocks = (a->cin + bk - 1) / bk;
  const int cin_pad = cblocks * bk;
  const long long ktot = (long long)ntaps * cin_pad;

  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  // ---- M tiling: flattened pixels for contiguous 1x1 stride-1, else a BW x BH patch per image ----
  const bool flat = (ntaps == 1 && a->stride == 1 && a->pad_t == 0 && a->pad_l == 0 && a->ho == a->h && a->wo == a->w);
  long long Wv_out, Hv_out, Nv_out;   // output extents as seen by the C map
  if (flat) {
    Wv_out = (long long)a->n * a->h * a->w; Hv_out = 1; Nv_out = 1;
    p.bw = 128; p.bh = 1;
  } else {
    Wv_out = a->wo; Hv_out = a->ho; Nv_out = a->n;
    // 128 output pixels per tile as a BW x BH patch of BI images (BW*BH*BI = 128): the box may span several images, which
    // cuts the padding of small maps (65x129, batch 4: 4 x 81 patches of 16x8 = 324 tiles -> 289 tiles of 8x4x4) and with
    // it the rounds of the persistent grid
    long long best = -1;
    p.bi = 1;
    for (int bi = 1; bi <= 16 && bi <= a->n * 2 - 1; bi *= 2)
      for (int bw = 128 / bi; bw >= 1; bw /= 2) {
        const int bh = 128 / bi / bw;
        if (bw > 256 || bh > 256) continue;
        const long long t = (long long)((a->wo + bw - 1) / bw) * ((a->ho + bh - 1) / bh) * ((a->n + bi - 1) / bi);
        if (best < 0 || t < best) { best = t; p.bw = bw; p.bh = bh; p.bi = bi; }
      }
    if (g_no_img_tiles && p.bi != 1) {                 // (tuning knob: single-image patches only)
      best = -1; p.bi = 1;
      for (int bw = 128; bw >= 1; bw /= 2) {
        const int bh = 128 / bw;
        const long long t = (long long)((a->wo + bw - 1) / bw) * ((a->ho + bh - 1) / bh) * a->n;
        if (best < 0 || t < best) { best = t; p.bw = bw; p.bh = bh; }
      }
    }
  }
  if (flat) p.bi = 1;
  p.n_img = (int)((Nv_out + p.bi - 1) / p.bi); p.ho = (int)Hv_out; p.wo = (int)Wv_out;
  if (Wv_out > 0x7fffffffLL) return set_error(-8, "conv_gemm: too many pixels");
  p.wtiles = (int)((Wv_out + p.bw - 1) / p.bw);
  p.htiles = (int)((Hv_out + p.bh - 1) / p.bh);
  // ---- N tiling ----
  p.cout = a->cout;
  p.bn = a->cout >= 256 ? 256 : ((a->cout + 15) & ~15);
  if (a->cout >= 256 && !g_no_bn128) {
    // OPT-IN experiment (segb200_set_option("gemm_bn128", 1)): 128-wide N tiles halve the wave-quantisation granularity of
    // the persistent grid (3x3 256->256 on a 65x129 map: 324 tiles = 3 rounds of 256 vs 5 rounds of 128 = 2.5), but they
    // raise the operand bytes per FLOP by 33 % and the kernel is bandwidth-delay bound on its smem ring: measured -45 %.
    const long long m_tiles = (long long)p.wtiles * p.htiles * p.n_img;
    const long long sms = a->max_ctas > 0 ? a->max_ctas : num_sms();
    const long long r256 = (m_tiles * ((a->cout + 255) / 256) + sms - 1) / sms * 256;
    const long long r128 = (m_tiles * ((a->cout + 127) / 128) + sms - 1) / sms * 128;
    if (r128 * 100 <= r256 * 95) p.bn = 128;
  }
  p.n_tiles = (a->cout + p.bn - 1) / p.bn;
  const long long total = (long long)p.wtiles * p.htiles * p.n_img * p.n_tiles;
  if (total > 0x7fffffffLL) return set_error(-8, "conv_gemm: too many tiles");
  p.total_tiles = (int)total;
  // Two tile streams per CTA (kDual) split the ring in two: with 48 KB k-blocks (64 channels x (128 + 256) rows) a stream would
  // be left with two slots, too shallow to cover the TMA latency.  The streams therefore walk HALF k-blocks -- 32 channels, 64-byte
  // swizzled rows, 24 KB -- over the same packed weights (legal whenever the linear k-block index still addresses them: one tap, or
  // cin a multiple of 64); their extra hand-shakes are what the second issuer hides.
  int grid0 = a->max_ctas > 0 ? a->max_ctas : num_sms();
  if (grid0 > p.total_tiles) grid0 = p.total_tiles;
  const bool use2_early = g_2cta != 0 && !a->y_f32 && (p.bn % 32) == 0 && p.bn >= 64 && (long long)p.wtiles * p.htiles * p.n_img >= 2 &&
                          (g_2cta == 1 || ktot >= 1024);
  const bool want_dual = g_dual != 0 && !use2_early && p.total_tiles >= 2 * grid0;
  if (want_dual && g_dual_subk && bk == 64 && p.bn > 128 && (ntaps == 1 || (a->cin & 63) == 0)) {
    bk = 32; bk_bytes = 64; cblocks = (a->cin + 31) / 32;
  }
  p.cblocks = cblocks; p.ntaps = ntaps; p.bk_bytes = bk_bytes;
  p.kmma_tail = (a->cin - (cblocks - 1) * bk + 15) / 16;
  p.a_stage_bytes = 128 * bk_bytes;
  // CTA-pair kernel (opt-in): needs an N tile that splits into two UMMA-legal halves, 16-bit output, at least one full pair
  const long long m_tiles_all = (long long)p.wtiles * p.htiles * p.n_img;
  const bool use2 = g_2cta != 0 && !a->y_f32 && (p.bn % 32) == 0 && p.bn >= 64 && m_tiles_all >= 2 &&
                    (g_2cta == 1 || ktot >= 1024);
  p.b_stage_bytes = ((((use2 ? p.bn / 2 : p.bn) * bk_bytes) + 1023) & ~1023);
  // ring size: 192 KB by default (1 CTA / SM owns the whole shared memory); segb200_set_option("gemm_ring_kb", n) shrinks
  // it so that a memory-/FMA-bound kernel of another stream can co-reside (measured: not worth it)
  int ring = g_ring_kb > 0 ? g_ring_kb * 1024 : kStageRegion;
  const int min_ring = (a->residual ? kResRegion : 0) + 2 * (p.a_stage_bytes + p.b_stage_bytes);
  if (ring < min_ring) ring = min_ring;
  if (ring > kStageRegion) ring = kStageRegion;
  ring = (ring + 1023) & ~1023;
  p.ring_bytes = ring;
  // B-stationary: a single N tile whose complete weight matrix is small stays in smem; the ring then streams A only,
  // halving the TMA instruction count of the small-K-block layers (3x3 on 32 channels, space-to-depth stems)
  const int nkb = ntaps * cblocks;
  const long long b_total = (long long)nkb * p.b_stage_bytes;
  const int avail = ring - (a->residual ? kResRegion : 0);
  p.b_resident = (!use2 && p.n_tiles == 1 && b_total <= 72 * 1024 && avail - b_total >= 4 * p.a_stage_bytes && !g_no_b_resident) ? 1 : 0;
  // k-blocks per ring stage: every full/empty hand-shake costs the single MMA-issuing thread several hundred cycles whatever the
  // size of the MMAs behind it (profiles/r2_gemm_decomposition.md), so small k-blocks (small Cin: stems, 3x3 convs on 16-64
  // channels) share one: as many as fit in 40 KB, at most 9 (a whole 3x3 kernel), never more than the conv has.  Measured on B200
  // (profiles/r2_call12_kgroup_sweep.log): HRNet-w18-small fp16 16x1024x2048 1 224 -> 1 706 img/s, ResNet101-DLv3+ forward 481 -> 501
  const int sub = p.a_stage_bytes + (p.b_resident ? 0 : p.b_stage_bytes);
  p.group = 1;
  if (!use2 && g_kgroup != 1) {
    int g = (g_kgroup_kb * 1024) / sub;
    const int cap = g_kgroup > 1 ? g_kgroup : 9;
    if (g > cap) g = cap;
    if (g > nkb) g = nkb;
    if (g > 1) p.group = g;
  }
  const int stage_sz = p.group * sub;
  p.num_stages = (int)((avail - (p.b_resident ? b_total : 0)) / stage_sz);
  if (p.num_stages > kMaxStages) p.num_stages = kMaxStages;
  if (p.num_stages < 2) return set_error(-8, "conv_gemm: shared-memory ring too small for this shape");
  p.dbg = g_dbg_counters;
  p.dbg_mode = g_dbg_mode;
  p.mma_pairs = (use2 && g_mma_pairs && p.num_stages >= 5) ? 1 : 0;
  p.out_f32 = a->y_f32 ? 1 : 0;
  p.act = a->act; p.scale = a->scale; p.shift = a->shift; p.residual = a->residual; p.res_ld = a->res_ld;

  // ---- A maps (parity views for stride 2) and the tap table ----
  CUtensorMap tmA[4], tmB, tmC, tmR;
  memset(tmA, 0, sizeof(tmA));
  bool used[4] = {false, false, false, false};
  for (int ky = 0; ky < a->kh; ++ky)
    for (int kx = 0; kx < a->kw; ++kx) {
      const int oh = ky * a->dilation - a->pad_t, ow = kx * a->dilation - a->pad_l;
      int ph = 0, pw = 0, ah = oh, aw = ow;
      if (a->stride == 2) { ah = floordiv2(oh, &ph); aw = floordiv2(ow, &pw); }
      if (ah < -128 || ah > 127 || aw < -128 || aw > 127) return set_error(-9, "conv_gemm: tap offset out of range");
      const int mid = ph * 2 + pw;
      used[mid] = true;
      p.taps[ky * a->kw + kx] = (uint32_t)mid | ((uint32_t)(aw + 128) << 8) | ((uint32_t)(ah + 128) << 16);
    }
  const char* xb = reinterpret_cast<const char*>(a->x);
  const uint32_t boxA[4] = {(uint32_t)bk, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bi};
  for (int mid = 0; mid < 4; ++mid) {
    if (!used[mid]) continue;
    int rc;
    if (flat) {
      const uint64_t dims[4] = {(uint64_t)a->cin, (uint64_t)Wv_out, 1, 1};
      const uint64_t str[3] = {(uint64_t)a->x_ld * 2, (uint64_t)a->x_ld * 2 * (uint64_t)Wv_out,
                               (uint64_t)a->x_ld * 2 * (uint64_t)Wv_out};
      rc = encode_map(&tmA[mid], a->dtype, 4, xb, dims, str, boxA, bk_bytes, "A/flat");
    } else {
      const int s = a->stride, ph = mid >> 1, pw = mid & 1;
      if (ph >= a->h || pw >= a->w) {   // degenerate parity view: point at pixel 0 with zero extent impossible -> 1x1 view never hit
        return set_error(-9, "conv_gemm: input too small for stride-2 parity view");
      }
      const uint64_t dims[4] = {(uint64_t)a->cin, (uint64_t)((a->w - pw + s - 1) / s), (uint64_t)((a->h - ph + s - 1) / s),
                                (uint64_t)a->n};
      const uint64_t str[3] = {(uint64_t)a->x_ld * 2 * s, (uint64_t)a->x_ld * 2 * a->w * s,
                               (uint64_t)a->x_ld * 2 * a->w * a->h};
      rc = encode_map(&tmA[mid], a->dtype, 4, xb + ((long long)ph * a->w + pw) * a->x_ld * 2, dims, str, boxA, bk_bytes,
                      "A/patch");
    }
    if (rc) return rc;
  }
  for (int mid = 0; mid < 4; ++mid)
    if (!used[mid]) for (int j = 0; j < 4; ++j) if (used[j]) { tmA[mid] = tmA[j]; break; }
  {
    const uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)a->cout};
    const uint64_t str[1] = {(uint64_t)ktot * 2};
    const uint32_t box[2] = {(uint32_t)bk, (uint32_t)(use2 ? p.bn / 2 : p.bn)};
    int rc = encode_map(&tmB, a->dtype, 2, a->wgt, dims, str, box, bk_bytes, "B");
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)a->cout, (uint64_t)Wv_out, (uint64_t)Hv_out, (uint64_t)Nv_out};
    const uint64_t esz = a->y_f32 ? 4 : 2;
    const uint64_t str[3] = {(uint64_t)a->y_ld * esz, (uint64_t)a->y_ld * esz * (uint64_t)Wv_out,
                             (uint64_t)a->y_ld * esz * (uint64_t)Wv_out * (uint64_t)Hv_out};
    const uint32_t box[4] = {a->y_f32 ? 32u : 64u, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bi};
    int rc = encode_map(&tmC, a->y_f32 ? (int)DT_F32 : a->dtype, 4, a->y, dims, str, box, 128, "C");
    if (rc) return rc;
    tmR = tmC;
    if (a->residual) {
      const uint64_t rstr[3] = {(uint64_t)a->res_ld * 2, (uint64_t)a->res_ld * 2 * (uint64_t)Wv_out,
                                (uint64_t)a->res_ld * 2 * (uint64_t)Wv_out * (uint64_t)Hv_out};
      rc = encode_map(&tmR, a->dtype, 4, a->residual, dims, rstr, box, 128, "R");
      if (rc) return rc;
    }
  }

  int grid = a->max_ctas > 0 ? a->max_ctas : num_sms();
  if (grid > p.total_tiles) grid = p.total_tiles;
  typedef void (*GemmFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                         const CUtensorMap, const ConvGemmParams);
  // [bf16][two epilogue groups][dual streams]
  static const GemmFn gemm_fns[2][2][2] = {{{conv_gemm_kernel<false, 1, false>, conv_gemm_kernel<false, 1, true>},
                                            {conv_gemm_kernel<false, 2, false>, conv_gemm_kernel<false, 2, true>}},
                                           {{conv_gemm_kernel<true, 1, false>, conv_gemm_kernel<true, 1, true>},
                                            {conv_gemm_kernel<true, 2, false>, conv_gemm_kernel<true, 2, true>}}};
  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    for (int i = 0; i < 8; ++i) cudaFuncSetAttribute(gemm_fns[i >> 2][(i >> 1) & 1][i & 1], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  });
  const int smem_bytes = p.ring_bytes + 2 * kEpiBufBytes + 2432;
  // HBM-bound shapes (short K loop: the epilogue paces the tile) get two epilogue groups, tensor-bound ones a single group
  const bool two_groups = ktot <= g_epi2_maxk && !a->y_f32;
  if (use2) {
    static std::once_flag attr2_once;
    std::call_once(attr2_once, [] {
      cudaFuncSetAttribute(conv_gemm2_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
      cudaFuncSetAttribute(conv_gemm2_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
      cudaFuncSetAttribute(conv_gemm2_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
      cudaFuncSetAttribute(conv_gemm2_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    });
    const long long pair_tiles = ((m_tiles_all + 1) / 2) * p.n_tiles;
    // SM pairs that can host a 2-CTA cluster at this smem size (odd-sized GPCs leave SMs unpaired): a persistent grid larger
    // than that would run its surplus clusters as a second wave
    static int max_pairs[2] = {0, 0};
    if (max_pairs[two_groups ? 1 : 0] == 0) {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3((unsigned)(num_sms() & ~1)); cfg.blockDim = dim3(two_groups ? 320 : 192); cfg.dynamicSmemBytes = kSmemBytes;
      cudaLaunchAttribute at;
      at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
      cfg.attrs = &at; cfg.numAttrs = 1;
      int nclus = 0;
      cudaError_t e = two_groups ? cudaOccupancyMaxActiveClusters(&nclus, conv_gemm2_kernel<true, 2>, &cfg)
                                 : cudaOccupancyMaxActiveClusters(&nclus, conv_gemm2_kernel<true, 1>, &cfg);
      if (e != cudaSuccess || nclus < 1) { cudaGetLastError(); nclus = num_sms() / 2; }
      max_pairs[two_groups ? 1 : 0] = nclus;
    }
    long long g2 = a->max_ctas > 0 ? (a->max_ctas & ~1) : 2LL * max_pairs[two_groups ? 1 : 0];
    if (g2 > 2 * pair_tiles) g2 = 2 * pair_tiles;
    if (g2 < 2) g2 = 2;
    const int grid2 = (int)g2;
    if (a->dtype == DT_BF16) {
      if (two_groups) conv_gemm2_kernel<true, 2><<<grid2, 320, smem_bytes, stream>>>(tmA[0], tmA[1], tmA[2], tmA[3], tmB, tmC, tmR, p);
      else conv_gemm2_kernel<true, 1><<<grid2, 192, smem_bytes, stream>>>(tmA[0], tmA[1], tmA[2], tmA[3], tmB, tmC, tmR, p);
    } else {
      if (two_groups) conv_gemm2_kernel<false, 2><<<grid2, 320, smem_bytes, stream>>>(tmA[0], tmA[1], tmA[2], tmA[3], tmB, tmC, tmR, p);
      else conv_gemm2_kernel<false, 1><<<grid2, 192, smem_bytes, stream>>>(tmA[0], tmA[1], tmA[2], tmA[3], tmB, tmC, tmR, p);
    }
    return check_launch("conv_gemm(2cta)");
  }
  // two tile streams per CTA: each needs >= 2 ring stages of its own and its own tiles; auto mode asks for a K loop of at least
  // g_dual_min_kb hand-shakes per tile (below that the tile is paced by its epilogue, not by the issuing thread)
  const int shakes = (nkb + p.group - 1) / p.group;
  const bool dual = want_dual && p.num_stages >= 4 && p.total_tiles >= 2 * grid && (g_dual == 2 || shakes >= g_dual_min_kb);
  if (dual) p.num_stages &= ~1;
  cudaError_t le = launch_kernel(gemm_fns[a->dtype == DT_BF16 ? 1 : 0][two_groups ? 1 : 0][dual ? 1 : 0], dim3((unsigned)grid),
                                 dim3((unsigned)((two_groups ? 320 : 192) + (dual ? 64 : 0))), (size_t)smem_bytes, stream, pdl_enabled() != 0,
                                 tmA[0], tmA[1], tmA[2], tmA[3], tmB, tmC, tmR, p);
  if (le != cudaSuccess) return set_error((int)le, "conv_gemm: launch failed: %s", cudaGetErrorString(le));
  return check_launch(dual ? "conv_gemm(dual)" : "conv_gemm");
}

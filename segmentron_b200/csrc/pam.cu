// segb200 -- position attention (DANet PAM_Module, modules/module.py:100-131) as a tiled softmax(Q K^T) V kernel on
// tcgen05 tensor cores.  The N x N attention matrix (N = H*W = 32768 at 1024x2048 / OS8: 4.3 GB in fp32 per image in the
// reference) is never materialised.
//
//   energy = Q K^T (no 1/sqrt(d) scaling), attention = softmax over keys, out = attention V, y = gamma*out + x
//
// Exact two-pass softmax (no running rescale of the output accumulator):
//   pass 1 (kPass = 1): per 128-query tile, S = Q K^T tile by tile -> row max m_i and l_i = sum_j exp(s_ij - m_i)
//   pass 2 (kPass = 2): per (128-query tile, 256-column half of d_v): S again, P = exp(S - m_i) / l_i in bf16/fp16
//                       -> shared memory (A operand), O[128 x 256] += P V accumulates in TMEM over all key tiles,
//                       epilogue y = gamma * (O + b_v) + x.
// Warp roles (320 threads): warp 0 = TMA producer (Q tile once; K tile [64 keys x 64] and V^T tile [256 x 64 keys] per
// key tile), warp 1 = tcgen05.mma issuer (S_j is issued before P_{j-1} V_{j-1} so the softmax of tile j overlaps the
// PV MMAs of tile j-1), warps 2..9 = softmax / epilogue: TWO warps per TMEM lane quadrant, each thread owns one query row (= TMEM
// lane) and HALF of the tile's 64 keys (round 2: with one warp per scheduler the exp / pack chain of a tile ran without any
// latency hiding and paced the kernel; the row statistics of pass 1 are combined across the two halves at the end).
// TMEM: O = columns [0,256), S double buffer = columns [256,384).  Smem: Q 16 KB, K 2 x 8 KB, P 2 x 16 KB, V 3 x 32 KB.
// V is consumed as V^T [d_v][N] (K-major along keys); the caller produces it with the same GEMM kernel (roles swapped),
// the value bias b_v is added in the epilogue (sum_j attention_ij = 1).
#include "common.cuh"
#include "../../include/segb200.h"

#include <mutex>

namespace segb200 {

constexpr int kPamQ = 128, kPamK = 64, kPamD = 64, kPamDV = 256;
// Query / key depth = kDB blocks of 64 channels (kDB = 1: DANet's PAM, C/8 = 64; kDB = 4: OCNet's BaseAttentionBlock, key_channels =
// 256, models/ocnet.py:72-113).  Q and K tiles are kDB sub-tiles of [rows x 64] (128-byte swizzled rows), S accumulates over them.
// Shared memory: Q kDB x 16 KB | K 2 x kDB x 8 KB | P 2 x 16 KB | V kVS x 32 KB (3 slots at depth 64, 2 at depth 256: 224 KB) | ctl.
template <int kDB> struct PamSmem {
  static constexpr int kVS = kDB == 1 ? 3 : 2;
  static constexpr int Q = 0, K = Q + kDB * 16384, P = K + 2 * kDB * 8192, V = P + 2 * 16384, Ctl = V + kVS * 32768, Bytes = Ctl + 256;
};
static_assert(PamSmem<1>::K == 16384 && PamSmem<1>::P == 32768 && PamSmem<1>::V == 65536, "depth-64 layout");
static_assert(PamSmem<4>::Bytes <= 232448, "depth-256 layout exceeds the shared memory of an SM");

struct PamCtl {
  uint64_t q_full, o_full;
  uint64_t k_full[2], k_empty[2], s_full[2], s_empty[2], p_full[2], p_empty[2];
  uint64_t v_full[3], v_empty[3];
  uint32_t tmem_base;
};

struct PamParams {
  int n_tok;            // N = H*W
  int ktiles;           // ceil(N / 64)
  float* stat_m;        // [B][N]
  float* stat_l;        // [B][N]
  const float* bias_v;  // [dv_total] or null
  const float* gamma;   // 1 float (device), or null = 1
  const void* x;        // residual [B][N][x_ld], or null = none
  void* y;              // [B][N][y_ld]
  long long x_ld, y_ld;
};

template <bool kBF16, int kPass, int kDB>
__global__ void __launch_bounds__(320, 1)
pam_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
           const __grid_constant__ CUtensorMap tmV, const PamParams p) {
  using H = Half2<kBF16>;
  using T = typename H::T;
  extern __shared__ __align__(1024) uint8_t smem[];
  using SM = PamSmem<kDB>;
  constexpr int kVS = SM::kVS;
  constexpr int kPamSmemQ = SM::Q, kPamSmemK = SM::K, kPamSmemP = SM::P, kPamSmemV = SM::V;
  PamCtl* ctl = reinterpret_cast<PamCtl*>(smem + SM::Ctl);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kPamQ;
  const int half = kPass == 2 ? blockIdx.y : 0;
  const int b = blockIdx.z;
  const int T_ = p.ktiles;

  if (warp == 1) {
    if (lane == 0) {
      mbar_init(&ctl->q_full, 1); mbar_init(&ctl->o_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&ctl->k_full[i], 1); mbar_init(&ctl->k_empty[i], 1);
        mbar_init(&ctl->s_full[i], 1); mbar_init(&ctl->s_empty[i], 256);
        mbar_init(&ctl->p_full[i], 256); mbar_init(&ctl->p_empty[i], 1);
      }
      for (int i = 0; i < kVS; ++i) { mbar_init(&ctl->v_full[i], 1); mbar_init(&ctl->v_empty[i], 1); }
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
  const uint32_t tmem_o = tmem_base, tmem_s = tmem_base + 256;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      mbar_expect_tx(&ctl->q_full, kDB * kPamQ * kPamD * 2);
      for (int db = 0; db < kDB; ++db) tma_load_3d(&tmQ, &ctl->q_full, smem + kPamSmemQ + db * 16384, db * kPamD, q0, b);
      for (int j = 0; j < T_; ++j) {
        const int kb = j & 1; const uint32_t kph = (j >> 1) & 1;
        mbar_wait(&ctl->k_empty[kb], kph ^ 1);
        mbar_expect_tx(&ctl->k_full[kb], kDB * kPamK * kPamD * 2);
        for (int db = 0; db < kDB; ++db)
          tma_load_3d(&tmK, &ctl->k_full[kb], smem + kPamSmemK + (kb * kDB + db) * 8192, db * kPamD, j * kPamK, b);
        if (kPass == 2) {
          const int vb = j % kVS; const uint32_t vph = (j / kVS) & 1;
          mbar_wait(&ctl->v_empty[vb], vph ^ 1);
          mbar_expect_tx(&ctl->v_full[vb], kPamDV * kPamK * 2);
          tma_load_3d(&tmV, &ctl->v_full[vb], smem + kPamSmemV + vb * 32768, j * kPamK, half * kPamDV, b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(kBF16, kPamK);
      const uint32_t idesc_o = make_idesc(kBF16, kPamDV);
      const uint64_t qdesc = make_kmajor_desc(smem_u32(smem + kPamSmemQ), 128);
      // descriptors of every buffer, built once (the issuing thread's instruction count between MMAs is what the pipe waits for)
      const uint64_t pdescs[2] = {make_kmajor_desc(smem_u32(smem + kPamSmemP), 128), make_kmajor_desc(smem_u32(smem + kPamSmemP + 16384), 128)};
      const uint64_t kdescs[2] = {make_kmajor_desc(smem_u32(smem + kPamSmemK), 128), make_kmajor_desc(smem_u32(smem + kPamSmemK + kDB * 8192), 128)};
      const uint64_t vdesc0 = make_kmajor_desc(smem_u32(smem + kPamSmemV), 128);
      int vb_pv = 0; uint32_t vph_pv = 0;              // V ring cursor of the PV side (kVS slots)
      auto do_pv = [&](int t) {
        const int pb = t & 1; const uint32_t pph = (t >> 1) & 1;
        const int vb = vb_pv; const uint32_t vph = vph_pv;
        if (++vb_pv == kVS) { vb_pv = 0; vph_pv ^= 1; }
        mbar_wait(&ctl->p_full[pb], pph);
        mbar_wait(&ctl->v_full[vb], vph);
        tc_fence_after();
        const uint64_t pdesc = pdescs[pb];
        const uint64_t vdesc = vdesc0 + (uint64_t)(vb * (32768 >> 4));
#pragma unroll
        for (int k = 0; k < kPamK / 16; ++k)
          umma_f16(tmem_o, pdesc + (uint64_t)(2 * k), vdesc + (uint64_t)(2 * k), idesc_o, (uint32_t)((t | k) != 0));
        umma_commit(&ctl->p_empty[pb]);
        umma_commit(&ctl->v_empty[vb]);
      };
      mbar_wait(&ctl->q_full, 0);
      for (int j = 0; j < T_; ++j) {
        const int kb = j & 1; const uint32_t kph = (j >> 1) & 1;
        mbar_wait(&ctl->k_full[kb], kph);
        mbar_wait(&ctl->s_empty[kb], kph ^ 1);
        tc_fence_after();
        const uint64_t kdesc = kdescs[kb];
#pragma unroll
        for (int db = 0; db < kDB; ++db)               // depth blocks: sub-tiles 16 KB (Q) / 8 KB (K) apart
#pragma unroll
          for (int k = 0; k < kPamD / 16; ++k)
            umma_f16(tmem_s + (uint32_t)(kb * kPamK), qdesc + (uint64_t)(db * (16384 >> 4) + 2 * k), kdesc + (uint64_t)(db * (8192 >> 4) + 2 * k),
                     idesc_s, (uint32_t)((db | k) != 0));
        umma_commit(&ctl->s_full[kb]);
        umma_commit(&ctl->k_empty[kb]);
        if (kPass == 2 && j >= 1) do_pv(j - 1);
      }
      if (kPass == 2) { do_pv(T_ - 1); umma_commit(&ctl->o_full); }
    }
    __syncwarp();
  } else {
    // ------------------------------ softmax / epilogue warps (2..9) ------------------------------
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    const int hk = (warp - 2) >> 2;                  // which half of the tile's 64 keys (and of the output columns) this warp owns
    const int row = quad * 32 + lane;
    const int qi = q0 + row;
    const bool row_ok = qi < p.n_tok;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    constexpr float kLog2e = 1.4426950408889634f;
    float m = kPass == 2 ? 0.f : -INFINITY, l = 0.f, inv_l = 0.f;
    if (kPass == 2 && row_ok) {
      m = p.stat_m[(long long)b * p.n_tok + qi];
      inv_l = 1.f / p.stat_l[(long long)b * p.n_tok + qi];
    }
    const float m2 = m * kLog2e;
    for (int j = 0; j < T_; ++j) {
      const int sb = j & 1; const uint32_t sph = (j >> 1) & 1;
      mbar_wait(&ctl->s_full[sb], sph);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32(tmem_s + (uint32_t)(sb * kPamK + hk * 32) + lane_off, v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&ctl->s_empty[sb]);
      const int nkeys = p.n_tok - j * kPamK - hk * 32;   // keys >= nkeys of this half-tile are padding (zero-filled by TMA)
      if (kPass == 1) {
        float tmax = -INFINITY;
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c < nkeys) tmax = fmaxf(tmax, __uint_as_float(v[c]));
        const float m_new = fmaxf(m, tmax);
        if (m_new > -INFINITY) {
          float sum = 0.f;
          const float mn2 = m_new * kLog2e;
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c < nkeys) sum += exp2f(fmaf(__uint_as_float(v[c]), kLog2e, -mn2));
          l = l * exp2f((m - m_new) * kLog2e) + sum;
          m = m_new;
        }
      } else {
        mbar_wait(&ctl->p_empty[sb], sph ^ 1);
        uint8_t* pbuf = smem + kPamSmemP + sb * 16384 + row * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {              // 4 x 16 B = this thread's 32 keys of the row, 128B-swizzled (A operand, K-major)
          uint32_t pk[4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int c = g * 8 + jj * 2;
            const float p0 = c < nkeys ? exp2f(fmaf(__uint_as_float(v[c]), kLog2e, -m2)) * inv_l : 0.f;
            const float p1 = c + 1 < nkeys ? exp2f(fmaf(__uint_as_float(v[c + 1]), kLog2e, -m2)) * inv_l : 0.f;
            pk[jj] = H::pack(p0, p1);
          }
          *reinterpret_cast<uint4*>(pbuf + (((hk * 4 + g) ^ (row & 7)) * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        fence_proxy_async();
        mbar_arrive(&ctl->p_full[sb]);
      }
    }
    if (kPass == 1) {
      // combine the two half-rows' (max, sum): stats live in the (idle in pass 1) P region of shared memory
      float2* xch = reinterpret_cast<float2*>(smem + kPamSmemP);
      if (hk == 1) xch[row] = make_float2(m, l);
      named_bar_sync(1, 256);
      if (hk == 0 && row_ok) {
        const float2 o = xch[row];
        const float mm = fmaxf(m, o.x);
        const float ll = (m > -INFINITY ? l * exp2f((m - mm) * kLog2e) : 0.f) + (o.x > -INFINITY ? o.y * exp2f((o.x - mm) * kLog2e) : 0.f);
        p.stat_m[(long long)b * p.n_tok + qi] = mm;
        p.stat_l[(long long)b * p.n_tok + qi] = ll;
      }
    } else {
      mbar_wait(&ctl->o_full, 0);
      tc_fence_after();
      const float gamma = p.gamma ? __ldg(p.gamma) : 1.f;
      const bool has_x = p.x != nullptr;
      const T* xr = reinterpret_cast<const T*>(p.x) + ((long long)b * p.n_tok + qi) * p.x_ld + half * kPamDV;
      T* yr = reinterpret_cast<T*>(p.y) + ((long long)b * p.n_tok + qi) * p.y_ld + half * kPamDV;
      for (int c0 = hk * (kPamDV / 2); c0 < (hk + 1) * (kPamDV / 2); c0 += 32) {     // each warp of the pair: half of the 256 output columns
        uint32_t o[32];
        tmem_ld_32x32(tmem_o + (uint32_t)c0 + lane_off, o);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint4 xv = has_x ? ldg_nc_v4(xr + c0 + g * 8) : make_uint4(0u, 0u, 0u, 0u);
            const uint32_t ux[4] = {xv.x, xv.y, xv.z, xv.w};
            uint32_t pk[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int c = c0 + g * 8 + jj * 2;
              const float b0 = p.bias_v ? __ldg(p.bias_v + half * kPamDV + c) : 0.f;
              const float b1 = p.bias_v ? __ldg(p.bias_v + half * kPamDV + c + 1) : 0.f;
              const float2 xf = H::unpack(ux[jj]);
              pk[jj] = H::pack(fmaf(gamma, __uint_as_float(o[g * 8 + jj * 2]) + b0, xf.x),
                               fmaf(gamma, __uint_as_float(o[g * 8 + jj * 2 + 1]) + b1, xf.y));
            }
            *reinterpret_cast<uint4*>(yr + c0 + g * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace segb200

using namespace segb200;

// q, k: [B][N][dk] (pitches q_ld, k_ld), dk in {64, 256};  vt: [B][dv][vt_ld] (V transposed, keys contiguous);  y = gamma * (softmax(q k^T) v + bias_v) + x
// with gamma == null -> 1 and x == null -> no residual.  Any softmax scale (OCNet: key_channels^-0.5) is folded into q by the caller.
static int attention_impl(const void* q, const void* k, const void* vt, const float* bias_v, const float* gamma, const void* x, void* y,
                          float* stat_m, float* stat_l, int batch, int n_tok, int dk, int dv, int q_ld, int k_ld, int vt_ld, int x_ld,
                          int y_ld, int dtype, cudaStream_t stream, const char* what) {
  if (!q || !k || !vt || !y || !stat_m || !stat_l) return set_error(-1, "%s: null pointer", what);
  if (dtype != DT_BF16 && dtype != DT_F16) return set_error(-2, "%s: dtype must be bf16 or f16", what);
  if (dk != 64 && dk != 256) return set_error(-4, "%s: query/key depth must be 64 or 256 (got %d)", what, dk);
  if (dv % kPamDV != 0 || dv < kPamDV) return set_error(-4, "%s: d_v must be a multiple of 256", what);
  if (q_ld < dk || k_ld < dk || (q_ld & 7) || (k_ld & 7) || (vt_ld & 7) || vt_ld < n_tok || (x && (x_ld & 7)) || (y_ld & 7))
    return set_error(-4, "%s: bad pitches", what);
  if (((uintptr_t)q & 15) || ((uintptr_t)k & 15) || ((uintptr_t)vt & 15) || ((uintptr_t)x & 15) || ((uintptr_t)y & 15))
    return set_error(-7, "%s: pointers must be 16-byte aligned", what);
  if (batch < 1 || n_tok < 1) return set_error(-6, "%s: empty", what);
  CUtensorMap tmQ, tmK, tmV;
  {
    const uint64_t dims[3] = {(uint64_t)dk, (uint64_t)n_tok, (uint64_t)batch};
    const uint64_t sq[2] = {(uint64_t)q_ld * 2, (uint64_t)q_ld * 2 * n_tok};
    const uint64_t sk[2] = {(uint64_t)k_ld * 2, (uint64_t)k_ld * 2 * n_tok};
    const uint32_t bq[3] = {(uint32_t)kPamD, (uint32_t)kPamQ, 1u}, bk[3] = {(uint32_t)kPamD, (uint32_t)kPamK, 1u};
    int rc = encode_map(&tmQ, dtype, 3, q, dims, sq, bq, 128, "attention/Q");
    if (rc) return rc;
    rc = encode_map(&tmK, dtype, 3, k, dims, sk, bk, 128, "attention/K");
    if (rc) return rc;
    const uint64_t dv_[3] = {(uint64_t)n_tok, (uint64_t)dv, (uint64_t)batch};
    const uint64_t sv[2] = {(uint64_t)vt_ld * 2, (uint64_t)vt_ld * 2 * dv};
    const uint32_t bv[3] = {(uint32_t)kPamK, (uint32_t)kPamDV, 1u};
    rc = encode_map(&tmV, dtype, 3, vt, dv_, sv, bv, 128, "attention/Vt");
    if (rc) return rc;
  }
  PamParams p;
  p.n_tok = n_tok; p.ktiles = (n_tok + kPamK - 1) / kPamK;
  p.stat_m = stat_m; p.stat_l = stat_l; p.bias_v = bias_v; p.gamma = gamma; p.x = x; p.y = y; p.x_ld = x_ld; p.y_ld = y_ld;
  typedef void (*Fn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const PamParams);
  // [depth 256][bf16][pass - 1]
  static const Fn fns[2][2][2] = {{{pam_kernel<false, 1, 1>, pam_kernel<false, 2, 1>}, {pam_kernel<true, 1, 1>, pam_kernel<true, 2, 1>}},
                                  {{pam_kernel<false, 1, 4>, pam_kernel<false, 2, 4>}, {pam_kernel<true, 1, 4>, pam_kernel<true, 2, 4>}}};
  static std::once_flag once;
  std::call_once(once, [] {
    for (int i = 0; i < 8; ++i)
      cudaFuncSetAttribute(fns[i >> 2][(i >> 1) & 1][i & 1], cudaFuncAttributeMaxDynamicSharedMemorySize, (i >> 2) ? PamSmem<4>::Bytes : PamSmem<1>::Bytes);
  });
  const int qtiles = (n_tok + kPamQ - 1) / kPamQ;
  const int deep = dk == 256 ? 1 : 0, bf = dtype == DT_BF16 ? 1 : 0;
  const int smem = deep ? PamSmem<4>::Bytes : PamSmem<1>::Bytes;
  dim3 g1((unsigned)qtiles, 1, (unsigned)batch), g2((unsigned)qtiles, (unsigned)(dv / kPamDV), (unsigned)batch);
  fns[deep][bf][0]<<<g1, 320, smem, stream>>>(tmQ, tmK, tmV, p);
  fns[deep][bf][1]<<<g2, 320, smem, stream>>>(tmQ, tmK, tmV, p);
  return check_launch(what);
}

extern "C" int segb200_pam_attention(const void* q, const void* k, const void* vt, const float* bias_v, const float* gamma,
                                     const void* x, void* y, float* stat_m, float* stat_l, int batch, int n_tok, int dv,
                                     int q_ld, int k_ld, int vt_ld, int x_ld, int y_ld, int dtype, void* stream_) {
  if (!gamma || !x) return set_error(-1, "pam_attention: null pointer");
  return attention_impl(q, k, vt, bias_v, gamma, x, y, stat_m, stat_l, batch, n_tok, 64, dv, q_ld, k_ld, vt_ld, x_ld, y_ld, dtype,
                        reinterpret_cast<cudaStream_t>(stream_), "pam_attention");
}

extern "C" int segb200_nonlocal_attention(const void* q, const void* k, const void* vt, const float* bias_v, const float* gamma,
                                          const void* x, void* y, float* stat_m, float* stat_l, int batch, int n_tok, int dk, int dv,
                                          int q_ld, int k_ld, int vt_ld, int x_ld, int y_ld, int dtype, void* stream_) {
  return attention_impl(q, k, vt, bias_v, gamma, x, y, stat_m, stat_l, batch, n_tok, dk, dv, q_ld, k_ld, vt_ld, x_ld, y_ld, dtype,
                        reinterpret_cast<cudaStream_t>(stream_), "nonlocal_attention");
}

// segb200 -- tensor-core rate micro-benchmark (diagnostics only; linked into libsegb200_dbg.so, never into the product library).
//
// One CTA (or CTA pair) per SM issues `iters` x 4 back-to-back tcgen05.mma (K = 16 each, operands resident in shared memory, no TMA,
// no epilogue) and reports SM cycles per MMA.  Answers "what does the tensor pipe sustain for THIS tile shape", which bounds the
// main loop of conv_gemm independently of L2 / TMA / epilogue effects (profiles/r2_gemm_decomposition.md):
//   variant 0: cta_group::1, M = 128, N = n      (A 4 KB + B n*32 B read from smem per MMA)
//   variant 1: cta_group::2, M = 256, N = n      (per SM: A 4 KB + B/2)
// mode (variant 0): 0 = back-to-back issue; 1 = + one tcgen05.commit per 4 MMAs; 2 = + mbarrier try_wait on a completed phase and
// tcgen05.fence before each group of 4; 3 = conv_gemm's full main-loop protocol (4-stage full/empty ring, a second warp standing in
// for the TMA producer, operands read from the ring's rotating stages) -- each step isolates what one element of the protocol costs.
// mode 8: TWO issuing warps, each running the mode-3 protocol on its own 2-stage ring and its own 256-column accumulator (two tiles
// in flight per CTA): does the tensor pipe interleave two instruction streams, hiding each issuer's hand-shake behind the other's MMAs?
#include "common.cuh"

namespace segb200 {

__device__ __forceinline__ uint32_t probe_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

template <int kPair>
__global__ void __launch_bounds__(256, 1) mma_probe_kernel(int n, int iters, int mode, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done;
  __shared__ uint64_t sfull[4], sempty[4], sdummy;
  __shared__ uint64_t dfull[2][2], dempty[2][2], ddone[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? probe_ctarank() : 0;
  for (int i = threadIdx.x; i < (mode == 8 ? 4 : 1) * (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // finite values
  if (threadIdx.x == 0) {
    mbar_init(&done, 1); mbar_init(&sdummy, 1);
    for (int i = 0; i < 4; ++i) { mbar_init(&sfull[i], 1); mbar_init(&sempty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&dfull[i >> 1][i & 1], 1); mbar_init(&dempty[i >> 1][i & 1], 1); }
    mbar_init(&ddone[0], 1); mbar_init(&ddone[1], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc(&tmem_base_s, 512);
      tmem_relinquish();
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (kPair) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  long long cycles = 0;
  if (!kPair && mode == 8) {
    // warps 0 / 3: issuers of streams 0 / 1;  warps 2 / 4: their stand-in producers
    if ((warp == 0 || warp == 3) && lane == 0) {
      const int sidx = warp == 0 ? 0 : 1;
      const uint32_t fmt = 1u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (((uint32_t)n >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t acc = tmem + (uint32_t)(sidx * 256);
      const long long t0 = clock64();
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&dfull[sidx][stage], phase);
        tc_fence_after();
        const uint32_t soff = (uint32_t)(sidx * 2 + stage) * 49152u;
        const uint64_t ad = make_kmajor_desc(smem_u32(smem) + soff, 128), bd = make_kmajor_desc(smem_u32(smem) + soff + 16384, 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(acc, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (uint32_t)((it & 7) | k));
        umma_commit(&dempty[sidx][stage]);
        if (++stage == 2) { stage = 0; phase ^= 1; }
      }
      umma_commit(&ddone[sidx]);
      mbar_wait(&ddone[sidx], 0);
      cycles = clock64() - t0;
      atomicMax(out, (unsigned long long)cycles);      // slowest issuer of the slowest CTA
      if (sidx == 0) atomicAdd(out + 1, 1ull);
    } else if ((warp == 2 || warp == 4) && lane == 0) {
      const int sidx = warp == 2 ? 0 : 1;
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&dempty[sidx][stage], phase ^ 1);
        mbar_arrive(&dfull[sidx][stage]);
        if (++stage == 2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 0 && lane == 0 && rank == 0) {
    const uint64_t adesc = make_kmajor_desc(smem_u32(smem), 128);
    const uint64_t bdesc = make_kmajor_desc(smem_u32(smem + 16384), 128);
    const uint32_t fmt = 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (((uint32_t)n >> 3) << 17) | (((kPair ? 256u : 128u) >> 4) << 24);
    const long long t0 = clock64();
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      uint32_t soff = 0;
      if (!kPair && mode == 3) {                       // the conv_gemm main-loop protocol: wait full[stage], MMAs on THAT stage, commit empty[stage]
        mbar_wait(&sfull[stage], phase);
        tc_fence_after();
        soff = (uint32_t)stage * 49152u;               // 4 stages of {A 16 KB, B 32 KB}
      } else if (!kPair && mode == 2) {
        mbar_wait(&sdummy, 1);                         // a barrier whose awaited phase already completed: pure try_wait + fence cost
        tc_fence_after();
      } else if (!kPair && mode == 4) {
        mbar_wait(&sdummy, 1);                         // wait only, no tcgen05.fence
      } else if (!kPair && mode == 5) {
        tc_fence_after();                              // fence only
      } else if (!kPair && mode == 7) {
        mbar_wait_lean(smem_u32(&sdummy), 1);          // lean wait (no watchdog loop) + fence
        tc_fence_after();
      } else if (!kPair && mode == 6) {                // ring protocol consumed in PAIRS of stages: 2 waits, one fence, 8 MMAs
        mbar_wait(&sfull[stage], phase);
        mbar_wait(&sfull[stage + 1], phase);
        tc_fence_after();
        const uint64_t a0 = make_kmajor_desc(smem_u32(smem) + (uint32_t)stage * 49152u, 128), b0 = a0 + (16384 >> 4);
        const uint64_t a1 = a0 + (49152 >> 4), b1 = b0 + (49152 >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem + (uint32_t)(((it >> 3) & 1) * 256), a0 + (uint64_t)(2 * k), b0 + (uint64_t)(2 * k), idesc, (uint32_t)((it & 7) | k));
        umma_commit(&sempty[stage]);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem + (uint32_t)(((it >> 3) & 1) * 256), a1 + (uint64_t)(2 * k), b1 + (uint64_t)(2 * k), idesc, 1u);
        umma_commit(&sempty[stage + 1]);
        stage += 2; if (stage == 4) { stage = 0; phase ^= 1; }
        ++it;                                          // this iteration consumed two k-blocks
        continue;
      }
      const uint64_t ad = make_kmajor_desc(smem_u32(smem) + soff, 128), bd = make_kmajor_desc(smem_u32(smem) + soff + 16384, 128);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (kPair)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem + (uint32_t)((it & 1) * 256)), "l"(adesc + (uint64_t)(2 * k)), "l"(bdesc + (uint64_t)(2 * k)), "r"(idesc), "r"(1u) : "memory");
        else
          umma_f16(tmem + (uint32_t)(((it >> 3) & 1) * 256), ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (uint32_t)((it & 7) | k));
      }
      if (!kPair && mode == 1) umma_commit(&sdummy);   // commit nobody waits for
      if (!kPair && (mode == 2 || mode == 4 || mode == 5 || mode == 7)) umma_commit(&sempty[it & 3]);
      if (!kPair && mode == 3) { umma_commit(&sempty[stage]); if (++stage == 4) { stage = 0; phase ^= 1; } }
    }
    if (kPair)
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(smem_u32(&done)), "h"((uint16_t)3) : "memory");
    else
      umma_commit(&done);
    mbar_wait(&done, 0);
    cycles = clock64() - t0;
    atomicAdd(out, (unsigned long long)cycles);
    atomicAdd(out + 1, 1ull);
  } else if (kPair && warp == 0 && lane == 0) {
    mbar_wait(&done, 0);          // the peer may not exit (its smem / TMEM are in use) before the pair's MMAs retire
  } else if (!kPair && (mode == 3 || mode == 6) && warp == 2 && lane == 0) {
    int stage = 0; uint32_t phase = 0;               // stand-in for the TMA producer: a slot becomes "full" as soon as it is empty
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&sempty[stage], phase ^ 1);
      mbar_arrive(&sfull[stage]);
      if (++stage == 4) { stage = 0; phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) {
    tc_fence_after();
    if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else tmem_dealloc(tmem, 512);
  }
}

}  // namespace segb200

using namespace segb200;

// out: 2 device u64 {sum of cycles over the issuing CTAs, number of issuing CTAs}.  Returns 0 / cudaError.
extern "C" int segb200_debug_mma_probe(int variant, int n, int iters, int mode, void* out2_u64, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!out2_u64 || n < 16 || n > 256 || (n & 15) || iters < 1) return set_error(-4, "mma_probe: bad arguments");
  const int smem = 200 * 1024;          // whole-SM shared memory: exactly one CTA per SM (each allocates all 512 TMEM columns)
  unsigned long long* out = reinterpret_cast<unsigned long long*>(out2_u64);
  if (variant == 0) {
    cudaFuncSetAttribute(mma_probe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    mma_probe_kernel<0><<<num_sms(), mode == 8 ? 160 : 128, smem, stream>>>(n, iters, mode, out);
  } else {
    cudaFuncSetAttribute(mma_probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(num_sms() & ~1)); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, mma_probe_kernel<1>, n, iters, mode, out);
    if (e != cudaSuccess) return set_error((int)e, "mma_probe: %s", cudaGetErrorString(e));
  }
  return check_launch("mma_probe");
}

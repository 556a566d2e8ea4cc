// segb200 -- SyncBatchNorm statistics exchange fused into the finalize kernels, over NVLink peer memory (sm_100a).
//
// Reference behaviour: `nn.SyncBatchNorm.convert_sync_batchnorm(model)` (tools/train.py:73-79): batch statistics over ALL ranks in
// forward, and the two backward sums (sum g, sum g*xhat) over all ranks in backward -- 2 small collectives per BatchNorm layer, 237
// per training step of DeepLabv3+/ResNet101, each a latency-bound NCCL launch serialised on the compute stream in round 1.
//
// Here the exchange is part of the kernel that needs its result.  Every rank owns a buffer in symmetric memory (allocated with
// torch.distributed._symmetric_memory; all peers' buffers are mapped into every process, device array `peers[world]`):
//     data : [slot][source rank][2][cmax] fp32         flags : [slot][source rank][32] u32
// A finalize kernel (one thread per channel, <= 16 CTAs)
//   1. reduces its rank's per-slab partial sums (fixed order, double),
//   2. STORES its [2][C] sums into EVERY peer's data[slot][my rank] -- plain st.global on peer-mapped pointers, i.e. NVLink P2P
//      writes -- then __threadfence_system() and a st.release.sys of the step's epoch into the peer's flags[slot][my rank][cta],
//   3. spins (ld.acquire.sys, watchdog) until flags[slot][q][cta] == epoch for every q in its OWN buffer,
//   4. sums the `world` contributions in rank order (bit-identical on all ranks) and finishes the BatchNorm arithmetic.
// No NCCL launch, no separate reduce kernel, ~one NVLink round trip of latency.  A slot is reused once per training step; the
// gradient all-reduce between two uses orders the ranks (no rank can be a whole step ahead), the epoch makes stale flags harmless.
#include "common.cuh"
#include "../../include/segb200.h"

namespace segb200 {

constexpr int kSyncThreads = 128;     // channels per CTA
constexpr int kSlabLanes = 4;         // threads per channel for the local slab reduction (fixed-order combine)
constexpr int kFlagsPerSrc = 32;      // CTAs per layer <= 32 (C <= 4096)

struct SyncGeo {
  float* const* peers;        // device array [world]: base of every rank's symmetric buffer (as mapped in this process)
  int world, rank, cmax;
  long long data_off;         // float offset of data[slot] from the buffer base
  long long flag_off;         // u32 offset of flags[slot] from the buffer base
  const unsigned* epoch;      // device word: the current step's epoch (> 0), bumped once per step by the host
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// publish (a, b) of channel `ch` to all ranks, wait for all ranks, return the rank-ordered sums
__device__ __forceinline__ void exchange2(const SyncGeo& g, int ch, bool valid, float a, float b, double& sa, double& sb) {
  const unsigned epoch = *g.epoch;
  const int cta = blockIdx.x;
  if (valid) {
    for (int p = 0; p < g.world; ++p) {
      float* dst = g.peers[p] + g.data_off + (long long)g.rank * 2 * g.cmax;
      dst[ch] = a;
      dst[g.cmax + ch] = b;
    }
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < g.world) {
    const int p = threadIdx.x;
    unsigned* f = reinterpret_cast<unsigned*>(g.peers[p]) + g.flag_off + (long long)g.rank * kFlagsPerSrc + cta;
    st_release_sys(f, epoch);
    const unsigned* mine = reinterpret_cast<const unsigned*>(g.peers[g.rank]) + g.flag_off + (long long)p * kFlagsPerSrc + cta;
    unsigned spins = 0;
    while (ld_acquire_sys(mine) != epoch) {
      if (spins > 4096u) __nanosleep(256);
      if (++spins > (1u << 26)) {       // ~20 s: a missing peer traps (launch failure) instead of hanging the box
        printf("segb200: SyncBN exchange timeout (rank %d waiting for rank %d, cta %d)\n", g.rank, p, cta);
        __trap();
      }
    }
  }
  __syncthreads();
  sa = 0.0; sb = 0.0;
  if (valid) {
    const float* src = g.peers[g.rank] + g.data_off;
    for (int q = 0; q < g.world; ++q) {
      const volatile float* s = src + (long long)q * 2 * g.cmax;
      sa += (double)s[ch];
      sb += (double)s[g.cmax + ch];
    }
  }
}

// sum the [slabs][2][c] partials of channel ch: kSlabLanes threads per channel stride over the slabs, combined in lane order
__device__ __forceinline__ void local_sums(const float* __restrict__ partial, int slabs, int c, int ch, bool valid, double& s, double& q) {
  __shared__ double red[kSlabLanes][kSyncThreads][2];
  const int t = threadIdx.x & (kSyncThreads - 1), lane = threadIdx.x / kSyncThreads;
  double a = 0.0, b = 0.0;
  if (valid)
    for (int sl = lane; sl < slabs; sl += kSlabLanes) {
      a += (double)partial[((long long)sl * 2) * c + ch];
      b += (double)partial[((long long)sl * 2 + 1) * c + ch];
    }
  red[lane][t][0] = a; red[lane][t][1] = b;
  __syncthreads();
  s = 0.0; q = 0.0;
#pragma unroll
  for (int l = 0; l < kSlabLanes; ++l) { s += red[l][t][0]; q += red[l][t][1]; }
}

// forward: local sum / sum of squares over the slabs -> exchange -> mean, invstd, scale, shift, running statistics
__global__ void __launch_bounds__(kSyncThreads * kSlabLanes)
bn_finalize_sync_kernel(const float* __restrict__ partial, int slabs, int c, double count_total, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float* __restrict__ running_mean, float* __restrict__ running_var,
                        float momentum, float eps, float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ scale,
                        float* __restrict__ shift, const SyncGeo g) {
  const int ch = blockIdx.x * kSyncThreads + (threadIdx.x & (kSyncThreads - 1));
  const bool lead = threadIdx.x < kSyncThreads;                 // lane 0 of each channel owns the exchange and the result
  const bool valid = ch < c;
  double s, q;
  local_sums(partial, slabs, c, ch, valid, s, q);
  double ts, tq;
  exchange2(g, ch, valid && lead, (float)s, (float)q, ts, tq);
  if (!valid || !lead) return;
  const double m = ts / count_total;
  double var = tq / count_total - m * m;
  if (var < 0.0) var = 0.0;
  const float is = (float)(1.0 / sqrt(var + (double)eps));
  mean[ch] = (float)m;
  invstd[ch] = is;
  const float ga = gamma != nullptr ? gamma[ch] : 1.f;
  const float sc = ga * is;
  scale[ch] = sc;
  shift[ch] = (beta != nullptr ? beta[ch] : 0.f) - (float)m * sc;
  if (running_mean != nullptr) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
  if (running_var != nullptr) {
    const double unb = count_total > 1.0 ? var * count_total / (count_total - 1.0) : var;
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unb;
  }
}

// backward: local (sum g, sum g*xhat) -> dgamma / dbeta (+)= LOCAL sums (DDP averages them with the other gradients) -> exchange ->
// sums[2][c] over all ranks for the data-gradient pass
__global__ void __launch_bounds__(kSyncThreads * kSlabLanes)
bn_bwd_finalize_sync_kernel(const float* __restrict__ partial, int slabs, int c, const float* __restrict__ mean,
                            const float* __restrict__ invstd, float* __restrict__ sums, float* __restrict__ dgamma,
                            float* __restrict__ dbeta, const SyncGeo g) {
  const int ch = blockIdx.x * kSyncThreads + (threadIdx.x & (kSyncThreads - 1));
  const bool lead = threadIdx.x < kSyncThreads;
  const bool valid = ch < c;
  double a, b;
  local_sums(partial, slabs, c, ch, valid, a, b);
  if (valid && lead) {
    b = (invstd != nullptr ? (double)invstd[ch] : 1.0) * (b - (mean != nullptr ? (double)mean[ch] : 0.0) * a);
    if (dgamma != nullptr) dgamma[ch] += (float)b;
    if (dbeta != nullptr) dbeta[ch] += (float)a;
  }
  double ta, tb;
  exchange2(g, ch, valid && lead, (float)a, (float)b, ta, tb);
  if (!valid || !lead) return;
  sums[ch] = (float)ta;
  sums[c + ch] = (float)tb;
}

__global__ void counter_add_kernel(unsigned* p, unsigned v) { *p += v; }

static int check_sync(const char* what, const void* peers, int world, int rank, int c, int cmax, const void* epoch) {
  if (!peers || !epoch) return set_error(-1, "%s: null exchange pointers", what);
  if (world < 1 || world > 64 || rank < 0 || rank >= world) return set_error(-4, "%s: bad world / rank (%d / %d)", what, world, rank);
  if (c < 1 || c > cmax || (c + kSyncThreads - 1) / kSyncThreads > kFlagsPerSrc)
    return set_error(-4, "%s: c = %d exceeds the exchange slot (cmax %d, at most %d channels)", what, c, cmax, kSyncThreads * kFlagsPerSrc);
  return 0;
}

}  // namespace segb200

using namespace segb200;

extern "C" int segb200_syncbn_slot_floats(int world, int cmax) { return world * 2 * cmax; }
extern "C" int segb200_syncbn_slot_flags(int world) { return world * kFlagsPerSrc; }

extern "C" int segb200_counter_add(void* counter_u32, int value, void* stream) {
  if (!counter_u32) return set_error(-1, "counter_add: null pointer");
  counter_add_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<unsigned*>(counter_u32), (unsigned)value);
  return check_launch("counter_add");
}

extern "C" int segb200_bn_finalize_sync(const float* partial, int slabs, int c, double count_total, const float* gamma,
                                        const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                                        float* mean, float* invstd, float* scale, float* shift, const void* peers_dev, int world,
                                        int rank, int cmax, long long data_off, long long flag_off, const void* epoch_dev,
                                        void* stream) {
  if (!partial || !mean || !invstd || !scale || !shift) return set_error(-1, "bn_finalize_sync: null pointer");
  if (slabs < 1 || count_total < 1.0) return set_error(-4, "bn_finalize_sync: bad sizes");
  int rc = check_sync("bn_finalize_sync", peers_dev, world, rank, c, cmax, epoch_dev);
  if (rc) return rc;
  SyncGeo g{reinterpret_cast<float* const*>(peers_dev), world, rank, cmax, data_off, flag_off, reinterpret_cast<const unsigned*>(epoch_dev)};
  bn_finalize_sync_kernel<<<(c + kSyncThreads - 1) / kSyncThreads, kSyncThreads * kSlabLanes, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      partial, slabs, c, count_total, gamma, beta, running_mean, running_var, momentum, eps, mean, invstd, scale, shift, g);
  return check_launch("bn_finalize_sync");
}

extern "C" int segb200_bn_bwd_finalize_sync(const float* partial, int slabs, int c, const float* mean, const float* invstd,
                                            float* sums, float* dgamma, float* dbeta, const void* peers_dev, int world, int rank,
                                            int cmax, long long data_off, long long flag_off, const void* epoch_dev, void* stream) {
  if (!partial || !sums) return set_error(-1, "bn_bwd_finalize_sync: null pointer");
  if (slabs < 1) return set_error(-4, "bn_bwd_finalize_sync: bad sizes");
  int rc = check_sync("bn_bwd_finalize_sync", peers_dev, world, rank, c, cmax, epoch_dev);
  if (rc) return rc;
  SyncGeo g{reinterpret_cast<float* const*>(peers_dev), world, rank, cmax, data_off, flag_off, reinterpret_cast<const unsigned*>(epoch_dev)};
  bn_bwd_finalize_sync_kernel<<<(c + kSyncThreads - 1) / kSyncThreads, kSyncThreads * kSlabLanes, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      partial, slabs, c, mean, invstd, sums, dgamma, dbeta, g);
  return check_launch("bn_bwd_finalize_sync");
}

// One-shot all-reduce (mean) of the flat head-gradient bucket over NVLink peer memory.
//
// Data parallelism of the fusion head (SURVEY.md §8 e) needs ONE collective per step: the sum of a 379 KB fp32 bucket
// over the ranks.  At that size a ring / tree all-reduce is pure latency (≈ 20-40 µs through NCCL at 8 GPUs, against a
// 210 µs step); with NVSwitch every GPU can read every peer at full bandwidth, so each rank simply reads the W copies
// and sums them itself: 8 x 379 KB over NVLink is a few microseconds.
//
// Every rank owns a symmetric STAGING buffer (two halves, used alternately) and a flag pad, both mapped into every
// peer (torch symmetric memory does the IPC plumbing, training.PeerAllReduce).  CTA c of rank r:
//   1. copies its slice of the local bucket into its own staging half,
//   2. publishes flag (r, c) = step on every peer's pad (release, system scope),
//   3. waits until flag (p, c) >= step arrived from every peer p (acquire, system scope),
//   4. sums slice c of all W staging buffers in rank order (so every rank gets the bit-identical result) and writes
//      mean back into its local bucket.
// Only CTA c of the peers has to be done with its copy, so there is no grid-wide barrier.  The staging half written at
// step k is not written again before step k + 2, and a rank can only reach step k + 2 after every peer has published
// its step k + 1 flags, i.e. finished reading step k: no trailing barrier either.  All CTAs must be co-resident
// (grid <= number of SMs).
// A peer that never publishes (crashed rank, torn-down mapping) must not hang the GPU: the wait is bounded
// (kSpinTimeoutNs of %globaltimer, ~2 s); on expiry the CTA records the missing peer in the status word that follows
// the flags in its own pad (mmrca_peer_allreduce_status reads it) and leaves its slice of the bucket untouched.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mmrca {
namespace peer {

constexpr int kMaxWorld = 8;
constexpr int kCtas = 48, kThreads = 256;
constexpr long long kSpinTimeoutNs = 2000000000LL;
constexpr int kStatusWords = 4;      // after the flags: [0] = 1 + rank of a peer that timed out (0: healthy), rest reserved

struct Args {
  float* flat;                       // local bucket, n floats (n % 4 == 0, 16-byte aligned): in/out
  const float* staging[kMaxWorld];   // staging buffer of every rank (mapped here), 2 halves of n_pad floats each
  uint32_t* pads[kMaxWorld];         // flag pad of every rank (mapped here): [2 parities][world][kCtas] uint32
  int n, n_pad, rank, world;
  uint32_t step;                     // 1, 2, 3, ... (monotone; the same on every rank)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {      // not through the (non-coherent) L1
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ long long globaltimer_ns_peer() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__global__ void __launch_bounds__(kThreads) allreduce_mean_kernel(const Args a) {
  __shared__ int s_dead;
  const int tid = threadIdx.x, cta = blockIdx.x;
  if (tid == 0) s_dead = 0;
  const int par = int(a.step & 1u);
  const int n4 = a.n / 4, per = (n4 + kCtas - 1) / kCtas, lo = cta * per, hi = min(n4, lo + per);
  float4* mine = reinterpret_cast<float4*>(const_cast<float*>(a.staging[a.rank])) + size_t(par) * (a.n_pad / 4);
  const float4* src = reinterpret_cast<const float4*>(a.flat);
  for (int i = lo + tid; i < hi; i += kThreads) mine[i] = src[i];
  __syncthreads();
  if (tid < a.world) {
    __threadfence_system();
    st_release_sys(a.pads[tid] + (size_t(par) * a.world + a.rank) * kCtas + cta, a.step);      // tell rank `tid`
    const uint32_t* f = a.pads[a.rank] + (size_t(par) * a.world + tid) * kCtas + cta;            // hear from rank `tid`
    const long long t0 = globaltimer_ns_peer();
    uint32_t spins = 0;
    while (int32_t(ld_acquire_sys(f) - a.step) < 0) {
      if ((++spins & 1023u) == 0 && globaltimer_ns_peer() - t0 > kSpinTimeoutNs) {
        a.pads[a.rank][size_t(2) * a.world * kCtas] = uint32_t(1 + tid);      // status word: peer `tid` never arrived
        s_dead = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (s_dead) return;
  // all W remote loads of an element group are in flight before the first one is consumed (a load-add-load-add loop
  // would pay one NVLink round trip per rank), two groups per thread interleaved
  const float inv = 1.0f / float(a.world);
  float4* dst = reinterpret_cast<float4*>(a.flat);
  const size_t half = size_t(par) * (a.n_pad / 4);
  for (int i0 = lo + tid; i0 < hi; i0 += 2 * kThreads) {
    const int i1 = i0 + kThreads;
    float4 v0[kMaxWorld], v1[kMaxWorld];
#pragma unroll
    for (int p = 0; p < kMaxWorld; ++p) {
      if (p < a.world) {
        v0[p] = ld_peer(reinterpret_cast<const float4*>(a.staging[p]) + half + i0);
        if (i1 < hi) v1[p] = ld_peer(reinterpret_cast<const float4*>(a.staging[p]) + half + i1);
      }
    }
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
#pragma unroll
    for (int p = 0; p < kMaxWorld; ++p) {
      if (p < a.world) {
        s0.x += v0[p].x; s0.y += v0[p].y; s0.z += v0[p].z; s0.w += v0[p].w;
        if (i1 < hi) { s1.x += v1[p].x; s1.y += v1[p].y; s1.z += v1[p].z; s1.w += v1[p].w; }
      }
    }
    dst[i0] = make_float4(s0.x * inv, s0.y * inv, s0.z * inv, s0.w * inv);
    if (i1 < hi) dst[i1] = make_float4(s1.x * inv, s1.y * inv, s1.z * inv, s1.w * inv);
  }
}

}  // namespace peer
}  // namespace mmrca

// Backward of the token-level attention blocks (mmrca_token.cuh): what loss.backward() does to SelfAttention.forward
// (CVPR_code/multimodal_model.py:51-68) / ReverseCrossAttention.forward (:82-108) on [B, L, K] token sequences.
//
//   tok_attn_bwd     per (sample, 128-query tile), from the forward's operand images (Q, K, V), its unnormalised attention
//                    weights P (bf16 image) and row sums: C = P V again (tensor core), LayerNorm + ReLU backward per row,
//                    d(weights) = E V^T, softmax backward in place (P -> dS), dV = P^T E, dQ = dS K, dK = dS^T Q - five
//                    tcgen05 chains, accumulators in TMEM.  dQ rows are the tile's own; dK / dV rows collect both query tiles.
//                    All three leave the kernel as bf16 ROWS of the flat gradient matrices G [B * L][columns] (dQ times
//                    1/sqrt(d_kq); dK | dV side by side), together with the bias gradients (column sums, from the fp32
//                    accumulators).  Two-tile samples: both query tiles add their dK | dV rows with 16-byte vector atomics.
//   tok_wgrad        dW^T[K tile of 128][gradient columns] += X^T G over the flat token dimension: BOTH operands are read as
//                    they lie in HBM - row-major [tokens][columns] is MN-major for this product - through tensor maps
//                    ({64 columns, 64 tokens} boxes, 128-byte swizzle), three-stage TMA pipeline, split-K over the tokens.
//   tok_dgrad        dX[B * L][K] = G [W_query; W_key; W_value]: A = G (K-major, TMA), B = the bf16 weights [columns][K] as they
//                    lie (MN-major, TMA), fp32 rows out.
#pragma once
#include "mmrca_token.cuh"

namespace mmrca {
namespace tok {


struct AttnBwdArgs {
  const void* q_img; const void* k_img; const void* v_img;     // the forward's operand images
  const void* p_img;        // [B * tiles][op_bytes(256)] unnormalised attention weights (forward, training)
  const float* sum;         // [B * tiles][128] softmax row sums
  const float* ln_g; const float* ln_b;
  const float* d_out;       // [B][L][DV]
  // bf16 gradient rows, token r = b * L + t: dQ (times 1/sqrt(d_kq)) -> g_q[r * ld_q + ..]; dK | dV -> g_kv[r * ld_kv + ..]
  __nv_bfloat16* g_q; int ld_q;
  __nv_bfloat16* g_kv; int ld_kv;
  int kv_atomic;            // two-tile samples: both query tiles ADD their dK | dV rows into the zeroed matrix with 16-byte vector
                            // atomics (red.global.add.noftz.v4.bf16x2: eight bf16 per operation; two addends per element, so the
                            // result does not depend on their order)
  float* g_bq; float* g_bk; float* g_bv;     // += column sums (atomics)
  float qscale;
  float* g_ln_g; float* g_ln_b;     // += (atomics)
  int L, tiles_per_sample, reverse;
};

template <int DKQ, int DV>
struct AttnBwdSmem {
  static constexpr uint32_t QB = htc::op_bytes(DKQ), VB = htc::op_bytes(DV), PB = htc::op_bytes(kMaxTiles * kTile);
  static constexpr uint32_t K = 0;                                           // kMaxTiles key tiles
  static constexpr uint32_t VQ = htc::al128(K + kMaxTiles * QB);             // value tiles; later the query tile
  static constexpr uint32_t VQ_BYTES = kMaxTiles * VB > QB ? kMaxTiles * VB : QB;
  static constexpr uint32_t P = htc::al128(VQ + VQ_BYTES);                   // P, overwritten by dS
  static constexpr uint32_t E = htc::al128(P + PB);                          // [128 x DV]: d(P_un V) rows, bf16
  static constexpr uint32_t LN = htc::al128(E + VB);                         // gamma, beta
  static constexpr uint32_t VSUM = LN + 2 * DV * 4;
  static constexpr uint32_t NS = 3 * DV + 2 * DKQ + DV;                      // column sums of a 32-row slab: d(gamma), d(beta),
  static constexpr uint32_t ACC = VSUM + DV * 4;                             // d(colsum V), d(b_query), d(b_key), d(b_value); x 4 slabs
  static constexpr uint32_t PART = ACC + 4 * NS * 4;
  static constexpr uint32_t BAR = htc::al128(PART + 2 * kTile * 8);
  static constexpr uint32_t BYTES = BAR + 64;
  static_assert(BYTES <= 232448, "token attention backward does not fit shared memory");
};

__device__ __forceinline__ void unpack_bf16x8(const uint4 u, float* v) {
  v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
  v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
  v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
  v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ void ld_chunks16(const uint8_t* op, int row, int col0, float (&v)[16]) {
  unpack_bf16x8(*reinterpret_cast<const uint4*>(op + uint32_t(col0 >> 3) * kCS + row_off(row)), v);
  unpack_bf16x8(*reinterpret_cast<const uint4*>(op + uint32_t((col0 >> 3) + 1) * kCS + row_off(row)), v + 8);
}
// dst[0 .. 8) (bf16) (+)= v: plain 16-byte store, or one vector atomic
__device__ __forceinline__ void put_bf16x8(__nv_bfloat16* dst, const float (&v)[8], bool atomic) {
  const uint4 u = pack_bf16x8(v);
  if (atomic)
    asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
  else
    *reinterpret_cast<uint4*>(dst) = u;
}
__device__ __forceinline__ float warp_sum32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Column sums over the 32 rows a warp holds (lane = row, v[0 .. N) = the row's values of N columns), WRITTEN to dst[0 .. N)
// (a per-warp slot: shared-memory float atomics are compare-and-swap loops, four warps on the same words crawl):
// instead of N butterflies of 5 shuffles, every step with an even count EXCHANGES halves - the lane with the step's bit set
// keeps the upper half of the columns, its partner the lower - so the count halves with the lane distance (48 -> 24 -> 12 ->
// 6 -> 3: 45 shuffles, then 3 for the last bit instead of 240).  Destroys v.
template <int N, int M, int F>
struct WarpColSum {
  static __device__ __forceinline__ void run(float* v, int lane, float* dst) {
    if constexpr (M == 0) {
      if ((lane & F) == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) dst[i] = v[i];
      }
    } else if constexpr (N % 2 == 0) {
      const bool up = (lane & M) != 0;
#pragma unroll
      for (int i = 0; i < N / 2; ++i) {
        const float send = up ? v[i] : v[i + N / 2], keep = up ? v[i + N / 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, M);
      }
      WarpColSum<N / 2, M / 2, F>::run(v, lane, dst + (up ? N / 2 : 0));
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], M);
      WarpColSum<N, M / 2, F | M>::run(v, lane, dst);
    }
  }
};
template <int N>
__device__ __forceinline__ void warp_col_sums(float (&v)[N], int lane, float* dst) { WarpColSum<N, 16, 0>::run(v, lane, dst); }

// 256 threads; thread (q = warp % 4, half = warp / 4, lane) owns row 32 q + lane of whatever the phase's rows are (queries
// for the row phases, keys when the dK / dV accumulators are drained) and the column half `half`.
template <int DKQ, int DV>
__global__ void __launch_bounds__(256, 1) tok_attn_bwd_kernel(const AttnBwdArgs a) {
  extern __shared__ __align__(128) uint8_t sm[];
  using S = AttnBwdSmem<DKQ, DV>;
  constexpr int HC = DV / 2, HQ = DKQ / 2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::BAR);      // [0] V, P loads; then Q  [1] MMAs  [2] K load  [3] the dQ chain
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* ln_s = reinterpret_cast<float*>(sm + S::LN);
  float* vsum = reinterpret_cast<float*>(sm + S::VSUM);
  float* slots = reinterpret_cast<float*>(sm + S::ACC);
  float2* part = reinterpret_cast<float2*>(sm + S::PART);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2, row = 32 * q + lane;
  float* myslot = slots + q * S::NS;
  constexpr int SLOT_B = 3 * DV;                   // bias sums start here
  const int tps = a.tiles_per_sample, L = a.L, ncols = tps * kTile;
  const int b = blockIdx.x / tps, mt = blockIdx.x - b * tps;
  const size_t tile_idx = size_t(b) * tps + mt;
  uint8_t *sk = sm + S::K, *sv = sm + S::VQ, *sq = sm + S::VQ, *sp = sm + S::P, *se = sm + S::E;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(&bars[0], uint32_t(tps) * (S::VB + 16 * kCS));
    for (int j = 0; j < tps; ++j)
      bulk_g2s(sv + j * S::VB, static_cast<const uint8_t*>(a.v_img) + (size_t(b) * tps + j) * S::VB, S::VB, &bars[0]);
    bulk_g2s(sp, static_cast<const uint8_t*>(a.p_img) + tile_idx * S::PB, uint32_t(tps) * 16 * kCS, &bars[0]);
    // the key tiles are not needed before the dQ / dK chains: their own barrier, behind the operands of the first products
    mbar_arrive_expect_tx(&bars[2], uint32_t(tps) * S::QB);
    for (int j = 0; j < tps; ++j)
      bulk_g2s(sk + j * S::QB, static_cast<const uint8_t*>(a.k_img) + (size_t(b) * tps + j) * S::QB, S::QB, &bars[2]);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < DV; i += 256) { ln_s[i] = a.ln_g[i]; ln_s[DV + i] = a.ln_b[i]; }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = uint32_t(32 * q) << 16;
  constexpr uint32_t COL_C = 0, COL_DV = 0, COL_DA = 192, COL_DQ = 0, COL_DK = 192;
  const int t = mt * kTile + row;                 // my query token
  const bool valid = t < L;
  const int kpad0 = L - (tps - 1) * kTile;        // first pad row of the last key tile
  const int qpad0 = min(kTile, L - mt * kTile);   // first pad row of my query tile
  // (global loads of my row issued before the wait for the operand tiles)
  const float sum = valid ? a.sum[tile_idx * kTile + row] : 1.0f;
  const float inv = valid ? 1.0f / sum : 0.f, rinv = 1.0f / float(L - 1);
  float dy[HC];                                      // d(out) of my columns, then d(pre-LayerNorm context)
  if (valid) {
    const float* src = a.d_out + (size_t(b) * L + t) * DV + HC * half;
#pragma unroll
    for (int e = 0; e < HC; e += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + e));
      dy[e] = v.x; dy[e + 1] = v.y; dy[e + 2] = v.z; dy[e + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int e = 0; e < HC; ++e) dy[e] = 0.f;
  }
  mbar_wait(&bars[0], 0);
  tc_fence_after_sync();
  // rows beyond the sample's L tokens were never written by the forward: they meet zero weights in the products below
  // and must be finite
  {
    uint8_t* vlast = sv + (tps - 1) * S::VB;
    for (int i = tid; i < (kTile - kpad0) * (DV / 8); i += 256) {
      const int r = kpad0 + i / (DV / 8), g = i % (DV / 8);
      *reinterpret_cast<uint4*>(vlast + uint32_t(g) * kCS + row_off(r)) = make_uint4(0u, 0u, 0u, 0u);
    }
    const int pg = ncols / 8;
    for (int i = tid; i < (kTile - qpad0) * pg; i += 256) {
      const int r = qpad0 + i / pg, g = i % pg;
      *reinterpret_cast<uint4*>(sp + uint32_t(g) * kCS + row_off(r)) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (a.reverse && tid < DV) {
      float acc0 = 0.f, acc1 = 0.f;
      for (int j = 0; j < L; j += 2) {
        const uint8_t* v0 = sv + (j >> 7) * S::VB + uint32_t(tid >> 3) * kCS + row_off(j & 127) + (tid & 7) * 2;
        acc0 += __uint_as_float(uint32_t(*reinterpret_cast<const uint16_t*>(v0)) << 16);
        if (j + 1 < L) {
          const uint8_t* v1 = sv + ((j + 1) >> 7) * S::VB + uint32_t(tid >> 3) * kCS + row_off((j + 1) & 127) + (tid & 7) * 2;
          acc1 += __uint_as_float(uint32_t(*reinterpret_cast<const uint16_t*>(v1)) << 16);
        }
      }
      vsum[tid] = acc0 + acc1;
    }
  }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  // ---- C = P V, as the forward did -----------------------------------------------------------------------------------------
  if (tid == 0) {
    for (int j = 0; j < tps; ++j)
      htc::mma_steps(tmem + COL_C, make_smem_desc(smem_u32(sp + uint32_t(16 * j) * kCS), kCS, kRS), 2 * kCS,
                     make_smem_desc(smem_u32(sv + j * S::VB), kRS, kCS), 2 * kRS, make_idesc_bf16(128, DV, 0, 1), 8, j > 0);
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after_sync();
  // ---- LayerNorm + ReLU backward of my row (:65-66, :105-106) -> E = d(P_un V) = d(context) / sum (reverse: * -1/(L-1)) ------
  {
    float x[HC];
    {
      uint32_t raw[HC];
      const uint32_t tc_ = tmem + lane_base + COL_C + HC * half;
#pragma unroll
      for (int j = 0; j < HC / 8; ++j) tmem_ld8_nw(tc_ + 8 * j, *reinterpret_cast<uint32_t(*)[8]>(&raw[8 * j]));
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < HC; ++e) {
        const float cun = __uint_as_float(raw[e]) * inv;
        x[e] = a.reverse ? (vsum[HC * half + e] - cun) * rinv : cun;
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int e = 0; e < HC; ++e) { s1 += x[e]; s2 = fmaf(x[e], x[e], s2); }
    part[half * kTile + row] = make_float2(s1, s2);
    __syncthreads();
    const float2 o = part[(half ^ 1) * kTile + row];
    const float mean = (s1 + o.x) * (1.0f / float(DV));
    const float rstd = rsqrtf(fmaxf((s2 + o.y) * (1.0f / float(DV)) - mean * mean, 0.f) + kLnEps);
    const float* gam = ln_s + HC * half;
    const float* bet = ln_s + DV + HC * half;
    float t1 = 0.f, t2 = 0.f;
    float tmp[HC];
#pragma unroll
    for (int e = 0; e < HC; ++e) {
      const float xh = (x[e] - mean) * rstd;
      const float g = fmaf(xh, gam[e], bet[e]) > 0.f ? dy[e] : 0.f;      // ReLU gate
      const float dxh = g * gam[e];
      t1 += dxh;
      t2 = fmaf(dxh, xh, t2);
      x[e] = xh;
      dy[e] = g;
      tmp[e] = g * xh;
    }
    // column sums over the tile's rows: d(gamma), d(beta)
    warp_col_sums(tmp, lane, myslot + HC * half);
#pragma unroll
    for (int e = 0; e < HC; ++e) tmp[e] = dy[e];
    warp_col_sums(tmp, lane, myslot + DV + HC * half);
    __syncthreads();
    part[half * kTile + row] = make_float2(t1, t2);
    __syncthreads();
    const float2 o2 = part[(half ^ 1) * kTile + row];
    const float m1 = (t1 + o2.x) * (1.0f / float(DV)), m2 = (t2 + o2.y) * (1.0f / float(DV));
    const float esc = a.reverse ? -inv * rinv : inv;
#pragma unroll
    for (int e = 0; e < HC; ++e)      // d(loss)/d(context row before the LayerNorm)
      dy[e] = valid ? rstd * (dy[e] * gam[e] - m1 - x[e] * m2) : 0.f;
#pragma unroll
    for (int g8 = 0; g8 < HC / 8; ++g8) {
      float ev[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) ev[e] = dy[8 * g8 + e] * esc;
      *reinterpret_cast<uint4*>(se + uint32_t((HC / 8) * half + g8) * kCS + row_off(row)) = pack_bf16x8(ev);
    }
    if (a.reverse) {                  // d(colsum V)
#pragma unroll
      for (int e = 0; e < HC; ++e) dy[e] *= rinv;
      warp_col_sums(dy, lane, myslot + 2 * DV + HC * half);
    }
  }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (a.reverse && tid < DV)      // d(colsum V) of the tile (vsum is dead: reuse); read after the softmax phase's barrier
    vsum[tid] = slots[2 * DV + tid] + slots[S::NS + 2 * DV + tid] + slots[2 * S::NS + 2 * DV + tid] + slots[3 * S::NS + 2 * DV + tid];
  // ---- d(weights) = E V^T (per key tile, N = 128) and dV = P^T E (per key tile, M = 128 keys) ------------------------------------
  if (tid == 0) {
    for (int j = 0; j < tps; ++j)
      htc::mma_steps(tmem + COL_DA + 128 * j, make_smem_desc(smem_u32(se), kCS, kRS), 2 * kCS,
                     make_smem_desc(smem_u32(sv + j * S::VB), kCS, kRS), 2 * kCS, make_idesc_bf16(128, 128, 0, 0), DV / 16, false);
    for (int j = 0; j < tps; ++j)
      htc::mma_steps(tmem + COL_DV + DV * j, make_smem_desc(smem_u32(sp + uint32_t(16 * j) * kCS), kRS, kCS), 2 * kRS,
                     make_smem_desc(smem_u32(se), kRS, kCS), 2 * kRS, make_idesc_bf16(128, DV, 1, 1), 8, false);
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 1);
  tc_fence_after_sync();
  // the value tiles are dead: the query tile takes their place
  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], S::QB);
    bulk_g2s(sq, static_cast<const uint8_t*>(a.q_img) + tile_idx * S::QB, S::QB, &bars[0]);
  }
  // ---- softmax backward in place: dS = P_un (dA' - dot / sum), dot = sum_j dA'_j P_un_j (dA' = dA / sum: E carries 1/sum) -----
  // (measured and dropped: dot = E . (P_un V) from the LayerNorm phase's registers, which saves the first pass (2 us per
  //  launch) - but the fp32 dot no longer matches the bf16 products term by term, the rows of dS stop summing to zero and
  //  d(b_key), analytically zero, picks up noise at 1.5 % of d(b_query)'s scale)
  {
    const uint32_t ta = tmem + lane_base + COL_DA;
    const int nchunks = ncols / 16;
    float dot = 0.f;
    {
      uint32_t nx[16];
      tmem_ld16_nw(ta + 16 * half, nx);
#pragma unroll 1
      for (int ch = half; ch < nchunks; ch += 2) {
        float p[16];
        ld_chunks16(sp, row, 16 * ch, p);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) dot = fmaf(__uint_as_float(nx[e]), p[e], dot);
        if (ch + 2 < nchunks) tmem_ld16_nw(ta + 16 * (ch + 2), nx);
      }
    }
    part[half * kTile + row].x = dot;
    __syncthreads();
    dot = (dot + part[(half ^ 1) * kTile + row].x) * inv;
    {
      uint32_t nx[16];
      tmem_ld16_nw(ta + 16 * half, nx);
#pragma unroll 1
      for (int ch = half; ch < nchunks; ch += 2) {
        float p[16];
        ld_chunks16(sp, row, 16 * ch, p);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) p[e] *= __uint_as_float(nx[e]) - dot;
        if (ch + 2 < nchunks) tmem_ld16_nw(ta + 16 * (ch + 2), nx);
        htc::st_chunks16(sp, row, 16 * ch, p);
      }
    }
  }
  // the query tile and the key tiles have landed; their rows beyond L must be finite (zero)
  mbar_wait(&bars[0], 1);
  mbar_wait(&bars[2], 0);
  for (int i = tid; i < (kTile - qpad0) * (DKQ / 8); i += 256) {
    const int r = qpad0 + i / (DKQ / 8), g = i % (DKQ / 8);
    *reinterpret_cast<uint4*>(sq + uint32_t(g) * kCS + row_off(r)) = make_uint4(0u, 0u, 0u, 0u);
  }
  {
    uint8_t* klast = sk + (tps - 1) * S::QB;
    for (int i = tid; i < (kTile - kpad0) * (DKQ / 8); i += 256) {
      const int r = kpad0 + i / (DKQ / 8), g = i % (DKQ / 8);
      *reinterpret_cast<uint4*>(klast + uint32_t(g) * kCS + row_off(r)) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  // ---- dK = dS^T Q (per key tile) into the columns the softmax backward has just released; runs while dV is drained --------------
  if (tid == 0) {
    for (int j = 0; j < tps; ++j)
      htc::mma_steps(tmem + COL_DK + DKQ * j, make_smem_desc(smem_u32(sp + uint32_t(16 * j) * kCS), kRS, kCS), 2 * kRS,
                     make_smem_desc(smem_u32(sq), kRS, kCS), 2 * kRS, make_idesc_bf16(128, DKQ, 1, 1), 8, false);
    umma_commit(&bars[1]);
  }
  // LayerNorm-affine gradients of this tile
  if (tid < DV) {
    atomicAdd(a.g_ln_g + tid, slots[tid] + slots[S::NS + tid] + slots[2 * S::NS + tid] + slots[3 * S::NS + tid]);
    atomicAdd(a.g_ln_b + tid, slots[DV + tid] + slots[S::NS + DV + tid] + slots[2 * S::NS + DV + tid] + slots[3 * S::NS + DV + tid]);
  }
  // ---- drain dV: my row is KEY 128 j + row of key tile j ------------------------------------------------------------------------
  {
    float vtot[HC];
#pragma unroll
    for (int e = 0; e < HC; ++e) vtot[e] = 0.f;
    for (int j = 0; j < tps; ++j) {
      uint32_t raw[HC];
      const uint32_t tv = tmem + lane_base + COL_DV + DV * j + HC * half;
#pragma unroll
      for (int g = 0; g < HC / 8; ++g) tmem_ld8_nw(tv + 8 * g, *reinterpret_cast<uint32_t(*)[8]>(&raw[8 * g]));
      tmem_wait_ld();
      const int key = j * kTile + row;
      const bool add_t = a.reverse && key < L;        // every valid key's value row feeds colsum(V)
      __nv_bfloat16* dst = a.g_kv + (size_t(b) * L + key) * a.ld_kv + DKQ + HC * half;
      float v[HC];
#pragma unroll
      for (int e = 0; e < HC; ++e) {
        v[e] = __uint_as_float(raw[e]) + (add_t ? vsum[HC * half + e] : 0.f);
        vtot[e] += v[e];
      }
      if (key < L) {
#pragma unroll
        for (int g8 = 0; g8 < HC / 8; ++g8) put_bf16x8(dst + 8 * g8, *reinterpret_cast<const float(*)[8]>(&v[8 * g8]), a.kv_atomic != 0);
      }
    }
    warp_col_sums(vtot, lane, myslot + SLOT_B + 2 * DKQ + HC * half);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  // ---- dQ = dS K (accumulated over the key tiles) into the drained dV columns; runs while dK is drained ------------------------------
  if (tid == 0) {
    for (int j = 0; j < tps; ++j)
      htc::mma_steps(tmem + COL_DQ, make_smem_desc(smem_u32(sp + uint32_t(16 * j) * kCS), kCS, kRS), 2 * kCS,
                     make_smem_desc(smem_u32(sk + j * S::QB), kRS, kCS), 2 * kRS, make_idesc_bf16(128, DKQ, 0, 1), 8, j > 0);
    umma_commit(&bars[3]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after_sync();
  {
  float ktot[HQ];
#pragma unroll
  for (int e = 0; e < HQ; ++e) ktot[e] = 0.f;
  for (int j = 0; j < tps; ++j) {
    const int key = j * kTile + row;
    __nv_bfloat16* dst = a.g_kv + (size_t(b) * L + key) * a.ld_kv + HQ * half;
    const uint32_t tk = tmem + lane_base + COL_DK + DKQ * j + HQ * half;
    uint32_t r[HQ];
#pragma unroll
    for (int c0 = 0; c0 < HQ; c0 += 16) tmem_ld16_nw(tk + c0, *reinterpret_cast<uint32_t(*)[16]>(&r[c0]));
    tmem_wait_ld();
    float v[HQ];
#pragma unroll
    for (int e = 0; e < HQ; ++e) { v[e] = __uint_as_float(r[e]); ktot[e] += v[e]; }
    if (key < L) {
#pragma unroll
      for (int g8 = 0; g8 < HQ / 8; ++g8) put_bf16x8(dst + 8 * g8, *reinterpret_cast<const float(*)[8]>(&v[8 * g8]), a.kv_atomic != 0);
    }
  }
  warp_col_sums(ktot, lane, myslot + SLOT_B + DKQ + HQ * half);
  }
  mbar_wait(&bars[3], 0);
  tc_fence_after_sync();
  {
    __nv_bfloat16* dst = a.g_q + (size_t(b) * L + t) * a.ld_q + HQ * half;
    const uint32_t tq = tmem + lane_base + COL_DQ + HQ * half;
    uint32_t r[HQ];
#pragma unroll
    for (int c0 = 0; c0 < HQ; c0 += 16) tmem_ld16_nw(tq + c0, *reinterpret_cast<uint32_t(*)[16]>(&r[c0]));
    tmem_wait_ld();
    float v[HQ];
#pragma unroll
    for (int e = 0; e < HQ; ++e) v[e] = __uint_as_float(r[e]) * a.qscale;      // scores = (q / sqrt(d_kq)) . k
    if (valid) {
#pragma unroll
      for (int g8 = 0; g8 < HQ / 8; ++g8)
        *reinterpret_cast<uint4*>(dst + 8 * g8) = pack_bf16x8(*reinterpret_cast<const float(*)[8]>(&v[8 * g8]));
    }
    warp_col_sums(v, lane, myslot + SLOT_B + HQ * half);
  }
  __syncthreads();
  for (int i = tid; i < 2 * DKQ + DV; i += 256) {
    const float* sl = slots + SLOT_B + i;
    atomicAdd(i < DKQ ? a.g_bq + i : (i < 2 * DKQ ? a.g_bk + (i - DKQ) : a.g_bv + (i - 2 * DKQ)),
              sl[0] + sl[S::NS] + sl[2 * S::NS] + sl[3 * S::NS]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- weight gradients ---------------------------------------------------------------------------------------------------------------
constexpr int kGradThreads = 192;                 // producer warp, MMA warp, 4 epilogue warps
constexpr int kGradStages = 3;
constexpr int kGradKT = 64;                       // tokens per pipeline stage
constexpr uint32_t kBoxBytes = 64 * 128;          // one {64 columns, 64 rows} bf16 box, 128-byte swizzled: 8 KB
constexpr int kMaxGBlocks = 6;                    // gradient columns <= 384

struct WgradOut { float* g_w; int n0, cols, ld; };      // gradient columns [n0, n0 + cols) -> dW [cols][K] += (row pitch ld, 0: K;
                                                        // n0, cols: multiples of 16)
struct WgradArgs {
  WgradOut out[3]; int nout, NG, K, rows;
  float* part;       // [splits][NG][K] fp32: every split writes its own slab (plain coalesced stores), tok_wgrad_reduce adds
                     // them into the gradients - measured against red.global.add straight from the epilogue (45 k scalar
                     // atomics per CTA): 58 -> 40 us for the ViT-L/16 block, and the sum order is fixed
};

// grid = (ceil(K / 128), splits of the 64-token chunks).  D[m = X column][n = gradient column] over k = token.
// tm_x: X [rows][K], tm_g: G [rows][NG]; both with {64, 64} boxes: per stage 2 boxes of X columns, ceil(NG / 64) of G.
__global__ void __launch_bounds__(kGradThreads, 1) tok_wgrad_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                                    const __grid_constant__ CUtensorMap tm_g, const WgradArgs a) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
  const int nblk = (a.NG + 63) / 64;
  const uint32_t stage = uint32_t(2 + nblk) * kBoxBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + kGradStages * stage);
  uint64_t* empty = full + kGradStages;
  uint64_t* accb = empty + kGradStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accb + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * 128;
  const int chunks = (a.rows + kGradKT - 1) / kGradKT;
  const int c0 = int((long long)chunks * blockIdx.y / gridDim.y), c1 = int((long long)chunks * (blockIdx.y + 1) / gridDim.y);
  const int n_it = c1 - c0;
  if (n_it <= 0) return;
  if (tid == 0) {
    for (int i = 0; i < 2 * kGradStages + 1; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_g);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  // accumulators: the gradient columns [0, n1) and, beyond 256 columns, [192, NG) - both starting on a 64-column box
  const int n1 = a.NG <= 256 ? a.NG : 192, n2 = a.NG - n1;
  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_it; ++it) {
        const int s = it % kGradStages, tok0 = (c0 + it) * kGradKT;
        if (it >= kGradStages) mbar_wait(&empty[s], uint32_t(it / kGradStages - 1) & 1u);
        mbar_arrive_expect_tx(&full[s], stage);
        uint8_t* st = sm + s * stage;
        tma_load_2d(st, &tm_x, m0, tok0, &full[s]);
        tma_load_2d(st + kBoxBytes, &tm_x, m0 + 64, tok0, &full[s]);
        for (int j = 0; j < nblk; ++j) tma_load_2d(st + (2 + j) * kBoxBytes, &tm_g, 64 * j, tok0, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc_bf16(128, n1, 1, 1), idesc2 = make_idesc_bf16(128, n2 > 0 ? n2 : 16, 1, 1);
      for (int it = 0; it < n_it; ++it) {
        const int s = it % kGradStages;
        mbar_wait(&full[s], uint32_t(it / kGradStages) & 1u);
        tc_fence_after_sync();
        const uint32_t st = smem_u32(sm + s * stage);
        const uint64_t ad = make_smem_desc_sw128_mn(st, kBoxBytes);
        const uint64_t bd1 = make_smem_desc_sw128_mn(st + 2 * kBoxBytes, kBoxBytes);
        const uint64_t bd2 = make_smem_desc_sw128_mn(st + 5 * kBoxBytes, kBoxBytes);
#pragma unroll
        for (int ks = 0; ks < kGradKT / 16; ++ks) {
          umma_bf16(tmem, desc_advance(ad, ks * 2048), desc_advance(bd1, ks * 2048), idesc1, (it | ks) ? 1u : 0u);
          if (n2 > 0) umma_bf16(tmem + 192, desc_advance(ad, ks * 2048), desc_advance(bd2, ks * 2048), idesc2, (it | ks) ? 1u : 0u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(accb);
    }
  } else {
    const int q = warp & 3, gm = m0 + 32 * q + lane;
    float* dst = a.part + size_t(blockIdx.y) * a.NG * a.K + gm;
    mbar_wait(accb, 0);
    tc_fence_after_sync();
    uint32_t nx[16];
    tmem_ld16_nw(tmem + (uint32_t(32 * q) << 16), nx);
#pragma unroll 1
    for (int c = 0; c < a.NG; c += 16) {
      uint32_t r[16];
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 16; ++e) r[e] = nx[e];
      if (c + 16 < a.NG) tmem_ld16_nw(tmem + (uint32_t(32 * q) << 16) + uint32_t(c + 16), nx);
      if (gm < a.K) {
#pragma unroll
        for (int e = 0; e < 16; ++e) dst[size_t(c + e) * a.K] = __uint_as_float(r[e]);      // lanes: consecutive floats
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// dW[n][m] += sum over the splits of part[split][n][m], four consecutive m per thread.  SL = 1: a thread walks all the slabs
// (large outputs: enough threads anyway); SL = 8: a block is 32 quads x 8 split lanes (small outputs with many slabs).
template <int SL>
__global__ void __launch_bounds__(256) tok_wgrad_reduce_kernel(const WgradArgs a, int splits) {
  __shared__ float4 red[SL > 1 ? SL : 1][32];
  const long long total4 = (long long)a.NG * a.K / 4;
  const float4* part = reinterpret_cast<const float4*>(a.part);
  const int qpb = 256 / SL;                        // quads per block
  const int ql = threadIdx.x % qpb, sl = threadIdx.x / qpb;
  for (long long base = (long long)blockIdx.x * qpb; base < total4; base += (long long)gridDim.x * qpb) {
    const long long i = base + ql;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < total4) {
#pragma unroll 6
      for (int s = sl; s < splits; s += SL) {
        const float4 v = __ldg(part + s * total4 + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    if (SL > 1) {
      red[sl][ql] = acc;
      __syncthreads();
      if (sl == 0) {
#pragma unroll
        for (int k = 1; k < SL; ++k) { const float4 v = red[k][ql]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
      }
    }
    if (sl == 0 && i < total4) {
      const int n = int(4 * i / a.K), m = int(4 * i - (long long)n * a.K);
#pragma unroll
      for (int o = 0; o < 3; ++o)
        if (o < a.nout && n >= a.out[o].n0 && n < a.out[o].n0 + a.out[o].cols) {
          float4* dst = reinterpret_cast<float4*>(a.out[o].g_w + size_t(n - a.out[o].n0) * (a.out[o].ld ? a.out[o].ld : a.K) + m);
          float4 g = *dst;
          g.x += acc.x; g.y += acc.y; g.z += acc.z; g.w += acc.w;
          *dst = g;
        }
    }
    if (SL > 1) __syncthreads();
  }
}

// ---- input gradients ------------------------------------------------------------------------------------------------------------------
struct DgradArgs { float* dx; int rows, K, NG; const float* bias; };     // dx [rows][K] = G [rows][NG] W (+ bias [K])
// grid = (ceil(rows / 128), ceil(K / 256)).  tm_g: G with {64 columns, 128 rows} boxes (K-major A); one stage per 64 columns
// of the contraction.  The second operand, as it lies in memory:
//   B_KMAJOR = false: W [NG][K] (contraction index outermost: MN-major B), {64 columns, 64 rows} boxes - the input gradient
//                     of a Linear layer, dX = dY W;
//   B_KMAJOR = true:  W [K][NG] (torch.nn.Linear's weight, contraction index contiguous: K-major B), one {64, bn rows} box
//                     per stage - the layer's forward, Y = X W^T + bias (the classic / normalized heads' projections).
template <bool B_KMAJOR>
__global__ void __launch_bounds__(kGradThreads, 1) tok_dgrad_kernel(const __grid_constant__ CUtensorMap tm_g,
                                                                    const __grid_constant__ CUtensorMap tm_w, const DgradArgs a) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
  const int n0 = blockIdx.y * 256, bn = min(256, a.K - n0), nb = (bn + 63) / 64;
  const uint32_t a_bytes = kTile * 128, b_bytes = B_KMAJOR ? uint32_t(bn) * 128 : uint32_t(nb) * kBoxBytes;
  const uint32_t stage = (a_bytes + b_bytes + 1023) & ~1023u;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + kGradStages * stage);
  uint64_t* empty = full + kGradStages;
  uint64_t* accb = empty + kGradStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accb + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * kTile, n_it = (a.NG + 63) / 64;
  if (tid == 0) {
    for (int i = 0; i < 2 * kGradStages + 1; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
    tma_prefetch_desc(&tm_g);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_it; ++it) {
        const int s = it % kGradStages;
        if (it >= kGradStages) mbar_wait(&empty[s], uint32_t(it / kGradStages - 1) & 1u);
        mbar_arrive_expect_tx(&full[s], a_bytes + b_bytes);
        uint8_t* st = sm + s * stage;
        tma_load_2d(st, &tm_g, 64 * it, row0, &full[s]);
        if (B_KMAJOR) {
          tma_load_2d(st + a_bytes, &tm_w, 64 * it, n0, &full[s]);
        } else {
          for (int j = 0; j < nb; ++j) tma_load_2d(st + a_bytes + j * kBoxBytes, &tm_w, n0 + 64 * j, 64 * it, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, bn, 0, B_KMAJOR ? 0 : 1);
      for (int it = 0; it < n_it; ++it) {
        const int s = it % kGradStages;
        mbar_wait(&full[s], uint32_t(it / kGradStages) & 1u);
        tc_fence_after_sync();
        const uint32_t st = smem_u32(sm + s * stage);
        const uint64_t ad = make_smem_desc_sw128(st);
        const uint64_t bd = B_KMAJOR ? make_smem_desc_sw128(st + a_bytes) : make_smem_desc_sw128_mn(st + a_bytes, kBoxBytes);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16(tmem, desc_advance(ad, ks * 32), desc_advance(bd, B_KMAJOR ? ks * 32 : ks * 2048), idesc, (it | ks) ? 1u : 0u);
        umma_commit(&empty[s]);
      }
      umma_commit(accb);
    }
  } else {
    const int q = warp & 3, r = row0 + 32 * q + lane;
    float* dst = a.dx + size_t(r) * a.K + n0;
    mbar_wait(accb, 0);
    tc_fence_after_sync();
#pragma unroll 1
    for (int c = 0; c < bn; c += 16) {
      uint32_t v[16];
      tmem_ld16_nw(tmem + (uint32_t(32 * q) << 16) + uint32_t(c), v);
      tmem_wait_ld();
      if (r < a.rows) {
        if (a.bias) {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __ldg(a.bias + n0 + c + e));
        }
#pragma unroll
        for (int e = 0; e < 16; e += 4) *reinterpret_cast<uint4*>(dst + c + e) = make_uint4(v[e], v[e + 1], v[e + 2], v[e + 3]);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

}  // namespace tok
}  // namespace mmrca

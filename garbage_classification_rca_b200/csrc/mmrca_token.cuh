// Token-level attention blocks (BASELINE.json configs[4]: ViT-L/16's 197 patch tokens x 1024, RoBERTa's 256 tokens x 768):
// SelfAttention (CVPR_code/multimodal_model.py:39-68) and ReverseCrossAttention (:71-108) for square L <= 256 on real
// token sequences [B, L, K] instead of the 16 pseudo-tokens of the pooled vector.  The reference classes are shape-generic
// (Linear on the last dimension, batched matmul, square-attention assert :93), so they define these shapes too (the parity tests check against them).
//
//   tok_proj   Q | K | V = X [W_query | W_key | W_value]^T + b : the one place on this path where the projection is a real
//              GEMM (K = 1024 / 768, N = 352).  Warp-specialised tcgen05 pipeline: one TMA producer thread
//              (cp.async.bulk.tensor through tensor maps, 128-byte swizzle: X is read as it lies in HBM, [B, L, K] bf16,
//              out-of-range token rows of a 128-row tile are zero-filled by the TMA unit), one MMA thread, four epilogue
//              warps; one CTA owns a 128-token tile and ALL N columns (two accumulators in TMEM), so X is read once.
//              Epilogue: + bias, Q pre-scaled by 1/sqrt(d_kq), bf16, written as operand IMAGES (canonical no-swizzle
//              core-matrix layout, one image per (sample, tile, Q|K|V)) that the attention kernel loads with one bulk copy.
//   tok_attn   per (sample, 128-query tile): S = Q K^T over the sample's <= 256 keys (accumulator [128 x 256] in TMEM),
//              softmax over the L valid keys (optionally the reverse weights (1 - A) / (L - 1)), C = P V, LayerNorm + ReLU.
//              One thread per query row for the softmax / LayerNorm; K / V tiles arrive by bulk copy (TMA engine).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mmrca_head_tc_bwd.cuh"

namespace mmrca {
namespace tok {

using namespace tc;
using htc::kCS;
using htc::kRS;
using htc::row_off;

constexpr int kTile = 128;            // tokens per tile
constexpr int kMaxTiles = 2;          // L <= 256
constexpr int kBK = 64;               // K per pipeline stage: one 128-byte swizzle span of bf16
constexpr int kStages = 3;
constexpr int kProjEpiGroups = 2;           // epilogue warps per TMEM lane quarter (4 measured: no change, the tile count per CTA - 394 tiles on 148 CTAs - sets the time)
constexpr int kProjThreads = 64 + 128 * kProjEpiGroups;      // producer warp, MMA warp, 4 x kProjEpiGroups epilogue warps

// ---- TMA tensor-map loads, 128B-swizzled shared-memory descriptors ------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// L2 prefetch of a box (no shared memory, nothing to wait for)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
// K-major operand whose rows are 128-byte swizzle spans (what a {64 bf16, rows} TMA box with SWIZZLE_128B lands):
// 8-row groups 1024 bytes apart, layout type 2 (SWIZZLE_128B); the leading-dimension offset is unused for this layout.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr >> 4) & 0x3FFF);
  d |= uint64_t(1) << 16;
  d |= uint64_t((1024u >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

// ---- projection GEMM -------------------------------------------------------------------------------------------------------
struct ProjSeg { void* img; int cols; float scale; };     // output columns [start, start + cols) -> image; values * scale
struct ProjArgs {
  const uint8_t* wblob;     // pre-swizzled weight images [k-block][N][128 B] (tok_wprep_kernel)
  ProjSeg seg[3];           // Q | K | V (cols == 0: absent)
  const float* bias;        // [N]
  int N, bn, nacc;          // N = bn * nacc total columns, bn <= 256 per accumulator
  int K, tiles_per_sample;
  int L, rows;              // tokens per sample, B * L: the GEMM's M dimension is the FLAT token index (no per-sample padding
                            // of the MMA work); the epilogue scatters row r to image (r / L, (r % L) / 128), row (r % L) % 128
};

__host__ __device__ constexpr uint32_t proj_stage_bytes(int n) { return uint32_t(kTile * kBK * 2 + n * kBK * 2); }

// Persistent: grid = min(tiles, SMs); a CTA walks the 128-row tiles blockIdx.x, + gridDim.x, ...  The producer's stage ring
// runs across tile boundaries, so the next tile's first stages land while the epilogue warps drain the accumulators; the
// MMA thread waits for the drain (tmem_empty) before it overwrites them.
__global__ void __launch_bounds__(kProjThreads, 1) tok_proj_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                                   const __grid_constant__ CUtensorMap tm_w, const ProjArgs a) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  // 1024-byte alignment of every operand buffer (SWIZZLE_128B)
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_bytes = kTile * kBK * 2, b_bytes = uint32_t(a.N) * kBK * 2, stage = a_bytes + b_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + kStages * stage);
  uint64_t* empty = full + kStages;
  uint64_t* accb = empty + kStages;          // accumulators complete (MMA -> epilogue)
  uint64_t* tmem_empty = accb + 1;           // accumulators drained (epilogue -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 2);      // [N]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles = (a.rows + kTile - 1) / kTile;
  const int n_it = (a.K + kBK - 1) / kBK;
  if (tid == 0) {
    for (int i = 0; i < 2 * kStages + 2; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < a.N; i += kProjThreads) bias_s[i] = __ldg(a.bias + i);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    if (lane == 0) {
      // (measured and dropped: L2 prefetch of the X boxes 8 k-blocks ahead, and a per-CTA rotation of the K order against
      //  hot L2 slices - neither moved the kernel: it runs at the L2 -> SM rate this tile shape needs)
      int git = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int it = 0; it < n_it; ++it, ++git) {
          const int s = git % kStages;
          if (git >= kStages) mbar_wait(&empty[s], uint32_t(git / kStages - 1) & 1u);
          mbar_arrive_expect_tx(&full[s], stage);
          uint8_t* sa = sm + s * stage;
          tma_load_2d(sa, &tm_x, it * kBK, tile * kTile, &full[s]);
          bulk_g2s(sa + a_bytes, a.wblob + size_t(it) * b_bytes, b_bytes, &full[s]);      // the whole [N x 64] weight slab
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kTile, a.bn, 0, 0);
      int git = 0, ti = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++ti) {
        if (ti > 0) { mbar_wait(tmem_empty, uint32_t(ti - 1) & 1u); tc_fence_after_sync(); }
        for (int it = 0; it < n_it; ++it, ++git) {
          const int s = git % kStages;
          mbar_wait(&full[s], uint32_t(git / kStages) & 1u);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(sm + s * stage);
          const uint64_t ad = make_smem_desc_sw128(sa);
#pragma unroll
          for (int ks = 0; ks < kBK / 16; ++ks) {
            for (int j = 0; j < a.nacc; ++j) {
              const uint64_t bd = make_smem_desc_sw128(sa + a_bytes + uint32_t(j) * uint32_t(a.bn) * kBK * 2);
              umma_bf16(tmem + uint32_t(j * a.bn), desc_advance(ad, ks * 32), desc_advance(bd, ks * 32), idesc, (it | ks) ? 1u : 0u);
            }
          }
          umma_commit(&empty[s]);
        }
        umma_commit(accb);
      }
    }
  } else {
    // ---- epilogue: 8 warps; warp w reads TMEM lanes 32 (w % 4) .. + 31 (its row of the tile) and every other 32-column
    //      chunk (half = (w - 2) / 4).  Measured on the first version (4 warps, per-chunk address arithmetic, segment
    //      table indexed dynamically): 9.5 us of fixed cost per 128-row tile against 13.7 us of main loop at K = 1024.
    const int q = warp & 3, half = (warp - 2) >> 2;      // (the chunk group of this warp, 0 .. kProjEpiGroups - 1)
    int ti = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++ti) {
      const int r_flat = tile * kTile + 32 * q + lane;
      const bool live = r_flat < a.rows;
      const int b = live ? r_flat / a.L : 0, t = r_flat - b * a.L, mt = t >> 7, row = t & 127;
      const size_t tile_idx = size_t(b) * a.tiles_per_sample + mt;
      const uint32_t roff = row_off(row);
      mbar_wait(accb, uint32_t(ti) & 1u);
      tc_fence_after_sync();
      int n0 = 0, chunk = 0;
#pragma unroll
      for (int sgi = 0; sgi < 3; ++sgi) {
        const ProjSeg sg = a.seg[sgi];                       // static index: stays in registers / constant bank
        if (sg.cols == 0) continue;
        uint8_t* img = static_cast<uint8_t*>(sg.img) + tile_idx * (size_t(sg.cols >> 3) * kCS) + roff;
        const float sc = sg.scale;
#pragma unroll 1
        for (int c = 0; c < sg.cols; c += 32, ++chunk) {
          if (chunk % kProjEpiGroups != half) continue;
          const bool two = c + 16 < sg.cols;
          uint32_t r0[16], r1[16];
          tmem_ld16_nw(tmem + (uint32_t(32 * q) << 16) + uint32_t(n0 + c), r0);
          if (two) tmem_ld16_nw(tmem + (uint32_t(32 * q) << 16) + uint32_t(n0 + c + 16), r1);
          tmem_wait_ld();
          const float* bs = bias_s + n0 + c;
          float v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = (__uint_as_float(r0[e]) + bs[e]) * sc;
          const float lo[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]};
          const float hi[8] = {v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]};
          uint8_t* dst = img + uint32_t(c >> 3) * kCS;
          if (live) {
            *reinterpret_cast<uint4*>(dst) = pack_bf16x8(lo);
            *reinterpret_cast<uint4*>(dst + kCS) = pack_bf16x8(hi);
          }
          if (two) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = (__uint_as_float(r1[e]) + bs[16 + e]) * sc;
            const float lo2[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]};
            const float hi2[8] = {v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]};
            if (live) {
              *reinterpret_cast<uint4*>(dst + 2 * kCS) = pack_bf16x8(lo2);
              *reinterpret_cast<uint4*>(dst + 3 * kCS) = pack_bf16x8(hi2);
            }
          }
        }
        n0 += sg.cols;
      }
      // the accumulators are drained: hand TMEM back to the MMA thread
      tc_fence_before_sync();
      named_bar_sync(1, 128 * kProjEpiGroups);
      if (tid == 64) mbar_arrive(tmem_empty);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// up to three fp32 -> bf16 casts in one launch (the stacked bf16 weights of the input-gradient GEMM)
struct Cast3Args { const float* src[3]; __nv_bfloat16* dst[3]; long long n[3]; };
__global__ void __launch_bounds__(256) cast3_bf16_kernel(const Cast3Args a) {
#pragma unroll 1
  for (int s = 0; s < 3; ++s) {
    const float* __restrict__ src = a.src[s];
    __nv_bfloat16* __restrict__ dst = a.dst[s];
    const long long n = a.n[s];
    for (long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * 1024) {
      if (i + 3 < n) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + i));
        *reinterpret_cast<uint2*>(dst + i) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
      } else {
        for (long long j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
      }
    }
  }
}

// dst[r][c] = bf16(src[r][c] / norm[r]) (norm null: plain cast), dst row pitch ld_dst: the hidden vectors of the classic /
// normalized heads, normalised and side by side, as the concat layer's GEMM operand
__global__ void __launch_bounds__(256) cast_rows_scaled_kernel(const float* __restrict__ src, const float* __restrict__ norm, int rows,
                                                               int width, __nv_bfloat16* __restrict__ dst, int ld_dst) {
  const int w8 = width / 8;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < (long long)rows * w8; i += (long long)gridDim.x * 256) {
    const int r = int(i / w8), c = int(i - (long long)r * w8) * 8;
    const float inv = norm ? 1.0f / __ldg(norm + r) : 1.0f;
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * width + c));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * width + c + 4));
    const float v[8] = {v0.x * inv, v0.y * inv, v0.z * inv, v0.w * inv, v1.x * inv, v1.y * inv, v1.z * inv, v1.w * inv};
    *reinterpret_cast<uint4*>(dst + (size_t)r * ld_dst + c) = pack_bf16x8(v);
  }
}

// W [n_rows][K] fp32 (rows n0 .. n0 + n_rows of the stacked [W_query; W_key; W_value]) -> bf16 blob laid out as the
// shared-memory IMAGE the projection's B operand wants: [k-block of 64][N rows][128 bytes, 16-byte chunks XOR-swizzled
// by (row % 8)] - what a {64, N} SWIZZLE_128B tensor-map box would land - so that a stage's whole weight slab is ONE
// contiguous bulk copy instead of N row requests through the tiled-TMA path.  Columns beyond K are zero.
struct WprepSeg { const float* w; const float* bias; float* bias_dst; uint8_t* blob; int n_rows, K, N, n0; };
struct WprepArgs { WprepSeg seg[3]; };
__global__ void __launch_bounds__(256) tok_wprep_kernel(const WprepArgs a) {
#pragma unroll 1
  for (int sgi = 0; sgi < 3; ++sgi) {
    const WprepSeg sg = a.seg[sgi];
    if (sg.n_rows == 0) continue;
    const int kblocks = (sg.K + kBK - 1) / kBK;
    const long long chunks = (long long)kblocks * sg.n_rows * 8;      // 16-byte chunks
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < chunks; i += (long long)gridDim.x * 256) {
      const int ch = int(i & 7);
      const long long t = i >> 3;
      const int n = int(t % sg.n_rows), kb = int(t / sg.n_rows);
      const int k0 = kb * kBK + ch * 8;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = k0 + e < sg.K ? __ldg(sg.w + (size_t)n * sg.K + k0 + e) : 0.f;
      const int row = sg.n0 + n;
      *reinterpret_cast<uint4*>(sg.blob + ((size_t)kb * sg.N + row) * 128 + uint32_t((ch ^ (row & 7)) * 16)) = pack_bf16x8(v);
    }
    if (blockIdx.x == 0)
      for (int i = threadIdx.x; i < sg.n_rows; i += 256) sg.bias_dst[i] = __ldg(sg.bias + i);
  }
}

// ---- attention ---------------------------------------------------------------------------------------------------------------
struct AttnArgs {
  const void* q_img; const void* k_img; const void* v_img;     // [B][tiles][cols/8 * kCS] operand images (tok_proj)
  const float* ln_g; const float* ln_b;
  float* out;               // [B][L][DV] fp32 (or, out_bf16 != 0, bf16: what a following block reads)
  int out_bf16;
  void* p_out;              // training: the unnormalised attention weights, [B * tiles][op_bytes(256)] bf16 images (null: not kept)
  float* sum_out;           // training: their row sums [B * tiles][128]
  int L, tiles_per_sample, reverse;
};

template <int DKQ, int DV>
struct AttnSmem {
  static constexpr uint32_t QB = htc::op_bytes(DKQ), VB = htc::op_bytes(DV), PB = htc::op_bytes(kMaxTiles * kTile);
  // The cross blocks (64 / 48) run TWO CTAs per SM: the attention weights P overwrite the key tiles (dead once the scores
  // are in TMEM), which brings a CTA under half the shared memory, and the context accumulator reuses the score columns
  // (256 TMEM columns per CTA).  The 128 / 96 blocks do not fit twice either way and keep separate regions.
  static constexpr bool kTwoPerSm = DKQ == 64;
  static constexpr uint32_t Q = 0;
  static constexpr uint32_t K = htc::al128(Q + QB);                     // kMaxTiles key tiles
  static constexpr uint32_t V = htc::al128(K + (kTwoPerSm ? PB : kMaxTiles * QB));      // kMaxTiles value tiles
  static constexpr uint32_t P = kTwoPerSm ? K : htc::al128(V + kMaxTiles * VB);         // [128 x 256] unnormalised attention weights
  static constexpr uint32_t LN = htc::al128(kTwoPerSm ? V + kMaxTiles * VB : P + PB);   // gamma, beta
  static constexpr uint32_t VSUM = LN + 2 * DV * 4;                     // column sums of V over the valid keys (reverse weights)
  static constexpr uint32_t PART = VSUM + DV * 4;                       // [2 halves][128 rows] float2 exchange
  static constexpr uint32_t BAR = htc::al128(PART + 2 * kTile * 8);
  static constexpr uint32_t BYTES = BAR + 64;
  static_assert(BYTES <= (kTwoPerSm ? 113 * 1024 : 232448), "token attention does not fit shared memory");
};

// 256 threads: two threads per query row (warp w reads TMEM lanes 32 (w % 4) .., half = w / 4 takes every other 16-column
// chunk).  Per-thread TMEM reads are software-pipelined (the next chunk is in flight while the current one is used): with
// one thread per row and a wait after every load the kernel spent most of its time on TMEM-load latency.
// Two passes over the scores: (1) row maximum, (2) p = exp(s - max) written UNNORMALISED as the bf16 P operand while its row
// sum accumulates; the 1 / sum lands on the context row after P V.  The reverse weights (1 - A) / (L - 1) (:95-99) follow
// from the same unnormalised product: ((1 - A) V)[c] = colsum(V)[c] - (P_un V)[c] / sum, with colsum(V) over the sample's
// L valid keys computed once per CTA.
template <int DKQ, int DV>
__global__ void __launch_bounds__(256, DKQ == 64 ? 2 : 1) tok_attn_kernel(const AttnArgs a) {
  extern __shared__ __align__(128) uint8_t sm[];
  using S = AttnSmem<DKQ, DV>;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::BAR);      // [0] loads, [1] MMAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  float* ln_s = reinterpret_cast<float*>(sm + S::LN);
  float* vsum = reinterpret_cast<float*>(sm + S::VSUM);
  float2* part = reinterpret_cast<float2*>(sm + S::PART);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2, row = 32 * q + lane;
  const int tps = a.tiles_per_sample;
  const int b = blockIdx.x / tps, mt = blockIdx.x - b * tps;
  uint8_t *sq = sm + S::Q, *sk = sm + S::K, *sv = sm + S::V, *sp = sm + S::P;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(&bars[0], S::QB + uint32_t(tps) * (S::QB + S::VB));
    bulk_g2s(sq, static_cast<const uint8_t*>(a.q_img) + (size_t(b) * tps + mt) * S::QB, S::QB, &bars[0]);
    for (int j = 0; j < tps; ++j) {
      bulk_g2s(sk + j * S::QB, static_cast<const uint8_t*>(a.k_img) + (size_t(b) * tps + j) * S::QB, S::QB, &bars[0]);
      bulk_g2s(sv + j * S::VB, static_cast<const uint8_t*>(a.v_img) + (size_t(b) * tps + j) * S::VB, S::VB, &bars[0]);
    }
  }
  constexpr uint32_t kTmemCols = S::kTwoPerSm ? 256 : 512;
  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  for (int i = tid; i < DV; i += 256) { ln_s[i] = a.ln_g[i]; ln_s[DV + i] = a.ln_b[i]; }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = uint32_t(32 * q) << 16;
  constexpr uint32_t COL_S = 0, COL_C = S::kTwoPerSm ? 0 : 256;      // (two per SM: the context reuses the score columns)
  const int L = a.L, ncols = tps * kTile;
  mbar_wait(&bars[0], 0);
  tc_fence_after_sync();
  // ---- S = Q K^T (Q carries the 1/sqrt(d_kq) factor): one N = 128 chain per key tile ------------------------------------
  if (tid == 0) {
    for (int j = 0; j < tps; ++j)
      htc::mma_steps(tmem + COL_S + 128 * j, make_smem_desc(smem_u32(sq), kCS, kRS), 2 * kCS,
                     make_smem_desc(smem_u32(sk + j * S::QB), kCS, kRS), 2 * kCS, make_idesc_bf16(128, 128, 0, 0), DKQ / 16, false);
    umma_commit(&bars[1]);
  }
  // while the scores are computed: the value rows beyond the sample's L tokens were never written by the projection
  // and meet P = 0 in the P V product, so they must be finite: zero them; and colsum(V) over the valid keys
  {
    const int pad0 = L - (tps - 1) * kTile;           // first pad row of the last value tile
    uint8_t* vlast = sv + (tps - 1) * S::VB;
    for (int i = tid; i < (kTile - pad0) * (DV / 8); i += 256) {
      const int r = pad0 + i / (DV / 8), g = i % (DV / 8);
      *reinterpret_cast<uint4*>(vlast + uint32_t(g) * kCS + row_off(r)) = make_uint4(0u, 0u, 0u, 0u);
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
  mbar_wait(&bars[1], 0);
  tc_fence_after_sync();
  // ---- softmax statistics of my query row (multimodal_model.py:58-60, :89-91); my chunks: 16-column chunks of parity `half`
  const uint32_t ts = tmem + lane_base + COL_S;
  const int nchunks = ncols / 16;
  float m = -INFINITY;
  {
    uint32_t nx[16];
    tmem_ld16_nw(ts + 16 * half, nx);
#pragma unroll 1
    for (int ch = half; ch < nchunks; ch += 2) {
      tmem_wait_ld();
      float s[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) s[e] = __uint_as_float(nx[e]);
      if (ch + 2 < nchunks) tmem_ld16_nw(ts + 16 * (ch + 2), nx);
      const int c0 = 16 * ch;
#pragma unroll
      for (int e = 0; e < 16; ++e) if (c0 + e < L) m = fmaxf(m, s[e]);
    }
  }
  part[half * kTile + row].x = m;
  __syncthreads();
  m = fmaxf(m, part[(half ^ 1) * kTile + row].x);
  float sum = 0.f;
  {
    uint32_t nx[16];
    tmem_ld16_nw(ts + 16 * half, nx);
#pragma unroll 1
    for (int ch = half; ch < nchunks; ch += 2) {
      tmem_wait_ld();
      float p[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) p[e] = __uint_as_float(nx[e]);
      if (ch + 2 < nchunks) tmem_ld16_nw(ts + 16 * (ch + 2), nx);
      const int c0 = 16 * ch;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        p[e] = c0 + e < L ? __expf(p[e] - m) : 0.f;
        sum += p[e];
      }
      htc::st_chunks16(sp, row, c0, p);
    }
  }
  part[half * kTile + row].y = sum;
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  sum += part[(half ^ 1) * kTile + row].y;
  if (a.p_out) {      // the backward's inputs (mmrca_token_bwd.cuh): P as it feeds the tensor core, and the row sums
    if (tid == 32) {
      bulk_s2g(static_cast<uint8_t*>(a.p_out) + (size_t(b) * tps + mt) * S::PB, sp, uint32_t(tps) * 16 * kCS);
      bulk_commit();
    }
    if (half == 0) a.sum_out[(size_t(b) * tps + mt) * kTile + row] = sum;
  }
  // ---- C = P V: A = P (K-major over the keys), B = V tiles read MN-major -----------------------------------------------------
  if (tid == 0) {
    for (int j = 0; j < tps; ++j)
      htc::mma_steps(tmem + COL_C, make_smem_desc(smem_u32(sp + uint32_t(16 * j) * kCS), kCS, kRS), 2 * kCS,
                     make_smem_desc(smem_u32(sv + j * S::VB), kRS, kCS), 2 * kRS, make_idesc_bf16(128, DV, 0, 1), 8, j > 0);
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 1);
  tc_fence_after_sync();
  // ---- context row -> LayerNorm + ReLU (:65-66, :105-106); my columns: [half * DV / 2, + DV / 2) -----------------------------
  {
    constexpr int HC = DV / 2;                       // 48 or 24
    const float inv = 1.0f / sum, rinv = 1.0f / float(L - 1);
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
    __syncthreads();                                  // (the softmax exchange above has been read by everyone)
    part[half * kTile + row] = make_float2(s1, s2);
    __syncthreads();
    const float2 o = part[(half ^ 1) * kTile + row];
    const float mean = (s1 + o.x) * (1.0f / float(DV));
    const float rstd = rsqrtf(fmaxf((s2 + o.y) * (1.0f / float(DV)) - mean * mean, 0.f) + kLnEps);
    const int t = mt * kTile + row;
    if (t < L) {
      const float* gam = ln_s + HC * half;
      const float* bet = ln_s + DV + HC * half;
#pragma unroll
      for (int e = 0; e < HC; ++e) x[e] = fmaxf(fmaf((x[e] - mean) * rstd, gam[e], bet[e]), 0.f);
      if (a.out_bf16) {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(a.out) + (size_t(b) * L + t) * DV + HC * half;
#pragma unroll
        for (int e = 0; e < HC; e += 8) *reinterpret_cast<uint4*>(dst + e) = pack_bf16x8(*reinterpret_cast<const float(*)[8]>(&x[e]));
      } else {
        float* dst = a.out + (size_t(b) * L + t) * DV + HC * half;
#pragma unroll
        for (int e = 0; e < HC; e += 4) *reinterpret_cast<float4*>(dst + e) = make_float4(x[e], x[e + 1], x[e + 2], x[e + 3]);
      }
    }
  }
  if (a.p_out && tid == 32) bulk_wait_all();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace tok
}  // namespace mmrca

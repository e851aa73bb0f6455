// bf16 tensor-core attention-block kernels (the 2e-2-absolute contract path).
//
// One CTA (128 threads = 4 warps, one thread per tile row) owns a tile of 8 samples = 128 rows, the
// UMMA M.  Every contraction of the block is a tcgen05.mma with fp32 accumulation in TMEM:
//   proj   [128 x DIN] . [DIN x (2 DKQ + DV)]        Q|K|V   (stacked W_query/W_key/W_value)
//   S      [128 x DKQ] . [DKQ x 128]                 scores of all 8 samples at once; only the 8
//                                                    diagonal 16x16 blocks are read back
//   ctx    [128 x 128] . [128 x DV]                  block-diagonal attention weights times V
// Between the MMAs each thread owns one row: it reads its accumulator row from TMEM, applies
// bias / 1/||x|| / softmax / reverse weights / LayerNorm / ReLU in registers (no shuffles) and writes
// the next bf16 operand into shared memory in the canonical no-swizzle core-matrix layout
// (mmrca_tc.cuh).  Weights arrive as a pre-packed bf16 blob through the bulk-copy engine (TMA).
// Q/K/V, the scores and the attention weights never leave the SM.
#pragma once
#include "mmrca_attn_fp32.cuh"
#include "mmrca_tc.cuh"

namespace mmrca {

constexpr int kTcThreads = 128;
constexpr int kTcRows = 128;                   // rows per tile (UMMA M)
constexpr int kTcG = kTcRows / kL;             // 8 samples per tile
constexpr uint32_t kOpLbo = 16 * 128 + 16;     // 128-row operand: bytes between adjacent 8-column groups (+16 pad)
constexpr uint32_t kOpSbo = 128;               // bytes between adjacent 8-row groups

template <int DIN_, int DKQ_, int DV_, bool SELF_>
struct TcCfg {
  static constexpr int DIN = DIN_, DKQ = DKQ_, DV = DV_;
  static constexpr bool SELF = SELF_;
  static constexpr int NKV = DKQ + DV, NALL = 2 * DKQ + DV;
  static constexpr uint32_t W_LBO = (NALL / 8) * 128;                 // packed weight blob: [DIN/8][NALL/8][8][8]
  static constexpr uint32_t W_BYTES = (DIN / 8) * W_LBO;
  static constexpr uint32_t X_BYTES = (DIN / 8) * kOpLbo;
  static constexpr uint32_t Q_BYTES = (DKQ / 8) * kOpLbo;
  static constexpr uint32_t V_BYTES = (DV / 8) * kOpLbo;
  static constexpr uint32_t P_BYTES = (kTcRows / 8) * kOpLbo;
  static_assert(DIN % 16 == 0 && DKQ % 16 == 0 && DV % 16 == 0 && NALL % 16 == 0, "UMMA K/N granularity");
  static constexpr uint32_t al(uint32_t v) { return (v + 127u) & ~127u; }
  static constexpr uint32_t OFF_W = 0;
  static constexpr uint32_t OFF_XQ = al(OFF_W + W_BYTES);
  static constexpr uint32_t OFF_XKV = SELF ? OFF_XQ : al(OFF_XQ + X_BYTES);
  static constexpr uint32_t OFF_Q = al(OFF_XKV + X_BYTES);
  static constexpr uint32_t OFF_K = al(OFF_Q + Q_BYTES);
  static constexpr uint32_t OFF_V = al(OFF_K + Q_BYTES);
  static constexpr uint32_t OFF_P = al(OFF_V + V_BYTES);
  static constexpr uint32_t OFF_BIAS = al(OFF_P + P_BYTES);          // fp32 [NALL]
  static constexpr uint32_t OFF_LN = OFF_BIAS + NALL * 4;             // fp32 gamma[DV], beta[DV]
  static constexpr uint32_t OFF_NORM = OFF_LN + 2 * DV * 4;           // fp32 [8] 1/||x|| per sample
  static constexpr uint32_t OFF_BAR = al(OFF_NORM + 8 * 4);           // 2 mbarriers + tmem base
  static constexpr uint32_t SMEM_BYTES = OFF_BAR + 64;
  static_assert(SMEM_BYTES <= 232448, "tile does not fit shared memory");
  // TMEM columns
  static constexpr uint32_t TM_PROJ = 0;        // [0, NALL): Q | K | V; later [0, DV): ctx
  static constexpr uint32_t TM_S = 384;         // [384, 512): scores
  static constexpr uint32_t TM_COLS = 512;
  static_assert(NALL <= 384, "projection accumulator overlaps the score columns");
};

// ---- weight pre-pack: fp32 [n][k] Linear weights -> bf16 canonical K-major blob --------------------------
struct PackJob { const float* wq; const float* wk; const float* wv; void* dst; int din, dkq, dv; };
struct PackArgs { PackJob job[4]; int njobs; };

__global__ void __launch_bounds__(256) pack_weights_kernel(const PackArgs a) {
  const PackJob& j = a.job[blockIdx.y];
  if (int(blockIdx.y) >= a.njobs) return;
  const int nall = 2 * j.dkq + j.dv, kcs = j.din / 8;
  const uint32_t lbo = (nall / 8) * 128;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nall * kcs; i += gridDim.x * blockDim.x) {
    const int n = i / kcs, kc = i - n * kcs;
    const float* src = n < j.dkq ? j.wq + size_t(n) * j.din
                                 : (n < 2 * j.dkq ? j.wk + size_t(n - j.dkq) * j.din : j.wv + size_t(n - 2 * j.dkq) * j.din);
    const float4 lo = __ldg(reinterpret_cast<const float4*>(src + kc * 8));
    const float4 hi = __ldg(reinterpret_cast<const float4*>(src + kc * 8 + 4));
    const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    *reinterpret_cast<uint4*>(static_cast<uint8_t*>(j.dst) + uint32_t(kc) * lbo + uint32_t(n >> 3) * 128u +
                              uint32_t(n & 7) * 16u) = tc::pack_bf16x8(v);
  }
}

struct TcAttnArgs {
  const float* xq;     // [B,16,DIN] fp32 (raw features when normalise)
  const float* xkv;    // [B,16,DIN] (== xq for self attention)
  const void* wblob;   // packed bf16 weights (pack_weights_kernel)
  const float* bq; const float* bk; const float* bv; const float* ln_g; const float* ln_b;
  float* out;          // [B,16,DV] fp32
  float* norms;        // [B] written when normalise
  int batch, reverse, normalise;
};

// Load this warp's two samples of the tile (fp32, coalesced 32-byte items), optionally accumulate the
// per-sample sum of squares, and store them as a bf16 K-major operand.
template <int DIN>
__device__ __forceinline__ void tc_stage_rows(uint8_t* op, const float* __restrict__ src, int b0, int batch,
                                              bool want_norm, float* inv_norm_s, float* norms_g) {
  constexpr int KCS = DIN / 8, ITEMS = kL * KCS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int g = warp * 2 + s, b = b0 + g;
    const bool valid = b < batch;
    const float* base = src + size_t(b) * kL * DIN;
    float ss = 0.f;
    for (int i = lane; i < ITEMS; i += 32) {
      const int row = i / KCS, kc = i - row * KCS;
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (valid) {
        const float4 lo = __ldg(reinterpret_cast<const float4*>(base + i * 8));
        const float4 hi = __ldg(reinterpret_cast<const float4*>(base + i * 8 + 4));
        v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
#pragma unroll
        for (int e = 0; e < 8; ++e) ss = fmaf(v[e], v[e], ss);
      }
      *reinterpret_cast<uint4*>(op + tc::core_off(g * kL + row, kc, kOpLbo, kOpSbo)) = tc::pack_bf16x8(v);
    }
    if (want_norm) {
      ss = warp_sum(ss);
      if (lane == 0) {
        const float nrm = sqrtf(ss);
        inv_norm_s[g] = valid ? 1.0f / nrm : 0.f;
        if (valid && norms_g) norms_g[b] = nrm;
      }
    } else if (lane == 0) {
      inv_norm_s[g] = 1.0f;
    }
  }
}

// One MMA phase boundary: make this thread's generic-proxy smem writes visible to the tensor core, sync the
// CTA, let thread 0 issue; everybody then waits on the mbarrier the commit arrives on.
__device__ __forceinline__ void tc_phase_sync() {
  tc::fence_proxy_async();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
}

template <class C>
__global__ void __launch_bounds__(kTcThreads, 1) attn_fwd_tc_kernel(const TcAttnArgs a) {
  extern __shared__ __align__(128) uint8_t sm_tc[];
  uint8_t* sm = sm_tc;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
  uint64_t* bar_mma = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
  float* bias_s = reinterpret_cast<float*>(sm + C::OFF_BIAS);
  float* ln_s = reinterpret_cast<float*>(sm + C::OFF_LN);
  float* inv_norm_s = reinterpret_cast<float*>(sm + C::OFF_NORM);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles = (a.batch + kTcG - 1) / kTcG;
  const bool reverse = a.reverse != 0;

  if (tid == 0) {
    tc::mbar_init(bar_w, 1);
    tc::mbar_init(bar_mma, 1);
    tc::mbar_fence_init();
    tc::mbar_arrive_expect_tx(bar_w, C::W_BYTES);
#pragma unroll 1
    for (int kc = 0; kc < C::DIN / 8; ++kc)
      tc::bulk_g2s(sm + C::OFF_W + kc * C::W_LBO, static_cast<const uint8_t*>(a.wblob) + kc * C::W_LBO, C::W_LBO, bar_w);
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, C::TM_COLS);
  for (int i = tid; i < C::NALL; i += kTcThreads)
    bias_s[i] = i < C::DKQ ? a.bq[i] : (i < 2 * C::DKQ ? a.bk[i - C::DKQ] : a.bv[i - 2 * C::DKQ]);
  for (int i = tid; i < C::DV; i += kTcThreads) { ln_s[i] = a.ln_g[i]; ln_s[C::DV + i] = a.ln_b[i]; }
  // the off-diagonal blocks of the attention-weight operand stay zero for the whole kernel
  for (uint32_t i = tid; i < C::P_BYTES / 16; i += kTcThreads)
    reinterpret_cast<uint4*>(sm + C::OFF_P)[i] = make_uint4(0u, 0u, 0u, 0u);
  tc_phase_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = uint32_t(warp * 32) << 16;
  tc::mbar_wait(bar_w, 0);
  uint32_t ph = 0;

  const uint64_t d_w = tc::make_smem_desc(tc::smem_u32(sm + C::OFF_W), C::W_LBO, 128);
  const uint64_t d_xq = tc::make_smem_desc(tc::smem_u32(sm + C::OFF_XQ), kOpLbo, kOpSbo);
  const uint64_t d_xkv = tc::make_smem_desc(tc::smem_u32(sm + C::OFF_XKV), kOpLbo, kOpSbo);
  const uint64_t d_q = tc::make_smem_desc(tc::smem_u32(sm + C::OFF_Q), kOpLbo, kOpSbo);
  const uint64_t d_k = tc::make_smem_desc(tc::smem_u32(sm + C::OFF_K), kOpLbo, kOpSbo);
  const uint64_t d_p = tc::make_smem_desc(tc::smem_u32(sm + C::OFF_P), kOpLbo, kOpSbo);
  // V as the MN-major B operand of P.V: 8-column groups of V are kOpLbo apart (stride), 8-row groups 128 apart (leading)
  const uint64_t d_v = tc::make_smem_desc(tc::smem_u32(sm + C::OFF_V), kOpSbo, kOpLbo);

  const int r = tid;                       // my row of the tile
  const int g = r >> 4;                    // my sample of the tile

  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int b0 = tile * kTcG;
    // ---- stage inputs ----------------------------------------------------------------------------------
    tc_stage_rows<C::DIN>(sm + C::OFF_XQ, a.xq, b0, a.batch, a.normalise != 0, inv_norm_s, a.norms);
    if (!C::SELF) {
      tc_stage_rows<C::DIN>(sm + C::OFF_XKV, a.xkv, b0, a.batch, false, inv_norm_s, nullptr);
    }
    tc_phase_sync();
    // ---- projection MMAs ----------------------------------------------------------------------------------
    if (tid == 0) {
      if constexpr (C::SELF) {
        constexpr int NH = C::NALL / 2;    // 352 = 2 x 176 (N <= 256 per instruction)
        static_assert(NH % 16 == 0, "N split");
        constexpr uint32_t idesc = tc::make_idesc_bf16(128, NH, 0, 0);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int ks = 0; ks < C::DIN / 16; ++ks)
            tc::umma_bf16(tmem + C::TM_PROJ + h * NH, tc::desc_advance(d_xq, ks * 2 * kOpLbo),
                          tc::desc_advance(d_w, ks * 2 * C::W_LBO + h * (NH / 8) * 128), idesc, ks > 0);
      } else {
        constexpr uint32_t idq = tc::make_idesc_bf16(128, C::DKQ, 0, 0);
        constexpr uint32_t idkv = tc::make_idesc_bf16(128, C::NKV, 0, 0);
#pragma unroll
        for (int ks = 0; ks < C::DIN / 16; ++ks)
          tc::umma_bf16(tmem + C::TM_PROJ, tc::desc_advance(d_xq, ks * 2 * kOpLbo),
                        tc::desc_advance(d_w, ks * 2 * C::W_LBO), idq, ks > 0);
#pragma unroll
        for (int ks = 0; ks < C::DIN / 16; ++ks)
          tc::umma_bf16(tmem + C::TM_PROJ + C::DKQ, tc::desc_advance(d_xkv, ks * 2 * kOpLbo),
                        tc::desc_advance(d_w, ks * 2 * C::W_LBO + (C::DKQ / 8) * 128), idkv, ks > 0);
      }
      tc::umma_commit(bar_mma);
    }
    tc::mbar_wait(bar_mma, ph); ph ^= 1;
    tc::tc_fence_after_sync();
    // ---- projection epilogue: scale by 1/||x||, add bias, bf16 -> Q / K / V operands ---------------------------
    {
      const float inv_q = inv_norm_s[g];
      const float inv_kv = C::SELF ? inv_q : 1.0f;
#pragma unroll 1
      for (int c0 = 0; c0 < C::NALL; c0 += 16) {
        float v[16];
        tc::tmem_ld16(tmem + lane_base + C::TM_PROJ + c0, v);
        const float inv = c0 < C::DKQ ? inv_q : inv_kv;
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = fmaf(v[e], inv, bias_s[c0 + e]);
        uint8_t* op;
        int cl;
        if (c0 < C::DKQ) { op = sm + C::OFF_Q; cl = c0; }
        else if (c0 < 2 * C::DKQ) { op = sm + C::OFF_K; cl = c0 - C::DKQ; }
        else { op = sm + C::OFF_V; cl = c0 - 2 * C::DKQ; }
        const float lo[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]};
        const float hi[8] = {v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]};
        *reinterpret_cast<uint4*>(op + tc::core_off(r, cl / 8, kOpLbo, kOpSbo)) = tc::pack_bf16x8(lo);
        *reinterpret_cast<uint4*>(op + tc::core_off(r, cl / 8 + 1, kOpLbo, kOpSbo)) = tc::pack_bf16x8(hi);
      }
    }
    tc_phase_sync();
    // ---- scores: S[128 x 128] = Q K^T (all 8 samples; the diagonal 16x16 blocks are the per-sample scores) ----
    if (tid == 0) {
      constexpr uint32_t ids = tc::make_idesc_bf16(128, 128, 0, 0);
#pragma unroll
      for (int ks = 0; ks < C::DKQ / 16; ++ks)
        tc::umma_bf16(tmem + C::TM_S, tc::desc_advance(d_q, ks * 2 * kOpLbo), tc::desc_advance(d_k, ks * 2 * kOpLbo),
                      ids, ks > 0);
      tc::umma_commit(bar_mma);
    }
    tc::mbar_wait(bar_mma, ph); ph ^= 1;
    tc::tc_fence_after_sync();
    // ---- softmax over my row's 16 scores (multimodal_model.py:58-60 / :89-98), weights -> bf16 operand --------
    {
      // a warp's 32 rows belong to two samples: load the warp's 32 score columns (tcgen05.ld takes one
      // warp-uniform address) and keep my sample's half
      float s32[32], s[16];
      tc::tmem_ld32(tmem + lane_base + C::TM_S + warp * 32, s32);
#pragma unroll
      for (int j = 0; j < 16; ++j) s[j] = (lane & 16) ? s32[16 + j] : s32[j];
      const float sq = sqrtf(float(C::DKQ));
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < 16; ++j) { s[j] = s[j] / sq; m = fmaxf(m, s[j]); }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) { s[j] = expf(s[j] - m); sum += s[j]; }
#pragma unroll
      for (int j = 0; j < 16; ++j) s[j] = attn_weight(s[j] / sum, reverse);
      const float lo[8] = {s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7]};
      const float hi[8] = {s[8], s[9], s[10], s[11], s[12], s[13], s[14], s[15]};
      *reinterpret_cast<uint4*>(sm + C::OFF_P + tc::core_off(r, 2 * g, kOpLbo, kOpSbo)) = tc::pack_bf16x8(lo);
      *reinterpret_cast<uint4*>(sm + C::OFF_P + tc::core_off(r, 2 * g + 1, kOpLbo, kOpSbo)) = tc::pack_bf16x8(hi);
    }
    tc_phase_sync();
    // ---- ctx[128 x DV] = P V (P block diagonal) ------------------------------------------------------------
    if (tid == 0) {
      constexpr uint32_t idc = tc::make_idesc_bf16(128, C::DV, 0, 1);
#pragma unroll
      for (int ks = 0; ks < kTcRows / 16; ++ks)
        tc::umma_bf16(tmem + C::TM_PROJ, tc::desc_advance(d_p, ks * 2 * kOpLbo), tc::desc_advance(d_v, ks * 2 * kOpSbo),
                      idc, ks > 0);
      tc::umma_commit(bar_mma);
    }
    tc::mbar_wait(bar_mma, ph); ph ^= 1;
    tc::tc_fence_after_sync();
    // ---- LayerNorm + ReLU on my row, store (multimodal_model.py:65-66 / :105-106) ------------------------------
    {
      float x[C::DV];
#pragma unroll
      for (int c0 = 0; c0 < C::DV; c0 += 16) {
        float v[16];
        tc::tmem_ld16(tmem + lane_base + C::TM_PROJ + c0, v);
#pragma unroll
        for (int e = 0; e < 16; ++e) x[c0 + e] = v[e];
      }
      float mean = 0.f;
#pragma unroll
      for (int c = 0; c < C::DV; ++c) mean += x[c];
      mean /= float(C::DV);
      float var = 0.f;
#pragma unroll
      for (int c = 0; c < C::DV; ++c) { const float d = x[c] - mean; var = fmaf(d, d, var); }
      const float rstd = rsqrtf(var / float(C::DV) + kLnEps);
      const int b = b0 + g;
      if (b < a.batch) {
        float4* dst = reinterpret_cast<float4*>(a.out + (size_t(b) * kL + (r & 15)) * C::DV);
#pragma unroll
        for (int c = 0; c < C::DV; c += 4) {
          float4 o;
          o.x = fmaxf((x[c + 0] - mean) * rstd * ln_s[c + 0] + ln_s[C::DV + c + 0], 0.f);
          o.y = fmaxf((x[c + 1] - mean) * rstd * ln_s[c + 1] + ln_s[C::DV + c + 1], 0.f);
          o.z = fmaxf((x[c + 2] - mean) * rstd * ln_s[c + 2] + ln_s[C::DV + c + 2], 0.f);
          o.w = fmaxf((x[c + 3] - mean) * rstd * ln_s[c + 3] + ln_s[C::DV + c + 3], 0.f);
          dst[c / 4] = o;
        }
      }
    }
    tc::tc_fence_before_sync();
    __syncthreads();    // TMEM and the operand buffers are reused by the next tile
    tc::tc_fence_after_sync();
  }
  if (warp == 0) tc::tmem_dealloc(tmem, C::TM_COLS);
}

}  // namespace mmrca

// bf16 tensor-core pipeline of the MM-RCA head (MMRCA_COMPUTE_BF16): forward kernels.
//
// Reference path: CVPR_code/multimodal_model.py:661-728 (MM_RCA.forward after the backbones), with
// SelfAttention (:39-68) and ReverseCrossAttention (:71-108).  Every contraction is a tcgen05.mma with fp32
// accumulation in TMEM; every per-row step (bias, softmax, LayerNorm, ReLU) runs on one thread per row.
//
// Algebra used here (exact in real arithmetic, it only changes where bf16 rounding happens):
//   scores: Q K^T = (Xq Wq^T + bq)(Xkv Wk^T + bk)^T.  The terms that are constant along a score ROW cancel in the
//   softmax, so   softmax(Q K^T / sqrt(d)) = softmax(Z Xkv^T)   with   Z = Xq M + u,
//   M = Wq^T Wk / sqrt(d)  [d_in x d_in],   u = Wk^T bq / sqrt(d)  [d_in].
//   d_in (48/80/96) is smaller than 2*d_kq (256/128), so Z replaces both Q and K: fewer projection columns,
//   a shorter score contraction, fewer accumulator read-backs.  W_key.bias drops out exactly (its gradient is
//   analytically zero in the reference too).
//   biases ride in the GEMM: every activation operand carries a constant-one column at k = d_in (and zeros up
//   to d_in+16), the weight blobs carry u / b_value in that row.
//   classifier (:719-726), cross-attention sources T_I / I_T: logits[b][c] = sum_{chunk r, j} F[(b,r)][j]
//   Wf[c][off + r*w + j] is one N=64 MMA  D[(b,r)][(r',c)] = sum_j F[(b,r)][j] Wf[c][off + r'*w + j];  the thread
//   that owns row (b,r) keeps columns (r,0..3) and a 16-lane shuffle sums the sample's chunks.
//   classifier, feature sources (normalised image / text features, 2048 of the 3584 concat columns): plain fp32
//   FMAs on the registers that already hold the sample while it is normalised (8 k FMA per sample), so the
//   part of the logits that decides the argmax carries no bf16 rounding at all.
//   dropout (:719): a counter-based keep mask regenerated wherever the concat is touched (mmrca_dropout.cuh).
//
// Tiles: 8 samples = 128 rows (row = 16*sample + chunk).  MMAs whose rows are tile rows and whose result is only
// converted (projections, classifier) use M=128: TMEM lane = row.  The attention core (scores, P V) uses two M=64
// MMAs per tile (samples 0-3 / 4-7): only the 4x4 sample blocks of a half are computed instead of 8x8, the
// score block of a row sits at a warp-uniform column, and P is 2 x [64 x 64] instead of [128 x 128].  The two
// M=64 accumulators interleave in TMEM lanes: warp q of a warpgroup holds rows 16q..16q+15 of half 0 in lanes
// 0-15 and of half 1 in lanes 16-31 ("s-mapping"), while an M=128 accumulator has row 32q+lane ("p-mapping").
//
// A CTA has two warpgroups, each an independent pipeline on its own tile with its own shared-memory operand
// buffers, TMEM columns and mbarrier: one warpgroup's read-backs overlap the other's MMAs.  Weights are packed
// once per step (prep_kernel) into bf16 blobs in the canonical no-swizzle core-matrix layout and land in shared
// memory through the bulk-copy engine (TMA).  Activations between the SA and CA kernels travel as bf16 operand
// IMAGES (the exact bytes the next kernel's MMA descriptor wants), written coalesced and reloaded by TMA.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mmrca_attn_fp32.cuh"
#include "mmrca_dropout.cuh"
#include "mmrca_tc.cuh"

namespace mmrca {
namespace htc {

using namespace tc;

constexpr int kWgThreads = 128;
constexpr int kCtaThreads = 256;
constexpr uint32_t kCS = 128 * 16 + 16;   // bytes between 8-column groups of a 128-row operand (+16: bank spread)
constexpr uint32_t kRS = 128;             // bytes between 8-row groups
constexpr uint32_t kPCS = 64 * 16 + 16;   // same for one 64-row half of P / dS
constexpr uint32_t kPHalf = 8 * kPCS;     // one [64 x 64] half
constexpr int kDV_SA = 96, kDV_CA = 48, kDKQ_SA = 128, kDKQ_CA = 64;
constexpr int kNCls = 64;                 // 16 chunks x 4 classes
constexpr int kClasses = 4;

__host__ __device__ constexpr uint32_t op_bytes(int cols) { return uint32_t(cols / 8) * kCS; }
__host__ __device__ constexpr uint32_t blob_bytes(int n, int k) { return uint32_t(k / 8) * uint32_t(n) * 16u; }
__host__ __device__ constexpr uint32_t al128(uint32_t v) { return (v + 127u) & ~127u; }

__device__ __forceinline__ uint32_t row_off(int r) { return uint32_t(r >> 3) * kRS + uint32_t(r & 7) * 16u; }

// ---------------------------------------------------------------------------------------------------------------
// prep: weights -> bf16 blobs, zero the step's accumulators.
// blob of a B operand Bm[n][k] (K-major): byte(n, k) = (k/8) * (N*16) + n*16 + (k%8)*2
// ---------------------------------------------------------------------------------------------------------------
struct PrepBlock {
  const float* wq; const float* bq; const float* wk; const float* wv; const float* bv;
  void* bz;      // [n = d_in][k = d_in + 16]:  M^T, row k = d_in holds u
  void* bvb;     // [n = d_v ][k = d_in + 16]:  W_value, row k = d_in holds b_value  (bf16 "hi" part)
  void* bvb_lo;  // same shape: bf16(W_value - hi).  hi + lo carries 16 mantissa bits: the rounding of W_value to ONE
                 // bf16 is a systematic (sample-independent) perturbation that does not average out over the batch and
                 // dominated the gradient error of this pipeline (tools/emulate_bf16.py)
  int din, dkq, dv;
};
// classifier source: columns [off, off + 16*w) of the concat; bc / bc_lo: hi / lo bf16 parts of the weight slice
struct PrepSrc { void* bc; void* bc_lo; int off, w; };
struct PrepArgs {
  PrepBlock blk[4];
  PrepSrc src[4];
  int nsrc;
  const float* wf; int D;                    // classifier [4][D]
  float* zero0; int nzero0;                  // fp32 region to clear (step accumulators)
  float* zero1;                              // one more float to clear (the loss accumulator), or null
  float* zero2; int nzero2;                  // MMRCA_FLAG_ZERO_GRADS: the caller's contiguous gradient bucket (16-byte aligned)
};

constexpr int kPrepZTasks = (80 + 16) / 8 + (48 + 16) / 8 + 2 * (96 + 16) / 8;   // one CTA per (block, 8-column group kc)
constexpr int kPrepCtas = kPrepZTasks + 16;

constexpr uint32_t kPrepSmemBytes = (128 * 80 + 128 * 8) * 4;     // W_key of the widest block + one 8-column slab of W_query

__device__ __forceinline__ void prep_body(const PrepArgs& a, int cta, int nctas, float* smf) {
  // (a) Z blobs.  CTA = (block, kc); thread = k'.  M[k][k'] = scale sum_n Wq[n][k] Wk[n][k'] for the 8 k of the group.
  //     W_key and the W_query slab are staged in shared memory first, every load of the CTA in flight at once: the
  //     weights come from HBM (the features have swept the L2 since the last step) and a dependent chain of d_kq
  //     global loads per thread would cost d_kq memory latencies.
  if (cta < kPrepZTasks) {
    int b = 0, kc = cta;
    for (; b < 4; ++b) {
      const int kcs = (a.blk[b].din + 16) / 8;
      if (kc < kcs) break;
      kc -= kcs;
    }
    const PrepBlock& B = a.blk[b];
    if (!B.bz) return;
    const int din = B.din, dkq = B.dkq;
    float* wk_s = smf;                  // [dkq][din]
    float* wq_s = smf + dkq * din;      // [dkq][8]: columns kc*8 .. kc*8+7 of W_query, or b_query in column 0
    for (int i = threadIdx.x; i < dkq * din / 4; i += blockDim.x)
      reinterpret_cast<float4*>(wk_s)[i] = __ldg(reinterpret_cast<const float4*>(B.wk) + i);
    for (int i = threadIdx.x; i < dkq * 2; i += blockDim.x) {
      const int n = i >> 1, h = i & 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kc * 8 < din) v = __ldg(reinterpret_cast<const float4*>(B.wq + size_t(n) * din + kc * 8 + 4 * h));
      else if (kc * 8 == din && h == 0) v.x = __ldg(B.bq + n);
      reinterpret_cast<float4*>(wq_s)[i] = v;
    }
    __syncthreads();
    const int kp = threadIdx.x;
    if (kp >= din) return;
    const float scale = rsqrtf(float(dkq));
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (kc * 8 <= din) {
#pragma unroll 4
      for (int n = 0; n < dkq; ++n) {
        const float wk = wk_s[n * din + kp];
        const float4 q0 = *reinterpret_cast<const float4*>(wq_s + n * 8);
        const float4 q1 = *reinterpret_cast<const float4*>(wq_s + n * 8 + 4);
        acc[0] = fmaf(q0.x, wk, acc[0]); acc[1] = fmaf(q0.y, wk, acc[1]); acc[2] = fmaf(q0.z, wk, acc[2]);
        acc[3] = fmaf(q0.w, wk, acc[3]); acc[4] = fmaf(q1.x, wk, acc[4]); acc[5] = fmaf(q1.y, wk, acc[5]);
        acc[6] = fmaf(q1.z, wk, acc[6]); acc[7] = fmaf(q1.w, wk, acc[7]);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= scale;
    *reinterpret_cast<uint4*>(static_cast<uint8_t*>(B.bz) + size_t(kc) * (din * 16) + kp * 16) = pack_bf16x8(acc);
    return;
  }
  const int gtid = (cta - kPrepZTasks) * blockDim.x + threadIdx.x, gsz = (nctas - kPrepZTasks) * blockDim.x;
  // (b) V blobs: thread per (n, kc)
  for (int b = 0; b < 4; ++b) {
    const PrepBlock& B = a.blk[b];
    if (!B.bvb) continue;
    const int kcs = (B.din + 16) / 8;
    for (int t = gtid; t < B.dv * kcs; t += gsz) {
      const int n = t / kcs, kc = t - n * kcs;
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (kc * 8 < B.din) {
        const float4 lo = __ldg(reinterpret_cast<const float4*>(B.wv + size_t(n) * B.din + kc * 8));
        const float4 hi = __ldg(reinterpret_cast<const float4*>(B.wv + size_t(n) * B.din + kc * 8 + 4));
        v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
      } else if (kc * 8 == B.din) {
        v[0] = __ldg(B.bv + n);
      }
      float lo[8];
      const uint4 hi = pack_bf16x8_hilo(v, lo);
      *reinterpret_cast<uint4*>(static_cast<uint8_t*>(B.bvb) + size_t(kc) * (B.dv * 16) + n * 16) = hi;
      *reinterpret_cast<uint4*>(static_cast<uint8_t*>(B.bvb_lo) + size_t(kc) * (B.dv * 16) + n * 16) = pack_bf16x8(lo);
    }
  }
  // (c) classifier blobs: Bc[n = r'*4 + c][k = j] = Wf[c][off + r'*w + j]
  for (int s = 0; s < a.nsrc; ++s) {
    const PrepSrc& S = a.src[s];
    const int kcs = S.w / 8;
    for (int t = gtid; t < kNCls * kcs; t += gsz) {
      const int n = t / kcs, kc = t - n * kcs, rp = n >> 2, c = n & 3;
      const float* p = a.wf + size_t(c) * a.D + S.off + rp * S.w + kc * 8;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __ldg(p + e);
      float lo[8];
      const uint4 hi = pack_bf16x8_hilo(v, lo);
      *reinterpret_cast<uint4*>(static_cast<uint8_t*>(S.bc) + size_t(kc) * (kNCls * 16) + n * 16) = hi;
      *reinterpret_cast<uint4*>(static_cast<uint8_t*>(S.bc_lo) + size_t(kc) * (kNCls * 16) + n * 16) = pack_bf16x8(lo);
    }
  }
  // (d) accumulators
  for (int i = gtid; i < a.nzero0; i += gsz) a.zero0[i] = 0.f;
  if (gtid == 0 && a.zero1) *a.zero1 = 0.f;
  for (int i = gtid; i < a.nzero2 / 4; i += gsz) reinterpret_cast<float4*>(a.zero2)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---------------------------------------------------------------------------------------------------------------
// features: L2-normalise (multimodal_model.py:662-665), emit the bf16 operand images the SA kernels load by TMA,
// the norms, and logits = bias + the classifier's feature-source terms in fp32 (:708-726, dropout :719 included).
// HBM-bound streaming: one warp per sample, persistent CTAs that keep the 4 x 2048 feature columns of the classifier
// weight in shared memory.  A 16-byte item is (chunk row r, 8-column group kc), kc fastest: a warp reads 1 KB of
// contiguous feature floats per instruction and its shared-memory weight reads are conflict-free.
// Image of a tile (8 samples = 128 rows): [KCS + 2 column groups][128 rows][8 bf16], column-group stride kCS;
// group KCS holds the constant-one column that carries the projection biases, group KCS + 1 zeros.
// ---------------------------------------------------------------------------------------------------------------
struct FeatSrc {
  const void* feat;       // [B][16 * din] fp32, or bf16 when FeatArgs::feat_bf16 (MMRCA_FLAG_FEATURES_BF16)
  void* x_tiles;          // [tiles][x_tile_bytes(din)] out
  float* norms;           // [B] out
  int cls_off;            // first concat column of this source
};
struct FeatArgs {
  FeatSrc src[2];         // 0: image (din 80), 1: text (din 48)
  float* logits;          // [B][4] out: bias (+ feature terms when with_features)
  const float* wf; const float* bf;
  int with_features;      // the classifier sees the features (not cross_attention_only)
  DropSpec drop;          // drop.D = concat width
  int batch;
  int feat_bf16;          // the features arrive as bf16 (half the HBM / PCIe bytes); norms and classifier terms stay fp32
};
__host__ __device__ constexpr uint32_t x_tile_bytes(int din) { return op_bytes(din + 16); }
constexpr int kFeatCols = 16 * 80 + 16 * 48;                 // 2048 feature columns of the concat
constexpr uint32_t kFeatSmemBytes = kClasses * kFeatCols * 4 > kPrepSmemBytes ? kClasses * kFeatCols * 4 : kPrepSmemBytes;

template <int DIN>
__device__ __forceinline__ void feat_load(const FeatSrc& S, bool bf16, int b, bool live, int lane, float (&v)[(kL * DIN / 8 + 31) / 32][8]) {
  constexpr int ITEMS = kL * DIN / 8, PER = (ITEMS + 31) / 32;
  if (bf16) {      // one 16-byte load per item
    const uint4* base = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(S.feat) + size_t(b) * (kL * DIN));
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int it = lane + 32 * k;
      uint4 raw = make_uint4(0u, 0u, 0u, 0u);
      if (live && (ITEMS % 32 == 0 || it < ITEMS)) raw = __ldg(base + it);
      v[k][0] = __uint_as_float(raw.x << 16); v[k][1] = __uint_as_float(raw.x & 0xffff0000u);
      v[k][2] = __uint_as_float(raw.y << 16); v[k][3] = __uint_as_float(raw.y & 0xffff0000u);
      v[k][4] = __uint_as_float(raw.z << 16); v[k][5] = __uint_as_float(raw.z & 0xffff0000u);
      v[k][6] = __uint_as_float(raw.w << 16); v[k][7] = __uint_as_float(raw.w & 0xffff0000u);
    }
    return;
  }
  const float* base = static_cast<const float*>(S.feat) + size_t(b) * (kL * DIN);
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int it = lane + 32 * k;
    float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
    if (live && (ITEMS % 32 == 0 || it < ITEMS)) {
      lo = __ldg(reinterpret_cast<const float4*>(base + it * 8));
      hi = __ldg(reinterpret_cast<const float4*>(base + it * 8 + 4));
    }
    v[k][0] = lo.x; v[k][1] = lo.y; v[k][2] = lo.z; v[k][3] = lo.w; v[k][4] = hi.x; v[k][5] = hi.y; v[k][6] = hi.z; v[k][7] = hi.w;
  }
}

// ws: this source's [4][16 * DIN] slice of the classifier weight in shared memory
template <int DIN>
__device__ __forceinline__ void feat_emit(const FeatArgs& a, const FeatSrc& S, const float* ws, int b, bool live, int lane,
                                          const float (&v)[(kL * DIN / 8 + 31) / 32][8], float (&acc)[kClasses]) {
  constexpr int KCS = DIN / 8, ITEMS = kL * KCS, PER = (ITEMS + 31) / 32;
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < PER; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) ss = fmaf(v[k][e], v[k][e], ss);
  const float nrm = sqrtf(warp_sum(ss));
  const float inv = live ? 1.0f / nrm : 0.f;        // no epsilon, like the reference
  if (lane == 0 && live) S.norms[b] = nrm;
  uint8_t* tile = static_cast<uint8_t*>(S.x_tiles) + size_t(b >> 3) * x_tile_bytes(DIN);
  const int r0 = (b & 7) * kL;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int it = lane + 32 * k;
    if (ITEMS % 32 != 0 && it >= ITEMS) continue;
    const int row = it / KCS, kc = it - row * KCS;
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = v[k][e] * inv;
    *reinterpret_cast<uint4*>(tile + uint32_t(kc) * kCS + row_off(r0 + row)) = pack_bf16x8(o);
    if (a.with_features && live) {
      if (a.drop.thresh) drop_apply8(a.drop, uint32_t(b), uint32_t(S.cls_off + it * 8), o);
#pragma unroll
      for (int cc = 0; cc < kClasses; ++cc) {
        // (staged as two planes per class - first / second four columns of every item - so that a warp's 16-byte reads are
        //  contiguous: with the natural order the 32-byte lane stride made every read a two-way bank conflict)
        const float4 w0 = reinterpret_cast<const float4*>(ws)[cc * (2 * ITEMS) + it];
        const float4 w1 = reinterpret_cast<const float4*>(ws)[cc * (2 * ITEMS) + ITEMS + it];
        acc[cc] = fmaf(o[0], w0.x, fmaf(o[1], w0.y, fmaf(o[2], w0.z, fmaf(o[3], w0.w, acc[cc]))));
        acc[cc] = fmaf(o[4], w1.x, fmaf(o[5], w1.y, fmaf(o[6], w1.z, fmaf(o[7], w1.w, acc[cc]))));
      }
    }
  }
  // the bias column groups of my 16 rows: lanes 0-15 the ones column, lanes 16-31 the zero group
  *reinterpret_cast<uint4*>(tile + uint32_t(KCS + (lane >> 4)) * kCS + row_off(r0 + (lane & 15))) =
      make_uint4(lane < 16 ? 0x00003F80u : 0u, 0u, 0u, 0u);
}

__global__ void __launch_bounds__(256) prep_feat_kernel(const PrepArgs pa, const FeatArgs fa, int prep_ctas) {
  extern __shared__ __align__(16) float wsm_f[];      // [4][1280] image columns, then [4][768] text columns
  if (int(blockIdx.x) < prep_ctas) { prep_body(pa, blockIdx.x, prep_ctas, wsm_f); return; }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = ((fa.batch + 7) / 8) * 8;      // whole tiles: the padding samples of the last tile become zero rows
  const int stride = (int(gridDim.x) - prep_ctas) * 8;
  int b = (int(blockIdx.x) - prep_ctas) * 8 + warp;
  // the first sample's features are requested before anything else: they travel while the classifier slice is staged
  float vi[5][8], vt[3][8];
  if (b < nb) {
    feat_load<80>(fa.src[0], fa.feat_bf16 != 0, b, b < fa.batch, lane, vi);
    feat_load<48>(fa.src[1], fa.feat_bf16 != 0, b, b < fa.batch, lane, vt);
  }
  if (fa.with_features) {
    constexpr int N4 = kClasses * kFeatCols / 4 / 256;      // 8 float4 per thread, requested four at a time
    static_assert(N4 * 256 * 4 == kClasses * kFeatCols && N4 % 4 == 0, "classifier slice staging");
#pragma unroll
    for (int u0 = 0; u0 < N4; u0 += 4) {
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int f = 4 * (int(threadIdx.x) + 256 * (u0 + u));
        int cc, j, off;
        if (f < kClasses * 1280) { cc = f / 1280; j = f - cc * 1280; off = fa.src[0].cls_off; }
        else { const int g = f - kClasses * 1280; cc = g / 768; j = g - cc * 768; off = fa.src[1].cls_off; }
        t[u] = __ldg(reinterpret_cast<const float4*>(fa.wf + size_t(cc) * fa.drop.D + off + j));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {      // float4 (class cc, item j / 8, half (j / 4) & 1) -> plane `half` of the class
        const int f = 4 * (int(threadIdx.x) + 256 * (u0 + u));
        const bool im = f < kClasses * 1280;
        const int g = im ? f : f - kClasses * 1280, w = im ? 1280 : 768, items = w / 8;
        const int cc = g / w, j = g - cc * w;
        reinterpret_cast<float4*>(wsm_f)[(im ? 0 : kClasses * 1280 / 4) + cc * (2 * items) + ((j >> 2) & 1) * items + (j >> 3)] = t[u];
      }
    }
  }
  __syncthreads();
  while (b < nb) {
    const bool live = b < fa.batch;
    float acc[kClasses] = {0.f, 0.f, 0.f, 0.f};
    feat_emit<80>(fa, fa.src[0], wsm_f, b, live, lane, vi, acc);
    feat_emit<48>(fa, fa.src[1], wsm_f + kClasses * 1280, b, live, lane, vt, acc);
    const int bn = b + stride;
    if (bn < nb) {      // the next sample's features travel while the classifier terms are reduced and written
      feat_load<80>(fa.src[0], fa.feat_bf16 != 0, bn, bn < fa.batch, lane, vi);
      feat_load<48>(fa.src[1], fa.feat_bf16 != 0, bn, bn < fa.batch, lane, vt);
    }
#pragma unroll
    for (int cc = 0; cc < kClasses; ++cc) acc[cc] = warp_sum(acc[cc]);
    if (lane < kClasses && live) {
      float r = acc[0];
      r = lane == 1 ? acc[1] : r; r = lane == 2 ? acc[2] : r; r = lane == 3 ? acc[3] : r;
      fa.logits[size_t(b) * kClasses + lane] = __ldg(fa.bf + lane) + r;
    }
    b = bn;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// shared pieces of the tile kernels
// ---------------------------------------------------------------------------------------------------------------
struct WgCtx {
  int wt, q, lane;        // thread in warpgroup, warp in warpgroup, lane
  int wg;                 // warpgroup in CTA
  uint32_t tmem;          // TMEM base of this warpgroup (lane 0, first column)
  uint32_t lane_base;     // (32*q) << 16
  uint64_t* bar;          // MMA-completion mbarrier of this warpgroup
  uint32_t ph;            // its phase
  // p-mapping: row of an M=128 accumulator
  int rp;
  // s-mapping: row of the interleaved M=64 accumulators
  int h, i, rs;
};

__device__ __forceinline__ WgCtx make_ctx(int wg_threads_base, uint32_t tmem, uint64_t* bar) {
  WgCtx c;
  c.wt = threadIdx.x - wg_threads_base;
  c.q = c.wt >> 5; c.lane = c.wt & 31; c.wg = threadIdx.x >> 7;
  c.tmem = tmem; c.lane_base = uint32_t(32 * c.q) << 16;
  c.bar = bar; c.ph = 0;
  c.rp = c.wt;
  c.h = c.lane >> 4; c.i = c.lane & 15; c.rs = 64 * c.h + 16 * c.q + c.i;
  return c;
}

// operands written by this warpgroup -> visible to the tensor core; warpgroup barrier
__device__ __forceinline__ void wg_sync_for_mma(const WgCtx& c) {
  fence_proxy_async();
  tc_fence_before_sync();
  named_bar_sync(1 + c.wg, kWgThreads);
  tc_fence_after_sync();
}
__device__ __forceinline__ void wg_wait_mma(WgCtx& c) {
  mbar_wait(c.bar, c.ph);
  c.ph ^= 1;
  tc_fence_after_sync();
}

// K-major operand (rows x K cols, chunk stride cs): D (+)= A B^T over `ksteps` K=16 steps
__device__ __forceinline__ void mma_steps(uint32_t d, uint64_t ad, uint32_t a_step, uint64_t bd, uint32_t b_step,
                                          uint32_t idesc, int ksteps, bool acc) {
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)        // every chain here has at most 8 K-steps: unrolled, descriptors by immediate adds
    if (ks < ksteps)
      umma_bf16(d, desc_advance(ad, ks * a_step), desc_advance(bd, ks * b_step), idesc, (acc || ks > 0) ? 1u : 0u);
}

// TMEM columns [c0, c0+16) of my row -> bf16 -> two 16-byte chunks of a 128-row operand
__device__ __forceinline__ void store_chunks16(uint8_t* op, int row, int col0, const uint32_t (&r)[16]) {
  float lo[8], hi[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { lo[e] = __uint_as_float(r[e]); hi[e] = __uint_as_float(r[8 + e]); }
  *reinterpret_cast<uint4*>(op + uint32_t(col0 >> 3) * kCS + row_off(row)) = pack_bf16x8(lo);
  *reinterpret_cast<uint4*>(op + uint32_t((col0 >> 3) + 1) * kCS + row_off(row)) = pack_bf16x8(hi);
}

// accumulator columns [col, col+NC) of my row (p-mapping) -> bf16 operand
template <int NC>
__device__ __forceinline__ void acc_to_operand(const WgCtx& c, uint32_t col, uint8_t* op) {
  static_assert(NC % 16 == 0, "16-column steps");
#pragma unroll
  for (int c0 = 0; c0 < NC; c0 += 32) {
    uint32_t r0[16], r1[16];
    tmem_ld16_nw(c.tmem + c.lane_base + col + c0, r0);
    if (c0 + 16 < NC) tmem_ld16_nw(c.tmem + c.lane_base + col + c0 + 16, r1);
    tmem_wait_ld();
    store_chunks16(op, c.rp, c0, r0);
    if (c0 + 16 < NC) store_chunks16(op, c.rp, c0 + 16, r1);
  }
}

// classifier read-back (p-mapping): D[(b,r)][(r',c)], keep r' == r, sum the sample's 16 chunks, add to logits
__device__ __forceinline__ void cls_readback(const WgCtx& c, uint32_t col, float* logits, int b, bool valid) {
  uint32_t v[4][16];
#pragma unroll
  for (int qq = 0; qq < 4; ++qq) tmem_ld16_nw(c.tmem + c.lane_base + col + 16 * qq, v[qq]);
  tmem_wait_ld();
  // my row's four columns (r, 0..3) = column 4r + cc of the 64: picked with bit masks (a chain of ?: on lane-dependent
  // conditions compiles to divergent branches, a thousand cycles of reconvergence per tile)
  const int r = c.rp & 15, grp = r >> 2, m = r & 3;
  const uint32_t g0 = grp == 0 ? ~0u : 0u, g1 = grp == 1 ? ~0u : 0u, g2 = grp == 2 ? ~0u : 0u, g3 = grp == 3 ? ~0u : 0u;
  const uint32_t m0 = m == 0 ? ~0u : 0u, m1 = m == 1 ? ~0u : 0u, m2 = m == 2 ? ~0u : 0u, m3 = m == 3 ? ~0u : 0u;
  uint32_t g[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) g[e] = (v[0][e] & g0) | (v[1][e] & g1) | (v[2][e] & g2) | (v[3][e] & g3);
  float s[4];
#pragma unroll
  for (int cc = 0; cc < 4; ++cc)
    s[cc] = __uint_as_float((g[cc] & m0) | (g[4 + cc] & m1) | (g[8 + cc] & m2) | (g[12 + cc] & m3));
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s[cc] += __shfl_xor_sync(0xffffffffu, s[cc], o);
  }
  if (r == 0 && valid) {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) red_add(logits + size_t(b) * kClasses + cc, s[cc]);
  }
}

// softmax over my row's 16 scores (s-mapping), optional reverse weights (multimodal_model.py:58-60, :89-98);
// p[] returns the weights that multiply V
__device__ __forceinline__ void softmax16(const WgCtx& c, uint32_t col_s, bool reverse, float (&p)[16]) {
  uint32_t r[16];
  tmem_ld16_nw(c.tmem + c.lane_base + col_s + 16 * c.q, r);
  tmem_wait_ld();
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < 16; ++j) { p[j] = __uint_as_float(r[j]); m = fmaxf(m, p[j]); }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) { p[j] = __expf(p[j] - m); sum += p[j]; }
  const float inv = 1.0f / sum;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float a = p[j] * inv;
    p[j] = reverse ? (1.0f - a) * (1.0f / float(kL - 1)) : a;
  }
}

// my row of one [64 x 64] half of P / dS: the sample's 16 columns are chunks 2q, 2q+1, the rest is zero
__device__ __forceinline__ void store_p_row(const WgCtx& c, uint8_t* pbuf, const float (&p)[16], bool zero_rest) {
  uint8_t* base = pbuf + c.h * kPHalf + row_off(16 * c.q + c.i);
  const float lo[8] = {p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]};
  const float hi[8] = {p[8], p[9], p[10], p[11], p[12], p[13], p[14], p[15]};
#pragma unroll
  for (int kc = 0; kc < 8; ++kc) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (kc == 2 * c.q) v = pack_bf16x8(lo);
    else if (kc == 2 * c.q + 1) v = pack_bf16x8(hi);
    else if (!zero_rest) continue;
    *reinterpret_cast<uint4*>(base + kc * kPCS) = v;
  }
}

// LayerNorm statistics of my context row (s-mapping accumulator columns [col, col+DV))
template <int DV>
__device__ __forceinline__ void ln_stats_tmem(const WgCtx& c, uint32_t col, float& mean, float& rstd) {
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int c0 = 0; c0 < DV; c0 += 16) {
    uint32_t r[16];
    tmem_ld16_nw(c.tmem + c.lane_base + col + c0, r);
    tmem_wait_ld();
#pragma unroll
    for (int e = 0; e < 16; ++e) { const float x = __uint_as_float(r[e]); s += x; ss = fmaf(x, x, ss); }
  }
  mean = s * (1.0f / float(DV));
  const float var = fmaxf(ss * (1.0f / float(DV)) - mean * mean, 0.f);
  rstd = rsqrtf(var + kLnEps);
}

// ---------------------------------------------------------------------------------------------------------------
// self-attention forward: features -> SA output image (bf16 operand layout) + the classifier's feature term
// ---------------------------------------------------------------------------------------------------------------
template <int DIN_>
struct SaCfg {
  static constexpr int DIN = DIN_, KE = DIN_ + 16, DV = kDV_SA;
  static constexpr uint32_t BZ_LBO = DIN * 16, BZ_BYTES = blob_bytes(DIN, KE);
  static constexpr uint32_t BV_LBO = DV * 16, BV_BYTES = blob_bytes(DV, KE);
  static constexpr uint32_t W_BYTES = BZ_BYTES + 2 * BV_BYTES;      // bz | bv (hi) | bv (lo)
  // TMEM columns (relative to the warpgroup base)
  static constexpr uint32_t COL_Z = 0, COL_V = DIN, COL_S = 0, COL_C = 64;
  static_assert(DIN + DV <= 256, "warpgroup TMEM budget");
};

constexpr uint32_t kSaTileBytes = op_bytes(kDV_SA);   // one SA output image: [128 x 96] bf16 = 12 chunk columns

struct SaRole {
  const void* x_tiles;    // [tiles][x_tile_bytes(DIN)]: normalised bf16 operand images (prep_feat_kernel)
  const float* ln_g; const float* ln_b;
  const void* blobs;      // bz | bv, contiguous
  void* out_tiles;        // [tiles][kSaTileBytes]
  void* v_tiles;          // training: V operand images [tiles][op_bytes(96)] for the backward (null: not kept)
  void* p_tiles;          // training: attention-weight images [tiles][2 * kPHalf]
  float2* ln_stats;       // training: (mean, rstd) of every context row, [tiles][128]
};
struct SaFwdArgs {
  SaRole role[2];         // 0: image (DIN 80), 1: text (DIN 48)
  int batch;
  long long* dbg;         // development: per-phase clock64 stamps of warpgroup 0 of CTA 0 (null in production)
};

// warpgroup buffers of the SA forward (sized for the wider role)
struct SaFwdSmem {
  static constexpr uint32_t X = 0;                                   // [128 x (80+16)]
  static constexpr uint32_t ZP = al128(X + op_bytes(96));             // Z [128 x 80]; later P (2 x [64 x 64])
  static constexpr uint32_t V = al128(ZP + op_bytes(80));             // [128 x 96]
  static constexpr uint32_t BYTES = al128(V + op_bytes(96));
  static_assert(op_bytes(80) >= 2 * kPHalf, "P aliases Z");
};
struct SaFwdLayout {
  static constexpr uint32_t W0 = 0;                                               // image blobs
  static constexpr uint32_t W1 = al128(W0 + SaCfg<80>::W_BYTES);                   // text blobs
  static constexpr uint32_t WG0 = al128(W1 + SaCfg<48>::W_BYTES);
  static constexpr uint32_t WG1 = WG0 + SaFwdSmem::BYTES;
  static constexpr uint32_t LN = WG1 + SaFwdSmem::BYTES;                            // 2 roles x (gamma, beta) x 96 fp32
  static constexpr uint32_t BAR = al128(LN + 2 * 2 * 96 * 4);                       // mbarriers + tmem slot
  static constexpr uint32_t BYTES = BAR + 128;
  static_assert(BYTES <= 232448, "SA forward does not fit shared memory");
};

// TMA of the X image of (tile, role) into this warpgroup's operand buffer, completion on bar_ld
__device__ __forceinline__ void sa_issue_x_load(const SaFwdArgs& a, int role, int tile, uint8_t* xop, uint64_t* bar_ld) {
  const uint32_t bytes = role == 0 ? x_tile_bytes(80) : x_tile_bytes(48);
  mbar_arrive_expect_tx(bar_ld, bytes);
  bulk_g2s(xop, static_cast<const uint8_t*>(a.role[role].x_tiles) + size_t(tile) * bytes, bytes, bar_ld);
}

template <class C>
__device__ __forceinline__ void sa_fwd_tile(WgCtx& c, const SaFwdArgs& a, const SaRole& R, uint8_t* wsm, uint8_t* bsm,
                                            const float* ln_s, int tile, uint64_t* bar_ld, uint32_t& ph_ld,
                                            int next_tile, int next_role, int stamp_n) {
#define FSTAMP(i) do { if (a.dbg && c.wt == 0 && c.wg == 0 && blockIdx.x == 0 && stamp_n + (i) < 250) a.dbg[stamp_n + (i)] = clock64(); } while (0)
  FSTAMP(0);
  uint8_t* xop = bsm + SaFwdSmem::X;
  uint8_t* zop = bsm + SaFwdSmem::ZP;
  uint8_t* vop = bsm + SaFwdSmem::V;
  // ---- X image of this tile (issued one tile ahead) ------------------------------------------------------------
  mbar_wait(bar_ld, ph_ld);
  ph_ld ^= 1;
  tc_fence_after_sync();
  FSTAMP(1);
  // ---- Z | V: two MMA chains over the same A operand ---------------------------------------------------------------
  if (c.wt == 0) {
    const uint64_t ax = make_smem_desc(smem_u32(xop), kCS, kRS);
    mma_steps(c.tmem + C::COL_Z, ax, 2 * kCS, make_smem_desc(smem_u32(wsm), C::BZ_LBO, 128), 2 * C::BZ_LBO,
              make_idesc_bf16(128, C::DIN, 0, 0), C::KE / 16, false);
    mma_steps(c.tmem + C::COL_V, ax, 2 * kCS, make_smem_desc(smem_u32(wsm + C::BZ_BYTES), C::BV_LBO, 128),
              2 * C::BV_LBO, make_idesc_bf16(128, C::DV, 0, 0), C::KE / 16, false);
    // V += X (W_value - hi)^T: the value projection sees W_value to 16 mantissa bits
    mma_steps(c.tmem + C::COL_V, ax, 2 * kCS, make_smem_desc(smem_u32(wsm + C::BZ_BYTES + C::BV_BYTES), C::BV_LBO, 128),
              2 * C::BV_LBO, make_idesc_bf16(128, C::DV, 0, 0), C::KE / 16, true);
    umma_commit(c.bar);
  }
  wg_wait_mma(c);
  FSTAMP(2);
  acc_to_operand<C::DIN>(c, C::COL_Z, zop);
  acc_to_operand<C::DV>(c, C::COL_V, vop);
  wg_sync_for_mma(c);
  FSTAMP(3);
  // ---- scores: two M=64 halves, S_h = Z_h X_h^T --------------------------------------------------------------
  if (c.wt == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      mma_steps(c.tmem + (uint32_t(16 * h) << 16) + C::COL_S, make_smem_desc(smem_u32(zop + h * 8 * kRS), kCS, kRS),
                2 * kCS, make_smem_desc(smem_u32(xop + h * 8 * kRS), kCS, kRS), 2 * kCS,
                make_idesc_bf16(64, 64, 0, 0), C::DIN / 16, false);
    umma_commit(c.bar);
    // training: the backward reloads V (and P, below) instead of recomputing the projections and the softmax
    if (R.v_tiles) { bulk_s2g(static_cast<uint8_t*>(R.v_tiles) + size_t(tile) * kSaTileBytes, vop, kSaTileBytes); bulk_commit(); }
  }
  wg_wait_mma(c);
  FSTAMP(4);
  // X is dead (Z, V and the scores have read it): the next tile's image lands while this one finishes
  if (c.wt == 0 && next_tile >= 0) sa_issue_x_load(a, next_role, next_tile, xop, bar_ld);
  {
    float p[16];
    softmax16(c, C::COL_S, false, p);     // SelfAttention has no reverse weights
    store_p_row(c, zop, p, true);         // P reuses Z's bytes: every chunk of my row is rewritten
  }
  wg_sync_for_mma(c);
  FSTAMP(5);
  // ---- context: C_h = P_h V_h -----------------------------------------------------------------------------------
  if (c.wt == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      mma_steps(c.tmem + (uint32_t(16 * h) << 16) + C::COL_C, make_smem_desc(smem_u32(zop + h * kPHalf), kPCS, kRS),
                2 * kPCS, make_smem_desc(smem_u32(vop + h * 8 * kRS), kRS, kCS), 2 * kRS,
                make_idesc_bf16(64, C::DV, 0, 1), 4, false);
    umma_commit(c.bar);
    if (R.p_tiles) { bulk_s2g(static_cast<uint8_t*>(R.p_tiles) + size_t(tile) * (2 * kPHalf), zop, 2 * kPHalf); bulk_commit(); }
  }
  wg_wait_mma(c);
  FSTAMP(6);
  // ---- LayerNorm + ReLU (multimodal_model.py:65-66) -> bf16 image row ------------------------------------------
  {
    float mean, rstd;
    ln_stats_tmem<C::DV>(c, C::COL_C, mean, rstd);
    if (R.ln_stats) R.ln_stats[size_t(tile) * 128 + c.rs] = make_float2(mean, rstd);
    uint8_t* dst = static_cast<uint8_t*>(R.out_tiles) + size_t(tile) * kSaTileBytes + row_off(c.rs);
    const float nmr = -mean * rstd;      // xhat = x rstd - mean rstd: one FMA per element
#pragma unroll
    for (int c0 = 0; c0 < C::DV; c0 += 16) {
      uint32_t r[16];
      tmem_ld16_nw(c.tmem + c.lane_base + C::COL_C + c0, r);
      tmem_wait_ld();
      float o[16];
#pragma unroll
      for (int e = 0; e < 16; ++e)
        o[e] = fmaxf(fmaf(fmaf(__uint_as_float(r[e]), rstd, nmr), ln_s[c0 + e], ln_s[96 + c0 + e]), 0.f);
      const float lo[8] = {o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]};
      const float hi[8] = {o[8], o[9], o[10], o[11], o[12], o[13], o[14], o[15]};
      *reinterpret_cast<uint4*>(dst + uint32_t(c0 >> 3) * kCS) = pack_bf16x8(lo);
      *reinterpret_cast<uint4*>(dst + uint32_t((c0 >> 3) + 1) * kCS) = pack_bf16x8(hi);
    }
  }
  FSTAMP(7);
  if (c.wt == 0 && R.v_tiles) bulk_wait_read();   // the V / P stores have read their buffers
  FSTAMP(8);
  tc_fence_before_sync();
  named_bar_sync(1 + c.wg, kWgThreads);   // TMEM columns and operand buffers are reused by the next tile
  tc_fence_after_sync();
  FSTAMP(9);
#undef FSTAMP
}

__global__ void __launch_bounds__(kCtaThreads, 1) sa_fwd_kernel(const SaFwdArgs a) {
  extern __shared__ __align__(128) uint8_t sm[];
  using L = SaFwdLayout;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::BAR);    // [0]: weights, [1], [2]: MMA, [3], [4]: X loads
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* ln_s = reinterpret_cast<float*>(sm + L::LN);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int wg = tid >> 7;
  const int tiles = (a.batch + 7) / 8;
  uint8_t* bsm = sm + (wg == 0 ? L::WG0 : L::WG1);
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(&bars[0], SaCfg<80>::W_BYTES + SaCfg<48>::W_BYTES);
    bulk_g2s(sm + L::W0, a.role[0].blobs, SaCfg<80>::W_BYTES, &bars[0]);
    bulk_g2s(sm + L::W1, a.role[1].blobs, SaCfg<48>::W_BYTES, &bars[0]);
  }
  __syncthreads();
  // the two warpgroups take the tile's two modalities, alternating: warpgroup wg starts with role wg
  if ((tid & 127) == 0 && int(blockIdx.x) < tiles) sa_issue_x_load(a, wg, blockIdx.x, bsm + SaFwdSmem::X, &bars[3 + wg]);
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < 2 * 96; i += kCtaThreads) {
    const int r = i / 96, k = i - r * 96;
    ln_s[r * 192 + k] = a.role[r].ln_g[k];
    ln_s[r * 192 + 96 + k] = a.role[r].ln_b[k];
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  mbar_wait(&bars[0], 0);
  WgCtx c = make_ctx(wg * kWgThreads, tmem + uint32_t(wg) * 256u, &bars[1 + wg]);
  uint32_t ph_ld = 0;
  int round = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++round) {
    const int role = (wg + round) & 1;
    const int next = tile + int(gridDim.x) < tiles ? tile + int(gridDim.x) : -1;
    if (role == 0) sa_fwd_tile<SaCfg<80>>(c, a, a.role[0], sm + L::W0, bsm, ln_s, tile, &bars[3 + wg], ph_ld, next, 1, 1 + 12 * round);
    else           sa_fwd_tile<SaCfg<48>>(c, a, a.role[1], sm + L::W1, bsm, ln_s + 192, tile, &bars[3 + wg], ph_ld, next, 0, 1 + 12 * round);
  }
  if ((tid & 127) == 0) bulk_wait_all();        // this thread's V / P stores are complete before the CTA retires
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// cross-attention forward (both directions): SA images -> classifier logits
// ---------------------------------------------------------------------------------------------------------------
struct CaCfg {
  static constexpr int DIN = kDV_SA, KE = DIN + 16, DV = kDV_CA;
  static constexpr uint32_t BZ_LBO = DIN * 16, BZ_BYTES = blob_bytes(DIN, KE);
  static constexpr uint32_t BV_LBO = DV * 16, BV_BYTES = blob_bytes(DV, KE);
  static constexpr uint32_t BC_LBO = kNCls * 16, BC_BYTES = blob_bytes(kNCls, DV);
  // blob of a direction: bz | bv | bc | bv (lo) | bc (lo).  The backward recomputes V with hi + lo but reads the classifier
  // slice (dOut = DL Wf^T) and W_value (dXkv) in one bf16: it loads the first W_BYTES_BWD bytes only.
  static constexpr uint32_t OFF_BV = BZ_BYTES, OFF_BC = OFF_BV + BV_BYTES, OFF_BV_LO = OFF_BC + BC_BYTES, OFF_BC_LO = OFF_BV_LO + BV_BYTES;
  static constexpr uint32_t W_BYTES = OFF_BC_LO + BC_BYTES, W_BYTES_BWD = OFF_BC_LO;
  static constexpr uint32_t COL_Z = 0, COL_V = DIN, COL_S = 0, COL_C = 64, COL_CLS = 112;
};

struct CaDir {
  const void* blobs;        // bz | bv | bc
  const float* ln_g; const float* ln_b;
};
struct CaFwdArgs {
  CaDir dir[2];             // 0: cross_attention_1 (q: text SA, kv: image SA); 1: cross_attention_2 (:683-686)
  const void* t_tiles;      // text SA images
  const void* i_tiles;      // image SA images
  float* logits;
  DropSpec drop;            // concat columns of direction d: [d * 768, d * 768 + 768)
  int batch, reverse;
  uint4* row_out[2];        // training: per direction [tiles][128] {keep bits lo, hi, mean, rstd} of every context row
                            // for the backward (dropout draws and LayerNorm statistics are not redone there); null: not kept
  long long* dbg;           // development: per-phase clock64 stamps of warpgroup 0 of CTA 0 (null in production)
};
struct CaFwdSmem {
  static constexpr uint32_t XQ = 0;                                   // [128 x 112]; Z, P, Out reuse its first 96 columns
  static constexpr uint32_t XKV = al128(XQ + op_bytes(112));
  static constexpr uint32_t V = al128(XKV + op_bytes(112));           // [128 x 48]
  static constexpr uint32_t BYTES = al128(V + op_bytes(48));
};
struct CaFwdLayout {      // a CTA works on ONE direction (blockIdx.y): its blob (hi + lo parts) + two warpgroup pipelines
  static constexpr uint32_t W0 = 0;
  static constexpr uint32_t WG0 = al128(W0 + CaCfg::W_BYTES);
  static constexpr uint32_t WG1 = WG0 + CaFwdSmem::BYTES;
  static constexpr uint32_t LN = WG1 + CaFwdSmem::BYTES;              // (gamma, beta) x 48
  static constexpr uint32_t BAR = al128(LN + 2 * 48 * 4);             // weights, 2 x MMA, 2 x tile-load barriers, tmem slot
  static constexpr uint32_t BYTES = BAR + 64;
  static_assert(BYTES <= 232448, "CA forward does not fit shared memory");
};

__device__ __forceinline__ void write_bias_columns(uint8_t* op, int kc0, int row) {
  *reinterpret_cast<uint4*>(op + uint32_t(kc0) * kCS + row_off(row)) = make_uint4(0x00003F80u, 0u, 0u, 0u);
  *reinterpret_cast<uint4*>(op + uint32_t(kc0 + 1) * kCS + row_off(row)) = make_uint4(0u, 0u, 0u, 0u);
}

// Input images of (tile, direction d) through the bulk-copy engine, one mbarrier phase per tile: the key/value image is
// requested first (it arms the barrier for both), the query image second.
__device__ __forceinline__ void ca_issue_kv(const CaFwdArgs& a, int d, int tile, uint8_t* xkv, uint64_t* bar_ld) {
  const uint8_t* ksrc = static_cast<const uint8_t*>(d == 0 ? a.i_tiles : a.t_tiles) + size_t(tile) * kSaTileBytes;
  mbar_arrive_expect_tx(bar_ld, 2 * kSaTileBytes);
  bulk_g2s(xkv, ksrc, kSaTileBytes, bar_ld);
}
__device__ __forceinline__ void ca_issue_q(const CaFwdArgs& a, int d, int tile, uint8_t* xq, uint64_t* bar_ld) {
  const uint8_t* qsrc = static_cast<const uint8_t*>(d == 0 ? a.t_tiles : a.i_tiles) + size_t(tile) * kSaTileBytes;
  bulk_g2s(xq, qsrc, kSaTileBytes, bar_ld);
}

__device__ __forceinline__ void ca_fwd_tile(WgCtx& c, const CaFwdArgs& a, int d, uint8_t* wsm, uint8_t* bsm,
                                            const float* ln_s, uint64_t* bar_ld, uint32_t& ph_ld, int tile, int next_tile, int stamp_n) {
  using C = CaCfg;
#define FSTAMP(i) do { if (a.dbg && c.wt == 0 && c.wg == 0 && blockIdx.x == 0 && stamp_n + (i) < 250) a.dbg[stamp_n + (i)] = clock64(); } while (0)
  FSTAMP(0);
  const int b0 = tile * 8;
  uint8_t* xq = bsm + CaFwdSmem::XQ;
  uint8_t* xkv = bsm + CaFwdSmem::XKV;
  uint8_t* vop = bsm + CaFwdSmem::V;
  // ---- SA images of this tile: requested during the previous tile (ca_issue_kv / ca_issue_q below) -----------------
  write_bias_columns(xq, C::DIN / 8, c.wt);     // (Z / P / Out reuse only the first 96 columns of xq, but the
  write_bias_columns(xkv, C::DIN / 8, c.wt);    //  image load of the next tile must not race with stale readers)
  mbar_wait(bar_ld, ph_ld);
  ph_ld ^= 1;
  wg_sync_for_mma(c);
  FSTAMP(1);
  if (c.wt == 0) {
    mma_steps(c.tmem + C::COL_Z, make_smem_desc(smem_u32(xq), kCS, kRS), 2 * kCS,
              make_smem_desc(smem_u32(wsm), C::BZ_LBO, 128), 2 * C::BZ_LBO, make_idesc_bf16(128, C::DIN, 0, 0),
              C::KE / 16, false);
    mma_steps(c.tmem + C::COL_V, make_smem_desc(smem_u32(xkv), kCS, kRS), 2 * kCS,
              make_smem_desc(smem_u32(wsm + C::OFF_BV), C::BV_LBO, 128), 2 * C::BV_LBO,
              make_idesc_bf16(128, C::DV, 0, 0), C::KE / 16, false);
    mma_steps(c.tmem + C::COL_V, make_smem_desc(smem_u32(xkv), kCS, kRS), 2 * kCS,      // += Xkv (W_value - hi)^T
              make_smem_desc(smem_u32(wsm + C::OFF_BV_LO), C::BV_LBO, 128), 2 * C::BV_LBO,
              make_idesc_bf16(128, C::DV, 0, 0), C::KE / 16, true);
    umma_commit(c.bar);
  }
  wg_wait_mma(c);
  FSTAMP(2);
  acc_to_operand<C::DIN>(c, C::COL_Z, xq);      // Z overwrites the query image (its projection is done)
  acc_to_operand<C::DV>(c, C::COL_V, vop);
  wg_sync_for_mma(c);
  FSTAMP(3);
  if (c.wt == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      mma_steps(c.tmem + (uint32_t(16 * h) << 16) + C::COL_S, make_smem_desc(smem_u32(xq + h * 8 * kRS), kCS, kRS),
                2 * kCS, make_smem_desc(smem_u32(xkv + h * 8 * kRS), kCS, kRS), 2 * kCS,
                make_idesc_bf16(64, 64, 0, 0), C::DIN / 16, false);
    umma_commit(c.bar);
  }
  wg_wait_mma(c);
  FSTAMP(4);
  // the key/value image is dead (V and the scores have read it): the next tile's lands while this one finishes
  if (c.wt == 0 && next_tile >= 0) {
    ca_issue_kv(a, d, next_tile, xkv, bar_ld);
    if (next_tile + 2 * int(gridDim.x) < (a.batch + 7) / 8) {      // and this warpgroup's tile after that: into the L2
      bulk_prefetch_l2(static_cast<const uint8_t*>(a.t_tiles) + size_t(next_tile + 2 * gridDim.x) * kSaTileBytes, kSaTileBytes);
      bulk_prefetch_l2(static_cast<const uint8_t*>(a.i_tiles) + size_t(next_tile + 2 * gridDim.x) * kSaTileBytes, kSaTileBytes);
    }
  }
  {
    float p[16];
    softmax16(c, C::COL_S, a.reverse != 0, p);
    store_p_row(c, xq, p, true);
  }
  wg_sync_for_mma(c);
  FSTAMP(5);
  if (c.wt == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      mma_steps(c.tmem + (uint32_t(16 * h) << 16) + C::COL_C, make_smem_desc(smem_u32(xq + h * kPHalf), kPCS, kRS),
                2 * kPCS, make_smem_desc(smem_u32(vop + h * 8 * kRS), kRS, kCS), 2 * kRS,
                make_idesc_bf16(64, C::DV, 0, 1), 4, false);
    umma_commit(c.bar);
  }
  wg_wait_mma(c);
  FSTAMP(6);
  // ---- LayerNorm + ReLU (:105-106) -> classifier operand ------------------------------------------------------
  {
    float mean, rstd;
    ln_stats_tmem<C::DV>(c, C::COL_C, mean, rstd);
    const uint32_t cat0 = uint32_t(d * kL * C::DV + (c.rs & 15) * C::DV);   // my row's first concat column (:689-716)
    // self.drop (:719) acts on the classifier's copy only: the keep bits of my 48 concat columns, drawn once
    const uint64_t keep = a.drop.thresh ? drop_bits<C::DV>(a.drop, uint32_t(b0 + (c.rs >> 4)), cat0) : ~uint64_t(0);
    const float dscale = a.drop.thresh ? a.drop.scale : 1.0f;
    const float nmr = -mean * rstd;
    if (a.row_out[d])
      a.row_out[d][size_t(tile) * 128 + c.rs] = make_uint4(uint32_t(keep), uint32_t(keep >> 32), __float_as_uint(mean), __float_as_uint(rstd));
#pragma unroll
    for (int c0 = 0; c0 < C::DV; c0 += 16) {
      uint32_t r[16];
      tmem_ld16_nw(c.tmem + c.lane_base + C::COL_C + c0, r);
      tmem_wait_ld();
      const uint32_t kb = uint32_t(keep >> c0);
      float o[16];
#pragma unroll
      for (int e = 0; e < 16; ++e)
        o[e] = fmaxf(fmaf(fmaf(__uint_as_float(r[e]), rstd, nmr), ln_s[c0 + e], ln_s[48 + c0 + e]), 0.f) * ((kb >> e) & 1u ? dscale : 0.f);
      const float lo[8] = {o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]};
      const float hi[8] = {o[8], o[9], o[10], o[11], o[12], o[13], o[14], o[15]};
      *reinterpret_cast<uint4*>(xq + uint32_t(c0 >> 3) * kCS + row_off(c.rs)) = pack_bf16x8(lo);
      *reinterpret_cast<uint4*>(xq + uint32_t((c0 >> 3) + 1) * kCS + row_off(c.rs)) = pack_bf16x8(hi);
    }
  }
  wg_sync_for_mma(c);
  FSTAMP(7);
  if (c.wt == 0) {
    mma_steps(c.tmem + C::COL_CLS, make_smem_desc(smem_u32(xq), kCS, kRS), 2 * kCS,
              make_smem_desc(smem_u32(wsm + C::OFF_BC), C::BC_LBO, 128), 2 * C::BC_LBO,
              make_idesc_bf16(128, kNCls, 0, 0), C::DV / 16, false);
    mma_steps(c.tmem + C::COL_CLS, make_smem_desc(smem_u32(xq), kCS, kRS), 2 * kCS,      // += F (Wf - hi)^T
              make_smem_desc(smem_u32(wsm + C::OFF_BC_LO), C::BC_LBO, 128), 2 * C::BC_LBO,
              make_idesc_bf16(128, kNCls, 0, 0), C::DV / 16, true);
    umma_commit(c.bar);
  }
  wg_wait_mma(c);
  FSTAMP(8);
  if (c.wt == 0 && next_tile >= 0) ca_issue_q(a, d, next_tile, xq, bar_ld);      // the classifier has read Out: xq is free
  cls_readback(c, C::COL_CLS, a.logits, b0 + (c.rp >> 4), b0 + (c.rp >> 4) < a.batch);
  FSTAMP(9);
  tc_fence_before_sync();
  named_bar_sync(1 + c.wg, kWgThreads);
  tc_fence_after_sync();
  FSTAMP(10);
#undef FSTAMP
}

__global__ void __launch_bounds__(kCtaThreads, 1) ca_fwd_kernel(const CaFwdArgs a) {
  extern __shared__ __align__(128) uint8_t sm[];
  using L = CaFwdLayout;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::BAR);   // [0] weights, [1],[2] MMA, [3],[4] tile loads
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* ln_s = reinterpret_cast<float*>(sm + L::LN);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int d = blockIdx.y;       // direction: 0 = cross_attention_1, 1 = cross_attention_2
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(&bars[0], CaCfg::W_BYTES);
    bulk_g2s(sm + L::W0, a.dir[d].blobs, CaCfg::W_BYTES, &bars[0]);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < 48; i += kCtaThreads) {
    ln_s[i] = a.dir[d].ln_g[i];
    ln_s[48 + i] = a.dir[d].ln_b[i];
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  mbar_wait(&bars[0], 0);
  const int wg = tid >> 7;
  WgCtx c = make_ctx(wg * kWgThreads, tmem + uint32_t(wg) * 256u, &bars[1 + wg]);
  uint8_t* bsm = sm + (wg == 0 ? L::WG0 : L::WG1);
  uint32_t ph_ld = 0;
  const int tiles = (a.batch + 7) / 8;
  // the CTA's two warpgroups are independent pipelines on their own tiles of this direction
  const int first = 2 * int(blockIdx.x) + wg, stride = 2 * int(gridDim.x);
  int round = 0;
  if ((tid & 127) == 0 && first < tiles) {
    ca_issue_kv(a, d, first, bsm + CaFwdSmem::XKV, &bars[3 + wg]);
    ca_issue_q(a, d, first, bsm + CaFwdSmem::XQ, &bars[3 + wg]);
  }
  for (int tile = first; tile < tiles; tile += stride, ++round)
    ca_fwd_tile(c, a, d, sm + L::W0, bsm, ln_s, &bars[3 + wg], ph_ld, tile, tile + stride < tiles ? tile + stride : -1,
                1 + 12 * round);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace htc
}  // namespace mmrca

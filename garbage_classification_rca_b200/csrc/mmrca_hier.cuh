// Hierarchical late-fusion head (--late_fusion=hierarchical): forward, CrossEntropyLoss hand-off and backward.
//
// Reference path: CVPR_code/multimodal_model.py:777-816 (Hierarchical.forward after the backbones and the two
// AvgPool2d): six L2 normalisations, two concats (image 1280 + 2560 + 2048 = 5888, text 3 x 768 = 2304), self.drop on
// each, Linear(5888 -> 512) and Linear(2304 -> 512) with ReLU, Linear(1024 -> 4).  Backward (main_both.py:112) with the
// backbones frozen: gradients of the three Linear layers.
//
// Unlike the RCA head this one is two real GEMMs (K = 5888 / 2304, N = 512, M = batch) plus their weight gradients
// (M = 512, N = 5888 / 2304, K = batch): tensor-core bound.  Both run as warp-specialised tcgen05 pipelines (one
// bulk-copy producer thread, one MMA thread, four epilogue warps, fp32 accumulators in TMEM) over bf16 operand IMAGES in
// the canonical no-swizzle core-matrix layout: a 128-sample tile of the (normalised, dropped-out) concat is stored as
// [8-column group][128 rows][8 bf16], so the slab a pipeline stage needs is one contiguous piece of memory for the
// forward GEMM (K-major A operand) and for the weight-gradient GEMM (the same bytes read as an MN-major B operand).
//
//   hier_prep    features -> norms -> dropout -> bf16 X images (HBM-bound), logits = bias
//   hier_wprep   W_image / W_text fp32 [512][K] -> bf16 K-major blobs [K/8][512][8]
//   hier_gemm    H = ReLU(X W^T + b) -> bf16 H [B][1024] + partial logits (fp32 FMAs in the epilogue, atomics)
//   (cross_entropy_kernel of the fp32 path gives dlogits)
//   hier_dh      dH = [H > 0] dlogits W_all -> bf16 dH image; dW_all, db_all, db_image, db_text (column sums)
//   hier_wgrad   dW = dH^T X  (both operands MN-major straight from the images), accumulated into the caller's grads
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mmrca_attn_fp32.cuh"
#include "mmrca_dropout.cuh"
#include "mmrca_tc.cuh"

namespace mmrca {
namespace hier {

using namespace tc;

constexpr int kHid = 512;                         // multimodal_model.py:294-295
constexpr int kDImg = 1280 + 2560 + 2048;         // 5888
constexpr int kDTxt = 3 * 768;                    // 2304
constexpr int kD = kDImg + kDTxt;                 // 8192: the virtual concat the dropout mask is indexed by
constexpr int kGImg = kDImg / 8, kGTxt = kDTxt / 8, kGHid = 2 * kHid / 8;      // 8-column groups: 736, 288, 128
constexpr int kTile = 128;                        // samples per tile
constexpr uint32_t kGrpBytes = kTile * 16;        // one 8-column group of a tile: 128 rows x 16 bytes
constexpr int kClasses = 4;

// ---------------------------------------------------------------------------------------------------------------
// features -> X images
// ---------------------------------------------------------------------------------------------------------------
struct PrepArgs {
  const float* seg[6];      // pooled [B,1280], stage-3 [B,2560], stage-6 [B,2048], text CLS last / layer 2 / layer 4 [B,768]
  void* x_img;              // [tiles][736][128][8] bf16
  void* x_txt;              // [tiles][288][128][8] bf16
  const uint8_t* mask;      // caller-drawn keep mask [B][8192] (image concat columns first) or null
  float mask_scale;
  DropSpec drop;            // seeded dropout (D = 8192) when mask == null and thresh != 0
  float* logits;            // [B][4]: initialised with the bias of final_hierarchical_all
  const float* b_all;
  int batch;
};

// One CTA per sample (padding samples of the last tile included: they become zero rows), 256 threads x 4 items of 8
// columns = the 8192 columns.  Every segment boundary is a multiple of 32 items, so a warp's items of one round belong
// to one segment and its sum of squares is one warp reduction + one shared-memory atomic.
__global__ void __launch_bounds__(256) hier_prep_kernel(const PrepArgs a) {
  __shared__ float s_ss[6];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const bool live = b < a.batch;
  if (tid < 6) s_ss[tid] = 0.f;
  __syncthreads();
  float v[4][8];
  int seg[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int item = tid + 256 * k;
    int s, first, len;
    const float* base;
    if (item < 160) { s = 0; first = 0; len = 1280; base = a.seg[0]; }
    else if (item < 480) { s = 1; first = 160; len = 2560; base = a.seg[1]; }
    else if (item < 736) { s = 2; first = 480; len = 2048; base = a.seg[2]; }
    else if (item < 832) { s = 3; first = 736; len = 768; base = a.seg[3]; }
    else if (item < 928) { s = 4; first = 832; len = 768; base = a.seg[4]; }
    else { s = 5; first = 928; len = 768; base = a.seg[5]; }
    seg[k] = s;
    float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
    if (live) {
      const float* p = base + size_t(b) * len + (item - first) * 8;
      lo = __ldg(reinterpret_cast<const float4*>(p));
      hi = __ldg(reinterpret_cast<const float4*>(p + 4));
    }
    v[k][0] = lo.x; v[k][1] = lo.y; v[k][2] = lo.z; v[k][3] = lo.w; v[k][4] = hi.x; v[k][5] = hi.y; v[k][6] = hi.z; v[k][7] = hi.w;
    float ss = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) ss = fmaf(v[k][e], v[k][e], ss);
    ss = warp_sum(ss);
    if (lane == 0) atomicAdd(&s_ss[s], ss);
  }
  __syncthreads();
  const int tile = b >> 7, row = b & 127;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int item = tid + 256 * k;
    const float inv = live ? 1.0f / sqrtf(s_ss[seg[k]]) : 0.f;       // no epsilon, like the reference (:777-789)
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = v[k][e] * inv;
    if (live) {
      if (a.mask) {
        const uint2 m = __ldg(reinterpret_cast<const uint2*>(a.mask + size_t(b) * kD + item * 8));
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] *= (((e < 4 ? m.x : m.y) >> (8 * (e & 3))) & 0xffu) ? a.mask_scale : 0.f;
      } else if (a.drop.thresh) {
        drop_apply8(a.drop, uint32_t(b), uint32_t(item * 8), o);
      }
    }
    uint8_t* dst = item < kGImg
                       ? static_cast<uint8_t*>(a.x_img) + ((size_t(tile) * kGImg + item) * kTile + row) * 16
                       : static_cast<uint8_t*>(a.x_txt) + ((size_t(tile) * kGTxt + (item - kGImg)) * kTile + row) * 16;
    *reinterpret_cast<uint4*>(dst) = pack_bf16x8(o);
  }
  if (tid < kClasses && live) a.logits[size_t(b) * kClasses + tid] = __ldg(a.b_all + tid);
}

// W [512][K] fp32 -> blob [K/8][512][8] bf16 (the K-major B operand of the forward GEMM).  CTA = 32 rows n x 32 column
// groups: a warp reads 1 KB of one weight row per instruction, the tile is transposed through shared memory and a warp
// writes the 32 consecutive 16-byte entries of one column group (reads and writes both coalesced).
// grid = (16 row blocks, 23 + 9 group blocks: image then text)
struct WPrepArgs { const float* w[2]; void* blob[2]; };
__global__ void __launch_bounds__(256) hier_wprep_kernel(const WPrepArgs a) {
  __shared__ uint4 t[32][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int gb = blockIdx.y, mod = 0;
  if (gb >= kGImg / 32) { gb -= kGImg / 32; mod = 1; }
  const int K = mod == 0 ? kDImg : kDTxt, n0 = blockIdx.x * 32, kg0 = gb * 32;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int nl = warp + 8 * j;
    const float* p = a.w[mod] + size_t(n0 + nl) * K + (kg0 + lane) * 8;
    const float4 lo = __ldg(reinterpret_cast<const float4*>(p)), hi = __ldg(reinterpret_cast<const float4*>(p + 4));
    const float o[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    t[lane][nl] = pack_bf16x8(o);      // [column group][row]
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int kl = warp + 8 * j;
    *reinterpret_cast<uint4*>(static_cast<uint8_t*>(a.blob[mod]) + (size_t(kg0 + kl) * kHid + n0 + lane) * 16) = t[kl][lane];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// warp-specialised tcgen05 GEMM pipelines.  192 threads: warp 0 = bulk-copy producer, warp 1 = MMA issuer (and TMEM
// owner), warps 2-5 = epilogue (warp w reads TMEM lanes 32 (w % 4) .. + 32).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kGemmThreads = 192;
constexpr int kBN = 256;                               // accumulator columns per CTA

// forward: H[128 x 256] = X tile [128 x K] * W[256 rows of 512][K]^T
struct GemmArgs {
  const void* x[2];          // X images (image, text)
  const void* wb[2];         // weight blobs
  const float* bias[2];      // [512]
  const float* w_all;        // [4][1024] fp32
  float* logits;             // [B][4], += partial logits
  void* h;                   // bf16 [tiles * 128][1024]: ReLU output, image half then text half
  int batch;
};
constexpr int kFwdStages = 4, kFwdKG = 8;                                   // 8 column groups = K 64 per stage
constexpr uint32_t kFwdABytes = kFwdKG * kGrpBytes;                         // 16 KB
constexpr uint32_t kFwdBBytes = kFwdKG * kBN * 16;                          // 32 KB
struct FwdSmem {
  static constexpr uint32_t A = 0, B = A + kFwdStages * kFwdABytes, W3 = B + kFwdStages * kFwdBBytes;
  static constexpr uint32_t BIAS = W3 + kClasses * kBN * 4, BAR = BIAS + kBN * 4, BYTES = BAR + 128;
  static_assert(BYTES <= 232448, "hierarchical forward GEMM does not fit shared memory");
};

__global__ void __launch_bounds__(kGemmThreads, 1) hier_gemm_kernel(const GemmArgs a) {
  extern __shared__ __align__(128) uint8_t sm[];
  using S = FwdSmem;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + S::BAR);      // [stages]
  uint64_t* empty = full + kFwdStages;                            // [stages]
  uint64_t* accb = empty + kFwdStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accb + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, nh = blockIdx.y, mod = blockIdx.z;
  const int KG = mod == 0 ? kGImg : kGTxt, n_it = KG / kFwdKG;
  if (tid == 0) {
    for (int i = 0; i < 2 * kFwdStages + 1; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kBN);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    if (lane == 0) {
      const uint8_t* xa = static_cast<const uint8_t*>(a.x[mod]) + size_t(tile) * KG * kGrpBytes;
      const uint8_t* wb = static_cast<const uint8_t*>(a.wb[mod]) + size_t(nh) * kBN * 16;
      for (int it = 0; it < n_it; ++it) {
        const int s = it % kFwdStages;
        if (it >= kFwdStages) mbar_wait(&empty[s], uint32_t(it / kFwdStages - 1) & 1u);
        mbar_arrive_expect_tx(&full[s], kFwdABytes + kFwdBBytes);
        bulk_g2s(sm + S::A + s * kFwdABytes, xa + size_t(it) * kFwdABytes, kFwdABytes, &full[s]);
#pragma unroll
        for (int g = 0; g < kFwdKG; ++g)
          bulk_g2s(sm + S::B + s * kFwdBBytes + g * (kBN * 16), wb + size_t(it * kFwdKG + g) * (kHid * 16), kBN * 16, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, kBN, 0, 0);
      for (int it = 0; it < n_it; ++it) {
        const int s = it % kFwdStages;
        mbar_wait(&full[s], uint32_t(it / kFwdStages) & 1u);
        tc_fence_after_sync();
        const uint64_t ad = make_smem_desc(smem_u32(sm + S::A + s * kFwdABytes), kGrpBytes, 128);
        const uint64_t bd = make_smem_desc(smem_u32(sm + S::B + s * kFwdBBytes), kBN * 16, 128);
#pragma unroll
        for (int ks = 0; ks < kFwdKG / 2; ++ks)
          umma_bf16(tmem, desc_advance(ad, ks * 2 * kGrpBytes), desc_advance(bd, ks * 2 * (kBN * 16)), idesc, (it | ks) ? 1u : 0u);
        umma_commit(&empty[s]);
      }
      umma_commit(accb);
    }
  } else {
    // ---- epilogue: bias + ReLU -> bf16 H, partial logits of final_hierarchical_all (:808-816) -------------------
    const int et = tid - 64;                                  // 0 .. 127
    float* w3s = reinterpret_cast<float*>(sm + S::W3);        // [4][256]
    float* bs = reinterpret_cast<float*>(sm + S::BIAS);       // [256]
    const int col0 = mod * kHid + nh * kBN;                   // first hidden column of this CTA
    for (int i = et; i < kClasses * kBN; i += 128) w3s[i] = __ldg(a.w_all + size_t(i / kBN) * (2 * kHid) + col0 + (i % kBN));
    for (int i = et; i < kBN; i += 128) bs[i] = __ldg(a.bias[mod] + nh * kBN + i);
    named_bar_sync(1, 128);
    const int q = warp & 3, row = 32 * q + lane, b = tile * kTile + row;
    mbar_wait(accb, 0);
    tc_fence_after_sync();
    float l[kClasses] = {0.f, 0.f, 0.f, 0.f};
    uint8_t* hrow = static_cast<uint8_t*>(a.h) + (size_t(b) * (2 * kHid) + col0) * 2;
#pragma unroll 1
    for (int c0 = 0; c0 < kBN; c0 += 32) {
      uint32_t r0[16], r1[16];
      tmem_ld16_nw(tmem + (uint32_t(32 * q) << 16) + c0, r0);
      tmem_ld16_nw(tmem + (uint32_t(32 * q) << 16) + c0 + 16, r1);
      tmem_wait_ld();
      float h[32];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        h[e] = fmaxf(__uint_as_float(r0[e]) + bs[c0 + e], 0.f);
        h[16 + e] = fmaxf(__uint_as_float(r1[e]) + bs[c0 + 16 + e], 0.f);
      }
#pragma unroll
      for (int e = 0; e < 32; ++e) {
#pragma unroll
        for (int c = 0; c < kClasses; ++c) l[c] = fmaf(h[e], w3s[c * kBN + c0 + e], l[c]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float o[8] = {h[8 * j], h[8 * j + 1], h[8 * j + 2], h[8 * j + 3], h[8 * j + 4], h[8 * j + 5], h[8 * j + 6], h[8 * j + 7]};
        *reinterpret_cast<uint4*>(hrow + (c0 + 8 * j) * 2) = pack_bf16x8(o);
      }
    }
    if (b < a.batch) {
#pragma unroll
      for (int c = 0; c < kClasses; ++c) red_add(a.logits + size_t(b) * kClasses + c, l[c]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kBN);
}

// ---------------------------------------------------------------------------------------------------------------
// dH = [H > 0] (dlogits W_all) -> bf16 dH image [tiles][128 groups][128 rows][8]; dW_all += dlogits^T H;
// db_all += sum_b dlogits; db_image | db_text += sum_b dH.   grid = (tiles, 8 blocks of 128 hidden columns)
// ---------------------------------------------------------------------------------------------------------------
struct DhArgs {
  const void* h;            // bf16 [tiles * 128][1024]
  const float* dlogits;     // [B][4]
  const float* w_all;       // [4][1024]
  void* dh;                 // dH image out
  float* g_w_all;           // [4][1024] += (null: skip)
  float* g_b_all;           // [4] +=
  float* g_b_hid[2];        // [512] += each
  int batch;
};
__global__ void __launch_bounds__(256) hier_dh_kernel(const DhArgs a) {
  __shared__ float part[8][16][41];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, cb = blockIdx.y;            // cb: block of 16 column groups
  const int g = tid & 15, rsub = tid >> 4;                  // my column group, my rows rsub + 16 j
  const int n0 = (cb * 16 + g) * 8;                         // my first hidden column
  float w3[kClasses][8];
#pragma unroll
  for (int c = 0; c < kClasses; ++c) {
    const float4 lo = __ldg(reinterpret_cast<const float4*>(a.w_all + size_t(c) * (2 * kHid) + n0));
    const float4 hi = __ldg(reinterpret_cast<const float4*>(a.w_all + size_t(c) * (2 * kHid) + n0 + 4));
    w3[c][0] = lo.x; w3[c][1] = lo.y; w3[c][2] = lo.z; w3[c][3] = lo.w; w3[c][4] = hi.x; w3[c][5] = hi.y; w3[c][6] = hi.z; w3[c][7] = hi.w;
  }
  float acc[40];                                            // [0,8): db_hidden; [8 + 8c, 16 + 8c): dW_all row c
#pragma unroll
  for (int i = 0; i < 40; ++i) acc[i] = 0.f;
  float dlsum[kClasses] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
  for (int j = 0; j < 8; ++j) {
    const int row = rsub + 16 * j, b = tile * kTile + row;
    float dh[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (b < a.batch) {
      const uint4 hv = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(a.h) + (size_t(b) * (2 * kHid) + n0) * 2));
      const float4 d4 = __ldg(reinterpret_cast<const float4*>(a.dlogits) + b);
      const float dl[kClasses] = {d4.x, d4.y, d4.z, d4.w};
      const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float h = __uint_as_float((e & 1) ? (hw[e >> 1] & 0xffff0000u) : (hw[e >> 1] << 16));
        float t = 0.f;
#pragma unroll
        for (int c = 0; c < kClasses; ++c) { t = fmaf(dl[c], w3[c][e], t); acc[8 + 8 * c + e] = fmaf(dl[c], h, acc[8 + 8 * c + e]); }
        dh[e] = h > 0.f ? t : 0.f;
        acc[e] += dh[e];
      }
      if (g == 0 && cb == 0) {
#pragma unroll
        for (int c = 0; c < kClasses; ++c) dlsum[c] += dl[c];
      }
    }
    *reinterpret_cast<uint4*>(static_cast<uint8_t*>(a.dh) + ((size_t(tile) * kGHid + cb * 16 + g) * kTile + row) * 16) = pack_bf16x8(dh);
  }
  // lanes l and l ^ 16 share the column group; then the 8 warps meet in shared memory
#pragma unroll
  for (int i = 0; i < 40; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
    if (lane < 16) part[warp][lane][i] = acc[i];
  }
  if (cb == 0) {      // db_all: threads with g == 0 hold the per-row dlogits sums (lanes 0 and 16 of every warp)
#pragma unroll
    for (int c = 0; c < kClasses; ++c) {
      float v = g == 0 ? dlsum[c] : 0.f;
      v = warp_sum(v);
      if (lane == 0 && a.g_b_all) red_add(a.g_b_all + c, v);
    }
  }
  __syncthreads();
  for (int o = tid; o < 16 * 40; o += 256) {
    const int gg = o / 40, i = o - gg * 40;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += part[w][gg][i];
    const int n = (cb * 16 + gg) * 8 + (i & 7);
    if (i < 8) red_add(a.g_b_hid[n >= kHid ? 1 : 0] + (n & (kHid - 1)), v);
    else if (a.g_w_all) red_add(a.g_w_all + size_t((i - 8) >> 3) * (2 * kHid) + n, v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradients: dW[512][K] += dH^T X, CTA = (128 rows of dW, 256 columns), K-loop over the sample tiles.
// A = dH image (MN-major: m = hidden column, k = sample), B = X image (MN-major: n = concat column, k = sample).
// ---------------------------------------------------------------------------------------------------------------
struct WgradArgs {
  const void* dh;            // dH image [tiles][128][128][8]
  const void* x[2];          // X images
  float* g_w[2];             // [512][5888], [512][2304]  +=
  int tiles;
};
constexpr int kWgStages = 2;
constexpr uint32_t kWgABytes = 16 * kGrpBytes;         // 128 hidden columns x 128 samples: 32 KB
constexpr uint32_t kWgBBytes = 32 * kGrpBytes;         // 256 concat columns x 128 samples: 64 KB
constexpr int kWgCtasImg = 4 * (kDImg / kBN), kWgCtasTxt = 4 * (kDTxt / kBN);      // 92 + 36
struct WgSmem {
  static constexpr uint32_t A = 0, B = A + kWgStages * kWgABytes, BAR = B + kWgStages * kWgBBytes, BYTES = BAR + 128;
  static_assert(BYTES <= 232448, "hierarchical weight-gradient GEMM does not fit shared memory");
};

__global__ void __launch_bounds__(kGemmThreads, 1) hier_wgrad_kernel(const WgradArgs a) {
  extern __shared__ __align__(128) uint8_t sm[];
  using S = WgSmem;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + S::BAR);
  uint64_t* empty = full + kWgStages;
  uint64_t* accb = empty + kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accb + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int cta = blockIdx.x, mod = 0;
  if (cta >= kWgCtasImg) { cta -= kWgCtasImg; mod = 1; }
  const int K = mod == 0 ? kDImg : kDTxt, KG = K / 8, nts = K / kBN;
  const int mt = cta / nts, nt = cta - mt * nts;
  if (tid == 0) {
    for (int i = 0; i < 2 * kWgStages + 1; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kBN);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    if (lane == 0) {
      const uint8_t* pa = static_cast<const uint8_t*>(a.dh) + size_t(mod * (kHid / 8) + mt * 16) * kGrpBytes;
      const uint8_t* pb = static_cast<const uint8_t*>(a.x[mod]) + size_t(nt * 32) * kGrpBytes;
      for (int it = 0; it < a.tiles; ++it) {
        const int s = it % kWgStages;
        if (it >= kWgStages) mbar_wait(&empty[s], uint32_t(it / kWgStages - 1) & 1u);
        mbar_arrive_expect_tx(&full[s], kWgABytes + kWgBBytes);
        bulk_g2s(sm + S::A + s * kWgABytes, pa + size_t(it) * kGHid * kGrpBytes, kWgABytes, &full[s]);
        bulk_g2s(sm + S::B + s * kWgBBytes, pb + size_t(it) * KG * kGrpBytes, kWgBBytes, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, kBN, 1, 1);
      for (int it = 0; it < a.tiles; ++it) {
        const int s = it % kWgStages;
        mbar_wait(&full[s], uint32_t(it / kWgStages) & 1u);
        tc_fence_after_sync();
        const uint64_t ad = make_smem_desc(smem_u32(sm + S::A + s * kWgABytes), 128, kGrpBytes);
        const uint64_t bd = make_smem_desc(smem_u32(sm + S::B + s * kWgBBytes), 128, kGrpBytes);
#pragma unroll
        for (int ks = 0; ks < kTile / 16; ++ks)
          umma_bf16(tmem, desc_advance(ad, ks * 256), desc_advance(bd, ks * 256), idesc, (it | ks) ? 1u : 0u);
        umma_commit(&empty[s]);
      }
      umma_commit(accb);
    }
  } else {
    const int q = warp & 3, row = 32 * q + lane;
    float* dst = a.g_w[mod] + size_t(mt * 128 + row) * K + nt * kBN;
    mbar_wait(accb, 0);
    tc_fence_after_sync();
#pragma unroll 1
    for (int c0 = 0; c0 < kBN; c0 += 32) {
      uint32_t r0[16], r1[16];
      tmem_ld16_nw(tmem + (uint32_t(32 * q) << 16) + c0, r0);
      tmem_ld16_nw(tmem + (uint32_t(32 * q) << 16) + c0 + 16, r1);
      float4 old[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) old[j] = *reinterpret_cast<const float4*>(dst + c0 + 4 * j);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t* r = j < 4 ? &r0[4 * j] : &r1[4 * (j - 4)];
        old[j].x += __uint_as_float(r[0]); old[j].y += __uint_as_float(r[1]);
        old[j].z += __uint_as_float(r[2]); old[j].w += __uint_as_float(r[3]);
        *reinterpret_cast<float4*>(dst + c0 + 4 * j) = old[j];
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kBN);
}

// ---------------------------------------------------------------------------------------------------------------
// feature gradients (fine-tune phase, main_both.py:687-694): d(dropped concat) = dH W, a GEMM with the SAME operand bytes
// as the forward - A = the dH image (K-major over the 512 hidden units of the modality), B = the forward's weight blob
// [k/8][512][8] read MN-major (n = concat column contiguous, K = hidden unit) - then, per sample, the dropout mask and the
// six L2-norm backwards (hier_dx_finish_kernel).  CTA = (sample tile, 256 concat columns); grid.y walks the image columns
// (23 blocks) then the text columns (9 blocks).
// ---------------------------------------------------------------------------------------------------------------
struct DxArgs {
  const void* dh;            // dH image [tiles][128 groups][128][8]
  const void* wb[2];         // weight blobs [K/8][512][8]
  float* dcat;               // out: d(dropped concat) fp32 [tiles * 128][8192] (image columns first)
};
constexpr int kDxStages = 2, kDxKG = 16;                                    // 16 hidden groups = K 128 per stage
constexpr uint32_t kDxABytes = kDxKG * kGrpBytes;                           // 32 KB
constexpr uint32_t kDxBBytes = (kBN / 8) * (kDxKG * 8) * 16;                // 32 column groups x 128 hidden x 16 B = 64 KB
struct DxSmem {
  static constexpr uint32_t A = 0, B = A + kDxStages * kDxABytes, BAR = B + kDxStages * kDxBBytes, BYTES = BAR + 128;
  static_assert(BYTES <= 232448, "hierarchical feature-gradient GEMM does not fit shared memory");
};

__global__ void __launch_bounds__(kGemmThreads, 1) hier_dx_kernel(const DxArgs a) {
  extern __shared__ __align__(128) uint8_t sm[];
  using S = DxSmem;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + S::BAR);
  uint64_t* empty = full + kDxStages;
  uint64_t* accb = empty + kDxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accb + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  int cb = blockIdx.y, mod = 0;
  if (cb >= kDImg / kBN) { cb -= kDImg / kBN; mod = 1; }
  constexpr int n_it = kHid / (kDxKG * 8);                                  // 4 K-stages
  if (tid == 0) {
    for (int i = 0; i < 2 * kDxStages + 1; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kBN);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    if (lane == 0) {
      const uint8_t* pa = static_cast<const uint8_t*>(a.dh) + (size_t(tile) * kGHid + size_t(mod) * (kHid / 8)) * kGrpBytes;
      const uint8_t* pb = static_cast<const uint8_t*>(a.wb[mod]) + size_t(cb) * (kBN / 8) * (kHid * 16);
      for (int it = 0; it < n_it; ++it) {
        const int s = it % kDxStages;
        if (it >= kDxStages) mbar_wait(&empty[s], uint32_t(it / kDxStages - 1) & 1u);
        mbar_arrive_expect_tx(&full[s], kDxABytes + kDxBBytes);
        bulk_g2s(sm + S::A + s * kDxABytes, pa + size_t(it) * kDxABytes, kDxABytes, &full[s]);
#pragma unroll 1
        for (int g = 0; g < kBN / 8; ++g)      // column group g: hidden units [128 it, 128 it + 128) are 2 KB contiguous
          bulk_g2s(sm + S::B + s * kDxBBytes + g * (kDxKG * 8 * 16), pb + (size_t(g) * kHid + size_t(it) * kDxKG * 8) * 16,
                   kDxKG * 8 * 16, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, kBN, 0, 1);
      for (int it = 0; it < n_it; ++it) {
        const int s = it % kDxStages;
        mbar_wait(&full[s], uint32_t(it / kDxStages) & 1u);
        tc_fence_after_sync();
        const uint64_t ad = make_smem_desc(smem_u32(sm + S::A + s * kDxABytes), kGrpBytes, 128);
        const uint64_t bd = make_smem_desc(smem_u32(sm + S::B + s * kDxBBytes), 128, kDxKG * 8 * 16);
#pragma unroll
        for (int ks = 0; ks < kDxKG / 2; ++ks)
          umma_bf16(tmem, desc_advance(ad, ks * 2 * kGrpBytes), desc_advance(bd, ks * 256), idesc, (it | ks) ? 1u : 0u);
        umma_commit(&empty[s]);
      }
      umma_commit(accb);
    }
  } else {
    const int q = warp & 3, row = 32 * q + lane;
    float* dst = a.dcat + size_t(tile * kTile + row) * kD + (mod ? kDImg : 0) + cb * kBN;
    mbar_wait(accb, 0);
    tc_fence_after_sync();
#pragma unroll 1
    for (int c0 = 0; c0 < kBN; c0 += 32) {
      uint32_t r0[16], r1[16];
      tmem_ld16_nw(tmem + (uint32_t(32 * q) << 16) + c0, r0);
      tmem_ld16_nw(tmem + (uint32_t(32 * q) << 16) + c0 + 16, r1);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        *reinterpret_cast<uint4*>(dst + c0 + 4 * j) = make_uint4(r0[4 * j], r0[4 * j + 1], r0[4 * j + 2], r0[4 * j + 3]);
        *reinterpret_cast<uint4*>(dst + c0 + 16 + 4 * j) = make_uint4(r1[4 * j], r1[4 * j + 1], r1[4 * j + 2], r1[4 * j + 3]);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kBN);
}

// per sample: d(x_seg) = (m . d - xn (xn . (m . d))) / ||x_seg|| for the six segments (m: the dropout multipliers of the forward).
// One CTA per sample, the thread / item / segment mapping of hier_prep_kernel.
struct DxFinishArgs {
  const float* seg[6];       // the forward's raw features
  const float* dcat;         // [tiles * 128][8192]
  const uint8_t* mask; float mask_scale; DropSpec drop;
  float* out[6];             // feature gradients, same shapes as seg
  int batch;
};
__global__ void __launch_bounds__(256) hier_dx_finish_kernel(const DxFinishArgs a) {
  __shared__ float s_ss[6], s_dot[6];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  if (tid < 6) { s_ss[tid] = 0.f; s_dot[tid] = 0.f; }
  __syncthreads();
  float v[4][8], d[4][8];
  int seg[4], first[4], len[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int item = tid + 256 * k;
    if (item < 160) { seg[k] = 0; first[k] = 0; len[k] = 1280; }
    else if (item < 480) { seg[k] = 1; first[k] = 160; len[k] = 2560; }
    else if (item < 736) { seg[k] = 2; first[k] = 480; len[k] = 2048; }
    else if (item < 832) { seg[k] = 3; first[k] = 736; len[k] = 768; }
    else if (item < 928) { seg[k] = 4; first[k] = 832; len[k] = 768; }
    else { seg[k] = 5; first[k] = 928; len[k] = 768; }
    const float* p = a.seg[seg[k]] + size_t(b) * len[k] + (item - first[k]) * 8;
    const float4 lo = __ldg(reinterpret_cast<const float4*>(p)), hi = __ldg(reinterpret_cast<const float4*>(p + 4));
    v[k][0] = lo.x; v[k][1] = lo.y; v[k][2] = lo.z; v[k][3] = lo.w; v[k][4] = hi.x; v[k][5] = hi.y; v[k][6] = hi.z; v[k][7] = hi.w;
    const float* q = a.dcat + size_t(b) * kD + item * 8;
    const float4 dlo = __ldg(reinterpret_cast<const float4*>(q)), dhi = __ldg(reinterpret_cast<const float4*>(q + 4));
    d[k][0] = dlo.x; d[k][1] = dlo.y; d[k][2] = dlo.z; d[k][3] = dlo.w; d[k][4] = dhi.x; d[k][5] = dhi.y; d[k][6] = dhi.z; d[k][7] = dhi.w;
    if (a.mask) {
      const uint2 m = __ldg(reinterpret_cast<const uint2*>(a.mask + size_t(b) * kD + item * 8));
#pragma unroll
      for (int e = 0; e < 8; ++e) d[k][e] *= (((e < 4 ? m.x : m.y) >> (8 * (e & 3))) & 0xffu) ? a.mask_scale : 0.f;
    } else if (a.drop.thresh) {
      drop_apply8(a.drop, uint32_t(b), uint32_t(item * 8), d[k]);
    }
    float ss = 0.f, dot = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) { ss = fmaf(v[k][e], v[k][e], ss); dot = fmaf(v[k][e], d[k][e], dot); }
    ss = warp_sum(ss); dot = warp_sum(dot);
    if (lane == 0) { atomicAdd(&s_ss[seg[k]], ss); atomicAdd(&s_dot[seg[k]], dot); }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int item = tid + 256 * k;
    const float n2 = s_ss[seg[k]], inv = 1.0f / sqrtf(n2);
    const float proj = s_dot[seg[k]] / n2;          // (x . d) / ||x||^2: xn (xn . d) = x proj
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = (d[k][e] - v[k][e] * proj) * inv;
    float* p = a.out[seg[k]] + size_t(b) * len[k] + (item - first[k]) * 8;
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
}

}  // namespace hier
}  // namespace mmrca

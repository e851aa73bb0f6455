// bf16 tensor-core pipeline of the MM-RCA head: cross-entropy, backward kernels, gradient finalisation.
//
// Backward of CVPR_code/multimodal_model.py:661-728 + CrossEntropyLoss (main_both.py:87-93,110-112), in the
// algebra of mmrca_head_tc.cuh (Z = Xq M + u replaces Q and K).  Per attention block and 128-row tile:
//   forward     CA: Z, V, S, P, C are recomputed on chip (only the block inputs, the dropout keep bits and the LayerNorm
//               statistics are read back; reloading Z / V / P was measured slower);
//               SA: V, P and the LayerNorm statistics are the forward's (TMA, a tile ahead), C = P V is one MMA
//   LayerNorm   dy = dOut [y > 0],  dC = rstd (dy g - mean(dy g) - xhat mean(dy g xhat))
//   attention   dP = dC V^T,  dV = P^T dC,  dS = softmax'(dP),  dZ = dS Xkv
//   inputs      dXq = dZ M^T,  dXkv = dS^T Z + dV Wv            (CA blocks: they feed the SA backward)
//   parameters  dM_ext += Xq_ext^T dZ,  dWv_ext += Xkv_ext^T dV   (the ones column of X_ext yields du, db_value)
//               dWf^T  += F^T DL        (classifier rows of a cross-attention source; DL = dlogits spread on the
//                                        chunk diagonal.  The rows of the feature sources: ce_feat_kernel, fp32)
//               [dgamma | dbeta] += [dy xhat | dy]^T 1
// Every contraction, including the reductions over the batch, is a tcgen05.mma.  The parameter gradients stay in
// TMEM for the whole persistent CTA and are flushed once at the end with lane-coalesced red.global; dW_query,
// dW_key, db_query follow from dM, du in finalize_kernel (products of size d_in^2, once per step).
// A CTA owns one tile at a time; its two warpgroups own the same rows and split the columns / the kinds of
// operand they emit.  The CA backward has a ninth warp that requests the next tile's images and issues the
// persistent dM / dWv MMAs (tcgen05.mma issue blocks while the tensor queue is full).
#pragma once
#include "mmrca_head_tc.cuh"

namespace mmrca {
namespace htc {

__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float (&f)[8]) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}

// ---------------------------------------------------------------------------------------------------------------
// CrossEntropyLoss(weight, label_smoothing), mean reduction (main_both.py:87-93), 4 classes, fused with the two
// reductions over the batch that need nothing but dlogits and the (normalised) features:
//   db_f[c]          += sum_b dlogits[b][c]
//   dWf[c][off + j]  += sum_b dlogits[b][c] drop(xn[b][j])               (feature sources of the concat, fp32 FMAs)
// grid = (column strips, sample slices).  A strip is one 8-column group kc of one modality's X image (10 image +
// 6 text strips): for a tile it is 2 KB of contiguous bf16, one 16-byte row per thread.  A slice is 16 tiles = 128
// samples = one cross-entropy evaluation per thread of the first four warps.  Thread (row r) accumulates the 4 x 8 products of its chunk
// r % 16 over the slice's tiles in registers; the 16 threads that share a chunk meet in shared memory and leave
// with one 16-byte atomic per class.  Every CTA recomputes the normaliser sum_b w[y_b] from the labels, so
// dlogits leave normalised in one pass.  labels == null: dlogits are given (autograd backward).
// ---------------------------------------------------------------------------------------------------------------
struct CeFeatArgs {
  const float* logits; const int64_t* labels; const float* cw; float eps; int batch;
  float* dlogits;     // [B][4]: written when labels != null, read otherwise
  float* loss;        // [1], zeroed beforehand (accumulated); CE mode only
  float* g_bf;        // [4] accumulated (null: skip)
  const void* x_img; const void* x_txt;           // X images written by prep_feat_kernel (null: no feature term)
  float* g_wf;        // [4][D]
  int off_img, off_txt;                           // first concat column of each feature source
  DropSpec drop;      // drop.D = D
};
constexpr int kCeSliceTiles = 16;      // 128 samples per slice: 10 + 6 strips x 32 slices = 512 CTAs at batch 4096
constexpr int kStripsImg = 80 / 8, kStripsTxt = 48 / 8;

__global__ void __launch_bounds__(256) ce_feat_kernel(const CeFeatArgs a) {
  __shared__ float red[8];
  __shared__ float s_den;
  __shared__ float4 dl_s[kCeSliceTiles * 8];
  __shared__ float part[8][16][33];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool ce = a.labels != nullptr;
  const int strip = blockIdx.x, tile0 = blockIdx.y * kCeSliceTiles, tiles = (a.batch + 7) / 8;
  float den = float(a.batch);
  if (ce && a.cw) {
    float w = 0.f;
    for (int b = tid; b < a.batch; b += 256) w += __ldg(a.cw + a.labels[b]);
    w = warp_sum(w);
    if (lane == 0) red[warp] = w;
    __syncthreads();
    if (tid == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; s_den = t; }
    __syncthreads();
    den = s_den;
  }
  // ---- dlogits of my sample --------------------------------------------------------------------------------------
  {
    const int b = tile0 * 8 + tid;
    float d[4] = {0.f, 0.f, 0.f, 0.f}, li = 0.f;
    if (tid < kCeSliceTiles * 8 && b < a.batch) {      // the slice's samples: one per thread of the first warps
      if (ce) {
        const float inv_den = 1.0f / den;
        float wc[4] = {1.f, 1.f, 1.f, 1.f};
        if (a.cw) { for (int c = 0; c < 4; ++c) wc[c] = __ldg(a.cw + c); }
        const float4 zz = __ldg(reinterpret_cast<const float4*>(a.logits) + b);
        float z[4] = {zz.x, zz.y, zz.z, zz.w};
        const float m = fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3]));
        const float se = expf(z[0] - m) + expf(z[1] - m) + expf(z[2] - m) + expf(z[3] - m);
        const float lse = m + logf(se);
        const int y = int(a.labels[b]);
        float t[4], tsum = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          t[c] = (a.eps * 0.25f) * wc[c] + (c == y ? (1.f - a.eps) * wc[c] : 0.f);
          z[c] -= lse;
          li -= t[c] * z[c];
          tsum += t[c];
        }
        li *= inv_den;
#pragma unroll
        for (int c = 0; c < 4; ++c) d[c] = (expf(z[c]) * tsum - t[c]) * inv_den;
        if (strip == 0 && a.dlogits) reinterpret_cast<float4*>(a.dlogits)[b] = make_float4(d[0], d[1], d[2], d[3]);
      } else {
        const float4 dd = __ldg(reinterpret_cast<const float4*>(a.dlogits) + b);
        d[0] = dd.x; d[1] = dd.y; d[2] = dd.z; d[3] = dd.w;
      }
    }
    if (tid < kCeSliceTiles * 8) dl_s[tid] = make_float4(d[0], d[1], d[2], d[3]);
    if (strip == 0) {       // the strips of a slice share its samples: one of them owns the loss and the bias gradient
      li = warp_sum(li);
#pragma unroll
      for (int c = 0; c < 4; ++c) d[c] = warp_sum(d[c]);
      if (lane == 0) {
        if (ce && a.loss) red_add(a.loss, li);
        if (a.g_bf) { for (int c = 0; c < 4; ++c) red_add(a.g_bf + c, d[c]); }
      }
    }
  }
  if (!a.g_wf) return;
  __syncthreads();
  // ---- my strip of the feature-source rows of dWf --------------------------------------------------------------------
  const bool is_img = strip < kStripsImg;
  const int kc = is_img ? strip : strip - kStripsImg, din = is_img ? 80 : 48;
  const uint32_t tile_bytes = is_img ? x_tile_bytes(80) : x_tile_bytes(48);
  const uint8_t* base = static_cast<const uint8_t*>(is_img ? a.x_img : a.x_txt) + uint32_t(kc) * kCS;
  const int r = tid & 127, half = tid >> 7, chunk = r & 15, g = r >> 4;
  const int col0 = (is_img ? a.off_img : a.off_txt) + chunk * din + kc * 8;      // my 8 concat columns
  float acc[4][8];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[c][e] = 0.f;
  uint4 raw[kCeSliceTiles / 2];
#pragma unroll
  for (int i = 0; i < kCeSliceTiles / 2; ++i) {       // the slice's 16 loads of this thread in flight together
    const int t = tile0 + 2 * i + half;
    raw[i] = t < tiles ? __ldg(reinterpret_cast<const uint4*>(base + size_t(t) * tile_bytes + row_off(r))) : make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int i = 0; i < kCeSliceTiles / 2; ++i) {
    const int tl = 2 * i + half, t = tile0 + tl;
    float x[8];
    unpack_bf16x8(raw[i], x);
    if (a.drop.thresh) drop_apply8(a.drop, uint32_t(t * 8 + g), uint32_t(col0), x);
    const float4 d = dl_s[tl * 8 + g];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[0][e] = fmaf(d.x, x[e], acc[0][e]); acc[1][e] = fmaf(d.y, x[e], acc[1][e]);
      acc[2][e] = fmaf(d.z, x[e], acc[2][e]); acc[3][e] = fmaf(d.w, x[e], acc[3][e]);
    }
  }
  // lanes l and l ^ 16 own the same chunk; then the 8 warps meet in shared memory
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[c][e] += __shfl_xor_sync(0xffffffffu, acc[c][e], 16);
      if (lane < 16) part[warp][lane][c * 8 + e] = acc[c][e];
    }
  __syncthreads();
  if (tid < 128) {       // thread -> (chunk, class, 4 columns)
    const int ch = tid >> 3, c = (tid >> 1) & 3, e0 = (tid & 1) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int w = 0; w < 8; ++w)
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] += part[w][ch][c * 8 + e0 + e];
    float* dst = a.g_wf + size_t(c) * a.drop.D + (is_img ? a.off_img : a.off_txt) + ch * din + kc * 8 + e0;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      red_add4(dst, make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) red_add(dst + e, v[e]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// shared pieces of the backward tile kernels (CTA = 256 threads, both warpgroups on the same rows)
// ---------------------------------------------------------------------------------------------------------------
struct BwCtx {
  int tid, w, q, lane;      // w: warpgroup (column / operand-kind split)
  uint32_t tmem, lane_base;
  uint64_t* bar; uint32_t ph;
  int rp;                   // p-mapping row
  int h, i, rs;             // s-mapping
};
__device__ __forceinline__ BwCtx make_bwctx(uint32_t tmem, uint64_t* bar) {
  BwCtx c;
  c.tid = threadIdx.x; c.w = c.tid >> 7; c.q = (c.tid >> 5) & 3; c.lane = c.tid & 31;
  c.tmem = tmem; c.lane_base = uint32_t(32 * c.q) << 16; c.bar = bar; c.ph = 0;
  c.rp = c.tid & 127;
  c.h = c.lane >> 4; c.i = c.lane & 15; c.rs = 64 * c.h + 16 * c.q + c.i;
  return c;
}
__device__ __forceinline__ void cta_sync_for_mma() {
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
}
__device__ __forceinline__ void cta_wait_mma(BwCtx& c) {
  mbar_wait(c.bar, c.ph);
  c.ph ^= 1;
  tc_fence_after_sync();
}
__device__ __forceinline__ void ld16f(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  tmem_ld16_nw(taddr, r);
  tmem_wait_ld();
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]);
}
__device__ __forceinline__ void st_chunks16(uint8_t* op, int row, int col0, const float (&v)[16]) {
  const float lo[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]};
  const float hi[8] = {v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]};
  *reinterpret_cast<uint4*>(op + uint32_t(col0 >> 3) * kCS + row_off(row)) = pack_bf16x8(lo);
  *reinterpret_cast<uint4*>(op + uint32_t((col0 >> 3) + 1) * kCS + row_off(row)) = pack_bf16x8(hi);
}
// f(c0, v[16]) for the accumulator columns [c_lo, c_hi) of my lane, 16 at a time; the next 16 are in flight
// while f runs (tcgen05.ld is asynchronous until tcgen05.wait::ld)
template <class F>
__device__ __forceinline__ void for_cols16(uint32_t taddr, int c_lo, int c_hi, F f) {
  if (c_lo >= c_hi) return;
  uint32_t nx[16];
  tmem_ld16_nw(taddr + c_lo, nx);
#pragma unroll 1
  for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
    tmem_wait_ld();
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(nx[e]);
    if (c0 + 16 < c_hi) tmem_ld16_nw(taddr + c0 + 16, nx);
    f(c0, v);
  }
}
// the same over two accumulators at once: f(c0, x[16], y[16])
template <class F>
__device__ __forceinline__ void for_cols16x2(uint32_t ta, uint32_t tb, int c_lo, int c_hi, F f) {
  if (c_lo >= c_hi) return;
  uint32_t na[16], nb[16];
  tmem_ld16_nw(ta + c_lo, na);
  tmem_ld16_nw(tb + c_lo, nb);
#pragma unroll 1
  for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
    tmem_wait_ld();
    float x[16], y[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) { x[e] = __uint_as_float(na[e]); y[e] = __uint_as_float(nb[e]); }
    if (c0 + 16 < c_hi) { tmem_ld16_nw(ta + c0 + 16, na); tmem_ld16_nw(tb + c0 + 16, nb); }
    f(c0, x, y);
  }
}
// accumulator columns [col + c_lo, col + c_hi) of `row`'s lane -> operand columns [c_lo, c_hi)
__device__ __forceinline__ void acc_cols_to_operand(const BwCtx& c, uint32_t col, int c_lo, int c_hi, uint8_t* op, int row) {
  for_cols16(c.tmem + c.lane_base + col, c_lo, c_hi, [&](int c0, const float (&v)[16]) { st_chunks16(op, row, c0, v); });
}
__device__ __forceinline__ void mbar_wait_ph(uint64_t* bar, uint32_t& ph) {
  mbar_wait(bar, ph);
  ph ^= 1;
  tc_fence_after_sync();
}

// DL[row (b,r)][(r',c)] = dlogits[b][c] [r' == r]: warpgroup w writes column groups 4w .. 4w+3 of the row
__device__ __forceinline__ float4 load_dl(const BwCtx& c, const float* __restrict__ dlogits, int b0, int batch) {
  const int b = b0 + (c.rp >> 4);
  return b < batch ? __ldg(reinterpret_cast<const float4*>(dlogits) + b) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void stage_dl(const BwCtx& c, uint8_t* dl, const float4 d) {
  const int r = c.rp & 15;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int kc = 4 * c.w + k;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (kc == (r >> 1)) {
      const uint32_t lo = pack_bf16x2(d.x, d.y), hi = pack_bf16x2(d.z, d.w);
      if (r & 1) { v.z = lo; v.w = hi; } else { v.x = lo; v.y = hi; }
    }
    *reinterpret_cast<uint4*>(dl + uint32_t(kc) * kCS + row_off(c.rp)) = v;
  }
}

// softmax backward on my row (s-mapping): p[] are the weights that multiplied V; ds[] = dL/dS
__device__ __forceinline__ void softmax_bwd16(const BwCtx& c, uint32_t col_dp, bool reverse, const float (&p)[16], float (&ds)[16]) {
  float dp[16];
  ld16f(c.tmem + c.lane_base + col_dp + 16 * c.q, dp);
  float dot0 = 0.f, dot1 = 0.f;
  float a[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    a[j] = reverse ? 1.0f - float(kL - 1) * p[j] : p[j];                 // A from (1 - A)/(L - 1)
    dp[j] = reverse ? dp[j] * (-1.0f / float(kL - 1)) : dp[j];           // dA
    if (j & 1) dot1 = fmaf(dp[j], a[j], dot1); else dot0 = fmaf(dp[j], a[j], dot0);
  }
  const float dot = dot0 + dot1;
#pragma unroll
  for (int j = 0; j < 16; ++j) ds[j] = a[j] * (dp[j] - dot);
}
// my row of a [64 x 64] half: warpgroup w writes the column groups of parity w (data group 2q+w, zeros elsewhere)
__device__ __forceinline__ void store_half_row_split(const BwCtx& c, uint8_t* buf, const float (&v)[16], bool zero_rest) {
  uint8_t* base = buf + c.h * kPHalf + row_off(16 * c.q + c.i);
  float x[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) x[e] = c.w ? v[8 + e] : v[e];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int kc = 2 * k + c.w;
    if (kc == 2 * c.q + c.w) *reinterpret_cast<uint4*>(base + kc * kPCS) = pack_bf16x8(x);
    else if (zero_rest) *reinterpret_cast<uint4*>(base + kc * kPCS) = make_uint4(0u, 0u, 0u, 0u);
  }
}
__device__ __forceinline__ void softmax16_bw(const BwCtx& c, uint32_t col_s, bool reverse, float (&p)[16]) {
  float s[16];
  ld16f(c.tmem + c.lane_base + col_s + 16 * c.q, s);
  float m0 = fmaxf(s[0], s[1]), m1 = fmaxf(s[2], s[3]);
#pragma unroll
  for (int j = 4; j < 16; j += 2) { m0 = fmaxf(m0, s[j]); m1 = fmaxf(m1, s[j + 1]); }
  const float m = fmaxf(m0, m1);
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 16; j += 2) { p[j] = __expf(s[j] - m); p[j + 1] = __expf(s[j + 1] - m); s0 += p[j]; s1 += p[j + 1]; }
  const float inv = 1.0f / (s0 + s1);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float a = p[j] * inv;
    p[j] = reverse ? (1.0f - a) * (1.0f / float(kL - 1)) : a;
  }
}
template <int DV>
__device__ __forceinline__ void ln_stats_bw(const BwCtx& c, uint32_t col, float& mean, float& rstd) {
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  for_cols16(c.tmem + c.lane_base + col, 0, DV, [&](int, const float (&x)[16]) {
#pragma unroll
    for (int e = 0; e < 16; e += 2) { s0 += x[e]; s1 += x[e + 1]; q0 = fmaf(x[e], x[e], q0); q1 = fmaf(x[e + 1], x[e + 1], q1); }
  });
  mean = (s0 + s1) * (1.0f / float(DV));
  const float var = fmaxf((q0 + q1) * (1.0f / float(DV)) - mean * mean, 0.f);
  rstd = rsqrtf(var + kLnEps);
}

// flush one persistent M=128 accumulator: lane = m.  dst(m, n) gives the address; columns split by warpgroup.
template <class F>
__device__ __forceinline__ void flush_acc(const BwCtx& c, uint32_t col, int ncols, int m_valid, F dst) {
  const int m = c.rp;
#pragma unroll 1
  for (int n0 = 16 * c.w; n0 < ncols; n0 += 32) {
    float v[16];
    ld16f(c.tmem + c.lane_base + col + n0, v);
    if (m < m_valid) {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float* p = dst(m, n0 + e);
        if (p) red_add(p, v[e]);
      }
    }
  }
}

__device__ __forceinline__ long long globaltimer_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define MMRCA_STAMP_NS(i) do { if (a.dbg && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) a.dbg[(i)] = globaltimer_ns(); } while (0)
#define MMRCA_STAMP(i) do { if (a.dbg && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && stamp_n + (i) < 256) a.dbg[stamp_n + (i)] = clock64(); } while (0)

// ---------------------------------------------------------------------------------------------------------------
// cross-attention backward, one direction per blockIdx.y
// ---------------------------------------------------------------------------------------------------------------
struct CaBwdDir {
  const void* blobs;              // bz | bv | bc
  const float* ln_g; const float* ln_b;
  float* gm;                      // [96][128] fp32: dM_ext^T (row n = k', column m = k; m = 96 holds du)
  float* g_wv; float* g_bv;       // [48][96], [48]
  float* g_ln_g; float* g_ln_b;   // [48]
  float* g_wf;                    // classifier rows of this source: &dWf[0][off], row stride D
  void* dxq_img; void* dxkv_img;  // [tiles][kSaTileBytes] out
  const uint4* row_in;            // [tiles][128] {keep bits lo, hi, mean, rstd} kept by the forward (ca_fwd_kernel)
};
struct CaBwdArgs {
  CaBwdDir dir[2];
  const void* t_tiles; const void* i_tiles;
  const float* dlogits;           // [B][4]
  int D;                          // concat width (row stride of dWf)
  DropSpec drop;                  // concat columns of direction d: [d * 768, d * 768 + 768)
  int batch, reverse;
  long long* dbg;                 // development: per-phase clock64 stamps of CTA (0, 0) (null in production)
};
struct CaBwdSmem {
  static constexpr uint32_t XQ = 0;                                  // [128 x 112]
  static constexpr uint32_t XKV = XQ + op_bytes(112);
  static constexpr uint32_t Z = XKV + op_bytes(112);                 // [128 x 96]
  static constexpr uint32_t V = Z + op_bytes(96);                    // [128 x 48]; later dV
  static constexpr uint32_t OUT = V + op_bytes(48);                  // [128 x 48]
  static constexpr uint32_t DC = OUT + op_bytes(48);                 // [128 x 48]
  static constexpr uint32_t DYX = DC + op_bytes(48);                 // [128 x 96]: dy*xhat | dy; later dZ
  static constexpr uint32_t DLS = DYX + op_bytes(96);                // DL [128 x 64]; later dS (2 x [64 x 64])
  static constexpr uint32_t P = DLS + 2 * kPHalf;                    // 2 x [64 x 64]
  static constexpr uint32_t ONES = P + 2 * kPHalf;                   // 512 B of ones: one K = 16 slice [2 k-groups][16 rows][8] of the all-ones
                                                                     // B operand of the column-sum MMAs; every K-step re-reads it
  static constexpr uint32_t W = al128(ONES + 512);                   // bz | bv | bc | bv (lo)
  static constexpr uint32_t LN = W + CaCfg::W_BYTES_BWD;             // gamma, beta [48] fp32
  static constexpr uint32_t PART = LN + 2 * 48 * 4;                  // [2 warpgroups][128 rows] (m1, m2) partial sums
  static constexpr uint32_t BAR = al128(PART + 2 * 128 * 8);
  static constexpr uint32_t BYTES = BAR + 128;
  static_assert(2 * kPHalf >= op_bytes(64), "dS aliases DL");
  static_assert(BYTES <= 232448, "CA backward does not fit shared memory");
};
struct CaBwdCols {   // TMEM columns
  static constexpr uint32_t Z = 0, V = 96, DOUT = 144, S = 0, C = 64, DP = 0, DV = 64, DZ = 112, DXQ = 0, DXKV = 96;
  static constexpr uint32_t G_M = 256, G_WV = 352, G_WF = 400, G_LN = 464;
};

// 8 worker warps + 1 loader warp.  The loader waits for the previous tile's dM / dWv MMAs (the last readers of the
// input images) and then asks the bulk-copy engine for the next tile, so the images land while the workers are still
// writing the previous tile's input gradients; the workers synchronise among themselves on a named barrier.
constexpr int kCaBwdThreads = kCtaThreads + 32;
__device__ __forceinline__ void wk_sync() { named_bar_sync(1, kCtaThreads); }
__device__ __forceinline__ void wk_sync_for_mma() {
  fence_proxy_async();
  tc_fence_before_sync();
  wk_sync();
  tc_fence_after_sync();
}
__global__ void __launch_bounds__(kCaBwdThreads, 1) ca_bwd_kernel(const CaBwdArgs a) {
  extern __shared__ __align__(128) uint8_t sm[];
  using S = CaBwdSmem; using T = CaBwdCols; using C = CaCfg;
  // mbarriers: [0] weights, [1] MMAs whose result is read back next, [2] tile load,
  //            [3] dWf / dgamma|dbeta MMAs (G2), [4] dM / dWv MMAs (G3): nobody reads those results until the flush,
  //            they are only waited for before one of their operand buffers is overwritten,
  //            [5] the second half of a split MMA phase: its read-back overlaps the first half's
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float* ln_s = reinterpret_cast<float*>(sm + S::LN);
  float2* part = reinterpret_cast<float2*>(sm + S::PART);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int d = blockIdx.y;
  const CaBwdDir& D = a.dir[d];
  const bool reverse = a.reverse != 0;
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(&bars[0], C::W_BYTES_BWD);
    bulk_g2s(sm + S::W, D.blobs, C::W_BYTES_BWD, &bars[0]);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < 48; i += kCaBwdThreads) { ln_s[i] = D.ln_g[i]; ln_s[48 + i] = D.ln_b[i]; }
  for (uint32_t i = tid; i < (2 * kPHalf + 512) / 16; i += kCaBwdThreads) {      // P zeros, then the ones operand
    const uint32_t off = i * 16;
    reinterpret_cast<uint4*>(sm + S::P)[i] =
        off < 2 * kPHalf ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  }
  if (tid < 128) { write_bias_columns(sm + S::XQ, C::DIN / 8, tid); write_bias_columns(sm + S::XKV, C::DIN / 8, tid); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  mbar_wait(&bars[0], 0);
  BwCtx c = make_bwctx(tmem, &bars[1]);
  uint32_t ph_ld = 0, ph_g2 = 0, ph_g3 = 0, ph_b = 0;
  uint8_t *xq = sm + S::XQ, *xkv = sm + S::XKV, *zb = sm + S::Z, *vb = sm + S::V, *ob = sm + S::OUT, *dcb = sm + S::DC,
          *dyx = sm + S::DYX, *dls = sm + S::DLS, *pb = sm + S::P, *ones = sm + S::ONES, *wsm = sm + S::W;
  const uint8_t* bz = wsm; const uint8_t* bv = wsm + C::OFF_BV; const uint8_t* bc = wsm + C::OFF_BC;
  const uint8_t* bv_lo = wsm + C::OFF_BV_LO;
  const int tiles = (a.batch + 7) / 8;
  if (warp == kCtaThreads / 32) {      // ---- loader warp ----------------------------------------------------------------
    uint32_t ph = 0, phb = 0;
    for (int tile = blockIdx.x, it = 0; tile < tiles; tile += gridDim.x, ++it) {
      if (it > 0) { mbar_wait(&bars[4], ph); ph ^= 1; }
      if ((tid & 31) == 0) {
        const uint8_t* qsrc = static_cast<const uint8_t*>(d == 0 ? a.t_tiles : a.i_tiles) + size_t(tile) * kSaTileBytes;
        const uint8_t* ksrc = static_cast<const uint8_t*>(d == 0 ? a.i_tiles : a.t_tiles) + size_t(tile) * kSaTileBytes;
        mbar_arrive_expect_tx(&bars[2], 2 * kSaTileBytes);
        bulk_g2s(xq, qsrc, kSaTileBytes, &bars[2]);
        bulk_g2s(xkv, ksrc, kSaTileBytes, &bars[2]);
        if (tile + int(gridDim.x) < tiles) {      // the tile after: into the L2 meanwhile
          bulk_prefetch_l2(qsrc + size_t(gridDim.x) * kSaTileBytes, kSaTileBytes);
          bulk_prefetch_l2(ksrc + size_t(gridDim.x) * kSaTileBytes, kSaTileBytes);
        }
      }
      __syncwarp();
      // The persistent dM / dWv MMAs (G3) are issued from here, once the tile's last read-back MMAs (dXkv, third
      // completion of bars[5] in a tile) are done: tcgen05.mma issue blocks while the tensor pipe's queue is full, so
      // issued by the workers' thread 0 they held its warp - and with it the end-of-tile barrier - for their whole run.
      for (int k = 0; k < 3; ++k) { mbar_wait(&bars[5], phb); phb ^= 1; }
      tc_fence_after_sync();
      if ((tid & 31) == 0) {
        mma_steps(tmem + T::G_M, make_smem_desc(smem_u32(xq), kRS, kCS), 2 * kRS, make_smem_desc(smem_u32(dyx), kRS, kCS), 2 * kRS,
                  make_idesc_bf16(128, C::DIN, 1, 1), 8, it > 0);
        mma_steps(tmem + T::G_WV, make_smem_desc(smem_u32(xkv), kRS, kCS), 2 * kRS, make_smem_desc(smem_u32(vb), kRS, kCS), 2 * kRS,
                  make_idesc_bf16(128, C::DV, 1, 1), 8, it > 0);
        umma_commit(&bars[4]);
      }
      __syncwarp();
    }
  } else {                             // ---- worker warps ---------------------------------------------------------------
  bool first = true;
  int ntiles = 0;
  float4 dl_next = make_float4(0.f, 0.f, 0.f, 0.f);
  uint4 row_next = make_uint4(0u, 0u, 0u, 0u);
  if (int(blockIdx.x) < tiles) {
    dl_next = load_dl(c, a.dlogits, blockIdx.x * 8, a.batch);
    row_next = __ldg(D.row_in + size_t(blockIdx.x) * 128 + c.rs);
  }
  int stamp_n = 0;
  if (a.dbg && tid == 0) a.dbg[300 + 2 * (blockIdx.x + gridDim.x * blockIdx.y)] = globaltimer_ns();
  MMRCA_STAMP(0); stamp_n = 1;
  MMRCA_STAMP_NS(250);
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, stamp_n += 16) {
    MMRCA_STAMP(0);
    // ---- P0: DL from dlogits; block inputs (SA images): requested by the loader warp once the previous tile's dM / dWv
    //      MMAs were done, which also makes their operand buffers (dZ, dV) free to overwrite below -------------------------
    stage_dl(c, dls, dl_next);
    const uint4 row_cur = row_next;      // my row's dropout keep bits and LayerNorm statistics from the forward
    if (tile + int(gridDim.x) < tiles) {      // a tile ahead
      dl_next = load_dl(c, a.dlogits, (tile + int(gridDim.x)) * 8, a.batch);
      row_next = __ldg(D.row_in + size_t(tile + int(gridDim.x)) * 128 + c.rs);
    }
    mbar_wait(&bars[2], ph_ld); ph_ld ^= 1;
    wk_sync_for_mma();
    MMRCA_STAMP(1);
    // ---- P1: Z, V (M=128) and dOut = DL Wf_src^T (two M=64 halves, s-mapping like C) ----------------------------
    if (tid == 0) {
      mma_steps(tmem + T::Z, make_smem_desc(smem_u32(xq), kCS, kRS), 2 * kCS, make_smem_desc(smem_u32(bz), C::BZ_LBO, 128),
                2 * C::BZ_LBO, make_idesc_bf16(128, C::DIN, 0, 0), C::KE / 16, false);
      umma_commit(c.bar);
      mma_steps(tmem + T::V, make_smem_desc(smem_u32(xkv), kCS, kRS), 2 * kCS, make_smem_desc(smem_u32(bv), C::BV_LBO, 128),
                2 * C::BV_LBO, make_idesc_bf16(128, C::DV, 0, 0), C::KE / 16, false);
      mma_steps(tmem + T::V, make_smem_desc(smem_u32(xkv), kCS, kRS), 2 * kCS, make_smem_desc(smem_u32(bv_lo), C::BV_LBO, 128),
                2 * C::BV_LBO, make_idesc_bf16(128, C::DV, 0, 0), C::KE / 16, true);      // the forward's hi + lo W_value
      for (int h = 0; h < 2; ++h)
        mma_steps(tmem + (uint32_t(16 * h) << 16) + T::DOUT, make_smem_desc(smem_u32(dls + h * 8 * kRS), kCS, kRS), 2 * kCS,
                  make_smem_desc(smem_u32(bc), 128, C::BC_LBO), 2 * 128, make_idesc_bf16(64, C::DV, 0, 1), kNCls / 16, false);
      umma_commit(&bars[5]);
    }
    cta_wait_mma(c);
    MMRCA_STAMP(2);
    // ---- P2: Z, V -> operands (columns split); Z is converted while the V / dOut MMAs run --------------------------
    acc_cols_to_operand(c, T::Z, 48 * c.w, 48 * c.w + 48, zb, c.rp);
    mbar_wait_ph(&bars[5], ph_b);
    if (c.w == 0) acc_cols_to_operand(c, T::V, 0, 32, vb, c.rp); else acc_cols_to_operand(c, T::V, 32, 48, vb, c.rp);
    wk_sync_for_mma();
    MMRCA_STAMP(3);
    // ---- P3: scores ----------------------------------------------------------------------------------------------------
    if (tid == 0) {
      for (int h = 0; h < 2; ++h)
        mma_steps(tmem + (uint32_t(16 * h) << 16) + T::S, make_smem_desc(smem_u32(zb + h * 8 * kRS), kCS, kRS), 2 * kCS,
                  make_smem_desc(smem_u32(xkv + h * 8 * kRS), kCS, kRS), 2 * kCS, make_idesc_bf16(64, 64, 0, 0), C::DIN / 16, false);
      umma_commit(c.bar);
    }
    cta_wait_mma(c);
    MMRCA_STAMP(4);
    float p[16];
    softmax16_bw(c, T::S, reverse, p);
    store_half_row_split(c, pb, p, false);
    wk_sync_for_mma();
    MMRCA_STAMP(5);
    // ---- P5: context ---------------------------------------------------------------------------------------------------
    if (tid == 0) {
      for (int h = 0; h < 2; ++h)
        mma_steps(tmem + (uint32_t(16 * h) << 16) + T::C, make_smem_desc(smem_u32(pb + h * kPHalf), kPCS, kRS), 2 * kPCS,
                  make_smem_desc(smem_u32(vb + h * 8 * kRS), kRS, kCS), 2 * kRS, make_idesc_bf16(64, C::DV, 0, 1), 4, false);
      umma_commit(c.bar);
    }
    cta_wait_mma(c);
    MMRCA_STAMP(6);
    // ---- P6: LayerNorm / ReLU backward.  Both warpgroups own the same rows and split a row's 48 columns (24 each):
    //      mean / rstd and the dropout keep bits are the forward's, everything else on my 24 columns, xhat and dxhat kept in
    //      registers between the two passes; the row sums m1 = mean(dxhat), m2 = mean(dxhat xhat) meet in shared memory.
    {
      constexpr int HC = C::DV / 2;
      const float mean = __uint_as_float(row_cur.z), rstd = __uint_as_float(row_cur.w), nmr = -mean * rstd;
      uint32_t xr[HC], gr[HC];
      {
        const uint32_t tc_ = c.tmem + c.lane_base + T::C + HC * c.w, tg_ = c.tmem + c.lane_base + T::DOUT + HC * c.w;
#pragma unroll
        for (int j = 0; j < HC / 8; ++j) {
          tmem_ld8_nw(tc_ + 8 * j, *reinterpret_cast<uint32_t(*)[8]>(&xr[8 * j]));
          tmem_ld8_nw(tg_ + 8 * j, *reinterpret_cast<uint32_t(*)[8]>(&gr[8 * j]));
        }
      }
      // self.drop (multimodal_model.py:719) sits between this block's output and the classifier: the keep bits of
      // my concat columns scale both what the classifier saw (Out, for dWf) and what it sends back (dOut)
      const bool dropping = a.drop.thresh != 0;
      const uint32_t keep = uint32_t((uint64_t(row_cur.x) | (uint64_t(row_cur.y) << 32)) >> (HC * c.w));
      const float dscale = dropping ? a.drop.scale : 1.0f;
      tmem_wait_ld();
      float xh[HC], dxh[HC];
      float m1a = 0.f, m1b = 0.f, m2a = 0.f, m2b = 0.f;
      const float* gam_s = ln_s + HC * c.w;
      const float* bet_s = ln_s + C::DV + HC * c.w;
#pragma unroll
      for (int j = 0; j < HC / 8; ++j) {
        float o[8], t1[8], t2[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int k = 8 * j + e;
          const float gam = gam_s[k];
          xh[k] = fmaf(__uint_as_float(xr[k]), rstd, nmr);
          const float y = fmaf(xh[k], gam, bet_s[k]);
          const float mk = (keep >> k) & 1u ? dscale : 0.f;
          const float dy = y > 0.f ? __uint_as_float(gr[k]) * mk : 0.f;
          dxh[k] = dy * gam;
          if (e & 1) { m1b += dxh[k]; m2b = fmaf(dxh[k], xh[k], m2b); } else { m1a += dxh[k]; m2a = fmaf(dxh[k], xh[k], m2a); }
          o[e] = fmaxf(y, 0.f) * mk; t1[e] = dy * xh[k]; t2[e] = dy;
        }
        const uint32_t kc = uint32_t(HC / 8 * c.w + j);
        *reinterpret_cast<uint4*>(ob + kc * kCS + row_off(c.rs)) = pack_bf16x8(o);
        *reinterpret_cast<uint4*>(dyx + kc * kCS + row_off(c.rs)) = pack_bf16x8(t1);
        *reinterpret_cast<uint4*>(dyx + (C::DV / 8 + kc) * kCS + row_off(c.rs)) = pack_bf16x8(t2);
      }
      part[c.w * 128 + c.rs] = make_float2(m1a + m1b, m2a + m2b);
      wk_sync();
      const float2 other = part[(c.w ^ 1) * 128 + c.rs];
      const float m1 = (m1a + m1b + other.x) * (1.0f / float(C::DV)), m2 = (m2a + m2b + other.y) * (1.0f / float(C::DV));
#pragma unroll
      for (int j = 0; j < HC / 8; ++j) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = rstd * (dxh[8 * j + e] - m1 - xh[8 * j + e] * m2);
        *reinterpret_cast<uint4*>(dcb + uint32_t(HC / 8 * c.w + j) * kCS + row_off(c.rs)) = pack_bf16x8(o);
      }
    }
    wk_sync_for_mma();
    MMRCA_STAMP(7);
    // ---- P7: dP (read back next); classifier-weight and LayerNorm-affine gradients (persistent, group G2) ------
    if (tid == 0) {
      for (int h = 0; h < 2; ++h)
        mma_steps(tmem + (uint32_t(16 * h) << 16) + T::DP, make_smem_desc(smem_u32(dcb + h * 8 * kRS), kCS, kRS), 2 * kCS,
                  make_smem_desc(smem_u32(vb + h * 8 * kRS), kCS, kRS), 2 * kCS, make_idesc_bf16(64, 64, 0, 0), C::DV / 16, false);
      umma_commit(c.bar);
      mma_steps(tmem + T::G_WF, make_smem_desc(smem_u32(ob), kRS, kCS), 2 * kRS, make_smem_desc(smem_u32(dls), kRS, kCS), 2 * kRS,
                make_idesc_bf16(64, kNCls, 1, 1), 8, !first);
      mma_steps(tmem + T::G_LN, make_smem_desc(smem_u32(dyx), kRS, kCS), 2 * kRS, make_smem_desc(smem_u32(ones), 256, 128), 0,
                make_idesc_bf16(128, 16, 1, 0), 8, !first);
      umma_commit(&bars[3]);
    }
    cta_wait_mma(c);
    MMRCA_STAMP(8);
    // ---- P8: softmax backward -> dS (reuses DL's bytes once G2 has read them) ----------------------------------------
    {
      float ds[16];
      softmax_bwd16(c, T::DP, reverse, p, ds);
      mbar_wait_ph(&bars[3], ph_g2);
      store_half_row_split(c, dls, ds, true);
    }
    wk_sync_for_mma();
    MMRCA_STAMP(9);
    // ---- P9: dZ = dS Xkv (converted while the next MMAs run), dV = P^T dC ---------------------------------------------
    if (tid == 0) {
      for (int h = 0; h < 2; ++h)
        mma_steps(tmem + (uint32_t(16 * h) << 16) + T::DZ, make_smem_desc(smem_u32(dls + h * kPHalf), kPCS, kRS), 2 * kPCS,
                  make_smem_desc(smem_u32(xkv + h * 8 * kRS), kRS, kCS), 2 * kRS, make_idesc_bf16(64, C::DIN, 0, 1), 4, false);
      umma_commit(c.bar);
      for (int h = 0; h < 2; ++h)
        mma_steps(tmem + (uint32_t(16 * h) << 16) + T::DV, make_smem_desc(smem_u32(pb + h * kPHalf), kRS, kPCS), 2 * kRS,
                  make_smem_desc(smem_u32(dcb + h * 8 * kRS), kRS, kCS), 2 * kRS, make_idesc_bf16(64, C::DV, 1, 1), 4, false);
      umma_commit(&bars[5]);
    }
    cta_wait_mma(c);
    MMRCA_STAMP(10);
    // ---- P10: dZ, dV -> operands (dZ over dy*xhat | dy, dV over V) ------------------------------------------------
    acc_cols_to_operand(c, T::DZ, 48 * c.w, 48 * c.w + 48, dyx, c.rs);
    mbar_wait_ph(&bars[5], ph_b);
    if (c.w == 0) acc_cols_to_operand(c, T::DV, 0, 32, vb, c.rs); else acc_cols_to_operand(c, T::DV, 32, 48, vb, c.rs);
    wk_sync_for_mma();
    MMRCA_STAMP(11);
    // ---- P11: gradients of the block inputs (read back next), then the parameter gradients (persistent, G3) ------
    if (tid == 0) {
      // dXq = dZ M^T  (B: the Z blob read along its other axis)
      mma_steps(tmem + T::DXQ, make_smem_desc(smem_u32(dyx), kCS, kRS), 2 * kCS, make_smem_desc(smem_u32(bz), 128, C::BZ_LBO), 2 * 128,
                make_idesc_bf16(128, C::DIN, 0, 1), C::DIN / 16, false);
      umma_commit(c.bar);
      // dXkv = dS^T Z + dV Wv  (runs while dXq is read back)
      for (int h = 0; h < 2; ++h) {
        mma_steps(tmem + (uint32_t(16 * h) << 16) + T::DXKV, make_smem_desc(smem_u32(dls + h * kPHalf), kRS, kPCS), 2 * kRS,
                  make_smem_desc(smem_u32(zb + h * 8 * kRS), kRS, kCS), 2 * kRS, make_idesc_bf16(64, C::DIN, 1, 1), 4, false);
        mma_steps(tmem + (uint32_t(16 * h) << 16) + T::DXKV, make_smem_desc(smem_u32(vb + h * 8 * kRS), kCS, kRS), 2 * kCS,
                  make_smem_desc(smem_u32(bv), 128, C::BV_LBO), 2 * 128, make_idesc_bf16(64, C::DIN, 0, 1), C::DV / 16, true);
      }
      umma_commit(&bars[5]);
    }
    cta_wait_mma(c);
    MMRCA_STAMP(12);
    // ---- P12: input gradients -> bf16 images for the SA backward ---------------------------------------------------
    {
      uint8_t* gq = static_cast<uint8_t*>(D.dxq_img) + size_t(tile) * kSaTileBytes;
      uint8_t* gk = static_cast<uint8_t*>(D.dxkv_img) + size_t(tile) * kSaTileBytes;
      // (global images use the same [column group][row] geometry as the shared-memory operands)
      acc_cols_to_operand(c, T::DXQ, 48 * c.w, 48 * c.w + 48, gq, c.rp);
      mbar_wait_ph(&bars[5], ph_b);
      acc_cols_to_operand(c, T::DXKV, 48 * c.w, 48 * c.w + 48, gk, c.rs);
    }
    first = false;
    ++ntiles;
    MMRCA_STAMP(13);
    tc_fence_before_sync();
    wk_sync();
    tc_fence_after_sync();
  }
  MMRCA_STAMP(0);
  // ---- flush the persistent accumulators ---------------------------------------------------------------------------
  if (!first) {
    ph_g3 = uint32_t(ntiles - 1) & 1u;
    mbar_wait_ph(&bars[4], ph_g3);
    flush_acc(c, T::G_M, C::DIN, C::DIN + 1, [&](int m, int n) { return D.gm + size_t(n) * 128 + m; });
    flush_acc(c, T::G_WV, C::DV, C::DIN + 1,
              [&](int m, int n) { return m < C::DIN ? D.g_wv + size_t(n) * C::DIN + m : D.g_bv + n; });
    // dWf^T: an M=64 accumulator keeps row m = 16q + lane in the lower half of each warp's lanes
    {
      const int j = 16 * c.q + c.lane;
#pragma unroll 1
      for (int n0 = 16 * c.w; n0 < kNCls; n0 += 32) {
        float v[16];
        ld16f(c.tmem + c.lane_base + T::G_WF + n0, v);
        if (c.lane < 16 && j < C::DV) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int n = n0 + e, rp = n >> 2, cls = n & 3;
            red_add(D.g_wf + size_t(cls) * a.D + rp * C::DV + j, v[e]);
          }
        }
      }
    }
    if (c.w == 0) {
      float v[16];
      ld16f(c.tmem + c.lane_base + T::G_LN, v);
      if (c.rp < C::DV) red_add(D.g_ln_g + c.rp, v[0]);
      else if (c.rp < 2 * C::DV) red_add(D.g_ln_b + (c.rp - C::DV), v[0]);
    }
  }
  wk_sync();
  MMRCA_STAMP(1);
  MMRCA_STAMP_NS(251);
  if (a.dbg && tid == 0) a.dbg[301 + 2 * (blockIdx.x + gridDim.x * blockIdx.y)] = globaltimer_ns();
  }   // workers
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// self-attention backward (features frozen: parameter gradients only), one modality per blockIdx.y.
// Nothing of the forward is recomputed except C = P V (one MMA, so that the LayerNorm backward sees exactly the
// fp32 context the forward normalised): X comes from prep_feat_kernel's image, V and P from the images the forward
// kept, all by TMA, one whole tile ahead into the other half of a double buffer.
// ---------------------------------------------------------------------------------------------------------------
struct SaBwdArgs {
  const void* x_tiles;            // [tiles][x_tile_bytes(DIN)]: normalised bf16 operand images (prep_feat_kernel)
  const void* v_tiles;            // [tiles][op_bytes(96)]  V operand images (sa_fwd_kernel)
  const void* p_tiles;            // [tiles][2 * kPHalf]    attention weights (sa_fwd_kernel)
  const float2* ln_stats;         // [tiles][128] (mean, rstd) of the forward's context rows (sa_fwd_kernel)
  const float* ln_g; const float* ln_b;
  const void* dout_a; const void* dout_b;   // dOut = a + b (images written by ca_bwd_kernel)
  float* gm;                      // [DIN][128]
  float* g_wv; float* g_bv; float* g_ln_g; float* g_ln_b;
  int batch;
  long long* dbg;                 // development: per-phase clock64 stamps of CTA 0 (null in production)
  // MMRCA_FLAG_FEATURE_GRADS (fine-tune phase): the tile's dZ, dS, dV operands are also left in HBM for sa_dx_kernel
  // (mmrca_head_tc_dx.cuh), which turns them into feature gradients; null: not kept
  void* dz_tiles; void* ds_tiles; void* dv_tiles;
};
template <int DIN_>
struct SaBwdSmem {
  using C = SaCfg<DIN_>;
  static constexpr uint32_t XB = op_bytes(C::KE), VB = op_bytes(96), PB = 2 * kPHalf;
  static constexpr uint32_t X = 0;                                   // 2 x [128 x (DIN+16)]   (TMA, double buffer)
  static constexpr uint32_t V = X + 2 * XB;                          // 2 x [128 x 96]         (TMA, double buffer)
  static constexpr uint32_t P = V + 2 * VB;                          // 2 x 2 x [64 x 64]      (TMA, double buffer)
  static constexpr uint32_t DC = P + 2 * PB;                         // [128 x 96] dC; later dV
  static constexpr uint32_t DYX = DC + op_bytes(96);                 // [128 x 192]: dOut images (TMA), then dy*xhat | dy, then dZ
  static constexpr uint32_t DLS = DYX + op_bytes(192);               // dS (2 x [64 x 64])
  static constexpr uint32_t ONES = DLS + 2 * kPHalf;
  static constexpr uint32_t LN = al128(ONES + 4096);                 // gamma, beta [96] fp32
  static constexpr uint32_t PART = LN + 2 * 96 * 4;                  // [2 warpgroups][128 rows] (m1, m2) partial sums
  static constexpr uint32_t BAR = al128(PART + 2 * 128 * 8);
  static constexpr uint32_t BYTES = BAR + 128;
  static_assert(BYTES <= 232448, "SA backward does not fit shared memory");
};
template <int DIN_>
struct SaBwdCols {
  static constexpr uint32_t C = 64, DP = 0, DV = 0, DZ = 96;                        // working: [0, 176)
  static constexpr uint32_t G_M = 176, G_WV = G_M + DIN_, G_LN = G_WV + 96;       // G_LN: 2 x 16 columns
  static_assert(G_LN + 32 <= 512, "TMEM budget");
};

template <int DIN_>
__device__ __forceinline__ void sa_bwd_body(const SaBwdArgs& a, uint8_t* sm) {
  using S = SaBwdSmem<DIN_>; using T = SaBwdCols<DIN_>; using C = SaCfg<DIN_>;
  constexpr int DIN = DIN_, DV = C::DV;
  // mbarriers: [0], [1] tile inputs X | V | P of buffer 0 / 1 (TMA), [2] MMAs read back next, [3] unused,
  //            [4] G2: dgamma|dbeta, [5] G3: dM, dWv  (G*: persistent accumulators, waited for only before an operand
  //            buffer is reused).  Thread 0 issues the MMAs, thread 32 (another warp) the TMA loads.
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float* ln_s = reinterpret_cast<float*>(sm + S::LN);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (a.dbg && tid == 0) a.dbg[300 + 2 * blockIdx.x] = globaltimer_ns();
  const int tiles = (a.batch + 7) / 8;
  constexpr uint32_t kXBytes = x_tile_bytes(DIN);
  auto issue_inputs = [&](int t, int buf) {
    mbar_arrive_expect_tx(&bars[buf], kXBytes + S::VB + S::PB);
    bulk_g2s(sm + S::X + buf * S::XB, static_cast<const uint8_t*>(a.x_tiles) + size_t(t) * kXBytes, kXBytes, &bars[buf]);
    bulk_g2s(sm + S::V + buf * S::VB, static_cast<const uint8_t*>(a.v_tiles) + size_t(t) * S::VB, S::VB, &bars[buf]);
    bulk_g2s(sm + S::P + buf * S::PB, static_cast<const uint8_t*>(a.p_tiles) + size_t(t) * S::PB, S::PB, &bars[buf]);
  };
  if (tid == 32) {
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
    if (int(blockIdx.x) < tiles) issue_inputs(blockIdx.x, 0);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  float2* part = reinterpret_cast<float2*>(sm + S::PART);
  for (int i = tid; i < DV; i += kCtaThreads) { ln_s[i] = a.ln_g[i]; ln_s[DV + i] = a.ln_b[i]; }
  for (uint32_t i = tid; i < 4096 / 16; i += kCtaThreads)
    reinterpret_cast<uint4*>(sm + S::ONES)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  BwCtx c = make_bwctx(tmem, &bars[2]);
  uint32_t ph_in[2] = {0, 0}, ph_g2 = 0, ph_g3 = 0;
  uint8_t *dcb = sm + S::DC, *dyx = sm + S::DYX, *dls = sm + S::DLS, *ones = sm + S::ONES;
  bool first = true;
  int buf = 0, stamp_n = 0;
  MMRCA_STAMP(0); stamp_n = 1;
  MMRCA_STAMP_NS(250);
  // my share of dOut = image a + image b (written by the CA backward): row rs, columns [48w, 48w + 48), and the forward's
  // LayerNorm statistics of the row, straight from L2 into registers.  They are requested most of a tile ahead (during
  // the previous tile's dV / dZ MMAs): their latency is never waited for.
  uint4 ra[6], rb[6];
  float2 st;
  auto load_dout = [&](int t) {
    const uint8_t* ga = static_cast<const uint8_t*>(a.dout_a) + size_t(t) * kSaTileBytes + uint32_t(6 * c.w) * kCS + row_off(c.rs);
    const uint8_t* gb = static_cast<const uint8_t*>(a.dout_b) + size_t(t) * kSaTileBytes + uint32_t(6 * c.w) * kCS + row_off(c.rs);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      ra[j] = __ldg(reinterpret_cast<const uint4*>(ga + j * kCS));
      rb[j] = __ldg(reinterpret_cast<const uint4*>(gb + j * kCS));
    }
    st = __ldg(a.ln_stats + size_t(t) * 128 + c.rs);
  };
  if (int(blockIdx.x) < tiles) load_dout(blockIdx.x);
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, buf ^= 1, stamp_n += 12) {
    uint8_t *xb = sm + S::X + buf * S::XB, *vb = sm + S::V + buf * S::VB, *pb = sm + S::P + buf * S::PB;
    MMRCA_STAMP(0);
    // ---- the previous tile's dM / dWv MMAs (G3) still read its X, dV (in dC's buffer) and dZ: once they are done,
    //      fetch the NEXT tile's inputs into the other buffers ------------------------------------------------------
    if (!first) mbar_wait_ph(&bars[5], ph_g3);
    MMRCA_STAMP(1);
    if (tid == 32 && tile + int(gridDim.x) < tiles) issue_inputs(tile + int(gridDim.x), buf ^ 1);
    if (tid == 64) {        // L2 prefetch: dOut images of the next tile (plain loads), TMA inputs of the one after
      const int t1 = tile + int(gridDim.x), t2 = t1 + int(gridDim.x);
      if (t1 < tiles) {
        bulk_prefetch_l2(static_cast<const uint8_t*>(a.dout_a) + size_t(t1) * kSaTileBytes, kSaTileBytes);
        bulk_prefetch_l2(static_cast<const uint8_t*>(a.dout_b) + size_t(t1) * kSaTileBytes, kSaTileBytes);
      }
      if (t2 < tiles) {
        bulk_prefetch_l2(static_cast<const uint8_t*>(a.x_tiles) + size_t(t2) * kXBytes, kXBytes);
        bulk_prefetch_l2(static_cast<const uint8_t*>(a.v_tiles) + size_t(t2) * S::VB, S::VB);
        bulk_prefetch_l2(static_cast<const uint8_t*>(a.p_tiles) + size_t(t2) * S::PB, S::PB);
      }
    }
    mbar_wait(&bars[buf], ph_in[buf]); ph_in[buf] ^= 1;
    tc_fence_after_sync();
    MMRCA_STAMP(2);
    // ---- context C = P V (two M=64 halves), exactly the forward's MMA ----------------------------------------------
    if (tid == 0) {
      for (int h = 0; h < 2; ++h)
        mma_steps(tmem + (uint32_t(16 * h) << 16) + T::C, make_smem_desc(smem_u32(pb + h * kPHalf), kPCS, kRS), 2 * kPCS,
                  make_smem_desc(smem_u32(vb + h * 8 * kRS), kRS, kCS), 2 * kRS, make_idesc_bf16(64, DV, 0, 1), 4, false);
      umma_commit(c.bar);
    }
    // my row of P (the forward's bf16 weights): chunks 2q, 2q+1 of my half
    float p[16];
    {
      const uint8_t* pr = pb + c.h * kPHalf + row_off(16 * c.q + c.i);
      float lo[8], hi[8];
      unpack_bf16x8(*reinterpret_cast<const uint4*>(pr + (2 * c.q) * kPCS), lo);
      unpack_bf16x8(*reinterpret_cast<const uint4*>(pr + (2 * c.q + 1) * kPCS), hi);
#pragma unroll
      for (int e = 0; e < 8; ++e) { p[e] = lo[e]; p[8 + e] = hi[e]; }
    }
    cta_wait_mma(c);
    MMRCA_STAMP(3);
    // ---- LayerNorm / ReLU backward.  The two warpgroups split the row's columns; the row sums m1 = mean(dxhat),
    //      m2 = mean(dxhat xhat) meet in shared memory.  xhat and dy of my 48 columns stay in registers between the
    //      two passes; mean / rstd are the forward's. ---------------------------------------------------------------
    {
      constexpr int HC = DV / 2;       // 48 columns per thread
      const float mean = st.x, rstd = st.y, nmr = -mean * rstd;
      float xh[HC], dy[HC];
      MMRCA_STAMP(4);
      {
        uint32_t raw[HC];
        const uint32_t tc_ = c.tmem + c.lane_base + T::C + HC * c.w;
#pragma unroll
        for (int j = 0; j < HC / 16; ++j) tmem_ld16_nw(tc_ + 16 * j, *reinterpret_cast<uint32_t(*)[16]>(&raw[16 * j]));
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          float fa[8], fb[8];
          unpack_bf16x8(ra[j], fa); unpack_bf16x8(rb[j], fb);
#pragma unroll
          for (int e = 0; e < 8; ++e) dy[8 * j + e] = fa[e] + fb[e];
        }
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < HC; ++e) xh[e] = fmaf(__uint_as_float(raw[e]), rstd, nmr);
      }
      float m1a = 0.f, m1b = 0.f, m2a = 0.f, m2b = 0.f;
      const float4* g4 = reinterpret_cast<const float4*>(ln_s + HC * c.w);
      const float4* b4 = reinterpret_cast<const float4*>(ln_s + DV + HC * c.w);
#pragma unroll
      for (int j = 0; j < HC / 4; ++j) {
        const float4 gq = g4[j], bq = b4[j];
        const float gam[4] = {gq.x, gq.y, gq.z, gq.w}, bet[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = 4 * j + e;
          const float y = fmaf(xh[k], gam[e], bet[e]);
          dy[k] = y > 0.f ? dy[k] : 0.f;
          const float dxh = dy[k] * gam[e];
          if (e & 1) { m1b += dxh; m2b = fmaf(dxh, xh[k], m2b); } else { m1a += dxh; m2a = fmaf(dxh, xh[k], m2a); }
        }
      }
      part[c.w * 128 + c.rs] = make_float2(m1a + m1b, m2a + m2b);
      if (tid == 96 && a.dz_tiles) bulk_wait_read();      // the previous tile's exported dZ / dS / dV have left shared memory
      __syncthreads();                                    // (every write to their buffers comes after this barrier)
      const float2 other = part[(c.w ^ 1) * 128 + c.rs];
      const float m1 = (m1a + m1b + other.x) * (1.0f / float(DV)), m2 = (m2a + m2b + other.y) * (1.0f / float(DV));
#pragma unroll
      for (int j = 0; j < HC / 16; ++j) {
        float o[16], t1[16], t2[16];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const float4 gq = g4[4 * j + q4];
          const float gam[4] = {gq.x, gq.y, gq.z, gq.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = 16 * j + 4 * q4 + e;
            o[4 * q4 + e] = rstd * (dy[k] * gam[e] - m1 - xh[k] * m2);
            t1[4 * q4 + e] = dy[k] * xh[k]; t2[4 * q4 + e] = dy[k];
          }
        }
        const int c0 = HC * c.w + 16 * j;
        st_chunks16(dcb, c.rs, c0, o);
        st_chunks16(dyx, c.rs, c0, t1);
        st_chunks16(dyx, c.rs, DV + c0, t2);
      }
    }
    MMRCA_STAMP(5);
    cta_sync_for_mma();
    MMRCA_STAMP(6);
    // ---- dP (read back next); then (G2) the LayerNorm-affine gradients: M = 192 = two M=128 MMAs ------------------
    if (tid == 0) {
      for (int h = 0; h < 2; ++h)
        mma_steps(tmem + (uint32_t(16 * h) << 16) + T::DP, make_smem_desc(smem_u32(dcb + h * 8 * kRS), kCS, kRS), 2 * kCS,
                  make_smem_desc(smem_u32(vb + h * 8 * kRS), kCS, kRS), 2 * kCS, make_idesc_bf16(64, 64, 0, 0), DV / 16, false);
      umma_commit(c.bar);
      for (int mt = 0; mt < 2; ++mt)
        mma_steps(tmem + T::G_LN + 16 * mt, make_smem_desc(smem_u32(dyx + mt * 16 * kCS), kRS, kCS), 2 * kRS,
                  make_smem_desc(smem_u32(ones), 256, 128), 512, make_idesc_bf16(128, 16, 1, 0), 8, !first);
      umma_commit(&bars[4]);
    }
    cta_wait_mma(c);
    MMRCA_STAMP(7);
    {
      float ds[16];
      softmax_bwd16(c, T::DP, false, p, ds);
      store_half_row_split(c, dls, ds, true);
    }
    cta_sync_for_mma();
    MMRCA_STAMP(8);
    if (tid == 0) {
      for (int h = 0; h < 2; ++h) {
        mma_steps(tmem + (uint32_t(16 * h) << 16) + T::DV, make_smem_desc(smem_u32(pb + h * kPHalf), kRS, kPCS), 2 * kRS,
                  make_smem_desc(smem_u32(dcb + h * 8 * kRS), kRS, kCS), 2 * kRS, make_idesc_bf16(64, DV, 1, 1), 4, false);
        mma_steps(tmem + (uint32_t(16 * h) << 16) + T::DZ, make_smem_desc(smem_u32(dls + h * kPHalf), kPCS, kRS), 2 * kPCS,
                  make_smem_desc(smem_u32(xb + h * 8 * kRS), kRS, kCS), 2 * kRS, make_idesc_bf16(64, DIN, 0, 1), 4, false);
      }
      umma_commit(c.bar);
    }
    if (tile + int(gridDim.x) < tiles) load_dout(tile + int(gridDim.x));
    cta_wait_mma(c);
    MMRCA_STAMP(9);
    mbar_wait_ph(&bars[4], ph_g2);                 // dgamma|dbeta have read dy*xhat | dy: its bytes become dZ
    // dC is dead (dP and dV have read it): its buffer takes dV
    if (c.w == 0) { acc_cols_to_operand(c, T::DZ, 0, DIN, dyx, c.rs); acc_cols_to_operand(c, T::DV, 0, 16, dcb, c.rs); }
    else acc_cols_to_operand(c, T::DV, 16, 96, dcb, c.rs);
    MMRCA_STAMP(10);
    cta_sync_for_mma();
    MMRCA_STAMP(11);
    if (tid == 96 && a.dz_tiles) {      // fine-tune phase: dZ, dS, dV also go to HBM (bulk stores, read asynchronously)
      bulk_s2g(static_cast<uint8_t*>(a.dz_tiles) + size_t(tile) * op_bytes(DIN), dyx, op_bytes(DIN));
      bulk_s2g(static_cast<uint8_t*>(a.ds_tiles) + size_t(tile) * (2 * kPHalf), dls, 2 * kPHalf);
      bulk_s2g(static_cast<uint8_t*>(a.dv_tiles) + size_t(tile) * op_bytes(96), dcb, op_bytes(96));
      bulk_commit();
    }
    if (tid == 0) {       // G3: nobody waits for these before the next tile is under way
      mma_steps(tmem + T::G_M, make_smem_desc(smem_u32(xb), kRS, kCS), 2 * kRS, make_smem_desc(smem_u32(dyx), kRS, kCS), 2 * kRS,
                make_idesc_bf16(128, DIN, 1, 1), 8, !first);
      mma_steps(tmem + T::G_WV, make_smem_desc(smem_u32(xb), kRS, kCS), 2 * kRS, make_smem_desc(smem_u32(dcb), kRS, kCS), 2 * kRS,
                make_idesc_bf16(128, DV, 1, 1), 8, !first);
      umma_commit(&bars[5]);
    }
    first = false;
  }
  MMRCA_STAMP(0);
  if (tid == 96 && a.dz_tiles) bulk_wait_all();      // the exported images are complete before the CTA retires
  if (!first) {
    mbar_wait_ph(&bars[5], ph_g3);
    flush_acc(c, T::G_M, DIN, DIN + 1, [&](int m, int n) { return a.gm + size_t(n) * 128 + m; });
    flush_acc(c, T::G_WV, DV, DIN + 1, [&](int m, int n) { return m < DIN ? a.g_wv + size_t(n) * DIN + m : a.g_bv + n; });
    {
      float v[16];
      ld16f(c.tmem + c.lane_base + T::G_LN + 16 * c.w, v);
      const int col = 128 * c.w + c.rp;             // row of the [dy*xhat | dy]^T 1 product
      if (col < DV) red_add(a.g_ln_g + col, v[0]);
      else if (col < 2 * DV) red_add(a.g_ln_b + (col - DV), v[0]);
    }
  }
  __syncthreads();
  MMRCA_STAMP(1);
  MMRCA_STAMP_NS(251);
  if (a.dbg && tid == 0) a.dbg[301 + 2 * blockIdx.x] = globaltimer_ns();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// Both modalities in one launch: blockIdx.y = 0 image (DIN 80), 1 text (DIN 48), gridDim.x CTAs each.  The two
// backward passes are independent, a tile costs about the same in both, and 2 x 512 tiles over 148 SMs is 7 rounds
// in one launch against 2 x 4 in two (plus one prologue and one accumulator flush instead of two).
struct SaBwdBothArgs { SaBwdArgs m[2]; };
constexpr uint32_t kSaBwdSmemBytes = SaBwdSmem<80>::BYTES > SaBwdSmem<48>::BYTES ? SaBwdSmem<80>::BYTES : SaBwdSmem<48>::BYTES;
__global__ void __launch_bounds__(kCtaThreads, 1) sa_bwd_kernel(const SaBwdBothArgs a) {
  extern __shared__ __align__(128) uint8_t sm[];
  if (blockIdx.y == 0) sa_bwd_body<80>(a.m[0], sm);
  else sa_bwd_body<48>(a.m[1], sm);
}

// ---------------------------------------------------------------------------------------------------------------
// dW_query, dW_key, db_query from dM, du (chain rule through M = Wq^T Wk / sqrt(d), u = Wk^T bq / sqrt(d))
// ---------------------------------------------------------------------------------------------------------------
struct FinBlock {
  const float* gm;     // [DIN][128]: gm[k'][k] = dM[k][k'], gm[k'][DIN] = du[k']
  const float* wq; const float* bq; const float* wk;
  float* g_wq; float* g_bq; float* g_wk;
  int din, dkq;
};
struct FinArgs { FinBlock blk[4]; int nblk; };
constexpr int kFinRows = 1;                                         // rows n of W_query / W_key per CTA
constexpr uint32_t kFinSmemBytes = (96 * 98 + 2 * kFinRows * 96 + kFinRows) * 4;
static_assert(kFinRows == 1, "finalize_kernel's thread mapping assumes one row per CTA");

// grid.x = sum over blocks of d_kq / 8.  CTA = (block, 8 rows n): dM (d_in x (d_in + 1), du in the last column) and the
// 8 rows of W_query / W_key are staged in shared memory with every load in flight at once, then each thread owns
// outputs (n, k): dWq[n][k] += s sum_k' dM[k][k'] Wk[n][k'],  dWk[n][k'] += s (sum_k Wq[n][k] dM[k][k'] + bq[n] du[k']),
// dbq[n] += s sum_k' du[k'] Wk[n][k'].  Every output has exactly one owner.  (Measured and dropped: 8 rows per CTA - 48 CTAs
// instead of 384, dM staged 8 times less often - ran 16.9 us against 10.7 us: the kernel is a latency chain, not traffic.)
__global__ void __launch_bounds__(256) finalize_kernel(const FinArgs a) {
  extern __shared__ __align__(16) float fsm[];
  int b = 0, grp = blockIdx.x;
  for (; b < a.nblk; ++b) {
    const int g = a.blk[b].dkq / kFinRows;
    if (grp < g) break;
    grp -= g;
  }
  if (b >= a.nblk) return;
  const FinBlock& B = a.blk[b];
  const int din = B.din, gs = din + 1, n0 = grp * kFinRows, tid = threadIdx.x;   // odd stride: conflict-free both ways
  float* g_s = fsm;                          // [din][gs]: g_s[k'][k] = dM[k][k'], g_s[k'][din] = du[k']
  float* wk_s = g_s + din * gs;              // [8][din]
  float* wq_s = wk_s + kFinRows * din;       // [8][din]
  float* bq_s = wq_s + kFinRows * din;       // [8]
  {   // dM rows are 128 floats apart: warp w takes rows w, w + 8, ... whole (one 16-byte load per lane, no index
      // arithmetic), every load of the thread in flight before the first shared-memory store
    const int warp = tid >> 5, lane = tid & 31;
    float4 v[12];
#pragma unroll
    for (int u = 0; u < 12; ++u) {
      const int kp = warp + 8 * u;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kp < din && 4 * lane <= din) v[u] = __ldg(reinterpret_cast<const float4*>(B.gm + size_t(kp) * 128) + lane);
    }
#pragma unroll
    for (int u = 0; u < 12; ++u) {
      const int kp = warp + 8 * u;
      if (kp < din && 4 * lane <= din) {
        float* d = g_s + kp * gs + 4 * lane;
        const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (4 * lane + j <= din) d[j] = e[j];
      }
    }
  }
  for (int i = tid; i < kFinRows * din; i += 256) {
    wk_s[i] = __ldg(B.wk + size_t(n0) * din + i);
    wq_s[i] = __ldg(B.wq + size_t(n0) * din + i);
  }
  if (tid < kFinRows) bq_s[tid] = __ldg(B.bq + n0 + tid);
  __syncthreads();
  const float s = rsqrtf(float(B.dkq));
  // two threads per output column k (each half of the contraction), the last warp takes db_query
  if (tid < 2 * din) {
    const int k = tid >> 1, half = tid & 1, j0 = half * (din / 2), j1 = j0 + din / 2;
    float q0 = 0.f, q1 = 0.f, k0 = half ? 0.f : bq_s[0] * g_s[k * gs + din], k1 = 0.f;
#pragma unroll 4
    for (int j = j0; j < j1; j += 2) {
      q0 = fmaf(g_s[j * gs + k], wk_s[j], q0);
      q1 = fmaf(g_s[(j + 1) * gs + k], wk_s[j + 1], q1);
      k0 = fmaf(wq_s[j], g_s[k * gs + j], k0);
      k1 = fmaf(wq_s[j + 1], g_s[k * gs + j + 1], k1);
    }
    float q = q0 + q1, kk = k0 + k1;
    q += __shfl_xor_sync(0xffffffffu, q, 1);
    kk += __shfl_xor_sync(0xffffffffu, kk, 1);
    if (half == 0) red_add(B.g_wq + size_t(n0) * din + k, q * s);      // accumulate (+=), not waited for
    else red_add(B.g_wk + size_t(n0) * din + k, kk * s);
  } else if (tid >= 224) {
    const int lane = tid & 31;
    float acc = 0.f;
    for (int j = lane; j < din; j += 32) acc = fmaf(g_s[j * gs + din], wk_s[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) red_add(B.g_bq + n0, acc * s);
  }
}

}  // namespace htc
}  // namespace mmrca

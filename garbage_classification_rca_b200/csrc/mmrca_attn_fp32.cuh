// fp32 SIMT attention-block kernels (the 1e-4-relative contract path).
//
// One kernel = one SelfAttention / ReverseCrossAttention block of the reference
// (CVPR_code/multimodal_model.py:39-108): stacked Q/K/V projection, 16x16 scores,
// softmax, optional reverse weights (1-A)/(L-1), P*V, LayerNorm, ReLU — everything
// between the block's input rows and its output rows stays in shared memory/registers.
// The backward kernel recomputes the block from its inputs (nothing but the block
// inputs is read back from HBM), produces the input gradients and the row-space
// projection gradients dY = [dQ|dK|dV]; the weight gradients dW = dY^T X are reduced
// over the batch by wgrad_kernel (mmrca_misc_fp32.cuh).
//
// Layout: a CTA owns a tile of G samples = R = 16*G rows.  Weights are staged once per
// (persistent) CTA, transposed to wT[k][n] with an odd row stride so that both the
// forward (lanes along n) and the transposed backward (lanes along k) reads are
// bank-conflict free.  Row buffers use strides == 4 (mod 8) floats: float4-aligned rows
// and conflict-free scalar reads when lanes walk interleaved rows.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mmrca {

constexpr int kL = 16;          // chunks ("patches") per sample, multimodal_model.py:249
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr float kLnEps = 1e-5f; // torch.nn.LayerNorm default
constexpr int kLdA = 20;        // row stride of the 16x16 attention matrices in smem

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int DIN_, int DKQ_, int DV_, bool SELF_, int G_, bool BWD_>
struct AttnCfg {
  static constexpr int DIN = DIN_, DKQ = DKQ_, DV = DV_, G = G_;
  static constexpr bool SELF = SELF_, BWD = BWD_;
  static constexpr int R = G * kL;
  static constexpr int NKV = DKQ + DV;
  static constexpr int NALL = 2 * DKQ + DV;
  static constexpr int LDW = NALL + 1;   // odd
  static constexpr int LDY = NALL + 4;   // == 4 (mod 8)
  static constexpr int LDC = DV + 4;     // == 4 (mod 8)
  static constexpr int RW = R / kWarps;  // rows per warp in the row-parallel phases
  static constexpr int CMV = (DV + 31) / 32;
  static constexpr int CMQ = (DKQ + 31) / 32;
  static constexpr int CMI = (DIN + 31) / 32;
  static_assert(R % kWarps == 0 && kL % RW == 0, "a warp's rows must stay inside one sample");
  static_assert(DIN % 4 == 0 && DKQ % 4 == 0 && DV % 4 == 0, "float4 rows");
  static_assert(LDY % 8 == 4 && LDC % 8 == 4, "row strides must be 4 mod 8");
  static_assert(G * 64 <= kThreads, "score phase uses 64 threads per sample");
  // shared-memory carve-up (floats)
  static constexpr int OFF_WT = 0;
  static constexpr int OFF_BIAS = OFF_WT + DIN * LDW + ((4 - (DIN * LDW) % 4) % 4);
  static constexpr int OFF_LNG = OFF_BIAS + NALL;
  static constexpr int OFF_LNB = OFF_LNG + DV;
  static constexpr int OFF_XQ = OFF_LNB + DV;
  static constexpr int OFF_XKV = OFF_XQ + R * DIN;
  static constexpr int OFF_Y = OFF_XKV + (SELF ? 0 : R * DIN);
  static constexpr int OFF_A = OFF_Y + R * LDY;
  static constexpr int OFF_DCTX = OFF_A + R * kLdA;
  static constexpr int OFF_DS = OFF_DCTX + (BWD ? R * LDC : 0);
  static constexpr int OFF_RED = OFF_DS + (BWD ? R * kLdA : 0);
  static constexpr int SMEM_FLOATS = OFF_RED + (BWD ? 2 * kWarps * 32 * CMV : 0) + 32;
  static constexpr size_t SMEM_BYTES = size_t(SMEM_FLOATS) * sizeof(float);
  static_assert(SMEM_BYTES <= 232448, "tile does not fit the 227 KB of shared memory per CTA");
};

struct AttnArgs {
  const float* xq;    // [B,16,DIN] (raw features when normalise)
  const float* xkv;   // [B,16,DIN] (== xq for self attention)
  const float* wq; const float* bq; const float* wk; const float* bk; const float* wv; const float* bv;
  const float* ln_g; const float* ln_b;
  float* out;         // fwd: [B,16,DV]
  float* norms;       // [B]: written by a normalising forward, read by a normalising backward
  int batch;
  int reverse;
  int normalise;      // per-sample L2 normalisation of the (self) input, multimodal_model.py:662-665
  // backward only
  const float* dout;  // [B,16,DV]
  float* dy;          // [B*16, NALL] row-space projection grads for wgrad_kernel
  float* dxq;         // [B,16,DIN] or null
  float* dxkv;        // [B,16,DIN] or null (non-self only)
  int acc_dxq, acc_dxkv;   // += instead of =
  float* g_ln_g; float* g_ln_b;  // accumulated
};

// ---- staging ---------------------------------------------------------------------------
template <class C>
__device__ __forceinline__ void stage_weights(float* sm, const AttnArgs& a) {
  // wT[k][n] = W[n][k]; consecutive threads read consecutive global addresses.
  auto put = [&](const float* w, const float* b, int n0, int nn) {
    for (int i = threadIdx.x; i < nn * C::DIN; i += kThreads) {
      int n = i / C::DIN, k = i - n * C::DIN;
      sm[C::OFF_WT + k * C::LDW + n0 + n] = w[i];
    }
    for (int i = threadIdx.x; i < nn; i += kThreads) sm[C::OFF_BIAS + n0 + i] = b[i];
  };
  put(a.wq, a.bq, 0, C::DKQ);
  put(a.wk, a.bk, C::DKQ, C::DKQ);
  put(a.wv, a.bv, 2 * C::DKQ, C::DV);
  for (int i = threadIdx.x; i < C::DV; i += kThreads) {
    sm[C::OFF_LNG + i] = a.ln_g[i];
    sm[C::OFF_LNB + i] = a.ln_b[i];
  }
}

// Load the tile's input rows (zero for samples past the batch) and optionally L2-normalise
// each sample's 16*DIN-wide vector in place (x / ||x||, no epsilon — multimodal_model.py:662-665).
template <class C>
__device__ __forceinline__ void load_rows(float* dst, const float* src, int b0, int batch) {
  constexpr int V4 = C::R * C::DIN / 4;
  const float4* s4 = reinterpret_cast<const float4*>(src + size_t(b0) * kL * C::DIN);
  float4* d4 = reinterpret_cast<float4*>(dst);
  const int valid4 = max(0, min(C::G, batch - b0)) * kL * C::DIN / 4;
  for (int i = threadIdx.x; i < V4; i += kThreads)
    d4[i] = (i < valid4) ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
}

template <class C>
__device__ __forceinline__ void normalise_rows(float* x, float* norms, int b0, int batch, bool recompute) {
  constexpr int PER = kL * C::DIN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int g = warp; g < C::G; g += kWarps) {
    if (b0 + g >= batch) continue;
    float* xs = x + g * PER;
    float nrm;
    if (recompute) {
      float ss = 0.f;
      for (int i = lane; i < PER; i += 32) ss = fmaf(xs[i], xs[i], ss);
      nrm = sqrtf(warp_sum(ss));
      if (lane == 0 && norms) norms[b0 + g] = nrm;
    } else {
      nrm = norms[b0 + g];
    }
    for (int i = lane; i < PER; i += 32) xs[i] = xs[i] / nrm;
  }
}

// ---- y[r][n0..n0+NCOLS) = x[r][:] . wT[:, n0..] + bias -----------------------------------
template <int R, int K, int NCOLS>
__device__ __forceinline__ void proj_gemm(const float* __restrict__ xs, int ldx,
                                          const float* __restrict__ wT, int ldw,
                                          const float* __restrict__ bias,
                                          float* __restrict__ y, int ldy) {
  constexpr int RW = R / kWarps, CJ = (NCOLS + 31) / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[RW][CJ];
#pragma unroll
  for (int j = 0; j < CJ; ++j) {
    const int n = lane + 32 * j;
    const float b = (n < NCOLS) ? bias[n] : 0.f;
#pragma unroll
    for (int i = 0; i < RW; ++i) acc[i][j] = b;
  }
  const float* xr = xs + warp * RW * ldx;
#pragma unroll 2
  for (int k0 = 0; k0 < K; k0 += 4) {
    float xv[RW][4];
#pragma unroll
    for (int i = 0; i < RW; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(xr + i * ldx + k0);
      xv[i][0] = t.x; xv[i][1] = t.y; xv[i][2] = t.z; xv[i][3] = t.w;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float w[CJ];
#pragma unroll
      for (int j = 0; j < CJ; ++j) {
        const int n = lane + 32 * j;
        w[j] = (NCOLS % 32 == 0 || n < NCOLS) ? wT[(k0 + kk) * ldw + n] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < RW; ++i)
#pragma unroll
        for (int j = 0; j < CJ; ++j) acc[i][j] = fmaf(xv[i][kk], w[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < RW; ++i)
#pragma unroll
    for (int j = 0; j < CJ; ++j) {
      const int n = lane + 32 * j;
      if (NCOLS % 32 == 0 || n < NCOLS) y[(warp * RW + i) * ldy + n] = acc[i][j];
    }
}

// ---- dx[r][k] (+)= sum_n dy[r][n0+n] * W[n][k]  (transposed read of wT) --------------------
template <int R, int K, int NCOLS>
__device__ __forceinline__ void dgrad_gemm(const float* __restrict__ dy, int ldy,
                                           const float* __restrict__ wT, int ldw,
                                           float* __restrict__ dx /*global [R][K]*/, int valid_rows,
                                           bool accumulate) {
  constexpr int RW = R / kWarps, CK = (K + 31) / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[RW][CK];
#pragma unroll
  for (int i = 0; i < RW; ++i)
#pragma unroll
    for (int m = 0; m < CK; ++m) acc[i][m] = 0.f;
  const float* dr = dy + warp * RW * ldy;
#pragma unroll 2
  for (int n0 = 0; n0 < NCOLS; n0 += 4) {
    float dv[RW][4];
#pragma unroll
    for (int i = 0; i < RW; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(dr + i * ldy + n0);
      dv[i][0] = t.x; dv[i][1] = t.y; dv[i][2] = t.z; dv[i][3] = t.w;
    }
#pragma unroll
    for (int nn = 0; nn < 4; ++nn) {
      float w[CK];
#pragma unroll
      for (int m = 0; m < CK; ++m) {
        const int k = lane + 32 * m;
        w[m] = (K % 32 == 0 || k < K) ? wT[k * ldw + n0 + nn] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < RW; ++i)
#pragma unroll
        for (int m = 0; m < CK; ++m) acc[i][m] = fmaf(dv[i][nn], w[m], acc[i][m]);
    }
  }
#pragma unroll
  for (int i = 0; i < RW; ++i) {
    const int r = warp * RW + i;
    if (r >= valid_rows) continue;
#pragma unroll
    for (int m = 0; m < CK; ++m) {
      const int k = lane + 32 * m;
      if (K % 32 == 0 || k < K) {
        float* p = dx + size_t(r) * K + k;
        *p = accumulate ? (*p + acc[i][m]) : acc[i][m];
      }
    }
  }
}

// ---- per-sample 16x16 block of dot products -------------------------------------------------
// 64 threads per sample: thread (ch, jb, ib) owns s[ii][jj] for rows i = ib+4*ii, j = jb+4*jj and
// the column subset c = ch + 4*t; the 4 column subsets are summed with two shuffles, so all
// four ch-lanes end with the full sums.
template <int DC>
__device__ __forceinline__ void block_dots(const float* __restrict__ A, int lda,
                                           const float* __restrict__ Bm, int ldb,
                                           int g, int ib, int jb, int ch, float (&s)[4][4]) {
#pragma unroll
  for (int ii = 0; ii < 4; ++ii)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) s[ii][jj] = 0.f;
  const float* a0 = A + (g * kL + ib) * lda + ch;
  const float* b0 = Bm + (g * kL + jb) * ldb + ch;
#pragma unroll 4
  for (int t = 0; t < DC / 4; ++t) {
    float av[4], bv[4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) av[ii] = a0[ii * 4 * lda + 4 * t];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) bv[jj] = b0[jj * 4 * ldb + 4 * t];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) s[ii][jj] = fmaf(av[ii], bv[jj], s[ii][jj]);
  }
#pragma unroll
  for (int ii = 0; ii < 4; ++ii)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      s[ii][jj] += __shfl_xor_sync(0xffffffffu, s[ii][jj], 1);
      s[ii][jj] += __shfl_xor_sync(0xffffffffu, s[ii][jj], 2);
    }
}

// scores -> softmax -> A in smem (stride kLdA).  multimodal_model.py:56-60 / :87-91.
template <class C>
__device__ __forceinline__ void scores_softmax(float* sm) {
  const int t = threadIdx.x;
  const int ch = t & 3, jb = (t >> 2) & 3, ib = (t >> 4) & 3, g = t >> 6;
  if (g < C::G) {   // warp-uniform: 64 threads = 2 whole warps per sample
    float s[4][4];
    const float* y = sm + C::OFF_Y;
    block_dots<C::DKQ>(y, C::LDY, y + C::DKQ, C::LDY, g, ib, jb, ch, s);
    const float sq = sqrtf(float(C::DKQ));
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      float m = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) { s[ii][jj] = s[ii][jj] / sq; m = fmaxf(m, s[ii][jj]); }
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
      float sum = 0.f;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) { s[ii][jj] = expf(s[ii][jj] - m); sum += s[ii][jj]; }
      sum += __shfl_xor_sync(0xffffffffu, sum, 4);
      sum += __shfl_xor_sync(0xffffffffu, sum, 8);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) s[ii][jj] = s[ii][jj] / sum;
    }
    float* as = sm + C::OFF_A + (g * kL) * kLdA;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
      if (ch == ii) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) as[(ib + 4 * ii) * kLdA + jb + 4 * jj] = s[ii][jj];
      }
  }
}

__device__ __forceinline__ float attn_weight(float a, bool reverse) {
  // multimodal_model.py:97-98: (1 - A) / (L - 1)
  return reverse ? (1.0f - a) / float(kL - 1) : a;
}

// ctx rows of this warp: acc[i][m] = sum_j P[r0+i][j] * V[g*16+j][lane+32m]
template <class C>
__device__ __forceinline__ void pv_rows(const float* sm, bool reverse, float (&acc)[C::RW][C::CMV]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = warp * C::RW, g = r0 / kL;
  const float* v = sm + C::OFF_Y + (g * kL) * C::LDY + 2 * C::DKQ;
  const float* as = sm + C::OFF_A + r0 * kLdA;
#pragma unroll
  for (int i = 0; i < C::RW; ++i)
#pragma unroll
    for (int m = 0; m < C::CMV; ++m) acc[i][m] = 0.f;
#pragma unroll
  for (int j0 = 0; j0 < kL; j0 += 4) {
    float p[C::RW][4];
#pragma unroll
    for (int i = 0; i < C::RW; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(as + i * kLdA + j0);
      p[i][0] = attn_weight(t.x, reverse); p[i][1] = attn_weight(t.y, reverse);
      p[i][2] = attn_weight(t.z, reverse); p[i][3] = attn_weight(t.w, reverse);
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      float vv[C::CMV];
#pragma unroll
      for (int m = 0; m < C::CMV; ++m) {
        const int c = lane + 32 * m;
        vv[m] = (C::DV % 32 == 0 || c < C::DV) ? v[(j0 + jj) * C::LDY + c] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < C::RW; ++i)
#pragma unroll
        for (int m = 0; m < C::CMV; ++m) acc[i][m] = fmaf(p[i][jj], vv[m], acc[i][m]);
    }
  }
}

// LayerNorm statistics of one row held across the warp (multimodal_model.py:65 / :105).
template <class C>
__device__ __forceinline__ void ln_stats(const float (&x)[C::CMV], float& mean, float& rstd) {
  const int lane = threadIdx.x & 31;
  float s = 0.f;
#pragma unroll
  for (int m = 0; m < C::CMV; ++m) s += (C::DV % 32 == 0 || lane + 32 * m < C::DV) ? x[m] : 0.f;
  mean = warp_sum(s) / float(C::DV);
  float q = 0.f;
#pragma unroll
  for (int m = 0; m < C::CMV; ++m) {
    const float d = (C::DV % 32 == 0 || lane + 32 * m < C::DV) ? (x[m] - mean) : 0.f;
    q = fmaf(d, d, q);
  }
  rstd = rsqrtf(warp_sum(q) / float(C::DV) + kLnEps);
}

// ---- forward kernel --------------------------------------------------------------------------
template <class C>
__device__ __forceinline__ void forward_tile_qkv(float* sm, const AttnArgs& a, int b0) {
  load_rows<C>(sm + C::OFF_XQ, a.xq, b0, a.batch);
  if (!C::SELF) load_rows<C>(sm + C::OFF_XKV, a.xkv, b0, a.batch);
  __syncthreads();
  if (a.normalise) {
    normalise_rows<C>(sm + C::OFF_XQ, a.norms, b0, a.batch, !C::BWD);
    __syncthreads();
  }
  if (C::SELF) {
    proj_gemm<C::R, C::DIN, C::NALL>(sm + C::OFF_XQ, C::DIN, sm + C::OFF_WT, C::LDW, sm + C::OFF_BIAS,
                                     sm + C::OFF_Y, C::LDY);
  } else {
    proj_gemm<C::R, C::DIN, C::DKQ>(sm + C::OFF_XQ, C::DIN, sm + C::OFF_WT, C::LDW, sm + C::OFF_BIAS,
                                    sm + C::OFF_Y, C::LDY);
    proj_gemm<C::R, C::DIN, C::NKV>(sm + C::OFF_XKV, C::DIN, sm + C::OFF_WT + C::DKQ, C::LDW,
                                    sm + C::OFF_BIAS + C::DKQ, sm + C::OFF_Y + C::DKQ, C::LDY);
  }
  __syncthreads();
  scores_softmax<C>(sm);
  __syncthreads();
}

template <class C>
__global__ void __launch_bounds__(kThreads, 1) attn_fwd_kernel(const AttnArgs a) {
  extern __shared__ __align__(16) float sm[];
  stage_weights<C>(sm, a);
  const int tiles = (a.batch + C::G - 1) / C::G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool reverse = a.reverse != 0;
  __syncthreads();
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int b0 = tile * C::G;
    forward_tile_qkv<C>(sm, a, b0);
    float ctx[C::RW][C::CMV];
    pv_rows<C>(sm, reverse, ctx);
#pragma unroll
    for (int i = 0; i < C::RW; ++i) {
      const int r = warp * C::RW + i;
      const int b = b0 + r / kL;
      float mean, rstd;
      ln_stats<C>(ctx[i], mean, rstd);
      if (b < a.batch) {
#pragma unroll
        for (int m = 0; m < C::CMV; ++m) {
          const int c = lane + 32 * m;
          if (C::DV % 32 == 0 || c < C::DV) {
            const float y = (ctx[i][m] - mean) * rstd * sm[C::OFF_LNG + c] + sm[C::OFF_LNB + c];
            a.out[(size_t(b) * kL + (r % kL)) * C::DV + c] = fmaxf(y, 0.f);   // ReLU, :66 / :106
          }
        }
      }
    }
    __syncthreads();   // smem tile is reused by the next iteration
  }
}

// ---- backward kernel (recompute + input grads + row-space projection grads) ---------------------
template <class C>
__global__ void __launch_bounds__(kThreads, 1) attn_bwd_kernel(const AttnArgs a) {
  extern __shared__ __align__(16) float sm[];
  stage_weights<C>(sm, a);
  const int tiles = (a.batch + C::G - 1) / C::G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool reverse = a.reverse != 0;
  float g_gamma[C::CMV], g_beta[C::CMV];
#pragma unroll
  for (int m = 0; m < C::CMV; ++m) g_gamma[m] = g_beta[m] = 0.f;
  __syncthreads();
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int b0 = tile * C::G;
    const int valid_rows = max(0, min(C::G, a.batch - b0)) * kL;
    forward_tile_qkv<C>(sm, a, b0);
    // -- LayerNorm/ReLU backward on this warp's rows -> dctx in smem
    {
      float ctx[C::RW][C::CMV];
      pv_rows<C>(sm, reverse, ctx);
#pragma unroll
      for (int i = 0; i < C::RW; ++i) {
        const int r = warp * C::RW + i;
        const int b = b0 + r / kL;
        float mean, rstd;
        ln_stats<C>(ctx[i], mean, rstd);
        float xhat[C::CMV], dxh[C::CMV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int m = 0; m < C::CMV; ++m) {
          const int c = lane + 32 * m;
          const bool ok = (C::DV % 32 == 0 || c < C::DV) && b < a.batch;
          xhat[m] = 0.f; dxh[m] = 0.f;
          if (ok) {
            xhat[m] = (ctx[i][m] - mean) * rstd;
            const float gam = sm[C::OFF_LNG + c];
            const float y = xhat[m] * gam + sm[C::OFF_LNB + c];
            const float go = __ldg(a.dout + (size_t(b) * kL + (r % kL)) * C::DV + c);
            const float dy = y > 0.f ? go : 0.f;
            g_gamma[m] = fmaf(dy, xhat[m], g_gamma[m]);
            g_beta[m] += dy;
            dxh[m] = dy * gam;
            s1 += dxh[m];
            s2 = fmaf(dxh[m], xhat[m], s2);
          }
        }
        s1 = warp_sum(s1) / float(C::DV);
        s2 = warp_sum(s2) / float(C::DV);
#pragma unroll
        for (int m = 0; m < C::CMV; ++m) {
          const int c = lane + 32 * m;
          if (C::DV % 32 == 0 || c < C::DV)
            sm[C::OFF_DCTX + r * C::LDC + c] = rstd * (dxh[m] - s1 - xhat[m] * s2);
        }
      }
    }
    __syncthreads();
    // -- dP = dctx V^T, softmax backward -> dS (already divided by sqrt(d_kq))
    {
      const int t = threadIdx.x;
      const int ch = t & 3, jb = (t >> 2) & 3, ib = (t >> 4) & 3, g = t >> 6;
      if (g < C::G) {
        float s[4][4];
        block_dots<C::DV>(sm + C::OFF_DCTX, C::LDC, sm + C::OFF_Y + 2 * C::DKQ, C::LDY, g, ib, jb, ch, s);
        const float sq = sqrtf(float(C::DKQ));
        const float* as = sm + C::OFF_A + (g * kL) * kLdA;
        float* ds = sm + C::OFF_DS + (g * kL) * kLdA;
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          float av[4], dot = 0.f;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            av[jj] = as[(ib + 4 * ii) * kLdA + jb + 4 * jj];
            if (reverse) s[ii][jj] = -s[ii][jj] / float(kL - 1);
            dot = fmaf(s[ii][jj], av[jj], dot);
          }
          dot += __shfl_xor_sync(0xffffffffu, dot, 4);
          dot += __shfl_xor_sync(0xffffffffu, dot, 8);
          if (ch == ii) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              ds[(ib + 4 * ii) * kLdA + jb + 4 * jj] = av[jj] * (s[ii][jj] - dot) / sq;
          }
        }
      }
    }
    __syncthreads();
    // -- dV (over V), dQ, dK (registers first: dQ needs K and dK needs Q)
    {
      const int r0 = warp * C::RW, g = r0 / kL, j_or_i0 = r0 % kL;
      const float* as = sm + C::OFF_A + (g * kL) * kLdA;
      const float* dss = sm + C::OFF_DS + (g * kL) * kLdA;
      float* yb = sm + C::OFF_Y + (g * kL) * C::LDY;
      float dv[C::RW][C::CMV];
#pragma unroll
      for (int i = 0; i < C::RW; ++i)
#pragma unroll
        for (int m = 0; m < C::CMV; ++m) dv[i][m] = 0.f;
#pragma unroll 4
      for (int i2 = 0; i2 < kL; ++i2) {     // dV[j][c] = sum_i P[i][j] dctx[i][c]
        float dc[C::CMV];
#pragma unroll
        for (int m = 0; m < C::CMV; ++m) {
          const int c = lane + 32 * m;
          dc[m] = (C::DV % 32 == 0 || c < C::DV) ? sm[C::OFF_DCTX + (g * kL + i2) * C::LDC + c] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < C::RW; ++i) {
          const float p = attn_weight(as[i2 * kLdA + j_or_i0 + i], reverse);
#pragma unroll
          for (int m = 0; m < C::CMV; ++m) dv[i][m] = fmaf(p, dc[m], dv[i][m]);
        }
      }
      float dq[C::RW][C::CMQ], dk[C::RW][C::CMQ];
#pragma unroll
      for (int i = 0; i < C::RW; ++i)
#pragma unroll
        for (int m = 0; m < C::CMQ; ++m) dq[i][m] = dk[i][m] = 0.f;
#pragma unroll 4
      for (int j = 0; j < kL; ++j) {        // dQ[i][c] = sum_j dS[i][j] K[j][c]; dK[j'][c] = sum_i dS[i][j'] Q[i][c]
        float kv[C::CMQ], qv[C::CMQ];
#pragma unroll
        for (int m = 0; m < C::CMQ; ++m) {
          const int c = lane + 32 * m;
          const bool ok = (C::DKQ % 32 == 0 || c < C::DKQ);
          kv[m] = ok ? yb[j * C::LDY + C::DKQ + c] : 0.f;
          qv[m] = ok ? yb[j * C::LDY + c] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < C::RW; ++i) {
          const float s_ij = dss[(j_or_i0 + i) * kLdA + j];   // row i (mine), col j
          const float s_ji = dss[j * kLdA + j_or_i0 + i];     // row j, col i (mine)
#pragma unroll
          for (int m = 0; m < C::CMQ; ++m) {
            dq[i][m] = fmaf(s_ij, kv[m], dq[i][m]);
            dk[i][m] = fmaf(s_ji, qv[m], dk[i][m]);
          }
        }
      }
      __syncthreads();   // every warp has finished reading Q, K, V
#pragma unroll
      for (int i = 0; i < C::RW; ++i) {
        float* yr = yb + (j_or_i0 + i) * C::LDY;
#pragma unroll
        for (int m = 0; m < C::CMQ; ++m) {
          const int c = lane + 32 * m;
          if (C::DKQ % 32 == 0 || c < C::DKQ) { yr[c] = dq[i][m]; yr[C::DKQ + c] = dk[i][m]; }
        }
#pragma unroll
        for (int m = 0; m < C::CMV; ++m) {
          const int c = lane + 32 * m;
          if (C::DV % 32 == 0 || c < C::DV) yr[2 * C::DKQ + c] = dv[i][m];
        }
      }
    }
    __syncthreads();
    // -- dY tile -> global (for wgrad_kernel), coalesced float4 rows
    {
      constexpr int N4 = C::NALL / 4;
      float4* dst = reinterpret_cast<float4*>(a.dy + size_t(b0) * kL * C::NALL);
      for (int i = threadIdx.x; i < valid_rows * N4; i += kThreads) {
        const int r = i / N4, c4 = i - r * N4;
        dst[i] = *reinterpret_cast<const float4*>(sm + C::OFF_Y + r * C::LDY + 4 * c4);
      }
    }
    // -- input gradients
    if (C::SELF) {
      if (a.dxq)
        dgrad_gemm<C::R, C::DIN, C::NALL>(sm + C::OFF_Y, C::LDY, sm + C::OFF_WT, C::LDW,
                                          a.dxq + size_t(b0) * kL * C::DIN, valid_rows, a.acc_dxq != 0);
    } else {
      if (a.dxq)
        dgrad_gemm<C::R, C::DIN, C::DKQ>(sm + C::OFF_Y, C::LDY, sm + C::OFF_WT, C::LDW,
                                         a.dxq + size_t(b0) * kL * C::DIN, valid_rows, a.acc_dxq != 0);
      if (a.dxkv)
        dgrad_gemm<C::R, C::DIN, C::NKV>(sm + C::OFF_Y + C::DKQ, C::LDY, sm + C::OFF_WT + C::DKQ, C::LDW,
                                         a.dxkv + size_t(b0) * kL * C::DIN, valid_rows, a.acc_dxkv != 0);
    }
    __syncthreads();
  }
  // -- LayerNorm affine grads: reduce the 8 warps' register partials, one atomic per column per CTA
  float* red = sm + C::OFF_RED;
#pragma unroll
  for (int m = 0; m < C::CMV; ++m) {
    red[(warp * C::CMV + m) * 32 + lane] = g_gamma[m];
    red[((kWarps + warp) * C::CMV + m) * 32 + lane] = g_beta[m];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C::DV; c += kThreads) {
    const int m = c / 32, l = c % 32;
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      sg += red[(w * C::CMV + m) * 32 + l];
      sb += red[((kWarps + w) * C::CMV + m) * 32 + l];
    }
    atomicAdd(a.g_ln_g + c, sg);
    atomicAdd(a.g_ln_b + c, sb);
  }
}

}  // namespace mmrca

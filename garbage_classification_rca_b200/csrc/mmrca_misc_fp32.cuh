// fp32 kernels around the attention blocks: batch-reduced weight gradients, the
// concat + dropout + classifier (multimodal_model.py:689-726), cross-entropy
// (main_both.py:87-93) and the L2-normalisation backward (multimodal_model.py:662-665).
// All of them are HBM-bound streaming kernels: coalesced float4 rows, register
// accumulators across a CTA's slab, one atomic per output element per CTA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mmrca_attn_fp32.cuh"

namespace mmrca {

// ---- dW[n][k] += sum_r dY[r][col0+n] * X[r][k] / norm[r/16];  db[n] += sum_r dY[r][col0+n] ----
// grid = (row slabs, ceil(N/64)); lanes along k, each warp owns 8 of the CTA's 64 columns.
constexpr int kWgRows = 32;   // rows per smem tile
constexpr int kWgCols = 64;   // dY columns per CTA

template <int K>
__global__ void __launch_bounds__(kThreads) wgrad_kernel(const float* __restrict__ dy, int ldy, int col0, int ncols,
                                                         const float* __restrict__ x, const float* __restrict__ norms,
                                                         int rows, float* __restrict__ dw, float* __restrict__ db) {
  constexpr int CK = (K + 31) / 32;
  constexpr int LDD = kWgCols + 4;
  __shared__ __align__(16) float xs[kWgRows * K];
  __shared__ __align__(16) float ds[kWgRows * LDD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb0 = blockIdx.y * kWgCols;             // first column (relative to col0) of this CTA
  const int nvalid = min(kWgCols, ncols - nb0);
  float acc[8][CK];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int m = 0; m < CK; ++m) acc[i][m] = 0.f;
  float bacc = 0.f;
  const int tiles = (rows + kWgRows - 1) / kWgRows;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int r0 = tile * kWgRows;
    for (int i = threadIdx.x; i < kWgRows * K / 4; i += kThreads) {
      const int r = (4 * i) / K;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < rows) {
        v = __ldg(reinterpret_cast<const float4*>(x + size_t(r0) * K) + i);
        if (norms) {
          const float n = __ldg(norms + (r0 + r) / kL);
          v.x = v.x / n; v.y = v.y / n; v.z = v.z / n; v.w = v.w / n;
        }
      }
      reinterpret_cast<float4*>(xs)[i] = v;
    }
    for (int i = threadIdx.x; i < kWgRows * kWgCols / 4; i += kThreads) {
      const int r = i / (kWgCols / 4), c4 = i - r * (kWgCols / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < rows && 4 * c4 < nvalid)
        v = __ldg(reinterpret_cast<const float4*>(dy + size_t(r0 + r) * ldy + col0 + nb0 + 4 * c4));
      *reinterpret_cast<float4*>(ds + r * LDD + 4 * c4) = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < kWgRows; ++r) {
      float xv[CK];
#pragma unroll
      for (int m = 0; m < CK; ++m) {
        const int k = lane + 32 * m;
        xv[m] = (K % 32 == 0 || k < K) ? xs[r * K + k] : 0.f;
      }
      const float4 d0 = *reinterpret_cast<const float4*>(ds + r * LDD + warp * 8);
      const float4 d1 = *reinterpret_cast<const float4*>(ds + r * LDD + warp * 8 + 4);
      const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int m = 0; m < CK; ++m) acc[i][m] = fmaf(dv[i], xv[m], acc[i][m]);
    }
    if (threadIdx.x < kWgCols) {
#pragma unroll 8
      for (int r = 0; r < kWgRows; ++r) bacc += ds[r * LDD + threadIdx.x];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = nb0 + warp * 8 + i;
    if (n < ncols) {
#pragma unroll
      for (int m = 0; m < CK; ++m) {
        const int k = lane + 32 * m;
        if (K % 32 == 0 || k < K) atomicAdd(dw + size_t(n) * K + k, acc[i][m]);
      }
    }
  }
  if (db && threadIdx.x < kWgCols && nb0 + threadIdx.x < ncols) atomicAdd(db + nb0 + threadIdx.x, bacc);
}

// ---- classifier over the (never materialised) concat ------------------------------------------
struct CatSeg {
  const float* src;    // [B, width]
  const float* norms;  // [B] or null: values are divided by norms[b] (the L2-normalised features)
  float* dst;          // backward: where d(concat segment) goes, or null
  int width;
  int dst_acc;
};
struct CatArgs {
  CatSeg seg[4];
  int nseg;
  int D;               // total concat width
  int batch;
  const uint8_t* mask; // [B, D] keep mask or null
  float scale;         // 1/(1-p)
  const float* wf;     // [NC, D]
  const float* bf;     // [NC]
  float* logits;       // fwd out [B, NC]
  const float* dlogits;// bwd in  [B, NC]
  float* g_wf;         // bwd acc [NC, D]
  float* g_bf;         // bwd acc [NC]
};

// forward: one warp per sample; weights are read through L1/L2 (57 KB, shared by all warps)
template <int NC>
__global__ void __launch_bounds__(kThreads) classifier_fwd_kernel(const CatArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpg = gridDim.x * kWarps;
  for (int b = blockIdx.x * kWarps + warp; b < a.batch; b += wpg) {
    float acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.f;
    int off = 0;
    for (int s = 0; s < a.nseg; ++s) {
      const CatSeg& sg = a.seg[s];
      const float nrm = sg.norms ? __ldg(sg.norms + b) : 1.f;
      const float4* src = reinterpret_cast<const float4*>(sg.src + size_t(b) * sg.width);
      for (int j = lane * 4; j < sg.width; j += 128) {
        float4 v = __ldg(src + j / 4);
        if (sg.norms) { v.x = v.x / nrm; v.y = v.y / nrm; v.z = v.z / nrm; v.w = v.w / nrm; }
        if (a.mask) {
          const uchar4 mk = *reinterpret_cast<const uchar4*>(a.mask + size_t(b) * a.D + off + j);
          v.x = mk.x ? v.x * a.scale : 0.f; v.y = mk.y ? v.y * a.scale : 0.f;
          v.z = mk.z ? v.z * a.scale : 0.f; v.w = mk.w ? v.w * a.scale : 0.f;
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(a.wf + size_t(c) * a.D + off + j));
          acc[c] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, acc[c]))));
        }
      }
      off += sg.width;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float t = warp_sum(acc[c]);
      if (lane == 0) a.logits[size_t(b) * NC + c] = t + __ldg(a.bf + c);
    }
  }
}

// backward: grid = (sample slabs, ceil(D/1024)); a thread owns 4 consecutive concat columns.
template <int NC>
__global__ void __launch_bounds__(kThreads) classifier_bwd_kernel(const CatArgs a) {
  const int j = blockIdx.y * (kThreads * 4) + threadIdx.x * 4;
  const bool col_ok = j < a.D;
  // locate the segment of column j (segment widths are multiples of 4)
  int s = 0, off = 0;
  if (col_ok) {
    while (s < a.nseg - 1 && j >= off + a.seg[s].width) { off += a.seg[s].width; ++s; }
  }
  const CatSeg sg = a.seg[s];
  const int jl = j - off;
  float4 w[NC], acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    w[c] = col_ok ? __ldg(reinterpret_cast<const float4*>(a.wf + size_t(c) * a.D + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
    acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int per = (a.batch + gridDim.x - 1) / gridDim.x;
  const int b_lo = blockIdx.x * per, b_hi = min(a.batch, b_lo + per);
  if (col_ok) {
    for (int b = b_lo; b < b_hi; ++b) {
      float4 v = __ldg(reinterpret_cast<const float4*>(sg.src + size_t(b) * sg.width + jl));
      if (sg.norms) {
        const float nrm = __ldg(sg.norms + b);
        v.x = v.x / nrm; v.y = v.y / nrm; v.z = v.z / nrm; v.w = v.w / nrm;
      }
      float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);
      if (a.mask) {
        const uchar4 m8 = *reinterpret_cast<const uchar4*>(a.mask + size_t(b) * a.D + j);
        mk = make_float4(m8.x ? a.scale : 0.f, m8.y ? a.scale : 0.f, m8.z ? a.scale : 0.f, m8.w ? a.scale : 0.f);
      }
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float dl = __ldg(a.dlogits + size_t(b) * NC + c);
        d.x = fmaf(dl, w[c].x, d.x); d.y = fmaf(dl, w[c].y, d.y);
        d.z = fmaf(dl, w[c].z, d.z); d.w = fmaf(dl, w[c].w, d.w);
        acc[c].x = fmaf(dl, v.x * mk.x, acc[c].x); acc[c].y = fmaf(dl, v.y * mk.y, acc[c].y);
        acc[c].z = fmaf(dl, v.z * mk.z, acc[c].z); acc[c].w = fmaf(dl, v.w * mk.w, acc[c].w);
      }
      if (sg.dst) {
        d.x *= mk.x; d.y *= mk.y; d.z *= mk.z; d.w *= mk.w;
        float4* p = reinterpret_cast<float4*>(sg.dst + size_t(b) * sg.width + jl);
        if (sg.dst_acc) { const float4 o = *p; d.x += o.x; d.y += o.y; d.z += o.z; d.w += o.w; }
        *p = d;
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float* g = a.g_wf + size_t(c) * a.D + j;
      atomicAdd(g + 0, acc[c].x); atomicAdd(g + 1, acc[c].y);
      atomicAdd(g + 2, acc[c].z); atomicAdd(g + 3, acc[c].w);
    }
  }
  if (blockIdx.y == 0 && threadIdx.x < NC) {
    float sb = 0.f;
    for (int b = b_lo; b < b_hi; ++b) sb += __ldg(a.dlogits + size_t(b) * NC + threadIdx.x);
    atomicAdd(a.g_bf + threadIdx.x, sb);
  }
}

// ---- CrossEntropyLoss(weight, label_smoothing), mean reduction — one CTA, two passes -------------
constexpr int kCeThreads = 1024;
constexpr int kCeMaxClasses = 16;

__device__ __forceinline__ float block_sum_1024(float v, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < kCeThreads / 32) ? red[threadIdx.x] : 0.f;
  if (warp == 0) { t = warp_sum(t); if (lane == 0) red[0] = t; }
  __syncthreads();
  return red[0];
}

__global__ void __launch_bounds__(kCeThreads) cross_entropy_kernel(const float* __restrict__ logits,
                                                                   const int64_t* __restrict__ labels,
                                                                   const float* __restrict__ cw, float eps,
                                                                   int batch, int nc,
                                                                   float* __restrict__ loss_out,
                                                                   float* __restrict__ dlogits) {
  __shared__ float red[32];
  float wsum = 0.f;
  for (int b = threadIdx.x; b < batch; b += kCeThreads) wsum += cw ? __ldg(cw + labels[b]) : 1.f;
  const float denom = block_sum_1024(wsum, red);
  float lsum = 0.f;
  for (int b = threadIdx.x; b < batch; b += kCeThreads) {
    float z[kCeMaxClasses];
    float m = -INFINITY;
    for (int c = 0; c < nc; ++c) { z[c] = logits[size_t(b) * nc + c]; m = fmaxf(m, z[c]); }
    float se = 0.f;
    for (int c = 0; c < nc; ++c) se += expf(z[c] - m);
    const float lse = m + logf(se);
    const int y = int(labels[b]);
    float tsum = 0.f, li = 0.f;
    for (int c = 0; c < nc; ++c) {
      const float wc = cw ? __ldg(cw + c) : 1.f;
      const float t = (eps / float(nc)) * wc + (c == y ? (1.f - eps) * wc : 0.f);
      z[c] = z[c] - lse;            // log p
      li -= t * z[c];
      tsum += t;
    }
    lsum += li;
    if (dlogits)
      for (int c = 0; c < nc; ++c) {
        const float wc = cw ? __ldg(cw + c) : 1.f;
        const float t = (eps / float(nc)) * wc + (c == y ? (1.f - eps) * wc : 0.f);
        dlogits[size_t(b) * nc + c] = (expf(z[c]) * tsum - t) / denom;
      }
  }
  const float total = block_sum_1024(lsum, red);
  if (threadIdx.x == 0 && loss_out) *loss_out = total / denom;
}

// ---- backward of x / ||x||: dx = (g - xn (xn.g)) / n, in place on g; one warp per sample ----------
__global__ void __launch_bounds__(kThreads) l2norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ norms,
                                                              float* __restrict__ g, int batch, int d) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = blockIdx.x * kWarps + warp; b < batch; b += gridDim.x * kWarps) {
    const float n = __ldg(norms + b);
    const float4* xs = reinterpret_cast<const float4*>(x + size_t(b) * d);
    float4* gs = reinterpret_cast<float4*>(g + size_t(b) * d);
    float dot = 0.f;
    for (int j = lane; j < d / 4; j += 32) {
      const float4 xv = __ldg(xs + j), gv = gs[j];
      dot += (xv.x / n) * gv.x + (xv.y / n) * gv.y + (xv.z / n) * gv.z + (xv.w / n) * gv.w;
    }
    dot = warp_sum(dot);
    for (int j = lane; j < d / 4; j += 32) {
      const float4 xv = __ldg(xs + j);
      float4 gv = gs[j];
      gv.x = (gv.x - (xv.x / n) * dot) / n; gv.y = (gv.y - (xv.y / n) * dot) / n;
      gv.z = (gv.z - (xv.z / n) * dot) / n; gv.w = (gv.w - (xv.w / n) * dot) / n;
      gs[j] = gv;
    }
  }
}

}  // namespace mmrca

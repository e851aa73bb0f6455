// Classic / Normalized late-fusion heads (--late_fusion=classic | normalized), fp32.
//
// Reference: CVPR_code/multimodal_model.py:489-579 (EffV2MediumAndDistilbertClassic / ...Normalized.forward after the
// backbones): image_to_hidden_size Linear(1280 -> H), text_to_hidden_size Linear(768 -> H)  [:521-522, :566-567],
// Normalized only: each divided by its row L2 norm, no epsilon  [:569-570], concat  [:524-525, :572-573],
// concat_layer Linear(2H -> H)  [:527, :575], self.drop  [:528, :576], fc_layer Linear(H -> n_classes)  [:529, :577].
// These are the "projections into the shared fusion dimension" (H = num_neurons_FC, 256).  No non-linearity sits
// between the Linear layers.  Backward (main_both.py:112): gradients of the four Linear layers and, for the fine-tune
// phase, of the pooled features.
//
// Arithmetic: one generic strided fp32 GEMM (64 x 64 x 16 shared-memory tiles, 4 x 4 register blocking) serves the
// forward (X W^T), the input gradients (dY W) and the weight gradients (dY^T X, accumulated into the caller's
// gradient); the L2 normalisation is folded into the GEMM operand loads as a per-row / per-contraction-index scale,
// so the normalised hidden vectors are never materialised.  Dropout + fc_layer + their backward are the concat
// classifier kernels of the MM-RCA fp32 path with a single 256-wide segment (mmrca_misc_fp32.cuh).  This head is the
// 1e-4-relative contract; a tensor-core version would reuse the hierarchical head's tcgen05 pipelines.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mmrca_attn_fp32.cuh"

namespace mmrca {
namespace fus {

constexpr int kTM = 64, kTN = 64, kTK = 16;

struct GemmArgs {
  const float* a; long long sa_m, sa_k;      // A(i, k) = a[i * sa_m + k * sa_k],  i < M, k < K
  const float* b; long long sb_k, sb_n;      // B(k, j) = b[k * sb_k + j * sb_n],  j < N
  float* c; long long ldc;                   // C(i, j) = c[i * ldc + j]
  const float* bias;                         // [N] added once (ignored when accumulate), or null
  const float* inv_m;                        // [M]: A(i, k) is divided by inv_m[i] (row L2 norms), or null
  const float* inv_k;                        // [K]: A(i, k) is divided by inv_k[k], or null
  int M, N, K, accumulate;
};

// C (+)= A B.  grid = (ceil(N / 64), ceil(M / 64), K splits), 256 threads; thread (ty, tx) owns rows 4 ty.. and columns
// 4 tx.. of the tile.  The tile loads pick the thread mapping that walks the operand's contiguous axis.  K splits > 1
// (accumulating GEMMs only: the weight gradients contract over the batch and have few output tiles) add their partial
// products with red.global.add.
__global__ void __launch_bounds__(256) sgemm_kernel(const GemmArgs g) {
  __shared__ __align__(16) float As[kTK][kTM + 4];
  __shared__ __align__(16) float Bs[kTK][kTN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * kTM, n0 = blockIdx.x * kTN;
  const bool a_kfast = g.sa_k == 1, b_nfast = g.sb_n == 1;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int kper = ((g.K + int(gridDim.z) - 1) / int(gridDim.z) + kTK - 1) / kTK * kTK;
  const int k_begin = int(blockIdx.z) * kper, k_end = min(g.K, k_begin + kper);
  for (int k0 = k_begin; k0 < k_end; k0 += kTK) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = tid + 256 * u;
      {
        const int i = a_kfast ? idx / kTK : idx % kTM, k = a_kfast ? idx % kTK : idx / kTM;
        float v = 0.f;
        if (m0 + i < g.M && k0 + k < k_end) {
          v = __ldg(g.a + (long long)(m0 + i) * g.sa_m + (long long)(k0 + k) * g.sa_k);
          if (g.inv_m) v = v / __ldg(g.inv_m + m0 + i);
          if (g.inv_k) v = v / __ldg(g.inv_k + k0 + k);
        }
        As[k][i] = v;
      }
      {
        const int j = b_nfast ? idx % kTN : idx / kTK, k = b_nfast ? idx / kTN : idx % kTK;
        float v = 0.f;
        if (n0 + j < g.N && k0 + k < k_end) v = __ldg(g.b + (long long)(k0 + k) * g.sb_k + (long long)(n0 + j) * g.sb_n);
        Bs[k][j] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][4 * ty]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][4 * tx]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + 4 * ty + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + 4 * tx + j;
      if (n >= g.N) continue;
      float* p = g.c + (long long)m * g.ldc + n;
      float v = acc[i][j];
      if (gridDim.z > 1) { atomicAdd(p, v); continue; }
      if (g.accumulate) v += *p;
      else if (g.bias) v += __ldg(g.bias + n);
      *p = v;
    }
  }
}

// out[n] += sum_r x[r * ld + n]   (bias gradients).  grid = (ceil(N / 32), row slabs): a warp sums one column group.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, long long ld, int rows, int N,
                                                     float* __restrict__ out) {
  __shared__ float part[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const int per = (rows + gridDim.y - 1) / gridDim.y, r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float s = 0.f;
  if (n < N)
    for (int r = r0 + warp; r < r1; r += 8) s += __ldg(x + (long long)r * ld + n);
  part[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w][lane];
    atomicAdd(out + n, t);
  }
}

}  // namespace fus
}  // namespace mmrca

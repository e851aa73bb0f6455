// Seeded inverted dropout of the concatenated feature vector (reference CVPR_code/multimodal_model.py:190
// `self.drop = nn.Dropout(drop_ratio)`, applied at :719 right before the final Linear).
//
// torch's Philox stream cannot be reproduced from inside another kernel, so the product path draws its own
// counter-based mask: element (sample b, concat column col) is kept iff a 16-bit draw derived from
// hash(seed, (b * D + col) / 2) is >= round(p * 65536); kept values are scaled by 1 / (1 - p).  The mask is a
// pure function of (seed, p, b, col): the forward and the backward kernels regenerate it on the fly (no bytes
// in HBM), mmrca_dropout_mask() materialises the very same mask for the fp32 kernels and for the parity tests.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mmrca {

struct DropSpec {
  uint32_t seed_lo, seed_hi;
  uint32_t thresh;   // drop when the 16-bit draw < thresh; 0 = dropout off (eval mode or p = 0)
  float scale;       // 1 / (1 - p)
  int D;             // concat width (even)
};

__host__ __device__ __forceinline__ uint32_t drop_hash(const DropSpec& s, uint32_t pair) {
  uint32_t h = pair ^ s.seed_lo;
  h ^= h >> 16; h *= 0x7feb352du;
  h ^= h >> 15; h *= 0x846ca68bu;
  h ^= h >> 16; h ^= s.seed_hi;
  h *= 0x9e3779b1u; h ^= h >> 15;
  return h;
}

// multipliers (0 or scale) of the two concat columns (col, col + 1), col even, of sample b
__device__ __forceinline__ void drop_pair(const DropSpec& s, uint32_t b, uint32_t col, float& m0, float& m1) {
  const uint32_t h = drop_hash(s, b * uint32_t(s.D >> 1) + (col >> 1));
  m0 = (h & 0xffffu) >= s.thresh ? s.scale : 0.f;
  m1 = (h >> 16) >= s.thresh ? s.scale : 0.f;
}

// keep-bits of `n` (multiple of 2, <= 64) consecutive concat columns starting at the even column col0
__device__ __forceinline__ uint64_t drop_bits(const DropSpec& s, uint32_t b, uint32_t col0, int n) {
  uint64_t bits = 0;
  const uint32_t p0 = b * uint32_t(s.D >> 1) + (col0 >> 1);
#pragma unroll 4
  for (int j = 0; j < n; j += 2) {
    const uint32_t h = drop_hash(s, p0 + uint32_t(j >> 1));
    bits |= uint64_t((h & 0xffffu) >= s.thresh ? 1u : 0u) << j;
    bits |= uint64_t((h >> 16) >= s.thresh ? 1u : 0u) << (j + 1);
  }
  return bits;
}

// uint8 keep-mask [B][D] of the same draws (1 = kept)
__global__ void __launch_bounds__(256) dropout_mask_kernel(const DropSpec s, int batch, uint8_t* __restrict__ out) {
  const size_t pairs = size_t(batch) * size_t(s.D >> 1);
  for (size_t p = size_t(blockIdx.x) * 256 + threadIdx.x; p < pairs; p += size_t(gridDim.x) * 256) {
    const uint32_t h = drop_hash(s, uint32_t(p));
    uchar2 v;
    v.x = (h & 0xffffu) >= s.thresh ? 1 : 0;
    v.y = (h >> 16) >= s.thresh ? 1 : 0;
    reinterpret_cast<uchar2*>(out)[p] = v;
  }
}

}  // namespace mmrca

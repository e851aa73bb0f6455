// Seeded inverted dropout of the concatenated feature vector (reference CVPR_code/multimodal_model.py:190
// `self.drop = nn.Dropout(drop_ratio)`, applied at :719 right before the final Linear).
//
// torch's Philox stream cannot be reproduced from inside another kernel, so the product path draws its own
// counter-based mask: element (sample b, concat column col) is kept iff a 16-bit draw derived from
// hash(seed, (b * D + col) / 4) (four draws per hash) is >= round(p * 65536); kept values are scaled by 1 / (1 - p).  The mask is a
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
  int D;             // concat width (multiple of 4)
};

// two 32-bit words (four 16-bit draws) for the quad of concat columns [4 q', 4 q' + 4) of a sample; quad = global index
__host__ __device__ __forceinline__ void drop_hash(const DropSpec& s, uint32_t quad, uint32_t& w0, uint32_t& w1) {
  uint32_t h = quad ^ s.seed_lo;
  h ^= h >> 16; h *= 0x7feb352du;
  h ^= h >> 15; h *= 0x846ca68bu;
  h ^= h >> 16; h ^= s.seed_hi;
  w0 = h * 0x9e3779b1u; w0 ^= w0 >> 15;      // two different multiplicative finalisers of the mixed word
  w1 = h * 0x85ebca6bu; w1 ^= w1 >> 13;
}

// multipliers (0 or scale) of the four concat columns col .. col + 3 (col % 4 == 0) of sample b
__device__ __forceinline__ void drop_quad(const DropSpec& s, uint32_t b, uint32_t col, float (&m)[4]) {
  uint32_t w0, w1;
  drop_hash(s, b * uint32_t(s.D >> 2) + (col >> 2), w0, w1);
  m[0] = (w0 & 0xffffu) >= s.thresh ? s.scale : 0.f;
  m[1] = (w0 >> 16) >= s.thresh ? s.scale : 0.f;
  m[2] = (w1 & 0xffffu) >= s.thresh ? s.scale : 0.f;
  m[3] = (w1 >> 16) >= s.thresh ? s.scale : 0.f;
}
// the same for eight columns (col % 8 == 0), applied in place
__device__ __forceinline__ void drop_apply8(const DropSpec& s, uint32_t b, uint32_t col, float (&x)[8]) {
  float m[4];
  drop_quad(s, b, col, m);
  x[0] *= m[0]; x[1] *= m[1]; x[2] *= m[2]; x[3] *= m[3];
  drop_quad(s, b, col + 4, m);
  x[4] *= m[0]; x[5] *= m[1]; x[6] *= m[2]; x[7] *= m[3];
}

// keep-bits of N (multiple of 4, <= 64) consecutive concat columns starting at column col0 (col0 % 4 == 0); fully
// unrolled so that the N / 4 hashes are independent instruction streams (a rolled loop serialises their multiply chains)
template <int N>
__device__ __forceinline__ uint64_t drop_bits(const DropSpec& s, uint32_t b, uint32_t col0) {
  static_assert(N % 4 == 0 && N <= 64, "quads");
  uint32_t lo = 0, hi = 0;
  const uint32_t q0 = b * uint32_t(s.D >> 2) + (col0 >> 2);
#pragma unroll
  for (int j = 0; j < N; j += 4) {
    uint32_t w0, w1;
    drop_hash(s, q0 + uint32_t(j >> 2), w0, w1);
    const uint32_t k = ((w0 & 0xffffu) >= s.thresh ? 1u : 0u) | ((w0 >> 16) >= s.thresh ? 2u : 0u) |
                       ((w1 & 0xffffu) >= s.thresh ? 4u : 0u) | ((w1 >> 16) >= s.thresh ? 8u : 0u);
    if (j < 32) lo |= k << j; else hi |= k << (j - 32);
  }
  return uint64_t(lo) | (uint64_t(hi) << 32);
}

// uint8 keep-mask [B][D] of the same draws (1 = kept)
__global__ void __launch_bounds__(256) dropout_mask_kernel(const DropSpec s, int batch, uint8_t* __restrict__ out) {
  const size_t quads = size_t(batch) * size_t(s.D >> 2);
  for (size_t q = size_t(blockIdx.x) * 256 + threadIdx.x; q < quads; q += size_t(gridDim.x) * 256) {
    uint32_t w0, w1;
    drop_hash(s, uint32_t(q), w0, w1);
    uchar4 v;
    v.x = (w0 & 0xffffu) >= s.thresh ? 1 : 0;
    v.y = (w0 >> 16) >= s.thresh ? 1 : 0;
    v.z = (w1 & 0xffffu) >= s.thresh ? 1 : 0;
    v.w = (w1 >> 16) >= s.thresh ? 1 : 0;
    reinterpret_cast<uchar4*>(out)[q] = v;
  }
}

}  // namespace mmrca

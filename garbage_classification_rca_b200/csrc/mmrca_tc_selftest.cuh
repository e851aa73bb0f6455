// Diagnostic kernel for the UMMA building blocks (mmrca_tc.cuh): one 128 x N x K bf16 GEMM with fp32
// accumulation through shared-memory descriptors -> tcgen05.mma -> TMEM -> tcgen05.ld, in all four
// operand-major combinations the head kernels use.  Exercised by tests/test_tc_blocks_gpu.py.
#pragma once
#include "mmrca_tc.cuh"

namespace mmrca {
namespace tc {

// mode bit 0: B is MN-major (source b[K][N]) instead of K-major (source b[N][K])
// mode bit 1: A is MN-major (source a[K][128]) instead of K-major (source a[128][K])
// mode bit 2: two M=64 MMAs (rows 0-63 and 64-127) instead of one M=128; the second accumulator sits 16 TMEM
//             lanes up, i.e. a warp's lanes 0-15 hold rows 16q.. of the first half and lanes 16-31 rows 64+16q..
// mode bit 3: B is the constant all-ones operand of the column-sum MMAs (N = 16): ONE K = 16 slice of it (512 bytes of
//             0x3F80, ordinary descriptor) that every K-step re-reads - the B descriptor is simply not advanced along K:
//             out[m][n] = sum_k A[m][k].  b is ignored.
// out[128][N] = A * B^T (logical A[128][K], B[N][K]).  N % 16 == 0, N <= 256, K % 16 == 0.
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(int mode, const float* __restrict__ a,
                                                              const float* __restrict__ b, float* __restrict__ out,
                                                              int N, int K) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool a_mn = mode & 2, b_mn = mode & 1;
  const int M = 128;
  // operand geometry (bytes)
  const uint32_t a_sbo = 128, a_lbo = (M / 8) * 128 + 16;
  const uint32_t b_sbo = 128, b_lbo = (N / 8) * 128 + 16;
  const uint32_t a_bytes = (K / 8) * a_lbo, b_off = (a_bytes + 127) & ~127u;
  uint8_t* sa = smem;
  uint8_t* sb = smem + b_off;

  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) {
    uint32_t cols = 32;
    while (cols < uint32_t(N)) cols <<= 1;
    tmem_alloc(&tmem_base, cols);
  }
  // ---- stage A ----
  if (!a_mn) {   // source a[128][K]: chunk (r, kc) = 8 consecutive k
    for (int i = tid; i < M * (K / 8); i += 128) {
      const int r = i / (K / 8), kc = i % (K / 8);
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = a[r * K + kc * 8 + e];
      *reinterpret_cast<uint4*>(sa + core_off(r, kc, a_lbo, a_sbo)) = pack_bf16x8(v);
    }
  } else {       // source a[K][128]: chunk (k, mc) = 8 consecutive m
    for (int i = tid; i < K * (M / 8); i += 128) {
      const int k = i / (M / 8), mc = i % (M / 8);
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = a[k * M + mc * 8 + e];
      *reinterpret_cast<uint4*>(sa + uint32_t(mc) * a_sbo + uint32_t(k >> 3) * a_lbo + uint32_t(k & 7) * 16u) =
          pack_bf16x8(v);
    }
  }
  // ---- stage B ----
  if (mode & 8) {
    for (int i = tid; i < 512 / 16; i += 128)
      reinterpret_cast<uint4*>(sb)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  } else if (!b_mn) {   // source b[N][K]
    for (int i = tid; i < N * (K / 8); i += 128) {
      const int n = i / (K / 8), kc = i % (K / 8);
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = b[n * K + kc * 8 + e];
      *reinterpret_cast<uint4*>(sb + core_off(n, kc, b_lbo, b_sbo)) = pack_bf16x8(v);
    }
  } else {       // source b[K][N]
    for (int i = tid; i < K * (N / 8); i += 128) {
      const int k = i / (N / 8), nc = i % (N / 8);
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = b[k * N + nc * 8 + e];
      *reinterpret_cast<uint4*>(sb + uint32_t(nc) * b_sbo + uint32_t(k >> 3) * b_lbo + uint32_t(k & 7) * 16u) =
          pack_bf16x8(v);
    }
  }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t taddr = tmem_base;
  const bool halves = mode & 4;
  if (tid == 0) {
    const bool bones = (mode & 8) != 0;
    const uint64_t ad = make_smem_desc(smem_u32(sa), a_lbo, a_sbo);
    const uint64_t bd = make_smem_desc(smem_u32(sb), bones ? 256u : b_lbo, b_sbo);
    const uint32_t b_lbo = bones ? 0u : (uint32_t(N) / 8) * 128 + 16;      // shadows: K-advance of the B descriptor (none for ones)
    if (!halves) {
      const uint32_t idesc = make_idesc_bf16(M, N, a_mn ? 1 : 0, b_mn ? 1 : 0);
      for (int ks = 0; ks < K / 16; ++ks)
        umma_bf16(taddr, desc_advance(ad, ks * 2 * a_lbo), desc_advance(bd, ks * 2 * b_lbo), idesc, ks > 0);
    } else {
      const uint32_t idesc = make_idesc_bf16(64, N, a_mn ? 1 : 0, b_mn ? 1 : 0);
      for (int h = 0; h < 2; ++h)     // 64 rows = 8 row groups (K-major) or 8 m-chunks (MN-major): 8 * a_sbo either way
        for (int ks = 0; ks < K / 16; ++ks)
          umma_bf16(taddr + (uint32_t(16 * h) << 16), desc_advance(ad, h * 8 * a_sbo + ks * 2 * a_lbo),
                    desc_advance(bd, ks * 2 * b_lbo), idesc, ks > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  const int r = halves ? 64 * (lane >> 4) + 16 * warp + (lane & 15) : warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(taddr + (uint32_t(warp * 32) << 16) + uint32_t(c0), v);
#pragma unroll
    for (int e = 0; e < 16; ++e) out[r * N + c0 + e] = v[e];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    uint32_t cols = 32;
    while (cols < uint32_t(N)) cols <<= 1;
    tmem_dealloc(taddr, cols);
  }
}

// mode 16: both operands MN-major in the 128-byte-swizzled layout a {64 columns, K rows} SWIZZLE_128B tensor-map box lands
// from a row-major [K][columns] source: per 64-column block, K rows of 128 bytes whose 16-byte chunks are XOR-swizzled by
// (row % 8); blocks K * 128 bytes apart (the descriptor's leading-dimension offset), 8-row groups 1024 bytes apart (stride
// offset), one K = 16 step = two row groups = 2048 bytes.  a: [K][128], b: [K][N]; out[128][N] = A^T B.
__global__ void __launch_bounds__(128, 1) umma_mn128_selftest_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                    float* __restrict__ out, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nblk = (N + 63) / 64;
  const uint32_t blk = uint32_t(K) * 128;
  uint8_t* sa = smem;
  uint8_t* sb = smem + 2 * blk;
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) {
    uint32_t cols = 32;
    while (cols < uint32_t(N)) cols <<= 1;
    tmem_alloc(&tmem_base, cols);
  }
  for (int i = tid; i < K * 16; i += 128) {          // A: 16 chunks of 8 columns per row
    const int k = i / 16, c = i % 16;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = a[k * 128 + c * 8 + e];
    *reinterpret_cast<uint4*>(sa + uint32_t(c >> 3) * blk + uint32_t(k) * 128 + uint32_t(((c & 7) ^ (k & 7)) * 16)) = pack_bf16x8(v);
  }
  for (int i = tid; i < K * nblk * 8; i += 128) {
    const int k = i / (nblk * 8), c = i % (nblk * 8);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = c * 8 + e < N ? b[k * N + c * 8 + e] : 0.f;
    *reinterpret_cast<uint4*>(sb + uint32_t(c >> 3) * blk + uint32_t(k) * 128 + uint32_t(((c & 7) ^ (k & 7)) * 16)) = pack_bf16x8(v);
  }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t taddr = tmem_base;
  if (tid == 0) {
    const uint64_t ad = make_smem_desc_sw128_mn(smem_u32(sa), blk), bd = make_smem_desc_sw128_mn(smem_u32(sb), blk);
    const uint32_t idesc = make_idesc_bf16(128, N, 1, 1);
    for (int ks = 0; ks < K / 16; ++ks) umma_bf16(taddr, desc_advance(ad, ks * 2048), desc_advance(bd, ks * 2048), idesc, ks > 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  const int r = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(taddr + (uint32_t(warp * 32) << 16) + uint32_t(c0), v);
#pragma unroll
    for (int e = 0; e < 16; ++e) out[r * N + c0 + e] = v[e];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    uint32_t cols = 32;
    while (cols < uint32_t(N)) cols <<= 1;
    tmem_dealloc(taddr, cols);
  }
}

inline size_t umma_selftest_smem_bytes(int N, int K) {
  const size_t a_lbo = (128 / 8) * 128 + 16, b_lbo = (size_t(N) / 8) * 128 + 16;
  return (((K / 8) * a_lbo + 127) & ~size_t(127)) + (K / 8) * b_lbo + 128;
}

}  // namespace tc
}  // namespace mmrca

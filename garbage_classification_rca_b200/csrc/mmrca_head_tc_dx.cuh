// bf16 tensor-core pipeline of the MM-RCA head: gradients of the pooled FEATURES (fine-tune phase, reference
// main_both.py:687-694: the backbones are unfrozen and loss.backward() continues into them).
//
// sa_bwd_kernel leaves, per tile and modality, the images dZ, dS, dV (MMRCA_FLAG_FEATURE_GRADS).  This kernel turns
// them into d(loss)/d(feature):
//   Z      = X M + u                                   (recomputed: one MMA chain, as in the forward)
//   d(xn)  = dZ M^T + dS^T Z + dV Wv                    (self attention: query source == key/value source; three MMA chains
//                                                        into the same pair of M=64 accumulators)
//          + mask * dlogits Wf[:, feature columns]      (the classifier's direct view of the normalised features,
//                                                        multimodal_model.py:708-726; fp32 FMAs; absent for cross_attention_only)
//   d(x)   = (d(xn) - xn (xn . d(xn))) / ||x||          (through x / ||x||, :662-665; the dot runs over the sample's 16 chunks)
// One warpgroup per CTA, persistent over the tiles; blockIdx.y selects the modality.
#pragma once
#include "mmrca_head_tc_bwd.cuh"

namespace mmrca {
namespace htc {

struct SaDxArgs {
  const void* x_tiles;      // [tiles][x_tile_bytes(DIN)]
  const void* dz_tiles;     // [tiles][op_bytes(DIN)]
  const void* ds_tiles;     // [tiles][2 * kPHalf]
  const void* dv_tiles;     // [tiles][op_bytes(96)]
  const void* blobs;        // bz | bv (hi) | ...
  const float* feat;        // raw features [B][16 * DIN] fp32
  const float* norms;       // [B]
  const float* dlogits;     // [B][4]
  const float* wf;          // classifier weight [4][D] or null (cross_attention_only: no direct term)
  int cls_off, D;           // first concat column of this modality's features; concat width
  DropSpec drop;
  float* d_feat;            // out [B][16 * DIN]
  int batch;
};
struct SaDxBothArgs { SaDxArgs m[2]; };

template <int DIN_>
struct SaDxSmem {
  using C = SaCfg<DIN_>;
  static constexpr uint32_t X = 0;
  static constexpr uint32_t DZ = al128(X + op_bytes(C::KE));
  static constexpr uint32_t DS = al128(DZ + op_bytes(DIN_));
  static constexpr uint32_t DV = al128(DS + 2 * kPHalf);
  static constexpr uint32_t Z = al128(DV + op_bytes(96));
  static constexpr uint32_t W = al128(Z + op_bytes(DIN_));                      // bz | bv (hi)
  static constexpr uint32_t WF = al128(W + C::BZ_BYTES + C::BV_BYTES);          // [4][16 * DIN] fp32
  static constexpr uint32_t BAR = al128(WF + kClasses * kL * DIN_ * 4);
  static constexpr uint32_t BYTES = BAR + 64;
  static_assert(BYTES <= 232448, "SA feature-gradient kernel does not fit shared memory");
};

template <int DIN_>
__device__ __forceinline__ void sa_dx_body(const SaDxArgs& a, uint8_t* sm) {
  using S = SaDxSmem<DIN_>; using C = SaCfg<DIN_>;
  constexpr int DIN = DIN_;
  constexpr uint32_t COL_Z = 0, COL_DX = 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::BAR);      // [0] weights, [1] tile loads, [2] MMAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  float* wf_s = reinterpret_cast<float*>(sm + S::WF);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tiles = (a.batch + 7) / 8;
  uint8_t *xop = sm + S::X, *dz = sm + S::DZ, *ds = sm + S::DS, *dv = sm + S::DV, *zop = sm + S::Z, *wsm = sm + S::W;
  const uint8_t* bz = wsm; const uint8_t* bv = wsm + C::BZ_BYTES;
  constexpr uint32_t kXBytes = x_tile_bytes(DIN), kDzBytes = op_bytes(DIN), kDsBytes = 2 * kPHalf, kDvBytes = op_bytes(96);
  auto issue_tile = [&](int t) {
    mbar_arrive_expect_tx(&bars[1], kXBytes + kDzBytes + kDsBytes + kDvBytes);
    bulk_g2s(xop, static_cast<const uint8_t*>(a.x_tiles) + size_t(t) * kXBytes, kXBytes, &bars[1]);
    bulk_g2s(dz, static_cast<const uint8_t*>(a.dz_tiles) + size_t(t) * kDzBytes, kDzBytes, &bars[1]);
    bulk_g2s(ds, static_cast<const uint8_t*>(a.ds_tiles) + size_t(t) * kDsBytes, kDsBytes, &bars[1]);
    bulk_g2s(dv, static_cast<const uint8_t*>(a.dv_tiles) + size_t(t) * kDvBytes, kDvBytes, &bars[1]);
  };
  if (tid == 0) {
    for (int i = 0; i < 3; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(&bars[0], C::BZ_BYTES + C::BV_BYTES);
    bulk_g2s(wsm, a.blobs, C::BZ_BYTES + C::BV_BYTES, &bars[0]);
    if (int(blockIdx.x) < tiles) issue_tile(blockIdx.x);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  if (a.wf) {
    for (int i = tid; i < kClasses * kL * DIN; i += kWgThreads) {
      const int cc = i / (kL * DIN), j = i - cc * (kL * DIN);
      wf_s[i] = __ldg(a.wf + size_t(cc) * a.D + a.cls_off + j);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  mbar_wait(&bars[0], 0);
  WgCtx c = make_ctx(0, tmem, &bars[2]);
  uint32_t ph_ld = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    mbar_wait(&bars[1], ph_ld);
    ph_ld ^= 1;
    tc_fence_after_sync();
    // ---- Z = X M + u ----------------------------------------------------------------------------------------------
    if (c.wt == 0) {
      mma_steps(c.tmem + COL_Z, make_smem_desc(smem_u32(xop), kCS, kRS), 2 * kCS, make_smem_desc(smem_u32(bz), C::BZ_LBO, 128),
                2 * C::BZ_LBO, make_idesc_bf16(128, DIN, 0, 0), C::KE / 16, false);
      umma_commit(c.bar);
    }
    wg_wait_mma(c);
    acc_to_operand<DIN>(c, COL_Z, zop);
    wg_sync_for_mma(c);
    // ---- d(xn) = dZ M^T + dS^T Z + dV Wv, two M = 64 halves ------------------------------------------------------------
    if (c.wt == 0) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t d = c.tmem + (uint32_t(16 * h) << 16) + COL_DX;
        mma_steps(d, make_smem_desc(smem_u32(dz + h * 8 * kRS), kCS, kRS), 2 * kCS, make_smem_desc(smem_u32(bz), 128, C::BZ_LBO),
                  2 * 128, make_idesc_bf16(64, DIN, 0, 1), DIN / 16, false);
        mma_steps(d, make_smem_desc(smem_u32(ds + h * kPHalf), kRS, kPCS), 2 * kRS, make_smem_desc(smem_u32(zop + h * 8 * kRS), kRS, kCS),
                  2 * kRS, make_idesc_bf16(64, DIN, 1, 1), 4, true);
        mma_steps(d, make_smem_desc(smem_u32(dv + h * 8 * kRS), kCS, kRS), 2 * kCS, make_smem_desc(smem_u32(bv), 128, C::BV_LBO),
                  2 * 128, make_idesc_bf16(64, DIN, 0, 1), C::DV / 16, true);
      }
      umma_commit(c.bar);
    }
    wg_wait_mma(c);
    // the input images are dead: the next tile's land while this one's epilogue runs
    if (c.wt == 0 && tile + int(gridDim.x) < tiles) issue_tile(tile + int(gridDim.x));
    // ---- epilogue: my row = (sample b, chunk r) in the s-mapping -----------------------------------------------------------
    {
      const int b = tile * 8 + (c.rs >> 4), r = c.rs & 15;
      const bool live = b < a.batch;
      const float nrm = live ? __ldg(a.norms + b) : 1.f;
      const float inv = 1.0f / nrm;
      const float4 dl = live && a.wf ? __ldg(reinterpret_cast<const float4*>(a.dlogits) + b) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float* frow = a.feat + size_t(live ? b : 0) * (kL * DIN) + r * DIN;
      const float* wrow = wf_s + r * DIN;
      const uint32_t col0 = uint32_t(a.cls_off + r * DIN);
      auto grad16 = [&](int c0, float (&g)[16], float (&xn)[16]) {
        ld16f(c.tmem + c.lane_base + COL_DX + c0, g);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const float4 f = live ? __ldg(reinterpret_cast<const float4*>(frow + c0) + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
          xn[4 * q4] = f.x * inv; xn[4 * q4 + 1] = f.y * inv; xn[4 * q4 + 2] = f.z * inv; xn[4 * q4 + 3] = f.w * inv;
          if (a.wf) {
            float m[4] = {1.f, 1.f, 1.f, 1.f};
            if (a.drop.thresh) drop_quad(a.drop, uint32_t(b), col0 + uint32_t(c0 + 4 * q4), m);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = c0 + 4 * q4 + e;
              const float t = fmaf(dl.x, wrow[j], fmaf(dl.y, wrow[kL * DIN + j], fmaf(dl.z, wrow[2 * kL * DIN + j], dl.w * wrow[3 * kL * DIN + j])));
              g[4 * q4 + e] = fmaf(m[e], t, g[4 * q4 + e]);
            }
          }
        }
      };
      float dot = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < DIN; c0 += 16) {
        float g[16], xn[16];
        grad16(c0, g, xn);
#pragma unroll
        for (int e = 0; e < 16; ++e) dot = fmaf(xn[e], g[e], dot);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);      // the sample's 16 chunks: 16 lanes
#pragma unroll 1
      for (int c0 = 0; c0 < DIN; c0 += 16) {
        float g[16], xn[16];
        grad16(c0, g, xn);
        if (live) {
          float* dst = a.d_feat + size_t(b) * (kL * DIN) + r * DIN + c0;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            *reinterpret_cast<float4*>(dst + 4 * q4) =
                make_float4((g[4 * q4] - xn[4 * q4] * dot) * inv, (g[4 * q4 + 1] - xn[4 * q4 + 1] * dot) * inv,
                            (g[4 * q4 + 2] - xn[4 * q4 + 2] * dot) * inv, (g[4 * q4 + 3] - xn[4 * q4 + 3] * dot) * inv);
        }
      }
    }
    tc_fence_before_sync();
    named_bar_sync(1, kWgThreads);      // TMEM columns and Z are reused by the next tile
    tc_fence_after_sync();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

constexpr uint32_t kSaDxSmemBytes = SaDxSmem<80>::BYTES > SaDxSmem<48>::BYTES ? SaDxSmem<80>::BYTES : SaDxSmem<48>::BYTES;
__global__ void __launch_bounds__(kWgThreads, 1) sa_dx_kernel(const SaDxBothArgs a) {
  extern __shared__ __align__(128) uint8_t sm[];
  if (blockIdx.y == 0) sa_dx_body<80>(a.m[0], sm);
  else sa_dx_body<48>(a.m[1], sm);
}

}  // namespace htc
}  // namespace mmrca

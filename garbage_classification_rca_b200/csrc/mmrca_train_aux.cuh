// Around the head: the feature hand-off from the stock backbones and the optimizer step over
// the flat head-parameter bucket.  Both are HBM-bound streaming kernels (coalesced 16-byte accesses, grids sized in
// multiples of the SM count); neither is a GEMM.
//
//   feature_handoff   text:  CLS row of the last hidden state  hidden[b][0][:]          (multimodal_model.py:651-658)
//                     image: global average pool of the final feature map + flatten     (:25-36: avgpool, flatten)
//                     -> [B, d_txt], [B, C] in bf16 (MMRCA_FLAG_FEATURES_BF16 input of the head) or fp32, one launch,
//                     no intermediate torch tensors (the strided [:, 0] view, .float(), .contiguous() copies).
//   sgd_step          torch.optim.SGD(lr, momentum, dampening, weight_decay, nesterov)  (main_both.py:548-549)
//   adamw_step        torch.optim.AdamW(lr, betas, eps, weight_decay)                   (main_both.py:545-546)
//                     over ONE contiguous fp32 parameter bucket and its gradient bucket (the bucket the backward kernels
//                     accumulate into and the data-parallel step all-reduces): one launch instead of 34 x per-tensor
//                     foreach kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mmrca {
namespace aux {

struct HandoffArgs {
  const void* hidden;          // [B][T][d_txt], row b starts at b * hidden_bstride elements; token 0 is read
  long long hidden_bstride;
  int hidden_bf16, d_txt;
  const void* fmap;            // [B][C][HW] (channels_last == 0) or [B][HW][C] (channels_last == 1), contiguous
  int fmap_bf16, channels, hw, channels_last;
  void* txt_out; void* img_out;   // [B][d_txt], [B][C]
  int out_bf16, batch;
};

__device__ __forceinline__ float ld_elem(const void* p, long long i, int bf16) {
  if (bf16) return __uint_as_float(uint32_t(static_cast<const uint16_t*>(p)[i]) << 16);
  return static_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_elem(void* p, long long i, float v, int bf16) {
  if (bf16) static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else static_cast<float*>(p)[i] = v;
}

// grid = (ceil(C / 32) + 1, B), 256 threads.  blockIdx.x == 0: the CLS row; else 32 channels of the pooled image vector.
// NCHW: a warp owns 4 channels in turn, lanes stride over the HW contiguous elements (coalesced), shuffle reduce.
// NHWC: thread (c, part) strides over pixels with the 32 channels of the block contiguous across the lanes.
__global__ void __launch_bounds__(256) feature_handoff_kernel(const HandoffArgs a) {
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (blockIdx.x == 0) {
    const long long base = (long long)b * a.hidden_bstride;
    for (int j = tid; j < a.d_txt; j += 256) st_elem(a.txt_out, (long long)b * a.d_txt + j, ld_elem(a.hidden, base + j, a.hidden_bf16), a.out_bf16);
    return;
  }
  const int c0 = (int(blockIdx.x) - 1) * 32;
  const float inv = 1.0f / float(a.hw);
  if (!a.channels_last) {
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + warp * 4 + k;
      if (c >= a.channels) break;
      const long long base = ((long long)b * a.channels + c) * a.hw;
      float s = 0.f;
      for (int i = lane; i < a.hw; i += 32) s += ld_elem(a.fmap, base + i, a.fmap_bf16);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) st_elem(a.img_out, (long long)b * a.channels + c, s * inv, a.out_bf16);
    }
  } else {
    __shared__ float part[8][33];
    const int c = c0 + lane;
    float s = 0.f;
    if (c < a.channels)
      for (int i = warp; i < a.hw; i += 8) s += ld_elem(a.fmap, ((long long)b * a.hw + i) * a.channels + c, a.fmap_bf16);
    part[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && c < a.channels) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += part[w][lane];
      st_elem(a.img_out, (long long)b * a.channels + c, t * inv, a.out_bf16);
    }
  }
}

// ---- optimizer steps over the flat bucket (n % 4 == 0, 16-byte aligned: FlatGrads / FlatParams guarantee both) ------------
struct SgdArgs {
  float* p; const float* g; float* buf;      // buf: momentum buffer (null when momentum == 0)
  long long n4;                               // float4 count
  float lr, momentum, dampening, weight_decay;
  int nesterov, first_step;                   // first_step: torch initialises the buffer with the gradient (no dampening)
};
__global__ void __launch_bounds__(256) sgd_step_kernel(const SgdArgs a) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < a.n4; i += (long long)gridDim.x * 256) {
    float4 p = reinterpret_cast<float4*>(a.p)[i];
    const float4 g4 = reinterpret_cast<const float4*>(a.g)[i];
    float g[4] = {g4.x, g4.y, g4.z, g4.w}, w[4] = {p.x, p.y, p.z, p.w};
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.buf && !a.first_step) b4 = reinterpret_cast<float4*>(a.buf)[i];
    float m[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float d = fmaf(a.weight_decay, w[e], g[e]);                // d_p = g + wd * p
      if (a.buf) {
        m[e] = a.first_step ? d : fmaf(a.momentum, m[e], (1.0f - a.dampening) * d);
        d = a.nesterov ? fmaf(a.momentum, m[e], d) : m[e];
      }
      w[e] = fmaf(-a.lr, d, w[e]);
    }
    reinterpret_cast<float4*>(a.p)[i] = make_float4(w[0], w[1], w[2], w[3]);
    if (a.buf) reinterpret_cast<float4*>(a.buf)[i] = make_float4(m[0], m[1], m[2], m[3]);
  }
}

struct AdamwArgs {
  float* p; const float* g; float* m; float* v;
  long long n4;
  float lr, beta1, beta2, eps, weight_decay;
  float bias1, bias2_sqrt;                    // 1 - beta1^t, sqrt(1 - beta2^t)  (computed on the host in double)
};
__global__ void __launch_bounds__(256) adamw_step_kernel(const AdamwArgs a) {
  const float step_size = a.lr / a.bias1;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < a.n4; i += (long long)gridDim.x * 256) {
    const float4 p4 = reinterpret_cast<float4*>(a.p)[i], g4 = reinterpret_cast<const float4*>(a.g)[i];
    const float4 m4 = reinterpret_cast<float4*>(a.m)[i], v4 = reinterpret_cast<float4*>(a.v)[i];
    float w[4] = {p4.x, p4.y, p4.z, p4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w};
    float m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      w[e] *= 1.0f - a.lr * a.weight_decay;                      // decoupled weight decay
      m[e] = fmaf(a.beta1, m[e], (1.0f - a.beta1) * g[e]);
      v[e] = fmaf(a.beta2, v[e], (1.0f - a.beta2) * g[e] * g[e]);
      const float denom = sqrtf(v[e]) / a.bias2_sqrt + a.eps;
      w[e] -= step_size * (m[e] / denom);
    }
    reinterpret_cast<float4*>(a.p)[i] = make_float4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<float4*>(a.m)[i] = make_float4(m[0], m[1], m[2], m[3]);
    reinterpret_cast<float4*>(a.v)[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

}  // namespace aux
}  // namespace mmrca

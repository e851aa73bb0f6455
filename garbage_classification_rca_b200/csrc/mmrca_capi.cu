// C ABI of the MM-RCA head (include/mmrca.h): argument checking, workspace carve-up and
// kernel launches.  No allocation, no synchronisation: everything is enqueued on the
// caller's stream.  There is no CPU path — without an sm_100 device every call fails.
#include "../../include/mmrca.h"

#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "mmrca_attn_fp32.cuh"
#include "mmrca_misc_fp32.cuh"
#include "mmrca_dropout.cuh"
#include "mmrca_tc_selftest.cuh"
#include "mmrca_head_tc.cuh"
#include "mmrca_head_tc_bwd.cuh"
#include "mmrca_head_tc_dx.cuh"
#include "mmrca_hier.cuh"
#include "mmrca_peer.cuh"
#include "mmrca_train_aux.cuh"
#include "mmrca_fusion_fp32.cuh"
#include "mmrca_token.cuh"
#include "mmrca_token_bwd.cuh"

namespace mmrca {

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;
static thread_local long long* g_dbg = nullptr;   // development aid, per calling thread (mmrca_dev_set_debug)
static thread_local int g_dbg_kernel = 0;         // 0: sa_bwd, 1: ca_bwd, 2: ca_fwd, 3: sa_fwd   // development: device buffer for per-phase clock stamps (mmrca_dev_set_debug)

// ---- optional per-kernel timing (mmrca_timing_begin / _end) -------------------------------------
struct TimingRec { const char* name; cudaEvent_t e0, e1; };
static thread_local TimingRec* g_trec = nullptr;
static thread_local int g_tcap = 0, g_tn = 0;

struct LaunchScope {   // counts the launch and, when timing is on, brackets it with events on its stream
  cudaStream_t st; int idx;
  LaunchScope(const char* name, cudaStream_t s) : st(s), idx(-1) {
    ++g_launches;
    if (g_trec && g_tn < g_tcap) {
      idx = g_tn++;
      g_trec[idx].name = name;
      cudaEventRecord(g_trec[idx].e0, st);
    }
  }
  ~LaunchScope() { if (idx >= 0) cudaEventRecord(g_trec[idx].e1, st); }
};

static int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}

#define MMRCA_CUDA(call)                                                                   \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) return fail(MMRCA_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

static size_t align_up_256(size_t v) { return (v + 255) & ~size_t(255); }

// Side streams for the independent GEMMs of one call (token-level blocks: the four weight / input gradient GEMMs of a
// cross block, its two projections): fork from the caller's stream with an event, join back with events - small
// latency-bound launches then overlap instead of queueing.  Per calling thread and device (two host threads never share
// events); the fork / join pattern is legal under stream capture.
struct SideStreams {
  bool ready = false;
  cudaStream_t s[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[3] = {nullptr, nullptr, nullptr};
};
static thread_local SideStreams g_side[16][4];      // [device][hash of the caller's stream]: calls on different streams (the image
                                                    // and the text branch of a step) mostly get their own side streams
static int side_streams(SideStreams** out, cudaStream_t st) {
  int dev = 0;
  MMRCA_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return fail(MMRCA_ERR_INVALID, "device index out of range%s%s");
  const uintptr_t h = reinterpret_cast<uintptr_t>(st);
  SideStreams& p = g_side[dev][((h >> 4) ^ (h >> 9) ^ (h >> 14)) & 3];
  if (!p.ready) {
    MMRCA_CUDA(cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming));
    for (int i = 0; i < 3; ++i) {
      MMRCA_CUDA(cudaStreamCreateWithFlags(&p.s[i], cudaStreamNonBlocking));
      MMRCA_CUDA(cudaEventCreateWithFlags(&p.join[i], cudaEventDisableTiming));
    }
    p.ready = true;
  }
  *out = &p;
  return MMRCA_OK;
}
static int side_fork(SideStreams* ss, cudaStream_t st, int n) {
  MMRCA_CUDA(cudaEventRecord(ss->fork, st));
  for (int i = 0; i < n; ++i) MMRCA_CUDA(cudaStreamWaitEvent(ss->s[i], ss->fork, 0));
  return MMRCA_OK;
}
static int side_join(SideStreams* ss, cudaStream_t st, int n) {
  for (int i = 0; i < n; ++i) {
    MMRCA_CUDA(cudaEventRecord(ss->join[i], ss->s[i]));
    MMRCA_CUDA(cudaStreamWaitEvent(st, ss->join[i], 0));
  }
  return MMRCA_OK;
}

struct DeviceInfo { int ok; int sms; };

static int device_info(DeviceInfo* out) {
  // one packed word per device, published with release / read with acquire: concurrent first calls from several
  // threads (nn.DataParallel) compute the same value and store it atomically
  static std::atomic<uint32_t> cache[64];       // bit 31: valid, bit 30: ok, low bits: SM count
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    cudaGetLastError();
    return fail(MMRCA_ERR_NO_DEVICE, "no CUDA device: the MM-RCA head has no CPU fallback%s%s");
  }
  uint32_t word = cache[dev].load(std::memory_order_acquire);
  if (!(word & 0x80000000u)) {
    int major = 0, sms = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    word = 0x80000000u | (major == 10 ? 0x40000000u : 0u) | (uint32_t(sms) & 0xffffu);
    cache[dev].store(word, std::memory_order_release);
  }
  out->ok = (word & 0x40000000u) ? 1 : 0;
  out->sms = int(word & 0xffffu);
  if (!out->ok) return fail(MMRCA_ERR_NO_DEVICE, "device is not compute capability 10.x (sm_100a kernels only)%s%s");
  return MMRCA_OK;
}

template <class K>
static int set_smem(K kernel, size_t bytes) {
  MMRCA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
  return MMRCA_OK;
}

// ---- attention block dispatch -------------------------------------------------------------------
template <int DIN, int DKQ, int DV, bool SELF, int GF, int GB>
static int launch_attn(bool backward, const AttnArgs& a, int sms, cudaStream_t st) {
  if (!backward) {
    using C = AttnCfg<DIN, DKQ, DV, SELF, GF, false>;
    int rc = set_smem(attn_fwd_kernel<C>, C::SMEM_BYTES);
    if (rc) return rc;
    const int tiles = (a.batch + C::G - 1) / C::G;
    LaunchScope ls(SELF ? (DIN == 48 ? "attn_fwd<48,128,96,self>" : DIN == 64 ? "attn_fwd<64,128,96,self>"
                                                                             : "attn_fwd<80,128,96,self>")
                        : "attn_fwd<96,64,48,cross>", st);
    attn_fwd_kernel<C><<<min(tiles, sms), kThreads, C::SMEM_BYTES, st>>>(a);
  } else {
    using C = AttnCfg<DIN, DKQ, DV, SELF, GB, true>;
    int rc = set_smem(attn_bwd_kernel<C>, C::SMEM_BYTES);
    if (rc) return rc;
    const int tiles = (a.batch + C::G - 1) / C::G;
    LaunchScope ls(SELF ? (DIN == 48 ? "attn_bwd<48,128,96,self>" : DIN == 64 ? "attn_bwd<64,128,96,self>"
                                                                             : "attn_bwd<80,128,96,self>")
                        : "attn_bwd<96,64,48,cross>", st);
    attn_bwd_kernel<C><<<min(tiles, sms), kThreads, C::SMEM_BYTES, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}

static int attn_dispatch(bool backward, bool self, int d_in, int d_kq, int d_v, const AttnArgs& a, int sms,
                         cudaStream_t st) {
  if (a.batch <= 0) return MMRCA_OK;
  if (self && d_kq == MMRCA_SA_DKQ && d_v == MMRCA_SA_DV) {
    if (d_in == 48) return launch_attn<48, 128, 96, true, 4, 4>(backward, a, sms, st);
    if (d_in == 64) return launch_attn<64, 128, 96, true, 4, 2>(backward, a, sms, st);
    if (d_in == 80) return launch_attn<80, 128, 96, true, 4, 2>(backward, a, sms, st);
  }
  if (!self && d_in == MMRCA_SA_DV && d_kq == MMRCA_CA_DKQ && d_v == MMRCA_CA_DV)
    return launch_attn<96, 64, 48, false, 4, 4>(backward, a, sms, st);
  return fail(MMRCA_ERR_INVALID,
              "unsupported attention block shape: need self (d_in in {48,64,80},128,96) or cross (96,64,48)%s%s");
}

static AttnArgs make_attn_args(const MmrcaAttnParams& p, const float* xq, const float* xkv, int batch, int reverse) {
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.xq = xq; a.xkv = xkv;
  a.wq = p.wq; a.bq = p.bq; a.wk = p.wk; a.bk = p.bk; a.wv = p.wv; a.bv = p.bv;
  a.ln_g = p.ln_g; a.ln_b = p.ln_b;
  a.batch = batch; a.reverse = reverse;
  return a;
}

static int launch_wgrad(int K, const float* dy, int ldy, int col0, int ncols, const float* x, const float* norms,
                        int rows, float* dw, float* db, int sms, cudaStream_t st) {
  const int ny = (ncols + kWgCols - 1) / kWgCols;
  const int tiles = (rows + kWgRows - 1) / kWgRows;
  int gx = (4 * sms + ny - 1) / ny;
  if (gx > tiles) gx = tiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx, ny);
  LaunchScope ls(K == 48 ? "wgrad<48>" : K == 64 ? "wgrad<64>" : K == 80 ? "wgrad<80>" : "wgrad<96>", st);
  switch (K) {
    case 48: wgrad_kernel<48><<<grid, kThreads, 0, st>>>(dy, ldy, col0, ncols, x, norms, rows, dw, db); break;
    case 64: wgrad_kernel<64><<<grid, kThreads, 0, st>>>(dy, ldy, col0, ncols, x, norms, rows, dw, db); break;
    case 80: wgrad_kernel<80><<<grid, kThreads, 0, st>>>(dy, ldy, col0, ncols, x, norms, rows, dw, db); break;
    case 96: wgrad_kernel<96><<<grid, kThreads, 0, st>>>(dy, ldy, col0, ncols, x, norms, rows, dw, db); break;
    default: return fail(MMRCA_ERR_INVALID, "unsupported wgrad K%s%s");
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}

// dW/db of one block from the row-space grads dy = [dQ|dK|dV] (written by attn_bwd_kernel)
static int attn_wgrads(bool self, int d_in, int d_kq, int d_v, const float* dy, const float* xq, const float* xkv,
                       const float* norms, int batch, const MmrcaAttnGrads& g, int sms, cudaStream_t st) {
  const int nall = 2 * d_kq + d_v, rows = batch * kL;
  int rc;
  // W_query/W_key/W_value are separate tensors in the state_dict, so three (or, for a shared
  // source, still three) reductions with their own destination pointers.
  if ((rc = launch_wgrad(d_in, dy, nall, 0, d_kq, xq, norms, rows, g.wq, g.bq, sms, st))) return rc;
  if ((rc = launch_wgrad(d_in, dy, nall, d_kq, d_kq, self ? xq : xkv, norms, rows, g.wk, g.bk, sms, st))) return rc;
  if ((rc = launch_wgrad(d_in, dy, nall, 2 * d_kq, d_v, self ? xq : xkv, norms, rows, g.wv, g.bv, sms, st))) return rc;
  return MMRCA_OK;
}

// ---- classifier dispatch --------------------------------------------------------------------------
template <int NC>
static void launch_classifier(bool backward, const CatArgs& a, int sms, cudaStream_t st) {
  LaunchScope ls(backward ? "classifier_bwd" : "classifier_fwd", st);
  if (!backward) {
    int grid = (a.batch + kWarps - 1) / kWarps;
    if (grid > 8 * sms) grid = 8 * sms;
    classifier_fwd_kernel<NC><<<grid, kThreads, 0, st>>>(a);
  } else {
    const int ny = (a.D + kThreads * 4 - 1) / (kThreads * 4);
    int gx = (4 * sms + ny - 1) / ny;
    if (gx > a.batch) gx = a.batch;
    classifier_bwd_kernel<NC><<<dim3(gx, ny), kThreads, 0, st>>>(a);
  }
}

static int classifier_dispatch(bool backward, int nc, const CatArgs& a, int sms, cudaStream_t st) {
  if (a.batch <= 0) return MMRCA_OK;
  switch (nc) {
    case 1: launch_classifier<1>(backward, a, sms, st); break;
    case 2: launch_classifier<2>(backward, a, sms, st); break;
    case 3: launch_classifier<3>(backward, a, sms, st); break;
    case 4: launch_classifier<4>(backward, a, sms, st); break;
    case 5: launch_classifier<5>(backward, a, sms, st); break;
    case 6: launch_classifier<6>(backward, a, sms, st); break;
    case 7: launch_classifier<7>(backward, a, sms, st); break;
    case 8: launch_classifier<8>(backward, a, sms, st); break;
    default: return fail(MMRCA_ERR_INVALID, "n_classes must be in [1, 8]%s%s");
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}

// ---- workspace --------------------------------------------------------------------------------------
struct Workspace {
  float *norm_img, *norm_txt, *t_sa, *i_sa, *t_i, *i_t;          // forward (kept for the backward)
  float *d_t_sa, *d_i_sa, *d_t_i, *d_i_t, *dy, *dlogits;          // training only
  uint8_t* mask;                                                    // seeded dropout materialised for the fp32 kernels
  // fused bf16 pipeline (mmrca_head_tc.cuh)
  void* fblob[4];                                                   // per block: bz | bv | bc blobs
  void* t_img; void* i_img;                                         // SA output images, [tiles][kSaTileBytes]
  void* x_img; void* x_txt;                                         // normalised feature images (SA inputs)
  void* dx_img[4];                                                  // training: dXq / dXkv images of CA1, then CA2
  void* sa_v[2]; void* sa_p[2];                                     // training: V / P images kept by the SA forward (img, txt)
  float2* sa_stats[2];                                              // training: LayerNorm (mean, rstd) per context row
  uint4* ca_row[2];                                                 // training: CA rows {keep bits, mean, rstd} per direction
  float* gm[4]; size_t gm_floats;                                   // training: dM_ext^T per block (contiguous)
  void* sa_dz[2]; void* sa_ds[2]; void* sa_dv[2];                   // training + feature gradients: dZ / dS / dV images of the SA backward
  float* step_loss = nullptr;                                       // train step: prep_feat clears gm and *step_loss (no memset nodes)
  float* zero_grads = nullptr; int zero_grads_n = 0;               // train step + MMRCA_FLAG_ZERO_GRADS: the gradient bucket to clear
  size_t bytes;
};

static size_t align_up(size_t v) { return align_up_256(v); }

static int concat_width(const MmrcaHeadDesc& d) {
  const int ca = 2 * kL * MMRCA_CA_DV;
  if (d.flags & MMRCA_FLAG_FEATURES_ONLY) return d.d_img + d.d_txt;
  if (d.flags & MMRCA_FLAG_CROSS_ATTENTION_ONLY) return ca;
  return ca + d.d_img + d.d_txt;
}

// seeded dropout of the concat (mmrca_dropout.cuh) from the desc; thresh == 0 when off
static DropSpec make_drop(const MmrcaHeadDesc& d) {
  DropSpec s;
  memset(&s, 0, sizeof(s));
  s.D = concat_width(d);
  s.scale = 1.0f;
  if (d.drop_p > 0.f) {
    const float p = d.drop_p < 1.f ? d.drop_p : 1.f;
    s.seed_lo = uint32_t(d.drop_seed & 0xffffffffu);
    s.seed_hi = uint32_t(d.drop_seed >> 32);
    s.thresh = uint32_t(p * 65536.0f + 0.5f);
    s.scale = p < 1.f ? 1.0f / (1.0f - p) : 0.f;
  }
  return s;
}

// The bf16 tensor-core pipeline covers the reference's literal dimensions (multimodal_model.py:249-258: 1280 / 768
// features, 4 classes) with frozen features; every other desc runs on the fp32 kernels (never on the CPU).  A pure
// function of the desc: it also decides which buffers the workspace holds.
static bool desc_is_tc(const MmrcaHeadDesc& d) {
  // (features_only with feature gradients is two streaming kernels' worth of work: it stays on the fp32 kernels)
  return d.compute != MMRCA_COMPUTE_FP32 && d.d_img == 1280 && d.d_txt == 768 && d.n_classes == 4 &&
         !((d.flags & MMRCA_FLAG_FEATURE_GRADS) && (d.flags & MMRCA_FLAG_FEATURES_ONLY));
}

static Workspace carve(const MmrcaHeadDesc& d, bool training, void* base) {
  Workspace w;
  memset(&w, 0, sizeof(w));
  char* p = static_cast<char*>(base);
  size_t off = 0;
  const size_t B = size_t(d.batch > 0 ? d.batch : 0);
  const bool tc = desc_is_tc(d);
  auto take = [&](size_t floats) { float* r = reinterpret_cast<float*>(p + off); off += align_up(floats * 4); return r; };
  w.norm_img = take(B); w.norm_txt = take(B);
  if (!tc) {      // fp32 SIMT kernels: block outputs in fp32 rows, a materialised dropout mask
    w.t_sa = take(B * kL * MMRCA_SA_DV); w.i_sa = take(B * kL * MMRCA_SA_DV);
    w.t_i = take(B * kL * MMRCA_CA_DV); w.i_t = take(B * kL * MMRCA_CA_DV);
    if (d.drop_p > 0.f) w.mask = reinterpret_cast<uint8_t*>(take((B * size_t(concat_width(d)) + 3) / 4));
  } else {        // bf16 pipeline: weight blobs and operand images
    const size_t tiles = (B + 7) / 8;
    w.fblob[0] = take(htc::SaCfg<80>::W_BYTES / 4);
    w.fblob[1] = take(htc::SaCfg<48>::W_BYTES / 4);
    w.fblob[2] = take(htc::CaCfg::W_BYTES / 4);
    w.fblob[3] = take(htc::CaCfg::W_BYTES / 4);
    w.t_img = take(tiles * htc::kSaTileBytes / 4);
    w.i_img = take(tiles * htc::kSaTileBytes / 4);
    w.x_img = take(tiles * htc::x_tile_bytes(80) / 4);
    w.x_txt = take(tiles * htc::x_tile_bytes(48) / 4);
  }
  if (training) {
    w.dlogits = take(B * size_t(d.n_classes > 0 ? d.n_classes : 0));
    if (!tc) {
      w.d_t_sa = take(B * kL * MMRCA_SA_DV); w.d_i_sa = take(B * kL * MMRCA_SA_DV);
      w.d_t_i = take(B * kL * MMRCA_CA_DV); w.d_i_t = take(B * kL * MMRCA_CA_DV);
      w.dy = take(B * kL * (2 * MMRCA_SA_DKQ + MMRCA_SA_DV));
    } else {
      const size_t tiles = (B + 7) / 8;
      for (int i = 0; i < 4; ++i) w.dx_img[i] = take(tiles * htc::kSaTileBytes / 4);
      for (int i = 0; i < 2; ++i) w.ca_row[i] = reinterpret_cast<uint4*>(take(tiles * (128 * 16 / 4)));
      for (int i = 0; i < 2; ++i) {
        w.sa_v[i] = take(tiles * htc::kSaTileBytes / 4); w.sa_p[i] = take(tiles * (2 * htc::kPHalf) / 4);
        w.sa_stats[i] = reinterpret_cast<float2*>(take(tiles * 256));
      }
      if (d.flags & MMRCA_FLAG_FEATURE_GRADS) {
        const int sdin[2] = {80, 48};
        for (int i = 0; i < 2; ++i) {
          w.sa_dz[i] = take(tiles * htc::op_bytes(sdin[i]) / 4); w.sa_ds[i] = take(tiles * (2 * htc::kPHalf) / 4);
          w.sa_dv[i] = take(tiles * htc::kSaTileBytes / 4);
        }
      }
      const int dins[4] = {80, 48, MMRCA_SA_DV, MMRCA_SA_DV};
      char* g0 = p + off;
      for (int i = 0; i < 4; ++i) { w.gm[i] = reinterpret_cast<float*>(p + off); off += size_t(dins[i]) * 128 * 4; }
      w.gm_floats = size_t(p + off - g0) / 4;
      off = align_up(off);
    }
  }
  w.bytes = off;
  return w;
}

static int check_desc(const MmrcaHeadDesc* d) {
  if (!d) return fail(MMRCA_ERR_INVALID, "null desc%s%s");
  if (d->batch < 0) return fail(MMRCA_ERR_INVALID, "negative batch%s%s");
  auto okdim = [](int v) { return v == 768 || v == 1024 || v == 1280; };
  if (!okdim(d->d_img) || !okdim(d->d_txt))
    return fail(MMRCA_ERR_INVALID, "d_img / d_txt must be one of 768, 1024, 1280 (16 chunks of 48/64/80)%s%s");
  if (d->n_classes < 1 || d->n_classes > 8) return fail(MMRCA_ERR_INVALID, "n_classes must be in [1, 8]%s%s");
  if ((d->flags & MMRCA_FLAG_FEATURES_ONLY) && (d->flags & MMRCA_FLAG_CROSS_ATTENTION_ONLY)) {
    // the reference's if/elif gives features_only precedence (multimodal_model.py:694-726)
  }
  if (d->compute != MMRCA_COMPUTE_FP32 && d->compute != MMRCA_COMPUTE_BF16 && d->compute != MMRCA_COMPUTE_BF16_FUSED)
    return fail(MMRCA_ERR_INVALID, "unknown compute mode%s%s");
  if (!(d->drop_p >= 0.f && d->drop_p <= 1.f)) return fail(MMRCA_ERR_INVALID, "drop_p must be in [0, 1]%s%s");
  if ((d->flags & MMRCA_FLAG_FEATURES_BF16) && !desc_is_tc(*d))
    return fail(MMRCA_ERR_INVALID, "MMRCA_FLAG_FEATURES_BF16 needs the bf16 pipeline (compute = MMRCA_COMPUTE_BF16, 1280 / 768 "
                                   "features, 4 classes, no feature gradients)%s%s");
  return MMRCA_OK;
}

// a caller-drawn dropout mask exists for parity with torch's Philox stream: fp32 kernels only
static int check_mask(const MmrcaHeadDesc& d, const uint8_t* mask) {
  if (mask && desc_is_tc(d))
    return fail(MMRCA_ERR_INVALID, "a caller-supplied drop_mask needs compute = MMRCA_COMPUTE_FP32 (the bf16 pipeline draws "
                                   "its seeded mask on chip: desc.drop_p / drop_seed)%s%s");
  return MMRCA_OK;
}

// concat order: multimodal_model.py:694-716
static CatArgs make_cat_args(const MmrcaHeadDesc& d, const Workspace& w, const float* img, const float* txt,
                             const uint8_t* mask, float scale, const float* wf, const float* bf) {
  CatArgs c;
  memset(&c, 0, sizeof(c));
  const bool fo = d.flags & MMRCA_FLAG_FEATURES_ONLY, co = !fo && (d.flags & MMRCA_FLAG_CROSS_ATTENTION_ONLY);
  int n = 0;
  if (!fo) {
    c.seg[n].src = w.t_i; c.seg[n].width = kL * MMRCA_CA_DV; ++n;
    c.seg[n].src = w.i_t; c.seg[n].width = kL * MMRCA_CA_DV; ++n;
  }
  if (!co) {
    c.seg[n].src = img; c.seg[n].norms = w.norm_img; c.seg[n].width = d.d_img; ++n;
    c.seg[n].src = txt; c.seg[n].norms = w.norm_txt; c.seg[n].width = d.d_txt; ++n;
  }
  c.nseg = n;
  c.D = concat_width(d);
  c.batch = d.batch;
  c.mask = mask; c.scale = scale;
  c.wf = wf; c.bf = bf;
  return c;
}

// L2 norms of the raw features when no attention block computes them (features_only skips the
// attention blocks: the reference runs and discards them, multimodal_model.py:676-692).
__global__ void __launch_bounds__(kThreads) l2norm_kernel(const float* __restrict__ x, float* __restrict__ norms,
                                                          int batch, int d) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = blockIdx.x * kWarps + warp; b < batch; b += gridDim.x * kWarps) {
    const float4* xs = reinterpret_cast<const float4*>(x + size_t(b) * d);
    float ss = 0.f;
    for (int j = lane; j < d / 4; j += 32) {
      const float4 v = __ldg(xs + j);
      ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss))));
    }
    ss = warp_sum(ss);
    if (lane == 0) norms[b] = sqrtf(ss);
  }
}

// ---- fused bf16 pipeline: forward -----------------------------------------------------------------------------
// blob of a self-attention block: bz | bv | bv (lo); of a cross-attention direction: bz | bv | bc | bv (lo) | bc (lo)
static htc::PrepBlock make_prep_block(const MmrcaAttnParams& p, void* blob, int din, int dkq, int dv, bool cross) {
  htc::PrepBlock b;
  b.wq = p.wq; b.bq = p.bq; b.wk = p.wk; b.wv = p.wv; b.bv = p.bv;
  b.bz = blob;
  b.bvb = static_cast<uint8_t*>(blob) + htc::blob_bytes(din, din + 16);
  b.bvb_lo = static_cast<uint8_t*>(blob) + (cross ? htc::CaCfg::OFF_BV_LO : htc::blob_bytes(din, din + 16) + htc::blob_bytes(dv, din + 16));
  b.din = din; b.dkq = dkq; b.dv = dv;
  return b;
}

// weights -> bf16 blobs; features -> normalised bf16 images, norms, logits = bias + fp32 feature terms
static int launch_prep_feat(const MmrcaHeadDesc& d, const MmrcaHeadParams& p, const float* img, const float* txt,
                            float* logits, const Workspace& w, int sms, cudaStream_t st) {
  const bool fo = d.flags & MMRCA_FLAG_FEATURES_ONLY, co = !fo && (d.flags & MMRCA_FLAG_CROSS_ATTENTION_ONLY);
  htc::PrepArgs a;
  memset(&a, 0, sizeof(a));
  // --features_only (multimodal_model.py:694-699, :721-722): the attention blocks do not reach the logits; only the
  // feature half of this kernel runs (no weight blobs), then ce_feat: two streaming kernels for forward + backward
  if (!fo) {
  a.blk[0] = make_prep_block(p.sa_img, w.fblob[0], 80, MMRCA_SA_DKQ, MMRCA_SA_DV, false);
  a.blk[1] = make_prep_block(p.sa_txt, w.fblob[1], 48, MMRCA_SA_DKQ, MMRCA_SA_DV, false);
  a.blk[2] = make_prep_block(p.ca1, w.fblob[2], MMRCA_SA_DV, MMRCA_CA_DKQ, MMRCA_CA_DV, true);
  a.blk[3] = make_prep_block(p.ca2, w.fblob[3], MMRCA_SA_DV, MMRCA_CA_DKQ, MMRCA_CA_DV, true);
  const int ca = kL * MMRCA_CA_DV;   // 768: width of T_I / I_T in the concat (multimodal_model.py:708-716)
  a.src[0].bc = static_cast<uint8_t*>(w.fblob[2]) + htc::CaCfg::OFF_BC;
  a.src[0].bc_lo = static_cast<uint8_t*>(w.fblob[2]) + htc::CaCfg::OFF_BC_LO;
  a.src[0].off = 0; a.src[0].w = MMRCA_CA_DV;
  a.src[1].bc = static_cast<uint8_t*>(w.fblob[3]) + htc::CaCfg::OFF_BC;
  a.src[1].bc_lo = static_cast<uint8_t*>(w.fblob[3]) + htc::CaCfg::OFF_BC_LO;
  a.src[1].off = ca; a.src[1].w = MMRCA_CA_DV;
  a.nsrc = 2;      // (the feature sources of the classifier run in fp32: prep_feat_kernel / ce_feat_kernel)
  }
  a.wf = p.wf; a.D = concat_width(d);
  htc::FeatArgs f;
  memset(&f, 0, sizeof(f));
  const int feat0 = fo ? 0 : 2 * kL * MMRCA_CA_DV;      // first concat column of the image features (:694-716)
  f.src[0].feat = img; f.src[0].x_tiles = w.x_img; f.src[0].norms = w.norm_img; f.src[0].cls_off = feat0;
  f.src[1].feat = txt; f.src[1].x_tiles = w.x_txt; f.src[1].norms = w.norm_txt; f.src[1].cls_off = feat0 + d.d_img;
  f.logits = logits; f.wf = p.wf; f.bf = p.bf; f.with_features = co ? 0 : 1;
  f.drop = make_drop(d); f.batch = d.batch;
  f.feat_bf16 = (d.flags & MMRCA_FLAG_FEATURES_BF16) ? 1 : 0;
  if (w.step_loss) { a.zero0 = w.gm[0]; a.nzero0 = int(w.gm_floats); a.zero1 = w.step_loss; }
  a.zero2 = w.zero_grads; a.nzero2 = w.zero_grads_n;
  const int prep_ctas = htc::kPrepCtas;
  const int feat_ctas = max(1, min((d.batch + 7) / 8, 2 * sms));
  {
    LaunchScope ls("prep_feat", st);
    htc::prep_feat_kernel<<<prep_ctas + feat_ctas, 256, htc::kFeatSmemBytes, st>>>(a, f, prep_ctas);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}

static int head_forward_fused(const MmrcaHeadDesc& d, const MmrcaHeadParams& p, const float* img, const float* txt,
                              float* logits, const Workspace& w, int sms, cudaStream_t st) {
  int rc;
  if ((rc = launch_prep_feat(d, p, img, txt, logits, w, sms, st))) return rc;
  if (d.flags & MMRCA_FLAG_FEATURES_ONLY) return MMRCA_OK;
  const int tiles = (d.batch + 7) / 8;
  const int grid = min(tiles, sms);
  {
    htc::SaFwdArgs a;
    memset(&a, 0, sizeof(a));
    a.role[0].x_tiles = w.x_img; a.role[0].ln_g = p.sa_img.ln_g; a.role[0].ln_b = p.sa_img.ln_b;
    a.role[0].blobs = w.fblob[0]; a.role[0].out_tiles = w.i_img; a.role[0].v_tiles = w.sa_v[0]; a.role[0].p_tiles = w.sa_p[0]; a.role[0].ln_stats = w.sa_stats[0];
    a.role[1].x_tiles = w.x_txt; a.role[1].ln_g = p.sa_txt.ln_g; a.role[1].ln_b = p.sa_txt.ln_b;
    a.role[1].blobs = w.fblob[1]; a.role[1].out_tiles = w.t_img; a.role[1].v_tiles = w.sa_v[1]; a.role[1].p_tiles = w.sa_p[1]; a.role[1].ln_stats = w.sa_stats[1];
    a.batch = d.batch;
    a.dbg = g_dbg_kernel == 3 ? g_dbg : nullptr;
    if ((rc = set_smem(htc::sa_fwd_kernel, htc::SaFwdLayout::BYTES))) return rc;
    LaunchScope ls("sa_fwd_bf16", st);
    htc::sa_fwd_kernel<<<grid, htc::kCtaThreads, htc::SaFwdLayout::BYTES, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  {
    htc::CaFwdArgs a;
    memset(&a, 0, sizeof(a));
    a.dir[0].blobs = w.fblob[2]; a.dir[0].ln_g = p.ca1.ln_g; a.dir[0].ln_b = p.ca1.ln_b;
    a.dir[1].blobs = w.fblob[3]; a.dir[1].ln_g = p.ca2.ln_g; a.dir[1].ln_b = p.ca2.ln_b;
    a.t_tiles = w.t_img; a.i_tiles = w.i_img;
    a.logits = logits; a.batch = d.batch; a.reverse = (d.flags & MMRCA_FLAG_REVERSE) ? 1 : 0;
    a.drop = make_drop(d);
    a.row_out[0] = w.ca_row[0]; a.row_out[1] = w.ca_row[1];
    a.dbg = g_dbg_kernel == 2 ? g_dbg : nullptr;
    if ((rc = set_smem(htc::ca_fwd_kernel, htc::CaFwdLayout::BYTES))) return rc;
    LaunchScope ls("ca_fwd_bf16", st);
    htc::ca_fwd_kernel<<<dim3(min((tiles + 1) / 2, max(1, sms / 2)), 2), htc::kCtaThreads, htc::CaFwdLayout::BYTES, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}

// ---- fused bf16 pipeline: backward --------------------------------------------------------------------------
static int head_backward_fused(const MmrcaHeadDesc& d, const MmrcaHeadParams& p, const float* img, const float* txt,
                               const float* dlogits, const MmrcaHeadGrads& g, float* d_img, float* d_txt,
                               const Workspace& w, int sms, cudaStream_t st) {
  const int tiles = (d.batch + 7) / 8, D = concat_width(d), ca = kL * MMRCA_CA_DV;
  int rc;
  if (d.flags & MMRCA_FLAG_FEATURES_ONLY) return MMRCA_OK;      // ce_feat has done everything there is to do
  if (!w.step_loss) MMRCA_CUDA(cudaMemsetAsync(w.gm[0], 0, w.gm_floats * sizeof(float), st));
  {
    htc::CaBwdArgs a;
    memset(&a, 0, sizeof(a));
    const MmrcaAttnParams* ap[2] = {&p.ca1, &p.ca2};
    const MmrcaAttnGrads* ag[2] = {&g.ca1, &g.ca2};
    for (int i = 0; i < 2; ++i) {
      a.dir[i].blobs = w.fblob[2 + i]; a.dir[i].ln_g = ap[i]->ln_g; a.dir[i].ln_b = ap[i]->ln_b;
      a.dir[i].gm = w.gm[2 + i]; a.dir[i].g_wv = ag[i]->wv; a.dir[i].g_bv = ag[i]->bv;
      a.dir[i].g_ln_g = ag[i]->ln_g; a.dir[i].g_ln_b = ag[i]->ln_b;
      a.dir[i].g_wf = g.wf + i * ca;
      a.dir[i].dxq_img = w.dx_img[2 * i]; a.dir[i].dxkv_img = w.dx_img[2 * i + 1];
      a.dir[i].row_in = w.ca_row[i];
    }
    a.t_tiles = w.t_img; a.i_tiles = w.i_img; a.dlogits = dlogits; a.D = D;
    a.batch = d.batch; a.reverse = (d.flags & MMRCA_FLAG_REVERSE) ? 1 : 0;
    a.drop = make_drop(d);
    a.dbg = g_dbg_kernel == 1 ? g_dbg : nullptr;
    if ((rc = set_smem(htc::ca_bwd_kernel, htc::CaBwdSmem::BYTES))) return rc;
    LaunchScope ls("ca_bwd_bf16", st);
    htc::ca_bwd_kernel<<<dim3(min(tiles, max(1, sms / 2)), 2), htc::kCaBwdThreads, htc::CaBwdSmem::BYTES, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  {
    // text SA output: query source of CA1, key/value source of CA2; image SA output: the other way round
    htc::SaBwdBothArgs both;
    memset(&both, 0, sizeof(both));
    {
      htc::SaBwdArgs& a = both.m[0];
      a.dbg = g_dbg_kernel == 0 ? g_dbg : nullptr;
      a.x_tiles = w.x_img; a.v_tiles = w.sa_v[0]; a.p_tiles = w.sa_p[0]; a.ln_stats = w.sa_stats[0]; a.ln_g = p.sa_img.ln_g; a.ln_b = p.sa_img.ln_b;
      a.dout_a = w.dx_img[1]; a.dout_b = w.dx_img[2];
      a.gm = w.gm[0]; a.g_wv = g.sa_img.wv; a.g_bv = g.sa_img.bv; a.g_ln_g = g.sa_img.ln_g; a.g_ln_b = g.sa_img.ln_b;
      a.batch = d.batch;
      if (d_img) { a.dz_tiles = w.sa_dz[0]; a.ds_tiles = w.sa_ds[0]; a.dv_tiles = w.sa_dv[0]; }
    }
    {
      htc::SaBwdArgs& a = both.m[1];
      a.x_tiles = w.x_txt; a.v_tiles = w.sa_v[1]; a.p_tiles = w.sa_p[1]; a.ln_stats = w.sa_stats[1]; a.ln_g = p.sa_txt.ln_g; a.ln_b = p.sa_txt.ln_b;
      a.dout_a = w.dx_img[0]; a.dout_b = w.dx_img[3];
      a.gm = w.gm[1]; a.g_wv = g.sa_txt.wv; a.g_bv = g.sa_txt.bv; a.g_ln_g = g.sa_txt.ln_g; a.g_ln_b = g.sa_txt.ln_b;
      a.batch = d.batch;
      if (d_txt) { a.dz_tiles = w.sa_dz[1]; a.ds_tiles = w.sa_ds[1]; a.dv_tiles = w.sa_dv[1]; }
    }
    if ((rc = set_smem(htc::sa_bwd_kernel, htc::kSaBwdSmemBytes))) return rc;
    {
      LaunchScope ls("sa_bwd_bf16", st);
      htc::sa_bwd_kernel<<<dim3(min(tiles, max(1, sms / 2)), 2), htc::kCtaThreads, htc::kSaBwdSmemBytes, st>>>(both);
    }
    MMRCA_CUDA(cudaGetLastError());
  }
  if (d_img && d_txt) {      // fine-tune phase: feature gradients from the images sa_bwd just left
    const bool co = (d.flags & MMRCA_FLAG_CROSS_ATTENTION_ONLY) != 0;
    htc::SaDxBothArgs both;
    memset(&both, 0, sizeof(both));
    const float* feats[2] = {img, txt};
    const float* norms[2] = {w.norm_img, w.norm_txt};
    const void* xt[2] = {w.x_img, w.x_txt};
    float* outs[2] = {d_img, d_txt};
    const int offs[2] = {2 * kL * MMRCA_CA_DV, 2 * kL * MMRCA_CA_DV + d.d_img};
    for (int i = 0; i < 2; ++i) {
      htc::SaDxArgs& a = both.m[i];
      a.x_tiles = xt[i]; a.dz_tiles = w.sa_dz[i]; a.ds_tiles = w.sa_ds[i]; a.dv_tiles = w.sa_dv[i]; a.blobs = w.fblob[i];
      a.feat = feats[i]; a.norms = norms[i]; a.dlogits = dlogits; a.wf = co ? nullptr : p.wf; a.cls_off = offs[i]; a.D = D;
      a.drop = make_drop(d); a.d_feat = outs[i]; a.batch = d.batch;
    }
    if ((rc = set_smem(htc::sa_dx_kernel, htc::kSaDxSmemBytes))) return rc;
    {
      LaunchScope ls("sa_dx_bf16", st);
      htc::sa_dx_kernel<<<dim3(min(tiles, max(1, sms / 2)), 2), htc::kWgThreads, htc::kSaDxSmemBytes, st>>>(both);
    }
    MMRCA_CUDA(cudaGetLastError());
  }
  {
    htc::FinArgs a;
    memset(&a, 0, sizeof(a));
    const MmrcaAttnParams* ap[4] = {&p.sa_img, &p.sa_txt, &p.ca1, &p.ca2};
    const MmrcaAttnGrads* ag[4] = {&g.sa_img, &g.sa_txt, &g.ca1, &g.ca2};
    const int dins[4] = {80, 48, MMRCA_SA_DV, MMRCA_SA_DV};
    const int dkqs[4] = {MMRCA_SA_DKQ, MMRCA_SA_DKQ, MMRCA_CA_DKQ, MMRCA_CA_DKQ};
    for (int i = 0; i < 4; ++i) {
      a.blk[i].gm = w.gm[i]; a.blk[i].wq = ap[i]->wq; a.blk[i].bq = ap[i]->bq; a.blk[i].wk = ap[i]->wk;
      a.blk[i].g_wq = ag[i]->wq; a.blk[i].g_bq = ag[i]->bq; a.blk[i].g_wk = ag[i]->wk;
      a.blk[i].din = dins[i]; a.blk[i].dkq = dkqs[i];
    }
    a.nblk = 4;
    int ctas = 0;
    for (int i = 0; i < 4; ++i) ctas += dkqs[i] / htc::kFinRows;
    LaunchScope ls("finalize_bf16", st);
    htc::finalize_kernel<<<ctas, 256, htc::kFinSmemBytes, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}

// cross-entropy (labels != null) or given dlogits, + classifier bias gradient + the feature-source rows of dWf
static int launch_ce_feat(const MmrcaHeadDesc& d, const float* logits, const int64_t* labels, const MmrcaCeDesc* ce,
                          float* loss, float* dlogits, const MmrcaHeadGrads& g, bool bias_grad, const Workspace& w,
                          cudaStream_t st) {
  // the reference's if / elif gives features_only precedence over cross_attention_only (multimodal_model.py:694-726)
  const bool fo = d.flags & MMRCA_FLAG_FEATURES_ONLY, co = !fo && (d.flags & MMRCA_FLAG_CROSS_ATTENTION_ONLY);
  if (labels && !w.step_loss) MMRCA_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  htc::CeFeatArgs a;
  memset(&a, 0, sizeof(a));
  a.logits = logits; a.labels = labels; a.cw = ce ? ce->class_weight : nullptr; a.eps = ce ? ce->label_smoothing : 0.f;
  a.batch = d.batch; a.dlogits = dlogits; a.loss = loss; a.g_bf = bias_grad ? g.bf : nullptr;
  a.drop = make_drop(d);
  if (!co && g.wf) {
    a.x_img = w.x_img; a.x_txt = w.x_txt; a.g_wf = g.wf;
    a.off_img = fo ? 0 : 2 * kL * MMRCA_CA_DV; a.off_txt = a.off_img + d.d_img;
  }
  {
    const int tiles = (d.batch + 7) / 8;
    const dim3 grid(a.g_wf ? htc::kStripsImg + htc::kStripsTxt : 1, (tiles + htc::kCeSliceTiles - 1) / htc::kCeSliceTiles);
    LaunchScope ls("ce_feat", st);
    htc::ce_feat_kernel<<<grid, 256, 0, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}

static int head_forward_impl(const MmrcaHeadDesc& d, const MmrcaHeadParams& p, const float* img, const float* txt,
                             const uint8_t* mask, float scale, float* logits, const Workspace& w, int sms,
                             cudaStream_t st) {
  const bool fo = d.flags & MMRCA_FLAG_FEATURES_ONLY;
  const int rev = (d.flags & MMRCA_FLAG_REVERSE) ? 1 : 0;
  int rc;
  if (d.batch == 0) return MMRCA_OK;
  if ((rc = check_mask(d, mask))) return rc;
  if (desc_is_tc(d)) return head_forward_fused(d, p, img, txt, logits, w, sms, st);
  if (!mask && d.drop_p > 0.f) {      // seeded dropout: the fp32 kernels read the materialised mask
    const DropSpec ds = make_drop(d);
    {
      LaunchScope ls("dropout_mask", st);
      dropout_mask_kernel<<<min(4 * sms, max(1, int((size_t(d.batch) * ds.D / 4 + 255) / 256))), 256, 0, st>>>(ds, d.batch, w.mask);
    }
    MMRCA_CUDA(cudaGetLastError());
    mask = w.mask; scale = ds.scale;
  }
  if (fo) {
    const int grid = min((d.batch + kWarps - 1) / kWarps, 8 * sms);
    { LaunchScope ls("l2norm", st); l2norm_kernel<<<grid, kThreads, 0, st>>>(img, w.norm_img, d.batch, d.d_img); }
    { LaunchScope ls("l2norm", st); l2norm_kernel<<<grid, kThreads, 0, st>>>(txt, w.norm_txt, d.batch, d.d_txt); }
    MMRCA_CUDA(cudaGetLastError());
  } else {
    AttnArgs a = make_attn_args(p.sa_txt, txt, txt, d.batch, 0);            // multimodal_model.py:677-678
    a.normalise = 1; a.norms = w.norm_txt; a.out = w.t_sa;
    if ((rc = attn_dispatch(false, true, d.d_txt / kL, MMRCA_SA_DKQ, MMRCA_SA_DV, a, sms, st))) return rc;
    a = make_attn_args(p.sa_img, img, img, d.batch, 0);                      // :679-680
    a.normalise = 1; a.norms = w.norm_img; a.out = w.i_sa;
    if ((rc = attn_dispatch(false, true, d.d_img / kL, MMRCA_SA_DKQ, MMRCA_SA_DV, a, sms, st))) return rc;
    a = make_attn_args(p.ca1, w.t_sa, w.i_sa, d.batch, rev);                 // :683-684
    a.out = w.t_i;
    if ((rc = attn_dispatch(false, false, MMRCA_SA_DV, MMRCA_CA_DKQ, MMRCA_CA_DV, a, sms, st))) return rc;
    a = make_attn_args(p.ca2, w.i_sa, w.t_sa, d.batch, rev);                 // :685-686
    a.out = w.i_t;
    if ((rc = attn_dispatch(false, false, MMRCA_SA_DV, MMRCA_CA_DKQ, MMRCA_CA_DV, a, sms, st))) return rc;
  }
  CatArgs c = make_cat_args(d, w, img, txt, mask, scale, p.wf, p.bf);
  c.logits = logits;
  return classifier_dispatch(false, d.n_classes, c, sms, st);
}

static int head_backward_impl(const MmrcaHeadDesc& d, const MmrcaHeadParams& p, const float* img, const float* txt,
                              const uint8_t* mask, float scale, const float* dlogits, const MmrcaHeadGrads& g,
                              float* d_img, float* d_txt, const Workspace& w, int sms, cudaStream_t st,
                              bool ce_done = false) {
  const bool fo = d.flags & MMRCA_FLAG_FEATURES_ONLY, co = !fo && (d.flags & MMRCA_FLAG_CROSS_ATTENTION_ONLY);
  const int rev = (d.flags & MMRCA_FLAG_REVERSE) ? 1 : 0;
  const bool want_feat = d_img != nullptr || d_txt != nullptr;
  int rc;
  if (d.batch == 0) return MMRCA_OK;
  if (want_feat && !(d_img && d_txt))
    return fail(MMRCA_ERR_INVALID, "d_img_feat and d_txt_feat must both be given or both be NULL%s%s");
  if ((rc = check_mask(d, mask))) return rc;
  if (desc_is_tc(d)) {
    if (want_feat && !(d.flags & MMRCA_FLAG_FEATURE_GRADS))
      return fail(MMRCA_ERR_INVALID, "feature gradients need MMRCA_FLAG_FEATURE_GRADS in the desc of the forward AND "
                                     "the backward (it sizes the workspace)%s%s");
    if (want_feat && (d.flags & MMRCA_FLAG_FEATURES_BF16))
      return fail(MMRCA_ERR_INVALID, "feature gradients are taken with respect to fp32 features (MMRCA_FLAG_FEATURES_BF16 is "
                                     "for frozen backbones)%s%s");
    // classifier bias gradient and the feature-source rows of dWf from the given dlogits (train_step has done
    // both inside its cross-entropy kernel)
    if (!ce_done && (rc = launch_ce_feat(d, nullptr, nullptr, nullptr, nullptr, const_cast<float*>(dlogits), g, g.bf != nullptr,
                                         w, st))) return rc;
    return head_backward_fused(d, p, img, txt, dlogits, g, want_feat ? d_img : nullptr, want_feat ? d_txt : nullptr, w, sms, st);
  }
  if (!mask && d.drop_p > 0.f) { mask = w.mask; scale = make_drop(d).scale; }   // materialised by the forward
  // 1. classifier: dWf, dbf, d(T_I), d(I_T) and the direct feature terms d(img_n), d(txt_n)
  CatArgs c = make_cat_args(d, w, img, txt, mask, scale, p.wf, p.bf);
  c.dlogits = dlogits; c.g_wf = g.wf; c.g_bf = g.bf;
  {
    int n = 0;
    if (!fo) { c.seg[n++].dst = w.d_t_i; c.seg[n++].dst = w.d_i_t; }
    if (!co) { c.seg[n++].dst = want_feat ? d_img : nullptr; c.seg[n++].dst = want_feat ? d_txt : nullptr; }
  }
  if ((rc = classifier_dispatch(true, d.n_classes, c, sms, st))) return rc;
  if (!fo) {
    // 2. cross_attention_1(text SA -> q, image SA -> k,v): d_t_sa = dq-path, d_i_sa = dkv-path
    AttnArgs a = make_attn_args(p.ca1, w.t_sa, w.i_sa, d.batch, rev);
    a.dout = w.d_t_i; a.dy = w.dy; a.dxq = w.d_t_sa; a.dxkv = w.d_i_sa; a.acc_dxq = 0; a.acc_dxkv = 0;
    a.g_ln_g = g.ca1.ln_g; a.g_ln_b = g.ca1.ln_b;
    if ((rc = attn_dispatch(true, false, MMRCA_SA_DV, MMRCA_CA_DKQ, MMRCA_CA_DV, a, sms, st))) return rc;
    if ((rc = attn_wgrads(false, MMRCA_SA_DV, MMRCA_CA_DKQ, MMRCA_CA_DV, w.dy, w.t_sa, w.i_sa, nullptr, d.batch,
                          g.ca1, sms, st))) return rc;
    // 3. cross_attention_2(image SA -> q, text SA -> k,v): accumulates into both
    a = make_attn_args(p.ca2, w.i_sa, w.t_sa, d.batch, rev);
    a.dout = w.d_i_t; a.dy = w.dy; a.dxq = w.d_i_sa; a.dxkv = w.d_t_sa; a.acc_dxq = 1; a.acc_dxkv = 1;
    a.g_ln_g = g.ca2.ln_g; a.g_ln_b = g.ca2.ln_b;
    if ((rc = attn_dispatch(true, false, MMRCA_SA_DV, MMRCA_CA_DKQ, MMRCA_CA_DV, a, sms, st))) return rc;
    if ((rc = attn_wgrads(false, MMRCA_SA_DV, MMRCA_CA_DKQ, MMRCA_CA_DV, w.dy, w.i_sa, w.t_sa, nullptr, d.batch,
                          g.ca2, sms, st))) return rc;
    // 4. self-attention blocks; their input grads land on the normalised features
    a = make_attn_args(p.sa_txt, txt, txt, d.batch, 0);
    a.normalise = 1; a.norms = w.norm_txt;
    a.dout = w.d_t_sa; a.dy = w.dy; a.dxq = want_feat ? d_txt : nullptr; a.acc_dxq = co ? 0 : 1;
    a.g_ln_g = g.sa_txt.ln_g; a.g_ln_b = g.sa_txt.ln_b;
    if ((rc = attn_dispatch(true, true, d.d_txt / kL, MMRCA_SA_DKQ, MMRCA_SA_DV, a, sms, st))) return rc;
    if ((rc = attn_wgrads(true, d.d_txt / kL, MMRCA_SA_DKQ, MMRCA_SA_DV, w.dy, txt, txt, w.norm_txt, d.batch,
                          g.sa_txt, sms, st))) return rc;
    a = make_attn_args(p.sa_img, img, img, d.batch, 0);
    a.normalise = 1; a.norms = w.norm_img;
    a.dout = w.d_i_sa; a.dy = w.dy; a.dxq = want_feat ? d_img : nullptr; a.acc_dxq = co ? 0 : 1;
    a.g_ln_g = g.sa_img.ln_g; a.g_ln_b = g.sa_img.ln_b;
    if ((rc = attn_dispatch(true, true, d.d_img / kL, MMRCA_SA_DKQ, MMRCA_SA_DV, a, sms, st))) return rc;
    if ((rc = attn_wgrads(true, d.d_img / kL, MMRCA_SA_DKQ, MMRCA_SA_DV, w.dy, img, img, w.norm_img, d.batch,
                          g.sa_img, sms, st))) return rc;
  }
  // 5. through x / ||x||
  if (want_feat) {
    const int grid = min((d.batch + kWarps - 1) / kWarps, 8 * sms);
    { LaunchScope ls("l2norm_bwd", st); l2norm_bwd_kernel<<<grid, kThreads, 0, st>>>(img, w.norm_img, d_img, d.batch, d.d_img); }
    { LaunchScope ls("l2norm_bwd", st); l2norm_bwd_kernel<<<grid, kThreads, 0, st>>>(txt, w.norm_txt, d_txt, d.batch, d.d_txt); }
    MMRCA_CUDA(cudaGetLastError());
  }
  return MMRCA_OK;
}

static int launch_ce(const float* logits, const int64_t* labels, const MmrcaCeDesc* ce, int batch, int nc,
                     float* loss, float* dlogits, cudaStream_t st) {
  if (nc < 1 || nc > kCeMaxClasses) return fail(MMRCA_ERR_INVALID, "n_classes must be in [1, 16]%s%s");
  if (batch <= 0) return fail(MMRCA_ERR_INVALID, "cross entropy needs batch > 0%s%s");
  {
    LaunchScope ls("cross_entropy", st);
    cross_entropy_kernel<<<1, kCeThreads, 0, st>>>(logits, labels, ce ? ce->class_weight : nullptr,
                                                   ce ? ce->label_smoothing : 0.f, batch, nc, loss, dlogits);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}

// ---- hierarchical head ----------------------------------------------------------------------------------------------
struct HierWorkspace {
  void* x_img; void* x_txt; void* wb_img; void* wb_txt; void* h; void* dh; float* dlogits;
  float* dcat;      // MMRCA_HIER_FEATURE_GRADS: d(dropped concat) fp32 [tiles * 128][8192]
  size_t bytes;
};
static HierWorkspace hier_carve(int batch, void* base, uint32_t flags = 0) {
  HierWorkspace w;
  memset(&w, 0, sizeof(w));
  const size_t tiles = size_t(batch + hier::kTile - 1) / hier::kTile;
  char* p = static_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) { void* r = p ? p + off : nullptr; off += align_up_256(bytes); return r; };
  w.x_img = take(tiles * hier::kGImg * hier::kGrpBytes);
  w.x_txt = take(tiles * hier::kGTxt * hier::kGrpBytes);
  w.wb_img = take(size_t(hier::kGImg) * hier::kHid * 16);
  w.wb_txt = take(size_t(hier::kGTxt) * hier::kHid * 16);
  w.h = take(tiles * hier::kTile * (2 * hier::kHid) * 2);
  w.dh = take(tiles * hier::kGHid * hier::kGrpBytes);
  w.dlogits = static_cast<float*>(take(tiles * hier::kTile * hier::kClasses * sizeof(float)));
  if (flags & MMRCA_HIER_FEATURE_GRADS) w.dcat = static_cast<float*>(take(tiles * hier::kTile * size_t(hier::kD) * sizeof(float)));
  w.bytes = off;
  return w;
}
static int hier_check(const MmrcaHierDesc* d) {
  if (!d) return fail(MMRCA_ERR_INVALID, "null descriptor%s%s");
  if (d->batch < 0) return fail(MMRCA_ERR_INVALID, "batch must be >= 0%s%s");
  if (d->n_classes != hier::kClasses) return fail(MMRCA_ERR_INVALID, "the hierarchical head is built for 4 classes%s%s");
  if (!(d->drop_p >= 0.f && d->drop_p <= 1.f)) return fail(MMRCA_ERR_INVALID, "drop_p must be in [0, 1]%s%s");
  return MMRCA_OK;
}
static DropSpec hier_drop(const MmrcaHierDesc& d) {
  DropSpec s;
  memset(&s, 0, sizeof(s));
  s.D = hier::kD;
  s.scale = 1.0f;
  if (d.drop_p > 0.f) {
    const float p = d.drop_p < 1.f ? d.drop_p : 1.f;
    s.seed_lo = uint32_t(d.drop_seed & 0xffffffffu);
    s.seed_hi = uint32_t(d.drop_seed >> 32);
    s.thresh = uint32_t(p * 65536.0f + 0.5f);
    s.scale = p < 1.f ? 1.0f / (1.0f - p) : 0.f;
  }
  return s;
}
static int hier_forward_impl(const MmrcaHierDesc& d, const MmrcaHierParams& p, const float* const* feats,
                             const uint8_t* mask, float scale, float* logits, const HierWorkspace& w, cudaStream_t st) {
  if (d.batch == 0) return MMRCA_OK;
  int rc;
  for (int i = 0; i < 6; ++i)
    if (!feats[i] || (reinterpret_cast<uintptr_t>(feats[i]) & 15))
      return fail(MMRCA_ERR_INVALID, "the six feature pointers must be non-null and 16-byte aligned%s%s");
  const int tiles = (d.batch + hier::kTile - 1) / hier::kTile;
  // the weight images (16 MB, a few microseconds) are built on a side stream next to the HBM-bound feature pass
  SideStreams* ss;
  if ((rc = side_streams(&ss, st)) || (rc = side_fork(ss, st, 1))) return rc;
  {
    hier::WPrepArgs a;
    a.w[0] = p.w_img; a.w[1] = p.w_txt; a.blob[0] = w.wb_img; a.blob[1] = w.wb_txt;
    LaunchScope ls("hier_wprep", ss->s[0]);
    hier::hier_wprep_kernel<<<dim3(hier::kHid / 32, hier::kGImg / 32 + hier::kGTxt / 32), 256, 0, ss->s[0]>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  {
    hier::PrepArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < 6; ++i) a.seg[i] = feats[i];
    a.x_img = w.x_img; a.x_txt = w.x_txt; a.mask = mask; a.mask_scale = scale; a.drop = hier_drop(d);
    a.logits = logits; a.b_all = p.b_all; a.batch = d.batch;
    LaunchScope ls("hier_prep", st);
    hier::hier_prep_kernel<<<tiles * hier::kTile, 256, 0, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  if ((rc = side_join(ss, st, 1))) return rc;
  {
    hier::GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.x[0] = w.x_img; a.x[1] = w.x_txt; a.wb[0] = w.wb_img; a.wb[1] = w.wb_txt; a.bias[0] = p.b_img; a.bias[1] = p.b_txt;
    a.w_all = p.w_all; a.logits = logits; a.h = w.h; a.batch = d.batch;
    if ((rc = set_smem(hier::hier_gemm_kernel, hier::FwdSmem::BYTES))) return rc;
    LaunchScope ls("hier_gemm", st);
    hier::hier_gemm_kernel<<<dim3(tiles, hier::kHid / hier::kBN, 2), hier::kGemmThreads, hier::FwdSmem::BYTES, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}
static int hier_backward_impl(const MmrcaHierDesc& d, const MmrcaHierParams& p, const float* dlogits,
                              const MmrcaHierGrads& g, const HierWorkspace& w, cudaStream_t st) {
  if (d.batch == 0) return MMRCA_OK;
  int rc;
  if (!g.w_img || !g.w_txt || !g.b_img || !g.b_txt)
    return fail(MMRCA_ERR_INVALID, "the hidden-layer gradient buffers must be given%s%s");
  if ((reinterpret_cast<uintptr_t>(g.w_img) | reinterpret_cast<uintptr_t>(g.w_txt)) & 15)
    return fail(MMRCA_ERR_INVALID, "weight-gradient buffers must be 16-byte aligned%s%s");
  const int tiles = (d.batch + hier::kTile - 1) / hier::kTile;
  {
    hier::DhArgs a;
    memset(&a, 0, sizeof(a));
    a.h = w.h; a.dlogits = dlogits; a.w_all = p.w_all; a.dh = w.dh; a.g_w_all = g.w_all; a.g_b_all = g.b_all;
    a.g_b_hid[0] = g.b_img; a.g_b_hid[1] = g.b_txt; a.batch = d.batch;
    LaunchScope ls("hier_dh", st);
    hier::hier_dh_kernel<<<dim3(tiles, hier::kGHid / 16), 256, 0, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  {
    hier::WgradArgs a;
    memset(&a, 0, sizeof(a));
    a.dh = w.dh; a.x[0] = w.x_img; a.x[1] = w.x_txt; a.g_w[0] = g.w_img; a.g_w[1] = g.w_txt; a.tiles = tiles;
    if ((rc = set_smem(hier::hier_wgrad_kernel, hier::WgSmem::BYTES))) return rc;
    LaunchScope ls("hier_wgrad", st);
    hier::hier_wgrad_kernel<<<hier::kWgCtasImg + hier::kWgCtasTxt, hier::kGemmThreads, hier::WgSmem::BYTES, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}

// ---- token-level attention blocks (mmrca_token.cuh) ------------------------------------------------------------------------
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
  // the driver entry point through the runtime: libmmrca.so does not link libcuda
  static std::atomic<void*> cached{nullptr};
  void* f = cached.load(std::memory_order_acquire);
  if (!f) {
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    cached.store(f, std::memory_order_release);
  }
  return reinterpret_cast<TensorMapEncodeFn>(f);
}
// bf16 [rows][cols] with a row pitch of ld elements (cols contiguous), box {64, box_rows}, 128-byte swizzle, out-of-range
// elements read as zero
static int make_tensor_map(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows) {
  TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc) return fail(MMRCA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver%s%s");
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {ld * 2};
  const cuuint32_t box[2] = {uint32_t(tok::kBK), box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MMRCA_ERR_CUDA, "cuTensorMapEncodeTiled failed (activations must be 16-byte aligned bf16, "
                                                     "d_in a multiple of 8)%s%s");
  return MMRCA_OK;
}
constexpr int kMaxSms = 160;      // bound on the split-K slabs of the token weight-gradient GEMM (B200: 148 SMs)
struct TokenWorkspace {
  __nv_bfloat16* w;       // stacked bf16 weights: self [2 d_kq + d_v][d_in]; cross [d_kq][d_in_q] then [d_kq + d_v][d_in_kv]
  float* bias;            // stacked [2 d_kq + d_v]
  void *q_img, *k_img, *v_img;
  // MMRCA_TOKEN_TRAINING: kept by the forward for the backward, and the backward's own buffers
  void* p_img; float* sum;
  __nv_bfloat16* g;             // gradient rows: self [rows][2 d_kq + d_v]; cross [rows][d_kq] then [rows][d_kq + d_v]
  float* wg_part;               // weight-gradient split-K slabs [splits][columns][d_in], splits * ceil(d_in / 128) <= SMs
  __nv_bfloat16* wbf;           // bf16 weights as they lie: self [2 d_kq + d_v][d_in]; cross [d_kq][d_in_q] then [d_kq + d_v][d_in_kv]
  size_t bytes;
};
static TokenWorkspace token_carve(const MmrcaTokenDesc& d, void* base) {
  TokenWorkspace w;
  memset(&w, 0, sizeof(w));
  char* p = static_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) { void* r = p ? p + off : nullptr; off += align_up_256(bytes); return r; };
  const size_t B = size_t(d.batch > 0 ? d.batch : 0), tps = size_t(d.seq_len + tok::kTile - 1) / tok::kTile;
  const size_t kq_pad = size_t(d.d_in_q + tok::kBK - 1) / tok::kBK * tok::kBK, kkv_pad = size_t(d.d_in_kv + tok::kBK - 1) / tok::kBK * tok::kBK;
  w.w = static_cast<__nv_bfloat16*>(take((size_t(d.d_kq) * kq_pad + size_t(d.d_kq + d.d_v) * kkv_pad) * 2));
  w.bias = static_cast<float*>(take(size_t(2 * d.d_kq + d.d_v) * 4));
  w.q_img = take(B * tps * htc::op_bytes(d.d_kq));
  w.k_img = take(B * tps * htc::op_bytes(d.d_kq));
  w.v_img = take(B * tps * htc::op_bytes(d.d_v));
  if (d.flags & MMRCA_TOKEN_TRAINING) {
    w.p_img = take(B * tps * htc::op_bytes(tok::kMaxTiles * tok::kTile));
    w.sum = static_cast<float*>(take(B * tps * tok::kTile * 4));
    const size_t rows = B * size_t(d.seq_len);
    w.g = static_cast<__nv_bfloat16*>(take(rows * size_t(2 * d.d_kq + d.d_v) * 2));
    w.wg_part = static_cast<float*>(take(size_t(kMaxSms) * size_t(2 * d.d_kq + d.d_v) * 128 * 4));
    w.wbf = static_cast<__nv_bfloat16*>(take((size_t(d.d_kq) * d.d_in_q + size_t(d.d_kq + d.d_v) * d.d_in_kv) * 2));
  }
  w.bytes = off;
  return w;
}
static int token_check(const MmrcaTokenDesc* d) {
  if (!d) return fail(MMRCA_ERR_INVALID, "null descriptor%s%s");
  if (d->batch < 0 || d->seq_len < 2 || d->seq_len > tok::kTile * tok::kMaxTiles)
    return fail(MMRCA_ERR_INVALID, "token attention: batch >= 0 and 2 <= seq_len <= 256%s%s");
  if (!((d->d_kq == 128 && d->d_v == 96) || (d->d_kq == 64 && d->d_v == 48)))
    return fail(MMRCA_ERR_INVALID, "token attention: (d_kq, d_v) must be (128, 96) or (64, 48) (multimodal_model.py:251-255)%s%s");
  if (d->d_in_q < 16 || (d->d_in_q & 7) || d->d_in_kv < 16 || (d->d_in_kv & 7))
    return fail(MMRCA_ERR_INVALID, "token attention: d_in must be a multiple of 8, >= 16%s%s");
  return MMRCA_OK;
}
static int launch_cast3(const tok::Cast3Args& a, int sms, cudaStream_t st) {
  const long long n = std::max(a.n[0], std::max(a.n[1], a.n[2]));
  {
    LaunchScope ls("cast_bf16", st);
    tok::cast3_bf16_kernel<<<int(std::min<long long>((n / 4 + 255) / 256 + 1, 4LL * sms)), 256, 0, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}
// the block's fp32 weights -> pre-swizzled bf16 images of the projection GEMM + the stacked bias, one launch
static int launch_tok_wprep(const tok::WprepArgs& a, int sms, cudaStream_t st) {
  long long chunks = 0;
  for (int i = 0; i < 3; ++i)
    chunks = std::max(chunks, (long long)((a.seg[i].K + tok::kBK - 1) / tok::kBK) * a.seg[i].n_rows * 8);
  {
    LaunchScope ls("tok_wprep", st);
    tok::tok_wprep_kernel<<<int(std::min<long long>((chunks + 255) / 256, 4LL * sms)), 256, 0, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}
// one projection GEMM: x [B][L][K] bf16, w [N][K] bf16 (rows: the segments back to back), bias [N]
static int launch_tok_proj(const MmrcaTokenDesc& d, const void* x, int K, const __nv_bfloat16* w, const float* bias, int N,
                           const tok::ProjSeg (&segs)[3], int sms, cudaStream_t st) {
  int rc;
  tok::ProjArgs a;
  memset(&a, 0, sizeof(a));
  a.wblob = reinterpret_cast<const uint8_t*>(w);
  for (int i = 0; i < 3; ++i) a.seg[i] = segs[i];
  a.bias = bias; a.N = N; a.K = K;
  a.nacc = N > 256 ? 2 : 1; a.bn = N / a.nacc;
  if (a.bn % 16 || a.bn * a.nacc != N) return fail(MMRCA_ERR_INVALID, "projection width must split into 16-column multiples%s%s");
  a.tiles_per_sample = (d.seq_len + tok::kTile - 1) / tok::kTile;
  a.L = d.seq_len; a.rows = d.batch * d.seq_len;
  CUtensorMap tx, tw;
  if ((rc = make_tensor_map(&tx, x, uint64_t(K), uint64_t(a.rows), uint64_t(K), tok::kTile))) return rc;
  tw = tx;      // (the weights arrive as pre-swizzled images through the bulk-copy engine: no second tensor map)
  const size_t smem = tok::kStages * size_t(tok::proj_stage_bytes(N)) + 128 + size_t(N) * 4 + 1024;
  if ((rc = set_smem(tok::tok_proj_kernel, smem))) return rc;
  {
    LaunchScope ls("tok_proj", st);
    tok::tok_proj_kernel<<<min((a.rows + tok::kTile - 1) / tok::kTile, sms), tok::kProjThreads, smem, st>>>(tx, tw, a);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}
template <int DKQ, int DV>
static int launch_tok_attn(const MmrcaTokenDesc& d, const tok::AttnArgs& a, cudaStream_t st) {
  int rc;
  using S = tok::AttnSmem<DKQ, DV>;
  if ((rc = set_smem(tok::tok_attn_kernel<DKQ, DV>, S::BYTES))) return rc;
  {
    LaunchScope ls(DKQ == 128 ? "tok_attn<128,96>" : "tok_attn<64,48>", st);
    tok::tok_attn_kernel<DKQ, DV><<<d.batch * a.tiles_per_sample, 256, S::BYTES, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}

template <int DKQ, int DV>
static int launch_tok_attn_bwd(const MmrcaTokenDesc& d, const tok::AttnBwdArgs& a, cudaStream_t st) {
  int rc;
  using S = tok::AttnBwdSmem<DKQ, DV>;
  if ((rc = set_smem(tok::tok_attn_bwd_kernel<DKQ, DV>, S::BYTES))) return rc;
  {
    LaunchScope ls(DKQ == 128 ? "tok_attn_bwd<128,96>" : "tok_attn_bwd<64,48>", st);
    tok::tok_attn_bwd_kernel<DKQ, DV><<<d.batch * a.tiles_per_sample, 256, S::BYTES, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}
// dW += G^T X for the gradient columns listed in a.out (x [rows][K] bf16, g [rows][NG] bf16 with row pitch ld_g)
static int launch_tok_wgrad(tok::WgradArgs a, const void* x, const __nv_bfloat16* g, int ld_g, float* part, int sms, cudaStream_t st,
                            int ld_x = 0) {
  int rc;
  CUtensorMap tx, tg;
  if ((rc = make_tensor_map(&tx, x, uint64_t(a.K), uint64_t(a.rows), uint64_t(ld_x ? ld_x : a.K), tok::kGradKT))) return rc;
  if ((rc = make_tensor_map(&tg, g, uint64_t(a.NG), uint64_t(a.rows), uint64_t(ld_g), tok::kGradKT))) return rc;
  const int nblk = (a.NG + 63) / 64, mtiles = (a.K + 127) / 128, chunks = (a.rows + tok::kGradKT - 1) / tok::kGradKT;
  const size_t smem = size_t(tok::kGradStages) * (2 + nblk) * tok::kBoxBytes + 128 + 1024;
  if ((rc = set_smem(tok::tok_wgrad_kernel, smem))) return rc;
  const int splits = std::max(1, std::min(chunks, std::min(sms, kMaxSms) / mtiles));
  a.part = part;
  {
    LaunchScope ls("tok_wgrad", st);
    tok::tok_wgrad_kernel<<<dim3(mtiles, splits), tok::kGradThreads, smem, st>>>(tx, tg, a);
  }
  MMRCA_CUDA(cudaGetLastError());
  {
    LaunchScope ls("tok_wgrad_reduce", st);
    const long long quads = (long long)a.NG * a.K / 4;
    if (quads >= 32768)
      tok::tok_wgrad_reduce_kernel<1><<<int(std::min<long long>((quads + 255) / 256, 16LL * sms)), 256, 0, st>>>(a, splits);
    else
      tok::tok_wgrad_reduce_kernel<8><<<int(std::min<long long>((quads + 31) / 32, 16LL * sms)), 256, 0, st>>>(a, splits);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}
// dx [rows][K] = G [rows][NG] W (+ bias); w bf16: [NG][K] (w_is_linear_weight = false: the input gradient dY W) or
// [K][NG] (true: torch.nn.Linear's weight, the forward X W^T + bias)
static int launch_tok_dgrad(float* dx, int rows, int K, int NG, const __nv_bfloat16* g, int ld_g, const __nv_bfloat16* w,
                            cudaStream_t st, bool w_is_linear_weight = false, const float* bias = nullptr, int ld_w = 0) {
  int rc;
  CUtensorMap tg, tw;
  if ((rc = make_tensor_map(&tg, g, uint64_t(NG), uint64_t(rows), uint64_t(ld_g), tok::kTile))) return rc;
  const int bn_max = std::min(256, K);
  if (w_is_linear_weight) rc = make_tensor_map(&tw, w, uint64_t(NG), uint64_t(K), uint64_t(ld_w ? ld_w : NG), uint32_t(bn_max));
  else rc = make_tensor_map(&tw, w, uint64_t(K), uint64_t(NG), uint64_t(ld_w ? ld_w : K), 64);
  if (rc) return rc;
  const int nb = (bn_max + 63) / 64;
  const size_t b_bytes = w_is_linear_weight ? size_t(bn_max) * 128 : size_t(nb) * tok::kBoxBytes;
  const size_t smem = size_t(tok::kGradStages) * ((tok::kTile * 128 + b_bytes + 1023) & ~size_t(1023)) + 128 + 1024;
  tok::DgradArgs a;
  a.dx = dx; a.rows = rows; a.K = K; a.NG = NG; a.bias = bias;
  const dim3 grid((rows + tok::kTile - 1) / tok::kTile, (K + 255) / 256);
  if (w_is_linear_weight) {
    if ((rc = set_smem(tok::tok_dgrad_kernel<true>, smem))) return rc;
    LaunchScope ls("tma_gemm_nt", st);
    tok::tok_dgrad_kernel<true><<<grid, tok::kGradThreads, smem, st>>>(tg, tw, a);
  } else {
    if ((rc = set_smem(tok::tok_dgrad_kernel<false>, smem))) return rc;
    LaunchScope ls("tok_dgrad", st);
    tok::tok_dgrad_kernel<false><<<grid, tok::kGradThreads, smem, st>>>(tg, tw, a);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}

// ---- classic / normalized fusion heads (mmrca_fusion_fp32.cuh) ---------------------------------------------------------
struct FusionWorkspace {
  float *h_img, *h_txt, *n_img, *n_txt, *c, *d_c, *d_h_img, *d_h_txt, *dlogits;
  uint8_t* mask;
  // MMRCA_FUSION_BF16: bf16 copies of the features, the two projection weights and d(hidden), split-K slabs of the weight gradients
  __nv_bfloat16 *x_bf[2], *w_bf[2], *dh_bf[2];
  __nv_bfloat16 *cat_bf, *wcat_bf, *dc_bf;      // [B][2H] (normalised) hidden vectors side by side, W_cat [H][2H], d(concat_layer output) [B][H]
  float* wg_part;
  size_t bytes;
};
static FusionWorkspace fusion_carve(const MmrcaFusionDesc& d, void* base) {
  FusionWorkspace w;
  memset(&w, 0, sizeof(w));
  char* p = static_cast<char*>(base);
  size_t off = 0;
  const size_t B = size_t(d.batch > 0 ? d.batch : 0), H = size_t(d.hidden > 0 ? d.hidden : 0);
  auto take = [&](size_t floats) { float* r = reinterpret_cast<float*>(p + off); off += align_up_256(floats * 4); return r; };
  w.h_img = take(B * H); w.h_txt = take(B * H); w.n_img = take(B); w.n_txt = take(B); w.c = take(B * H);
  w.mask = reinterpret_cast<uint8_t*>(take((B * H + 3) / 4));
  w.d_c = take(B * H); w.d_h_img = take(B * H); w.d_h_txt = take(B * H);
  w.dlogits = take(B * size_t(d.n_classes > 0 ? d.n_classes : 0));
  if (d.flags & MMRCA_FUSION_BF16) {
    const size_t dins[2] = {size_t(d.d_img > 0 ? d.d_img : 0), size_t(d.d_txt > 0 ? d.d_txt : 0)};
    for (int m = 0; m < 2; ++m) {
      w.x_bf[m] = reinterpret_cast<__nv_bfloat16*>(take((B * dins[m] + 1) / 2));
      w.w_bf[m] = reinterpret_cast<__nv_bfloat16*>(take((H * dins[m] + 1) / 2));
      w.dh_bf[m] = reinterpret_cast<__nv_bfloat16*>(take((B * H + 1) / 2));
    }
    w.cat_bf = reinterpret_cast<__nv_bfloat16*>(take(B * H));
    w.wcat_bf = reinterpret_cast<__nv_bfloat16*>(take(H * H));
    w.dc_bf = reinterpret_cast<__nv_bfloat16*>(take((B * H + 1) / 2));
    w.wg_part = take(2 * size_t(kMaxSms) * H * 128);      // one slab set per modality chain (they run side by side)
  }
  w.bytes = off;
  return w;
}
static int fusion_check(const MmrcaFusionDesc* d) {
  if (!d) return fail(MMRCA_ERR_INVALID, "null descriptor%s%s");
  if (d->batch < 0 || d->d_img <= 0 || d->d_txt <= 0 || d->hidden <= 0 || (d->hidden & 3))
    return fail(MMRCA_ERR_INVALID, "fusion head: batch >= 0, positive widths, hidden a multiple of 4%s%s");
  if (d->n_classes < 1 || d->n_classes > 8) return fail(MMRCA_ERR_INVALID, "n_classes must be in [1, 8]%s%s");
  if (!(d->drop_p >= 0.f && d->drop_p <= 1.f)) return fail(MMRCA_ERR_INVALID, "drop_p must be in [0, 1]%s%s");
  if ((d->flags & MMRCA_FUSION_BF16) && ((d->hidden & 15) || d->hidden > 256 || (d->d_img & 15) || (d->d_txt & 15)))
    return fail(MMRCA_ERR_INVALID, "fusion head, bf16: hidden a multiple of 16, <= 256; feature widths multiples of 16%s%s");
  return MMRCA_OK;
}
static int launch_sgemm(const fus::GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return MMRCA_OK;
  const int tiles = ((g.N + fus::kTN - 1) / fus::kTN) * ((g.M + fus::kTM - 1) / fus::kTM);
  // accumulating GEMMs with few output tiles and a long contraction (weight gradients: K = batch): split K over ~4 waves
  int splits = 1;
  if (g.accumulate && tiles < 600) splits = std::max(1, std::min((600 + tiles - 1) / tiles, g.K / 256));
  {
    LaunchScope ls("fusion_sgemm", st);
    fus::sgemm_kernel<<<dim3((g.N + fus::kTN - 1) / fus::kTN, (g.M + fus::kTM - 1) / fus::kTM, splits), 256, 0, st>>>(g);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}
static fus::GemmArgs gemm_args(const float* a, long long sa_m, long long sa_k, const float* b, long long sb_k, long long sb_n,
                               float* c, long long ldc, int M, int N, int K, const float* bias, bool acc) {
  fus::GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.a = a; g.sa_m = sa_m; g.sa_k = sa_k; g.b = b; g.sb_k = sb_k; g.sb_n = sb_n; g.c = c; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.bias = bias; g.accumulate = acc ? 1 : 0;
  return g;
}
static int launch_colsum(const float* x, int ld, int rows, int n, float* out, int sms, cudaStream_t st) {
  if (!out || rows <= 0) return MMRCA_OK;
  {
    LaunchScope ls("fusion_colsum", st);
    fus::colsum_kernel<<<dim3((n + 31) / 32, max(1, min((rows + 63) / 64, 2 * sms))), 256, 0, st>>>(x, ld, rows, n, out);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}
static DropSpec fusion_drop(const MmrcaFusionDesc& d) {
  MmrcaHeadDesc h;
  memset(&h, 0, sizeof(h));
  h.drop_p = d.drop_p; h.drop_seed = d.drop_seed;
  DropSpec s = make_drop(h);
  s.D = d.hidden;
  if (d.drop_p <= 0.f) s.thresh = 0;
  return s;
}
static CatArgs fusion_cat(const MmrcaFusionDesc& d, const MmrcaFusionParams& p, const FusionWorkspace& w,
                          const uint8_t* mask, float scale) {
  CatArgs c;
  memset(&c, 0, sizeof(c));
  c.seg[0].src = w.c; c.seg[0].width = d.hidden; c.nseg = 1; c.D = d.hidden; c.batch = d.batch;
  c.mask = mask; c.scale = scale; c.wf = p.w_fc; c.bf = p.b_fc;
  return c;
}
static int fusion_forward_impl(const MmrcaFusionDesc& d, const MmrcaFusionParams& p, const float* img, const float* txt,
                               const uint8_t*& mask, float& scale, float* logits, const FusionWorkspace& w, int sms,
                               cudaStream_t st) {
  const int B = d.batch, H = d.hidden;
  const bool nrm = (d.flags & MMRCA_FUSION_NORMALIZED) != 0;
  int rc;
  if (B == 0) return MMRCA_OK;
  // image_to_hidden_size / text_to_hidden_size (multimodal_model.py:521-522, :566-567)
  if (d.flags & MMRCA_FUSION_BF16) {
    // the two projections on the tensor cores: bf16 copies of the features and of the weights as they lie, TMA-fed tcgen05
    // GEMM with the bias in the epilogue (mmrca_token_bwd.cuh, tok_dgrad_kernel<true>); everything after them is fp32
    const tok::Cast3Args c1 = {{img, txt, p.w_img}, {w.x_bf[0], w.x_bf[1], w.w_bf[0]},
                               {(long long)B * d.d_img, (long long)B * d.d_txt, (long long)H * d.d_img}};
    if ((rc = launch_cast3(c1, sms, st))) return rc;
    const tok::Cast3Args c2 = {{p.w_txt, nullptr, nullptr}, {w.w_bf[1], nullptr, nullptr}, {(long long)H * d.d_txt, 0, 0}};
    if ((rc = launch_cast3(c2, sms, st))) return rc;
    SideStreams* ss;
    if ((rc = side_streams(&ss, st)) || (rc = side_fork(ss, st, 1))) return rc;
    if ((rc = launch_tok_dgrad(w.h_img, B, H, d.d_img, w.x_bf[0], d.d_img, w.w_bf[0], st, true, p.b_img))) return rc;
    if ((rc = launch_tok_dgrad(w.h_txt, B, H, d.d_txt, w.x_bf[1], d.d_txt, w.w_bf[1], ss->s[0], true, p.b_txt))) return rc;
    if ((rc = side_join(ss, st, 1))) return rc;
  } else {
    if ((rc = launch_sgemm(gemm_args(img, d.d_img, 1, p.w_img, 1, d.d_img, w.h_img, H, B, H, d.d_img, p.b_img, false), st))) return rc;
    if ((rc = launch_sgemm(gemm_args(txt, d.d_txt, 1, p.w_txt, 1, d.d_txt, w.h_txt, H, B, H, d.d_txt, p.b_txt, false), st))) return rc;
  }
  if (nrm) {      // :569-570, no epsilon
    const int grid = min((B + kWarps - 1) / kWarps, 8 * sms);
    { LaunchScope ls("l2norm", st); l2norm_kernel<<<grid, kThreads, 0, st>>>(w.h_img, w.n_img, B, H); }
    { LaunchScope ls("l2norm", st); l2norm_kernel<<<grid, kThreads, 0, st>>>(w.h_txt, w.n_txt, B, H); }
    MMRCA_CUDA(cudaGetLastError());
  }
  if (d.flags & MMRCA_FUSION_BF16) {
    // concat + concat_layer on the tensor cores: the (normalised) hidden vectors side by side as one bf16 [B, 2H] operand
    const int grid = int(std::min<long long>(((long long)B * H / 8 + 255) / 256, 8LL * sms));
    {
      LaunchScope ls("cast_rows", st);
      tok::cast_rows_scaled_kernel<<<grid, 256, 0, st>>>(w.h_img, nrm ? w.n_img : nullptr, B, H, w.cat_bf, 2 * H);
      tok::cast_rows_scaled_kernel<<<grid, 256, 0, st>>>(w.h_txt, nrm ? w.n_txt : nullptr, B, H, w.cat_bf + H, 2 * H);
    }
    MMRCA_CUDA(cudaGetLastError());
    const tok::Cast3Args cw = {{p.w_cat, nullptr, nullptr}, {w.wcat_bf, nullptr, nullptr}, {(long long)H * 2 * H, 0, 0}};
    if ((rc = launch_cast3(cw, sms, st))) return rc;
    if ((rc = launch_tok_dgrad(w.c, B, H, 2 * H, w.cat_bf, 2 * H, w.wcat_bf, st, true, p.b_cat))) return rc;
  } else {
    // concat + concat_layer (:524-527, :572-575): two accumulating GEMMs over the halves of W_cat
    fus::GemmArgs g = gemm_args(w.h_img, H, 1, p.w_cat, 1, 2 * H, w.c, H, B, H, H, p.b_cat, false);
    g.inv_m = nrm ? w.n_img : nullptr;
    if ((rc = launch_sgemm(g, st))) return rc;
    g = gemm_args(w.h_txt, H, 1, p.w_cat + H, 1, 2 * H, w.c, H, B, H, H, nullptr, true);
    g.inv_m = nrm ? w.n_txt : nullptr;
    if ((rc = launch_sgemm(g, st))) return rc;
  }
  // self.drop + fc_layer (:528-529, :576-577)
  if (!mask && d.drop_p > 0.f) {
    const DropSpec ds = fusion_drop(d);
    {
      LaunchScope ls("dropout_mask", st);
      dropout_mask_kernel<<<min(4 * sms, max(1, int((size_t(B) * H / 4 + 255) / 256))), 256, 0, st>>>(ds, B, w.mask);
    }
    MMRCA_CUDA(cudaGetLastError());
    mask = w.mask; scale = ds.scale;
  }
  CatArgs c = fusion_cat(d, p, w, mask, scale);
  c.logits = logits;
  return classifier_dispatch(false, d.n_classes, c, sms, st);
}
static int fusion_backward_impl(const MmrcaFusionDesc& d, const MmrcaFusionParams& p, const float* img, const float* txt,
                                const uint8_t* mask, float scale, const float* dlogits, const MmrcaFusionGrads& g,
                                float* d_img, float* d_txt, const FusionWorkspace& w, int sms, cudaStream_t st) {
  const int B = d.batch, H = d.hidden;
  const bool nrm = (d.flags & MMRCA_FUSION_NORMALIZED) != 0;
  int rc;
  if (B == 0) return MMRCA_OK;
  if (!mask && d.drop_p > 0.f) { mask = w.mask; scale = fusion_drop(d).scale; }      // materialised by the forward
  // fc_layer + dropout backward: dW_fc, db_fc, d_c
  CatArgs c = fusion_cat(d, p, w, mask, scale);
  c.dlogits = dlogits; c.g_wf = g.w_fc; c.g_bf = g.b_fc; c.seg[0].dst = w.d_c;
  if ((rc = classifier_dispatch(true, d.n_classes, c, sms, st))) return rc;
  // concat_layer: db, dW (two halves), d(hidden halves)
  if ((rc = launch_colsum(w.d_c, H, B, H, g.b_cat, sms, st))) return rc;
  const float* hs[2] = {w.h_img, w.h_txt};
  const float* ns[2] = {w.n_img, w.n_txt};
  float* dhs[2] = {w.d_h_img, w.d_h_txt};
  const float* xs[2] = {img, txt};
  const int dins[2] = {d.d_img, d.d_txt};
  float* gws[2] = {g.w_img, g.w_txt};
  float* gbs[2] = {g.b_img, g.b_txt};
  const float* ws[2] = {p.w_img, p.w_txt};
  float* dxs[2] = {d_img, d_txt};
  if (d.flags & MMRCA_FUSION_BF16) {
    // Tensor-core path.  After d_c the two modalities are independent chains (concat_layer half -> normalisation backward ->
    // projection): the text chain runs on a side stream next to the image chain (every launch here is short and
    // latency-bound), with its own split-K slabs.
    const tok::Cast3Args cd = {{w.d_c, nullptr, nullptr}, {w.dc_bf, nullptr, nullptr}, {(long long)B * H, 0, 0}};
    if ((rc = launch_cast3(cd, sms, st))) return rc;
    SideStreams* ss;
    if ((rc = side_streams(&ss, st)) || (rc = side_fork(ss, st, 1))) return rc;
    for (int m = 0; m < 2; ++m) {
      cudaStream_t sm = m ? ss->s[0] : st;
      float* part = w.wg_part + size_t(m) * kMaxSms * H * 128;
      // dW_cat[:, half m] += d_c^T (normalised hidden half m); d(hidden half m) = d_c W_cat[:, half m]: bf16 TMA GEMMs over the
      // forward's cat_bf / wcat_bf (operands addressed in place through their row pitch 2H)
      tok::WgradArgs a;
      memset(&a, 0, sizeof(a));
      a.out[0] = {g.w_cat + m * H, 0, H, 2 * H}; a.nout = 1; a.NG = H; a.K = H; a.rows = B;
      if ((rc = launch_tok_wgrad(a, w.cat_bf + m * H, w.dc_bf, H, part, sms, sm, 2 * H))) return rc;
      if ((rc = launch_tok_dgrad(dhs[m], B, H, H, w.dc_bf, H, w.wcat_bf + m * H, sm, false, nullptr, 2 * H))) return rc;
      if (nrm) {      // through h / ||h||
        const int grid = min((B + kWarps - 1) / kWarps, 8 * sms);
        { LaunchScope ls("l2norm_bwd", sm); l2norm_bwd_kernel<<<grid, kThreads, 0, sm>>>(hs[m], ns[m], dhs[m], B, H); }
        MMRCA_CUDA(cudaGetLastError());
      }
      // the projection: db, dW += d(hidden)^T X over the bf16 copies (the forward's x_bf), optionally d(features) in fp32
      const tok::Cast3Args ch = {{dhs[m], nullptr, nullptr}, {w.dh_bf[m], nullptr, nullptr}, {(long long)B * H, 0, 0}};
      if ((rc = launch_cast3(ch, sms, sm))) return rc;
      if ((rc = launch_colsum(dhs[m], H, B, H, gbs[m], sms, sm))) return rc;
      memset(&a, 0, sizeof(a));
      a.out[0] = {gws[m], 0, H, 0}; a.nout = 1; a.NG = H; a.K = dins[m]; a.rows = B;
      if ((rc = launch_tok_wgrad(a, w.x_bf[m], w.dh_bf[m], H, part, sms, sm))) return rc;
      if (dxs[m] && (rc = launch_sgemm(gemm_args(dhs[m], H, 1, ws[m], dins[m], 1, dxs[m], dins[m], B, dins[m], H, nullptr, false), sm)))
        return rc;
    }
    return side_join(ss, st, 1);
  }
  for (int m = 0; m < 2; ++m) {
    fus::GemmArgs a = gemm_args(w.d_c, 1, H, hs[m], H, 1, g.w_cat + m * H, 2 * H, H, H, B, nullptr, true);
    a.inv_k = nrm ? ns[m] : nullptr;
    if ((rc = launch_sgemm(a, st))) return rc;
    if ((rc = launch_sgemm(gemm_args(w.d_c, H, 1, p.w_cat + m * H, 2 * H, 1, dhs[m], H, B, H, H, nullptr, false), st))) return rc;
    if (nrm) {      // through h / ||h||
      const int grid = min((B + kWarps - 1) / kWarps, 8 * sms);
      { LaunchScope ls("l2norm_bwd", st); l2norm_bwd_kernel<<<grid, kThreads, 0, st>>>(hs[m], ns[m], dhs[m], B, H); }
      MMRCA_CUDA(cudaGetLastError());
    }
  }
  // the two projections: db, dW, optionally d(features)
  for (int m = 0; m < 2; ++m) {
    if ((rc = launch_colsum(dhs[m], H, B, H, gbs[m], sms, st))) return rc;
    if ((rc = launch_sgemm(gemm_args(dhs[m], 1, H, xs[m], dins[m], 1, gws[m], dins[m], H, dins[m], B, nullptr, true), st))) return rc;
    if (dxs[m] && (rc = launch_sgemm(gemm_args(dhs[m], H, 1, ws[m], dins[m], 1, dxs[m], dins[m], B, dins[m], H, nullptr, false), st)))
      return rc;
  }
  return MMRCA_OK;
}

}  // namespace mmrca

using namespace mmrca;

extern "C" {

int mmrca_query(int what) {
  switch (what) {
    case MMRCA_QUERY_ABI_VERSION: return MMRCA_ABI_VERSION;
    case MMRCA_QUERY_DEVICE_OK: { DeviceInfo di; return device_info(&di) == MMRCA_OK ? 1 : 0; }
    case MMRCA_QUERY_SM_COUNT: { DeviceInfo di; return device_info(&di) == MMRCA_OK ? di.sms : 0; }
    case MMRCA_QUERY_KERNEL_LAUNCHES: return g_launches;
    case MMRCA_QUERY_RESET_LAUNCHES: { int v = g_launches; g_launches = 0; return v; }
    case MMRCA_QUERY_HAS_BF16: return 1;
    default: return -1;
  }
}

const char* mmrca_last_error(void) { return g_err; }

int mmrca_timing_begin(int32_t max_records) {
  if (g_trec) return -fail(MMRCA_ERR_INVALID, "timing already active on this thread%s%s");
  if (max_records <= 0) return -fail(MMRCA_ERR_INVALID, "max_records must be positive%s%s");
  TimingRec* r = new TimingRec[max_records];
  for (int i = 0; i < max_records; ++i) {
    if (cudaEventCreate(&r[i].e0) != cudaSuccess || cudaEventCreate(&r[i].e1) != cudaSuccess) {
      delete[] r;
      return -fail(MMRCA_ERR_CUDA, "cudaEventCreate failed%s%s");
    }
  }
  g_trec = r; g_tcap = max_records; g_tn = 0;
  return 0;
}

int mmrca_timing_end(MmrcaKernelTime* out, int32_t max_out) {
  if (!g_trec) return -fail(MMRCA_ERR_INVALID, "timing is not active on this thread%s%s");
  const int n = g_tn;
  int rc = n;
  for (int i = 0; i < n; ++i) {
    float ms = 0.f;
    if (cudaEventSynchronize(g_trec[i].e1) != cudaSuccess ||
        cudaEventElapsedTime(&ms, g_trec[i].e0, g_trec[i].e1) != cudaSuccess) {
      rc = -fail(MMRCA_ERR_CUDA, "reading a timing event failed%s%s");
      break;
    }
    if (out && i < max_out) { out[i].name = g_trec[i].name; out[i].ms = ms; }
  }
  for (int i = 0; i < g_tcap; ++i) { cudaEventDestroy(g_trec[i].e0); cudaEventDestroy(g_trec[i].e1); }
  delete[] g_trec;
  g_trec = nullptr; g_tcap = g_tn = 0;
  return rc;
}

size_t mmrca_head_workspace_bytes(const MmrcaHeadDesc* desc, int training) {
  if (!desc) return 0;
  return carve(*desc, training != 0, nullptr).bytes;
}

long long mmrca_head_workspace_offset(const MmrcaHeadDesc* desc, int training, int what) {
  if (!desc) return -1;
  Workspace w = carve(*desc, training != 0, nullptr);
  const void* p = nullptr;
  switch (what) {
    case MMRCA_WS_TEXT_SA_IMAGE: p = w.t_img; break;
    case MMRCA_WS_IMAGE_SA_IMAGE: p = w.i_img; break;
    default: return -1;
  }
  return p ? static_cast<long long>(reinterpret_cast<const char*>(p) - static_cast<const char*>(nullptr)) : -1;
}

int mmrca_head_forward(const MmrcaHeadDesc* desc, const MmrcaHeadParams* params, const float* img_feat,
                       const float* txt_feat, const uint8_t* drop_mask, float drop_scale, float* logits,
                       void* workspace, size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = check_desc(desc))) return rc;
  if (!params || !img_feat || !txt_feat || !logits || !workspace)
    return fail(MMRCA_ERR_INVALID, "null pointer argument%s%s");
  DeviceInfo di;
  if ((rc = device_info(&di))) return rc;
  // MMRCA_FLAG_TRAINING: a backward will follow, the forward keeps what it reloads (training-size workspace)
  const bool training = (desc->flags & MMRCA_FLAG_TRAINING) != 0;
  Workspace w = carve(*desc, training, workspace);
  if (workspace_bytes < w.bytes)
    return fail(MMRCA_ERR_WORKSPACE, training ? "workspace too small (MMRCA_FLAG_TRAINING needs the training size)%s%s"
                                              : "workspace too small%s%s");
  return head_forward_impl(*desc, *params, img_feat, txt_feat, drop_mask, drop_scale, logits, w, di.sms,
                           static_cast<cudaStream_t>(stream));
}

int mmrca_head_backward(const MmrcaHeadDesc* desc, const MmrcaHeadParams* params, const float* img_feat,
                        const float* txt_feat, const uint8_t* drop_mask, float drop_scale, const float* dlogits,
                        const MmrcaHeadGrads* grads, float* d_img_feat, float* d_txt_feat, void* workspace,
                        size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = check_desc(desc))) return rc;
  if (!params || !img_feat || !txt_feat || !dlogits || !grads || !workspace)
    return fail(MMRCA_ERR_INVALID, "null pointer argument%s%s");
  DeviceInfo di;
  if ((rc = device_info(&di))) return rc;
  Workspace w = carve(*desc, true, workspace);
  if (workspace_bytes < w.bytes) return fail(MMRCA_ERR_WORKSPACE, "workspace too small (training size needed)%s%s");
  return head_backward_impl(*desc, *params, img_feat, txt_feat, drop_mask, drop_scale, dlogits, *grads, d_img_feat,
                            d_txt_feat, w, di.sms, static_cast<cudaStream_t>(stream));
}

int mmrca_cross_entropy(const float* logits, const int64_t* labels, const MmrcaCeDesc* ce, int32_t batch,
                        int32_t n_classes, float* loss_out, float* dlogits, void* stream) {
  if (!logits || !labels) return fail(MMRCA_ERR_INVALID, "null pointer argument%s%s");
  DeviceInfo di;
  int rc;
  if ((rc = device_info(&di))) return rc;
  return launch_ce(logits, labels, ce, batch, n_classes, loss_out, dlogits, static_cast<cudaStream_t>(stream));
}

int mmrca_head_train_step(const MmrcaHeadDesc* desc, const MmrcaHeadParams* params, const float* img_feat,
                          const float* txt_feat, const uint8_t* drop_mask, float drop_scale, const int64_t* labels,
                          const MmrcaCeDesc* ce, float* logits, float* loss_out, const MmrcaHeadGrads* grads,
                          float* d_img_feat, float* d_txt_feat, void* workspace, size_t workspace_bytes,
                          void* stream) {
  int rc;
  if ((rc = check_desc(desc))) return rc;
  if (!params || !img_feat || !txt_feat || !labels || !logits || !loss_out || !grads || !workspace)
    return fail(MMRCA_ERR_INVALID, "null pointer argument%s%s");
  DeviceInfo di;
  if ((rc = device_info(&di))) return rc;
  Workspace w = carve(*desc, true, workspace);
  if (workspace_bytes < w.bytes) return fail(MMRCA_ERR_WORKSPACE, "workspace too small (training size needed)%s%s");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (desc_is_tc(*desc) && !drop_mask) w.step_loss = loss_out;
  if (desc->flags & MMRCA_FLAG_ZERO_GRADS) {
    // the gradient tensors are one contiguous bucket in MmrcaHeadGrads order (functional.FlatGrads): [sa_img.wq, bf + C)
    float* lo = grads->sa_img.wq;
    float* hi = grads->bf ? grads->bf + desc->n_classes : nullptr;
    const long long n = lo && hi ? (long long)(hi - lo) : -1;
    if (n <= 0 || n > (1 << 24) || (reinterpret_cast<uintptr_t>(lo) & 15))
      return fail(MMRCA_ERR_INVALID, "MMRCA_FLAG_ZERO_GRADS needs the gradients in one contiguous, 16-byte aligned bucket in "
                                     "MmrcaHeadGrads order (sa_img.wq first, bf last)%s%s");
    const long long n4 = (n + 3) & ~3LL;      // the bucket is padded to 4 floats per tensor
    if (w.step_loss) { w.zero_grads = lo; w.zero_grads_n = int(n4); }
    else MMRCA_CUDA(cudaMemsetAsync(lo, 0, size_t(n4) * sizeof(float), st));
  }
  if ((rc = head_forward_impl(*desc, *params, img_feat, txt_feat, drop_mask, drop_scale, logits, w, di.sms, st)))
    return rc;
  if (desc_is_tc(*desc)) {
    if ((rc = launch_ce_feat(*desc, logits, labels, ce, loss_out, w.dlogits, *grads, grads->bf != nullptr, w, st)))
      return rc;
    return head_backward_impl(*desc, *params, img_feat, txt_feat, drop_mask, drop_scale, w.dlogits, *grads, d_img_feat,
                              d_txt_feat, w, di.sms, st, true);
  }
  if ((rc = launch_ce(logits, labels, ce, desc->batch, desc->n_classes, loss_out, w.dlogits, st))) return rc;
  return head_backward_impl(*desc, *params, img_feat, txt_feat, drop_mask, drop_scale, w.dlogits, *grads, d_img_feat,
                            d_txt_feat, w, di.sms, st);
}

size_t mmrca_token_attention_workspace_bytes(const MmrcaTokenDesc* desc) {
  if (token_check(desc)) return 0;
  return token_carve(*desc, nullptr).bytes;
}

int mmrca_token_attention_forward(const MmrcaTokenDesc* desc, const MmrcaAttnParams* p, const void* x_q, const void* x_kv,
                                  float* out, void* workspace, size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = token_check(desc))) return rc;
  if (!p || !x_q || !out || !workspace) return fail(MMRCA_ERR_INVALID, "null pointer argument%s%s");
  const bool self = x_kv == nullptr || x_kv == x_q;
  if (self && desc->d_in_q != desc->d_in_kv) return fail(MMRCA_ERR_INVALID, "self attention: d_in_q must equal d_in_kv%s%s");
  if ((reinterpret_cast<uintptr_t>(x_q) & 15) || (!self && (reinterpret_cast<uintptr_t>(x_kv) & 15)))
    return fail(MMRCA_ERR_INVALID, "token activations must be 16-byte aligned%s%s");
  DeviceInfo di;
  if ((rc = device_info(&di))) return rc;
  const TokenWorkspace w = token_carve(*desc, workspace);
  if (workspace_bytes < w.bytes) return fail(MMRCA_ERR_WORKSPACE, "workspace too small%s%s");
  if (desc->batch == 0) return MMRCA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int dkq = desc->d_kq, dv = desc->d_v, kq = desc->d_in_q, kkv = desc->d_in_kv;
  // bf16 weight images ([k-block][N][128 B swizzled], tok_wprep_kernel) and stacked fp32 biases.  Self attention: one
  // image of the stacked [W_query; W_key; W_value] (N = 2 d_kq + d_v); cross: one of W_query (x_q's width), one of
  // [W_key; W_value] (x_kv's width).
  const size_t kq_pad = size_t(kq + tok::kBK - 1) / tok::kBK * tok::kBK;
  __nv_bfloat16* wq = w.w;
  __nv_bfloat16* wk = wq + size_t(dkq) * kq_pad;      // cross: the second image starts here
  if (!(desc->flags & MMRCA_TOKEN_WEIGHTS_READY)) {
    tok::WprepArgs wa;
    memset(&wa, 0, sizeof(wa));
    uint8_t *bq_ = reinterpret_cast<uint8_t*>(wq), *bk_ = reinterpret_cast<uint8_t*>(wk);
    if (self) {
      const int N = 2 * dkq + dv;
      wa.seg[0] = {p->wq, p->bq, w.bias, bq_, dkq, kq, N, 0};
      wa.seg[1] = {p->wk, p->bk, w.bias + dkq, bq_, dkq, kq, N, dkq};
      wa.seg[2] = {p->wv, p->bv, w.bias + 2 * dkq, bq_, dv, kq, N, 2 * dkq};
    } else {
      wa.seg[0] = {p->wq, p->bq, w.bias, bq_, dkq, kq, dkq, 0};
      wa.seg[1] = {p->wk, p->bk, w.bias + dkq, bk_, dkq, kkv, dkq + dv, 0};
      wa.seg[2] = {p->wv, p->bv, w.bias + 2 * dkq, bk_, dv, kkv, dkq + dv, dkq};
    }
    if ((rc = launch_tok_wprep(wa, di.sms, st))) return rc;
  }
  const float qscale = 1.0f / sqrtf(float(dkq));      // scores / sqrt(d_kq) (:58-60, :89-91) folded into Q
  if (self) {
    const tok::ProjSeg segs[3] = {{w.q_img, dkq, qscale}, {w.k_img, dkq, 1.0f}, {w.v_img, dv, 1.0f}};
    if ((rc = launch_tok_proj(*desc, x_q, kq, wq, w.bias, 2 * dkq + dv, segs, di.sms, st))) return rc;
  } else {
    // the query and the key / value projection read different activations: side by side on two streams
    SideStreams* ss;
    if ((rc = side_streams(&ss, st)) || (rc = side_fork(ss, st, 1))) return rc;
    const tok::ProjSeg sq[3] = {{w.q_img, dkq, qscale}, {nullptr, 0, 1.0f}, {nullptr, 0, 1.0f}};
    if ((rc = launch_tok_proj(*desc, x_q, kq, wq, w.bias, dkq, sq, di.sms, st))) return rc;
    const tok::ProjSeg skv[3] = {{w.k_img, dkq, 1.0f}, {w.v_img, dv, 1.0f}, {nullptr, 0, 1.0f}};
    if ((rc = launch_tok_proj(*desc, x_kv, kkv, wk, w.bias + dkq, dkq + dv, skv, di.sms, ss->s[0]))) return rc;
    if ((rc = side_join(ss, st, 1))) return rc;
  }
  tok::AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.q_img = w.q_img; a.k_img = w.k_img; a.v_img = w.v_img; a.ln_g = p->ln_g; a.ln_b = p->ln_b; a.out = out;
  a.L = desc->seq_len; a.tiles_per_sample = (desc->seq_len + tok::kTile - 1) / tok::kTile; a.reverse = desc->reverse ? 1 : 0;
  a.p_out = w.p_img; a.sum_out = w.sum;      // null unless MMRCA_TOKEN_TRAINING
  a.out_bf16 = (desc->flags & MMRCA_TOKEN_OUT_BF16) ? 1 : 0;
  return dkq == 128 ? launch_tok_attn<128, 96>(*desc, a, st) : launch_tok_attn<64, 48>(*desc, a, st);
}

int mmrca_token_attention_backward(const MmrcaTokenDesc* desc, const MmrcaAttnParams* p, const void* x_q, const void* x_kv,
                                   const float* d_out, const MmrcaAttnGrads* grads, float* d_x_q, float* d_x_kv,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = token_check(desc))) return rc;
  if (!(desc->flags & MMRCA_TOKEN_TRAINING))
    return fail(MMRCA_ERR_INVALID, "token attention backward needs MMRCA_TOKEN_TRAINING on the descriptor of forward and backward%s%s");
  if (!p || !x_q || !d_out || !grads || !workspace) return fail(MMRCA_ERR_INVALID, "null pointer argument%s%s");
  const bool self = x_kv == nullptr || x_kv == x_q;
  if (self && desc->d_in_q != desc->d_in_kv) return fail(MMRCA_ERR_INVALID, "self attention: d_in_q must equal d_in_kv%s%s");
  if ((desc->d_in_q & 15) || (desc->d_in_kv & 15))
    return fail(MMRCA_ERR_INVALID, "token attention backward: d_in must be a multiple of 16%s%s");
  if ((reinterpret_cast<uintptr_t>(x_q) & 15) || (!self && (reinterpret_cast<uintptr_t>(x_kv) & 15)) ||
      (reinterpret_cast<uintptr_t>(d_out) & 15) || (reinterpret_cast<uintptr_t>(d_x_q) & 15) || (reinterpret_cast<uintptr_t>(d_x_kv) & 15))
    return fail(MMRCA_ERR_INVALID, "token activations and gradients must be 16-byte aligned%s%s");
  if ((reinterpret_cast<uintptr_t>(grads->wq) | reinterpret_cast<uintptr_t>(grads->wk) | reinterpret_cast<uintptr_t>(grads->wv)) & 15)
    return fail(MMRCA_ERR_INVALID, "the weight-gradient tensors must be 16-byte aligned%s%s");
  if (!grads->wq || !grads->wk || !grads->wv || !grads->bq || !grads->bk || !grads->bv || !grads->ln_g || !grads->ln_b)
    return fail(MMRCA_ERR_INVALID, "null gradient pointer%s%s");
  DeviceInfo di;
  if ((rc = device_info(&di))) return rc;
  const TokenWorkspace w = token_carve(*desc, workspace);
  if (workspace_bytes < w.bytes) return fail(MMRCA_ERR_WORKSPACE, "workspace too small%s%s");
  if (desc->batch == 0) return MMRCA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int dkq = desc->d_kq, dv = desc->d_v, kq = desc->d_in_q, kkv = desc->d_in_kv, L = desc->seq_len;
  const int tps = (L + tok::kTile - 1) / tok::kTile, rows = desc->batch * L;
  // gradient rows: self attention keeps dQ | dK | dV side by side (one GEMM operand); cross attention keeps dQ apart from
  // dK | dV (they meet different activations)
  __nv_bfloat16* gq = w.g;
  __nv_bfloat16* gkv = self ? w.g + dkq : w.g + size_t(rows) * dkq;
  const int ldq = self ? 2 * dkq + dv : dkq, ldkv = self ? 2 * dkq + dv : dkq + dv;
  {
    tok::AttnBwdArgs a;
    memset(&a, 0, sizeof(a));
    a.q_img = w.q_img; a.k_img = w.k_img; a.v_img = w.v_img; a.p_img = w.p_img; a.sum = w.sum;
    a.ln_g = p->ln_g; a.ln_b = p->ln_b; a.d_out = d_out;
    a.g_q = gq; a.ld_q = ldq;
    a.g_kv = gkv; a.ld_kv = ldkv; a.kv_atomic = tps > 1 ? 1 : 0;
    if (tps > 1)      // both query tiles of a sample add into the dK | dV columns
      MMRCA_CUDA(self ? cudaMemsetAsync(w.g, 0, size_t(rows) * ldq * 2, st) : cudaMemsetAsync(gkv, 0, size_t(rows) * ldkv * 2, st));
    a.g_bq = grads->bq; a.g_bk = grads->bk; a.g_bv = grads->bv; a.qscale = 1.0f / sqrtf(float(dkq));
    a.g_ln_g = grads->ln_g; a.g_ln_b = grads->ln_b;
    a.L = L; a.tiles_per_sample = tps; a.reverse = desc->reverse ? 1 : 0;
    if ((rc = dkq == 128 ? launch_tok_attn_bwd<128, 96>(*desc, a, st) : launch_tok_attn_bwd<64, 48>(*desc, a, st))) return rc;
  }
  // weight and input gradients: up to four independent GEMMs (cross block: dW_query | dW_key, dW_value | d x_q | d x_kv), each a
  // short latency-bound launch: forked onto side streams, joined before returning to the caller's stream
  SideStreams* ss;
  if ((rc = side_streams(&ss, st))) return rc;
  tok::WgradArgs a;
  memset(&a, 0, sizeof(a));
  a.rows = rows;
  if (self) {
    if (d_x_kv) return fail(MMRCA_ERR_INVALID, "self attention: the input gradient goes to d_x_q%s%s");
    const int nside = d_x_q ? 1 : 0;
    if ((rc = side_fork(ss, st, nside))) return rc;
    a.out[0] = {grads->wq, 0, dkq, 0}; a.out[1] = {grads->wk, dkq, dkq, 0}; a.out[2] = {grads->wv, 2 * dkq, dv, 0}; a.nout = 3;
    a.NG = 2 * dkq + dv; a.K = kq;
    if ((rc = launch_tok_wgrad(a, x_q, w.g, ldq, w.wg_part, di.sms, st))) return rc;
    if (d_x_q) {      // the bf16 weights as they lie ([gradient column][d_in]: MN-major for this product)
      const tok::Cast3Args ca = {{p->wq, p->wk, p->wv}, {w.wbf, w.wbf + size_t(dkq) * kq, w.wbf + size_t(2 * dkq) * kq},
                                 {(long long)dkq * kq, (long long)dkq * kq, (long long)dv * kq}};
      if ((rc = launch_cast3(ca, di.sms, ss->s[0]))) return rc;
      if ((rc = launch_tok_dgrad(d_x_q, rows, kq, 2 * dkq + dv, w.g, ldq, w.wbf, ss->s[0]))) return rc;
    }
    if ((rc = side_join(ss, st, nside))) return rc;
  } else {
    __nv_bfloat16* wkv = w.wbf + size_t(dkq) * kq;
    if (d_x_q || d_x_kv) {
      const tok::Cast3Args ca = {{p->wq, p->wk, p->wv}, {w.wbf, wkv, wkv + size_t(dkq) * kkv},
                                 {d_x_q ? (long long)dkq * kq : 0, d_x_kv ? (long long)dkq * kkv : 0, d_x_kv ? (long long)dv * kkv : 0}};
      if ((rc = launch_cast3(ca, di.sms, st))) return rc;
    }
    if ((rc = side_fork(ss, st, 3))) return rc;
    a.out[0] = {grads->wq, 0, dkq, 0}; a.nout = 1; a.NG = dkq; a.K = kq;
    if ((rc = launch_tok_wgrad(a, x_q, gq, ldq, w.wg_part, di.sms, st))) return rc;
    a.out[0] = {grads->wk, 0, dkq, 0}; a.out[1] = {grads->wv, dkq, dv, 0}; a.nout = 2; a.NG = dkq + dv; a.K = kkv;
    if ((rc = launch_tok_wgrad(a, x_kv, gkv, ldkv, w.wg_part + size_t(kMaxSms) * dkq * 128, di.sms, ss->s[0]))) return rc;
    if (d_x_q && (rc = launch_tok_dgrad(d_x_q, rows, kq, dkq, gq, ldq, w.wbf, ss->s[1]))) return rc;
    if (d_x_kv && (rc = launch_tok_dgrad(d_x_kv, rows, kkv, dkq + dv, gkv, ldkv, wkv, ss->s[2]))) return rc;
    if ((rc = side_join(ss, st, 3))) return rc;
  }
  return MMRCA_OK;
}

size_t mmrca_fusion_workspace_bytes(const MmrcaFusionDesc* desc) {
  if (fusion_check(desc)) return 0;
  return fusion_carve(*desc, nullptr).bytes;
}

static int fusion_common(const MmrcaFusionDesc* desc, const void* a, const void* b, const void* c, const void* e, void* ws,
                         size_t ws_bytes, FusionWorkspace* w, DeviceInfo* di) {
  int rc;
  if ((rc = fusion_check(desc))) return rc;
  if (!a || !b || !c || !e || !ws) return fail(MMRCA_ERR_INVALID, "null pointer argument%s%s");
  if ((desc->flags & MMRCA_FUSION_BF16) && ((reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15))
    return fail(MMRCA_ERR_INVALID, "fusion head, bf16: the feature tensors must be 16-byte aligned%s%s");
  if ((rc = device_info(di))) return rc;
  *w = fusion_carve(*desc, ws);
  if (ws_bytes < w->bytes) return fail(MMRCA_ERR_WORKSPACE, "workspace too small%s%s");
  return MMRCA_OK;
}

int mmrca_fusion_forward(const MmrcaFusionDesc* desc, const MmrcaFusionParams* params, const float* img_feat,
                         const float* txt_feat, const uint8_t* drop_mask, float drop_scale, float* logits, void* workspace,
                         size_t workspace_bytes, void* stream) {
  FusionWorkspace w; DeviceInfo di;
  int rc;
  if ((rc = fusion_common(desc, params, img_feat, txt_feat, logits, workspace, workspace_bytes, &w, &di))) return rc;
  return fusion_forward_impl(*desc, *params, img_feat, txt_feat, drop_mask, drop_scale, logits, w, di.sms,
                             static_cast<cudaStream_t>(stream));
}

int mmrca_fusion_backward(const MmrcaFusionDesc* desc, const MmrcaFusionParams* params, const float* img_feat,
                          const float* txt_feat, const uint8_t* drop_mask, float drop_scale, const float* dlogits,
                          const MmrcaFusionGrads* grads, float* d_img_feat, float* d_txt_feat, void* workspace,
                          size_t workspace_bytes, void* stream) {
  FusionWorkspace w; DeviceInfo di;
  int rc;
  if ((rc = fusion_common(desc, params, img_feat, txt_feat, dlogits, workspace, workspace_bytes, &w, &di))) return rc;
  if (!grads) return fail(MMRCA_ERR_INVALID, "null pointer argument%s%s");
  return fusion_backward_impl(*desc, *params, img_feat, txt_feat, drop_mask, drop_scale, dlogits, *grads, d_img_feat,
                              d_txt_feat, w, di.sms, static_cast<cudaStream_t>(stream));
}

int mmrca_fusion_train_step(const MmrcaFusionDesc* desc, const MmrcaFusionParams* params, const float* img_feat,
                            const float* txt_feat, const uint8_t* drop_mask, float drop_scale, const int64_t* labels,
                            const MmrcaCeDesc* ce, float* logits, float* loss_out, const MmrcaFusionGrads* grads,
                            float* d_img_feat, float* d_txt_feat, void* workspace, size_t workspace_bytes, void* stream) {
  FusionWorkspace w; DeviceInfo di;
  int rc;
  if ((rc = fusion_common(desc, params, img_feat, txt_feat, logits, workspace, workspace_bytes, &w, &di))) return rc;
  if (!grads || !labels || !loss_out) return fail(MMRCA_ERR_INVALID, "null pointer argument%s%s");
  if (desc->batch == 0) return MMRCA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((rc = fusion_forward_impl(*desc, *params, img_feat, txt_feat, drop_mask, drop_scale, logits, w, di.sms, st))) return rc;
  if ((rc = launch_ce(logits, labels, ce, desc->batch, desc->n_classes, loss_out, w.dlogits, st))) return rc;
  return fusion_backward_impl(*desc, *params, img_feat, txt_feat, drop_mask, drop_scale, w.dlogits, *grads, d_img_feat,
                              d_txt_feat, w, di.sms, st);
}

size_t mmrca_hier_workspace_bytes(const MmrcaHierDesc* desc) {
  if (hier_check(desc)) return 0;
  return hier_carve(desc->batch, nullptr, desc->flags).bytes;
}

int mmrca_hier_forward(const MmrcaHierDesc* desc, const MmrcaHierParams* params, const float* const* feats,
                       const uint8_t* drop_mask, float drop_scale, float* logits, void* workspace,
                       size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = hier_check(desc))) return rc;
  if (!params || !feats || !logits || !workspace) return fail(MMRCA_ERR_INVALID, "null pointer argument%s%s");
  DeviceInfo di;
  if ((rc = device_info(&di))) return rc;
  const HierWorkspace w = hier_carve(desc->batch, workspace, desc->flags);
  if (workspace_bytes < w.bytes) return fail(MMRCA_ERR_WORKSPACE, "workspace too small%s%s");
  return hier_forward_impl(*desc, *params, feats, drop_mask, drop_scale, logits, w, static_cast<cudaStream_t>(stream));
}

int mmrca_hier_backward(const MmrcaHierDesc* desc, const MmrcaHierParams* params, const float* dlogits,
                        const MmrcaHierGrads* grads, void* workspace, size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = hier_check(desc))) return rc;
  if (!params || !dlogits || !grads || !workspace) return fail(MMRCA_ERR_INVALID, "null pointer argument%s%s");
  DeviceInfo di;
  if ((rc = device_info(&di))) return rc;
  const HierWorkspace w = hier_carve(desc->batch, workspace, desc->flags);
  if (workspace_bytes < w.bytes) return fail(MMRCA_ERR_WORKSPACE, "workspace too small%s%s");
  return hier_backward_impl(*desc, *params, dlogits, *grads, w, static_cast<cudaStream_t>(stream));
}

int mmrca_hier_backward_features(const MmrcaHierDesc* desc, const MmrcaHierParams* params, const float* const* feats,
                                 const uint8_t* drop_mask, float drop_scale, float* const* d_feats, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = hier_check(desc))) return rc;
  if (!params || !feats || !d_feats || !workspace) return fail(MMRCA_ERR_INVALID, "null pointer argument%s%s");
  if (!(desc->flags & MMRCA_HIER_FEATURE_GRADS))
    return fail(MMRCA_ERR_INVALID, "feature gradients need MMRCA_HIER_FEATURE_GRADS in the desc of the forward and the backward%s%s");
  for (int i = 0; i < 6; ++i)
    if (!feats[i] || !d_feats[i] || ((reinterpret_cast<uintptr_t>(feats[i]) | reinterpret_cast<uintptr_t>(d_feats[i])) & 15))
      return fail(MMRCA_ERR_INVALID, "the six feature / gradient pointers must be non-null and 16-byte aligned%s%s");
  DeviceInfo di;
  if ((rc = device_info(&di))) return rc;
  const HierWorkspace w = hier_carve(desc->batch, workspace, desc->flags);
  if (workspace_bytes < w.bytes) return fail(MMRCA_ERR_WORKSPACE, "workspace too small%s%s");
  if (desc->batch == 0) return MMRCA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int tiles = (desc->batch + hier::kTile - 1) / hier::kTile;
  {   // d(dropped concat) = dH W: the dH image of mmrca_hier_backward and the forward's weight blobs
    hier::DxArgs a;
    a.dh = w.dh; a.wb[0] = w.wb_img; a.wb[1] = w.wb_txt; a.dcat = w.dcat;
    if ((rc = set_smem(hier::hier_dx_kernel, hier::DxSmem::BYTES))) return rc;
    LaunchScope ls("hier_dx", st);
    hier::hier_dx_kernel<<<dim3(tiles, hier::kD / hier::kBN), hier::kGemmThreads, hier::DxSmem::BYTES, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  {
    hier::DxFinishArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < 6; ++i) { a.seg[i] = feats[i]; a.out[i] = d_feats[i]; }
    a.dcat = w.dcat; a.mask = drop_mask; a.mask_scale = drop_scale; a.drop = hier_drop(*desc); a.batch = desc->batch;
    LaunchScope ls("hier_dx_finish", st);
    hier::hier_dx_finish_kernel<<<desc->batch, 256, 0, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}

int mmrca_hier_train_step(const MmrcaHierDesc* desc, const MmrcaHierParams* params, const float* const* feats,
                          const uint8_t* drop_mask, float drop_scale, const int64_t* labels, const MmrcaCeDesc* ce,
                          float* logits, float* loss_out, const MmrcaHierGrads* grads, void* workspace,
                          size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = hier_check(desc))) return rc;
  if (!params || !feats || !labels || !logits || !loss_out || !grads || !workspace)
    return fail(MMRCA_ERR_INVALID, "null pointer argument%s%s");
  DeviceInfo di;
  if ((rc = device_info(&di))) return rc;
  const HierWorkspace w = hier_carve(desc->batch, workspace, desc->flags);
  if (workspace_bytes < w.bytes) return fail(MMRCA_ERR_WORKSPACE, "workspace too small%s%s");
  if (desc->batch == 0) return MMRCA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((rc = hier_forward_impl(*desc, *params, feats, drop_mask, drop_scale, logits, w, st))) return rc;
  {   // CrossEntropyLoss + dlogits: the multi-CTA kernel of the bf16 pipeline in its cross-entropy-only mode
    MMRCA_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
    htc::CeFeatArgs a;
    memset(&a, 0, sizeof(a));
    a.logits = logits; a.labels = labels; a.cw = ce ? ce->class_weight : nullptr; a.eps = ce ? ce->label_smoothing : 0.f;
    a.batch = desc->batch; a.dlogits = w.dlogits; a.loss = loss_out;
    const int tiles8 = (desc->batch + 7) / 8;
    LaunchScope ls("ce_feat", st);
    htc::ce_feat_kernel<<<dim3(1, (tiles8 + htc::kCeSliceTiles - 1) / htc::kCeSliceTiles), 256, 0, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  return hier_backward_impl(*desc, *params, w.dlogits, *grads, w, st);
}

int mmrca_peer_allreduce_mean(float* flat, int32_t n, int32_t n_pad, const void* const* staging, void* const* pads,
                              int32_t rank, int32_t world, uint32_t step, void* stream) {
  if (!flat || !staging || !pads) return fail(MMRCA_ERR_INVALID, "null pointer argument%s%s");
  if (n <= 0 || (n & 3) || n_pad < n || (n_pad & 3) || (reinterpret_cast<uintptr_t>(flat) & 15))
    return fail(MMRCA_ERR_INVALID, "bucket must be 16-byte aligned with a multiple of 4 floats, n_pad >= n%s%s");
  if (world < 1 || world > peer::kMaxWorld || rank < 0 || rank >= world || step == 0)
    return fail(MMRCA_ERR_INVALID, "world must be in [1, 8], rank in [0, world), step >= 1%s%s");
  DeviceInfo di;
  int rc;
  if ((rc = device_info(&di))) return rc;
  if (di.sms < peer::kCtas) return fail(MMRCA_ERR_INVALID, "the peer all-reduce needs its CTAs co-resident%s%s");
  peer::Args a;
  memset(&a, 0, sizeof(a));
  a.flat = flat; a.n = n; a.n_pad = n_pad; a.rank = rank; a.world = world; a.step = step;
  for (int i = 0; i < world; ++i) {
    if (!staging[i] || !pads[i]) return fail(MMRCA_ERR_INVALID, "null peer buffer%s%s");
    a.staging[i] = static_cast<const float*>(staging[i]); a.pads[i] = static_cast<uint32_t*>(pads[i]);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    LaunchScope ls("peer_allreduce", st);
    peer::allreduce_mean_kernel<<<peer::kCtas, peer::kThreads, 0, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}
int mmrca_peer_allreduce_pad_bytes(int32_t world) {
  return (2 * world * peer::kCtas + peer::kStatusWords) * int(sizeof(uint32_t));
}
int mmrca_peer_allreduce_status(const void* own_pad, int32_t world, void* stream) {
  if (!own_pad || world < 1 || world > peer::kMaxWorld) return -fail(MMRCA_ERR_INVALID, "bad pad / world%s%s");
  uint32_t word = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cudaMemcpyAsync(&word, static_cast<const uint32_t*>(own_pad) + size_t(2) * world * peer::kCtas, sizeof(word),
                      cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
    return -fail(MMRCA_ERR_CUDA, "reading the peer all-reduce status word failed%s%s");
  return int(word);
}

int mmrca_feature_handoff(const void* hidden, int32_t hidden_bf16, int64_t hidden_batch_stride, int32_t d_txt,
                          const void* fmap, int32_t fmap_bf16, int32_t channels, int32_t hw, int32_t channels_last,
                          int32_t batch, void* txt_out, void* img_out, int32_t out_bf16, void* stream) {
  if (!hidden || !fmap || !txt_out || !img_out) return fail(MMRCA_ERR_INVALID, "null pointer argument%s%s");
  if (batch < 0 || d_txt <= 0 || channels <= 0 || hw <= 0 || hidden_batch_stride < d_txt)
    return fail(MMRCA_ERR_INVALID, "bad hand-off shape%s%s");
  DeviceInfo di;
  int rc;
  if ((rc = device_info(&di))) return rc;
  if (batch == 0) return MMRCA_OK;
  if (batch > 65535) return fail(MMRCA_ERR_INVALID, "hand-off batch must be <= 65535%s%s");
  aux::HandoffArgs a;
  memset(&a, 0, sizeof(a));
  a.hidden = hidden; a.hidden_bstride = hidden_batch_stride; a.hidden_bf16 = hidden_bf16 ? 1 : 0; a.d_txt = d_txt;
  a.fmap = fmap; a.fmap_bf16 = fmap_bf16 ? 1 : 0; a.channels = channels; a.hw = hw; a.channels_last = channels_last ? 1 : 0;
  a.txt_out = txt_out; a.img_out = img_out; a.out_bf16 = out_bf16 ? 1 : 0; a.batch = batch;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    LaunchScope ls("feature_handoff", st);
    aux::feature_handoff_kernel<<<dim3((channels + 31) / 32 + 1, batch), 256, 0, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}

static int check_bucket(const void* p, int64_t n) {
  if (!p || n <= 0 || (n & 3) || (reinterpret_cast<uintptr_t>(p) & 15))
    return fail(MMRCA_ERR_INVALID, "optimizer buckets must be non-null, 16-byte aligned, with a multiple of 4 floats%s%s");
  return MMRCA_OK;
}

int mmrca_sgd_step(float* params, const float* grads, float* momentum_buf, int64_t n, float lr, float momentum,
                   float dampening, float weight_decay, int32_t nesterov, int32_t first_step, void* stream) {
  int rc;
  if ((rc = check_bucket(params, n)) || (rc = check_bucket(grads, n))) return rc;
  if (momentum != 0.f && (rc = check_bucket(momentum_buf, n))) return rc;
  DeviceInfo di;
  if ((rc = device_info(&di))) return rc;
  aux::SgdArgs a;
  a.p = params; a.g = grads; a.buf = momentum != 0.f ? momentum_buf : nullptr; a.n4 = n / 4;
  a.lr = lr; a.momentum = momentum; a.dampening = dampening; a.weight_decay = weight_decay;
  a.nesterov = nesterov ? 1 : 0; a.first_step = first_step ? 1 : 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    LaunchScope ls("sgd_step", st);
    aux::sgd_step_kernel<<<int(std::min<long long>((a.n4 + 255) / 256, 2LL * di.sms)), 256, 0, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}

int mmrca_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                     float beta1, float beta2, float eps, float weight_decay, int32_t step, void* stream) {
  int rc;
  if ((rc = check_bucket(params, n)) || (rc = check_bucket(grads, n)) || (rc = check_bucket(exp_avg, n)) ||
      (rc = check_bucket(exp_avg_sq, n))) return rc;
  if (step < 1) return fail(MMRCA_ERR_INVALID, "AdamW step counts from 1%s%s");
  DeviceInfo di;
  if ((rc = device_info(&di))) return rc;
  aux::AdamwArgs a;
  a.p = params; a.g = grads; a.m = exp_avg; a.v = exp_avg_sq; a.n4 = n / 4;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bias1 = float(1.0 - pow(double(beta1), double(step)));
  a.bias2_sqrt = float(sqrt(1.0 - pow(double(beta2), double(step))));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    LaunchScope ls("adamw_step", st);
    aux::adamw_step_kernel<<<int(std::min<long long>((a.n4 + 255) / 256, 2LL * di.sms)), 256, 0, st>>>(a);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}

size_t mmrca_attention_forward_scratch_bytes(int32_t, int32_t, int32_t, int32_t) { return 0; }

int mmrca_attention_forward(const MmrcaAttnParams* p, const float* x_q, const float* x_kv, int32_t batch,
                            int32_t d_in, int32_t d_kq, int32_t d_v, int32_t reverse, int32_t normalise,
                            float* norms_out, float* out, void* scratch, size_t scratch_bytes, int32_t compute,
                            void* stream) {
  (void)scratch; (void)scratch_bytes;
  if (!p || !x_q || !x_kv || !out) return fail(MMRCA_ERR_INVALID, "null pointer argument%s%s");
  if (compute != MMRCA_COMPUTE_FP32)
    return fail(MMRCA_ERR_INVALID, "stand-alone attention blocks run in fp32 only: the bf16 tensor-core pipeline exists "
                                   "as the fused head (mmrca_head_*)%s%s");
  if (normalise && (x_q != x_kv || !norms_out))
    return fail(MMRCA_ERR_INVALID, "normalise needs x_kv == x_q and a norms_out buffer%s%s");
  DeviceInfo di;
  int rc;
  if ((rc = device_info(&di))) return rc;
  AttnArgs a = make_attn_args(*p, x_q, x_kv, batch, reverse ? 1 : 0);
  a.normalise = normalise ? 1 : 0; a.norms = norms_out; a.out = out;
  return attn_dispatch(false, x_q == x_kv, d_in, d_kq, d_v, a, di.sms, static_cast<cudaStream_t>(stream));
}

size_t mmrca_attention_backward_scratch_bytes(int32_t batch, int32_t d_kq, int32_t d_v) {
  return align_up(size_t(batch > 0 ? batch : 0) * kL * size_t(2 * d_kq + d_v) * sizeof(float));
}

int mmrca_attention_backward(const MmrcaAttnParams* p, const float* x_q, const float* x_kv, const float* d_out,
                             int32_t batch, int32_t d_in, int32_t d_kq, int32_t d_v, int32_t reverse,
                             const MmrcaAttnGrads* grads, float* d_x_q, float* d_x_kv, void* scratch,
                             size_t scratch_bytes, int32_t compute, void* stream) {
  if (!p || !x_q || !x_kv || !d_out || !grads || !scratch) return fail(MMRCA_ERR_INVALID, "null pointer argument%s%s");
  if (compute != MMRCA_COMPUTE_FP32)
    return fail(MMRCA_ERR_INVALID, "stand-alone attention blocks run in fp32 only%s%s");
  if (scratch_bytes < mmrca_attention_backward_scratch_bytes(batch, d_kq, d_v))
    return fail(MMRCA_ERR_WORKSPACE, "scratch too small%s%s");
  const bool self = x_q == x_kv;
  if (self && d_x_kv) return fail(MMRCA_ERR_INVALID, "self attention: pass d_x_kv = NULL, d_x_q receives the sum%s%s");
  DeviceInfo di;
  int rc;
  if ((rc = device_info(&di))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AttnArgs a = make_attn_args(*p, x_q, x_kv, batch, reverse ? 1 : 0);
  a.dout = d_out; a.dy = static_cast<float*>(scratch); a.dxq = d_x_q; a.dxkv = d_x_kv;
  a.g_ln_g = grads->ln_g; a.g_ln_b = grads->ln_b;
  if ((rc = attn_dispatch(true, self, d_in, d_kq, d_v, a, di.sms, st))) return rc;
  return attn_wgrads(self, d_in, d_kq, d_v, a.dy, x_q, x_kv, nullptr, batch, *grads, di.sms, st);
}

int mmrca_dropout_mask(uint64_t seed, float p, int32_t batch, int32_t width, uint8_t* mask_out, void* stream) {
  if (!mask_out || batch < 0 || width <= 0 || (width & 3) || !(p >= 0.f && p <= 1.f))
    return fail(MMRCA_ERR_INVALID, "dropout mask needs a width that is a multiple of 4, p in [0, 1] and an output buffer%s%s");
  DeviceInfo di;
  int rc;
  if ((rc = device_info(&di))) return rc;
  if (batch == 0) return MMRCA_OK;
  MmrcaHeadDesc d;
  memset(&d, 0, sizeof(d));
  d.drop_p = p; d.drop_seed = seed;
  DropSpec ds = make_drop(d);
  ds.D = width;
  if (p == 0.f) ds.thresh = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    LaunchScope ls("dropout_mask", st);
    dropout_mask_kernel<<<min(4 * di.sms, max(1, int((size_t(batch) * width / 4 + 255) / 256))), 256, 0, st>>>(ds, batch, mask_out);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}

int mmrca_dev_set_debug(void* device_buffer_1024_int64, int32_t kernel) {
  g_dbg = static_cast<long long*>(device_buffer_1024_int64); g_dbg_kernel = kernel;
  return MMRCA_OK;
}

int mmrca_dev_umma_selftest(int32_t mode, const float* a, const float* b, float* out, int32_t n, int32_t k,
                            void* stream) {
  if (!a || !b || !out) return fail(MMRCA_ERR_INVALID, "null pointer argument%s%s");
  if (n < 16 || n > 256 || n % 16 || k < 16 || k % 16 || mode < 0 || mode > 16)
    return fail(MMRCA_ERR_INVALID, "selftest needs 16 <= n <= 256, n % 16 == 0, k % 16 == 0, mode in [0,16]%s%s");
  DeviceInfo di;
  int rc;
  if ((rc = device_info(&di))) return rc;
  if (mode == 16) {
    const size_t smem16 = size_t(2 + (n + 63) / 64) * k * 128 + 1024;
    if (smem16 > 200 * 1024) return fail(MMRCA_ERR_INVALID, "selftest operands do not fit shared memory%s%s");
    if ((rc = set_smem(tc::umma_mn128_selftest_kernel, smem16))) return rc;
    cudaStream_t st16 = static_cast<cudaStream_t>(stream);
    {
      LaunchScope ls("umma_mn128_selftest", st16);
      tc::umma_mn128_selftest_kernel<<<1, 128, smem16, st16>>>(a, b, out, n, k);
    }
    MMRCA_CUDA(cudaGetLastError());
    return MMRCA_OK;
  }
  const size_t smem = tc::umma_selftest_smem_bytes(n, k);
  if (smem > 200 * 1024) return fail(MMRCA_ERR_INVALID, "selftest operands do not fit shared memory%s%s");
  if ((rc = set_smem(tc::umma_selftest_kernel, smem))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    LaunchScope ls("umma_selftest", st);
    tc::umma_selftest_kernel<<<1, 128, smem, st>>>(mode, a, b, out, n, k);
  }
  MMRCA_CUDA(cudaGetLastError());
  return MMRCA_OK;
}

}  // extern "C"

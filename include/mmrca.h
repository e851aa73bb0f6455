/*
 * mmrca.h — C ABI of the B200-native MM-RCA late-fusion head.
 *
 * This is the drop-in boundary for the hot path of espiriki/Garbage_Classification_RCA:
 * everything MM_RCA.forward does after the two backbones returned their pooled
 * features (CVPR_code/multimodal_model.py:661-728), the CrossEntropyLoss applied to
 * its logits (main_both.py:87-93,110) and loss.backward() through both
 * (main_both.py:112).  The reference is pure Python/PyTorch, so "the FFI a maintainer
 * would bind" is a ctypes stub called from MM_RCA.forward / a torch.autograd.Function;
 * INTEGRATION.md shows it.
 *
 * Conventions
 *   - plain C, no torch types: raw DEVICE pointers + sizes + a cudaStream_t passed as void*;
 *   - every call only enqueues work on `stream`: no allocation, no synchronisation, no host
 *     callbacks; the caller owns every buffer (PyTorch tensors in the Python host layer);
 *   - re-entrant per device/stream (nn.DataParallel calls forward from N threads);
 *   - return 0 on success, non-zero MMRCA_ERR_* otherwise; mmrca_last_error() gives the
 *     thread-local message.  There is NO CPU fallback: without an sm_100 device the calls fail.
 *   - all matrices are row-major, weights in torch.nn.Linear layout [out_features, in_features].
 *
 * Fixed by the reference ctor literals (multimodal_model.py:249-258): 16 chunks per sample,
 * self-attention 128 (q/k) / 96 (v), cross-attention 64 / 48.
 */
#ifndef MMRCA_H_
#define MMRCA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMRCA_ABI_VERSION 4

#define MMRCA_NUM_PATCHES 16 /* multimodal_model.py:249 */
#define MMRCA_SA_DKQ 128     /* :251 */
#define MMRCA_SA_DV 96       /* :252 */
#define MMRCA_CA_DKQ 64      /* :254 */
#define MMRCA_CA_DV 48       /* :255 */

/* error codes */
#define MMRCA_OK 0
#define MMRCA_ERR_INVALID 1     /* bad argument / unsupported shape */
#define MMRCA_ERR_CUDA 2        /* a CUDA runtime call failed */
#define MMRCA_ERR_NO_DEVICE 3   /* no sm_100 device: there is no fallback path */
#define MMRCA_ERR_WORKSPACE 4   /* workspace too small */

/* MmrcaHeadDesc.flags — the reference's constructor switches (multimodal_model.py:166-168) */
#define MMRCA_FLAG_REVERSE 1u              /* --reverse: (1-A)/(L-1) weights, :95-99 */
#define MMRCA_FLAG_FEATURES_ONLY 2u        /* --features_only: concat = [img, txt], :694-699 */
#define MMRCA_FLAG_CROSS_ATTENTION_ONLY 4u /* --cross_attention_only: concat = [T_I, I_T], :701-706 */
#define MMRCA_FLAG_FEATURE_GRADS 256u      /* the backward will be asked for d_img_feat / d_txt_feat (fine-tune phase,
                                              main_both.py:687-694): set it for the forward too */
#define MMRCA_FLAG_FEATURES_BF16 1024u     /* img_feat / txt_feat point to bf16 (not fp32) arrays of the same shape: cast the
                                              pointers.  What a backbone under bf16 autocast hands over; halves the bytes a
                                              host -> device hand-off ships.  Norms, the fp32 classifier terms and everything
                                              downstream are computed from the bf16 values in fp32.  bf16 pipeline only. */
#define MMRCA_FLAG_ZERO_GRADS 2048u        /* mmrca_head_train_step only: clear the gradients first (optimizer.zero_grad() folded
                                              into the step's first kernel).  The gradient tensors must form ONE contiguous,
                                              16-byte aligned bucket in MmrcaHeadGrads order, sa_img.wq first, bf last, each
                                              tensor padded to a multiple of 4 floats (what the Python layer's FlatGrads is). */
#define MMRCA_FLAG_TRAINING 512u           /* mmrca_head_forward only: a mmrca_head_backward on the same workspace follows
                                              (autograd), so the forward keeps what the backward reloads and needs the
                                              training-size workspace.  Without it the forward is inference: nothing is
                                              kept, whatever the size of the workspace.  (ABI v4; v3 inferred this from
                                              the workspace size.) */

/* MmrcaHeadDesc.compute */
#define MMRCA_COMPUTE_FP32 0 /* fp32 SIMT kernels: the 1e-4-relative contract */
#define MMRCA_COMPUTE_BF16 1 /* bf16 tcgen05 tensor-core pipeline, fp32 accumulate: the 2e-2-absolute contract.
                                Covers the reference's literal shapes (1280 / 768 features, 4 classes) with seeded
                                dropout (MmrcaHeadDesc.drop_p / drop_seed), frozen features or feature gradients
                                (MMRCA_FLAG_FEATURE_GRADS); other widths (and features_only with feature gradients) run
                                the fp32 kernels (the workspace is laid out accordingly: it is a function of the desc
                                alone); a caller-supplied drop_mask is an error in this mode (pass MMRCA_COMPUTE_FP32);
                                features_only is two streaming kernels (normalise + fp32 classifier, then
                                cross-entropy + dWf). */
#define MMRCA_COMPUTE_BF16_FUSED 2 /* alias of MMRCA_COMPUTE_BF16 (kept for ABI v1 callers) */

/* mmrca_query() selectors */
#define MMRCA_QUERY_ABI_VERSION 0
#define MMRCA_QUERY_DEVICE_OK 1      /* 1 if the current device is compute capability 10.x */
#define MMRCA_QUERY_SM_COUNT 2
#define MMRCA_QUERY_KERNEL_LAUNCHES 3 /* kernels launched by this library on this thread since last reset */
#define MMRCA_QUERY_RESET_LAUNCHES 4
#define MMRCA_QUERY_HAS_BF16 5       /* 1 if the bf16 tensor-core path is compiled in */

/* One SelfAttention / ReverseCrossAttention block (multimodal_model.py:39-108):
 * W_query/W_key/W_value Linear weights+biases and the LayerNorm affine. */
typedef struct MmrcaAttnParams {
  const float* wq; /* [d_kq, d_in] */
  const float* bq; /* [d_kq] */
  const float* wk; /* [d_kq, d_in] */
  const float* bk; /* [d_kq] */
  const float* wv; /* [d_v, d_in] */
  const float* bv; /* [d_v] */
  const float* ln_g; /* [d_v] */
  const float* ln_b; /* [d_v] */
} MmrcaAttnParams;

typedef struct MmrcaAttnGrads {
  float* wq; float* bq; float* wk; float* bk; float* wv; float* bv; float* ln_g; float* ln_b;
} MmrcaAttnGrads;

/* Parameters read by MM_RCA.forward (34 tensors, 94 820 scalars in the full variant). */
typedef struct MmrcaHeadParams {
  MmrcaAttnParams sa_img; /* self_attention_image, :266 */
  MmrcaAttnParams sa_txt; /* self_attention_text,  :268 */
  MmrcaAttnParams ca1;    /* cross_attention_1 (Q: text SA, K/V: image SA), :271, :683 */
  MmrcaAttnParams ca2;    /* cross_attention_2 (Q: image SA, K/V: text SA), :275, :685 */
  const float* wf;        /* classifier selected by flags, [n_classes, D]: final_with_everything (:290) /
                             final_features_only_linear (:282) / cross_attention_only_linear (:286) */
  const float* bf;        /* [n_classes] */
} MmrcaHeadParams;

/* Gradient destinations, same layout.  The backward ACCUMULATES (+=) into them, like
 * loss.backward() accumulates into .grad; the caller zeroes them (optimizer.zero_grad()). */
typedef struct MmrcaHeadGrads {
  MmrcaAttnGrads sa_img, sa_txt, ca1, ca2;
  float* wf;
  float* bf;
} MmrcaHeadGrads;

typedef struct MmrcaHeadDesc {
  int32_t batch;     /* B */
  int32_t d_img;     /* pooled image feature width (1280, :258); multiple of 16*4 */
  int32_t d_txt;     /* pooled text feature width (768, :257) */
  int32_t n_classes; /* 4 */
  uint32_t flags;    /* MMRCA_FLAG_* */
  int32_t compute;   /* MMRCA_COMPUTE_* */
  /* self.drop = nn.Dropout(drop_ratio) on the concat (multimodal_model.py:190, :719), train mode only.
   * drop_p > 0 and drop_mask == NULL: the library draws the keep mask itself, a pure function of
   * (drop_seed, drop_p, sample, concat column) that forward and backward regenerate on chip; kept values are
   * scaled by 1/(1-drop_p).  mmrca_dropout_mask() materialises the same mask.  The backward must be given the
   * same drop_p / drop_seed as its forward.  drop_p == 0 (eval): no dropout. */
  float drop_p;
  uint64_t drop_seed;
} MmrcaHeadDesc;

/* Cross-entropy configuration: torch.nn.CrossEntropyLoss(weight=, label_smoothing=), mean
 * reduction (main_both.py:87-93). class_weight may be NULL. */
typedef struct MmrcaCeDesc {
  const float* class_weight; /* device [n_classes] or NULL */
  float label_smoothing;
} MmrcaCeDesc;

int mmrca_query(int what);
const char* mmrca_last_error(void);

/* Bytes of device scratch the calls below need for this desc (monotone in batch).
 * `training` != 0 adds the backward buffers.  The forward leaves the per-block activations
 * (SA / CA outputs, feature norms) in the workspace; a backward for the same inputs must be
 * given the same, unmodified workspace. */
size_t mmrca_head_workspace_bytes(const MmrcaHeadDesc* desc, int training);

/* Test hook: byte offset of an intermediate inside the workspace (-1 if this desc has none).
 * The SA "images" are the bf16 [tiles][12 column groups][128 rows][8] operand layout of the fused pipeline. */
#define MMRCA_WS_TEXT_SA_IMAGE 0
#define MMRCA_WS_IMAGE_SA_IMAGE 1
long long mmrca_head_workspace_offset(const MmrcaHeadDesc* desc, int training, int what);

/* MM_RCA.forward from the pooled features on (multimodal_model.py:661-728).
 *   img_feat [B, d_img], txt_feat [B, d_txt] fp32
 *   drop_mask: caller-drawn uint8 [B, D] keep-mask of self.drop (:719), kept values scaled by drop_scale =
 *              1/(1-p); overrides desc->drop_p.  NULL: eval (drop_p == 0) or seeded dropout (drop_p > 0)
 *   logits   [B, n_classes] fp32 out */
int mmrca_head_forward(const MmrcaHeadDesc* desc, const MmrcaHeadParams* params,
                       const float* img_feat, const float* txt_feat,
                       const uint8_t* drop_mask, float drop_scale,
                       float* logits, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of mmrca_head_forward given dL/dlogits (what autograd hands to the head when
 * loss.backward() runs, main_both.py:112).  Attention internals are recomputed on chip.
 *   grads: accumulated into (+=);  d_img_feat / d_txt_feat: written if non-NULL (fine-tune phase,
 *   main_both.py:687-694), may be NULL when the backbones are frozen. */
int mmrca_head_backward(const MmrcaHeadDesc* desc, const MmrcaHeadParams* params,
                        const float* img_feat, const float* txt_feat,
                        const uint8_t* drop_mask, float drop_scale,
                        const float* dlogits, const MmrcaHeadGrads* grads,
                        float* d_img_feat, float* d_txt_feat,
                        void* workspace, size_t workspace_bytes, void* stream);

/* CrossEntropyLoss forward + dL/dlogits (main_both.py:87-93,110): loss_out [1], dlogits [B, C]
 * (either may be NULL).  labels int64 [B]. */
int mmrca_cross_entropy(const float* logits, const int64_t* labels, const MmrcaCeDesc* ce,
                        int32_t batch, int32_t n_classes, float* loss_out, float* dlogits,
                        void* stream);

/* One training step of the head: forward + CrossEntropyLoss + backward in one call
 * (run_one_epoch body, main_both.py:106-112, restricted to the fusion head).
 * Writes logits [B, C] and loss [1], accumulates grads, optionally writes feature grads. */
int mmrca_head_train_step(const MmrcaHeadDesc* desc, const MmrcaHeadParams* params,
                          const float* img_feat, const float* txt_feat,
                          const uint8_t* drop_mask, float drop_scale,
                          const int64_t* labels, const MmrcaCeDesc* ce,
                          float* logits, float* loss_out, const MmrcaHeadGrads* grads,
                          float* d_img_feat, float* d_txt_feat,
                          void* workspace, size_t workspace_bytes, void* stream);

/* uint8 keep-mask [batch, width] (1 = kept) that the seeded dropout of a head with concat width `width` draws for
 * (seed, p): what the kernels regenerate on chip.  For the fp32 kernels, tests and debugging. */
int mmrca_dropout_mask(uint64_t seed, float p, int32_t batch, int32_t width, uint8_t* mask_out, void* stream);

/* Stand-alone attention block = SelfAttention.forward (multimodal_model.py:51-68, x_kv == x_q,
 * reverse = 0) or ReverseCrossAttention.forward (:82-108).  x_q, x_kv: [B, 16, d_in];
 * out: [B, 16, d_v].  (d_in, d_kq, d_v) must be one of the head's blocks:
 * (d_in in {48, 64, 80}, 128, 96) or (96, 64, 48).  `normalise` != 0 applies the per-sample
 * L2 normalisation of :662-665 to x_q (== x_kv) first and writes the norms to norms_out [B].
 * compute must be MMRCA_COMPUTE_FP32 (the bf16 tensor-core pipeline exists only as the fused head);
 * scratch is unused (mmrca_attention_forward_scratch_bytes() returns 0; kept for ABI stability). */
size_t mmrca_attention_forward_scratch_bytes(int32_t d_in, int32_t d_kq, int32_t d_v, int32_t compute);
int mmrca_attention_forward(const MmrcaAttnParams* p, const float* x_q, const float* x_kv,
                            int32_t batch, int32_t d_in, int32_t d_kq, int32_t d_v,
                            int32_t reverse, int32_t normalise, float* norms_out,
                            float* out, void* scratch, size_t scratch_bytes, int32_t compute, void* stream);

/* Backward of mmrca_attention_forward: recomputes the block from x_q / x_kv, then
 *   d_x_q, d_x_kv [B,16,d_in]: written (NULL to skip; for x_kv == x_q pass d_x_kv = NULL and d_x_q
 *   receives the sum);  grads: accumulated;  scratch: device buffer of
 *   mmrca_attention_backward_scratch_bytes(batch, d_kq, d_v) bytes. */
size_t mmrca_attention_backward_scratch_bytes(int32_t batch, int32_t d_kq, int32_t d_v);
int mmrca_attention_backward(const MmrcaAttnParams* p, const float* x_q, const float* x_kv,
                             const float* d_out, int32_t batch, int32_t d_in, int32_t d_kq,
                             int32_t d_v, int32_t reverse, const MmrcaAttnGrads* grads,
                             float* d_x_q, float* d_x_kv, void* scratch, size_t scratch_bytes,
                             int32_t compute, void* stream);

/* ---- Hierarchical late-fusion head: Hierarchical.forward after the backbones and the two AvgPool2d
 * (multimodal_model.py:777-816) + CrossEntropyLoss + backward with frozen backbones.  bf16 tcgen05 GEMMs, fp32
 * accumulate: the 2e-2-absolute logits contract.  Fixed by the reference: 4 classes, hidden width 512, image
 * concat 1280 + 2560 + 2048, text concat 3 x 768 (:294-296). ---- */
typedef struct MmrcaHierParams {
  const float* w_img; const float* b_img; /* final_hierarchical_image  [512, 5888], [512]  (:294) */
  const float* w_txt; const float* b_txt; /* final_hierarchical_text   [512, 2304], [512]  (:295) */
  const float* w_all; const float* b_all; /* final_hierarchical_all    [4, 1024],  [4]     (:296) */
} MmrcaHierParams;
typedef struct MmrcaHierGrads { /* accumulated into (+=), 16-byte aligned; w_all / b_all may be NULL */
  float* w_img; float* b_img; float* w_txt; float* b_txt; float* w_all; float* b_all;
} MmrcaHierGrads;
#define MMRCA_HIER_FEATURE_GRADS 1u /* the backward will be asked for the six feature gradients (fine-tune phase): sizes the workspace */
typedef struct MmrcaHierDesc {
  int32_t batch;
  int32_t n_classes;  /* 4 */
  float drop_p;       /* self.drop on both concats (:805-806): seeded mask over the virtual concat [image 5888 | text 2304] */
  uint32_t flags;     /* MMRCA_HIER_* */
  uint64_t drop_seed;
} MmrcaHierDesc;
size_t mmrca_hier_workspace_bytes(const MmrcaHierDesc* desc);
/* feats[6]: pooled image [B,1280], AvgPool(7)+flatten of the 160-channel stage [B,2560], AvgPool(6)+flatten of the
 * 512-channel stage [B,2048], text CLS of the last layer / hidden_states[2] / hidden_states[4] [B,768] each (fp32,
 * 16-byte aligned).  drop_mask: caller-drawn uint8 keep mask [B, 8192] (image concat columns first), kept values
 * scaled by drop_scale; NULL: eval (drop_p == 0) or the seeded mask (mmrca_dropout_mask(seed, p, B, 8192)). */
int mmrca_hier_forward(const MmrcaHierDesc* desc, const MmrcaHierParams* params, const float* const* feats,
                       const uint8_t* drop_mask, float drop_scale, float* logits, void* workspace,
                       size_t workspace_bytes, void* stream);
/* needs the unmodified workspace of the forward.  d_feats: NULL (frozen backbones) or six output pointers shaped like
 * feats (fine-tune phase, main_both.py:687-694; needs MMRCA_HIER_FEATURE_GRADS in desc->flags of forward and backward and the
 * forward's feats / drop_mask / drop_scale again): d(loss)/d(feature) through the dropout, the concat and the six L2 norms. */
int mmrca_hier_backward(const MmrcaHierDesc* desc, const MmrcaHierParams* params, const float* dlogits,
                        const MmrcaHierGrads* grads, void* workspace, size_t workspace_bytes, void* stream);
int mmrca_hier_backward_features(const MmrcaHierDesc* desc, const MmrcaHierParams* params, const float* const* feats,
                                 const uint8_t* drop_mask, float drop_scale, float* const* d_feats, void* workspace,
                                 size_t workspace_bytes, void* stream);
int mmrca_hier_train_step(const MmrcaHierDesc* desc, const MmrcaHierParams* params, const float* const* feats,
                          const uint8_t* drop_mask, float drop_scale, const int64_t* labels, const MmrcaCeDesc* ce,
                          float* logits, float* loss_out, const MmrcaHierGrads* grads, void* workspace,
                          size_t workspace_bytes, void* stream);

/* ---- Token-level attention blocks (BASELINE.json configs[4]): SelfAttention.forward (multimodal_model.py:51-68) and
 * ReverseCrossAttention.forward (:82-108) on real token sequences x [B, L, d_in] with 2 <= L <= 256 (ViT-L/16: 197 x 1024,
 * RoBERTa: 256 x 768) instead of the head's 16 pseudo-tokens.  Square attention like the reference (:93): x_q and x_kv share L.
 * bf16 tensor-core path (2e-2 absolute contract): the Q | K | V projection is a TMA-fed tcgen05 GEMM over the tokens, the
 * attention core one kernel per (sample, 128-query tile).
 *   x_q, x_kv: bf16 [B, L, d_in_q] / [B, L, d_in_kv], 16-byte aligned; x_kv == NULL or == x_q: self attention.
 *   (d_kq, d_v) in {(128, 96), (64, 48)}; p: fp32 parameters in torch.nn.Linear layout; out: fp32 [B, L, d_v]. ---- */
#define MMRCA_TOKEN_WEIGHTS_READY 1u /* the workspace already holds this block's bf16 weights (an earlier call with the same,
                                       unchanged parameters on the same workspace): skip their conversion */
#define MMRCA_TOKEN_TRAINING 2u      /* the forward keeps the attention weights for mmrca_token_attention_backward (larger
                                       workspace; set on the descriptor of the workspace query, forward and backward) */
#define MMRCA_TOKEN_OUT_BF16 4u      /* forward: `out` is bf16 [B, L, d_v] (the activation format of a following block) instead of fp32 */
typedef struct MmrcaTokenDesc {
  int32_t batch, seq_len, d_in_q, d_in_kv, d_kq, d_v;
  int32_t reverse;    /* (1 - A) / (L - 1) weights (:95-99) */
  uint32_t flags;     /* MMRCA_TOKEN_* */
} MmrcaTokenDesc;
size_t mmrca_token_attention_workspace_bytes(const MmrcaTokenDesc* desc);
int mmrca_token_attention_forward(const MmrcaTokenDesc* desc, const MmrcaAttnParams* p, const void* x_q, const void* x_kv,
                                  float* out, void* workspace, size_t workspace_bytes, void* stream);
/* Backward of the block (what loss.backward() does to SelfAttention.forward / ReverseCrossAttention.forward, multimodal_model.py:51-68,
 * :82-108): needs the unmodified workspace of the MMRCA_TOKEN_TRAINING forward and the same x_q / x_kv.
 *   d_out: fp32 [B, L, d_v];  grads: accumulated into (+=, device atomics);  d_x_q / d_x_kv: fp32 [B, L, d_in], WRITTEN
 *   (NULL to skip; self attention: d_x_q receives the sum over the three projections, d_x_kv must be NULL).
 *   d_in a multiple of 16.  bf16 operands, fp32 accumulation. */
int mmrca_token_attention_backward(const MmrcaTokenDesc* desc, const MmrcaAttnParams* p, const void* x_q, const void* x_kv,
                                   const float* d_out, const MmrcaAttnGrads* grads, float* d_x_q, float* d_x_kv,
                                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- Classic / Normalized late-fusion heads (--late_fusion=classic | normalized): everything after the backbones in
 * EffV2MediumAndDistilbertClassic / ...Normalized.forward (multimodal_model.py:489-579) - the image and text projections
 * into the shared fusion dimension H = num_neurons_FC (:521-522, :566-567), for "normalized" their row L2 normalisation
 * (:569-570, no epsilon), concat + concat_layer (:524-527), self.drop (:528), fc_layer (:529) - plus CrossEntropyLoss and the
 * backward, fp32 (the 1e-4 contract).  (The reference hands image_to_hidden_size the extractor's TUPLE, which raises a
 * TypeError; the pooled vector, its third element, is what is meant and what img_feat is.) ---- */
#define MMRCA_FUSION_NORMALIZED 1u
#define MMRCA_FUSION_BF16 2u /* the two projections and their weight gradients as bf16 tcgen05 GEMMs (fp32 accumulate): the
                                2e-2-absolute logits contract instead of 1e-4 relative; hidden % 16 == 0, <= 256, feature
                                widths % 16 == 0, feature tensors and gradient tensors 16-byte aligned.  The backward needs the
                                unmodified workspace of the forward. */
typedef struct MmrcaFusionParams {
  const float* w_img; const float* b_img; /* image_to_hidden_size [H, d_img], [H]   (:199-201) */
  const float* w_txt; const float* b_txt; /* text_to_hidden_size  [H, d_txt], [H]   (:203-206) */
  const float* w_cat; const float* b_cat; /* concat_layer         [H, 2H],    [H]   (:208-210) */
  const float* w_fc;  const float* b_fc;  /* fc_layer             [n_classes, H], [n_classes] (:212) */
} MmrcaFusionParams;
typedef struct MmrcaFusionGrads { /* accumulated into (+=) */
  float* w_img; float* b_img; float* w_txt; float* b_txt; float* w_cat; float* b_cat; float* w_fc; float* b_fc;
} MmrcaFusionGrads;
typedef struct MmrcaFusionDesc {
  int32_t batch, d_img, d_txt, hidden, n_classes;
  uint32_t flags;     /* MMRCA_FUSION_NORMALIZED | MMRCA_FUSION_BF16 */
  float drop_p;       /* self.drop on the concat_layer output [B, H]: seeded mask = mmrca_dropout_mask(seed, p, B, H) */
  uint64_t drop_seed;
} MmrcaFusionDesc;
size_t mmrca_fusion_workspace_bytes(const MmrcaFusionDesc* desc);
/* drop_mask: caller-drawn uint8 keep mask [B, H] (kept values scaled by drop_scale) or NULL (eval / seeded) */
int mmrca_fusion_forward(const MmrcaFusionDesc* desc, const MmrcaFusionParams* params, const float* img_feat,
                         const float* txt_feat, const uint8_t* drop_mask, float drop_scale, float* logits,
                         void* workspace, size_t workspace_bytes, void* stream);
/* needs the unmodified workspace of the forward; d_img_feat / d_txt_feat: written if non-NULL (fine-tune phase) */
int mmrca_fusion_backward(const MmrcaFusionDesc* desc, const MmrcaFusionParams* params, const float* img_feat,
                          const float* txt_feat, const uint8_t* drop_mask, float drop_scale, const float* dlogits,
                          const MmrcaFusionGrads* grads, float* d_img_feat, float* d_txt_feat, void* workspace,
                          size_t workspace_bytes, void* stream);
int mmrca_fusion_train_step(const MmrcaFusionDesc* desc, const MmrcaFusionParams* params, const float* img_feat,
                            const float* txt_feat, const uint8_t* drop_mask, float drop_scale, const int64_t* labels,
                            const MmrcaCeDesc* ce, float* logits, float* loss_out, const MmrcaFusionGrads* grads,
                            float* d_img_feat, float* d_txt_feat, void* workspace, size_t workspace_bytes, void* stream);

/* One-shot all-reduce (mean over `world` ranks) of the flat head-gradient bucket over NVLink peer memory: the single
 * collective of a data-parallel step (SURVEY.md §8 e; the reference never ran multi-GPU, stock DDP would call NCCL
 * here).  staging[i] / pads[i]: rank i's symmetric staging buffer (2 * n_pad floats) and flag pad
 * (mmrca_peer_allreduce_pad_bytes(world) bytes, zeroed once), mapped into this process (torch symmetric memory or CUDA
 * IPC; host arrays of device pointers).  step = 1, 2, 3, ...: the same sequence on every rank.  Every rank calls it once
 * per step on its stream; flat is reduced in place, bit-identical on all ranks. */
int mmrca_peer_allreduce_mean(float* flat, int32_t n, int32_t n_pad, const void* const* staging, void* const* pads,
                              int32_t rank, int32_t world, uint32_t step, void* stream);
int mmrca_peer_allreduce_pad_bytes(int32_t world);
/* The wait for a peer's flag is bounded (~2 s): a rank that never arrives does not hang the GPU; the kernel then leaves the
 * bucket unreduced and records the peer in the status word behind the flags of this rank's own pad.  This call reads it
 * (the one peer function that synchronises `stream`): 0 = healthy, 1 + r = rank r timed out, < 0 = -MMRCA_ERR_*. */
int mmrca_peer_allreduce_status(const void* own_pad, int32_t world, void* stream);

/* ---- around the head ---- */
/* Feature hand-off from the stock backbones, one launch: txt_out[b][:] = hidden[b][0][:] (the CLS row of the text
 * backbone's last hidden state, multimodal_model.py:651-658) and img_out[b][c] = mean over the hw positions of the image
 * backbone's final feature map (avgpool + flatten, :25-36).  hidden: [B][T][d_txt] with row b at b * hidden_batch_stride
 * elements; fmap: contiguous [B][channels][hw] (channels_last = 0) or [B][hw][channels] (channels_last = 1); *_bf16: the
 * array holds bf16 instead of fp32.  bf16 outputs are what MMRCA_FLAG_FEATURES_BF16 takes. */
int mmrca_feature_handoff(const void* hidden, int32_t hidden_bf16, int64_t hidden_batch_stride, int32_t d_txt,
                          const void* fmap, int32_t fmap_bf16, int32_t channels, int32_t hw, int32_t channels_last,
                          int32_t batch, void* txt_out, void* img_out, int32_t out_bf16, void* stream);
/* Optimizer step over ONE contiguous fp32 parameter bucket and its gradient bucket (n floats, n % 4 == 0, 16-byte
 * aligned): torch.optim.SGD / torch.optim.AdamW semantics (main_both.py:544-552: lr, weight_decay = --reg).
 * sgd: momentum_buf may be NULL when momentum == 0; first_step != 0: the momentum buffer is initialised with the
 * gradient (torch's first step).  adamw: step = 1, 2, ... (bias correction). */
int mmrca_sgd_step(float* params, const float* grads, float* momentum_buf, int64_t n, float lr, float momentum,
                   float dampening, float weight_decay, int32_t nesterov, int32_t first_step, void* stream);
int mmrca_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                     float beta1, float beta2, float eps, float weight_decay, int32_t step, void* stream);

/* Per-kernel timing for roofline reports: between begin and end every kernel this library launches on
 * the calling thread is bracketed by a pair of CUDA events on ITS launch stream (up to max_records
 * launches).  mmrca_timing_end synchronises those events (the only call here that blocks), writes up to
 * max_out records and returns the number of launches recorded, or a negative MMRCA_ERR_* code. */
typedef struct MmrcaKernelTime {
  const char* name; /* static string, e.g. "attn_bwd<80,128,96,self>" */
  float ms;
} MmrcaKernelTime;
int mmrca_timing_begin(int32_t max_records);
int mmrca_timing_end(MmrcaKernelTime* out, int32_t max_out);

/* Development aid: a device buffer of 1024 int64 that receives per-phase clock64() stamps of CTA 0 of the profiled
 * tile kernel (kernel 0: SA backward, 1: CA backward) on its next launches (NULL switches it off; off by default). */
int mmrca_dev_set_debug(void* device_buffer_1024_int64, int32_t kernel);

/* Diagnostic: one 128 x N x K bf16 tcgen05 GEMM (fp32 accumulate in TMEM) through the library's operand
 * staging.  mode bit 0: b is [K][N] (MN-major) instead of [N][K]; bit 1: a is [K][128] instead of [128][K].
 * mode 16: a is [K][128], b is [K][N], both staged as the 128-byte-swizzled MN-major layout a SWIZZLE_128B tensor-map box lands.
 * out[128][N] = A B^T.  N % 16 == 0, 16 <= N <= 256, K % 16 == 0, operands must fit shared memory. */
int mmrca_dev_umma_selftest(int32_t mode, const float* a, const float* b, float* out, int32_t n, int32_t k,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMRCA_H_ */

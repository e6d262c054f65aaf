/*
 * toad_b200.h -- C ABI of libtoad_b200.so: the B200 (sm_100a) implementation of
 * the attention-MIL hot path of mahmoodlab/TOAD.
 *
 * The reference has no FFI of its own (it is pure PyTorch); the boundary it
 * offers is the nn.Module surface of models/model_toad.py.  Each entry point
 * below replaces the ATen/cuBLAS call chain of one reference method and is
 * what a ctypes binding in the reference's models/model_toad.py would call
 * (see INTEGRATION.md):
 *
 *   toad_fwd              <- TOAD_fc_mtl_concat.forward        models/model_toad.py:90-116
 *                            (TOAD_FLAG_ATTENTION_ONLY: the early return at :92-94)
 *   toad_bwd              <- autograd of that forward under the training loss,
 *                            utils/core_utils_mtl_concat.py:213-215,231
 *   toad_attn_gated_fwd   <- Attn_Net_Gated.forward            models/model_toad.py:36-41
 *   toad_topk             <- torch.topk over results['A'][t]   (SURVEY.md F6 / config 5;
 *                            the k=1 uses at models/model_toad.py:102,106)
 *   toad_linear_bf16x3    <- one nn.Linear(+ReLU), the building block (models/model_toad.py:59,62)
 *   toad_resnet_fwd       <- ResNet_Baseline.forward            models/resnet_custom.py:96-109
 *
 * Conventions
 *   - All pointers are DEVICE pointers to fp32 (or int64 where stated),
 *     row-major and contiguous; weights are [out_features, in_features] exactly
 *     as nn.Linear stores them.  The library reads parameters in place.
 *   - The caller owns every buffer, including the workspace (query its size
 *     with the *_workspace_bytes function; 256-byte aligned).  The library
 *     keeps no global mutable state and never frees caller memory.
 *   - Every call is asynchronous on `stream`; no implicit synchronisation.
 *   - Return 0 on success, a negative TOAD_ERR_* for argument errors, or a
 *     positive cudaError_t.  No C++ exception crosses this boundary.
 *   - There is no CPU fallback: without a CUDA device the calls fail.
 */
#ifndef TOAD_B200_H
#define TOAD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOAD_ABI_VERSION 3

typedef void* toad_stream_t; /* cudaStream_t */

enum {
  TOAD_OK = 0,
  TOAD_ERR_ARG = -1,         /* null pointer / bad size */
  TOAD_ERR_WORKSPACE = -2,   /* workspace too small or misaligned */
  TOAD_ERR_UNSUPPORTED = -3, /* dims outside what the kernels are built for */
  TOAD_ERR_DRIVER = -4       /* could not obtain a driver entry point (TMA descriptor) */
};

/* forward flags */
#define TOAD_FLAG_ATTENTION_ONLY 1u /* stop after the raw attention scores (model_toad.py:92-94) */
#define TOAD_FLAG_SIMT_FP32 2u      /* fp32 CUDA-core GEMMs instead of the tcgen05 split-bf16 path */
#define TOAD_FLAG_SAVE_ACTS 4u      /* also store h1,h,a,b (fp32) for toad_bwd */
#define TOAD_FLAG_DROPOUT 16u        /* training-mode nn.Dropout on h1, h, a, b (model_toad.py:27-29,60-64); needs `saved` */
#define TOAD_FLAG_TC_SINGLE_CTA 8u  /* debug: every tcgen05 GEMM with cta_group::1 (one CTA per 128-row tile) */
#define TOAD_FLAG_REUSE_WEIGHT_PLANES 128u /* the bf16 (hi,lo) weight planes left in `workspace` by an earlier call with the
                                            * same dims (any n_patches), tensor-core path and UNCHANGED parameters are still
                                            * valid: skip the 3 weight-split launches (eval loops; the caller tracks versions) */
#define TOAD_FLAG_BWD_TRANSPOSED 256u /* debug (toad_bwd): wgrads on K-major operands from explicitly transposed planes
                                       * instead of the MN-major operand path */
#define TOAD_FLAG_FC2_WIDE 64u      /* debug: fc2 on 256 x 512 pair tiles like fc1 */
#define TOAD_FLAG_TC_PAIR_ALL 32u   /* debug: every tcgen05 GEMM as CTA pairs (cta_group::2), including the fp32-fed fc1 */

/* Layer widths of TOAD_fc_mtl_concat (model_toad.py:56): "big" = {1024,512,384}, "small" = {1024,512,256}. */
typedef struct {
  int32_t in_dim;    /* 1024 */
  int32_t hid_dim;   /* 512  */
  int32_t attn_dim;  /* 384 or 256 */
  int32_t n_tasks;   /* 2 */
  int32_t n_classes; /* rows of `classifier` */
} toad_dims_t;

/* The 14 parameter tensors, in state_dict order (model_toad.py:59-73). */
typedef struct {
  const float* w1;    /* attention_net.0.weight              [hid, in]   */
  const float* b1;    /* attention_net.0.bias                [hid]       */
  const float* w2;    /* attention_net.2.weight              [hid, hid]  */
  const float* b2;    /* attention_net.2.bias                [hid]       */
  const float* wa;    /* attention_net.4.attention_a.0.weight [D, hid]   */
  const float* ba;    /* attention_net.4.attention_a.0.bias   [D]        */
  const float* wb;    /* attention_net.4.attention_b.0.weight [D, hid]   */
  const float* bb;    /* attention_net.4.attention_b.0.bias   [D]        */
  const float* wc;    /* attention_net.4.attention_c.weight   [n_tasks, D] */
  const float* bc;    /* attention_net.4.attention_c.bias     [n_tasks]  */
  const float* wcls;  /* classifier.weight                   [n_classes, hid+1] */
  const float* bcls;  /* classifier.bias                     [n_classes] */
  const float* wsite; /* site_classifier.weight              [2, hid+1]  */
  const float* bsite; /* site_classifier.bias                [2]         */
} toad_params_t;

/* Outputs of the forward = the reference's results_dict (model_toad.py:109-116). */
typedef struct {
  float* a_raw;        /* 'A'          [n_tasks, N]  pre-softmax scores, contiguous */
  float* features;     /* 'features'   [n_tasks, hid+1]  (M with sex appended) */
  float* logits;       /* 'logits'     [n_classes] */
  float* y_prob;       /* 'Y_prob'     [n_classes] */
  int64_t* y_hat;      /* 'Y_hat'      [1] */
  float* site_logits;  /* 'site_logits'[2] */
  float* site_prob;    /* 'site_prob'  [2] */
  int64_t* site_hat;   /* 'site_hat'   [1] */
  float* softmax_stats;/* [n_tasks][2] = (row max, sum of exp) of a_raw; needed by toad_bwd */
} toad_fwd_out_t;

/* Activations kept for the backward (TOAD_FLAG_SAVE_ACTS) as the next layer saw them (i.e. after dropout
 * when TOAD_FLAG_DROPOUT is set), plus the dropout configuration: element i of activation L in
 * {1: h1, 2: h, 3: a, 4: b} is kept iff toad_dropout_hash(seed, L, i) >= p * 2^32 and kept values are
 * scaled by 1/(1-p).  The mask is a pure function of (seed, L, i): bitwise parity with torch's Philox
 * stream is impossible, distributional parity is what holds.
 * Which of h1 / h is needed depends on the path:
 *   TOAD_FLAG_SIMT_FP32: fp32 h1, h (planes ignored);
 *   tensor-core path:    the (hi, lo) bf16 planes h1_hi/lo, h_hi/lo -- hi = bf16(v), lo = bf16(v - hi), the operand
 *                        format of the split-bf16 GEMMs, so the forward keeps exactly what fc2 / the gate consumed and
 *                        the backward contracts over them again without a conversion pass; fp32 h1 / h are optional
 *                        extras (written when non-NULL).
 * a, b are fp32 on both paths. */
typedef struct {
  float* h1; /* [N, hid] relu(fc1) */
  float* h;  /* [N, hid] relu(fc2) */
  float* a;  /* [N, D] tanh branch  */
  float* b;  /* [N, D] sigmoid branch */
  uint64_t dropout_seed;
  float dropout_p; /* 0.25 in the reference; used only with TOAD_FLAG_DROPOUT (forward) / when > 0 (backward) */
  void* h1_hi; /* [N, hid] bf16 */
  void* h1_lo;
  void* h_hi;
  void* h_lo;
} toad_saved_t;

/* The mask hash (host-callable, for tests): high 32 bits of splitmix64(seed, layer, index). */
uint32_t toad_dropout_hash(uint64_t seed, uint32_t layer, uint64_t index);

int toad_abi_version(void);
/* Hash of the sources this binary was built from (toad_b200/build.py passes -DTOAD_BUILD_ID): the Python binding
 * compares it with the sources on disk and rebuilds (or refuses to run) a stale library. */
const char* toad_build_id(void);
const char* toad_error_string(int code);

/* Number of fp32 elements of the flat parameter/gradient buffer and the offset of
 * each of the 14 tensors in it (state_dict order); offsets[14] = total. */
int toad_param_offsets(const toad_dims_t* dims, int64_t offsets[15]);

int toad_fwd_workspace_bytes(const toad_dims_t* dims, int64_t n_patches, uint32_t flags, size_t* bytes);

/* x: [N, in_dim]; sex: device float[1] (core_utils_mtl_concat.py:204).
 * `saved` may be NULL unless TOAD_FLAG_SAVE_ACTS.  With TOAD_FLAG_ATTENTION_ONLY
 * only out->a_raw is written. */
int toad_fwd(const toad_dims_t* dims, const toad_params_t* params, const float* x, int64_t n_patches,
             const float* sex, const toad_fwd_out_t* out, const toad_saved_t* saved, void* workspace,
             size_t workspace_bytes, uint32_t flags, toad_stream_t stream);

/* Batched eval forward (collate_MIL_mtl_concat, utils/utils.py:30-35, concatenates the bags of a batch the same
 * way): n_slides <= 16 slides stored back to back in x [offsets[n_slides], in_dim], slide s = rows
 * [offsets[s], offsets[s+1]) (offsets: HOST array, offsets[0] = 0, no empty slide).  The three trunk GEMMs run once
 * over all rows -- small bags fill the GPU together and share one set of launches --, the pooling / heads kernel runs
 * with one grid row per slide.  `out` points at slide-major blocks: a_raw [n_tasks, n_total] (slide s = columns
 * offsets[s]..), features [S, n_tasks, hid+1], logits / y_prob [S, n_classes], y_hat / site_hat [S] int64,
 * site_logits / site_prob [S, 2], softmax_stats [S, n_tasks, 2]; sex [S].  Per-row results are bit-identical to
 * toad_fwd on each slide; pooled results differ only by fp32 summation order. */
int toad_fwd_batch_workspace_bytes(const toad_dims_t* dims, int64_t n_total, int32_t n_slides, uint32_t flags,
                                   size_t* bytes);
int toad_fwd_batch(const toad_dims_t* dims, const toad_params_t* params, const float* x, const int64_t* offsets,
                   int32_t n_slides, const float* sex, const toad_fwd_out_t* out, void* workspace,
                   size_t workspace_bytes, uint32_t flags, toad_stream_t stream);

/* Per-stage device timing of toad_fwd (diagnostics; bench.py's roofline leg).  A profile handle
 * owns CUDA events for up to max_calls forwards; toad_fwd_profiled records an event between the
 * stages on `stream` (no synchronisation); toad_profile_read waits for the recorded events,
 * returns the summed milliseconds per stage and the number of calls, and resets the handle.
 * Stages: 0 weight split, 1 fc1 GEMM, 2 fc2 GEMM, 3 gated-attention GEMM(s), 4 pooling tail. */
#define TOAD_N_STAGES 5
int toad_profile_create(void** prof, int32_t max_calls);
int toad_profile_destroy(void* prof);
int toad_profile_read(void* prof, double stage_ms[TOAD_N_STAGES], int32_t* n_calls);
int toad_fwd_profiled(const toad_dims_t* dims, const toad_params_t* params, const float* x, int64_t n_patches,
                      const float* sex, const toad_fwd_out_t* out, const toad_saved_t* saved, void* workspace,
                      size_t workspace_bytes, uint32_t flags, toad_stream_t stream, void* prof);

int toad_bwd_workspace_bytes(const toad_dims_t* dims, int64_t n_patches, uint32_t flags, size_t* bytes);

/* flags: 0 = the five large contractions (2 dgrad, 3 wgrad with split-K) on the tcgen05 split-bf16 GEMM;
 * TOAD_FLAG_SIMT_FP32 = all of them on fp32 CUDA cores.
 * Gradients of sum(dlogits*logits) + sum(dsite_logits*site_logits) w.r.t. the 14
 * parameters, written (not accumulated) into grad_flat in toad_param_offsets order.
 * fwd_out / saved are the buffers the forward (with TOAD_FLAG_SAVE_ACTS) filled. */
int toad_bwd(const toad_dims_t* dims, const toad_params_t* params, const float* x, int64_t n_patches,
             const toad_fwd_out_t* fwd_out, const toad_saved_t* saved, const float* dlogits,
             const float* dsite_logits, float* grad_flat, void* workspace, size_t workspace_bytes,
             uint32_t flags, toad_stream_t stream);

/* The loss of the reference training loop (utils/core_utils_mtl_concat.py:213-215) and its gradient in one launch:
 * loss3 = {w_cls*CE(logits,label) + w_site*CE(site_logits,site), CE(logits,label), CE(site_logits,site)} and
 * dlogits[n_classes], dsite_logits[2] = d loss3[0] / d(logits, site_logits), ready for toad_bwd.  label / site are
 * device int64[1] (the loop's label.to(device)); a target outside its range gives NaN for that head (the
 * reference trips a device assert).  The reference uses w_cls = 0.75, w_site = 0.25. */
int toad_ce_loss_grad(const float* logits, const float* site_logits, int32_t n_classes, const int64_t* label,
                      const int64_t* site, float w_cls, float w_site, float* loss3, float* dlogits,
                      float* dsite_logits, toad_stream_t stream);

/* torch.optim.Adam as the reference builds it (utils/utils.py:65: lr, weight_decay = L2 term added to the
 * gradient, betas, eps, amsgrad off) applied to all 14 parameter tensors in place, one launch.
 * grad_flat / exp_avg / exp_avg_sq: flat fp32 buffers in toad_param_offsets order (the moments start at zero);
 * step = 1 for the first update; the hyper-parameters are doubles (python floats on the caller's side: 1 - beta
 * and the bias corrections are formed in double, as torch does); grad_scale multiplies the gradient first (1/world_size after a summing all-reduce). */
int toad_adam_step(const toad_dims_t* dims, const toad_params_t* params, const float* grad_flat, float* exp_avg,
                   float* exp_avg_sq, int64_t step, double lr, double beta1, double beta2, double eps,
                   double weight_decay, float grad_scale, toad_stream_t stream);

/* Standalone gated attention head: A[N, n_tasks] = Wc(tanh(Wa x + ba) * sigmoid(Wb x + bb)) + bc. */
/* Training-side state of the standalone block: the two gate branches as the score contraction saw them (after
 * dropout when TOAD_FLAG_DROPOUT is set; element i of the tanh / sigmoid branch uses dropout layers 3 / 4 of
 * toad_dropout_hash, index row*D + column -- the same convention as inside toad_fwd). */
typedef struct {
  float* a;              /* [n, D] */
  float* b;              /* [n, D] */
  uint64_t dropout_seed;
  float dropout_p;       /* nn.Dropout(0.25) in the reference (model_toad.py:27-29) */
} toad_attn_saved_t;

int toad_attn_gated_workspace_bytes(int32_t L, int32_t D, int32_t n_tasks, int64_t n, uint32_t flags, size_t* bytes);
/* saved: NULL for inference; given, the tensor-core path stores a / b (and applies dropout with TOAD_FLAG_DROPOUT). */
int toad_attn_gated_fwd(int32_t L, int32_t D, int32_t n_tasks, const float* wa, const float* ba,
                        const float* wb, const float* bb, const float* wc, const float* bc,
                        const float* x, int64_t n, float* A_out, const toad_attn_saved_t* saved, void* workspace,
                        size_t workspace_bytes, uint32_t flags, toad_stream_t stream);
/* Backward of the standalone block (autograd of model_toad.py:36-41): given dA [n, n_tasks] writes the gradients of
 * the six parameters and, when dx != NULL, of the input x [n, L].  n_tasks <= 2, L % 256 == 0 (tensor-core wgrad /
 * dgrad tiles).  `saved` is what toad_attn_gated_fwd filled for the same x. */
int toad_attn_gated_bwd_workspace_bytes(int32_t L, int32_t D, int32_t n_tasks, int64_t n, size_t* bytes);
int toad_attn_gated_bwd(int32_t L, int32_t D, int32_t n_tasks, const float* wa, const float* wb, const float* wc,
                        const float* x, int64_t n, const toad_attn_saved_t* saved, const float* dA, float* d_wa,
                        float* d_ba, float* d_wb, float* d_bb, float* d_wc, float* d_bc, float* dx, void* workspace,
                        size_t workspace_bytes, toad_stream_t stream);

/* top-k of one score row: values descending, ties -> lower index first (k <= 2048).  One cooperative launch
 * over all SMs; `workspace` (toad_topk_workspace_bytes, 16-byte aligned) holds the radix histograms and is
 * zeroed by the call itself. */
int toad_topk_workspace_bytes(int64_t n, int32_t k, size_t* bytes);
int toad_topk(const float* scores, int64_t n, int32_t k, float* out_vals, int64_t* out_idx,
              void* workspace, size_t workspace_bytes, toad_stream_t stream);

/* out[i, :] = table[idx[i], :], rows of row_bytes bytes (a multiple of 4): e.g. the `coords` [N, 2] rows of the
 * top-k patches for a heatmap (the reference's h5 bags keep coords next to features,
 * datasets/dataset_mtl_concat.py:377-383).  Out-of-range indices give zero rows. */
int toad_gather_rows(const void* table, int64_t n_rows, int32_t row_bytes, const int64_t* idx, int32_t k, void* out,
                     toad_stream_t stream);

/* y[M, N] = act(x[M, K] . w[N, K]^T + bias[N]) on the tcgen05 3-pass split-bf16 path.
 * relu != 0 applies ReLU; bias may be NULL.  K % 64 == 0, N % 64 == 0.
 * variant: bit 0 = feed x as pre-split (hi,lo) bf16 planes through TMA instead of converting
 * fp32 rows in the kernel; bit 1 = CTA pairs (cta_group::2, 256-row tiles); bit 6 = 512-wide tiles (with bit 1); bits 4-5 = force the N tile (1: 64, 2: 128, 3: 256; 0: largest that
 * divides N).  Both paths give the same result; the knob exists so tests cover every tile shape. */
int toad_linear_workspace_bytes(int64_t m, int32_t n, int32_t k, size_t* bytes);
int toad_linear_bf16x3(const float* x, const float* w, const float* bias, float* y, int64_t m, int32_t n,
                       int32_t k, int32_t relu, int32_t variant, void* workspace, size_t workspace_bytes,
                       toad_stream_t stream);

/* ---- resnet50_baseline: ResNet-50 truncated after layer3 + global average pool, eval mode
 * (models/resnet_custom.py:57-109; Bottleneck_Baseline :19-55).  BatchNorm (running statistics) is
 * folded into the convolution weights by toad_resnet_prepare; activations are NHWC internally.
 *
 * tensors: HOST array of TOAD_RESNET_N_TENSORS device pointers (fp32), the module's state_dict in
 * order with the `num_batches_tracked` entries skipped, i.e. for each of the 43 convolutions
 * {conv.weight [Co,Ci,k,k], bn.weight, bn.bias, bn.running_mean, bn.running_var}.
 * x: [B, 3, H, W] fp32 NCHW, H and W multiples of 16, W <= 512 (224 x 224, 256 x 256, ...);  out: [B, 1024] fp32.
 *
 * flags (the same value for prepare / workspace_bytes / fwd of one prepared block):
 *   0                       activations stored between layers as ONE fp16 NHWC plane (2 B / element), weights as fp16
 *                           (hi, lo) pairs, two tensor-core passes per product, fp32 accumulation.  Activation rounding is
 *                           2^-11 relative, the input precision of the TF32 convolutions cuDNN runs for the reference by
 *                           default; features land within ~4e-4 of the feature scale of the fp64 reference.
 *   TOAD_RESNET_FLAG_EXACT  activations as (hi, lo) bf16 plane pairs (4 B / element), three passes: fp32-class accuracy
 *                           (~7e-5 of the feature scale) at twice the HBM traffic and 1.5x the tensor work. */
#define TOAD_RESNET_N_TENSORS 215
#define TOAD_RESNET_FLAG_EXACT 1u
int toad_resnet_prepared_bytes(size_t* bytes);
int toad_resnet_prepare(const float* const* tensors, int32_t n_tensors, void* prepared, size_t prepared_bytes,
                        uint32_t flags, toad_stream_t stream);
int toad_resnet_workspace_bytes(int32_t batch, int32_t height, int32_t width, uint32_t flags, size_t* bytes);
int toad_resnet_fwd(const void* prepared, const float* x, int32_t batch, int32_t height, int32_t width, float* out,
                    void* workspace, size_t workspace_bytes, uint32_t flags, toad_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TOAD_B200_H */

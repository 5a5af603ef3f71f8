/*
 * dml_b200.h -- C ABI of libdml_b200.so: hand-written sm_100a CUDA kernels for the DMLNet
 * per-pixel metric-learning hot path (distance head -> labels / EDS / MMSP scores -> exact
 * AUROC / AUPR / FPR@95; DCE+VL loss forward/backward; per-class masked mean).
 *
 * The reference (Jun-CEN/Open-World-Semantic-Segmentation) is 100 % Python and has no FFI;
 * this header is what a ctypes binding of its operator surface binds.  Each entry point
 * cites the reference lines it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - tensors are dense, row-major, in the stated layout; fp32 unless stated;
 *   - the library never allocates: callers pass outputs and workspaces
 *     (`*_workspace_bytes` tells how much);
 *   - kernels are enqueued on `stream` (a cudaStream_t) of the CURRENT device and the
 *     call returns without synchronising; no global mutable state => re-entrant from
 *     one host thread per GPU (nn.DataParallel's threading model);
 *   - return value: DML_OK (0) or a negative DML_ERR_* code; `dml_error_string` decodes.
 */
#ifndef DML_B200_H_
#define DML_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DML_B200_ABI_VERSION 4

#if defined(__GNUC__)
#define DML_API __attribute__((visibility("default")))
#else
#define DML_API
#endif

typedef void* dml_stream_t; /* cudaStream_t */

enum {
  DML_OK = 0,
  DML_ERR_INVALID_ARG = -1,   /* bad shape / null pointer / unsupported combination */
  DML_ERR_UNSUPPORTED_DIM = -2, /* embedding dim or class count outside the compiled range */
  DML_ERR_CUDA = -3,          /* a CUDA runtime call failed (see dml_last_cuda_error) */
  DML_ERR_WORKSPACE = -4      /* workspace too small */
};

DML_API int dml_abi_version(void);
DML_API const char* dml_error_string(int code);
/* cudaError_t of the most recent failed runtime call made by this thread's entry point. */
DML_API int dml_last_cuda_error(void);
/* Largest supported embedding dim D and class count K (compile-time unrolled kernels). */
DML_API int dml_max_dim(void);
/* Number of CUDA kernels this library has enqueued so far in this process (benchmark bookkeeping). */
DML_API unsigned long long dml_kernel_launches(void);

/* ------------------------------------------------------------------------------------ *
 * (a) Fused distance + score head.
 *
 * Replaces, in one pass over the NCHW embedding:
 *   anomaly/models/models.py:636-657                (distance logits, prototypes :614-618)
 *   DeepLabV3Plus-Pytorch/network/utils.py:89-118   (same, full resolution, NHWC `features`)
 *   anomaly/eval_ood_traditional.py:218             (argmax label)
 *   anomaly/eval_ood_traditional.py:276-278,288-290 (msp, maxlogit)
 *   anomaly/eval_ood_traditional.py:301-304         (EDS raw sum + clamp; :305 normalise is
 *                                                    dml_scores_finalize / fused key-gen)
 *   anomaly/eval_ood_traditional.py:434             (max-softmax for MMSP)
 *   DeepLabV3Plus-Pytorch/test_embedding.py:339-350 (argmax, max-softmax, dis_sum + clamp)
 *   DeepLabV3Plus-Pytorch/test_embedding.py:428-433,445 (NPM novel-prototype override, fp64)
 *   DeepLabV3Plus-Pytorch/metrics/stream_metrics.py:49-55, anomaly/utils.py:128-156
 *                                                   (confusion counts, when `gt` is given)
 *
 *   z[b,k,p] = - sum_d (x[b,d,p] - mu[k,d])^2
 * Prototypes: `mu == NULL` => mu = diag_m * I_K (requires K == D; the only case the
 * reference ships) computed with a cancellation-free leave-one-out sum of squares;
 * otherwise `mu` is a dense [K,D] table (direct form).
 * Every output pointer may be NULL (not produced).
 * ------------------------------------------------------------------------------------ */
typedef struct dml_head_params {
  uint32_t struct_bytes;      /* sizeof(dml_head_params), ABI guard */
  int32_t B, D, K, H, W;      /* x is [B,D,H,W] */
  const float* x;
  const float* mu;            /* [K,D] or NULL */
  float diag_m;               /* prototype magnitude when mu == NULL (reference: 3) */
  int32_t input_is_logits;    /* != 0: x already holds the logits z [B,K,H,W] (anomaly path, where the
                                 distances are taken at stride 8 and upsampled/averaged first,
                                 anomaly/eval_ood_traditional.py:198-210); scores only, K == D */
  int32_t score_first_class;  /* 0, or 1 for OOD.exclude_back (scores skip channel 0;
                                 labels/logits still use all classes) */
  float eds_clamp;            /* > 0: eds = min(eds, clamp) (400 anomaly / 1000 DeepLab); <= 0: none */

  /* NPM novel prototypes (optional): [n_novel, D] float64, override label where
     z_nov > novel_thr && z_nov > max_k z_k; label written = novel_label_base + j. */
  const double* mu_novel;
  int32_t n_novel;
  int32_t novel_label_base;
  double novel_thr;

  /* outputs */
  float* logits;              /* [B,K,H,W] */
  uint8_t* label_u8;          /* [B,H,W] argmax_k z (first max), after NPM override */
  int64_t* label_i64;         /* same as int64 (torch.max parity) */
  float* maxlogit;            /* [B,H,W] max_k z over score classes */
  float* eds;                 /* [B,H,W] -sum_k z over score classes, clamped */
  float* msp;                 /* [B,H,W] max_k softmax_k(z) over score classes */
  float* features_nhwc;       /* [B,H,W,D] contiguous copy of x (DeepLab return value) */
  double* novel_dist;         /* [n_novel,B,H,W] z_nov (float64) */
  float* minmax;              /* [B,4] per image (eds_min, eds_max, msp_min, msp_max); needs eds/msp
                                 to be computed (their map pointers may still be NULL) */
  int32_t want_eds_minmax, want_msp_minmax;

  /* fused confusion counts (optional) */
  const uint8_t* gt_u8;       /* [B,H,W] ground truth; or */
  const int64_t* gt_i64;      /* [B,H,W] */
  unsigned long long* confusion; /* [conf_rows, conf_cols] += count; gt outside [0,conf_rows) ignored */
  int32_t conf_rows, conf_cols;
  int32_t reference_order;    /* != 0 (mu == NULL, D < 16, logits requested): parity mode -- every z_k is rounded exactly
                                 like the reference's torch-CPU op sequence (anomaly/models/models.py:649-651: subtract,
                                 square, sum over the embedding dim in torch's CPU order), so that the distance logits are
                                 bit-identical to the reference's on identical embeddings.  Slower; for verification and for
                                 the tiny stride-8 maps of the anomaly path. */
} dml_head_params;

DML_API int dml_head_forward(const dml_head_params* p, dml_stream_t stream);

/* ------------------------------------------------------------------------------------ *
 * (a2) Fused multi-scale bilinear upsample + average + score head (anomaly path).
 *
 * Replaces, without materialising any full-resolution [B,K,H,W] tensor:
 *   anomaly/models/models.py:659-661         F.interpolate(z_s, segSize, 'bilinear', align_corners=False)
 *   anomaly/eval_ood_traditional.py:192-208  scores = scores + scores_tmp / len(imgSizes), per scale in order
 *   anomaly/eval_ood_traditional.py:212-218,276-305,434  (tmp_scores / argmax / msp / maxlogit / EDS as in
 *                                             dml_head_forward with input_is_logits)
 * z[s] are the stride-8 distance logits of scale s ([B,K,h[s],w[s]], e.g. dml_head_forward(logits=...) on
 * the decoder's embedding); the kernel evaluates, per output pixel and class,
 *   scores[k] = sum_s ( h0*(w0*z00 + w1*z01) + h1*(w0*z10 + w1*z11) ) / n_scales
 * with torch's upsample_bilinear2d source-index / lambda arithmetic in fp32 and the accumulation order of
 * the reference loop.  `reciprocal_average` == 0: correctly rounded division (torch CPU semantics);
 * != 0: multiplication by fl(1/n_scales) (what torch's CUDA div-by-scalar kernel computes).
 * `scores` (optional) receives the averaged full-resolution maps; with every score output NULL the call is
 * the plain "upsample + average" of :209-210 (`ft1`, fed with the low-resolution embeddings).
 * Remaining fields have the meaning they have in dml_head_params.
 * ------------------------------------------------------------------------------------ */
#define DML_MAX_SCALES 8
typedef struct dml_multiscale_params {
  uint32_t struct_bytes;      /* sizeof(dml_multiscale_params), ABI guard */
  int32_t B, K, H, W;         /* output maps are [B,*,H,W] */
  int32_t n_scales;           /* 1..DML_MAX_SCALES */
  const float* z[DML_MAX_SCALES]; /* [B,K,h[s],w[s]] */
  int32_t h[DML_MAX_SCALES], w[DML_MAX_SCALES];
  int32_t reciprocal_average;
  int32_t score_first_class;
  float eds_clamp;
  float* scores;              /* [B,K,H,W] or NULL */
  uint8_t* label_u8;
  int64_t* label_i64;
  float* maxlogit;
  float* eds;
  float* msp;
  float* minmax;              /* [B,4] */
  int32_t want_eds_minmax, want_msp_minmax;
  const uint8_t* gt_u8;
  const int64_t* gt_i64;
  unsigned long long* confusion;
  int32_t conf_rows, conf_cols;
} dml_multiscale_params;

DML_API int dml_multiscale_head_forward(const dml_multiscale_params* p, dml_stream_t stream);

/* Per-image min-max normalise + mix (anomaly/eval_ood_traditional.py:101-106,305,435,447-448):
 *   eds_n = (eds - min)/(max - min), msp_n likewise, c = 1/(1+exp(lambda (eds_n - thr))),
 *   mix = c*eds_n + (1-c)*msp_n.  In-place allowed (out == in).  Outputs may be NULL.
 * `complement` != 0 writes 1 - eds_n instead (DeepLabV3Plus-Pytorch/test_embedding.py:370). */
DML_API int dml_scores_finalize(const float* eds, const float* msp, const float* minmax /*[B,4]*/,
                        int32_t B, int64_t pixels_per_image, float lambda, float thr, int32_t complement,
                        float* eds_norm, float* msp_norm, float* mix, dml_stream_t stream);

/* Confusion counts on their own (stream_metrics.py:49-55; anomaly/utils.py:128-156):
 * confusion[gt*cols + pred] += 1 for gt in [0,rows).  Exactly one of gt_u8/gt_i64 and one of
 * pred_u8/pred_i64 is non-NULL. */
DML_API int dml_confusion(const uint8_t* gt_u8, const int64_t* gt_i64, const uint8_t* pred_u8,
                  const int64_t* pred_i64, int64_t n, int32_t rows, int32_t cols,
                  unsigned long long* confusion, dml_stream_t stream);

/* PLM merge (DeepLabV3Plus-Pytorch/test_self_distillation.py:292-297):
 * base[p] = novel_label where head[p] == novel_label. */
DML_API int dml_plm_merge(uint8_t* base_u8, int64_t* base_i64, const uint8_t* head_u8, const int64_t* head_i64,
                  int64_t n, int32_t novel_label, dml_stream_t stream);

/* ------------------------------------------------------------------------------------ *
 * (f-2) Final 1x1 classifier conv fused into the distance head (inference).
 *   anomaly/models/models.py:609 (conv_last[4] = nn.Conv2d(512, num_class, 1)) + :636-657 (distance block)
 *   DeepLabV3Plus-Pytorch/network/utils.py:23 (classifier[3] = nn.Conv2d(256, num_classes, 1))
 * features [B,C,H,W] (the post-ReLU output of the layer before), weight [K,C] (the conv's [K,C,1,1] weight), bias [K]
 * or NULL.  embedding[b,k,p] = sum_c weight[k,c] features[b,c,p] + bias[k] (fp32 FFMA, channel order);
 * logits[b,k,p] = -sum_d (embedding_d - diag_m [d = k])^2.  Either output may be NULL.  K <= 32. */
DML_API int dml_conv1x1_head_forward(const float* features, const float* weight, const float* bias, float diag_m, int32_t B,
                             int32_t C, int32_t K, int32_t H, int32_t W, float* embedding, float* logits, dml_stream_t stream);

/* ------------------------------------------------------------------------------------ *
 * (b) Fused DCE + VL (+ Inter) loss, forward and backward, straight from the embedding.
 *   anomaly/models/models.py:42-78 (CE + alpha*VL, ignore -1)
 *   DeepLabV3Plus-Pytorch/utils/loss.py:34-82 (line 79 form; shipped early return = alpha=beta=0)
 *   loss = (CE + alpha*VL + beta*Inter)/n  -- see SURVEY.md appendix A.6.
 * Forward writes 4 doubles: loss, CE, VL, Inter and the valid-pixel count as double in [4].
 * `partials` is a 16-byte aligned workspace of dml_loss_workspace_bytes(B,H,W).
 * Backward recomputes the distances and writes dL/dx [B,D,H,W] scaled by *grad_out (device scalar).
 * x_is_logits != 0: `x` already holds the logits z [B,K,H,W] (the reference criterion's signature,
 * utils/loss.py:34 `forward(logit, target, features_in)`); mu must be NULL and dx is dL/dz.
 * ------------------------------------------------------------------------------------ */
DML_API size_t dml_loss_workspace_bytes(int32_t B, int32_t H, int32_t W);
DML_API int dml_loss_forward(const float* x, int32_t x_is_logits, const float* mu, float diag_m, const uint8_t* target_u8,
                     const int64_t* target_i64, int64_t ignore_index, int32_t B, int32_t D, int32_t K,
                     int32_t H, int32_t W, double alpha, double beta, void* partials,
                     double* out5, dml_stream_t stream);
DML_API int dml_loss_backward(const float* x, int32_t x_is_logits, const float* mu, float diag_m, const uint8_t* target_u8,
                      const int64_t* target_i64, int64_t ignore_index, int32_t B, int32_t D, int32_t K,
                      int32_t H, int32_t W, double alpha, double beta, const double* out5,
                      const float* grad_out, float* dx, dml_stream_t stream);

/* ------------------------------------------------------------------------------------ *
 * (c) Masked per-class mean (few-shot novel prototype generation).
 *   DeepLabV3Plus-Pytorch/test_embedding.py:413-419, utils/loss.py:65-67
 * sums[b,c,:] = sum_{p: label[b,p]==c} x[b,:,p] (float64), counts[b,c] = #pixels.
 * x is NCHW (`nhwc`=0) or NHWC (`nhwc`=1).  Labels outside [0,n_cls) are skipped.
 * ------------------------------------------------------------------------------------ */
DML_API size_t dml_class_sums_workspace_bytes(int32_t B, int32_t D, int32_t n_cls, int64_t pixels_per_image);
DML_API int dml_class_sums(const float* x, int32_t nhwc, const uint8_t* label_u8, const int64_t* label_i64,
                   int32_t B, int32_t D, int64_t pixels_per_image, int32_t n_cls, void* workspace,
                   double* sums /*[B,n_cls,D]*/, long long* counts /*[B,n_cls]*/, dml_stream_t stream);

/* ------------------------------------------------------------------------------------ *
 * (d) Exact, tie-aware ranking metrics (AUROC, AUPR, FPR@recall) on the GPU.
 *   anomaly/anom_utils.py:25-65 (fpr_and_fdr_at_recall), :67-78 (get_measures -> sklearn
 *   roc_auc_score / average_precision_score), :95-116; anomaly/eval_ood_traditional.py:128-148.
 *
 * Pipeline: key generation (order-preserving u32 key of the score, positive flag packed
 * in bit 0) -> LSD radix sort (8-bit digits, decoupled look-back) per segment -> segmented
 * scan over distinct-score groups -> AUROC / AP / FPR reductions.  Segments are
 * independent (one per image for the reference's per-image semantics, one for pooled).
 * ------------------------------------------------------------------------------------ */
typedef struct dml_ood_result {
  double auroc, aupr, fpr;    /* NaN when the segment has a single class */
  long long n_pos, n_neg;
  long long n_nan;            /* NaN scores seen (Python raises ValueError like sklearn) */
  long long n_groups;         /* distinct score values */
} dml_ood_result;

/* Statistics of the order-preserving ("sortable") u32 image of the ranking key, used to choose
 * `key_base` for arbitrary scores: seg_stats (device, [n_seg,4] int64) = (min, max, n_nan, 0). */
DML_API int dml_ood_keystats(const float* values, int32_t score_kind, int32_t n_seg, int64_t seg_len,
                     long long* seg_stats, dml_stream_t stream);

/* Key generation.  score_kind: 0 = `conf` map, ranked as score = -conf (positives expected
 * at LOW conf; anomaly/eval_ood_traditional.py:139-141); 1 = plain score (higher = positive).
 * If `minmax` != NULL the kernel first applies the per-segment normalisation
 * (v - min)/(max - min) with min/max = minmax[seg*4 + minmax_slot*2 + {0,1}] in fp32, exactly as
 * NumPy does (fusing eval_ood_traditional.py:305 into key-gen), and optionally stores it to
 * `conf_out`.  Positives: gt label in `out_label_mask` (bit l set for label l, labels 0..63;
 * exactly one of gt_u8 / gt_i64 / pos_u8 is non-NULL; pos_u8 != 0 marks positives directly).
 * Packed key = ((sortable(key) - key_base) << 1) | positive; all keys of a call must lie in
 * [key_base, key_base + 2^31) -- true with key_base = 0x80000000 for any non-negative conf
 * (e.g. normalised maps in [0,1]); violations are counted, not silently wrapped.
 * seg_stats (device, [n_seg,4] int64) = (n_pos, n_nan, n_out_of_window, 0).
 * Optional fused score outputs (need minmax, minmax_slot == 0): with `msp` = raw max-softmax map,
 * `msp_norm_out` = its min-max normalisation (MMSP, eval_ood_traditional.py:434-435) and `mix_out` =
 * c*eds_n + (1-c)*mmsp with c = 1/(1+exp(lambda (eds_n - thr))) (:104-106,447-448).
 * `sort_workspace` != NULL (the workspace later passed to dml_ood_eval_segments for the same n_seg / seg_len):
 * the digit histograms of all radix passes are accumulated here, while the keys are still in registers, and
 * dml_ood_eval_segments is then called with hist_precomputed = 1 (the sort's separate counting read of the
 * keys disappears; key-gen is HBM-bound, the histogram shared-atomic-bound, so the two overlap). */
DML_API int dml_ood_keygen(const float* values, const float* minmax, int32_t minmax_slot, float* conf_out,
                   const uint8_t* gt_u8, const int64_t* gt_i64, uint64_t out_label_mask,
                   const uint8_t* pos_u8, int32_t score_kind, uint32_t key_base, int32_t n_seg,
                   int64_t seg_len, uint32_t* keys, long long* seg_stats, const float* msp,
                   float* msp_norm_out, float* mix_out, float lambda, float thr, void* sort_workspace,
                   size_t sort_workspace_bytes, dml_stream_t stream);

DML_API size_t dml_ood_workspace_bytes(int32_t n_seg, int64_t seg_len);
/* Sort `keys` (n_seg segments of seg_len packed keys; clobbered) and evaluate every segment.
 * `seg_stats` is dml_ood_keygen's output (n_pos / n_nan per segment, device).
 * `hist_precomputed` != 0: dml_ood_keygen already left the digit histograms in `workspace`.
 * `results` is a DEVICE array of n_seg dml_ood_result.  recall_level: 0.95 in the reference. */
DML_API int dml_ood_eval_segments(uint32_t* keys, const long long* seg_stats, int32_t n_seg, int64_t seg_len,
                          double recall_level, void* workspace, size_t workspace_bytes, int32_t hist_precomputed,
                          dml_ood_result* results, dml_stream_t stream);

/* Pooled metric over the pairs of MANY dml_ood_keygen / dml_ood_eval_segments calls without generating or counting
 * the keys a second time (the reference pools nothing; BASELINE.json's north star asks for the exact full-set
 * AUROC / FPR95 next to the per-image mean of anomaly/eval_ood_traditional.py:569,641).  Call between
 * dml_ood_keygen(..., sort_workspace = seg_workspace) and dml_ood_eval_segments(..., hist_precomputed = 1) of a
 * batch: the batch's per-segment digit histograms (still raw counts at that point) are added to the histogram slot
 * of `pooled_workspace`, the workspace of ONE segment of `pooled_len` keys (dml_ood_workspace_bytes(1, pooled_len)),
 * and the batch's seg_stats rows (n_pos, n_nan, n_out_of_window, 0) to the 4 int64 of `pooled_stats`;
 * `reset` != 0 overwrites instead of adding (first batch of a pooled evaluation).  `pooled_workspace` may be NULL
 * (counts only: the multi-GPU path re-partitions the keys across ranks) and so may `pooled_stats`.
 * When every batch wrote its keys into consecutive slices of one buffer (all with the same key_base / score_kind),
 * that buffer -- in whatever order the per-batch sorts left it -- is then evaluated with
 *   dml_ood_eval_segments(all_keys, pooled_stats, 1, pooled_len, level, pooled_workspace, bytes, 1, result). */
DML_API int dml_ood_pool_histograms(const void* seg_workspace, size_t seg_workspace_bytes, const long long* seg_stats,
                            int32_t n_seg, int64_t seg_len, void* pooled_workspace, size_t pooled_workspace_bytes,
                            int64_t pooled_len, long long* pooled_stats, int32_t reset, dml_stream_t stream);

/* Second FPR@recall convention, used by the reference's softmax-baseline evaluator
 * (DeepLabV3Plus-Pytorch/test.py:241-244): fpr[tpr >= recall_level][0] on
 * sklearn.metrics.roc_curve(y, score) with its default drop_intermediate=True, i.e. the FIRST kept ROC point
 * whose recall reaches the level (dml_ood_result.fpr is the closest-recall rule of anomaly/anom_utils.py:57-65).
 * Call right after dml_ood_eval_segments with the same keys / seg_stats / workspace (it reads the sorted keys
 * and the per-tile carries the scan left there).  fpr_out: DEVICE [n_seg] float64, NaN for single-class segments. */
DML_API int dml_ood_roc_fpr(const uint32_t* keys, const long long* seg_stats, int32_t n_seg, int64_t seg_len,
                    double recall_level, const void* workspace, size_t workspace_bytes, double* fpr_out,
                    dml_stream_t stream);

/* Building blocks for the multi-GPU pooled metric (locally sorted shards + NCCL exchange on the
 * Python side).  dml_ood_sort: radix sort only, bits [begin_bit, end_bit); `*sorted_out` (host
 * pointer-to-pointer) receives the device buffer holding the result (keys, or inside workspace).
 * dml_ood_scan_range: group scan over an already sorted key range that starts on a score-group
 * boundary.  range_info (device, 4 int64) = (pos_before, idx_before, total_pos, total_n) of the
 * global ranking.  partial_out (device, 80 bytes) = { u64 auroc_num; f64 ap_sum; i64 a_idx, a_tps, a_fps;
 * i64 b_tps, b_idx, b_fps; i64 n_groups; i64 reserved }: a = the last group whose cumulative positives tps
 * stay <= T* (the largest tps with float64 recall tps/P <= recall_level; a_idx = -1: none), b = the smallest
 * tps > T* and the latest group having it (b_tps = INT64_MAX: none).  Ranges combine by
 * (+, +, max idx, (min tps, max idx), +); FPR = fps of whichever of a / b has the smaller
 * |tps/P - recall_level| in float64 (tie -> b), divided by N. */
DML_API int dml_ood_sort(uint32_t* keys, int32_t n_seg, int64_t seg_len, int32_t begin_bit, int32_t end_bit,
                 void* workspace, size_t workspace_bytes, uint32_t** sorted_out, dml_stream_t stream);
/* positions[j] = first index with sorted_keys[i] >= queries[j] (cuts a sorted shard at the range splitters);
 * count = number of keys with the positive bit set.  All pointers device. */
DML_API int dml_ood_lower_bound(const uint32_t* sorted_keys, int64_t n, const uint32_t* queries, int32_t n_queries,
                        long long* positions, dml_stream_t stream);
DML_API int dml_ood_count_positive(const uint32_t* keys, int64_t n, long long* count, dml_stream_t stream);
DML_API int dml_ood_scan_range(const uint32_t* sorted_keys, int64_t n, const long long* range_info,
                       double recall_level, void* workspace, size_t workspace_bytes, void* partial_out,
                       dml_stream_t stream);

/* ------------------------------------------------------------------------------------ *
 * (d') Exact metrics without sorting the negatives ("minority rank" path).
 *
 * Same inputs, outputs and semantics as dml_ood_keygen + dml_ood_eval_segments (anomaly/anom_utils.py:25-78,
 * anomaly/eval_ood_traditional.py:128-148), for the usual case that the positives (OOD pixels) are few: per segment
 * the positives' keys are gathered and sorted in shared memory, and ONE pass over all pairs -- which also writes the
 * normalised conf / MMSP / mix maps and, if `keys_out` != NULL, the packed ranking keys (for a later pooled metric) --
 * locates every negative among the segment's distinct positive scores and counts it; a scan over the positive groups
 * yields AUROC / AUPR / FPR@recall.  AUROC and FPR are bit-identical to the sort path (integer counting), AUPR differs
 * by float64 summation order only; dml_ood_result.n_groups is -1 (negative-only score groups are not enumerated).
 * `pos_capacity` (1 .. 32768): positives per segment this call can hold.  A segment with more is NOT evaluated:
 * its result is NaN and seg_stats[seg][3] = 1 -- evaluate it with dml_ood_keygen + dml_ood_eval_segments instead.
 * seg_stats (device, [n_seg,4] int64) = (n_pos, n_nan, n_out_of_window, overflow flag).
 * workspace: dml_ood_rank_workspace_bytes(n_seg, pos_capacity).  All pointers device. */
DML_API size_t dml_ood_rank_workspace_bytes(int32_t n_seg, int32_t pos_capacity);
DML_API int dml_ood_rank_segments(const float* values, const float* minmax, int32_t minmax_slot, float* conf_out,
                          const uint8_t* gt_u8, const int64_t* gt_i64, uint64_t out_label_mask, const uint8_t* pos_u8,
                          int32_t score_kind, uint32_t key_base, int32_t n_seg, int64_t seg_len, uint32_t* keys_out,
                          long long* seg_stats, const float* msp, float* msp_norm_out, float* mix_out, float lambda,
                          float thr, int32_t pos_capacity, double recall_level, void* workspace, size_t workspace_bytes,
                          dml_ood_result* results, dml_stream_t stream);

/* Appends the positives' score keys (key >> 1) of every segment of the LAST dml_ood_rank_segments call on
 * `rank_workspace` (same n_seg / pos_capacity) to `out` at *count (device, running total, in/out): a pooled evaluation
 * over many batches then needs no pass over all keys to find its positives.  Segments flagged as overflowed export
 * nothing (the caller notices *count < the pooled n_pos and falls back to dml_ood_pos_compact). */
DML_API int dml_ood_rank_export_positives(const void* rank_workspace, size_t workspace_bytes, int32_t n_seg, int32_t pos_capacity,
                                  uint32_t* out, int64_t out_capacity, long long* count, dml_stream_t stream);

/* ------------------------------------------------------------------------------------ *
 * (d'') Pooled metric on the minority-rank idea: building blocks for ONE ranking of up to 2^32 - 1 pairs (the
 * full-set AUROC / AUPR / FPR@95 of the north star; reference semantics anomaly/anom_utils.py:25-78 over all pixels)
 * whose negatives are never sorted and -- across GPUs -- never exchanged:
 *   dml_ood_pos_compact    packed keys (bit 0 = positive) -> score keys (key >> 1) of the positives, any order;
 *                          *count (device) = number of positives (may exceed `capacity`: then only `capacity` were kept)
 *   dml_ood_sort           sort them (bits [0, 31))
 *   dml_ood_unique_counts  sorted score keys -> distinct values + multiplicities, *n_unique (device)
 *   dml_ood_bucket_rank    counters (device, 2 * n_groups + 2 uint64, zeroed by the call):
 *                          [2g] = negatives strictly between group g-1 and g, [2g + 1] = negatives tied with group g,
 *                          [2 * n_groups] = negatives above every positive.  Negatives are grouped into buckets of
 *                          <= 12288 consecutive positive groups by one counting + one non-stable scatter pass, then
 *                          located in shared memory.  DML_ERR_UNSUPPORTED_DIM: more than 4096 * 12288 distinct positive
 *                          scores (use the sort path).  Counters of several key sets (ranks) simply add.
 *   dml_ood_pooled_scan    scan over the groups -> *result (device); n_groups == 0 / one class empty -> NaN.
 * The host reads *count and *n_unique between the calls (two small synchronisations per pooled evaluation). */
DML_API int dml_ood_pos_compact(const uint32_t* keys, int64_t n, uint32_t* pos_keys_out, int64_t capacity, long long* count,
                        dml_stream_t stream);
/* Concatenates the key lists of `n_slots` fixed-capacity slot records, in slot order, into `out` (device):
 * record i starts at slots + i * stride_words (u32 words, 8-byte aligned, stride even); its words 0..1 hold the int64
 * number of keys (clamped to [0, slot_capacity]), its keys start at word `hdr_words`.  The records are the per-batch
 * positives' lists that the ranks all-gather while the per-image pass is still running (the reference has no
 * counterpart: anomaly/eval_ood_traditional.py:128-148 evaluates image by image on one device); the host reads the
 * headers once to size `out`.  Nothing is written when the lists exceed out_capacity. */
DML_API int dml_ood_slots_gather(const uint32_t* slots, int32_t n_slots, int64_t stride_words, int32_t hdr_words,
                         int64_t slot_capacity, uint32_t* out, int64_t out_capacity, dml_stream_t stream);
DML_API size_t dml_ood_unique_workspace_bytes(int64_t n);
DML_API int dml_ood_unique_counts(const uint32_t* sorted_keys, int64_t n, uint32_t* values_out, uint32_t* counts_out,
                          long long* n_unique, void* workspace, size_t workspace_bytes, dml_stream_t stream);
DML_API size_t dml_ood_bucket_rank_workspace_bytes(int64_t n, int64_t n_groups);
DML_API int dml_ood_bucket_rank(const uint32_t* keys, int64_t n, const uint32_t* group_scores, int64_t n_groups, uint32_t key_base,
                        unsigned long long* counters, void* workspace, size_t workspace_bytes, dml_stream_t stream);
DML_API size_t dml_ood_pooled_scan_workspace_bytes(int64_t n_groups);
DML_API int dml_ood_pooled_scan(const uint32_t* group_counts, const unsigned long long* counters, int64_t n_groups,
                        int64_t total_pos, int64_t total_n, int64_t n_nan, double recall_level, void* workspace,
                        size_t workspace_bytes, dml_ood_result* result, dml_stream_t stream);

/* "partition" exchange mode of the multi-GPU pooled metric: scatter UNSORTED packed keys into n_buckets
 * contiguous key ranges (one per rank), ship range r to rank r, sort only what is received -- one partition
 * pass instead of one full local sort.  bucket(key) = #{ j : key >= bounds[j] } with `bounds` a DEVICE
 * array of n_buckets-1 ascending keys; out = [bucket 0 | bucket 1 | ...] (order inside a bucket
 * unspecified), counts (device, [n_buckets] int64) = exact bucket sizes.  1 <= n_buckets <= DML_MAX_PARTITIONS. */
#define DML_MAX_PARTITIONS 16
DML_API size_t dml_ood_partition_workspace_bytes(int32_t n_buckets);
DML_API int dml_ood_partition(const uint32_t* keys, int64_t n, const uint32_t* bounds, int32_t n_buckets, uint32_t* out,
                      long long* counts, void* workspace, size_t workspace_bytes, dml_stream_t stream);

/* ------------------------------------------------------------------------------------ *
 * (f-4) Input side of the multi-scale evaluation: ValDataset's resize + normalisation loop
 *   anomaly/dataset.py:11-21,281-297  imresize(img, (w, h), 'bilinear') = PIL.Image.resize(.., BILINEAR) per scale
 *   anomaly/dataset.py:65-70          img_transform: float32 / 255, HWC -> CHW, Normalize(mean, std)
 * Bit-exact with Pillow's 8-bit two-pass resampling (integer coefficients, uint8 intermediate) followed by the
 * reference's float32 arithmetic ((v / 255) - mean) / std.
 *   dml_resize_ksize      taps per output index of one axis (host)
 *   dml_resize_coeffs     host: bounds[2 i] = first tap, bounds[2 i + 1] = tap count, coeffs[i * ksize + t] = weight 2^22
 *   dml_resize_bilinear_normalize   device: image uint8 [B, H, W, 3] -> out float32 [B, 3, out_h, out_w]; the four
 *                         tables are DEVICE copies of dml_resize_coeffs' output; mean / stdv are HOST float[3];
 *                         max_rows_per_tile >= the input rows any 8 consecutive output rows touch
 *                         (max over i of bounds_y[2 min(i + 7, out_h - 1)] + count - bounds_y[2 i]).
 * ------------------------------------------------------------------------------------ */
DML_API int32_t dml_resize_ksize(int32_t in_size, int32_t out_size);
DML_API int dml_resize_coeffs(int32_t in_size, int32_t out_size, int32_t* bounds, int32_t* coeffs, int32_t ksize);
DML_API int dml_resize_bilinear_normalize(const uint8_t* image, int32_t B, int32_t H, int32_t W, const int32_t* bounds_x,
                                  const int32_t* coeffs_x, int32_t ksize_x, const int32_t* bounds_y, const int32_t* coeffs_y,
                                  int32_t ksize_y, int32_t out_h, int32_t out_w, int32_t max_rows_per_tile, const float* mean,
                                  const float* stdv, float* out, dml_stream_t stream);

/* ------------------------------------------------------------------------------------ *
 * (f-4) Synchronised batch normalisation, one process per GPU (NCHW fp32):
 *   anomaly/lib/nn/modules/batchnorm.py:57-88,121-139 (_SynchronizedBatchNorm.forward / _compute_mean_std)
 *   dml_bn_stats           forward (dy = mean = NULL): sums[c] = sum x, sums[C + c] = sum x^2 over (B, HW);
 *                          backward: sums[c] = sum dy, sums[C + c] = sum dy (x - mean[c]).  Device doubles [2C]; the caller
 *                          all-reduces them across ranks (NCCL) before the next call.
 *   dml_bn_finalize        all-reduced forward sums + global element count (*count: DEVICE double, all-reduced with the
 *                          sums, so no host synchronisation sits between the kernels; must exceed 1) -> mean, inv_std = clamp(var, eps)^-1/2,
 *                          clamped[c] = 1 where the clamp is active; with tmp_running_mean != NULL also the reference's
 *                          moving averages (_tmp_running_mean / _tmp_running_var / _running_iter -> running_mean / running_var)
 *   dml_bn_apply           y = (x - mean) (inv_std weight) + bias       (weight / bias may be NULL)
 *   dml_bn_backward_apply  dx from the all-reduced backward sums and the global count
 * ------------------------------------------------------------------------------------ */
DML_API size_t dml_bn_workspace_bytes(int32_t B, int32_t C, int64_t HW);
DML_API int dml_bn_stats(const float* x, const float* dy, const float* mean, int32_t B, int32_t C, int64_t HW, double* sums,
                 void* workspace, size_t workspace_bytes, dml_stream_t stream);
DML_API int dml_bn_finalize(const double* sums, const double* count, float eps, float momentum, int32_t C, float* tmp_running_mean,
                    float* tmp_running_var, float* running_iter, float* running_mean, float* running_var, float* mean,
                    float* inv_std, uint8_t* clamped, dml_stream_t stream);
DML_API int dml_bn_apply(const float* x, const float* mean, const float* inv_std, const float* weight, const float* bias, int32_t B,
                 int32_t C, int64_t HW, float* y, dml_stream_t stream);
DML_API int dml_bn_backward_apply(const float* x, const float* dy, const float* mean, const float* inv_std, const float* weight,
                          const uint8_t* clamped, const double* sums, const double* count, int32_t B, int32_t C, int64_t HW, float* dx,
                          dml_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DML_B200_H_ */

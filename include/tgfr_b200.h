/*
 * tgfr_b200.h -- C ABI of libtgfr_b200.so: the FCAM contrastive-loss / margin-head hot path of
 * Mahedi-61/Text_Guided_Face_Recognition, written for NVIDIA B200 (sm_100a).
 *
 * The reference has no FFI layer: its boundary is Python symbols.  Every
 * entry point below replaces the device work of one of those symbols and is what the Python
 * mirror in text_guided_face_recognition_b200/models/ binds with ctypes.  Conventions:
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the name ends in _host;
 *   - float tensors are fp32, label/class-id tensors int64, caption lengths int32;
 *   - strides are in ELEMENTS; "canonical" layouts are ctx [Bc,R,D], words [Bq,T,D];
 *   - `stream` is a cudaStream_t passed as void* (the caller's current stream);
 *   - every function returns 0 on success, a negative TGFR_E_* code on failure, and never
 *     falls back to a CPU path; tgfr_last_error() returns a static description.
 *   - functions are re-entrant per device (no global mutable state besides the error string).
 */
#ifndef TGFR_B200_H_
#define TGFR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGFR_OK 0
#define TGFR_E_INVALID (-1)   /* bad argument / unsupported shape */
#define TGFR_E_CUDA (-2)      /* CUDA runtime error (see tgfr_last_error) */
#define TGFR_E_WORKSPACE (-3) /* workspace too small */
#define TGFR_E_ARCH (-4)      /* device is not sm_100 */

/* precision selector of the word-region kernels and of the margin-head contractions */
#define TGFR_PREC_FP32 0      /* SIMT fp32 everywhere (bit-for-bit deterministic forward) */
#define TGFR_PREC_TC 1        /* tcgen05 tensor-core contractions (fp16 operands), fp32 accumulate/statistics */

int tgfr_version(void);
const char* tgfr_last_error(void);
/* 0 if the current device can run this library (compute capability 10.x), else TGFR_E_ARCH. */
int tgfr_device_check(void);

/* ------------------------------------------------------------------------------------------
 * Word-region attention loss.  Replaces models/attention.py:10-43 (func_attention) as called
 * B times from models/losses.py:73-114 (words_loss): all Bc x Bq (face, caption) pairs at once.
 *   sim[b, i] = gamma3 * log sum_t exp(gamma2 * cos(q_it, W_bit))           (losses.py:104-122)
 * cap_lens: int32 [Bq] (LSTM path, losses.py:82) or NULL (all captions use T words).
 * attn_diag: optional [Bc, T, R]; row b receives A2 of pair (b, b + diag_off) (losses.py:97),
 *            zero-filled beyond the caption's length; NULL to skip.  TGFR_PREC_TC emits it from the tensor-core
 *            forward (fp16-operand scores: within 2e-3 of a map's largest entry); with the environment variable
 *            TGFR_ATTN_MAPS=fp32 it comes from the exact fp32 kernel.
 * ------------------------------------------------------------------------------------------ */
int tgfr_wordregion_fwd(const float* ctx, int64_t ctx_sb, int64_t ctx_sr, int64_t ctx_sd,
                        const float* words, int64_t w_sb, int64_t w_st, int64_t w_sd,
                        const int32_t* cap_lens, int Bc, int Bq, int T, int R, int D,
                        float gamma1, float gamma2, float gamma3, float eps,
                        float* sim, float* attn_diag, int diag_off,
                        int precision, void* workspace, size_t workspace_bytes,
                        void* saved, size_t saved_bytes, void* stream);

/* Gradients of sum(gsim * sim): dctx [Bc,R,D] and dwords [Bq,T,D] (contiguous, OVERWRITTEN;
 * either may be NULL to skip it -- the reference's training scripts only need dctx because the
 * text side is detached, utils/dataset_utils.py:42-46).  With the forward's `saved` buffer the face-side gradient is
 * computed from what the forward left there; without it (and for dwords) the attention is recomputed on chip. */
int tgfr_wordregion_bwd(const float* ctx, int64_t ctx_sb, int64_t ctx_sr, int64_t ctx_sd,
                        const float* words, int64_t w_sb, int64_t w_st, int64_t w_sd,
                        const int32_t* cap_lens, int Bc, int Bq, int T, int R, int D,
                        float gamma1, float gamma2, float gamma3, float eps,
                        const float* gsim, float* dctx, float* dwords,
                        int precision, void* workspace, size_t workspace_bytes,
                        const void* saved, size_t saved_bytes, void* stream);

/* Bytes of scratch the two calls above need for the given shape/precision (0 is possible). */
size_t tgfr_wordregion_workspace_bytes(int Bc, int Bq, int T, int R, int D, int precision);
/* Optional forward -> backward buffer (TGFR_PREC_TC): when `saved` (tgfr_wordregion_saved_bytes bytes, 256-byte
 * aligned) is given to the forward call it stores, per pair, either the fp16 word-softmax / attention records (default:
 * the backward then runs no score GEMM and no exponential) or, with the environment variable TGFR_WORDREGION_SAVE=wu
 * set when the size is queried, only the fp16 attended-word tiles (2.5x fewer bytes, nothing per (caption, word, region); the backward recomputes the
 * scores).
 * The two layouts differ in size and both calls recognise the layout by `saved_bytes`, so pass exactly what
 * tgfr_wordregion_saved_bytes returned.  saved = NULL on either side selects the fully recomputing (memory-lean) path;
 * results agree to fp16 rounding. */
size_t tgfr_wordregion_saved_bytes(int Bc, int Bq, int T, int R, int D, int precision);

/* Stand-alone func_attention(query, context, gamma1) for B independent (query, context) pairs
 * (models/attention.py:10-43): wc [B,T,D] canonical, attn [B,T,R]. */
int tgfr_attention_fwd(const float* ctx, int64_t ctx_sb, int64_t ctx_sr, int64_t ctx_sd,
                       const float* query, int64_t q_sb, int64_t q_st, int64_t q_sd,
                       int B, int T, int R, int D, float gamma1,
                       float* wc, float* attn, void* stream);
/* Backward of the above for upstream gradients g_wc [B,T,D] and g_attn [B,T,R] (either may be
 * NULL): dctx [B,R,D], dquery [B,T,D], both contiguous and overwritten. */
int tgfr_attention_bwd(const float* ctx, int64_t ctx_sb, int64_t ctx_sr, int64_t ctx_sd,
                       const float* query, int64_t q_sb, int64_t q_st, int64_t q_sd,
                       int B, int T, int R, int D, float gamma1,
                       const float* g_wc, const float* g_attn,
                       float* dctx, float* dquery, void* stream);

/* ------------------------------------------------------------------------------------------
 * Cosine score matrix.  Replaces models/losses.py:38-48 (sent_loss), :338-343 (global_loss) and
 * :292-294 (ClipLoss.get_logits).
 *   normalise != 0: scores[i,j] = scale * <x_i,y_j> / max(|x_i||y_j|, eps);  else scale*<x_i,y_j>.
 * class_ids_x/_y (int64, may be NULL): entries with class_ids_x[i]==class_ids_y[j] and
 * (i + diag_off) != j are set to -inf (losses.py:20-30, 47-48).  xnorm/ynorm: outputs [Bx]/[By]
 * kept for the backward pass.
 * ------------------------------------------------------------------------------------------ */
int tgfr_cosine_scores_fwd(const float* x, int64_t x_sr, const float* y, int64_t y_sr,
                           int Bx, int By, int D, float scale, int normalise, float eps,
                           const int64_t* class_ids_x, const int64_t* class_ids_y, int diag_off,
                           float* scores, float* xnorm, float* ynorm, void* stream);
/* dx [Bx,D], dy [By,D] (contiguous, overwritten; either may be NULL) from gscores [Bx,By]. */
int tgfr_cosine_scores_bwd(const float* x, int64_t x_sr, const float* y, int64_t y_sr,
                           int Bx, int By, int D, float scale, int normalise, float eps,
                           const float* xnorm, const float* ynorm, const float* gscores,
                           float* dx, float* dy, void* workspace, size_t workspace_bytes, void* stream);
size_t tgfr_cosine_workspace_bytes(int Bx, int By, int D);

/* ------------------------------------------------------------------------------------------
 * Two-direction cross entropy over a [Bx,By] score block whose row b is paired with column
 * b + diag_off (nn.CrossEntropyLoss x2 at losses.py:49-53, 128-132, 348-350).
 * stats layout (floats): rowlse[Bx] | colmax[By] | colsum[By] | diag[Bx].  A rank that holds a
 * row block of the global matrix all-gathers colmax/colsum and combines them before _finish.
 * ------------------------------------------------------------------------------------------ */
int tgfr_pair_ce_stats(const float* scores, int Bx, int By, int diag_off,
                       float* rowlse, float* colmax, float* colsum, float* diag, void* stream);
/* losses[0] = sum_b (rowlse[b] - diag[b]) * inv_b ; collse[j] = colmax[j] + log(colsum[j]);
 * losses[1] = sum over this rank's diagonal columns of (collse[b+diag_off] - diag[b]) * inv_b. */
int tgfr_pair_ce_finish(const float* rowlse, const float* colmax, const float* colsum,
                        const float* diag, int Bx, int By, int diag_off, float inv_b,
                        float* losses, float* collse, void* stream);
/* gscores = (g0 * (softmax_row - I) + g1 * (softmax_col - I)) * inv_b; g0/g1 are DEVICE scalars. */
int tgfr_pair_ce_bwd(const float* scores, const float* rowlse, const float* collse,
                     const float* g0, const float* g1, int Bx, int By, int diag_off, float inv_b,
                     float* gscores, void* stream);

/* ------------------------------------------------------------------------------------------
 * Margin heads.  cos = normalize(x) . normalize(w_c): models/metrics.py:44, models/magface.py:92-94.
 * w_sc / w_sk: element strides of weight along the class / feature axis
 *   (ArcMarginProduct.weight [C,Din]: w_sc=Din, w_sk=1;  MagLinear.weight [Din,C]: w_sc=1, w_sk=C).
 * class_off: first class id owned by this rank (class-sharded partial FC); labels are global.
 * ------------------------------------------------------------------------------------------ */
/* out[b,c] = s * cos[b,c]  (clamp_cos != 0: cos clamped to [-1,1] first, magface.py:94);
 * xnorm[B], wnorm[C] receive the raw L2 norms.  precision = TGFR_PREC_TC runs the contraction on the
 * tensor cores and needs tgfr_margin_workspace_bytes(B, C, Din, TGFR_PREC_TC) bytes of 256-byte aligned
 * workspace (TGFR_PREC_FP32: workspace may be NULL). */
int tgfr_cos_logits_fwd(const float* x, int64_t x_sr, const float* w, int64_t w_sc, int64_t w_sk,
                        int B, int C, int Din, float s, int clamp_cos,
                        float* out, int64_t out_sr, float* xnorm, float* wnorm,
                        int precision, void* workspace, size_t workspace_bytes, void* stream);
/* ArcFace margin on the label column, in place (metrics.py:45-57); cos_t[B] keeps the target
 * cosine (NaN if the label is not owned by this rank). */
int tgfr_arc_margin_apply(float* logits, int64_t sr, const int64_t* labels, int B, int C,
                          int class_off, float s, float m, int easy_margin, float* cos_t, void* stream);
/* dx [B,Din], dw (same strides as w) from glogits [B,C]; dx may be NULL. dx/dw overwritten. */
int tgfr_arc_margin_bwd(const float* x, int64_t x_sr, const float* w, int64_t w_sc, int64_t w_sk,
                        const float* xnorm, const float* wnorm, const int64_t* labels,
                        const float* cos_t, const float* glogits, int64_t g_sr,
                        int B, int C, int Din, int class_off, float s, float m, int easy_margin,
                        float* dx, float* dw, int precision, void* workspace, size_t workspace_bytes, void* stream);
size_t tgfr_margin_workspace_bytes(int B, int C, int Din, int precision);

/* Fused ArcFace + cross entropy (metrics.py:42-60 followed by nn.CrossEntropyLoss / FocalLoss, losses.py:313-325)
 * on the tensor cores: the [B,C] logits are never written.  The forward leaves, for THIS class shard
 * (columns class_off .. class_off + C - 1), the online-softmax statistics rowmax[B], rowsum[B] (sum of
 * exp(logit - rowmax)), the target logit tgt[B] (0 when the label lives on another shard) and cos_t[B] (NaN then);
 * the caller merges shards (all-reduce), and tgfr_focal_finish turns (rowmax, rowsum, tgt) into the loss, the
 * per-row log-sum-exp `lse` and the focal factor.  The backward takes lse, the focal factor `coef` and the upstream
 * gradient `gout` (device scalars, either may be NULL = 1) and returns dx [B,Din] (may be NULL) and dw (w's strides).
 * `saved` (tgfr_arc_fused_saved_bytes) carries the fp16 normalised operands from the forward to the backward;
 * workspace: tgfr_arc_fused_workspace_bytes; both 256-byte aligned. */
size_t tgfr_arc_fused_workspace_bytes(int B, int C, int Din);
size_t tgfr_arc_fused_saved_bytes(int B, int C, int Din);
int tgfr_arc_fused_fwd(const float* x, int64_t x_sr, const float* w, int64_t w_sc, int64_t w_sk,
                       const int64_t* labels, int B, int C, int Din, int class_off, float s, float m, int easy_margin,
                       float* xnorm, float* wnorm, float* rowmax, float* rowsum, float* tgt, float* cos_t,
                       void* workspace, size_t workspace_bytes, void* saved, size_t saved_bytes, void* stream);
int tgfr_arc_fused_bwd(const float* x, int64_t x_sr, const float* w, int64_t w_sc, int64_t w_sk,
                       const int64_t* labels, const float* xnorm, const float* wnorm, const float* lse,
                       const float* coef, const float* gout, int B, int C, int Din, int class_off, float s, float m,
                       int easy_margin, float* dx, float* dw, void* workspace, size_t workspace_bytes,
                       const void* saved, size_t saved_bytes, void* stream);

/* MagFace: cos_m[b,c] from cos_s (= scale*cos) and per-row margins (magface.py:95-106). */
int tgfr_mag_margin_fwd(const float* cos_s, const float* margin, int B, int C, float scale,
                        int easy_margin, float* cos_m_s, void* stream);
/* Given upstream g_cos, g_cosm (either may be NULL): gtotal[b,c] (gradient w.r.t. scale*cos) and
 * gmargin[b] (gradient w.r.t. the per-row margin). */
int tgfr_mag_margin_bwd(const float* cos_s, const float* margin, const float* g_cos, const float* g_cosm,
                        int B, int C, float scale, int easy_margin, float* gtotal, float* gmargin,
                        void* stream);
/* Backward of tgfr_cos_logits_fwd alone (no label margin): dx, dw from gout [B,C]. */
int tgfr_cos_logits_bwd(const float* x, int64_t x_sr, const float* w, int64_t w_sc, int64_t w_sk,
                        const float* xnorm, const float* wnorm, const float* out, int64_t out_sr,
                        const float* gout, int64_t g_sr, int B, int C, int Din, float s, int clamp_cos,
                        float* dx, float* dw, int precision, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row-wise cross entropy over dense logits [B,C] (nn.CrossEntropyLoss inside FocalLoss,
 * losses.py:319-325, and F.cross_entropy at magface.py:135).  Online softmax, one pass.
 * rowmax/rowsum: per-row max and sum exp(l - max) of THIS rank's class shard; tgt[b] = logit of
 * the label column (0 if not owned).
 * ------------------------------------------------------------------------------------------ */
int tgfr_ce_rows_stats(const float* logits, int64_t sr, const int64_t* labels, int B, int C,
                       int class_off, float* rowmax, float* rowsum, float* tgt, void* stream);
/* out[0] = logp = mean_b(rowmax + log rowsum - tgt); out[1] = focal loss (1-exp(-logp))^gamma*logp;
 * out[2] = d focal / d logp.  lse[b] = rowmax + log(rowsum). */
int tgfr_focal_finish(const float* rowmax, const float* rowsum, const float* tgt, int B, float gamma,
                      float* out, float* lse, void* stream);
/* glogits[b,c] = coef * gout * (exp(l - lse[b]) - [c == label_b]) / B; coef, gout DEVICE scalars
 * (either may be NULL = 1). */
int tgfr_ce_rows_bwd(const float* logits, int64_t sr, const int64_t* labels, const float* lse,
                     const float* coef, const float* gout, int B, int C, int class_off,
                     float* glogits, int64_t g_sr, void* stream);
/* cosine_similarity(x1, x2, dim=1, eps) of models/losses.py:12-16 for [N, D] rows (element strides sr / sd):
 * out[n] = sum(x1 x2) / max(|x1| |x2|, eps); stats [N, 3] = (sum(x1 x2), |x1|, |x2|) is what the backward reads.
 * dx1 / dx2 are contiguous [N, D] (either may be NULL). */
int tgfr_cosine_rows_fwd(const float* x1, int64_t x1_sr, int64_t x1_sd, const float* x2, int64_t x2_sr,
                         int64_t x2_sd, int64_t N, int D, float eps, float* out, float* stats, void* stream);
int tgfr_cosine_rows_bwd(const float* x1, int64_t x1_sr, int64_t x1_sd, const float* x2, int64_t x2_sr,
                         int64_t x2_sd, int64_t N, int D, float eps, const float* stats, const float* gout,
                         float* dx1, float* dx2, void* stream);
/* Cross-shard merge of online-softmax statistics after one all-gather: gathered [n][K][M] holds
 * every shard's (max, sum exp(. - max)[, target logit]) rows (K = 2 or 3); out [K][M] = (global max, rescaled
 * sum[, sum of the target logits -- non-owners hold 0]). */
int tgfr_merge_softmax_stats(const float* gathered, int n, int K, int M, float* out, void* stream);
/* MagLoss (models/magface.py:131-135): the blended logits output[b,c] = (c == label_b) ? cos_m[b,c] : cos_s[b,c]
 * are never materialised -- the online-softmax row statistics read the two logit tensors (row stride sr) directly;
 * one_hot [B,C] (the tensor the reference returns; NULL = skip) is written in the same pass.  Finish with
 * tgfr_focal_finish(gamma = 0).  The backward writes both dense gradients: g_cos is zero on the label column,
 * g_cosm is zero off it (contiguous [B,C]); gout is a DEVICE scalar (NULL = 1). */
int tgfr_mag_ce_stats(const float* cos_s, const float* cos_m, int64_t sr, const int64_t* labels, int B, int C,
                      float* rowmax, float* rowsum, float* tgt, float* one_hot, void* stream);
int tgfr_mag_ce_bwd(const float* cos_s, const float* cos_m, int64_t sr, const int64_t* labels, const float* lse,
                    const float* gout, int B, int C, float* g_cos, float* g_cosm, void* stream);

/* ------------------------------------------------------------------------------------------
 * TextHeading (models/models.py:170-232): BERT tokens [B, L = bert_words_num - 1, E]
 * (contiguous fp32) -> words [B, T = bert_words_num - 2, F] (unit rows, the layout tgfr_wordregion_* reads; the
 * reference returns its transpose [B, F, T]) and sent [B, F].  w2 / w3 / w4: the Conv2d(1, F, (K, E)) weights
 * [F, 1, K, E] as contiguous [F, K*E]; b2 / b3 / b4: their biases (NULL = none).  `saved`
 * (tgfr_texthead_saved_bytes) carries the three ReLU outputs to the backward, which returns the weight / bias
 * gradients for upstream gwords [B, T, F] and gsent [B, F] (either may be NULL); the tokens get no gradient
 * (frozen BERT).  The last word is detached exactly as in the reference (models.py:206).  The six products run on
 * tcgen05 as error-compensated fp16 hi/lo splits of the fp32 operands (three accumulated terms, ~22 significant bits,
 * fp32 accumulation) when E % 8 == 0 and F % 8 == 0, else (or with TGFR_TEXTHEAD_PRECISION=fp32) on the fp32 SIMT GEMM.
 * ------------------------------------------------------------------------------------------ */
size_t tgfr_texthead_saved_bytes(int B, int L, int E, int F);
size_t tgfr_texthead_workspace_bytes(int B, int L, int E, int F);
int tgfr_texthead_fwd(const float* tokens, const float* w2, const float* w3, const float* w4,
                      const float* b2, const float* b3, const float* b4, int B, int L, int E, int F,
                      int bert_words_num, float* words, float* sent, void* saved, size_t saved_bytes, void* stream);
int tgfr_texthead_bwd(const float* tokens, const float* gwords, const float* gsent, int B, int L, int E, int F,
                      int bert_words_num, float* dw2, float* dw3, float* dw4, float* db2, float* db3, float* db4,
                      void* workspace, size_t workspace_bytes, const void* saved, size_t saved_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Verification / identification scoring (utils/modules.py:40-88,150-166).
 * tgfr_pair_cosine: scores[i] = sum_k (x1[i,k] / max(|x1_i|, eps)) (x2[i,k] / max(|x2_i|, eps)), i.e.
 *   nn.CosineSimilarity(dim=1, eps=1e-6)(out1, out2) of utils/modules.py:150-151; strides in elements.
 * tgfr_roc_curve: the integer part of sklearn.metrics.roc_curve(y_true, y_score) (utils/modules.py:54; scikit-learn
 *   1.9.0 _ranking.py roc_curve / confusion_matrix_at_thresholds): scores sorted descending, one point per distinct
 *   score with tps = positives (label == 1) at or above it and fps = 1 + index - tps, then (drop_intermediate != 0)
 *   only the end points and the points whose second difference of fps or tps is non-zero.  thresholds / fps / tps
 *   need room for N entries; counts = int64[3] on the device: points written, 1 if a score was NaN, distinct scores.
 *   The caller prepends the (inf, 0, 0) point and divides by the last entries (host side, float64, as sklearn does).
 *   Bit-exact with scikit-learn for the same fp32 scores.  N < 2^31.
 * tgfr_row_argmax: index[r] = first position of the maximum of scores[r, :] (np.argmax, utils/modules.py:84-85).
 * ------------------------------------------------------------------------------------------ */
int tgfr_pair_cosine(const float* x1, int64_t x1_sr, int64_t x1_sd, const float* x2, int64_t x2_sr, int64_t x2_sd,
                     int64_t N, int D, float eps, float* scores, void* stream);
size_t tgfr_roc_workspace_bytes(int64_t N);
int tgfr_roc_curve(const float* scores, const int64_t* labels, int64_t N, int drop_intermediate, float* thresholds,
                   int64_t* fps, int64_t* tps, int64_t* counts, void* workspace, size_t workspace_bytes, void* stream);
int tgfr_row_argmax(const float* scores, int64_t sr, int rows, int cols, int64_t* index, void* stream);

/* ------------------------------------------------------------------------------------------
 * FCFM fusion net `Working`, eval-mode forward (models/fusion_nets.py:217-258): the fused
 * 640-d embedding [ Linear(maxpool(LayerNorm(SelfAttention(image, text)))) | LayerNorm(gl_img) | LayerNorm(sent) ]
 * that tgfr_pair_cosine scores.  img logical [B,256,14,14], word logical [B,256,T] (1 <= T <= 64), both through
 * element strides; gl_img / sent [B,256] (row stride, unit element stride); out [B,640].  params_host: HOST array of
 * tgfr_fcfm_working_num_params() = 26 device pointers, the module's state_dict in this order: conv.weight, conv.bias,
 * bn_img.{weight,bias,running_mean,running_var}, projection.{weight,bias}, bn_word.{weight,bias,running_mean,
 * running_var}, sa.query_proj.{weight,bias}, sa.key_proj.{weight,bias}, sa.value_proj.{weight,bias}, ln.{weight,bias},
 * linear.{weight,bias}, ln_gl_image.{weight,bias}, ln_sent.{weight,bias} (all contiguous fp32).  BatchNorm uses the
 * running statistics (the evaluation path of utils/modules.py:141-147); training mode: tgfr_fcfm_train_fwd / _bwd below.
 * ------------------------------------------------------------------------------------------ */
int tgfr_fcfm_working_num_params(void);
int tgfr_fcfm_working_fwd(const float* img, int64_t img_sb, int64_t img_sc, int64_t img_sh, int64_t img_sw, const float* word,
                          int64_t word_sb, int64_t word_sd, int64_t word_st, const float* gl_img, int64_t gl_sr,
                          const float* sent, int64_t sent_sr, const float* const* params_host, int n_params, int B, int T,
                          float* out, int64_t out_sr, void* stream);
/* The same forward with the 3x3 convolution (fusion_nets.py:235, 95 % of the arithmetic) on the tensor cores: one
 * implicit GEMM per 4096 samples over fp16 hi / lo copies of the image (three split terms, fp32-class accuracy), the
 * rest in the per-sample kernel.  For the large batches of verification (src/test.py: every pair of the list in one
 * call).  workspace: tgfr_fcfm_working_workspace_bytes(B) bytes, 256-byte aligned; same results within 1e-5. */
size_t tgfr_fcfm_working_workspace_bytes(int B);
int tgfr_fcfm_working_fwd_tc(const float* img, int64_t img_sb, int64_t img_sc, int64_t img_sh, int64_t img_sw,
                             const float* word, int64_t word_sb, int64_t word_sd, int64_t word_st, const float* gl_img,
                             int64_t gl_sr, const float* sent, int64_t sent_sr, const float* const* params_host,
                             int n_params, int B, int T, float* out, int64_t out_sr, void* workspace,
                             size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Self-tests of the tcgen05 / TMA building blocks (used by tests/test_gpu_tc.py only).
 * tgfr_debug_umma: out[128,N] = A * B^T on one CTA with fp16 operands a (a_mn ? [K,128] : [128,K])
 * and b (b_mn ? [K,N] : [N,K]); manual_a stages A with the hand-written 128B swizzle.
 * tgfr_debug_tma_reduce: out[r,c] += 1000 r + c through a swizzled tile and a TMA reduce-add.
 * ------------------------------------------------------------------------------------------ */
int tgfr_debug_umma(const void* a, const void* b, float* out, int N, int K, int a_mn, int b_mn, int manual_a,
                    void* stream);
int tgfr_debug_tma_reduce(float* out, int rows, int cols, void* stream);
/* CTA pair (cta_group::2) probe: out[256,N] = A[256,K] * B[N,K]^T on a cluster of two CTAs (M = 256 MMAs issued by
 * the leader, each CTA holding its 128 rows of A, half of B and its half of D in TMEM). */
int tgfr_debug_umma_2cta(const void* a, const void* b, float* out, int N, int K, void* stream);
/* Phase trace of the tensor-core word-region kernels: dev_buf = int64[16*32] (or NULL to switch
 * it off); CTA 0 stamps clock64() per pipeline phase of its first 16 units (tools/trace_wordregion.py). */
int tgfr_debug_set_trace(void* dev_buf);

/* ------------------------------------------------------------------------------------------
 * IMIM, the local branch of ImageHeading (models/models.py:380-405; SelfAttention models/fusion_nets.py:82-118;
 * SURVEY.md 8(f) row f3): x [B,256,P = 14*14] (element strides x_sb / x_sc / x_sp) -> out [B, P, 256] contiguous =
 * the reference's result in its memory order (logical [B,256,14,14], unit L2 norm over the channels at every
 * position) -- the layout tgfr_wordregion_* reads.  params: 16 device pointers in the order
 *   bn_img.weight, bn_img.bias, sa.query_proj.weight [256,256], .bias, sa.key_proj.weight, .bias, sa.value_proj.weight,
 *   .bias, ln.weight [256*P], ln.bias, conv1x1_1.weight [128,256], .bias, conv1x1_2.weight [256,128], .bias,
 *   project_local.projection.weight [256,256], .bias          (the reference's state_dict tensors, contiguous fp32).
 * training != 0: BatchNorm uses the batch statistics and updates running_mean / running_var (momentum, unbiased
 * variance; either may be NULL); else it normalises with them.  `saved` (tgfr_imim_saved_bytes) carries the
 * activations to tgfr_imim_bwd, which takes gout [B,P,256] and the forward's `out`, and writes dparams (16 device
 * pointers, same order and shapes) and, if not NULL, dx [B,256,P] contiguous. */
size_t tgfr_imim_saved_bytes(int B, int P);
size_t tgfr_imim_workspace_bytes(int B, int P);
int tgfr_imim_num_params(void);
int tgfr_imim_fwd(const float* x, int64_t x_sb, int64_t x_sc, int64_t x_sp, const void* const* params, int n_params,
                  int B, int P, int training, float momentum, float eps, float* running_mean, float* running_var,
                  float* out, void* saved, size_t saved_bytes, void* stream);
int tgfr_imim_bwd(const float* gout, const float* out, const float* x, int64_t x_sb, int64_t x_sc, int64_t x_sp,
                  const void* const* params, int n_params, int B, int P, int training, const void* saved,
                  size_t saved_bytes, void* const* dparams, float* dx, void* workspace, size_t workspace_bytes,
                  void* stream);

/* The contraction IMIM's layers run on (csrc/gemm_tc.cu gemm_tc_pair), with fp32 operands: every 1x1 convolution /
 * nn.Linear / torch.bmm of models/models.py:380-405 and models/fusion_nets.py:97-115 is one of these three shapes.
 *   mode 0  C = A B^T  (A [M,K], B [N,K])    mode 1  C = A B  (A [M,K], B [K,N])    mode 2  C = A^T B  (A [K,M], B [K,N])
 *   C [batch, M, ldc] = relu?( alpha * product + bias[col] )
 * Both operands are split into fp16 hi + lo with one power-of-two scale each and the tensor cores accumulate
 * A_hi B_lo + A_lo B_hi + A_hi B_hi in fp32 (nterms = 3, ~22 significant bits) or A_hi B_hi only (nterms = 1).
 * batch > 1: densely packed samples (lda / ldb = the row length), no split-K.  splits > 1: split-K over `splits` CTAs
 * per tile (no bias / relu).  workspace: tgfr_matmul_split_workspace_bytes, 256-byte aligned. */
size_t tgfr_matmul_split_workspace_bytes(int mode, int M, int N, int K, int batch);
int tgfr_matmul_split(int mode, const float* a, int64_t lda, const float* b, int64_t ldb, float* c, int64_t ldc, int M,
                      int N, int K, int batch, float alpha, const float* bias, int relu, int splits, int nterms,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ProjectionHead (models/models.py:96-119), the global branch of ImageHeading: out [M,N] = normalize(x W^T + b) for
 * x [M,K] (row stride x_sr), weight [N,K], bias [N] (NULL = none); znorm [M] = |x W^T + b| for the backward, which
 * writes dweight [N,K], dbias [N] and, if not NULL, dx [M,K]; dz_scratch is [M,N] floats. */
int tgfr_proj_head_fwd(const float* x, int64_t x_sr, const float* weight, const float* bias, int M, int N, int K,
                       float* out, float* znorm, void* stream);
int tgfr_proj_head_bwd(const float* gout, const float* out, const float* znorm, const float* x, int64_t x_sr,
                       const float* weight, int M, int N, int K, float* dz_scratch, float* dx, float* dweight,
                       float* dbias, void* stream);

/* ------------------------------------------------------------------------------------------
 * FCFM fusion net `Working`, TRAINING-mode forward + backward (models/fusion_nets.py:217-258 under autograd, as the
 * fusion training step src/fusion_bert.py:205-233 runs it).  tgfr_fcfm_working_fwd above is the
 * evaluation forward.  img [B,256,14,14] (element strides), word [B,256,T] (strides word_sb / word_sc, T contiguous),
 * gl_img / sent [B,256] (row strides) -> out [B,640] (row stride out_sr).  params: 22 device pointers in the order
 *   conv.weight [36,2304], conv.bias, bn_img.weight, bn_img.bias, projection.weight [36,256], projection.bias,
 *   bn_word.weight, bn_word.bias, sa.query_proj.weight [36,36], .bias, sa.key_proj.weight, .bias, sa.value_proj.weight,
 *   .bias, ln.weight [36*36], ln.bias, linear.weight [128,324], linear.bias, ln_gl_image.weight [256], .bias,
 *   ln_sent.weight [256], .bias;
 * running_stats: {bn_img.running_mean, bn_img.running_var, bn_word.running_mean, bn_word.running_var} (entries may be
 * NULL when training; updated with `momentum` then).  The backward writes dparams (22 pointers, same order / shapes) and,
 * where not NULL, dimg [B,256,14,14], dword [B,256,T], dgl_img / dsent [B,256] (all contiguous). */
size_t tgfr_fcfm_train_saved_bytes(int B, int T);
size_t tgfr_fcfm_train_workspace_bytes(int B, int T);
int tgfr_fcfm_train_fwd(const float* img, int64_t img_sb, int64_t img_sc, int64_t img_sh, int64_t img_sw,
                        const float* word, int64_t word_sb, int64_t word_sc, const float* gl_img, int64_t gl_sr,
                        const float* sent, int64_t sent_sr, const void* const* params, int n_params, int B, int T,
                        int training, float momentum, float eps, void* const* running_stats, float* out,
                        int64_t out_sr, void* saved, size_t saved_bytes, void* stream);
int tgfr_fcfm_train_bwd(const float* gout, int64_t gout_sr, const float* word, int64_t word_sb, int64_t word_sc,
                        const float* gl_img, int64_t gl_sr, const float* sent, int64_t sent_sr,
                        const void* const* params, int n_params, int B, int T, int training, const void* saved,
                        size_t saved_bytes, void* const* dparams, float* dimg, float* dword, float* dgl_img,
                        float* dsent, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TGFR_B200_H_ */

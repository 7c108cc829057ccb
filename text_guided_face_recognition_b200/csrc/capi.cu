// extern "C" surface of libtgfr_b200.so (declared in include/tgfr_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace tgfr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// wordregion_simt.cu
int wordregion_fwd_simt(const float*, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, int64_t,
                        const int32_t*, int, int, int, int, int, float, float, float, float, float*, float*, int,
                        cudaStream_t);
int wordregion_bwd_simt(const float*, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, int64_t,
                        const int32_t*, int, int, int, int, int, float, float, float, float, const float*, float*,
                        float*, cudaStream_t);
int attention_fwd_simt(const float*, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, int64_t,
                       const int32_t*, int, int, int, int, float, float*, float*, cudaStream_t);
int attention_bwd_simt(const float*, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, int64_t, int, int,
                       int, int, float, const float*, const float*, float*, float*, cudaStream_t);
// wordregion_tc.cu
size_t wordregion_tc_workspace_bytes(int Bc, int Bq, int T, int R, int D);
size_t wordregion_tc_saved_bytes(int Bc, int Bq, int T, int R, int D);
int wordregion_tc_saved_mode(int Bc, int Bq, int T, int R, int D, const void* saved, size_t saved_bytes);
bool wordregion_tc_recompute_ok(int Bq, int T, int R, int D);
int wordregion_fwd_tc(const float*, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, int64_t,
                      const int32_t*, int, int, int, int, int, float, float, float, float, float*, float*, int, void*, size_t,
                      void*, size_t, cudaStream_t);
int wordregion_bwd_tc(const float*, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, int64_t,
                      const int32_t*, int, int, int, int, int, float, float, float, const float*, float*, float*, void*,
                      size_t, const void*, size_t, cudaStream_t);
int wordregion_tc_set_trace(void*);
// dense_simt.cu
int cosine_scores_fwd(const float*, int64_t, const float*, int64_t, int, int, int, float, int, float, const int64_t*,
                      const int64_t*, int, float*, float*, float*, cudaStream_t);
size_t cosine_workspace_bytes(int, int, int);
int cosine_scores_bwd(const float*, int64_t, const float*, int64_t, int, int, int, float, int, float, const float*,
                      const float*, const float*, float*, float*, void*, size_t, cudaStream_t);
int pair_ce_stats(const float*, int, int, int, float*, float*, float*, float*, cudaStream_t);
int pair_ce_finish(const float*, const float*, const float*, const float*, int, int, int, float, float*, float*,
                   cudaStream_t);
int pair_ce_bwd(const float*, const float*, const float*, const float*, const float*, int, int, int, float, float*,
                cudaStream_t);
int cos_logits_fwd(const float*, int64_t, const float*, int64_t, int64_t, int, int, int, float, int, float*, int64_t,
                   float*, float*, int, void*, size_t, cudaStream_t);
int arc_margin_apply(float*, int64_t, const int64_t*, int, int, int, float, float, int, float*, cudaStream_t);
size_t margin_workspace_bytes(int, int, int, int);
int margin_bwd(const float*, int64_t, const float*, int64_t, int64_t, const float*, const float*, const int64_t*,
               const float*, const float*, int64_t, int, int, int, int, float, float, int, float*, float*, int, void*,
               size_t, cudaStream_t);
size_t arc_fused_workspace_bytes(int, int, int);
size_t arc_fused_saved_bytes(int, int, int);
int arc_fused_fwd(const float*, int64_t, const float*, int64_t, int64_t, const int64_t*, int, int, int, int, float, float,
                  int, float*, float*, float*, float*, float*, float*, void*, size_t, void*, size_t, cudaStream_t);
int arc_fused_bwd(const float*, int64_t, const float*, int64_t, int64_t, const int64_t*, const float*, const float*,
                  const float*, const float*, const float*, int, int, int, int, float, float, int, float*, float*, void*,
                  size_t, const void*, size_t, cudaStream_t);
int mag_margin_fwd(const float*, const float*, int, int, float, int, float*, cudaStream_t);
int mag_margin_bwd(const float*, const float*, const float*, const float*, int, int, float, int, float*, float*,
                   cudaStream_t);
int ce_rows_stats(const float*, int64_t, const int64_t*, int, int, int, float*, float*, float*, cudaStream_t);
int focal_finish(const float*, const float*, const float*, int, float, float*, float*, cudaStream_t);
int ce_rows_bwd(const float*, int64_t, const int64_t*, const float*, const float*, const float*, int, int, int,
                float*, int64_t, cudaStream_t);
int cosine_rows_ref_fwd(const float*, int64_t, int64_t, const float*, int64_t, int64_t, int64_t, int, float, float*,
                        float*, cudaStream_t);
int cosine_rows_ref_bwd(const float*, int64_t, int64_t, const float*, int64_t, int64_t, int64_t, int, float,
                        const float*, const float*, float*, float*, cudaStream_t);
int merge_softmax_stats(const float*, int, int, int, float*, cudaStream_t);
int mag_ce_stats(const float*, const float*, int64_t, const int64_t*, int, int, float*, float*, float*, float*,
                 cudaStream_t);
int mag_ce_bwd(const float*, const float*, int64_t, const int64_t*, const float*, const float*, int, int, float*,
               float*, cudaStream_t);

// texthead.cu
size_t texthead_saved_bytes(int, int, int, int);
size_t texthead_workspace_bytes(int, int, int, int);
int texthead_fwd(const float*, const float* const*, const float* const*, int, int, int, int, int, float*, float*, void*,
                 size_t, cudaStream_t);
int texthead_bwd(const float*, const float*, const float*, int, int, int, int, int, float* const*, float* const*, void*,
                 size_t, const void*, size_t, cudaStream_t);

// imim.cu
size_t matmul_split_workspace_bytes(int mode, int M, int N, int K, int batch);
int matmul_split(int mode, const float* A, int64_t lda, const float* Bm, int64_t ldb, float* C, int64_t ldc, int M, int N, int K,
                 int batch, float alpha, const float* bias, int relu, int splits, int nterms, void* ws, size_t ws_bytes,
                 cudaStream_t st);
size_t imim_saved_bytes(int, int);
size_t imim_workspace_bytes(int, int);
int imim_fwd(const float*, int64_t, int64_t, int64_t, const float* const*, int, int, int, float, float, float*, float*, float*,
             void*, size_t, cudaStream_t);
int imim_bwd(const float*, const float*, const float*, int64_t, int64_t, int64_t, const float* const*, int, int, int, const void*,
             size_t, float* const*, float*, void*, size_t, cudaStream_t);

int proj_head_fwd(const float*, int64_t, const float*, const float*, int, int, int, float*, float*, cudaStream_t);
int proj_head_bwd(const float*, const float*, const float*, const float*, int64_t, const float*, int, int, int, float*, float*,
                  float*, float*, cudaStream_t);

// fcfm_train.cu
size_t fcfm_train_saved_bytes(int, int);
size_t fcfm_train_workspace_bytes(int, int);
int fcfm_train_fwd(const float*, int64_t, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, const float*, int64_t,
                   const float*, int64_t, const float* const*, int, int, int, float, float, float* const*, float*, int64_t, void*,
                   size_t, cudaStream_t);
int fcfm_train_bwd(const float*, int64_t, const float*, int64_t, int64_t, const float*, int64_t, const float*, int64_t,
                   const float* const*, int, int, int, const void*, size_t, float* const*, float*, float*, float*, float*, void*,
                   size_t, cudaStream_t);

// scoring.cu
int pair_cosine(const float*, int64_t, int64_t, const float*, int64_t, int64_t, int64_t, int, float, float*, cudaStream_t);
int row_argmax(const float*, int64_t, int, int, int64_t*, cudaStream_t);
size_t roc_workspace_bytes(int64_t);
int roc_curve(const float*, const int64_t*, int64_t, int, float*, int64_t*, int64_t*, int64_t*, void*, size_t, cudaStream_t);

// fcfm.cu
int fcfm_working_fwd(const float*, int64_t, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, int64_t, const float*,
                     int64_t, const float*, int64_t, const float* const*, int, int, float*, int64_t, cudaStream_t);
int fcfm_working_num_params();
size_t fcfm_working_workspace_bytes(int B);
int fcfm_working_fwd_tc(const float*, int64_t, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, int64_t, const float*,
                        int64_t, const float*, int64_t, const float* const*, int, int, float*, int64_t, void*, size_t, cudaStream_t);

// tc_selftest.cu
int debug_umma(const void*, const void*, float*, int, int, int, int, int, cudaStream_t);
int debug_tma_reduce(float*, int, int, cudaStream_t);
int debug_umma_2cta(const void*, const void*, float*, int, int, cudaStream_t);

}  // namespace tgfr

using namespace tgfr;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int tgfr_version(void) { return 100; }
const char* tgfr_last_error(void) { return g_err; }

int tgfr_device_check(void) {
  int dev = 0;
  TGFR_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  TGFR_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libtgfr_b200 is built for sm_100a only", dev, prop.major, prop.minor);
    return TGFR_E_ARCH;
  }
  return TGFR_OK;
}

size_t tgfr_wordregion_workspace_bytes(int Bc, int Bq, int T, int R, int D, int precision) {
  if (precision == TGFR_PREC_TC) return wordregion_tc_workspace_bytes(Bc, Bq, T, R, D);
  return 0;
}
size_t tgfr_wordregion_saved_bytes(int Bc, int Bq, int T, int R, int D, int precision) {
  if (precision == TGFR_PREC_TC) return wordregion_tc_saved_bytes(Bc, Bq, T, R, D);
  return 0;
}

int tgfr_wordregion_fwd(const float* ctx, int64_t ctx_sb, int64_t ctx_sr, int64_t ctx_sd, const float* words,
                        int64_t w_sb, int64_t w_st, int64_t w_sd, const int32_t* cap_lens, int Bc, int Bq, int T,
                        int R, int D, float gamma1, float gamma2, float gamma3, float eps, float* sim,
                        float* attn_diag, int diag_off, int precision, void* workspace, size_t workspace_bytes,
                        void* saved, size_t saved_bytes, void* stream) {
  TGFR_REQUIRE(ctx && words && sim, "wordregion_fwd: NULL tensor");
  if (precision == TGFR_PREC_TC) {
    // The B diagonal attention maps: emitted by the tensor-core forward itself (their scores are the fp16-operand ones,
    // like the loss), or -- TGFR_ATTN_MAPS=fp32 -- by the exact fp32 kernel on the matching pairs only.
    const char* am = getenv("TGFR_ATTN_MAPS");
    const bool maps_fp32 = am && strcmp(am, "fp32") == 0;
    TGFR_REQUIRE(!attn_diag || (diag_off >= 0 && diag_off + Bc <= Bq), "wordregion_fwd: diagonal outside the caption range");
    if (int rc = wordregion_fwd_tc(ctx, ctx_sb, ctx_sr, ctx_sd, words, w_sb, w_st, w_sd, cap_lens, Bc, Bq, T, R, D,
                                   gamma1, gamma2, gamma3, eps, sim, maps_fp32 ? nullptr : attn_diag, diag_off, workspace,
                                   workspace_bytes, saved, saved_bytes, ST(stream)))
      return rc;
    if (!attn_diag || !maps_fp32) return TGFR_OK;
    const int n = Bc;
    return attention_fwd_simt(ctx, ctx_sb, ctx_sr, ctx_sd, words + (int64_t)diag_off * w_sb, w_sb, w_st, w_sd,
                              cap_lens ? cap_lens + diag_off : nullptr, n, T, R, D, gamma1, nullptr, attn_diag,
                              ST(stream));
  }
  TGFR_REQUIRE(precision == TGFR_PREC_FP32, "wordregion_fwd: unknown precision %d", precision);
  return wordregion_fwd_simt(ctx, ctx_sb, ctx_sr, ctx_sd, words, w_sb, w_st, w_sd, cap_lens, Bc, Bq, T, R, D, gamma1,
                             gamma2, gamma3, eps, sim, attn_diag, diag_off, ST(stream));
}

int tgfr_wordregion_bwd(const float* ctx, int64_t ctx_sb, int64_t ctx_sr, int64_t ctx_sd, const float* words,
                        int64_t w_sb, int64_t w_st, int64_t w_sd, const int32_t* cap_lens, int Bc, int Bq, int T,
                        int R, int D, float gamma1, float gamma2, float gamma3, float eps, const float* gsim,
                        float* dctx, float* dwords, int precision, void* workspace, size_t workspace_bytes,
                        const void* saved, size_t saved_bytes, void* stream) {
  TGFR_REQUIRE(ctx && words && gsim, "wordregion_bwd: NULL tensor");
  if (precision == TGFR_PREC_TC) {
    float* dctx_tc = dctx;
    float* dwords_tc = dwords;
    if (!wordregion_tc_recompute_ok(Bq, T, R, D)) {
      // The recompute kernel (text-side gradient; face-side gradient without forward records) has no shared-memory
      // plan for this shape (T = 30 with R = 196, D = 256).  The face-side gradient from the forward's records -- the
      // reference's live path -- still runs on the tensor cores; the rest runs the exact fp32 CUDA kernels.
      const bool have_saved = wordregion_tc_saved_mode(Bc, Bq, T, R, D, saved, saved_bytes) != 0;
      float* dctx_simt = (dctx && !have_saved) ? dctx : nullptr;
      if (dctx_simt || dwords) {
        if (int rc = wordregion_bwd_simt(ctx, ctx_sb, ctx_sr, ctx_sd, words, w_sb, w_st, w_sd, cap_lens, Bc, Bq, T, R, D,
                                         gamma1, gamma2, gamma3, eps, gsim, dctx_simt, dwords, ST(stream)))
          return rc;
      }
      if (dctx_simt) dctx_tc = nullptr;
      dwords_tc = nullptr;
      if (!dctx_tc) return TGFR_OK;
    }
    return wordregion_bwd_tc(ctx, ctx_sb, ctx_sr, ctx_sd, words, w_sb, w_st, w_sd, cap_lens, Bc, Bq, T, R, D, gamma1,
                             gamma2, gamma3, gsim, dctx_tc, dwords_tc, workspace, workspace_bytes, saved, saved_bytes,
                             ST(stream));
  }
  TGFR_REQUIRE(precision == TGFR_PREC_FP32, "wordregion_bwd: unknown precision %d", precision);
  return wordregion_bwd_simt(ctx, ctx_sb, ctx_sr, ctx_sd, words, w_sb, w_st, w_sd, cap_lens, Bc, Bq, T, R, D, gamma1,
                             gamma2, gamma3, eps, gsim, dctx, dwords, ST(stream));
}

int tgfr_attention_fwd(const float* ctx, int64_t ctx_sb, int64_t ctx_sr, int64_t ctx_sd, const float* query,
                       int64_t q_sb, int64_t q_st, int64_t q_sd, int B, int T, int R, int D, float gamma1, float* wc,
                       float* attn, void* stream) {
  TGFR_REQUIRE(ctx && query, "attention_fwd: NULL tensor");
  return attention_fwd_simt(ctx, ctx_sb, ctx_sr, ctx_sd, query, q_sb, q_st, q_sd, nullptr, B, T, R, D, gamma1, wc,
                            attn, ST(stream));
}

int tgfr_attention_bwd(const float* ctx, int64_t ctx_sb, int64_t ctx_sr, int64_t ctx_sd, const float* query,
                       int64_t q_sb, int64_t q_st, int64_t q_sd, int B, int T, int R, int D, float gamma1,
                       const float* g_wc, const float* g_attn, float* dctx, float* dquery, void* stream) {
  TGFR_REQUIRE(ctx && query, "attention_bwd: NULL tensor");
  return attention_bwd_simt(ctx, ctx_sb, ctx_sr, ctx_sd, query, q_sb, q_st, q_sd, B, T, R, D, gamma1, g_wc, g_attn,
                            dctx, dquery, ST(stream));
}

int tgfr_cosine_scores_fwd(const float* x, int64_t x_sr, const float* y, int64_t y_sr, int Bx, int By, int D,
                           float scale, int normalise, float eps, const int64_t* class_ids_x,
                           const int64_t* class_ids_y, int diag_off, float* scores, float* xnorm, float* ynorm,
                           void* stream) {
  TGFR_REQUIRE(x && y && scores, "cosine_scores_fwd: NULL tensor");
  return cosine_scores_fwd(x, x_sr, y, y_sr, Bx, By, D, scale, normalise, eps, class_ids_x, class_ids_y, diag_off,
                           scores, xnorm, ynorm, ST(stream));
}

int tgfr_cosine_scores_bwd(const float* x, int64_t x_sr, const float* y, int64_t y_sr, int Bx, int By, int D,
                           float scale, int normalise, float eps, const float* xnorm, const float* ynorm,
                           const float* gscores, float* dx, float* dy, void* workspace, size_t workspace_bytes,
                           void* stream) {
  TGFR_REQUIRE(x && y && gscores, "cosine_scores_bwd: NULL tensor");
  return cosine_scores_bwd(x, x_sr, y, y_sr, Bx, By, D, scale, normalise, eps, xnorm, ynorm, gscores, dx, dy,
                           workspace, workspace_bytes, ST(stream));
}

size_t tgfr_cosine_workspace_bytes(int Bx, int By, int D) { return cosine_workspace_bytes(Bx, By, D); }

int tgfr_pair_ce_stats(const float* scores, int Bx, int By, int diag_off, float* rowlse, float* colmax,
                       float* colsum, float* diag, void* stream) {
  return pair_ce_stats(scores, Bx, By, diag_off, rowlse, colmax, colsum, diag, ST(stream));
}
int tgfr_pair_ce_finish(const float* rowlse, const float* colmax, const float* colsum, const float* diag, int Bx,
                        int By, int diag_off, float inv_b, float* losses, float* collse, void* stream) {
  return pair_ce_finish(rowlse, colmax, colsum, diag, Bx, By, diag_off, inv_b, losses, collse, ST(stream));
}
int tgfr_pair_ce_bwd(const float* scores, const float* rowlse, const float* collse, const float* g0, const float* g1,
                     int Bx, int By, int diag_off, float inv_b, float* gscores, void* stream) {
  return pair_ce_bwd(scores, rowlse, collse, g0, g1, Bx, By, diag_off, inv_b, gscores, ST(stream));
}

int tgfr_cos_logits_fwd(const float* x, int64_t x_sr, const float* w, int64_t w_sc, int64_t w_sk, int B, int C,
                        int Din, float s, int clamp_cos, float* out, int64_t out_sr, float* xnorm, float* wnorm,
                        int precision, void* workspace, size_t workspace_bytes, void* stream) {
  TGFR_REQUIRE(x && w && out && xnorm && wnorm, "cos_logits_fwd: NULL tensor");
  TGFR_REQUIRE(precision == TGFR_PREC_FP32 || precision == TGFR_PREC_TC, "cos_logits_fwd: unknown precision %d", precision);
  return cos_logits_fwd(x, x_sr, w, w_sc, w_sk, B, C, Din, s, clamp_cos, out, out_sr, xnorm, wnorm, precision, workspace,
                        workspace_bytes, ST(stream));
}
int tgfr_arc_margin_apply(float* logits, int64_t sr, const int64_t* labels, int B, int C, int class_off, float s,
                          float m, int easy_margin, float* cos_t, void* stream) {
  TGFR_REQUIRE(logits && labels && cos_t, "arc_margin_apply: NULL tensor");
  return arc_margin_apply(logits, sr, labels, B, C, class_off, s, m, easy_margin, cos_t, ST(stream));
}
int tgfr_arc_margin_bwd(const float* x, int64_t x_sr, const float* w, int64_t w_sc, int64_t w_sk, const float* xnorm,
                        const float* wnorm, const int64_t* labels, const float* cos_t, const float* glogits,
                        int64_t g_sr, int B, int C, int Din, int class_off, float s, float m, int easy_margin,
                        float* dx, float* dw, int precision, void* workspace, size_t workspace_bytes, void* stream) {
  TGFR_REQUIRE(x && w && xnorm && wnorm && labels && cos_t && glogits, "arc_margin_bwd: NULL tensor");
  TGFR_REQUIRE(precision == TGFR_PREC_FP32 || precision == TGFR_PREC_TC, "arc_margin_bwd: unknown precision %d", precision);
  return margin_bwd(x, x_sr, w, w_sc, w_sk, xnorm, wnorm, labels, cos_t, glogits, g_sr, B, C, Din, class_off, s, m,
                    easy_margin, dx, dw, precision, workspace, workspace_bytes, ST(stream));
}
size_t tgfr_margin_workspace_bytes(int B, int C, int Din, int precision) {
  return margin_workspace_bytes(B, C, Din, precision);
}

size_t tgfr_arc_fused_workspace_bytes(int B, int C, int Din) { return arc_fused_workspace_bytes(B, C, Din); }
size_t tgfr_arc_fused_saved_bytes(int B, int C, int Din) { return arc_fused_saved_bytes(B, C, Din); }
int tgfr_arc_fused_fwd(const float* x, int64_t x_sr, const float* w, int64_t w_sc, int64_t w_sk, const int64_t* labels,
                       int B, int C, int Din, int class_off, float s, float m, int easy_margin, float* xnorm,
                       float* wnorm, float* rowmax, float* rowsum, float* tgt, float* cos_t, void* workspace,
                       size_t workspace_bytes, void* saved, size_t saved_bytes, void* stream) {
  TGFR_REQUIRE(x && w && labels && xnorm && wnorm && rowmax && rowsum && tgt && cos_t, "arc_fused_fwd: NULL tensor");
  return arc_fused_fwd(x, x_sr, w, w_sc, w_sk, labels, B, C, Din, class_off, s, m, easy_margin, xnorm, wnorm, rowmax,
                       rowsum, tgt, cos_t, workspace, workspace_bytes, saved, saved_bytes, ST(stream));
}
int tgfr_arc_fused_bwd(const float* x, int64_t x_sr, const float* w, int64_t w_sc, int64_t w_sk, const int64_t* labels,
                       const float* xnorm, const float* wnorm, const float* lse, const float* coef, const float* gout,
                       int B, int C, int Din, int class_off, float s, float m, int easy_margin, float* dx, float* dw,
                       void* workspace, size_t workspace_bytes, const void* saved, size_t saved_bytes, void* stream) {
  TGFR_REQUIRE(x && w && labels && xnorm && wnorm && lse, "arc_fused_bwd: NULL tensor");
  return arc_fused_bwd(x, x_sr, w, w_sc, w_sk, labels, xnorm, wnorm, lse, coef, gout, B, C, Din, class_off, s, m,
                       easy_margin, dx, dw, workspace, workspace_bytes, saved, saved_bytes, ST(stream));
}

int tgfr_mag_margin_fwd(const float* cos_s, const float* margin, int B, int C, float scale, int easy_margin,
                        float* cos_m_s, void* stream) {
  return mag_margin_fwd(cos_s, margin, B, C, scale, easy_margin, cos_m_s, ST(stream));
}
int tgfr_mag_margin_bwd(const float* cos_s, const float* margin, const float* g_cos, const float* g_cosm, int B,
                        int C, float scale, int easy_margin, float* gtotal, float* gmargin, void* stream) {
  return mag_margin_bwd(cos_s, margin, g_cos, g_cosm, B, C, scale, easy_margin, gtotal, gmargin, ST(stream));
}
int tgfr_cos_logits_bwd(const float* x, int64_t x_sr, const float* w, int64_t w_sc, int64_t w_sk, const float* xnorm,
                        const float* wnorm, const float* out, int64_t out_sr, const float* gout, int64_t g_sr, int B,
                        int C, int Din, float s, int clamp_cos, float* dx, float* dw, int precision, void* workspace,
                        size_t workspace_bytes, void* stream) {
  (void)out; (void)out_sr; (void)clamp_cos;
  TGFR_REQUIRE(x && w && xnorm && wnorm && gout, "cos_logits_bwd: NULL tensor");
  TGFR_REQUIRE(precision == TGFR_PREC_FP32 || precision == TGFR_PREC_TC, "cos_logits_bwd: unknown precision %d", precision);
  return margin_bwd(x, x_sr, w, w_sc, w_sk, xnorm, wnorm, nullptr, nullptr, gout, g_sr, B, C, Din, 0, s, 0.f, 0, dx,
                    dw, precision, workspace, workspace_bytes, ST(stream));
}

int tgfr_ce_rows_stats(const float* logits, int64_t sr, const int64_t* labels, int B, int C, int class_off,
                       float* rowmax, float* rowsum, float* tgt, void* stream) {
  return ce_rows_stats(logits, sr, labels, B, C, class_off, rowmax, rowsum, tgt, ST(stream));
}
int tgfr_focal_finish(const float* rowmax, const float* rowsum, const float* tgt, int B, float gamma, float* out,
                      float* lse, void* stream) {
  return focal_finish(rowmax, rowsum, tgt, B, gamma, out, lse, ST(stream));
}
int tgfr_ce_rows_bwd(const float* logits, int64_t sr, const int64_t* labels, const float* lse, const float* coef,
                     const float* gout, int B, int C, int class_off, float* glogits, int64_t g_sr, void* stream) {
  return ce_rows_bwd(logits, sr, labels, lse, coef, gout, B, C, class_off, glogits, g_sr, ST(stream));
}
int tgfr_cosine_rows_fwd(const float* x1, int64_t x1_sr, int64_t x1_sd, const float* x2, int64_t x2_sr, int64_t x2_sd,
                         int64_t N, int D, float eps, float* out, float* stats, void* stream) {
  TGFR_REQUIRE(N >= 0 && D >= 1, "cosine_rows: bad shape N=%lld D=%d", (long long)N, D);
  return cosine_rows_ref_fwd(x1, x1_sr, x1_sd, x2, x2_sr, x2_sd, N, D, eps, out, stats, ST(stream));
}
int tgfr_cosine_rows_bwd(const float* x1, int64_t x1_sr, int64_t x1_sd, const float* x2, int64_t x2_sr, int64_t x2_sd,
                         int64_t N, int D, float eps, const float* stats, const float* gout, float* dx1, float* dx2,
                         void* stream) {
  return cosine_rows_ref_bwd(x1, x1_sr, x1_sd, x2, x2_sr, x2_sd, N, D, eps, stats, gout, dx1, dx2, ST(stream));
}
int tgfr_merge_softmax_stats(const float* gathered, int n, int K, int M, float* out, void* stream) {
  return merge_softmax_stats(gathered, n, K, M, out, ST(stream));
}
int tgfr_mag_ce_stats(const float* cos_s, const float* cos_m, int64_t sr, const int64_t* labels, int B, int C,
                      float* rowmax, float* rowsum, float* tgt, float* one_hot, void* stream) {
  return mag_ce_stats(cos_s, cos_m, sr, labels, B, C, rowmax, rowsum, tgt, one_hot, ST(stream));
}
int tgfr_mag_ce_bwd(const float* cos_s, const float* cos_m, int64_t sr, const int64_t* labels, const float* lse,
                    const float* gout, int B, int C, float* g_cos, float* g_cosm, void* stream) {
  return mag_ce_bwd(cos_s, cos_m, sr, labels, lse, gout, B, C, g_cos, g_cosm, ST(stream));
}

size_t tgfr_matmul_split_workspace_bytes(int mode, int M, int N, int K, int batch) {
  return matmul_split_workspace_bytes(mode, M, N, K, batch);
}
int tgfr_matmul_split(int mode, const float* a, int64_t lda, const float* b, int64_t ldb, float* c, int64_t ldc, int M, int N,
                      int K, int batch, float alpha, const float* bias, int relu, int splits, int nterms, void* workspace,
                      size_t workspace_bytes, void* stream) {
  TGFR_REQUIRE(a && b && c, "matmul_split: NULL tensor");
  return matmul_split(mode, a, lda, b, ldb, c, ldc, M, N, K, batch, alpha, bias, relu, splits, nterms, workspace,
                      workspace_bytes, ST(stream));
}

size_t tgfr_imim_saved_bytes(int B, int P) { return imim_saved_bytes(B, P); }
size_t tgfr_imim_workspace_bytes(int B, int P) { return imim_workspace_bytes(B, P); }
int tgfr_imim_num_params(void) { return 16; }
int tgfr_imim_fwd(const float* x, int64_t x_sb, int64_t x_sc, int64_t x_sp, const void* const* params, int n_params, int B,
                  int P, int training, float momentum, float eps, float* running_mean, float* running_var, float* out,
                  void* saved, size_t saved_bytes, void* stream) {
  TGFR_REQUIRE(x && params && out, "imim_fwd: NULL tensor");
  TGFR_REQUIRE(n_params == 16, "imim_fwd: expected 16 parameter tensors, got %d", n_params);
  TGFR_REQUIRE(training || (running_mean && running_var), "imim_fwd: evaluation mode needs the running statistics");
  return imim_fwd(x, x_sb, x_sc, x_sp, reinterpret_cast<const float* const*>(params), B, P, training, momentum, eps,
                  running_mean, running_var, out, saved, saved_bytes, ST(stream));
}
int tgfr_imim_bwd(const float* gout, const float* out, const float* x, int64_t x_sb, int64_t x_sc, int64_t x_sp,
                  const void* const* params, int n_params, int B, int P, int training, const void* saved, size_t saved_bytes,
                  void* const* dparams, float* dx, void* workspace, size_t workspace_bytes, void* stream) {
  TGFR_REQUIRE(gout && out && x && params && dparams, "imim_bwd: NULL tensor");
  TGFR_REQUIRE(n_params == 16, "imim_bwd: expected 16 parameter tensors, got %d", n_params);
  return imim_bwd(gout, out, x, x_sb, x_sc, x_sp, reinterpret_cast<const float* const*>(params), B, P, training, saved,
                  saved_bytes, reinterpret_cast<float* const*>(dparams), dx, workspace, workspace_bytes, ST(stream));
}

int tgfr_proj_head_fwd(const float* x, int64_t x_sr, const float* weight, const float* bias, int M, int N, int K, float* out,
                       float* znorm, void* stream) {
  TGFR_REQUIRE(x && weight && out && znorm, "proj_head_fwd: NULL tensor");
  return proj_head_fwd(x, x_sr, weight, bias, M, N, K, out, znorm, ST(stream));
}
int tgfr_proj_head_bwd(const float* gout, const float* out, const float* znorm, const float* x, int64_t x_sr,
                       const float* weight, int M, int N, int K, float* dz_scratch, float* dx, float* dweight, float* dbias,
                       void* stream) {
  TGFR_REQUIRE(gout && out && znorm && x && weight && dz_scratch && dweight && dbias, "proj_head_bwd: NULL tensor");
  return proj_head_bwd(gout, out, znorm, x, x_sr, weight, M, N, K, dz_scratch, dx, dweight, dbias, ST(stream));
}

size_t tgfr_fcfm_train_saved_bytes(int B, int T) { return fcfm_train_saved_bytes(B, T); }
size_t tgfr_fcfm_train_workspace_bytes(int B, int T) { return fcfm_train_workspace_bytes(B, T); }
int tgfr_fcfm_train_fwd(const float* img, int64_t img_sb, int64_t img_sc, int64_t img_sh, int64_t img_sw, const float* word,
                        int64_t word_sb, int64_t word_sc, const float* gl_img, int64_t gl_sr, const float* sent, int64_t sent_sr,
                        const void* const* params, int n_params, int B, int T, int training, float momentum, float eps,
                        void* const* running_stats, float* out, int64_t out_sr, void* saved, size_t saved_bytes, void* stream) {
  TGFR_REQUIRE(img && word && gl_img && sent && params && out && running_stats, "fcfm_train_fwd: NULL tensor");
  TGFR_REQUIRE(n_params == 22, "fcfm_train_fwd: expected 22 parameter tensors, got %d", n_params);
  TGFR_REQUIRE(B >= 1 && T >= 1, "fcfm_train_fwd: empty shape");
  return fcfm_train_fwd(img, img_sb, img_sc, img_sh, img_sw, word, word_sb, word_sc, gl_img, gl_sr, sent, sent_sr,
                        reinterpret_cast<const float* const*>(params), B, T, training, momentum, eps,
                        reinterpret_cast<float* const*>(running_stats), out, out_sr, saved, saved_bytes, ST(stream));
}
int tgfr_fcfm_train_bwd(const float* gout, int64_t gout_sr, const float* word, int64_t word_sb, int64_t word_sc,
                        const float* gl_img, int64_t gl_sr, const float* sent, int64_t sent_sr, const void* const* params,
                        int n_params, int B, int T, int training, const void* saved, size_t saved_bytes, void* const* dparams,
                        float* dimg, float* dword, float* dgl_img, float* dsent, void* workspace, size_t workspace_bytes,
                        void* stream) {
  TGFR_REQUIRE(gout && word && gl_img && sent && params && dparams, "fcfm_train_bwd: NULL tensor");
  TGFR_REQUIRE(n_params == 22, "fcfm_train_bwd: expected 22 parameter tensors, got %d", n_params);
  return fcfm_train_bwd(gout, gout_sr, word, word_sb, word_sc, gl_img, gl_sr, sent, sent_sr,
                        reinterpret_cast<const float* const*>(params), B, T, training, saved, saved_bytes,
                        reinterpret_cast<float* const*>(dparams), dimg, dword, dgl_img, dsent, workspace, workspace_bytes,
                        ST(stream));
}

int tgfr_pair_cosine(const float* x1, int64_t x1_sr, int64_t x1_sd, const float* x2, int64_t x2_sr, int64_t x2_sd, int64_t N,
                     int D, float eps, float* scores, void* stream) {
  TGFR_REQUIRE(N >= 0 && D >= 1, "pair_cosine: bad shape N=%lld D=%d", (long long)N, D);
  TGFR_REQUIRE(N == 0 || (x1 && x2 && scores), "pair_cosine: NULL tensor");
  return pair_cosine(x1, x1_sr, x1_sd, x2, x2_sr, x2_sd, N, D, eps, scores, ST(stream));
}
int tgfr_row_argmax(const float* scores, int64_t sr, int rows, int cols, int64_t* index, void* stream) {
  TGFR_REQUIRE(rows >= 0 && cols >= 1, "row_argmax: bad shape %d x %d", rows, cols);
  TGFR_REQUIRE(rows == 0 || (scores && index), "row_argmax: NULL tensor");
  return row_argmax(scores, sr, rows, cols, index, ST(stream));
}
size_t tgfr_roc_workspace_bytes(int64_t N) { return roc_workspace_bytes(N); }
int tgfr_roc_curve(const float* scores, const int64_t* labels, int64_t N, int drop_intermediate, float* thresholds,
                   int64_t* fps, int64_t* tps, int64_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
  TGFR_REQUIRE(N >= 0 && N < (int64_t(1) << 31), "roc_curve: N = %lld outside [0, 2^31)", (long long)N);
  TGFR_REQUIRE(counts && workspace, "roc_curve: NULL counts / workspace");
  TGFR_REQUIRE(N == 0 || (scores && labels && thresholds && fps && tps), "roc_curve: NULL tensor");
  return roc_curve(scores, labels, N, drop_intermediate, thresholds, fps, tps, counts, workspace, workspace_bytes,
                   ST(stream));
}

int tgfr_fcfm_working_num_params(void) { return fcfm_working_num_params(); }
int tgfr_fcfm_working_fwd(const float* img, int64_t img_sb, int64_t img_sc, int64_t img_sh, int64_t img_sw, const float* word,
                          int64_t word_sb, int64_t word_sd, int64_t word_st, const float* gl_img, int64_t gl_sr,
                          const float* sent, int64_t sent_sr, const float* const* params_host, int n_params, int B, int T,
                          float* out, int64_t out_sr, void* stream) {
  TGFR_REQUIRE(params_host && n_params == fcfm_working_num_params(), "fcfm_working_fwd: expected %d parameter pointers, got %d",
               fcfm_working_num_params(), n_params);
  TGFR_REQUIRE(B == 0 || (img && word && gl_img && sent && out), "fcfm_working_fwd: NULL tensor");
  return fcfm_working_fwd(img, img_sb, img_sc, img_sh, img_sw, word, word_sb, word_sd, word_st, gl_img, gl_sr, sent, sent_sr,
                          params_host, B, T, out, out_sr, ST(stream));
}

size_t tgfr_fcfm_working_workspace_bytes(int B) { return fcfm_working_workspace_bytes(B); }
int tgfr_fcfm_working_fwd_tc(const float* img, int64_t img_sb, int64_t img_sc, int64_t img_sh, int64_t img_sw, const float* word,
                             int64_t word_sb, int64_t word_sd, int64_t word_st, const float* gl_img, int64_t gl_sr,
                             const float* sent, int64_t sent_sr, const float* const* params_host, int n_params, int B, int T,
                             float* out, int64_t out_sr, void* workspace, size_t workspace_bytes, void* stream) {
  TGFR_REQUIRE(params_host && n_params == fcfm_working_num_params(), "fcfm_working_fwd: expected %d parameter pointers, got %d",
               fcfm_working_num_params(), n_params);
  TGFR_REQUIRE(B == 0 || (img && word && gl_img && sent && out), "fcfm_working_fwd: NULL tensor");
  return fcfm_working_fwd_tc(img, img_sb, img_sc, img_sh, img_sw, word, word_sb, word_sd, word_st, gl_img, gl_sr, sent, sent_sr,
                             params_host, B, T, out, out_sr, workspace, workspace_bytes, ST(stream));
}

int tgfr_debug_umma(const void* a, const void* b, float* out, int N, int K, int a_mn, int b_mn, int manual_a,
                    void* stream) {
  return debug_umma(a, b, out, N, K, a_mn, b_mn, manual_a, ST(stream));
}
int tgfr_debug_set_trace(void* dev_buf) { return wordregion_tc_set_trace(dev_buf); }
int tgfr_debug_umma_2cta(const void* a, const void* b, float* out, int N, int K, void* stream) {
  TGFR_REQUIRE(a && b && out, "debug_umma_2cta: NULL tensor");
  return debug_umma_2cta(a, b, out, N, K, ST(stream));
}
int tgfr_debug_tma_reduce(float* out, int rows, int cols, void* stream) {
  return debug_tma_reduce(out, rows, cols, ST(stream));
}


size_t tgfr_texthead_saved_bytes(int B, int L, int E, int F) { return texthead_saved_bytes(B, L, E, F); }
size_t tgfr_texthead_workspace_bytes(int B, int L, int E, int F) { return texthead_workspace_bytes(B, L, E, F); }
int tgfr_texthead_fwd(const float* tokens, const float* w2, const float* w3, const float* w4, const float* b2,
                      const float* b3, const float* b4, int B, int L, int E, int F, int bert_words_num, float* words,
                      float* sent, void* saved, size_t saved_bytes, void* stream) {
  TGFR_REQUIRE(tokens && w2 && w3 && w4 && words && sent, "texthead_fwd: NULL tensor");
  const float* w[3] = {w2, w3, w4};
  const float* b[3] = {b2, b3, b4};
  return texthead_fwd(tokens, w, b, B, L, E, F, bert_words_num, words, sent, saved, saved_bytes, ST(stream));
}
int tgfr_texthead_bwd(const float* tokens, const float* gwords, const float* gsent, int B, int L, int E, int F,
                      int bert_words_num, float* dw2, float* dw3, float* dw4, float* db2, float* db3, float* db4,
                      void* workspace, size_t workspace_bytes, const void* saved, size_t saved_bytes, void* stream) {
  TGFR_REQUIRE(tokens && dw2 && dw3 && dw4 && db2 && db3 && db4, "texthead_bwd: NULL tensor");
  float* dw[3] = {dw2, dw3, dw4};
  float* db[3] = {db2, db3, db4};
  return texthead_bwd(tokens, gwords, gsent, B, L, E, F, bert_words_num, dw, db, workspace, workspace_bytes, saved,
                      saved_bytes, ST(stream));
}
}  // extern "C"

// IMIM / ImageHeading local branch (reference models/models.py:380-405 + SelfAttention models/fusion_nets.py:82-118):
// the producer of the word-region loss's region features, forward AND backward.
//
//   x [B,256,14,14] -> BatchNorm2d (batch statistics when training) -> 196 x 196 self-attention over the positions
//   (1x1 query / key / value projections, softmax(k_i . q_j / 16) over j, response = attention . value)
//   -> LayerNorm([256,14,14]) -> relu(conv1x1 256->128) -> relu(conv1x1 128->256) -> Linear 256->256
//   -> L2 normalisation (twice, as the reference does) -> [B,14,14,256] in memory (= the channels-last layout the
//   word-region kernels read, logical [B,256,14,14]).
//
// Layout: every activation is position-major [B*196, C] (channels contiguous), so each 1x1 convolution / linear layer is
// one row-major product over M = B*196 rows and the attention is a strided-batched product per sample.  The
// normalisations (BatchNorm statistics / apply + transpose, softmax, LayerNorm, ReLU masks, L2 norm) are fused
// element-wise kernels.  The contractions (85 % of the fp32 step) run on the tensor cores: every operand is split once
// into fp16 hi + lo with a power-of-two scale (gemm_tc_split_operand) and each product accumulates
// A_hi B_hi + A_hi B_lo + A_lo B_hi in one TMEM accumulator (gemm_tc_pair in gemm_tc.cu: TMA-fed tcgen05 tiles, NT / NN /
// TN through K-major / MN-major descriptors, a strided batch per sample for the attention, split-K with vector
// reductions for the weight gradients whose K is B*196) -- ~22 significant bits, the same fp32-class contract as
// TextHeading.  The forward keeps its fp16 operand copies in `saved`, so the backward splits only the gradients.
// TGFR_IMIM_PRECISION=fp32 selects the register-blocked fp32 SIMT kernel (sgemm_kernel) for every product.  Measured
// against the float64 run of the reference module (tools/imim_precision.py, B = 128): both modes 5e-7 on the output and
// on every tile of every gradient, except where a ReLU pre-activation within rounding of zero flips its mask (one entry
// in 6.4 M at B = 128, worth 3e-4 of the gradient norm -- the fp32 reference differs from its own float64 run by as
// much).  The output is the unit-norm operand of the word-region scores and its gradient trains image_head
// (src/train_encoders_bert.py:263-265).
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "nn_blocks.cuh"

namespace tgfr {

struct TcOperand {          // gemm_tc.cu
  const __half *hi, *lo;
  int64_t ld;
  const float* scale;
};
int gemm_tc_pair(int mode, const TcOperand& A, const TcOperand& B, float* C, int64_t ldc, int64_t c_bs, int M, int N, int K,
                 int batch, float alpha, const float* bias, int relu, int splits, int nterms, cudaStream_t st,
                 float* amax = nullptr);
int gemm_tc_split_operand(const float* src, int64_t ld, int rows, int cols, float* scale, __half* hi, __half* lo, int ld_out,
                          cudaStream_t st, bool have_max = false);

namespace {

constexpr int kC = 256, kC2 = 128;

// 0 = fp32 SIMT products (TGFR_IMIM_PRECISION=fp32), 3 = hi / lo split on the tensor cores (default).  A single fp16
// term is not offered: the K = B*196 reductions of the backward amplify its 2^-11 operand rounding to 2 % gradients.
int imim_terms() {
  const char* e = getenv("TGFR_IMIM_PRECISION");
  return (e && strcmp(e, "fp32") == 0) ? 0 : 3;
}

// fp16 hi / lo copies of one fp32 matrix: [rows, ld] each, lo right behind hi
struct H16 {
  __half* hi;
  __half* lo;
  float* scale;
  int ld;
  TcOperand op(int64_t col0 = 0) const { return TcOperand{hi + col0, lo + col0, ld, scale}; }
};
struct HalfPool {            // bump allocator over a 256-byte aligned region; scales come from a float region
  uint8_t* base;
  float* scales;
  size_t off = 0;
  int nscale = 0;
  H16 take(size_t rows, int ld) {
    H16 h;
    const size_t bytes = (rows * ld * sizeof(__half) + 255) / 256 * 256;
    h.hi = base ? reinterpret_cast<__half*>(base + off) : nullptr;
    h.lo = base ? reinterpret_cast<__half*>(base + off + bytes) : nullptr;
    h.scale = scales ? scales + 2 * nscale : nullptr;
    h.ld = ld;
    off += 2 * bytes;
    ++nscale;
    return h;
  }
};
int split_into(const float* src, int64_t ld, int rows, int cols, const H16& h, cudaStream_t st) {
  return gemm_tc_split_operand(src, ld, rows, cols, h.scale, h.hi, h.lo, h.ld, st);
}
// The large activations / gradients skip the max-abs pass: the kernel that produces the tensor leaves max |x| in h.scale[0]
// (zero_scale before it, split_ready after it).
int zero_scale(const H16& h, cudaStream_t st) {
  TGFR_CUDA_OK(cudaMemsetAsync(h.scale, 0, 2 * sizeof(float), st));
  return TGFR_OK;
}
int split_ready(const float* src, int64_t ld, int rows, int cols, const H16& h, cudaStream_t st) {
  static const bool fused = !(getenv("TGFR_IMIM_FUSED_MAX") && atoi(getenv("TGFR_IMIM_FUSED_MAX")) == 0);   // 0: A/B switch
  return gemm_tc_split_operand(src, ld, rows, cols, h.scale, h.hi, h.lo, h.ld, st, fused);
}
constexpr int kScaleSlots = 32;

// the forward's operand copies (kept for the backward) / the backward's gradient copies
struct ImimHalf {
  H16 xn, qkv, prob, y, h1, h2, wqkv, w10, w12, w14;
  size_t bytes;
};
ImimHalf imim_half(int B, int P, void* base, float* scales) {
  HalfPool pool{reinterpret_cast<uint8_t*>(base), scales};
  const size_t M = (size_t)B * P;
  const int Pp = (P + 7) & ~7;
  ImimHalf h;
  h.xn = pool.take(M, kC); h.qkv = pool.take(M, 3 * kC); h.prob = pool.take(M, Pp); h.y = pool.take(M, kC);
  h.h1 = pool.take(M, kC2); h.h2 = pool.take(M, kC); h.wqkv = pool.take(3 * kC, kC); h.w10 = pool.take(kC2, kC);
  h.w12 = pool.take(kC, kC2); h.w14 = pool.take(kC, kC);
  h.bytes = pool.off;
  return h;
}
struct ImimGradHalf {
  H16 g256, g128, gp, g768;
  size_t bytes;
};
ImimGradHalf imim_grad_half(int B, int P, void* base, float* scales) {
  HalfPool pool{reinterpret_cast<uint8_t*>(base), scales};
  const size_t M = (size_t)B * P;
  const int Pp = (P + 7) & ~7;
  ImimGradHalf h;
  h.g256 = pool.take(M, kC); h.g128 = pool.take(M, kC2); h.gp = pool.take(M, Pp); h.g768 = pool.take(M, 3 * kC);
  h.bytes = pool.off;
  return h;
}

// saved / workspace layout (floats), M = B * P
struct ImimLayout {
  size_t mean, invstd, xn, qkv, prob, o, ln_mu, ln_rstd, y, h1, h2, znorm, wqkv, bqkv, scales, bn_part, ln_part, ln_wt, ln_bt, total;
};
ImimLayout imim_layout(int B, int P) {
  ImimLayout L;
  const size_t M = (size_t)B * P;
  size_t off = 0;
  auto take = [&](size_t n) { const size_t o = off; off += (n + 63) / 64 * 64; return o; };
  L.mean = take(kC); L.invstd = take(kC); L.xn = take(M * kC); L.qkv = take(M * 3 * kC); L.prob = take((size_t)B * P * P);
  L.o = take(M * kC); L.ln_mu = take(B); L.ln_rstd = take(B); L.y = take(M * kC); L.h1 = take(M * kC2); L.h2 = take(M * kC);
  L.znorm = take(M); L.wqkv = take((size_t)3 * kC * kC); L.bqkv = take(3 * kC);
  L.scales = take(2 * kScaleSlots);
  L.bn_part = take((size_t)kC * kBnSplit * 2); L.ln_part = take((size_t)B * kLnSplit * 2);
  L.ln_wt = take((size_t)P * kC); L.ln_bt = take((size_t)P * kC);       // ln.weight / ln.bias position-major
  L.total = off;                                      // floats; the fp16 operand copies (imim_half) follow
  return L;
}

}  // namespace

// parameter order of tgfr_imim_fwd / dparams of tgfr_imim_bwd (the reference's state_dict names):
//  0 bn_img.weight  1 bn_img.bias  2 sa.query_proj.weight [256,256]  3 .bias  4 sa.key_proj.weight  5 .bias
//  6 sa.value_proj.weight  7 .bias  8 ln.weight [256*196]  9 ln.bias  10 conv1x1_1.weight [128,256]  11 .bias
//  12 conv1x1_2.weight [256,128]  13 .bias  14 project_local.projection.weight [256,256]  15 .bias

size_t imim_saved_bytes(int B, int P) {
  return imim_layout(B, P).total * sizeof(float) + imim_half(B, P, nullptr, nullptr).bytes;
}
static size_t imim_ws_floats(int B, int P) {
  // backward scratch: dZ/dH2 [M,256], dH1 [M,128], dY/dO [M,256], dP [B,P,P], dQKV [M,768], dxn [M,256], scales
  const size_t M = (size_t)B * P;
  return (M * kC * 3 + M * kC2 + (size_t)B * P * P + M * 3 * kC + (size_t)3 * kC * kC + 3 * kC + (size_t)B * kLnSplit * 2 + 1024 +
          63) / 64 * 64;
}

// LayerNorm([256,14,14]) forward / backward on position-major data, kLnSplit blocks per sample
int ln_forward(const float* o, int B, int P, const float* w, const float* bia, float* wt, float* bt, float* part, float* y,
               float* mu, float* rstd, cudaStream_t st, float* amax = nullptr) {
  const int n = P * kC;
  ln_affine_t_kernel<<<ceil_div(n, 256), 256, 0, st>>>(w, bia, P, kC, wt, bt);
  TGFR_LAUNCH_OK();
  ln_stats_part_kernel<<<dim3(kLnSplit, B), 512, 0, st>>>(o, (int64_t)n, n, part);
  TGFR_LAUNCH_OK();
  ln_apply_kernel<<<dim3(ceil_div(n, 1024), B), 256, 0, st>>>(o, (int64_t)n, P, kC, wt, bt, 1, part, y, (int64_t)n, mu, rstd, amax);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}
// dY -> dO in place; d ln.weight / d ln.bias first (they need dY).  wt = the forward's position-major copy of ln.weight
int ln_backward(float* dY, const float* o, int B, int P, const float* wt, const float* mu, const float* rstd, float* part,
                float* dw, float* db, cudaStream_t st, float* amax = nullptr) {
  const int n = P * kC;
  TGFR_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * n, st));
  TGFR_CUDA_OK(cudaMemsetAsync(db, 0, sizeof(float) * n, st));
  ln_bwd_params_split_kernel<<<dim3(ceil_div(n, 128), B >= 32 ? 8 : 1), 128, 0, st>>>(dY, (int64_t)n, o, (int64_t)n, B, P, kC, mu, rstd,
                                                                                     dw, db);
  TGFR_LAUNCH_OK();
  ln_bwd_part_kernel<<<dim3(kLnSplit, B), 512, 0, st>>>(dY, (int64_t)n, o, (int64_t)n, P, kC, wt, 1, mu, rstd, part);
  TGFR_LAUNCH_OK();
  ln_bwd_apply_kernel<<<dim3(ceil_div(n, 1024), B), 256, 0, st>>>(dY, (int64_t)n, o, (int64_t)n, P, kC, wt, 1, mu, rstd, part, dY,
                                                                 (int64_t)n, amax);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}
int bn_backward_sums(const float* dxn, const float* x, int64_t sb, int64_t sc, int64_t sp, int B, int P, const float* mean,
                     const float* invstd, float* dgamma, float* dbeta, cudaStream_t st) {
  TGFR_CUDA_OK(cudaMemsetAsync(dgamma, 0, sizeof(float) * kC, st));
  TGFR_CUDA_OK(cudaMemsetAsync(dbeta, 0, sizeof(float) * kC, st));
  bn_bwd_sums_tile_kernel<<<dim3(ceil_div(P, 32), kC / 32, B < 16 ? B : 16), dim3(32, 8), 0, st>>>(dxn, x, sb, sc, sp, B, kC, P, mean,
                                                                                                  invstd, dgamma, dbeta);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}
size_t imim_workspace_bytes(int B, int P) {
  return imim_ws_floats(B, P) * sizeof(float) + imim_grad_half(B, P, nullptr, nullptr).bytes;
}

int imim_fwd(const float* x, int64_t sb, int64_t sc, int64_t sp, const float* const* prm, int B, int P, int training,
             float momentum, float eps, float* run_mean, float* run_var, float* out, void* saved, size_t saved_bytes,
             cudaStream_t st) {
  TGFR_REQUIRE(B >= 1 && P >= 1, "imim_fwd: empty batch");
  const ImimLayout L = imim_layout(B, P);
  TGFR_REQUIRE(saved && saved_bytes >= imim_saved_bytes(B, P), "imim_fwd: saved buffer too small (%zu < %zu)", saved_bytes,
               imim_saved_bytes(B, P));
  TGFR_REQUIRE((reinterpret_cast<uintptr_t>(saved) & 255) == 0, "imim_fwd: saved buffer must be 256-byte aligned");
  float* S = reinterpret_cast<float*>(saved);
  const int M = B * P;
  const int nt = imim_terms();
  const ImimHalf H = imim_half(B, P, S + L.total, S + L.scales);
  if (training) {
    bn_stats_part_kernel<<<dim3(kC, kBnSplit), 256, 0, st>>>(x, sb, sc, sp, B, P, S + L.bn_part);
    TGFR_LAUNCH_OK();
    bn_stats_merge_kernel<<<ceil_div(kC, 128), 128, 0, st>>>(S + L.bn_part, kC, B, P, eps, momentum, run_mean, run_var, S + L.mean,
                                                            S + L.invstd);
  } else {
    bn_stats_kernel<<<kC, 256, 0, st>>>(x, sb, sc, sp, B, P, eps, momentum, training, run_mean, run_var, S + L.mean, S + L.invstd);
  }
  TGFR_LAUNCH_OK();
  if (nt) { if (int rc = zero_scale(H.xn, st)) return rc; }
  bn_apply_t_kernel<<<dim3(ceil_div(P, 32), kC / 32, B), dim3(32, 8), 0, st>>>(x, sb, sc, sp, kC, P, S + L.mean, S + L.invstd,
                                                                              prm[0], prm[1], S + L.xn, nt ? H.xn.scale : nullptr);
  TGFR_LAUNCH_OK();
  // one product for the three projections: W_qkv [768,256] = [Wq; Wk; Wv]
  for (int k = 0; k < 3; ++k) {
    TGFR_CUDA_OK(cudaMemcpyAsync(S + L.wqkv + (size_t)k * kC * kC, prm[2 + 2 * k], sizeof(float) * kC * kC, cudaMemcpyDeviceToDevice, st));
    TGFR_CUDA_OK(cudaMemcpyAsync(S + L.bqkv + (size_t)k * kC, prm[3 + 2 * k], sizeof(float) * kC, cudaMemcpyDeviceToDevice, st));
  }
  if (nt) {
    // the same graph on the tensor cores: split each operand once, products on the hi / lo pairs
    if (int rc = split_ready(S + L.xn, kC, M, kC, H.xn, st)) return rc;
    if (int rc = split_into(S + L.wqkv, kC, 3 * kC, kC, H.wqkv, st)) return rc;
    if (int rc = zero_scale(H.qkv, st)) return rc;
    if (int rc = gemm_tc_pair(0, H.xn.op(), H.wqkv.op(), S + L.qkv, 3 * kC, 0, M, 3 * kC, kC, 1, 1.f, S + L.bqkv, 0, 1, nt, st, H.qkv.scale))
      return rc;
    if (int rc = split_ready(S + L.qkv, 3 * kC, M, 3 * kC, H.qkv, st)) return rc;
    // attention[b,i,j] = softmax_j(k_i . q_j / sqrt(256))      (fusion_nets.py:97-105)
    if (int rc = gemm_tc_pair(0, H.qkv.op(kC), H.qkv.op(0), S + L.prob, P, (int64_t)P * P, P, P, kC, B, 1.f / 16.f, nullptr, 0, 1, nt, st))
      return rc;
    if (int rc = zero_scale(H.prob, st)) return rc;
    softmax_rows_kernel<<<ceil_div(B * P, 8), 256, 0, st>>>(S + L.prob, B * P, P, H.prob.scale);
    TGFR_LAUNCH_OK();
    if (int rc = split_ready(S + L.prob, P, M, P, H.prob, st)) return rc;
    // response = attention . value                              (:115)
    if (int rc = gemm_tc_pair(1, H.prob.op(), H.qkv.op(2 * kC), S + L.o, kC, (int64_t)P * kC, P, kC, P, B, 1.f, nullptr, 0, 1, nt, st))
      return rc;
    if (int rc = zero_scale(H.y, st)) return rc;
    if (int rc = ln_forward(S + L.o, B, P, prm[8], prm[9], S + L.ln_wt, S + L.ln_bt, S + L.ln_part, S + L.y, S + L.ln_mu, S + L.ln_rstd, st,
                            H.y.scale))
      return rc;
    if (int rc = split_ready(S + L.y, kC, M, kC, H.y, st)) return rc;
    if (int rc = split_into(prm[10], kC, kC2, kC, H.w10, st)) return rc;
    if (int rc = zero_scale(H.h1, st)) return rc;
    if (int rc = gemm_tc_pair(0, H.y.op(), H.w10.op(), S + L.h1, kC2, 0, M, kC2, kC, 1, 1.f, prm[11], 1, 1, nt, st, H.h1.scale)) return rc;
    if (int rc = split_ready(S + L.h1, kC2, M, kC2, H.h1, st)) return rc;
    if (int rc = split_into(prm[12], kC2, kC, kC2, H.w12, st)) return rc;
    if (int rc = zero_scale(H.h2, st)) return rc;
    if (int rc = gemm_tc_pair(0, H.h1.op(), H.w12.op(), S + L.h2, kC, 0, M, kC, kC2, 1, 1.f, prm[13], 1, 1, nt, st, H.h2.scale)) return rc;
    if (int rc = split_ready(S + L.h2, kC, M, kC, H.h2, st)) return rc;
    if (int rc = split_into(prm[14], kC, kC, kC, H.w14, st)) return rc;
    if (int rc = gemm_tc_pair(0, H.h2.op(), H.w14.op(), out, kC, 0, M, kC, kC, 1, 1.f, prm[15], 0, 1, nt, st)) return rc;
    l2norm2_rows_kernel<<<ceil_div(M, 8), 256, 0, st>>>(out, M, kC, out, S + L.znorm);
    TGFR_LAUNCH_OK();
    return TGFR_OK;
  }
  if (int rc = sgemm(0, S + L.xn, kC, 0, S + L.wqkv, kC, 0, S + L.qkv, 3 * kC, 0, M, 3 * kC, kC, 1, 1.f, S + L.bqkv, 0, 1, st)) return rc;
  // attention[b,i,j] = softmax_j(k_i . q_j / sqrt(256))      (fusion_nets.py:97-105)
  const float* Q = S + L.qkv;
  const float* K = S + L.qkv + kC;
  const float* V = S + L.qkv + 2 * kC;
  if (int rc = sgemm(0, K, 3 * kC, (int64_t)P * 3 * kC, Q, 3 * kC, (int64_t)P * 3 * kC, S + L.prob, P, (int64_t)P * P, P, P, kC, B,
                     1.f / 16.f, nullptr, 0, 1, st)) return rc;
  softmax_rows_kernel<<<ceil_div(B * P, 8), 256, 0, st>>>(S + L.prob, B * P, P);
  TGFR_LAUNCH_OK();
  // response = attention . value                              (:115)
  if (int rc = sgemm(1, S + L.prob, P, (int64_t)P * P, V, 3 * kC, (int64_t)P * 3 * kC, S + L.o, kC, (int64_t)P * kC, P, kC, P, B, 1.f,
                     nullptr, 0, 1, st)) return rc;
  if (int rc = ln_forward(S + L.o, B, P, prm[8], prm[9], S + L.ln_wt, S + L.ln_bt, S + L.ln_part, S + L.y, S + L.ln_mu, S + L.ln_rstd, st)) return rc;
  if (int rc = sgemm(0, S + L.y, kC, 0, prm[10], kC, 0, S + L.h1, kC2, 0, M, kC2, kC, 1, 1.f, prm[11], 1, 1, st)) return rc;
  if (int rc = sgemm(0, S + L.h1, kC2, 0, prm[12], kC2, 0, S + L.h2, kC, 0, M, kC, kC2, 1, 1.f, prm[13], 1, 1, st)) return rc;
  // the projection lands in `out`, then is normalised in place (Z itself is not needed again: dZ uses out and |Z|)
  if (int rc = sgemm(0, S + L.h2, kC, 0, prm[14], kC, 0, out, kC, 0, M, kC, kC, 1, 1.f, prm[15], 0, 1, st)) return rc;
  l2norm2_rows_kernel<<<ceil_div(M, 8), 256, 0, st>>>(out, M, kC, out, S + L.znorm);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int imim_bwd(const float* gout, const float* out, const float* x, int64_t sb, int64_t sc, int64_t sp, const float* const* prm,
             int B, int P, int training, const void* saved, size_t saved_bytes, float* const* dprm, float* dx, void* ws,
             size_t ws_bytes, cudaStream_t st) {
  const ImimLayout L = imim_layout(B, P);
  TGFR_REQUIRE(saved && saved_bytes >= imim_saved_bytes(B, P), "imim_bwd: saved buffer too small");
  TGFR_REQUIRE(ws && ws_bytes >= imim_workspace_bytes(B, P), "imim_bwd: workspace too small");
  TGFR_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "imim_bwd: workspace must be 256-byte aligned");
  const float* S = reinterpret_cast<const float*>(saved);
  const int M = B * P;
  const int nt = imim_terms();
  // the forward's operand copies (read-only here) and the gradient copies of this call
  const ImimHalf H = imim_half(B, P, const_cast<float*>(S) + L.total, const_cast<float*>(S) + L.scales);
  float* Wf = reinterpret_cast<float*>(ws);
  const ImimGradHalf G = imim_grad_half(B, P, Wf + imim_ws_floats(B, P), Wf + imim_ws_floats(B, P) - 2 * kScaleSlots);
  const size_t Mz = (size_t)M;
  float* W = reinterpret_cast<float*>(ws);
  float* dZ = W;                              // [M,256]  dZ, then dH2
  float* dH2 = dZ + Mz * kC;                  // [M,256]
  float* dH1 = dH2 + Mz * kC;                 // [M,128]
  float* dY = dH1 + Mz * kC2;                 // [M,256]  dY, then dO
  float* dP = dY + Mz * kC;                   // [B,P,P]
  float* dQKV = dP + (size_t)B * P * P;       // [M,768]
  float* dwqkv = dQKV + Mz * 3 * kC;          // [768,256]
  float* dbqkv = dwqkv + (size_t)3 * kC * kC; // [768]
  float* ln_part = dbqkv + 3 * kC;            // [B, kLnSplit, 2]
  float* dxn = dZ;                            // [M,256]  (dZ is dead by then)
  const int splits = M >= 4096 ? 32 : (M >= 512 ? 8 : 1);      // weight gradients: K = M rows
  // projection + L2 norm
  if (nt) { if (int rc = zero_scale(G.g256, st)) return rc; }
  l2norm_rows_bwd_kernel<<<ceil_div(M, 8), 256, 0, st>>>(gout, out, S + L.znorm, M, kC, dZ, nt ? G.g256.scale : nullptr);
  TGFR_LAUNCH_OK();
  if (nt) {
    const int64_t sqh = (int64_t)P * 3 * kC;
    const int ks = M >= 8192 ? 37 : (M >= 512 ? 8 : 1);       // split-K of the weight gradients: 4 tiles x 37 = one wave
    H16 gz = G.g256;
    if (int rc = split_ready(dZ, kC, M, kC, gz, st)) return rc;
    if (int rc = gemm_tc_pair(2, gz.op(), H.h2.op(), dprm[14], kC, 0, kC, kC, M, 1, 1.f, nullptr, 0, ks, nt, st)) return rc;
    if (int rc = colsum(dZ, kC, M, kC, dprm[15], st)) return rc;
    if (int rc = gemm_tc_pair(1, gz.op(), H.w14.op(), dH2, kC, 0, M, kC, kC, 1, 1.f, nullptr, 0, 1, nt, st)) return rc;
    if (int rc = zero_scale(gz, st)) return rc;              // (the two products above have consumed gz and its scale)
    relu_mask_kernel<<<1184, 256, 0, st>>>(dH2, S + L.h2, (int64_t)Mz * kC, gz.scale);
    TGFR_LAUNCH_OK();
    // conv1x1_2
    if (int rc = split_ready(dH2, kC, M, kC, gz, st)) return rc;
    if (int rc = gemm_tc_pair(2, gz.op(), H.h1.op(), dprm[12], kC2, 0, kC, kC2, M, 1, 1.f, nullptr, 0, ks, nt, st)) return rc;
    if (int rc = colsum(dH2, kC, M, kC, dprm[13], st)) return rc;
    if (int rc = gemm_tc_pair(1, gz.op(), H.w12.op(), dH1, kC2, 0, M, kC2, kC, 1, 1.f, nullptr, 0, 1, nt, st)) return rc;
    if (int rc = zero_scale(G.g128, st)) return rc;
    relu_mask_kernel<<<1184, 256, 0, st>>>(dH1, S + L.h1, (int64_t)Mz * kC2, G.g128.scale);
    TGFR_LAUNCH_OK();
    // conv1x1_1
    if (int rc = split_ready(dH1, kC2, M, kC2, G.g128, st)) return rc;
    if (int rc = gemm_tc_pair(2, G.g128.op(), H.y.op(), dprm[10], kC, 0, kC2, kC, M, 1, 1.f, nullptr, 0, ks, nt, st)) return rc;
    if (int rc = colsum(dH1, kC2, M, kC2, dprm[11], st)) return rc;
    if (int rc = gemm_tc_pair(1, G.g128.op(), H.w10.op(), dY, kC, 0, M, kC, kC2, 1, 1.f, nullptr, 0, 1, nt, st)) return rc;
    // LayerNorm
    if (int rc = zero_scale(gz, st)) return rc;
    if (int rc = ln_backward(dY, S + L.o, B, P, S + L.ln_wt, S + L.ln_mu, S + L.ln_rstd, ln_part, dprm[8], dprm[9], st, gz.scale))
      return rc;                                                                                    // dY -> dO
    // attention: O = P V;  S = K Q^T / 16
    if (int rc = split_ready(dY, kC, M, kC, gz, st)) return rc;                                    // dO
    if (int rc = gemm_tc_pair(0, gz.op(), H.qkv.op(2 * kC), dP, P, (int64_t)P * P, P, P, kC, B, 1.f, nullptr, 0, 1, nt, st)) return rc;   // dP = dO V^T
    if (int rc = zero_scale(G.g768, st)) return rc;          // max |dQKV| is collected by the three products that write its slices
    if (int rc = gemm_tc_pair(2, H.prob.op(), gz.op(), dQKV + 2 * kC, 3 * kC, sqh, P, kC, P, B, 1.f, nullptr, 0, 1, nt, st, G.g768.scale))
      return rc;                                                                                    // dV = P^T dO
    if (int rc = zero_scale(G.gp, st)) return rc;
    softmax_rows_bwd_kernel<<<ceil_div(B * P, 8), 256, 0, st>>>(S + L.prob, dP, B * P, P, 1.f / 16.f, G.gp.scale);   // dP -> dS
    TGFR_LAUNCH_OK();
    if (int rc = split_ready(dP, P, M, P, G.gp, st)) return rc;
    if (int rc = gemm_tc_pair(1, G.gp.op(), H.qkv.op(0), dQKV + kC, 3 * kC, sqh, P, kC, P, B, 1.f, nullptr, 0, 1, nt, st, G.g768.scale))
      return rc;                                                                                    // dK = dS Q
    if (int rc = gemm_tc_pair(2, G.gp.op(), H.qkv.op(kC), dQKV, 3 * kC, sqh, P, kC, P, B, 1.f, nullptr, 0, 1, nt, st, G.g768.scale))
      return rc;                                                                                    // dQ = dS^T K
    // projections: d W_qkv [768,256] = dQKV^T xn, biases, d xn = dQKV W_qkv
    if (int rc = split_ready(dQKV, 3 * kC, M, 3 * kC, G.g768, st)) return rc;
    if (int rc = gemm_tc_pair(2, G.g768.op(), H.xn.op(), dwqkv, kC, 0, 3 * kC, kC, M, 1, 1.f, nullptr, 0, M >= 8192 ? 24 : ks, nt, st)) return rc;
    if (int rc = colsum(dQKV, 3 * kC, M, 3 * kC, dbqkv, st)) return rc;
    for (int k = 0; k < 3; ++k) {
      TGFR_CUDA_OK(cudaMemcpyAsync(dprm[2 + 2 * k], dwqkv + (size_t)k * kC * kC, sizeof(float) * kC * kC, cudaMemcpyDeviceToDevice, st));
      TGFR_CUDA_OK(cudaMemcpyAsync(dprm[3 + 2 * k], dbqkv + (size_t)k * kC, sizeof(float) * kC, cudaMemcpyDeviceToDevice, st));
    }
    if (int rc = gemm_tc_pair(1, G.g768.op(), H.wqkv.op(), dxn, kC, 0, M, kC, 3 * kC, 1, 1.f, nullptr, 0, 1, nt, st)) return rc;
    if (int rc = bn_backward_sums(dxn, x, sb, sc, sp, B, P, S + L.mean, S + L.invstd, dprm[0], dprm[1], st)) return rc;
    if (dx) {
      bn_bwd_dx_kernel<<<dim3(ceil_div(P, 32), kC / 32, B), dim3(32, 8), 0, st>>>(dxn, x, sb, sc, sp, B, kC, P, S + L.mean, S + L.invstd,
                                                                                 prm[0], dprm[0], dprm[1], training, dx, (int64_t)kC * P, P, 1);
      TGFR_LAUNCH_OK();
    }
    return TGFR_OK;
  }
  if (int rc = sgemm(2, dZ, kC, 0, S + L.h2, kC, 0, dprm[14], kC, 0, kC, kC, M, 1, 1.f, nullptr, 0, splits, st)) return rc;
  if (int rc = colsum(dZ, kC, M, kC, dprm[15], st)) return rc;
  if (int rc = sgemm(1, dZ, kC, 0, prm[14], kC, 0, dH2, kC, 0, M, kC, kC, 1, 1.f, nullptr, 0, 1, st)) return rc;
  relu_mask_kernel<<<1184, 256, 0, st>>>(dH2, S + L.h2, (int64_t)Mz * kC);
  TGFR_LAUNCH_OK();
  // conv1x1_2
  if (int rc = sgemm(2, dH2, kC, 0, S + L.h1, kC2, 0, dprm[12], kC2, 0, kC, kC2, M, 1, 1.f, nullptr, 0, splits, st)) return rc;
  if (int rc = colsum(dH2, kC, M, kC, dprm[13], st)) return rc;
  if (int rc = sgemm(1, dH2, kC, 0, prm[12], kC2, 0, dH1, kC2, 0, M, kC2, kC, 1, 1.f, nullptr, 0, 1, st)) return rc;
  relu_mask_kernel<<<1184, 256, 0, st>>>(dH1, S + L.h1, (int64_t)Mz * kC2);
  TGFR_LAUNCH_OK();
  // conv1x1_1
  if (int rc = sgemm(2, dH1, kC2, 0, S + L.y, kC, 0, dprm[10], kC, 0, kC2, kC, M, 1, 1.f, nullptr, 0, splits, st)) return rc;
  if (int rc = colsum(dH1, kC2, M, kC2, dprm[11], st)) return rc;
  if (int rc = sgemm(1, dH1, kC2, 0, prm[10], kC, 0, dY, kC, 0, M, kC, kC2, 1, 1.f, nullptr, 0, 1, st)) return rc;
  // LayerNorm
  if (int rc = ln_backward(dY, S + L.o, B, P, S + L.ln_wt, S + L.ln_mu, S + L.ln_rstd, ln_part, dprm[8], dprm[9], st)) return rc;   // dY -> dO
  // attention: O = P V;  S = K Q^T / 16
  const float* Q = S + L.qkv;
  const float* K = S + L.qkv + kC;
  const float* V = S + L.qkv + 2 * kC;
  float* dQ = dQKV;
  float* dK = dQKV + kC;
  float* dV = dQKV + 2 * kC;
  const int64_t sq = (int64_t)P * 3 * kC;
  if (int rc = sgemm(0, dY, kC, (int64_t)P * kC, V, 3 * kC, sq, dP, P, (int64_t)P * P, P, P, kC, B, 1.f, nullptr, 0, 1, st)) return rc;     // dP = dO V^T
  if (int rc = sgemm(2, S + L.prob, P, (int64_t)P * P, dY, kC, (int64_t)P * kC, dV, 3 * kC, sq, P, kC, P, B, 1.f, nullptr, 0, 1, st)) return rc;   // dV = P^T dO
  softmax_rows_bwd_kernel<<<ceil_div(B * P, 8), 256, 0, st>>>(S + L.prob, dP, B * P, P, 1.f / 16.f);          // dP -> dS
  TGFR_LAUNCH_OK();
  if (int rc = sgemm(1, dP, P, (int64_t)P * P, Q, 3 * kC, sq, dK, 3 * kC, sq, P, kC, P, B, 1.f, nullptr, 0, 1, st)) return rc;   // dK = dS Q
  if (int rc = sgemm(2, dP, P, (int64_t)P * P, K, 3 * kC, sq, dQ, 3 * kC, sq, P, kC, P, B, 1.f, nullptr, 0, 1, st)) return rc;   // dQ = dS^T K
  // projections: d W_qkv [768,256] = dQKV^T xn, biases, d xn = dQKV W_qkv
  if (int rc = sgemm(2, dQKV, 3 * kC, 0, S + L.xn, kC, 0, dwqkv, kC, 0, 3 * kC, kC, M, 1, 1.f, nullptr, 0, splits, st)) return rc;
  if (int rc = colsum(dQKV, 3 * kC, M, 3 * kC, dbqkv, st)) return rc;
  for (int k = 0; k < 3; ++k) {
    TGFR_CUDA_OK(cudaMemcpyAsync(dprm[2 + 2 * k], dwqkv + (size_t)k * kC * kC, sizeof(float) * kC * kC, cudaMemcpyDeviceToDevice, st));
    TGFR_CUDA_OK(cudaMemcpyAsync(dprm[3 + 2 * k], dbqkv + (size_t)k * kC, sizeof(float) * kC, cudaMemcpyDeviceToDevice, st));
  }
  if (int rc = sgemm(1, dQKV, 3 * kC, 0, S + L.wqkv, kC, 0, dxn, kC, 0, M, kC, 3 * kC, 1, 1.f, nullptr, 0, 1, st)) return rc;
  // BatchNorm: xn = gamma xhat + beta
  if (int rc = bn_backward_sums(dxn, x, sb, sc, sp, B, P, S + L.mean, S + L.invstd, dprm[0], dprm[1], st)) return rc;
  if (dx) {
    // d xhat = dxn gamma; the sums above are over dxn (not dxn gamma): scale inside the kernel
    bn_bwd_dx_kernel<<<dim3(ceil_div(P, 32), kC / 32, B), dim3(32, 8), 0, st>>>(dxn, x, sb, sc, sp, B, kC, P, S + L.mean, S + L.invstd,
                                                                               prm[0], dprm[0], dprm[1], training, dx, (int64_t)kC * P, P, 1);
    TGFR_LAUNCH_OK();
  }
  return TGFR_OK;
}

// ProjectionHead (models/models.py:96-119): out = normalize(x W^T + b) for x [M,K], W [N,K]; znorm [M] for the backward
int proj_head_fwd(const float* x, int64_t ldx, const float* w, const float* b, int M, int N, int K, float* out, float* znorm,
                  cudaStream_t st) {
  if (int rc = sgemm(0, x, ldx, 0, w, K, 0, out, N, 0, M, N, K, 1, 1.f, b, 0, 1, st)) return rc;
  l2norm_rows_kernel<<<ceil_div(M, 8), 256, 0, st>>>(out, M, N, out, znorm);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}
// dz scratch [M,N]; dx [M,K] (may be NULL), dw [N,K], db [N]
int proj_head_bwd(const float* gout, const float* out, const float* znorm, const float* x, int64_t ldx, const float* w, int M,
                  int N, int K, float* dz, float* dx, float* dw, float* db, cudaStream_t st) {
  l2norm_rows_bwd_kernel<<<ceil_div(M, 8), 256, 0, st>>>(gout, out, znorm, M, N, dz);
  TGFR_LAUNCH_OK();
  const int splits = M >= 4096 ? 32 : (M >= 512 ? 8 : 1);
  if (int rc = sgemm(2, dz, N, 0, x, ldx, 0, dw, K, 0, N, K, M, 1, 1.f, nullptr, 0, splits, st)) return rc;
  if (int rc = colsum(dz, N, M, N, db, st)) return rc;
  if (dx) return sgemm(1, dz, N, 0, w, K, 0, dx, K, 0, M, K, N, 1, 1.f, nullptr, 0, 1, st);
  return TGFR_OK;
}

}  // namespace tgfr

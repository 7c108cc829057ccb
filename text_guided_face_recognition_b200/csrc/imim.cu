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
// element-wise kernels; the contractions run through ONE register-blocked fp32 kernel (sgemm_kernel: NT / NN / TN,
// strided batch, split-K with atomics for the weight gradients whose K is B*196).  fp32 throughout: the output is the
// unit-norm operand of the word-region scores and its gradient trains image_head (src/train_encoders_bert.py:263-265).
#include "common.cuh"
#include "nn_blocks.cuh"

namespace tgfr {
namespace {

constexpr int kC = 256, kC2 = 128;

// saved / workspace layout (floats), M = B * P
struct ImimLayout {
  size_t mean, invstd, xn, qkv, prob, o, ln_mu, ln_rstd, y, h1, h2, znorm, wqkv, bqkv, total;
};
ImimLayout imim_layout(int B, int P) {
  ImimLayout L;
  const size_t M = (size_t)B * P;
  size_t off = 0;
  auto take = [&](size_t n) { const size_t o = off; off += (n + 63) / 64 * 64; return o; };
  L.mean = take(kC); L.invstd = take(kC); L.xn = take(M * kC); L.qkv = take(M * 3 * kC); L.prob = take((size_t)B * P * P);
  L.o = take(M * kC); L.ln_mu = take(B); L.ln_rstd = take(B); L.y = take(M * kC); L.h1 = take(M * kC2); L.h2 = take(M * kC);
  L.znorm = take(M); L.wqkv = take((size_t)3 * kC * kC); L.bqkv = take(3 * kC);
  L.total = off;
  return L;
}

}  // namespace

// parameter order of tgfr_imim_fwd / dparams of tgfr_imim_bwd (the reference's state_dict names):
//  0 bn_img.weight  1 bn_img.bias  2 sa.query_proj.weight [256,256]  3 .bias  4 sa.key_proj.weight  5 .bias
//  6 sa.value_proj.weight  7 .bias  8 ln.weight [256*196]  9 ln.bias  10 conv1x1_1.weight [128,256]  11 .bias
//  12 conv1x1_2.weight [256,128]  13 .bias  14 project_local.projection.weight [256,256]  15 .bias
constexpr int kImimParams = 16;

size_t imim_saved_bytes(int B, int P) { return imim_layout(B, P).total * sizeof(float); }
size_t imim_workspace_bytes(int B, int P) {
  // backward scratch: dZ/dH2 [M,256], dH1 [M,128], dY/dO [M,256], dP [B,P,P], dQKV [M,768], dxn [M,256]
  const size_t M = (size_t)B * P;
  return (M * kC * 3 + M * kC2 + (size_t)B * P * P + M * 3 * kC + (size_t)3 * kC * kC + 3 * kC + 1024) * sizeof(float);
}

int imim_fwd(const float* x, int64_t sb, int64_t sc, int64_t sp, const float* const* prm, int B, int P, int training,
             float momentum, float eps, float* run_mean, float* run_var, float* out, void* saved, size_t saved_bytes,
             cudaStream_t st) {
  TGFR_REQUIRE(B >= 1 && P >= 1, "imim_fwd: empty batch");
  const ImimLayout L = imim_layout(B, P);
  TGFR_REQUIRE(saved && saved_bytes >= L.total * sizeof(float), "imim_fwd: saved buffer too small (%zu < %zu)", saved_bytes,
               L.total * sizeof(float));
  float* S = reinterpret_cast<float*>(saved);
  const int M = B * P;
  bn_stats_kernel<<<kC, 256, 0, st>>>(x, sb, sc, sp, B, P, eps, momentum, training, run_mean, run_var, S + L.mean, S + L.invstd);
  TGFR_LAUNCH_OK();
  bn_apply_t_kernel<<<dim3(ceil_div(P, 32), kC / 32, B), dim3(32, 8), 0, st>>>(x, sb, sc, sp, kC, P, S + L.mean, S + L.invstd,
                                                                              prm[0], prm[1], S + L.xn);
  TGFR_LAUNCH_OK();
  // one product for the three projections: W_qkv [768,256] = [Wq; Wk; Wv]
  for (int k = 0; k < 3; ++k) {
    TGFR_CUDA_OK(cudaMemcpyAsync(S + L.wqkv + (size_t)k * kC * kC, prm[2 + 2 * k], sizeof(float) * kC * kC, cudaMemcpyDeviceToDevice, st));
    TGFR_CUDA_OK(cudaMemcpyAsync(S + L.bqkv + (size_t)k * kC, prm[3 + 2 * k], sizeof(float) * kC, cudaMemcpyDeviceToDevice, st));
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
  ln_fwd_kernel<<<B, 1024, 0, st>>>(S + L.o, (int64_t)P * kC, P, kC, prm[8], prm[9], S + L.y, (int64_t)P * kC, S + L.ln_mu, S + L.ln_rstd);
  TGFR_LAUNCH_OK();
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
  TGFR_REQUIRE(saved && saved_bytes >= L.total * sizeof(float), "imim_bwd: saved buffer too small");
  TGFR_REQUIRE(ws && ws_bytes >= imim_workspace_bytes(B, P), "imim_bwd: workspace too small");
  const float* S = reinterpret_cast<const float*>(saved);
  const int M = B * P;
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
  float* dxn = dZ;                            // [M,256]  (dZ is dead by then)
  const int splits = M >= 4096 ? 32 : (M >= 512 ? 8 : 1);      // weight gradients: K = M rows
  // projection + L2 norm
  l2norm_rows_bwd_kernel<<<ceil_div(M, 8), 256, 0, st>>>(gout, out, S + L.znorm, M, kC, dZ);
  TGFR_LAUNCH_OK();
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
  ln_bwd_params_kernel<<<ceil_div(P * kC, 256), 256, 0, st>>>(dY, (int64_t)P * kC, S + L.o, (int64_t)P * kC, B, P, kC, S + L.ln_mu, S + L.ln_rstd, dprm[8], dprm[9]);
  TGFR_LAUNCH_OK();
  ln_bwd_dx_kernel<<<B, 1024, 0, st>>>(dY, (int64_t)P * kC, S + L.o, (int64_t)P * kC, P, kC, prm[8], S + L.ln_mu, S + L.ln_rstd, dY, (int64_t)P * kC);     // dY -> dO
  TGFR_LAUNCH_OK();
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
  bn_bwd_sums_kernel<<<kC, 256, 0, st>>>(dxn, x, sb, sc, sp, B, kC, P, S + L.mean, S + L.invstd, dprm[0], dprm[1]);
  TGFR_LAUNCH_OK();
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

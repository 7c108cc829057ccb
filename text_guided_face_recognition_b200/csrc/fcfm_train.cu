// FCFM fusion net `Working` (reference models/fusion_nets.py:217-258), TRAINING-mode forward and backward
// (the fusion training step src/fusion_bert.py:205-233 back-propagates through it).
// The evaluation forward is the one-launch kernel of csrc/fcfm.cu; BatchNorm with batch statistics couples the samples,
// so the training path is a sequence of batch-wide launches built from csrc/nn_blocks.cuh:
//
//   img [B,256,14,14] -> im2col -> relu(col Wc^T + bc) [B*144, 36] -> 2x2 max-pool -> BatchNorm(36) -> xni [B*36, 36]
//   word [B,256,T] -> Linear(256 -> 36) per word -> gram w^T w / 6 [B,36,36] (viewed [B,36,6,6]) -> BatchNorm(36) -> xnw
//   q = Wq xnw, k = Wk xni, v = Wv xni;  attention = softmax_j(k_i . q_j / 6);  response = attention v
//   -> LayerNorm([36,6,6]) -> 2x2 max-pool -> flatten [B,324] -> Linear(324 -> 128) -> out[:, 0:128]
//   LayerNorm(256) of the global image feature -> out[:, 128:384];  LayerNorm(256) of the sentence feature -> out[:, 384:640]
//
// Activations are position-major ([B * positions, channels]).  Gradients go to all 26 parameter tensors and the four inputs.
#include "common.cuh"
#include "nn_blocks.cuh"

namespace tgfr {
namespace {

constexpr int kCh = 36;        // channel_dim the reference hard-codes (fusion_nets.py:220)
constexpr int kPos = 36;       // 6 x 6 positions
constexpr int kIn = 256;
constexpr int kK9 = kIn * 9;   // im2col row length

// col[(b*144 + oy*12 + ox), c*9 + ky*3 + kx] = img[b, c, oy+ky, ox+kx]
__global__ void im2col3_kernel(const float* __restrict__ img, int64_t sb, int64_t sc, int64_t sh, int64_t sw, int B,
                               float* __restrict__ col) {
  const int64_t total = (int64_t)B * 144 * kK9;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(k % kK9);
    const int64_t row = k / kK9;
    const int b = (int)(row / 144), pos = (int)(row - (int64_t)b * 144);
    const int oy = pos / 12, ox = pos - oy * 12;
    const int c = j / 9, t = j - c * 9, ky = t / 3, kx = t - ky * 3;
    col[k] = img[b * sb + c * sc + (oy + ky) * sh + (ox + kx) * sw];
  }
}
// dimg[b,c,y,x] (contiguous) = sum_{ky,kx} dcol[(b, y-ky, x-kx), c*9 + ky*3 + kx]   (gather: no atomics)
__global__ void col2im3_kernel(const float* __restrict__ dcol, int B, float* __restrict__ dimg) {
  const int64_t total = (int64_t)B * kIn * 196;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(k % 14), y = (int)((k / 14) % 14), c = (int)((k / 196) % kIn), b = (int)(k / (196 * kIn));
    float s = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int oy = y - ky;
      if (oy < 0 || oy >= 12) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ox = x - kx;
        if (ox < 0 || ox >= 12) continue;
        s += dcol[((int64_t)b * 144 + oy * 12 + ox) * kK9 + c * 9 + ky * 3 + kx];
      }
    }
    dimg[k] = s;
  }
}

// 2x2 max-pool of position-major x [B, H*W, C] (H, W even) -> y, first maximum wins (torch's rule);
// channel_major_out: y[b, c*(H/2*W/2) + q] (the flatten order of a [B,C,h,w] tensor), else y[b, q, c]
__global__ void maxpool2_fwd_kernel(const float* __restrict__ x, int B, int H, int W, int C, int channel_major_out,
                                    float* __restrict__ y, uint8_t* __restrict__ idx) {
  const int h2 = H / 2, w2 = W / 2, Q = h2 * w2;
  const int64_t total = (int64_t)B * Q * C;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(k % C), q = (int)((k / C) % Q), b = (int)(k / ((int64_t)C * Q));
    const int qy = q / w2, qx = q - qy * w2;
    float best = -INFINITY;
    int bi = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int pp = (2 * qy + (t >> 1)) * W + 2 * qx + (t & 1);
      const float v = x[((int64_t)b * H * W + pp) * C + c];
      if (t == 0 || v > best) {
        best = v;
        bi = t;
      }
    }
    idx[k] = (uint8_t)bi;
    if (channel_major_out) y[(int64_t)b * C * Q + c * Q + q] = best;
    else y[k] = best;
  }
}
// dx (position-major [B, H*W, C], fully written) from dy in the layout maxpool2_fwd_kernel produced
__global__ void maxpool2_bwd_kernel(const float* __restrict__ dy, int64_t dy_bstride, const uint8_t* __restrict__ idx, int B,
                                    int H, int W, int C, int channel_major_in, float* __restrict__ dx) {
  const int h2 = H / 2, w2 = W / 2, Q = h2 * w2;
  const int64_t total = (int64_t)B * Q * C;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(k % C), q = (int)((k / C) % Q), b = (int)(k / ((int64_t)C * Q));
    const int qy = q / w2, qx = q - qy * w2;
    const float g = channel_major_in ? dy[(int64_t)b * dy_bstride + c * Q + q] : dy[(int64_t)b * dy_bstride + (int64_t)q * C + c];
    const int bi = idx[k];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int pp = (2 * qy + (t >> 1)) * W + 2 * qx + (t & 1);
      dx[((int64_t)b * H * W + pp) * C + c] = (t == bi) ? g : 0.f;
    }
  }
}
__global__ void transpose2d_kernel(const float* __restrict__ a, int rows, int cols, float* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= rows * cols) return;
  const int r = k / cols, c = k - r * cols;
  out[c * rows + r] = a[k];
}
// s[b] = g[b] + g[b]^T for B matrices of n x n
__global__ void symmetrize_kernel(const float* __restrict__ g, int B, int n, float* __restrict__ s) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (int64_t)B * n * n) return;
  const int j = (int)(k % n), i = (int)((k / n) % n);
  const int64_t b = k / (n * n);
  s[k] = g[k] + g[(b * n + j) * n + i];
}
int blocks_for(int64_t n) { return (int)((n + 255) / 256 < 4736 ? (n + 255) / 256 : 4736); }

struct FcfmLayout {      // offsets in floats
  size_t col, yc, xi, bni, xni, wprojT, wproj, gram, bnw, xnw, q, k, v, prob, o, lnmu, lnrstd, y, pooled, glmu, glrstd, smu,
      srstd, idx1, idx2, total;
};
FcfmLayout fcfm_layout(int B, int T) {
  FcfmLayout L;
  size_t off = 0;
  auto take = [&](size_t n) { const size_t o = off; off += (n + 63) / 64 * 64; return o; };
  const size_t M1 = (size_t)B * 144, M = (size_t)B * kPos;
  L.col = take(M1 * kK9); L.yc = take(M1 * kCh); L.xi = take(M * kCh); L.bni = take(2 * kCh); L.xni = take(M * kCh);
  L.wprojT = take((size_t)kIn * kCh); L.wproj = take((size_t)B * T * kCh); L.gram = take(M * kCh); L.bnw = take(2 * kCh);
  L.xnw = take(M * kCh); L.q = take(M * kCh); L.k = take(M * kCh); L.v = take(M * kCh); L.prob = take((size_t)B * kPos * kPos);
  L.o = take(M * kCh); L.lnmu = take(B); L.lnrstd = take(B); L.y = take(M * kCh); L.pooled = take((size_t)B * 324);
  L.glmu = take(B); L.glrstd = take(B); L.smu = take(B); L.srstd = take(B);
  L.idx1 = take((M * kCh + 3) / 4); L.idx2 = take(((size_t)B * 324 + 3) / 4);
  L.total = off;
  return L;
}

}  // namespace

// parameter order (the reference's state_dict names; running statistics are separate arguments):
//  0 conv.weight [36,2304]  1 conv.bias  2 bn_img.weight  3 bn_img.bias  4 projection.weight [36,256]  5 projection.bias
//  6 bn_word.weight  7 bn_word.bias  8 sa.query_proj.weight [36,36]  9 .bias  10 sa.key_proj.weight  11 .bias
//  12 sa.value_proj.weight  13 .bias  14 ln.weight [36*36]  15 ln.bias  16 linear.weight [128,324]  17 linear.bias
//  18 ln_gl_image.weight [256]  19 .bias  20 ln_sent.weight [256]  21 .bias
size_t fcfm_train_saved_bytes(int B, int T) { return fcfm_layout(B, T).total * sizeof(float); }
size_t fcfm_train_workspace_bytes(int B, int T) {
  const size_t M1 = (size_t)B * 144, M = (size_t)B * kPos;
  // dcol [M1,2304], dyc [M1,36], 8 x [M,36], dprob [B,36,36], dwproj [B*T,36], dpooled [B,324], dwpT [256,36]
  return (M1 * kK9 + M1 * kCh + 8 * M * kCh + (size_t)B * kPos * kPos + (size_t)B * T * kCh + (size_t)B * 324 + (size_t)kIn * kCh +
          4096) * sizeof(float);
}

int fcfm_train_fwd(const float* img, int64_t isb, int64_t isc, int64_t ish, int64_t isw, const float* word, int64_t wsb,
                   int64_t wsc, const float* gl, int64_t gsr, const float* sent, int64_t ssr, const float* const* prm, int B,
                   int T, int training, float momentum, float eps, float* const* run_stats, float* out, int64_t out_sr,
                   void* saved, size_t saved_bytes, cudaStream_t st) {
  const FcfmLayout L = fcfm_layout(B, T);
  TGFR_REQUIRE(saved && saved_bytes >= L.total * sizeof(float), "fcfm_train_fwd: saved buffer too small");
  float* S = reinterpret_cast<float*>(saved);
  uint8_t* idx1 = reinterpret_cast<uint8_t*>(S + L.idx1);
  uint8_t* idx2 = reinterpret_cast<uint8_t*>(S + L.idx2);
  const int M1 = B * 144, M = B * kPos;
  // image branch: conv3x3 + relu (fusion_nets.py:235), 2x2 max-pool, BatchNorm (:236)
  im2col3_kernel<<<blocks_for((int64_t)M1 * kK9), 256, 0, st>>>(img, isb, isc, ish, isw, B, S + L.col);
  TGFR_LAUNCH_OK();
  if (int rc = sgemm(0, S + L.col, kK9, 0, prm[0], kK9, 0, S + L.yc, kCh, 0, M1, kCh, kK9, 1, 1.f, prm[1], 1, 1, st)) return rc;
  maxpool2_fwd_kernel<<<blocks_for((int64_t)M * kCh), 256, 0, st>>>(S + L.yc, B, 12, 12, kCh, 0, S + L.xi, idx1);
  TGFR_LAUNCH_OK();
  bn_stats_kernel<<<kCh, 256, 0, st>>>(S + L.xi, (int64_t)kPos * kCh, 1, kCh, B, kPos, eps, momentum, training, run_stats[0],
                                      run_stats[1], S + L.bni, S + L.bni + kCh);
  TGFR_LAUNCH_OK();
  bn_apply_t_kernel<<<dim3(ceil_div(kPos, 32), ceil_div(kCh, 32), B), dim3(32, 8), 0, st>>>(
      S + L.xi, (int64_t)kPos * kCh, 1, kCh, kCh, kPos, S + L.bni, S + L.bni + kCh, prm[2], prm[3], S + L.xni);
  TGFR_LAUNCH_OK();
  // word branch: projection (:239), gram / sqrt(36) (:240), BatchNorm (:242)
  transpose2d_kernel<<<ceil_div(kCh * kIn, 256), 256, 0, st>>>(prm[4], kCh, kIn, S + L.wprojT);
  TGFR_LAUNCH_OK();
  if (int rc = sgemm(2, word, wsc, wsb, S + L.wprojT, kCh, 0, S + L.wproj, kCh, (int64_t)T * kCh, T, kCh, kIn, B, 1.f, prm[5], 0, 1, st))
    return rc;
  if (int rc = sgemm(2, S + L.wproj, kCh, (int64_t)T * kCh, S + L.wproj, kCh, (int64_t)T * kCh, S + L.gram, kCh, (int64_t)kPos * kCh,
                     kCh, kCh, T, B, 1.f / 6.f, nullptr, 0, 1, st)) return rc;
  bn_stats_kernel<<<kCh, 256, 0, st>>>(S + L.gram, (int64_t)kPos * kCh, 1, kCh, B, kPos, eps, momentum, training, run_stats[2],
                                      run_stats[3], S + L.bnw, S + L.bnw + kCh);
  TGFR_LAUNCH_OK();
  bn_apply_t_kernel<<<dim3(ceil_div(kPos, 32), ceil_div(kCh, 32), B), dim3(32, 8), 0, st>>>(
      S + L.gram, (int64_t)kPos * kCh, 1, kCh, kCh, kPos, S + L.bnw, S + L.bnw + kCh, prm[6], prm[7], S + L.xnw);
  TGFR_LAUNCH_OK();
  // image-text attention (:247; SelfAttention fusion_nets.py:92-117 with x = image, y = words, scale = 1)
  if (int rc = sgemm(0, S + L.xnw, kCh, 0, prm[8], kCh, 0, S + L.q, kCh, 0, M, kCh, kCh, 1, 1.f, prm[9], 0, 1, st)) return rc;
  if (int rc = sgemm(0, S + L.xni, kCh, 0, prm[10], kCh, 0, S + L.k, kCh, 0, M, kCh, kCh, 1, 1.f, prm[11], 0, 1, st)) return rc;
  if (int rc = sgemm(0, S + L.xni, kCh, 0, prm[12], kCh, 0, S + L.v, kCh, 0, M, kCh, kCh, 1, 1.f, prm[13], 0, 1, st)) return rc;
  const int64_t sm = (int64_t)kPos * kCh;
  if (int rc = sgemm(0, S + L.k, kCh, sm, S + L.q, kCh, sm, S + L.prob, kPos, (int64_t)kPos * kPos, kPos, kPos, kCh, B, 1.f / 6.f, nullptr,
                     0, 1, st)) return rc;
  softmax_rows_kernel<<<ceil_div(M, 8), 256, 0, st>>>(S + L.prob, M, kPos);
  TGFR_LAUNCH_OK();
  if (int rc = sgemm(1, S + L.prob, kPos, (int64_t)kPos * kPos, S + L.v, kCh, sm, S + L.o, kCh, sm, kPos, kCh, kPos, B, 1.f, nullptr, 0, 1, st))
    return rc;
  // LayerNorm([36,6,6]) (:248), 2x2 max-pool + flatten (:249-250), Linear(324 -> 128) (:254)
  ln_fwd_kernel<<<B, 256, 0, st>>>(S + L.o, sm, kPos, kCh, prm[14], prm[15], S + L.y, sm, S + L.lnmu, S + L.lnrstd);
  TGFR_LAUNCH_OK();
  maxpool2_fwd_kernel<<<blocks_for((int64_t)B * 324), 256, 0, st>>>(S + L.y, B, 6, 6, kCh, 1, S + L.pooled, idx2);
  TGFR_LAUNCH_OK();
  if (int rc = sgemm(0, S + L.pooled, 324, 0, prm[16], 324, 0, out, out_sr, 0, B, 128, 324, 1, 1.f, prm[17], 0, 1, st)) return rc;
  // LayerNorm(256) of the two global features (:255-256), concatenated behind the fused part (:257)
  ln_fwd_kernel<<<B, 256, 0, st>>>(gl, gsr, 1, kIn, prm[18], prm[19], out + 128, out_sr, S + L.glmu, S + L.glrstd);
  TGFR_LAUNCH_OK();
  ln_fwd_kernel<<<B, 256, 0, st>>>(sent, ssr, 1, kIn, prm[20], prm[21], out + 384, out_sr, S + L.smu, S + L.srstd);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int fcfm_train_bwd(const float* gout, int64_t g_sr, const float* word, int64_t wsb, int64_t wsc, const float* gl, int64_t gsr,
                   const float* sent, int64_t ssr, const float* const* prm, int B, int T, int training, const void* saved,
                   size_t saved_bytes, float* const* dprm, float* dimg, float* dword, float* dgl, float* dsent, void* ws,
                   size_t ws_bytes, cudaStream_t st) {
  const FcfmLayout L = fcfm_layout(B, T);
  TGFR_REQUIRE(saved && saved_bytes >= L.total * sizeof(float), "fcfm_train_bwd: saved buffer too small");
  TGFR_REQUIRE(ws && ws_bytes >= fcfm_train_workspace_bytes(B, T), "fcfm_train_bwd: workspace too small");
  const float* S = reinterpret_cast<const float*>(saved);
  const uint8_t* idx1 = reinterpret_cast<const uint8_t*>(S + L.idx1);
  const uint8_t* idx2 = reinterpret_cast<const uint8_t*>(S + L.idx2);
  const int M1 = B * 144, M = B * kPos;
  const size_t Mz = (size_t)M * kCh;
  float* W = reinterpret_cast<float*>(ws);
  float* dcol = W;                                   W += (size_t)M1 * kK9;
  float* dyc = W;                                    W += (size_t)M1 * kCh;
  float* dy = W;                                     W += Mz;          // dy -> do
  float* dq = W;                                     W += Mz;
  float* dk = W;                                     W += Mz;
  float* dv = W;                                     W += Mz;
  float* dxnw = W;                                   W += Mz;          // -> dgram (after BatchNorm backward)
  float* dxni = W;                                   W += Mz;          // -> dxi
  float* tmp = W;                                    W += Mz;
  float* dsym = W;                                   W += Mz;
  float* dprob = W;                                  W += (size_t)B * kPos * kPos;
  float* dwproj = W;                                 W += (size_t)B * T * kCh;
  float* dpooled = W;                                W += (size_t)B * 324;
  float* dwpT = W;                                   W += (size_t)kIn * kCh;
  const int64_t sm = (int64_t)kPos * kCh;
  const int splitM = M >= 4096 ? 16 : (M >= 512 ? 4 : 1);
  // the two global LayerNorms
  ln_bwd_params_kernel<<<1, 256, 0, st>>>(gout + 128, g_sr, gl, gsr, B, 1, kIn, S + L.glmu, S + L.glrstd, dprm[18], dprm[19]);
  TGFR_LAUNCH_OK();
  ln_bwd_params_kernel<<<1, 256, 0, st>>>(gout + 384, g_sr, sent, ssr, B, 1, kIn, S + L.smu, S + L.srstd, dprm[20], dprm[21]);
  TGFR_LAUNCH_OK();
  if (dgl) {
    ln_bwd_dx_kernel<<<B, 256, 0, st>>>(gout + 128, g_sr, gl, gsr, 1, kIn, prm[18], S + L.glmu, S + L.glrstd, dgl, kIn);
    TGFR_LAUNCH_OK();
  }
  if (dsent) {
    ln_bwd_dx_kernel<<<B, 256, 0, st>>>(gout + 384, g_sr, sent, ssr, 1, kIn, prm[20], S + L.smu, S + L.srstd, dsent, kIn);
    TGFR_LAUNCH_OK();
  }
  // linear (:254)
  if (int rc = sgemm(2, gout, g_sr, 0, S + L.pooled, 324, 0, dprm[16], 324, 0, 128, 324, B, 1, 1.f, nullptr, 0, 1, st)) return rc;
  if (int rc = colsum(gout, g_sr, B, 128, dprm[17], st)) return rc;
  if (int rc = sgemm(1, gout, g_sr, 0, prm[16], 324, 0, dpooled, 324, 0, B, 324, 128, 1, 1.f, nullptr, 0, 1, st)) return rc;
  // max-pool, LayerNorm
  maxpool2_bwd_kernel<<<blocks_for((int64_t)B * 324), 256, 0, st>>>(dpooled, 324, idx2, B, 6, 6, kCh, 1, dy);
  TGFR_LAUNCH_OK();
  ln_bwd_params_kernel<<<ceil_div(kPos * kCh, 256), 256, 0, st>>>(dy, sm, S + L.o, sm, B, kPos, kCh, S + L.lnmu, S + L.lnrstd, dprm[14],
                                                                  dprm[15]);
  TGFR_LAUNCH_OK();
  ln_bwd_dx_kernel<<<B, 256, 0, st>>>(dy, sm, S + L.o, sm, kPos, kCh, prm[14], S + L.lnmu, S + L.lnrstd, dy, sm);     // dy -> do
  TGFR_LAUNCH_OK();
  // attention: o = P v;  S = k q^T / 6
  if (int rc = sgemm(0, dy, kCh, sm, S + L.v, kCh, sm, dprob, kPos, (int64_t)kPos * kPos, kPos, kPos, kCh, B, 1.f, nullptr, 0, 1, st)) return rc;
  if (int rc = sgemm(2, S + L.prob, kPos, (int64_t)kPos * kPos, dy, kCh, sm, dv, kCh, sm, kPos, kCh, kPos, B, 1.f, nullptr, 0, 1, st)) return rc;
  softmax_rows_bwd_kernel<<<ceil_div(M, 8), 256, 0, st>>>(S + L.prob, dprob, M, kPos, 1.f / 6.f);
  TGFR_LAUNCH_OK();
  if (int rc = sgemm(1, dprob, kPos, (int64_t)kPos * kPos, S + L.q, kCh, sm, dk, kCh, sm, kPos, kCh, kPos, B, 1.f, nullptr, 0, 1, st)) return rc;
  if (int rc = sgemm(2, dprob, kPos, (int64_t)kPos * kPos, S + L.k, kCh, sm, dq, kCh, sm, kPos, kCh, kPos, B, 1.f, nullptr, 0, 1, st)) return rc;
  // 1x1 projections
  if (int rc = sgemm(2, dq, kCh, 0, S + L.xnw, kCh, 0, dprm[8], kCh, 0, kCh, kCh, M, 1, 1.f, nullptr, 0, splitM, st)) return rc;
  if (int rc = colsum(dq, kCh, M, kCh, dprm[9], st)) return rc;
  if (int rc = sgemm(2, dk, kCh, 0, S + L.xni, kCh, 0, dprm[10], kCh, 0, kCh, kCh, M, 1, 1.f, nullptr, 0, splitM, st)) return rc;
  if (int rc = colsum(dk, kCh, M, kCh, dprm[11], st)) return rc;
  if (int rc = sgemm(2, dv, kCh, 0, S + L.xni, kCh, 0, dprm[12], kCh, 0, kCh, kCh, M, 1, 1.f, nullptr, 0, splitM, st)) return rc;
  if (int rc = colsum(dv, kCh, M, kCh, dprm[13], st)) return rc;
  if (int rc = sgemm(1, dq, kCh, 0, prm[8], kCh, 0, dxnw, kCh, 0, M, kCh, kCh, 1, 1.f, nullptr, 0, 1, st)) return rc;
  if (int rc = sgemm(1, dk, kCh, 0, prm[10], kCh, 0, dxni, kCh, 0, M, kCh, kCh, 1, 1.f, nullptr, 0, 1, st)) return rc;
  if (int rc = sgemm(1, dv, kCh, 0, prm[12], kCh, 0, dxni, kCh, 0, M, kCh, kCh, 1, 1.f, nullptr, 0, 1, st, 1)) return rc;   // +=
  // word branch: BatchNorm, gram, projection
  bn_bwd_sums_kernel<<<kCh, 256, 0, st>>>(dxnw, S + L.gram, sm, 1, kCh, B, kCh, kPos, S + L.bnw, S + L.bnw + kCh, dprm[6], dprm[7]);
  TGFR_LAUNCH_OK();
  bn_bwd_dx_kernel<<<dim3(ceil_div(kPos, 32), ceil_div(kCh, 32), B), dim3(32, 8), 0, st>>>(
      dxnw, S + L.gram, sm, 1, kCh, B, kCh, kPos, S + L.bnw, S + L.bnw + kCh, prm[6], dprm[6], dprm[7], training, tmp, sm, 1, kCh);
  TGFR_LAUNCH_OK();
  symmetrize_kernel<<<ceil_div(M * kCh, 256), 256, 0, st>>>(tmp, B, kCh, dsym);
  TGFR_LAUNCH_OK();
  if (int rc = sgemm(1, S + L.wproj, kCh, (int64_t)T * kCh, dsym, kCh, sm, dwproj, kCh, (int64_t)T * kCh, T, kCh, kCh, B, 1.f / 6.f, nullptr, 0,
                     1, st)) return rc;
  // d projection.weight^T [256,36] = sum_b word[b] [256,T] dwproj[b] [T,36];  bias;  d word[b] [256,T] = W^T dwproj[b]^T
  if (int rc = sgemm(1, word, wsc, wsb, dwproj, kCh, (int64_t)T * kCh, dwpT, kCh, 0, kIn, kCh, T, B, 1.f, nullptr, 0, 1, st)) return rc;
  transpose2d_kernel<<<ceil_div(kCh * kIn, 256), 256, 0, st>>>(dwpT, kIn, kCh, dprm[4]);
  TGFR_LAUNCH_OK();
  if (int rc = colsum(dwproj, kCh, B * T, kCh, dprm[5], st)) return rc;
  if (dword)
    if (int rc = sgemm(0, S + L.wprojT, kCh, 0, dwproj, kCh, (int64_t)T * kCh, dword, T, (int64_t)kIn * T, kIn, T, kCh, B, 1.f, nullptr, 0, 1,
                       st)) return rc;
  // image branch: BatchNorm, max-pool, relu, convolution
  bn_bwd_sums_kernel<<<kCh, 256, 0, st>>>(dxni, S + L.xi, sm, 1, kCh, B, kCh, kPos, S + L.bni, S + L.bni + kCh, dprm[2], dprm[3]);
  TGFR_LAUNCH_OK();
  bn_bwd_dx_kernel<<<dim3(ceil_div(kPos, 32), ceil_div(kCh, 32), B), dim3(32, 8), 0, st>>>(
      dxni, S + L.xi, sm, 1, kCh, B, kCh, kPos, S + L.bni, S + L.bni + kCh, prm[2], dprm[2], dprm[3], training, tmp, sm, 1, kCh);
  TGFR_LAUNCH_OK();
  maxpool2_bwd_kernel<<<blocks_for((int64_t)M * kCh), 256, 0, st>>>(tmp, sm, idx1, B, 12, 12, kCh, 0, dyc);
  TGFR_LAUNCH_OK();
  relu_mask_kernel<<<blocks_for((int64_t)M1 * kCh), 256, 0, st>>>(dyc, S + L.yc, (int64_t)M1 * kCh);
  TGFR_LAUNCH_OK();
  const int split1 = M1 >= 8192 ? 32 : (M1 >= 1024 ? 8 : 1);
  if (int rc = sgemm(2, dyc, kCh, 0, S + L.col, kK9, 0, dprm[0], kK9, 0, kCh, kK9, M1, 1, 1.f, nullptr, 0, split1, st)) return rc;
  if (int rc = colsum(dyc, kCh, M1, kCh, dprm[1], st)) return rc;
  if (dimg) {
    if (int rc = sgemm(1, dyc, kCh, 0, prm[0], kK9, 0, dcol, kK9, 0, M1, kK9, kCh, 1, 1.f, nullptr, 0, 1, st)) return rc;
    col2im3_kernel<<<blocks_for((int64_t)B * kIn * 196), 256, 0, st>>>(dcol, B, dimg);
    TGFR_LAUNCH_OK();
  }
  return TGFR_OK;
}

}  // namespace tgfr

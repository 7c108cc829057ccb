// FCFM fusion net `Working`, eval-mode forward (reference models/fusion_nets.py:217-258):
// the producer of the 640-d fused embeddings that verification scoring (scoring.cu) consumes in BASELINE configs[4].
//
//   img  [B,256,14,14] -> conv3x3 (256 -> 36, valid) + ReLU + maxpool2 -> [36,6,6] -> BatchNorm (running statistics)
//   word [B,256,T]     -> Linear(256 -> 36) per word -> gram matrix / 6 -> [36,6,6] -> BatchNorm
//   SelfAttention(image, text): 1x1 query(text) / key(image) / value(image), softmax(key^T query / 6) value
//   -> LayerNorm([36,6,6]) -> maxpool2 -> [324] -> Linear(324 -> 128)
//   out = [ that | LayerNorm(gl_img) | LayerNorm(sent) ]  = [B, 640]
//
// One CTA per sample, everything after the convolution lives in shared memory (a sample's intermediate state is
// 36 x 36 floats per tensor), one launch instead of the reference's ~25.  fp32 throughout: the embeddings feed
// verification decisions.  The convolution is the only real work (12 MFLOP per sample): input channels are staged in
// chunks of 16 (image rows + the matching 36 x 16 x 9 weights), a thread keeps 9 output pixels of two output channels
// in registers, so an image value feeds two FMAs and a weight nine (11 shared-memory loads per 18 FMAs).
#include "common.cuh"

namespace tgfr {
namespace {

constexpr int kFT = 288;                 // 18 output-channel pairs x 16 pixel groups of 9
constexpr int kC = 36, kHW = 36, kCin = 256, kChunk = 16, kPix = 196, kConvPix = 144, kMaxT = 64;
constexpr float kEps = 1e-5f;            // BatchNorm2d / LayerNorm default eps

enum {
  P_CONV_W, P_CONV_B, P_BNI_W, P_BNI_B, P_BNI_M, P_BNI_V, P_PROJ_W, P_PROJ_B, P_BNW_W, P_BNW_B, P_BNW_M, P_BNW_V,
  P_Q_W, P_Q_B, P_K_W, P_K_B, P_V_W, P_V_B, P_LN_W, P_LN_B, P_LIN_W, P_LIN_B, P_LNG_W, P_LNG_B, P_LNS_W, P_LNS_B, P_NUM
};
struct FcfmParams {
  const float* p[P_NUM];
};

constexpr int kSmemFloats = kChunk * kPix + kC * kChunk * 9 + kC * kConvPix + 7 * kC * kHW + kMaxT * kC + 64;

// out[c][p] = bias[c] + sum_c' W[c][c'] in[c'][p]   (a 1x1 convolution over the 36 positions)
__device__ __forceinline__ void conv1x1_36(const float* __restrict__ W, const float* __restrict__ bias, const float* in,
                                           float* out) {
  for (int idx = threadIdx.x; idx < kC * kHW; idx += kFT) {
    const int c = idx / kHW, p = idx - c * kHW;
    float acc = __ldg(bias + c);
    for (int k = 0; k < kC; ++k) acc = fmaf(__ldg(W + c * kC + k), in[k * kHW + p], acc);
    out[idx] = acc;
  }
}

__global__ void __launch_bounds__(kFT) fcfm_working_fwd_kernel(const float* __restrict__ img, int64_t isb, int64_t isc,
                                                               int64_t ish, int64_t isw, const float* __restrict__ word,
                                                               int64_t wsb, int64_t wsd, int64_t wst,
                                                               const float* __restrict__ gl, int64_t gl_sr,
                                                               const float* __restrict__ sent, int64_t se_sr,
                                                               const FcfmParams P, int T, float* __restrict__ out,
                                                               int64_t out_sr) {
  extern __shared__ float sm[];
  float* s_in = sm;                               // [16][196]   image channels of the current chunk
  float* s_w = s_in + kChunk * kPix;              // [36][16][9] their weights
  float* s_conv = s_w + kC * kChunk * 9;          // [36][144]   relu(conv)
  float* s_x = s_conv + kC * kConvPix;            // [36][36]    image branch
  float* s_y = s_x + kC * kHW;                    // [36][36]    text branch
  float* s_q = s_y + kC * kHW;
  float* s_k = s_q + kC * kHW;
  float* s_v = s_k + kC * kHW;
  float* s_att = s_v + kC * kHW;
  float* s_r = s_att + kC * kHW;
  float* s_proj = s_r + kC * kHW;                 // [T][36]
  float* s_red = s_proj + kMaxT * kC;             // [64] reduction scratch
  const int tid = threadIdx.x;
  const int b = blockIdx.x;

  // ---- 1. conv3x3 (valid) + ReLU                                                       fusion_nets.py:235
  // thread = (pair of output channels, group of 9 output pixels): an image value feeds two FMAs, a weight nine
  {
    const int oc0 = (tid >> 4) * 2, sub = tid & 15;
    int off[9];
    float acc0[9], acc1[9];
    const float bias0 = __ldg(P.p[P_CONV_B] + oc0), bias1 = __ldg(P.p[P_CONV_B] + oc0 + 1);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const int p = sub * 9 + i;
      off[i] = (p / 12) * 14 + (p % 12);
      acc0[i] = bias0;
      acc1[i] = bias1;
    }
    const float* ib = img + (int64_t)b * isb;
    for (int c0 = 0; c0 < kCin; c0 += kChunk) {
      __syncthreads();
      for (int idx = tid; idx < kChunk * kPix; idx += kFT) {
        int ci, pix;
        if (isc == 1) {            // channels-last memory (IMIM's output): the channel index is the fast one
          ci = idx & (kChunk - 1);
          pix = idx >> 4;
        } else {
          ci = idx / kPix;
          pix = idx - ci * kPix;
        }
        const int h = pix / 14, w = pix - h * 14;
        s_in[ci * kPix + pix] = __ldg(ib + (int64_t)(c0 + ci) * isc + h * ish + w * isw);
      }
      for (int idx = tid; idx < kC * kChunk * 9; idx += kFT) {
        const int o = idx / (kChunk * 9), r = idx - o * (kChunk * 9);
        s_w[idx] = __ldg(P.p[P_CONV_W] + ((int64_t)o * kCin + c0) * 9 + r);
      }
      __syncthreads();
      for (int ci = 0; ci < kChunk; ++ci) {
        const float* w0 = s_w + (oc0 * kChunk + ci) * 9;
        const float* w1 = w0 + kChunk * 9;
        const float* xin = s_in + ci * kPix;
#pragma unroll
        for (int kk = 0; kk < 9; ++kk) {
          const float wv0 = w0[kk], wv1 = w1[kk];
          const int d = (kk / 3) * 14 + (kk % 3);
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            const float x = xin[off[i] + d];
            acc0[i] = fmaf(wv0, x, acc0[i]);
            acc1[i] = fmaf(wv1, x, acc1[i]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      s_conv[oc0 * kConvPix + sub * 9 + i] = fmaxf(acc0[i], 0.f);
      s_conv[(oc0 + 1) * kConvPix + sub * 9 + i] = fmaxf(acc1[i], 0.f);
    }
  }
  __syncthreads();
  // ---- maxpool2 + BatchNorm (eval)                                                     :235-236
  for (int idx = tid; idx < kC * kHW; idx += kFT) {
    const int c = idx / kHW, q = idx - c * kHW, qy = q / 6, qx = q - qy * 6;
    const float* src = s_conv + c * kConvPix + (2 * qy) * 12 + 2 * qx;
    const float m = fmaxf(fmaxf(src[0], src[1]), fmaxf(src[12], src[13]));
    s_x[idx] = (m - __ldg(P.p[P_BNI_M] + c)) / sqrtf(__ldg(P.p[P_BNI_V] + c) + kEps) * __ldg(P.p[P_BNI_W] + c) +
               __ldg(P.p[P_BNI_B] + c);
  }
  // ---- 2. words: Linear(256 -> 36), gram / sqrt(36), BatchNorm                           :239-242
  {
    const float* wb = word + (int64_t)b * wsb;
    for (int idx = tid; idx < T * kC; idx += kFT) {
      const int t = idx / kC, c = idx - t * kC;
      float acc = __ldg(P.p[P_PROJ_B] + c);
      const float* wr = P.p[P_PROJ_W] + c * kCin;
      for (int d = 0; d < kCin; ++d) acc = fmaf(__ldg(wb + (int64_t)d * wsd + (int64_t)t * wst), __ldg(wr + d), acc);
      s_proj[idx] = acc;
    }
  }
  __syncthreads();
  for (int idx = tid; idx < kC * kHW; idx += kFT) {
    const int c1 = idx / kHW, c2 = idx - c1 * kHW;
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc = fmaf(s_proj[t * kC + c1], s_proj[t * kC + c2], acc);
    acc = acc / 6.f;
    s_y[idx] = (acc - __ldg(P.p[P_BNW_M] + c1)) / sqrtf(__ldg(P.p[P_BNW_V] + c1) + kEps) * __ldg(P.p[P_BNW_W] + c1) +
               __ldg(P.p[P_BNW_B] + c1);
  }
  __syncthreads();
  // ---- 3. SelfAttention(x = image, y = text)                                            :82-118, called at :247
  conv1x1_36(P.p[P_Q_W], P.p[P_Q_B], s_y, s_q);
  conv1x1_36(P.p[P_K_W], P.p[P_K_B], s_x, s_k);
  conv1x1_36(P.p[P_V_W], P.p[P_V_B], s_x, s_v);
  __syncthreads();
  for (int idx = tid; idx < kHW * kHW; idx += kFT) {      // att[i][j] = sum_c key[c][i] query[c][j] / sqrt(36)
    const int i = idx / kHW, j = idx - i * kHW;
    float acc = 0.f;
    for (int c = 0; c < kC; ++c) acc = fmaf(s_k[c * kHW + i], s_q[c * kHW + j], acc);
    s_att[idx] = acc / 6.f;
  }
  __syncthreads();
  if (tid < kHW) {                                         // softmax over j
    float* row = s_att + tid * kHW;
    float mx = row[0];
    for (int j = 1; j < kHW; ++j) mx = fmaxf(mx, row[j]);
    float sum = 0.f;
    for (int j = 0; j < kHW; ++j) {
      row[j] = expf(row[j] - mx);
      sum += row[j];
    }
    const float inv = 1.f / sum;
    for (int j = 0; j < kHW; ++j) row[j] *= inv;
  }
  __syncthreads();
  float lsum = 0.f;
  for (int idx = tid; idx < kC * kHW; idx += kFT) {        // response[c][i] = sum_j att[i][j] value[c][j]
    const int c = idx / kHW, i = idx - c * kHW;
    float acc = 0.f;
    for (int j = 0; j < kHW; ++j) acc = fmaf(s_att[i * kHW + j], s_v[c * kHW + j], acc);
    s_r[idx] = acc;
    lsum += acc;
  }
  // ---- 4. LayerNorm([36,6,6])                                                           :248
  const float mean = block_sum(lsum, s_red) * (1.f / (kC * kHW));
  float lvar = 0.f;
  for (int idx = tid; idx < kC * kHW; idx += kFT) {
    const float d = s_r[idx] - mean;
    lvar = fmaf(d, d, lvar);
  }
  const float rstd = 1.f / sqrtf(block_sum(lvar, s_red) * (1.f / (kC * kHW)) + kEps);
  for (int idx = tid; idx < kC * kHW; idx += kFT)
    s_q[idx] = (s_r[idx] - mean) * rstd * __ldg(P.p[P_LN_W] + idx) + __ldg(P.p[P_LN_B] + idx);
  __syncthreads();
  // ---- 5. maxpool2 -> [324] -> Linear(324 -> 128)                                       :249-254
  for (int idx = tid; idx < kC * 9; idx += kFT) {
    const int c = idx / 9, q = idx - c * 9, qy = q / 3, qx = q - qy * 3;
    const float* src = s_q + c * kHW + (2 * qy) * 6 + 2 * qx;
    s_k[idx] = fmaxf(fmaxf(src[0], src[1]), fmaxf(src[6], src[7]));
  }
  __syncthreads();
  float* ob = out + (int64_t)b * out_sr;
  if (tid < 128) {
    float acc = __ldg(P.p[P_LIN_B] + tid);
    const float* wr = P.p[P_LIN_W] + tid * 324;
    for (int i = 0; i < 324; ++i) acc = fmaf(__ldg(wr + i), s_k[i], acc);
    ob[tid] = acc;
  }
  // ---- 6. LayerNorm(256) of the global image feature and of the sentence feature        :255-257
  for (int which = 0; which < 2; ++which) {
    const float* src = which ? sent + (int64_t)b * se_sr : gl + (int64_t)b * gl_sr;
    const float x = tid < 256 ? __ldg(src + tid) : 0.f;
    const float mu = block_sum(x, s_red) * (1.f / 256.f);
    const float d = tid < 256 ? x - mu : 0.f;
    const float rs = 1.f / sqrtf(block_sum(d * d, s_red) * (1.f / 256.f) + kEps);
    if (tid < 256)
      ob[128 + which * 256 + tid] = d * rs * __ldg(P.p[which ? P_LNS_W : P_LNG_W] + tid) + __ldg(P.p[which ? P_LNS_B : P_LNG_B] + tid);
  }
}

}  // namespace

int fcfm_working_fwd(const float* img, int64_t isb, int64_t isc, int64_t ish, int64_t isw, const float* word, int64_t wsb,
                     int64_t wsd, int64_t wst, const float* gl, int64_t gl_sr, const float* sent, int64_t se_sr,
                     const float* const* params, int B, int T, float* out, int64_t out_sr, cudaStream_t st) {
  TGFR_REQUIRE(B >= 0 && T >= 1 && T <= kMaxT, "fcfm_working_fwd: need 1 <= T <= %d words, got %d", kMaxT, T);
  if (B == 0) return TGFR_OK;
  FcfmParams P;
  for (int k = 0; k < P_NUM; ++k) {
    TGFR_REQUIRE(params[k] != nullptr, "fcfm_working_fwd: parameter %d is NULL", k);
    P.p[k] = params[k];
  }
  constexpr int smem = kSmemFloats * (int)sizeof(float);
  static bool attr_done[64] = {};
  int dev = 0;
  TGFR_CUDA_OK(cudaGetDevice(&dev));
  if (!attr_done[dev & 63]) {
    TGFR_CUDA_OK(cudaFuncSetAttribute(fcfm_working_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done[dev & 63] = true;
  }
  fcfm_working_fwd_kernel<<<B, kFT, smem, st>>>(img, isb, isc, ish, isw, word, wsb, wsd, wst, gl, gl_sr, sent, se_sr, P, T, out,
                                                out_sr);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int fcfm_working_num_params() { return P_NUM; }

}  // namespace tgfr

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
//
// Large batches (verification: 60 000 samples per call) take the convolution out of that kernel and run it on the
// tensor cores as ONE implicit GEMM per chunk of 4096 samples (fcfm_working_fwd_tc): the image is re-laid channels-last
// as fp16 hi / lo [B*196, 256] (one transpose pass), the weights as [36, 9*256] tap-major, and gemm_tc_conv3x3 walks the
// nine taps as row offsets of its TMA box (no im2col matrix), three split terms per tap for fp32-class accuracy, bias +
// ReLU in the epilogue.  The per-sample kernel then starts from the [196, 36] conv rows.
#include <cuda_fp16.h>

#include "common.cuh"

namespace tgfr {

struct TcOperand {          // gemm_tc.cu
  const __half *hi, *lo;
  int64_t ld;
  const float* scale;
};
int gemm_tc_conv3x3(const TcOperand& img, const TcOperand& w, float* C, int64_t ldc, int rows, int N, int Cin, int width,
                    const float* bias, int relu, int nterms, cudaStream_t st);
int gemm_tc_split_operand(const float* src, int64_t ld, int rows, int cols, float* scale, __half* hi, __half* lo, int ld_out,
                          cudaStream_t st, bool have_max);

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
                                                               int64_t out_sr, const float* __restrict__ conv, int b0) {
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
  const int b = blockIdx.x + b0;

  // ---- 1. conv3x3 (valid) + ReLU                                                       fusion_nets.py:235
  if (conv) {
    // already computed on the tensor cores (bias + ReLU applied): rows = the 196 anchor positions of this chunk's sample
    const float* cb = conv + (int64_t)blockIdx.x * kPix * kC;
    for (int idx = tid; idx < kConvPix * kC; idx += kFT) {
      const int pix = idx / kC, c = idx - pix * kC;
      const int y = pix / 12, x = pix - y * 12;
      s_conv[c * kConvPix + pix] = __ldg(cb + (y * 14 + x) * kC + c);
    }
  } else {
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
  // 32 words at a time are staged in the (now dead) convolution buffers as [t][256 + 4]; thread = (output channel c, word
  // lane g): the weight row of c is read once as float4s for up to four words, the word values are shared-memory broadcasts
  {
    const float* wb = word + (int64_t)b * wsb;
    float* s_word = sm;                                    // kChunk * kPix + kC * kChunk * 9 = 32 x 260 floats
    constexpr int kWP = kCin + 4;
    const int c = tid % kC, g = tid / kC;                  // 288 threads = 36 channels x 8 word lanes
    const float4* wr = reinterpret_cast<const float4*>(P.p[P_PROJ_W] + c * kCin);
    const float bias = __ldg(P.p[P_PROJ_B] + c);
    for (int t0 = 0; t0 < T; t0 += 32) {
      const int nT = min(32, T - t0);
      __syncthreads();
      for (int idx = tid; idx < nT * kCin; idx += kFT) {
        int tt, d;
        if (wsd == 1) {                                    // [B, T, 256] memory: the feature index is the fast one
          tt = idx >> 8;
          d = idx & (kCin - 1);
        } else {
          d = idx / nT;
          tt = idx - d * nT;
        }
        s_word[tt * kWP + d] = __ldg(wb + (int64_t)d * wsd + (int64_t)(t0 + tt) * wst);
      }
      __syncthreads();
      float acc[4] = {bias, bias, bias, bias};
      for (int d4 = 0; d4 < kCin / 4; ++d4) {
        const float4 w4 = __ldg(wr + d4);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int tt = g + 8 * u;
          if (tt < nT) {
            const float4 x4 = *reinterpret_cast<const float4*>(s_word + tt * kWP + 4 * d4);
            acc[u] = fmaf(x4.x, w4.x, acc[u]);
            acc[u] = fmaf(x4.y, w4.y, acc[u]);
            acc[u] = fmaf(x4.z, w4.z, acc[u]);
            acc[u] = fmaf(x4.w, w4.w, acc[u]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (g + 8 * u < nT) s_proj[(t0 + g + 8 * u) * kC + c] = acc[u];
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

// ---- operands of the tensor-core convolution ----
// max |img| over a chunk of samples (element strides) into scale[0] (zeroed by the caller)
__global__ void fcfm_img_maxabs_kernel(const float* __restrict__ img, int64_t isb, int64_t isc, int64_t ish, int64_t isw, int nb,
                                       float* __restrict__ scale) {
  __shared__ float scratch[32];
  const int64_t n = (int64_t)nb * kCin * kPix;
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int pix = (int)(i % kPix);
    const int64_t r = i / kPix;
    const int c = (int)(r % kCin);
    const int64_t b = r / kCin;
    m = fmaxf(m, fabsf(__ldg(img + b * isb + c * isc + (pix / 14) * ish + (pix % 14) * isw)));
  }
  m = block_max(m, scratch);
  if (threadIdx.x == 0) atomicMax(reinterpret_cast<int*>(scale), __float_as_int(m));
}
// img [nb,256,14,14] (element strides) -> hi / lo fp16 [nb*196, 256] channels-last, scaled by the power of two that puts
// the largest entry near 2^12; 32 channels x 32 positions per block through shared memory; scale[1] = 2^-e
__global__ void fcfm_img_split_kernel(const float* __restrict__ img, int64_t isb, int64_t isc, int64_t ish, int64_t isw,
                                      float* __restrict__ scale, __half* __restrict__ hi, __half* __restrict__ lo) {
  __shared__ float tile[32][33];
  const float mx = scale[0];
  const float sc = (mx > 0.f) ? exp2f(floorf(log2f(4096.f / mx))) : 1.f;
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0 && threadIdx.y == 0) scale[1] = 1.f / sc;
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const float* ib = img + (int64_t)b * isb;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, pix = p0 + threadIdx.x;
    tile[j][threadIdx.x] = pix < kPix ? __ldg(ib + (int64_t)c * isc + (pix / 14) * ish + (pix % 14) * isw) * sc : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int pix = p0 + j;
    if (pix < kPix) {
      const float x = tile[threadIdx.x][j];
      const __half h = __float2half_rn(x);
      const int64_t o = ((int64_t)b * kPix + pix) * kCin + c0 + threadIdx.x;
      hi[o] = h;
      lo[o] = __float2half_rn(x - __half2float(h));
    }
  }
}
// dense samples (NCHW-contiguous or channels-last: 50 176 consecutive floats each): the order does not matter for a maximum
__global__ void fcfm_img_maxabs_flat_kernel(const float* __restrict__ img, int64_t isb, int nb, float* __restrict__ scale) {
  __shared__ float scratch[32];
  constexpr int q = kCin * kPix / 4;
  const int64_t n = (int64_t)nb * q;
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / q;
    const float4 v = __ldg(reinterpret_cast<const float4*>(img + b * isb) + (i - b * q));
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  m = block_max(m, scratch);
  if (threadIdx.x == 0) atomicMax(reinterpret_cast<int*>(scale), __float_as_int(m));
}
// channels-last memory (IMIM's output layout): the copies keep the order, 16 bytes in / 8 + 8 bytes out per thread
__global__ void fcfm_img_split_cl_kernel(const float* __restrict__ img, int64_t isb, int nb, float* __restrict__ scale,
                                         __half* __restrict__ hi, __half* __restrict__ lo) {
  const float mx = scale[0];
  const float sc = (mx > 0.f) ? exp2f(floorf(log2f(4096.f / mx))) : 1.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) scale[1] = 1.f / sc;
  constexpr int q = kCin * kPix / 4;
  const int64_t n = (int64_t)nb * q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / q;
    const float4 v = __ldg(reinterpret_cast<const float4*>(img + b * isb) + (i - b * q));
    const float x[4] = {v.x * sc, v.y * sc, v.z * sc, v.w * sc};
    __half h[4], l[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      h[u] = __float2half_rn(x[u]);
      l[u] = __float2half_rn(x[u] - __half2float(h[u]));
    }
    *reinterpret_cast<uint2*>(hi + 4 * i) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(lo + 4 * i) = *reinterpret_cast<const uint2*>(l);
  }
}
// conv.weight [36, 256, 3, 3] -> [36, 9, 256] (tap-major K of the implicit GEMM)
__global__ void fcfm_conv_w_taps_kernel(const float* __restrict__ w, float* __restrict__ wt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kC * 9 * kCin) return;
  const int c = i % kCin, tap = (i / kCin) % 9, o = i / (9 * kCin);
  wt[i] = __ldg(w + ((int64_t)o * kCin + c) * 9 + tap);
}

constexpr int kConvChunk = 4096;          // samples per implicit-GEMM launch (0.8 GB of fp16 image copies)
struct FcfmTcLayout {
  size_t scales, wt, whi, wlo, ihi, ilo, conv, total;
};
FcfmTcLayout fcfm_tc_layout(int B) {
  FcfmTcLayout L;
  const size_t chunk = (size_t)(B < kConvChunk ? B : kConvChunk);
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  L.scales = take(256);
  L.wt = take((size_t)kC * 9 * kCin * 4);
  L.whi = take((size_t)kC * 9 * kCin * 2);
  L.wlo = take((size_t)kC * 9 * kCin * 2);
  L.ihi = take(chunk * kPix * kCin * 2);
  L.ilo = take(chunk * kPix * kCin * 2);
  L.conv = take(chunk * kPix * kC * 4);
  L.total = off;
  return L;
}

}  // namespace

size_t fcfm_working_workspace_bytes(int B) { return B > 0 ? fcfm_tc_layout(B).total : 0; }

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
  TGFR_REQUIRE((reinterpret_cast<uintptr_t>(P.p[P_PROJ_W]) & 15) == 0, "fcfm_working_fwd: projection.weight must be 16-byte aligned");
  constexpr int smem = kSmemFloats * (int)sizeof(float);
  static bool attr_done[64] = {};
  int dev = 0;
  TGFR_CUDA_OK(cudaGetDevice(&dev));
  if (!attr_done[dev & 63]) {
    TGFR_CUDA_OK(cudaFuncSetAttribute(fcfm_working_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done[dev & 63] = true;
  }
  fcfm_working_fwd_kernel<<<B, kFT, smem, st>>>(img, isb, isc, ish, isw, word, wsb, wsd, wst, gl, gl_sr, sent, se_sr, P, T, out,
                                                out_sr, nullptr, 0);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

// the same forward with the convolution on the tensor cores (see the header); ws = fcfm_working_workspace_bytes(B) bytes
int fcfm_working_fwd_tc(const float* img, int64_t isb, int64_t isc, int64_t ish, int64_t isw, const float* word, int64_t wsb,
                        int64_t wsd, int64_t wst, const float* gl, int64_t gl_sr, const float* sent, int64_t se_sr,
                        const float* const* params, int B, int T, float* out, int64_t out_sr, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
  TGFR_REQUIRE(B >= 0 && T >= 1 && T <= kMaxT, "fcfm_working_fwd: need 1 <= T <= %d words, got %d", kMaxT, T);
  if (B == 0) return TGFR_OK;
  const FcfmTcLayout L = fcfm_tc_layout(B);
  TGFR_REQUIRE(ws && ws_bytes >= L.total && (reinterpret_cast<uintptr_t>(ws) & 255) == 0,
               "fcfm_working_fwd: workspace too small or not 256-byte aligned (%zu < %zu)", ws_bytes, L.total);
  FcfmParams P;
  for (int k = 0; k < P_NUM; ++k) {
    TGFR_REQUIRE(params[k] != nullptr, "fcfm_working_fwd: parameter %d is NULL", k);
    P.p[k] = params[k];
  }
  TGFR_REQUIRE((reinterpret_cast<uintptr_t>(P.p[P_PROJ_W]) & 15) == 0, "fcfm_working_fwd: projection.weight must be 16-byte aligned");
  constexpr int smem = kSmemFloats * (int)sizeof(float);
  static bool attr_done[64] = {};
  int dev = 0;
  TGFR_CUDA_OK(cudaGetDevice(&dev));
  if (!attr_done[dev & 63]) {
    TGFR_CUDA_OK(cudaFuncSetAttribute(fcfm_working_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done[dev & 63] = true;
  }
  uint8_t* base = reinterpret_cast<uint8_t*>(ws);
  float* scales = reinterpret_cast<float*>(base + L.scales);
  float* wt = reinterpret_cast<float*>(base + L.wt);
  __half* whi = reinterpret_cast<__half*>(base + L.whi);
  __half* wlo = reinterpret_cast<__half*>(base + L.wlo);
  __half* ihi = reinterpret_cast<__half*>(base + L.ihi);
  __half* ilo = reinterpret_cast<__half*>(base + L.ilo);
  float* conv = reinterpret_cast<float*>(base + L.conv);
  fcfm_conv_w_taps_kernel<<<ceil_div(kC * 9 * kCin, 256), 256, 0, st>>>(P.p[P_CONV_W], wt);
  TGFR_LAUNCH_OK();
  if (int rc = gemm_tc_split_operand(wt, 9 * kCin, kC, 9 * kCin, scales + 2, whi, wlo, 9 * kCin, st, false)) return rc;
  const TcOperand ow{whi, wlo, 9 * kCin, scales + 2};
  for (int b0 = 0; b0 < B; b0 += kConvChunk) {
    const int nb = B - b0 < kConvChunk ? B - b0 : kConvChunk;
    const float* ic = img + (int64_t)b0 * isb;
    TGFR_CUDA_OK(cudaMemsetAsync(scales, 0, 2 * sizeof(float), st));
    const int64_t n = (int64_t)nb * kCin * kPix;
    const bool aligned = (reinterpret_cast<uintptr_t>(ic) & 15) == 0 && (isb & 3) == 0;
    const bool chlast = isc == 1 && isw == kCin && ish == 14 * kCin;
    const bool nchw = isw == 1 && ish == 14 && isc == kPix;
    const int blocks = (int)((n / 4 + 255) / 256 < 148 * 8 ? (n / 4 + 255) / 256 : 148 * 8);
    if (aligned && (chlast || nchw))
      fcfm_img_maxabs_flat_kernel<<<blocks, 256, 0, st>>>(ic, isb, nb, scales);
    else
      fcfm_img_maxabs_kernel<<<blocks, 256, 0, st>>>(ic, isb, isc, ish, isw, nb, scales);
    TGFR_LAUNCH_OK();
    if (aligned && chlast)
      fcfm_img_split_cl_kernel<<<blocks, 256, 0, st>>>(ic, isb, nb, scales, ihi, ilo);
    else
      fcfm_img_split_kernel<<<dim3(ceil_div(kPix, 32), kCin / 32, nb), dim3(32, 8), 0, st>>>(ic, isb, isc, ish, isw, scales, ihi, ilo);
    TGFR_LAUNCH_OK();
    const TcOperand oi{ihi, ilo, kCin, scales};
    if (int rc = gemm_tc_conv3x3(oi, ow, conv, kC, nb * kPix, kC, kCin, 14, P.p[P_CONV_B], 1, 3, st)) return rc;
    fcfm_working_fwd_kernel<<<nb, kFT, smem, st>>>(img, isb, isc, ish, isw, word, wsb, wsd, wst, gl, gl_sr, sent, se_sr, P, T, out,
                                                   out_sr, conv, b0);
    TGFR_LAUNCH_OK();
  }
  return TGFR_OK;
}

int fcfm_working_num_params() { return P_NUM; }

}  // namespace tgfr

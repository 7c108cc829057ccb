// TextHeading (reference models/models.py:170-232): the BERT-token -> word / sentence feature head that produces
// the `words_emb` / `sent_emb` inputs of the FCAM losses.
//
//   Bert_Word_Mapping : three n-gram convolutions Conv2d(1, F, (K, E)), K = 2, 3, 4, + ReLU over tokens [B, L, E].
//                       A window of K consecutive tokens is K E CONTIGUOUS floats, so each convolution is one
//                       strided product  act_K [B L, F] = relu(Win_K . W_K^T + b_K)  with Win_K[i, :] = tokens[i E ...]
//                       (row pitch E, row length K E: overlapping rows, no im2col copy).  Windows that run over the
//                       end of a caption (j > L - K) are computed and never read.
//   get_each_word_feature : word t = max over the convolutions that have position t (all three for t < seq,
//                       K = 2, 3 for t = seq, K = 2 alone for t = seq + 1), L2-normalised -> words [B, T, F] in the
//                       layout the word-region kernels read.  The reference builds this with a Python double loop
//                       of B x T stack / amax calls (models.py:197-213).
//   get_word_feature  : max over positions per convolution, mean of the three, L2-normalised -> sent [B, F].
//
// Backward (the head is trained; the BERT tokens are not): normalisation backward, the amax routing (ties share the
// gradient evenly, as torch.amax does; max_pool1d takes the first maximum), ReLU mask ->  G_K [B L, F], then
// dW_K = G_K^T . Win_K (the same overlapping view as the B operand) and db_K = column sums.
// Reference quirk kept: the last word is copied with `torch.cuda.FloatTensor(a[i, seq+1])` (models.py:206), which
// detaches it -- it receives no gradient.
//
// This first version runs the products on the exact fp32 SIMT GEMM (dense_simt.cu): the max routing is a discrete
// decision, and fp16-operand rounding would flip near ties.
#include "common.cuh"

namespace tgfr {

int sgemm_strided(const float* A, int64_t sAm, int64_t sAk, const float* Bm, int64_t sBk, int64_t sBn, float* C,
                  int64_t sCm, int64_t sCn, int M, int N, int K, const float* bias, int relu, cudaStream_t st);

namespace {

constexpr int kConvs = 3;              // K = 2, 3, 4
constexpr float kNormEps = 1e-12f;     // F.normalize

struct Acts {
  const float* a[kConvs];              // act_K [B L, F]
};
struct Grads {
  float* g[kConvs];                    // G_K [B L, F] (pre-zeroed)
  float* db[kConvs];                   // [F] (pre-zeroed)
};

// one block per caption b, thread f owns feature f (strided when F > blockDim)
__global__ void texthead_combine_fwd_kernel(const Acts acts, int L, int F, int T, int seq, float* __restrict__ words,
                                            float* __restrict__ sent) {
  __shared__ float scratch[32];
  const int b = blockIdx.x;
  const int64_t row0 = (int64_t)b * L;
  for (int t = 0; t < T; ++t) {
    float ss = 0.f;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
      float v = acts.a[0][(row0 + t) * F + f];
      if (t <= seq) v = fmaxf(v, acts.a[1][(row0 + t) * F + f]);
      if (t < seq) v = fmaxf(v, acts.a[2][(row0 + t) * F + f]);
      ss = fmaf(v, v, ss);
    }
    ss = block_sum(ss, scratch);
    const float inv = 1.f / fmaxf(sqrtf(ss), kNormEps);
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
      float v = acts.a[0][(row0 + t) * F + f];
      if (t <= seq) v = fmaxf(v, acts.a[1][(row0 + t) * F + f]);
      if (t < seq) v = fmaxf(v, acts.a[2][(row0 + t) * F + f]);
      words[((int64_t)b * T + t) * F + f] = v * inv;
    }
  }
  // sentence feature: max over the positions of each convolution, mean of the three
  float ss = 0.f;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float o = 0.f;
    for (int k = 0; k < kConvs; ++k) {
      float m = -INFINITY;
      for (int j = 0; j < L - 1 - k; ++j) m = fmaxf(m, acts.a[k][(row0 + j) * F + f]);
      o += m;
    }
    o *= (1.f / 3.f);
    sent[(int64_t)b * F + f] = o;          // un-normalised for the moment
    ss = fmaf(o, o, ss);
  }
  ss = block_sum(ss, scratch);
  const float inv = 1.f / fmaxf(sqrtf(ss), kNormEps);
  for (int f = threadIdx.x; f < F; f += blockDim.x) sent[(int64_t)b * F + f] *= inv;
}

__global__ void texthead_combine_bwd_kernel(const Acts acts, int L, int F, int T, int seq,
                                            const float* __restrict__ gwords, const float* __restrict__ gsent,
                                            const Grads out) {
  __shared__ float scratch[32];
  const int b = blockIdx.x;
  const int64_t row0 = (int64_t)b * L;
  if (gwords) {
    for (int t = 0; t < T - 1; ++t) {      // the last word is detached in the reference (models.py:206)
      float ss = 0.f, gd = 0.f;
      for (int f = threadIdx.x; f < F; f += blockDim.x) {
        float v = acts.a[0][(row0 + t) * F + f];
        if (t <= seq) v = fmaxf(v, acts.a[1][(row0 + t) * F + f]);
        if (t < seq) v = fmaxf(v, acts.a[2][(row0 + t) * F + f]);
        ss = fmaf(v, v, ss);
        gd = fmaf(gwords[((int64_t)b * T + t) * F + f], v, gd);
      }
      ss = block_sum(ss, scratch);
      gd = block_sum(gd, scratch);
      const float nrm = sqrtf(ss);
      const int ncand = t < seq ? 3 : 2;   // t == seq: K = 2, 3 (t = seq + 1 never reaches here)
      for (int f = threadIdx.x; f < F; f += blockDim.x) {
        float a[kConvs], v = -INFINITY;
        for (int k = 0; k < ncand; ++k) {
          a[k] = acts.a[k][(row0 + t) * F + f];
          v = fmaxf(v, a[k]);
        }
        const float g = gwords[((int64_t)b * T + t) * F + f];
        // d (v / max(|v|, eps)): projection when the clamp is inactive, plain scaling otherwise
        const float dv = nrm > kNormEps ? (g - gd * v / (nrm * nrm)) / nrm : g / kNormEps;
        int nt = 0;
        for (int k = 0; k < ncand; ++k) nt += (a[k] == v);
        const float share = dv / (float)nt;                       // torch.amax: ties share the gradient evenly
        for (int k = 0; k < ncand; ++k)
          if (a[k] == v && a[k] > 0.f) out.g[k][(row0 + t) * F + f] += share;   // ReLU mask
      }
    }
  }
  if (gsent) {
    float ss = 0.f, gd = 0.f;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
      float o = 0.f;
      for (int k = 0; k < kConvs; ++k) {
        float m = -INFINITY;
        for (int j = 0; j < L - 1 - k; ++j) m = fmaxf(m, acts.a[k][(row0 + j) * F + f]);
        o += m;
      }
      o *= (1.f / 3.f);
      ss = fmaf(o, o, ss);
      gd = fmaf(gsent[(int64_t)b * F + f], o, gd);
    }
    ss = block_sum(ss, scratch);
    gd = block_sum(gd, scratch);
    const float nrm = sqrtf(ss);
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
      float m[kConvs], o = 0.f;
      int jm[kConvs];
      for (int k = 0; k < kConvs; ++k) {
        m[k] = -INFINITY;
        jm[k] = 0;
        for (int j = 0; j < L - 1 - k; ++j) {
          const float x = acts.a[k][(row0 + j) * F + f];
          if (x > m[k]) { m[k] = x; jm[k] = j; }                   // max_pool1d: the first maximum
        }
        o += m[k];
      }
      o *= (1.f / 3.f);
      const float g = gsent[(int64_t)b * F + f];
      const float dout = nrm > kNormEps ? (g - gd * o / (nrm * nrm)) / nrm : g / kNormEps;
      for (int k = 0; k < kConvs; ++k)
        if (m[k] > 0.f) out.g[k][(row0 + jm[k]) * F + f] += dout * (1.f / 3.f);
    }
  }
  // bias gradients: this caption's column sums (thread f is the only writer of column f within the block)
  __syncthreads();
  for (int f = threadIdx.x; f < F; f += blockDim.x)
    for (int k = 0; k < kConvs; ++k) {
      float s = 0.f;
      for (int j = 0; j < L - 1 - k; ++j) s += out.g[k][(row0 + j) * F + f];
      if (s != 0.f) atomicAdd(out.db[k] + f, s);
    }
}

}  // namespace

// saved: act_2 | act_3 | act_4, each [B L, F] fp32;  workspace (backward): G_2 | G_3 | G_4
size_t texthead_saved_bytes(int B, int L, int F) { return sizeof(float) * kConvs * (size_t)B * L * F; }
size_t texthead_workspace_bytes(int B, int L, int F) { return sizeof(float) * kConvs * (size_t)B * L * F; }

static int check(int B, int L, int E, int F, int words_num) {
  TGFR_REQUIRE(B > 0 && E > 0 && F > 0, "texthead: empty shape");
  TGFR_REQUIRE(words_num >= 5 && L == words_num - 1, "texthead: tokens must hold bert_words_num - 1 = %d positions, got %d",
               words_num - 1, L);
  return TGFR_OK;
}

int texthead_fwd(const float* tokens, const float* const* w, const float* const* bias, int B, int L, int E, int F,
                 int words_num, float* words, float* sent, void* saved, size_t saved_bytes, cudaStream_t st) {
  if (int rc = check(B, L, E, F, words_num)) return rc;
  TGFR_REQUIRE(saved && saved_bytes >= texthead_saved_bytes(B, L, F), "texthead_fwd: saved buffer too small");
  const int seq = words_num - 4, T = seq + 2;
  float* act = reinterpret_cast<float*>(saved);
  Acts acts{};
  for (int k = 0; k < kConvs; ++k) {
    const int K = k + 2, rows = B * L - K + 1;          // every token position that still has K tokens after it
    float* a = act + (size_t)k * B * L * F;
    acts.a[k] = a;
    // the last K - 1 rows are never produced: keep them defined
    TGFR_CUDA_OK(cudaMemsetAsync(a + (size_t)rows * F, 0, sizeof(float) * (size_t)(K - 1) * F, st));
    if (int rc = sgemm_strided(tokens, E, 1, w[k], 1, (int64_t)K * E, a, F, 1, rows, F, K * E, bias[k], 1, st)) return rc;
  }
  const int threads = F >= 256 ? 256 : ((F + 31) & ~31);
  texthead_combine_fwd_kernel<<<B, threads, 0, st>>>(acts, L, F, T, seq, words, sent);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int texthead_bwd(const float* tokens, const float* gwords, const float* gsent, int B, int L, int E, int F, int words_num,
                 float* const* dw, float* const* db, void* ws, size_t ws_bytes, const void* saved, size_t saved_bytes,
                 cudaStream_t st) {
  if (int rc = check(B, L, E, F, words_num)) return rc;
  TGFR_REQUIRE(saved && saved_bytes >= texthead_saved_bytes(B, L, F), "texthead_bwd: saved buffer too small");
  TGFR_REQUIRE(ws && ws_bytes >= texthead_workspace_bytes(B, L, F), "texthead_bwd: workspace too small");
  const int seq = words_num - 4, T = seq + 2;
  const float* act = reinterpret_cast<const float*>(saved);
  float* G = reinterpret_cast<float*>(ws);
  TGFR_CUDA_OK(cudaMemsetAsync(G, 0, texthead_workspace_bytes(B, L, F), st));
  Acts acts{};
  Grads gr{};
  for (int k = 0; k < kConvs; ++k) {
    acts.a[k] = act + (size_t)k * B * L * F;
    gr.g[k] = G + (size_t)k * B * L * F;
    gr.db[k] = db[k];
    TGFR_CUDA_OK(cudaMemsetAsync(db[k], 0, sizeof(float) * F, st));
  }
  const int threads = F >= 256 ? 256 : ((F + 31) & ~31);
  texthead_combine_bwd_kernel<<<B, threads, 0, st>>>(acts, L, F, T, seq, gwords, gsent, gr);
  TGFR_LAUNCH_OK();
  for (int k = 0; k < kConvs; ++k) {
    const int K = k + 2, rows = B * L - K + 1;
    // dW_K [F, K E] = G_K^T [F, rows] . Win_K [rows, K E]
    if (int rc = sgemm_strided(gr.g[k], 1, F, tokens, E, 1, dw[k], (int64_t)K * E, 1, F, K * E, rows, nullptr, 0, st))
      return rc;
  }
  return TGFR_OK;
}

}  // namespace tgfr

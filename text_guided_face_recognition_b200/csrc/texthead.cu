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
// Products.  The max routing is a discrete decision, so plain fp16-operand tensor-core products (2^-11 operand
// rounding) are not good enough: they would flip near ties.  The tensor-core path therefore splits every fp32 operand
// into fp16 hi + lo (x 2^e = hi + lo, about 22 significant bits) and accumulates the three terms hi hi + hi lo + lo hi
// in one TMEM accumulator (gemm_tc.cu, nterms = 3) -- the accuracy of an fp32 SGEMM at 3x the tensor-core work, which is
// still ~10x faster than the fp32 SIMT GEMM.  The sliding windows stay views: the TMA tensor map of Win_K has row pitch
// E and row length K E (overlapping rows), K-major for the forward and MN-major for dW.  The three convolutions are one
// launch each way (blockIdx.z).  Shapes with E % 8 or F % 8 != 0, or TGFR_TEXTHEAD_PRECISION=fp32, run the exact fp32
// SIMT GEMM (dense_simt.cu).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace tgfr {

int sgemm_strided(const float* A, int64_t sAm, int64_t sAk, const float* Bm, int64_t sBk, int64_t sBn, float* C,
                  int64_t sCm, int64_t sCn, int M, int N, int K, const float* bias, int relu, cudaStream_t st);
// gemm_tc.cu
struct Split3Product {
  const __half *a_hi, *a_lo, *b_hi, *b_lo;
  int64_t lda, ldb, ldc;
  int M, N, K;
  float* C;
  const float* bias;
};
int gemm_tc_split3_batched(const Split3Product* prods, int n, int a_mn, int a_overlap, int b_mn, int b_overlap, float alpha,
                           const float* dscale, const float* dscale2, int relu, cudaStream_t st);
int head_split_f16(int nblocks, const float* const* src, const int64_t* ld, const int* rows, const int* cols, float* scale,
                   __half* const* hi, __half* const* lo, const int* ld_out, cudaStream_t st, bool have_max = false);

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

// grid (T + 1, B): block (t, b) builds word t of caption b, block (T, b) its sentence feature; thread f owns
// feature f (strided when F > blockDim)
__global__ void texthead_combine_fwd_kernel(const Acts acts, int L, int F, int T, int seq, float* __restrict__ words,
                                            float* __restrict__ sent) {
  __shared__ float scratch[32];
  const int b = blockIdx.y;
  const int64_t row0 = (int64_t)b * L;
  if ((int)blockIdx.x < T) {
    const int t = blockIdx.x;
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
    return;
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

// ---------------------------------------------------------------------------------------------
// all-tensor-core forward: exact re-evaluation of the near-tied activations
//
// The split-operand tensor-core products are good to ~3e-6 (the tensor core truncates when it aligns its fp32
// accumulator), the fp32 SIMT GEMM to ~5e-7.  Outputs do not care, the arg-max routing of the backward does: two
// activations closer than the error can swap.  So every activation that is within kTieTol of the maximum it competes
// for (the per-word max over the convolutions, the per-convolution max over the positions) is listed and recomputed
// as a float64 dot product -- correctly rounded, i.e. the routing then follows the exact values more closely than
// the SIMT GEMM's does.  About one activation in 10^4 is listed.
// ---------------------------------------------------------------------------------------------
constexpr float kTieTol = 4e-5f;
constexpr uint32_t kTieCap = 1u << 16;

struct TieList {
  uint32_t* count;      // [1] entries wanted (may exceed kTieCap: the rest keeps its tensor-core value)
  uint2* items;         // [kTieCap] (row, feature | conv << 24)
};
__device__ __forceinline__ void tie_push(const TieList& tl, int k, int64_t row, int f) {
  const uint32_t at = atomicAdd(tl.count, 1u);
  if (at < kTieCap) tl.items[at] = make_uint2((uint32_t)row, (uint32_t)f | ((uint32_t)k << 24));
}
// grid (seq + 2, B): block (t <= seq, b) checks word t, block (seq + 1, b) the sentence maxima
__global__ void texthead_mark_ties_kernel(const Acts acts, int L, int F, int T, int seq, const TieList tl) {
  const int b = blockIdx.y;
  const int64_t row0 = (int64_t)b * L;
  const bool sentence = (int)blockIdx.x == seq + 1;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    if (!sentence) {                                       // word t: max over the convolutions that reach position t
      const int t = blockIdx.x;
      const int nk = t < seq ? 3 : 2;
      float v[kConvs], top = -INFINITY;
      for (int k = 0; k < nk; ++k) {
        v[k] = acts.a[k][(row0 + t) * F + f];
        top = fmaxf(top, v[k]);
      }
      if (top <= 0.f) continue;                            // relu: nothing flows through a zero maximum
      const float tol = kTieTol * fmaxf(1.f, top);
      int close = 0;
      for (int k = 0; k < nk; ++k) close += (top - v[k] < tol);
      if (close > 1)
        for (int k = 0; k < nk; ++k)
          if (top - v[k] < tol) tie_push(tl, k, row0 + t, f);
      continue;
    }
    for (int k = 0; k < kConvs; ++k) {                     // sentence: max over the positions of convolution k
      float top = -INFINITY;
      for (int j = 0; j < L - 1 - k; ++j) top = fmaxf(top, acts.a[k][(row0 + j) * F + f]);
      if (top <= 0.f) continue;
      const float tol = kTieTol * fmaxf(1.f, top);
      int close = 0;
      for (int j = 0; j < L - 1 - k; ++j) close += (top - acts.a[k][(row0 + j) * F + f] < tol);
      if (close > 1)
        for (int j = 0; j < L - 1 - k; ++j)
          if (top - acts.a[k][(row0 + j) * F + f] < tol) tie_push(tl, k, row0 + j, f);
    }
  }
}
struct RefineArgs {
  const float* w[kConvs];
  const float* bias[kConvs];
  float* act[kConvs];
};
// a warp per listed activation: relu(b_K[f] + <tokens[row .. row + K), W_K[f]>) in float64
__global__ void texthead_refine_ties_kernel(const TieList tl, const float* __restrict__ tokens, const RefineArgs ra, int E,
                                            int F) {
  const int lane = threadIdx.x & 31;
  const uint32_t n = min(*tl.count, kTieCap);
  for (uint32_t it = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); it < n; it += gridDim.x * (blockDim.x >> 5)) {
    const uint2 e = tl.items[it];
    const int k = (int)(e.y >> 24), f = (int)(e.y & 0xffffffu), len = (k + 2) * E;
    const float* x = tokens + (int64_t)e.x * E;
    const float* wk = ra.w[k] + (int64_t)f * len;
    double acc = 0.0;
    for (int j = lane; j < len; j += 32) acc = fma((double)__ldg(x + j), (double)__ldg(wk + j), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      if (ra.bias[k]) acc += (double)__ldg(ra.bias[k] + f);
      ra.act[k][(int64_t)e.x * F + f] = fmaxf((float)acc, 0.f);
    }
  }
}

__global__ void texthead_combine_bwd_kernel(const Acts acts, int L, int F, int T, int seq,
                                            const float* __restrict__ gwords, const float* __restrict__ gsent,
                                            const Grads out) {
  // grid (T, B): block (t < T - 1, b) routes the gradient of word t, block (T - 1, b) the sentence feature's.  The two
  // can meet in one element of G_K: both add atomically into the zeroed buffer (two addends: the sum does not depend
  // on the order).  The last word is detached in the reference (models.py:206) and has no block.
  __shared__ float scratch[32];
  const int b = blockIdx.y;
  const int64_t row0 = (int64_t)b * L;
  if ((int)blockIdx.x < T - 1) {
    if (gwords) {
      const int t = blockIdx.x;
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
          if (a[k] == v && a[k] > 0.f) atomicAdd(&out.g[k][(row0 + t) * F + f], share);   // ReLU mask
      }
    }
    return;
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
        if (m[k] > 0.f) atomicAdd(&out.g[k][(row0 + jm[k]) * F + f], dout * (1.f / 3.f));
    }
  }
}

// bias gradients: the column sums of G_K, caption by caption (grid (kConvs, B), after the routing kernel)
__global__ void texthead_bias_grad_kernel(const Grads out, int L, int F) {
  const int k = blockIdx.x, b = blockIdx.y;
  const int64_t row0 = (int64_t)b * L;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < L - 1 - k; ++j) s += out.g[k][(row0 + j) * F + f];
    if (s != 0.f) atomicAdd(out.db[k] + f, s);
  }
}

}  // namespace

// saved: act_2 | act_3 | act_4, each [B L, F] fp32; tensor-core path adds [scales: tokens, weights | tokens hi | lo | W hi | W lo]
// workspace (backward): G_2 | G_3 | G_4 fp32; tensor-core path adds [scale | G hi | G lo]
namespace {
// TGFR_TEXTHEAD_PRECISION = tc (default) | mixed | fp32.
//   tc   : all six products on tcgen05 with hi / lo split operands.  The activations are good to ~3e-6 (the tensor
//          core truncates when it aligns its fp32 accumulator; the SIMT GEMM reaches ~5e-7), which on its own flips the
//          arg-max of about two near-tied (caption, word, feature) triples in 720 k at B = 128 -- a 1e-3-sized change
//          of dW -- so the near-tied activations are re-evaluated exactly (texthead_refine_ties_kernel)
//   mixed: forward products on the exact fp32 SIMT GEMM, dW products on tcgen05
//   fp32 : everything on the SIMT GEMM
int texthead_mode(int E, int F) {      // 0 fp32, 1 mixed, 2 tc
  if ((E & 7) != 0 || (F & 7) != 0) return 0;
  const char* e = getenv("TGFR_TEXTHEAD_PRECISION");
  if (e && (e[0] == 'f' || e[0] == 's' || e[0] == '0')) return 0;
  if (e && e[0] == 'm') return 1;
  return 2;
}
bool use_tc(int E, int F) { return texthead_mode(E, F) != 0; }
struct TextheadLayout {
  size_t act, scales, tok_hi, tok_lo, w_hi, w_lo, ties, saved_total;   // saved
  size_t g, gscale, g_hi, g_lo, ws_total;                          // workspace
};
TextheadLayout texthead_layout(int B, int L, int E, int F) {
  TextheadLayout t{};
  const size_t acts = sizeof(float) * kConvs * (size_t)B * L * F;
  size_t o = 0;
  t.act = o; o += align_up(acts, 256);
  if (use_tc(E, F)) {
    t.scales = o; o += 256;
    t.tok_hi = o; o += align_up(2 * (size_t)B * L * E, 256);
    t.tok_lo = o; o += align_up(2 * (size_t)B * L * E, 256);
    t.w_hi = o; o += align_up(2 * (size_t)F * 9 * E, 256);          // K = 2 + 3 + 4 rows of E per feature
    t.w_lo = o; o += align_up(2 * (size_t)F * 9 * E, 256);
    t.ties = o; o += 256 + sizeof(uint2) * (size_t)kTieCap;         // count | items (forward scratch)
  }
  t.saved_total = o;
  o = 0;
  t.g = o; o += align_up(acts, 256);
  if (use_tc(E, F)) {
    t.gscale = o; o += 256;
    t.g_hi = o; o += align_up(2 * (size_t)kConvs * B * L * F, 256);
    t.g_lo = o; o += align_up(2 * (size_t)kConvs * B * L * F, 256);
  }
  t.ws_total = o;
  return t;
}
}  // namespace

size_t texthead_saved_bytes(int B, int L, int E, int F) { return texthead_layout(B, L, E, F).saved_total; }
size_t texthead_workspace_bytes(int B, int L, int E, int F) { return texthead_layout(B, L, E, F).ws_total; }

static int check(int B, int L, int E, int F, int words_num) {
  TGFR_REQUIRE(B > 0 && E > 0 && F > 0, "texthead: empty shape");
  TGFR_REQUIRE(words_num >= 5 && L == words_num - 1, "texthead: tokens must hold bert_words_num - 1 = %d positions, got %d",
               words_num - 1, L);
  return TGFR_OK;
}

int texthead_fwd(const float* tokens, const float* const* w, const float* const* bias, int B, int L, int E, int F,
                 int words_num, float* words, float* sent, void* saved, size_t saved_bytes, cudaStream_t st) {
  if (int rc = check(B, L, E, F, words_num)) return rc;
  const TextheadLayout lay = texthead_layout(B, L, E, F);
  TGFR_REQUIRE(saved && saved_bytes >= lay.saved_total, "texthead_fwd: saved buffer too small");
  const int seq = words_num - 4, T = seq + 2;
  uint8_t* sv = reinterpret_cast<uint8_t*>(saved);
  float* act = reinterpret_cast<float*>(sv + lay.act);
  const bool tc = use_tc(E, F);
  TGFR_REQUIRE(!tc || (reinterpret_cast<uintptr_t>(saved) & 255) == 0, "texthead_fwd: saved must be 256-byte aligned");
  Acts acts{};
  Split3Product prods[kConvs];
  const bool tc_fwd = texthead_mode(E, F) == 2;
  if (tc) {
    // fp16 hi / lo splits: the tokens (kept in `saved` for the backward) and the three weight tensors (one scale)
    float* scales = reinterpret_cast<float*>(sv + lay.scales);
    __half* tok_hi = reinterpret_cast<__half*>(sv + lay.tok_hi);
    __half* tok_lo = reinterpret_cast<__half*>(sv + lay.tok_lo);
    {
      const float* src[1] = {tokens};
      const int64_t ld[1] = {E};
      const int rows[1] = {B * L}, cols[1] = {E}, ldo[1] = {E};
      __half* hi[1] = {tok_hi};
      __half* lo[1] = {tok_lo};
      if (int rc = head_split_f16(1, src, ld, rows, cols, scales, hi, lo, ldo, st)) return rc;
    }
    const float* src[kConvs];
    int64_t ld[kConvs];
    int rows[kConvs], cols[kConvs], ldo[kConvs];
    __half* hi[kConvs];
    __half* lo[kConvs];
    size_t off = 0;
    for (int k = 0; k < kConvs; ++k) {
      const int K = k + 2;
      src[k] = w[k]; ld[k] = (int64_t)K * E; rows[k] = F; cols[k] = K * E; ldo[k] = K * E;
      hi[k] = reinterpret_cast<__half*>(sv + lay.w_hi) + off;
      lo[k] = reinterpret_cast<__half*>(sv + lay.w_lo) + off;
      off += (size_t)F * K * E;
    }
    if (tc_fwd)
      if (int rc = head_split_f16(kConvs, src, ld, rows, cols, scales + 2, hi, lo, ldo, st)) return rc;
    for (int k = 0; k < kConvs; ++k) {
      const int K = k + 2;
      prods[k] = Split3Product{tok_hi, tok_lo, hi[k], lo[k], E, (int64_t)K * E, F, B * L - K + 1, F, K * E,
                               act + (size_t)k * B * L * F, bias[k]};
    }
  }
  for (int k = 0; k < kConvs; ++k) {
    const int K = k + 2, rows = B * L - K + 1;          // every token position that still has K tokens after it
    float* a = act + (size_t)k * B * L * F;
    acts.a[k] = a;
    // the last K - 1 rows are never produced: keep them defined
    TGFR_CUDA_OK(cudaMemsetAsync(a + (size_t)rows * F, 0, sizeof(float) * (size_t)(K - 1) * F, st));
    if (!tc_fwd)
      if (int rc = sgemm_strided(tokens, E, 1, w[k], 1, (int64_t)K * E, a, F, 1, rows, F, K * E, bias[k], 1, st)) return rc;
  }
  if (tc_fwd) {   // act_K = relu(Win_K W_K^T + b_K): A = the overlapping token windows (K-major), B = W_K (K-major)
    const float* scales = reinterpret_cast<const float*>(sv + lay.scales);
    if (int rc = gemm_tc_split3_batched(prods, kConvs, 0, 1, 0, 0, 1.f, scales, scales + 2, 1, st)) return rc;
    // near-tied activations again, exactly (see texthead_mark_ties_kernel)
    TieList tl{reinterpret_cast<uint32_t*>(sv + lay.ties), reinterpret_cast<uint2*>(sv + lay.ties + 256)};
    TGFR_CUDA_OK(cudaMemsetAsync(tl.count, 0, sizeof(uint32_t), st));
    const int mthreads = F >= 256 ? 256 : ((F + 31) & ~31);
    texthead_mark_ties_kernel<<<dim3(seq + 2, B), mthreads, 0, st>>>(acts, L, F, T, seq, tl);
    TGFR_LAUNCH_OK();
    RefineArgs ra{};
    for (int k = 0; k < kConvs; ++k) {
      ra.w[k] = w[k];
      ra.bias[k] = bias[k];
      ra.act[k] = act + (size_t)k * B * L * F;
    }
    texthead_refine_ties_kernel<<<296, 256, 0, st>>>(tl, tokens, ra, E, F);
    TGFR_LAUNCH_OK();
  }
  const int threads = F >= 256 ? 256 : ((F + 31) & ~31);
  texthead_combine_fwd_kernel<<<dim3(T + 1, B), threads, 0, st>>>(acts, L, F, T, seq, words, sent);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int texthead_bwd(const float* tokens, const float* gwords, const float* gsent, int B, int L, int E, int F, int words_num,
                 float* const* dw, float* const* db, void* ws, size_t ws_bytes, const void* saved, size_t saved_bytes,
                 cudaStream_t st) {
  if (int rc = check(B, L, E, F, words_num)) return rc;
  const TextheadLayout lay = texthead_layout(B, L, E, F);
  TGFR_REQUIRE(saved && saved_bytes >= lay.saved_total, "texthead_bwd: saved buffer too small");
  TGFR_REQUIRE(ws && ws_bytes >= lay.ws_total, "texthead_bwd: workspace too small");
  const int seq = words_num - 4, T = seq + 2;
  const uint8_t* sv = reinterpret_cast<const uint8_t*>(saved);
  uint8_t* wsb = reinterpret_cast<uint8_t*>(ws);
  const float* act = reinterpret_cast<const float*>(sv + lay.act);
  float* G = reinterpret_cast<float*>(wsb + lay.g);
  const bool tc = use_tc(E, F);
  TGFR_REQUIRE(!tc || ((reinterpret_cast<uintptr_t>(saved) | reinterpret_cast<uintptr_t>(ws)) & 255) == 0,
               "texthead_bwd: saved / workspace must be 256-byte aligned");
  TGFR_CUDA_OK(cudaMemsetAsync(G, 0, sizeof(float) * kConvs * (size_t)B * L * F, st));
  Acts acts{};
  Grads gr{};
  for (int k = 0; k < kConvs; ++k) {
    acts.a[k] = act + (size_t)k * B * L * F;
    gr.g[k] = G + (size_t)k * B * L * F;
    gr.db[k] = db[k];
    TGFR_CUDA_OK(cudaMemsetAsync(db[k], 0, sizeof(float) * F, st));
  }
  const int threads = F >= 256 ? 256 : ((F + 31) & ~31);
  texthead_combine_bwd_kernel<<<dim3(T, B), threads, 0, st>>>(acts, L, F, T, seq, gwords, gsent, gr);
  TGFR_LAUNCH_OK();
  texthead_bias_grad_kernel<<<dim3(kConvs, B), threads, 0, st>>>(gr, L, F);
  TGFR_LAUNCH_OK();
  if (tc) {
    // dW_K [F, K E] = G_K^T . Win_K: A = G_K read MN-major, B = the overlapping token windows read MN-major
    float* gscale = reinterpret_cast<float*>(wsb + lay.gscale);
    __half* g_hi = reinterpret_cast<__half*>(wsb + lay.g_hi);
    __half* g_lo = reinterpret_cast<__half*>(wsb + lay.g_lo);
    {
      const float* src[1] = {G};
      const int64_t ld[1] = {F};
      const int rows[1] = {kConvs * B * L}, cols[1] = {F}, ldo[1] = {F};
      __half* hi[1] = {g_hi};
      __half* lo[1] = {g_lo};
      if (int rc = head_split_f16(1, src, ld, rows, cols, gscale, hi, lo, ldo, st)) return rc;
    }
    const __half* tok_hi = reinterpret_cast<const __half*>(sv + lay.tok_hi);
    const __half* tok_lo = reinterpret_cast<const __half*>(sv + lay.tok_lo);
    const float* scales = reinterpret_cast<const float*>(sv + lay.scales);
    Split3Product prods[kConvs];
    for (int k = 0; k < kConvs; ++k) {
      const int K = k + 2;
      const size_t goff = (size_t)k * B * L * F;
      prods[k] = Split3Product{g_hi + goff, g_lo + goff, tok_hi, tok_lo, F, E, (int64_t)K * E, F, K * E, B * L - K + 1,
                               dw[k], nullptr};
    }
    return gemm_tc_split3_batched(prods, kConvs, 1, 0, 1, 1, 1.f, gscale, scales, 0, st);
  }
  for (int k = 0; k < kConvs; ++k) {
    const int K = k + 2, rows = B * L - K + 1;
    // dW_K [F, K E] = G_K^T [F, rows] . Win_K [rows, K E]
    if (int rc = sgemm_strided(gr.g[k], 1, F, tokens, E, 1, dw[k], (int64_t)K * E, 1, F, K * E, rows, nullptr, 0, st))
      return rc;
  }
  return TGFR_OK;
}

}  // namespace tgfr

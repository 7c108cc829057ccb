// fp32 SIMT kernels for the dense part of the path: cosine score matrices (sent/global/Clip
// losses), the two-direction B x B cross entropy, the margin heads (ArcFace / MagFace) and the
// row-wise cross entropy / focal loss.  The contraction is a strided 64x64x16 register-tiled
// SGEMM with fused normalisation scales; everything else is bandwidth-bound glue around it.
#include <cuda_fp16.h>

#include <stdlib.h>

#include "common.cuh"

namespace tgfr {

namespace {

// --------------------------------------------------------------------------------------------
// row norms: one warp per row, arbitrary strides
// --------------------------------------------------------------------------------------------
__global__ void rownorm_kernel(const float* __restrict__ x, int64_t s_row, int64_t s_col, int rows, int cols,
                               float* __restrict__ norm) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* p = x + (int64_t)row * s_row;
  float acc = 0.f;
  for (int c = lane; c < cols; c += 32) {
    const float v = __ldg(p + (int64_t)c * s_col);
    acc = fmaf(v, v, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) norm[row] = sqrtf(acc);
}

// column norms of a row-major [rows, cols] matrix (MagLinear.weight [Din, C]): coalesced over cols
__global__ void colnorm_kernel(const float* __restrict__ x, int64_t s_row, int rows, int cols,
                               float* __restrict__ norm) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float acc = 0.f;
  for (int r = 0; r < rows; ++r) {
    const float v = __ldg(x + (int64_t)r * s_row + c);
    acc = fmaf(v, v, acc);
  }
  norm[c] = sqrtf(acc);
}

}  // namespace

int launch_norms(const float* x, int64_t s_vec, int64_t s_elem, int nvec, int len, float* norm, cudaStream_t st) {
  if (s_vec == 1 && s_elem != 1) {
    colnorm_kernel<<<ceil_div(nvec, 256), 256, 0, st>>>(x, s_elem, len, nvec, norm);
  } else {
    rownorm_kernel<<<ceil_div(nvec, 8), 256, 0, st>>>(x, s_vec, s_elem, nvec, len, norm);
  }
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

namespace {

// --------------------------------------------------------------------------------------------
// strided SGEMM:  C[m,n] = epilogue( sum_k A(m,k) * B(k,n) * bk[k] )
// --------------------------------------------------------------------------------------------
enum EpiMode {
  kEpiScale = 0,   // alpha * acc * inv(am[m]) * inv(bn[n])       (am/bn are norms, clamped at norm_eps)
  kEpiCosine = 1,  // alpha * acc / max(am[m]*bn[n], eps), -inf on class collisions
  kEpiClamp = 2,   // alpha * clamp(acc * inv(am[m]) * inv(bn[n]), -1, 1)
  kEpiBiasRelu = 3,  // max(acc + bias[n], 0)   (bias may be NULL; relu optional: alpha < 0 switches it off)
};

struct Gemm {
  const float* A; int64_t sAm, sAk;
  const float* B; int64_t sBk, sBn;
  float* C; int64_t sCm, sCn;
  int M, N, K;
  float alpha;
  const float* bk;      // optional norms over k: B(k,n) is divided by max(bk[k], norm_eps)
  const float* am;      // optional norms over m
  const float* bn;      // optional norms over n
  float norm_eps;       // F.normalize eps (1e-12)
  float eps;            // cosine clamp
  int mode;
  const int64_t* ids_m; // optional class ids for the -inf mask (kEpiCosine)
  const int64_t* ids_n;
  int diag_off;
  const float* bias;    // kEpiBiasRelu: per-column bias
};

constexpr int BM = 64, BN = 64, BK = 16, GT = 256;

__device__ __forceinline__ float gemm_epilogue(const Gemm& g, float acc, int m, int n, float na, float nb) {
  float v;
  if (g.mode == kEpiBiasRelu) {
    v = acc + (g.bias ? __ldg(g.bias + n) : 0.f);
    return g.alpha < 0.f ? v : fmaxf(v, 0.f);
  }
  if (g.mode == kEpiCosine) {
    v = acc / fmaxf(na * nb, g.eps) * g.alpha;
    if (g.ids_m && g.ids_n && (m + g.diag_off) != n && __ldg(g.ids_m + m) == __ldg(g.ids_n + n)) v = -INFINITY;
  } else {
    v = acc;
    if (g.am) v /= fmaxf(na, g.norm_eps);
    if (g.bn) v /= fmaxf(nb, g.norm_eps);
    if (g.mode == kEpiClamp) v = fminf(fmaxf(v, -1.f), 1.f);
    v *= g.alpha;
  }
  return v;
}

// Latency-bound shapes (the B x B sentence-loss matrices: 2*B*B*D flops on < 1 MB): one warp per 4 x 4 output
// tile, lanes split K, so that a 128 x 128 product already fills the GPU with 128 CTAs instead of four 64 x 64 tiles.
__global__ void __launch_bounds__(256) sgemm_small_kernel(const Gemm g) {
  const int lane = threadIdx.x & 31;
  const int tiles_n = (g.N + 3) >> 2;
  const int tile = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int m0 = (tile / tiles_n) << 2, n0 = (tile % tiles_n) << 2;
  if (m0 >= g.M) return;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k = lane; k < g.K; k += 32) {
    float a[4], b[4];
    const float bden = g.bk ? fmaxf(__ldg(g.bk + k), g.norm_eps) : 1.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[i] = (m0 + i < g.M) ? __ldg(g.A + (int64_t)(m0 + i) * g.sAm + (int64_t)k * g.sAk) : 0.f;
      b[i] = (n0 + i < g.N) ? __ldg(g.B + (int64_t)k * g.sBk + (int64_t)(n0 + i) * g.sBn) : 0.f;
      if (g.bk) b[i] = b[i] / bden;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = warp_sum(acc[i][j]);
  if (lane < 16) {
    const int i = lane >> 2, j = lane & 3;
    const int m = m0 + i, n = n0 + j;
    if (m < g.M && n < g.N) {
      float v = 0.f;
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          if (ii == i && jj == j) v = acc[ii][jj];
      const float na = g.am ? __ldg(g.am + m) : 1.f, nb = g.bn ? __ldg(g.bn + n) : 1.f;
      g.C[(int64_t)m * g.sCm + (int64_t)n * g.sCn] = gemm_epilogue(g, v, m, n, na, nb);
    }
  }
}

__global__ void __launch_bounds__(GT) sgemm_kernel(const Gemm g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kfast = (g.sAk == 1), b_nfast = (g.sBn == 1);

  for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int mm, kk;
      if (a_kfast) { kk = tid & 15; mm = (tid >> 4) + 16 * i; }
      else         { mm = tid & 63; kk = (tid >> 6) + 4 * i; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < g.M && k < g.K) ? __ldg(g.A + (int64_t)m * g.sAm + (int64_t)k * g.sAk) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int nn, kk;
      if (b_nfast) { nn = tid & 63; kk = (tid >> 6) + 4 * i; }
      else         { kk = tid & 15; nn = (tid >> 4) + 16 * i; }
      const int n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < g.N && k < g.K) {
        v = __ldg(g.B + (int64_t)k * g.sBk + (int64_t)n * g.sBn);
        if (g.bk) v /= fmaxf(__ldg(g.bk + k), g.norm_eps);
      }
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
    const float na = g.am ? __ldg(g.am + m) : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      const float nb = g.bn ? __ldg(g.bn + n) : 1.f;
      g.C[(int64_t)m * g.sCm + (int64_t)n * g.sCn] = gemm_epilogue(g, acc[i][j], m, n, na, nb);
    }
  }
}

int launch_gemm(const Gemm& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return TGFR_OK;
  if ((int64_t)g.M * g.N <= 512 * 512 && g.K <= 1024) {
    const int tiles = ceil_div(g.M, 4) * ceil_div(g.N, 4);
    sgemm_small_kernel<<<ceil_div(tiles, 8), 256, 0, st>>>(g);
    TGFR_LAUNCH_OK();
    return TGFR_OK;
  }
  const dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, BM));
  sgemm_kernel<<<grid, GT, 0, st>>>(g);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

// --------------------------------------------------------------------------------------------
// d(normalize(v))/dv projection:  out = (g - <g, v^> v^) / |v|,  one warp per vector
// --------------------------------------------------------------------------------------------
__global__ void normalize_bwd_kernel(const float* __restrict__ ghat, int64_t g_sv, const float* __restrict__ v,
                                     int64_t v_sv, int64_t v_se, const float* __restrict__ norm, float norm_eps,
                                     int nvec, int len, float* __restrict__ out, int64_t o_sv, int64_t o_se) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= nvec) return;
  const float n = fmaxf(norm[row], norm_eps), inv = 1.f / n;
  const float* gp = ghat + (int64_t)row * g_sv;
  const float* vp = v + (int64_t)row * v_sv;
  float dot = 0.f;
  for (int c = lane; c < len; c += 32) dot = fmaf(gp[c], __ldg(vp + (int64_t)c * v_se) * inv, dot);
  dot = warp_sum(dot);
  float* op = out + (int64_t)row * o_sv;
  for (int c = lane; c < len; c += 32) {
    const float vh = __ldg(vp + (int64_t)c * v_se) * inv;
    op[(int64_t)c * o_se] = (gp[c] - dot * vh) * inv;
  }
}

// --------------------------------------------------------------------------------------------
// two-direction cross entropy over a [Bx, By] block
// --------------------------------------------------------------------------------------------
__global__ void pair_ce_stats_kernel(const float* __restrict__ sc, int Bx, int By, int diag_off,
                                     float* __restrict__ rowlse, float* __restrict__ colmax,
                                     float* __restrict__ colsum, float* __restrict__ diag) {
  __shared__ float scratch[32];
  const int id = blockIdx.x;
  const bool is_row = id < Bx;
  const int n = is_row ? By : Bx;
  const float* base = is_row ? sc + (int64_t)id * By : sc + (id - Bx);
  const int64_t stride = is_row ? 1 : By;
  float m = -INFINITY;
  for (int k = threadIdx.x; k < n; k += blockDim.x) m = fmaxf(m, base[k * stride]);
  m = block_max(m, scratch);
  const float mm = (m == -INFINITY) ? 0.f : m;
  float s = 0.f;
  for (int k = threadIdx.x; k < n; k += blockDim.x) s += expf(base[k * stride] - mm);
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) {
    if (is_row) {
      rowlse[id] = mm + logf(s);
      const int j = id + diag_off;
      diag[id] = (j >= 0 && j < By) ? sc[(int64_t)id * By + j] : 0.f;
    } else {
      colmax[id - Bx] = mm;
      colsum[id - Bx] = s;
    }
  }
}

__global__ void pair_ce_finish_kernel(const float* __restrict__ rowlse, const float* __restrict__ colmax,
                                      const float* __restrict__ colsum, const float* __restrict__ diag, int Bx,
                                      int By, int diag_off, float inv_b, float* __restrict__ losses,
                                      float* __restrict__ collse) {
  __shared__ float scratch[32];
  for (int j = threadIdx.x; j < By; j += blockDim.x) collse[j] = colmax[j] + logf(colsum[j]);
  __syncthreads();
  float l0 = 0.f, l1 = 0.f;
  for (int b = threadIdx.x; b < Bx; b += blockDim.x) {
    l0 += rowlse[b] - diag[b];
    const int j = b + diag_off;
    if (j >= 0 && j < By) l1 += collse[j] - diag[b];
  }
  l0 = block_sum(l0, scratch);
  l1 = block_sum(l1, scratch);
  if (threadIdx.x == 0) {
    losses[0] = l0 * inv_b;
    losses[1] = l1 * inv_b;
  }
}

__global__ void pair_ce_bwd_kernel(const float* __restrict__ sc, const float* __restrict__ rowlse,
                                   const float* __restrict__ collse, const float* __restrict__ g0p,
                                   const float* __restrict__ g1p, int Bx, int By, int diag_off, float inv_b,
                                   float* __restrict__ gs) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)Bx * By) return;
  const int b = (int)(idx / By), j = (int)(idx - (int64_t)b * By);
  const float g0 = g0p ? *g0p : 1.f, g1 = g1p ? *g1p : 1.f;
  const float s = sc[idx];
  const float on = (j == b + diag_off) ? 1.f : 0.f;
  const float pr = expf(s - rowlse[b]), pc = expf(s - collse[j]);
  gs[idx] = (g0 * (pr - on) + g1 * (pc - on)) * inv_b;
}

// --------------------------------------------------------------------------------------------
// row-wise cross entropy over dense logits, focal loss
// --------------------------------------------------------------------------------------------
__global__ void ce_rows_stats_kernel(const float* __restrict__ lg, int64_t sr, const int64_t* __restrict__ labels,
                                     int C, int class_off, float* __restrict__ rowmax, float* __restrict__ rowsum,
                                     float* __restrict__ tgt) {
  __shared__ float scratch[32];
  const int b = blockIdx.x;
  const float* p = lg + (int64_t)b * sr;
  float m = -INFINITY, s = 0.f;     // online softmax: one pass over the row
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = p[c];
    if (v > m) { s = s * expf(m - v) + 1.f; m = v; }
    else       { s += expf(v - m); }
  }
  const float bm = block_max(m, scratch);
  const float bmm = (bm == -INFINITY) ? 0.f : bm;
  s = (m == -INFINITY) ? 0.f : s * expf(m - bmm);
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) {
    rowmax[b] = bmm;
    rowsum[b] = s;
    const int64_t y = labels[b] - class_off;
    tgt[b] = (y >= 0 && y < C) ? p[y] : 0.f;
  }
}

__global__ void focal_finish_kernel(const float* __restrict__ rowmax, const float* __restrict__ rowsum,
                                    const float* __restrict__ tgt, int B, float gamma, float* __restrict__ out,
                                    float* __restrict__ lse) {
  __shared__ float scratch[32];
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float l = rowmax[b] + logf(rowsum[b]);
    lse[b] = l;
    acc += l - tgt[b];
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) {
    const float logp = acc / (float)B;          // CrossEntropyLoss(mean), losses.py:322
    const float pt = expf(-logp);               // :323
    const float om = 1.f - pt;
    float loss, dl;
    if (gamma == 0.f) { loss = logp; dl = 1.f; }
    else {
      loss = powf(om, gamma) * logp;            // :324
      dl = powf(om, gamma) + gamma * powf(om, gamma - 1.f) * pt * logp;
    }
    out[0] = logp; out[1] = loss; out[2] = dl;
  }
}

__global__ void ce_rows_bwd_kernel(const float* __restrict__ lg, int64_t sr, const int64_t* __restrict__ labels,
                                   const float* __restrict__ lse, const float* __restrict__ coef,
                                   const float* __restrict__ gout, int B, int C, int class_off,
                                   float* __restrict__ gl, int64_t g_sr) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float k = (coef ? *coef : 1.f) * (gout ? *gout : 1.f) / (float)B;
  const float on = ((int64_t)c + class_off == labels[b]) ? 1.f : 0.f;
  gl[(int64_t)b * g_sr + c] = k * (expf(lg[(int64_t)b * sr + c] - lse[b]) - on);
}

// --------------------------------------------------------------------------------------------
// cosine_similarity(x1, x2, dim=1, eps) of models/losses.py:12-16: sum(x1 x2) / max(|x1| |x2|, eps), one warp per row;
// stats[row] = (w12, |x1|, |x2|) for the backward.  (The pair-scoring kernel of scoring.cu clamps each norm
// separately, like nn.CosineSimilarity; this one clamps the product, like the reference's free function.)
// --------------------------------------------------------------------------------------------
__global__ void cosine_rows_ref_fwd_kernel(const float* __restrict__ x1, int64_t s1r, int64_t s1d,
                                           const float* __restrict__ x2, int64_t s2r, int64_t s2d, int64_t n, int D,
                                           float eps, float* __restrict__ out, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* a = x1 + row * s1r;
  const float* b = x2 + row * s2r;
  float w12 = 0.f, w1 = 0.f, w2 = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float u = a[(int64_t)d * s1d], v = b[(int64_t)d * s2d];
    w12 = fmaf(u, v, w12);
    w1 = fmaf(u, u, w1);
    w2 = fmaf(v, v, w2);
  }
  w12 = warp_sum(w12);
  w1 = sqrtf(warp_sum(w1));
  w2 = sqrtf(warp_sum(w2));
  if (lane == 0) {
    out[row] = w12 / fmaxf(w1 * w2, eps);
    stats[3 * row + 0] = w12;
    stats[3 * row + 1] = w1;
    stats[3 * row + 2] = w2;
  }
}

__global__ void cosine_rows_ref_bwd_kernel(const float* __restrict__ x1, int64_t s1r, int64_t s1d,
                                           const float* __restrict__ x2, int64_t s2r, int64_t s2d, int64_t n, int D,
                                           float eps, const float* __restrict__ stats, const float* __restrict__ g,
                                           float* __restrict__ d1, float* __restrict__ d2) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float w12 = stats[3 * row], w1 = stats[3 * row + 1], w2 = stats[3 * row + 2], go = g[row];
  const float prod = w1 * w2;
  const bool live = prod > eps;                    // clamp(min=eps) passes no gradient to the norms below eps
  const float inv = 1.f / fmaxf(prod, eps);
  // d/dx1 = g (x2 inv - [live] w12 inv x1 / w1^2),  symmetric in x2
  const float k1 = (live && w1 > 0.f) ? w12 * inv / (w1 * w1) : 0.f;
  const float k2 = (live && w2 > 0.f) ? w12 * inv / (w2 * w2) : 0.f;
  const float* a = x1 + row * s1r;
  const float* b = x2 + row * s2r;
  for (int d = lane; d < D; d += 32) {
    const float u = a[(int64_t)d * s1d], v = b[(int64_t)d * s2d];
    if (d1) d1[row * D + d] = go * (v * inv - k1 * u);
    if (d2) d2[row * D + d] = go * (u * inv - k2 * v);
  }
}

// --------------------------------------------------------------------------------------------
// merge of per-shard online-softmax statistics after ONE all-gather: in [n][K][M] with K = 2 (max, sum-exp) or
// 3 (+ a value that is simply summed: the target logit, owned by one shard) -> out [K][M]
// --------------------------------------------------------------------------------------------
__global__ void merge_softmax_stats_kernel(const float* __restrict__ in, int n, int K, int M, float* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  float gmax = -INFINITY;
  for (int r = 0; r < n; ++r) gmax = fmaxf(gmax, in[((int64_t)r * K) * M + j]);
  float gsum = 0.f, gt = 0.f;
  for (int r = 0; r < n; ++r) {
    const float m = in[((int64_t)r * K) * M + j], sm = in[((int64_t)r * K + 1) * M + j];
    gsum += (sm == 0.f) ? 0.f : sm * expf(m - gmax);
    if (K > 2) gt += in[((int64_t)r * K + 2) * M + j];
  }
  out[j] = gmax;
  out[M + j] = gsum;
  if (K > 2) out[2 * M + j] = gt;
}

// --------------------------------------------------------------------------------------------
// MagLoss blend + cross entropy (magface.py:131-135): output[b,c] = (c == label_b) ? cos_m[b,c] : cos[b,c]
// is never materialised -- the row statistics read the two logit tensors directly and, as a by-product,
// write the one_hot [B,C] tensor the reference returns.
// --------------------------------------------------------------------------------------------
__global__ void mag_ce_stats_kernel(const float* __restrict__ lc, const float* __restrict__ lm, int64_t sr,
                                    const int64_t* __restrict__ labels, int C, float* __restrict__ rowmax,
                                    float* __restrict__ rowsum, float* __restrict__ tgt, float* __restrict__ one_hot) {
  __shared__ float scratch[32];
  const int b = blockIdx.x;
  const float* pc = lc + (int64_t)b * sr;
  const float* pm = lm + (int64_t)b * sr;
  const int64_t y = labels[b];
  float m = -INFINITY, s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = (c == y) ? pm[c] : pc[c];
    if (one_hot) one_hot[(int64_t)b * C + c] = (c == y) ? 1.f : 0.f;
    if (v > m) { s = s * expf(m - v) + 1.f; m = v; }
    else       { s += expf(v - m); }
  }
  const float bm = block_max(m, scratch);
  const float bmm = (bm == -INFINITY) ? 0.f : bm;
  s = (m == -INFINITY) ? 0.f : s * expf(m - bmm);
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) {
    rowmax[b] = bmm;
    rowsum[b] = s;
    tgt[b] = (y >= 0 && y < C) ? pm[y] : 0.f;
  }
}

// g_cos[b,c] = k (p - 0) off the label column, 0 on it; g_cosm[b,c] = k (p - 1) on the label column, 0 elsewhere
__global__ void mag_ce_bwd_kernel(const float* __restrict__ lc, const float* __restrict__ lm, int64_t sr,
                                  const int64_t* __restrict__ labels, const float* __restrict__ lse,
                                  const float* __restrict__ gout, int B, int C, float* __restrict__ g_cos,
                                  float* __restrict__ g_cosm) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float k = (gout ? *gout : 1.f) / (float)B;
  const bool on = (int64_t)c == labels[b];
  const int64_t o = (int64_t)b * sr + c;
  const float v = on ? lm[o] : lc[o];
  const float g = k * (expf(v - lse[b]) - (on ? 1.f : 0.f));
  g_cos[(int64_t)b * C + c] = on ? 0.f : g;
  g_cosm[(int64_t)b * C + c] = on ? g : 0.f;
}

// --------------------------------------------------------------------------------------------
// ArcFace margin on the label column (metrics.py:45-57) and its backward correction
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ float arc_phi(float c, float cm, float sm, float th, float mm, int easy, float* dphi) {
  const float one_m = 1.f - c * c;
  const float sine = sqrtf(fminf(fmaxf(one_m, 0.f), 1.f));
  const float phi = c * cm - sine * sm;
  const bool use = easy ? (c > 0.f) : (c > th);
  if (dphi) {
    const bool inside = one_m >= 0.f && one_m <= 1.f;
    *dphi = use ? (cm + (inside ? c / sine : 0.f) * sm) : 1.f;
  }
  return use ? phi : (easy ? c : c - mm);
}

__global__ void arc_apply_kernel(float* __restrict__ lg, int64_t sr, const int64_t* __restrict__ labels, int B,
                                 int C, int class_off, float s, float cm, float sm, float th, float mm, int easy,
                                 float* __restrict__ cos_t) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t y = labels[b] - class_off;
  if (y < 0 || y >= C) { cos_t[b] = nanf(""); return; }
  float* p = lg + (int64_t)b * sr + y;
  const float c = *p / s;
  cos_t[b] = c;
  *p = s * arc_phi(c, cm, sm, th, mm, easy, nullptr);
}

// dXhat[b,:] += k_b * w^_y ; dWhat[y,:] += k_b * x^_b   with k_b = (phi'(cos_t) - 1) * s * g[b,y]
__global__ void arc_fix_kernel(const float* __restrict__ x, int64_t x_sr, const float* __restrict__ w, int64_t w_sc,
                               int64_t w_sk, const float* __restrict__ xnorm, const float* __restrict__ wnorm,
                               const int64_t* __restrict__ labels, const float* __restrict__ cos_t,
                               const float* __restrict__ g, int64_t g_sr, int C, int Din, int class_off, float s,
                               float cm, float sm, float th, float mm, int easy, float* __restrict__ dxh,
                               float* __restrict__ dwh) {
  const int b = blockIdx.x;
  const int64_t y = labels[b] - class_off;
  if (y < 0 || y >= C) return;
  float dphi;
  arc_phi(cos_t[b], cm, sm, th, mm, easy, &dphi);
  const float k = (dphi - 1.f) * s * g[(int64_t)b * g_sr + y];
  if (k == 0.f) return;
  const float ix = 1.f / fmaxf(xnorm[b], 1e-12f), iw = 1.f / fmaxf(wnorm[y], 1e-12f);
  for (int d = threadIdx.x; d < Din; d += blockDim.x) {
    if (dxh) dxh[(int64_t)b * Din + d] += k * w[y * w_sc + (int64_t)d * w_sk] * iw;
    atomicAdd(dwh + y * Din + d, k * x[(int64_t)b * x_sr + d] * ix);
  }
}

// --------------------------------------------------------------------------------------------
// MagFace margin (magface.py:95-106), elementwise over [B, C]
// --------------------------------------------------------------------------------------------
__global__ void mag_fwd_kernel(const float* __restrict__ cs, const float* __restrict__ margin, int B, int C,
                               float scale, int easy, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mar = margin[b];
  float cm, sm;
  sincosf(mar, &sm, &cm);
  const float ct = cs[(int64_t)b * C + c] / scale;
  const float st = sqrtf(1.f - ct * ct);
  float v = ct * cm - st * sm;
  if (easy) v = (ct > 0.f) ? v : ct;
  else {
    const float th = cosf(3.14159265358979323846f - mar), mm = sinf(3.14159265358979323846f - mar) * mar;
    v = (ct > th) ? v : ct - mm;
  }
  out[(int64_t)b * C + c] = scale * v;
}

__global__ void mag_bwd_kernel(const float* __restrict__ cs, const float* __restrict__ margin,
                               const float* __restrict__ g_cos, const float* __restrict__ g_cosm, int B, int C,
                               float scale, int easy, float* __restrict__ gtotal, float* __restrict__ gmargin) {
  __shared__ float scratch[32];
  const int b = blockIdx.x;
  const float mar = margin[b];
  float cm, sm;
  sincosf(mar, &sm, &cm);
  const float pi = 3.14159265358979323846f;
  const float th = cosf(pi - mar);
  const float alt_dm = cosf(pi - mar) * mar - sinf(pi - mar);   // d/dmar (cos - sin(pi-mar)*mar)
  float gm = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int64_t idx = (int64_t)b * C + c;
    const float ct = cs[idx] / scale;
    const float st = sqrtf(1.f - ct * ct);
    const bool use = easy ? (ct > 0.f) : (ct > th);
    const float gc = g_cos ? g_cos[idx] : 0.f, gcm = g_cosm ? g_cosm[idx] : 0.f;
    float dsel_dcos = 1.f, dsel_dm = easy ? 0.f : alt_dm;
    if (use) {
      dsel_dcos = cm + (ct / st) * sm;
      dsel_dm = -ct * sm - st * cm;
    }
    // both outputs are scale * f(cos); gtotal is the gradient w.r.t. (scale * cos)
    gtotal[idx] = gc + (gcm != 0.f ? gcm * dsel_dcos : 0.f);
    gm += (gcm != 0.f) ? gcm * scale * dsel_dm : 0.f;
  }
  gm = block_sum(gm, scratch);
  if (threadIdx.x == 0) gmargin[b] = gm;
}

}  // namespace

// ============================================================================================
// host-side entry points used by capi.cu
// ============================================================================================
int cosine_scores_fwd(const float* x, int64_t x_sr, const float* y, int64_t y_sr, int Bx, int By, int D, float scale,
                      int normalise, float eps, const int64_t* ids_x, const int64_t* ids_y, int diag_off,
                      float* scores, float* xnorm, float* ynorm, cudaStream_t st) {
  TGFR_REQUIRE(Bx > 0 && By > 0 && D > 0, "cosine_scores: empty shape");
  TGFR_REQUIRE(normalise || !(ids_x && ids_y), "cosine_scores: the class-id mask needs normalise != 0");
  if (normalise) {
    TGFR_REQUIRE(xnorm && ynorm, "cosine_scores: xnorm/ynorm required when normalise != 0");
    if (int rc = launch_norms(x, x_sr, 1, Bx, D, xnorm, st)) return rc;
    if (int rc = launch_norms(y, y_sr, 1, By, D, ynorm, st)) return rc;
  }
  Gemm g{};
  g.A = x; g.sAm = x_sr; g.sAk = 1;
  g.B = y; g.sBk = 1; g.sBn = y_sr;
  g.C = scores; g.sCm = By; g.sCn = 1;
  g.M = Bx; g.N = By; g.K = D; g.alpha = scale; g.norm_eps = 1e-12f; g.eps = eps;
  if (normalise) { g.mode = kEpiCosine; g.am = xnorm; g.bn = ynorm; }
  else g.mode = kEpiScale;
  g.ids_m = ids_x; g.ids_n = ids_y; g.diag_off = diag_off;
  return launch_gemm(g, st);
}

size_t cosine_workspace_bytes(int Bx, int By, int D) {
  return sizeof(float) * ((size_t)Bx * D + (size_t)By * D);
}

int cosine_scores_bwd(const float* x, int64_t x_sr, const float* y, int64_t y_sr, int Bx, int By, int D, float scale,
                      int normalise, float eps, const float* xnorm, const float* ynorm, const float* gs, float* dx,
                      float* dy, void* ws, size_t ws_bytes, cudaStream_t st) {
  (void)eps;
  TGFR_REQUIRE(ws_bytes >= cosine_workspace_bytes(Bx, By, D), "cosine_scores_bwd: workspace too small");
  float* dxh = reinterpret_cast<float*>(ws);
  float* dyh = dxh + (size_t)Bx * D;
  if (dx) {
    Gemm g{};  // dXhat = scale * G . Yhat
    g.A = gs; g.sAm = By; g.sAk = 1;
    g.B = y; g.sBk = y_sr; g.sBn = 1;
    g.C = normalise ? dxh : dx; g.sCm = D; g.sCn = 1;
    g.M = Bx; g.N = D; g.K = By; g.alpha = scale; g.norm_eps = 1e-12f; g.mode = kEpiScale;
    g.bk = normalise ? ynorm : nullptr;
    if (int rc = launch_gemm(g, st)) return rc;
    if (normalise) {
      normalize_bwd_kernel<<<ceil_div(Bx, 8), 256, 0, st>>>(dxh, D, x, x_sr, 1, xnorm, 1e-12f, Bx, D, dx, D, 1);
      TGFR_LAUNCH_OK();
    }
  }
  if (dy) {
    Gemm g{};  // dYhat = scale * G^T . Xhat
    g.A = gs; g.sAm = 1; g.sAk = By;
    g.B = x; g.sBk = x_sr; g.sBn = 1;
    g.C = normalise ? dyh : dy; g.sCm = D; g.sCn = 1;
    g.M = By; g.N = D; g.K = Bx; g.alpha = scale; g.norm_eps = 1e-12f; g.mode = kEpiScale;
    g.bk = normalise ? xnorm : nullptr;
    if (int rc = launch_gemm(g, st)) return rc;
    if (normalise) {
      normalize_bwd_kernel<<<ceil_div(By, 8), 256, 0, st>>>(dyh, D, y, y_sr, 1, ynorm, 1e-12f, By, D, dy, D, 1);
      TGFR_LAUNCH_OK();
    }
  }
  return TGFR_OK;
}

int pair_ce_stats(const float* scores, int Bx, int By, int diag_off, float* rowlse, float* colmax, float* colsum,
                  float* diag, cudaStream_t st) {
  TGFR_REQUIRE(Bx > 0 && By > 0, "pair_ce_stats: empty shape");
  pair_ce_stats_kernel<<<Bx + By, 128, 0, st>>>(scores, Bx, By, diag_off, rowlse, colmax, colsum, diag);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int pair_ce_finish(const float* rowlse, const float* colmax, const float* colsum, const float* diag, int Bx, int By,
                   int diag_off, float inv_b, float* losses, float* collse, cudaStream_t st) {
  pair_ce_finish_kernel<<<1, 256, 0, st>>>(rowlse, colmax, colsum, diag, Bx, By, diag_off, inv_b, losses, collse);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int pair_ce_bwd(const float* scores, const float* rowlse, const float* collse, const float* g0, const float* g1,
                int Bx, int By, int diag_off, float inv_b, float* gscores, cudaStream_t st) {
  const int64_t n = (int64_t)Bx * By;
  pair_ce_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scores, rowlse, collse, g0, g1, Bx, By, diag_off,
                                                                  inv_b, gscores);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

// plain strided fp32 product with an optional bias / ReLU epilogue (texthead.cu):
// C[m,n] = act(sum_k A[m sAm + k sAk] B[k sBk + n sBn] + bias[n]); rows of A may overlap (sAm < K sAk)
int sgemm_strided(const float* A, int64_t sAm, int64_t sAk, const float* Bm, int64_t sBk, int64_t sBn, float* C,
                  int64_t sCm, int64_t sCn, int M, int N, int K, const float* bias, int relu, cudaStream_t st) {
  Gemm g{};
  g.A = A; g.sAm = sAm; g.sAk = sAk;
  g.B = Bm; g.sBk = sBk; g.sBn = sBn;
  g.C = C; g.sCm = sCm; g.sCn = sCn;
  g.M = M; g.N = N; g.K = K; g.alpha = relu ? 1.f : -1.f; g.norm_eps = 1e-12f; g.mode = kEpiBiasRelu; g.bias = bias;
  return launch_gemm(g, st);
}

// gemm_tc.cu
bool head_tc_supported(int B, int C, int Din);
int gemm_tc(const __half* A, int a_mn, int64_t lda, const __half* Bm, int b_mn, int64_t ldb, int M, int N, int K,
            float alpha, const float* dscale, int clamp, float* C, int64_t ldc, int splits, cudaStream_t st);
int head_normalize_f16(const float* x, int64_t s_vec, int64_t s_elem, int nvec, int len, const float* norm, __half* out,
                       int ld_out, cudaStream_t st);
int head_scale_f16(const float* g, int64_t ld, int rows, int cols, float* scale, __half* out, int ld_out, cudaStream_t st);
int head_normalize_bwd_pair(const float*, const float*, int64_t, const float*, float*, int, const float*, const float*, int64_t,
                            int64_t, const float*, float*, int64_t, int64_t, int, int, cudaStream_t);
int head_prepare_operands(const float* x, int64_t x_sr, const float* w, int64_t w_sc, int64_t w_sk, int B, int C, int Din,
                          float* xnorm, float* wnorm, __half* x16, __half* w16, int Dp, cudaStream_t st);
int gemm_tc_arc_ce(const __half* x16, int64_t ldx, const __half* w16, int64_t ldw, int M, int N, int K, float s, float m,
                   int easy, const int64_t* labels, int class_off, int grad, float* part, float* rowmax, float* rowsum,
                   float* tgt, float* cos_t, const float* lse, const float* coef, const float* gout, float* scale,
                   __half* g16, int ld_g, cudaStream_t st, float* cos_out = nullptr, int ld_cos = 0);
int arc_ce_grad_from_cos(const float* cosm, int ld_cos, int M, int N, float s, float m, int easy, const int64_t* labels,
                         int class_off, const float* lse, const float* coef, const float* gout, float* scale, __half* g16,
                         int ld_g, cudaStream_t st);

namespace {
struct HeadWs {   // workspace layout of the head calls (fp32 part first, everything 256-byte aligned)
  size_t dxh, dwh, x16, w16, g16, scale, total;
  int Dp, Cp;
};
HeadWs head_ws(int B, int C, int Din, int precision) {
  HeadWs h{};
  h.Dp = (Din + 7) & ~7;
  h.Cp = (C + 7) & ~7;
  size_t o = 0;
  h.dxh = o; o += align_up(sizeof(float) * (size_t)B * Din, 256);
  h.dwh = o; o += align_up(sizeof(float) * (size_t)C * Din, 256);
  if (precision == TGFR_PREC_TC) {
    h.x16 = o; o += align_up(2 * (size_t)B * h.Dp, 256);
    h.w16 = o; o += align_up(2 * (size_t)C * h.Dp, 256);
    h.g16 = o; o += align_up(2 * (size_t)B * h.Cp, 256);
    h.scale = o; o += 256;
  }
  h.total = o;
  return h;
}
}  // namespace

int cos_logits_fwd(const float* x, int64_t x_sr, const float* w, int64_t w_sc, int64_t w_sk, int B, int C, int Din,
                   float s, int clamp_cos, float* out, int64_t out_sr, float* xnorm, float* wnorm, int precision,
                   void* ws, size_t ws_bytes, cudaStream_t st) {
  TGFR_REQUIRE(B > 0 && C > 0 && Din > 0, "cos_logits: empty shape");
  if (precision != TGFR_PREC_TC) {
    if (int rc = launch_norms(x, x_sr, 1, B, Din, xnorm, st)) return rc;
    if (int rc = launch_norms(w, w_sc, w_sk, C, Din, wnorm, st)) return rc;
  }
  if (precision == TGFR_PREC_TC) {
    TGFR_REQUIRE(head_tc_supported(B, C, Din), "cos_logits(tc): unsupported shape B=%d C=%d Din=%d", B, C, Din);
    const HeadWs h = head_ws(B, C, Din, precision);
    TGFR_REQUIRE(ws && ws_bytes >= h.total, "cos_logits(tc): workspace too small (%zu < %zu)", ws_bytes, h.total);
    TGFR_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "cos_logits(tc): workspace must be 256-byte aligned");
    uint8_t* base = reinterpret_cast<uint8_t*>(ws);
    __half* x16 = reinterpret_cast<__half*>(base + h.x16);
    __half* w16 = reinterpret_cast<__half*>(base + h.w16);
    if (int rc = head_prepare_operands(x, x_sr, w, w_sc, w_sk, B, C, Din, xnorm, wnorm, x16, w16, h.Dp, st)) return rc;
    return gemm_tc(x16, 0, h.Dp, w16, 0, h.Dp, B, C, h.Dp, s, nullptr, clamp_cos, out, out_sr, 1, st);
  }
  Gemm g{};
  g.A = x; g.sAm = x_sr; g.sAk = 1;
  g.B = w; g.sBk = w_sk; g.sBn = w_sc;
  g.C = out; g.sCm = out_sr; g.sCn = 1;
  g.M = B; g.N = C; g.K = Din; g.alpha = s; g.norm_eps = 1e-12f;
  g.am = xnorm; g.bn = wnorm; g.mode = clamp_cos ? kEpiClamp : kEpiScale;
  return launch_gemm(g, st);
}

int arc_margin_apply(float* logits, int64_t sr, const int64_t* labels, int B, int C, int class_off, float s, float m,
                     int easy, float* cos_t, cudaStream_t st) {
  const float pi = 3.14159265358979323846f;
  arc_apply_kernel<<<ceil_div(B, 128), 128, 0, st>>>(logits, sr, labels, B, C, class_off, s, cosf(m), sinf(m),
                                                     cosf(pi - m), sinf(pi - m) * m, easy, cos_t);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

size_t margin_workspace_bytes(int B, int C, int Din, int precision) { return head_ws(B, C, Din, precision).total; }

// shared by ArcFace (labels != NULL) and the plain cosine head (MagFace)
int margin_bwd(const float* x, int64_t x_sr, const float* w, int64_t w_sc, int64_t w_sk, const float* xnorm,
               const float* wnorm, const int64_t* labels, const float* cos_t, const float* g, int64_t g_sr, int B,
               int C, int Din, int class_off, float s, float m, int easy, float* dx, float* dw, int precision,
               void* ws, size_t ws_bytes, cudaStream_t st) {
  const HeadWs h = head_ws(B, C, Din, precision);
  TGFR_REQUIRE(ws && ws_bytes >= h.total, "margin_bwd: workspace too small (%zu < %zu)", ws_bytes, h.total);
  TGFR_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "margin_bwd: workspace must be 256-byte aligned");
  TGFR_REQUIRE(dw != nullptr, "margin_bwd: dw must not be NULL");
  uint8_t* base = reinterpret_cast<uint8_t*>(ws);
  float* dxh = reinterpret_cast<float*>(base + h.dxh);
  float* dwh = reinterpret_cast<float*>(base + h.dwh);
  if (precision == TGFR_PREC_TC) {
    TGFR_REQUIRE(head_tc_supported(B, C, Din), "margin_bwd(tc): unsupported shape B=%d C=%d Din=%d", B, C, Din);
    __half* x16 = reinterpret_cast<__half*>(base + h.x16);
    __half* w16 = reinterpret_cast<__half*>(base + h.w16);
    __half* g16 = reinterpret_cast<__half*>(base + h.g16);
    float* scale = reinterpret_cast<float*>(base + h.scale);
    if (int rc = head_normalize_f16(x, x_sr, 1, B, Din, xnorm, x16, h.Dp, st)) return rc;
    if (int rc = head_normalize_f16(w, w_sc, w_sk, C, Din, wnorm, w16, h.Dp, st)) return rc;
    if (int rc = head_scale_f16(g, g_sr, B, C, scale, g16, h.Cp, st)) return rc;     // g * 2^e as fp16, scale[1] = 2^-e
    if (dx) {   // dXhat[b,:] = s * sum_c g[b,c] w^_c : A = g16 K-major, B = w^ read MN-major, split over the classes
      const int tiles = ceil_div(B, 128) * ceil_div(Din, 128);
      const int splits = tiles >= 148 ? 1 : 296 / tiles;      // every (tile, split) CTA co-resident at 2 per SM: no tail wave
      if (int rc = gemm_tc(g16, 0, h.Cp, w16, 1, h.Dp, B, Din, C, s, scale, 0, dxh, Din, splits, st)) return rc;
    }
    // dWhat[c,:] = s * sum_b g[b,c] x^_b : A = g16 read MN-major, B = x^ read MN-major
    if (int rc = gemm_tc(g16, 1, h.Cp, x16, 1, h.Dp, C, Din, B, s, scale, 0, dwh, Din, 1, st)) return rc;
  } else if (dx) {
    Gemm a{};  // dXhat[b,:] = s * sum_c g[b,c] w^_c
    a.A = g; a.sAm = g_sr; a.sAk = 1;
    a.B = w; a.sBk = w_sc; a.sBn = w_sk;
    a.C = dxh; a.sCm = Din; a.sCn = 1;
    a.M = B; a.N = Din; a.K = C; a.alpha = s; a.norm_eps = 1e-12f; a.mode = kEpiScale; a.bk = wnorm;
    if (int rc = launch_gemm(a, st)) return rc;
  }
  if (precision != TGFR_PREC_TC) {
    Gemm b{};  // dWhat[c,:] = s * sum_b g[b,c] x^_b
    b.A = g; b.sAm = 1; b.sAk = g_sr;
    b.B = x; b.sBk = x_sr; b.sBn = 1;
    b.C = dwh; b.sCm = Din; b.sCn = 1;
    b.M = C; b.N = Din; b.K = B; b.alpha = s; b.norm_eps = 1e-12f; b.mode = kEpiScale; b.bk = xnorm;
    if (int rc = launch_gemm(b, st)) return rc;
  }
  if (labels) {
    const float pi = 3.14159265358979323846f;
    arc_fix_kernel<<<B, 128, 0, st>>>(x, x_sr, w, w_sc, w_sk, xnorm, wnorm, labels, cos_t, g, g_sr, C, Din, class_off,
                                      s, cosf(m), sinf(m), cosf(pi - m), sinf(pi - m) * m, easy, dx ? dxh : nullptr,
                                      dwh);
    TGFR_LAUNCH_OK();
  }
  const int fused = head_normalize_bwd_pair(dxh, x, x_sr, xnorm, dx, B, dwh, w, w_sc, w_sk, wnorm, dw, w_sc, w_sk, C, Din, st);
  if (fused <= 0) return fused;
  if (dx) {
    normalize_bwd_kernel<<<ceil_div(B, 8), 256, 0, st>>>(dxh, Din, x, x_sr, 1, xnorm, 1e-12f, B, Din, dx, Din, 1);
    TGFR_LAUNCH_OK();
  }
  normalize_bwd_kernel<<<ceil_div(C, 8), 256, 0, st>>>(dwh, Din, w, w_sc, w_sk, wnorm, 1e-12f, C, Din, dw, w_sc, w_sk);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

// --------------------------------------------------------------------------------------------
// fused ArcFace + cross entropy (tensor cores only): the [B, C] logits are never written.
//   forward : norms, fp16 normalised operands (kept in `saved` for the backward), cos-theta GEMM whose epilogue
//             applies the margin and reduces online-softmax partials; (rowmax, rowsum, tgt) describe THIS class
//             shard -- the caller merges shards (all-reduce) and runs focal_finish for the loss and lse.
//   backward: the same GEMM with the softmax-gradient epilogue -> g16, then the dX^ / dW^ GEMMs and the
//             normalisation backward exactly as margin_bwd.
// workspace: [part: 2 B ceil(C/128) fp32 | dxh | dwh | g16 | scale]; saved: [x16 | w16]
// --------------------------------------------------------------------------------------------
namespace {
struct FusedWs {
  size_t part, dxh, dwh, g16, scale, total, sv_x16, sv_w16, sv_cos, sv_total;
  int Dp, Cp, nt;
};
FusedWs fused_ws(int B, int C, int Din) {
  FusedWs f{};
  f.Dp = (Din + 7) & ~7;
  f.Cp = (C + 7) & ~7;
  f.nt = ceil_div(C, 128);
  size_t o = 0;
  f.part = o; o += align_up(sizeof(float) * 2 * (size_t)B * f.nt, 256);
  f.dxh = o; o += align_up(sizeof(float) * (size_t)B * Din, 256);
  f.dwh = o; o += align_up(sizeof(float) * (size_t)C * Din, 256);
  f.g16 = o; o += align_up(2 * (size_t)B * f.Cp, 256);
  f.scale = o; o += 256;
  f.total = o;
  f.sv_x16 = 0;
  f.sv_w16 = align_up(2 * (size_t)B * f.Dp, 256);
  f.sv_cos = f.sv_w16 + align_up(2 * (size_t)C * f.Dp, 256);                  // cos-theta [B, Cp] fp32 for the backward
  f.sv_total = f.sv_cos + align_up(sizeof(float) * (size_t)B * f.Cp, 256);
  return f;
}
}  // namespace

size_t arc_fused_workspace_bytes(int B, int C, int Din) { return fused_ws(B, C, Din).total; }
size_t arc_fused_saved_bytes(int B, int C, int Din) { return fused_ws(B, C, Din).sv_total; }

int arc_fused_fwd(const float* x, int64_t x_sr, const float* w, int64_t w_sc, int64_t w_sk, const int64_t* labels, int B,
                  int C, int Din, int class_off, float s, float m, int easy, float* xnorm, float* wnorm, float* rowmax,
                  float* rowsum, float* tgt, float* cos_t, void* ws, size_t ws_bytes, void* saved, size_t saved_bytes,
                  cudaStream_t st) {
  TGFR_REQUIRE(B > 0 && C > 0 && Din > 0, "arc_fused_fwd: empty shape");
  TGFR_REQUIRE(head_tc_supported(B, C, Din), "arc_fused_fwd: unsupported shape B=%d C=%d Din=%d", B, C, Din);
  const FusedWs f = fused_ws(B, C, Din);
  TGFR_REQUIRE(ws && ws_bytes >= f.total, "arc_fused_fwd: workspace too small (%zu < %zu)", ws_bytes, f.total);
  TGFR_REQUIRE(saved && saved_bytes >= f.sv_total, "arc_fused_fwd: saved buffer too small (%zu < %zu)", saved_bytes, f.sv_total);
  TGFR_REQUIRE(((reinterpret_cast<uintptr_t>(ws) | reinterpret_cast<uintptr_t>(saved)) & 255) == 0,
               "arc_fused_fwd: workspace / saved must be 256-byte aligned");
  uint8_t* base = reinterpret_cast<uint8_t*>(ws);
  uint8_t* sv = reinterpret_cast<uint8_t*>(saved);
  __half* x16 = reinterpret_cast<__half*>(sv + f.sv_x16);
  __half* w16 = reinterpret_cast<__half*>(sv + f.sv_w16);
  if (int rc = head_prepare_operands(x, x_sr, w, w_sc, w_sk, B, C, Din, xnorm, wnorm, x16, w16, f.Dp, st)) return rc;
  return gemm_tc_arc_ce(x16, f.Dp, w16, f.Dp, B, C, f.Dp, s, m, easy, labels, class_off, 0,
                        reinterpret_cast<float*>(base + f.part), rowmax, rowsum, tgt, cos_t, nullptr, nullptr, nullptr,
                        nullptr, nullptr, 0, st, reinterpret_cast<float*>(sv + f.sv_cos), f.Cp);
}

// per-device side stream + fork / join events of the head backward (TGFR_HEAD_OVERLAP=0 disables the overlap)
static int head_side_stream(cudaStream_t st, cudaStream_t* side, cudaEvent_t* join) {
  static cudaStream_t streams[64] = {};
  static cudaEvent_t forks[64] = {}, joins[64] = {};
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("TGFR_HEAD_OVERLAP");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  *side = st;
  *join = nullptr;
  if (!enabled) return TGFR_OK;
  int dev = 0;
  TGFR_CUDA_OK(cudaGetDevice(&dev));
  dev &= 63;
  if (!streams[dev]) {
    TGFR_CUDA_OK(cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking));
    TGFR_CUDA_OK(cudaEventCreateWithFlags(&forks[dev], cudaEventDisableTiming));
    TGFR_CUDA_OK(cudaEventCreateWithFlags(&joins[dev], cudaEventDisableTiming));
  }
  TGFR_CUDA_OK(cudaEventRecord(forks[dev], st));
  TGFR_CUDA_OK(cudaStreamWaitEvent(streams[dev], forks[dev], 0));
  *side = streams[dev];
  *join = joins[dev];
  return TGFR_OK;
}

int arc_fused_bwd(const float* x, int64_t x_sr, const float* w, int64_t w_sc, int64_t w_sk, const int64_t* labels,
                  const float* xnorm, const float* wnorm, const float* lse, const float* coef, const float* gout, int B,
                  int C, int Din, int class_off, float s, float m, int easy, float* dx, float* dw, void* ws,
                  size_t ws_bytes, const void* saved, size_t saved_bytes, cudaStream_t st) {
  const FusedWs f = fused_ws(B, C, Din);
  TGFR_REQUIRE(ws && ws_bytes >= f.total, "arc_fused_bwd: workspace too small (%zu < %zu)", ws_bytes, f.total);
  TGFR_REQUIRE(saved && saved_bytes >= f.sv_total, "arc_fused_bwd: saved buffer too small");
  TGFR_REQUIRE(dw != nullptr, "arc_fused_bwd: dw must not be NULL");
  uint8_t* base = reinterpret_cast<uint8_t*>(ws);
  const uint8_t* sv = reinterpret_cast<const uint8_t*>(saved);
  const __half* x16 = reinterpret_cast<const __half*>(sv + f.sv_x16);
  const __half* w16 = reinterpret_cast<const __half*>(sv + f.sv_w16);
  float* dxh = reinterpret_cast<float*>(base + f.dxh);
  float* dwh = reinterpret_cast<float*>(base + f.dwh);
  __half* g16 = reinterpret_cast<__half*>(base + f.g16);
  float* scale = reinterpret_cast<float*>(base + f.scale);
  // the softmax gradient (fp16 operand of the two gradient products) from the cos-theta matrix the forward kept: one
  // element-wise pass over 21 MB that sit in L2 instead of a second cos-theta GEMM (TGFR_HEAD_RECOMPUTE=1: the GEMM)
  static const bool recompute = getenv("TGFR_HEAD_RECOMPUTE") && atoi(getenv("TGFR_HEAD_RECOMPUTE")) == 1;
  if (recompute) {
    if (int rc = gemm_tc_arc_ce(x16, f.Dp, w16, f.Dp, B, C, f.Dp, s, m, easy, labels, class_off, 1, nullptr, nullptr,
                                nullptr, nullptr, nullptr, lse, coef, gout, scale, g16, f.Cp, st))
      return rc;
  } else {
    if (int rc = arc_ce_grad_from_cos(reinterpret_cast<const float*>(sv + f.sv_cos), f.Cp, B, C, s, m, easy, labels, class_off, lse,
                                      coef, gout, scale, g16, f.Cp, st))
      return rc;
  }
  // dX^ = g w^ and dW^ = g^T x^ are independent given g16 and each is latency bound at this size (tensor pipe < 20 %):
  // the first runs on a per-device side stream beside the second (fork / join by events; capturable into a graph)
  cudaStream_t side = st;
  cudaEvent_t ev_join = nullptr;
  if (dx) {
    if (int rc = head_side_stream(st, &side, &ev_join)) return rc;
    const int tiles = ceil_div(B, 128) * ceil_div(Din, 128);
    const int splits = tiles >= 148 ? 1 : 148 / tiles;      // half the slots: the dW^ product runs beside it
    if (int rc = gemm_tc(g16, 0, f.Cp, w16, 1, f.Dp, B, Din, C, s, scale, 0, dxh, Din, splits < 1 ? 1 : splits, side)) return rc;
    if (side != st) TGFR_CUDA_OK(cudaEventRecord(ev_join, side));
  }
  if (int rc = gemm_tc(g16, 1, f.Cp, x16, 1, f.Dp, C, Din, B, s, scale, 0, dwh, Din, 1, st)) return rc;
  if (dx && side != st) TGFR_CUDA_OK(cudaStreamWaitEvent(st, ev_join, 0));
  const int fused = head_normalize_bwd_pair(dxh, x, x_sr, xnorm, dx, B, dwh, w, w_sc, w_sk, wnorm, dw, w_sc, w_sk, C, Din, st);
  if (fused <= 0) return fused;
  if (dx) {
    normalize_bwd_kernel<<<ceil_div(B, 8), 256, 0, st>>>(dxh, Din, x, x_sr, 1, xnorm, 1e-12f, B, Din, dx, Din, 1);
    TGFR_LAUNCH_OK();
  }
  normalize_bwd_kernel<<<ceil_div(C, 8), 256, 0, st>>>(dwh, Din, w, w_sc, w_sk, wnorm, 1e-12f, C, Din, dw, w_sc, w_sk);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int mag_margin_fwd(const float* cos_s, const float* margin, int B, int C, float scale, int easy, float* out,
                   cudaStream_t st) {
  mag_fwd_kernel<<<dim3(ceil_div(C, 256), B), 256, 0, st>>>(cos_s, margin, B, C, scale, easy, out);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int mag_margin_bwd(const float* cos_s, const float* margin, const float* g_cos, const float* g_cosm, int B, int C,
                   float scale, int easy, float* gtotal, float* gmargin, cudaStream_t st) {
  mag_bwd_kernel<<<B, 256, 0, st>>>(cos_s, margin, g_cos, g_cosm, B, C, scale, easy, gtotal, gmargin);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int ce_rows_stats(const float* logits, int64_t sr, const int64_t* labels, int B, int C, int class_off, float* rowmax,
                  float* rowsum, float* tgt, cudaStream_t st) {
  TGFR_REQUIRE(B > 0 && C > 0, "ce_rows_stats: empty shape");
  ce_rows_stats_kernel<<<B, 256, 0, st>>>(logits, sr, labels, C, class_off, rowmax, rowsum, tgt);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int focal_finish(const float* rowmax, const float* rowsum, const float* tgt, int B, float gamma, float* out,
                 float* lse, cudaStream_t st) {
  focal_finish_kernel<<<1, 256, 0, st>>>(rowmax, rowsum, tgt, B, gamma, out, lse);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int ce_rows_bwd(const float* logits, int64_t sr, const int64_t* labels, const float* lse, const float* coef,
                const float* gout, int B, int C, int class_off, float* glogits, int64_t g_sr, cudaStream_t st) {
  ce_rows_bwd_kernel<<<dim3(ceil_div(C, 256), B), 256, 0, st>>>(logits, sr, labels, lse, coef, gout, B, C, class_off,
                                                                glogits, g_sr);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int mag_ce_stats(const float* cos_s, const float* cos_m, int64_t sr, const int64_t* labels, int B, int C, float* rowmax,
                 float* rowsum, float* tgt, float* one_hot, cudaStream_t st) {
  TGFR_REQUIRE(B > 0 && C > 0, "mag_ce_stats: empty shape");
  mag_ce_stats_kernel<<<B, 256, 0, st>>>(cos_s, cos_m, sr, labels, C, rowmax, rowsum, tgt, one_hot);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int mag_ce_bwd(const float* cos_s, const float* cos_m, int64_t sr, const int64_t* labels, const float* lse,
               const float* gout, int B, int C, float* g_cos, float* g_cosm, cudaStream_t st) {
  mag_ce_bwd_kernel<<<dim3(ceil_div(C, 256), B), 256, 0, st>>>(cos_s, cos_m, sr, labels, lse, gout, B, C, g_cos, g_cosm);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int cosine_rows_ref_fwd(const float* x1, int64_t s1r, int64_t s1d, const float* x2, int64_t s2r, int64_t s2d, int64_t n,
                        int D, float eps, float* out, float* stats, cudaStream_t st) {
  if (n == 0) return TGFR_OK;
  cosine_rows_ref_fwd_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(x1, s1r, s1d, x2, s2r, s2d, n, D, eps, out, stats);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int cosine_rows_ref_bwd(const float* x1, int64_t s1r, int64_t s1d, const float* x2, int64_t s2r, int64_t s2d, int64_t n,
                        int D, float eps, const float* stats, const float* g, float* d1, float* d2, cudaStream_t st) {
  if (n == 0) return TGFR_OK;
  cosine_rows_ref_bwd_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(x1, s1r, s1d, x2, s2r, s2d, n, D, eps, stats, g, d1,
                                                                     d2);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int merge_softmax_stats(const float* in, int n, int K, int M, float* out, cudaStream_t st) {
  TGFR_REQUIRE(n >= 1 && (K == 2 || K == 3) && M >= 1, "merge_softmax_stats: bad shape n=%d K=%d M=%d", n, K, M);
  merge_softmax_stats_kernel<<<ceil_div(M, 256), 256, 0, st>>>(in, n, K, M, out);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

}  // namespace tgfr

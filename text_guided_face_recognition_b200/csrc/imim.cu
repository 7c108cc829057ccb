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

namespace tgfr {
namespace {

constexpr int kC = 256, kC2 = 128;
constexpr float kLnEps = 1e-5f;

// ------------------------------------------------------------------------------------------------------------
// fp32 GEMM, row-major:  C[b] (+)= alpha * op(A[b]) * op(B[b]) (+ bias) (relu)
//   mode 0 (NT): A [M,K] lda, B [N,K] ldb          mode 1 (NN): A [M,K], B [K,N]          mode 2 (TN): A [K,M], B [K,N]
// blockIdx.z = batch * splits + split.
// ------------------------------------------------------------------------------------------------------------
struct SgemmP {
  const float *A, *B;
  float* C;
  const float* bias;
  int64_t lda, ldb, ldc, sa, sb, sc;   // leading dimensions and batch strides (elements)
  int M, N, K, splits, relu, atomic;
  float alpha;
};

// 128 x 128 tile, 8-deep K slices (double-buffered in shared memory, global loads of slice k+1 in flight while slice k
// is multiplied), 256 threads, 8 x 8 outputs per thread as 2 x 2 blocks of 4 x 4 (conflict-free float4 reads).
template <int MODE>
__global__ void __launch_bounds__(256) sgemm_kernel(const SgemmP p) {
  constexpr int BM = 128, BN = 128, BK = 8;
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int bz = blockIdx.z / p.splits, sp = blockIdx.z - bz * p.splits;
  const float* A = p.A + (int64_t)bz * p.sa;
  const float* Bm = p.B + (int64_t)bz * p.sb;
  float* C = p.C + (int64_t)bz * p.sc;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kper = ((p.K + p.splits - 1) / p.splits + BK - 1) / BK * BK;
  const int k_begin = sp * kper, k_end = min(p.K, k_begin + kper);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // staging: every thread moves 4 elements of A and 4 of B per slice
  //   k-contiguous operand ([rows, K]):  row = tid >> 1, k = (tid & 1) * 4 .. +3
  //   row-contiguous operand ([K, rows]): k = tid >> 5, rows (tid & 31) * 4 .. +3
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
    if (MODE == 2) {
      const int kk = tid >> 5, mm = (tid & 31) * 4, gk = k0 + kk;
#pragma unroll
      for (int j = 0; j < 4; ++j) ra[j] = (gk < k_end && m0 + mm + j < p.M) ? A[(int64_t)gk * p.lda + m0 + mm + j] : 0.f;
    } else {
      const int mm = tid >> 1, kk = (tid & 1) * 4, gm = m0 + mm;
#pragma unroll
      for (int j = 0; j < 4; ++j) ra[j] = (gm < p.M && k0 + kk + j < k_end) ? A[(int64_t)gm * p.lda + k0 + kk + j] : 0.f;
    }
    if (MODE == 0) {
      const int nn = tid >> 1, kk = (tid & 1) * 4, gn = n0 + nn;
#pragma unroll
      for (int j = 0; j < 4; ++j) rb[j] = (gn < p.N && k0 + kk + j < k_end) ? Bm[(int64_t)gn * p.ldb + k0 + kk + j] : 0.f;
    } else {
      const int kk = tid >> 5, nn = (tid & 31) * 4, gk = k0 + kk;
#pragma unroll
      for (int j = 0; j < 4; ++j) rb[j] = (gk < k_end && n0 + nn + j < p.N) ? Bm[(int64_t)gk * p.ldb + n0 + nn + j] : 0.f;
    }
  };
  auto stash = [&](int buf) {
    if (MODE == 2) {
      const int kk = tid >> 5, mm = (tid & 31) * 4;
      *reinterpret_cast<float4*>(&As[buf][kk][mm]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
    } else {
      const int mm = tid >> 1, kk = (tid & 1) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) As[buf][kk + j][mm] = ra[j];
    }
    if (MODE == 0) {
      const int nn = tid >> 1, kk = (tid & 1) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[buf][kk + j][nn] = rb[j];
    } else {
      const int kk = tid >> 5, nn = (tid & 31) * 4;
      *reinterpret_cast<float4*>(&Bs[buf][kk][nn]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    }
  };
  float acc[8][8] = {};
  int buf = 0;
  if (k_begin < k_end) {
    fetch(k_begin);
    stash(0);
  }
  __syncthreads();
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    const bool more = k0 + BK < k_end;
    if (more) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + (j >> 2) * 64 + tx * 4 + (j & 3);
      if (gn >= p.N) continue;
      float v = p.alpha * acc[i][j];
      if (p.bias && sp == 0) v += p.bias[gn];
      if (p.atomic) atomicAdd(&C[(int64_t)gm * p.ldc + gn], v);
      else C[(int64_t)gm * p.ldc + gn] = p.relu ? fmaxf(v, 0.f) : v;
    }
  }
}

int sgemm(int mode, const float* A, int64_t lda, int64_t sa, const float* B, int64_t ldb, int64_t sb, float* C, int64_t ldc,
          int64_t sc, int M, int N, int K, int batch, float alpha, const float* bias, int relu, int splits, cudaStream_t st) {
  SgemmP p{A, B, C, bias, lda, ldb, ldc, sa, sb, sc, M, N, K, splits < 1 ? 1 : splits, relu, splits > 1, alpha};
  if (p.atomic) {
    TGFR_REQUIRE(batch == 1 && ldc == N, "sgemm: split-K needs one dense output");
    TGFR_CUDA_OK(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st));
  }
  const dim3 grid(ceil_div(N, 128), ceil_div(M, 128), batch * p.splits);
  if (mode == 0) sgemm_kernel<0><<<grid, 256, 0, st>>>(p);
  else if (mode == 1) sgemm_kernel<1><<<grid, 256, 0, st>>>(p);
  else sgemm_kernel<2><<<grid, 256, 0, st>>>(p);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

// ------------------------------------------------------------------------------------------------------------
// BatchNorm2d over x [B, C, P] (element strides): per-channel statistics, apply + transpose to [B*P, C], backward
// ------------------------------------------------------------------------------------------------------------
__global__ void bn_stats_kernel(const float* __restrict__ x, int64_t sb, int64_t sc, int64_t sp, int B, int P, float eps,
                                float momentum, int training, float* __restrict__ run_mean, float* __restrict__ run_var,
                                float* __restrict__ mean, float* __restrict__ invstd) {
  __shared__ float scratch[32];
  const int c = blockIdx.x;
  if (!training) {
    if (threadIdx.x == 0) {
      mean[c] = run_mean[c];
      invstd[c] = rsqrtf(run_var[c] + eps);
    }
    return;
  }
  const int n = B * P;
  float s = 0.f;
  for (int k = threadIdx.x; k < n; k += blockDim.x) s += x[(k / P) * sb + c * sc + (k % P) * sp];
  const float mu = block_sum(s, scratch) / (float)n;
  float v = 0.f;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const float d = x[(k / P) * sb + c * sc + (k % P) * sp] - mu;
    v = fmaf(d, d, v);
  }
  const float var = block_sum(v, scratch) / (float)n;          // biased variance normalises (as torch does)
  if (threadIdx.x == 0) {
    mean[c] = mu;
    invstd[c] = rsqrtf(var + eps);
    if (run_mean) {                                            // running statistics: unbiased variance, momentum update
      run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mu;
      run_var[c] = (1.f - momentum) * run_var[c] + momentum * var * ((float)n / (float)max(n - 1, 1));
    }
  }
}

// xn[(b*P + p), c] = (x[b,c,p] - mean[c]) invstd[c] gamma[c] + beta[c]; a 32 x 32 tile transposed through shared memory
__global__ void bn_apply_t_kernel(const float* __restrict__ x, int64_t sb, int64_t sc, int64_t sp, int C, int P,
                                  const float* __restrict__ mean, const float* __restrict__ invstd,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ xn) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, pp = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && pp < P) ? x[b * sb + c * sc + pp * sp] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int pp = p0 + j, c = c0 + threadIdx.x;
    if (pp < P && c < C) xn[((int64_t)b * P + pp) * C + c] = (tile[threadIdx.x][j] - mean[c]) * invstd[c] * gamma[c] + beta[c];
  }
}

// per-channel sums for the BatchNorm backward: s1[c] = sum dxn, s2[c] = sum dxn * xhat  (xhat recomputed from x)
__global__ void bn_bwd_sums_kernel(const float* __restrict__ dxn, const float* __restrict__ x, int64_t sb, int64_t sc,
                                   int64_t sp, int B, int C, int P, const float* __restrict__ mean,
                                   const float* __restrict__ invstd, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float scratch[32];
  const int c = blockIdx.x, n = B * P;
  float s1 = 0.f, s2 = 0.f;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const float g = dxn[(int64_t)k * C + c];
    const float xh = (x[(k / P) * sb + c * sc + (k % P) * sp] - mean[c]) * invstd[c];
    s1 += g;
    s2 = fmaf(g, xh, s2);
  }
  s1 = block_sum(s1, scratch);
  s2 = block_sum(s2, scratch);
  if (threadIdx.x == 0) {
    dbeta[c] = s1;
    dgamma[c] = s2;
  }
}

// dx[b,c,p] (contiguous [B,C,P]) = gamma invstd (dxn - [training] (s1 + xhat s2) / n)
__global__ void bn_bwd_dx_kernel(const float* __restrict__ dxn, const float* __restrict__ x, int64_t sb, int64_t sc,
                                 int64_t sp, int B, int C, int P, const float* __restrict__ mean,
                                 const float* __restrict__ invstd, const float* __restrict__ gamma,
                                 const float* __restrict__ dgamma, const float* __restrict__ dbeta, int training,
                                 float* __restrict__ dx) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int pp = p0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (pp < P && c < C) ? dxn[((int64_t)b * P + pp) * C + c] : 0.f;
  }
  __syncthreads();
  const float inv_n = 1.f / (float)(B * P);
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, pp = p0 + threadIdx.x;
    if (c < C && pp < P) {
      float g = tile[threadIdx.x][j];
      if (training) {
        const float xh = (x[b * sb + c * sc + pp * sp] - mean[c]) * invstd[c];
        g -= (dbeta[c] + xh * dgamma[c]) * inv_n;
      }
      dx[((int64_t)b * C + c) * P + pp] = g * gamma[c] * invstd[c];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// softmax over the last axis of [rows, n] in place, and its backward dS = scale * P (dP - sum_j P dP)
// ------------------------------------------------------------------------------------------------------------
__global__ void softmax_rows_kernel(float* __restrict__ s, int rows, int n) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* p = s + (int64_t)row * n;
  float m = -INFINITY;
  for (int j = lane; j < n; j += 32) m = fmaxf(m, p[j]);
  m = warp_max(m);
  float sum = 0.f;
  for (int j = lane; j < n; j += 32) {
    const float e = expf(p[j] - m);
    p[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int j = lane; j < n; j += 32) p[j] *= inv;
}
__global__ void softmax_rows_bwd_kernel(const float* __restrict__ prob, float* __restrict__ dp, int rows, int n, float scale) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* p = prob + (int64_t)row * n;
  float* d = dp + (int64_t)row * n;
  float inner = 0.f;
  for (int j = lane; j < n; j += 32) inner = fmaf(p[j], d[j], inner);
  inner = warp_sum(inner);
  for (int j = lane; j < n; j += 32) d[j] = scale * p[j] * (d[j] - inner);
}

// ------------------------------------------------------------------------------------------------------------
// LayerNorm([C, 14, 14]) per sample on position-major data o [B, P, C]; affine weight / bias indexed [c * P + p]
// ------------------------------------------------------------------------------------------------------------
__global__ void ln_fwd_kernel(const float* __restrict__ o, int P, int C, const float* __restrict__ w,
                              const float* __restrict__ bia, float* __restrict__ y, float* __restrict__ mu_out,
                              float* __restrict__ rstd_out) {
  __shared__ float scratch[32];
  const int b = blockIdx.x, n = P * C;
  const float* src = o + (int64_t)b * n;
  float s = 0.f;
  for (int k = threadIdx.x; k < n; k += blockDim.x) s += src[k];
  const float mu = block_sum(s, scratch) / (float)n;
  float v = 0.f;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const float d = src[k] - mu;
    v = fmaf(d, d, v);
  }
  const float rstd = rsqrtf(block_sum(v, scratch) / (float)n + kLnEps);
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const int pp = k / C, c = k - pp * C;
    y[(int64_t)b * n + k] = (src[k] - mu) * rstd * w[c * P + pp] + bia[c * P + pp];
  }
  if (threadIdx.x == 0) {
    mu_out[b] = mu;
    rstd_out[b] = rstd;
  }
}
// dO = rstd (dh - mean(dh) - yhat mean(dh yhat)), dh = dY w;  dY is overwritten with dO
__global__ void ln_bwd_dx_kernel(float* __restrict__ dy, const float* __restrict__ o, int P, int C,
                                 const float* __restrict__ w, const float* __restrict__ mu_in,
                                 const float* __restrict__ rstd_in) {
  __shared__ float scratch[32];
  const int b = blockIdx.x, n = P * C;
  const float mu = mu_in[b], rstd = rstd_in[b];
  float* g = dy + (int64_t)b * n;
  const float* src = o + (int64_t)b * n;
  float s1 = 0.f, s2 = 0.f;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const int pp = k / C, c = k - pp * C;
    const float dh = g[k] * w[c * P + pp];
    s1 += dh;
    s2 = fmaf(dh, (src[k] - mu) * rstd, s2);
  }
  s1 = block_sum(s1, scratch) / (float)n;
  s2 = block_sum(s2, scratch) / (float)n;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const int pp = k / C, c = k - pp * C;
    const float dh = g[k] * w[c * P + pp];
    g[k] = rstd * (dh - s1 - (src[k] - mu) * rstd * s2);
  }
}
// d ln.weight[c*P+p] = sum_b dY yhat, d ln.bias = sum_b dY  (one thread per (p, c); must run BEFORE ln_bwd_dx_kernel)
__global__ void ln_bwd_params_kernel(const float* __restrict__ dy, const float* __restrict__ o, int B, int P, int C,
                                     const float* __restrict__ mu, const float* __restrict__ rstd, float* __restrict__ dw,
                                     float* __restrict__ db) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, n = P * C;
  if (k >= n) return;
  const int pp = k / C, c = k - pp * C;
  float a = 0.f, s = 0.f;
  for (int b = 0; b < B; ++b) {
    const float g = dy[(int64_t)b * n + k];
    a = fmaf(g, (o[(int64_t)b * n + k] - mu[b]) * rstd[b], a);
    s += g;
  }
  dw[c * P + pp] = a;
  db[c * P + pp] = s;
}

// ------------------------------------------------------------------------------------------------------------
// row-wise helpers on [M, C]: L2 normalisation (applied twice, models.py:119 + :403) and its backward; ReLU mask;
// column sums (bias gradients)
// ------------------------------------------------------------------------------------------------------------
__global__ void l2norm2_rows_kernel(const float* __restrict__ z, int M, int C, float* __restrict__ out,
                                    float* __restrict__ znorm) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* p = z + (int64_t)row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s = fmaf(p[c], p[c], s);
  const float n1 = fmaxf(sqrtf(warp_sum(s)), 1e-12f);
  float s2 = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = p[c] / n1;
    s2 = fmaf(v, v, s2);
  }
  const float n2 = fmaxf(sqrtf(warp_sum(s2)), 1e-12f);
  for (int c = lane; c < C; c += 32) out[(int64_t)row * C + c] = (p[c] / n1) / n2;
  if (lane == 0) znorm[row] = n1 * n2;
}
__global__ void l2norm_rows_kernel(const float* __restrict__ z, int M, int C, float* __restrict__ out,
                                   float* __restrict__ znorm) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* p = z + (int64_t)row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s = fmaf(p[c], p[c], s);
  const float n1 = fmaxf(sqrtf(warp_sum(s)), 1e-12f);
  for (int c = lane; c < C; c += 32) out[(int64_t)row * C + c] = p[c] / n1;
  if (lane == 0) znorm[row] = n1;
}
// dZ = (g - (g . o) o) / |Z|   (o = the unit output rows)
__global__ void l2norm_rows_bwd_kernel(const float* __restrict__ g, const float* __restrict__ o,
                                       const float* __restrict__ znorm, int M, int C, float* __restrict__ dz) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* gp = g + (int64_t)row * C;
  const float* op = o + (int64_t)row * C;
  float d = 0.f;
  for (int c = lane; c < C; c += 32) d = fmaf(gp[c], op[c], d);
  d = warp_sum(d);
  const float inv = 1.f / znorm[row];
  for (int c = lane; c < C; c += 32) dz[(int64_t)row * C + c] = (gp[c] - d * op[c]) * inv;
}
__global__ void relu_mask_kernel(float* __restrict__ g, const float* __restrict__ act, int64_t n) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    if (!(act[k] > 0.f)) g[k] = 0.f;
}
// out[c] = sum_m g[m, ld*.. + c]: blocks of 256 rows, atomics into a zeroed vector
__global__ void colsum_kernel(const float* __restrict__ g, int64_t ld, int M, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int m0 = blockIdx.y * 256, m1 = min(M, m0 + 256);
  float s = 0.f;
  for (int m = m0; m < m1; ++m) s += g[(int64_t)m * ld + c];
  atomicAdd(&out[c], s);
}

int colsum(const float* g, int64_t ld, int M, int C, float* out, cudaStream_t st) {
  TGFR_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * C, st));
  colsum_kernel<<<dim3(ceil_div(C, 128), ceil_div(M, 256)), 128, 0, st>>>(g, ld, M, C, out);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

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
  ln_fwd_kernel<<<B, 1024, 0, st>>>(S + L.o, P, kC, prm[8], prm[9], S + L.y, S + L.ln_mu, S + L.ln_rstd);
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
  ln_bwd_params_kernel<<<ceil_div(P * kC, 256), 256, 0, st>>>(dY, S + L.o, B, P, kC, S + L.ln_mu, S + L.ln_rstd, dprm[8], dprm[9]);
  TGFR_LAUNCH_OK();
  ln_bwd_dx_kernel<<<B, 1024, 0, st>>>(dY, S + L.o, P, kC, prm[8], S + L.ln_mu, S + L.ln_rstd);     // dY -> dO
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
                                                                               prm[0], dprm[0], dprm[1], training, dx);
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

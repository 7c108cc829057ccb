// Word-region attention loss, fp32 SIMT path (TGFR_PREC_FP32).
//
// One CTA per (face b, caption i) pair; the T x R attention tile lives entirely in registers and
// shared memory, so the B x B x T x R tensor of models/losses.py:96 never reaches HBM.
//
//   phase A  S[r,t] = <c_r, q_t>        thread-owns-region (r = threadIdx.x), region rows staged
//            A1 = softmax_t(S)           through shared memory in 32-feature chunks (coalesced),
//            E  = exp(g1*A1)             softmax over words is thread-local (attention.py:27-36)
//   phase B  Wu[t,d] = sum_r E[r,t] c_r[d], Z[t] = sum_r E[r,t]   warp-owns-words (attention.py:41)
//   phase C  cos_t, exp(g2 cos_t), log-sum  (losses.py:104-109)
// Backward recomputes A and B, then runs the same two contraction shapes for the gradients.
//
// Limits of this path: T <= 32, R <= 256, D <= 256 and D % 4 == 0 (the reference uses T <= 30,
// R = 196, D = 256).
#include "common.cuh"

namespace tgfr {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kChunk = 32;    // features per staged chunk of the region tile
constexpr int kCStride = 36;  // padded row stride (floats) of the staged chunk: conflict-free LDS.128
constexpr int kMaxR = 256;
constexpr int kMaxD = 256;

enum Mode { kLoss = 0, kAttention = 1 };

struct Params {
  const float* ctx;
  int64_t csb, csr, csd;
  const float* words;
  int64_t wsb, wst, wsd;
  const int32_t* cap_lens;
  int Bc, Bq, T, R, D;
  float g1, g2, g3, eps;
  // forward outputs
  float* sim;    // [Bc,Bq]            (kLoss)
  float* attn;   // [Bc,T,R]           (kLoss: diagonal pairs; kAttention: all)
  float* wc;     // [B,T,D]            (kAttention)
  int diag_off;
  // backward
  const float* gsim;   // [Bc,Bq]      (kLoss)
  const float* g_wc;   // [B,T,D]      (kAttention, may be null)
  const float* g_attn; // [B,T,R]      (kAttention, may be null)
  float* dctx;         // [Bc,R,D] accumulated with atomics (pre-zeroed), may be null
  float* dwords;       // [Bq,T,D] accumulated with atomics (pre-zeroed), may be null
};

struct Smem {
  float* q;    // [TP][qs]   words of the caption (zero rows beyond its length)
  float* v;    // [TP][qs]   backward: d loss / d w_t (or upstream g_wc)
  float* e;    // [R][es]    E = exp(g1*A1); later dS
  float* cs;   // [kMaxR][kCStride] staging of a 32-feature chunk of the region tile
  float* red;  // [64] small reductions: red[0..TP) per-word values, red[32..64) scratch
  float* z;    // [32] Z[t]
  float* rs;   // [32] rowsum_t
};

template <int TP>
__host__ __device__ inline size_t smem_floats(int R, int D, bool bwd) {
  const int qs = D + 4, es = TP + 1;
  size_t n = (size_t)TP * qs + (size_t)R * es + (size_t)kMaxR * kCStride + 64 + 32 + 32;
  if (bwd) n += (size_t)TP * qs;
  return n;
}

template <int TP>
__device__ inline Smem carve(float* base, int R, int D, bool bwd) {
  const int qs = D + 4, es = TP + 1;
  Smem s;
  s.q = base;
  base += TP * qs;
  s.v = base;
  if (bwd) base += TP * qs;
  s.cs = base;  // keep 16-byte alignment: TP*qs is a multiple of 4 floats
  base += kMaxR * kCStride;
  s.e = base;
  base += R * es;
  s.red = base;
  base += 64;
  s.z = base;
  base += 32;
  s.rs = base;
  return s;
}

// rows [0,n) of a [*, D] matrix with element strides (st, sd) -> dst[TP][qs]; other rows zeroed.
template <int TP>
__device__ inline void load_rows(float* dst, const float* src, int64_t st, int64_t sd, int n, int D) {
  const int qs = D + 4;
  for (int idx = threadIdx.x; idx < TP * D; idx += kThreads) {
    const int t = idx / D, d = idx - t * D;
    dst[t * qs + d] = (t < n) ? __ldg(src + t * st + d * sd) : 0.f;
  }
}

// acc[t] = sum_d C[r][d] * V[t][d] for this thread's region r (all threads take part in staging).
template <int TP>
__device__ inline void region_gemm(float (&acc)[TP], const float* __restrict__ vs, const float* __restrict__ cb,
                                   int64_t csr, int64_t csd, float* cs, int R, int D) {
  const int qs = D + 4;
  const int r = threadIdx.x;
#pragma unroll
  for (int t = 0; t < TP; ++t) acc[t] = 0.f;
  for (int d0 = 0; d0 < D; d0 += kChunk) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < R * kChunk; idx += kThreads) {
      const int rr = idx >> 5, dd = idx & 31;
      cs[rr * kCStride + dd] = (d0 + dd < D) ? __ldg(cb + rr * csr + (int64_t)(d0 + dd) * csd) : 0.f;
    }
    __syncthreads();
    if (r < R) {
      const int nq = min(kChunk, D - d0) >> 2;
      for (int dq = 0; dq < nq; ++dq) {
        const float4 c4 = *reinterpret_cast<const float4*>(cs + r * kCStride + 4 * dq);
#pragma unroll
        for (int t = 0; t < TP; ++t) {
          const float4 v4 = *reinterpret_cast<const float4*>(vs + t * qs + d0 + 4 * dq);
          acc[t] = fmaf(c4.x, v4.x, acc[t]);
          acc[t] = fmaf(c4.y, v4.y, acc[t]);
          acc[t] = fmaf(c4.z, v4.z, acc[t]);
          acc[t] = fmaf(c4.w, v4.w, acc[t]);
        }
      }
    }
  }
}

// acc[j][k] = sum_r X[r][t] * C[r][d]  with t = warp + 8 j, d = lane + 32 k; zsum[j] = sum_r X[r][t].
template <int TP>
__device__ inline void word_gemm(float (&acc)[TP / 8][8], float (&zsum)[TP / 8], const float* __restrict__ xs,
                                 const float* __restrict__ cb, int64_t csr, int64_t csd, int R, int D) {
  constexpr int NJ = TP / 8;
  const int es = TP + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    zsum[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[j][k] = 0.f;
  }
#pragma unroll 2
  for (int r = 0; r < R; ++r) {
    float c[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int d = lane + 32 * k;
      c[k] = (d < D) ? __ldg(cb + r * csr + (int64_t)d * csd) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float x = xs[r * es + warp + 8 * j];
      zsum[j] += x;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[j][k] = fmaf(x, c[k], acc[j][k]);
    }
  }
}

// Phase A: scores, word softmax, E -> shared memory (rows of padded words give E = 0).
template <int TP>
__device__ inline void phase_a(const Params& p, const Smem& s, const float* cb, int Ti) {
  float acc[TP];
  region_gemm<TP>(acc, s.q, cb, p.csr, p.csd, s.cs, p.R, p.D);
  const int r = threadIdx.x;
  if (r < p.R) {
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < TP; ++t)
      if (t < Ti) m = fmaxf(m, acc[t]);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < TP; ++t) {
      acc[t] = (t < Ti) ? expf(acc[t] - m) : 0.f;
      sum += acc[t];
    }
    const float inv = 1.f / sum;
    const int es = TP + 1;
#pragma unroll
    for (int t = 0; t < TP; ++t) s.e[r * es + t] = (t < Ti) ? expf(p.g1 * (acc[t] * inv)) : 0.f;
  }
  __syncthreads();
}

template <int TP, int MODE>
__global__ void __launch_bounds__(kThreads, 2) wr_fwd_kernel(const Params p) {
  extern __shared__ __align__(16) float smem_raw[];
  const Smem s = carve<TP>(smem_raw, p.R, p.D, false);
  constexpr int NJ = TP / 8;
  const int i = blockIdx.x, b = (MODE == kLoss) ? blockIdx.y : blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qs = p.D + 4;
  int Ti = p.T;
  if (p.cap_lens) Ti = min(max(p.cap_lens[i], 1), p.T);
  const float* cb = p.ctx + (int64_t)b * p.csb;

  load_rows<TP>(s.q, p.words + (int64_t)i * p.wsb, p.wst, p.wsd, Ti, p.D);
  phase_a<TP>(p, s, cb, Ti);

  float wu[NJ][8], zs[NJ];
  word_gemm<TP>(wu, zs, s.e, cb, p.csr, p.csd, p.R, p.D);

#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int t = warp + 8 * j;
    if (t >= Ti) {
      if (lane == 0 && t < 32) { s.red[t] = 0.f; s.z[t] = 1.f; }
      continue;
    }
    const float invz = 1.f / zs[j];
    if (MODE == kAttention) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int d = lane + 32 * k;
        if (d < p.D && p.wc) p.wc[((int64_t)b * p.T + t) * p.D + d] = wu[j][k] * invz;
      }
      if (lane == 0) s.z[t] = zs[j];
    } else {
      float dot = 0.f, w2 = 0.f, q2 = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int d = lane + 32 * k;
        const float q = (d < p.D) ? s.q[t * qs + d] : 0.f;
        const float w = wu[j][k] * invz;
        dot = fmaf(q, w, dot);
        w2 = fmaf(w, w, w2);
        q2 = fmaf(q, q, q2);
      }
      dot = warp_sum(dot);
      w2 = warp_sum(w2);
      q2 = warp_sum(q2);
      const float den = fmaxf(sqrtf(q2) * sqrtf(w2), p.eps);  // losses.py:12-16
      if (lane == 0) {
        s.red[t] = expf(p.g2 * (dot / den));                  // losses.py:107
        s.z[t] = zs[j];
      }
    }
  }
  __syncthreads();
  if (MODE == kLoss && threadIdx.x == 0) {
    float sum = 0.f;
    for (int t = 0; t < Ti; ++t) sum += s.red[t];
    p.sim[(int64_t)b * p.Bq + i] = p.g3 * logf(sum);          // losses.py:108-109, 122
  }
  const bool want_attn = p.attn != nullptr && (MODE == kAttention || i == b + p.diag_off);
  if (want_attn) {
    const int es = TP + 1;
    float* out = p.attn + (int64_t)b * p.T * p.R;
    for (int idx = threadIdx.x; idx < p.T * p.R; idx += kThreads) {
      const int t = idx / p.R, r = idx - t * p.R;
      out[idx] = (t < Ti) ? s.e[r * es + t] / s.z[t] : 0.f;
    }
  }
}

template <int TP, int MODE>
__global__ void __launch_bounds__(kThreads, 1) wr_bwd_kernel(const Params p) {
  extern __shared__ __align__(16) float smem_raw[];
  const Smem s = carve<TP>(smem_raw, p.R, p.D, true);
  constexpr int NJ = TP / 8;
  const int i = blockIdx.x, b = (MODE == kLoss) ? blockIdx.y : blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qs = p.D + 4, es = TP + 1;
  int Ti = p.T;
  if (MODE == kLoss && p.cap_lens) Ti = min(max(p.cap_lens[i], 1), p.T);
  const float* cb = p.ctx + (int64_t)b * p.csb;
  const bool need_dq = p.dwords != nullptr, need_dc = p.dctx != nullptr;

  load_rows<TP>(s.q, p.words + (int64_t)i * p.wsb, p.wst, p.wsd, Ti, p.D);
  if (MODE == kAttention) {
    if (p.g_wc) load_rows<TP>(s.v, p.g_wc + (int64_t)b * p.T * p.D, p.D, 1, Ti, p.D);
    else load_rows<TP>(s.v, p.words, 0, 0, 0, p.D);  // zeros
  }
  phase_a<TP>(p, s, cb, Ti);

  {
    float wu[NJ][8], zs[NJ];
    word_gemm<TP>(wu, zs, s.e, cb, p.csr, p.csd, p.R, p.D);
    if (MODE == kAttention) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int t = warp + 8 * j;
        if (lane == 0 && t < 32) s.z[t] = (t < Ti) ? zs[j] : 1.f;
      }
      __syncthreads();
    } else {
      // cos_t and exp(g2 cos_t); keep per-word scalars in registers for the second half
      float cosv[NJ], denv[NJ], q2v[NJ], w2v[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int t = warp + 8 * j;
        cosv[j] = 0.f; denv[j] = 1.f; q2v[j] = 1.f; w2v[j] = 1.f;
        if (t >= Ti) {
          if (lane == 0 && t < 32) { s.red[t] = 0.f; s.z[t] = 1.f; }
          continue;
        }
        const float invz = 1.f / zs[j];
        float dot = 0.f, w2 = 0.f, q2 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int d = lane + 32 * k;
          const float q = (d < p.D) ? s.q[t * qs + d] : 0.f;
          const float w = wu[j][k] * invz;
          dot = fmaf(q, w, dot);
          w2 = fmaf(w, w, w2);
          q2 = fmaf(q, q, q2);
        }
        dot = warp_sum(dot);
        w2 = warp_sum(w2);
        q2 = warp_sum(q2);
        const float prod = sqrtf(q2) * sqrtf(w2);
        denv[j] = fmaxf(prod, p.eps);
        cosv[j] = dot / denv[j];
        q2v[j] = (prod > p.eps) ? q2 : INFINITY;   // clamp active -> no gradient through the norms
        w2v[j] = (prod > p.eps) ? w2 : INFINITY;
        if (lane == 0) {
          s.red[t] = expf(p.g2 * cosv[j]);
          s.z[t] = zs[j];
        }
      }
      __syncthreads();
      float total = 0.f;
      for (int t = 0; t < Ti; ++t) total += s.red[t];
      const float gscale = p.gsim[(int64_t)b * p.Bq + i] * p.g3 * p.g2 / total;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int t = warp + 8 * j;
        if (t >= TP) continue;
        if (t >= Ti) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int d = lane + 32 * k;
            if (d < p.D) s.v[t * qs + d] = 0.f;
          }
          continue;
        }
        const float dcos = gscale * s.red[t];
        const float invz = 1.f / zs[j];
        const float a = dcos / denv[j], bw = dcos * cosv[j] / w2v[j], bq = dcos * cosv[j] / q2v[j];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int d = lane + 32 * k;
          if (d >= p.D) continue;
          const float q = s.q[t * qs + d];
          const float w = wu[j][k] * invz;
          s.v[t * qs + d] = a * q - bw * w;                     // d loss / d w_t
          if (need_dq) atomicAdd(p.dwords + ((int64_t)i * p.T + t) * p.D + d, a * w - bq * q);
        }
      }
    }
  }
  if (threadIdx.x < 32) s.rs[threadIdx.x] = 0.f;
  // (region_gemm starts with a __syncthreads, which also publishes s.v / s.z / s.rs)

  float da[TP];
  region_gemm<TP>(da, s.v, cb, p.csr, p.csd, s.cs, p.R, p.D);   // dA2[t] = <dw_t, c_r>
  const int r = threadIdx.x;
  float a2[TP];
#pragma unroll
  for (int t = 0; t < TP; ++t) {
    a2[t] = (r < p.R && t < Ti) ? s.e[r * es + t] / s.z[t] : 0.f;
    if (MODE == kAttention && p.g_attn && r < p.R && t < Ti)
      da[t] += p.g_attn[((int64_t)b * p.T + t) * p.R + r];
  }
#pragma unroll
  for (int t = 0; t < TP; ++t) {
    const float part = warp_sum(a2[t] * da[t]);
    if (lane == 0 && t < Ti) atomicAdd(&s.rs[t], part);
  }
  __syncthreads();
  {
    // dZ2 = A2 (dA2 - rowsum), dA1 = g1 dZ2, dS = A1 (dA1 - sum_t A1 dA1)   (thread-local in t)
    float inner = 0.f;
    const float inv_g1 = (p.g1 != 0.f) ? 1.f / p.g1 : 0.f;
#pragma unroll
    for (int t = 0; t < TP; ++t) {
      const bool live = r < p.R && t < Ti;
      const float a1 = live ? logf(s.e[r * es + t]) * inv_g1 : 0.f;
      const float da1 = live ? p.g1 * a2[t] * (da[t] - s.rs[t]) : 0.f;
      inner = fmaf(a1, da1, inner);
      da[t] = da1;
    }
#pragma unroll
    for (int t = 0; t < TP; ++t) {
      const bool live = r < p.R && t < Ti;
      const float a1 = live ? logf(s.e[r * es + t]) * inv_g1 : 0.f;
      da[t] = a1 * (da[t] - inner);                             // dS[r][t]
    }
  }
  // dC[r][:] += sum_t A2[r,t] dw_t + dS[r,t] q_t, 32 features at a time through the staging tile
  if (need_dc) {
    float* dcb = p.dctx + (int64_t)b * p.R * p.D;
    for (int d0 = 0; d0 < p.D; d0 += kChunk) {
      const int nq = min(kChunk, p.D - d0) >> 2;
      __syncthreads();
      if (r < p.R) {
        for (int dq = 0; dq < nq; ++dq) {
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int t = 0; t < TP; ++t) {
            const float4 v4 = *reinterpret_cast<const float4*>(s.v + t * qs + d0 + 4 * dq);
            const float4 q4 = *reinterpret_cast<const float4*>(s.q + t * qs + d0 + 4 * dq);
            o.x = fmaf(a2[t], v4.x, fmaf(da[t], q4.x, o.x));
            o.y = fmaf(a2[t], v4.y, fmaf(da[t], q4.y, o.y));
            o.z = fmaf(a2[t], v4.z, fmaf(da[t], q4.z, o.z));
            o.w = fmaf(a2[t], v4.w, fmaf(da[t], q4.w, o.w));
          }
          *reinterpret_cast<float4*>(s.cs + r * kCStride + 4 * dq) = o;
        }
      }
      __syncthreads();
      for (int idx = threadIdx.x; idx < p.R * nq; idx += kThreads) {
        const int rr = idx / nq, dq = idx - rr * nq;
        const float4 o = *reinterpret_cast<const float4*>(s.cs + rr * kCStride + 4 * dq);
        red_add_v4(dcb + (int64_t)rr * p.D + d0 + 4 * dq, o.x, o.y, o.z, o.w);
      }
    }
  }
  // dQ[t][:] += sum_r dS[r,t] c_r
  if (need_dq) {
    __syncthreads();
    if (r < p.R) {
#pragma unroll
      for (int t = 0; t < TP; ++t) s.e[r * es + t] = da[t];
    }
    __syncthreads();
    float dq[NJ][8], unused[NJ];
    word_gemm<TP>(dq, unused, s.e, cb, p.csr, p.csd, p.R, p.D);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int t = warp + 8 * j;
      if (t >= Ti) continue;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int d = lane + 32 * k;
        if (d < p.D) atomicAdd(p.dwords + ((int64_t)i * p.T + t) * p.D + d, dq[j][k]);
      }
    }
  }
}


// Attention maps only (the B diagonal maps of the tensor-core path, or a func_attention caller that drops the
// context): phase A of the loss kernel with the region tile staged by cp.async, three buffers deep, so that the
// eight 32-feature rounds overlap their loads with the FMAs instead of serialising load -> sync -> compute
// (a 128-CTA launch has nothing else to hide that latency with).  The dot products run in the same order as
// region_gemm, so the maps are bit-identical to the fp32 loss path's.
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int TP>
__global__ void __launch_bounds__(kThreads, 1) attn_only_kernel(const Params p) {
  extern __shared__ __align__(16) float smem_raw[];
  const int qs = p.D + 4, es = TP + 1;
  float* sq = smem_raw;                              // [TP][qs] words of the caption (zero rows beyond its length)
  float* cs0 = sq + TP * qs;                         // [3][kMaxR][kCStride] staged 32-feature chunks of the region tile
  float* se = cs0 + 3 * kMaxR * kCStride;            // [R][es]  E = exp(g1 A1)
  float* sz = se + p.R * es;                         // [32]     Z[t] = sum_r E[r,t]
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int Ti = p.T;
  if (p.cap_lens) Ti = min(max(p.cap_lens[b], 1), p.T);
  const float* cb = p.ctx + (int64_t)b * p.csb;
  // 16-byte copies need contiguous, aligned feature rows; other layouts stage with plain loads
  const bool vec = p.csd == 1 && (p.csr & 3) == 0 && (p.D & 31) == 0 && ((reinterpret_cast<uintptr_t>(cb) & 15) == 0);
  const int nchunk = (p.D + kChunk - 1) / kChunk;

  auto stage = [&](int k) {
    float* cs = cs0 + (k % 3) * (kMaxR * kCStride);
    const int d0 = k * kChunk;
    if (vec) {
      for (int idx = threadIdx.x; idx < p.R * 8; idx += kThreads) {
        const int rr = idx >> 3, q4 = idx & 7;
        cp_async16(cs + rr * kCStride + 4 * q4, cb + (int64_t)rr * p.csr + d0 + 4 * q4);
      }
    } else {
      for (int idx = threadIdx.x; idx < p.R * kChunk; idx += kThreads) {
        const int rr = idx >> 5, dd = idx & 31;
        cs[rr * kCStride + dd] = (d0 + dd < p.D) ? __ldg(cb + rr * p.csr + (int64_t)(d0 + dd) * p.csd) : 0.f;
      }
    }
    cp_async_commit();
  };

  stage(0);
  if (nchunk > 1) stage(1);
  load_rows<TP>(sq, p.words + (int64_t)b * p.wsb, p.wst, p.wsd, Ti, p.D);
  const int r = threadIdx.x;
  float acc[TP];
#pragma unroll
  for (int t = 0; t < TP; ++t) acc[t] = 0.f;
  for (int k = 0; k < nchunk; ++k) {
    if (k + 2 < nchunk) {
      stage(k + 2);                 // two rounds ahead: that buffer's last readers passed the barrier closing round k-1
      cp_async_wait<2>();
    } else if (k + 1 < nchunk) {
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (r < p.R) {
      const float* cs = cs0 + (k % 3) * (kMaxR * kCStride);
      const int d0 = k * kChunk;
      const int nq = min(kChunk, p.D - d0) >> 2;
      for (int dq = 0; dq < nq; ++dq) {
        const float4 c4 = *reinterpret_cast<const float4*>(cs + r * kCStride + 4 * dq);
#pragma unroll
        for (int t = 0; t < TP; ++t) {
          const float4 v4 = *reinterpret_cast<const float4*>(sq + t * qs + d0 + 4 * dq);
          acc[t] = fmaf(c4.x, v4.x, acc[t]);
          acc[t] = fmaf(c4.y, v4.y, acc[t]);
          acc[t] = fmaf(c4.z, v4.z, acc[t]);
          acc[t] = fmaf(c4.w, v4.w, acc[t]);
        }
      }
    }
    __syncthreads();
  }
  if (r < p.R) {                    // word softmax and E, as phase_a
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < TP; ++t)
      if (t < Ti) m = fmaxf(m, acc[t]);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < TP; ++t) {
      acc[t] = (t < Ti) ? expf(acc[t] - m) : 0.f;
      sum += acc[t];
    }
    const float inv = 1.f / sum;
#pragma unroll
    for (int t = 0; t < TP; ++t) se[r * es + t] = (t < Ti) ? expf(p.g1 * (acc[t] * inv)) : 0.f;
  }
  __syncthreads();
  for (int t = warp; t < 32; t += kWarps) {
    float z = 0.f;
    if (t < Ti)
      for (int rr = lane; rr < p.R; rr += 32) z += se[rr * es + t];
    z = warp_sum(z);
    if (lane == 0) sz[t] = (t < Ti) ? z : 1.f;
  }
  __syncthreads();
  float* out = p.attn + (int64_t)b * p.T * p.R;
  for (int idx = threadIdx.x; idx < p.T * p.R; idx += kThreads) {
    const int t = idx / p.R, rr = idx - t * p.R;
    out[idx] = (t < Ti) ? se[rr * es + t] / sz[t] : 0.f;
  }
}

template <int TP>
int launch_attn_only(const Params& p, cudaStream_t st) {
  const size_t bytes = ((size_t)TP * (p.D + 4) + 3 * (size_t)kMaxR * kCStride + (size_t)p.R * (TP + 1) + 32) * sizeof(float);
  auto k = attn_only_kernel<TP>;
  {
    static bool attr_done[64] = {};   // once per instantiation and device
    int dev = 0;
    TGFR_CUDA_OK(cudaGetDevice(&dev));
    if (!attr_done[dev & 63]) {
      TGFR_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
      attr_done[dev & 63] = true;
    }
  }
  k<<<p.Bc, kThreads, bytes, st>>>(p);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

template <int TP, int MODE>
int launch_fwd(const Params& p, cudaStream_t st) {
  const size_t bytes = smem_floats<TP>(p.R, p.D, false) * sizeof(float);
  auto k = wr_fwd_kernel<TP, MODE>;
  {
    static bool attr_done[64] = {};   // once per instantiation and device
    int dev = 0;
    TGFR_CUDA_OK(cudaGetDevice(&dev));
    if (!attr_done[dev & 63]) {
      TGFR_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
      attr_done[dev & 63] = true;
    }
  }
  const dim3 grid(MODE == kLoss ? p.Bq : p.Bc, MODE == kLoss ? p.Bc : 1);
  k<<<grid, kThreads, bytes, st>>>(p);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

template <int TP, int MODE>
int launch_bwd(const Params& p, cudaStream_t st) {
  const size_t bytes = smem_floats<TP>(p.R, p.D, true) * sizeof(float);
  auto k = wr_bwd_kernel<TP, MODE>;
  {
    static bool attr_done[64] = {};   // once per instantiation and device
    int dev = 0;
    TGFR_CUDA_OK(cudaGetDevice(&dev));
    if (!attr_done[dev & 63]) {
      TGFR_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
      attr_done[dev & 63] = true;
    }
  }
  const dim3 grid(MODE == kLoss ? p.Bq : p.Bc, MODE == kLoss ? p.Bc : 1);
  k<<<grid, kThreads, bytes, st>>>(p);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int check_shape(const Params& p) {
  TGFR_REQUIRE(p.Bc > 0 && p.Bq > 0 && p.T > 0 && p.R > 0 && p.D > 0, "wordregion: empty shape");
  TGFR_REQUIRE(p.T <= 32, "wordregion(fp32): T=%d > 32 words is not supported", p.T);
  TGFR_REQUIRE(p.R <= kMaxR, "wordregion(fp32): R=%d > %d regions is not supported", p.R, kMaxR);
  TGFR_REQUIRE(p.D <= kMaxD && p.D % 4 == 0, "wordregion(fp32): D=%d must be <= %d and a multiple of 4", p.D, kMaxD);
  return TGFR_OK;
}

template <int MODE>
int dispatch_fwd(const Params& p, cudaStream_t st) {
  if (p.T <= 8) return launch_fwd<8, MODE>(p, st);
  if (p.T <= 16) return launch_fwd<16, MODE>(p, st);
  if (p.T <= 24) return launch_fwd<24, MODE>(p, st);
  return launch_fwd<32, MODE>(p, st);
}
template <int MODE>
int dispatch_bwd(const Params& p, cudaStream_t st) {
  if (p.T <= 8) return launch_bwd<8, MODE>(p, st);
  if (p.T <= 16) return launch_bwd<16, MODE>(p, st);
  if (p.T <= 24) return launch_bwd<24, MODE>(p, st);
  return launch_bwd<32, MODE>(p, st);
}

}  // namespace

int wordregion_fwd_simt(const float* ctx, int64_t csb, int64_t csr, int64_t csd, const float* words, int64_t wsb,
                        int64_t wst, int64_t wsd, const int32_t* cap_lens, int Bc, int Bq, int T, int R, int D,
                        float g1, float g2, float g3, float eps, float* sim, float* attn, int diag_off,
                        cudaStream_t st) {
  Params p{};
  p.ctx = ctx; p.csb = csb; p.csr = csr; p.csd = csd;
  p.words = words; p.wsb = wsb; p.wst = wst; p.wsd = wsd;
  p.cap_lens = cap_lens; p.Bc = Bc; p.Bq = Bq; p.T = T; p.R = R; p.D = D;
  p.g1 = g1; p.g2 = g2; p.g3 = g3; p.eps = eps;
  p.sim = sim; p.attn = attn; p.diag_off = diag_off;
  if (int rc = check_shape(p)) return rc;
  return dispatch_fwd<kLoss>(p, st);
}

int wordregion_bwd_simt(const float* ctx, int64_t csb, int64_t csr, int64_t csd, const float* words, int64_t wsb,
                        int64_t wst, int64_t wsd, const int32_t* cap_lens, int Bc, int Bq, int T, int R, int D,
                        float g1, float g2, float g3, float eps, const float* gsim, float* dctx, float* dwords,
                        cudaStream_t st) {
  Params p{};
  p.ctx = ctx; p.csb = csb; p.csr = csr; p.csd = csd;
  p.words = words; p.wsb = wsb; p.wst = wst; p.wsd = wsd;
  p.cap_lens = cap_lens; p.Bc = Bc; p.Bq = Bq; p.T = T; p.R = R; p.D = D;
  p.g1 = g1; p.g2 = g2; p.g3 = g3; p.eps = eps;
  p.gsim = gsim; p.dctx = dctx; p.dwords = dwords;
  if (int rc = check_shape(p)) return rc;
  if (dctx) TGFR_CUDA_OK(cudaMemsetAsync(dctx, 0, sizeof(float) * (size_t)Bc * R * D, st));
  if (dwords) TGFR_CUDA_OK(cudaMemsetAsync(dwords, 0, sizeof(float) * (size_t)Bq * T * D, st));
  if (!dctx && !dwords) return TGFR_OK;
  return dispatch_bwd<kLoss>(p, st);
}

int attention_fwd_simt(const float* ctx, int64_t csb, int64_t csr, int64_t csd, const float* query, int64_t qsb,
                       int64_t qst, int64_t qsd, const int32_t* cap_lens, int B, int T, int R, int D, float g1,
                       float* wc, float* attn, cudaStream_t st) {
  Params p{};
  p.ctx = ctx; p.csb = csb; p.csr = csr; p.csd = csd;
  p.words = query; p.wsb = qsb; p.wst = qst; p.wsd = qsd;
  p.cap_lens = cap_lens;
  p.Bc = B; p.Bq = B; p.T = T; p.R = R; p.D = D; p.g1 = g1; p.g2 = 1.f; p.g3 = 1.f; p.eps = 1e-8f;
  p.wc = wc; p.attn = attn;
  if (int rc = check_shape(p)) return rc;
  TGFR_REQUIRE(wc != nullptr || attn != nullptr, "attention_fwd: no output requested");
  if (wc == nullptr) {                      // attention maps only: the warp-per-region kernel
    if (p.T <= 8) return launch_attn_only<8>(p, st);
    if (p.T <= 16) return launch_attn_only<16>(p, st);
    if (p.T <= 24) return launch_attn_only<24>(p, st);
    return launch_attn_only<32>(p, st);
  }
  return dispatch_fwd<kAttention>(p, st);
}

int attention_bwd_simt(const float* ctx, int64_t csb, int64_t csr, int64_t csd, const float* query, int64_t qsb,
                       int64_t qst, int64_t qsd, int B, int T, int R, int D, float g1, const float* g_wc,
                       const float* g_attn, float* dctx, float* dquery, cudaStream_t st) {
  Params p{};
  p.ctx = ctx; p.csb = csb; p.csr = csr; p.csd = csd;
  p.words = query; p.wsb = qsb; p.wst = qst; p.wsd = qsd;
  p.Bc = B; p.Bq = B; p.T = T; p.R = R; p.D = D; p.g1 = g1; p.g2 = 1.f; p.g3 = 1.f; p.eps = 1e-8f;
  p.g_wc = g_wc; p.g_attn = g_attn; p.dctx = dctx; p.dwords = dquery;
  if (int rc = check_shape(p)) return rc;
  if (dctx) TGFR_CUDA_OK(cudaMemsetAsync(dctx, 0, sizeof(float) * (size_t)B * R * D, st));
  if (dquery) TGFR_CUDA_OK(cudaMemsetAsync(dquery, 0, sizeof(float) * (size_t)B * T * D, st));
  if (!dctx && !dquery) return TGFR_OK;
  return dispatch_bwd<kAttention>(p, st);
}

}  // namespace tgfr

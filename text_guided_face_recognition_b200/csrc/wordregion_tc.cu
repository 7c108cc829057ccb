// Word-region attention loss on the 5th-generation tensor cores (TGFR_PREC_TC).
//
// One persistent CTA per SM walks a contiguous range of "units".  A unit is one face b and a group
// of `nc` captions whose words fill the 128 columns of the accumulator tile:
//
//   GEMM-1  S[r, w]  = <c_r, q_w>             M = regions (2 tiles of 128 TMEM lanes), N = 128 words,
//                                             K = D; A = C_b (K-major), B = Q_g (K-major)
//   epi-1   A1 = softmax over the words of each caption -- thread-local: a thread owns one region
//           row of the accumulator;  E = exp(g1*(A1-1)) -> fp16 -> shared memory (128B swizzle)
//   GEMM-2  Wu[w, d] = sum_r E[r,w] c_r[d]    M = 128 words, N = D, K = regions; A = E^T (MN-major view
//                                             of the tile epi-1 wrote), B = C_b (MN-major view of the
//                                             very same shared-memory tile GEMM-1 used K-major)
//   epi-2   cos(q_w, Wu_w) (scale invariant, so the region-softmax denominator is never needed),
//           exp(g2 cos), log-sum per caption -> sim[b, i]
//
// Operands are fp16 (11-bit significand = TF32's) with fp32 accumulation; softmax / cosine /
// log-sum-exp statistics stay in fp32.  C_b (<= 104 KB) stays resident in shared memory while the
// CTA works on face b; Q_g is streamed by TMA and prefetched as soon as GEMM-1 has consumed it.
// The B x B x T x R attention tensor never leaves the SM.
//
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread) + TMEM allocator,
// warps 2-9 = epilogue (4 warps per 128-lane region tile; the two groups split D in epi-2).
#include "common.cuh"
#include "tc.cuh"

namespace tgfr {
namespace {

using namespace tc;

constexpr int kThreadsTC = 320;
constexpr int kEpiThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;

enum Bar { kCFull = 0, kQFull, kQEmpty, kSFull, kEFull, kWuFull, kWuEmpty, kNumBars };

struct TcParams {
  const __half* q16;     // [Bq*Tp, D]
  const float* qnorm;    // [Bq*Tp]
  const int* lens;       // [Bq]
  float* sim;            // [Bc, Bq]
  int Bc, Bq, Tp, R, Rp, D, nc, G, nw_rows, n_tiles, total_units;
  uint32_t c_panel, q_panel, e_panel, off_q, off_e, off_misc;
  float k1, k2, g3;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// ---------------------------------------------------------------------------------------------
// prep: fp32 (any strides) -> fp16 canonical copies + exact fp32 word norms + caption lengths
// ---------------------------------------------------------------------------------------------
__global__ void wr_tc_prep_kernel(const float* __restrict__ ctx, int64_t csb, int64_t csr, int64_t csd,
                                  const float* __restrict__ words, int64_t wsb, int64_t wst, int64_t wsd,
                                  const int32_t* __restrict__ cap_lens, int Bc, int Bq, int T, int Tp, int R, int D,
                                  __half* __restrict__ c16, __half* __restrict__ q16, float* __restrict__ qnorm,
                                  int* __restrict__ lens) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int64_t n_ctx = (int64_t)Bc * R, n_w = (int64_t)Bq * Tp;
  if (row < n_ctx) {
    const int b = (int)(row / R), r = (int)(row - (int64_t)b * R);
    const float* src = ctx + b * csb + r * csr;
    __half* dst = c16 + row * D;
    for (int d = lane; d < D; d += 32) dst[d] = __float2half_rn(__ldg(src + (int64_t)d * csd));
  } else if (row < n_ctx + n_w) {
    const int64_t wrow = row - n_ctx;
    const int i = (int)(wrow / Tp), t = (int)(wrow - (int64_t)i * Tp);
    int len = T;
    if (cap_lens) len = min(max(cap_lens[i], 1), T);
    if (t == 0 && lane == 0) lens[i] = len;
    __half* dst = q16 + wrow * D;
    float acc = 0.f;
    if (t < len) {
      const float* src = words + i * wsb + t * wst;
      for (int d = lane; d < D; d += 32) {
        const float v = __ldg(src + (int64_t)d * wsd);
        acc = fmaf(v, v, acc);
        dst[d] = __float2half_rn(v);
      }
    } else {
      for (int d = lane; d < D; d += 32) dst[d] = __float2half_rn(0.f);
    }
    acc = warp_sum(acc);
    if (lane == 0) qnorm[wrow] = sqrtf(acc);
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsTC, 1)
wr_tc_fwd_kernel(const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ CUtensorMap tm_q, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_c = smem;
  uint8_t* s_q = smem + p.off_q;
  uint8_t* s_e = smem + p.off_e;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.off_misc + 64);
  float* part_d = reinterpret_cast<float*>(smem + p.off_misc + 128);   // [2][128]
  float* part_n = part_d + 256;                                        // [2][128]
  float* exs = part_n + 256;                                           // [128]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int u0 = (int)((int64_t)blockIdx.x * p.total_units / gridDim.x);
  const int u1 = (int)((int64_t)(blockIdx.x + 1) * p.total_units / gridDim.x);
  const int kchunks = p.D >> 6;

  if (tid == 0) {
    mbar_init(&bars[kCFull], 1);
    mbar_init(&bars[kQFull], 1);
    mbar_init(&bars[kQEmpty], 1);
    mbar_init(&bars[kSFull], 1);
    mbar_init(&bars[kEFull], kEpiThreads);
    mbar_init(&bars[kWuFull], 1);
    mbar_init(&bars[kWuEmpty], kEpiThreads);
    fence_barrier_init();
    tma_prefetch_desc(&tm_c);
    tma_prefetch_desc(&tm_q);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int prev_b = -1, n = 0;
      for (int u = u0; u < u1; ++u, ++n) {
        const int b = u / p.G, g = u - b * p.G;
        // GEMM-1(n-1) has retired => GEMM-2(n-2) has too, so kWuFull is at most one phase behind
        // the parity tested below (waiting on it before this point could alias two phases).
        mbar_wait(&bars[kQEmpty], (n & 1) ^ 1);
        if (b != prev_b) {
          if (n > 0) mbar_wait(&bars[kWuFull], (n - 1) & 1);      // every MMA reading the old C_b has retired
          mbar_arrive_expect_tx(&bars[kCFull], kchunks * p.c_panel);
          for (int kc = 0; kc < kchunks; ++kc) tma_load_3d(s_c + kc * p.c_panel, &tm_c, &bars[kCFull], kc * 64, 0, b);
          prev_b = b;
        }
        mbar_arrive_expect_tx(&bars[kQFull], kchunks * p.q_panel);
        for (int kc = 0; kc < kchunks; ++kc)
          tma_load_3d(s_q + kc * p.q_panel, &tm_q, &bars[kQFull], kc * 64, g * p.nw_rows, 0);
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc_f16(128, 128, false, false);
      const uint32_t idesc2 = make_idesc_f16(128, p.D, true, true);
      const uint32_t a_c = smem_u32(s_c), a_q = smem_u32(s_q), a_e = smem_u32(s_e);
      int prev_b = -1, n = 0, m = -1;
      for (int u = u0; u < u1; ++u, ++n) {
        const int b = u / p.G;
        if (b != prev_b) {
          ++m;
          mbar_wait(&bars[kCFull], m & 1);
          prev_b = b;
        }
        mbar_wait(&bars[kQFull], n & 1);
        tc_fence_after();
        for (int t = 0; t < p.n_tiles; ++t) {
          for (int k16 = 0; k16 < (p.D >> 4); ++k16) {
            const uint64_t ad = make_smem_desc(a_c + (k16 >> 2) * p.c_panel + t * (128 * 128) + (k16 & 3) * 32, 16, 1024);
            const uint64_t bd = make_smem_desc(a_q + (k16 >> 2) * p.q_panel + (k16 & 3) * 32, 16, 1024);
            umma_ss(tmem + t * 128, ad, bd, idesc1, k16 > 0);
          }
        }
        umma_commit(&bars[kQEmpty]);
        umma_commit(&bars[kSFull]);
        mbar_wait(&bars[kEFull], n & 1);
        mbar_wait(&bars[kWuEmpty], (n & 1) ^ 1);
        tc_fence_after();
        for (int j = 0; j < (p.Rp >> 4); ++j) {
          const uint64_t ad = make_smem_desc(a_e + j * 2048, p.e_panel, 1024);
          const uint64_t bd = make_smem_desc(a_c + j * 2048, p.c_panel, 1024);
          umma_ss(tmem + 256, ad, bd, idesc2, j > 0);
        }
        umma_commit(&bars[kWuFull]);
      }
    }
  } else {
    // ======================================= epilogue =======================================
    const int tile = (warp - 2) >> 2;            // region tile in epi-1, D half in epi-2
    const int quarter = warp & 3;                // TMEM lane quarter this warp may touch
    const int lrow = quarter * 32 + lane;        // lane row 0..127
    const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
    const int r = tile * 128 + lrow;
    const bool warp_has_rows = tile < p.n_tiles && (tile * 128 + quarter * 32) < p.Rp;
    const int dhalf = p.D >> 1;
    int n = 0;
    for (int u = u0; u < u1; ++u, ++n) {
      const int b = u / p.G, g = u - b * p.G;
      // ---------------- epi-1: word softmax, E -> shared memory ----------------
      mbar_wait(&bars[kSFull], n & 1);
      tc_fence_after();
      if (warp_has_rows) {
        for (int c = 0; c < p.nc; ++c) {
          const int i = g * p.nc + c;
          if (i >= p.Bq) break;
          const int len = __ldg(p.lens + i);
          uint32_t v[32];
          const uint32_t col = tmem + t_lane + tile * 128 + c * p.Tp;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (8 * j < p.Tp) tmem_ld8(col + 8 * j, v + 8 * j);
          tmem_ld_wait();
          float mx = -INFINITY;
#pragma unroll
          for (int t = 0; t < 32; ++t)
            if (t < len) mx = fmaxf(mx, __uint_as_float(v[t]));
          float sum = 0.f;
          float e[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            e[t] = (t < len) ? fast_exp2((__uint_as_float(v[t]) - mx) * kLog2e) : 0.f;
            sum += e[t];
          }
          const float inv = 1.f / sum;
#pragma unroll
          for (int t = 0; t < 32; ++t) e[t] = (t < len) ? fast_exp2(p.k1 * (e[t] * inv) - p.k1) : 0.f;
          if (r < p.Rp) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (8 * j < p.Tp) {
                const int w0 = c * p.Tp + 8 * j;
                uint4 pk;
                pk.x = pack_half2(e[8 * j + 0], e[8 * j + 1]);
                pk.y = pack_half2(e[8 * j + 2], e[8 * j + 3]);
                pk.z = pack_half2(e[8 * j + 4], e[8 * j + 5]);
                pk.w = pack_half2(e[8 * j + 6], e[8 * j + 7]);
                *reinterpret_cast<uint4*>(s_e + (w0 >> 6) * p.e_panel + sw128_offset(r, (w0 & 63) >> 3)) = pk;
              }
            }
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bars[kEFull]);

      // ---------------- epi-2: cosine, exp, log-sum ----------------
      mbar_wait(&bars[kWuFull], n & 1);
      tc_fence_after();
      const int w = lrow;
      const int c = w / p.Tp, t = w - c * p.Tp;
      const int i = g * p.nc + c;
      const bool valid = (w < p.nw_rows) && (i < p.Bq) && (t < __ldg(p.lens + min(i, p.Bq - 1)));
      const int64_t qrow = (int64_t)min(i, p.Bq - 1) * p.Tp + t;
      float dot = 0.f, n2 = 0.f;
      for (int ch = 0; ch < (dhalf >> 5); ++ch) {
        uint32_t v[32];
        tmem_ld32(tmem + t_lane + 256 + tile * dhalf + 32 * ch, v);
        tmem_ld_wait();
        if (valid) {
          const uint4* qp = reinterpret_cast<const uint4*>(p.q16 + qrow * p.D + tile * dhalf + 32 * ch);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const uint4 qv = __ldg(qp + cc);
            const __half2* qh = reinterpret_cast<const __half2*>(&qv);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 qf = __half22float2(qh[k]);
              const float w0 = __uint_as_float(v[8 * cc + 2 * k]), w1 = __uint_as_float(v[8 * cc + 2 * k + 1]);
              dot = fmaf(qf.x, w0, dot);
              dot = fmaf(qf.y, w1, dot);
              n2 = fmaf(w0, w0, n2);
              n2 = fmaf(w1, w1, n2);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&bars[kWuEmpty]);
      part_d[tile * 128 + w] = dot;
      part_n[tile * 128 + w] = n2;
      epi_bar_sync();
      if (tile == 0) {
        float ex = 0.f;
        if (valid) {
          const float dd = part_d[w] + part_d[128 + w], nn = part_n[w] + part_n[128 + w];
          const float den = fmaxf(__ldg(p.qnorm + qrow) * sqrtf(nn), 1e-30f);
          ex = fast_exp2(p.k2 * (dd / den));
        }
        exs[w] = ex;
      }
      epi_bar_sync();
      if (tile == 0 && w < p.nc) {
        const int ii = g * p.nc + w;
        if (ii < p.Bq) {
          float s = 0.f;
          for (int tt = 0; tt < p.Tp; ++tt) s += exs[w * p.Tp + tt];
          p.sim[(int64_t)b * p.Bq + ii] = p.g3 * logf(s);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

struct TcPlan {
  int Tp, Rp, nc, G, nw_rows, n_tiles;
  uint32_t c_panel, q_panel, e_panel, off_q, off_e, off_misc, smem_bytes;
  size_t ws_c16, ws_q16, ws_qnorm, ws_lens, ws_total;
};

int make_plan(int Bc, int Bq, int T, int R, int D, TcPlan* pl) {
  TGFR_REQUIRE(D % 64 == 0 && D >= 64 && D <= 256, "wordregion(tc): D=%d must be 64, 128, 192 or 256", D);
  TGFR_REQUIRE(R >= 1 && R <= 256, "wordregion(tc): R=%d regions (max 256)", R);
  TGFR_REQUIRE(T >= 1 && T <= 32, "wordregion(tc): T=%d words (max 32)", T);
  pl->Tp = (T + 7) & ~7;
  pl->Rp = (R + 15) & ~15;
  pl->n_tiles = (pl->Rp + 127) / 128;
  pl->nc = 128 / pl->Tp;
  pl->nw_rows = pl->nc * pl->Tp;
  pl->G = (Bq + pl->nc - 1) / pl->nc;
  const int kch = D / 64;
  pl->c_panel = (uint32_t)pl->Rp * 128u;
  pl->q_panel = (uint32_t)pl->nw_rows * 128u;
  pl->e_panel = (uint32_t)pl->Rp * 128u;
  pl->off_q = kch * pl->c_panel;
  pl->off_e = pl->off_q + kch * pl->q_panel;
  pl->off_misc = pl->off_e + 2 * pl->e_panel;
  pl->smem_bytes = pl->off_misc + 4096 + 1024;        // misc + alignment slack
  TGFR_REQUIRE(pl->smem_bytes <= 232448, "wordregion(tc): shared memory plan needs %u bytes", pl->smem_bytes);
  pl->ws_c16 = 0;
  pl->ws_q16 = align_up((size_t)Bc * R * D * 2, 256);
  pl->ws_qnorm = pl->ws_q16 + align_up((size_t)Bq * pl->Tp * D * 2, 256);
  pl->ws_lens = pl->ws_qnorm + align_up((size_t)Bq * pl->Tp * 4, 256);
  pl->ws_total = pl->ws_lens + align_up((size_t)Bq * 4, 256);
  return TGFR_OK;
}

}  // namespace

size_t wordregion_tc_workspace_bytes(int Bc, int Bq, int T, int R, int D) {
  TcPlan pl;
  if (make_plan(Bc, Bq, T, R, D, &pl) != TGFR_OK) return 0;
  return pl.ws_total;
}

int wordregion_fwd_tc(const float* ctx, int64_t csb, int64_t csr, int64_t csd, const float* words, int64_t wsb,
                      int64_t wst, int64_t wsd, const int32_t* cap_lens, int Bc, int Bq, int T, int R, int D, float g1,
                      float g2, float g3, float eps, float* sim, void* ws, size_t ws_bytes, cudaStream_t st) {
  (void)eps;
  TcPlan pl;
  if (int rc = make_plan(Bc, Bq, T, R, D, &pl)) return rc;
  TGFR_REQUIRE(ws != nullptr && ws_bytes >= pl.ws_total, "wordregion(tc): workspace too small (%zu < %zu)", ws_bytes,
               pl.ws_total);
  TGFR_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "wordregion(tc): workspace must be 256-byte aligned");
  uint8_t* base = reinterpret_cast<uint8_t*>(ws);
  __half* c16 = reinterpret_cast<__half*>(base + pl.ws_c16);
  __half* q16 = reinterpret_cast<__half*>(base + pl.ws_q16);
  float* qnorm = reinterpret_cast<float*>(base + pl.ws_qnorm);
  int* lens = reinterpret_cast<int*>(base + pl.ws_lens);

  const int64_t rows = (int64_t)Bc * R + (int64_t)Bq * pl.Tp;
  wr_tc_prep_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(ctx, csb, csr, csd, words, wsb, wst, wsd, cap_lens, Bc, Bq,
                                                               T, pl.Tp, R, D, c16, q16, qnorm, lens);
  TGFR_LAUNCH_OK();

  CUtensorMap tm_c, tm_q;
  if (int rc = make_tmap_3d(&tm_c, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, c16, D, R, Bc, 64, pl.Rp, 1)) return rc;
  if (int rc = make_tmap_3d(&tm_q, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, q16, D, (uint64_t)Bq * pl.Tp, 1, 64, pl.nw_rows, 1))
    return rc;

  TcParams p{};
  p.q16 = q16; p.qnorm = qnorm; p.lens = lens; p.sim = sim;
  p.Bc = Bc; p.Bq = Bq; p.Tp = pl.Tp; p.R = R; p.Rp = pl.Rp; p.D = D; p.nc = pl.nc; p.G = pl.G;
  p.nw_rows = pl.nw_rows; p.n_tiles = pl.n_tiles; p.total_units = Bc * pl.G;
  p.c_panel = pl.c_panel; p.q_panel = pl.q_panel; p.e_panel = pl.e_panel;
  p.off_q = pl.off_q; p.off_e = pl.off_e; p.off_misc = pl.off_misc;
  p.k1 = g1 * kLog2e; p.k2 = g2 * kLog2e; p.g3 = g3;

  int dev = 0, sms = 0;
  TGFR_CUDA_OK(cudaGetDevice(&dev));
  TGFR_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = p.total_units < sms ? p.total_units : sms;
  TGFR_CUDA_OK(cudaFuncSetAttribute(wr_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
  wr_tc_fwd_kernel<<<grid, kThreadsTC, pl.smem_bytes, st>>>(tm_c, tm_q, p);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

}  // namespace tgfr

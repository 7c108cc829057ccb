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
// The fp32 B x B x T x R attention tensor of the reference is never formed in HBM.  What the forward leaves for the
// backward is the caller's choice (TGFR_WORDREGION_SAVE; the layouts are told apart by the buffer size):
//   records (default)   fp16 (A1 | E) per (face, caption, word, region) + V planes: the backward (rec::wr_tc_bwd2_kernel)
//                       runs no score GEMM and no exponential                 (kLayRec; its kernels: namespace rec below)
//   wu                  fp16 Wu planes + (alpha, beta) per word: 2.5x fewer bytes, the backward (wr_tc_bwd3_kernel)
//                       recomputes S and E on the tensor cores / MUFU
//   none                wr_tc_bwd_kernel recomputes everything (also the d words path)
//
// Roles in the forward (wr_tc_fwd3_kernel, 896 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected
// thread) + TMEM allocator, warps 4-11 = group A (epi-1: word softmax, E), warps 12-27 = group B (epi-2: cosine,
// log-sum, saved state); registers are re-dealt with setmaxnreg (40 / 104 / 64), group A runs one unit ahead of B.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc.cuh"

namespace tgfr {
namespace {

using namespace tc;

constexpr int kThreadsTC = 320;
constexpr int kEpiThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;

enum Bar { kCFull = 0, kQFull, kQEmpty, kSFull0, kSFull1, kEFull0, kEFull1, kWuFull, kWuEmpty, kNumBars };

// what the forward leaves for the backward (TGFR_WORDREGION_SAVE; see the header of this file)
enum { kLayNone = 0, kLayWu = 1, kLayRec = 2 };
namespace rec { constexpr float kSV = 256.f; }   // power-of-two scale of the saved V tile (keeps small word weights in the fp16 normal range)

struct TcParams {
  const __half* q16;     // [Bq*Tp, D]
  const float* qnorm;    // [Bq*Tp]
  const int* lens;       // [Bq]
  float* sim;            // [Bc, Bq]
  float* attn_out;       // [Bc, T, R] diagonal attention maps (un-normalised here) or NULL
  float* attn_z;         // [Bc, 32] their region sums
  int diag_off, T;
  // Wu layout: what the record-free backward (wr_tc_bwd3_kernel) needs beside its own recomputation of S, per unit u = b * G + g:
  __half* sv_wu;         // [total_units][D/8][nw_rows][8]  Wu_w as fp16 planes (the un-normalised context of word w)
  float* sv_ab;          // [total_units][2][128]  alpha_w = g2 g3 p_w / (|q_w| |Wu_w|), beta_w = g2 g3 p_w cos_w / |Wu_w|^2:
                         //   d sim[b,i] / d Wu_w = alpha_w q_w - beta_w Wu_w   (0 for padding words and missing captions)
  // records layout (rec::wr_tc_bwd2_kernel reads these instead of recomputing), per unit u = b * G + g:
  __half* sv_v;          // [total_units][D/8][nw_rows][8]  V_w = kSV p_w (q^_w - cos_w w^_w): d sim / d Wu up to a per-caption scalar
  uint8_t* sv_rec;       // [total_units][nc][Tp/4 chunks of 8 fp16: A1 x Tp | E x Tp][Rp]   word softmax, exp(g1 (A1 - 1))
  float* sv_inw;         // [total_units][128]           1 / |Wu_w| (0 for padding words and missing captions)
  uint32_t rec_stride;   // bytes of one unit's records
  int Bc, Bq, Tp, R, Rp, D, nc, G, nw_rows, n_tiles, total_units;
  int uniform_len;       // > 0: every caption has this many words (no cap_lens given); else read `lens`
  uint32_t c_panel, q_panel, e_panel, off_q, off_e, off_misc;
  float k1, k2, g3, g23;
};

// optional phase trace (tgfr_debug_set_trace): clock64 stamps of CTA 0's first units
__device__ long long* g_trace = nullptr;
#define TGFR_TRACE(n, ev)                                                      \
  do {                                                                         \
    long long* _t = g_trace;                                                   \
    if (_t && blockIdx.x == 0 && (n) < 16) _t[(n) * 32 + (ev)] = clock64();    \
  } while (0)

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
  // range guard: lens[Bq] / lens[Bq + 1] collect max |ctx| / max |words| (float bits; zeroed by the host before the
  // launch).  The fp16 copies carry no per-tensor scale, so the forward poisons `sim` with NaN when either tensor leaves
  // the range fp16 holds well (range_bad below) instead of returning a silently degraded loss.
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int64_t n_ctx = (int64_t)Bc * R, n_w = (int64_t)Bq * Tp;
  int* flags = lens + Bq;
  if (row < n_ctx) {
    const int b = (int)(row / R), r = (int)(row - (int64_t)b * R);
    const float* src = ctx + b * csb + r * csr;
    __half* dst = c16 + row * D;
    float m = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float v = __ldg(src + (int64_t)d * csd);
      m = fmaxf(m, fabsf(v));
      dst[d] = __float2half_rn(v);
    }
    m = warp_max(m);
    if (lane == 0 && __float_as_int(m) > __ldcg(flags)) atomicMax(flags, __float_as_int(m));
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
      float m = 0.f;
      for (int d = lane; d < D; d += 32) {
        const float v = __ldg(src + (int64_t)d * wsd);
        acc = fmaf(v, v, acc);
        m = fmaxf(m, fabsf(v));
        dst[d] = __float2half_rn(v);
      }
      m = warp_max(m);
      if (lane == 0 && __float_as_int(m) > __ldcg(flags + 1)) atomicMax(flags + 1, __float_as_int(m));
    } else {
      for (int d = lane; d < D; d += 32) dst[d] = __float2half_rn(0.f);
    }
    acc = warp_sum(acc);
    if (lane == 0) qnorm[wrow] = sqrtf(acc);
  }
}
// largest entry of either operand outside [2^-9, 65504]: beyond fp16, or so small that typical entries are subnormal
// (unit-norm rows of up to 256 features have a largest entry >= 1/16)
__device__ __forceinline__ bool range_bad(const int* __restrict__ lens, int Bq) {
  const float mc = __int_as_float(__ldg(lens + Bq)), mq = __int_as_float(__ldg(lens + Bq + 1));
  return !(mc >= 0.001953125f && mc <= 65504.f && mq >= 0.001953125f && mq <= 65504.f);
}

// The diagonal pair's attention map (losses.py:97; the B maps words_loss returns) straight from epi-1: for the caption
// i = b + diag_off of face b the un-normalised region-softmax numerators E[r, t] = exp(g1 (A1 - 1)) go to attn [b, t, r]
// and their sums over the regions to z [b, t] (zeroed by the host); attn_normalize_kernel divides afterwards.  e[] holds
// exp(s - max) of the word softmax, kinv = k1 / sum, nk1 = -k1.  One caption in Bq per face: the cost is nil, and the
// separate attention-only kernel (36 us at B = 128) goes away.  The maps inherit the fp16-operand scores (TF32 class).
// (not inlined: one caption in Bq takes this path, and inlined it costs the softmax loop registers -- spills)
template <int TP>
__device__ __noinline__ void emit_diag_attention(const float* e, float kinv, float nk1, int len, int T, int R, int r, int lane,
                                                 float* __restrict__ attn_b, float* __restrict__ z_b) {
  for (int t = 0; t < TP; ++t) {
    if (t < len && t < T) {                                       // warp-uniform: len belongs to the caption
      const float val = (r < R) ? fast_exp2(fmaf(e[t], kinv, nk1)) : 0.f;
      if (r < R) attn_b[(int64_t)t * R + r] = val;
      const float s = warp_sum(val);
      if (lane == 0) atomicAdd(z_b + t, s);
    }
  }
}
// attn [b, t, r] /= z [b, t] for the words of caption b + diag_off, zeros beyond its length (and for a face without one)
__global__ void attn_normalize_kernel(float* __restrict__ attn, const float* __restrict__ z, const int* __restrict__ lens, int Bc,
                                      int Bq, int T, int R, int ztp, int diag_off, int uniform_len) {
  const int64_t n = (int64_t)Bc * T * R;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bt = k / R;
    const int b = (int)(bt / T), t = (int)(bt - (int64_t)b * T);
    const int i = b + diag_off;
    int len = 0;
    if (i >= 0 && i < Bq) len = uniform_len > 0 ? uniform_len : __ldg(lens + i);
    attn[k] = (t < len) ? attn[k] / __ldg(z + (int64_t)b * ztp + t) : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// forward, pipelined (wr_tc_fwd3_kernel): the same unit walk and the same arithmetic as wr_tc_fwd_kernel, but the two
// epilogues belong to different warps, so that a unit's cosine / log-sum (and V) work overlaps the NEXT unit's
// word softmax, and the tensor core runs one region tile ahead of the softmax warps:
//
//   warp 0          TMA producer            warp 1   MMA issuer + TMEM allocator            warps 2-3   idle
//   group A (epi-1) kF3NA warps per TMEM lane quarter: word softmax of a (region tile, caption) task, E -> shared
//                   memory (and the A1 | E records).  MUFU bound: 2 exponentials per (region, word).
//   group B (epi-2) kF3NB warps per lane quarter (a thread = one word row x D / kF3NB features, its slice of q_w held
//                   in registers): cos(q_w, Wu_w), exp, log-sum -> sim, and the V tile.  FMA / TMEM-read bound.
//   Registers follow the roles (setmaxnreg; 896 x 72 at launch): warps 0-3 release 32 each and the cosine warps 8 each,
//   exactly what lets the softmax warps grow to 104 (a request beyond the CTA's pool would spin forever).
//
//   MMA order per unit n:  [E tile 0 of n written]  GEMM-2(n) K steps over tile 0 | GEMM-1(n+1) tile 0   (whichever
//                          has its other input first: Wu drained by group B / Q(n+1) landed)
//                          [E tile 1 of n written]  GEMM-2(n) K steps over tile 1,  GEMM-1(n+1) tile 1
//   so S tile t of unit n+1 overwrites S tile t of unit n as soon as group A has read it (tile-level ping-pong of the
//   two 128-column accumulators), E tile t is rewritten as soon as GEMM-2's instalment over it has retired, and Wu is
//   the only per-unit resource group B has to hand back (f3WuEmpty).  A change of face drains the pipeline (the C
//   tile is single-buffered: 104 KB).
// ---------------------------------------------------------------------------------------------
enum Bar3 { f3CFull = 0, f3QFull, f3QEmpty, f3SFull0, f3SFull1, f3E0Full, f3E1Full, f3E0Free, f3WuFull, f3WuEmpty, f3Num };

constexpr int kF3NA = 2;
constexpr int kF3NB = 4;
constexpr int kF3ThreadsA = 128 * kF3NA;
constexpr int kF3ThreadsB = 128 * kF3NB;
constexpr int kF3Threads = 128 + kF3ThreadsA + kF3ThreadsB;   // 896: 72 registers per thread at launch

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Wait of an epilogue warp: 24 warps wait on barriers here, and a polling warp takes issue slots from the working
// ones (a first profile showed more than half of all issued instructions in try_wait loops), so a failed poll
// sleeps before the next one.  A wait that lasts seconds is a protocol bug: trap instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait_lazy(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t it = 0;; ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    asm volatile("nanosleep.u32 128;" ::: "memory");
    if (it > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void f3b_bar_sync() { asm volatile("bar.sync 2, %0;" ::"n"(kF3ThreadsB) : "memory"); }
template <int N> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

template <int TP, int LAY>
__global__ void __launch_bounds__(kF3Threads, 1)
wr_tc_fwd3_kernel(const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ CUtensorMap tm_q, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_c = smem;
  uint8_t* s_q = smem + p.off_q;
  uint8_t* s_e = smem + p.off_e;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.off_misc + 128);
  float* part_d = reinterpret_cast<float*>(smem + p.off_misc + 256);   // [kF3NB][128]
  float* part_n = part_d + kF3NB * 128;                                // [kF3NB][128]
  float* exs = part_n + kF3NB * 128;                                   // [128]
  float* cosw = exs + 128;                                             // [128] SAVE: cos_w
  float* inww = cosw + 128;                                            // [128] SAVE: 1 / |Wu_w|

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int u0 = (int)((int64_t)blockIdx.x * p.total_units / gridDim.x);
  const int u1 = (int)((int64_t)(blockIdx.x + 1) * p.total_units / gridDim.x);
  const int kchunks = p.D >> 6;

  if (tid == 0) {
    mbar_init(&bars[f3CFull], 1);
    mbar_init(&bars[f3QFull], 1);
    mbar_init(&bars[f3QEmpty], 1);
    mbar_init(&bars[f3SFull0], 1);
    mbar_init(&bars[f3SFull1], 1);
    mbar_init(&bars[f3E0Full], kF3ThreadsA);
    mbar_init(&bars[f3E1Full], kF3ThreadsA);
    mbar_init(&bars[f3E0Free], 1);
    mbar_init(&bars[f3WuFull], 1);
    mbar_init(&bars[f3WuEmpty], kF3ThreadsB);
    fence_barrier_init();
    tma_prefetch_desc(&tm_c);
    tma_prefetch_desc(&tm_q);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    reg_dealloc<40>();
    if (warp == 0) {
      // ===================================== TMA producer =====================================
      if (lane == 0) {
        int prev_b = -1, n = 0;
        for (int u = u0; u < u1; ++u, ++n) {
          const int b = u / p.G, g = u - b * p.G;
          mbar_wait_lazy(&bars[f3QEmpty], (n & 1) ^ 1);               // GEMM-1(n-1) has read the Q tile
          if (b != prev_b) {
            if (n > 0) mbar_wait_lazy(&bars[f3WuFull], (n - 1) & 1);  // every MMA reading the old C_b has retired
            mbar_arrive_expect_tx(&bars[f3CFull], kchunks * p.c_panel);
            for (int kc = 0; kc < kchunks; ++kc) tma_load_3d(s_c + kc * p.c_panel, &tm_c, &bars[f3CFull], kc * 64, 0, b);
            prev_b = b;
          }
          mbar_arrive_expect_tx(&bars[f3QFull], kchunks * p.q_panel);
          for (int kc = 0; kc < kchunks; ++kc)
            tma_load_3d(s_q + kc * p.q_panel, &tm_q, &bars[f3QFull], kc * 64, g * p.nw_rows, 0);
        }
      }
    } else if (warp == 1) {
      // ====================================== MMA issuer ======================================
      if (lane == 0 && u0 < u1) {
        const uint32_t idesc1 = make_idesc_f16(128, 128, false, false);
        const uint32_t idesc2 = make_idesc_f16(128, p.D, true, true);
        const uint32_t a_c = smem_u32(s_c), a_q = smem_u32(s_q), a_e = smem_u32(s_e);
        const int j0 = min(p.Rp >> 4, 8), j1 = p.Rp >> 4;
        auto gemm1 = [&](int t) {                                   // S_t = C_t . Q^T
          for (int k16 = 0; k16 < (p.D >> 4); ++k16) {
            const uint64_t ad = make_smem_desc(a_c + (k16 >> 2) * p.c_panel + t * (128 * 128) + (k16 & 3) * 32, 16, 1024);
            const uint64_t bd = make_smem_desc(a_q + (k16 >> 2) * p.q_panel + (k16 & 3) * 32, 16, 1024);
            umma_ss(tmem + t * 128, ad, bd, idesc1, k16 > 0);
          }
        };
        auto gemm2 = [&](int ja, int jb) {                          // Wu (+)= E^T . C over K steps [ja, jb)
          for (int j = ja; j < jb; ++j) {
            const uint64_t ad = make_smem_desc(a_e + j * 2048, p.e_panel, 1024);
            const uint64_t bd = make_smem_desc(a_c + j * 2048, p.c_panel, 1024);
            umma_ss(tmem + 256, ad, bd, idesc2, j > 0);
          }
        };
        int m = 0;
        mbar_wait(&bars[f3CFull], 0);
        mbar_wait(&bars[f3QFull], 0);
        tc_fence_after();
        gemm1(0);
        umma_commit(&bars[f3SFull0]);
        if (p.n_tiles == 2) gemm1(1);
        umma_commit(&bars[f3SFull1]);
        umma_commit(&bars[f3QEmpty]);
        int n = 0;
        for (int u = u0; u < u1; ++u, ++n) {
          const int b = u / p.G;
          const bool has_next = u + 1 < u1;
          const bool boundary = has_next && ((u + 1) / p.G != b);
          mbar_wait_lazy(&bars[f3E0Full], n & 1);                    // E tile 0 of unit n written, S tile 0 read
          TGFR_TRACE(n, 17);
          bool g1_pending = has_next && !boundary, g2_pending = true;
          while (g1_pending || g2_pending) {
            if (g2_pending && mbar_test(&bars[f3WuEmpty], (n & 1) ^ 1)) {        // group B has drained Wu(n-1)
              tc_fence_after();
              gemm2(0, j0);
              umma_commit(&bars[f3E0Free]);
              g2_pending = false;
            } else if (g1_pending && mbar_test(&bars[f3QFull], (n + 1) & 1)) {    // Q(n+1) has landed
              tc_fence_after();
              gemm1(0);
              umma_commit(&bars[f3SFull0]);
              g1_pending = false;
            } else {
              asm volatile("nanosleep.u32 64;" ::: "memory");
            }
          }
          TGFR_TRACE(n, 18);
          mbar_wait_lazy(&bars[f3E1Full], n & 1);                    // E tile 1 written, S tile 1 read
          tc_fence_after();
          gemm2(j0, j1);
          umma_commit(&bars[f3WuFull]);
          TGFR_TRACE(n, 19);
          if (has_next) {
            if (boundary) {                                          // new face: the C tile is reloaded once Wu(n) is complete
              ++m;
              mbar_wait(&bars[f3CFull], m & 1);
              mbar_wait(&bars[f3QFull], (n + 1) & 1);
              tc_fence_after();
              gemm1(0);
              umma_commit(&bars[f3SFull0]);
            }
            if (p.n_tiles == 2) gemm1(1);
            umma_commit(&bars[f3SFull1]);
            umma_commit(&bars[f3QEmpty]);
          }
        }
      }
    }
  } else if (warp < 4 + 4 * kF3NA) {
    // ======================================= group A: epi-1 =======================================
    reg_alloc<104>();
    const int grp = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int lrow = quarter * 32 + lane;
    const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
    const uint32_t rx = (uint32_t)(lrow & 7);
    int n = 0;
    for (int u = u0; u < u1; ++u, ++n) {
      const int g = u % p.G;
      for (int tile = 0; tile < 2; ++tile) {
        // every thread follows both tile barriers (even without rows there): an arrival may never run a phase ahead
        mbar_wait_lazy(&bars[tile ? f3SFull1 : f3SFull0], n & 1);
        const bool has = tile < p.n_tiles && (tile * 128 + quarter * 32) < p.Rp;
        if (has) {
          tc_fence_after();
          // GEMM-2(n-1)'s instalment over this tile's rows of E has retired
          mbar_wait_lazy(&bars[tile ? f3WuFull : f3E0Free], (n & 1) ^ 1);
          if (tid == 128 && tile == 0) TGFR_TRACE(n, 2);
          const int r = tile * 128 + lrow;
          uint8_t* const e_row = s_e + (uint32_t)r * 128u;
          const int diag_i = p.attn_out != nullptr ? u / p.G + p.diag_off : -1;
          for (int c = (grp + tile) % kF3NA; c < p.nc; c += kF3NA) {
            const int i = g * p.nc + c;
            if (i >= p.Bq) {
              if constexpr (LAY == kLayRec) {                               // missing captions: zero records for the backward
                if (r < p.Rp) {
                  uint4* rdst = reinterpret_cast<uint4*>(p.sv_rec + (int64_t)u * p.rec_stride) + (int64_t)c * (TP / 4) * p.Rp + r;
#pragma unroll
                  for (int j = 0; j < TP / 4; ++j) rdst[(int64_t)j * p.Rp] = make_uint4(0, 0, 0, 0);
                }
              }
              continue;
            }
            const int len = p.uniform_len > 0 ? p.uniform_len : __ldg(p.lens + i);
            uint32_t v[TP];
            const uint32_t col = tmem + t_lane + tile * 128 + c * TP;
#pragma unroll
            for (int j = 0; j < TP / 8; ++j) tmem_ld8(col + 8 * j, v + 8 * j);
            tmem_ld_wait();
            float e[TP];
#pragma unroll
            for (int t = 0; t < TP; ++t) e[t] = __uint_as_float(v[t]);
            if (len < TP) {                                       // padding words: the last 8 columns, or a ragged caption
              if (len > TP - 8) {
#pragma unroll
                for (int t = TP - 8; t < TP; ++t) e[t] = (t < len) ? e[t] : -INFINITY;
              } else {
#pragma unroll
                for (int t = 0; t < TP; ++t) e[t] = (t < len) ? e[t] : -INFINITY;
              }
            }
            float mxp[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
            for (int t = 0; t < TP; ++t) mxp[t & 3] = fmaxf(mxp[t & 3], e[t]);
            const float mx = fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3]));
            const float nmx = -mx * kLog2e;
            float sump[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int t = 0; t < TP; ++t) {
              e[t] = fast_exp2(fmaf(e[t], kLog2e, nmx));
              sump[t & 3] += e[t];
            }
            const float sum = (sump[0] + sump[1]) + (sump[2] + sump[3]);
            const float nk1 = -p.k1;
            if (i == diag_i) {                                       // this face's own caption: its map is an output
              float ecopy[TP];
#pragma unroll
              for (int t = 0; t < TP; ++t) ecopy[t] = e[t];
              emit_diag_attention<TP>(ecopy, p.k1 / sum, nk1, len, p.T, p.R, r, lane, p.attn_out + (int64_t)(u / p.G) * p.T * p.R,
                                      p.attn_z + (int64_t)(u / p.G) * 32);
            }
            uint32_t pe[TP / 2];
            uint32_t pa[LAY == kLayRec ? TP / 2 : 1];
            if constexpr (LAY == kLayRec) {
              const bool live_row = r < p.R;
              const float inv = live_row ? 1.f / sum : 0.f;
#pragma unroll
              for (int t = 0; t < TP; t += 2) {
                const float a0 = e[t] * inv, a1 = e[t + 1] * inv;
                pa[t >> 1] = pack_half2(a0, a1);
                const uint32_t pk = pack_half2(fast_exp2(fmaf(a0, p.k1, nk1)), fast_exp2(fmaf(a1, p.k1, nk1)));
                pe[t >> 1] = live_row ? pk : 0u;
              }
            } else {
              const float kinv = p.k1 / sum;
#pragma unroll
              for (int t = 0; t < TP; t += 2)
                pe[t >> 1] = pack_half2(fast_exp2(fmaf(e[t], kinv, nk1)), fast_exp2(fmaf(e[t + 1], kinv, nk1)));
            }
            if (r < p.Rp) {
              // E first: GEMM-2 waits for it; the records (global stores, LSU bound) follow
              const uint32_t w0 = (uint32_t)(c * TP);
#pragma unroll
              for (int j = 0; j < TP / 8; ++j) {
                const uint32_t ww = w0 + 8u * j;
                *reinterpret_cast<uint4*>(e_row + (ww >> 6) * p.e_panel + ((((ww & 63u) >> 3) ^ rx) << 4)) =
                    make_uint4(pe[4 * j], pe[4 * j + 1], pe[4 * j + 2], pe[4 * j + 3]);
              }
              if constexpr (LAY == kLayRec) {
                uint4* rdst = reinterpret_cast<uint4*>(p.sv_rec + (int64_t)u * p.rec_stride) + (int64_t)c * (TP / 4) * p.Rp + r;
#pragma unroll
                for (int j = 0; j < TP / 8; ++j) {
                  rdst[(int64_t)j * p.Rp] = make_uint4(pa[4 * j], pa[4 * j + 1], pa[4 * j + 2], pa[4 * j + 3]);
                  rdst[(int64_t)(TP / 8 + j) * p.Rp] = make_uint4(pe[4 * j], pe[4 * j + 1], pe[4 * j + 2], pe[4 * j + 3]);
                }
              }
            }
          }
          fence_proxy_async();
          tc_fence_before();
        }
        mbar_arrive(&bars[tile ? f3E1Full : f3E0Full]);
        if (tid == 128) TGFR_TRACE(n, 3 + tile);
      }
    }
  } else {
    // ======================================= group B: epi-2 =======================================
    reg_dealloc<64>();
    const int grp = (warp - 4 - 4 * kF3NA) >> 2;         // which slice of the features
    const int quarter = warp & 3;
    const int w = quarter * 32 + lane;                    // word row = TMEM lane
    const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
    const int dq = p.D / kF3NB;                           // features per thread (16 .. 64)
    const int nch = dq >> 3;                              // 8-column TMEM loads per pass (2 .. 8)
    const int c = w / TP, t = w - c * TP;
    const uint32_t wu_col = tmem + t_lane + 256 + grp * dq;
    int n = 0;
    for (int u = u0; u < u1; ++u, ++n) {
      const int b = u / p.G, g = u - b * p.G;
      const int i = g * p.nc + c;
      const int len = (i < p.Bq) ? (p.uniform_len > 0 ? p.uniform_len : __ldg(p.lens + i)) : 0;
      const bool valid = (w < p.nw_rows) && (t < len);
      const int64_t qrow = (int64_t)min(i, p.Bq - 1) * p.Tp + t;
      // this thread's slice of q_w is fetched before the wait: its latency hides under GEMM-2, and it serves both passes
      uint4 qreg[8];
      {
        const uint4* qp = reinterpret_cast<const uint4*>(p.q16 + qrow * p.D + grp * dq);
#pragma unroll
        for (int k = 0; k < 8; ++k) qreg[k] = (valid && k < nch) ? __ldg(qp + k) : make_uint4(0, 0, 0, 0);
      }
      const float nq = valid ? fmaxf(__ldg(p.qnorm + qrow), 1e-30f) : 1.f;
      mbar_wait_lazy(&bars[f3WuFull], n & 1);
      if (tid == 128 + kF3ThreadsA) TGFR_TRACE(n, 5);
      tc_fence_after();
      float dot = 0.f, n2 = 0.f, dot1 = 0.f, n21 = 0.f;
      uint32_t vb[2][8];                          // the next 8 columns load while the current ones are consumed
      // Wu layout: Wu_w leaves for the backward as fp16 planes [d / 8][word][8 halfs] in this same pass (a warp's store
      // covers 512 contiguous bytes; the tile is both a K-major and an MN-major no-swizzle UMMA operand, tc.cuh)
      uint4* wdst = nullptr;
      if constexpr (LAY == kLayWu)
        wdst = reinterpret_cast<uint4*>(p.sv_wu + (int64_t)u * p.nw_rows * p.D) + (int64_t)(grp * (dq >> 3)) * p.nw_rows +
               min(w, p.nw_rows - 1);
      const int64_t wstep = p.nw_rows;
      tmem_ld8(wu_col, vb[0]);
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        if (ch < nch) {
          tmem_ld_wait();
          if (ch + 1 < nch) tmem_ld8(wu_col + 8 * (ch + 1), vb[(ch + 1) & 1]);
          const uint32_t(&v)[8] = vb[ch & 1];
          const __half2* qh = reinterpret_cast<const __half2*>(&qreg[ch]);
          uint32_t o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 qf = __half22float2(qh[k]);
            const float w0 = __uint_as_float(v[2 * k]), w1 = __uint_as_float(v[2 * k + 1]);
            dot = fmaf(qf.x, w0, dot);
            dot1 = fmaf(qf.y, w1, dot1);
            n2 = fmaf(w0, w0, n2);
            n21 = fmaf(w1, w1, n21);
            if constexpr (LAY == kLayWu) o[k] = valid ? pack_half2(w0, w1) : 0u;   // padding words / missing captions: exact zeros
          }
          if constexpr (LAY == kLayWu)
            if (w < p.nw_rows) wdst[ch * wstep] = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      dot += dot1;
      n2 += n21;
      if constexpr (LAY != kLayRec) {
        tc_fence_before();
        mbar_arrive(&bars[f3WuEmpty]);
      }
      if (tid == 128 + kF3ThreadsA) TGFR_TRACE(n, 6);
      part_d[grp * 128 + w] = dot;
      part_n[grp * 128 + w] = n2;
      f3b_bar_sync();
      if (grp == 0) {
        float ex = 0.f, cs = 0.f, inw = 0.f;
        if (valid) {
          float dd = 0.f, nn = 0.f;
#pragma unroll
          for (int k = 0; k < kF3NB; ++k) {
            dd += part_d[k * 128 + w];
            nn += part_n[k * 128 + w];
          }
          const float nW = fmaxf(sqrtf(nn), 1e-30f);
          cs = dd / (nq * nW);
          ex = fast_exp2(p.k2 * cs);
          inw = 1.f / nW;
        }
        exs[w] = ex;
        if constexpr (LAY != kLayNone) {
          cosw[w] = cs;
          inww[w] = inw;
        }
        if constexpr (LAY == kLayRec) p.sv_inw[(int64_t)u * 128 + w] = inw;
      }
      f3b_bar_sync();
      if constexpr (LAY == kLayWu) {
        if (grp == 1) {
          // d sim[b,i] / d Wu_w = alpha_w q_w - beta_w Wu_w with p_w = softmax over the caption's words of g2 cos
          float al = 0.f, be = 0.f;
          if (valid) {
            float ssum = 0.f;
#pragma unroll
            for (int tt = 0; tt < TP; ++tt) ssum += exs[c * TP + tt];
            const float pw = p.g23 * exs[w] / ssum;
            al = pw * inww[w] / nq;
            be = pw * cosw[w] * inww[w] * inww[w];
          }
          p.sv_ab[(int64_t)u * 256 + w] = al;
          p.sv_ab[(int64_t)u * 256 + 128 + w] = be;
        }
        if (tid == 128 + kF3ThreadsA) TGFR_TRACE(n, 7);
      }
      if constexpr (LAY == kLayRec) {
        // second pass over Wu: V_w = kSV p_w (q_w / |q_w| - cos_w Wu_w / |Wu_w|) -> fp16 planes of the saved V tile
        float c1 = 0.f, c2 = 0.f;
        if (valid) {
          float ssum = 0.f;
#pragma unroll
          for (int tt = 0; tt < TP; ++tt) ssum += exs[c * TP + tt];
          const float pw = rec::kSV * exs[w] / ssum;
          c1 = pw / nq;
          c2 = pw * cosw[w] * inww[w];
        }
        uint4* vdst = reinterpret_cast<uint4*>(p.sv_v + (int64_t)u * p.nw_rows * p.D) +
                      (int64_t)(grp * (dq >> 3)) * p.nw_rows + min(w, p.nw_rows - 1);
        const int64_t vstep = p.nw_rows;
        tmem_ld8(wu_col, vb[0]);
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          if (ch < nch) {
            tmem_ld_wait();
            if (ch + 1 < nch) tmem_ld8(wu_col + 8 * (ch + 1), vb[(ch + 1) & 1]);
            const uint32_t(&v)[8] = vb[ch & 1];
            const __half2* qh = reinterpret_cast<const __half2*>(&qreg[ch]);
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 qf = __half22float2(qh[k]);
              // padding words / missing captions: exact zeros (their Wu rows come from unwritten E columns)
              o[k] = valid ? pack_half2(fmaf(-c2, __uint_as_float(v[2 * k]), c1 * qf.x),
                                        fmaf(-c2, __uint_as_float(v[2 * k + 1]), c1 * qf.y))
                           : 0u;
            }
            if (w < p.nw_rows) vdst[ch * vstep] = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
        tc_fence_before();
        mbar_arrive(&bars[f3WuEmpty]);
        if (tid == 128 + kF3ThreadsA) TGFR_TRACE(n, 7);
      }
      if (grp == 0 && w < p.nc) {
        const int ii = g * p.nc + w;
        if (ii < p.Bq) {
          float sacc = 0.f;
          for (int tt = 0; tt < TP; ++tt) sacc += exs[w * TP + tt];
          p.sim[(int64_t)b * p.Bq + ii] = range_bad(p.lens, p.Bq) ? __int_as_float(0x7fc00000) : p.g3 * logf(sacc);
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

// ---------------------------------------------------------------------------------------------
// backward (d ctx): recompute S / E / Wu, then
//   epi-2   dW^[w,:] = sigma * dcos_w (q^_w - cos_w w^_w)  -> fp16 -> shared memory (over the E tile)
//   GEMM-3  dE^[r,w] = <c_r, dW^_w>                          (same shape as GEMM-1, B = dW^)
//   epi-3   dS = A1 (dA1 - sum_t A1 dA1), dA1 = g1 E^ dE^, E^ = E / |Wu_w|   (thread-local), written as
//           fp16 A operands *in TMEM*, in place over the S / dE^ accumulators they were computed from
//   GEMM-5/6  dC[r, d] = sum_w dS[r,w] q_w[d] + E^[r,w] dW^_w[d]   A from TMEM (lanes = regions),
//           B = MN-major views of the resident Q tile and of the dW^ tile, accumulated in the TMEM
//           columns the packed operands freed; 64 features per block, two rounds of four blocks
//   drain   TMEM -> registers (x 1/sigma) -> swizzled fp32 staging -> TMA reduce-add into d ctx
// sigma is a per-unit power of two that keeps the fp16 gradient operands in the normal range.
// ---------------------------------------------------------------------------------------------
enum BarB { bCFull = 0, bQFull, bSFull0, bSFull1, bEFull, bWuFull, bDwFull, bDeFull0, bDeFull1, bDsFull, bDc1, bDc2, bDc3, bDr0, bDr1, bDr2, bDr3, bNum };

struct TcBwdParams {
  const float* qnorm;    // [Bq*Tp]
  const int* lens;       // [Bq]
  const float* gsim;     // [Bc, Bq]
  int Bc, Bq, Tp, R, Rp, D, nc, G, nw_rows, n_tiles, total_units;
  uint32_t c_panel, q_panel, e_panel, off_q, off_x, off_misc;
  float k1, k2, g1, g23;   // g23 = gamma2 * gamma3
};

__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}

// DQ = false: d ctx (GEMM-5/6 above).  DQ = true: d words -- epi-3 writes dS' = dS + ca_w E^ (the second term is the
// direct cosine gradient through Wu = E^T C) as fp16 into shared memory in the E layout, GEMM-4
// dQ[w,d] = sum_r dS'[r,w] c_r[d] reuses GEMM-2's descriptors, and the drain subtracts the direct term in q_w and
// reduce-adds into the padded [Bq*Tp, D] gradient (tm_dc is then the map of that buffer).
// This kernel recomputes everything from C and Q (no forward records).  It serves d words and, when the forward kept
// nothing, d ctx; with the forward's Wu tiles d ctx runs in wr_tc_bwd3_kernel below.
template <int TP, bool DQ>
__global__ void __launch_bounds__(kThreadsTC, 1)
wr_tc_bwd_kernel(const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ CUtensorMap tm_q,
                 const __grid_constant__ CUtensorMap tm_dc, const TcBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_c = smem;
  uint8_t* s_q = smem + p.off_q;
  uint8_t* s_x = smem + p.off_x;               // E (GEMM-2) then dW^ (GEMM-3 / GEMM-6); 1 KB of zeros follows it
  uint8_t* misc = smem + p.off_misc;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 128);
  float2* part = reinterpret_cast<float2*>(misc + 256);      // [128] half-1 partials, then (ca, cb)
  float* exs = reinterpret_cast<float*>(misc + 256 + 1024);  // [128]
  float* invnw = exs + 128;                                  // [128] 1/|Wu_w| (0 for padding words)
  float* cqs = invnw + 128;                                  // [128] DQ: coefficient of q_w in the direct term

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int u0 = (int)((int64_t)blockIdx.x * p.total_units / gridDim.x);
  const int u1 = (int)((int64_t)(blockIdx.x + 1) * p.total_units / gridDim.x);
  const int kchunks = p.D >> 6;

  if (tid == 0) {
    mbar_init(&bars[bCFull], 1);
    mbar_init(&bars[bQFull], 1);
    mbar_init(&bars[bSFull0], 1);
    mbar_init(&bars[bSFull1], 1);
    mbar_init(&bars[bEFull], kEpiThreads);
    mbar_init(&bars[bWuFull], 1);
    mbar_init(&bars[bDwFull], kEpiThreads);
    mbar_init(&bars[bDeFull0], 1);
    mbar_init(&bars[bDeFull1], 1);
    mbar_init(&bars[bDsFull], kEpiThreads);
    for (int k = bDc1; k <= bDc3; ++k) mbar_init(&bars[k], 1);
    for (int k = bDr0; k <= bDr3; ++k) mbar_init(&bars[k], kEpiThreads);
    fence_barrier_init();
    tma_prefetch_desc(&tm_c);
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_dc);
  }
  // Operand tiles are read a few rows beyond what TMA / the epilogue write (K and N padding of the
  // MMA shapes); those rows only ever multiply zeros, but they must hold finite values.
  for (uint32_t k = tid; k < (p.off_misc >> 4); k += kThreadsTC) reinterpret_cast<uint4*>(smem)[k] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
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
        if (n > 0) mbar_wait(&bars[bDr3], (n - 1) & 1);   // previous unit fully drained: Q / X / C are free
        if (b != prev_b) {
          mbar_arrive_expect_tx(&bars[bCFull], kchunks * p.c_panel);
          for (int kc = 0; kc < kchunks; ++kc) tma_load_3d(s_c + kc * p.c_panel, &tm_c, &bars[bCFull], kc * 64, 0, b);
          prev_b = b;
        }
        mbar_arrive_expect_tx(&bars[bQFull], kchunks * p.q_panel);
        for (int kc = 0; kc < kchunks; ++kc)
          tma_load_3d(s_q + kc * p.q_panel, &tm_q, &bars[bQFull], kc * 64, g * p.nw_rows, 0);
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc_f16(128, 128, false, false);   // S, dE^
      const uint32_t idesc2 = make_idesc_f16(128, p.D, true, true);     // Wu
      const uint32_t a_c = smem_u32(s_c), a_q = smem_u32(s_q), a_x = smem_u32(s_x);
      int prev_b = -1, n = 0, m = -1;
      for (int u = u0; u < u1; ++u, ++n) {
        const int b = u / p.G;
        if (b != prev_b) {
          ++m;
          mbar_wait(&bars[bCFull], m & 1);
          prev_b = b;
        }
        mbar_wait(&bars[bQFull], n & 1);
        TGFR_TRACE(n, 17);
        tc_fence_after();
        // GEMM-1: S_t = C_t . Q^T
        for (int t = 0; t < p.n_tiles; ++t) {
          for (int k16 = 0; k16 < (p.D >> 4); ++k16) {
            const uint64_t ad = make_smem_desc(a_c + (k16 >> 2) * p.c_panel + t * (128 * 128) + (k16 & 3) * 32, 16, 1024);
            const uint64_t bd = make_smem_desc(a_q + (k16 >> 2) * p.q_panel + (k16 & 3) * 32, 16, 1024);
            umma_ss(tmem + t * 128, ad, bd, idesc1, k16 > 0);
          }
          if (t == 0) umma_commit(&bars[bSFull0]);
        }
        umma_commit(&bars[bSFull1]);
        // GEMM-2: Wu = E^T . C
        mbar_wait(&bars[bEFull], n & 1);
        TGFR_TRACE(n, 18);
        tc_fence_after();
        for (int j = 0; j < (p.Rp >> 4); ++j) {
          const uint64_t ad = make_smem_desc(a_x + j * 2048, p.e_panel, 1024);
          const uint64_t bd = make_smem_desc(a_c + j * 2048, p.c_panel, 1024);
          umma_ss(tmem + 256, ad, bd, idesc2, j > 0);
        }
        umma_commit(&bars[bWuFull]);
        // GEMM-3: dE^_t = C_t . dW^^T
        mbar_wait(&bars[bDwFull], n & 1);
        TGFR_TRACE(n, 19);
        tc_fence_after();
        for (int t = 0; t < p.n_tiles; ++t) {
          for (int k16 = 0; k16 < (p.D >> 4); ++k16) {
            const uint64_t ad = make_smem_desc(a_c + (k16 >> 2) * p.c_panel + t * (128 * 128) + (k16 & 3) * 32, 16, 1024);
            const uint64_t bd = make_smem_desc(a_x + (k16 >> 2) * p.q_panel + (k16 & 3) * 32, 16, 1024);
            umma_ss(tmem + 256 + t * 128, ad, bd, idesc1, k16 > 0);
          }
          if (t == 0) umma_commit(&bars[bDeFull0]);
        }
        umma_commit(&bars[bDeFull1]);
        if constexpr (DQ) {
          // GEMM-4: dQ[w,d] = sum_r dS'[r,w] c_r[d]  (A = the tile epi-3 wrote, MN-major; B = C, MN-major)
          mbar_wait(&bars[bDsFull], n & 1);
          TGFR_TRACE(n, 20);
          tc_fence_after();
          for (int j = 0; j < (p.Rp >> 4); ++j) {
            const uint64_t ad = make_smem_desc(a_x + j * 2048, p.e_panel, 1024);
            const uint64_t bd = make_smem_desc(a_c + j * 2048, p.c_panel, 1024);
            umma_ss(tmem + 256, ad, bd, idesc2, j > 0);
          }
          umma_commit(&bars[bDc1]);
          TGFR_TRACE(n, 21);
        } else {
          // GEMM-5/6: four rounds (tile, feature half) of 128-column dC blocks ping-ponging between the two
          // 128-column holes the packed operands leave free: [64,192) for tile 0, [320,448) for tile 1
          mbar_wait(&bars[bDsFull], n & 1);
          TGFR_TRACE(n, 20);
          tc_fence_after();
          for (int rd = 0; rd < 4; ++rd) {
            const int t = rd & 1, half = rd >> 1;
            if (rd >= 2) {                                   // the hole's previous block has been drained
              mbar_wait(&bars[bDr0 + rd - 2], n & 1);
              tc_fence_after();
            }
            const int ncols = min(p.D - half * 128, 128);
            if (t < p.n_tiles && ncols > 0) {
              const uint32_t idesc5 = make_idesc_f16(128, ncols, false, true);   // A in TMEM, B MN-major
              const uint32_t dcol = tmem + (t ? 320 : 64);
              for (int k16 = 0; k16 < 8; ++k16) {
                const uint64_t bx = make_smem_desc(a_x + 2 * half * p.q_panel + k16 * 2048, p.q_panel, 1024);
                umma_ts(dcol, tmem + 256 + t * 192 + 8 * k16, bx, idesc5, k16 > 0);      // E^ . dW^
              }
              for (int k16 = 0; k16 < 8; ++k16) {
                const uint64_t bq = make_smem_desc(a_q + 2 * half * p.q_panel + k16 * 2048, p.q_panel, 1024);
                umma_ts(dcol, tmem + t * 192 + 8 * k16, bq, idesc5, true);                // dS . Q
              }
            }
            if (rd >= 1) {
              umma_commit(&bars[bDc1 + rd - 1]);
              TGFR_TRACE(n, 20 + rd);
            }
          }
        }
        // the next unit's GEMM-1 overwrites the accumulator holes: wait until they are drained
        mbar_wait(&bars[bDr3], n & 1);
        TGFR_TRACE(n, 23);
      }
    }
  } else {
    // ======================================= epilogue =======================================
    const int tile = (warp - 2) >> 2;
    const int quarter_w = warp & 3;
    const int lrow = quarter_w * 32 + lane;
    const uint32_t t_lane = (uint32_t)(quarter_w * 32) << 16;
    const int r = tile * 128 + lrow;
    const bool warp_has_rows = tile < p.n_tiles && (tile * 128 + quarter_w * 32) < p.Rp;
    const int dhalf = p.D >> 1;
    // drain boxes (4 KB each): lane quarters 0-2 own two, quarter 3 (half the rows when R <= 224) one -- 28 KB per
    // warp group, which is what operand panels 0/1 leave free while rounds 2/3 still read panels 2/3
    uint8_t* const stage = (tile == 0 ? s_q : s_x + 1024) + quarter_w * 8192;
    const int nbuf = quarter_w == 3 ? 1 : 2;
    int cur = 0;
    int n = 0;
    for (int u = u0; u < u1; ++u, ++n) {
      const int b = u / p.G, g = u - b * p.G;
      // per-unit power-of-two scale from the largest |d loss / d sim| of the unit's captions
      float gmax = 0.f;
      for (int c = 0; c < p.nc; ++c) {
        const int i = g * p.nc + c;
        if (i < p.Bq) gmax = fmaxf(gmax, fabsf(__ldg(p.gsim + (int64_t)b * p.Bq + i)));
      }
      const float sigma = (gmax > 0.f) ? exp2f(floorf(log2f(4096.f / (p.g23 * gmax)))) : 1.f;
      const float inv_sigma = 1.f / sigma;

      // ---------------- epi-1: word softmax, E -> shared memory ----------------
      mbar_wait(&bars[tile == 0 ? bSFull0 : bSFull1], n & 1);
      if (tid == 64) TGFR_TRACE(n, 2);
      tc_fence_after();
      if (warp_has_rows) {
        const bool live_row = r < p.R;
        for (int c = 0; c < p.nc; ++c) {
          const int i = g * p.nc + c;
          const int len = (i < p.Bq) ? __ldg(p.lens + i) : 0;      // missing captions: zero columns
          uint32_t v[TP];
          const uint32_t col = tmem + t_lane + tile * 128 + c * TP;
#pragma unroll
          for (int j = 0; j < TP / 8; ++j) tmem_ld8(col + 8 * j, v + 8 * j);
          tmem_ld_wait();
          float e[TP];
          // four independent max / sum chains: with two warps per scheduler the dependent chain is the cost
          float mxp[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
          for (int t = 0; t < TP; ++t) {
            e[t] = (t < len) ? __uint_as_float(v[t]) : -INFINITY;
            mxp[t & 3] = fmaxf(mxp[t & 3], e[t]);
          }
          const float mx = fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3]));
          const float nmx = -mx * kLog2e;
          float sump[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int t = 0; t < TP; ++t) {
            e[t] = fast_exp2(fmaf(e[t], kLog2e, nmx));
            sump[t & 3] += e[t];
          }
          const float sum = (sump[0] + sump[1]) + (sump[2] + sump[3]);
          const float inv = (len > 0 && live_row) ? 1.f / sum : 0.f, nk1 = -p.k1;
          uint32_t pa[TP / 2], pe[TP / 2];
#pragma unroll
          for (int t = 0; t < TP; t += 2) {
            const float a0 = e[t] * inv, a1 = e[t + 1] * inv;
            e[t] = (t < len && live_row) ? fast_exp2(fmaf(a0, p.k1, nk1)) : 0.f;
            e[t + 1] = (t + 1 < len && live_row) ? fast_exp2(fmaf(a1, p.k1, nk1)) : 0.f;
            pa[t >> 1] = pack_half2(a0, a1);
            pe[t >> 1] = pack_half2(e[t], e[t + 1]);
          }
          // A1 and E (fp16) replace the caption's score columns: epi-3 reads them back instead of recomputing
#pragma unroll
          for (int j = 0; j < TP / 8; ++j) {
            tmem_st4(col + 4 * j, pa[4 * j], pa[4 * j + 1], pa[4 * j + 2], pa[4 * j + 3]);
            tmem_st4(col + TP / 2 + 4 * j, pe[4 * j], pe[4 * j + 1], pe[4 * j + 2], pe[4 * j + 3]);
          }
          if (r < p.Rp) {
#pragma unroll
            for (int j = 0; j < TP / 8; ++j) {
              const int w0 = c * TP + 8 * j;
              *reinterpret_cast<uint4*>(s_x + (w0 >> 6) * p.e_panel + sw128_offset(r, (w0 & 63) >> 3)) =
                  make_uint4(pe[4 * j], pe[4 * j + 1], pe[4 * j + 2], pe[4 * j + 3]);
            }
          }
        }
        tmem_st_wait();
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bars[bEFull]);
      if (tid == 64) TGFR_TRACE(n, 3);

      // ---------------- epi-2: cosine, softmax over words, dW^ -> shared memory ----------------
      mbar_wait(&bars[bWuFull], n & 1);
      if (tid == 64) TGFR_TRACE(n, 4);
      tc_fence_after();
      const int w = lrow;
      const int cw = w / TP, tw = w - cw * TP;
      const int iw = g * p.nc + cw;
      const bool valid = (w < p.nw_rows) && (iw < p.Bq) && (tw < __ldg(p.lens + min(iw, p.Bq - 1)));
      const int64_t qrow = (int64_t)min(iw, p.Bq - 1) * p.Tp + tw;
      float dot = 0.f, n2 = 0.f;
      {
        float dp[4] = {0.f, 0.f, 0.f, 0.f}, np[4] = {0.f, 0.f, 0.f, 0.f};
        for (int ch = 0; ch < (dhalf >> 5); ++ch) {
          uint32_t v[32];
          tmem_ld32(tmem + t_lane + 256 + tile * dhalf + 32 * ch, v);
          tmem_ld_wait();
          const int d0 = tile * dhalf + 32 * ch;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const uint4 qv = *reinterpret_cast<const uint4*>(s_q + (d0 >> 6) * p.q_panel +
                                                             sw128_offset(w, ((d0 & 63) >> 3) + cc));
            const __half2* qh = reinterpret_cast<const __half2*>(&qv);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 qf = __half22float2(qh[k]);
              const float w0 = __uint_as_float(v[8 * cc + 2 * k]), w1 = __uint_as_float(v[8 * cc + 2 * k + 1]);
              dp[(2 * k) & 3] = fmaf(qf.x, w0, dp[(2 * k) & 3]);
              dp[(2 * k + 1) & 3] = fmaf(qf.y, w1, dp[(2 * k + 1) & 3]);
              np[(2 * k) & 3] = fmaf(w0, w0, np[(2 * k) & 3]);
              np[(2 * k + 1) & 3] = fmaf(w1, w1, np[(2 * k + 1) & 3]);
            }
          }
        }
        dot = (dp[0] + dp[1]) + (dp[2] + dp[3]);
        n2 = (np[0] + np[1]) + (np[2] + np[3]);
      }
      if (tile == 1) part[w] = make_float2(dot, n2);
      epi_bar_sync();
      float cosv = 0.f, nW = 1.f, nq = 1.f;
      if (tile == 0) {
        float ex = 0.f;
        if (valid) {
          const float2 o = part[w];
          dot += o.x;
          n2 += o.y;
          nW = fmaxf(sqrtf(n2), 1e-30f);
          nq = fmaxf(__ldg(p.qnorm + qrow), 1e-30f);
          cosv = dot / (nq * nW);
          ex = fast_exp2(p.k2 * cosv);
        }
        exs[w] = ex;
      }
      epi_bar_sync();
      if (tile == 0) {
        float ca = 0.f, cb = 0.f, inw = 0.f;
        if (valid) {
          float ssum = 0.f;
          for (int tt = 0; tt < TP; ++tt) ssum += exs[cw * TP + tt];
          const float dcos = __ldg(p.gsim + (int64_t)b * p.Bq + iw) * p.g23 * (exs[w] / ssum) * sigma;
          ca = dcos / nq;                 // multiplies q_w
          cb = dcos * cosv / nW;          // multiplies Wu_w
          inw = 1.f / nW;
        }
        part[w] = make_float2(ca, cb);
        invnw[w] = inw;
        if constexpr (DQ) cqs[w] = valid ? ca * cosv / nq : 0.f;
      }
      epi_bar_sync();
      {
        const float2 cab = part[w];
        for (int ch = 0; ch < (dhalf >> 5); ++ch) {
          uint32_t v[32];
          tmem_ld32(tmem + t_lane + 256 + tile * dhalf + 32 * ch, v);
          tmem_ld_wait();
          const int d0 = tile * dhalf + 32 * ch;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const uint32_t off = (d0 >> 6) * p.q_panel + sw128_offset(w, ((d0 & 63) >> 3) + cc);
            const uint4 qv = *reinterpret_cast<const uint4*>(s_q + off);
            const __half2* qh = reinterpret_cast<const __half2*>(&qv);
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 qf = __half22float2(qh[k]);
              const float w0 = __uint_as_float(v[8 * cc + 2 * k]), w1 = __uint_as_float(v[8 * cc + 2 * k + 1]);
              o[k] = valid ? pack_half2(cab.x * qf.x - cab.y * w0, cab.x * qf.y - cab.y * w1) : 0u;
            }
            if (w < p.nw_rows) *reinterpret_cast<uint4*>(s_x + off) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bars[bDwFull]);
      if (tid == 64) TGFR_TRACE(n, 5);

      // ---------------- epi-3: dS and E^ as fp16 A operands, in place in TMEM ----------------
      // (DQ: the tile written below is GEMM-3's B operand, so every GEMM-3 MMA must have retired)
      mbar_wait(&bars[(tile == 0 && !DQ) ? bDeFull0 : bDeFull1], n & 1);
      if (tid == 64) TGFR_TRACE(n, 6);
      tc_fence_after();
      if (warp_has_rows) {
        const uint32_t s_base = tmem + t_lane + tile * 128;
        const uint32_t e_base = tmem + t_lane + 256 + tile * 128;
        // packed fp16 operands: tile 0 in the lower half of its accumulators (captions in ascending order),
        // tile 1 in the upper half (descending order), so that a thread never overwrites a column it has
        // not read yet and the free halves [64,192) and [320,448) are contiguous
        const uint32_t s_pack = s_base + tile * 64, e_pack = e_base + tile * 64;
        const bool live_row = r < p.R;
        for (int cc = 0; cc < p.nc; ++cc) {
          const int c = tile ? p.nc - 1 - cc : cc;
          const int i = g * p.nc + c;
          const int len = (i < p.Bq) ? __ldg(p.lens + i) : 0;
          (void)len;
          uint32_t vs[TP], vd[TP];
#pragma unroll
          for (int j = 0; j < TP / 8; ++j) {
            tmem_ld8(s_base + c * TP + 8 * j, vs + 8 * j);       // [A1 fp16 x TP | E fp16 x TP]
            tmem_ld8(e_base + c * TP + 8 * j, vd + 8 * j);       // dE^ fp32
          }
          tmem_ld_wait();
          float a1[TP], eh[TP], da[TP];
          float innerp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int t = 0; t < TP; t += 2) {
            const float2 af = __half22float2(*reinterpret_cast<const __half2*>(&vs[t >> 1]));
            const float2 ef = __half22float2(*reinterpret_cast<const __half2*>(&vs[TP / 2 + (t >> 1)]));
            const float2 nw = *reinterpret_cast<const float2*>(invnw + c * TP + t);   // 0 for padding words
            a1[t] = af.x;
            a1[t + 1] = af.y;
            eh[t] = ef.x * nw.x;
            eh[t + 1] = ef.y * nw.y;
            da[t] = p.g1 * eh[t] * __uint_as_float(vd[t]);
            da[t + 1] = p.g1 * eh[t + 1] * __uint_as_float(vd[t + 1]);
            innerp[t & 2] = fmaf(a1[t], da[t], innerp[t & 2]);
            innerp[(t & 2) + 1] = fmaf(a1[t + 1], da[t + 1], innerp[(t & 2) + 1]);
          }
          const float inner = (innerp[0] + innerp[1]) + (innerp[2] + innerp[3]);
          if constexpr (DQ) {
            if (r < p.Rp) {
#pragma unroll
              for (int j = 0; j < TP / 8; ++j) {
                uint32_t ds[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const int t0 = 8 * j + 2 * k;
                  const float ca0 = part[c * TP + t0].x, ca1 = part[c * TP + t0 + 1].x;
                  const float e0 = live_row ? eh[t0] : 0.f, e1 = live_row ? eh[t0 + 1] : 0.f;
                  ds[k] = pack_half2(fmaf(ca0, e0, a1[t0] * (da[t0] - inner)), fmaf(ca1, e1, a1[t0 + 1] * (da[t0 + 1] - inner)));
                }
                const int w0 = c * TP + 8 * j;
                *reinterpret_cast<uint4*>(s_x + (w0 >> 6) * p.e_panel + sw128_offset(r, (w0 & 63) >> 3)) =
                    make_uint4(ds[0], ds[1], ds[2], ds[3]);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < TP / 8; ++j) {
              uint32_t ds[4], ee[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int t0 = 8 * j + 2 * k;
                ds[k] = pack_half2(a1[t0] * (da[t0] - inner), a1[t0 + 1] * (da[t0 + 1] - inner));
                ee[k] = pack_half2(eh[t0], eh[t0 + 1]);
              }
              tmem_st4(s_pack + ((c * TP) >> 1) + 4 * j, ds[0], ds[1], ds[2], ds[3]);
              tmem_st4(e_pack + ((c * TP) >> 1) + 4 * j, ee[0], ee[1], ee[2], ee[3]);
            }
          }
        }
        if constexpr (!DQ) {
          // zero the K padding (words nw..127) of both operands
          for (int col = p.nw_rows >> 1; col < 64; col += 4) {
            tmem_st4(s_pack + col, 0u, 0u, 0u, 0u);
            tmem_st4(e_pack + col, 0u, 0u, 0u, 0u);
          }
        }
        tmem_st_wait();
      }
      if constexpr (DQ) fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bars[bDsFull]);
      if (tid == 64) TGFR_TRACE(n, 7);

      if constexpr (DQ) {
        // ---------------- drain (DQ): (dQ block - cq_w q_w) / sigma -> staging ring -> TMA reduce-add ----------------
        mbar_wait(&bars[bDc1], n & 1);
        if (tid == 64) TGFR_TRACE(n, 8);
        tc_fence_after();
        if (gmax > 0.f && quarter_w * 32 < p.nw_rows) {
          // the dS' tile is dead once GEMM-4 has retired: two 4 KB boxes per warp, one for the last warp (60 KB)
          uint8_t* const stage_q = s_x + 1024 + (warp - 2) * 8192;
          const int nbq = warp == 9 ? 1 : 2;
          int curq = 0;
          const bool row_ok = lrow < p.nw_rows;                          // lanes beyond the group hold no words
          const float cq = cqs[lrow];
          const int nchq = dhalf >> 5;
#pragma unroll 1
          for (int ch = 0; ch < nchq; ++ch) {
            const int d0 = tile * dhalf + 32 * ch;
            uint32_t v[32];
            tmem_ld32(tmem + t_lane + 256 + d0, v);
            tmem_ld_wait();
            uint8_t* const bufq = stage_q + curq * 4096;
            if (lane == 0) {
              if (nbq == 2) tma_wait_group_read<1>();
              else tma_wait_group_read<0>();
            }
            __syncwarp();
#pragma unroll
            for (int h8 = 0; h8 < 4; ++h8) {
              const int dd = d0 + 8 * h8;
              const uint4 qv = *reinterpret_cast<const uint4*>(s_q + (dd >> 6) * p.q_panel + sw128_offset(lrow, (dd & 63) >> 3));
              const __half2* qh = reinterpret_cast<const __half2*>(&qv);
              float o[8];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 qf = __half22float2(qh[k]);
                o[2 * k] = row_ok ? (__uint_as_float(v[8 * h8 + 2 * k]) - cq * qf.x) * inv_sigma : 0.f;
                o[2 * k + 1] = row_ok ? (__uint_as_float(v[8 * h8 + 2 * k + 1]) - cq * qf.y) * inv_sigma : 0.f;
              }
              *reinterpret_cast<float4*>(bufq + sw128_offset(lane, 2 * h8)) = make_float4(o[0], o[1], o[2], o[3]);
              *reinterpret_cast<float4*>(bufq + sw128_offset(lane, 2 * h8 + 1)) = make_float4(o[4], o[5], o[6], o[7]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_reduce_add_3d(&tm_dc, bufq, d0, g * p.nw_rows + quarter_w * 32, 0);
              tma_commit_group();
            }
            curq ^= nbq - 1;
          }
        }
        if (lane == 0) tma_wait_group_read<0>();
        __syncwarp();
        tc_fence_before();
        mbar_arrive(&bars[bDr3]);
        if (tid == 64) TGFR_TRACE(n, 12);
      } else {
        // ---------------- drain: dC blocks -> per-warp staging ring -> TMA reduce-add ----------------
        // Every warp drains 32 lanes x 64 columns of each block through its own 4 KB box (32 rows x 32 floats,
        // 128B swizzle) that overlays operand panels 0/1 of Q (warps 2-5) and dW^ (warps 6-9); those panels are
        // dead once rounds 0 and 1 have retired (bDc1).  The two warps of a scheduler alternate, which hides the
        // wait for the TMA engine to read a box back.
        for (int rd = 0; rd < 4; ++rd) {
          const int t = rd & 1, half = rd >> 1;
          if (rd != 1) {
            mbar_wait(&bars[rd == 0 ? bDc1 : bDc1 + rd - 1], n & 1);
            tc_fence_after();
          }
          if (rd == 0 && tid == 64) TGFR_TRACE(n, 8);
          const int col0 = half * 128 + tile * 64;             // first feature this warp drains
          const int row0 = t * 128 + quarter_w * 32;
          if (t < p.n_tiles && row0 < p.R && col0 < p.D && gmax > 0.f) {
            const uint32_t dcol = tmem + t_lane + (t ? 320 : 64) + tile * 64;
#pragma unroll 1
            for (int ch = 0; ch < 2; ++ch) {
              uint32_t v[32];
              tmem_ld32(dcol + 32 * ch, v);
              tmem_ld_wait();
              uint8_t* const buf = stage + cur * 4096;
              if (lane == 0) {                                  // the box written nbuf stores ago has been read
                if (nbuf == 2) tma_wait_group_read<1>();
                else tma_wait_group_read<0>();
              }
              __syncwarp();
#pragma unroll
              for (int c16 = 0; c16 < 8; ++c16) {
                float4 o;
                o.x = __uint_as_float(v[4 * c16 + 0]) * inv_sigma;
                o.y = __uint_as_float(v[4 * c16 + 1]) * inv_sigma;
                o.z = __uint_as_float(v[4 * c16 + 2]) * inv_sigma;
                o.w = __uint_as_float(v[4 * c16 + 3]) * inv_sigma;
                *reinterpret_cast<float4*>(buf + sw128_offset(lane, c16)) = o;
              }
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_reduce_add_3d(&tm_dc, buf, col0 + 32 * ch, row0, b);
                tma_commit_group();
              }
              cur ^= nbuf - 1;
            }
          }
          if (rd == 3) {                                        // Q / X are handed back to the producer
            if (lane == 0) tma_wait_group_read<0>();
            __syncwarp();
          }
          tc_fence_before();
          mbar_arrive(&bars[bDr0 + rd]);
          if (tid == 64) TGFR_TRACE(n, 9 + rd);
        }
      }
    }
    if (lane == 0) tma_wait_group<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// backward (d ctx) without attention records: wr_tc_bwd3_kernel
//
// The forward (SAVE) leaves, per unit (face b, caption group g), only the un-normalised word contexts Wu_w as fp16 planes
// and two scalars per word with   d sim[b,i] / d Wu_w = alpha_w q_w - beta_w Wu_w   (w a word of caption i).  Nothing of
// size B x B x T x R is ever written: the word softmax A1 and E = exp(g1 (A1 - 1)) are recomputed here from one extra
// score product, and the gradient of a pair stays linear in G[b,i] = d loss / d sim[b,i]:
//   GEMM-1'  S[r,w]  = <c_r, q_w>                               M = 128 regions (one tile), N = 128 words, K = D
//   GEMM-3'  X[r,w]  = <c_r, Wu_w>                              same shape, B = the Wu planes (K-major, no swizzle)
//   epi      A1 = softmax over the caption's words of S, E = exp(g1 (A1 - 1))        (2 exponentials per element)
//            ga_w = G sigma alpha_w, gb_w = G sigma beta_w;   dE = ga S - gb X;   dA1 = g1 E dE;
//            dS = A1 (dA1 - sum_t A1 dA1);   A_Q = dS + ga E,   A_W = -gb E   -> fp16 A operands, IN PLACE over the
//            caption's own S / X columns (a thread only overwrites columns it has read itself)
//   GEMM-5'/6'  dC[r,:] += sum_w A_Q[r,w] q_w + A_W[r,w] Wu_w   A from TMEM, B = MN-major views of the Q and Wu tiles;
//            a caption's Tp words take ceil(Tp / 16) K steps (the tail of the last step is zero in A)
// A work item is (face b, region tile t, caption group g): the 128 x D fp32 block of d ctx of one (b, t) stays in
// 256 TMEM columns while the CTA walks the caption groups and leaves once, by TMA reduce-add.  CTAs 2k and 2k+1 walk the
// same (b, g) units, one region tile each, so that the second reader of a Wu tile finds it in L2.
// sigma_b is a per-face power of two that keeps the fp16 operands in the normal range.
//
// TMEM columns: [0,256) d ctx block, [256,384) S -> A_Q, [384,512) X -> A_W.
// ---------------------------------------------------------------------------------------------
enum BarC { cCFull = 0, cVFull, cQFull, cSxFull, cOpsFull, cVFree, cAccDone, cDrained, cNum };
constexpr int kB3NE = 3;                                   // epilogue warps per TMEM lane quarter
constexpr int kB3EpiThreads = 128 * kB3NE;
constexpr int kB3Threads = 64 + kB3EpiThreads;
__device__ __forceinline__ void b3_bar_sync() { asm volatile("bar.sync 3, %0;" ::"n"(kB3EpiThreads) : "memory"); }

struct TcBwd3Params {
  const __half* wu;      // [total_units][D/8][nw_rows][8] fp16 Wu planes
  const float* ab;       // [total_units][2][128]  alpha_w, beta_w
  const float* gsim;     // [Bc, Bq]
  const int* lens;       // [Bq]
  int Bc, Bq, R, Rp, D, nc, G, nw_rows, n_tiles, total_units, uniform_len;
  uint32_t q_panel, off_q, off_v, off_misc;
  float k1, g1;          // k1 = g1 log2(e)
};

template <int TP>
__global__ void __launch_bounds__(kB3Threads, 1)
wr_tc_bwd3_kernel(const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ CUtensorMap tm_q,
                  const __grid_constant__ CUtensorMap tm_dc, const TcBwd3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_c = smem;                         // C tile: D/64 panels of 128 rows x 128 B (K-major A of GEMM-1' / 3')
  uint8_t* s_q = smem + p.off_q;               // Q tile: D/64 panels of nw_rows x 128 B
  uint8_t* s_v = smem + p.off_v;               // Wu tile: D/8 planes of nw_rows x 16 B (no swizzle); 1 KB of zeros follows
  uint8_t* misc = smem + p.off_misc;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 128);
  float* gab = reinterpret_cast<float*>(misc + 256);           // [2 (item parity)][2 (ga | gb)][128]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // two region tiles: CTAs (2k, 2k+1) share a range of units and take one tile each (the grid is even)
  const int t = (p.n_tiles == 2) ? (int)(blockIdx.x & 1) : 0;
  const int share = (p.n_tiles == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nshare = (p.n_tiles == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int u0 = (int)((int64_t)share * p.total_units / nshare);
  const int u1 = (int)((int64_t)(share + 1) * p.total_units / nshare);
  const int kchunks = p.D >> 6;
  constexpr uint32_t kCPanel = 128 * 128;
  constexpr int KS = (TP + 15) / 16;           // K steps of 16 words per caption in GEMM-5' / 6'

  if (tid == 0) {
    mbar_init(&bars[cCFull], 1);
    mbar_init(&bars[cVFull], 1);
    mbar_init(&bars[cQFull], 1);
    mbar_init(&bars[cSxFull], 1);
    mbar_init(&bars[cOpsFull], kB3EpiThreads);
    mbar_init(&bars[cVFree], 1);
    mbar_init(&bars[cAccDone], 1);
    mbar_init(&bars[cDrained], kB3EpiThreads);
    fence_barrier_init();
    tma_prefetch_desc(&tm_c);
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_dc);
  }
  // operand tiles are read a few rows beyond what TMA writes (K / N padding of the MMA shapes): keep them finite
  for (uint32_t k = tid; k < (p.off_misc >> 4); k += kB3Threads) reinterpret_cast<uint4*>(smem)[k] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int prev_b = -1, n = 0, m = -1;
      for (int u = u0; u < u1; ++u, ++n) {
        const int b = u / p.G, g = u - b * p.G;
        if (u + 1 < u1)                                             // the next Wu tile towards L2, one item ahead
          for (int kc = 0; kc < kchunks; ++kc)
            bulk_prefetch_l2(p.wu + (int64_t)(u + 1) * p.nw_rows * p.D + (int64_t)kc * (p.q_panel >> 1), p.q_panel);
        const bool new_face = b != prev_b;
        if (new_face) {
          if (n > 0) mbar_wait(&bars[cAccDone], (n - 1) & 1);       // every MMA of the previous face has retired ...
          if (m >= 0) mbar_wait(&bars[cDrained], m & 1);            // ... and its block is drained (the boxes overlay the tiles)
          ++m;
          mbar_arrive_expect_tx(&bars[cCFull], kchunks * kCPanel);
          for (int kc = 0; kc < kchunks; ++kc) tma_load_3d(s_c + kc * kCPanel, &tm_c, &bars[cCFull], kc * 64, t * 128, b);
          prev_b = b;
        }
        if (n > 0 && !new_face) mbar_wait(&bars[cVFree], (n - 1) & 1);   // GEMM-6' of the previous item has read the Wu tile
        mbar_arrive_expect_tx(&bars[cVFull], kchunks * p.q_panel);
        for (int kc = 0; kc < kchunks; ++kc)                        // the tile is one contiguous image: four bulk copies
          bulk_load(s_v + kc * p.q_panel, p.wu + (int64_t)u * p.nw_rows * p.D + (int64_t)kc * (p.q_panel >> 1), p.q_panel,
                    &bars[cVFull]);
        if (n > 0 && !new_face) mbar_wait(&bars[cAccDone], (n - 1) & 1); // GEMM-5' of the previous item has read the Q tile
        mbar_arrive_expect_tx(&bars[cQFull], kchunks * p.q_panel);
        for (int kc = 0; kc < kchunks; ++kc)
          tma_load_3d(s_q + kc * p.q_panel, &tm_q, &bars[cQFull], kc * 64, g * p.nw_rows, 0);
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc_f16(128, 128, false, false);    // S, X
      const uint32_t idesc5 = make_idesc_f16(128, p.D, false, true);     // d ctx block: A in TMEM, B MN-major
      const uint32_t a_c = smem_u32(s_c), a_q = smem_u32(s_q), a_v = smem_u32(s_v);
      const uint32_t v_plane = (uint32_t)p.nw_rows * 16u;                // bytes between consecutive 8-feature planes of Wu
      int prev_b = -1, n = 0, m = -1;
      for (int u = u0; u < u1; ++u, ++n) {
        const int b = u / p.G;
        const bool first = b != prev_b;
        if (first) {
          ++m;
          mbar_wait(&bars[cCFull], m & 1);
          prev_b = b;
        }
        mbar_wait(&bars[cVFull], n & 1);
        TGFR_TRACE(n, 17);
        tc_fence_after();
        for (int k16 = 0; k16 < (p.D >> 4); ++k16) {                     // GEMM-3': X = C_t Wu^T
          const uint64_t ad = make_smem_desc(a_c + (k16 >> 2) * kCPanel + (k16 & 3) * 32, 16, 1024);
          const uint64_t bd = make_smem_desc_ns(a_v + k16 * 2 * v_plane, v_plane, 128);   // K-major over d
          umma_ss(tmem + 384, ad, bd, idesc1, k16 > 0);
        }
        mbar_wait(&bars[cQFull], n & 1);
        tc_fence_after();
        for (int k16 = 0; k16 < (p.D >> 4); ++k16) {                     // GEMM-1': S = C_t Q^T
          const uint64_t ad = make_smem_desc(a_c + (k16 >> 2) * kCPanel + (k16 & 3) * 32, 16, 1024);
          const uint64_t bd = make_smem_desc(a_q + (k16 >> 2) * p.q_panel + (k16 & 3) * 32, 16, 1024);
          umma_ss(tmem + 256, ad, bd, idesc1, k16 > 0);
        }
        umma_commit(&bars[cSxFull]);
        mbar_wait(&bars[cOpsFull], n & 1);
        TGFR_TRACE(n, 18);
        tc_fence_after();
        bool acc = !first;
        for (int c = 0; c < p.nc; ++c)                                   // GEMM-6': A_W . Wu  (the first MMA of a (b, t) range overwrites)
          for (int j = 0; j < KS; ++j) {
            const uint32_t row0 = (uint32_t)(c * TP + 16 * j);
            const uint64_t bv = make_smem_desc_ns(a_v + row0 * 16u, 128, v_plane);        // MN-major over d, K = words
            umma_ts(tmem, tmem + 384 + c * TP + 8 * j, bv, idesc5, acc);
            acc = true;
          }
        umma_commit(&bars[cVFree]);
        for (int c = 0; c < p.nc; ++c)                                   // GEMM-5': A_Q . Q
          for (int j = 0; j < KS; ++j) {
            const uint32_t row0 = (uint32_t)(c * TP + 16 * j);
            const uint64_t bq = make_smem_desc(a_q + row0 * 128u, p.q_panel, 1024);
            umma_ts(tmem, tmem + 256 + c * TP + 8 * j, bq, idesc5, true);
          }
        umma_commit(&bars[cAccDone]);
        TGFR_TRACE(n, 19);
      }
    }
  } else {
    // ======================================= epilogue =======================================
    const int h = (warp - 2) >> 2;               // warp group 0..2: which captions in the epilogue (c = h, h + 3, ...);
                                                 // groups 0 / 1 also drain one half of D each
    const int quarter = warp & 3;                // TMEM lane quarter this warp may touch
    const int lrow = quarter * 32 + lane;
    const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
    int prev_b = -1, n = 0;
    float sigma = 1.f, inv_sigma = 1.f;
    for (int u = u0; u < u1; ++u, ++n) {
      const int b = u / p.G, g = u - b * p.G;
      const bool last = (u + 1 == u1) || ((u + 1) / p.G != b);
      if (b != prev_b) {                         // per-face power-of-two scale from the largest |d loss / d sim[b, :]|
        float gmax = 0.f;
        for (int i = lane; i < p.Bq; i += 32) gmax = fmaxf(gmax, fabsf(__ldg(p.gsim + (int64_t)b * p.Bq + i)));
        gmax = warp_max(gmax);
        // |A_Q| <~ G sigma alpha (1 + g1) with alpha <= g2 g3 / (|q| |Wu|): aim the largest entries at ~2^9
        sigma = (gmax > 0.f) ? exp2f(floorf(log2f(16.f / gmax))) : 1.f;
        inv_sigma = 1.f / sigma;
        prev_b = b;
      }
      float* const ga = gab + (n & 1) * 256;
      float* const gb = ga + 128;
      if (h == 0) {
        const int w = lrow, c = w / TP, i = g * p.nc + c;
        float al = 0.f, be = 0.f;
        if (w < p.nw_rows && i < p.Bq) {
          const float gs = __ldg(p.gsim + (int64_t)b * p.Bq + i) * sigma;
          al = __ldg(p.ab + (int64_t)u * 256 + w) * gs;
          be = __ldg(p.ab + (int64_t)u * 256 + 128 + w) * gs;
        }
        ga[w] = al;
        gb[w] = be;
      }
      const int r = t * 128 + lrow;
      const bool warp_has_rows = (t * 128 + quarter * 32) < p.Rp;
      const bool live_row = r < p.R;
      b3_bar_sync();                             // ga / gb visible; the other buffer is free for the next item
      mbar_wait(&bars[cSxFull], n & 1);          // GEMM-1' / 3' retired (and with them every MMA of the previous item)
      if (tid == 64) TGFR_TRACE(n, 2);
      tc_fence_after();
      if (warp_has_rows) {
        for (int c = h; c < p.nc; c += kB3NE) {
          const int i = g * p.nc + c;
          const uint32_t s_col = tmem + t_lane + 256 + c * TP, x_col = tmem + t_lane + 384 + c * TP;
          uint32_t pq[TP / 2], pw[TP / 2];
          if (i < p.Bq) {                    // warp-uniform: the TMEM loads / stores below are warp-collective
            const int len = p.uniform_len > 0 ? p.uniform_len : __ldg(p.lens + i);
            uint32_t vs[TP], vx[TP];
#pragma unroll
            for (int j = 0; j < TP / 8; ++j) {
              tmem_ld8(s_col + 8 * j, vs + 8 * j);
              tmem_ld8(x_col + 8 * j, vx + 8 * j);
            }
            tmem_ld_wait();
            float e[TP];
            float mxp[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
            for (int tt = 0; tt < TP; ++tt) {
              e[tt] = (tt < len) ? __uint_as_float(vs[tt]) : -INFINITY;
              mxp[tt & 3] = fmaxf(mxp[tt & 3], e[tt]);
            }
            const float mx = fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3]));
            const float nmx = -mx * kLog2e;
            float sump[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int tt = 0; tt < TP; ++tt) {
              e[tt] = fast_exp2(fmaf(e[tt], kLog2e, nmx));          // exp(S - max); 0 beyond the caption's length
              sump[tt & 3] += e[tt];
            }
            const float inv = 1.f / ((sump[0] + sump[1]) + (sump[2] + sump[3]));
            const float nk1 = -p.k1;
            // with u = ga E, v = gb E, d = u S - v X (= dE E / g1):  dA1 = g1 d,  inner = sum_t A1 dA1 = g1 sum_t A1 d,
            //   A_Q = A1 (dA1 - inner) + u = a1g d + u - a1g inner'   (a1g = g1 A1, inner' = sum_t A1 d),   A_W = -v
            float innerp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int tt = 0; tt < TP; tt += 2) {
              const float2 gaw = *reinterpret_cast<const float2*>(ga + c * TP + tt);       // 0 for padding words
              const float2 gbw = *reinterpret_cast<const float2*>(gb + c * TP + tt);
              const float a0 = e[tt] * inv, a1 = e[tt + 1] * inv;
              const float e0 = fast_exp2(fmaf(a0, p.k1, nk1)), e1 = fast_exp2(fmaf(a1, p.k1, nk1));   // E = exp(g1 (A1 - 1))
              const float u0 = e0 * gaw.x, u1 = e1 * gaw.y, v0 = e0 * gbw.x, v1 = e1 * gbw.y;
              const float d0 = fmaf(u0, __uint_as_float(vs[tt]), -(v0 * __uint_as_float(vx[tt])));
              const float d1 = fmaf(u1, __uint_as_float(vs[tt + 1]), -(v1 * __uint_as_float(vx[tt + 1])));
              innerp[tt & 2] = fmaf(a0, d0, innerp[tt & 2]);
              innerp[(tt & 2) + 1] = fmaf(a1, d1, innerp[(tt & 2) + 1]);
              const float g0 = a0 * p.g1, g1v = a1 * p.g1;
              e[tt] = g0;
              e[tt + 1] = g1v;
              vs[tt] = __float_as_uint(fmaf(g0, d0, u0));
              vs[tt + 1] = __float_as_uint(fmaf(g1v, d1, u1));
              pw[tt >> 1] = live_row ? pack_half2(-v0, -v1) : 0u;
            }
            const float ninner = -((innerp[0] + innerp[1]) + (innerp[2] + innerp[3]));
#pragma unroll
            for (int tt = 0; tt < TP; tt += 2) {
              const float q0 = fmaf(ninner, e[tt], __uint_as_float(vs[tt]));
              const float q1 = fmaf(ninner, e[tt + 1], __uint_as_float(vs[tt + 1]));
              pq[tt >> 1] = live_row ? pack_half2(q0, q1) : 0u;      // rows beyond R: zero operands
            }
          } else {
#pragma unroll
            for (int j = 0; j < TP / 2; ++j) pq[j] = pw[j] = 0u;     // missing captions: zero operands
          }
          // packed operands over the caption's own columns: TP / 2 columns of data, zero up to 8 KS
#pragma unroll
          for (int j = 0; j < 2 * KS; ++j) {
            uint32_t a4[4], b4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              a4[k] = (4 * j + k < TP / 2) ? pq[(4 * j + k) < TP / 2 ? 4 * j + k : 0] : 0u;
              b4[k] = (4 * j + k < TP / 2) ? pw[(4 * j + k) < TP / 2 ? 4 * j + k : 0] : 0u;
            }
            tmem_st4(s_col + 4 * j, a4[0], a4[1], a4[2], a4[3]);
            tmem_st4(x_col + 4 * j, b4[0], b4[1], b4[2], b4[3]);
          }
        }
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(&bars[cOpsFull]);
      if (tid == 64) TGFR_TRACE(n, 3);

      if (last) {
        // ---------------- drain: the (b, t) block / sigma -> per-warp 4 KB boxes -> TMA reduce-add into d ctx ----------------
        mbar_wait(&bars[cAccDone], n & 1);
        if (tid == 64) TGFR_TRACE(n, 8);
        tc_fence_after();
        const int row0 = t * 128 + quarter * 32;
        const int dhalf = p.D >> 1;
        if (row0 < p.R && h < 2) {
          uint8_t* const stage = smem + (warp - 2) * 8192;           // every operand tile is dead by now
          int cur = 0;
#pragma unroll 1
          for (int ch = 0; ch < (dhalf >> 5); ++ch) {
            const int col0 = h * dhalf + 32 * ch;
            uint32_t v[32];
            tmem_ld32(tmem + t_lane + col0, v);
            tmem_ld_wait();
            uint8_t* const buf = stage + cur * 4096;
            if (lane == 0) tma_wait_group_read<1>();                 // the box written two stores ago has been read
            __syncwarp();
#pragma unroll
            for (int c16 = 0; c16 < 8; ++c16) {
              float4 o;
              o.x = __uint_as_float(v[4 * c16 + 0]) * inv_sigma;
              o.y = __uint_as_float(v[4 * c16 + 1]) * inv_sigma;
              o.z = __uint_as_float(v[4 * c16 + 2]) * inv_sigma;
              o.w = __uint_as_float(v[4 * c16 + 3]) * inv_sigma;
              *reinterpret_cast<float4*>(buf + sw128_offset(lane, c16)) = o;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_reduce_add_3d(&tm_dc, buf, col0, row0, b);
              tma_commit_group();
            }
            cur ^= 1;
          }
          if (lane == 0) tma_wait_group_read<0>();
          __syncwarp();
        }
        tc_fence_before();
        mbar_arrive(&bars[cDrained]);
        if (tid == 64) TGFR_TRACE(n, 9);
      }
    }
    if (lane == 0) tma_wait_group<0>();
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
  size_t ws_c16, ws_q16, ws_qnorm, ws_lens, ws_dq, ws_attz, ws_total;
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
  pl->smem_bytes = pl->off_misc + 6144 + 1024;        // misc + alignment slack
  TGFR_REQUIRE(pl->smem_bytes <= 232448, "wordregion(tc): shared memory plan needs %u bytes", pl->smem_bytes);
  pl->ws_c16 = 0;
  pl->ws_q16 = align_up((size_t)Bc * R * D * 2, 256);
  pl->ws_qnorm = pl->ws_q16 + align_up((size_t)Bq * pl->Tp * D * 2, 256);
  pl->ws_lens = pl->ws_qnorm + align_up((size_t)Bq * pl->Tp * 4, 256);
  pl->ws_dq = pl->ws_lens + align_up((size_t)Bq * 4 + 64, 256);                  // lens + the two range-guard words                       // padded fp32 d words [Bq*Tp, D]
  pl->ws_attz = pl->ws_dq + align_up((size_t)Bq * pl->Tp * D * 4, 256);           // region sums of the diagonal attention maps
  pl->ws_total = pl->ws_attz + align_up((size_t)Bc * 32 * 4, 256);
  return TGFR_OK;
}

// =============================================================================================
// namespace rec: the RECORD-BASED variant of the saved state (TGFR_WORDREGION_SAVE=records, the default).
// The forward additionally writes, per unit, the word softmax and E = exp(g1 (A1 - 1)) as fp16 rows (A1 | E records)
// and V_w = kSV p_w (q^_w - cos_w w^_w) instead of Wu_w; the backward (wr_tc_bwd2_kernel) then needs no score product
// and no exponential.  It is the faster of the two backward paths at the price of HBM traffic and memory
// (B x B x T x R fp16 x 2); the record-free path (wr_tc_fwd3_kernel<.., true> + wr_tc_bwd3_kernel, TGFR_WORDREGION_SAVE=wu)
// writes nothing of that size.  Both are parity-tested (tests/test_gpu_tc.py, bwd_mode fixture).
// =============================================================================================
namespace rec {







// ---------------------------------------------------------------------------------------------
// backward (d ctx) from the forward's records: wr_tc_bwd2_kernel
//
// The forward (SAVE) leaves, per unit (face b, caption group g): A1 and E as fp16 rows, 1/|Wu_w|, and the tile
//   V_w = kSV p_w (q^_w - cos_w w^_w),    d sim[b,i] / d Wu_w = g2 g3 V_w / (kSV |Wu_w|)   (w a word of caption i),
// so the gradient of a pair is linear in G[b,i] = d loss / d sim[b,i] and needs no cosine / softmax work here:
//   GEMM-3   dE~[r,w] = <c_r, V_w>                               M = 128 regions (one tile), N = 128 words, K = D
//   epi-3    kappa_w = G[b,i] g2 g3 sigma_b / (kSV |Wu_w|);  Ek = E kappa;  dA1 = g1 Ek dE~;
//            dS = A1 (dA1 - sum_t A1 dA1)  -> (dS | Ek) as fp16 A operands in TMEM
//   GEMM-6/5 dC[r,:] += sum_w Ek[r,w] V_w + dS[r,w] q_w          A from TMEM, B = MN-major views of the V and Q tiles
// A work item is (face b, region tile t, caption group g): the 128 x D fp32 block of d ctx of one (b, t) stays in
// 256 TMEM columns while the CTA walks the caption groups and leaves once, by TMA reduce-add (a face's groups may
// be split between two CTAs).  CTAs 2k and 2k+1 walk the same (b, g) units, one region tile each, so that the
// second reader of a V tile finds it in L2.  Per item the CTA streams the Q and V tiles (bulk copies, the next V
// prefetched into L2 one item ahead) and the tile's A1 | E rows (coalesced loads); there is no GEMM-1, GEMM-2,
// exponential or per-unit drain.
// sigma_b is a per-face power of two that keeps the fp16 operands in the normal range.
//
// TMEM columns: [0,256) d ctx block, [256,384) dE~, [384,448) dS (fp16 pairs), [448,512) Ek (fp16 pairs).
// ---------------------------------------------------------------------------------------------
enum BarC { cCFull = 0, cVFull, cQFull, cDeFull, cOpsFull, cVFree, cAccDone, cDrained, cNum };

struct TcBwd2Params {
  const __half* v;       // [total_units][D/8][nw_rows][8] fp16 V tiles
  const uint8_t* rec;    // [total_units][nc][Tp/4 chunks][Rp] x 16 bytes: A1 x Tp | E x Tp as fp16
  const float* inw;      // [total_units][128]
  const float* gsim;     // [Bc, Bq]
  uint32_t rec_stride;   // bytes per unit
  int Bc, Bq, R, Rp, D, nc, G, nw_rows, n_tiles, total_units;
  uint32_t q_panel, off_q, off_v, off_misc;
  float g1, g23;
};

template <int TP>
__global__ void __launch_bounds__(kThreadsTC, 1)
wr_tc_bwd2_kernel(const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ CUtensorMap tm_q,
                  const __grid_constant__ CUtensorMap tm_dc, const TcBwd2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_c = smem;                         // C tile: D/64 panels of 128 rows x 128 B (K-major, GEMM-3's A)
  uint8_t* s_q = smem + p.off_q;               // Q tile: D/64 panels of nw_rows x 128 B
  uint8_t* s_v = smem + p.off_v;               // V tile: D/8 planes of nw_rows x 16 B (no swizzle); 1 KB of zeros follows
  uint8_t* misc = smem + p.off_misc;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 128);
  float* kap = reinterpret_cast<float*>(misc + 256);           // [2][128] kappa_w, double buffered over items

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // two region tiles: CTAs (2k, 2k+1) share a range of units and take one tile each (the grid is even)
  const int t = (p.n_tiles == 2) ? (int)(blockIdx.x & 1) : 0;
  const int share = (p.n_tiles == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nshare = (p.n_tiles == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int u0 = (int)((int64_t)share * p.total_units / nshare);
  const int u1 = (int)((int64_t)(share + 1) * p.total_units / nshare);
  const int kchunks = p.D >> 6;
  constexpr uint32_t kCPanel = 128 * 128;

  if (tid == 0) {
    mbar_init(&bars[cCFull], 1);
    mbar_init(&bars[cVFull], 1);
    mbar_init(&bars[cQFull], 1);
    mbar_init(&bars[cDeFull], 1);
    mbar_init(&bars[cOpsFull], kEpiThreads);
    mbar_init(&bars[cVFree], 1);
    mbar_init(&bars[cAccDone], 1);
    mbar_init(&bars[cDrained], kEpiThreads);
    fence_barrier_init();
    tma_prefetch_desc(&tm_c);
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_dc);
  }
  // operand tiles are read a few rows beyond what TMA writes (K / N padding of the MMA shapes): keep them finite
  for (uint32_t k = tid; k < (p.off_misc >> 4); k += kThreadsTC) reinterpret_cast<uint4*>(smem)[k] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int prev_b = -1, n = 0, m = -1;
      const uint32_t rec_rows = (uint32_t)min(p.Rp - t * 128, 128);
      for (int u = u0; u < u1; ++u, ++n) {
        const int b = u / p.G, g = u - b * p.G;
        // towards L2, one item ahead: the tile's A1 | E rows and the V tile (HBM traffic spreads over the item)
        for (int un = (n == 0 ? u : u + 1); un <= u + 1 && un < u1; ++un) {
          const uint8_t* base = p.rec + (int64_t)un * p.rec_stride;
          for (int cj = 0; cj < p.nc * (TP / 4); ++cj)
            bulk_prefetch_l2(base + ((int64_t)cj * p.Rp + t * 128) * 16, rec_rows * 16);
          if (un > u)
            for (int kc = 0; kc < kchunks; ++kc)
              bulk_prefetch_l2(p.v + (int64_t)un * p.nw_rows * p.D + (int64_t)kc * (p.q_panel >> 1), p.q_panel);
        }
        if (n > 0) mbar_wait(&bars[cVFree], (n - 1) & 1);           // GEMM-3 and GEMM-6 of the previous item retired
        if (b != prev_b) {
          if (m >= 0) mbar_wait(&bars[cDrained], m & 1);            // drain boxes overlay the operand tiles
          ++m;
          mbar_arrive_expect_tx(&bars[cCFull], kchunks * kCPanel);
          for (int kc = 0; kc < kchunks; ++kc) tma_load_3d(s_c + kc * kCPanel, &tm_c, &bars[cCFull], kc * 64, t * 128, b);
          prev_b = b;
        }
        mbar_arrive_expect_tx(&bars[cVFull], kchunks * p.q_panel);
        for (int kc = 0; kc < kchunks; ++kc)                        // the tile is one contiguous image: four bulk copies
          bulk_load(s_v + kc * p.q_panel, p.v + (int64_t)u * p.nw_rows * p.D + (int64_t)kc * (p.q_panel >> 1), p.q_panel,
                    &bars[cVFull]);
        if (n > 0) mbar_wait(&bars[cAccDone], (n - 1) & 1);         // GEMM-5 of the previous item retired
        mbar_arrive_expect_tx(&bars[cQFull], kchunks * p.q_panel);
        for (int kc = 0; kc < kchunks; ++kc)
          tma_load_3d(s_q + kc * p.q_panel, &tm_q, &bars[cQFull], kc * 64, g * p.nw_rows, 0);
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (lane == 0) {
      const uint32_t idesc3 = make_idesc_f16(128, 128, false, false);    // dE~
      const uint32_t idesc5 = make_idesc_f16(128, p.D, false, true);     // d ctx block: A in TMEM, B MN-major
      const uint32_t a_c = smem_u32(s_c), a_q = smem_u32(s_q), a_v = smem_u32(s_v);
      const uint32_t v_plane = (uint32_t)p.nw_rows * 16u;                // bytes between consecutive 8-feature planes of V
      int prev_b = -1, n = 0, m = -1;
      for (int u = u0; u < u1; ++u, ++n) {
        const int b = u / p.G;
        const bool first = b != prev_b;
        if (first) {
          ++m;
          mbar_wait(&bars[cCFull], m & 1);
          prev_b = b;
        }
        mbar_wait(&bars[cVFull], n & 1);
        TGFR_TRACE(n, 17);
        tc_fence_after();
        for (int k16 = 0; k16 < (p.D >> 4); ++k16) {                     // GEMM-3
          const uint64_t ad = make_smem_desc(a_c + (k16 >> 2) * kCPanel + (k16 & 3) * 32, 16, 1024);
          const uint64_t bd = make_smem_desc_ns(a_v + k16 * 2 * v_plane, v_plane, 128);   // K-major over d
          umma_ss(tmem + 256, ad, bd, idesc3, k16 > 0);
        }
        umma_commit(&bars[cDeFull]);
        mbar_wait(&bars[cOpsFull], n & 1);
        TGFR_TRACE(n, 18);
        tc_fence_after();
        for (int k16 = 0; k16 < 8; ++k16) {                              // GEMM-6: Ek . V  (the first MMA of a (b, t) range overwrites)
          const uint64_t bv = make_smem_desc_ns(a_v + k16 * 256, 128, v_plane);           // MN-major over d, K = words
          umma_ts(tmem, tmem + 448 + 8 * k16, bv, idesc5, !(first && k16 == 0));
        }
        umma_commit(&bars[cVFree]);
        mbar_wait(&bars[cQFull], n & 1);
        tc_fence_after();
        for (int k16 = 0; k16 < 8; ++k16) {                              // GEMM-5: dS . Q
          const uint64_t bq = make_smem_desc(a_q + k16 * 2048, p.q_panel, 1024);
          umma_ts(tmem, tmem + 384 + 8 * k16, bq, idesc5, true);
        }
        umma_commit(&bars[cAccDone]);
        TGFR_TRACE(n, 19);
      }
    }
  } else {
    // ======================================= epilogue =======================================
    const int h = (warp - 2) >> 2;               // warp group: which captions in epi-3, which half of D in the drain
    const int quarter = warp & 3;                // TMEM lane quarter this warp may touch
    const int lrow = quarter * 32 + lane;
    const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
    const int nc0 = (p.nc + 1) >> 1;
    const int c_lo = h ? nc0 : 0, c_hi = h ? p.nc : nc0;
    constexpr int NB = TP <= 24 ? 3 : 2;         // captions whose rows are fetched per batch (register budget)
    if (h == 0) {                                // K padding (words nw_rows..127) of both operands: zero, once
      for (int col = p.nw_rows >> 1; col < 64; col += 4) {
        tmem_st4(tmem + t_lane + 384 + col, 0u, 0u, 0u, 0u);
        tmem_st4(tmem + t_lane + 448 + col, 0u, 0u, 0u, 0u);
      }
      tmem_st_wait();
    }
    int prev_b = -1, n = 0;
    float sigma = 1.f, inv_sigma = 1.f;
    for (int u = u0; u < u1; ++u, ++n) {
      const int b = u / p.G, g = u - b * p.G;
      const bool last = (u + 1 == u1) || ((u + 1) / p.G != b);
      if (b != prev_b) {                         // per-face power-of-two scale from the largest |d loss / d sim[b, :]|
        float gmax = 0.f;
        for (int i = lane; i < p.Bq; i += 32) gmax = fmaxf(gmax, fabsf(__ldg(p.gsim + (int64_t)b * p.Bq + i)));
        gmax = warp_max(gmax);
        sigma = (gmax > 0.f) ? exp2f(floorf(log2f(4096.f / (p.g23 * gmax)))) : 1.f;
        inv_sigma = 1.f / sigma;
        prev_b = b;
      }
      float* const kp = kap + (n & 1) * 128;
      if (h == 0) {
        const int w = lrow, c = w / TP, i = g * p.nc + c;
        float k = 0.f;
        if (w < p.nw_rows && i < p.Bq)
          k = __ldg(p.inw + (int64_t)u * 128 + w) * __ldg(p.gsim + (int64_t)b * p.Bq + i) * (p.g23 * sigma * (1.f / kSV));
        kp[w] = k;
      }
      // this thread's A1 | E rows of its captions: issued before the waits so that the latency hides under GEMM-3
      const int r = t * 128 + lrow;
      const bool warp_has_rows = (t * 128 + quarter * 32) < p.Rp;
      const uint4* rsrc = reinterpret_cast<const uint4*>(p.rec + (int64_t)u * p.rec_stride) + min(r, p.Rp - 1);
      const int64_t cstride = (int64_t)p.Rp * (TP / 4);                  // uint4 per caption; chunk j is j * Rp further
      epi_bar_sync();                            // kappa visible; the other buffer is free for the next item
      bool waited = false;
      for (int c0 = c_lo; c0 < c_hi; c0 += NB) {
        uint4 rc[NB][TP / 4];
        if (warp_has_rows) {
#pragma unroll
          for (int cc = 0; cc < NB; ++cc)
#pragma unroll
            for (int j = 0; j < TP / 4; ++j) rc[cc][j] = __ldg(rsrc + min(c0 + cc, c_hi - 1) * cstride + (int64_t)j * p.Rp);
        }
        if (!waited) {
          mbar_wait(&bars[cDeFull], n & 1);      // GEMM-3 retired (and with it every MMA of the previous item)
          if (tid == 64) TGFR_TRACE(n, 2);
          tc_fence_after();
          waited = true;
        }
        if (warp_has_rows) {
#pragma unroll
          for (int cc = 0; cc < NB; ++cc) {
            const int c = c0 + cc;
            if (c < c_hi) {
              uint32_t vd[TP];
#pragma unroll
              for (int j = 0; j < TP / 8; ++j) tmem_ld8(tmem + t_lane + 256 + c * TP + 8 * j, vd + 8 * j);
              tmem_ld_wait();
              const uint32_t* ra = reinterpret_cast<const uint32_t*>(&rc[cc][0]);
              float a1[TP], ek[TP], da[TP];
              float innerp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int tt = 0; tt < TP; tt += 2) {
                const float2 af = __half22float2(*reinterpret_cast<const __half2*>(&ra[tt >> 1]));
                const float2 ef = __half22float2(*reinterpret_cast<const __half2*>(&ra[TP / 2 + (tt >> 1)]));
                const float2 kk = *reinterpret_cast<const float2*>(kp + c * TP + tt);
                a1[tt] = af.x;
                a1[tt + 1] = af.y;
                ek[tt] = ef.x * kk.x;
                ek[tt + 1] = ef.y * kk.y;
                da[tt] = p.g1 * ek[tt] * __uint_as_float(vd[tt]);
                da[tt + 1] = p.g1 * ek[tt + 1] * __uint_as_float(vd[tt + 1]);
                innerp[tt & 2] = fmaf(a1[tt], da[tt], innerp[tt & 2]);
                innerp[(tt & 2) + 1] = fmaf(a1[tt + 1], da[tt + 1], innerp[(tt & 2) + 1]);
              }
              const float inner = (innerp[0] + innerp[1]) + (innerp[2] + innerp[3]);
#pragma unroll
              for (int j = 0; j < TP / 8; ++j) {
                uint32_t ds[4], ee[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const int t0 = 8 * j + 2 * k;
                  ds[k] = pack_half2(a1[t0] * (da[t0] - inner), a1[t0 + 1] * (da[t0 + 1] - inner));
                  ee[k] = pack_half2(ek[t0], ek[t0 + 1]);
                }
                tmem_st4(tmem + t_lane + 384 + ((c * TP) >> 1) + 4 * j, ds[0], ds[1], ds[2], ds[3]);
                tmem_st4(tmem + t_lane + 448 + ((c * TP) >> 1) + 4 * j, ee[0], ee[1], ee[2], ee[3]);
              }
            }
          }
        }
      }
      if (!waited) {                             // a warp group without captions (nc = 1) still follows the barrier
        mbar_wait(&bars[cDeFull], n & 1);
        tc_fence_after();
      }
      if (warp_has_rows) tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bars[cOpsFull]);
      if (tid == 64) TGFR_TRACE(n, 3);

      if (last) {
        // ---------------- drain: the (b, t) block / sigma -> per-warp 4 KB boxes -> TMA reduce-add into d ctx ----------------
        mbar_wait(&bars[cAccDone], n & 1);
        if (tid == 64) TGFR_TRACE(n, 8);
        tc_fence_after();
        const int row0 = t * 128 + quarter * 32;
        const int dhalf = p.D >> 1;
        if (row0 < p.R) {
          uint8_t* const stage = smem + (warp - 2) * 8192;           // every operand tile is dead by now
          int cur = 0;
#pragma unroll 1
          for (int ch = 0; ch < (dhalf >> 5); ++ch) {
            const int col0 = h * dhalf + 32 * ch;
            uint32_t v[32];
            tmem_ld32(tmem + t_lane + col0, v);
            tmem_ld_wait();
            uint8_t* const buf = stage + cur * 4096;
            if (lane == 0) tma_wait_group_read<1>();                 // the box written two stores ago has been read
            __syncwarp();
#pragma unroll
            for (int c16 = 0; c16 < 8; ++c16) {
              float4 o;
              o.x = __uint_as_float(v[4 * c16 + 0]) * inv_sigma;
              o.y = __uint_as_float(v[4 * c16 + 1]) * inv_sigma;
              o.z = __uint_as_float(v[4 * c16 + 2]) * inv_sigma;
              o.w = __uint_as_float(v[4 * c16 + 3]) * inv_sigma;
              *reinterpret_cast<float4*>(buf + sw128_offset(lane, c16)) = o;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_reduce_add_3d(&tm_dc, buf, col0, row0, b);
              tma_commit_group();
            }
            cur ^= 1;
          }
          if (lane == 0) tma_wait_group_read<0>();
          __syncwarp();
        }
        tc_fence_before();
        mbar_arrive(&bars[cDrained]);
        if (tid == 64) TGFR_TRACE(n, 9);
      }
    }
    if (lane == 0) tma_wait_group<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}


// what the forward leaves for wr_tc_bwd2_kernel, in one caller-owned buffer:
//   [V tiles: total_units x (D/8 planes x nw_rows x 8 fp16)][records: total_units x (nc x Tp/4 chunks x Rp x 16 B)]
//   [1/|Wu|: total_units x 128 fp32][the fp16 operand copies of ctx and words, word norms, caption lengths]
// (the last group is what wr_tc_prep_kernel produces: the backward then skips its own conversion pass)
struct SavedLayout {
  size_t off_v, off_rec, off_inw, off_c16, off_q16, off_qnorm, off_lens, total;
  uint32_t rec_stride;
};
SavedLayout saved_layout(const TcPlan& pl, int Bc, int Bq, int R, int D) {
  SavedLayout L;
  const size_t units = (size_t)Bc * pl.G;
  L.rec_stride = (uint32_t)align_up((size_t)pl.nc * pl.Rp * pl.Tp * 4, 128);
  L.off_v = 0;
  L.off_rec = align_up(units * pl.nw_rows * D * 2, 1024);
  L.off_inw = L.off_rec + align_up(units * L.rec_stride, 1024);
  L.off_c16 = L.off_inw + align_up(units * 128 * 4, 256);
  L.off_q16 = L.off_c16 + align_up((size_t)Bc * R * D * 2, 256);
  L.off_qnorm = L.off_q16 + align_up((size_t)Bq * pl.Tp * D * 2, 256);
  L.off_lens = L.off_qnorm + align_up((size_t)Bq * pl.Tp * 4, 256);
  L.total = L.off_lens + align_up((size_t)Bq * 4 + 64, 256);
  return L;
}


}  // namespace rec

// what the forward leaves for wr_tc_bwd3_kernel, in one caller-owned buffer:
//   [Wu tiles: total_units x (D/8 planes x nw_rows x 8 fp16)][alpha | beta: total_units x 2 x 128 fp32]
//   [the fp16 operand copies of ctx and words, word norms, caption lengths]
// (the last group is what wr_tc_prep_kernel produces: the backward then skips its own conversion pass)
struct SavedLayout {
  size_t off_wu, off_ab, off_c16, off_q16, off_qnorm, off_lens, total;
};
SavedLayout saved_layout(const TcPlan& pl, int Bc, int Bq, int R, int D) {
  SavedLayout L;
  const size_t units = (size_t)Bc * pl.G;
  L.off_wu = 0;
  L.off_ab = align_up(units * pl.nw_rows * D * 2, 1024);
  L.off_c16 = L.off_ab + align_up(units * 256 * 4, 256);
  L.off_q16 = L.off_c16 + align_up((size_t)Bc * R * D * 2, 256);
  L.off_qnorm = L.off_q16 + align_up((size_t)Bq * pl.Tp * D * 2, 256);
  L.off_lens = L.off_qnorm + align_up((size_t)Bq * pl.Tp * 4, 256);
  L.total = L.off_lens + align_up((size_t)Bq * 4 + 64, 256);
  return L;
}

enum { kSavedNone = 0, kSavedRecords = 1, kSavedWu = 2 };
// the record layout is strictly larger than the Wu layout, so the size of the caller's buffer names the layout
inline int saved_mode_of(const void* saved, size_t saved_bytes, size_t wu_total, size_t rec_total) {
  if (saved == nullptr) return kSavedNone;
  if (saved_bytes >= rec_total) return kSavedRecords;
  if (saved_bytes >= wu_total) return kSavedWu;
  return kSavedNone;
}

// shared-memory plan of wr_tc_bwd2_kernel / wr_tc_bwd3_kernel: one 128-row C tile, the Q and V tiles, 1 KB of zeros, misc
struct TcBwd3Plan {
  uint32_t q_panel, off_q, off_v, off_misc, smem_bytes;
};
int make_bwd3_plan(const TcPlan& fp, int D, TcBwd3Plan* pl) {
  const int kch = D / 64;
  pl->q_panel = fp.q_panel;
  pl->off_q = kch * 128u * 128u;
  pl->off_v = pl->off_q + kch * pl->q_panel;
  pl->off_misc = pl->off_v + kch * pl->q_panel + 1024;
  if (pl->off_misc < 65536) pl->off_misc = 65536;               // the drain boxes (8 warps x 8 KB) overlay the tiles
  pl->smem_bytes = pl->off_misc + 4096 + 1024;                  // misc (barriers, ga | gb double buffer) + alignment slack
  TGFR_REQUIRE(pl->smem_bytes <= 232448, "wordregion(tc): backward shared memory plan needs %u bytes", pl->smem_bytes);
  return TGFR_OK;
}

struct TcBwdPlan {
  int Tp, Rp, c_rows, nc, G, nw_rows, n_tiles;
  uint32_t c_panel, q_panel, e_panel, off_q, off_x, off_misc, smem_bytes;
};

int make_bwd_plan(int Bq, int T, int R, int D, TcBwdPlan* pl) {
  TGFR_REQUIRE(D % 64 == 0 && D >= 64 && D <= 256, "wordregion(tc): D=%d must be 64, 128, 192 or 256", D);
  TGFR_REQUIRE(R >= 1 && R <= 256, "wordregion(tc): R=%d regions (max 256)", R);
  TGFR_REQUIRE(T >= 1 && T <= 32, "wordregion(tc): T=%d words (max 32)", T);
  const int kch = D / 64;
  pl->Tp = (T + 7) & ~7;
  pl->Rp = (R + 15) & ~15;
  pl->c_rows = (R + 7) & ~7;
  pl->n_tiles = (pl->Rp + 127) / 128;
  pl->c_panel = (uint32_t)pl->c_rows * 128u;
  pl->e_panel = (uint32_t)pl->Rp * 128u;
  for (int nc = 128 / pl->Tp; nc >= 1; --nc) {
    const uint32_t q_panel = (uint32_t)nc * pl->Tp * 128u;
    uint32_t q_bytes = kch * q_panel, x_bytes = kch * q_panel + 1024;
    if (q_bytes < 28672) q_bytes = 28672;                       // drain boxes (7 x 4 KB per warp group)
    if (x_bytes < 2 * pl->e_panel) x_bytes = 2 * pl->e_panel;
    if (x_bytes < 61440 + 1024) x_bytes = 61440 + 1024;         // DQ drain: 15 x 4 KB over the dead dS' tile
    // with more than two feature panels the boxes must fit in panels 0/1, which are dead while rounds 2/3 run
    if (kch > 2 && 2 * q_panel < 28672 + 1024) continue;
    const uint32_t off_q = kch * pl->c_panel, off_x = off_q + q_bytes, off_misc = off_x + x_bytes;
    const uint32_t total = off_misc + 4096 + 1024;
    if (total <= 232448) {
      pl->nc = nc; pl->nw_rows = nc * pl->Tp; pl->q_panel = q_panel;
      pl->off_q = off_q; pl->off_x = off_x; pl->off_misc = off_misc; pl->smem_bytes = total;
      pl->G = (Bq + nc - 1) / nc;
      return TGFR_OK;
    }
  }
  set_error("wordregion(tc): no caption grouping fits in shared memory (T=%d R=%d D=%d)", T, R, D);
  return TGFR_E_INVALID;
}

// dwords[i, t, :] = dq_pad[i * Tp + t, :]  (the padded rows t >= T are dropped)
__global__ void wr_tc_unpad_kernel(const float* __restrict__ dq_pad, float* __restrict__ dwords, int Bq, int T, int Tp, int D) {
  const int64_t n4 = (int64_t)Bq * T * (D >> 2);
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n4; k += (int64_t)gridDim.x * blockDim.x) {
    const int d4 = (int)(k % (D >> 2));
    const int64_t row = k / (D >> 2);
    const int i = (int)(row / T), t = (int)(row - (int64_t)i * T);
    reinterpret_cast<float4*>(dwords)[k] = reinterpret_cast<const float4*>(dq_pad)[((int64_t)i * Tp + t) * (D >> 2) + d4];
  }
}

}  // namespace

int wordregion_bwd_tc(const float* ctx, int64_t csb, int64_t csr, int64_t csd, const float* words, int64_t wsb,
                      int64_t wst, int64_t wsd, const int32_t* cap_lens, int Bc, int Bq, int T, int R, int D, float g1,
                      float g2, float g3, const float* gsim, float* dctx, float* dwords, void* ws, size_t ws_bytes,
                      const void* saved, size_t saved_bytes, cudaStream_t st) {
  TcPlan fp;
  if (int rc = make_plan(Bc, Bq, T, R, D, &fp)) return rc;     // workspace layout is shared with the forward
  const SavedLayout L = saved_layout(fp, Bc, Bq, R, D);
  const rec::SavedLayout LR = rec::saved_layout(fp, Bc, Bq, R, D);
  const int sv_mode = saved_mode_of(saved, saved_bytes, L.total, LR.total);
  const bool saved_ok = sv_mode != kSavedNone;
  TcBwdPlan pl{};
  pl.Tp = fp.Tp;
  if ((dctx && !saved_ok) || dwords) {                          // the recompute kernel has its own shared-memory plan
    if (int rc = make_bwd_plan(Bq, T, R, D, &pl)) return rc;
  }
  TGFR_REQUIRE(ws != nullptr && ws_bytes >= fp.ws_total, "wordregion(tc): workspace too small (%zu < %zu)", ws_bytes,
               fp.ws_total);
  TGFR_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "wordregion(tc): workspace must be 256-byte aligned");
  TGFR_REQUIRE(!dctx || (reinterpret_cast<uintptr_t>(dctx) & 15) == 0, "wordregion(tc): dctx must be 16-byte aligned");
  TGFR_REQUIRE(!dwords || (reinterpret_cast<uintptr_t>(dwords) & 15) == 0, "wordregion(tc): dwords must be 16-byte aligned");
  if (!dctx && !dwords) return TGFR_OK;
  uint8_t* base = reinterpret_cast<uint8_t*>(ws);
  __half* c16 = reinterpret_cast<__half*>(base + fp.ws_c16);
  __half* q16 = reinterpret_cast<__half*>(base + fp.ws_q16);
  float* qnorm = reinterpret_cast<float*>(base + fp.ws_qnorm);
  int* lens = reinterpret_cast<int*>(base + fp.ws_lens);
  float* dq_pad = reinterpret_cast<float*>(base + fp.ws_dq);

  const bool have_saved = saved_ok;
  if (have_saved) {
    // the forward kept its fp16 operand copies next to its saved state: no second conversion pass
    uint8_t* sv = const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(saved));
    const bool r = sv_mode == kSavedRecords;
    c16 = reinterpret_cast<__half*>(sv + (r ? LR.off_c16 : L.off_c16));
    q16 = reinterpret_cast<__half*>(sv + (r ? LR.off_q16 : L.off_q16));
    qnorm = reinterpret_cast<float*>(sv + (r ? LR.off_qnorm : L.off_qnorm));
    lens = reinterpret_cast<int*>(sv + (r ? LR.off_lens : L.off_lens));
  } else {
    const int64_t rows = (int64_t)Bc * R + (int64_t)Bq * pl.Tp;
    TGFR_CUDA_OK(cudaMemsetAsync(lens + Bq, 0, 2 * sizeof(int), st));
    wr_tc_prep_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(ctx, csb, csr, csd, words, wsb, wst, wsd, cap_lens, Bc,
                                                                 Bq, T, pl.Tp, R, D, c16, q16, qnorm, lens);
    TGFR_LAUNCH_OK();
  }

  CUtensorMap tm_c{}, tm_q{};
  if ((dctx && !saved_ok) || dwords) {                          // maps of the recompute kernel (its own caption grouping)
    if (int rc = make_tmap_3d(&tm_c, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, c16, D, R, Bc, 64, pl.c_rows, 1)) return rc;
    if (int rc = make_tmap_3d(&tm_q, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, q16, D, (uint64_t)Bq * pl.Tp, 1, 64, pl.nw_rows, 1))
      return rc;
  }

  TcBwdParams p{};
  p.qnorm = qnorm; p.lens = lens; p.gsim = gsim;
  p.Bc = Bc; p.Bq = Bq; p.Tp = pl.Tp; p.R = R; p.Rp = pl.Rp; p.D = D; p.nc = pl.nc; p.G = pl.G;
  p.nw_rows = pl.nw_rows; p.n_tiles = pl.n_tiles; p.total_units = Bc * pl.G;
  p.c_panel = pl.c_panel; p.q_panel = pl.q_panel; p.e_panel = pl.e_panel;
  p.off_q = pl.off_q; p.off_x = pl.off_x; p.off_misc = pl.off_misc;
  p.k1 = g1 * kLog2e; p.k2 = g2 * kLog2e; p.g1 = g1; p.g23 = g2 * g3;

  int dev = 0, sms = 0;
  TGFR_CUDA_OK(cudaGetDevice(&dev));
  TGFR_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int grid = p.total_units < sms ? p.total_units : sms;
  int grid_dbg = 0;
  if (const char* dg = getenv("TGFR_DEBUG_GRID")) grid_dbg = atoi(dg);   // profiling aid: fewer CTAs -> no L2 contention
  if (grid_dbg > 0 && grid_dbg < grid) grid = grid_dbg;

#define TGFR_LAUNCH_BWD(TPV, DQV, TM)                                                                             \
  case TPV: {                                                                                                     \
    static bool attr_done[64] = {};   /* once per instantiation and device, to the architectural maximum */       \
    bool& attr_set = attr_done[dev & 63];                                                                         \
    if (!attr_set) {                                                                                              \
      TGFR_CUDA_OK(cudaFuncSetAttribute(wr_tc_bwd_kernel<TPV, DQV>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                        232448));                                                                 \
      attr_set = true;                                                                                            \
    }                                                                                                             \
    wr_tc_bwd_kernel<TPV, DQV><<<grid, kThreadsTC, pl.smem_bytes, st>>>(tm_c, tm_q, TM, p);                        \
  } break;
#define TGFR_LAUNCH_BWD2(TPV)                                                                                     \
  case TPV: {                                                                                                     \
    static bool attr_done[64] = {};                                                                               \
    bool& attr_set = attr_done[dev & 63];                                                                         \
    if (!attr_set) {                                                                                              \
      TGFR_CUDA_OK(cudaFuncSetAttribute(rec::wr_tc_bwd2_kernel<TPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        232448));                                                                 \
      attr_set = true;                                                                                            \
    }                                                                                                             \
    rec::wr_tc_bwd2_kernel<TPV><<<grid2, kThreadsTC, pl2.smem_bytes, st>>>(tm_c2, tm_q2, tm_dc, p2);               \
  } break;
#define TGFR_LAUNCH_BWD3(TPV)                                                                                     \
  case TPV: {                                                                                                     \
    static bool attr_done[64] = {};                                                                               \
    bool& attr_set = attr_done[dev & 63];                                                                         \
    if (!attr_set) {                                                                                              \
      TGFR_CUDA_OK(cudaFuncSetAttribute(wr_tc_bwd3_kernel<TPV>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                        232448));                                                                 \
      attr_set = true;                                                                                            \
    }                                                                                                             \
    wr_tc_bwd3_kernel<TPV><<<grid2, kB3Threads, pl2.smem_bytes, st>>>(tm_c2, tm_q2, tm_dc, p2);                     \
  } break;
  if (dctx) {
    CUtensorMap tm_dc;
    TGFR_CUDA_OK(cudaMemsetAsync(dctx, 0, sizeof(float) * (size_t)Bc * R * D, st));
    if (int rc = make_tmap_3d(&tm_dc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dctx, D, R, Bc, 32, 32, 1)) return rc;
    if (have_saved) {
      // (b, tile, caption group) items, d ctx accumulated in TMEM -- from the forward's records (wr_tc_bwd2_kernel)
      // or from its Wu tiles and the S it recomputes (wr_tc_bwd3_kernel); both share one shared-memory plan
      TcBwd3Plan pl2;
      if (int rc = make_bwd3_plan(fp, D, &pl2)) return rc;
      const uint8_t* sv = reinterpret_cast<const uint8_t*>(saved);
      CUtensorMap tm_c2, tm_q2;
      if (int rc = make_tmap_3d(&tm_c2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, c16, D, R, Bc, 64, 128, 1)) return rc;
      if (int rc = make_tmap_3d(&tm_q2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, q16, D, (uint64_t)Bq * fp.Tp, 1, 64, fp.nw_rows, 1))
        return rc;
      // one CTA per SM; with two region tiles CTAs (2k, 2k+1) share a unit range, one tile each
      int shares = fp.n_tiles == 2 ? sms / 2 : sms;
      if (grid_dbg > 0 && grid_dbg < shares) shares = grid_dbg;
      if (shares > Bc * fp.G) shares = Bc * fp.G;
      const int grid2 = shares * fp.n_tiles;
      if (sv_mode == kSavedRecords) {
        rec::TcBwd2Params p2{};
        p2.v = reinterpret_cast<const __half*>(sv + LR.off_v);
        p2.rec = sv + LR.off_rec;
        p2.inw = reinterpret_cast<const float*>(sv + LR.off_inw);
        p2.gsim = gsim;
        p2.rec_stride = LR.rec_stride;
        p2.Bc = Bc; p2.Bq = Bq; p2.R = R; p2.Rp = fp.Rp; p2.D = D; p2.nc = fp.nc; p2.G = fp.G;
        p2.nw_rows = fp.nw_rows; p2.n_tiles = fp.n_tiles; p2.total_units = Bc * fp.G;
        p2.q_panel = pl2.q_panel; p2.off_q = pl2.off_q; p2.off_v = pl2.off_v; p2.off_misc = pl2.off_misc;
        p2.g1 = g1; p2.g23 = g2 * g3;
        switch (fp.Tp) {
          TGFR_LAUNCH_BWD2(8)
          TGFR_LAUNCH_BWD2(16)
          TGFR_LAUNCH_BWD2(24)
          TGFR_LAUNCH_BWD2(32)
          default:
            set_error("wordregion(tc): unsupported padded caption length %d", fp.Tp);
            return TGFR_E_INVALID;
        }
      } else {
      TcBwd3Params p2{};
      p2.wu = reinterpret_cast<const __half*>(sv + L.off_wu);
      p2.ab = reinterpret_cast<const float*>(sv + L.off_ab);
      p2.gsim = gsim;
      p2.lens = lens;
      p2.Bc = Bc; p2.Bq = Bq; p2.R = R; p2.Rp = fp.Rp; p2.D = D; p2.nc = fp.nc; p2.G = fp.G;
      p2.nw_rows = fp.nw_rows; p2.n_tiles = fp.n_tiles; p2.total_units = Bc * fp.G;
      p2.uniform_len = cap_lens ? 0 : T;
      p2.q_panel = pl2.q_panel; p2.off_q = pl2.off_q; p2.off_v = pl2.off_v; p2.off_misc = pl2.off_misc;
      p2.k1 = g1 * kLog2e; p2.g1 = g1;
      switch (fp.Tp) {
        TGFR_LAUNCH_BWD3(8)
        TGFR_LAUNCH_BWD3(16)
        TGFR_LAUNCH_BWD3(24)
        TGFR_LAUNCH_BWD3(32)
        default:
          set_error("wordregion(tc): unsupported padded caption length %d", fp.Tp);
          return TGFR_E_INVALID;
      }
      }
    } else {
      switch (pl.Tp) {
        TGFR_LAUNCH_BWD(8, false, tm_dc)
        TGFR_LAUNCH_BWD(16, false, tm_dc)
        TGFR_LAUNCH_BWD(24, false, tm_dc)
        TGFR_LAUNCH_BWD(32, false, tm_dc)
        default:
          set_error("wordregion(tc): unsupported padded caption length %d", pl.Tp);
          return TGFR_E_INVALID;
      }
    }
    TGFR_LAUNCH_OK();
  }
  if (dwords) {
    CUtensorMap tm_dq;
    TGFR_CUDA_OK(cudaMemsetAsync(dq_pad, 0, sizeof(float) * (size_t)Bq * pl.Tp * D, st));
    if (int rc = make_tmap_3d(&tm_dq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dq_pad, D, (uint64_t)Bq * pl.Tp, 1, 32, 32, 1))
      return rc;
    switch (pl.Tp) {
      TGFR_LAUNCH_BWD(8, true, tm_dq)
      TGFR_LAUNCH_BWD(16, true, tm_dq)
      TGFR_LAUNCH_BWD(24, true, tm_dq)
      TGFR_LAUNCH_BWD(32, true, tm_dq)
      default:
        set_error("wordregion(tc): unsupported padded caption length %d", pl.Tp);
        return TGFR_E_INVALID;
    }
    TGFR_LAUNCH_OK();
    const int64_t n4 = (int64_t)Bq * T * (D >> 2);
    wr_tc_unpad_kernel<<<(unsigned)((n4 + 255) / 256 < 2048 ? (n4 + 255) / 256 : 2048), 256, 0, st>>>(dq_pad, dwords, Bq, T,
                                                                                                        pl.Tp, D);
    TGFR_LAUNCH_OK();
  }
#undef TGFR_LAUNCH_BWD
#undef TGFR_LAUNCH_BWD3
#undef TGFR_LAUNCH_BWD2
  return TGFR_OK;
}

// does the recompute kernel (d words; d ctx without forward records) have a shared-memory plan for this shape?
// (T = 30 with R = 196, D = 256 has none: the C tile, two word tiles and the drain boxes exceed 227 KB)
bool wordregion_tc_recompute_ok(int Bq, int T, int R, int D) {
  TcBwdPlan pl;
  const bool ok = make_bwd_plan(Bq, T, R, D, &pl) == TGFR_OK;
  return ok;
}

int wordregion_tc_set_trace(void* dev_buf) {
  long long* p = reinterpret_cast<long long*>(dev_buf);
  TGFR_CUDA_OK(cudaMemcpyToSymbol(g_trace, &p, sizeof(p)));
  return TGFR_OK;
}

// bytes of the forward -> backward state (0 if the shape has no plan).  Two layouts, told apart by their size:
//   records (default)              A1 | E per (caption, word, region) + V tiles: the backward runs no exponential
//   TGFR_WORDREGION_SAVE=wu        Wu tiles + alpha | beta per word: 2.5x fewer bytes, the backward recomputes S and E
size_t wordregion_tc_saved_bytes(int Bc, int Bq, int T, int R, int D) {
  TcPlan pl;
  if (make_plan(Bc, Bq, T, R, D, &pl) != TGFR_OK) return 0;
  const char* m = getenv("TGFR_WORDREGION_SAVE");
  if (m && (strcmp(m, "wu") == 0 || strcmp(m, "WU") == 0)) return saved_layout(pl, Bc, Bq, R, D).total;
  return rec::saved_layout(pl, Bc, Bq, R, D).total;
}

// which layout a buffer of `saved_bytes` holds: 0 none / too small, 1 records, 2 Wu tiles
int wordregion_tc_saved_mode(int Bc, int Bq, int T, int R, int D, const void* saved, size_t saved_bytes) {
  TcPlan pl;
  if (make_plan(Bc, Bq, T, R, D, &pl) != TGFR_OK) return kSavedNone;
  return saved_mode_of(saved, saved_bytes, saved_layout(pl, Bc, Bq, R, D).total, rec::saved_layout(pl, Bc, Bq, R, D).total);
}

size_t wordregion_tc_workspace_bytes(int Bc, int Bq, int T, int R, int D) {
  TcPlan pl;
  if (make_plan(Bc, Bq, T, R, D, &pl) != TGFR_OK) return 0;
  return pl.ws_total;
}

// attn_diag != NULL: the forward also emits the attention map of every face's own caption (b, b + diag_off) -> [Bc, T, R]
int wordregion_fwd_tc(const float* ctx, int64_t csb, int64_t csr, int64_t csd, const float* words, int64_t wsb,
                      int64_t wst, int64_t wsd, const int32_t* cap_lens, int Bc, int Bq, int T, int R, int D, float g1,
                      float g2, float g3, float eps, float* sim, float* attn_diag, int diag_off, void* ws, size_t ws_bytes,
                      void* saved, size_t saved_bytes, cudaStream_t st) {
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
  const bool save = saved != nullptr;
  const SavedLayout L = saved_layout(pl, Bc, Bq, R, D);
  const rec::SavedLayout LR = rec::saved_layout(pl, Bc, Bq, R, D);
  const int sv_mode = saved_mode_of(saved, saved_bytes, L.total, LR.total);
  const bool records = sv_mode == kSavedRecords;
  if (save) {
    TGFR_REQUIRE(sv_mode != kSavedNone, "wordregion(tc): saved buffer too small (%zu < %zu)", saved_bytes, L.total);
    TGFR_REQUIRE((reinterpret_cast<uintptr_t>(saved) & 255) == 0, "wordregion(tc): saved buffer must be 256-byte aligned");
    uint8_t* sv = reinterpret_cast<uint8_t*>(saved);     // the operand copies live with the saved state: the backward reuses them
    c16 = reinterpret_cast<__half*>(sv + (records ? LR.off_c16 : L.off_c16));
    q16 = reinterpret_cast<__half*>(sv + (records ? LR.off_q16 : L.off_q16));
    qnorm = reinterpret_cast<float*>(sv + (records ? LR.off_qnorm : L.off_qnorm));
    lens = reinterpret_cast<int*>(sv + (records ? LR.off_lens : L.off_lens));
  }

  const int64_t rows = (int64_t)Bc * R + (int64_t)Bq * pl.Tp;
  TGFR_CUDA_OK(cudaMemsetAsync(lens + Bq, 0, 2 * sizeof(int), st));
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
  p.k1 = g1 * kLog2e; p.k2 = g2 * kLog2e; p.g3 = g3; p.g23 = g2 * g3;
  p.uniform_len = cap_lens ? 0 : T;
  p.T = T; p.diag_off = diag_off;
  if (attn_diag) {
    p.attn_out = attn_diag;
    p.attn_z = reinterpret_cast<float*>(base + pl.ws_attz);
    TGFR_CUDA_OK(cudaMemsetAsync(p.attn_z, 0, sizeof(float) * (size_t)Bc * 32, st));
  }
  if (save) {
    uint8_t* sv = reinterpret_cast<uint8_t*>(saved);
    if (records) {
      p.sv_v = reinterpret_cast<__half*>(sv + LR.off_v);
      p.sv_rec = sv + LR.off_rec;
      p.sv_inw = reinterpret_cast<float*>(sv + LR.off_inw);
      p.rec_stride = LR.rec_stride;
    } else {
      p.sv_wu = reinterpret_cast<__half*>(sv + L.off_wu);
      p.sv_ab = reinterpret_cast<float*>(sv + L.off_ab);
    }
  }

  int dev = 0, sms = 0;
  TGFR_CUDA_OK(cudaGetDevice(&dev));
  TGFR_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int grid = p.total_units < sms ? p.total_units : sms;
  if (const char* dg = getenv("TGFR_DEBUG_GRID")) {             // profiling aid: fewer CTAs -> no L2 contention
    const int v = atoi(dg);
    if (v > 0 && v < grid) grid = v;
  }
#define TGFR_LAUNCH_FWD1(TPV, LAYV)                                                                               \
  {                                                                                                              \
    static bool attr_done[64] = {};                                                                              \
    bool& attr_set = attr_done[dev & 63];                                                                        \
    if (!attr_set) {                                                                                             \
      TGFR_CUDA_OK(cudaFuncSetAttribute(wr_tc_fwd3_kernel<TPV, LAYV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        232448));                                                                \
      attr_set = true;                                                                                           \
    }                                                                                                            \
    wr_tc_fwd3_kernel<TPV, LAYV><<<grid, kF3Threads, pl.smem_bytes, st>>>(tm_c, tm_q, p);                          \
  }
#define TGFR_LAUNCH_FWD(TPV)                          \
  case TPV:                                           \
    if (records) TGFR_LAUNCH_FWD1(TPV, kLayRec)       \
    else if (save) TGFR_LAUNCH_FWD1(TPV, kLayWu)      \
    else TGFR_LAUNCH_FWD1(TPV, kLayNone)              \
    break;
  switch (pl.Tp) {
    TGFR_LAUNCH_FWD(8)
    TGFR_LAUNCH_FWD(16)
    TGFR_LAUNCH_FWD(24)
    TGFR_LAUNCH_FWD(32)
    default:
      set_error("wordregion(tc): unsupported padded caption length %d", pl.Tp);
      return TGFR_E_INVALID;
  }
#undef TGFR_LAUNCH_FWD
#undef TGFR_LAUNCH_FWD1
  TGFR_LAUNCH_OK();
  if (attn_diag) {
    const int64_t n = (int64_t)Bc * T * R;
    attn_normalize_kernel<<<(unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, st>>>(attn_diag, p.attn_z, lens, Bc, Bq, T,
                                                                                                    R, 32, diag_off, p.uniform_len);
    TGFR_LAUNCH_OK();
  }
  return TGFR_OK;
}

}  // namespace tgfr

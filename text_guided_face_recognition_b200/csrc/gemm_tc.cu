// Tensor-core GEMM of the margin heads (TGFR_PREC_TC): C[M,N] (fp32) (+)= alpha * sum_k A(m,k) B(n,k)
// with fp16 operands staged by TMA and fp32 accumulation in TMEM (tcgen05.mma kind::f16).
//
// Three contractions of the head run through it (models/metrics.py:44, its autograd, models/magface.py:92-94):
//   cos    [B,C]   = x^ . w^T          A = x^ [B,Din] K-major,            B = w^ [C,Din] K-major
//   dX^    [B,Din] = g  . w^           A = g  [B,C]   K-major,            B = w^ [C,Din] read MN-major (split-K, atomics)
//   dW^    [C,Din] = g^T. x^           A = g  [B,C]   read MN-major,      B = x^ [B,Din] read MN-major
// so no operand is ever transposed in memory: the UMMA shared-memory descriptors select the major-ness.
//
// The cos-theta GEMM also has two fused ArcFace epilogues that never write the [B, C] logits (kEpiCeStats,
// kEpiCeGrad below): margin + online-softmax statistics in the forward, margin + softmax gradient as the fp16
// operand of the two gradient GEMMs in the backward (metrics.py:45-57 + losses.py:321-325 in one pass each).
//
// One 128 x BN output tile per CTA, TMA pipeline over K in steps of 64, two or three CTAs per SM so that one CTA's
// epilogue overlaps another's main loop.  A CTA lives ~10 us whatever it does (set-up, pipeline fill, epilogue), so a
// partial second wave costs as much as the first: the host picks the tile width / residency that makes EVERY tile
// co-resident (B = 512, C = 10177: 128-wide tiles are 320 CTAs for 296 slots; 160-wide tiles are 256; the MN-major
// products, whose panels are 64 wide, run 128-wide tiles on 2 stages at three CTAs per SM = 444 slots).  Roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM
// allocator, warps 2-5 epilogue (TMEM -> registers -> warp-private shared-memory transpose -> coalesced row stores,
// which is what the unaligned row pitch of a [B, 10177] logits matrix needs).
#include <string.h>

#include "common.cuh"
#include "tc.cuh"

namespace tgfr {
namespace {

using namespace tc;

constexpr int kGemmThreads = 192;
constexpr int kBM = 128, kBK = 64;
constexpr uint32_t kTileBytes = kBM * kBK * 2;           // 16 KB of A per stage; B: BN x 128 bytes
__host__ __device__ constexpr uint32_t gemm_stage_bytes(int bn) { return kTileBytes + (uint32_t)bn * kBK * 2; }
__host__ __device__ constexpr uint32_t gemm_smem_bytes(int bn, int stages) {
  return (uint32_t)stages * gemm_stage_bytes(bn) + 1024 /*barriers*/ + 1024 /*alignment*/;
}

enum Epi { kEpiStore = 0, kEpiCeStats = 1, kEpiCeGrad = 2 };

// ArcFace margin on the label column (metrics.py:45-57): phi(c) and d phi / d c
__device__ __forceinline__ float arc_phi_tc(float c, float cm, float sm, float th, float mm, int easy, float* dphi) {
  const float one_m = 1.f - c * c;
  const float sine = sqrtf(fminf(fmaxf(one_m, 0.f), 1.f));
  const bool use = easy ? (c > 0.f) : (c > th);
  if (dphi) {
    const bool inside = one_m >= 0.f && one_m <= 1.f;
    *dphi = use ? (cm + (inside ? c / sine : 0.f) * sm) : 1.f;
  }
  return use ? (c * cm - sine * sm) : (easy ? c : c - mm);
}

struct CeParams {            // fused ArcFace + cross-entropy epilogues (C = cos-theta tile, alpha = s)
  const int64_t* labels;     // [M]
  int class_off;             // global class index of column 0 (class-sharded heads)
  float cm, sm, th, mm;      // cos m, sin m, cos(pi - m), sin(pi - m) m
  int easy;
  // kEpiCeStats: per (row, column tile) online-softmax partials, target logit and cos(theta_y) of the owner tile
  float* pmax;               // [M, n_tiles]
  float* psum;               // [M, n_tiles]
  float* tgt;                // [M] pre-zeroed
  float* cos_t;              // [M] pre-filled with NaN
  // kEpiCeGrad: g16[b, c] = fp16(2^e k (softmax - onehot) (phi' on the label column)), k = coef gout / M
  const float* lse;          // [M] global log-sum-exp of the rows
  const float* coef;         // device scalar (focal factor) or NULL
  const float* gout;         // device scalar (upstream gradient) or NULL
  float* scale;              // scale[1] = 2^-e for the gradient GEMMs (written by block (0,0))
  __half* g16;               // [M, ld_g]
  int ld_g;
  // kEpiCeStats, optional: the cos-theta tile itself, fp32 [M, ld_cos] (ld_cos % 4 == 0), so that the backward derives the
  // softmax gradient with an element-wise pass (arc_ce_grad_from_cos) instead of a second GEMM
  float* cos_out;
  int ld_cos;
};

struct GemmTcParams {
  float* C;
  int64_t ldc;
  int M, N, K;
  int kt_per_split;      // 64-wide K steps per blockIdx.z
  float alpha;
  const float* dscale;   // optional device scalar: alpha is multiplied by dscale[1] (power-of-two gradient scaling)
  int clamp;             // clamp the accumulator to [-1, 1] before scaling (magface.py:94)
  int atomic;            // accumulate with atomics (split-K; C pre-zeroed)
  int a_mn, b_mn;
  CeParams ce;
  // plain-store epilogue extras (TextHeading products): out = relu?(alpha acc + bias[col])
  const float* bias;     // [N] or NULL
  int relu;
  const float* dscale2;  // second optional device scalar (alpha *= dscale2[1]): one power-of-two scale per operand
  // error-compensated products: nterms = 3 accumulates A_hi B_hi + A_hi B_lo + A_lo B_hi (fp16 hi / lo splits of fp32
  // operands, ~22 significant bits) in the one TMEM accumulator by walking K three times with different tensor maps
  int nterms;            // 1 or 3
  // batched = 1: blockIdx.z selects one of up to three independent products (own maps, shapes and outputs; no split-K)
  int batched;
  // strided = 1: blockIdx.z is the sample of a strided batch (third tensor-map coordinate; C advances by c_bs; no split-K)
  int strided;
  int64_t c_bs;
  // conv_c != 0: implicit 3x3 convolution over a channels-last image matrix A [rows = positions, conv_c channels]: K runs
  // over 9 taps x conv_c channels, and tap (dy, dx) reads A shifted down by dy * conv_w + dx rows (valid convolution:
  // the outputs of the anchors whose window leaves the image are garbage the caller ignores)
  int conv_c, conv_w;
  // optional: max |output| over the whole product, as float bits via atomicMax (zeroed by the caller) -- the operand scale of
  // the NEXT split product comes out of this product's epilogue instead of a separate pass over the result
  float* amax;
  struct Z {
    float* C;
    const float* bias;
    int64_t ldc;
    int M, N, K;
  } z[3];
};

struct GemmMaps {        // [product z][0 = hi (or the only operand), 1 = lo]
  CUtensorMap a[3][2];
  CUtensorMap b[3][2];
};

// The epilogue of one 128 x BN tile whose accumulator sits at TMEM address `tmem`: warp quarter q owns rows
// m0 + 32 q .. + 31.  `stage` is that warp's private 32 x 33 float transpose buffer; (bx, by) the tile's position in
// the nx-wide tile grid.  Shared by the one-tile-per-CTA kernel and the persistent kernel.
template <int EPI, int BN>
__device__ __forceinline__ void gemm_epilogue(const GemmTcParams& p, uint32_t tmem, float* stage, int m0, int n0, int bx, int by,
                                              int nx, int q, int lane, int warp, int Mz, int Nz, float* Cz, int64_t ldcz,
                                              const float* biasz) {
  const int row_base = m0 + q * 32;
  const float alpha = p.alpha * (p.dscale ? __ldg(p.dscale + 1) : 1.f) * (p.dscale2 ? __ldg(p.dscale2 + 1) : 1.f);
  if constexpr (EPI != kEpiStore) {
    // a thread owns row (row_base + lane) of the tile: logits = s cos, phi on the label column
    const int row = row_base + lane;
    const bool row_ok = row < p.M;
    const int64_t ycol = row_ok ? __ldg(p.ce.labels + row) - p.ce.class_off : -1;     // label column in this shard
    float kscale = 0.f, lse = 0.f;
    if constexpr (EPI == kEpiCeGrad) {
      // power-of-two scale from the row-independent factor k: entries are k (p - onehot) phi', |.| <= k max(1, phi')
      const float k = (p.ce.coef ? __ldg(p.ce.coef) : 1.f) * (p.ce.gout ? __ldg(p.ce.gout) : 1.f) / (float)p.M;
      const float ak = fabsf(k);
      const float sc = (ak > 0.f) ? exp2f(floorf(log2f(256.f / ak))) : 1.f;
      kscale = k * sc;
      if (bx == 0 && by == 0 && warp == 2 && lane == 0) p.ce.scale[1] = 1.f / sc;
      lse = row_ok ? __ldg(p.ce.lse + row) : 0.f;
    }
    float mx = -INFINITY, sum = 0.f;
#pragma unroll 1
    for (int ch = 0; ch < BN / 32; ++ch) {
      const int c0 = n0 + 32 * ch;
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 32 * ch, v);      // warp-collective even past the last column
      tmem_ld_wait();
      if constexpr (EPI == kEpiCeStats) {
        if (p.ce.cos_out != nullptr && row_ok) {                         // a thread owns 32 consecutive columns of its row
          float* cd = p.ce.cos_out + (int64_t)row * p.ce.ld_cos + c0;
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4)
            if (c0 + 4 * q4 < p.ce.ld_cos)
              *reinterpret_cast<uint4*>(cd + 4 * q4) = make_uint4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
        }
      }
      float lg[32];
      float dph = 1.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float c = __uint_as_float(v[j]);
        float l = c * p.alpha;
        if ((int64_t)(c0 + j) == ycol && c0 + j < p.N) {   // a label of the NEXT class shard can fall into this tile's padding
          l = p.alpha * arc_phi_tc(c, p.ce.cm, p.ce.sm, p.ce.th, p.ce.mm, p.ce.easy, &dph);
          if constexpr (EPI == kEpiCeStats) {
            p.ce.cos_t[row] = c;
            p.ce.tgt[row] = l;
          }
        }
        lg[j] = (c0 + j < p.N) ? l : -INFINITY;
      }
      if constexpr (EPI == kEpiCeStats) {
        float cmx = lg[0];
#pragma unroll
        for (int j = 1; j < 32; ++j) cmx = fmaxf(cmx, lg[j]);
        if (cmx > mx) {
          sum *= expf(mx - cmx);          // exp(-inf) = 0 on the first chunk
          mx = cmx;
        }
        if (mx > -INFINITY) {
#pragma unroll
          for (int j = 0; j < 32; ++j) sum += expf(lg[j] - mx);
        }
      } else {
        // a thread owns 32 consecutive entries of ITS row: four 16-byte stores, no transpose (ld_g % 8 == 0)
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float g2[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            float g = 0.f;
            if (c0 + j + u < p.N) {
              const bool on = (int64_t)(c0 + j + u) == ycol;
              g = kscale * (expf(lg[j + u] - lse) - (on ? 1.f : 0.f)) * (on ? dph : 1.f);
              g = fminf(fmaxf(g, -65504.f), 65504.f);
            }
            g2[u] = g;
          }
          const __half2 h2 = __floats2half2_rn(g2[0], g2[1]);
          pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        if (row_ok) {
          __half* dst = p.ce.g16 + (int64_t)row * p.ce.ld_g + c0;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            if (c0 + 8 * q4 < p.ce.ld_g)                           // columns [N, ld_g) are the zero K padding
              *reinterpret_cast<uint4*>(dst + 8 * q4) = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
        }
      }
    }
    if constexpr (EPI == kEpiCeStats) {
      if (row_ok) {
        p.ce.pmax[(int64_t)row * nx + bx] = mx;
        p.ce.psum[(int64_t)row * nx + bx] = sum;
      }
    }
  } else {
  const bool direct = (ldcz & 3) == 0 && (reinterpret_cast<uintptr_t>(Cz) & 15) == 0;
  float amax_t = 0.f;
#pragma unroll 1
  for (int ch = 0; ch < BN / 32; ++ch) {
    const int c0 = n0 + 32 * ch;
    if (c0 >= Nz) break;
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 32 * ch, v);
    tmem_ld_wait();
    if (direct) {
      // a thread owns 32 consecutive outputs of ITS row (128 contiguous bytes): 16-byte stores or vector reductions
      const int row = row_base + lane;
      if (row < Mz) {
        float* dst = Cz + (int64_t)row * ldcz + c0;
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4) {
          float o[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float x = __uint_as_float(v[4 * g4 + u]);
            if (p.clamp) x = fminf(fmaxf(x, -1.f), 1.f);
            x *= alpha;
            if (biasz && c0 + 4 * g4 + u < Nz) x += __ldg(biasz + c0 + 4 * g4 + u);
            if (p.relu) x = fmaxf(x, 0.f);
            o[u] = x;
            if (c0 + 4 * g4 + u < Nz) amax_t = fmaxf(amax_t, fabsf(x));
          }
          const int c = c0 + 4 * g4;
          if (c + 3 < Nz) {
            if (p.atomic) red_add_v4(dst + 4 * g4, o[0], o[1], o[2], o[3]);
            else *reinterpret_cast<float4*>(dst + 4 * g4) = make_float4(o[0], o[1], o[2], o[3]);
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (c + u < Nz) {
                if (p.atomic) atomicAdd(dst + 4 * g4 + u, o[u]);
                else dst[4 * g4 + u] = o[u];
              }
          }
        }
      }
      continue;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float x = __uint_as_float(v[j]);
      if (p.clamp) x = fminf(fmaxf(x, -1.f), 1.f);
      x *= alpha;
      if (biasz && c0 + j < Nz) x += __ldg(biasz + c0 + j);
      if (p.relu) x = fmaxf(x, 0.f);
      stage[lane * 33 + j] = x;
      if (row_base + lane < Mz && c0 + j < Nz) amax_t = fmaxf(amax_t, fabsf(x));
    }
    __syncwarp();
    const int col = c0 + lane;
    if (col < Nz) {
      const int nrows = min(32, Mz - row_base);
      for (int rr = 0; rr < nrows; ++rr) {
        float* dst = Cz + (int64_t)(row_base + rr) * ldcz + col;
        const float val = stage[rr * 33 + lane];
        if (p.atomic) atomicAdd(dst, val);
        else *dst = val;
      }
    }
    __syncwarp();
  }
  if (p.amax) {
    amax_t = warp_max(amax_t);
    if (lane == 0 && __float_as_int(amax_t) > __ldcg(reinterpret_cast<const int*>(p.amax)))
      atomicMax(reinterpret_cast<int*>(p.amax), __float_as_int(amax_t));
  }
  }
}

template <int EPI, int BN, int STAGES, int MINB>
__global__ void __launch_bounds__(kGemmThreads, MINB)
gemm_tc_kernel(const __grid_constant__ GemmMaps maps, const GemmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kStages = STAGES, kBN = BN;
  constexpr uint32_t kStageBytes = gemm_stage_bytes(BN);
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "tile width: whole 32-column epilogue chunks");
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty = full + kStages;
  uint64_t* accum = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
  const int zi = p.batched ? (int)blockIdx.z : 0;
  const int Mz = p.batched ? p.z[zi].M : p.M, Nz = p.batched ? p.z[zi].N : p.N, Kz = p.batched ? p.z[zi].K : p.K;
  const int zc = p.strided ? (int)blockIdx.z : 0;
  float* const Cz = (p.batched ? p.z[zi].C : p.C) + (p.strided ? (int64_t)blockIdx.z * p.c_bs : 0);
  const int64_t ldcz = p.batched ? p.z[zi].ldc : p.ldc;
  const float* const biasz = p.batched ? p.z[zi].bias : p.bias;
  if (m0 >= Mz || n0 >= Nz) return;                        // batched products share one grid: tiles outside this one
  const int kt_total = (Kz + kBK - 1) / kBK;
  const int kt0 = (p.batched || p.strided) ? 0 : blockIdx.z * p.kt_per_split;
  const int kt1 = (p.batched || p.strided) ? kt_total : min(kt_total, kt0 + p.kt_per_split);
  const int nk1 = kt1 - kt0;                               // >= 1 by construction of the grid
  const int nkt = nk1 * (p.nterms == 3 ? 3 : 1);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
    tma_prefetch_desc(&maps.a[zi][0]);
    tma_prefetch_desc(&maps.b[zi][0]);
  }
  constexpr uint32_t kTmemCols = BN <= 128 ? 128 : 256;
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < nkt; ++it) {
        const int s = it % kStages, use = it / kStages;
        if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
        uint8_t* sa = smem + s * kStageBytes;
        uint8_t* sb = sa + kTileBytes;
        // three terms: hi lo, lo hi, then hi hi.  The tensor core truncates (toward -inf) when it adds into the fp32
        // accumulator, a bias proportional to the accumulator's magnitude per step: the two small cross terms go first,
        // while the accumulator is 2^-11 of its final size, so only the K / 16 hi hi steps pay it (measured on the
        // K = 25 088 column sums of IMIM's backward: 3e-4 -> 1e-4 relative).
        const int term = it / nk1;
        const int k0 = (kt0 + it - term * nk1) * kBK;
        const bool three = p.nterms == 3;
        const CUtensorMap* const ta = &maps.a[zi][three && term == 1 ? 1 : 0];
        const CUtensorMap* const tb = &maps.b[zi][three && term == 0 ? 1 : 0];
        mbar_arrive_expect_tx(&full[s], kStageBytes);
        if (p.a_mn) {                                        // memory [K, M]: two panels of 64 M-elements x 64 K-rows
          tma_load_3d(sa, ta, &full[s], m0, k0, zc);
          tma_load_3d(sa + 8192, ta, &full[s], m0 + 64, k0, zc);
        } else if (p.conv_c) {                               // tap-shifted rows of the image matrix
          const int tap = k0 / p.conv_c;
          tma_load_3d(sa, ta, &full[s], k0 - tap * p.conv_c, m0 + (tap / 3) * p.conv_w + tap % 3, zc);
        } else {                                             // memory [M, K]: 128 rows x 64 K-elements
          tma_load_3d(sa, ta, &full[s], k0, m0, zc);
        }
        if (p.b_mn) {
          tma_load_3d(sb, tb, &full[s], n0, k0, zc);
          tma_load_3d(sb + 8192, tb, &full[s], n0 + 64, k0, zc);
        } else {
          tma_load_3d(sb, tb, &full[s], k0, n0, zc);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(kBM, kBN, p.a_mn != 0, p.b_mn != 0);
      for (int it = 0; it < nkt; ++it) {
        const int s = it % kStages, use = it / kStages;
        mbar_wait(&full[s], use & 1);
        tc_fence_after();
        const uint32_t a = smem_u32(smem + s * kStageBytes), b = a + kTileBytes;
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          const uint64_t ad = p.a_mn ? make_smem_desc(a + k16 * 2048, 8192, 1024) : make_smem_desc(a + k16 * 32, 16, 1024);
          const uint64_t bd = p.b_mn ? make_smem_desc(b + k16 * 2048, 8192, 1024) : make_smem_desc(b + k16 * 32, 16, 1024);
          umma_ss(tmem, ad, bd, idesc, (it | k16) != 0);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(accum);
    }
  } else {
    // epilogue: warp (2..5) owns TMEM lane quarter warp % 4, i.e. rows m0 + 32 q .. + 31
    const int q = warp & 3;
    mbar_wait(accum, 0);
    tc_fence_after();
    // the operand stages are dead now: warp-private 32 x 33 float transpose buffers live in stage 0
    gemm_epilogue<EPI, BN>(p, tmem, reinterpret_cast<float*>(smem) + q * (32 * 33), m0, n0, (int)blockIdx.x, (int)blockIdx.y,
                           (int)gridDim.x, q, lane, warp, Mz, Nz, Cz, ldcz, biasz);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Persistent form of the same GEMM: one CTA per SM walks the tile list (x fastest, so concurrently running CTAs share the
// A rows and stream different B tiles), the TMA ring runs on across tile boundaries and the accumulator is double
// buffered in TMEM -- tile i+1's main loop runs under tile i's epilogue, and the per-CTA set-up (TMEM allocation, barrier
// initialisation, pipeline fill) is paid once per SM instead of once per tile.  Same roles, same epilogues
// (gemm_epilogue), same split-K / multi-product / strided-batch meaning of the tile grid's z axis.
// ------------------------------------------------------------------------------------------------------------
constexpr uint32_t kEpiStageBytes = 4 * 32 * 33 * sizeof(float) + 512;      // warp-private transpose buffers, padded
__host__ __device__ constexpr uint32_t gemm_persist_smem_bytes(int bn, int stages) {
  return (uint32_t)stages * gemm_stage_bytes(bn) + 1024 /*barriers*/ + kEpiStageBytes + 1024 /*alignment*/;
}

struct TileInfo {
  int bx, by, zi, zc, m0, n0, Mz, Nz, kt0, nk1, nkt;
  bool skip;
  float* Cz;
  int64_t ldcz;
  const float* biasz;
};
template <int BN>
__device__ __forceinline__ TileInfo tile_info(const GemmTcParams& p, int t, int nx, int ny) {
  TileInfo ti;
  ti.bx = t % nx;
  const int r = t / nx;
  ti.by = r % ny;
  const int bz = r / ny;
  ti.zi = p.batched ? bz : 0;
  ti.zc = p.strided ? bz : 0;
  ti.Mz = p.batched ? p.z[ti.zi].M : p.M;
  ti.Nz = p.batched ? p.z[ti.zi].N : p.N;
  const int Kz = p.batched ? p.z[ti.zi].K : p.K;
  ti.Cz = (p.batched ? p.z[ti.zi].C : p.C) + (p.strided ? (int64_t)bz * p.c_bs : 0);
  ti.ldcz = p.batched ? p.z[ti.zi].ldc : p.ldc;
  ti.biasz = p.batched ? p.z[ti.zi].bias : p.bias;
  ti.m0 = ti.by * kBM;
  ti.n0 = ti.bx * BN;
  ti.skip = ti.m0 >= ti.Mz || ti.n0 >= ti.Nz;
  const int kt_total = (Kz + kBK - 1) / kBK;
  ti.kt0 = (p.batched || p.strided) ? 0 : bz * p.kt_per_split;
  const int kt1 = (p.batched || p.strided) ? kt_total : min(kt_total, ti.kt0 + p.kt_per_split);
  ti.nk1 = kt1 - ti.kt0;
  ti.nkt = ti.nk1 * (p.nterms == 3 ? 3 : 1);
  return ti;
}

template <int EPI, int BN, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_persist_kernel(const __grid_constant__ GemmMaps maps, const GemmTcParams p, const int nx, const int ny, const int nz) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t kStageBytes = gemm_stage_bytes(BN);
  constexpr uint32_t kAccCols = BN <= 128 ? 128 : 256;
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "tile width: whole 32-column epilogue chunks");
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;       // [2] accumulator buffer written (MMA -> epilogue)
  uint64_t* acc_empty = acc_full + 2;        // [2] accumulator buffer drained (epilogue -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * kStageBytes + 1024);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles = nx * ny * nz;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 4);           // one arrival per epilogue warp
    }
    fence_barrier_init();
    tma_prefetch_desc(&maps.a[0][0]);
    tma_prefetch_desc(&maps.b[0][0]);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * kAccCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;                                         // ring position, running across tiles
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const TileInfo ti = tile_info<BN>(p, t, nx, ny);
        if (ti.skip) continue;
        const bool three = p.nterms == 3;
        for (int k = 0; k < ti.nkt; ++k, ++it) {
          const uint32_t s = it % STAGES, use = it / STAGES;
          if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
          uint8_t* sa = smem + s * kStageBytes;
          uint8_t* sb = sa + kTileBytes;
          const int term = k / ti.nk1;                         // hi lo, lo hi, then hi hi (see gemm_tc_kernel)
          const int k0 = (ti.kt0 + k - term * ti.nk1) * kBK;
          const CUtensorMap* const ta = &maps.a[ti.zi][three && term == 1 ? 1 : 0];
          const CUtensorMap* const tb = &maps.b[ti.zi][three && term == 0 ? 1 : 0];
          mbar_arrive_expect_tx(&full[s], kStageBytes);
          if (p.a_mn) {
            tma_load_3d(sa, ta, &full[s], ti.m0, k0, ti.zc);
            tma_load_3d(sa + 8192, ta, &full[s], ti.m0 + 64, k0, ti.zc);
          } else if (p.conv_c) {
            const int tap = k0 / p.conv_c;
            tma_load_3d(sa, ta, &full[s], k0 - tap * p.conv_c, ti.m0 + (tap / 3) * p.conv_w + tap % 3, ti.zc);
          } else {
            tma_load_3d(sa, ta, &full[s], k0, ti.m0, ti.zc);
          }
          if (p.b_mn) {
            tma_load_3d(sb, tb, &full[s], ti.n0, k0, ti.zc);
            tma_load_3d(sb + 8192, tb, &full[s], ti.n0 + 64, k0, ti.zc);
          } else {
            tma_load_3d(sb, tb, &full[s], k0, ti.n0, ti.zc);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(kBM, BN, p.a_mn != 0, p.b_mn != 0);
      uint32_t it = 0, li = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const TileInfo ti = tile_info<BN>(p, t, nx, ny);
        if (ti.skip) continue;
        const uint32_t buf = li & 1, ub = li >> 1;
        mbar_wait(&acc_empty[buf], (ub & 1) ^ 1);              // a fresh barrier passes the wait on parity 1 at once
        tc_fence_after();
        const uint32_t acc = tmem + buf * kAccCols;
        for (int k = 0; k < ti.nkt; ++k, ++it) {
          const uint32_t s = it % STAGES, use = it / STAGES;
          mbar_wait(&full[s], use & 1);
          tc_fence_after();
          const uint32_t a = smem_u32(smem + s * kStageBytes), b = a + kTileBytes;
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
            const uint64_t ad = p.a_mn ? make_smem_desc(a + k16 * 2048, 8192, 1024) : make_smem_desc(a + k16 * 32, 16, 1024);
            const uint64_t bd = p.b_mn ? make_smem_desc(b + k16 * 2048, 8192, 1024) : make_smem_desc(b + k16 * 32, 16, 1024);
            umma_ss(acc, ad, bd, idesc, (k | k16) != 0);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[buf]);
        ++li;
      }
    }
  } else {
    const int q = warp & 3;
    uint32_t li = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      const TileInfo ti = tile_info<BN>(p, t, nx, ny);
      if (ti.skip) continue;
      const uint32_t buf = li & 1, ub = li >> 1;
      mbar_wait(&acc_full[buf], ub & 1);
      tc_fence_after();
      gemm_epilogue<EPI, BN>(p, tmem + buf * kAccCols, epi_stage + q * (32 * 33), ti.m0, ti.n0, ti.bx, ti.by, nx, q, lane, warp, ti.Mz,
                             ti.Nz, ti.Cz, ti.ldcz, ti.biasz);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      ++li;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem, 2 * kAccCols);
  }
}

// one warp per vector: out16[v, :] = x_v / max(norm[v], 1e-12) as fp16 (zero padded to ld_out)
__global__ void normalize_rows_f16_kernel(const float* __restrict__ x, int64_t s_vec, int64_t s_elem, int nvec, int len,
                                          const float* __restrict__ norm, __half* __restrict__ out, int ld_out) {
  const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (v >= nvec) return;
  const float* src = x + (int64_t)v * s_vec;
  const float inv = 1.f / fmaxf(__ldg(norm + v), 1e-12f);
  __half* dst = out + (int64_t)v * ld_out;
  for (int k = lane; k < ld_out; k += 32) dst[k] = __float2half_rn(k < len ? __ldg(src + (int64_t)k * s_elem) * inv : 0.f);
}

// weight stored [len, nvec] (vector index contiguous: MagLinear.weight [Din, C]): 32 x 32 tiles through shared memory,
// norms come from colnorm (already computed); out16 [nvec, ld_out]
__global__ void normalize_cols_f16_kernel(const float* __restrict__ x, int64_t s_elem, int nvec, int len,
                                          const float* __restrict__ norm, __half* __restrict__ out, int ld_out) {
  __shared__ float tile[32][33];
  const int v0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 256 threads: 8 rows per pass
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, v = v0 + tx;
    tile[r][tx] = (k < len && v < nvec) ? __ldg(x + (int64_t)k * s_elem + v) : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int v = v0 + r, k = k0 + tx;
    if (v < nvec && k < ld_out) out[(int64_t)v * ld_out + k] = __float2half_rn(tile[tx][r] / fmaxf(__ldg(norm + v), 1e-12f));
  }
}

// norms + fp16 normalised operands of BOTH head operands in one launch (rows contiguous, len % 4 == 0, len <= 1024):
// a warp per vector, the vector held in registers, so x and w are read once (the two-kernel form read them twice
// and took four launches): norm[v] = |x_v|, out16[v, :] = x_v / max(|x_v|, 1e-12), zero padded to ld_out
__global__ void __launch_bounds__(256) norm_f16_pair_kernel(const float* __restrict__ x, int64_t x_sr, int nx,
                                                            const float* __restrict__ w, int64_t w_sr, int nw, int len,
                                                            float* __restrict__ xnorm, float* __restrict__ wnorm,
                                                            __half* __restrict__ x16, __half* __restrict__ w16, int ld_out) {
  const int lane = threadIdx.x & 31;
  const int nv4 = len >> 2;
  for (int v = blockIdx.x * 8 + (threadIdx.x >> 5); v < nx + nw; v += gridDim.x * 8) {
    const bool is_x = v < nx;
    const int r = is_x ? v : v - nx;
    const float4* src = reinterpret_cast<const float4*>(is_x ? x + (int64_t)r * x_sr : w + (int64_t)r * w_sr);
    float4 e[8];
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int idx = k * 32 + lane;
      e[k] = idx < nv4 ? __ldg(src + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      acc = fmaf(e[k].x, e[k].x, acc);
      acc = fmaf(e[k].y, e[k].y, acc);
      acc = fmaf(e[k].z, e[k].z, acc);
      acc = fmaf(e[k].w, e[k].w, acc);
    }
    const float nrm = sqrtf(warp_sum(acc));
    if (lane == 0) (is_x ? xnorm : wnorm)[r] = nrm;
    const float inv = 1.f / fmaxf(nrm, 1e-12f);
    __half* dst = (is_x ? x16 : w16) + (int64_t)r * ld_out;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int idx = k * 32 + lane;
      if (4 * idx < ld_out) {                                   // ld_out % 8 == 0 and len % 4 == 0: whole 8-byte groups
        const __half2 lo = __floats2half2_rn(e[k].x * inv, e[k].y * inv), hi = __floats2half2_rn(e[k].z * inv, e[k].w * inv);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo);
        pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(dst + 4 * idx) = pk;
      }
    }
  }
}

// d(normalize(v)) / dv projections of BOTH head operands in one launch: out = (g - <g, v^> v^) / |v| with
// v^ = v / max(|v|, eps); rows contiguous, len % 4 == 0, len <= 1024: a warp per vector, everything in registers
__global__ void __launch_bounds__(256) normalize_bwd_pair_kernel(const float* __restrict__ gx, const float* __restrict__ x,
                                                                 int64_t x_sr, const float* __restrict__ xnorm,
                                                                 float* __restrict__ dx, int nx, const float* __restrict__ gw,
                                                                 const float* __restrict__ w, int64_t w_sr,
                                                                 const float* __restrict__ wnorm, float* __restrict__ dw,
                                                                 int64_t dw_sr, int nw, int len, float eps) {
  const int lane = threadIdx.x & 31;
  const int nv4 = len >> 2;
  for (int v = blockIdx.x * 8 + (threadIdx.x >> 5); v < nx + nw; v += gridDim.x * 8) {
    const bool is_x = v < nx;
    const int r = is_x ? v : v - nx;
    const float4* gp = reinterpret_cast<const float4*>((is_x ? gx : gw) + (int64_t)r * len);
    const float4* vp = reinterpret_cast<const float4*>(is_x ? x + (int64_t)r * x_sr : w + (int64_t)r * w_sr);
    float4* op = reinterpret_cast<float4*>(is_x ? dx + (int64_t)r * len : dw + (int64_t)r * dw_sr);
    const float inv = 1.f / fmaxf(__ldg((is_x ? xnorm : wnorm) + r), eps);
    float4 g[8], vh[8];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int idx = k * 32 + lane;
      if (idx < nv4) {
        g[k] = __ldg(gp + idx);
        const float4 t = __ldg(vp + idx);
        vh[k] = make_float4(t.x * inv, t.y * inv, t.z * inv, t.w * inv);
        dot = fmaf(g[k].x, vh[k].x, dot);
        dot = fmaf(g[k].y, vh[k].y, dot);
        dot = fmaf(g[k].z, vh[k].z, dot);
        dot = fmaf(g[k].w, vh[k].w, dot);
      }
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int idx = k * 32 + lane;
      if (idx < nv4)
        op[idx] = make_float4((g[k].x - dot * vh[k].x) * inv, (g[k].y - dot * vh[k].y) * inv, (g[k].z - dot * vh[k].z) * inv,
                              (g[k].w - dot * vh[k].w) * inv);
    }
  }
}

__global__ void maxabs_kernel(const float* __restrict__ g, int64_t ld, int rows, int cols, float* __restrict__ out) {
  __shared__ float scratch[32];
  float m = 0.f;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {        // one row per block pass: coalesced, no index division
    const float* src = g + (int64_t)r * ld;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) m = fmaxf(m, fabsf(__ldg(src + c)));
  }
  m = block_max(m, scratch);
  if (threadIdx.x == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));      // m >= 0: int order = float order
}

// g16[r, c] = fp16(g[r, c] * 2^e) with 2^e chosen so that the largest entry lands near 2^13; scale[1] = 2^-e
__global__ void scale_to_f16_kernel(const float* __restrict__ g, int64_t ld, int rows, int cols, float* __restrict__ scale,
                                    __half* __restrict__ out, int ld_out) {
  const float mx = scale[0];
  const float sc = (mx > 0.f) ? exp2f(floorf(log2f(8192.f / mx))) : 1.f;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) scale[1] = 1.f / sc;
  const int r = blockIdx.y;
  const float* src = g + (int64_t)r * ld;
  __half* dst = out + (int64_t)r * ld_out;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ld_out; c += gridDim.x * blockDim.x)
    dst[c] = __float2half_rn(c < cols ? __ldg(src + c) * sc : 0.f);
}

// x 2^e = hi + lo with fp16 hi, lo (about 22 significant bits): the operands of the error-compensated products.
// 2^e puts the largest entry near 2^12, so lo (2^-11 of hi) stays in the fp16 normal range for entries down to
// 2^-15 of the maximum; scale[0] = max|x| (maxabs_kernel), scale[1] = 2^-e
__global__ void split_f16_kernel(const float* __restrict__ g, int64_t ld, int rows, int cols, float* __restrict__ scale,
                                 __half* __restrict__ hi, __half* __restrict__ lo, int ld_out) {
  const float mx = scale[0];
  const float sc = (mx > 0.f) ? exp2f(floorf(log2f(4096.f / mx))) : 1.f;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) scale[1] = 1.f / sc;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) {
    const float* src = g + (int64_t)r * ld;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ld_out; c += gridDim.x * blockDim.x) {
      const float x = c < cols ? __ldg(src + c) * sc : 0.f;
      const __half h = __float2half_rn(x);
      hi[(int64_t)r * ld_out + c] = h;
      lo[(int64_t)r * ld_out + c] = __float2half_rn(x - __half2float(h));
    }
  }
}

// 16-byte variants for densely aligned matrices (cols, ld multiples of 4; ld_out a multiple of 8): one float4 per thread
// and step, a grid that fills the machine whatever the row length
__global__ void __launch_bounds__(256) maxabs4_kernel(const float* __restrict__ g, int64_t ld, int rows, int cols, float* __restrict__ out) {
  __shared__ float scratch[32];
  const int q = cols >> 2;
  const int64_t n = (int64_t)rows * q;
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / q;
    const int c = (int)(i - r * q) << 2;
    const float4 v = __ldg(reinterpret_cast<const float4*>(g + r * ld + c));
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  m = block_max(m, scratch);
  if (threadIdx.x == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}
__global__ void __launch_bounds__(256) split4_f16_kernel(const float* __restrict__ g, int64_t ld, int rows, int cols, float* __restrict__ scale,
                                                         __half* __restrict__ hi, __half* __restrict__ lo, int ld_out) {
  const float mx = scale[0];
  const float sc = (mx > 0.f) ? exp2f(floorf(log2f(4096.f / mx))) : 1.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) scale[1] = 1.f / sc;
  const int q = ld_out >> 2;
  const int64_t n = (int64_t)rows * q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / q;
    const int c = (int)(i - r * q) << 2;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < cols) v = __ldg(reinterpret_cast<const float4*>(g + r * ld + c));       // cols % 4 == 0: whole quads only
    const float x[4] = {v.x * sc, v.y * sc, v.z * sc, v.w * sc};
    __half h[4], l[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      h[u] = __float2half_rn(x[u]);
      l[u] = __float2half_rn(x[u] - __half2float(h[u]));
    }
    *reinterpret_cast<uint2*>(hi + r * ld_out + c) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(lo + r * ld_out + c) = *reinterpret_cast<const uint2*>(l);
  }
}

// merge the per-tile online-softmax partials of a row: (max, sum exp(. - max)) over nt column tiles
__global__ void ce_merge_partials_kernel(const float* __restrict__ pmax, const float* __restrict__ psum, int rows, int nt,
                                         float* __restrict__ rowmax, float* __restrict__ rowsum) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float m = -INFINITY;
  for (int k = lane; k < nt; k += 32) m = fmaxf(m, pmax[(int64_t)r * nt + k]);
  m = warp_max(m);
  float s = 0.f;
  for (int k = lane; k < nt; k += 32) {
    const float pm = pmax[(int64_t)r * nt + k];
    if (pm > -INFINITY) s += psum[(int64_t)r * nt + k] * expf(pm - m);
  }
  s = warp_sum(s);
  if (lane == 0) {
    rowmax[r] = (m == -INFINITY) ? 0.f : m;
    rowsum[r] = s;
  }
}

}  // namespace

int launch_norms(const float* x, int64_t s_vec, int64_t s_elem, int nvec, int len, float* norm, cudaStream_t st);   // dense_simt.cu

// ---- host API (used by dense_simt.cu's head entry points when precision == TGFR_PREC_TC) ------------------------
bool head_tc_supported(int B, int C, int Din) { return B >= 1 && C >= 1 && Din >= 8; }

// fp16 operand: `mn` = 0: memory [rows, K] (K contiguous, pitch ld);  1: memory [K, rows] (rows contiguous, pitch ld)
// `batch` samples lie ld * (mn ? K : rows) elements apart (contiguous samples)
static int operand_map(CUtensorMap* tm, const __half* ptr, int mn, int rows, int K, int64_t ld, int box_rows = 128,
                       bool overlap = false, int batch = 1) {
  if (mn)
    return make_tmap_3d(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptr, (uint64_t)rows, (uint64_t)K, (uint64_t)batch, 64, 64, 1, 128,
                        (uint64_t)ld, overlap);
  return make_tmap_3d(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptr, (uint64_t)K, (uint64_t)rows, (uint64_t)batch, 64, box_rows, 1, 128,
                      (uint64_t)ld, overlap);
}

template <int EPI, int BN, int STAGES, int MINB>
static int launch_gemm(dim3 grid, const CUtensorMap& tm_a, const CUtensorMap& tm_b, const GemmTcParams& p, cudaStream_t st);

static int sm_count();
// TGFR_GEMM_PERSIST=1 selects the persistent kernel.  Measured (B200, same runs, tools/time_head.py / time_imim.py):
// fused head step 143 us vs 117 us, IMIM fwd+bwd 1.41 ms vs 1.35 ms for the one-tile-per-CTA kernel -- at 2-8 tiles per
// SM these products are bound by their epilogues (160 exponentials per thread in the cross-entropy ones, 64 KB of
// stores in the plain one), and two or three co-resident CTAs bring 8-12 epilogue warps per SM where one persistent
// CTA has 4; the set-up it saves is smaller than that.  It stays selectable (and parity-tested) as the base for a
// version with two epilogue warpgroups; the default is the co-resident plan.
static bool gemm_persistent() {
  const char* e = getenv("TGFR_GEMM_PERSIST");
  return e && atoi(e) == 1;
}
template <int EPI, int BN, int STAGES>
static int launch_gemm_persist(dim3 tiles, const GemmMaps& maps, const GemmTcParams& p, cudaStream_t st) {
  constexpr uint32_t smem = gemm_persist_smem_bytes(BN, STAGES);
  static_assert(smem <= 232448, "persistent GEMM: shared memory plan");
  static bool attr_done[64] = {};
  int dev = 0;
  TGFR_CUDA_OK(cudaGetDevice(&dev));
  if (!attr_done[dev & 63]) {
    TGFR_CUDA_OK(cudaFuncSetAttribute(gemm_tc_persist_kernel<EPI, BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dev & 63] = true;
  }
  const int n = (int)(tiles.x * tiles.y * tiles.z), sms = sm_count();
  gemm_tc_persist_kernel<EPI, BN, STAGES><<<n < sms ? n : sms, kGemmThreads, smem, st>>>(maps, p, (int)tiles.x, (int)tiles.y, (int)tiles.z);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

template <int EPI, int BN, int STAGES, int MINB>
static int launch_gemm(dim3 grid, const GemmMaps& maps, const GemmTcParams& p, cudaStream_t st) {
  if (gemm_persistent()) return launch_gemm_persist<EPI, BN, (BN <= 128 ? 6 : 5)>(grid, maps, p, st);
  constexpr uint32_t smem = gemm_smem_bytes(BN, STAGES);
  static bool attr_done[64] = {};
  int dev = 0;
  TGFR_CUDA_OK(cudaGetDevice(&dev));
  if (!attr_done[dev & 63]) {
    TGFR_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<EPI, BN, STAGES, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dev & 63] = true;
  }
  gemm_tc_kernel<EPI, BN, STAGES, MINB><<<grid, kGemmThreads, smem, st>>>(maps, p);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

template <int EPI, int BN, int STAGES, int MINB>
static int launch_gemm(dim3 grid, const CUtensorMap& tm_a, const CUtensorMap& tm_b, const GemmTcParams& p, cudaStream_t st) {
  GemmMaps maps;
  memset(&maps, 0, sizeof(maps));
  maps.a[0][0] = tm_a;
  maps.b[0][0] = tm_b;
  return launch_gemm<EPI, BN, STAGES, MINB>(grid, maps, p, st);
}

static int sm_count() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 148;
}

// tile plan: 0 = 128 wide, 3 stages, 2 CTAs / SM;  1 = 160 wide (K-major B only);  2 = 128 wide, 2 stages, 3 CTAs / SM
// (plain-store epilogue only: its register count allows three residents).  TGFR_HEAD_TILE = 0 | 1 | 2 forces one.
static int pick_tile_plan(int M, int N, int splits, int b_mn, int cover, bool allow3) {
  const int sms = sm_count();
  const int mt = (M + kBM - 1) / kBM;
  const int t128 = ((N + 127) / 128) * mt * splits, t160 = ((N + 159) / 160) * mt * splits;
  const bool ok160 = !b_mn && ((N + 159) / 160) * 160 >= cover;
  if (const char* e = getenv("TGFR_HEAD_TILE")) {
    const int v = atoi(e);
    if (v == 1 && ok160) return 1;
    if (v == 2 && allow3) return 2;
    if (v == 0) return 0;
  }
  if (t128 <= 2 * sms) return 0;
  if (ok160 && t160 <= 2 * sms) return 1;
  if (allow3 && t128 <= 3 * sms) return 2;
  return 0;
}

int gemm_tc(const __half* A, int a_mn, int64_t lda, const __half* Bm, int b_mn, int64_t ldb, int M, int N, int K,
            float alpha, const float* dscale, int clamp, float* C, int64_t ldc, int splits, cudaStream_t st) {
  const int kt_total = (K + kBK - 1) / kBK;
  if (splits < 1) splits = 1;
  if (splits > kt_total) splits = kt_total;
  const int per = (kt_total + splits - 1) / splits;
  splits = (kt_total + per - 1) / per;                       // no empty split
  const int plan = pick_tile_plan(M, N, splits, b_mn, 0, true);
  const int bn = plan == 1 ? 160 : 128;
  CUtensorMap tm_a, tm_b;
  if (int rc = operand_map(&tm_a, A, a_mn, M, K, lda)) return rc;
  if (int rc = operand_map(&tm_b, Bm, b_mn, N, K, ldb, bn)) return rc;
  GemmTcParams p{};
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.kt_per_split = per; p.alpha = alpha; p.dscale = dscale; p.clamp = clamp;
  p.atomic = splits > 1; p.a_mn = a_mn; p.b_mn = b_mn;
  if (splits > 1) TGFR_CUDA_OK(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * ldc, st));
  const dim3 grid((N + bn - 1) / bn, (M + kBM - 1) / kBM, splits);
  if (plan == 1) return launch_gemm<kEpiStore, 160, 3, 2>(grid, tm_a, tm_b, p, st);
  if (plan == 2) return launch_gemm<kEpiStore, 128, 2, 3>(grid, tm_a, tm_b, p, st);
  return launch_gemm<kEpiStore, 128, 3, 2>(grid, tm_a, tm_b, p, st);
}

// g16[b, c] = fp16(2^e k (softmax - onehot) (phi' on the label column)) from the saved cos-theta matrix: the element-wise
// twin of the kEpiCeGrad epilogue (same arithmetic), eight columns per thread
__global__ void __launch_bounds__(256) ce_grad_from_cos_kernel(const float* __restrict__ cosm, int ld_cos, const int64_t* __restrict__ labels,
                                                               int class_off, int M, int N, float alpha, float cm, float sm, float th,
                                                               float mm, int easy, const float* __restrict__ lse,
                                                               const float* __restrict__ coef, const float* __restrict__ gout,
                                                               float* __restrict__ scale, __half* __restrict__ g16, int ld_g) {
  const float k = (coef ? __ldg(coef) : 1.f) * (gout ? __ldg(gout) : 1.f) / (float)M;
  const float ak = fabsf(k);
  const float sc = (ak > 0.f) ? exp2f(floorf(log2f(256.f / ak))) : 1.f;
  const float kscale = k * sc;
  if (blockIdx.x == 0 && threadIdx.x == 0) scale[1] = 1.f / sc;
  const int q = ld_g >> 3;
  const int64_t n = (int64_t)M * q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / q), c0 = (int)(i - (int64_t)row * q) << 3;
    const int64_t ycol = __ldg(labels + row) - class_off;
    const float ls = __ldg(lse + row);
    float c[8];
    if (c0 + 7 < ld_cos) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(cosm + (int64_t)row * ld_cos + c0));
      const float4 b = __ldg(reinterpret_cast<const float4*>(cosm + (int64_t)row * ld_cos + c0 + 4));
      c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
    } else {
#pragma unroll
      for (int u = 0; u < 8; ++u) c[u] = c0 + u < ld_cos ? __ldg(cosm + (int64_t)row * ld_cos + c0 + u) : 0.f;
    }
    uint32_t pk[4];
#pragma unroll
    for (int u = 0; u < 8; u += 2) {
      float g2[2];
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        float g = 0.f;
        const int col = c0 + u + v;
        if (col < N) {
          const bool on = (int64_t)col == ycol;
          float dph = 1.f;
          const float l = on ? alpha * arc_phi_tc(c[u + v], cm, sm, th, mm, easy, &dph) : alpha * c[u + v];
          g = kscale * (expf(l - ls) - (on ? 1.f : 0.f)) * (on ? dph : 1.f);
          g = fminf(fmaxf(g, -65504.f), 65504.f);
        }
        g2[v] = g;
      }
      const __half2 h2 = __floats2half2_rn(g2[0], g2[1]);
      pk[u >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
    }
    *reinterpret_cast<uint4*>(g16 + (int64_t)row * ld_g + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}
int arc_ce_grad_from_cos(const float* cosm, int ld_cos, int M, int N, float s, float m, int easy, const int64_t* labels,
                         int class_off, const float* lse, const float* coef, const float* gout, float* scale, __half* g16,
                         int ld_g, cudaStream_t st) {
  TGFR_REQUIRE((ld_g & 7) == 0 && (ld_cos & 3) == 0, "arc_ce_grad_from_cos: pitches must be multiples of 8 / 4");
  const float pi = 3.14159265358979323846f;
  const int64_t n = (int64_t)M * (ld_g >> 3);
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  ce_grad_from_cos_kernel<<<blocks, 256, 0, st>>>(cosm, ld_cos, labels, class_off, M, N, s, cosf(m), sinf(m), cosf(pi - m),
                                                  sinf(pi - m) * m, easy, lse, coef, gout, scale, g16, ld_g);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

// cos-theta GEMM (x16 [M, K] . w16 [N, K]^T, alpha = s) with a fused ArcFace cross-entropy epilogue.
//   grad == 0: rowmax / rowsum [M] (of this class shard), tgt, cos_t (owner rows only; others 0 / NaN);
//              part = 2 * M * ceil(N / 128) floats of scratch
//   grad != 0: g16 [M, ld_g] and scale[1] from lse / coef / gout
int gemm_tc_arc_ce(const __half* x16, int64_t ldx, const __half* w16, int64_t ldw, int M, int N, int K, float s, float m,
                   int easy, const int64_t* labels, int class_off, int grad, float* part, float* rowmax, float* rowsum,
                   float* tgt, float* cos_t, const float* lse, const float* coef, const float* gout, float* scale,
                   __half* g16, int ld_g, cudaStream_t st, float* cos_out, int ld_cos) {
  const int plan = pick_tile_plan(M, N, 1, 0, grad ? ld_g : 0, false);
  const int bn = plan == 1 ? 160 : 128;
  CUtensorMap tm_a, tm_b;
  if (int rc = operand_map(&tm_a, x16, 0, M, K, ldx)) return rc;
  if (int rc = operand_map(&tm_b, w16, 0, N, K, ldw, bn)) return rc;
  const float pi = 3.14159265358979323846f;
  const int nt = (N + bn - 1) / bn;
  GemmTcParams p{};
  p.M = M; p.N = N; p.K = K; p.kt_per_split = (K + kBK - 1) / kBK; p.alpha = s;
  p.ce.labels = labels; p.ce.class_off = class_off;
  p.ce.cm = cosf(m); p.ce.sm = sinf(m); p.ce.th = cosf(pi - m); p.ce.mm = sinf(pi - m) * m; p.ce.easy = easy;
  p.ce.pmax = part; p.ce.psum = part ? part + (size_t)M * nt : nullptr; p.ce.tgt = tgt; p.ce.cos_t = cos_t;
  p.ce.lse = lse; p.ce.coef = coef; p.ce.gout = gout; p.ce.scale = scale; p.ce.g16 = g16; p.ce.ld_g = ld_g;
  p.ce.cos_out = grad ? nullptr : cos_out; p.ce.ld_cos = ld_cos;
  const dim3 grid(nt, (M + kBM - 1) / kBM, 1);
  if (!grad) {
    TGFR_CUDA_OK(cudaMemsetAsync(tgt, 0, sizeof(float) * M, st));
    TGFR_CUDA_OK(cudaMemsetAsync(cos_t, 0xFF, sizeof(float) * M, st));          // NaN: label outside this shard
    if (int rc = plan == 1 ? launch_gemm<kEpiCeStats, 160, 3, 2>(grid, tm_a, tm_b, p, st)
                           : launch_gemm<kEpiCeStats, 128, 3, 2>(grid, tm_a, tm_b, p, st))
      return rc;
    ce_merge_partials_kernel<<<(M + 7) / 8, 256, 0, st>>>(p.ce.pmax, p.ce.psum, M, nt, rowmax, rowsum);
    TGFR_LAUNCH_OK();
  } else {
    if (int rc = plan == 1 ? launch_gemm<kEpiCeGrad, 160, 3, 2>(grid, tm_a, tm_b, p, st)
                           : launch_gemm<kEpiCeGrad, 128, 3, 2>(grid, tm_a, tm_b, p, st))
      return rc;
  }
  return TGFR_OK;
}

// out16[v, :] = x_v / max(norm[v], 1e-12) for nvec vectors of length len with element strides (s_vec, s_elem)
int head_normalize_f16(const float* x, int64_t s_vec, int64_t s_elem, int nvec, int len, const float* norm, __half* out,
                       int ld_out, cudaStream_t st) {
  if (s_vec == 1 && s_elem != 1) {        // vector index contiguous ([len, nvec] storage): tiled transpose
    const dim3 grid((nvec + 31) / 32, (ld_out + 31) / 32);
    normalize_cols_f16_kernel<<<grid, 256, 0, st>>>(x, s_elem, nvec, len, norm, out, ld_out);
  } else {
    normalize_rows_f16_kernel<<<(nvec + 7) / 8, 256, 0, st>>>(x, s_vec, s_elem, nvec, len, norm, out, ld_out);
  }
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

// norms and fp16 operands of x [B, Din] (row pitch x_sr) and w [C, Din] (strides w_sc, w_sk) for the cos-theta GEMM
int head_prepare_operands(const float* x, int64_t x_sr, const float* w, int64_t w_sc, int64_t w_sk, int B, int C, int Din,
                          float* xnorm, float* wnorm, __half* x16, __half* w16, int Dp, cudaStream_t st) {
  const bool fused = w_sk == 1 && (Din & 3) == 0 && Din <= 1024 && (x_sr & 3) == 0 && (w_sc & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15) == 0 && (Dp & 7) == 0;
  if (fused) {
    const int blocks = (B + C + 7) / 8;
    norm_f16_pair_kernel<<<blocks < 148 * 8 ? blocks : 148 * 8, 256, 0, st>>>(x, x_sr, B, w, w_sc, C, Din, xnorm, wnorm, x16,
                                                                             w16, Dp);
    TGFR_LAUNCH_OK();
    return TGFR_OK;
  }
  if (int rc = launch_norms(x, x_sr, 1, B, Din, xnorm, st)) return rc;
  if (int rc = launch_norms(w, w_sc, w_sk, C, Din, wnorm, st)) return rc;
  if (int rc = head_normalize_f16(x, x_sr, 1, B, Din, xnorm, x16, Dp, st)) return rc;
  return head_normalize_f16(w, w_sc, w_sk, C, Din, wnorm, w16, Dp, st);
}

// dx [B, Din] (contiguous) and dw (row pitch dw_sr) from the GEMM outputs gx [B, Din] / gw [C, Din]; gx / dx may be NULL.
// Returns 1 if the shapes do not fit the fused kernel (the caller then runs normalize_bwd_kernel per operand).
int head_normalize_bwd_pair(const float* gx, const float* x, int64_t x_sr, const float* xnorm, float* dx, int B, const float* gw,
                            const float* w, int64_t w_sc, int64_t w_sk, const float* wnorm, float* dw, int64_t dw_sc,
                            int64_t dw_sk, int C, int Din, cudaStream_t st) {
  const uintptr_t al = reinterpret_cast<uintptr_t>(gw) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(dw) |
                       (dx ? reinterpret_cast<uintptr_t>(gx) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) : 0);
  const bool ok = w_sk == 1 && dw_sk == 1 && (Din & 3) == 0 && Din <= 1024 && (x_sr & 3) == 0 && (w_sc & 3) == 0 &&
                  (dw_sc & 3) == 0 && (al & 15) == 0;
  if (!ok) return 1;
  const int nx = dx ? B : 0;
  const int blocks = (nx + C + 7) / 8;
  normalize_bwd_pair_kernel<<<blocks < 148 * 8 ? blocks : 148 * 8, 256, 0, st>>>(gx, x, x_sr, xnorm, dx, nx, gw, w, w_sc, wnorm, dw,
                                                                                dw_sc, C, Din, 1e-12f);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

// max|x| over several fp32 blocks into scale[0] (scale[0..1] zeroed first), then their hi / lo splits with ONE scale
int head_split_f16(int nblocks, const float* const* src, const int64_t* ld, const int* rows, const int* cols, float* scale,
                   __half* const* hi, __half* const* lo, const int* ld_out, cudaStream_t st, bool have_max) {
  if (!have_max) TGFR_CUDA_OK(cudaMemsetAsync(scale, 0, 2 * sizeof(float), st));
  auto vec_ok = [&](int k) {
    return (cols[k] & 3) == 0 && (ld[k] & 3) == 0 && (ld_out[k] & 7) == 0 && (reinterpret_cast<uintptr_t>(src[k]) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(hi[k]) & 7) == 0 && (reinterpret_cast<uintptr_t>(lo[k]) & 7) == 0;
  };
  auto blocks_for = [](int64_t quads) { return (int)(quads + 255) / 256 < 148 * 8 ? (int)((quads + 255) / 256) : 148 * 8; };
  for (int k = 0; k < nblocks && !have_max; ++k) {
    if (vec_ok(k))
      maxabs4_kernel<<<blocks_for((int64_t)rows[k] * (cols[k] >> 2)), 256, 0, st>>>(src[k], ld[k], rows[k], cols[k], scale);
    else
      maxabs_kernel<<<rows[k] < 1184 ? rows[k] : 1184, 256, 0, st>>>(src[k], ld[k], rows[k], cols[k], scale);
    TGFR_LAUNCH_OK();
  }
  for (int k = 0; k < nblocks; ++k) {
    if (vec_ok(k))
      split4_f16_kernel<<<blocks_for((int64_t)rows[k] * (ld_out[k] >> 2)), 256, 0, st>>>(src[k], ld[k], rows[k], cols[k], scale, hi[k],
                                                                                        lo[k], ld_out[k]);
    else
      split_f16_kernel<<<dim3((ld_out[k] + 1023) / 1024, rows[k] < 32768 ? rows[k] : 32768), 256, 0, st>>>(src[k], ld[k], rows[k], cols[k],
                                                                                                          scale, hi[k], lo[k], ld_out[k]);
    TGFR_LAUNCH_OK();
  }
  return TGFR_OK;
}

// up to three independent error-compensated products in one launch (blockIdx.z = product):
//   C_z [M_z, N_z] = relu?( alpha dscale[1] dscale2[1] (A_hi B_lo^T + A_lo B_hi^T + A_hi B_hi^T) + bias_z )
// a_mn / b_mn as in gemm_tc; a_overlap / b_overlap: the operand's rows share memory (pitch < row length)
struct TcOperand {          // an fp32 matrix as fp16 hi / lo (gemm_tc_split_operand); lo may be NULL for nterms == 1
  const __half *hi, *lo;
  int64_t ld;               // row pitch of the fp16 copies (multiple of 8)
  const float* scale;       // device float[2]: scale[1] = 2^-e undoes the power-of-two scaling of the copies
};
struct Split3Product {
  const __half *a_hi, *a_lo, *b_hi, *b_lo;
  int64_t lda, ldb, ldc;
  int M, N, K;
  float* C;
  const float* bias;
};
int gemm_tc_split3_batched(const Split3Product* prods, int n, int a_mn, int a_overlap, int b_mn, int b_overlap, float alpha,
                           const float* dscale, const float* dscale2, int relu, cudaStream_t st) {
  TGFR_REQUIRE(n >= 1 && n <= 3, "gemm_tc_split3_batched: 1..3 products");
  GemmMaps maps;
  memset(&maps, 0, sizeof(maps));
  GemmTcParams p{};
  p.alpha = alpha; p.dscale = dscale; p.dscale2 = dscale2; p.relu = relu; p.a_mn = a_mn; p.b_mn = b_mn;
  p.nterms = 3; p.batched = 1;
  int mt = 0, nt = 0;
  for (int z = 0; z < n; ++z) {
    const Split3Product& q = prods[z];
    if (int rc = operand_map(&maps.a[z][0], q.a_hi, a_mn, q.M, q.K, q.lda, 128, a_overlap != 0)) return rc;
    if (int rc = operand_map(&maps.a[z][1], q.a_lo, a_mn, q.M, q.K, q.lda, 128, a_overlap != 0)) return rc;
    if (int rc = operand_map(&maps.b[z][0], q.b_hi, b_mn, q.N, q.K, q.ldb, 128, b_overlap != 0)) return rc;
    if (int rc = operand_map(&maps.b[z][1], q.b_lo, b_mn, q.N, q.K, q.ldb, 128, b_overlap != 0)) return rc;
    p.z[z].C = q.C; p.z[z].bias = q.bias; p.z[z].ldc = q.ldc; p.z[z].M = q.M; p.z[z].N = q.N; p.z[z].K = q.K;
    mt = mt > (q.M + kBM - 1) / kBM ? mt : (q.M + kBM - 1) / kBM;
    nt = nt > (q.N + 127) / 128 ? nt : (q.N + 127) / 128;
  }
  return launch_gemm<kEpiStore, 128, 3, 2>(dim3(nt, mt, n), maps, p, st);
}

// General error-compensated product on fp16 hi / lo operand pairs (the contractions of IMIM, csrc/imim.cu):
//   mode 0  C = A B^T   A [M, K], B [N, K]          mode 1  C = A B   A [M, K], B [K, N]
//   mode 2  C = A^T B   A [K, M], B [K, N]
//   C [M, N] (pitch ldc) = relu?( alpha A.scale[1] B.scale[1] (sum of nterms products) + bias[col] )
// batch > 1: contiguous samples of both operands (ld x rows-in-memory apart), C advances by c_bs per sample, no split-K.
// splits > 1: split-K with vector reductions into a zeroed C (no bias / relu).
int gemm_tc_pair(int mode, const TcOperand& A, const TcOperand& B, float* C, int64_t ldc, int64_t c_bs, int M, int N, int K,
                 int batch, float alpha, const float* bias, int relu, int splits, int nterms, cudaStream_t st, float* amax) {
  TGFR_REQUIRE(mode >= 0 && mode <= 2 && (nterms == 1 || nterms == 3), "gemm_tc_pair: bad mode / nterms");
  TGFR_REQUIRE(batch >= 1 && (batch == 1 || splits <= 1), "gemm_tc_pair: a strided batch cannot split K");
  TGFR_REQUIRE(splits <= 1 || (!bias && !relu), "gemm_tc_pair: split-K has no bias / relu epilogue");
  const int a_mn = mode == 2, b_mn = mode != 0;
  const int kt_total = (K + kBK - 1) / kBK;
  if (splits < 1) splits = 1;
  if (splits > kt_total) splits = kt_total;
  const int per = (kt_total + splits - 1) / splits;
  splits = (kt_total + per - 1) / per;
  GemmMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (int rc = operand_map(&maps.a[0][0], A.hi, a_mn, M, K, A.ld, 128, false, batch)) return rc;
  if (int rc = operand_map(&maps.b[0][0], B.hi, b_mn, N, K, B.ld, 128, false, batch)) return rc;
  if (nterms == 3) {
    if (int rc = operand_map(&maps.a[0][1], A.lo, a_mn, M, K, A.ld, 128, false, batch)) return rc;
    if (int rc = operand_map(&maps.b[0][1], B.lo, b_mn, N, K, B.ld, 128, false, batch)) return rc;
  }
  GemmTcParams p{};
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.kt_per_split = per; p.alpha = alpha;
  p.dscale = A.scale; p.dscale2 = B.scale; p.bias = bias; p.relu = relu;
  p.atomic = splits > 1; p.a_mn = a_mn; p.b_mn = b_mn; p.nterms = nterms;
  p.strided = batch > 1; p.c_bs = c_bs;
  TGFR_REQUIRE(!amax || splits <= 1, "gemm_tc_pair: max |output| is not defined for split-K partial sums");
  p.amax = amax;
  if (splits > 1) TGFR_CUDA_OK(cudaMemset2DAsync(C, sizeof(float) * (size_t)ldc, 0, sizeof(float) * (size_t)N, (size_t)M, st));
  const dim3 grid((N + 127) / 128, (M + kBM - 1) / kBM, batch > 1 ? batch : splits);
  TGFR_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "gemm_tc_pair: %u row tiles x %u samples / splits exceed one launch grid",
               grid.y, grid.z);
  const int tiles = (int)(grid.x * grid.y * grid.z);
  if (tiles > 2 * sm_count() && tiles <= 3 * sm_count()) return launch_gemm<kEpiStore, 128, 2, 3>(grid, maps, p, st);
  return launch_gemm<kEpiStore, 128, 3, 2>(grid, maps, p, st);
}

// x [rows, cols] fp32 (pitch ld) -> hi / lo fp16 [rows, ld_out] with one power-of-two scale: scale[0] = max|x|, scale[1] = 2^-e
// have_max: scale[0] already holds max |x| (left there by the kernel that produced x: GemmTcParams::amax and the `amax`
// arguments of the element-wise kernels in nn_blocks.cuh), so the max-abs pass is skipped
int gemm_tc_split_operand(const float* src, int64_t ld, int rows, int cols, float* scale, __half* hi, __half* lo, int ld_out,
                          cudaStream_t st, bool have_max) {
  const float* srcs[1] = {src};
  const int64_t lds[1] = {ld};
  const int rws[1] = {rows}, cls[1] = {cols}, ldo[1] = {ld_out};
  __half* his[1] = {hi};
  __half* los[1] = {lo};
  return head_split_f16(1, srcs, lds, rws, cls, scale, his, los, ldo, st, have_max);
}

// Valid 3x3 convolution as ONE product (FCFM's conv, models/fusion_nets.py:235): img = channels-last hi / lo copies
// [rows = B * H * W positions, Cin] (pitch img.ld), w = hi / lo copies of the weights re-laid as [N, 9 * Cin] (tap-major),
// C [rows, N] (pitch ldc) = relu?(conv + bias) at every anchor position (the caller uses the anchors whose 3 x 3 window
// fits).  No im2col matrix exists anywhere: a tap is a row offset of the TMA box.
int gemm_tc_conv3x3(const TcOperand& img, const TcOperand& w, float* C, int64_t ldc, int rows, int N, int Cin, int width,
                    const float* bias, int relu, int nterms, cudaStream_t st) {
  TGFR_REQUIRE(Cin % kBK == 0 && N >= 1 && N <= 64 && (nterms == 1 || nterms == 3), "gemm_tc_conv3x3: Cin %% 64, N <= 64");
  GemmMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (int rc = operand_map(&maps.a[0][0], img.hi, 0, rows, Cin, img.ld, 128)) return rc;
  if (int rc = operand_map(&maps.b[0][0], w.hi, 0, N, 9 * Cin, w.ld, 64)) return rc;
  if (nterms == 3) {
    if (int rc = operand_map(&maps.a[0][1], img.lo, 0, rows, Cin, img.ld, 128)) return rc;
    if (int rc = operand_map(&maps.b[0][1], w.lo, 0, N, 9 * Cin, w.ld, 64)) return rc;
  }
  GemmTcParams p{};
  p.C = C; p.ldc = ldc; p.M = rows; p.N = N; p.K = 9 * Cin; p.kt_per_split = 9 * Cin / kBK; p.alpha = 1.f;
  p.dscale = img.scale; p.dscale2 = w.scale; p.bias = bias; p.relu = relu; p.nterms = nterms;
  p.conv_c = Cin; p.conv_w = width;
  const dim3 grid(1, (rows + kBM - 1) / kBM, 1);
  TGFR_REQUIRE(grid.y <= 65535, "gemm_tc_conv3x3: %u row tiles exceed one launch grid (chunk the batch)", grid.y);
  constexpr uint32_t smem = gemm_smem_bytes(64, 4);
  static bool attr_done[64] = {};
  int dev = 0;
  TGFR_CUDA_OK(cudaGetDevice(&dev));
  if (!attr_done[dev & 63]) {
    TGFR_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<kEpiStore, 64, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dev & 63] = true;
  }
  gemm_tc_kernel<kEpiStore, 64, 4, 2><<<grid, kGemmThreads, smem, st>>>(maps, p);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

// fp32 front end of gemm_tc_pair (C ABI tgfr_matmul_split): splits both operands into the workspace, then one product.
// Memory shapes per sample: A [M, K] (modes 0, 1) or [K, M] (mode 2); B [N, K] (mode 0) or [K, N] (modes 1, 2); samples of a
// batch are contiguous (pitch x rows apart).  C [batch, M, ldc].
static size_t matmul_split_plan(int mode, int M, int N, int K, int batch, int* ra, int* ca, int* rb, int* cb) {
  *ra = mode == 2 ? K : M; *ca = mode == 2 ? M : K;
  *rb = mode == 0 ? N : K; *cb = mode == 0 ? K : N;
  const size_t a = ((size_t)batch * *ra * ((*ca + 7) & ~7) * sizeof(__half) + 255) / 256 * 256;
  const size_t b = ((size_t)batch * *rb * ((*cb + 7) & ~7) * sizeof(__half) + 255) / 256 * 256;
  return 256 + 2 * a + 2 * b;
}
size_t matmul_split_workspace_bytes(int mode, int M, int N, int K, int batch) {
  int ra, ca, rb, cb;
  return matmul_split_plan(mode, M, N, K, batch, &ra, &ca, &rb, &cb);
}
int matmul_split(int mode, const float* A, int64_t lda, const float* Bm, int64_t ldb, float* C, int64_t ldc, int M, int N, int K,
                 int batch, float alpha, const float* bias, int relu, int splits, int nterms, void* ws, size_t ws_bytes,
                 cudaStream_t st) {
  TGFR_REQUIRE(mode >= 0 && mode <= 2 && M >= 1 && N >= 1 && K >= 1 && batch >= 1, "matmul_split: bad shape / mode");
  int ra, ca, rb, cb;
  const size_t need = matmul_split_plan(mode, M, N, K, batch, &ra, &ca, &rb, &cb);
  TGFR_REQUIRE(ws && ws_bytes >= need && (reinterpret_cast<uintptr_t>(ws) & 255) == 0,
               "matmul_split: workspace too small or not 256-byte aligned (%zu < %zu)", ws_bytes, need);
  const int lda16 = (ca + 7) & ~7, ldb16 = (cb + 7) & ~7;
  const size_t a = ((size_t)batch * ra * lda16 * sizeof(__half) + 255) / 256 * 256;
  const size_t b = ((size_t)batch * rb * ldb16 * sizeof(__half) + 255) / 256 * 256;
  uint8_t* base = reinterpret_cast<uint8_t*>(ws);
  float* scales = reinterpret_cast<float*>(base);
  __half* ahi = reinterpret_cast<__half*>(base + 256);
  __half* alo = reinterpret_cast<__half*>(base + 256 + a);
  __half* bhi = reinterpret_cast<__half*>(base + 256 + 2 * a);
  __half* blo = reinterpret_cast<__half*>(base + 256 + 2 * a + b);
  TGFR_REQUIRE(batch == 1 || (lda == ca && ldb == cb), "matmul_split: a batch needs densely packed operands");
  if (int rc = gemm_tc_split_operand(A, lda, batch * ra, ca, scales, ahi, alo, lda16, st, false)) return rc;
  if (int rc = gemm_tc_split_operand(Bm, ldb, batch * rb, cb, scales + 2, bhi, blo, ldb16, st, false)) return rc;
  const TcOperand oa{ahi, alo, lda16, scales}, ob{bhi, blo, ldb16, scales + 2};
  return gemm_tc_pair(mode, oa, ob, C, ldc, (int64_t)M * ldc, M, N, K, batch, alpha, bias, relu, splits, nterms, st, nullptr);
}

// g [rows, cols] fp32 (pitch ld) -> g16 [rows, ld_out] fp16 scaled by a power of two; scale[0] = max|g|, scale[1] = 1/2^e
int head_scale_f16(const float* g, int64_t ld, int rows, int cols, float* scale, __half* out, int ld_out, cudaStream_t st) {
  TGFR_CUDA_OK(cudaMemsetAsync(scale, 0, 2 * sizeof(float), st));
  maxabs_kernel<<<rows < 1184 ? rows : 1184, 256, 0, st>>>(g, ld, rows, cols, scale);
  TGFR_LAUNCH_OK();
  scale_to_f16_kernel<<<dim3((ld_out + 1023) / 1024, rows), 256, 0, st>>>(g, ld, rows, cols, scale, out, ld_out);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

}  // namespace tgfr

// Verification / identification scoring (reference utils/modules.py:40-88,150-166).
//
//   scores[i]  = cosine(out1[i], out2[i])                 utils/modules.py:150-151 (nn.CosineSimilarity(dim=1, eps=1e-6))
//   ROC counts = sklearn.metrics.roc_curve(y_true, y_score) utils/modules.py:54: scores sorted descending, one point per
//                DISTINCT score, tps = positives at or above it, fps = 1 + index - tps, collinear points dropped
//   ident      = argmax over each subject's row of scores  utils/modules.py:84-85
//
// All of it is HBM-bound integer / byte work (no contraction): coalesced 16-byte loads, shared-memory staging,
// warp-aggregated counters.  The ROC is exact integer arithmetic on the order-preserving bit pattern of the fp32
// scores, so its output is bit-identical to scikit-learn's for the same scores:
//   1. key = ~monotone(score) (ascending key order = descending score order), label byte = (label == 1)
//   2. LSD radix sort of (key, label byte): 4 passes of 8 bits; per pass a per-tile digit histogram, one exclusive
//      scan over [digit][tile], and a stable scatter (ranks from __match_any_sync, per-warp digit counters)
//   3. one point per distinct key: tile counts of (positives, group ends), scan, emit (threshold, fps, tps)
//   4. drop_intermediate: keep the end points and every point whose second difference of fps or tps is non-zero
//      (same three steps: count, scan, compact)
#include "common.cuh"

namespace tgfr {
namespace {

constexpr int kThreads = 256;
constexpr int kItems = 16;
constexpr int kTile = kThreads * kItems;      // 4096 keys per block
constexpr int kWarps = kThreads / 32;
constexpr int kSeg = kTile / kWarps;          // 512 consecutive keys per warp in the scatter

// ---------------------------------------------------------------------------------------------
// pairwise cosine: one warp per pair, rows held in registers (NV float4 per lane and vector)
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(kThreads) cosine_rows_vec_kernel(const float* __restrict__ x1, int64_t s1,
                                                                   const float* __restrict__ x2, int64_t s2, int64_t n,
                                                                   int D, float eps, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  const int nv = D >> 2;
  for (int64_t row = warp0; row < n; row += nwarps) {
    const float4* a = reinterpret_cast<const float4*>(x1 + row * s1);
    const float4* b = reinterpret_cast<const float4*>(x2 + row * s2);
    float4 va[NV], vb[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int idx = k * 32 + lane;
      va[k] = idx < nv ? __ldg(a + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      vb[k] = idx < nv ? __ldg(b + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      sa += va[k].x * va[k].x + va[k].y * va[k].y + va[k].z * va[k].z + va[k].w * va[k].w;
      sb += vb[k].x * vb[k].x + vb[k].y * vb[k].y + vb[k].z * vb[k].z + vb[k].w * vb[k].w;
    }
    const float na = fmaxf(sqrtf(warp_sum(sa)), eps), nb = fmaxf(sqrtf(warp_sum(sb)), eps);
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      dot += __fdiv_rn(va[k].x, na) * __fdiv_rn(vb[k].x, nb) + __fdiv_rn(va[k].y, na) * __fdiv_rn(vb[k].y, nb) +
             __fdiv_rn(va[k].z, na) * __fdiv_rn(vb[k].z, nb) + __fdiv_rn(va[k].w, na) * __fdiv_rn(vb[k].w, nb);
    }
    dot = warp_sum(dot);
    if (lane == 0) out[row] = dot;
  }
}

// any D, any alignment, element strides: two passes over the rows (the second hits L1 / L2)
__global__ void __launch_bounds__(kThreads) cosine_rows_generic_kernel(const float* __restrict__ x1, int64_t s1, int64_t e1,
                                                                       const float* __restrict__ x2, int64_t s2, int64_t e2,
                                                                       int64_t n, int D, float eps, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  for (int64_t row = warp0; row < n; row += nwarps) {
    const float* a = x1 + row * s1;
    const float* b = x2 + row * s2;
    float sa = 0.f, sb = 0.f;
    for (int k = lane; k < D; k += 32) {
      const float u = __ldg(a + k * e1), v = __ldg(b + k * e2);
      sa += u * u;
      sb += v * v;
    }
    const float na = fmaxf(sqrtf(warp_sum(sa)), eps), nb = fmaxf(sqrtf(warp_sum(sb)), eps);
    float dot = 0.f;
    for (int k = lane; k < D; k += 32) dot += __fdiv_rn(__ldg(a + k * e1), na) * __fdiv_rn(__ldg(b + k * e2), nb);
    dot = warp_sum(dot);
    if (lane == 0) out[row] = dot;
  }
}

// ---------------------------------------------------------------------------------------------
// identification: first index of the row maximum (np.argmax: a NaN is the maximum)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool better(float a, int ia, float b, int ib) {
  const bool an = a != a, bn = b != b;
  if (an != bn) return an;
  if (an || a == b) return ia < ib;
  return a > b;
}
__global__ void __launch_bounds__(kThreads) row_argmax_kernel(const float* __restrict__ x, int64_t sr, int rows, int cols,
                                                              int64_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  for (int64_t r = warp0; r < rows; r += nwarps) {
    const float* src = x + r * sr;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = lane; c < cols; c += 32) {
      const float v = __ldg(src + c);
      if (bi == 0x7fffffff || better(v, c, best, bi)) {
        best = v;
        bi = c;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || better(ov, oi, best, bi))) {
        best = ov;
        bi = oi;
      }
    }
    if (lane == 0) out[r] = bi;
  }
}

// ---------------------------------------------------------------------------------------------
// ROC: keys
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t desc_key(float f) {
  uint32_t u = __float_as_uint(f);
  if (u == 0x80000000u) u = 0u;                                      // -0.0 and +0.0 are one threshold
  const uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // unsigned order = float order
  return ~asc;                                                       // ascending keys = descending scores
}
__device__ __forceinline__ float key_score(uint32_t k) {
  const uint32_t asc = ~k;
  return __uint_as_float((asc & 0x80000000u) ? (asc & 0x7fffffffu) : ~asc);
}

__global__ void __launch_bounds__(kThreads) roc_keys_kernel(const float* __restrict__ scores, const int64_t* __restrict__ labels,
                                                            int64_t n, uint32_t* __restrict__ keys, uint8_t* __restrict__ lab,
                                                            uint32_t* __restrict__ info) {
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const float s = __ldg(scores + i);
    bad |= (s != s);
    keys[i] = desc_key(s);
    lab[i] = __ldg(labels + i) == 1 ? 1 : 0;
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(info + 2, 1u);
}

// ---------------------------------------------------------------------------------------------
// ROC: radix sort pass = histogram, scan, stable scatter
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift,
                                                              uint32_t* __restrict__ hist, int nblk) {
  __shared__ uint32_t h[257];
  const int lane = threadIdx.x & 31;
  for (int k = threadIdx.x; k < 257; k += kThreads) h[k] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kTile;
#pragma unroll 4
  for (int k = threadIdx.x; k < kTile; k += kThreads) {
    const int64_t i = base + k;
    const uint32_t d = i < n ? ((__ldg(keys + i) >> shift) & 255u) : 256u;
    // scores of one sign share their top digits: one add per warp then; mixed digits go to (mostly) distinct counters
    if (__all_sync(0xffffffffu, d == __shfl_sync(0xffffffffu, d, 0))) {
      if (lane == 0) atomicAdd(&h[d], 32u);
    } else {
      atomicAdd(&h[d], 1u);
    }
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];     // [digit][tile]
}

// exclusive block scan of one value per thread (kThreads threads); *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* wsum, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  uint32_t woff = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    const uint32_t c = wsum[w];
    if (w < warp) woff += c;
    tot += c;
  }
  __syncthreads();
  *total = tot;
  return woff + inc - v;
}

// in-place exclusive scan of a[0, len) by ONE block; for two interleaved sequences (stride 2) pass pairs = 1: even
// and odd entries are scanned independently.  Used directly for short arrays and for the chunk sums of long ones.
__global__ void __launch_bounds__(kThreads) scan_excl_kernel(uint32_t* __restrict__ a, int64_t len, int pairs,
                                                             uint32_t* __restrict__ totals) {
  __shared__ uint32_t wsum[kWarps];
  const int nseq = pairs ? 2 : 1;
  for (int q = 0; q < nseq; ++q) {
    uint32_t carry = 0;
    for (int64_t base = 0; base < len; base += kThreads * 4) {
      uint32_t v[4];
      uint32_t s = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t i = base + (int64_t)threadIdx.x * 4 + j;
        v[j] = i < len ? a[i * nseq + q] : 0u;
        s += v[j];
      }
      uint32_t tot;
      uint32_t off = carry + block_excl_scan(s, wsum, &tot);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t i = base + (int64_t)threadIdx.x * 4 + j;
        if (i < len) a[i * nseq + q] = off;
        off += v[j];
      }
      carry += tot;
    }
    if (totals && threadIdx.x == 0) totals[q] = carry;
  }
}

// long arrays (the [digit][tile] histogram of a big sort): chunk sums, scan of the sums (above), chunk scans
constexpr int kScanChunk = kThreads * 16;
__global__ void __launch_bounds__(kThreads) scan_chunk_sums_kernel(const uint32_t* __restrict__ a, int64_t len,
                                                                   uint32_t* __restrict__ sums) {
  __shared__ uint32_t wsum[kWarps];
  const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * 16;
  uint32_t s = 0;
  if (base + 16 <= len) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(a + base) + v);
      s += q.x + q.y + q.z + q.w;
    }
  } else {
    for (int j = 0; j < 16; ++j)
      if (base + j < len) s += a[base + j];
  }
  uint32_t tot;
  block_excl_scan(s, wsum, &tot);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(kThreads) scan_chunks_kernel(uint32_t* __restrict__ a, int64_t len,
                                                               const uint32_t* __restrict__ sums) {
  __shared__ uint32_t wsum[kWarps];
  const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * 16;
  uint32_t v[16];
  uint32_t s = 0;
  const bool full = base + 16 <= len;
  if (full) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint4 q = *(reinterpret_cast<const uint4*>(a + base) + k);
      v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = base + j < len ? a[base + j] : 0u;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) s += v[j];
  uint32_t tot;
  uint32_t off = __ldg(sums + blockIdx.x) + block_excl_scan(s, wsum, &tot);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const uint32_t x = v[j];
    v[j] = off;
    off += x;
  }
  if (full) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      *(reinterpret_cast<uint4*>(a + base) + k) = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (base + j < len) a[base + j] = v[j];
  }
}

// stable scatter of one tile.  Ranks come from __match_any_sync and per-warp digit counters; the tile is first put in
// digit order in shared memory so that the global writes are runs of consecutive addresses (one run per digit)
// instead of 4-byte scatters.
__global__ void __launch_bounds__(kThreads) radix_scatter_kernel(const uint32_t* __restrict__ kin, const uint8_t* __restrict__ lin,
                                                                 uint32_t* __restrict__ kout, uint8_t* __restrict__ lout,
                                                                 int64_t n, int shift, const uint32_t* __restrict__ hist,
                                                                 int nblk) {
  __shared__ uint32_t cnt[kWarps][257];
  __shared__ uint32_t lstart[256], gstart[256], wsum[kWarps];
  __shared__ uint32_t skeys[kTile];
  __shared__ uint8_t slab[kTile];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < kWarps * 257; k += kThreads) (&cnt[0][0])[k] = 0;
  __syncthreads();
  const int64_t tbase = (int64_t)blockIdx.x * kTile;
  const int64_t wbase = tbase + (int64_t)warp * kSeg;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t key[kSeg / 32];
  uint16_t rank[kSeg / 32];
  uint8_t lb[kSeg / 32];
#pragma unroll
  for (int r = 0; r < kSeg / 32; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    key[r] = i < n ? __ldg(kin + i) : 0u;
    lb[r] = i < n ? __ldg(lin + i) : (uint8_t)0;
    const uint32_t d = i < n ? ((key[r] >> shift) & 255u) : 256u;
    const uint32_t mask = __match_any_sync(0xffffffffu, d);
    const uint32_t before = cnt[warp][d];
    __syncwarp();
    if (lane == __ffs(mask) - 1) cnt[warp][d] = before + (uint32_t)__popc(mask);
    __syncwarp();
    rank[r] = (uint16_t)(before + (uint32_t)__popc(mask & lt));      // stable: earlier keys of the digit come first
  }
  __syncthreads();
  {   // thread d: each warp's offset inside digit d, digit d's start inside the tile and in the output
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const uint32_t c = cnt[w][d];
      cnt[w][d] = run;
      run += c;
    }
    uint32_t tot;
    lstart[d] = block_excl_scan(run, wsum, &tot);
    gstart[d] = __ldg(hist + (int64_t)d * nblk + blockIdx.x);
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSeg / 32; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    if (i < n) {
      const uint32_t d = (key[r] >> shift) & 255u;
      const uint32_t slot = lstart[d] + cnt[warp][d] + rank[r];
      skeys[slot] = key[r];
      slab[slot] = lb[r];
    }
  }
  __syncthreads();
  const int valid = (int)min((int64_t)kTile, n - tbase);
  for (int slot = threadIdx.x; slot < valid; slot += kThreads) {
    const uint32_t k = skeys[slot];
    const uint32_t d = (k >> shift) & 255u;
    const uint32_t pos = gstart[d] + (uint32_t)slot - lstart[d];
    kout[pos] = k;
    lout[pos] = slab[slot];
  }
}

// ---------------------------------------------------------------------------------------------
// ROC: one point per distinct score
// ---------------------------------------------------------------------------------------------
// thread t of a tile owns keys [base + 16 t, base + 16 t + 16): returns (positives << 16 | group ends) and the
// per-item bits (bit j: label, bit 16 + j: key j is the last of its group)
__device__ __forceinline__ uint32_t curve_thread_counts(const uint32_t* __restrict__ keys, const uint8_t* __restrict__ lab,
                                                        int64_t n, int64_t i0, uint32_t* bits, uint32_t* mykeys) {
  uint32_t b = 0, npos = 0, nend = 0;
  if (i0 < n) {
    const int cnt = (int)min((int64_t)kItems, n - i0);
    uint32_t k[kItems + 1];
    uint8_t l[kItems];
    if (cnt == kItems) {
#pragma unroll
      for (int v = 0; v < kItems / 4; ++v) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(keys + i0) + v);
        k[4 * v] = q.x; k[4 * v + 1] = q.y; k[4 * v + 2] = q.z; k[4 * v + 3] = q.w;
      }
      const uint4 lq = __ldg(reinterpret_cast<const uint4*>(lab + i0));
      const uint32_t lw[4] = {lq.x, lq.y, lq.z, lq.w};
#pragma unroll
      for (int j = 0; j < kItems; ++j) l[j] = (uint8_t)((lw[j >> 2] >> (8 * (j & 3))) & 0xffu);
    } else {
#pragma unroll
      for (int j = 0; j < kItems; ++j) {
        k[j] = j < cnt ? __ldg(keys + i0 + j) : 0u;
        l[j] = j < cnt ? __ldg(lab + i0 + j) : (uint8_t)0;
      }
    }
    const bool has_next = i0 + cnt < n;
    k[kItems] = has_next ? __ldg(keys + i0 + cnt) : 0u;
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
      if (j < cnt) {
        const uint32_t nxt = (j + 1 < cnt) ? k[j + 1] : k[kItems];
        const bool end = (j + 1 == cnt && !has_next) || nxt != k[j];
        b |= (uint32_t)(l[j] != 0) << j;
        b |= (uint32_t)end << (16 + j);
        npos += l[j] != 0;
        nend += end;
      }
      if (mykeys) mykeys[j] = k[j];
    }
  }
  *bits = b;
  return (npos << 16) | nend;
}

__global__ void __launch_bounds__(kThreads) curve_count_kernel(const uint32_t* __restrict__ keys, const uint8_t* __restrict__ lab,
                                                               int64_t n, uint32_t* __restrict__ tile_counts) {
  __shared__ uint32_t wsum[kWarps];
  uint32_t bits;
  const int64_t i0 = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
  const uint32_t c = curve_thread_counts(keys, lab, n, i0, &bits, nullptr);
  uint32_t tot;
  block_excl_scan(c, wsum, &tot);
  if (threadIdx.x == 0) {
    tile_counts[2 * blockIdx.x] = tot >> 16;           // positives in the tile
    tile_counts[2 * blockIdx.x + 1] = tot & 0xffffu;   // groups that end in the tile
  }
}

__global__ void __launch_bounds__(kThreads) curve_emit_kernel(const uint32_t* __restrict__ keys, const uint8_t* __restrict__ lab,
                                                              int64_t n, const uint32_t* __restrict__ tile_offs,
                                                              float* __restrict__ thr, int64_t* __restrict__ fps,
                                                              int64_t* __restrict__ tps) {
  __shared__ uint32_t wsum[kWarps];
  uint32_t bits, k[kItems];
  const int64_t i0 = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
  const uint32_t c = curve_thread_counts(keys, lab, n, i0, &bits, k);
  uint32_t tot;
  const uint32_t off = block_excl_scan(c, wsum, &tot);
  int64_t pos = (int64_t)tile_offs[2 * blockIdx.x] + (off >> 16);
  int64_t j = (int64_t)tile_offs[2 * blockIdx.x + 1] + (off & 0xffffu);
#pragma unroll
  for (int q = 0; q < kItems; ++q) {
    pos += (bits >> q) & 1u;
    if ((bits >> (16 + q)) & 1u) {
      thr[j] = key_score(k[q]);
      tps[j] = pos;
      fps[j] = i0 + q + 1 - pos;
      ++j;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// ROC: drop_intermediate (sklearn roc_curve: np.diff(fps, 2) | np.diff(tps, 2), end points always kept)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool keep_point(const int64_t* __restrict__ fps, const int64_t* __restrict__ tps, int64_t j,
                                           int64_t m) {
  if (m <= 2 || j == 0 || j == m - 1) return true;
  const int64_t d2f = fps[j + 1] - 2 * fps[j] + fps[j - 1];
  const int64_t d2t = tps[j + 1] - 2 * tps[j] + tps[j - 1];
  return d2f != 0 || d2t != 0;
}
__global__ void __launch_bounds__(kThreads) drop_count_kernel(const int64_t* __restrict__ fps, const int64_t* __restrict__ tps,
                                                              const uint32_t* __restrict__ info, uint32_t* __restrict__ tile_counts) {
  __shared__ uint32_t wsum[kWarps];
  const int64_t m = info[1];
  const int64_t j0 = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
  uint32_t c = 0;
  for (int q = 0; q < kItems; ++q)
    if (j0 + q < m) c += keep_point(fps, tps, j0 + q, m);
  uint32_t tot;
  block_excl_scan(c, wsum, &tot);
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(kThreads) drop_emit_kernel(const float* __restrict__ thr, const int64_t* __restrict__ fps,
                                                             const int64_t* __restrict__ tps, const uint32_t* __restrict__ info,
                                                             const uint32_t* __restrict__ tile_offs, float* __restrict__ thr_out,
                                                             int64_t* __restrict__ fps_out, int64_t* __restrict__ tps_out) {
  __shared__ uint32_t wsum[kWarps];
  const int64_t m = info[1];
  const int64_t j0 = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
  uint32_t c = 0, bits = 0;
  for (int q = 0; q < kItems; ++q)
    if (j0 + q < m && keep_point(fps, tps, j0 + q, m)) {
      bits |= 1u << q;
      ++c;
    }
  uint32_t tot;
  int64_t o = (int64_t)tile_offs[blockIdx.x] + block_excl_scan(c, wsum, &tot);
  for (int q = 0; q < kItems; ++q)
    if ((bits >> q) & 1u) {
      thr_out[o] = thr[j0 + q];
      fps_out[o] = fps[j0 + q];
      tps_out[o] = tps[j0 + q];
      ++o;
    }
}

__global__ void roc_finish_kernel(const uint32_t* __restrict__ info, int64_t* __restrict__ count_out) {
  count_out[0] = info[0];      // points returned
  count_out[1] = info[2];      // 1: a score was NaN (the curve is meaningless; the caller raises)
  count_out[2] = info[1];      // distinct scores (points before drop_intermediate)
}

struct RocPlan {
  int nblk;
  size_t keys0, keys1, lab0, lab1, hist, sums, tiles, info, thr, fps, tps, total;
};
RocPlan roc_plan(int64_t n) {
  RocPlan p;
  p.nblk = (int)((n + kTile - 1) / kTile);
  if (p.nblk < 1) p.nblk = 1;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t at = o;
    o += align_up(bytes, 256);
    return at;
  };
  const size_t nn = (size_t)(n > 0 ? n : 1);
  p.keys0 = take(nn * 4 + 64);
  p.keys1 = take(nn * 4 + 64);
  p.lab0 = take(nn + 64);
  p.lab1 = take(nn + 64);
  p.hist = take((size_t)256 * p.nblk * 4);
  p.sums = take((((size_t)256 * p.nblk + kScanChunk - 1) / kScanChunk) * 4);
  p.tiles = take((size_t)2 * p.nblk * 4);
  p.info = take(64);
  p.thr = take(nn * 4);
  p.fps = take(nn * 8);
  p.tps = take(nn * 8);
  p.total = o;
  return p;
}

int stream_grid(int64_t work_items, int per_block) {
  int64_t g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > 148 * 16) g = 148 * 16;
  return (int)g;
}

}  // namespace

int pair_cosine(const float* x1, int64_t s1r, int64_t s1d, const float* x2, int64_t s2r, int64_t s2d, int64_t n, int D,
                float eps, float* out, cudaStream_t st) {
  if (n == 0) return TGFR_OK;
  const int grid = stream_grid(n, kWarps);
  const bool vec = s1d == 1 && s2d == 1 && (D & 3) == 0 && D <= 1024 && (s1r & 3) == 0 && (s2r & 3) == 0 &&
                   (reinterpret_cast<uintptr_t>(x1) & 15) == 0 && (reinterpret_cast<uintptr_t>(x2) & 15) == 0;
  if (vec) {
    const int nv = (D / 4 + 31) / 32;
    switch (nv) {
#define TGFR_COS_CASE(NV)                                                                              \
  case NV:                                                                                             \
    cosine_rows_vec_kernel<NV><<<grid, kThreads, 0, st>>>(x1, s1r, x2, s2r, n, D, eps, out);            \
    break;
      TGFR_COS_CASE(1) TGFR_COS_CASE(2) TGFR_COS_CASE(3) TGFR_COS_CASE(4)
      TGFR_COS_CASE(5) TGFR_COS_CASE(6) TGFR_COS_CASE(7) TGFR_COS_CASE(8)
#undef TGFR_COS_CASE
      default:
        set_error("pair_cosine: internal (nv = %d)", nv);
        return TGFR_E_INVALID;
    }
  } else {
    cosine_rows_generic_kernel<<<grid, kThreads, 0, st>>>(x1, s1r, s1d, x2, s2r, s2d, n, D, eps, out);
  }
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int row_argmax(const float* x, int64_t sr, int rows, int cols, int64_t* out, cudaStream_t st) {
  if (rows == 0) return TGFR_OK;
  row_argmax_kernel<<<stream_grid(rows, kWarps), kThreads, 0, st>>>(x, sr, rows, cols, out);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

size_t roc_workspace_bytes(int64_t n) { return roc_plan(n).total; }

int roc_curve(const float* scores, const int64_t* labels, int64_t n, int drop_intermediate, float* thr_out, int64_t* fps_out,
              int64_t* tps_out, int64_t* count_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  const RocPlan pl = roc_plan(n);
  if (ws_bytes < pl.total) {
    set_error("roc_curve: workspace %zu < %zu bytes", ws_bytes, pl.total);
    return TGFR_E_WORKSPACE;
  }
  uint8_t* base = static_cast<uint8_t*>(ws);
  uint32_t* keys[2] = {reinterpret_cast<uint32_t*>(base + pl.keys0), reinterpret_cast<uint32_t*>(base + pl.keys1)};
  uint8_t* lab[2] = {base + pl.lab0, base + pl.lab1};
  uint32_t* hist = reinterpret_cast<uint32_t*>(base + pl.hist);
  uint32_t* tiles = reinterpret_cast<uint32_t*>(base + pl.tiles);
  uint32_t* sums = reinterpret_cast<uint32_t*>(base + pl.sums);
  uint32_t* info = reinterpret_cast<uint32_t*>(base + pl.info);        // [0] points out, [1] distinct scores, [2] NaN flag
  TGFR_CUDA_OK(cudaMemsetAsync(info, 0, 64, st));
  if (n == 0) {
    roc_finish_kernel<<<1, 1, 0, st>>>(info, count_out);
    TGFR_LAUNCH_OK();
    return TGFR_OK;
  }
  const int nblk = pl.nblk;
  roc_keys_kernel<<<stream_grid(n, kThreads * 4), kThreads, 0, st>>>(scores, labels, n, keys[0], lab[0], info);
  TGFR_LAUNCH_OK();
  int cur = 0;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 8 * pass;
    radix_hist_kernel<<<nblk, kThreads, 0, st>>>(keys[cur], n, shift, hist, nblk);
    TGFR_LAUNCH_OK();
    const int64_t hlen = (int64_t)256 * nblk;
    if (hlen <= 4 * kScanChunk) {
      scan_excl_kernel<<<1, kThreads, 0, st>>>(hist, hlen, 0, nullptr);
      TGFR_LAUNCH_OK();
    } else {
      const int nchunks = (int)((hlen + kScanChunk - 1) / kScanChunk);
      scan_chunk_sums_kernel<<<nchunks, kThreads, 0, st>>>(hist, hlen, sums);
      TGFR_LAUNCH_OK();
      scan_excl_kernel<<<1, kThreads, 0, st>>>(sums, nchunks, 0, nullptr);
      TGFR_LAUNCH_OK();
      scan_chunks_kernel<<<nchunks, kThreads, 0, st>>>(hist, hlen, sums);
      TGFR_LAUNCH_OK();
    }
    radix_scatter_kernel<<<nblk, kThreads, 0, st>>>(keys[cur], lab[cur], keys[cur ^ 1], lab[cur ^ 1], n, shift, hist, nblk);
    TGFR_LAUNCH_OK();
    cur ^= 1;
  }
  float* thr = drop_intermediate ? reinterpret_cast<float*>(base + pl.thr) : thr_out;
  int64_t* fps = drop_intermediate ? reinterpret_cast<int64_t*>(base + pl.fps) : fps_out;
  int64_t* tps = drop_intermediate ? reinterpret_cast<int64_t*>(base + pl.tps) : tps_out;
  curve_count_kernel<<<nblk, kThreads, 0, st>>>(keys[cur], lab[cur], n, tiles);
  TGFR_LAUNCH_OK();
  scan_excl_kernel<<<1, kThreads, 0, st>>>(tiles, nblk, 1, info + 4);    // info[4] positives, info[5] distinct scores
  TGFR_LAUNCH_OK();
  curve_emit_kernel<<<nblk, kThreads, 0, st>>>(keys[cur], lab[cur], n, tiles, thr, fps, tps);
  TGFR_LAUNCH_OK();
  TGFR_CUDA_OK(cudaMemcpyAsync(info + 1, info + 5, 4, cudaMemcpyDeviceToDevice, st));
  if (drop_intermediate) {
    drop_count_kernel<<<nblk, kThreads, 0, st>>>(fps, tps, info, tiles);
    TGFR_LAUNCH_OK();
    scan_excl_kernel<<<1, kThreads, 0, st>>>(tiles, nblk, 0, info);
    TGFR_LAUNCH_OK();
    drop_emit_kernel<<<nblk, kThreads, 0, st>>>(thr, fps, tps, info, tiles, thr_out, fps_out, tps_out);
    TGFR_LAUNCH_OK();
  } else {
    TGFR_CUDA_OK(cudaMemcpyAsync(info, info + 1, 4, cudaMemcpyDeviceToDevice, st));
  }
  roc_finish_kernel<<<1, 1, 0, st>>>(info, count_out);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

}  // namespace tgfr

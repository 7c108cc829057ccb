// Self-tests of the tcgen05 / TMA building blocks (tc.cuh), callable through the C ABI so that
// the GPU test-suite pins the descriptor encodings independently of the fused kernels.
#include "common.cuh"
#include "tc.cuh"

namespace tgfr {
namespace {

using namespace tc;

// D[128, N] = A[128, K] * B[N, K]^T on one CTA.  Operand layouts in GLOBAL memory (fp16):
//   a_mn == 0: A_g[128][K] (K contiguous)      a_mn == 1: A_g[K][128] (M contiguous)
//   b_mn == 0: B_g[N][K]                        b_mn == 1: B_g[K][N]
//   b_mn == 2: B_g[K/8][N][8]  planes of 8-half K chunks, brought in by ONE bulk copy, read K-major without swizzle
//   b_mn == 3: B_g[N/8][K][8]  planes of 8-half N chunks, one bulk copy, read MN-major without swizzle
// manual_a (needs a_mn == 1, K <= 128): A is written to shared memory by the threads with the
// hand-computed 128B swizzle instead of by TMA (the pattern the fused kernels use for E / dS).
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __half* __restrict__ a_g, float* __restrict__ out, int N, int K, int a_mn, int b_mn,
                     const __half* __restrict__ b_g, int manual_a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = base;                    // up to 128 x 256 halfs = 64 KB
  uint8_t* sb = base + 65536;            // up to 256 x 256 halfs = 128 KB
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  const uint32_t a_bytes = 128u * K * 2u, b_bytes = (uint32_t)N * K * 2u;
  if (manual_a == 2) {
    // A operand in TMEM: lane = row, 32-bit column j holds halfs (2j, 2j+1) of that row (a_g is [128][K])
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a_g + (size_t)tid * K);
    for (int c0 = 0; c0 < K / 2; c0 += 16) {
      uint32_t v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = src[c0 + j];
      tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c0, v);
    }
    tmem_st_wait();
    tc_fence_before();
  } else if (manual_a) {
    // thread k owns row k of A_g[K][128]: two 64-half panels of K rows each
    if (tid < K) {
      const uint4* src = reinterpret_cast<const uint4*>(a_g + (size_t)tid * 128);
      for (int p = 0; p < 2; ++p)
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(sa + p * (K * 128) + sw128_offset(tid, c)) = src[p * 8 + c];
    }
    fence_proxy_async();
  }
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    mbar_arrive_expect_tx(&bar_load, (manual_a ? 0u : a_bytes) + b_bytes);
    if (!manual_a) {
      if (!a_mn) for (int kc = 0; kc < K / 64; ++kc) tma_load_3d(sa + kc * (128 * 128), &tmap_a, &bar_load, kc * 64, 0, 0);
      else       for (int p = 0; p < 2; ++p)         tma_load_3d(sa + p * (K * 128), &tmap_a, &bar_load, p * 64, 0, 0);
    }
    if (b_mn >= 2) bulk_load(sb, b_g, b_bytes, &bar_load);
    else if (!b_mn) for (int kc = 0; kc < K / 64; ++kc) tma_load_3d(sb + kc * (N * 128), &tmap_b, &bar_load, kc * 64, 0, 0);
    else       for (int p = 0; p < N / 64; ++p)     tma_load_3d(sb + p * (K * 128), &tmap_b, &bar_load, p * 64, 0, 0);
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, N, a_mn != 0, b_mn == 1 || b_mn == 3);
    for (int k16 = 0; k16 < K / 16; ++k16) {
      uint64_t ad, bd;
      if (!a_mn) ad = make_smem_desc(smem_u32(sa) + (k16 >> 2) * (128 * 128) + (k16 & 3) * 32, 16, 1024);
      else       ad = make_smem_desc(smem_u32(sa) + k16 * 2048, K * 128, 1024);
      if (b_mn == 2) {
        bd = make_smem_desc_ns(smem_u32(sb) + k16 * 2 * (N * 16), N * 16, 128);
      } else if (b_mn == 3) {
        bd = make_smem_desc_ns(smem_u32(sb) + k16 * 256, 128, K * 16);
      } else if (!b_mn) bd = make_smem_desc(smem_u32(sb) + (k16 >> 2) * (N * 128) + (k16 & 3) * 32, 16, 1024);
      else       bd = make_smem_desc(smem_u32(sb) + k16 * 2048, K * 128, 1024);
      if (manual_a == 2) umma_ts(tmem, tmem + 256 + 8 * k16, bd, idesc, k16 > 0);
      else umma_ss(tmem, ad, bd, idesc, k16 > 0);
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    float* o = out + (size_t)(warp * 32 + lane) * N + c0;
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// CTA pair (cta_group::2): D[256, N] = A[256, K] * B[N, K]^T on a cluster of two CTAs.  CTA r loads rows
// [128 r, 128 r + 128) of A and rows [N/2 r, N/2 r + N/2) of B (both K-major, TMA, 128B swizzle); the leader issues
// M = 256 MMAs; each CTA reads its 128 rows of D from its own TMEM.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma_2cta_selftest_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                          float* __restrict__ out, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = base;                    // 128 x K halfs (<= 64 KB)
  uint8_t* sb = base + 65536;            // N/2 x K halfs (<= 64 KB)
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int nh = N / 2;

  if (tid == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc_2cta(&tmem_base_s, 256);
  tc_fence_before();
  cluster_sync_all();                    // barriers initialised and TMEM allocated in both CTAs
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    mbar_arrive_expect_tx(&bar_load, (128u + (uint32_t)nh) * K * 2u);
    for (int kc = 0; kc < K / 64; ++kc) {
      tma_load_3d(sa + kc * (128 * 128), &tmap_a, &bar_load, kc * 64, (int)rank * 128, 0);
      tma_load_3d(sb + kc * (nh * 128), &tmap_b, &bar_load, kc * 64, (int)rank * nh, 0);
    }
    mbar_wait(&bar_load, 0);
  }
  __syncthreads();
  cluster_sync_all();                    // both CTAs' operands have landed
  if (rank == 0 && tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(256, N, false, false);
    for (int k16 = 0; k16 < K / 16; ++k16) {
      const uint64_t ad = make_smem_desc(smem_u32(sa) + (k16 >> 2) * (128 * 128) + (k16 & 3) * 32, 16, 1024);
      const uint64_t bd = make_smem_desc(smem_u32(sb) + (k16 >> 2) * (nh * 128) + (k16 & 3) * 32, 16, 1024);
      umma_ss_2cta(tmem, ad, bd, idesc, k16 > 0);
    }
    umma_commit_2cta(&bar_mma, 0b11);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    float* o = out + (size_t)(rank * 128 + warp * 32 + lane) * N + c0;
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  cluster_sync_all();                    // both CTAs are done with TMEM before the pair's allocation goes away
  if (warp == 0) tmem_dealloc_2cta(tmem, 256);
}

// out[r][c] += f(r, c) through a swizzled fp32 staging tile and cp.reduce.async.bulk.tensor
__global__ void __launch_bounds__(128, 1)
tma_reduce_selftest_kernel(const __grid_constant__ CUtensorMap tmap_out, int rows, int cols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  const int r = threadIdx.x;                        // one row per thread
  for (int c0 = 0; c0 < cols; c0 += 32) {
    uint8_t* tile = base + (c0 / 32) * (128 * 128);
    for (int c16 = 0; c16 < 8; ++c16) {
      float4 v;
      const int c = c0 + 4 * c16;
      v.x = (float)(r * 1000 + c); v.y = v.x + 1.f; v.z = v.x + 2.f; v.w = v.x + 3.f;
      *reinterpret_cast<float4*>(tile + sw128_offset(r, c16)) = v;
    }
  }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int c0 = 0; c0 < cols; c0 += 32) tma_reduce_add_3d(&tmap_out, base + (c0 / 32) * (128 * 128), c0, 0, 0);
    tma_commit_group();
    tma_wait_group<0>();
  }
  (void)rows;
}

}  // namespace

int debug_umma(const void* a, const void* b, float* out, int N, int K, int a_mn, int b_mn, int manual_a,
               cudaStream_t st) {
  TGFR_REQUIRE(N % 64 == 0 && N >= 64 && N <= 256 && K % 64 == 0 && K >= 64 && K <= 256, "debug_umma: bad N/K");
  TGFR_REQUIRE(manual_a != 1 || (a_mn && K <= 128), "debug_umma: manual_a=1 needs a_mn and K <= 128");
  TGFR_REQUIRE(manual_a != 2 || (!a_mn && N <= 256), "debug_umma: manual_a=2 (A in TMEM) needs a K-major A");
  CUtensorMap ta, tb;
  if (!a_mn) { if (int rc = make_tmap_3d(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a, K, 128, 1, 64, 128, 1)) return rc; }
  else       { if (int rc = make_tmap_3d(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a, 128, K, 1, 64, K, 1)) return rc; }
  if (b_mn != 1) { if (int rc = make_tmap_3d(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, b, K, N, 1, 64, N, 1)) return rc; }
  else           { if (int rc = make_tmap_3d(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, b, N, K, 1, 64, K, 1)) return rc; }
  const int smem = 65536 + 131072 + 1024;
  TGFR_CUDA_OK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_selftest_kernel<<<1, 128, smem, st>>>(ta, tb, reinterpret_cast<const __half*>(a), out, N, K, a_mn, b_mn,
                                             reinterpret_cast<const __half*>(b), manual_a);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int debug_umma_2cta(const void* a, const void* b, float* out, int N, int K, cudaStream_t st) {
  TGFR_REQUIRE(N % 32 == 0 && N >= 32 && N <= 256 && K % 64 == 0 && K >= 64 && K <= 256, "debug_umma_2cta: bad N/K");
  CUtensorMap ta, tb;
  if (int rc = make_tmap_3d(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a, K, 256, 1, 64, 128, 1)) return rc;
  if (int rc = make_tmap_3d(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, b, K, N, 1, 64, N / 2, 1)) return rc;
  const int smem = 65536 + 65536 + 1024;
  TGFR_CUDA_OK(cudaFuncSetAttribute(umma_2cta_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_2cta_selftest_kernel<<<2, 128, smem, st>>>(ta, tb, out, N, K);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

int debug_tma_reduce(float* out, int rows, int cols, cudaStream_t st) {
  TGFR_REQUIRE(rows == 128 && cols % 32 == 0 && cols <= 256, "debug_tma_reduce: rows must be 128, cols % 32 == 0");
  CUtensorMap t;
  if (int rc = make_tmap_3d(&t, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, cols, rows, 1, 32, 128, 1)) return rc;
  const int smem = (cols / 32) * 16384 + 1024;
  TGFR_CUDA_OK(cudaFuncSetAttribute(tma_reduce_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tma_reduce_selftest_kernel<<<1, 128, smem, st>>>(t, rows, cols);
  TGFR_LAUNCH_OK();
  return TGFR_OK;
}

}  // namespace tgfr

// Micro-benchmark of the word-softmax task of the word-region forward (24 scores per thread -> softmax -> exp(k(a-1))
// -> packed fp16), W warps per sub-partition, no TMEM / barriers: what the arithmetic alone costs per task.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned pk(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<unsigned*>(&h); }
template <int VARIANT>
__global__ void __launch_bounds__(896, 1) k2(const float* in, uint4* out, long long* clk, int iters, int mode, volatile int* flag) {
  // 896 threads like the real kernel: warps 4-11 run the task loop (2 per sub-partition), the other 20 warps wait
  // (mode 0: exit at once; 1: nanosleep loop until the workers finish; 2: same + setmaxnreg like the real kernel)
  constexpr int TP = 24;
  __shared__ uint4 sm[256 * 3];
  __shared__ int done;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) done = 0;
  __syncthreads();
  if (warp < 4 || warp >= 12) {
    if (mode == 2) { if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;"); else asm volatile("setmaxnreg.dec.sync.aligned.u32 64;"); }
    if (mode >= 1) while (*(volatile int*)&done < 8) asm volatile("nanosleep.u32 256;");
    return;
  }
  if (mode == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
  const int tix = threadIdx.x - 128;
  float base[TP];
  for (int t = 0; t < TP; ++t) base[t] = in[(tix * TP + t) & 4095];
  const float k1 = 5.77f, nk1 = -5.77f, L2E = 1.4426950408889634f;
  long long t0 = clock64();
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
    float e[TP];
#pragma unroll
    for (int t = 0; t < TP; ++t) e[t] = base[t] + 1e-6f * it;
    float mxp[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
    for (int t = 0; t < TP; ++t) mxp[t & 3] = fmaxf(mxp[t & 3], e[t]);
    const float mx = fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3]));
    const float nmx = -mx * L2E;
    float sump[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < TP; ++t) { e[t] = ex2f(fmaf(e[t], L2E, nmx)); sump[t & 3] += e[t]; }
    const float sum = (sump[0] + sump[1]) + (sump[2] + sump[3]);
    unsigned pe[TP / 2], pa[TP / 2];
    const float inv = 1.f / sum;
#pragma unroll
    for (int t = 0; t < TP; t += 2) {
      const float a0 = e[t] * inv, a1 = e[t + 1] * inv;
      pa[t >> 1] = pk(a0, a1);
      pe[t >> 1] = pk(ex2f(fmaf(a0, k1, nk1)), ex2f(fmaf(a1, k1, nk1)));
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) sm[tix * 3 + j] = make_uint4(pe[4 * j], pe[4 * j + 1], pe[4 * j + 2], pe[4 * j + 3]);
#pragma unroll
    for (int j = 0; j < TP / 2; ++j) acc ^= pa[j];
  }
  long long t1 = clock64();
  out[blockIdx.x * 256 + tix] = make_uint4(acc, sm[tix].x, 0, 0);
  if (tix == 0 && blockIdx.x == 0) *clk = t1 - t0;
  __syncwarp();
  if ((threadIdx.x & 31) == 0) atomicAdd(&done, 1);
}

template <int VARIANT>
__global__ void k(const float* in, uint4* out, long long* clk, int iters) {
  constexpr int TP = 24;
  __shared__ uint4 sm[1024 * 3];
  float base[TP];
  for (int t = 0; t < TP; ++t) base[t] = in[(threadIdx.x * TP + t) & 4095];
  const float k1 = 5.77f, nk1 = -5.77f, L2E = 1.4426950408889634f;
  __syncthreads();
  long long t0 = clock64();
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
    float e[TP];
#pragma unroll
    for (int t = 0; t < TP; ++t) e[t] = base[t] + 1e-6f * it;
    float mxp[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
    for (int t = 0; t < TP; ++t) mxp[t & 3] = fmaxf(mxp[t & 3], e[t]);
    const float mx = fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3]));
    const float nmx = -mx * L2E;
    float sump[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < TP; ++t) { e[t] = ex2f(fmaf(e[t], L2E, nmx)); sump[t & 3] += e[t]; }
    const float sum = (sump[0] + sump[1]) + (sump[2] + sump[3]);
    unsigned pe[TP / 2], pa[TP / 2];
    if (VARIANT == 0) {            // as in the kernel (SAVE): A1 and E packed
      const float inv = 1.f / sum;
#pragma unroll
      for (int t = 0; t < TP; t += 2) {
        const float a0 = e[t] * inv, a1 = e[t + 1] * inv;
        pa[t >> 1] = pk(a0, a1);
        pe[t >> 1] = pk(ex2f(fmaf(a0, k1, nk1)), ex2f(fmaf(a1, k1, nk1)));
      }
    } else {                        // no A1: E only
      const float kinv = k1 / sum;
#pragma unroll
      for (int t = 0; t < TP; t += 2) { pa[t >> 1] = 0; pe[t >> 1] = pk(ex2f(fmaf(e[t], kinv, nk1)), ex2f(fmaf(e[t + 1], kinv, nk1))); }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) sm[threadIdx.x * 3 + j] = make_uint4(pe[4 * j], pe[4 * j + 1], pe[4 * j + 2], pe[4 * j + 3]);
#pragma unroll
    for (int j = 0; j < TP / 2; ++j) acc ^= pa[j];
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = make_uint4(acc, sm[threadIdx.x].x, 0, 0);
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
  float* in; uint4* out; long long* clk;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 1 << 22); cudaMalloc(&clk, 8);
  float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 0.001f * (i % 977) - 0.4f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const int iters = 2000;
  for (int variant = 0; variant < 2; ++variant)
    for (int warps = 4; warps <= 32; warps *= 2) {
      long long c = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (variant == 0) k<0><<<148, warps * 32>>>(in, out, clk, iters); else k<1><<<148, warps * 32>>>(in, out, clk, iters);
        cudaDeviceSynchronize();
        cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
      }
      printf("variant %d warps/SMSP %d: %.0f cycles per task per warp, %.0f per task per sub-partition\n", variant, warps / 4,
             (double)c / iters, (double)c / iters / (warps / 4));
    }
  for (int mode = 0; mode < 3; ++mode) {
    long long c = 0;
    for (int rep = 0; rep < 2; ++rep) {
      k2<0><<<148, 896>>>(in, out, clk, iters, mode, nullptr);
      cudaDeviceSynchronize();
      cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    }
    printf("896-thread CTA, 8 workers, mode %d (%s): %.0f cycles per task per warp\n", mode,
           mode == 0 ? "others exit" : mode == 1 ? "others sleep-poll" : "others sleep-poll + setmaxnreg 40/104/64", (double)c / iters);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

// Micro-benchmark: MUFU.EX2 issue rate per SM sub-partition on this GPU (fp32 and packed f16x2), 1 / 2 / 4 warps per
// sub-partition, independent chains.  nvcc -arch=sm_100a -O3 -o mufu mufu.cu && ./mufu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned ex2h2(unsigned x) { unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
template <int MODE>
__global__ void k(float* out, long long* clk, int iters) {
  float a[8];
  unsigned h[8];
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xB800B800u + threadIdx.x + i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = ex2f(a[i]) - 1.0f;
      else if (MODE == 1) h[i] = ex2h2(h[i]) ^ 0x80008000u;
      else { float x = a[i]; float p = fmaf(x, 0.0555f, 0.2402f); p = fmaf(p, x, 0.6931f); p = fmaf(p, x, 1.0f); a[i] = p - 1.0f; }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
  float* out; long long* clk; cudaMalloc(&out, 1 << 20); cudaMalloc(&clk, 8);
  const int iters = 4096;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps = 4; warps <= 32; warps *= 2) {
      long long c = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, warps * 32>>>(out, clk, iters);
        if (mode == 1) k<1><<<1, warps * 32>>>(out, clk, iters);
        if (mode == 2) k<2><<<1, warps * 32>>>(out, clk, iters);
        cudaDeviceSynchronize();
        cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
      }
      const double per = (double)c / (iters * 8.0 * (warps / 4));   // clocks per warp-instruction per sub-partition
      printf("mode %d (%s) warps/SMSP %d: %.2f clk per warp-op on one sub-partition\n", mode,
             mode == 0 ? "ex2.f32 + fadd" : mode == 1 ? "ex2.f16x2 + xor" : "3 ffma + fadd", warps / 4, per);
    }
  return 0;
}

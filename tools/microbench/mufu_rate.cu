// Microbenchmark: ex2.approx throughput per SM (and FFMA for scale), all SMs busy.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, long long* cyc) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      else x[i] = fmaf(x[i], 0.999f, -0.001f);
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int mode = 0; mode < 2; ++mode) {
    if (mode == 0) k<0><<<148, 512>>>(out, iters, cyc); else k<1><<<148, 512>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double ops = (double)iters * 8 * 512;  // per SM
    printf("%s: %.2f lanes/clk/SM\n", mode == 0 ? "ex2.approx" : "ffma", ops / h);
  }
}

// Microbenchmark: per-SM throughput of the instructions in the attention softmax loops — ex2.approx (MUFU),
// cvt.rn.bf16x2.f32 (F2FP pack), packed fp32x2 FMA, and ex2 + pack interleaved (do they share a pipe?).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, long long* cyc) {
  float x[8];
  unsigned p[8];
  for (int i = 0; i < 8; ++i) { x[i] = -0.001f * (threadIdx.x + i); p[i] = i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0 || MODE == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 1 || MODE == 2)
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p[i]) : "f"(x[i]), "f"(__uint_as_float(p[(i + 1) & 7])));
    }
    if (MODE == 3) {  // four independent packed chains
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        float2 v = __ffma2_rn(make_float2(x[i], x[i + 1]), make_float2(0.999f, 0.998f), make_float2(-0.001f, -0.002f));
        x[i] = v.x;
        x[i + 1] = v.y;
      }
    }
    if (MODE == 4) {  // the same work as scalar FFMA
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], 0.999f, -0.001f);
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float(p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  const char* names[5] = {"ex2.approx", "cvt.bf16x2 (F2FP)", "ex2 + F2FP pairs", "fma.f32x2 (4 instr = 8 fma per iter; rate in fma)", "ffma scalar (8 per iter)"};
  for (int mode = 0; mode < 5; ++mode) {
    if (mode == 0) k<0><<<148, 512>>>(out, iters, cyc);
    if (mode == 1) k<1><<<148, 512>>>(out, iters, cyc);
    if (mode == 2) k<2><<<148, 512>>>(out, iters, cyc);
    if (mode == 3) k<3><<<148, 512>>>(out, iters, cyc);
    if (mode == 4) k<4><<<148, 512>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double ops = (double)iters * 8 * 512;  // per SM
    printf("%s: %.2f lane-instr/clk/SM\n", names[mode], ops / h);
  }
}

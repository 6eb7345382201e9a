// Microbenchmark: sustained cycles per tcgen05.mma (M=128, K=16, bf16) by N, operand mode, number of independent
// accumulators and number of issuing warps.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../plainlm_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace plm;

// MODE: 0 SS K/K, 1 SS K/MN, 2 SS MN/MN, 3 TS (A in TMEM) B MN
template <int N, int MODE, int CHAINS, int WARPS>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc<512>(&slot); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const int warp = threadIdx.x >> 5;
  if (warp < WARPS && (threadIdx.x & 31) == 0) {
    constexpr int a_mn = (MODE == 2), b_mn = (MODE >= 1);
    constexpr uint32_t idesc = make_idesc_bf16(128, N, a_mn, b_mn);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32768);
    const uint64_t ad0 = a_mn ? make_smem_desc_sw128(a_addr, 8192, 1024) : make_smem_desc_sw128(a_addr, 16, 1024);
    const uint64_t bd0 = b_mn ? make_smem_desc_sw128(b_addr, 8192, 1024) : make_smem_desc_sw128(b_addr, 16, 1024);
    constexpr uint32_t astep = a_mn ? (2048 >> 4) : (32 >> 4), bstep = b_mn ? (2048 >> 4) : (32 >> 4);
    // each warp owns its accumulators: warp w uses columns [w*128, w*128+128) (N<=64*CHAINS<=128) or all (N=256, 1 warp)
    const uint32_t dbase = tm + (N == 256 ? 0 : warp * 128);
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t d = dbase + ((k % CHAINS) * N) % 128;
        if (MODE == 3) umma_ts(d, tm + 384 + (k & 3) * 8, bd0 + (k & 3) * bstep, idesc, 1u);
        else umma_ss(d, ad0 + (k & 3) * astep, bd0 + (k & 3) * bstep, idesc, 1u);
      }
    }
    long long t1 = clock64();
    umma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    long long t2 = clock64();
    if (blockIdx.x == 0 && warp == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int N, int MODE, int CHAINS, int WARPS>
void run(long long* dout) {
  const int iters = 8192;
  auto kern = mma_rate_kernel<N, MODE, CHAINS, WARPS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  kern<<<148, 128, 65536>>>(iters, dout);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[] = {"SS K/K  ", "SS K/MN ", "SS MN/MN", "TS B-MN "};
  printf("%s N=%3d chains=%d warps=%d : issue %.1f cyc/mma/warp, complete %.1f -> %.1f cyc per MMA on the pipe (nominal %d) %s\n",
         names[MODE], N, CHAINS, WARPS, (double)h[0] / iters, (double)h[1] / iters, (double)h[1] / iters / WARPS,
         128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 16);
  run<64, 0, 1, 1>(dout);
  run<64, 0, 2, 1>(dout);
  run<64, 0, 1, 2>(dout);
  run<64, 0, 1, 3>(dout);
  run<128, 0, 1, 1>(dout);
  run<128, 0, 1, 2>(dout);
  run<256, 0, 1, 1>(dout);
  run<64, 1, 1, 1>(dout);
  run<64, 1, 1, 2>(dout);
  run<64, 2, 1, 1>(dout);
  run<64, 2, 1, 2>(dout);
  run<128, 2, 1, 1>(dout);
  run<256, 2, 1, 1>(dout);
  run<64, 3, 1, 1>(dout);
  run<64, 3, 2, 1>(dout);
  run<64, 3, 1, 2>(dout);
  run<128, 3, 1, 1>(dout);
  run<256, 3, 1, 1>(dout);
  return 0;
}

// RMSNorm forward / backward (models/components.py:16-28) as single-pass bandwidth kernels.
// One warp owns one row; the row lives in registers (d = 128 * VPL, float4 per lane per 128 columns), reductions are
// warp shuffles, every global access is a 128-bit vector.  Forward: 4 B/elt read + 2 B/elt write.
// Backward: reads dy (2) + x (4) [+ dx_in (4)] through a cp.async shared-memory ring, writes dx (4) [+ bf16 copy (2)];
// dw partial sums stay in registers across the rows a warp visits and are reduced once per block.
#include "common.cuh"
#include "ptx.cuh"

namespace plm {

constexpr int NORM_WARPS = 8;
constexpr int NORM_BWD_MAX_BLOCKS = 148;  // rmsnorm_bwd: one persistent block per SM

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int VPL>
__global__ void __launch_bounds__(NORM_WARPS * 32)
rmsnorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, __nv_bfloat16* __restrict__ y,
                   float* __restrict__ rstd, int64_t rows, float eps) {
  constexpr int D = VPL * 128;
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * NORM_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * D);
  float4 v[VPL];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = __ldcs(xr + i * 32 + lane);
    ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  ss = warp_sum(ss);
  const float r = rsqrtf(ss / D + eps);
  if (lane == 0) rstd[row] = r;
  const float4* wr = reinterpret_cast<const float4*>(w);
  uint2* yr = reinterpret_cast<uint2*>(y + row * D);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float4 ww = __ldg(wr + i * 32 + lane);
    uint2 o;
    o.x = pack_bf16x2(v[i].x * r * ww.x, v[i].y * r * ww.y);
    o.y = pack_bf16x2(v[i].z * r * ww.z, v[i].w * r * ww.w);
    yr[i * 32 + lane] = o;
  }
}

// Backward.  One row per warp and step, one persistent block (up to 6 warps) per SM.  The register-resident version of this
// kernel was latency-bound (ncu: 79 % of stall cycles long-scoreboard, 4.2 TB/s): what a warp can have in flight was
// capped by the registers that receive its loads.  Here every warp streams its rows through a private shared-memory ring
// with cp.async (x, dx_in, dy of the next STAGES-1 rows are in flight while one row is computed), so ~160 of the SM's
// 227 KB of shared memory are outstanding loads.
// Measured at d = 1024 (16384 rows): 4 warps x 5 stages 53 us, 6 x 3 47.5 us, 7 x 3 47 us, 8 x 2 48 us
// (register-resident predecessor: 61 us).
constexpr int NORM_BWD_SMEM_BUDGET = 220 * 1024;
template <int VPL>
struct NormBwdCfg {
  static constexpr int D = VPL * 128;
  static constexpr int ROW_BYTES = D * 10;  // x fp32 | dx_in fp32 | dy bf16
  static constexpr int WFIT = NORM_BWD_SMEM_BUDGET / (2 * ROW_BYTES);  // warps that fit with a 2-deep ring
  static constexpr int WARPS = WFIT > 6 ? 6 : WFIT;
  static constexpr int FIT = NORM_BWD_SMEM_BUDGET / (WARPS * ROW_BYTES);
  static constexpr int STAGES = FIT > 8 ? 8 : FIT;
  static constexpr int SMEM = WARPS * STAGES * ROW_BYTES;
  static_assert(WARPS >= 1 && STAGES >= 2 && SMEM <= 227 * 1024, "rmsnorm_bwd ring does not fit");
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int VPL>
__global__ void __launch_bounds__(NormBwdCfg<VPL>::WARPS * 32, 1)
rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
                   const float* __restrict__ rstd, const float* dx_in, float* dx_out,
                   __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dw_partial, int64_t rows) {
  using Cfg = NormBwdCfg<VPL>;
  constexpr int D = Cfg::D;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NORM_BWD_WARPS = Cfg::WARPS;
  extern __shared__ __align__(16) uint8_t ring[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint8_t* wring = ring + warp * (STAGES * Cfg::ROW_BYTES);
  const float4* wr = reinterpret_cast<const float4*>(w);
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * NORM_BWD_WARPS + warp;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * NORM_BWD_WARPS;
  float4 dw[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) dw[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  // one cp.async group per row (empty past the end, so the group arithmetic stays uniform)
  auto issue = [&](int64_t row, int st) {
    if (row < rows) {
      uint8_t* base = wring + st * Cfg::ROW_BYTES;
      const float4* xs = reinterpret_cast<const float4*>(x + row * D);
#pragma unroll
      for (int i = 0; i < VPL; ++i) cp_async16(base + (i * 32 + lane) * 16, xs + i * 32 + lane);
      if (dx_in) {
        const float4* ds = reinterpret_cast<const float4*>(dx_in + row * D);
#pragma unroll
        for (int i = 0; i < VPL; ++i) cp_async16(base + D * 4 + (i * 32 + lane) * 16, ds + i * 32 + lane);
      }
      const uint4* ys = reinterpret_cast<const uint4*>(dy + row * D);
      for (int pc = lane; pc < VPL * 16; pc += 32) cp_async16(base + D * 8 + pc * 16, ys + pc);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) issue(row0 + s * stride, s);

  int st = 0;
  for (int64_t row = row0; row < rows; row += stride) {
    __syncwarp();  // every lane is done with the stage that is refilled now (it was consumed one step ago)
    issue(row + (STAGES - 1) * stride, (st + STAGES - 1) % STAGES);
    cp_async_wait<STAGES - 1>();  // this row's group has landed
    __syncwarp();                 // ... for every lane (dy pieces are copied and read by different lanes)
    const uint8_t* base = wring + st * Cfg::ROW_BYTES;
    const float r = rstd[row];
    float4 xv[VPL];
    uint2 dv[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      xv[i] = *reinterpret_cast<const float4*>(base + (i * 32 + lane) * 16);
      dv[i] = *reinterpret_cast<const uint2*>(base + D * 8 + (i * 32 + lane) * 8);
    }
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float4 ww = __ldg(wr + i * 32 + lane);
      dot += bf16_lo(dv[i].x) * ww.x * xv[i].x + bf16_hi(dv[i].x) * ww.y * xv[i].y +
             bf16_lo(dv[i].y) * ww.z * xv[i].z + bf16_hi(dv[i].y) * ww.w * xv[i].w;
    }
    dot = warp_sum(dot) * r * (1.0f / D);  // mean_j (dy_j w_j xhat_j)
    float4* dxo = reinterpret_cast<float4*>(dx_out + row * D);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float4 ww = __ldg(wr + i * 32 + lane);
      const float d0 = bf16_lo(dv[i].x), d1 = bf16_hi(dv[i].x), d2 = bf16_lo(dv[i].y), d3 = bf16_hi(dv[i].y);
      const float h0 = xv[i].x * r, h1 = xv[i].y * r, h2 = xv[i].z * r, h3 = xv[i].w * r;
      dw[i].x += d0 * h0;
      dw[i].y += d1 * h1;
      dw[i].z += d2 * h2;
      dw[i].w += d3 * h3;
      float4 o;
      o.x = r * (d0 * ww.x - h0 * dot);
      o.y = r * (d1 * ww.y - h1 * dot);
      o.z = r * (d2 * ww.z - h2 * dot);
      o.w = r * (d3 * ww.w - h3 * dot);
      if (dx_in) {
        const float4 a = *reinterpret_cast<const float4*>(base + D * 4 + (i * 32 + lane) * 16);
        o.x += a.x;
        o.y += a.y;
        o.z += a.z;
        o.w += a.w;
      }
      __stcs(dxo + i * 32 + lane, o);
      if (dx_bf16) {
        uint2 bq;
        bq.x = pack_bf16x2(o.x, o.y);
        bq.y = pack_bf16x2(o.z, o.w);
        reinterpret_cast<uint2*>(dx_bf16 + row * D)[i * 32 + lane] = bq;
      }
    }
    st = (st + 1) % STAGES;
  }
  cp_async_wait<0>();
  // block reduction of the dw partials through the (now idle) ring, fixed order (warp 0..3) => deterministic
  __syncthreads();
  float4* red = reinterpret_cast<float4*>(ring);
#pragma unroll
  for (int i = 0; i < VPL; ++i) red[warp * (D / 4) + i * 32 + lane] = dw[i];
  __syncthreads();
  float4* out = reinterpret_cast<float4*>(dw_partial + static_cast<int64_t>(blockIdx.x) * D);
  for (int c = threadIdx.x; c < D / 4; c += NORM_BWD_WARPS * 32) {
    float4 sacc = red[c];
#pragma unroll
    for (int k = 1; k < NORM_BWD_WARPS; ++k) {
      const float4 t = red[k * (D / 4) + c];
      sacc.x += t.x;
      sacc.y += t.y;
      sacc.z += t.z;
      sacc.w += t.w;
    }
    out[c] = sacc;
  }
}

// dw[c] += sum_b partial[b, c].  Block = 32 columns x 8 row groups; fixed summation order => deterministic.
__global__ void __launch_bounds__(256)
colsum_accum_kernel(const float* __restrict__ partial, float* __restrict__ dw, int nblocks, int d) {
  // blockIdx.y selects one of `batch` independent reductions laid out back to back ([batch][nblocks][d] -> [batch][d])
  partial += static_cast<int64_t>(blockIdx.y) * nblocks * d;
  dw += static_cast<int64_t>(blockIdx.y) * d;
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s0 = 0.f, s1 = 0.f;
  if (c < d) {
    int b = ty;
    for (; b + 8 < nblocks; b += 16) {
      s0 += partial[static_cast<int64_t>(b) * d + c];
      s1 += partial[static_cast<int64_t>(b + 8) * d + c];
    }
    if (b < nblocks) s0 += partial[static_cast<int64_t>(b) * d + c];
  }
  red[ty][tx] = s0 + s1;
  __syncthreads();
  if (ty == 0 && c < d) {
    float s = red[0][tx];
#pragma unroll
    for (int k = 1; k < 8; ++k) s += red[k][tx];
    dw[c] += s;
  }
}

static int norm_bwd_blocks(int64_t rows) {  // independent of d: the caller sizes dw_partial from it
  int64_t b = (rows + 3) / 4;
  if (b > NORM_BWD_MAX_BLOCKS) b = NORM_BWD_MAX_BLOCKS;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace plm

#define PLM_VPL_SWITCH(vpl, CALL)                                          \
  switch (vpl) {                                                           \
    case 1: { constexpr int V = 1; CALL; } break;                          \
    case 2: { constexpr int V = 2; CALL; } break;                          \
    case 3: { constexpr int V = 3; CALL; } break;                          \
    case 4: { constexpr int V = 4; CALL; } break;                          \
    case 6: { constexpr int V = 6; CALL; } break;                          \
    case 8: { constexpr int V = 8; CALL; } break;                          \
    case 12: { constexpr int V = 12; CALL; } break;                        \
    case 16: { constexpr int V = 16; CALL; } break;                        \
    default:                                                               \
      return plm::fail(PLM_ERR_UNSUPPORTED, "rmsnorm: d=%d not in {128,256,384,512,768,1024,1536,2048}", d); \
  }

extern "C" {

int plm_rmsnorm_fwd(const float* x, const float* w, void* y_bf16, float* rstd, int64_t rows, int32_t d, float eps,
                    plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(x);
  PLM_REQUIRE(x && w && y_bf16 && rstd, "rmsnorm_fwd: null pointer");
  PLM_REQUIRE(rows >= 0 && d > 0, "rmsnorm_fwd: bad size");
  PLM_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y_bf16), "rmsnorm_fwd: misaligned pointer");
  if (d % 128 != 0) return fail(PLM_ERR_UNSUPPORTED, "rmsnorm: d=%d must be a multiple of 128", d);
  if (rows == 0) return PLM_OK;
  const int blocks = static_cast<int>((rows + NORM_WARPS - 1) / NORM_WARPS);
  PLM_VPL_SWITCH(d / 128, (rmsnorm_fwd_kernel<V><<<blocks, NORM_WARPS * 32, 0, stream>>>(
                              x, w, static_cast<__nv_bfloat16*>(y_bf16), rstd, rows, eps)));
  return check_launch("rmsnorm_fwd");
}

int plm_rmsnorm_bwd_blocks(int64_t rows) { return plm::norm_bwd_blocks(rows); }

int plm_rmsnorm_bwd(const void* dy_bf16, const float* x, const float* w, const float* rstd, const float* dx_in,
                    float* dx_out, void* dx_out_bf16, float* dw_partial, int64_t rows, int32_t d,
                    plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(x);
  PLM_REQUIRE(dy_bf16 && x && w && rstd && dx_out && dw_partial, "rmsnorm_bwd: null pointer");
  PLM_REQUIRE(rows > 0 && d > 0, "rmsnorm_bwd: bad size");
  PLM_REQUIRE(aligned16(dy_bf16) && aligned16(x) && aligned16(w) && aligned16(dx_out) && aligned16(dw_partial) &&
                  (!dx_in || aligned16(dx_in)) && (!dx_out_bf16 || aligned16(dx_out_bf16)),
              "rmsnorm_bwd: misaligned pointer");
  if (d % 128 != 0) return fail(PLM_ERR_UNSUPPORTED, "rmsnorm: d=%d must be a multiple of 128", d);
  const int blocks = norm_bwd_blocks(rows);
  {
    cudaError_t e = cudaSuccess;
    PLM_VPL_SWITCH(d / 128, (e = cudaFuncSetAttribute(rmsnorm_bwd_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      NormBwdCfg<V>::SMEM)));
    if (e != cudaSuccess) return fail(PLM_ERR_CUDA, "rmsnorm_bwd smem attribute: %s", cudaGetErrorString(e));
  }
  PLM_VPL_SWITCH(d / 128, (rmsnorm_bwd_kernel<V><<<blocks, NormBwdCfg<V>::WARPS * 32, NormBwdCfg<V>::SMEM, stream>>>(
                              static_cast<const __nv_bfloat16*>(dy_bf16), x, w, rstd, dx_in, dx_out,
                              static_cast<__nv_bfloat16*>(dx_out_bf16), dw_partial, rows)));
  return check_launch("rmsnorm_bwd");
}

int plm_colsum_accum_batched(const float* partial, float* dw, int32_t nblocks, int32_t d, int32_t batch,
                             plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(partial);
  PLM_REQUIRE(partial && dw && nblocks > 0 && d > 0 && batch > 0 && batch <= 65535, "colsum_accum_batched: bad argument");
  colsum_accum_kernel<<<dim3((d + 31) / 32, batch), 256, 0, stream>>>(partial, dw, nblocks, d);
  return check_launch("colsum_accum_batched");
}

int plm_colsum_accum(const float* partial, float* dw, int32_t nblocks, int32_t d, plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(partial);
  PLM_REQUIRE(partial && dw && nblocks > 0 && d > 0, "colsum_accum: bad argument");
  colsum_accum_kernel<<<(d + 31) / 32, 256, 0, stream>>>(partial, dw, nblocks, d);
  return check_launch("colsum_accum");
}

}  // extern "C"

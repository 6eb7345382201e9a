// Bandwidth kernels of the plainLM train step that are neither GEMM nor attention:
// SwiGLU gate fwd/bwd (models/components.py:55-56), embedding gather / scatter-add (models/transformer.py:110),
// fp32<->bf16 flat casts (autocast weight casts, gradient bucket pack/unpack), stand-alone RoPE
// (models/embeddings.py:15-30) and the document-segment map (data/datasets/data_prep_utils.py:7-23).
// All are 128-bit vectorised, one pass over their data.
#include "common.cuh"
#include "ptx.cuh"

namespace plm {

// ------------------------------------------------------------------------------------------- SwiGLU
// u = [a | z] per row (2F columns); h = silu(a) * z.  One thread per 8 hidden elements.
__global__ void __launch_bounds__(256)
swiglu_fwd_kernel(const uint4* __restrict__ u, uint4* __restrict__ h, int64_t rows, int F8) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * F8) return;
  const int64_t r = idx / F8;
  const int c = static_cast<int>(idx - r * F8);
  const uint4 a = __ldcs(u + r * (2 * F8) + c);
  const uint4 z = __ldcs(u + r * (2 * F8) + F8 + c);
  const uint32_t av[4] = {a.x, a.y, a.z, a.w};
  const uint32_t zv[4] = {z.x, z.y, z.z, z.w};
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a0 = bf16_lo(av[i]), a1 = bf16_hi(av[i]);
    const float z0 = bf16_lo(zv[i]), z1 = bf16_hi(zv[i]);
    o[i] = pack_bf16x2(a0 * sigmoidf_fast(a0) * z0, a1 * sigmoidf_fast(a1) * z1);
  }
  h[idx] = make_uint4(o[0], o[1], o[2], o[3]);
}

// du = [dh * z * (s + a s (1-s)) | dh * a s],  s = sigmoid(a)
__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(const uint4* __restrict__ dh, const uint4* __restrict__ u, uint4* __restrict__ du, int64_t rows,
                  int F8) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * F8) return;
  const int64_t r = idx / F8;
  const int c = static_cast<int>(idx - r * F8);
  const uint4 a = __ldcs(u + r * (2 * F8) + c);
  const uint4 z = __ldcs(u + r * (2 * F8) + F8 + c);
  const uint4 g = __ldcs(dh + idx);
  const uint32_t av[4] = {a.x, a.y, a.z, a.w};
  const uint32_t zv[4] = {z.x, z.y, z.z, z.w};
  const uint32_t gv[4] = {g.x, g.y, g.z, g.w};
  uint32_t da[4], dz[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a0 = bf16_lo(av[i]), a1 = bf16_hi(av[i]);
    const float z0 = bf16_lo(zv[i]), z1 = bf16_hi(zv[i]);
    const float g0 = bf16_lo(gv[i]), g1 = bf16_hi(gv[i]);
    const float s0 = sigmoidf_fast(a0), s1 = sigmoidf_fast(a1);
    da[i] = pack_bf16x2(g0 * z0 * (s0 + a0 * s0 * (1.f - s0)), g1 * z1 * (s1 + a1 * s1 * (1.f - s1)));
    dz[i] = pack_bf16x2(g0 * a0 * s0, g1 * a1 * s1);
  }
  du[r * (2 * F8) + c] = make_uint4(da[0], da[1], da[2], da[3]);
  du[r * (2 * F8) + F8 + c] = make_uint4(dz[0], dz[1], dz[2], dz[3]);
}

// ------------------------------------------------------------------------------------------- plain MLP activations
// models/components.py:31-40 (MLP: silu) and :59-70 (MLPReluSquared: relu(u)^2).  KIND 0 = silu, 1 = relu squared.
template <int KIND>
__device__ __forceinline__ float act_f(float a) {
  if (KIND == 0) return a * sigmoidf_fast(a);
  const float r = fmaxf(a, 0.f);
  return r * r;
}
template <int KIND>
__device__ __forceinline__ float act_df(float a) {
  if (KIND == 0) {
    const float s = sigmoidf_fast(a);
    return s + a * s * (1.f - s);
  }
  return 2.f * fmaxf(a, 0.f);
}

template <int KIND>
__global__ void __launch_bounds__(256) act_fwd_kernel(const uint4* __restrict__ u, uint4* __restrict__ h, int64_t n8) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n8) return;
  const uint4 a = __ldcs(u + idx);
  const uint32_t av[4] = {a.x, a.y, a.z, a.w};
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = pack_bf16x2(act_f<KIND>(bf16_lo(av[i])), act_f<KIND>(bf16_hi(av[i])));
  h[idx] = make_uint4(o[0], o[1], o[2], o[3]);
}

template <int KIND>
__global__ void __launch_bounds__(256)
act_bwd_kernel(const uint4* __restrict__ dh, const uint4* __restrict__ u, uint4* __restrict__ du, int64_t n8) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n8) return;
  const uint4 a = __ldcs(u + idx);
  const uint4 g = __ldcs(dh + idx);
  const uint32_t av[4] = {a.x, a.y, a.z, a.w};
  const uint32_t gv[4] = {g.x, g.y, g.z, g.w};
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    o[i] = pack_bf16x2(bf16_lo(gv[i]) * act_df<KIND>(bf16_lo(av[i])), bf16_hi(gv[i]) * act_df<KIND>(bf16_hi(av[i])));
  du[idx] = make_uint4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------------------------------------- embedding
__global__ void __launch_bounds__(256)
embed_fwd_kernel(const int64_t* __restrict__ ids, const float4* __restrict__ W, float4* __restrict__ x, int64_t rows,
                 int d4, int64_t vocab) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  int64_t id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);  // ids are validated on the host; clamp keeps loads in bounds
  const float4* src = W + id * d4;
  float4* dst = x + row * d4;
  for (int i = lane; i < d4; i += 32) __stcs(dst + i, __ldg(src + i));
}

__global__ void __launch_bounds__(256)
embed_bwd_kernel(const int64_t* __restrict__ ids, const float4* __restrict__ dx, float* __restrict__ dW, int64_t rows,
                 int d4, int64_t vocab) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  int64_t id = ids[row];
  if (id < 0 || id >= vocab) return;
  float* dst = dW + id * d4 * 4;
  const float4* src = dx + row * d4;
  for (int i = lane; i < d4; i += 32) {
    const float4 v = __ldcs(src + i);
    red_add_f32x4(dst + i * 4, v.x, v.y, v.z, v.w);
  }
}

// ------------------------------------------------------------------------------------------- casts
__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n, float scale) {
  const int64_t n8 = n >> 3;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    uint4 o;
    o.x = pack_bf16x2(a.x * scale, a.y * scale);
    o.y = pack_bf16x2(a.z * scale, a.w * scale);
    o.z = pack_bf16x2(b.x * scale, b.y * scale);
    o.w = pack_bf16x2(b.z * scale, b.w * scale);
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const int64_t i = (n8 << 3) + threadIdx.x;
    dst[i] = __float2bfloat16_rn(src[i] * scale);
  }
}

__global__ void __launch_bounds__(256)
cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t n, float scale) {
  const int64_t n8 = n >> 3;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(src) + i);
    float4 a, b;
    a.x = bf16_lo(v.x) * scale;
    a.y = bf16_hi(v.x) * scale;
    a.z = bf16_lo(v.y) * scale;
    a.w = bf16_hi(v.y) * scale;
    b.x = bf16_lo(v.z) * scale;
    b.y = bf16_hi(v.z) * scale;
    b.z = bf16_lo(v.w) * scale;
    b.w = bf16_hi(v.w) * scale;
    reinterpret_cast<float4*>(dst)[2 * i] = a;
    reinterpret_cast<float4*>(dst)[2 * i + 1] = b;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const int64_t i = (n8 << 3) + threadIdx.x;
    dst[i] = __bfloat162float(src[i]) * scale;
  }
}

// ------------------------------------------------------------------------------------------- RoPE (stand-alone)
// qkv: [rows, 3*H*hd]; rotates the first 2*H*hd columns in place. One thread per 8 columns (4 pairs).
__global__ void __launch_bounds__(256)
rope_qk_kernel(uint4* __restrict__ qkv, const float* __restrict__ table, int64_t rows, int T, int H, int hd,
               float dir) {
  const int c8 = (2 * H * hd) >> 3;  // vectors per row to rotate
  const int ld8 = (3 * H * hd) >> 3;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * c8) return;
  const int64_t r = idx / c8;
  const int c = static_cast<int>(idx - r * c8);
  const int pos = static_cast<int>(r % T);
  const int pair0 = ((c * 8) % hd) >> 1;
  const float4* tab = reinterpret_cast<const float4*>(table + (static_cast<int64_t>(pos) * (hd >> 1) + pair0) * 2);
  const float4 cs0 = __ldg(tab), cs1 = __ldg(tab + 1);
  uint4 v = qkv[r * ld8 + c];
  const float cosv[4] = {cs0.x, cs0.z, cs1.x, cs1.z};
  const float sinv[4] = {cs0.y * dir, cs0.w * dir, cs1.y * dir, cs1.w * dir};
  uint32_t in[4] = {v.x, v.y, v.z, v.w}, o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = bf16_lo(in[i]), b = bf16_hi(in[i]);
    o[i] = pack_bf16x2(a * cosv[i] - b * sinv[i], b * cosv[i] + a * sinv[i]);
  }
  qkv[r * ld8 + c] = make_uint4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------------------------------------- document segments
// One block per sequence. lengths of sequence b: lengths[offsets[b] .. offsets[b+1]) summing to T+1.
// seg_start[b*T + t] = start position of the document containing t, for t < T.
__global__ void __launch_bounds__(256)
seg_start_kernel(const int32_t* __restrict__ lengths, const int32_t* __restrict__ offsets,
                 int32_t* __restrict__ seg_start, int T) {
  const int b = blockIdx.x;
  const int lo = offsets[b], hi = offsets[b + 1];
  // each thread scans documents serially; documents are few (<= T+1) and this runs once per micro-batch
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    int start = 0;
    for (int k = lo; k < hi; ++k) {
      const int len = lengths[k];
      if (t < start + len) break;
      start += len;
    }
    seg_start[static_cast<int64_t>(b) * T + t] = start;
  }
}

}  // namespace plm

extern "C" {

int plm_swiglu_fwd(const void* u, void* h, int64_t rows, int32_t F, plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(u);
  PLM_REQUIRE(u && h && rows >= 0 && F > 0, "swiglu_fwd: bad argument");
  PLM_REQUIRE(F % 8 == 0 && aligned16(u) && aligned16(h), "swiglu_fwd: F %% 8 and 16-byte alignment required");
  if (rows == 0) return PLM_OK;
  const int64_t n = rows * (F / 8);
  swiglu_fwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      static_cast<const uint4*>(u), static_cast<uint4*>(h), rows, F / 8);
  return check_launch("swiglu_fwd");
}

int plm_swiglu_bwd(const void* dh, const void* u, void* du, int64_t rows, int32_t F, plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(u);
  PLM_REQUIRE(dh && u && du && rows >= 0 && F > 0, "swiglu_bwd: bad argument");
  PLM_REQUIRE(F % 8 == 0 && aligned16(u) && aligned16(dh) && aligned16(du),
              "swiglu_bwd: F %% 8 and 16-byte alignment required");
  if (rows == 0) return PLM_OK;
  const int64_t n = rows * (F / 8);
  swiglu_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      static_cast<const uint4*>(dh), static_cast<const uint4*>(u), static_cast<uint4*>(du), rows, F / 8);
  return check_launch("swiglu_bwd");
}

int plm_act_fwd(const void* u, void* h, int64_t n, int32_t kind, plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(u);
  PLM_REQUIRE(u && h && n >= 0, "act_fwd: bad argument");
  PLM_REQUIRE(kind == PLM_ACT_SILU || kind == PLM_ACT_RELU2, "act_fwd: unknown activation %d", kind);
  PLM_REQUIRE(n % 8 == 0 && aligned16(u) && aligned16(h), "act_fwd: n %% 8 and 16-byte alignment required");
  if (n == 0) return PLM_OK;
  const int64_t n8 = n / 8;
  const unsigned grid = static_cast<unsigned>((n8 + 255) / 256);
  if (kind == PLM_ACT_SILU)
    act_fwd_kernel<0><<<grid, 256, 0, stream>>>(static_cast<const uint4*>(u), static_cast<uint4*>(h), n8);
  else
    act_fwd_kernel<1><<<grid, 256, 0, stream>>>(static_cast<const uint4*>(u), static_cast<uint4*>(h), n8);
  return check_launch("act_fwd");
}

int plm_act_bwd(const void* dh, const void* u, void* du, int64_t n, int32_t kind, plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(u);
  PLM_REQUIRE(dh && u && du && n >= 0, "act_bwd: bad argument");
  PLM_REQUIRE(kind == PLM_ACT_SILU || kind == PLM_ACT_RELU2, "act_bwd: unknown activation %d", kind);
  PLM_REQUIRE(n % 8 == 0 && aligned16(u) && aligned16(dh) && aligned16(du),
              "act_bwd: n %% 8 and 16-byte alignment required");
  if (n == 0) return PLM_OK;
  const int64_t n8 = n / 8;
  const unsigned grid = static_cast<unsigned>((n8 + 255) / 256);
  if (kind == PLM_ACT_SILU)
    act_bwd_kernel<0><<<grid, 256, 0, stream>>>(static_cast<const uint4*>(dh), static_cast<const uint4*>(u),
                                                static_cast<uint4*>(du), n8);
  else
    act_bwd_kernel<1><<<grid, 256, 0, stream>>>(static_cast<const uint4*>(dh), static_cast<const uint4*>(u),
                                                static_cast<uint4*>(du), n8);
  return check_launch("act_bwd");
}

int plm_embed_fwd(const int64_t* ids, const float* W, float* x, int64_t rows, int32_t d, int64_t vocab,
                  plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(W);
  PLM_REQUIRE(ids && W && x && rows >= 0 && d > 0 && vocab > 0, "embed_fwd: bad argument");
  PLM_REQUIRE(d % 4 == 0 && aligned16(W) && aligned16(x), "embed_fwd: d %% 4 and 16-byte alignment required");
  if (rows == 0) return PLM_OK;
  embed_fwd_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
      ids, reinterpret_cast<const float4*>(W), reinterpret_cast<float4*>(x), rows, d / 4, vocab);
  return check_launch("embed_fwd");
}

int plm_embed_bwd(const int64_t* ids, const float* dx, float* dW, int64_t rows, int32_t d, int64_t vocab,
                  plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(dx);
  PLM_REQUIRE(ids && dx && dW && rows >= 0 && d > 0 && vocab > 0, "embed_bwd: bad argument");
  PLM_REQUIRE(d % 4 == 0 && aligned16(dW) && aligned16(dx), "embed_bwd: d %% 4 and 16-byte alignment required");
  if (rows == 0) return PLM_OK;
  embed_bwd_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
      ids, reinterpret_cast<const float4*>(dx), dW, rows, d / 4, vocab);
  return check_launch("embed_bwd");
}

int plm_cast_f32_bf16(const float* src, void* dst, int64_t n, float scale, plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(src);
  PLM_REQUIRE(src && dst && n >= 0, "cast_f32_bf16: bad argument");
  PLM_REQUIRE(aligned16(src) && aligned16(dst), "cast_f32_bf16: misaligned pointer");
  if (n == 0) return PLM_OK;
  int64_t blocks = ((n >> 3) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  cast_f32_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst), n,
                                                                          scale);
  return check_launch("cast_f32_bf16");
}

int plm_cast_bf16_f32(const void* src, float* dst, int64_t n, float scale, plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(src);
  PLM_REQUIRE(src && dst && n >= 0, "cast_bf16_f32: bad argument");
  PLM_REQUIRE(aligned16(src) && aligned16(dst), "cast_bf16_f32: misaligned pointer");
  if (n == 0) return PLM_OK;
  int64_t blocks = ((n >> 3) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  cast_bf16_f32_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), dst,
                                                                          n, scale);
  return check_launch("cast_bf16_f32");
}

int plm_rope_qk(void* qkv, const float* rope_table, int64_t rows, int32_t T, int32_t H, int32_t hd, int32_t dir,
                plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(qkv);
  PLM_REQUIRE(qkv && rope_table && rows >= 0 && T > 0 && H > 0 && hd > 0, "rope_qk: bad argument");
  PLM_REQUIRE(hd % 8 == 0 && aligned16(qkv) && aligned16(rope_table), "rope_qk: hd %% 8 and alignment required");
  PLM_REQUIRE(dir == 1 || dir == -1, "rope_qk: dir must be +1 or -1");
  if (rows == 0) return PLM_OK;
  const int64_t n = rows * ((2 * H * hd) / 8);
  rope_qk_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(static_cast<uint4*>(qkv), rope_table,
                                                                               rows, T, H, hd, static_cast<float>(dir));
  return check_launch("rope_qk");
}

int plm_seg_start_from_lengths(const int32_t* lengths, const int32_t* offsets, int32_t* seg_start, int32_t B,
                               int32_t T, plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(lengths);
  PLM_REQUIRE(lengths && offsets && seg_start && B > 0 && T > 0, "seg_start: bad argument");
  seg_start_kernel<<<B, 256, 0, stream>>>(lengths, offsets, seg_start, T);
  return check_launch("seg_start");
}

}  // extern "C"

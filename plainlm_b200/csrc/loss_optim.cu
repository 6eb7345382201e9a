// Cross-entropy (engine/engine.py:81,110-112), gradient-norm (engine/engine.py:126-128) and the parameter updates
// (optim/init_optim.py:13-21 -> torch fused AdamW; optim/signSGD.py:21-46) as flat-buffer bandwidth kernels.
#include "common.cuh"
#include "ptx.cuh"

#include <math.h>

namespace plm {

// ------------------------------------------------------------------------------------------- cross-entropy
struct MS {
  float m, s;
};
__device__ __forceinline__ MS ms_combine(MS a, MS b) {
  MS r;
  r.m = fmaxf(a.m, b.m);
  if (r.m == -INFINITY) {
    r.s = 0.f;
    return r;
  }
  r.s = a.s * __expf(a.m - r.m) + b.s * __expf(b.m - r.m);
  return r;
}

__device__ __forceinline__ bool ce_ignored(int64_t t, int V) { return t < 0 || t >= V; }

// One block per row: online max / sum-exp over V bf16 logits; row_lse = logsumexp, row_loss = lse - logit[target].
__global__ void __launch_bounds__(256)
ce_stats_kernel(const __nv_bfloat16* __restrict__ logits, const int64_t* __restrict__ targets,
                float* __restrict__ row_loss, float* __restrict__ row_lse, int V, int64_t ldl) {
  __shared__ MS red[8];
  const int64_t row = blockIdx.x;
  const uint4* lp = reinterpret_cast<const uint4*>(logits + row * ldl);
  MS acc = {-INFINITY, 0.f};
  const int V8 = V >> 3;
  for (int i = threadIdx.x; i < V8; i += blockDim.x) {
    const uint4 v = __ldg(lp + i);
    float x[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y),
                  bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
    float mx = x[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) mx = fmaxf(mx, x[k]);
    const float nm = fmaxf(acc.m, mx);
    float s = acc.s * __expf(acc.m - nm);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += __expf(x[k] - nm);
    acc.m = nm;
    acc.s = s;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MS other;
    other.m = __shfl_xor_sync(0xffffffffu, acc.m, o);
    other.s = __shfl_xor_sync(0xffffffffu, acc.s, o);
    acc = ms_combine(acc, other);
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    MS t = red[0];
    for (int k = 1; k < 8; ++k) t = ms_combine(t, red[k]);
    const float lse = t.m + logf(t.s);
    row_lse[row] = lse;
    const int64_t tgt = targets[row];
    row_loss[row] = ce_ignored(tgt, V) ? 0.f : lse - __bfloat162float(logits[row * ldl + tgt]);
  }
}

// Single block: stats[0] = sum of row losses, stats[1] = #valid rows, stats[2] = mean. Fixed order.
__global__ void __launch_bounds__(1024)
ce_reduce_kernel(const float* __restrict__ row_loss, const int64_t* __restrict__ targets, float* __restrict__ stats,
                 int64_t rows, int V) {
  __shared__ float ssum[32];
  __shared__ float scnt[32];
  float s = 0.f, c = 0.f;
  for (int64_t i = threadIdx.x; i < rows; i += blockDim.x) {
    s += row_loss[i];
    c += ce_ignored(targets[i], V) ? 0.f : 1.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) {
    ssum[threadIdx.x >> 5] = s;
    scnt[threadIdx.x >> 5] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tc = 0.f;
    for (int k = 0; k < 32; ++k) {
      ts += ssum[k];
      tc += scnt[k];
    }
    stats[0] = ts;
    stats[1] = tc;
    stats[2] = tc > 0.f ? ts / tc : nanf("");  // torch returns nan when every target is ignored
  }
}

// dlogits = (softmax - onehot) * grad_scale / n_valid, written over the logits (bf16).
__global__ void __launch_bounds__(256)
ce_grad_kernel(__nv_bfloat16* __restrict__ logits, const int64_t* __restrict__ targets,
               const float* __restrict__ row_lse, const float* __restrict__ stats, int V, int64_t ldl,
               float grad_scale) {
  const int64_t row = blockIdx.x;
  uint4* lp = reinterpret_cast<uint4*>(logits + row * ldl);
  const int64_t tgt = targets[row];
  const int V8 = V >> 3;
  if (ce_ignored(tgt, V)) {
    for (int i = threadIdx.x; i < V8; i += blockDim.x) lp[i] = make_uint4(0, 0, 0, 0);
    return;
  }
  const float lse = row_lse[row];
  const float scale = grad_scale / stats[1];
  const int t8 = static_cast<int>(tgt >> 3), tk = static_cast<int>(tgt & 7);
  auto convert = [&](uint4 v, int i) -> uint4 {
    float x[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y),
                  bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float p = __expf(x[k] - lse);
      if (i == t8 && k == tk) p -= 1.0f;
      x[k] = p * scale;
    }
    uint4 o;
    o.x = pack_bf16x2(x[0], x[1]);
    o.y = pack_bf16x2(x[2], x[3]);
    o.z = pack_bf16x2(x[4], x[5]);
    o.w = pack_bf16x2(x[6], x[7]);
    return o;
  };
  // two 16-byte loads in flight per thread and iteration (one was latency-bound at 5.5 TB/s)
  int i = threadIdx.x;
  for (; i + static_cast<int>(blockDim.x) < V8; i += 2 * blockDim.x) {
    const uint4 v0 = lp[i];
    const uint4 v1 = lp[i + blockDim.x];
    lp[i] = convert(v0, i);
    lp[i + blockDim.x] = convert(v1, i + static_cast<int>(blockDim.x));
  }
  if (i < V8) lp[i] = convert(lp[i], i);
}

// Fused LM-head path: combine the per-(column tile, row) base-2 statistics the GEMM epilogue left in `partial`
// ([tiles, rows] of (max, sum 2^(x log2e - max))) into row_lse (natural log) and row_loss.  One thread per row; for a
// fixed tile consecutive threads read consecutive float2 -> coalesced.  Fixed order: deterministic.
__global__ void __launch_bounds__(256)
ce_finalize_kernel(const float2* __restrict__ partial, const float* __restrict__ tgt_logit,
                   const int64_t* __restrict__ targets, float* __restrict__ row_loss, float* __restrict__ row_lse,
                   int64_t rows, int tiles, int V) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  float m = -INFINITY, s = 0.f;
  for (int j = 0; j < tiles; ++j) {
    const float2 t = __ldg(partial + static_cast<int64_t>(j) * rows + row);
    const float nm = fmaxf(m, t.x);
    s = s * exp2f(m - nm) + t.y * exp2f(t.x - nm);
    m = nm;
  }
  const float lse = (m + log2f(s)) * 0.6931471805599453f;
  row_lse[row] = lse;
  const int64_t tgt = targets[row];
  row_loss[row] = ce_ignored(tgt, V) ? 0.f : lse - tgt_logit[row];
}

// ------------------------------------------------------------------------------------------- sum of squares
constexpr int SUMSQ_BLOCKS = 592;  // 4 per SM; must be <= PLM_SUMSQ_WORKSPACE

__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partial) {
  __shared__ float red[8];
  const int64_t n4 = n >> 2;
  // contiguous chunk per block, fixed thread->element mapping => deterministic
  const int64_t per = (n4 + gridDim.x - 1) / gridDim.x;
  const int64_t lo = per * blockIdx.x, hi = min(n4, lo + per);
  float s = 0.f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(g) + i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[(n4 << 2) + threadIdx.x];
    s += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    partial[blockIdx.x] = t;
  }
}

// Data-parallel tail (plainlm_b200/dp.py): the all-reduced bf16 wire buffer -> fp32 gradients AND the squared gradient
// norm in ONE pass (2 B read + 4 B written per element) instead of an unpack pass per bucket plus a separate sumsq pass.
// Same block -> chunk mapping and reduction order as sumsq_partial_kernel: bit-identical norm on every rank.
__global__ void __launch_bounds__(256)
unpack_sumsq_partial_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t n, float scale,
                            float* __restrict__ partial) {
  __shared__ float red[8];
  const int64_t n8 = n >> 3;
  const int64_t per = (n8 + gridDim.x - 1) / gridDim.x;
  const int64_t lo = per * blockIdx.x, hi = min(n8, lo + per);
  float s = 0.f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
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
    s += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
    reinterpret_cast<float4*>(dst)[2 * i] = a;
    reinterpret_cast<float4*>(dst)[2 * i + 1] = b;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const float v = __bfloat162float(src[(n8 << 3) + threadIdx.x]) * scale;
    dst[(n8 << 3) + threadIdx.x] = v;
    s += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    partial[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(32)
sumsq_final_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ out, int accumulate) {
  float s = 0.f;
  for (int i = threadIdx.x; i < nblocks; i += 32) s += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) out[0] = accumulate ? out[0] + s : s;
}

// ------------------------------------------------------------------------------------------- optimizers
// Non-finite gradients never reach the master weights: when the squared gradient norm handed to an update kernel is
// NaN or Inf (a NaN loss makes every gradient NaN), the kernel leaves p, the optimizer state and the bf16 shadow
// untouched.  The host raises the reference's 'Train loss is nan' (engine/engine.py:116-117) at its next check.
__device__ __forceinline__ bool grads_poisoned(const float* gnorm_sq) {
  return gnorm_sq != nullptr && !isfinite(*gnorm_sq);
}

__device__ __forceinline__ float clip_coef(const float* gnorm_sq, float max_norm) {
  if (gnorm_sq == nullptr || max_norm <= 0.f) return 1.0f;
  const float norm = sqrtf(*gnorm_sq);
  const float c = max_norm / (norm + 1e-6f);  // torch.nn.utils.clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max=1)
  return c < 1.0f ? c : 1.0f;
}

struct AdamArgs {
  float lr, one_minus_b1, b2, one_minus_b2, eps, lr_wd, step_size, bc2_sqrt, max_norm;
};

__device__ __forceinline__ void adamw_elem(float& p, float g, float& m, float& v, const AdamArgs& a) {
  p -= a.lr_wd * p;
  m = m + a.one_minus_b1 * (g - m);
  v = a.b2 * v + a.one_minus_b2 * g * g;
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p -= a.step_size * m / denom;
}

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             __nv_bfloat16* __restrict__ pb, int64_t n, AdamArgs a, const float* __restrict__ gnorm_sq) {
  if (grads_poisoned(gnorm_sq)) return;
  const float clip = clip_coef(gnorm_sq, a.max_norm);
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adamw_elem(pv.x, gv.x * clip, mv.x, vv.x, a);
    adamw_elem(pv.y, gv.y * clip, mv.y, vv.y, a);
    adamw_elem(pv.z, gv.z * clip, mv.z, vv.z, a);
    adamw_elem(pv.w, gv.w * clip, mv.w, vv.w, a);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (pb) {
      uint2 o;
      o.x = pack_bf16x2(pv.x, pv.y);
      o.y = pack_bf16x2(pv.z, pv.w);
      reinterpret_cast<uint2*>(pb)[i] = o;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    float pv = p[i], mv = m[i], vv = v[i];
    adamw_elem(pv, g[i] * clip, mv, vv, a);
    p[i] = pv;
    m[i] = mv;
    v[i] = vv;
    if (pb) pb[i] = __float2bfloat16_rn(pv);
  }
}

__device__ __forceinline__ float signf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

__device__ __forceinline__ void signsgd_elem(float& p, float g, float& m, float lr, float mu, float omd, float decay,
                                             int first) {
  p *= decay;
  const float m0 = first ? g : m;  // optim/signSGD.py:38-39: m initialised to a clone of the gradient
  m = m0 * mu + omd * g;
  p -= lr * signf(m);
}

__global__ void __launch_bounds__(256)
signsgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
               __nv_bfloat16* __restrict__ pb, int64_t n, float lr, float mu, float omd, float decay, int first,
               const float* __restrict__ gnorm_sq, float max_norm) {
  if (grads_poisoned(gnorm_sq)) return;
  const float clip = clip_coef(gnorm_sq, max_norm);
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 mv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(m)[i];
    signsgd_elem(pv.x, gv.x * clip, mv.x, lr, mu, omd, decay, first);
    signsgd_elem(pv.y, gv.y * clip, mv.y, lr, mu, omd, decay, first);
    signsgd_elem(pv.z, gv.z * clip, mv.z, lr, mu, omd, decay, first);
    signsgd_elem(pv.w, gv.w * clip, mv.w, lr, mu, omd, decay, first);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    if (pb) {
      uint2 o;
      o.x = pack_bf16x2(pv.x, pv.y);
      o.y = pack_bf16x2(pv.z, pv.w);
      reinterpret_cast<uint2*>(pb)[i] = o;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    float pv = p[i], mv = first ? 0.f : m[i];
    signsgd_elem(pv, g[i] * clip, mv, lr, mu, omd, decay, first);
    p[i] = pv;
    m[i] = mv;
    if (pb) pb[i] = __float2bfloat16_rn(pv);
  }
}

// ---- NAdam with decoupled weight decay (torch.optim.NAdam(decoupled_weight_decay=True), optim/init_optim.py:23-32)
// and SGD with momentum (torch.optim.SGD, init_optim.py:34-41) share one flat kernel shape.
struct NAdamArgs {
  float decay, one_minus_b1, b2, one_minus_b2, eps, inv_bc2, c_grad, c_mom, max_norm;
};
struct NAdamOp {
  NAdamArgs a;
  __device__ __forceinline__ void operator()(float& p, float g, float& m, float& v) const {
    p *= a.decay;
    m = m + a.one_minus_b1 * (g - m);
    v = a.b2 * v + a.one_minus_b2 * g * g;
    const float denom = sqrtf(v * a.inv_bc2) + a.eps;
    p += a.c_grad * g / denom;  // torch: two addcdiv_ calls, grad first
    p += a.c_mom * m / denom;
  }
};
struct SgdArgs {
  float lr, momentum, one_minus_damp, wd, max_norm;
  int first, use_momentum;
};
struct SgdOp {
  SgdArgs a;
  __device__ __forceinline__ void operator()(float& p, float g, float& buf, float&) const {
    g = g + a.wd * p;  // coupled (L2) weight decay
    if (a.use_momentum) {
      buf = a.first ? g : a.momentum * buf + a.one_minus_damp * g;  // torch: buffer starts as a clone of the gradient
      g = buf;
    }
    p -= a.lr * g;
  }
};

template <class Op, bool HAS_V>
__global__ void __launch_bounds__(256)
flat_opt_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                __nv_bfloat16* __restrict__ pb, int64_t n, Op op, float max_norm, const float* __restrict__ gnorm_sq) {
  if (grads_poisoned(gnorm_sq)) return;
  const float clip = clip_coef(gnorm_sq, max_norm);
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = HAS_V ? reinterpret_cast<float4*>(v)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    op(pv.x, gv.x * clip, mv.x, vv.x);
    op(pv.y, gv.y * clip, mv.y, vv.y);
    op(pv.z, gv.z * clip, mv.z, vv.z);
    op(pv.w, gv.w * clip, mv.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    if (HAS_V) reinterpret_cast<float4*>(v)[i] = vv;
    if (pb) {
      uint2 o;
      o.x = pack_bf16x2(pv.x, pv.y);
      o.y = pack_bf16x2(pv.z, pv.w);
      reinterpret_cast<uint2*>(pb)[i] = o;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    float pv = p[i], mv = m[i], vv = HAS_V ? v[i] : 0.f;
    op(pv, g[i] * clip, mv, vv);
    p[i] = pv;
    m[i] = mv;
    if (HAS_V) v[i] = vv;
    if (pb) pb[i] = __float2bfloat16_rn(pv);
  }
}

static unsigned flat_grid(int64_t n4) {
  int64_t blocks = (n4 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<unsigned>(blocks);
}

}  // namespace plm

extern "C" {

int plm_ce_fwd_bwd(void* logits, const int64_t* targets, float* row_loss, float* row_lse, float* stats, int64_t rows,
                   int32_t V, int64_t ldl, float grad_scale, int32_t write_grad, plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(logits);
  PLM_REQUIRE(logits && targets && row_loss && row_lse && stats, "ce: null pointer");
  PLM_REQUIRE(rows > 0 && V > 0 && ldl >= V, "ce: bad size");
  PLM_REQUIRE(V % 8 == 0 && ldl % 8 == 0 && aligned16(logits), "ce: V, ldl must be multiples of 8, logits aligned");
  PLM_REQUIRE(rows < (1ll << 31), "ce: too many rows");
  __nv_bfloat16* lg = static_cast<__nv_bfloat16*>(logits);
  ce_stats_kernel<<<static_cast<unsigned>(rows), 256, 0, stream>>>(lg, targets, row_loss, row_lse, V, ldl);
  int rc = check_launch("ce_stats");
  if (rc != PLM_OK) return rc;
  ce_reduce_kernel<<<1, 1024, 0, stream>>>(row_loss, targets, stats, rows, V);
  rc = check_launch("ce_reduce");
  if (rc != PLM_OK) return rc;
  if (write_grad) {
    ce_grad_kernel<<<static_cast<unsigned>(rows), 256, 0, stream>>>(lg, targets, row_lse, stats, V, ldl, grad_scale);
    rc = check_launch("ce_grad");
  }
  return rc;
}

int plm_lmhead_ce_fwd(const void* h, const void* W, const int64_t* targets, void* logits, int64_t ldl, float* partial,
                      float* tgt_logit, float* row_loss, float* row_lse, float* stats, int64_t rows, int32_t d, int32_t V,
                      int64_t ldh, int64_t ldw, plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(h);
  PLM_REQUIRE(h && W && targets && partial && tgt_logit && row_loss && row_lse && stats, "lmhead_ce: null pointer");
  PLM_REQUIRE(rows > 0 && rows < (1ll << 31) && V > 0 && d > 0, "lmhead_ce: bad size");
  plm_gemm_args g = {};
  g.A = h;
  g.B = W;
  g.C = logits;  // NULL: loss-only forward, nothing of size [rows, V] is written
  g.M = rows;
  g.N = V;
  g.K = d;
  g.lda = ldh;
  g.ldb = ldw;
  g.ldc = logits ? ldl : ((static_cast<int64_t>(V) + 7) & ~7ll);
  g.a_kmajor = 1;
  g.b_kmajor = 1;
  g.epilogue = PLM_EPI_BF16_CE;
  g.splits = 1;
  g.ce_targets = targets;
  g.ce_partial = partial;
  g.ce_tgt_logit = tgt_logit;
  int rc = plm_gemm_bf16(&g, stream_);
  if (rc != PLM_OK) return rc;
  const int tiles = plm_lmhead_ce_tiles(V);
  ce_finalize_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const float2*>(partial), tgt_logit, targets, row_loss, row_lse, rows, tiles, V);
  rc = check_launch("ce_finalize");
  if (rc != PLM_OK) return rc;
  ce_reduce_kernel<<<1, 1024, 0, stream>>>(row_loss, targets, stats, rows, V);
  return check_launch("ce_reduce");
}

int plm_ce_grad(void* logits, const int64_t* targets, const float* row_lse, const float* stats, int64_t rows, int32_t V,
                int64_t ldl, float grad_scale, plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(logits);
  PLM_REQUIRE(logits && targets && row_lse && stats, "ce_grad: null pointer");
  PLM_REQUIRE(rows > 0 && rows < (1ll << 31) && V > 0 && ldl >= V, "ce_grad: bad size");
  PLM_REQUIRE(V % 8 == 0 && ldl % 8 == 0 && aligned16(logits), "ce_grad: V, ldl must be multiples of 8, logits aligned");
  ce_grad_kernel<<<static_cast<unsigned>(rows), 256, 0, stream>>>(static_cast<__nv_bfloat16*>(logits), targets, row_lse,
                                                                  stats, V, ldl, grad_scale);
  return check_launch("ce_grad");
}

int plm_sumsq(const float* g, int64_t n, float* workspace, float* out, int32_t accumulate, plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(g);
  static_assert(SUMSQ_BLOCKS <= PLM_SUMSQ_WORKSPACE, "workspace too small");
  PLM_REQUIRE(g && workspace && out && n >= 0, "sumsq: bad argument");
  PLM_REQUIRE(aligned16(g), "sumsq: misaligned pointer");
  sumsq_partial_kernel<<<SUMSQ_BLOCKS, 256, 0, stream>>>(g, n, workspace);
  int rc = check_launch("sumsq_partial");
  if (rc != PLM_OK) return rc;
  sumsq_final_kernel<<<1, 32, 0, stream>>>(workspace, SUMSQ_BLOCKS, out, accumulate);
  return check_launch("sumsq_final");
}

int plm_unpack_sumsq(const void* src_bf16, float* dst, int64_t n, float scale, float* workspace, float* out,
                     plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(src_bf16);
  PLM_REQUIRE(src_bf16 && dst && workspace && out && n >= 0, "unpack_sumsq: bad argument");
  PLM_REQUIRE(aligned16(src_bf16) && aligned16(dst), "unpack_sumsq: misaligned pointer");
  unpack_sumsq_partial_kernel<<<SUMSQ_BLOCKS, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src_bf16), dst, n, scale,
                                                                 workspace);
  int rc = check_launch("unpack_sumsq_partial");
  if (rc != PLM_OK) return rc;
  sumsq_final_kernel<<<1, 32, 0, stream>>>(workspace, SUMSQ_BLOCKS, out, 0);
  return check_launch("sumsq_final");
}

int plm_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, float bc1, float bc2, const float* gnorm_sq,
                   float max_norm, plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(p);
  PLM_REQUIRE(p && g && m && v && n >= 0, "adamw: bad argument");
  PLM_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v), "adamw: misaligned pointer");
  PLM_REQUIRE(!p_bf16 || (reinterpret_cast<uintptr_t>(p_bf16) & 7) == 0, "adamw: misaligned bf16 shadow");
  PLM_REQUIRE(bc1 > 0.f && bc2 > 0.f, "adamw: bias corrections must be positive (step >= 1)");
  if (n == 0) return PLM_OK;
  AdamArgs a;
  a.lr = lr;
  a.one_minus_b1 = static_cast<float>(1.0 - static_cast<double>(beta1));
  a.b2 = beta2;
  a.one_minus_b2 = static_cast<float>(1.0 - static_cast<double>(beta2));
  a.eps = eps;
  a.lr_wd = lr * weight_decay;
  a.step_size = lr / bc1;
  a.bc2_sqrt = sqrtf(bc2);
  a.max_norm = max_norm;
  adamw_kernel<<<flat_grid(n >> 2), 256, 0, stream>>>(p, g, m, v, static_cast<__nv_bfloat16*>(p_bf16), n, a, gnorm_sq);
  return check_launch("adamw");
}

int plm_signsgd_step(float* p, const float* g, float* m, void* p_bf16, int64_t n, float lr, float momentum,
                     float dampening, float weight_decay, int32_t first_step, const float* gnorm_sq, float max_norm,
                     plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(p);
  PLM_REQUIRE(p && g && m && n >= 0, "signsgd: bad argument");
  PLM_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m), "signsgd: misaligned pointer");
  PLM_REQUIRE(!p_bf16 || (reinterpret_cast<uintptr_t>(p_bf16) & 7) == 0, "signsgd: misaligned bf16 shadow");
  if (n == 0) return PLM_OK;
  signsgd_kernel<<<flat_grid(n >> 2), 256, 0, stream>>>(p, g, m, static_cast<__nv_bfloat16*>(p_bf16), n, lr, momentum,
                                                         static_cast<float>(1.0 - static_cast<double>(dampening)),
                                                         static_cast<float>(1.0 - static_cast<double>(lr) * weight_decay),
                                                         first_step ? 1 : 0,
                                                         gnorm_sq, max_norm);
  return check_launch("signsgd");
}

int plm_nadamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, float bc2, float c_grad, float c_mom,
                    const float* gnorm_sq, float max_norm, plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(p);
  PLM_REQUIRE(p && g && m && v && n >= 0, "nadamw: bad argument");
  PLM_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v), "nadamw: misaligned pointer");
  PLM_REQUIRE(!p_bf16 || (reinterpret_cast<uintptr_t>(p_bf16) & 7) == 0, "nadamw: misaligned bf16 shadow");
  PLM_REQUIRE(bc2 > 0.f, "nadamw: bias correction must be positive (step >= 1)");
  if (n == 0) return PLM_OK;
  NAdamOp op;
  op.a.decay = static_cast<float>(1.0 - static_cast<double>(lr) * weight_decay);
  op.a.one_minus_b1 = static_cast<float>(1.0 - static_cast<double>(beta1));
  op.a.b2 = beta2;
  op.a.one_minus_b2 = static_cast<float>(1.0 - static_cast<double>(beta2));
  op.a.eps = eps;
  op.a.inv_bc2 = 1.0f / bc2;
  op.a.c_grad = c_grad;
  op.a.c_mom = c_mom;
  flat_opt_kernel<NAdamOp, true><<<flat_grid(n >> 2), 256, 0, stream>>>(p, g, m, v, static_cast<__nv_bfloat16*>(p_bf16),
                                                                         n, op, max_norm, gnorm_sq);
  return check_launch("nadamw");
}

int plm_sgd_step(float* p, const float* g, float* buf, void* p_bf16, int64_t n, float lr, float momentum,
                 float dampening, float weight_decay, int32_t first_step, const float* gnorm_sq, float max_norm,
                 plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(p);
  PLM_REQUIRE(p && g && buf && n >= 0, "sgd: bad argument");
  PLM_REQUIRE(aligned16(p) && aligned16(g) && aligned16(buf), "sgd: misaligned pointer");
  PLM_REQUIRE(!p_bf16 || (reinterpret_cast<uintptr_t>(p_bf16) & 7) == 0, "sgd: misaligned bf16 shadow");
  if (n == 0) return PLM_OK;
  SgdOp op;
  op.a.lr = lr;
  op.a.momentum = momentum;
  op.a.one_minus_damp = static_cast<float>(1.0 - static_cast<double>(dampening));
  op.a.wd = weight_decay;
  op.a.first = first_step ? 1 : 0;
  op.a.use_momentum = momentum != 0.f ? 1 : 0;
  flat_opt_kernel<SgdOp, false><<<flat_grid(n >> 2), 256, 0, stream>>>(p, g, buf, nullptr,
                                                                        static_cast<__nv_bfloat16*>(p_bf16), n, op,
                                                                        max_norm, gnorm_sq);
  return check_launch("sgd");
}

}  // extern "C"

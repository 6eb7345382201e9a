// Flash-attention backward on tcgen05 (backward of models/transformer.py:53-63, reached from engine/engine.py:120).
//
// One CTA per (128-key tile j, head, batch), looping over the query tiles i >= j that can see it. 576 threads:
//   warps 0..15 compute: warp = (TMEM lane quarter, column quarter); a thread owns 32 columns of key row r of the
//               transposed score tile — four warps per scheduler so TMEM / smem / MUFU latencies overlap
//   warp 16     MMA issuer (one thread, all tcgen05.mma)      warp 17 TMA producer (K,V once; Q_i,dO_i 3-stage ring)
// Five GEMMs per (j, i) pair, all on the tensor core, all 512 TMEM columns in use:
//   S^T  = K Q_i^T        (cols   0..127)      dP^T = V dO_i^T       (cols 128..255)
//   dV  += P^T dO_i       (cols 256..319)      dK  += dS^T Q_i       (cols 320..383)
//   dQ_i = dS K           (cols 384..447) -> fp32 red.add into dq_acc (finalised by dq_finalize_kernel)
//   P^T as packed bf16    (cols 448..511) -> A operand of the dV GEMM read straight from tensor memory
// dS^T is written once to swizzled smem as a K-major A operand (dK) and re-read MN-major for the dQ GEMM; Q_i / dO_i /
// K are consumed as MN-major B operands straight from their TMA boxes — no transposes anywhere.  The issue order
// (dV_i, S_{i+1} | dK_i, dQ_i, dP_{i+1}) keeps the tensor pipe busy while the compute warps do the exp / dS math of
// the neighbouring step.  dK and dQ are rotated back through RoPE in the epilogues.
#include "common.cuh"
#include "ptx.cuh"

#include <mutex>

namespace plm {

constexpr int AB_T = 128;   // tile edge (keys per CTA, queries per step)
constexpr int AB_HD = 64;
constexpr int AB_CWARPS = 16;    // compute warps: (TMEM lane quarter) x (column quarter)
constexpr int AB_THREADS = (AB_CWARPS + 2) * 32;  // + MMA issuer warp + TMA producer warp
constexpr int AB_STAGES = 3;       // Q_i / dO_i ring
constexpr int AB_TILE = AB_T * AB_HD * 2;  // 16 KB
// K, V, (Q,dO) x AB_STAGES, dS^T (2 blocks), vectors (lse2, delta, seg) x2 stages, barriers
constexpr int AB_VEC_BYTES = 2 * 3 * AB_T * 4;
constexpr int AB_SMEM = AB_TILE * (2 + 2 * AB_STAGES + 2) + AB_VEC_BYTES + 256;

__device__ __forceinline__ float ex2b(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// delta[b,h,t] = sum_c dO[t,h,c] * O[t,h,c].  8 lanes per (row, head): each lane 8 elements.
__global__ void __launch_bounds__(256)
attn_delta_kernel(const uint4* __restrict__ o, const uint4* __restrict__ dout, float* __restrict__ delta, int64_t rows,
                  int T, int H) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // one per 8 elements
  const int64_t total = rows * H * 8;
  float s = 0.f;
  if (idx < total) {
    const uint4 a = __ldg(o + idx), g = __ldg(dout + idx);
    s = bf16_lo(a.x) * bf16_lo(g.x) + bf16_hi(a.x) * bf16_hi(g.x) + bf16_lo(a.y) * bf16_lo(g.y) +
        bf16_hi(a.y) * bf16_hi(g.y) + bf16_lo(a.z) * bf16_lo(g.z) + bf16_hi(a.z) * bf16_hi(g.z) +
        bf16_lo(a.w) * bf16_lo(g.w) + bf16_hi(a.w) * bf16_hi(g.w);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if (idx < total && (threadIdx.x & 7) == 0) {
    const int64_t rh = idx >> 3;  // row * H + h
    const int64_t row = rh / H;
    const int h = static_cast<int>(rh - row * H);
    const int64_t b = row / T;
    const int t = static_cast<int>(row - b * T);
    delta[(b * H + h) * T + t] = s;
  }
}

// dq_acc fp32 [rows, d] -> inverse RoPE -> bf16 into dqkv[:, 0:d].  One thread per 8 columns.
__global__ void __launch_bounds__(256)
dq_finalize_kernel(const float* __restrict__ dq_acc, const float* __restrict__ table, __nv_bfloat16* __restrict__ dqkv,
                   int64_t rows, int T, int d, int hd, float scale) {
  const int d8 = d >> 3;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * d8) return;
  const int64_t r = idx / d8;
  const int c = static_cast<int>(idx - r * d8) * 8;
  const float4 a = __ldcs(reinterpret_cast<const float4*>(dq_acc + r * d + c));
  const float4 b = __ldcs(reinterpret_cast<const float4*>(dq_acc + r * d + c) + 1);
  float v[8] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale, b.x * scale, b.y * scale, b.z * scale, b.w * scale};
  if (table) {
    const int pos = static_cast<int>(r % T);
    const int pair0 = (c % hd) >> 1;
    const float4* tab = reinterpret_cast<const float4*>(table + (static_cast<int64_t>(pos) * (hd >> 1) + pair0) * 2);
    const float4 cs0 = __ldg(tab), cs1 = __ldg(tab + 1);
    const float cosv[4] = {cs0.x, cs0.z, cs1.x, cs1.z};
    const float sinv[4] = {cs0.y, cs0.w, cs1.y, cs1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // transpose of the forward rotation = rotation by -theta
      const float x0 = v[2 * i], x1 = v[2 * i + 1];
      v[2 * i] = x0 * cosv[i] + x1 * sinv[i];
      v[2 * i + 1] = x1 * cosv[i] - x0 * sinv[i];
    }
  }
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]);
  o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]);
  o.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(dqkv + r * (3 * d) + c) = o;
}

// p = exp2(s * scale_log2 - lse2), optionally masked.  MASK: 0 = none (interior tile), 1 = causal/document/ragged.
template <bool MASK>
__device__ __forceinline__ void bwd_p_chunk(const uint32_t (&t)[32], float (&p)[32], const float* lse2, const int32_t* sg,
                                            int kj, int qpos0, float scale_log2) {
#pragma unroll
  for (int q4 = 0; q4 < 8; ++q4) {
    const float4 l = *reinterpret_cast<const float4*>(lse2 + q4 * 4);
    const float lv[4] = {l.x, l.y, l.z, l.w};
    int4 g = make_int4(0, 0, 0, 0);
    if (MASK) g = *reinterpret_cast<const int4*>(sg + q4 * 4);
    const int gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int q = q4 * 4 + e;
      float v = ex2b(fmaf(__uint_as_float(t[q]), scale_log2, -lv[e]));
      if (MASK) {
        const int qi = qpos0 + q;
        if (kj > qi || kj < gv[e]) v = 0.f;
      }
      p[q] = v;
    }
  }
}

__device__ __forceinline__ void store_bf16_row32(uint8_t* row_base, int r, int chunk0, const float (&v)[32]) {
  // 32 consecutive K-columns of one A-operand row -> four 16-byte chunks of the 128B-swizzled row
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 o;
    o.x = pack_bf16x2(v[8 * c + 0], v[8 * c + 1]);
    o.y = pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
    o.z = pack_bf16x2(v[8 * c + 4], v[8 * c + 5]);
    o.w = pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
    *reinterpret_cast<uint4*>(row_base + ((((chunk0 + c) & 7) ^ (r & 7)) << 4)) = o;
  }
}

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const float* __restrict__ lse, const float* __restrict__ delta, const int32_t* __restrict__ seg_start,
                const float* __restrict__ rope, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dq_acc, int T,
                int H, float scale, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = smem + AB_TILE;
  uint8_t* sQ = smem + 2 * AB_TILE;                    // [AB_STAGES]
  uint8_t* sDO = smem + (2 + AB_STAGES) * AB_TILE;     // [AB_STAGES]
  uint8_t* sDS = smem + (2 + 2 * AB_STAGES) * AB_TILE; // dS^T: 2 blocks (q 0..63 | 64..127), each [128 kv rows x 128 B]
  float* sLse = reinterpret_cast<float*>(smem + (4 + 2 * AB_STAGES) * AB_TILE);  // [2][128]  lse * log2(e)
  float* sDelta = sLse + 2 * AB_T;                                                 // [2][128]
  int32_t* sSeg = reinterpret_cast<int32_t*>(sDelta + 2 * AB_T);                   // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (4 + 2 * AB_STAGES) * AB_TILE + AB_VEC_BYTES);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;                // [AB_STAGES]
  uint64_t* qdo_empty = bars + 1 + AB_STAGES;   // [AB_STAGES]
  uint64_t* s_full = bars + 1 + 2 * AB_STAGES;
  uint64_t* dp_full = s_full + 1;
  uint64_t* p_ready = s_full + 2;
  uint64_t* ds_ready = s_full + 3;
  uint64_t* dq_full = s_full + 4;
  uint64_t* dq_empty = s_full + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 6);

  if ((smem_u32(smem) & 1023u) != 0) return;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nq = (T + AB_T - 1) / AB_T;  // ragged tail: T need not be a multiple of 128
  const int j = blockIdx.x;  // key tile
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int d = H * AB_HD;
  const int64_t seq0 = static_cast<int64_t>(b) * T;
  const int64_t krow0 = seq0 + j * AB_T;

  // last query tile that can see this key tile: seg_start is non-decreasing in t
  int i_hi = nq - 1;
  if (seg_start) {
    const int k_last = j * AB_T + AB_T - 1;
    while (i_hi > j && seg_start[seq0 + i_hi * AB_T] > k_last) --i_hi;
  }
  const int n_it = i_hi - j + 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(kv_full, 1);
    for (int s = 0; s < AB_STAGES; ++s) {
      mbar_init(&qdo_full[s], 1);
      mbar_init(&qdo_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(p_ready, AB_CWARPS);
    mbar_init(ds_ready, AB_CWARPS);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, AB_CWARPS);
    fence_barrier_init();
  }
  if (warp == AB_CWARPS) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S^T 0..127 | dP^T 128..255 | dV 256..319 | dK 320..383 | dQ 384..447 | P^T (bf16 pairs) 448..511
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320,
                 tDQ = tmem_base + 384, tP = tmem_base + 448;

  if (warp == AB_CWARPS + 1) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * AB_TILE);
      tma_load_2d(sK, &tmQKV, kv_full, d + h * AB_HD, static_cast<int>(krow0));
      tma_load_2d(sV, &tmQKV, kv_full, 2 * d + h * AB_HD, static_cast<int>(krow0));
      for (int it = 0; it < n_it; ++it) {
        const int st = it % AB_STAGES, use = it / AB_STAGES;
        mbar_wait(&qdo_empty[st], (use & 1) ^ 1);
        const int qr = static_cast<int>(seq0 + (j + it) * AB_T);
        mbar_arrive_expect_tx(&qdo_full[st], 2 * AB_TILE);
        tma_load_2d(sQ + st * AB_TILE, &tmQKV, &qdo_full[st], h * AB_HD, qr);
        tma_load_2d(sDO + st * AB_TILE, &tmDO, &qdo_full[st], h * AB_HD, qr);
      }
    }
  } else if (warp == AB_CWARPS) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer.  Tensor-pipe order per step `it`:
      //   dV_it (needs P^T_it) , S^T_{it+1} | dK_it , dQ_it (need dS^T_it) , dP^T_{it+1}
      // so the exp-heavy half of step it+1 overlaps the dK/dQ/dP MMAs of step it, and its dS half overlaps dV/S.
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // K-major x K-major, N = 128 queries
      constexpr uint32_t idesc_kn = make_idesc_bf16(128, 64, 0, 1);   // A K-major (smem or TMEM), B MN-major, N = 64
      constexpr uint32_t idesc_nn = make_idesc_bf16(128, 64, 1, 1);   // A MN-major, B MN-major (dQ)
      mbar_wait(kv_full, 0);
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), ds_addr = smem_u32(sDS);
      auto issue_s = [&](int it) {
        const int st = it % AB_STAGES;
        mbar_wait(&qdo_full[st], (it / AB_STAGES) & 1);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(sQ + st * AB_TILE);
#pragma unroll
        for (int k = 0; k < AB_HD / 16; ++k)
          umma_ss(tS, make_smem_desc_sw128(k_addr + k * 32, 16, 1024), make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                  idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
      };
      auto issue_dp = [&](int it) {
        const int st = it % AB_STAGES;
        const uint32_t do_addr = smem_u32(sDO + st * AB_TILE);
#pragma unroll
        for (int k = 0; k < AB_HD / 16; ++k)
          umma_ss(tDP, make_smem_desc_sw128(v_addr + k * 32, 16, 1024),
                  make_smem_desc_sw128(do_addr + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(dp_full);
      };
      issue_s(0);
      issue_dp(0);
      for (int it = 0; it < n_it; ++it) {
        const int st = it % AB_STAGES;
        const uint32_t q_addr = smem_u32(sQ + st * AB_TILE), do_addr = smem_u32(sDO + st * AB_TILE);
        // ---- dV += P^T dO   (A = P^T from tensor memory: 8 columns per 16-query K-step)
        mbar_wait(p_ready, it & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AB_T / 16; ++k)
          umma_ts(tDV, tP + k * 8, make_smem_desc_sw128(do_addr + k * 2048, AB_TILE, 1024), idesc_kn,
                  (it > 0 || k > 0) ? 1u : 0u);
        if (it + 1 < n_it) issue_s(it + 1);  // S^T region is free: every compute warp read it before p_ready
        // ---- dK += dS^T Q ; dQ = dS K
        mbar_wait(ds_ready, it & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AB_T / 16; ++k)
          umma_ss(tDK, make_smem_desc_sw128(ds_addr + (k >> 2) * AB_TILE + (k & 3) * 32, 16, 1024),
                  make_smem_desc_sw128(q_addr + k * 2048, AB_TILE, 1024), idesc_kn, (it > 0 || k > 0) ? 1u : 0u);
        if (it > 0) {
          mbar_wait(dq_empty, (it - 1) & 1);
          tc_fence_after();
        }
        // A = dS^T buffer read MN-major (M = queries, 64 per block, blocks AB_TILE apart), B = K MN-major
#pragma unroll
        for (int k = 0; k < AB_T / 16; ++k)
          umma_ss(tDQ, make_smem_desc_sw128(ds_addr + k * 2048, AB_TILE, 1024),
                  make_smem_desc_sw128(k_addr + k * 2048, AB_TILE, 1024), idesc_nn, k > 0 ? 1u : 0u);
        umma_commit(&qdo_empty[st]);
        umma_commit(dq_full);
        if (it + 1 < n_it) issue_dp(it + 1);  // dP^T region is free: every compute warp read it before ds_ready
      }
    }
  } else {
    // ------------------------------------------------------------ compute warps (16): warp = (lane quarter, column quarter)
    // Four warps per scheduler: TMEM / shared-memory / MUFU latencies of one warp hide behind the other three.
    const int quarter = warp & 3;
    const int cq = warp >> 2;                // query columns [32*cq, +32) of S^T / dP^T; hd cols [16*cq, +16) of dQ/dK/dV
    const int r = quarter * 32 + lane;       // key row within the tile (S^T lane) / query row within the tile (dQ lane)
    const int kj = j * AB_T + r;             // key position
    const int ct = threadIdx.x;              // 0..511
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int64_t vec_base = (static_cast<int64_t>(b) * H + h) * T;

    // per-query vectors of a step: threads 0..127 fetch lse, 128..255 delta, 256..383 seg_start
    auto fetch = [&](int i) -> uint32_t {
      const int q = i * AB_T + (ct & 127);
      const bool ok = q < T;
      if (ct < 128) return __float_as_uint(ok ? lse[vec_base + q] * 1.4426950408889634f : 0.f);
      if (ct < 256) return __float_as_uint(ok ? delta[vec_base + q] : 0.f);
      if (ct < 384) return static_cast<uint32_t>(ok ? (seg_start ? seg_start[seq0 + q] : 0) : 0x7fffffff);
      return 0u;
    };
    auto publish = [&](int st, uint32_t v) {
      if (ct < 128) sLse[st * AB_T + ct] = __uint_as_float(v);
      else if (ct < 256) sDelta[st * AB_T + ct - 128] = __uint_as_float(v);
      else if (ct < 384) sSeg[st * AB_T + ct - 256] = static_cast<int32_t>(v);
    };
    auto dq_flush = [&](int i_tile) {  // dQ of query tile i_tile: lane r now means QUERY row r; 16 head-dim cols/thread
      float* dst = dq_acc + (seq0 + i_tile * AB_T + r) * d + h * AB_HD + cq * 16;
      const bool dq_ok = i_tile * AB_T + r < T;
      uint32_t t[16];
      tmem_ld16(tDQ + lane_off + cq * 16, t);
      tmem_ld_wait();
      if (dq_ok) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          red_add_f32x4(dst + q4 * 4, __uint_as_float(t[4 * q4]), __uint_as_float(t[4 * q4 + 1]),
                        __uint_as_float(t[4 * q4 + 2]), __uint_as_float(t[4 * q4 + 3]));
      }
    };
    publish(0, fetch(j));

    for (int it = 0; it < n_it; ++it) {
      const int i = j + it;
      const int st = it & 1;
      named_bar_sync(1, AB_CWARPS * 32);  // vectors of this step visible; everyone is done with the other buffer
      uint32_t nvec = 0;
      if (it + 1 < n_it) nvec = fetch(i + 1);  // prefetch: latency hidden behind this step's work
      const float* lse2 = sLse + st * AB_T + cq * 32;
      const float* dl = sDelta + st * AB_T + cq * 32;
      const int32_t* sg = sSeg + st * AB_T + cq * 32;
      const int qpos0 = i * AB_T + cq * 32;
      const bool need_mask = (i == j) || (sSeg[st * AB_T + AB_T - 1] > j * AB_T);

      // ---- P^T (32 query columns per thread): registers for dS^T, packed bf16 pairs into tensor memory for dV
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      float p[32];
      {
        uint32_t t[32];
        tmem_ld32(tS + lane_off + cq * 32, t);
        tmem_ld_wait();
        if (need_mask)
          bwd_p_chunk<true>(t, p, lse2, sg, kj, qpos0, scale_log2);
        else
          bwd_p_chunk<false>(t, p, lse2, sg, kj, qpos0, scale_log2);
        uint32_t w[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) w[e] = pack_bf16x2(p[2 * e], p[2 * e + 1]);
        tmem_st16(tP + lane_off + cq * 16, w);
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);

      // ---- dS^T = P^T o (dP^T - delta)      (the softmax scale is applied once, in the dK / dQ epilogues)
      mbar_wait(dp_full, it & 1);
      tc_fence_after();
      {
        uint32_t t[32];
        tmem_ld32(tDP + lane_off + cq * 32, t);
        tmem_ld_wait();
        float ds[32];
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 dv = *reinterpret_cast<const float4*>(dl + q4 * 4);
          ds[4 * q4 + 0] = p[4 * q4 + 0] * (__uint_as_float(t[4 * q4 + 0]) - dv.x);
          ds[4 * q4 + 1] = p[4 * q4 + 1] * (__uint_as_float(t[4 * q4 + 1]) - dv.y);
          ds[4 * q4 + 2] = p[4 * q4 + 2] * (__uint_as_float(t[4 * q4 + 2]) - dv.z);
          ds[4 * q4 + 3] = p[4 * q4 + 3] * (__uint_as_float(t[4 * q4 + 3]) - dv.w);
        }
        // the previous step's dK/dQ MMAs must be done reading the dS^T buffer before it is overwritten
        if (it > 0) mbar_wait(dq_full, (it - 1) & 1);
        store_bf16_row32(sDS + (cq >> 1) * AB_TILE + r * 128, r, (cq & 1) * 4, ds);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_ready);

      // ---- dQ of the previous step (after the hand-off: the red.adds drain while the tensor pipe runs dK_it)
      if (it > 0) {
        tc_fence_after();
        dq_flush(i - 1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_empty);
      }
      if (it + 1 < n_it) publish(st ^ 1, nvec);
    }

    // ---- tail: dQ of the last step, then dV / dK of this key tile (16 head-dim columns per thread)
    mbar_wait(dq_full, (n_it - 1) & 1);
    tc_fence_after();
    dq_flush(j + n_it - 1);
    const bool k_ok = kj < T;
    {
      uint32_t t[16];
      tmem_ld16(tDV + lane_off + cq * 16, t);
      tmem_ld_wait();
      if (k_ok) {
        __nv_bfloat16* dv_out = dqkv + (krow0 + r) * (3 * d) + 2 * d + h * AB_HD + cq * 16;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(t[8 * g + 0]), __uint_as_float(t[8 * g + 1]));
          v.y = pack_bf16x2(__uint_as_float(t[8 * g + 2]), __uint_as_float(t[8 * g + 3]));
          v.z = pack_bf16x2(__uint_as_float(t[8 * g + 4]), __uint_as_float(t[8 * g + 5]));
          v.w = pack_bf16x2(__uint_as_float(t[8 * g + 6]), __uint_as_float(t[8 * g + 7]));
          *reinterpret_cast<uint4*>(dv_out + g * 8) = v;
        }
      }
    }
    {
      uint32_t t[16];
      tmem_ld16(tDK + lane_off + cq * 16, t);
      tmem_ld_wait();
      if (k_ok) {
        float v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(t[e]) * scale;
        if (rope) {  // rotate back: transpose of models/embeddings.py:15-30
          const float4* tab =
              reinterpret_cast<const float4*>(rope + (static_cast<int64_t>(kj) * (AB_HD >> 1) + cq * 8) * 2);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 cs = __ldg(tab + q4);
            const float x0 = v[4 * q4 + 0], x1 = v[4 * q4 + 1], y0 = v[4 * q4 + 2], y1 = v[4 * q4 + 3];
            v[4 * q4 + 0] = x0 * cs.x + x1 * cs.y;
            v[4 * q4 + 1] = x1 * cs.x - x0 * cs.y;
            v[4 * q4 + 2] = y0 * cs.z + y1 * cs.w;
            v[4 * q4 + 3] = y1 * cs.z - y0 * cs.w;
          }
        }
        __nv_bfloat16* dk_out = dqkv + (krow0 + r) * (3 * d) + d + h * AB_HD + cq * 16;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint4 o;
          o.x = pack_bf16x2(v[8 * g + 0], v[8 * g + 1]);
          o.y = pack_bf16x2(v[8 * g + 2], v[8 * g + 3]);
          o.z = pack_bf16x2(v[8 * g + 4], v[8 * g + 5]);
          o.w = pack_bf16x2(v[8 * g + 6], v[8 * g + 7]);
          *reinterpret_cast<uint4*>(dk_out + g * 8) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == AB_CWARPS) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace plm

extern "C" int plm_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                            const int32_t* seg_start, const float* rope_table, void* dqkv, float* delta,
                            float* dq_acc, int32_t B, int32_t T, int32_t H, int32_t hd, plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(qkv);
  PLM_REQUIRE(qkv && out && dout && lse && dqkv && delta && dq_acc, "attn_bwd: null pointer");
  PLM_REQUIRE(B > 0 && T > 0 && H > 0, "attn_bwd: bad size");
  if (hd != AB_HD) return fail(PLM_ERR_UNSUPPORTED, "attn_bwd: head_dim %d unsupported (need 64)", hd);
  PLM_REQUIRE(aligned16(qkv) && aligned16(out) && aligned16(dout) && aligned16(dqkv) && aligned16(dq_acc) &&
                  (!rope_table || aligned16(rope_table)),
              "attn_bwd: misaligned pointer");
  PLM_REQUIRE(static_cast<int64_t>(B) * T < (1ll << 31) && B <= 65535 && H <= 65535, "attn_bwd: size too large");

  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
  });
  if (attr_err != cudaSuccess) return fail(PLM_ERR_CUDA, "attn_bwd smem attribute: %s", cudaGetErrorString(attr_err));

  const int d = H * hd;
  const int64_t rows = static_cast<int64_t>(B) * T;
  CUtensorMap tmQKV, tmDO;
  int rc = make_tmap_bf16_2d(&tmQKV, qkv, rows, 3ull * d, 3ull * d, AB_T, 64);
  if (rc != PLM_OK) return rc;
  rc = make_tmap_bf16_2d(&tmDO, dout, rows, d, d, AB_T, 64);
  if (rc != PLM_OK) return rc;

  cudaError_t e = cudaMemsetAsync(dq_acc, 0, static_cast<size_t>(rows) * d * sizeof(float), stream);
  if (e != cudaSuccess) return fail(PLM_ERR_CUDA, "attn_bwd memset: %s", cudaGetErrorString(e));

  {
    const int64_t n = rows * H * 8;
    attn_delta_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
        static_cast<const uint4*>(out), static_cast<const uint4*>(dout), delta, rows, T, H);
    rc = check_launch("attn_delta");
    if (rc != PLM_OK) return rc;
  }
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  dim3 grid((T + AB_T - 1) / AB_T, H, B);
  attn_bwd_kernel<<<grid, AB_THREADS, AB_SMEM, stream>>>(tmQKV, tmDO, lse, delta, seg_start, rope_table,
                                                         static_cast<__nv_bfloat16*>(dqkv), dq_acc, T, H, scale,
                                                         scale * 1.4426950408889634f);
  rc = check_launch("attn_bwd");
  if (rc != PLM_OK) return rc;
  {
    const int64_t n = rows * (d / 8);
    dq_finalize_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
        dq_acc, rope_table, static_cast<__nv_bfloat16*>(dqkv), rows, T, d, hd, scale);
    rc = check_launch("dq_finalize");
  }
  return rc;
}

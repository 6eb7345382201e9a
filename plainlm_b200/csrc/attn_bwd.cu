// Flash-attention backward on tcgen05 (backward of models/transformer.py:53-63, reached from engine/engine.py:120).
//
// One CTA per (128-key tile j, head, batch), looping over the query tiles i >= j that can see it. 768 threads:
//   warps 0..15  compute: warp = (TMEM lane quarter, column quarter); a thread owns 32 columns of key row r of the
//                transposed score tile — four warps per scheduler so TMEM / smem / MUFU latencies overlap
//   warps 16-18  MMA issuers (S^T,dP^T,dQ | dV | dK: one thread each)   warp 19 TMA producer (K,V; Q_i,dO_i 3-stage ring)
//   warps 20-23  dQ drain (one per TMEM lane quarter): tensor memory -> swizzled smem -> TMA bulk reduce-add into dq_acc
// 768 threads leave 80 registers per thread; setmaxnreg moves registers from warps 16-23 (48) to the compute warps (96).
// Five GEMMs per (j, i) pair, all on the tensor core, all 512 TMEM columns in use:
//   S^T  = K Q_i^T        (cols   0..127)      dP^T = V dO_i^T       (cols 128..255)
//   dV  += P^T dO_i       (cols 256..319)      dK  += dS^T Q_i       (cols 320..383)
//   dQ_i = dS K           (cols 384..447) -> fp32 reduce-add into dq_acc (finalised by dq_finalize_kernel)
//   P^T as packed bf16    (cols 448..511) -> A operand of the dV GEMM read straight from tensor memory
// dS^T is written once to swizzled smem as a K-major A operand (dK) and re-read MN-major for the dQ GEMM; Q_i / dO_i /
// K are consumed as MN-major B operands straight from their TMA boxes — no transposes anywhere.  S^T / dP^T of step
// i+1 are issued as soon as the compute warps have READ step i's tiles out of tensor memory, so the tensor pipe works
// on the neighbouring step while the exp / dS math runs.  dK and dQ are rotated back through RoPE in the epilogues.
// Round 1 measured ~3000 cycles per step.  Its ncu source page (profiles/r2_attn_bwd_instruction_mix.md) showed the
// kernel ISSUE-bound: 10.2 k warp-instructions per step = 2560 issue cycles per scheduler, of which only 2.1 k were the
// exp / dS math: 36 % were mbarrier polling loops of waiting warps, the rest clock64 trace hooks and per-step address
// arithmetic.  This version sleeps in hardware while waiting (ptx.cuh: mbar_try_wait), has no trace hooks and keeps
// its loop-invariant addresses in registers.
#include "common.cuh"
#include "ptx.cuh"

#include <cstdlib>
#include <mutex>

namespace plm {

constexpr int AB_T = 128;   // tile edge (keys per CTA, queries per step)
constexpr int AB_HD = 64;
constexpr int AB_CWARPS = 16;    // compute warps: (TMEM lane quarter) x (column quarter)
constexpr int AB_W_MMA_S = AB_CWARPS;       // issues S^T, dP^T (early in a step) and dQ (late in the step)
constexpr int AB_W_MMA_DV = AB_CWARPS + 1;  // issues dV
constexpr int AB_W_MMA_DK = AB_CWARPS + 2;  // issues dK
constexpr int AB_W_TMA = AB_CWARPS + 3;     // TMA producer
constexpr int AB_W_DRAIN = AB_CWARPS + 4;   // four dQ drain warps, one per TMEM lane quarter
constexpr int AB_THREADS = (AB_CWARPS + 8) * 32;
constexpr int AB_REGS_COMPUTE = 96;  // setmaxnreg: compute warpgroups take registers from the auxiliary ones
constexpr int AB_REGS_AUX = 48;
// persistent kernel: a fourth issuer (dQ) and three padding warps make 28 warps = 7 warpgroups (setmaxnreg is per
// warpgroup): 896 threads x 72 registers at launch = 64512 = 512 x 96 (compute) + 384 x 40 (everything else)
constexpr int ABP_W_MMA_S = AB_CWARPS;
constexpr int ABP_W_MMA_DV = AB_CWARPS + 1;
constexpr int ABP_W_MMA_DK = AB_CWARPS + 2;
constexpr int ABP_W_TMA = AB_CWARPS + 3;
constexpr int ABP_W_DRAIN = AB_CWARPS + 4;   // 20..23
constexpr int ABP_W_MMA_DQ = AB_CWARPS + 8;  // 24 (25..27 pad the warpgroup)
constexpr int ABP_THREADS = (AB_CWARPS + 12) * 32;
constexpr int ABP_REGS_AUX = 40;
constexpr int AB_STAGES = 3;       // Q_i / dO_i ring
constexpr int AB_TILE = AB_T * AB_HD * 2;  // 16 KB
// K, V, (Q,dO) x AB_STAGES, dS^T (2 blocks), dQ staging (fp32 128 x 64), vectors (lse2, delta, seg) x stages, barriers
constexpr int AB_VEC_BYTES = AB_STAGES * 3 * AB_T * 4;
constexpr int AB_DQ_STAGE = AB_T * AB_HD * 4;  // 32 KB: per lane quarter two [32 rows x 128 B] blocks (128-byte swizzle)
constexpr int AB_SMEM = AB_TILE * (2 + 2 * AB_STAGES + 2) + AB_DQ_STAGE + AB_VEC_BYTES + 512;

__device__ __forceinline__ float ex2b(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// delta[b,h,t] = sum_c dO[t,h,c] * O[t,h,c].  8 lanes per (row, head): each lane 8 elements.  The same pass zeroes the
// fp32 dQ accumulator (same [rows, H*64] element grid) that the main kernel reduce-adds into: one launch instead of a
// memset node + a kernel per layer.
__global__ void __launch_bounds__(256)
attn_delta_kernel(const uint4* __restrict__ o, const uint4* __restrict__ dout, float* __restrict__ delta,
                  float4* __restrict__ dq_acc, int64_t rows, int T, int H) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // one per 8 elements
  const int64_t total = rows * H * 8;
  float s = 0.f;
  if (idx < total) {
    dq_acc[2 * idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    dq_acc[2 * idx + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
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
  const float2 sc2 = make_float2(scale_log2, scale_log2);
#pragma unroll
  for (int q4 = 0; q4 < 8; ++q4) {
    const float4 l = *reinterpret_cast<const float4*>(lse2 + q4 * 4);
    // packed fp32x2 FMA (sm_100): two exp2 arguments per instruction
    const float2 a0 = __ffma2_rn(make_float2(__uint_as_float(t[4 * q4 + 0]), __uint_as_float(t[4 * q4 + 1])), sc2,
                                 make_float2(-l.x, -l.y));
    const float2 a1 = __ffma2_rn(make_float2(__uint_as_float(t[4 * q4 + 2]), __uint_as_float(t[4 * q4 + 3])), sc2,
                                 make_float2(-l.z, -l.w));
    float v[4] = {ex2b(a0.x), ex2b(a0.y), ex2b(a1.x), ex2b(a1.y)};
    if (MASK) {
      const int4 g = *reinterpret_cast<const int4*>(sg + q4 * 4);
      const int gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qi = qpos0 + q4 * 4 + e;
        if (kj > qi || kj < gv[e]) v[e] = 0.f;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) p[4 * q4 + e] = v[e];
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
                const __grid_constant__ CUtensorMap tmDQ,
                const float* __restrict__ lse, const float* __restrict__ delta, const int32_t* __restrict__ seg_start,
                const float* __restrict__ rope, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dq_acc, int T,
                int H, float scale, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = smem + AB_TILE;
  uint8_t* sQ = smem + 2 * AB_TILE;                    // [AB_STAGES]
  uint8_t* sDO = smem + (2 + AB_STAGES) * AB_TILE;     // [AB_STAGES]
  uint8_t* sDS = smem + (2 + 2 * AB_STAGES) * AB_TILE; // dS^T: 2 blocks (q 0..63 | 64..127), each [128 kv rows x 128 B]
  uint8_t* sDQ = smem + (4 + 2 * AB_STAGES) * AB_TILE;  // dQ_i staging for the TMA reduce-add
  float* sLse = reinterpret_cast<float*>(sDQ + AB_DQ_STAGE);  // [AB_STAGES][128]  lse * log2(e)
  float* sDelta = sLse + AB_STAGES * AB_T;                                         // [AB_STAGES][128]
  int32_t* sSeg = reinterpret_cast<int32_t*>(sDelta + AB_STAGES * AB_T);           // [AB_STAGES][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDQ + AB_DQ_STAGE + AB_VEC_BYTES);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;                // [AB_STAGES]  TMA bytes of Q_i, dO_i + the 32 staging lanes
  uint64_t* qdo_empty = bars + 1 + AB_STAGES;   // [AB_STAGES]
  uint64_t* s_full = bars + 1 + 2 * AB_STAGES;  // S^T and dP^T of a step are in tensor memory
  uint64_t* dq_full = s_full + 1;               // dQ MMAs of a step complete
  uint64_t* dq_empty = s_full + 2;              // dQ of a step has been read out of tensor memory
  uint64_t* sdp_free = s_full + 3;              // compute warps have read S^T and dP^T out of tensor memory
  uint64_t* pds_ready = s_full + 4;  // P^T (tensor memory) and dS^T (smem) of a step are written (16 warps)
  uint64_t* mma_done = s_full + 5;   // the dV, dK and dQ GEMMs of a step are complete (3 commits): P^T columns and the
                                     // dS^T buffer may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 6);

  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128-byte-swizzle layout contract violated: fail the launch loudly

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
    tma_prefetch_desc(&tmDQ);
    mbar_init(kv_full, 1);
    for (int s = 0; s < AB_STAGES; ++s) {
      mbar_init(&qdo_full[s], 1 + 32);  // expect_tx arrive + one arrive per staging lane
      mbar_init(&qdo_empty[s], 2);      // released by the dV issuer (dO_i) and by the dK issuer (Q_i)
    }
    mbar_init(s_full, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 4);  // the four drain warps
    mbar_init(sdp_free, AB_CWARPS);
    mbar_init(pds_ready, AB_CWARPS);
    mbar_init(mma_done, 3);
    fence_barrier_init();
  }
  if (warp == AB_W_MMA_S) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // 768 threads leave 80 registers per thread; the compute warpgroups need 96 and take them from the auxiliary ones
  // (setmaxnreg at the top of each role's branch)
  // TMEM columns: S^T 0..127 | dP^T 128..255 | dV 256..319 | dK 320..383 | dQ 384..447 | P^T (bf16 pairs) 448..511
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320,
                 tDQ = tmem_base + 384, tP = tmem_base + 448;

  if (warp == AB_W_TMA) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AB_REGS_AUX));
    // ------------------------------------------------------------ producer warp: lane 0 drives TMA (K,V once; Q_i,dO_i
    // ring), all 32 lanes stage the per-query vectors (lse*log2e, delta, seg_start) of the step next to its tiles.
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * AB_TILE);
      tma_load_2d(sK, &tmQKV, kv_full, d + h * AB_HD, static_cast<int>(krow0));
      tma_load_2d(sV, &tmQKV, kv_full, 2 * d + h * AB_HD, static_cast<int>(krow0));
    }
    const int64_t vec_base = (static_cast<int64_t>(b) * H + h) * T;
    int st = 0;
    for (int it = 0; it < n_it; ++it) {
      mbar_wait(&qdo_empty[st], ((it / AB_STAGES) & 1) ^ 1);
      if (lane == 0) {
        const int qr = static_cast<int>(seq0 + (j + it) * AB_T);
        mbar_arrive_expect_tx(&qdo_full[st], 2 * AB_TILE);
        tma_load_2d(sQ + st * AB_TILE, &tmQKV, &qdo_full[st], h * AB_HD, qr);
        tma_load_2d(sDO + st * AB_TILE, &tmDO, &qdo_full[st], h * AB_HD, qr);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qq = e * 32 + lane;
        const int q = (j + it) * AB_T + qq;
        const bool ok = q < T;
        sLse[st * AB_T + qq] = ok ? lse[vec_base + q] * 1.4426950408889634f : 0.f;
        sDelta[st * AB_T + qq] = ok ? delta[vec_base + q] : 0.f;
        sSeg[st * AB_T + qq] = ok ? (seg_start ? seg_start[seq0 + q] : 0) : 0x7fffffff;  // q >= T: nothing allowed
      }
      mbar_arrive(&qdo_full[st]);  // release: the vectors are visible to whoever acquires the barrier
      if (++st == AB_STAGES) st = 0;
    }
  } else if (warp == AB_W_MMA_S) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AB_REGS_AUX));
    // ------------------------------------------------------------ MMA issuer 1: S^T = K Q^T, dP^T = V dO^T, dQ = dS K.
    // S^T / dP^T of step it+1 only wait for step it's tiles to be READ out of tensor memory (sdp_free, early in the
    // step); dQ of step it waits for its dS^T (pds_ready, late in the step) and for the previous dQ to have left tensor
    // memory (dq_empty, signalled by the drain warps) — the two jobs never compete for this thread.  dQ: A = the dS^T buffer read MN-major (M =
    // queries, 64 per block, blocks AB_TILE apart), B = K MN-major.
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // K-major x K-major, N = 128 queries
      constexpr uint32_t idesc_nn = make_idesc_bf16(128, 64, 1, 1);   // A MN-major, B MN-major (dQ)
      mbar_wait(kv_full, 0);
      // descriptors are built once; per K-step only the 14-bit start-address field advances (tight issue loop)
      const uint64_t k_desc = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t v_desc = make_smem_desc_sw128(smem_u32(sV), 16, 1024);
      const uint64_t q_desc0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t do_desc0 = make_smem_desc_sw128(smem_u32(sDO), 16, 1024);
      const uint64_t k_desc_mn = make_smem_desc_sw128(smem_u32(sK), AB_TILE, 1024);
      const uint64_t ds_desc_mn = make_smem_desc_sw128(smem_u32(sDS), AB_TILE, 1024);
      int st = 0;
      for (int it = 0; it <= n_it; ++it) {
        if (it < n_it) {  // S^T / dP^T of step `it`
          const uint64_t q_desc = q_desc0 + st * (AB_TILE >> 4), do_desc = do_desc0 + st * (AB_TILE >> 4);
          mbar_wait(&qdo_full[st], (it / AB_STAGES) & 1);
          if (it > 0) mbar_wait(sdp_free, (it - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < AB_HD / 16; ++k) {
            umma_ss(tS, k_desc + k * 2, q_desc + k * 2, idesc_s, k > 0 ? 1u : 0u);
            umma_ss(tDP, v_desc + k * 2, do_desc + k * 2, idesc_s, k > 0 ? 1u : 0u);
          }
          umma_commit(s_full);
          if (++st == AB_STAGES) st = 0;
        }
        if (it > 0) {  // dQ of step it-1
          const int pit = it - 1;
          mbar_wait(pds_ready, pit & 1);
          if (pit > 0) mbar_wait(dq_empty, (pit - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < AB_T / 16; ++k)
            umma_ss(tDQ, ds_desc_mn + k * (2048 >> 4), k_desc_mn + k * (2048 >> 4), idesc_nn, k > 0 ? 1u : 0u);
          umma_commit(dq_full);
          umma_commit(mma_done);
        }
      }
    }
  } else if (warp == AB_W_MMA_DV) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AB_REGS_AUX));
    // ------------------------------------------------------------ MMA issuer 2: dV += P^T dO (A = P^T from tensor memory)
    if (lane == 0) {
      constexpr uint32_t idesc_kn = make_idesc_bf16(128, 64, 0, 1);   // A K-major (TMEM), B MN-major, N = 64
      const uint64_t do_desc0 = make_smem_desc_sw128(smem_u32(sDO), AB_TILE, 1024);
      int st = 0;
      for (int it = 0; it < n_it; ++it) {
        const uint64_t do_desc = do_desc0 + st * (AB_TILE >> 4);
        mbar_wait(pds_ready, it & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AB_T / 16; ++k)
          umma_ts(tDV, tP + k * 8, do_desc + k * (2048 >> 4), idesc_kn, (it > 0 || k > 0) ? 1u : 0u);
        umma_commit(mma_done);
        umma_commit(&qdo_empty[st]);
        if (++st == AB_STAGES) st = 0;
      }
    }
  } else if (warp == AB_W_MMA_DK) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AB_REGS_AUX));
    // ------------------------------------------------------------ MMA issuer 3: dK += dS^T Q
    if (lane == 0) {
      constexpr uint32_t idesc_kn = make_idesc_bf16(128, 64, 0, 1);   // A K-major, B MN-major, N = 64
      const uint64_t ds_desc_k = make_smem_desc_sw128(smem_u32(sDS), 16, 1024);        // K-major view of dS^T
      const uint64_t q_desc0 = make_smem_desc_sw128(smem_u32(sQ), AB_TILE, 1024);
      int st = 0;
      for (int it = 0; it < n_it; ++it) {
        const uint64_t q_desc = q_desc0 + st * (AB_TILE >> 4);
        mbar_wait(pds_ready, it & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AB_T / 16; ++k)
          umma_ss(tDK, ds_desc_k + ((k >> 2) * AB_TILE + (k & 3) * 32) / 16, q_desc + k * (2048 >> 4), idesc_kn,
                  (it > 0 || k > 0) ? 1u : 0u);
        umma_commit(mma_done);
        umma_commit(&qdo_empty[st]);
        if (++st == AB_STAGES) st = 0;
      }
    }
  } else if (warp >= AB_W_DRAIN) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AB_REGS_AUX));
    // ------------------------------------------------------------ dQ drain warps (one per TMEM lane quarter), off the
    // compute warps' serial chain: dQ_i of a step leaves tensor memory (-> dq_empty: the next dQ GEMM may start), is
    // staged as two 128B-swizzled [32 rows x 32] fp32 blocks and added into dq_acc by two TMA bulk reduce-adds.
    // Lane r = QUERY row r.  Rows beyond T hold exact zeros (P is masked to 0 there), so adding them is harmless.
    const int quarter = warp & 3;
    uint8_t* blk = sDQ + quarter * (AB_DQ_STAGE / 4);
    for (int it = 0; it < n_it; ++it) {
      mbar_wait(dq_full, it & 1);
      tc_fence_after();
      if (lane == 0) bulk_wait_group_read<0>();  // the previous step's reduce-adds have finished reading the blocks
      __syncwarp();
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {  // 16 head-dim columns at a time (these warps run on 48 registers)
        uint32_t t[16];
        tmem_ld16(tDQ + (static_cast<uint32_t>(quarter * 32) << 16) + qq * 16, t);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(blk + (qq >> 1) * 4096 + lane * 128 + ((((qq & 1) * 4 + c) ^ (lane & 7)) << 4)) =
              make_uint4(t[4 * c], t[4 * c + 1], t[4 * c + 2], t[4 * c + 3]);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(dq_empty);
        const int row0 = static_cast<int>(seq0 + (j + it) * AB_T + quarter * 32);
        tma_reduce_add_2d(&tmDQ, blk, h * AB_HD, row0);
        tma_reduce_add_2d(&tmDQ, blk + 4096, h * AB_HD + 32, row0);
        bulk_commit_group();
      }
    }
    if (lane == 0) bulk_wait_group<0>();  // the reduce-adds have landed
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(AB_REGS_COMPUTE));
    // ------------------------------------------------------------ compute warps (16): warp = (lane quarter, column quarter)
    const int quarter = warp & 3;
    const int cq = warp >> 2;                // query columns [32*cq, +32) of S^T / dP^T; hd cols [16*cq, +16) of dQ/dK/dV
    const int r = quarter * 32 + lane;       // key row within the tile (S^T lane) / query row within the tile (dQ lane)
    const int kj = j * AB_T + r;             // key position
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;

    int st = 0;
    for (int it = 0; it < n_it; ++it) {
      const int i = j + it;
      // one hand-off in: tiles of this step are in tensor memory (s_full), its vectors are staged (qdo_full)
      mbar_wait(&qdo_full[st], (it / AB_STAGES) & 1);
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      const float* lse2 = sLse + st * AB_T + cq * 32;
      const float* dl = sDelta + st * AB_T + cq * 32;
      const int32_t* sg = sSeg + st * AB_T + cq * 32;
      const int qpos0 = i * AB_T + cq * 32;
      const bool need_mask = (i == j) || (sSeg[st * AB_T + AB_T - 1] > j * AB_T);

      uint32_t ts[32], tdp[32];
      tmem_ld32(tS + lane_off + cq * 32, ts);
      tmem_ld32(tDP + lane_off + cq * 32, tdp);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_free);  // S^T / dP^T of the next step may be issued while we do the math

      // P^T = exp2(S^T * scale*log2e - lse*log2e), masked;  dS^T = P^T o (dP^T - delta)  (softmax scale applied once,
      // in the dK / dQ epilogues)
      float p[32];
      // (A per-scheduler FIFO lock making the exp2 bursts of the four warps of a scheduler exclusive was measured and
      // removed: 0.344 -> 0.390 ms; see attn_fwd.cu.)
      if (need_mask)
        bwd_p_chunk<true>(ts, p, lse2, sg, kj, qpos0, scale_log2);
      else
        bwd_p_chunk<false>(ts, p, lse2, sg, kj, qpos0, scale_log2);
      uint32_t w[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) w[e] = pack_bf16x2(p[2 * e], p[2 * e + 1]);
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        const float4 dv = *reinterpret_cast<const float4*>(dl + q4 * 4);
        const float2 d0 = __fadd2_rn(make_float2(__uint_as_float(tdp[4 * q4 + 0]), __uint_as_float(tdp[4 * q4 + 1])),
                                     make_float2(-dv.x, -dv.y));
        const float2 d1 = __fadd2_rn(make_float2(__uint_as_float(tdp[4 * q4 + 2]), __uint_as_float(tdp[4 * q4 + 3])),
                                     make_float2(-dv.z, -dv.w));
        const float2 r0 = __fmul2_rn(make_float2(p[4 * q4 + 0], p[4 * q4 + 1]), d0);
        const float2 r1 = __fmul2_rn(make_float2(p[4 * q4 + 2], p[4 * q4 + 3]), d1);
        p[4 * q4 + 0] = r0.x;
        p[4 * q4 + 1] = r0.y;
        p[4 * q4 + 2] = r1.x;
        p[4 * q4 + 3] = r1.y;
      }
      // the previous step's dV MMAs must be done with this warp's P^T columns, its dK / dQ MMAs with its dS^T block
      if (it > 0) {
        mbar_wait(mma_done, (it - 1) & 1);
        tc_fence_after();
      }
      tmem_st16(tP + lane_off + cq * 16, w);
      store_bf16_row32(sDS + (cq >> 1) * AB_TILE + r * 128, r, (cq & 1) * 4, p);
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_ready);  // one hand-off out
      if (++st == AB_STAGES) st = 0;
    }

    // ---- tail: dV / dK of this key tile (16 head-dim columns per thread)
    mbar_wait(mma_done, (n_it - 1) & 1);
    tc_fence_after();
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
  if (warp == AB_W_MMA_S) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// =====================================================================================================================
// Persistent variant (variant 1).  Per-CTA costs of the kernel above — launch, TMEM allocation, barrier setup, the K/V
// and first Q/dO round trips, the pipeline fill and the dK/dV epilogue — are paid 2048 times for 17408 steps (B = 8,
// H = 16, T = 2048): measured ~4600 cycles per step against ~3000 in steady state.  Here one CTA per SM walks the
// (key tile, head, batch) items in snake order over the heaviest-first list; tensor memory, barriers and the Q/dO ring
// live across items (all parities run on global step counters), the producer prefetches the next item's Q/dO tiles
// while the current one drains, and the dK/dV epilogue of item n overlaps the K/V fetch of item n+1.
__global__ void __launch_bounds__(ABP_THREADS, 1)
attn_bwd_persistent_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmDQ,
                const float* __restrict__ lse, const float* __restrict__ delta, const int32_t* __restrict__ seg_start,
                const float* __restrict__ rope, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dq_acc, int T,
                int H, int BH, float scale, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = smem + AB_TILE;
  uint8_t* sQ = smem + 2 * AB_TILE;                    // [AB_STAGES]
  uint8_t* sDO = smem + (2 + AB_STAGES) * AB_TILE;     // [AB_STAGES]
  uint8_t* sDS = smem + (2 + 2 * AB_STAGES) * AB_TILE; // dS^T: 2 blocks (q 0..63 | 64..127), each [128 kv rows x 128 B]
  uint8_t* sDQ = smem + (4 + 2 * AB_STAGES) * AB_TILE;  // dQ_i staging for the TMA reduce-add
  float* sLse = reinterpret_cast<float*>(sDQ + AB_DQ_STAGE);  // [AB_STAGES][128]  lse * log2(e)
  float* sDelta = sLse + AB_STAGES * AB_T;                                         // [AB_STAGES][128]
  int32_t* sSeg = reinterpret_cast<int32_t*>(sDelta + AB_STAGES * AB_T);           // [AB_STAGES][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDQ + AB_DQ_STAGE + AB_VEC_BYTES);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;                // [AB_STAGES]  TMA bytes of Q_i, dO_i + the 32 staging lanes
  uint64_t* qdo_empty = bars + 1 + AB_STAGES;   // [AB_STAGES]
  uint64_t* s_full = bars + 1 + 2 * AB_STAGES;  // S^T and dP^T of a step are in tensor memory
  uint64_t* dq_full = s_full + 1;               // dQ MMAs of a step complete
  uint64_t* dq_empty = s_full + 2;              // dQ of a step has been read out of tensor memory
  uint64_t* sdp_free = s_full + 3;              // compute warps have read S^T and dP^T out of tensor memory
  uint64_t* pds_ready = s_full + 4;  // P^T (tensor memory) and dS^T (smem) of a step are written (16 warps)
  uint64_t* mma_done = s_full + 5;   // the dV, dK and dQ GEMMs of a step are complete (3 commits): P^T columns and the
                                     // dS^T buffer may be overwritten
  uint64_t* kv_empty = s_full + 6;   // every MMA that reads the item's K / V tiles has completed (persistent kernel)
  uint64_t* acc_free = s_full + 7;   // the dV / dK accumulators of the item have been read out (16 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 8);

  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128-byte-swizzle layout contract violated: fail the launch loudly

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nq = (T + AB_T - 1) / AB_T;  // ragged tail: T need not be a multiple of 128
  const int d = H * AB_HD;
  // Persistent schedule: one CTA per SM walks the (key tile, head, batch) items in snake order over the heaviest-first
  // list (key tile 0 sees every query tile).  k-th item of this CTA, or n_it = 0 when the list is exhausted.
  const int n_items = nq * BH;
  const int rounds = (n_items + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  struct Item {
    int j, h, b, n_it;
    int64_t seq0, krow0;
  };
  auto get_item = [&](int k, Item& it_) -> bool {
    const int G = static_cast<int>(gridDim.x), c = static_cast<int>(blockIdx.x);
    const int idx = k * G + ((k & 1) ? (G - 1 - c) : c);
    if (idx >= n_items) return false;
    it_.j = idx / BH;
    const int bh = idx - it_.j * BH;
    it_.b = bh / H;
    it_.h = bh - it_.b * H;
    it_.seq0 = static_cast<int64_t>(it_.b) * T;
    it_.krow0 = it_.seq0 + it_.j * AB_T;
    int i_hi = nq - 1;  // last query tile that can see this key tile: seg_start is non-decreasing in t
    if (seg_start) {
      const int k_last = it_.j * AB_T + AB_T - 1;
      while (i_hi > it_.j && __ldg(seg_start + it_.seq0 + i_hi * AB_T) > k_last) --i_hi;
    }
    it_.n_it = i_hi - it_.j + 1;
    return true;
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQ);
    mbar_init(kv_full, 1);
    for (int s = 0; s < AB_STAGES; ++s) {
      mbar_init(&qdo_full[s], 1 + 32);  // expect_tx arrive + one arrive per staging lane
      mbar_init(&qdo_empty[s], 2);      // released by the dV issuer (dO_i) and by the dK issuer (Q_i)
    }
    mbar_init(s_full, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 4);  // the four drain warps
    mbar_init(sdp_free, AB_CWARPS);
    mbar_init(pds_ready, AB_CWARPS);
    mbar_init(mma_done, 3);
    mbar_init(kv_empty, 2);  // the S^T/dP^T issuer and the dQ issuer
    mbar_init(acc_free, AB_CWARPS);
    fence_barrier_init();
  }
  if (warp == ABP_W_MMA_S) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // 768 threads leave 80 registers per thread; the compute warpgroups need 96 and take them from the auxiliary ones
  // (setmaxnreg at the top of each role's branch)
  // TMEM columns: S^T 0..127 | dP^T 128..255 | dV 256..319 | dK 320..383 | dQ 384..447 | P^T (bf16 pairs) 448..511
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320,
                 tDQ = tmem_base + 384, tP = tmem_base + 448;

  if (warp == ABP_W_TMA) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ABP_REGS_AUX));
    // ------------------------------------------------------------ producer warp: lane 0 drives the Q_i, dO_i ring (K, V
    // are fetched by the S^T issuer, which alone knows when they are free), all 32 lanes stage the per-query vectors (lse*log2e, delta, seg_start) of the step next to its tiles.
    uint32_t g = 0;      // global step counter (ring slot and parities run across items)
    for (int k = 0; k < rounds; ++k) {
      Item im;
      if (!get_item(k, im)) continue;
      const int64_t vec_base = (static_cast<int64_t>(im.b) * H + im.h) * T;
      for (int it = 0; it < im.n_it; ++it, ++g) {
        const uint32_t st = g % AB_STAGES;
        mbar_wait(&qdo_empty[st], ((g / AB_STAGES) & 1) ^ 1);
        if (lane == 0) {
          const int qr = static_cast<int>(im.seq0 + (im.j + it) * AB_T);
          mbar_arrive_expect_tx(&qdo_full[st], 2 * AB_TILE);
          tma_load_2d(sQ + st * AB_TILE, &tmQKV, &qdo_full[st], im.h * AB_HD, qr);
          tma_load_2d(sDO + st * AB_TILE, &tmDO, &qdo_full[st], im.h * AB_HD, qr);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qq = e * 32 + lane;
          const int q = (im.j + it) * AB_T + qq;
          const bool ok = q < T;
          sLse[st * AB_T + qq] = ok ? lse[vec_base + q] * 1.4426950408889634f : 0.f;
          sDelta[st * AB_T + qq] = ok ? delta[vec_base + q] : 0.f;
          sSeg[st * AB_T + qq] = ok ? (seg_start ? seg_start[im.seq0 + q] : 0) : 0x7fffffff;  // q >= T: nothing allowed
        }
        mbar_arrive(&qdo_full[st]);  // release: the vectors are visible to whoever acquires the barrier
      }
    }
  } else if (warp == ABP_W_MMA_S) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ABP_REGS_AUX));
    // ------------------------------------------------------------ MMA issuer 1: S^T = K Q^T, dP^T = V dO^T.
    // S^T / dP^T of step g+1 only wait for step g's tiles to be READ out of tensor memory (sdp_free, early in the
    // step).  (In the non-persistent kernel this thread also issues dQ; a tcgen05.mma costs its issuer ~90 cycles, and
    // with 16 per step the next step's scores arrived ~650 cycles after the compute warps wanted them.)
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // K-major x K-major, N = 128 queries
      const uint64_t k_desc = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t v_desc = make_smem_desc_sw128(smem_u32(sV), 16, 1024);
      const uint64_t q_desc0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t do_desc0 = make_smem_desc_sw128(smem_u32(sDO), 16, 1024);
      uint32_t g = 0;    // S^T / dP^T pairs issued so far
      uint32_t items = 0;
      for (int k = 0; k < rounds; ++k) {
        Item im;
        if (!get_item(k, im)) continue;
        // K and V of this item are fetched by this thread: it (and the dQ issuer) issue every GEMM that reads them,
        // and both commit to kv_empty behind their last one of an item.  The producer warp meanwhile runs ahead with
        // the Q / dO ring of this item.
        if (items > 0) mbar_wait(kv_empty, (items - 1) & 1);
        mbar_arrive_expect_tx(kv_full, 2 * AB_TILE);
        tma_load_2d(sK, &tmQKV, kv_full, d + im.h * AB_HD, static_cast<int>(im.krow0));
        tma_load_2d(sV, &tmQKV, kv_full, 2 * d + im.h * AB_HD, static_cast<int>(im.krow0));
        mbar_wait(kv_full, items & 1);
        ++items;
        for (int it = 0; it < im.n_it; ++it, ++g) {
          const uint32_t st = g % AB_STAGES;
          const uint64_t q_desc = q_desc0 + st * (AB_TILE >> 4), do_desc = do_desc0 + st * (AB_TILE >> 4);
          mbar_wait(&qdo_full[st], (g / AB_STAGES) & 1);
          if (g > 0) mbar_wait(sdp_free, (g - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < AB_HD / 16; ++kk) {
            umma_ss(tS, k_desc + kk * 2, q_desc + kk * 2, idesc_s, kk > 0 ? 1u : 0u);
            umma_ss(tDP, v_desc + kk * 2, do_desc + kk * 2, idesc_s, kk > 0 ? 1u : 0u);
          }
          umma_commit(s_full);
        }
        umma_commit(kv_empty);
      }
    }
  } else if (warp == ABP_W_MMA_DV) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ABP_REGS_AUX));
    // ------------------------------------------------------------ MMA issuer 2: dV += P^T dO (A = P^T from tensor memory)
    if (lane == 0) {
      constexpr uint32_t idesc_kn = make_idesc_bf16(128, 64, 0, 1);   // A K-major (TMEM), B MN-major, N = 64
      const uint64_t do_desc0 = make_smem_desc_sw128(smem_u32(sDO), AB_TILE, 1024);
      uint32_t g = 0, items = 0;
      for (int k = 0; k < rounds; ++k) {
        Item im;
        if (!get_item(k, im)) continue;
        for (int it = 0; it < im.n_it; ++it, ++g) {
          const uint32_t st = g % AB_STAGES;
          const uint64_t do_desc = do_desc0 + st * (AB_TILE >> 4);
          mbar_wait(pds_ready, g & 1);
          if (it == 0 && items > 0) mbar_wait(acc_free, (items - 1) & 1);  // the previous item's dV has been read out
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < AB_T / 16; ++kk)
            umma_ts(tDV, tP + kk * 8, do_desc + kk * (2048 >> 4), idesc_kn, (it > 0 || kk > 0) ? 1u : 0u);
          umma_commit(mma_done);
          umma_commit(&qdo_empty[st]);
        }
        ++items;
      }
    }
  } else if (warp == ABP_W_MMA_DK) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ABP_REGS_AUX));
    // ------------------------------------------------------------ MMA issuer 3: dK += dS^T Q
    if (lane == 0) {
      constexpr uint32_t idesc_kn = make_idesc_bf16(128, 64, 0, 1);   // A K-major, B MN-major, N = 64
      const uint64_t ds_desc_k = make_smem_desc_sw128(smem_u32(sDS), 16, 1024);        // K-major view of dS^T
      const uint64_t q_desc0 = make_smem_desc_sw128(smem_u32(sQ), AB_TILE, 1024);
      uint32_t g = 0, items = 0;
      for (int k = 0; k < rounds; ++k) {
        Item im;
        if (!get_item(k, im)) continue;
        for (int it = 0; it < im.n_it; ++it, ++g) {
          const uint32_t st = g % AB_STAGES;
          const uint64_t q_desc = q_desc0 + st * (AB_TILE >> 4);
          mbar_wait(pds_ready, g & 1);
          if (it == 0 && items > 0) mbar_wait(acc_free, (items - 1) & 1);  // the previous item's dK has been read out
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < AB_T / 16; ++kk)
            umma_ss(tDK, ds_desc_k + ((kk >> 2) * AB_TILE + (kk & 3) * 32) / 16, q_desc + kk * (2048 >> 4), idesc_kn,
                    (it > 0 || kk > 0) ? 1u : 0u);
          umma_commit(mma_done);
          umma_commit(&qdo_empty[st]);
        }
        ++items;
      }
    }
  } else if (warp >= ABP_W_DRAIN && warp < ABP_W_DRAIN + 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ABP_REGS_AUX));
    // ------------------------------------------------------------ dQ drain warps (one per TMEM lane quarter), off the
    // compute warps' serial chain: dQ_i of a step leaves tensor memory (-> dq_empty: the next dQ GEMM may start), is
    // staged as two 128B-swizzled [32 rows x 32] fp32 blocks and added into dq_acc by two TMA bulk reduce-adds.
    // Lane r = QUERY row r.  Rows beyond T hold exact zeros (P is masked to 0 there), so adding them is harmless.
    const int quarter = warp & 3;
    uint8_t* blk = sDQ + quarter * (AB_DQ_STAGE / 4);
    uint32_t g = 0;
    for (int k = 0; k < rounds; ++k) {
      Item im;
      if (!get_item(k, im)) continue;
      for (int it = 0; it < im.n_it; ++it, ++g) {
        mbar_wait(dq_full, g & 1);
        tc_fence_after();
        if (lane == 0) bulk_wait_group_read<0>();  // the previous step's reduce-adds have finished reading the blocks
        __syncwarp();
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {  // 16 head-dim columns at a time (these warps run on 48 registers)
          uint32_t t[16];
          tmem_ld16(tDQ + (static_cast<uint32_t>(quarter * 32) << 16) + qq * 16, t);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(blk + (qq >> 1) * 4096 + lane * 128 + ((((qq & 1) * 4 + c) ^ (lane & 7)) << 4)) =
                make_uint4(t[4 * c], t[4 * c + 1], t[4 * c + 2], t[4 * c + 3]);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(dq_empty);
          const int row0 = static_cast<int>(im.seq0 + (im.j + it) * AB_T + quarter * 32);
          tma_reduce_add_2d(&tmDQ, blk, im.h * AB_HD, row0);
          tma_reduce_add_2d(&tmDQ, blk + 4096, im.h * AB_HD + 32, row0);
          bulk_commit_group();
        }
      }
    }
    if (lane == 0) bulk_wait_group<0>();  // the reduce-adds have landed
  } else if (warp >= ABP_W_MMA_DQ) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ABP_REGS_AUX));  // dQ issuer + padding warps
if (warp == ABP_W_MMA_DQ && lane == 0) {
      // ---------------------------------------------------------- MMA issuer 4: dQ = dS K.  A = the dS^T buffer read
      // MN-major (M = queries, 64 per block, blocks AB_TILE apart), B = K MN-major.  Waits for the step's dS^T
      // (pds_ready) and for the previous dQ to have left tensor memory (dq_empty, signalled by the drain warps).
      constexpr uint32_t idesc_nn = make_idesc_bf16(128, 64, 1, 1);   // A MN-major, B MN-major
      const uint64_t k_desc_mn = make_smem_desc_sw128(smem_u32(sK), AB_TILE, 1024);
      const uint64_t ds_desc_mn = make_smem_desc_sw128(smem_u32(sDS), AB_TILE, 1024);
      uint32_t gq = 0;
      for (int k = 0; k < rounds; ++k) {
        Item im;
        if (!get_item(k, im)) continue;
        for (int it = 0; it < im.n_it; ++it, ++gq) {
          mbar_wait(pds_ready, gq & 1);
          if (gq > 0) mbar_wait(dq_empty, (gq - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < AB_T / 16; ++kk)
            umma_ss(tDQ, ds_desc_mn + kk * (2048 >> 4), k_desc_mn + kk * (2048 >> 4), idesc_nn, kk > 0 ? 1u : 0u);
          umma_commit(dq_full);
          umma_commit(mma_done);
        }
        umma_commit(kv_empty);
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(AB_REGS_COMPUTE));
    // ------------------------------------------------------------ compute warps (16): warp = (lane quarter, column quarter)
    const int quarter = warp & 3;
    const int cq = warp >> 2;                // query columns [32*cq, +32) of S^T / dP^T; hd cols [16*cq, +16) of dQ/dK/dV
    const int r = quarter * 32 + lane;       // key row within the tile (S^T lane) / query row within the tile (dQ lane)
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t g = 0;  // global step counter

    for (int k = 0; k < rounds; ++k) {
      Item im;
      if (!get_item(k, im)) continue;
      const int j = im.j;
      const int kj = j * AB_T + r;             // key position
      for (int it = 0; it < im.n_it; ++it, ++g) {
        const int i = j + it;
        const uint32_t st = g % AB_STAGES;
        // one hand-off in: tiles of this step are in tensor memory (s_full), its vectors are staged (qdo_full)
        mbar_wait(&qdo_full[st], (g / AB_STAGES) & 1);
        mbar_wait(s_full, g & 1);
        tc_fence_after();
        const float* lse2 = sLse + st * AB_T + cq * 32;
        const float* dl = sDelta + st * AB_T + cq * 32;
        const int32_t* sg = sSeg + st * AB_T + cq * 32;
        const int qpos0 = i * AB_T + cq * 32;
        const bool need_mask = (i == j) || (sSeg[st * AB_T + AB_T - 1] > j * AB_T);

        uint32_t ts[32], tdp[32];
        tmem_ld32(tS + lane_off + cq * 32, ts);
        tmem_ld32(tDP + lane_off + cq * 32, tdp);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sdp_free);  // S^T / dP^T of the next step may be issued while we do the math

        // P^T = exp2(S^T * scale*log2e - lse*log2e), masked;  dS^T = P^T o (dP^T - delta)  (softmax scale applied once,
        // in the dK / dQ epilogues)
        float p[32];
        if (need_mask)
          bwd_p_chunk<true>(ts, p, lse2, sg, kj, qpos0, scale_log2);
        else
          bwd_p_chunk<false>(ts, p, lse2, sg, kj, qpos0, scale_log2);
        uint32_t w[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) w[e] = pack_bf16x2(p[2 * e], p[2 * e + 1]);
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 dv = *reinterpret_cast<const float4*>(dl + q4 * 4);
          const float2 d0 = __fadd2_rn(make_float2(__uint_as_float(tdp[4 * q4 + 0]), __uint_as_float(tdp[4 * q4 + 1])),
                                       make_float2(-dv.x, -dv.y));
          const float2 d1 = __fadd2_rn(make_float2(__uint_as_float(tdp[4 * q4 + 2]), __uint_as_float(tdp[4 * q4 + 3])),
                                       make_float2(-dv.z, -dv.w));
          const float2 r0 = __fmul2_rn(make_float2(p[4 * q4 + 0], p[4 * q4 + 1]), d0);
          const float2 r1 = __fmul2_rn(make_float2(p[4 * q4 + 2], p[4 * q4 + 3]), d1);
          p[4 * q4 + 0] = r0.x;
          p[4 * q4 + 1] = r0.y;
          p[4 * q4 + 2] = r1.x;
          p[4 * q4 + 3] = r1.y;
        }
        // the previous step's dV MMAs must be done with this warp's P^T columns, its dK / dQ MMAs with its dS^T block
        // (the previous step may belong to the previous item: its tail below has already waited for the same phase)
        if (g > 0) {
          mbar_wait(mma_done, (g - 1) & 1);
          tc_fence_after();
        }
        tmem_st16(tP + lane_off + cq * 16, w);
        store_bf16_row32(sDS + (cq >> 1) * AB_TILE + r * 128, r, (cq & 1) * 4, p);
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pds_ready);  // one hand-off out
      }

      // ---- item tail: dV / dK of this key tile (16 head-dim columns per thread)
      mbar_wait(mma_done, (g - 1) & 1);
      tc_fence_after();
      const bool k_ok = kj < T;
      uint32_t tv[16], tk[16];
      tmem_ld16(tDV + lane_off + cq * 16, tv);
      tmem_ld16(tDK + lane_off + cq * 16, tk);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free);  // the next item's first dV / dK GEMMs may overwrite the accumulators
      if (k_ok) {
        __nv_bfloat16* dv_out = dqkv + (im.krow0 + r) * (3 * d) + 2 * d + im.h * AB_HD + cq * 16;
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(tv[8 * gg + 0]), __uint_as_float(tv[8 * gg + 1]));
          v.y = pack_bf16x2(__uint_as_float(tv[8 * gg + 2]), __uint_as_float(tv[8 * gg + 3]));
          v.z = pack_bf16x2(__uint_as_float(tv[8 * gg + 4]), __uint_as_float(tv[8 * gg + 5]));
          v.w = pack_bf16x2(__uint_as_float(tv[8 * gg + 6]), __uint_as_float(tv[8 * gg + 7]));
          *reinterpret_cast<uint4*>(dv_out + gg * 8) = v;
        }
        float v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(tk[e]) * scale;
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
        __nv_bfloat16* dk_out = dqkv + (im.krow0 + r) * (3 * d) + d + im.h * AB_HD + cq * 16;
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
          uint4 o;
          o.x = pack_bf16x2(v[8 * gg + 0], v[8 * gg + 1]);
          o.y = pack_bf16x2(v[8 * gg + 2], v[8 * gg + 3]);
          o.z = pack_bf16x2(v[8 * gg + 4], v[8 * gg + 5]);
          o.w = pack_bf16x2(v[8 * gg + 6], v[8 * gg + 7]);
          *reinterpret_cast<uint4*>(dk_out + gg * 8) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == ABP_W_MMA_S) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace plm

#ifndef PLM_ATTN_BWD_DEFAULT_VARIANT
#define PLM_ATTN_BWD_DEFAULT_VARIANT 0
#endif

extern "C" int plm_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                            const int32_t* seg_start, const float* rope_table, void* dqkv, float* delta,
                            float* dq_acc, int32_t B, int32_t T, int32_t H, int32_t hd, plm_stream_t stream) {
  return plm_attn_bwd_variant(qkv, out, dout, lse, seg_start, rope_table, dqkv, delta, dq_acc, B, T, H, hd, -1, stream);
}

// variant: 0 = one CTA per (key tile, head, batch), v >= 1 = item-walking CTAs, v per SM (1 = persistent); < 0 = default.
extern "C" int plm_attn_bwd_variant(const void* qkv, const void* out, const void* dout, const float* lse,
                                    const int32_t* seg_start, const float* rope_table, void* dqkv, float* delta,
                                    float* dq_acc, int32_t B, int32_t T, int32_t H, int32_t hd, int32_t variant,
                                    plm_stream_t stream_) {
  using namespace plm;
  if (variant < 0) {
    // diagnostics: PLM_ATTN_BWD_VARIANT overrides the compiled default; read ONCE per process, never per launch
    static std::once_flag vonce;
    static int dflt = PLM_ATTN_BWD_DEFAULT_VARIANT;
    std::call_once(vonce, [] {
      const char* v = getenv("PLM_ATTN_BWD_VARIANT");
      if (v && *v) dflt = atoi(v);
    });
    variant = dflt;
  }
  PLM_REQUIRE(variant >= 0 && variant <= 16, "attn_bwd: variant %d out of range", variant);
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
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(attn_bwd_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
  });
  if (attr_err != cudaSuccess) return fail(PLM_ERR_CUDA, "attn_bwd smem attribute: %s", cudaGetErrorString(attr_err));

  const int d = H * hd;
  const int64_t rows = static_cast<int64_t>(B) * T;
  CUtensorMap tmQKV, tmDO;
  int rc = make_tmap_bf16_2d(&tmQKV, qkv, rows, 3ull * d, 3ull * d, AB_T, 64);
  if (rc != PLM_OK) return rc;
  rc = make_tmap_bf16_2d(&tmDO, dout, rows, d, d, AB_T, 64);
  if (rc != PLM_OK) return rc;
  CUtensorMap tmDQ;  // dq_acc fp32 [rows, d]: boxes of 32 rows x 32 columns (half of one drain warp's share) for the reduce-adds
  rc = make_tmap_f32_2d(&tmDQ, dq_acc, rows, d, d, 32, 32);
  if (rc != PLM_OK) return rc;

  {
    const int64_t n = rows * H * 8;
    attn_delta_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
        static_cast<const uint4*>(out), static_cast<const uint4*>(dout), delta, reinterpret_cast<float4*>(dq_acc), rows, T,
        H);
    rc = check_launch("attn_delta");
    if (rc != PLM_OK) return rc;
  }
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  dim3 grid((T + AB_T - 1) / AB_T, H, B);
  if (variant >= 1) {
    // variant v: v CTAs per SM over the kernel's lifetime (v = 1: fully persistent).  With v > 1 the CTAs retire in v
    // waves, which lets GEMMs queued on another stream (the weight-gradient side stream) take over SMs between waves
    // while each CTA still amortises its set-up over ~n_items / (v SMs) items.
    const long long n_items = static_cast<long long>(grid.x) * H * B;
    const long long want = static_cast<long long>(variant) * sm_count();
    const int ctas = static_cast<int>(n_items < want ? n_items : want);
    attn_bwd_persistent_kernel<<<ctas, ABP_THREADS, AB_SMEM, stream>>>(
        tmQKV, tmDO, tmDQ, lse, delta, seg_start, rope_table, static_cast<__nv_bfloat16*>(dqkv), dq_acc, T, H, B * H,
        scale, scale * 1.4426950408889634f);
  } else {
    attn_bwd_kernel<<<grid, AB_THREADS, AB_SMEM, stream>>>(tmQKV, tmDO, tmDQ, lse, delta, seg_start, rope_table,
                                                           static_cast<__nv_bfloat16*>(dqkv), dq_acc, T, H, scale,
                                                           scale * 1.4426950408889634f);
  }
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

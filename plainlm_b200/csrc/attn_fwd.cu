// Causal / document-masked flash-attention forward on tcgen05 (models/transformer.py:53-63), round-2 design.
//
// Head dim 64 makes this kernel MUFU-bound, not tensor-bound: a 128x128 score tile needs 16384 exp2 (1024 cycles of the
// SM's 16/clk MUFU pipe) against 512 cycles of tensor pipe.  The round-1 kernel (one 128-query tile per CTA, two CTAs
// per SM, one thread per query row) kept the MUFU pipe only 71 % busy — two softmax warps per scheduler cannot cover
// each other's serial TMEM-load / max / store / barrier latencies — and spent ~7 k cycles per CTA outside its tile loop.
// This kernel is organised around FOUR independent softmax streams per scheduler and no per-tile launch cost:
//
//   * persistent: one CTA per SM walks a static, snake-ordered (heaviest first) list of work items; an item is a PAIR of
//     adjacent 128-query tiles of one (batch, head), so K/V are fetched once for 256 queries (half the L2 traffic);
//   * split-KV inside the CTA: keys are consumed in 64-key subtiles; even subtiles go to stream `a`, odd ones to stream
//     `b` of each query tile.  A stream is 4 warps (one thread per query row, 64 score registers) with its OWN running
//     max, row sum and O accumulator, so streams never talk to each other inside the key loop; the two partial results
//     of a query tile are merged once, in the epilogue (flash-decoding style).  2 query tiles x 2 streams = 16 softmax
//     warps = 4 per scheduler;
//   * TMEM (512 columns): per stream 64 columns S (fp32 scores, overwritten in place by P as packed bf16 pairs — the A
//     operand of the P·V MMA is read straight from tensor memory) + 64 columns O;
//   * two MMA-issuer threads (one per query tile): S_next = Q K^T is issued right behind the P·V of the same stream —
//     tcgen05.mma executes in issue order, so S may overwrite the P it follows and "S ready" also means "previous P·V
//     done" (the lazy O rescale needs no extra barrier);
//   * one TMA producer thread: Q tiles of the next item and a 6-deep ring of (K, V) subtiles run ahead of the MMAs;
//   * optional FMA-pipe exp2 (template POLY: that many of every 4 score pairs use a degree-3 polynomial instead of
//     MUFU.EX2; max relative error 7.5e-5, far below the bf16 rounding of P).
// Document masking never touches a dense mask: a row attends keys in [seg_start[row], row]; subtiles entirely before
// the item's first document are skipped.
#include "common.cuh"
#include "ptx.cuh"

#include <cstdlib>
#include <mutex>

#ifndef PLM_ATTN_FWD_DEFAULT_VARIANT
#define PLM_ATTN_FWD_DEFAULT_VARIANT 11
#endif

namespace plm {

constexpr int AF_BQ = 128;  // queries per tile (TMEM lanes)
constexpr int AF_BK = 64;   // keys per subtile
constexpr int AF_HD = 64;   // head dim
constexpr int AF_STAGES = 6;
constexpr int AF_QTILE_BYTES = AF_BQ * AF_HD * 2;  // 16 KB
constexpr int AF_SUB_BYTES = AF_BK * AF_HD * 2;    // 8 KB
constexpr int AF_STAGE_BYTES = 2 * AF_SUB_BYTES;   // K subtile + V subtile
constexpr int AF_SM_WARPS = 16;                    // softmax warps: stream = warp / 4, TMEM lane quarter = warp % 4
constexpr int AF_W_ISSUE = AF_SM_WARPS;            // warps 16, 17: MMA issuers of query tile 0 / 1
constexpr int AF_W_TMA = AF_SM_WARPS + 2;          // warp 18: TMA producer, TMEM allocation (warp 19 pads the warpgroup)
constexpr int AF_THREADS = (AF_SM_WARPS + 4) * 32;
constexpr int AF_REGS_SOFTMAX = 104;  // setmaxnreg moves registers inside the CTA pool (640 x 96 at launch): 512 x 104 + 128 x 64 = 61440
constexpr int AF_REGS_AUX = 64;
constexpr int AF_OFF_RING = 2 * AF_QTILE_BYTES;
constexpr int AF_OFF_XCH = AF_OFF_RING + AF_STAGES * AF_STAGE_BYTES;
constexpr int AF_OFF_BARS = AF_OFF_XCH + 2 * 4 * AF_BQ * 8;  // (max, sum) exchange, double-buffered by item parity
constexpr int AF_NBARS = 2 + 2 + 2 * AF_STAGES + 4 + 4 + 4 + 2;
constexpr int AF_SMEM = AF_OFF_BARS + AF_NBARS * 8 + 16;

__device__ __forceinline__ float af_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float af_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float af_max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ void af_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// exp2 of two arguments (<= ~2^7, -inf allowed) on the FMA / ALU pipes: round to nearest integer with the 1.5 * 2^23
// trick, degree-3 minimax polynomial of 2^f on [-0.5, 0.5] (max relative error 7.5e-5), exponent added as an integer.
__device__ __forceinline__ float2 af_exp2_poly2(float2 x) {
  const float kMagic = 12582912.0f;
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
  const float2 magic = make_float2(kMagic, kMagic);
  const float2 t = __fadd2_rn(x, magic);
  const float2 xi = __fadd2_rn(t, make_float2(-kMagic, -kMagic));
  const float2 f = __fadd2_rn(x, make_float2(-xi.x, -xi.y));
  float2 p = __ffma2_rn(make_float2(0.0551716648f, 0.0551716648f), f, make_float2(0.2426111251f, 0.2426111251f));
  p = __ffma2_rn(p, f, make_float2(0.6932609677f, 0.6932609677f));
  p = __ffma2_rn(p, f, make_float2(0.9999280572f, 0.9999280572f));
  float2 r;
  r.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return r;
}

// One work item: query tiles qt0 = 2m (slot 0) and 2m + 1 (slot 1) of (b, h).  Every role derives it the same way.
struct AfItem {
  int b, h, qt0;
  int j_lo;      // first 64-key subtile any row of the item can see
  int n0, n1;    // subtiles consumed by slot 0 / slot 1 (0: slot inactive — odd number of query tiles)
  int n_ring;    // subtiles the producer loads (= max(n))
  int64_t row0;  // global row of the item's first query (b * T + qt0 * 128)
};

struct AfSched {
  int n_items, BH, H, T, nq, n_pairs;
};

// k-th item of this CTA in the snake order over the heaviest-first item list; false when the list is exhausted.
__device__ __forceinline__ bool af_item(const AfSched& sc, int k, const int32_t* __restrict__ seg_start, AfItem& it) {
  const int G = static_cast<int>(gridDim.x), c = static_cast<int>(blockIdx.x);
  const int idx = k * G + ((k & 1) ? (G - 1 - c) : c);
  if (idx >= sc.n_items) return false;
  const int m = sc.n_pairs - 1 - idx / sc.BH;
  const int bh = idx % sc.BH;
  it.b = bh / sc.H;
  it.h = bh - it.b * sc.H;
  it.qt0 = 2 * m;
  it.row0 = static_cast<int64_t>(it.b) * sc.T + static_cast<int64_t>(it.qt0) * AF_BQ;
  it.j_lo = seg_start ? (__ldg(seg_start + it.row0) / AF_BK) : 0;
  it.n0 = 2 * it.qt0 + 2 - it.j_lo;
  it.n1 = (it.qt0 + 1 < sc.nq) ? 2 * it.qt0 + 4 - it.j_lo : 0;
  it.n_ring = it.n1 > it.n0 ? it.n1 : it.n0;
  return true;
}
// Number of rounds of the snake schedule this CTA must walk (rows of the item list).
__device__ __forceinline__ int af_rounds(const AfSched& sc) {
  return (sc.n_items + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
}

template <int POLY>
__global__ void __launch_bounds__(AF_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const int32_t* __restrict__ seg_start,
                __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int B, int T, int H, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                   // [2][16 KB]
  uint8_t* sRing = smem + AF_OFF_RING;  // [AF_STAGES][K 8 KB | V 8 KB]
  float2* sXch = reinterpret_cast<float2*>(smem + AF_OFF_XCH);  // [2][4 streams][128 rows] (running max, row sum)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AF_OFF_BARS);
  uint64_t* q_full = bars;                         // [2]  Q tile of the slot has landed
  uint64_t* q_empty = bars + 2;                    // [2]  every S MMA of the slot's item has completed
  uint64_t* kv_full = bars + 4;                    // [AF_STAGES]
  uint64_t* kv_empty = bars + 4 + AF_STAGES;       // [AF_STAGES]  released by BOTH issuers
  uint64_t* s_full = bars + 4 + 2 * AF_STAGES;     // [4]  scores of the stream's next subtile are in tensor memory
  uint64_t* p_full = s_full + 4;                   // [4]  P of the stream's subtile is in tensor memory (4 warps)
  uint64_t* pv_last = s_full + 8;                  // [4]  the stream's last P·V of the item has completed
  uint64_t* o_free = s_full + 12;                  // [2]  the epilogue has read both O accumulators of the slot (8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + AF_NBARS);

  if ((smem_u32(smem) & 1023u) != 0) __trap();  // layout contract of the 128-byte swizzle: fail the launch loudly

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  AfSched sc;
  sc.H = H;
  sc.T = T;
  sc.BH = B * H;
  sc.nq = (T + AF_BQ - 1) / AF_BQ;
  sc.n_pairs = (sc.nq + 1) / 2;
  sc.n_items = sc.n_pairs * sc.BH;
  const int rounds = af_rounds(sc);
  const int d = H * AF_HD;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&o_free[i], 8);
    }
    for (int i = 0; i < AF_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_last[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == AF_W_TMA) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // stream s: scores / P at columns [64 s, 64 s + 64), O accumulator at columns [256 + 64 s, +64)

  if (warp >= AF_SM_WARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AF_REGS_AUX));
    if (warp == AF_W_TMA) {
      // ---------------------------------------------------------------- TMA producer
      if (lane == 0) {
        uint32_t ring = 0;
        uint32_t qcnt0 = 0, qcnt1 = 0;
        for (int k = 0; k < rounds; ++k) {
          AfItem it;
          if (!af_item(sc, k, seg_start, it)) continue;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if ((q ? it.n1 : it.n0) == 0) continue;
            mbar_wait(&q_empty[q], ((q ? qcnt1 : qcnt0) & 1) ^ 1);
            if (q) ++qcnt1; else ++qcnt0;
            mbar_arrive_expect_tx(&q_full[q], AF_QTILE_BYTES);
            const int qr = static_cast<int>(it.row0) + q * AF_BQ;
            tma_load_2d(sQ + q * AF_QTILE_BYTES, &tmQKV, &q_full[q], it.h * AF_HD, qr);
            tma_load_2d(sQ + q * AF_QTILE_BYTES + AF_SUB_BYTES, &tmQKV, &q_full[q], it.h * AF_HD, qr + 64);
          }
          const int kr0 = it.b * T + it.j_lo * AF_BK;
          for (int jj = 0; jj < it.n_ring; ++jj, ++ring) {
            const uint32_t st = ring % AF_STAGES, ph = (ring / AF_STAGES) & 1;
            mbar_wait(&kv_empty[st], ph ^ 1);
            mbar_arrive_expect_tx(&kv_full[st], AF_STAGE_BYTES);
            uint8_t* dst = sRing + st * AF_STAGE_BYTES;
            tma_load_2d(dst, &tmQKV, &kv_full[st], d + it.h * AF_HD, kr0 + jj * AF_BK);
            tma_load_2d(dst + AF_SUB_BYTES, &tmQKV, &kv_full[st], 2 * d + it.h * AF_HD, kr0 + jj * AF_BK);
          }
        }
      }
    } else if (warp == AF_W_ISSUE || warp == AF_W_ISSUE + 1) {
      // ---------------------------------------------------------------- MMA issuer of query tile q
      if (lane == 0) {
        const int q = warp - AF_W_ISSUE;
        constexpr uint32_t idesc_s = make_idesc_bf16(128, AF_BK, 0, 0);  // S = Q K^T: both K-major, N = 64 keys
        constexpr uint32_t idesc_o = make_idesc_bf16(128, AF_HD, 0, 1);  // O += P V: A from TMEM, V MN-major, N = 64
        const uint64_t q_desc = make_smem_desc_sw128(smem_u32(sQ + q * AF_QTILE_BYTES), 16, 1024);
        const uint64_t k_desc0 = make_smem_desc_sw128(smem_u32(sRing), 16, 1024);
        const uint64_t v_desc0 = make_smem_desc_sw128(smem_u32(sRing + AF_SUB_BYTES), AF_SUB_BYTES, 1024);
        const uint32_t tS0 = tmem_base + q * 128, tO0 = tmem_base + 256 + q * 128;
        uint32_t ring0 = 0;        // ring position of the item's first subtile (same sequence as the producer)
        uint32_t items = 0;        // active items of this slot so far (q_full / o_free parity)
        uint32_t cnt_a = 0, cnt_b = 0;  // subtiles of stream a / b so far (p_full parity)
        auto issue_s = [&](int half, uint32_t rpos) {
          const uint32_t st = rpos % AF_STAGES, ph = (rpos / AF_STAGES) & 1;
          mbar_wait(&kv_full[st], ph);
          tc_fence_after();
          const uint64_t k_desc = k_desc0 + st * (AF_STAGE_BYTES >> 4);
#pragma unroll
          for (int kk = 0; kk < AF_HD / 16; ++kk)
            umma_ss(tS0 + half * 64, q_desc + kk * 2, k_desc + kk * 2, idesc_s, kk > 0 ? 1u : 0u);
          umma_commit(&s_full[2 * q + half]);
        };
        for (int k = 0; k < rounds; ++k) {
          AfItem it;
          if (!af_item(sc, k, seg_start, it)) continue;
          const int n = q ? it.n1 : it.n0;
          if (n > 0) {
            mbar_wait(&q_full[q], items & 1);
            tc_fence_after();
            issue_s(0, ring0);
            if (n > 1) issue_s(1, ring0 + 1);
            if (n <= 2) umma_commit(&q_empty[q]);  // every S MMA of this item has been issued
            for (int jj = 0; jj < n; ++jj) {
              const int half = jj & 1;
              mbar_wait(&p_full[2 * q + half], (half ? cnt_b : cnt_a) & 1);
              if (half) ++cnt_b; else ++cnt_a;
              if (jj == 0) mbar_wait(&o_free[q], (items & 1) ^ 1);  // the previous item's epilogue has read O
              tc_fence_after();
              const uint32_t st = (ring0 + jj) % AF_STAGES;
              const uint64_t v_desc = v_desc0 + st * (AF_STAGE_BYTES >> 4);
              const uint32_t tP = tS0 + half * 64, tO = tO0 + half * 64;
#pragma unroll
              for (int kk = 0; kk < AF_BK / 16; ++kk)
                umma_ts(tO, tP + kk * 8, v_desc + kk * (2048 >> 4), idesc_o, (jj >= 2 || kk > 0) ? 1u : 0u);
              umma_commit(&kv_empty[st]);  // this slot is done with the stage (its S MMA ran earlier, in order)
              if (jj + 2 < n) {
                issue_s(half, ring0 + jj + 2);  // overwrites the P just consumed: tcgen05.mma executes in issue order
                if (jj + 2 == n - 1) umma_commit(&q_empty[q]);
              } else {
                umma_commit(&pv_last[2 * q + half]);
              }
            }
            ++items;
          }
          // subtiles only the other slot consumes (or all of them when this slot is inactive): release them once loaded
          for (int jj = n; jj < it.n_ring; ++jj) {
            const uint32_t rpos = ring0 + jj, st = rpos % AF_STAGES, ph = (rpos / AF_STAGES) & 1;
            mbar_wait(&kv_full[st], ph);
            mbar_arrive(&kv_empty[st]);
          }
          ring0 += it.n_ring;
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(AF_REGS_SOFTMAX));
    // ------------------------------------------------------------------ softmax warps
    const int s = warp >> 2;        // stream
    const int q = s >> 1;           // query tile (slot)
    const int half = s & 1;         // parity of the subtiles this stream consumes
    const int quarter = warp & 3;   // TMEM lane quarter
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + s * 64 + lane_off;
    const uint32_t tO = tmem_base + 256 + s * 64 + lane_off;
    uint32_t cnt = 0;    // subtiles consumed so far (s_full parity)
    uint32_t items = 0;  // active items so far (pv_last parity)
    const float2 sc2 = make_float2(scale_log2, scale_log2);

    for (int k = 0; k < rounds; ++k) {
      AfItem it;
      if (!af_item(sc, k, seg_start, it)) continue;
      const int n = q ? it.n1 : it.n0;
      if (n == 0) continue;
      const int qi = (it.qt0 + q) * AF_BQ + r;  // position within the sequence
      const bool row_ok = qi < T;               // ragged tail: T need not be a multiple of 128
      const int64_t grow = it.row0 + q * AF_BQ + r;
      const int seg_lo = (seg_start && row_ok) ? __ldg(seg_start + grow) : 0;
      float m_run = -INFINITY, l_run = 0.f;

      for (int jj = half; jj < n; jj += 2) {
        const int kbase = (it.j_lo + jj) * AF_BK;
        mbar_wait(&s_full[s], cnt & 1);
        ++cnt;
        tc_fence_after();
        uint32_t t[AF_BK];
        tmem_ld32(tS, *reinterpret_cast<uint32_t(*)[32]>(&t[0]));
        tmem_ld32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&t[32]));
        tmem_ld_wait();
        const bool masked = (kbase + AF_BK - 1 > qi) || (kbase < seg_lo);  // key kj allowed iff seg_lo <= kj <= qi
        if (masked) {
#pragma unroll
          for (int i = 0; i < AF_BK; ++i) {
            const int kj = kbase + i;
            if (kj > qi || kj < seg_lo) t[i] = 0xff800000u;  // -inf
          }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < AF_BK; i += 8) {  // four independent chains of 3-input maxima
          mx0 = af_max3(mx0, __uint_as_float(t[i]), __uint_as_float(t[i + 1]));
          mx1 = af_max3(mx1, __uint_as_float(t[i + 2]), __uint_as_float(t[i + 3]));
          mx2 = af_max3(mx2, __uint_as_float(t[i + 4]), __uint_as_float(t[i + 5]));
          mx3 = af_max3(mx3, __uint_as_float(t[i + 6]), __uint_as_float(t[i + 7]));
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
        // running max with lazy rescale: the reference only moves when the max grows by more than 2^8
        const bool grow_row = mx > m_run + 8.0f;
        const bool any_grow = __any_sync(0xffffffffu, grow_row);
        float alpha = 1.0f;
        if (any_grow) {
          const float m_new = fmaxf(m_run, mx);
          alpha = (m_new == -INFINITY) ? 1.0f : af_ex2(m_run - m_new);  // m_run = -inf -> 0
          m_run = m_new;
          l_run *= alpha;
        }
        const float m_use = (m_run == -INFINITY) ? 0.f : m_run;
        const float2 nm2 = make_float2(-m_use, -m_use);
        float2 ps0 = make_float2(0.f, 0.f), ps1 = make_float2(0.f, 0.f);
        // exp2 pass, in place: P (packed bf16 pairs) overwrites the first half of the score registers
#pragma unroll
        for (int c8 = 0; c8 < AF_BK / 8; ++c8) {
          float2 e[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 a = __ffma2_rn(
                make_float2(__uint_as_float(t[c8 * 8 + 2 * i]), __uint_as_float(t[c8 * 8 + 2 * i + 1])), sc2, nm2);
            if (i >= 4 - POLY)
              e[i] = af_exp2_poly2(a);
            else
              e[i] = make_float2(af_ex2(a.x), af_ex2(a.y));
          }
          ps0 = __fadd2_rn(ps0, __fadd2_rn(e[0], e[1]));
          ps1 = __fadd2_rn(ps1, __fadd2_rn(e[2], e[3]));
#pragma unroll
          for (int i = 0; i < 4; ++i) t[c8 * 4 + i] = pack_bf16x2(e[i].x, e[i].y);  // slots < 8 * c8: already consumed
        }
        // O of this stream is quiescent here: S of this subtile was issued BEHIND the previous P·V of the stream
        if (any_grow && jj >= 2) {
#pragma unroll
          for (int c = 0; c < AF_HD / 16; ++c) {
            uint32_t o[16];
            tmem_ld16(tO + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tO + c * 16, o);
          }
        }
        tmem_st32(tS, *reinterpret_cast<const uint32_t(*)[32]>(&t[0]));  // P: 64 keys = 32 columns of bf16 pairs
        tmem_st_wait();
        l_run += (ps0.x + ps0.y) + (ps1.x + ps1.y);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[s]);
      }

      // ---- epilogue: merge the two streams of the query tile, O / l -> bf16 out[b, t, h, :], lse
      float2* xch = sXch + (items & 1) * (4 * AF_BQ);  // a slot is rewritten two items later: behind o_free
      xch[s * AF_BQ + r] = make_float2(m_run, l_run);
      af_bar_sync(1 + q, 256);
      const float2 other = xch[(s ^ 1) * AF_BQ + r];
      const float m_a = half ? other.x : m_run, l_a = half ? other.y : l_run;
      const float m_b = half ? m_run : other.x, l_b = half ? l_run : other.y;
      const float m_all = fmaxf(m_a, m_b);
      const float m_use = (m_all == -INFINITY) ? 0.f : m_all;
      const float w_a = af_ex2(m_a - m_use), w_b = af_ex2(m_b - m_use);  // -inf -> 0
      const float l_all = l_a * w_a + l_b * w_b;
      const float inv_l = l_all > 0.f ? 1.0f / l_all : 0.f;
      const float f_a = w_a * inv_l, f_b = w_b * inv_l;
      mbar_wait(&pv_last[2 * q], items & 1);
      mbar_wait(&pv_last[2 * q + 1], items & 1);
      ++items;
      tc_fence_after();
      // this thread finishes head-dim columns [32 half, 32 half + 32) of its row from BOTH accumulators
      uint32_t oa[32], ob[32];
      tmem_ld32(tmem_base + 256 + (2 * q) * 64 + lane_off + half * 32, oa);
      tmem_ld32(tmem_base + 256 + (2 * q + 1) * 64 + lane_off + half * 32, ob);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[q]);
      if (row_ok) {
        __nv_bfloat16* orow = out + grow * d + it.h * AF_HD + half * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 v;
          uint32_t* vv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = 8 * i + 2 * e;
            vv[e] = pack_bf16x2(__uint_as_float(oa[c]) * f_a + __uint_as_float(ob[c]) * f_b,
                                __uint_as_float(oa[c + 1]) * f_a + __uint_as_float(ob[c + 1]) * f_b);
          }
          *reinterpret_cast<uint4*>(orow + i * 8) = v;
        }
        if (half == 0)
          lse[(static_cast<int64_t>(it.b) * H + it.h) * T + qi] = (m_use + af_lg2(l_all)) * 0.6931471805599453f;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == AF_W_TMA) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int POLY>
static int launch_attn_fwd(const CUtensorMap& tm, const int32_t* seg_start, void* out, float* lse, int B, int T, int H,
                           float scale_log2, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(attn_fwd_kernel<POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM);
  });
  if (attr_err != cudaSuccess) return fail(PLM_ERR_CUDA, "attn_fwd smem attribute: %s", cudaGetErrorString(attr_err));
  const int nq = (T + AF_BQ - 1) / AF_BQ;
  const long long n_items = static_cast<long long>((nq + 1) / 2) * B * H;
  const int grid = static_cast<int>(n_items < sm_count() ? n_items : sm_count());
  attn_fwd_kernel<POLY><<<grid, AF_THREADS, AF_SMEM, stream>>>(tm, seg_start, static_cast<__nv_bfloat16*>(out), lse, B,
                                                               T, H, scale_log2);
  return check_launch("attn_fwd");
}


// =====================================================================================================================
// Three DECOUPLED streams (variants 10..12).  The four-stream kernel above single-buffers S (P overwrites it), so every
// subtile of a stream walks the whole chain  P store -> p_full -> P·V issue -> S issue -> MMA -> s_full -> TMEM load
// (measured: ~3500 cycles per round of four subtiles against a 2048-cycle MUFU floor).  Here a stream owns
// 160 TMEM columns — S (64) | P (32) | O (64) — so the next S = Q K^T is issued as soon as the softmax warps have READ
// the current one, and runs under their max / exp2 / P-store work: the softmax warps never wait for the tensor pipe.
// 3 x 160 = 480 columns allow three streams: an item is a group of up to THREE adjacent 128-query tiles of one
// (batch, head) sharing one ring of (K, V) subtiles; stream s = query tile t0 + s (no split-KV, no merge).
// 16 warps: 0..11 softmax (stream = warp / 4), 12..14 one MMA issuer per stream, 15 TMA producer.  512 threads leave
// 128 registers per thread: no setmaxnreg.
constexpr int A3_STREAMS = 3;
constexpr int A3_STAGES = 10;  // deep enough for a stream to run 4+ subtiles ahead of the slowest one (see rotation)
constexpr int A3_W_ISSUE = 4 * A3_STREAMS;   // warps 12, 13, 14
constexpr int A3_W_TMA = A3_W_ISSUE + A3_STREAMS;  // warp 15
constexpr int A3_THREADS = (A3_W_TMA + 1) * 32;
constexpr int A3_OFF_RING = A3_STREAMS * AF_QTILE_BYTES;
constexpr int A3_OFF_BARS = A3_OFF_RING + A3_STAGES * AF_STAGE_BYTES;
constexpr int A3_NBARS = 2 * A3_STAGES + 7 * A3_STREAMS;
constexpr int A3_SMEM = A3_OFF_BARS + A3_NBARS * 8 + 16;
constexpr int A3_TCOLS = 160;  // TMEM columns per stream: S at +0, P at +64, O at +96

struct A3Item {
  int b, h, t0;   // first query tile of the group
  int j_lo;       // first 64-key subtile any row of the item can see
  int cnt;        // active streams (query tiles) of the item: 1..3
  int rot;        // stream s works on tile t0 + (s + rot) % cnt: the longest tile of a group rotates over the streams
  int n_ring;     // subtiles the producer loads = subtiles of the item's last tile
  int64_t row0;   // global row of the item's first query
};
struct A3Sched {
  int n_items, BH, H, T, nq;
};
// position (tile index within the group) stream s works on, -1: inactive.  A group's tiles need n, n + 2, n + 4
// subtiles; with a fixed assignment stream 2 would do 4 subtiles more than stream 0 in EVERY item and the others would
// idle 12 % of the time.  Rotating the assignment with the item counter equalises the streams' totals, and the deep
// (K, V) ring lets a stream that finished its tile run ahead into the next item instead of waiting.
__device__ __forceinline__ int a3_pos(const A3Item& it, int s) {
  if (s >= it.cnt) return -1;
  const int p = s + it.rot;
  return p >= it.cnt ? p - it.cnt : p;
}
// subtiles the tile at position `pos` consumes (0: inactive)
__device__ __forceinline__ int a3_n(const A3Item& it, int pos) { return pos >= 0 ? 2 * (it.t0 + pos) + 2 - it.j_lo : 0; }

// k-th item of this CTA in the snake order over the heaviest-first item list (groups are cut from the END of the
// sequence, so only the lightest group of a (batch, head) can be short); false when the list is exhausted.
__device__ __forceinline__ bool a3_item(const A3Sched& sc, int k, const int32_t* __restrict__ seg_start, A3Item& it) {
  const int G = static_cast<int>(gridDim.x), c = static_cast<int>(blockIdx.x);
  const int idx = k * G + ((k & 1) ? (G - 1 - c) : c);
  if (idx >= sc.n_items) return false;
  const int g = idx / sc.BH;
  const int bh = idx - g * sc.BH;
  it.b = bh / sc.H;
  it.h = bh - it.b * sc.H;
  const int lo = sc.nq - A3_STREAMS * (g + 1);
  it.t0 = lo < 0 ? 0 : lo;
  it.cnt = lo < 0 ? A3_STREAMS + lo : A3_STREAMS;
  it.rot = it.cnt == A3_STREAMS ? k % A3_STREAMS : 0;
  it.row0 = static_cast<int64_t>(it.b) * sc.T + static_cast<int64_t>(it.t0) * AF_BQ;
  it.j_lo = seg_start ? (__ldg(seg_start + it.row0) / AF_BK) : 0;
  it.n_ring = 2 * (it.t0 + it.cnt - 1) + 2 - it.j_lo;
  return true;
}

template <int POLY>
__global__ void __launch_bounds__(A3_THREADS, 1)
attn_fwd3_kernel(const __grid_constant__ CUtensorMap tmQKV, const int32_t* __restrict__ seg_start,
                 __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int B, int T, int H, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                   // [3][16 KB]
  uint8_t* sRing = smem + A3_OFF_RING;  // [A3_STAGES][K 8 KB | V 8 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A3_OFF_BARS);
  uint64_t* kv_full = bars;                        // [A3_STAGES]
  uint64_t* kv_empty = bars + A3_STAGES;           // [A3_STAGES]  released by all three issuers
  uint64_t* q_full = bars + 2 * A3_STAGES;         // [3]  per stream from here on
  uint64_t* q_empty = q_full + A3_STREAMS;         //      every S MMA of the stream's item has completed
  uint64_t* s_full = q_full + 2 * A3_STREAMS;      //      scores of the next subtile are in tensor memory
  uint64_t* s_empty = q_full + 3 * A3_STREAMS;     //      ... and have been read out of it (4 warps)
  uint64_t* p_full = q_full + 4 * A3_STREAMS;      //      P of the subtile is in tensor memory (4 warps)
  uint64_t* pv_done = q_full + 5 * A3_STREAMS;     //      the subtile's P·V has completed: P may be overwritten, O read
  uint64_t* o_free = q_full + 6 * A3_STREAMS;      //      the epilogue has read O (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + A3_NBARS);

  if ((smem_u32(smem) & 1023u) != 0) __trap();

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  A3Sched sc;
  sc.H = H;
  sc.T = T;
  sc.BH = B * H;
  sc.nq = (T + AF_BQ - 1) / AF_BQ;
  sc.n_items = ((sc.nq + A3_STREAMS - 1) / A3_STREAMS) * sc.BH;
  const int rounds = (sc.n_items + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int d = H * AF_HD;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int i = 0; i < A3_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], A3_STREAMS);
    }
    for (int i = 0; i < A3_STREAMS; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_free[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == A3_W_TMA) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == A3_W_TMA) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t ring = 0;  // Q tiles are fetched by each stream's own issuer: this thread never waits for a stream
      for (int k = 0; k < rounds; ++k) {
        A3Item it;
        if (!a3_item(sc, k, seg_start, it)) continue;
        const int kr0 = it.b * T + it.j_lo * AF_BK;
        for (int jj = 0; jj < it.n_ring; ++jj, ++ring) {
          const uint32_t st = ring % A3_STAGES, ph = (ring / A3_STAGES) & 1;
          mbar_wait(&kv_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&kv_full[st], AF_STAGE_BYTES);
          uint8_t* dst = sRing + st * AF_STAGE_BYTES;
          tma_load_2d(dst, &tmQKV, &kv_full[st], d + it.h * AF_HD, kr0 + jj * AF_BK);
          tma_load_2d(dst + AF_SUB_BYTES, &tmQKV, &kv_full[st], 2 * d + it.h * AF_HD, kr0 + jj * AF_BK);
        }
      }
    }
  } else if (warp >= A3_W_ISSUE) {
    // -------------------------------------------------------------------- MMA issuer of stream s
    if (lane == 0) {
      const int s = warp - A3_W_ISSUE;
      constexpr uint32_t idesc_s = make_idesc_bf16(128, AF_BK, 0, 0);  // S = Q K^T: both K-major, N = 64 keys
      constexpr uint32_t idesc_o = make_idesc_bf16(128, AF_HD, 0, 1);  // O += P V: A from TMEM, V MN-major, N = 64
      const uint64_t q_desc = make_smem_desc_sw128(smem_u32(sQ + s * AF_QTILE_BYTES), 16, 1024);
      const uint64_t k_desc0 = make_smem_desc_sw128(smem_u32(sRing), 16, 1024);
      const uint64_t v_desc0 = make_smem_desc_sw128(smem_u32(sRing + AF_SUB_BYTES), AF_SUB_BYTES, 1024);
      const uint32_t tS = tmem_base + s * A3_TCOLS, tP = tS + 64, tO = tS + 96;
      uint32_t ring0 = 0;  // ring position of the item's first subtile (same sequence as the producer)
      uint32_t items = 0;  // active items of this stream so far (q_full / q_empty / o_free parity)
      uint32_t cs = 0;     // S tiles issued so far (s_empty parity)
      uint32_t cp = 0;     // P·V issued so far (p_full parity)
      // requests the Q tile of this stream's next active item after item k (the buffer must be free: q_empty)
      auto request_next_q = [&](int k_from) {
        for (int kn = k_from; kn < rounds; ++kn) {
          A3Item nx;
          if (!a3_item(sc, kn, seg_start, nx)) continue;
          const int pos = a3_pos(nx, s);
          if (pos < 0) continue;
          mbar_arrive_expect_tx(&q_full[s], AF_QTILE_BYTES);
          const int qr = static_cast<int>(nx.row0) + pos * AF_BQ;
          tma_load_2d(sQ + s * AF_QTILE_BYTES, &tmQKV, &q_full[s], nx.h * AF_HD, qr);
          tma_load_2d(sQ + s * AF_QTILE_BYTES + AF_SUB_BYTES, &tmQKV, &q_full[s], nx.h * AF_HD, qr + 64);
          return;
        }
      };
      request_next_q(0);
      auto issue_s = [&](uint32_t rpos) {
        const uint32_t st = rpos % A3_STAGES, ph = (rpos / A3_STAGES) & 1;
        mbar_wait(&kv_full[st], ph);
        mbar_wait(&s_empty[s], (cs & 1) ^ 1);  // the softmax warps have read the previous scores out of tensor memory
        ++cs;
        tc_fence_after();
        const uint64_t k_desc = k_desc0 + st * (AF_STAGE_BYTES >> 4);
#pragma unroll
        for (int kk = 0; kk < AF_HD / 16; ++kk) umma_ss(tS, q_desc + kk * 2, k_desc + kk * 2, idesc_s, kk > 0 ? 1u : 0u);
        umma_commit(&s_full[s]);
      };
      for (int k = 0; k < rounds; ++k) {
        A3Item it;
        if (!a3_item(sc, k, seg_start, it)) continue;
        const int n = a3_n(it, a3_pos(it, s));
        if (n > 0) {
          mbar_wait(&q_full[s], items & 1);
          tc_fence_after();
          issue_s(ring0);
          for (int jj = 0; jj < n; ++jj) {
            // S of the NEXT subtile goes out as soon as the current one has been read, ahead of this subtile's P·V
            if (jj + 1 < n) issue_s(ring0 + jj + 1);
            if (jj + 1 == n - 1) umma_commit(&q_empty[s]);  // every S MMA of this item has been issued (n >= 2)
            mbar_wait(&p_full[s], cp & 1);
            ++cp;
            if (jj == 0) mbar_wait(&o_free[s], (items & 1) ^ 1);  // the previous item's epilogue has read O
            tc_fence_after();
            const uint32_t st = (ring0 + jj) % A3_STAGES;
            const uint64_t v_desc = v_desc0 + st * (AF_STAGE_BYTES >> 4);
#pragma unroll
            for (int kk = 0; kk < AF_BK / 16; ++kk)
              umma_ts(tO, tP + kk * 8, v_desc + kk * (2048 >> 4), idesc_o, (jj > 0 || kk > 0) ? 1u : 0u);
            umma_commit(&kv_empty[st]);  // this stream is done with the stage (its S MMA ran earlier, in order)
            umma_commit(&pv_done[s]);
            if (jj == n - 2) {
              // the item's last S was issued a whole subtile ago: its Q buffer is free (or about to be) — fetch the
              // next item's Q tile now, under this item's last subtile and epilogue
              mbar_wait(&q_empty[s], items & 1);
              request_next_q(k + 1);
            }
          }
          ++items;
        }
        // subtiles only later tiles of the group consume (or all of them when this stream is inactive)
        for (int jj = n; jj < it.n_ring; ++jj) {
          const uint32_t rpos = ring0 + jj, st = rpos % A3_STAGES, ph = (rpos / A3_STAGES) & 1;
          mbar_wait(&kv_full[st], ph);
          mbar_arrive(&kv_empty[st]);
        }
        ring0 += it.n_ring;
      }
    }
  } else {
    // ---------------------------------------------------------------------- softmax warps
    const int s = warp >> 2;       // stream = query tile within the group
    const int quarter = warp & 3;  // TMEM lane quarter
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + s * A3_TCOLS + lane_off, tP = tS + 64, tO = tS + 96;
    uint32_t c = 0;  // subtiles consumed so far (s_full / pv_done parity)
    const float2 sc2 = make_float2(scale_log2, scale_log2);

    for (int k = 0; k < rounds; ++k) {
      A3Item it;
      if (!a3_item(sc, k, seg_start, it)) continue;
      const int pos = a3_pos(it, s);
      const int n = a3_n(it, pos);
      if (n == 0) continue;
      const int qi = (it.t0 + pos) * AF_BQ + r;  // position within the sequence
      const bool row_ok = qi < T;                // ragged tail: T need not be a multiple of 128
      const int64_t grow = it.row0 + pos * AF_BQ + r;
      const int seg_lo = (seg_start && row_ok) ? __ldg(seg_start + grow) : 0;
      float m_run = -INFINITY, l_run = 0.f;

      for (int jj = 0; jj < n; ++jj, ++c) {
        const int kbase = (it.j_lo + jj) * AF_BK;
        mbar_wait(&s_full[s], c & 1);
        tc_fence_after();
        uint32_t t[AF_BK];
        tmem_ld32(tS, *reinterpret_cast<uint32_t(*)[32]>(&t[0]));
        tmem_ld32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&t[32]));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[s]);  // S may be overwritten by the next Q K^T
        const bool masked = (kbase + AF_BK - 1 > qi) || (kbase < seg_lo);  // key kj allowed iff seg_lo <= kj <= qi
        if (masked) {
#pragma unroll
          for (int i = 0; i < AF_BK; ++i) {
            const int kj = kbase + i;
            if (kj > qi || kj < seg_lo) t[i] = 0xff800000u;  // -inf
          }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < AF_BK; i += 8) {
          mx0 = af_max3(mx0, __uint_as_float(t[i]), __uint_as_float(t[i + 1]));
          mx1 = af_max3(mx1, __uint_as_float(t[i + 2]), __uint_as_float(t[i + 3]));
          mx2 = af_max3(mx2, __uint_as_float(t[i + 4]), __uint_as_float(t[i + 5]));
          mx3 = af_max3(mx3, __uint_as_float(t[i + 6]), __uint_as_float(t[i + 7]));
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
        const bool grow_row = mx > m_run + 8.0f;  // lazy rescale: the reference only moves for growth beyond 2^8
        const bool any_grow = __any_sync(0xffffffffu, grow_row);
        float alpha = 1.0f;
        if (any_grow) {
          const float m_new = fmaxf(m_run, mx);
          alpha = (m_new == -INFINITY) ? 1.0f : af_ex2(m_run - m_new);
          m_run = m_new;
          l_run *= alpha;
        }
        const float neg_m = (m_run == -INFINITY) ? 0.f : -m_run;
        // (A per-scheduler FIFO lock that made the exp2 passes of the three warps sharing a MUFU pipe exclusive was
        // measured and removed: 0.119 -> 0.137 ms.  The warps interleave on the pipe better than a lock hand-off, whose
        // ~200-cycle wake-up leaves the pipe idle three times per round.)
        const float2 nm2 = make_float2(neg_m, neg_m);
        float2 ps0 = make_float2(0.f, 0.f), ps1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int c8 = 0; c8 < AF_BK / 8; ++c8) {
          float2 e[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 a = __ffma2_rn(
                make_float2(__uint_as_float(t[c8 * 8 + 2 * i]), __uint_as_float(t[c8 * 8 + 2 * i + 1])), sc2, nm2);
            if (i >= 4 - POLY)
              e[i] = af_exp2_poly2(a);
            else
              e[i] = make_float2(af_ex2(a.x), af_ex2(a.y));
          }
          ps0 = __fadd2_rn(ps0, __fadd2_rn(e[0], e[1]));
          ps1 = __fadd2_rn(ps1, __fadd2_rn(e[2], e[3]));
#pragma unroll
          for (int i = 0; i < 4; ++i) t[c8 * 4 + i] = pack_bf16x2(e[i].x, e[i].y);
        }
        const float psum = (ps0.x + ps0.y) + (ps1.x + ps1.y);
        // the previous P·V must be complete before its P is overwritten or O rescaled
        mbar_wait(&pv_done[s], (c & 1) ^ 1);
        tc_fence_after();
        if (any_grow && jj > 0) {
#pragma unroll
          for (int cc = 0; cc < AF_HD / 16; ++cc) {
            uint32_t o[16];
            tmem_ld16(tO + cc * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tO + cc * 16, o);
          }
        }
        tmem_st32(tP, *reinterpret_cast<const uint32_t(*)[32]>(&t[0]));
        tmem_st_wait();
        l_run += psum;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[s]);
      }

      // ---- epilogue: O / l -> bf16 out[b, t, h, :], lse
      mbar_wait(&pv_done[s], (c & 1) ^ 1);  // the item's last P·V
      tc_fence_after();
      const float inv_l = l_run > 0.f ? 1.0f / l_run : 0.f;
      __nv_bfloat16* orow = out + grow * d + it.h * AF_HD;
#pragma unroll
      for (int cc = 0; cc < AF_HD / 32; ++cc) {
        uint32_t o[32];
        tmem_ld32(tO + cc * 32, o);
        tmem_ld_wait();
        if (cc == AF_HD / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_free[s]);
        }
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l);
            v.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l);
            v.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l);
            v.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + cc * 32 + i * 8) = v;
          }
        }
      }
      const float m_fin = (m_run == -INFINITY) ? 0.f : m_run;
      if (row_ok) lse[(static_cast<int64_t>(it.b) * H + it.h) * T + qi] = (m_fin + af_lg2(l_run)) * 0.6931471805599453f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == A3_W_TMA) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int POLY>
static int launch_attn_fwd3(const CUtensorMap& tm, const int32_t* seg_start, void* out, float* lse, int B, int T, int H,
                            float scale_log2, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(attn_fwd3_kernel<POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM);
  });
  if (attr_err != cudaSuccess) return fail(PLM_ERR_CUDA, "attn_fwd smem attribute: %s", cudaGetErrorString(attr_err));
  const int nq = (T + AF_BQ - 1) / AF_BQ;
  const long long n_items = static_cast<long long>((nq + A3_STREAMS - 1) / A3_STREAMS) * B * H;
  const int grid = static_cast<int>(n_items < sm_count() ? n_items : sm_count());
  attn_fwd3_kernel<POLY><<<grid, A3_THREADS, A3_SMEM, stream>>>(tm, seg_start, static_cast<__nv_bfloat16*>(out), lse, B,
                                                                T, H, scale_log2);
  return check_launch("attn_fwd");
}

}  // namespace plm

// variant: 0..2 = four coupled streams, 10..12 = three decoupled streams; the last digit is the number of score pairs
// out of every 4 whose exp2 runs on the FMA pipe; < 0 = library default.
extern "C" int plm_attn_fwd_variant(const void* qkv, const int32_t* seg_start, void* out, float* lse, int32_t B,
                                    int32_t T, int32_t H, int32_t hd, int32_t variant, plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(qkv);
  PLM_REQUIRE(qkv && out && lse, "attn_fwd: null pointer");
  PLM_REQUIRE(B > 0 && T > 0 && H > 0, "attn_fwd: bad size");
  if (hd != AF_HD) return fail(PLM_ERR_UNSUPPORTED, "attn_fwd: head_dim %d unsupported (need 64)", hd);
  PLM_REQUIRE(aligned16(qkv) && aligned16(out), "attn_fwd: misaligned pointer");
  PLM_REQUIRE(static_cast<int64_t>(B) * T < (1ll << 31) - 4096 && static_cast<int64_t>(B) * H < (1 << 24),
              "attn_fwd: size too large");
  const int d = H * hd;
  CUtensorMap tm;  // boxes of 64 rows x 64 columns: a K or V subtile, half a Q tile
  int rc = make_tmap_bf16_2d(&tm, qkv, static_cast<uint64_t>(B) * T, 3ull * d, 3ull * d, AF_BK, 64);
  if (rc != PLM_OK) return rc;
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(hd));
  if (variant < 0) {
    // diagnostics: PLM_ATTN_FWD_VARIANT overrides the compiled default; read ONCE per process, never per launch
    static std::once_flag vonce;
    static int dflt = PLM_ATTN_FWD_DEFAULT_VARIANT;
    std::call_once(vonce, [] {
      const char* v = getenv("PLM_ATTN_FWD_VARIANT");
      if (v && *v) dflt = atoi(v);
    });
    variant = dflt;
  }
  switch (variant) {
    case 0: return launch_attn_fwd<0>(tm, seg_start, out, lse, B, T, H, scale_log2, stream);
    case 1: return launch_attn_fwd<1>(tm, seg_start, out, lse, B, T, H, scale_log2, stream);
    case 2: return launch_attn_fwd<2>(tm, seg_start, out, lse, B, T, H, scale_log2, stream);
    case 10: return launch_attn_fwd3<0>(tm, seg_start, out, lse, B, T, H, scale_log2, stream);
    case 11: return launch_attn_fwd3<1>(tm, seg_start, out, lse, B, T, H, scale_log2, stream);
    case 12: return launch_attn_fwd3<2>(tm, seg_start, out, lse, B, T, H, scale_log2, stream);
    default: return fail(PLM_ERR_INVALID, "attn_fwd: variant %d out of range", variant);
  }
}

extern "C" int plm_attn_fwd(const void* qkv, const int32_t* seg_start, void* out, float* lse, int32_t B, int32_t T,
                            int32_t H, int32_t hd, plm_stream_t stream) {
  return plm_attn_fwd_variant(qkv, seg_start, out, lse, B, T, H, hd, -1, stream);
}

// ROUND-1 forward kernel, kept only as the A/B baseline of tools/gpu_kernel_check.py (entry point plm_attn_fwd_v1);
// the product path is attn_fwd.cu.
// Causal / document-masked flash-attention forward on tcgen05 (models/transformer.py:53-63).
//
// One CTA per (128-query tile, head, batch); 160 threads:
//   warps 0..3  softmax: thread r owns query row r (TMEM lane r) — row max / sum need no shuffles
//   warp 4      control: one thread issues the TMA loads (Q once, K/V double-buffered) and all tcgen05.mma
// Per 128-key tile:  S = Q K^T  (TMEM cols 0..127)  ->  softmax in registers  ->  P (bf16 pairs) to TMEM cols 192..255
//                    ->  O += P V  (TMEM cols 128..191; A = P read from tensor memory, V consumed MN-major straight
//                        from its TMA box).
// O stays in TMEM for the whole row of tiles; it is rescaled only when the running max grows by more than 2^8
// (lazy rescale), so the common path never round-trips O through registers.  Two CTAs are resident per SM
// (80 KB smem, 256 TMEM columns each): one CTA's softmax overlaps the other's MMAs.
// Document masking never touches a dense mask: a row attends keys in [seg_start[row], row]; key tiles entirely
// before the tile's first document are skipped.
#include "common.cuh"
#include "ptx.cuh"

#include <cstdlib>
#include <mutex>

namespace plm {

constexpr int ATT_BQ = 128;   // queries per CTA
constexpr int ATT_BK = 128;   // keys per tile
constexpr int ATT_HD = 64;    // head dim
constexpr int ATT_THREADS = 160;
constexpr int ATT_TILE_BYTES = ATT_BK * ATT_HD * 2;  // 16 KB
constexpr int ATT_FWD_SMEM = ATT_TILE_BYTES * (1 + 2 + 2) + 128;  // Q, K x2, V x2 + barriers (P lives in tensor memory)

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float max3f(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_v1_kernel(const __grid_constant__ CUtensorMap tmQKV, const int32_t* __restrict__ seg_start,
                __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int T, int H, float scale_log2,
                unsigned long long* __restrict__ trace) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT_TILE_BYTES;
  uint8_t* sV = smem + 3 * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * ATT_TILE_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* pv_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  if ((smem_u32(smem) & 1023u) != 0) __trap();  // layout contract violated: fail the launch loudly

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // diagnostics (PLM_ATTN_FWD_TRACE): one CTA stamps clock64() at its phase boundaries for four steady-state tiles
  const bool tr_on = trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == gridDim.z / 2;
#define AF_TR(slot_)                                                        \
  do {                                                                      \
    if (tr_on && it >= 6 && it < 10 && lane == 0) trace[(slot_)] = clock64(); \
  } while (0)
  if (tr_on && threadIdx.x == 0) {
    trace[120] = clock64();
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    trace[122] = gt;
  }
  const int qt = gridDim.x - 1 - blockIdx.x;  // heavy (late) tiles first
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int d = H * ATT_HD;
  const int64_t row0 = static_cast<int64_t>(b) * T + qt * ATT_BQ;

  int j_lo = 0;
  if (seg_start) j_lo = seg_start[row0] / ATT_BK;
  const int j_hi = qt;
  const int n_it = j_hi - j_lo + 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    mbar_init(&kv_full[0], 1);
    mbar_init(&kv_full[1], 1);
    mbar_init(&kv_empty[0], 1);
    mbar_init(&kv_empty[1], 1);
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc<256>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;
  const uint32_t tO = tmem_base + 128;
  const uint32_t tP = tmem_base + 192;  // P as packed bf16 pairs: lane = query row, 64 columns = 128 keys

  if (warp == 4) {
    if (lane == 0) {
      // ------------------------------------------------------------ control thread: TMA + MMA issue
      constexpr uint32_t idesc_s = make_idesc_bf16(128, ATT_BK, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, ATT_HD, 0, 1);
      mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_2d(sQ, &tmQKV, q_full, h * ATT_HD, static_cast<int>(row0));
      {
        const int kr = static_cast<int>(static_cast<int64_t>(b) * T + j_lo * ATT_BK);
        mbar_arrive_expect_tx(&kv_full[0], 2 * ATT_TILE_BYTES);
        tma_load_2d(sK, &tmQKV, &kv_full[0], d + h * ATT_HD, kr);
        tma_load_2d(sV, &tmQKV, &kv_full[0], 2 * d + h * ATT_HD, kr);
      }
      // every later K/V tile of this CTA goes to L2 now: the smem ring is only two deep
      for (int itp = 2; itp < n_it; ++itp) {
        const int kr = static_cast<int>(static_cast<int64_t>(b) * T + (j_lo + itp) * ATT_BK);
        tma_prefetch_l2_2d(&tmQKV, d + h * ATT_HD, kr);
        tma_prefetch_l2_2d(&tmQKV, 2 * d + h * ATT_HD, kr);
      }
      mbar_wait(q_full, 0);
      // descriptors are built once; per K-step only the 14-bit start-address field advances (tight issue loop)
      const uint64_t q_desc = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t k_desc0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t v_desc0 = make_smem_desc_sw128(smem_u32(sV), ATT_TILE_BYTES, 1024);  // MN-major view
      auto issue_s = [&](int st) {
        const uint64_t k_desc = k_desc0 + st * (ATT_TILE_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < ATT_HD / 16; ++k) umma_ss(tS, q_desc + k * 2, k_desc + k * 2, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
      };
      auto load_kv = [&](int it_next) {
        const int nst = it_next & 1;
        const int kr = static_cast<int>(static_cast<int64_t>(b) * T + (j_lo + it_next) * ATT_BK);
        mbar_arrive_expect_tx(&kv_full[nst], 2 * ATT_TILE_BYTES);
        tma_load_2d(sK + nst * ATT_TILE_BYTES, &tmQKV, &kv_full[nst], d + h * ATT_HD, kr);
        tma_load_2d(sV + nst * ATT_TILE_BYTES, &tmQKV, &kv_full[nst], 2 * d + h * ATT_HD, kr);
      };
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_s(0);
      if (n_it > 1) load_kv(1);
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        // S of the NEXT key tile goes out as soon as the softmax warps have read the current one out of tensor memory,
        // ahead of this tile's P·V: the next softmax never waits for the tensor pipe.
        if (it + 1 < n_it) {
          mbar_wait(&kv_full[st ^ 1], ((it + 1) >> 1) & 1);
          mbar_wait(s_empty, it & 1);
          AF_TR(64 + (it - 6) * 4 + 0);
          tc_fence_after();
          issue_s(st ^ 1);
          AF_TR(64 + (it - 6) * 4 + 1);
        }
        mbar_wait(p_full, it & 1);
        AF_TR(64 + (it - 6) * 4 + 2);
        tc_fence_after();
        const uint64_t v_desc = v_desc0 + st * (ATT_TILE_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < ATT_BK / 16; ++k)
          umma_ts(tO, tP + k * 8, v_desc + k * (2048 >> 4), idesc_o, (it > 0 || k > 0) ? 1u : 0u);
        umma_commit(&kv_empty[st]);
        umma_commit(pv_done);
        AF_TR(64 + (it - 6) * 4 + 3);
        if (it + 2 < n_it) {  // refill this K/V stage for tile it+2 once P·V has drained it
          mbar_wait(&kv_empty[st], (it >> 1) & 1);
          load_kv(it + 2);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warps
    const int r = warp * 32 + lane;          // row within the tile == TMEM lane
    const int qi = qt * ATT_BQ + r;          // position within the sequence
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const bool row_ok = qi < T;                // ragged tail: T need not be a multiple of 128
    const int seg_lo = (seg_start && row_ok) ? seg_start[row0 + r] : 0;
    float m_run = -INFINITY, l_run = 0.f;

    for (int it = 0; it < n_it; ++it) {
      const int j = j_lo + it;
      const bool trw = tr_on && warp == 0;
#define AF_TRS(k_)                                                                   \
  do {                                                                               \
    if (trw && it >= 6 && it < 10 && lane == 0) trace[(it - 6) * 8 + (k_)] = clock64(); \
  } while (0)
      AF_TRS(0);
      mbar_wait(s_full, it & 1);
      AF_TRS(1);
      tc_fence_after();
      // The whole score row (128 fp32) is pulled into registers with four back-to-back tcgen05.ld and ONE wait (a load
      // per pass and per 32-column chunk serialises eight TMEM round trips per tile), and tensor memory is handed back
      // at once: the next Q K^T runs under this tile's max / exp2 / P-store work.  The softmax scale is folded into the
      // exp2 argument.
      const int kbase = j * ATT_BK;
      const bool masked = (kbase + ATT_BK - 1 > qi) || (kbase < seg_lo);  // key kj allowed iff seg_lo <= kj <= qi
      uint32_t t[ATT_BK];
#pragma unroll
      for (int c = 0; c < ATT_BK / 32; ++c)
        tmem_ld32(tS + lane_off + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&t[c * 32]));
      tmem_ld_wait();
      AF_TRS(2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);  // S may be overwritten by the next Q K^T
      if (masked) {
#pragma unroll
        for (int i = 0; i < ATT_BK; ++i) {
          const int kj = kbase + i;
          if (kj > qi || kj < seg_lo) t[i] = 0xff800000u;  // -inf
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < ATT_BK; i += 8) {  // four independent chains of 3-input maxima (FMNMX3): 64 instructions
        mx0 = max3f(mx0, __uint_as_float(t[i]), __uint_as_float(t[i + 1]));
        mx1 = max3f(mx1, __uint_as_float(t[i + 2]), __uint_as_float(t[i + 3]));
        mx2 = max3f(mx2, __uint_as_float(t[i + 4]), __uint_as_float(t[i + 5]));
        mx3 = max3f(mx3, __uint_as_float(t[i + 6]), __uint_as_float(t[i + 7]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
      AF_TRS(3);
      // running max / lazy rescale decision (registers only; O itself is rescaled after the exp pass, below)
      const bool grow = mx > m_run + 8.0f;
      const bool any_grow = __any_sync(0xffffffffu, grow);
      float alpha = 1.0f;
      if (any_grow) {
        const float m_new = fmaxf(m_run, mx);
        alpha = (m_new == -INFINITY) ? 1.0f : ex2(m_run - m_new);  // m_run = -inf -> 0
        m_run = m_new;
        l_run *= alpha;
      }
      const float m_use = (m_run == -INFINITY) ? 0.f : m_run;
      const float2 sc2 = make_float2(scale_log2, scale_log2);
      const float2 nm2 = make_float2(-m_use, -m_use);
      float2 ps0 = make_float2(0.f, 0.f), ps1 = make_float2(0.f, 0.f);
      // exp2 pass, in place: P (packed bf16 pairs) overwrites the first half of the score registers, so that nothing
      // here depends on the previous tile's P·V yet
#pragma unroll
      for (int c16 = 0; c16 < ATT_BK / 8; ++c16) {
        float2 e[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 a = __ffma2_rn(
              make_float2(__uint_as_float(t[c16 * 8 + 2 * i]), __uint_as_float(t[c16 * 8 + 2 * i + 1])), sc2, nm2);
          e[i] = make_float2(ex2(a.x), ex2(a.y));
        }
        ps0 = __fadd2_rn(ps0, __fadd2_rn(e[0], e[1]));
        ps1 = __fadd2_rn(ps1, __fadd2_rn(e[2], e[3]));
#pragma unroll
        for (int i = 0; i < 4; ++i) t[c16 * 4 + i] = pack_bf16x2(e[i].x, e[i].y);  // slots < 8*c16: already consumed
      }
      AF_TRS(4);
      // the previous P·V must be complete before O is rescaled or the P tile in smem is overwritten
      if (it > 0) {
        mbar_wait(pv_done, (it - 1) & 1);
        tc_fence_after();
        if (any_grow) {
#pragma unroll
          for (int c = 0; c < ATT_HD / 16; ++c) {  // 16 columns at a time: the score row is live in registers
            uint32_t o[16];
            tmem_ld16(tO + lane_off + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tO + lane_off + c * 16, o);
          }
          tmem_st_wait();
        }
      }
      // P goes to tensor memory (A operand of the P·V MMA, read in place): no smem round trip — hd = 64 MMAs are
      // shared-memory-bandwidth bound, and the P tile was 44 % of this kernel's smem traffic
      tmem_st32(tP + lane_off, *reinterpret_cast<const uint32_t(*)[32]>(&t[0]));
      tmem_st32(tP + lane_off + 32, *reinterpret_cast<const uint32_t(*)[32]>(&t[32]));
      tmem_st_wait();
      const float psum = (ps0.x + ps0.y) + (ps1.x + ps1.y);
      l_run += psum;
      AF_TRS(5);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      AF_TRS(6);
    }

    // ---- epilogue: O / l -> bf16 out[b, t, h, :], lse
    mbar_wait(pv_done, (n_it - 1) & 1);
    tc_fence_after();
    const float inv_l = l_run > 0.f ? 1.0f / l_run : 0.f;
    __nv_bfloat16* orow = out + (row0 + r) * d + h * ATT_HD;
#pragma unroll
    for (int c = 0; c < ATT_HD / 32; ++c) {
      uint32_t t[32];
      tmem_ld32(tO + lane_off + c * 32, t);
      tmem_ld_wait();
      if (!row_ok) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(t[8 * i + 0]) * inv_l, __uint_as_float(t[8 * i + 1]) * inv_l);
        v.y = pack_bf16x2(__uint_as_float(t[8 * i + 2]) * inv_l, __uint_as_float(t[8 * i + 3]) * inv_l);
        v.z = pack_bf16x2(__uint_as_float(t[8 * i + 4]) * inv_l, __uint_as_float(t[8 * i + 5]) * inv_l);
        v.w = pack_bf16x2(__uint_as_float(t[8 * i + 6]) * inv_l, __uint_as_float(t[8 * i + 7]) * inv_l);
        *reinterpret_cast<uint4*>(orow + c * 32 + i * 8) = v;
      }
    }
    const float m_use = (m_run == -INFINITY) ? 0.f : m_run;
    if (row_ok) lse[(static_cast<int64_t>(b) * H + h) * T + qi] = (m_use + lg2(l_run)) * 0.6931471805599453f;
  }

  tc_fence_before();
  __syncthreads();
  if (tr_on && threadIdx.x == 0) {
    trace[121] = clock64();
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    trace[123] = gt;
  }
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace plm

extern "C" int plm_attn_fwd_v1(const void* qkv, const int32_t* seg_start, void* out, float* lse, int32_t B, int32_t T,
                            int32_t H, int32_t hd, plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(qkv);
  PLM_REQUIRE(qkv && out && lse, "attn_fwd_v1: null pointer");
  PLM_REQUIRE(B > 0 && T > 0 && H > 0, "attn_fwd_v1: bad size");
  if (hd != ATT_HD) return fail(PLM_ERR_UNSUPPORTED, "attn_fwd_v1: head_dim %d unsupported (need 64)", hd);
    PLM_REQUIRE(aligned16(qkv) && aligned16(out), "attn_fwd_v1: misaligned pointer");
  PLM_REQUIRE(static_cast<int64_t>(B) * T < (1ll << 31) && B <= 65535 && H <= 65535, "attn_fwd_v1: size too large");

  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(attn_fwd_v1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_FWD_SMEM);
  });
  if (attr_err != cudaSuccess) return fail(PLM_ERR_CUDA, "attn_fwd_v1 smem attribute: %s", cudaGetErrorString(attr_err));

  const int d = H * hd;
  CUtensorMap tm;
  int rc = make_tmap_bf16_2d(&tm, qkv, static_cast<uint64_t>(B) * T, 3ull * d, 3ull * d, ATT_BK, 64);
  if (rc != PLM_OK) return rc;
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(hd));
  dim3 grid((T + ATT_BQ - 1) / ATT_BQ, H, B);
  unsigned long long* trace = nullptr;  // (the clock64 trace hook of round 1 is no longer reachable from the ABI)
  attn_fwd_v1_kernel<<<grid, ATT_THREADS, ATT_FWD_SMEM, stream>>>(tm, seg_start, static_cast<__nv_bfloat16*>(out), lse, T,
                                                               H, scale_log2, trace);
  return check_launch("attn_fwd_v1");
}

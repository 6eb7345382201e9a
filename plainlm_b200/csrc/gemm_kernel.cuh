// tcgen05 GEMM kernel for every nn.Linear of the plainLM train step (models/transformer.py:42,67,114;
// models/components.py:55-56) — forward, dgrad and wgrad — replacing the cuBLASLt calls PyTorch makes under autocast.
//
// One persistent CTA per SM, 192 threads (320 for the GLU-backward kind: eight epilogue warps); launched as 2-CTA
// clusters whenever there is more than one M block:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128B-swizzled boxes, STAGES-deep mbarrier ring)
//   warp 1      MMA issuer     (tcgen05.mma kind::f16, bf16 x bf16 -> fp32 in TMEM; owns TMEM alloc/dealloc).  In PAIR
//               mode the leader CTA issues ONE tcgen05.mma.cta_group::2 (M = 256) per K-step for both CTAs.
//   warps 2..5  epilogue       (tcgen05.ld 32x32b, one accumulator row per thread; fused math; 128B-swizzled smem staging
//               tile; one TMA bulk store / reduce-add per 128-byte-wide column chunk).  Kinds: plain bf16 / fp32 / fp32
//               reduce-add, RoPE (table rows cached in smem through TMA), fp32 residual, SwiGLU (fc1), cross-entropy
//               statistics (LM head), GLU backward (fc2 dgrad; its a / z operand tiles arrive through TMA too).
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the main loop of
// tile i+1.  Tile = 128 x BN x 64 per CTA with BN in {128, 256}.  Operands may be K-major or MN-major (UMMA descriptor +
// instruction-descriptor major bits), which is what lets dgrad and wgrad read activations/weights in place.
//
// The epilogue kind is a TEMPLATE parameter: every kind gets its own kernel (own register allocation, no dead
// branches), instantiated in its own translation unit (gemm_epi_*.cu) so the library builds in parallel.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

#include <mutex>

namespace plm {

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#ifdef PLM_GEMM_DEBUG
#define PLM_DBG(p_) ((p_).debug)
#else
#define PLM_DBG(p_) 0
#endif

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_EPI_WARPS = 4;   // one per TMEM lane quarter ...
// ... or two: the GLU backward runs EIGHT epilogue warps, the two warps of a lane quarter splitting every 64-column
// chunk into its 32-column halves (105.5 -> 102.0 us at the 420M shape; that epilogue is bound by shared-memory
// bandwidth — its a/z/da/dz tiles add 512 KB per output tile to the 1 MB the main loop already moves — not by issue
// slots).  The cross-entropy kind has the same two-warp code path (EW == 8) but measured SLOWER with it when the logits
// are stored (1.289 -> 1.337 ms; faster only in the loss-only form, 1.357 -> 1.253 ms), so it stays at four.
template <int EPI>
constexpr int epi_warps() {
  return (EPI == PLM_EPI_BF16_GLU_BWD) ? 2 * GEMM_EPI_WARPS : GEMM_EPI_WARPS;
}
constexpr int EPI_BUF_BYTES = 128 * 128;  // staging tile: 128 rows x 128 B (64 bf16 or 32 fp32 columns)

struct GemmParams {
  void* C;
  const float* R;
  const float* rope;
  int64_t M, N, K;
  int64_t ldc;
  int epilogue;
  int splits;
  int rope_cols, rope_T, head_dim;
  int rope_cached;  // PLM_EPI_BF16_ROPE: 1 = the tile's 128 (cos,sin) rows are TMA-loaded into shared memory once per row block
                    // (head_dim == 64, rope_T % 128 == 0); 0 = coalesced global fetch per sub-chunk
  int glu_F;  // PLM_EPI_BF16_SWIGLU: F = N/2; tile n covers gate columns [128n, 128n+128) and up columns F + the same
  int num_m, num_n, kblocks;
  int debug;      // only read when the library is compiled with -DPLM_GEMM_DEBUG (timing experiments: 1 = skip epilogue
                  // operand loads, 2 = skip stores, 4 = skip B tile loads, 8 = skip A tile loads, 16 = skip only the
                  // epilogue's global operand loads; results are then garbage).  Production builds fold it to 0.
  int n_fastest;  // tile rasterisation: 0 = consecutive tiles walk M (B tile reused), 1 = walk N (A tile reused)
  // PLM_EPI_BF16_CE (LM head fused with the cross-entropy statistics, models/transformer.py:114 + engine/engine.py:110-112)
  const int64_t* ce_targets;  // [M]
  float2* ce_partial;         // [num_n, M]: per (column tile, row) base-2 (max, sum of 2^(x log2e - max))
  float* ce_tgt_logit;        // [M]: the (bf16-rounded) logit of the row's target column
  int ce_store;               // 0: loss-only forward, the logits tile is not written to C
};

// PAIR: the two CTAs of a cluster run ONE tcgen05.mma.cta_group::2 (M = 256) per K-step; each CTA's smem then holds
// only its half of the B tile (N/2 rows), so a stage shrinks from 48 to 32 KB and the ring deepens from 4 to 6.
// PAIR with BN = 128 (256 x 128 per pair) exists for outputs whose 256-wide tiling quantises badly over 74 pairs
// (N = 1024: 256 tiles = 3.46 waves; 512 narrow tiles = 6.92).
// The RoPE epilogue keeps the (cos,sin) rows of its 128-row block in shared memory (32 KB at head_dim 64) and pays for
// them with one operand stage; so does the GLU-backward epilogue for its second pair of (a, z) staging tiles.
constexpr int ROPE_TABLE_BYTES = 128 * 64 * 4;  // 128 positions x (32 pairs x (cos, sin)) fp32: two 128-byte-wide TMA boxes
template <int BN, bool PAIR, int EPI>
struct GemmCfg {
  static constexpr bool ROPE = (EPI == PLM_EPI_BF16_ROPE);
  static constexpr bool GLUB = (EPI == PLM_EPI_BF16_GLU_BWD);  // three (a, z) staging pairs instead of two staging tiles
  static constexpr bool WIDE = (BN == 256 && !PAIR);  // 48 KB stages
  static constexpr int STAGES = WIDE ? (GLUB ? 2 : ROPE ? 3 : 4) : (GLUB ? 4 : ROPE ? 5 : 6);
  static constexpr int GLUB_PAIRS = 3;
  static constexpr int EPI_BUFS = GLUB ? 2 * GLUB_PAIRS : 2;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int AUX_BYTES = ROPE ? ROPE_TABLE_BYTES : 0;
  static constexpr int BAR_BYTES = (2 * STAGES + 8) * 8 + 16;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BUFS * EPI_BUF_BYTES + AUX_BYTES + BAR_BYTES + 1024;  // + alignment slack
  static_assert(SMEM_BYTES <= 232448, "over the 227 KB shared-memory limit of sm_100");
};

// Work item -> (m block, n block, k range).  With CL = 2 a work item is a PAIR of M-adjacent tiles processed by the two
// CTAs of a cluster in lock-step; `rank` selects the CTA's half of the pair.
template <int CL>
__device__ __forceinline__ void decode_work(const GemmParams& p, int w, int rank, int& m_blk, int& n_blk, int& kb0,
                                            int& kb1) {
  const int num_mg = (p.num_m + CL - 1) / CL;
  const int tiles = num_mg * p.num_n;
  const int tile = w % tiles;
  const int split = w / tiles;
  int mg;
  if (p.n_fastest) {
    n_blk = tile % p.num_n;
    mg = tile / p.num_n;
  } else {
    mg = tile % num_mg;
    n_blk = tile / num_mg;
  }
  m_blk = mg * CL + rank;
  const int per = (p.kblocks + p.splits - 1) / p.splits;
  kb0 = split * per;
  kb1 = min(p.kblocks, kb0 + per);
}

// EPI: PLM_EPI_BF16 | PLM_EPI_BF16_ROPE | PLM_EPI_F32 (also serves PLM_EPI_ATOMIC_F32: same code, the bulk store becomes a
// bulk reduce-add) | PLM_EPI_RESID_F32 | PLM_EPI_BF16_SWIGLU | PLM_EPI_BF16_CE | PLM_EPI_BF16_GLU_BWD
template <int EPI, int BN, bool A_K, bool B_K, bool PAIR>
__global__ void __launch_bounds__((2 + epi_warps<EPI>()) * 32, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2, const GemmParams p) {
  using Cfg = GemmCfg<BN, PAIR, EPI>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int CL = PAIR ? 2 : 1;
  static_assert(!PAIR || BN == 256, "the CTA-pair MMA is instantiated for 256-wide tiles only");
  const int rank = (CL == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const int cluster_id = blockIdx.x / CL;
  const int num_clusters = gridDim.x / CL;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* sEpi = smem + STAGES * Cfg::STAGE_BYTES;  // staging tiles of the TMA-store epilogue
  uint8_t* sRope = sEpi + Cfg::EPI_BUFS * EPI_BUF_BYTES;  // RoPE kind: (cos,sin) rows of the current 128-row block
  uint64_t* full = reinterpret_cast<uint64_t*>(sRope + Cfg::AUX_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* rope_full = tempty + 2;
  uint64_t* ufull = rope_full + 1;  // [3] GLU backward: the (a, z) chunk pair has landed in its staging tiles
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ufull + 3);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total = ((p.num_m + CL - 1) / CL) * p.num_n * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    tma_prefetch_desc(&tmC2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);  // PAIR: the leader's single commit covers the pair's MMA
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], PAIR ? 2 * epi_warps<EPI>() : epi_warps<EPI>());  // PAIR: both CTAs' epilogues drain first
    }
    mbar_init(rope_full, 1);
    for (int b = 0; b < 3; ++b) mbar_init(&ufull[b], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) {
      tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
      tmem_relinquish_pair();
    } else {
      tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();  // the peer's barriers must be initialised before anything can arrive on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int w = cluster_id; w < total; w += num_clusters) {
        int m_blk, n_blk, kb0, kb1;
        decode_work<CL>(p, w, rank, m_blk, n_blk, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* a_dst = sA + s * Cfg::A_BYTES;
          uint8_t* b_dst = sB + s * Cfg::B_BYTES;
          if (PAIR) {
            // Both CTAs land their A tile and their half of the B tile (N rows [128 rank, +128) of the tile; with the
            // SwiGLU epilogue: gate rows / up rows) in their OWN smem; all bytes are signalled on the LEADER's barrier,
            // which its MMA thread waits on.
            if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);
            if (A_K) {
              tma_load_2d_pair(a_dst, &tmA, &full[s], kb * BK, m_blk * BM);
            } else {
#pragma unroll
              for (int g = 0; g < BM / 64; ++g)
                tma_load_2d_pair(a_dst + g * (BK * 128), &tmA, &full[s], m_blk * BM + g * 64, kb * BK);
            }
            if (B_K) {
              const int b_row = (EPI == PLM_EPI_BF16_SWIGLU) ? (rank ? p.glu_F : 0) + n_blk * (BN / 2)
                                                             : n_blk * BN + rank * (BN / 2);
              tma_load_2d_pair(b_dst, &tmB, &full[s], kb * BK, b_row);
            } else {
#pragma unroll
              for (int g = 0; g < BN / 128; ++g)
                tma_load_2d_pair(b_dst + g * (BK * 128), &tmB, &full[s], n_blk * BN + (rank * (BN / 128) + g) * 64,
                                 kb * BK);
            }
          } else {
            mbar_arrive_expect_tx(&full[s],
                                  ((PLM_DBG(p) & 8) ? 0 : Cfg::A_BYTES) + ((PLM_DBG(p) & 4) ? 0 : Cfg::B_BYTES));
            if (PLM_DBG(p) & 8) {
            } else if (A_K) {
              tma_load_2d(a_dst, &tmA, &full[s], kb * BK, m_blk * BM);
            } else {
#pragma unroll
              for (int g = 0; g < BM / 64; ++g)
                tma_load_2d(a_dst + g * (BK * 128), &tmA, &full[s], m_blk * BM + g * 64, kb * BK);
            }
            if (PLM_DBG(p) & 4) {
            } else if (B_K && EPI == PLM_EPI_BF16_SWIGLU) {  // gate rows then up rows (the map's box is BN/2 rows here)
              tma_load_2d(b_dst, &tmB, &full[s], kb * BK, n_blk * (BN / 2));
              tma_load_2d(b_dst + (BN / 2) * 128, &tmB, &full[s], kb * BK, p.glu_F + n_blk * (BN / 2));
            } else if (B_K) {
              tma_load_2d(b_dst, &tmB, &full[s], kb * BK, n_blk * BN);
            } else {
#pragma unroll
              for (int g = 0; g < BN / 64; ++g)
                tma_load_2d(b_dst + g * (BK * 128), &tmB, &full[s], n_blk * BN + g * 64, kb * BK);
            }
          }
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && (!PAIR || rank == 0)) {  // PAIR: the leader CTA issues for both
      constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 2 * BM : BM, BN, A_K ? 0 : 1, B_K ? 0 : 1);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int w = cluster_id; w < total; w += num_clusters, ++it) {
        int m_blk, n_blk, kb0, kb1;
        decode_work<CL>(p, w, rank, m_blk, n_blk, kb0, kb1);
        const int a = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty[a], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + s * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + s * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = A_K ? make_smem_desc_sw128(a_addr + k * 32, 16, 1024)
                                       : make_smem_desc_sw128(a_addr + k * 2048, BK * 128, 1024);
            const uint64_t bdesc = B_K ? make_smem_desc_sw128(b_addr + k * 32, 16, 1024)
                                       : make_smem_desc_sw128(b_addr + k * 2048, BK * 128, 1024);
            if (PAIR)
              umma_ss_pair(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else
              umma_ss(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (PAIR)
            umma_commit_pair(&empty[s]);
          else
            umma_commit(&empty[s]);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        if (PAIR)
          umma_commit_pair(&tfull[a]);  // both CTAs' epilogues read their own 128 rows of the accumulator
        else
          umma_commit(&tfull[a]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    // TMEM -> registers (one accumulator row per thread) -> fused math -> 128B-swizzled smem staging tile ->
    // ONE TMA store (or fp32 reduce-add) per 128-byte-wide column chunk.  Register<->global accesses with the
    // row-per-lane layout touch 32 cache lines per warp instruction and made the epilogue longer than the main loop,
    // so nothing here uses them: outputs leave through TMA, and the operands that come from global memory (fp32
    // residual rows, RoPE (cos,sin) rows) are fetched with a coalesced layout (8 lanes per 128-byte row segment, one
    // sub-chunk ahead) and transposed to row-per-lane through the staging tile.  (Streaming the residual tile in through
    // TMA instead — loaded into the staging tile four chunks ahead, added in place, stored again — was built and
    // measured in round 2: bit-identical, no faster at K = 1024 and 7 % slower at K = 2816 because the extra staging
    // tiles cost two operand stages; these GEMMs are bound by the L2 -> SM operand feed, not by their epilogue.)
    constexpr bool OUT_BF16 = (EPI == PLM_EPI_BF16 || EPI == PLM_EPI_BF16_ROPE || EPI == PLM_EPI_BF16_CE);
    constexpr int EW = epi_warps<EPI>();
    const int half = (warp - 2) >> 2;          // EW == 8: which 32-column half of every 64-column chunk this warp takes
    const int q = warp & 3;                    // TMEM lane quarter this warp may read
    const int r_tile = q * 32 + lane;          // row within the tile
    const bool elected = (threadIdx.x == 64);  // warp 2, lane 0 issues the bulk stores
    const bool is_rope = (EPI == PLM_EPI_BF16_ROPE) && !(PLM_DBG(p) & 1);
    const bool is_resid = (EPI == PLM_EPI_RESID_F32) && !(PLM_DBG(p) & 1);
    const int piece = lane & 7;                // coalesced layout: 16-byte piece of a row's 128-byte segment
    const int lrow0 = q * 32 + (lane >> 3);    // ... of rows lrow0 + 4j, j = 0..7
    const uint32_t own_off = r_tile * 128;
    const int own_sw = r_tile & 7;
    auto release_accumulator = [&](int a) {  // last TMEM read of this tile done: hand the buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR)
          mbar_arrive_cluster(&tempty[a], 0);  // the leader's MMA thread owns the pair's accumulators
        else
          mbar_arrive(&tempty[a]);
      }
    };
    int it = 0;
    uint32_t chunk_no = 0;
    // RoPE, cached table: every epilogue thread tracks (uniformly) which 128-row block's (cos,sin) rows sit in sRope and
    // how many loads have been issued; the elected thread issues the loads — the first here, the others right after the
    // last chunk of the previous tile, i.e. a whole main loop ahead of their use.
    // GLU backward, elected thread only: the (a, z) chunk loads run one chunk ahead of the chunk being processed
    int pf_w = cluster_id, pf_c = 0;
    uint32_t pf_issued = 0;
    const bool rope_cached = (EPI == PLM_EPI_BF16_ROPE) && p.rope_cached && !(PLM_DBG(p) & 1);
    int rope_m = -1;          // row block whose table rows are (being) loaded
    uint32_t rope_loads = 0;  // loads issued so far (parity of rope_full)
    auto rope_tile_needs = [&](int w2, int& m2) -> bool {  // does work item w2 rotate anything, and which row block?
      if (w2 >= total) return false;
      int n2, k0, k1;
      decode_work<CL>(p, w2, rank, m2, n2, k0, k1);
      return static_cast<int64_t>(n2) * BN < p.rope_cols;
    };
    auto rope_issue = [&](int m2) {  // elected thread: 128 positions x 64 floats as two 128-byte-wide boxes
      const int pos0 = static_cast<int>((static_cast<int64_t>(m2) * BM) % p.rope_T);
      mbar_arrive_expect_tx(rope_full, ROPE_TABLE_BYTES);
      tma_load_2d(sRope, &tmC2, rope_full, 0, pos0);
      tma_load_2d(sRope + ROPE_TABLE_BYTES / 2, &tmC2, rope_full, 32, pos0);
    };
    if (rope_cached) {
      int m2;
      if (rope_tile_needs(cluster_id, m2)) {
        if (elected) rope_issue(m2);
        rope_m = m2;
        ++rope_loads;
      }
    }
    for (int w = cluster_id; w < total; w += num_clusters, ++it) {
      int m_blk, n_blk, kb0, kb1;
      decode_work<CL>(p, w, rank, m_blk, n_blk, kb0, kb1);
      const int a = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int64_t row_base = static_cast<int64_t>(m_blk) * BM;
      const int64_t tile_col0 = static_cast<int64_t>(n_blk) * BN;
      const int64_t cols_left = p.N - tile_col0;
      const int n_sub = static_cast<int>(cols_left < BN ? (cols_left + 31) >> 5 : BN / 32);  // 32-column sub-chunks
      int pos_j[8];
      if (is_rope) {
#pragma unroll
        for (int j = 0; j < 8; ++j) pos_j[j] = static_cast<int>((row_base + lrow0 + 4 * j) % p.rope_T);
      }
      // coalesced fetch of sub-chunk `sc`'s global operand: dst[j] = piece `piece` of row lrow0 + 4j
      auto fetch_aux = [&](float4(&dst)[8], int sc) {
        if (PLM_DBG(p) & 16) return;  // timing experiment: keep the staging + math, drop the global loads
        const int64_t col0 = tile_col0 + sc * 32;
        if (is_resid) {
          const int64_t colp = col0 + piece * 4;
          if (colp < p.N) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int64_t row = row_base + lrow0 + 4 * j;
              if (row < p.M) dst[j] = __ldg(reinterpret_cast<const float4*>(p.R + row * p.ldc + colp));
            }
          }
        } else if (is_rope && col0 < p.rope_cols) {
          const int off = static_cast<int>(col0 % p.head_dim) + piece * 4;  // (cos,sin) pairs: 2 floats per 2 columns
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = __ldg(reinterpret_cast<const float4*>(p.rope + static_cast<int64_t>(pos_j[j]) * p.head_dim + off));
        }
      };
      auto stage_aux = [&](uint8_t* buf, const float4(&src)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rr = lrow0 + 4 * j;
          *reinterpret_cast<float4*>(buf + rr * 128 + ((piece ^ (rr & 7)) << 4)) = src[j];
        }
      };
      if (is_resid) {  // pull residual row segments into L2 well ahead of their use: this tile's on the first
                       // iteration, and always the NEXT tile's (its main loop has not even finished yet)
        auto l2_prefetch_rows = [&](int mb, int nb) {
          const int64_t row = static_cast<int64_t>(mb) * BM + r_tile;
          const int64_t c0 = static_cast<int64_t>(nb) * BN;
          if (row < p.M) {
            for (int sc = 0; sc < BN / 32 && c0 + sc * 32 < p.N; ++sc)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.R + row * p.ldc + c0 + sc * 32));
          }
        };
        if (it == 0) l2_prefetch_rows(m_blk, n_blk);
        if (w + num_clusters < total) {
          int m2, n2, k0, k1;
          decode_work<CL>(p, w + num_clusters, rank, m2, n2, k0, k1);
          l2_prefetch_rows(m2, n2);
        }
      }
      float4 nxt[8];
      if ((EPI == PLM_EPI_BF16_ROPE && !rope_cached) || EPI == PLM_EPI_RESID_F32) fetch_aux(nxt, 0);
      if (rope_cached && tile_col0 < p.rope_cols) {
        if (rope_m != m_blk) {  // not prefetched (the previous tile's look-ahead saw another row block): load it now
          if (elected) rope_issue(m_blk);
          rope_m = m_blk;
          ++rope_loads;
        }
        mbar_wait(rope_full, (rope_loads - 1) & 1);
      }
      // cross-entropy statistics of this row over this tile's columns (base 2), and where its target column sits
      float ce_m = -INFINITY, ce_s = 0.f;
      int tgt_local = -1;
      if (EPI == PLM_EPI_BF16_CE) {
        const int64_t row = row_base + r_tile;
        if (row < p.M) {
          const int64_t t = p.ce_targets[row];
          if (t >= tile_col0 && t < tile_col0 + BN && t < p.N) tgt_local = static_cast<int>(t - tile_col0);
        }
      }
      mbar_wait(&tfull[a], aph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * BN;
      const int r0 = static_cast<int>(row_base);
      if constexpr (EPI == PLM_EPI_BF16_SWIGLU) {
        // accumulator columns [0,128) = gate a, [128,256) = up z.  Per 64-column group: write a and z to their places in
        // u, then h = silu(a) * z (from the bf16-rounded values, as the stand-alone kernel computes it) to C2.
        auto emit = [&](const uint32_t(&o)[32], const CUtensorMap* map, int c0) {
          uint8_t* buf = sEpi + (chunk_no & 1) * EPI_BUF_BYTES;
          if (elected) bulk_wait_group_read<1>();
          named_bar_sync(1, EW * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<uint4*>(buf + own_off + ((i ^ own_sw) << 4)) =
                make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
          fence_proxy_async_smem();
          named_bar_sync(1, EW * 32);
          if (elected && !(PLM_DBG(p) & 2)) {
            tma_store_2d(map, buf, c0, r0);
            bulk_commit_group();
          }
          ++chunk_no;
        };
        const int gate_col0 = n_blk * (BN / 2);
#pragma unroll 1
        for (int j = 0; j < BN / 128; ++j) {
          uint32_t oa[32], oz[32];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t r[32];
            tmem_ld32(t_row + j * 64 + h * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              oa[16 * h + i] = pack_bf16x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
            tmem_ld32(t_row + BN / 2 + j * 64 + h * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              oz[16 * h + i] = pack_bf16x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
          }
          if (j == BN / 128 - 1) release_accumulator(a);
          emit(oa, &tmC, gate_col0 + j * 64);
          emit(oz, &tmC, p.glu_F + gate_col0 + j * 64);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float a0 = bf16_lo(oa[i]), a1 = bf16_hi(oa[i]);
            const float z0 = bf16_lo(oz[i]), z1 = bf16_hi(oz[i]);
            oa[i] = pack_bf16x2(a0 * sigmoidf_fast(a0) * z0, a1 * sigmoidf_fast(a1) * z1);
          }
          emit(oa, &tmC2, gate_col0 + j * 64);
        }
      } else if constexpr (EPI == PLM_EPI_BF16_GLU_BWD) {
        // fc2's input-gradient GEMM fused with the GLU backward (models/components.py:55-56): the accumulator tile is
        // dg = d(silu(a) z); per 64-column chunk the matching a and z tiles of u = [a | z] are TMA-loaded into a staging
        // pair (three pairs in rotation) one chunk ahead, each thread turns its row IN PLACE into da = dg z (s + a s (1 - s)) and dz = dg a s
        // (s = sigmoid(a) = 0.5 tanh(a / 2) + 0.5: one MUFU op per element; packed f32x2 arithmetic), and the pair
        // leaves through two TMA stores into du = [da | dz].  dg itself is never written.
        constexpr int NCH = BN / 64;
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          constexpr uint32_t NP = Cfg::GLUB_PAIRS;
          const uint32_t pi = chunk_no % NP;
          uint8_t* aBuf = sEpi + pi * (2 * EPI_BUF_BYTES);
          uint8_t* zBuf = aBuf + EPI_BUF_BYTES;
          if (elected) {
            // load #L reuses the pair chunk L - 3 was stored from; at chunk n the stores 0..n-1 are committed and
            // L <= n + 1, so "all but the most recent group have finished reading" frees it without a stall (with two
            // pairs the store issued a moment ago had to drain first: 108.7 us per launch at the 420M shape)
            while (pf_issued <= chunk_no + 1 && pf_w < total) {
              if (pf_issued >= NP) bulk_wait_group_read<1>();
              int m2, n2, k0, k1;
              decode_work<CL>(p, pf_w, rank, m2, n2, k0, k1);
              const uint32_t b2 = pf_issued % NP;
              uint8_t* dst = sEpi + b2 * (2 * EPI_BUF_BYTES);
              mbar_arrive_expect_tx(&ufull[b2], 2 * EPI_BUF_BYTES);
              tma_load_2d(dst, &tmC2, &ufull[b2], n2 * BN + pf_c * 64, m2 * BM);
              tma_load_2d(dst + EPI_BUF_BYTES, &tmC2, &ufull[b2], p.glu_F + n2 * BN + pf_c * 64, m2 * BM);
              ++pf_issued;
              if (++pf_c == NCH) {
                pf_c = 0;
                pf_w += num_clusters;
              }
            }
          }
          mbar_wait(&ufull[pi], (chunk_no / NP) & 1);
          {
            const int h = half;  // eight epilogue warps: this warp's 32-column half of the chunk
            uint32_t r[32];
            tmem_ld32(t_row + c * 64 + h * 32, r);
            tmem_ld_wait();
            if (c == NCH - 1) release_accumulator(a);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t off = own_off + (((4 * h + i) ^ own_sw) << 4);
              const uint4 av = *reinterpret_cast<const uint4*>(aBuf + off);
              const uint4 zv = *reinterpret_cast<const uint4*>(zBuf + off);
              const uint32_t aw[4] = {av.x, av.y, av.z, av.w};
              const uint32_t zw[4] = {zv.x, zv.y, zv.z, zv.w};
              uint32_t da[4], dz[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 a2 = make_float2(bf16_lo(aw[k]), bf16_hi(aw[k]));
                const float2 z2 = make_float2(bf16_lo(zw[k]), bf16_hi(zw[k]));
                const float2 g2 = make_float2(__uint_as_float(r[8 * i + 2 * k]), __uint_as_float(r[8 * i + 2 * k + 1]));
                const float2 half2 = make_float2(0.5f, 0.5f);
                const float2 ha = __fmul2_rn(a2, half2);
                const float2 t2 = make_float2(tanh_approx(ha.x), tanh_approx(ha.y));
                const float2 s2 = __ffma2_rn(t2, half2, half2);                                   // sigmoid(a)
                const float2 as2 = __fmul2_rn(a2, s2);                                            // silu(a)
                const float2 oms = __ffma2_rn(s2, make_float2(-1.f, -1.f), make_float2(1.f, 1.f));  // 1 - s
                const float2 q2 = __ffma2_rn(as2, oms, s2);                                       // s + a s (1 - s)
                const float2 dz2 = __fmul2_rn(g2, as2);
                const float2 da2 = __fmul2_rn(__fmul2_rn(g2, z2), q2);
                da[k] = pack_bf16x2(da2.x, da2.y);
                dz[k] = pack_bf16x2(dz2.x, dz2.y);
              }
              *reinterpret_cast<uint4*>(aBuf + off) = make_uint4(da[0], da[1], da[2], da[3]);
              *reinterpret_cast<uint4*>(zBuf + off) = make_uint4(dz[0], dz[1], dz[2], dz[3]);
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(1, EW * 32);
          if (elected && !(PLM_DBG(p) & 2)) {
            const int c0 = static_cast<int>(tile_col0) + c * 64;
            tma_store_2d(&tmC, aBuf, c0, r0);
            tma_store_2d(&tmC, zBuf, p.glu_F + c0, r0);
            bulk_commit_group();
          }
          ++chunk_no;
        }
      } else if constexpr (OUT_BF16) {
        const int n_chunks = (n_sub + 1) >> 1;  // 64 bf16 columns per staging row
#pragma unroll 1
        for (int c = 0; c < n_chunks; ++c) {
          uint8_t* buf = sEpi + (chunk_no & 1) * EPI_BUF_BYTES;
          if (elected) bulk_wait_group_read<1>();  // the store that last read this staging tile has drained it
          named_bar_sync(1, EW * 32);
          uint32_t o[32];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (EW == 8 && h != half) continue;  // eight epilogue warps: each takes one 32-column half of the chunk
            const int sc = 2 * c + h;
            if (sc < n_sub) {  // warp-uniform
              const bool rope_on = is_rope && (tile_col0 + sc * 32 < p.rope_cols);
              if (rope_on && !rope_cached) {
                stage_aux(buf, nxt);
                __syncwarp();
              }
              if (EPI == PLM_EPI_BF16_ROPE && !rope_cached && sc + 1 < n_sub) fetch_aux(nxt, sc + 1);
              uint32_t r[32];
              tmem_ld32(t_row + sc * 32, r);
              tmem_ld_wait();
              if (rope_on) {
                // this row's 16 (cos,sin) pairs: from the cached table (box 0 = pairs 0..15, box 1 = pairs 16..31 of the
                // head; a 32-column sub-chunk starts at column 0 or 32 of a 64-wide head) or from the staged fetch
                const uint8_t* cs_row = rope_cached
                                            ? sRope + (((tile_col0 + sc * 32) & 32) ? ROPE_TABLE_BYTES / 2 : 0) + own_off
                                            : buf + own_off;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 cs = *reinterpret_cast<const float4*>(cs_row + ((i ^ own_sw) << 4));
                  const float a0 = __uint_as_float(r[4 * i + 0]), b0 = __uint_as_float(r[4 * i + 1]);
                  const float a1 = __uint_as_float(r[4 * i + 2]), b1 = __uint_as_float(r[4 * i + 3]);
                  r[4 * i + 0] = __float_as_uint(a0 * cs.x - b0 * cs.y);
                  r[4 * i + 1] = __float_as_uint(b0 * cs.x + a0 * cs.y);
                  r[4 * i + 2] = __float_as_uint(a1 * cs.z - b1 * cs.w);
                  r[4 * i + 3] = __float_as_uint(b1 * cs.z + a1 * cs.w);
                }
                if (!rope_cached) __syncwarp();  // every lane has read its staged (cos,sin) row before the tile is overwritten
              }
#pragma unroll
              for (int i = 0; i < 16; ++i)
                o[16 * h + i] = pack_bf16x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
              if constexpr (EPI == PLM_EPI_BF16_CE) {
                // online max / sum-exp over the bf16-ROUNDED logits (what the reference's CrossEntropyLoss sees after
                // autocast), base 2; columns past N (zero-filled B rows of the ragged last tile) count as -inf
                constexpr float LOG2E = 1.4426950408889634f;
                const int64_t left = cols_left - sc * 32;
                const int valid = left < 32 ? static_cast<int>(left) : 32;  // warp-uniform
                float x[32];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  x[2 * i] = bf16_lo(o[16 * h + i]);
                  x[2 * i + 1] = bf16_hi(o[16 * h + i]);
                }
                if (valid < 32) {
#pragma unroll
                  for (int i = 0; i < 32; ++i)
                    if (i >= valid) x[i] = -INFINITY;
                }
                // four independent max / sum chains: a single 32-deep dependent chain per reduction made this epilogue
                // longer than the main loop (ncu: tensor pipe 82 % active against 95 % for the plain LM-head GEMM)
                float mx0 = x[0], mx1 = x[1], mx2 = x[2], mx3 = x[3];
#pragma unroll
                for (int i = 4; i < 32; i += 4) {
                  mx0 = fmaxf(mx0, x[i]);
                  mx1 = fmaxf(mx1, x[i + 1]);
                  mx2 = fmaxf(mx2, x[i + 2]);
                  mx3 = fmaxf(mx3, x[i + 3]);
                }
                const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                const float m_new = fmaxf(ce_m, mx * LOG2E);
                float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  acc0 += ex2_approx(fmaf(x[i], LOG2E, -m_new));
                  acc1 += ex2_approx(fmaf(x[i + 1], LOG2E, -m_new));
                  acc2 += ex2_approx(fmaf(x[i + 2], LOG2E, -m_new));
                  acc3 += ex2_approx(fmaf(x[i + 3], LOG2E, -m_new));
                }
                const float acc = (acc0 + acc1) + (acc2 + acc3);
                ce_s = ce_s * ex2_approx(ce_m - m_new) + acc;
                ce_m = m_new;
              }
            }
          }
          if (c == n_chunks - 1) release_accumulator(a);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (EW == 8 && (i >> 2) != half) continue;  // this warp staged only its half of the row
            *reinterpret_cast<uint4*>(buf + own_off + ((i ^ own_sw) << 4)) =
                make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
          }
          if constexpr (EPI == PLM_EPI_BF16_CE) {
            // the row's target column is in the part of this chunk THIS thread staged: pick it out of the staged row
            if ((tgt_local >> 6) == c && (EW != 8 || ((tgt_local >> 5) & 1) == half)) {
              const int cc = tgt_local & 63;
              const uint16_t bits =
                  *reinterpret_cast<const uint16_t*>(buf + own_off + (((cc >> 3) ^ own_sw) << 4) + (cc & 7) * 2);
              p.ce_tgt_logit[row_base + r_tile] = __uint_as_float(static_cast<uint32_t>(bits) << 16);
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(1, EW * 32);
          if (elected && !(PLM_DBG(p) & 2) && (EPI != PLM_EPI_BF16_CE || p.ce_store)) {
            tma_store_2d(&tmC, buf, static_cast<int>(tile_col0) + c * 64, r0);
            bulk_commit_group();
          }
          ++chunk_no;
        }
        if constexpr (EPI == PLM_EPI_BF16_CE) {
          const int64_t row = row_base + r_tile;
          // EW == 8: two threads per row, each with the statistics of its column halves -> two partials per tile
          const int64_t slot = (EW == 8) ? 2 * static_cast<int64_t>(n_blk) + half : static_cast<int64_t>(n_blk);
          if (row < p.M) p.ce_partial[slot * p.M + row] = make_float2(ce_m, ce_s);
        }
        if (rope_cached) {  // every thread is past its last read of sRope (the chunk barrier above): fetch the next
                            // rotating tile's rows now, a whole main loop ahead of their use
          int m2;
          if (rope_tile_needs(w + num_clusters, m2) && m2 != rope_m) {
            if (elected) rope_issue(m2);
            rope_m = m2;
            ++rope_loads;
          }
        }
      } else {
#pragma unroll 1
        for (int sc = 0; sc < n_sub; ++sc) {  // 32 fp32 columns per staging row
          uint8_t* buf = sEpi + (chunk_no & 1) * EPI_BUF_BYTES;
          if (elected) bulk_wait_group_read<1>();
          named_bar_sync(1, EW * 32);
          if (is_resid) {
            stage_aux(buf, nxt);
            __syncwarp();
          }
          if (EPI == PLM_EPI_RESID_F32 && sc + 1 < n_sub) fetch_aux(nxt, sc + 1);
          uint32_t r[32];
          tmem_ld32(t_row + sc * 32, r);
          tmem_ld_wait();
          if (sc == n_sub - 1) release_accumulator(a);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4* slot = reinterpret_cast<float4*>(buf + own_off + ((i ^ own_sw) << 4));
            float4 v;
            v.x = __uint_as_float(r[4 * i + 0]);
            v.y = __uint_as_float(r[4 * i + 1]);
            v.z = __uint_as_float(r[4 * i + 2]);
            v.w = __uint_as_float(r[4 * i + 3]);
            if (is_resid) {  // the slot holds this row's residual piece; only this thread touches it from here on
              const float4 x = *slot;
              v.x += x.x;
              v.y += x.y;
              v.z += x.z;
              v.w += x.w;
            }
            *slot = v;
          }
          fence_proxy_async_smem();
          named_bar_sync(1, EW * 32);
          if (elected && !(PLM_DBG(p) & 2)) {
            const int c0 = static_cast<int>(tile_col0) + sc * 32;
            if (EPI == PLM_EPI_F32 && p.epilogue == PLM_EPI_ATOMIC_F32)
              tma_reduce_add_2d(&tmC, buf, c0, r0);
            else
              tma_store_2d(&tmC, buf, c0, r0);
            bulk_commit_group();
          }
          ++chunk_no;
        }
      }
    }
    if (elected) bulk_wait_group<0>();  // every store has landed before the CTA may exit
  }

  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();  // no CTA may exit while its peer can still signal into it
  if (warp == 1) {
    tc_fence_after();
    if (PAIR)
      tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
    else
      tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int EPI, int BN, bool A_K, bool B_K, bool PAIR>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmC2,
                       const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, PAIR, EPI>;
  constexpr int CL = PAIR ? 2 : 1;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_kernel<EPI, BN, A_K, B_K, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return fail(PLM_ERR_CUDA, "gemm smem attribute: %s", cudaGetErrorString(attr_err));
  const int total = ((p.num_m + CL - 1) / CL) * p.num_n * p.splits;
  int clusters = sm_count() / CL;
  if (total < clusters) clusters = total;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3((2 + epi_warps<EPI>()) * 32);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_kernel<EPI, BN, A_K, B_K, PAIR>, tmA, tmB, tmC, tmC2, p);
  if (e != cudaSuccess) return fail(PLM_ERR_CUDA, "gemm_kernel launch: %s", cudaGetErrorString(e));
  return check_launch("gemm_kernel");
}

// Tile configuration chosen by the host: bn in {128, 256}; pair = CTA-pair MMA (needs bn == 256).
struct GemmTile {
  int bn;
  bool pair;
};

// One entry point per epilogue kind, defined in gemm_epi_*.cu through PLM_DEFINE_GEMM_EPI.  The general kinds accept
// every operand major-ness; the fused forward kinds (RoPE, SwiGLU, cross-entropy) exist for K-major A and B only.
template <int EPI>
int gemm_launch_epi(bool a_k, bool b_k, GemmTile tile, const CUtensorMap& tmA, const CUtensorMap& tmB,
                    const CUtensorMap& tmC, const CUtensorMap& tmC2, const GemmParams& p, cudaStream_t stream);

#define PLM_GEMM_TILES_(EPI_, AK_, BK_)                                                             \
  if (tile.pair) return launch_gemm<EPI_, 256, AK_, BK_, true>(tmA, tmB, tmC, tmC2, p, stream);     \
  if (tile.bn == 256) return launch_gemm<EPI_, 256, AK_, BK_, false>(tmA, tmB, tmC, tmC2, p, stream); \
  return launch_gemm<EPI_, 128, AK_, BK_, false>(tmA, tmB, tmC, tmC2, p, stream);

#define PLM_DEFINE_GEMM_EPI_GENERAL(EPI_)                                                                          \
  template <>                                                                                                      \
  int gemm_launch_epi<EPI_>(bool a_k, bool b_k, GemmTile tile, const CUtensorMap& tmA, const CUtensorMap& tmB,     \
                            const CUtensorMap& tmC, const CUtensorMap& tmC2, const GemmParams& p,                  \
                            cudaStream_t stream) {                                                                 \
    if (a_k && b_k) { PLM_GEMM_TILES_(EPI_, true, true) }                                                          \
    if (a_k && !b_k) { PLM_GEMM_TILES_(EPI_, true, false) }                                                        \
    if (!a_k && b_k) { PLM_GEMM_TILES_(EPI_, false, true) }                                                        \
    PLM_GEMM_TILES_(EPI_, false, false)                                                                            \
  }

// fused backward kind: A K-major (the incoming gradient), B MN-major (the weight read in place), 256-wide tiles only
#define PLM_DEFINE_GEMM_EPI_DGRAD(EPI_)                                                                            \
  template <>                                                                                                      \
  int gemm_launch_epi<EPI_>(bool a_k, bool b_k, GemmTile tile, const CUtensorMap& tmA, const CUtensorMap& tmB,     \
                            const CUtensorMap& tmC, const CUtensorMap& tmC2, const GemmParams& p,                  \
                            cudaStream_t stream) {                                                                 \
    if (!(a_k && !b_k) || tile.bn != 256)                                                                          \
      return fail(PLM_ERR_UNSUPPORTED, "gemm: this fused epilogue needs a K-major A, an MN-major B and N %% 256 == 0"); \
    if (tile.pair) return launch_gemm<EPI_, 256, true, false, true>(tmA, tmB, tmC, tmC2, p, stream);               \
    return launch_gemm<EPI_, 256, true, false, false>(tmA, tmB, tmC, tmC2, p, stream);                             \
  }

#define PLM_DEFINE_GEMM_EPI_FORWARD(EPI_)                                                                          \
  template <>                                                                                                      \
  int gemm_launch_epi<EPI_>(bool a_k, bool b_k, GemmTile tile, const CUtensorMap& tmA, const CUtensorMap& tmB,     \
                            const CUtensorMap& tmC, const CUtensorMap& tmC2, const GemmParams& p,                  \
                            cudaStream_t stream) {                                                                 \
    if (!(a_k && b_k)) return fail(PLM_ERR_UNSUPPORTED, "gemm: this fused epilogue needs K-major A and B");        \
    PLM_GEMM_TILES_(EPI_, true, true)                                                                              \
  }

}  // namespace plm

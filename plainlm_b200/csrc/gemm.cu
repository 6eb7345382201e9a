// tcgen05 GEMM for every nn.Linear of the plainLM train step (models/transformer.py:42,67,114;
// models/components.py:55-56) — forward, dgrad and wgrad — replacing the cuBLASLt calls PyTorch makes under autocast.
//
// One persistent CTA per SM, 192 threads; launched as 2-CTA clusters whenever there is more than one M block:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128B-swizzled boxes, STAGES-deep mbarrier ring)
//   warp 1      MMA issuer     (tcgen05.mma kind::f16, bf16 x bf16 -> fp32 in TMEM; owns TMEM alloc/dealloc).  In PAIR
//               mode the leader CTA issues ONE tcgen05.mma.cta_group::2 (M = 256) per K-step for both CTAs.
//   warps 2..5  epilogue       (tcgen05.ld 32x32b, one accumulator row per thread; fused math; 128B-swizzled smem staging
//               tile; one TMA bulk store / reduce-add per 128-byte-wide column chunk)
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the main loop of
// tile i+1.  Tile = 128 x BN x 64 per CTA with BN in {128, 256}.  Operands may be K-major or MN-major (UMMA descriptor +
// instruction-descriptor major bits), which is what lets dgrad and wgrad read activations/weights in place.
#include "common.cuh"
#include "ptx.cuh"

#include <mutex>

namespace plm {

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

#ifdef PLM_GEMM_DEBUG
#define PLM_DBG(p_) ((p_).debug)
#else
#define PLM_DBG(p_) 0
#endif

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_EPI_WARPS = 4;   // one per TMEM lane quarter
constexpr int EPI_BUF_BYTES = 128 * 128;  // staging tile: 128 rows x 128 B (64 bf16 or 32 fp32 columns)
constexpr int GEMM_THREADS = (2 + GEMM_EPI_WARPS) * 32;

struct GemmParams {
  void* C;
  const float* R;
  const float* rope;
  int64_t M, N, K;
  int64_t ldc;
  int epilogue;
  int splits;
  int rope_cols, rope_T, head_dim;
  int glu_F;  // PLM_EPI_BF16_SWIGLU: F = N/2; tile n covers gate columns [128n, 128n+128) and up columns F + the same
  int num_m, num_n, kblocks;
  int debug;      // only read when the library is compiled with -DPLM_GEMM_DEBUG (timing experiments: 1 = skip epilogue
                  // operand loads, 2 = skip stores, 4 = skip B tile loads, 8 = skip A tile loads, 16 = skip only the
                  // epilogue's global operand loads; results are then garbage).  Production builds fold it to 0.
  int n_fastest;  // tile rasterisation: 0 = consecutive tiles walk M (B tile reused), 1 = walk N (A tile reused)
};

// PAIR: the two CTAs of a cluster run ONE tcgen05.mma.cta_group::2 (M = 256) per K-step; each CTA's smem then holds
// only its half of the B tile (N/2 rows), so a stage shrinks from 48 to 32 KB and the ring deepens from 4 to 6.
template <int BN, bool PAIR = false>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256 && !PAIR) ? 4 : 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * EPI_BUF_BYTES + BAR_BYTES + 1024;  // + alignment slack
};

// Work item -> (m block, n block, k range).  With CL = 2 a work item is a PAIR of M-adjacent tiles processed by the two
// CTAs of a cluster in lock-step (they share — and multicast — the B tile); `rank` selects the CTA's half of the pair.
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

template <int BN, bool A_K, bool B_K, int CL, bool PAIR>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2, const GemmParams p) {
  using Cfg = GemmCfg<BN, PAIR>;
  constexpr int STAGES = Cfg::STAGES;
  static_assert(CL == 1 || CL == 2, "cluster of 1 or 2 CTAs");
  static_assert(!PAIR || (CL == 2 && BN == 256), "the CTA-pair MMA needs a 2-CTA cluster and 256-wide tiles");
  const int rank = (CL == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const int cluster_id = blockIdx.x / CL;
  const int num_clusters = gridDim.x / CL;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* sEpi = smem + STAGES * Cfg::STAGE_BYTES;  // 2 staging tiles for the TMA-store epilogue
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + 2 * EPI_BUF_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

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
      // CL = 2: the stage is overwritten in BOTH CTAs, so both CTAs' MMAs must have consumed it (two multicast commits);
      // PAIR: the leader's single commit covers the pair's MMA
      mbar_init(&empty[s], PAIR ? 1 : CL);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], PAIR ? 2 * GEMM_EPI_WARPS : GEMM_EPI_WARPS);  // PAIR: both CTAs' epilogues drain first
    }
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
              const int b_row = p.glu_F ? (rank ? p.glu_F : 0) + n_blk * (BN / 2) : n_blk * BN + rank * (BN / 2);
              tma_load_2d_pair(b_dst, &tmB, &full[s], kb * BK, b_row);
            } else {
#pragma unroll
              for (int g = 0; g < BN / 128; ++g)
                tma_load_2d_pair(b_dst + g * (BK * 128), &tmB, &full[s], n_blk * BN + (rank * (BN / 128) + g) * 64,
                                 kb * BK);
            }
            if (++s == STAGES) {
              s = 0;
              ph ^= 1;
            }
            continue;
          }
          mbar_arrive_expect_tx(&full[s], ((PLM_DBG(p) & 8) ? 0 : Cfg::A_BYTES) + ((PLM_DBG(p) & 4) ? 0 : Cfg::B_BYTES));
          if (PLM_DBG(p) & 8) {
          } else if (A_K) {
            tma_load_2d(a_dst, &tmA, &full[s], kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int g = 0; g < BM / 64; ++g)
              tma_load_2d(a_dst + g * (BK * 128), &tmA, &full[s], m_blk * BM + g * 64, kb * BK);
          }
          if (PLM_DBG(p) & 4) {
          } else if (CL == 1) {
            if (B_K && p.glu_F) {  // gate rows then up rows (the tensor map's box is BN/2 rows here)
              tma_load_2d(b_dst, &tmB, &full[s], kb * BK, n_blk * (BN / 2));
              tma_load_2d(b_dst + (BN / 2) * 128, &tmB, &full[s], kb * BK, p.glu_F + n_blk * (BN / 2));
            } else if (B_K) {
              tma_load_2d(b_dst, &tmB, &full[s], kb * BK, n_blk * BN);
            } else {
#pragma unroll
              for (int g = 0; g < BN / 64; ++g)
                tma_load_2d(b_dst + g * (BK * 128), &tmB, &full[s], n_blk * BN + g * 64, kb * BK);
            }
          } else {
            // each CTA fetches HALF of the shared B tile and multicasts it into both CTAs' stage buffers:
            // L2 -> SM traffic per CTA drops from A + B to A + B/2
            if (B_K) {
              const int b_row = p.glu_F ? (rank ? p.glu_F : 0) + n_blk * (BN / 2) : n_blk * BN + rank * (BN / 2);
              tma_load_2d_mc(b_dst + rank * (BN / 2) * 128, &tmB, &full[s], kb * BK, b_row, 0x3);
            } else {
#pragma unroll
              for (int g = 0; g < BN / 128; ++g) {
                const int gg = rank * (BN / 128) + g;
                tma_load_2d_mc(b_dst + gg * (BK * 128), &tmB, &full[s], n_blk * BN + gg * 64, kb * BK, 0x3);
              }
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
          else if (CL == 2)
            umma_commit_mc(&empty[s], 0x3);
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
    // sub-chunk ahead) and transposed to row-per-lane through the staging tile.
    const int q = warp & 3;                    // TMEM lane quarter this warp may read
    const int r_tile = q * 32 + lane;          // row within the tile
    const bool elected = (threadIdx.x == 64);  // warp 2, lane 0 issues the bulk stores
    const bool out_bf16 = (p.epilogue == PLM_EPI_BF16 || p.epilogue == PLM_EPI_BF16_ROPE);
    const bool is_rope = (p.epilogue == PLM_EPI_BF16_ROPE) && !(PLM_DBG(p) & 1);
    const bool is_resid = (p.epilogue == PLM_EPI_RESID_F32) && !(PLM_DBG(p) & 1);
    const int piece = lane & 7;                // coalesced layout: 16-byte piece of a row's 128-byte segment
    const int lrow0 = q * 32 + (lane >> 3);    // ... of rows lrow0 + 4j, j = 0..7
    const uint32_t own_off = r_tile * 128;
    const int own_sw = r_tile & 7;
    int it = 0;
    uint32_t chunk_no = 0;
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
      fetch_aux(nxt, 0);
      mbar_wait(&tfull[a], aph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * BN;
      const int r0 = static_cast<int>(row_base);
      if (p.epilogue == PLM_EPI_BF16_SWIGLU) {
        // accumulator columns [0,128) = gate a, [128,256) = up z.  Per 64-column group: write a and z to their places in
        // u, then h = silu(a) * z (from the bf16-rounded values, as the stand-alone kernel computes it) to C2.
        auto emit = [&](const uint32_t(&o)[32], const CUtensorMap* map, int c0) {
          uint8_t* buf = sEpi + (chunk_no & 1) * EPI_BUF_BYTES;
          if (elected) bulk_wait_group_read<1>();
          named_bar_sync(1, GEMM_EPI_WARPS * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<uint4*>(buf + own_off + ((i ^ own_sw) << 4)) =
                make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
          fence_proxy_async_smem();
          named_bar_sync(1, GEMM_EPI_WARPS * 32);
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
          if (j == BN / 128 - 1) {  // last read of this accumulator: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR)
                mbar_arrive_cluster(&tempty[a], 0);  // the leader's MMA thread owns the pair's accumulators
              else
                mbar_arrive(&tempty[a]);
            }
          }
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
      } else if (out_bf16) {
        const int n_chunks = (n_sub + 1) >> 1;  // 64 bf16 columns per staging row
#pragma unroll 1
        for (int c = 0; c < n_chunks; ++c) {
          uint8_t* buf = sEpi + (chunk_no & 1) * EPI_BUF_BYTES;
          if (elected) bulk_wait_group_read<1>();  // the store that last read this staging tile has drained it
          named_bar_sync(1, GEMM_EPI_WARPS * 32);
          uint32_t o[32];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int sc = 2 * c + h;
            if (sc < n_sub) {  // warp-uniform
              const bool rope_on = is_rope && (tile_col0 + sc * 32 < p.rope_cols);
              if (rope_on) {
                stage_aux(buf, nxt);
                __syncwarp();
              }
              if (sc + 1 < n_sub) fetch_aux(nxt, sc + 1);
              uint32_t r[32];
              tmem_ld32(t_row + sc * 32, r);
              tmem_ld_wait();
              if (rope_on) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 cs = *reinterpret_cast<const float4*>(buf + own_off + ((i ^ own_sw) << 4));
                  const float a0 = __uint_as_float(r[4 * i + 0]), b0 = __uint_as_float(r[4 * i + 1]);
                  const float a1 = __uint_as_float(r[4 * i + 2]), b1 = __uint_as_float(r[4 * i + 3]);
                  r[4 * i + 0] = __float_as_uint(a0 * cs.x - b0 * cs.y);
                  r[4 * i + 1] = __float_as_uint(b0 * cs.x + a0 * cs.y);
                  r[4 * i + 2] = __float_as_uint(a1 * cs.z - b1 * cs.w);
                  r[4 * i + 3] = __float_as_uint(b1 * cs.z + a1 * cs.w);
                }
                __syncwarp();  // every lane has read its (cos,sin) row before the tile is overwritten
              }
#pragma unroll
              for (int i = 0; i < 16; ++i)
                o[16 * h + i] = pack_bf16x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
            }
          }
          if (c == n_chunks - 1) {  // last read of this accumulator: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR)
                mbar_arrive_cluster(&tempty[a], 0);  // the leader's MMA thread owns the pair's accumulators
              else
                mbar_arrive(&tempty[a]);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<uint4*>(buf + own_off + ((i ^ own_sw) << 4)) =
                make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
          fence_proxy_async_smem();
          named_bar_sync(1, GEMM_EPI_WARPS * 32);
          if (elected && !(PLM_DBG(p) & 2)) {
            tma_store_2d(&tmC, buf, static_cast<int>(tile_col0) + c * 64, r0);
            bulk_commit_group();
          }
          ++chunk_no;
        }
      } else {
#pragma unroll 1
        for (int sc = 0; sc < n_sub; ++sc) {  // 32 fp32 columns per staging row
          uint8_t* buf = sEpi + (chunk_no & 1) * EPI_BUF_BYTES;
          if (elected) bulk_wait_group_read<1>();
          named_bar_sync(1, GEMM_EPI_WARPS * 32);
          if (is_resid) {
            stage_aux(buf, nxt);
            __syncwarp();
          }
          if (sc + 1 < n_sub) fetch_aux(nxt, sc + 1);
          uint32_t r[32];
          tmem_ld32(t_row + sc * 32, r);
          tmem_ld_wait();
          if (sc == n_sub - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR)
                mbar_arrive_cluster(&tempty[a], 0);  // the leader's MMA thread owns the pair's accumulators
              else
                mbar_arrive(&tempty[a]);
            }
          }
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
          named_bar_sync(1, GEMM_EPI_WARPS * 32);
          if (elected && !(PLM_DBG(p) & 2)) {
            const int c0 = static_cast<int>(tile_col0) + sc * 32;
            if (p.epilogue == PLM_EPI_ATOMIC_F32)
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
  if (CL == 2) cluster_sync_all();  // no CTA may exit while its peer can still multicast into it
  if (warp == 1) {
    tc_fence_after();
    if (PAIR)
      tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
    else
      tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BN, bool A_K, bool B_K, int CL, bool PAIR = false>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmC2,
                       const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, PAIR>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_kernel<BN, A_K, B_K, CL, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return fail(PLM_ERR_CUDA, "gemm smem attribute: %s", cudaGetErrorString(attr_err));
  const int total = ((p.num_m + CL - 1) / CL) * p.num_n * p.splits;
  int clusters = sm_count() / CL;
  if (total < clusters) clusters = total;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_kernel<BN, A_K, B_K, CL, PAIR>, tmA, tmB, tmC, tmC2, p);
  if (e != cudaSuccess) return fail(PLM_ERR_CUDA, "gemm_kernel launch: %s", cudaGetErrorString(e));
  return check_launch("gemm_kernel");
}

// Tuning overrides for A/B measurements (tools/gpu_kernel_check.py, tests): process-global, set through
// plm_gemm_set_tuning(), never read from the environment on the launch path.  Defaults = automatic choices.
struct GemmTuning {
  int bn = 0;        // 128 / 256 forces the tile width
  int raster = -1;   // 0 / 1 forces the tile walk order
  int cluster = 0;   // 1 / 2 forces the cluster size
  int pair = 1;      // 0: two cta_group::1 MMAs sharing a multicast B tile instead of the CTA-pair MMA
  int debug = 0;     // only honoured by -DPLM_GEMM_DEBUG builds
};
static GemmTuning g_gemm_tuning;
static const GemmTuning& gemm_env() { return g_gemm_tuning; }

}  // namespace plm

extern "C" int plm_gemm_set_tuning(int32_t bn, int32_t raster, int32_t cluster, int32_t pair, int32_t debug) {
  if (!(bn == 0 || bn == 128 || bn == 256) || raster < -1 || raster > 1 || cluster < 0 || cluster > 2)
    return plm::fail(PLM_ERR_INVALID, "gemm_set_tuning: bad value");
  plm::g_gemm_tuning.bn = bn;
  plm::g_gemm_tuning.raster = raster;
  plm::g_gemm_tuning.cluster = cluster;
  plm::g_gemm_tuning.pair = pair;
  plm::g_gemm_tuning.debug = debug;
  return PLM_OK;
}

extern "C" int plm_gemm_bf16(const plm_gemm_args* a, plm_stream_t stream_) {
  using namespace plm;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PLM_ENSURE_CONTEXT(a ? a->A : nullptr);
  PLM_REQUIRE(a != nullptr, "gemm: null args");
  PLM_REQUIRE(a->A && a->B && a->C, "gemm: null operand");
  PLM_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "gemm: non-positive size M=%lld N=%lld K=%lld", (long long)a->M,
              (long long)a->N, (long long)a->K);
  PLM_REQUIRE(a->M < (1ll << 31) && a->N < (1ll << 31) && a->K < (1ll << 31), "gemm: size too large");
  PLM_REQUIRE(aligned16(a->A) && aligned16(a->B) && aligned16(a->C), "gemm: operands must be 16-byte aligned");
  PLM_REQUIRE(a->N % 8 == 0 && a->lda % 8 == 0 && a->ldb % 8 == 0 && a->ldc % 8 == 0,
              "gemm: N and leading dimensions must be multiples of 8");
  PLM_REQUIRE(a->epilogue >= PLM_EPI_BF16 && a->epilogue <= PLM_EPI_BF16_SWIGLU, "gemm: bad epilogue %d", a->epilogue);
  const bool glu = a->epilogue == PLM_EPI_BF16_SWIGLU;
  if (glu) {
    PLM_REQUIRE(a->b_kmajor != 0 && a->N % 256 == 0, "gemm: the SwiGLU epilogue needs a K-major B and N/2 %% 128 == 0");
    PLM_REQUIRE(a->C2 && aligned16(a->C2) && a->ldc2 % 8 == 0 && a->ldc2 >= a->N / 2,
                "gemm: SwiGLU output C2 missing/misaligned");
  }
  if (a->epilogue == PLM_EPI_RESID_F32) PLM_REQUIRE(a->R && aligned16(a->R), "gemm: residual pointer missing/misaligned");
  if (a->epilogue == PLM_EPI_BF16_ROPE) {
    PLM_REQUIRE(a->rope_table && aligned16(a->rope_table), "gemm: rope table missing/misaligned");
    PLM_REQUIRE(a->head_dim >= 32 && a->head_dim % 32 == 0 && a->rope_T > 0 && a->rope_cols % a->head_dim == 0,
                "gemm: bad rope geometry");
  }
  const bool a_k = a->a_kmajor != 0, b_k = a->b_kmajor != 0;
  PLM_REQUIRE(a->lda >= (a_k ? a->K : a->M) && a->ldb >= (b_k ? a->K : a->N) && a->ldc >= a->N,
              "gemm: leading dimension too small");

  GemmParams p;
  p.C = a->C;
  p.R = a->R;
  p.rope = a->rope_table;
  p.M = a->M;
  p.N = a->N;
  p.K = a->K;
  p.ldc = a->ldc;
  p.epilogue = a->epilogue;
  p.rope_cols = a->rope_cols;
  p.rope_T = a->rope_T;
  p.head_dim = a->head_dim;
  p.glu_F = glu ? static_cast<int>(a->N / 2) : 0;
  p.debug = gemm_env().debug;
  p.kblocks = static_cast<int>((a->K + BK - 1) / BK);
  p.num_m = static_cast<int>((a->M + BM - 1) / BM);

  // Tile width: 128x256 tiles feed the tensor core best (one 128x256x16 MMA reads 12 KB of smem per 128 cycles;
  // a 128x128x16 MMA reads 8 KB per 64 cycles, which is the smem bandwidth limit).  BN=128 only for narrow N.
  const int sms = sm_count();
  int bn = a->N <= 128 ? 128 : 256;
  // short-K GEMMs with the fp32 residual epilogue are bound by their epilogue (8 B/element in, 4 B out per 2K flop):
  // narrower tiles give twice as many epilogue streams per wave and finer wave quantisation (measured 64 -> 55 us at
  // 16384 x 1024 x 1024)
  if (a->epilogue == PLM_EPI_RESID_F32 && a->K <= 1024) bn = 128;
  {
    const int forced = gemm_env().bn;
    if (forced == 128 || forced == 256) bn = forced;
  }
  if (glu) bn = 256;  // 128 gate + 128 up columns per tile
  // Rasterisation: the ~148 tiles in flight should share the operand that does NOT fit in L2.  Walking M keeps one
  // B tile hot and streams A once per N-block (fine when A fits in L2); walking N reads each A row-block once.
  {
    const double a_bytes = 2.0 * static_cast<double>(a->M) * static_cast<double>(a->K);
    const double b_bytes = 2.0 * static_cast<double>(a->N) * static_cast<double>(a->K);
    p.n_fastest = (b_bytes < a_bytes) ? 1 : 0;  // measured with CTA pairs: walking N wins whenever the weights are the
                                                // smaller operand (fc1 151 -> 147 us, out-proj 55 -> 52 us), even if A fits in L2
    const int forced = gemm_env().raster;
    if (forced == 0 || forced == 1) p.n_fastest = forced;
  }
  p.num_n = static_cast<int>((a->N + bn - 1) / bn);
  // 2-CTA clusters (M-adjacent tiles sharing a multicast B tile) whenever there is more than one M block
  int cl = (p.num_m >= 2 && bn == 256) ? 2 : 1;
  {
    const int forced = gemm_env().cluster;
    if (forced == 1 || (forced == 2 && bn == 256)) cl = forced;
  }

  int splits = a->splits;
  if (a->epilogue != PLM_EPI_ATOMIC_F32) {
    PLM_REQUIRE(splits <= 1, "gemm: split-K needs the atomic epilogue");
    splits = 1;
  } else if (splits <= 0) {
    // Pick the split count that fills whole waves of the persistent grid: efficiency = items / (waves * SMs), with a
    // small penalty per split (each split adds one fp32 reduce-add pass over the output tile).
    const int tiles = ((p.num_m + cl - 1) / cl) * p.num_n * cl;
    const int max_splits = p.kblocks / 8 > 0 ? (p.kblocks / 8 < 32 ? p.kblocks / 8 : 32) : 1;  // >= 8 k-blocks each
    double best = -1.0;
    splits = 1;
    for (int s = 1; s <= max_splits; ++s) {
      const int items = tiles * s;
      const int waves = (items + sms - 1) / sms;
      const double eff = static_cast<double>(items) / (static_cast<double>(waves) * sms) - 0.004 * (s - 1);
      if (eff > best + 1e-9) {
        best = eff;
        splits = s;
      }
    }
  }
  if (splits > p.kblocks) splits = p.kblocks;
  {
    const int per = (p.kblocks + splits - 1) / splits;
    splits = (p.kblocks + per - 1) / per;  // no empty split
  }
  p.splits = splits;

  CUtensorMap tmA, tmB, tmC, tmC2;
  int rc;
  if (a->epilogue == PLM_EPI_BF16 || a->epilogue == PLM_EPI_BF16_ROPE || glu)
    rc = make_tmap_bf16_2d(&tmC, a->C, a->M, a->N, a->ldc, BM, 64);
  else
    rc = make_tmap_f32_2d(&tmC, a->C, a->M, a->N, a->ldc, BM, 32);
  if (rc != PLM_OK) return rc;
  tmC2 = tmC;
  if (glu) {
    rc = make_tmap_bf16_2d(&tmC2, a->C2, a->M, a->N / 2, a->ldc2, BM, 64);
    if (rc != PLM_OK) return rc;
  }
  if (a_k)
    rc = make_tmap_bf16_2d(&tmA, a->A, a->M, a->K, a->lda, BM, 64);
  else
    rc = make_tmap_bf16_2d(&tmA, a->A, a->K, a->M, a->lda, BK, 64);
  if (rc != PLM_OK) return rc;
  if (b_k)
    rc = make_tmap_bf16_2d(&tmB, a->B, a->N, a->K, a->ldb, glu ? bn / 2 : bn / cl, 64);  // cluster: each CTA fetches half the rows
  else
    rc = make_tmap_bf16_2d(&tmB, a->B, a->K, a->N, a->ldb, BK, 64);
  if (rc != PLM_OK) return rc;

#define PLM_DISPATCH(BN_, CL_)                                                        \
  if (a_k && b_k) return launch_gemm<BN_, true, true, CL_>(tmA, tmB, tmC, tmC2, p, stream);      \
  if (a_k && !b_k) return launch_gemm<BN_, true, false, CL_>(tmA, tmB, tmC, tmC2, p, stream);    \
  if (!a_k && b_k) return launch_gemm<BN_, false, true, CL_>(tmA, tmB, tmC, tmC2, p, stream);    \
  return launch_gemm<BN_, false, false, CL_>(tmA, tmB, tmC, tmC2, p, stream);
  // CTA-pair MMA (tcgen05.mma.cta_group::2) for every clustered launch; PLM_GEMM_PAIR=0 falls back to two
  // cta_group::1 MMAs sharing a multicast B tile
  const bool pair = cl == 2 && bn == 256 && gemm_env().pair != 0;
  if (pair) {
    if (a_k && b_k) return launch_gemm<256, true, true, 2, true>(tmA, tmB, tmC, tmC2, p, stream);
    if (a_k && !b_k) return launch_gemm<256, true, false, 2, true>(tmA, tmB, tmC, tmC2, p, stream);
    if (!a_k && b_k) return launch_gemm<256, false, true, 2, true>(tmA, tmB, tmC, tmC2, p, stream);
    return launch_gemm<256, false, false, 2, true>(tmA, tmB, tmC, tmC2, p, stream);
  }
  if (bn == 256 && cl == 2) {
    PLM_DISPATCH(256, 2)
  } else if (bn == 256) {
    PLM_DISPATCH(256, 1)
  } else {
    PLM_DISPATCH(128, 1)
  }
#undef PLM_DISPATCH
}

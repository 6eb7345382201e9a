// Host side of plm_gemm_bf16: argument checks, tile / rasterisation / split-K choices, tensor maps, and the dispatch to
// the per-epilogue kernels of gemm_kernel.cuh (instantiated in gemm_epi_*.cu).
#include "gemm_kernel.cuh"

namespace plm {

// Tuning overrides for A/B measurements (tools/gpu_kernel_check.py, tests): process-global, set through
// plm_gemm_set_tuning(), never read from the environment on the launch path.  Defaults = automatic choices.
struct GemmTuning {
  int bn = 0;        // 128 / 256 forces the tile width
  int raster = -1;   // 0 / 1 forces the tile walk order
  int cluster = 0;   // 1 forces single-CTA launches, 2 CTA pairs (where the tile width allows)
  int pair = 1;      // 0: never use the CTA-pair MMA (single-CTA tiles only)
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
  const bool ce = a->epilogue == PLM_EPI_BF16_CE;
  PLM_REQUIRE(a->A && a->B && (a->C || ce), "gemm: null operand");
  PLM_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "gemm: non-positive size M=%lld N=%lld K=%lld", (long long)a->M,
              (long long)a->N, (long long)a->K);
  PLM_REQUIRE(a->M < (1ll << 31) && a->N < (1ll << 31) && a->K < (1ll << 31), "gemm: size too large");
  PLM_REQUIRE(aligned16(a->A) && aligned16(a->B) && aligned16(a->C), "gemm: operands must be 16-byte aligned");
  PLM_REQUIRE(a->N % 8 == 0 && a->lda % 8 == 0 && a->ldb % 8 == 0 && a->ldc % 8 == 0,
              "gemm: N and leading dimensions must be multiples of 8");
  PLM_REQUIRE(a->epilogue >= PLM_EPI_BF16 && a->epilogue <= PLM_EPI_BF16_GLU_BWD, "gemm: bad epilogue %d", a->epilogue);
  const bool glu = a->epilogue == PLM_EPI_BF16_SWIGLU;
  if (glu) {
    PLM_REQUIRE(a->b_kmajor != 0 && a->N % 256 == 0, "gemm: the SwiGLU epilogue needs a K-major B and N/2 %% 128 == 0");
    PLM_REQUIRE(a->C2 && aligned16(a->C2) && a->ldc2 % 8 == 0 && a->ldc2 >= a->N / 2,
                "gemm: SwiGLU output C2 missing/misaligned");
  }
  const bool glub = a->epilogue == PLM_EPI_BF16_GLU_BWD;
  if (glub) {
    PLM_REQUIRE(a->a_kmajor != 0 && a->b_kmajor == 0 && a->N % 256 == 0,
                "gemm: the GLU-backward epilogue needs a K-major A, an MN-major B and N %% 256 == 0");
    PLM_REQUIRE(a->C2 && aligned16(a->C2) && a->ldc2 % 8 == 0 && a->ldc2 >= 2 * a->N && a->ldc >= 2 * a->N,
                "gemm: GLU-backward input u / output du missing, misaligned or narrower than 2N");
  }
  if (ce) {
    PLM_REQUIRE(a->ce_targets && a->ce_partial && a->ce_tgt_logit && aligned16(a->ce_partial),
                "gemm: the cross-entropy epilogue needs targets, a partial-statistics workspace and a target-logit buffer");
    PLM_REQUIRE(a->a_kmajor != 0 && a->b_kmajor != 0, "gemm: the cross-entropy epilogue needs K-major A and B");
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
  p.rope_cached = 0;
  p.glu_F = glu ? static_cast<int>(a->N / 2) : (glub ? static_cast<int>(a->N) : 0);
  p.debug = gemm_env().debug;
  p.ce_targets = a->ce_targets;
  p.ce_partial = reinterpret_cast<float2*>(a->ce_partial);
  p.ce_tgt_logit = a->ce_tgt_logit;
  p.ce_store = a->C != nullptr;
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
  if (glub) bn = 256;  // a / z chunk pairs are cut from 256-wide tiles
  if (ce) bn = 256;   // the partial-statistics workspace is laid out for 256-column tiles (plm_lmhead_ce_partials)
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
  // CTA pairs (M-adjacent tiles, one tcgen05.mma.cta_group::2 per K-step) whenever there is more than one M block
  int cl = (p.num_m >= 2 && bn == 256) ? 2 : 1;
  {
    const int forced = gemm_env().cluster;
    if (forced == 1 || (forced == 2 && bn == 256)) cl = forced;
    if (gemm_env().pair == 0) cl = 1;
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
  const void* c_base = a->C ? a->C : a->A;  // loss-only cross-entropy forward: the map is encoded but never stored through
  if (glub)  // du = [da | dz]: 2N columns
    rc = make_tmap_bf16_2d(&tmC, a->C, a->M, 2 * a->N, a->ldc, BM, 64);
  else if (a->epilogue == PLM_EPI_BF16 || a->epilogue == PLM_EPI_BF16_ROPE || glu || ce)
    rc = make_tmap_bf16_2d(&tmC, c_base, a->M, a->N, a->C ? a->ldc : a->lda, BM, 64);
  else
    rc = make_tmap_f32_2d(&tmC, a->C, a->M, a->N, a->ldc, BM, 32);
  if (rc != PLM_OK) return rc;
  tmC2 = tmC;
  if (a->epilogue == PLM_EPI_BF16_ROPE && a->head_dim == 64 && a->rope_T % BM == 0) {
    // the (cos,sin) rows of a 128-row block are contiguous table rows: cache them in shared memory through TMA
    // (table viewed [rope_T, 64] fp32, boxes of 128 positions x 32 floats)
    rc = make_tmap_f32_2d(&tmC2, a->rope_table, a->rope_T, 64, 64, BM, 32);
    if (rc != PLM_OK) return rc;
    p.rope_cached = 1;
  }
  if (glu) {
    rc = make_tmap_bf16_2d(&tmC2, a->C2, a->M, a->N / 2, a->ldc2, BM, 64);
    if (rc != PLM_OK) return rc;
  }
  if (glub) {  // u = [a | z], loaded chunk-wise into the staging tiles
    rc = make_tmap_bf16_2d(&tmC2, a->C2, a->M, 2 * a->N, a->ldc2, BM, 64);
    if (rc != PLM_OK) return rc;
  }
  if (a_k)
    rc = make_tmap_bf16_2d(&tmA, a->A, a->M, a->K, a->lda, BM, 64);
  else
    rc = make_tmap_bf16_2d(&tmA, a->A, a->K, a->M, a->lda, BK, 64);
  if (rc != PLM_OK) return rc;
  if (b_k)
    rc = make_tmap_bf16_2d(&tmB, a->B, a->N, a->K, a->ldb, glu ? bn / 2 : bn / cl, 64);  // pair: each CTA fetches half the rows
  else
    rc = make_tmap_bf16_2d(&tmB, a->B, a->K, a->N, a->ldb, BK, 64);
  if (rc != PLM_OK) return rc;

  const GemmTile tile = {bn, cl == 2};
  switch (a->epilogue) {
    case PLM_EPI_BF16:
      return gemm_launch_epi<PLM_EPI_BF16>(a_k, b_k, tile, tmA, tmB, tmC, tmC2, p, stream);
    case PLM_EPI_BF16_ROPE:
      return gemm_launch_epi<PLM_EPI_BF16_ROPE>(a_k, b_k, tile, tmA, tmB, tmC, tmC2, p, stream);
    case PLM_EPI_F32:
    case PLM_EPI_ATOMIC_F32:  // same kernel: the bulk store becomes a bulk reduce-add
      return gemm_launch_epi<PLM_EPI_F32>(a_k, b_k, tile, tmA, tmB, tmC, tmC2, p, stream);
    case PLM_EPI_RESID_F32:
      return gemm_launch_epi<PLM_EPI_RESID_F32>(a_k, b_k, tile, tmA, tmB, tmC, tmC2, p, stream);
    case PLM_EPI_BF16_SWIGLU:
      return gemm_launch_epi<PLM_EPI_BF16_SWIGLU>(a_k, b_k, tile, tmA, tmB, tmC, tmC2, p, stream);
    case PLM_EPI_BF16_GLU_BWD:
      return gemm_launch_epi<PLM_EPI_BF16_GLU_BWD>(a_k, b_k, tile, tmA, tmB, tmC, tmC2, p, stream);
    default:
      return gemm_launch_epi<PLM_EPI_BF16_CE>(a_k, b_k, tile, tmA, tmB, tmC, tmC2, p, stream);
  }
}

/* Number of 256-column tiles plm_gemm_bf16 cuts N into for PLM_EPI_BF16_CE = rows of the partial-statistics workspace. */
extern "C" int plm_lmhead_ce_tiles(int64_t N) { return static_cast<int>((N + 255) / 256); }

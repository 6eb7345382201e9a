/* plainlm_b200.h — C ABI of libplainlm_b200.so
 *
 * B200 (sm_100a) kernels for ONE hot path of Niccolo-Ajroldi/plainLM: the data-parallel transformer training
 * step (engine/engine.py:93-141 -> models/transformer.py:108-114 -> optim).  The reference reaches all of this
 * arithmetic through PyTorch library calls; each entry point below names the reference call site it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer owned by the caller (workspaces included);
 *    nothing here allocates, frees or synchronises.  `stream` is a cudaStream_t passed as void*.
 *  - every function enqueues on `stream` and returns PLM_OK (0) or a negative plm_status; the message of the last
 *    failure on the calling thread is returned by plm_last_error().  There is NO fallback path: unsupported shapes
 *    return PLM_ERR_UNSUPPORTED.
 *  - functions are stateless and re-entrant (callable from the autograd worker thread).
 *  - "bf16" buffers are raw uint16 bfloat16 bit patterns; row-major everywhere; `ld*` are leading dimensions in
 *    ELEMENTS.  All pointers must be 16-byte aligned and all leading dimensions multiples of 8 elements.
 */
#ifndef PLAINLM_B200_H_
#define PLAINLM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLM_OK 0
#define PLM_ERR_INVALID (-1)     /* bad argument (null pointer, misalignment, negative size)            */
#define PLM_ERR_CUDA (-2)        /* CUDA runtime / driver error while encoding a descriptor or launching */
#define PLM_ERR_UNSUPPORTED (-3) /* shape outside what the sm_100a kernels implement                     */

#define PLM_ABI_VERSION 5

typedef void* plm_stream_t;

int plm_abi_version(void);
const char* plm_last_error(void);
/* PLM_OK iff the current CUDA device is compute capability 10.x (B200). */
int plm_device_check(void);

/* ------------------------------------------------------------------------------------------------ GEMM
 * C[M,N] = op(A) * op(B)^T-style contraction on tcgen05 tensor cores (TMA-fed, TMEM accumulators, fp32 accumulate).
 * Replaces every nn.Linear forward/backward GEMM: models/transformer.py:42,67,114, models/components.py:55-56
 * (cuBLASLt under autocast in the reference).
 *
 *   a_kmajor = 1: A is stored [M, K] (K contiguous, lda >= K)      a_kmajor = 0: A is stored [K, M] (M contiguous)
 *   b_kmajor = 1: B is stored [N, K] (K contiguous, ldb >= K)      b_kmajor = 0: B is stored [K, N] (N contiguous)
 *       forward  y = x W^T   : a_kmajor=1 (x[M,K])     b_kmajor=1 (W[N,K])
 *       dgrad    dx = dy W   : a_kmajor=1 (dy[M,N'])   b_kmajor=0 (W[N',K'] read as [K=N', N=K'])
 *       wgrad    dW = dy^T x : a_kmajor=0 (dy[Mtok,N'])b_kmajor=0 (x[Mtok,K'])   (contraction over tokens)
 *
 * epilogue:
 *   PLM_EPI_BF16        C (bf16)  = acc
 *   PLM_EPI_BF16_ROPE   C (bf16)  = acc, columns [0, rope_cols) rotated pairwise by the RoPE table
 *                       (models/embeddings.py:15-30; table = precompute_freqs_cis(...)[0] viewed [rope_T, head_dim/2, 2]
 *                       fp32 (cos, sin); position of row r is r % rope_T; pair index of column c is (c % head_dim)/2)
 *   PLM_EPI_F32         C (fp32)  = acc
 *   PLM_EPI_RESID_F32   C (fp32)  = R (fp32, same ld as C) + acc         (models/transformer.py:81-82 residual add)
 *   PLM_EPI_ATOMIC_F32  C (fp32) += acc through TMA bulk reduce-adds performed at L2 (split-K capable; fp32 .grad
 *                       accumulation across micro-steps)
 *   PLM_EPI_BF16_SWIGLU C (bf16)  = acc  AND  C2 (bf16 [M, N/2], ld = ldc2) = silu(C[:, :N/2]) * C[:, N/2:]
 *                       (models/components.py:55-56: the GLU gate applied to fc1's output u = [a | z] while the tile is
 *                       still on chip; silu is evaluated on the bf16-rounded a, z exactly like plm_swiglu_fwd).
 *                       Needs b_kmajor = 1 and (N/2) % 128 == 0.
 *   PLM_EPI_BF16_CE     C (bf16)  = acc (skipped when C == NULL)  AND  the cross-entropy statistics of the bf16-ROUNDED
 *                       logits: ce_partial[tile, row] = base-2 (max, sum 2^(x log2e - max)) over the tile's valid columns,
 *                       ce_tgt_logit[row] = C[row, ce_targets[row]] (rows whose target is ignored are left untouched).
 *                       The LM head of models/transformer.py:114 fused with the forward half of
 *                       engine/engine.py:110-112; normally reached through plm_lmhead_ce_fwd.  K-major A and B only.
 *   PLM_EPI_BF16_GLU_BWD  fc2's input-gradient GEMM fused with the GLU backward (models/components.py:55-56, autograd of
 *                       silu(a) * z): acc = dg [M, N = F]; C (bf16 [M, 2F], ld = ldc) = du = [dg z (s + a s (1 - s)) | dg a s]
 *                       with s = sigmoid(a), a = C2[:, :F], z = C2[:, F:], C2 = the saved fc1 output u (bf16 [M, 2F],
 *                       ld = ldc2).  dg is never written.  Needs a_kmajor = 1, b_kmajor = 0 and N %% 256 == 0.
 * The fused forward epilogues (ROPE, SWIGLU, CE) exist for a_kmajor = b_kmajor = 1 only (PLM_ERR_UNSUPPORTED otherwise).
 * splits > 1 partitions K and is only legal with PLM_EPI_ATOMIC_F32.  splits <= 0 lets the library choose.
 */
#define PLM_EPI_BF16 0
#define PLM_EPI_BF16_ROPE 1
#define PLM_EPI_F32 2
#define PLM_EPI_RESID_F32 3
#define PLM_EPI_ATOMIC_F32 4
#define PLM_EPI_BF16_SWIGLU 5
#define PLM_EPI_BF16_CE 6
#define PLM_EPI_BF16_GLU_BWD 7

typedef struct plm_gemm_args {
  const void* A; /* bf16 */
  const void* B; /* bf16 */
  void* C;       /* bf16 or fp32 depending on epilogue */
  const float* R;          /* residual, PLM_EPI_RESID_F32 only */
  const float* rope_table; /* PLM_EPI_BF16_ROPE only */
  int64_t M, N, K;
  int64_t lda, ldb, ldc;
  int32_t a_kmajor, b_kmajor;
  int32_t epilogue;
  int32_t splits;
  int32_t rope_cols, rope_T, head_dim;
  void* C2;     /* PLM_EPI_BF16_SWIGLU: output bf16 [M, N/2]; PLM_EPI_BF16_GLU_BWD: INPUT u, bf16 [M, 2N] */
  int64_t ldc2; /* leading dimension of C2 (elements) */
  const int64_t* ce_targets; /* PLM_EPI_BF16_CE: int64 [M] */
  float* ce_partial;         /* PLM_EPI_BF16_CE: fp32 [plm_lmhead_ce_tiles(N), M, 2] */
  float* ce_tgt_logit;       /* PLM_EPI_BF16_CE: fp32 [M] */
} plm_gemm_args;

int plm_gemm_bf16(const plm_gemm_args* args, plm_stream_t stream);
/* Diagnostics / A-B measurement only: process-global overrides of the automatic tile choices of plm_gemm_bf16
 * (bn: 0 auto | 128 | 256; raster: -1 auto | 0 walk M | 1 walk N; cluster: 0 auto | 1 | 2; pair: 1 = CTA-pair MMA when
 * clustered, 0 = two cta_group::1 MMAs; debug: timing-experiment mask, only honoured by -DPLM_GEMM_DEBUG builds).
 * plm_gemm_set_tuning(0, -1, 0, 1, 0) restores the defaults.  Nothing on the launch path reads the environment. */
int plm_gemm_set_tuning(int32_t bn, int32_t raster, int32_t cluster, int32_t pair, int32_t debug);

/* ------------------------------------------------------------------------------------------------ attention
 * Causal / document-masked flash attention (models/transformer.py:53-63, F.scaled_dot_product_attention) on tcgen05.
 * qkv: bf16 [B*T, 3*H*hd], columns = q | k | v, each [H, hd] (the reference's w_qkv split, transformer.py:42-45),
 *      q and k ALREADY rotated (RoPE is applied by the QKV GEMM epilogue).
 * seg_start: NULL for plain causal; else int32 [B*T]: first position (within the sequence) of the document that
 *      position t belongs to.  allowed(i, j) <=> seg_start[i] <= j <= i — identical to
 *      data/datasets/data_prep_utils.py:7-23 + engine/engine.py:19-23 (block-diagonal causal mask).
 * out: bf16 [B*T, H*hd] (already in the [B,T,H*hd] layout w_out consumes: no transpose/contiguous copy).
 * lse: fp32 [B, H, T] natural-log logsumexp of the scaled scores.
 * Supported: hd == 64; any T >= 1 (ragged last tile).
 */
int plm_attn_fwd(const void* qkv, const int32_t* seg_start, void* out, float* lse, int32_t B, int32_t T, int32_t H,
                 int32_t hd, plm_stream_t stream);
/* Diagnostics / A-B measurement only (tools/gpu_kernel_check.py): the same forward with an explicit kernel variant
 * (number of score pairs out of every 4 whose exp2 runs as an FMA-pipe polynomial instead of MUFU.EX2: 0..2; < 0 = the
 * default plm_attn_fwd uses), and the round-1 kernel (one query tile per CTA) kept as the baseline to beat. */
int plm_attn_fwd_variant(const void* qkv, const int32_t* seg_start, void* out, float* lse, int32_t B, int32_t T,
                         int32_t H, int32_t hd, int32_t variant, plm_stream_t stream);
int plm_attn_fwd_v1(const void* qkv, const int32_t* seg_start, void* out, float* lse, int32_t B, int32_t T, int32_t H,
                    int32_t hd, plm_stream_t stream);

/* Backward.  dout: bf16 [B*T, H*hd].  dqkv: bf16 [B*T, 3*H*hd] (fully overwritten); dq and dk are rotated back by the
 * inverse RoPE (transpose of models/embeddings.py:15-30) so dqkv is the gradient of the QKV GEMM's un-rotated output.
 * rope_table may be NULL (no inverse rotation).  Workspaces: delta fp32 [B,H,T]; dq_acc fp32 [B*T, H*hd]
 * (zeroed by this call). */
int plm_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const int32_t* seg_start,
                 const float* rope_table, void* dqkv, float* delta, float* dq_acc, int32_t B, int32_t T, int32_t H,
                 int32_t hd, plm_stream_t stream);

/* Diagnostics / A-B measurement only: the same backward with an explicit kernel variant (0 = one CTA per (key tile, head,
 * batch), v in 1..16 = CTAs that walk a list of (key tile, head, batch) items, v CTAs per SM (1 = persistent); < 0 = the
 * default plm_attn_bwd uses). */
int plm_attn_bwd_variant(const void* qkv, const void* out, const void* dout, const float* lse,
                         const int32_t* seg_start, const float* rope_table, void* dqkv, float* delta, float* dq_acc,
                         int32_t B, int32_t T, int32_t H, int32_t hd, int32_t variant, plm_stream_t stream);

/* Stand-alone RoPE on the q|k columns of a qkv buffer, in place (dir = +1 forward, -1 inverse). Used by tests and
 * by callers that bypass the fused GEMM epilogue. */
int plm_rope_qk(void* qkv, const float* rope_table, int64_t rows, int32_t T, int32_t H, int32_t hd, int32_t dir,
                plm_stream_t stream);

/* ------------------------------------------------------------------------------------------------ RMSNorm
 * models/components.py:16-28.  y = (x * rsqrt(mean(x^2) + eps)) * w, fp32 math, emitted as bf16 for the next GEMM.
 * rstd: fp32 [rows] saved for backward.  d % 8 == 0, d <= 8192. */
int plm_rmsnorm_fwd(const float* x, const float* w, void* y_bf16, float* rstd, int64_t rows, int32_t d, float eps,
                    plm_stream_t stream);
/* dx_out = (dx_in ? dx_in : 0) + rmsnorm_backward(dy); dx_out_bf16 (nullable) is a bf16 copy of dx_out for the next
 * dgrad/wgrad GEMMs.  dw_partial: fp32 [plm_rmsnorm_bwd_blocks(rows), d] per-block partial sums of dy * xhat. */
int plm_rmsnorm_bwd_blocks(int64_t rows);
int plm_rmsnorm_bwd(const void* dy_bf16, const float* x, const float* w, const float* rstd, const float* dx_in,
                    float* dx_out, void* dx_out_bf16, float* dw_partial, int64_t rows, int32_t d,
                    plm_stream_t stream);
/* dw[d] += sum over blocks of dw_partial (deterministic order). */
int plm_colsum_accum(const float* partial, float* dw, int32_t nblocks, int32_t d, plm_stream_t stream);
/* `batch` such reductions in one launch: partial [batch][nblocks][d] -> dw [batch][d] (both contiguous).  The runtime
 * keeps the partials of all 2L+1 norms and folds them once per micro-step (their gradients are contiguous in the flat
 * gradient buffer, in backward order). */
int plm_colsum_accum_batched(const float* partial, float* dw, int32_t nblocks, int32_t d, int32_t batch,
                             plm_stream_t stream);

/* ------------------------------------------------------------------------------------------------ SwiGLU
 * models/components.py:55-56: u = [a | z] (bf16 [rows, 2F]); h = silu(a) * z (bf16 [rows, F]). */
int plm_swiglu_fwd(const void* u, void* h, int64_t rows, int32_t F, plm_stream_t stream);
/* du = [dh * z * silu'(a) | dh * silu(a)] (bf16 [rows, 2F]). */
int plm_swiglu_bwd(const void* dh, const void* u, void* du, int64_t rows, int32_t F, plm_stream_t stream);

/* ------------------------------------------------------------------------------------------------ plain MLPs
 * models/components.py:31-40 (MLP: h = silu(u)) and :59-70 (MLPReluSquared: h = relu(u)^2); u, h, dh, du bf16 [n].
 * bwd: du = dh * act'(u). */
#define PLM_ACT_SILU 0
#define PLM_ACT_RELU2 1
int plm_act_fwd(const void* u, void* h, int64_t n, int32_t kind, plm_stream_t stream);
int plm_act_bwd(const void* dh, const void* u, void* du, int64_t n, int32_t kind, plm_stream_t stream);

/* ------------------------------------------------------------------------------------------------ embedding
 * models/transformer.py:110 nn.Embedding: x[r, :] = W[ids[r], :] (fp32 residual stream); ids are int64. */
int plm_embed_fwd(const int64_t* ids, const float* W, float* x, int64_t rows, int32_t d, int64_t vocab,
                  plm_stream_t stream);
/* dW[ids[r], :] += dx[r, :] (fp32 red.add; dW is the fp32 .grad buffer). */
int plm_embed_bwd(const int64_t* ids, const float* dx, float* dW, int64_t rows, int32_t d, int64_t vocab,
                  plm_stream_t stream);

/* ------------------------------------------------------------------------------------------------ cross-entropy
 * engine/engine.py:81,110-112: CrossEntropyLoss(mean, ignore_index=-100) on logits.view(-1, V), divided by
 * grad-accumulation steps.  logits: bf16 [rows, ldl] (first V columns valid), overwritten IN PLACE by
 * dlogits = (softmax - onehot) * grad_scale / n_valid (bf16).  row_loss: fp32 [rows] (0 for ignored rows);
 * n_valid is computed on device: stats[0] = sum of row losses, stats[1] = number of non-ignored rows,
 * stats[2] = mean loss (what the engine returns).  Two launches: statistics, then gradient. */
int plm_ce_fwd_bwd(void* logits, const int64_t* targets, float* row_loss, float* row_lse, float* stats, int64_t rows,
                   int32_t V, int64_t ldl, float grad_scale, int32_t write_grad, plm_stream_t stream);

/* LM head FUSED with the cross-entropy forward: models/transformer.py:114 (lm_head) + engine/engine.py:81,110-112.
 * One tcgen05 GEMM h[rows, d] x W[V, d]^T whose epilogue (a) rounds each logits tile to bf16 and stores it to `logits`
 * (bf16 [rows, ldl]; pass NULL for a loss-only forward — eval — and nothing of size [rows, V] is ever written), and
 * (b) reduces, per row and 256-column tile, the online-softmax statistics of the rounded logits and picks out the target
 * logit while the tile is on chip.  A finalize launch folds the plm_lmhead_ce_tiles(V) partials per row into row_lse /
 * row_loss, a third one takes the fixed-order mean (stats as for plm_ce_fwd_bwd).  The loss needs NO pass over the
 * [rows, V] logits.  Workspaces: partial fp32 [2 * plm_lmhead_ce_tiles(V) * rows] (16-byte aligned), tgt_logit fp32 [rows].
 * h: bf16 [rows, ldh]; W: bf16 [V, ldw]. */
int plm_lmhead_ce_tiles(int64_t V);
int plm_lmhead_ce_fwd(const void* h, const void* W, const int64_t* targets, void* logits, int64_t ldl, float* partial,
                      float* tgt_logit, float* row_loss, float* row_lse, float* stats, int64_t rows, int32_t d, int32_t V,
                      int64_t ldh, int64_t ldw, plm_stream_t stream);
/* Backward half: dlogits = (softmax - onehot) * grad_scale / n_valid written IN PLACE over the bf16 logits (one read +
 * one write pass), from the row_lse / stats plm_lmhead_ce_fwd (or plm_ce_fwd_bwd) produced.  The two LM-head gradient
 * GEMMs (dgrad, wgrad) then consume dlogits through plm_gemm_bf16. */
int plm_ce_grad(void* logits, const int64_t* targets, const float* row_lse, const float* stats, int64_t rows, int32_t V,
                int64_t ldl, float grad_scale, plm_stream_t stream);

/* ------------------------------------------------------------------------------------------------ optimizer path
 * out[0] (+)= sum of squares of a flat fp32 buffer: first pass of torch.nn.utils.clip_grad_norm_
 * (engine/engine.py:126-128).  Deterministic (fixed-order two-stage reduction) so every data-parallel rank derives
 * the bit-identical clip coefficient.  workspace: fp32 [PLM_SUMSQ_WORKSPACE]. */
#define PLM_SUMSQ_WORKSPACE 1024
int plm_sumsq(const float* g, int64_t n, float* workspace, float* out, int32_t accumulate, plm_stream_t stream);
/* Data-parallel tail: dst[i] = float(src_bf16[i]) * scale for the whole all-reduced wire buffer AND out[0] = sum dst[i]^2
 * in one pass (replaces one plm_cast_bf16_f32 per bucket + plm_sumsq; what DDP's bucket copy-back + clip_grad_norm_'s
 * norm pass do at engine/engine.py:104-105,126-128).  Same fixed-order reduction as plm_sumsq. */
int plm_unpack_sumsq(const void* src_bf16, float* dst, int64_t n, float scale, float* workspace, float* out,
                     plm_stream_t stream);

/* AdamW as built by optim/init_optim.py:13-21 (torch fused AdamW semantics), over a flat range:
 *   g' = g * clip,  clip = (max_norm > 0 && gnorm_sq) ? min(1, max_norm / (sqrt(*gnorm_sq) + 1e-6)) : 1
 *   p *= 1 - lr*wd;  m += (g' - m)(1-b1);  v = b2 v + (1-b2) g'^2;  p -= (lr / bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
 * bc1 = 1 - b1^t, bc2 = 1 - b2^t computed by the caller.  p_bf16 (nullable) receives the bf16 shadow of p. */
int plm_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, float bc1, float bc2, const float* gnorm_sq,
                   float max_norm, plm_stream_t stream);

/* optim/signSGD.py:21-46:  p *= 1 - lr*wd;  m = (first ? g : m) * mu + (1-damp) g';  p -= lr * sign(m). */
int plm_signsgd_step(float* p, const float* g, float* m, void* p_bf16, int64_t n, float lr, float momentum,
                     float dampening, float weight_decay, int32_t first_step, const float* gnorm_sq, float max_norm,
                     plm_stream_t stream);

/* torch.optim.NAdam(decoupled_weight_decay=True) as built at optim/init_optim.py:23-32 ("nadamw"), flat:
 *   p *= 1 - lr*wd;  m += (1-beta1)(g - m);  v = beta2 v + (1-beta2) g^2;  denom = sqrt(v / bc2) + eps;
 *   p += c_grad * g / denom;  p += c_mom * m / denom
 * with the host-side scalars of step t (torch/optim/nadam.py _single_tensor_nadam): bc2 = 1 - beta2^t,
 * mu_t = beta1 (1 - 0.5 * 0.96^(t * momentum_decay)), c_grad = -lr (1 - mu_t) / (1 - prod_{s<=t} mu_s),
 * c_mom = -lr mu_{t+1} / (1 - prod_{s<=t+1} mu_s).  Gradient clipping / bf16 shadow as in plm_adamw_step. */
int plm_nadamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, float bc2, float c_grad, float c_mom,
                    const float* gnorm_sq, float max_norm, plm_stream_t stream);

/* torch.optim.SGD(momentum, dampening, weight_decay) as built at optim/init_optim.py:34-41, flat:
 *   g += wd * p;  buf = first_step ? g : momentum * buf + (1 - dampening) * g;  p -= lr * buf
 * (momentum == 0: p -= lr * g, buf untouched).  Gradient clipping / bf16 shadow as in plm_adamw_step. */
int plm_sgd_step(float* p, const float* g, float* buf, void* p_bf16, int64_t n, float lr, float momentum,
                 float dampening, float weight_decay, int32_t first_step, const float* gnorm_sq, float max_norm,
                 plm_stream_t stream);

/* fp32 -> bf16 shadow copy of a flat buffer (initial weight cast; autocast's per-step cast in the reference). */
int plm_cast_f32_bf16(const float* src, void* dst, int64_t n, float scale, plm_stream_t stream);
/* bf16 -> fp32 with scale (gradient bucket unpack after the bf16 all-reduce; scale = 1/world). */
int plm_cast_bf16_f32(const void* src, float* dst, int64_t n, float scale, plm_stream_t stream);

/* Document segmentation on device: lengths (int32, concatenated per sequence, each sequence's lengths sum to T+1),
 * offsets int32 [B+1] into `lengths`; writes seg_start int32 [B*T].  Same meaning as plm host helper and as
 * data/datasets/data_prep_utils.py:7-23 cropped to [:T,:T] (engine/engine.py:23). */
int plm_seg_start_from_lengths(const int32_t* lengths, const int32_t* offsets, int32_t* seg_start, int32_t B,
                               int32_t T, plm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PLAINLM_B200_H_ */

"""Tensor-level wrappers over the C ABI: torch tensors in, device pointers + current CUDA stream out.

PyTorch is used only for device memory and streams; every function here ends in exactly one `plm_*` call.
All functions require CUDA tensors and raise otherwise (no CPU path).
"""

import ctypes

import torch

from . import _lib
from ._lib import GemmArgs, check


def _stream():
  return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, dtype=None, name='tensor'):
  if t is None:
    return None
  if not t.is_cuda:
    raise RuntimeError(f'plainlm_b200: {name} must be a CUDA tensor (no CPU path)')
  if dtype is not None and t.dtype != dtype:
    raise TypeError(f'plainlm_b200: {name} must be {dtype}, got {t.dtype}')
  return ctypes.c_void_p(t.data_ptr())


def _rowmajor_ld(t, name):
  if t.dim() != 2 or t.stride(1) != 1:
    raise ValueError(f'plainlm_b200: {name} must be 2-D with unit inner stride, got strides {t.stride()}')
  return t.stride(0)


bf16, f32 = torch.bfloat16, torch.float32


# ---------------------------------------------------------------------------------------------- GEMM
def gemm(a, b, out, *, a_kmajor=True, b_kmajor=True, epilogue=_lib.EPI_BF16, residual=None, rope_table=None,
         rope_cols=0, rope_T=0, head_dim=0, splits=1, out2=None):
  """out[M,N] = contraction of a and b (see include/plainlm_b200.h, plm_gemm_bf16).

  a: [M,K] if a_kmajor else [K,M];  b: [N,K] if b_kmajor else [K,N];  bf16, unit inner stride.
  """
  lib = _lib.load()
  lda, ldb, ldc = _rowmajor_ld(a, 'a'), _rowmajor_ld(b, 'b'), _rowmajor_ld(out, 'out')
  M, K = (a.shape[0], a.shape[1]) if a_kmajor else (a.shape[1], a.shape[0])
  N, Kb = (b.shape[0], b.shape[1]) if b_kmajor else (b.shape[1], b.shape[0])
  n_out = 2 * N if epilogue == _lib.EPI_BF16_GLU_BWD else N  # GLU backward writes du = [da | dz]
  if K != Kb or out.shape[0] != M or out.shape[1] != n_out:
    raise ValueError(f'plainlm_b200.gemm: shape mismatch a={tuple(a.shape)} b={tuple(b.shape)} out={tuple(out.shape)}')
  out_dtype = bf16 if epilogue in (_lib.EPI_BF16, _lib.EPI_BF16_ROPE, _lib.EPI_BF16_SWIGLU, _lib.EPI_BF16_CE,
                                   _lib.EPI_BF16_GLU_BWD) else f32
  args = GemmArgs()
  args.A, args.B, args.C = _ptr(a, bf16, 'a'), _ptr(b, bf16, 'b'), _ptr(out, out_dtype, 'out')
  if residual is not None:
    if residual.shape != out.shape or _rowmajor_ld(residual, 'residual') != ldc:
      raise ValueError('plainlm_b200.gemm: residual must match out in shape and leading dimension')
  args.R = _ptr(residual, f32, 'residual')
  args.rope_table = _ptr(rope_table, f32, 'rope_table')
  args.M, args.N, args.K = M, N, K
  args.lda, args.ldb, args.ldc = lda, ldb, ldc
  args.a_kmajor, args.b_kmajor = int(a_kmajor), int(b_kmajor)
  args.epilogue, args.splits = epilogue, splits
  args.rope_cols, args.rope_T, args.head_dim = rope_cols, rope_T, head_dim
  if epilogue == _lib.EPI_BF16_SWIGLU:
    if out2 is None or out2.shape[0] != M or out2.shape[1] * 2 != N:
      raise ValueError('plainlm_b200.gemm: the SwiGLU epilogue needs out2 of shape [M, N/2]')
    args.C2, args.ldc2 = _ptr(out2, bf16, 'out2'), _rowmajor_ld(out2, 'out2')
  if epilogue == _lib.EPI_BF16_GLU_BWD:  # out2 = the saved fc1 output u = [a | z] (an INPUT of this epilogue)
    if out2 is None or tuple(out2.shape) != (M, 2 * N):
      raise ValueError('plainlm_b200.gemm: the GLU-backward epilogue needs out2 = u of shape [M, 2N]')
    args.C2, args.ldc2 = _ptr(out2, bf16, 'out2'), _rowmajor_ld(out2, 'out2')
  check(lib.plm_gemm_bf16(ctypes.byref(args), _stream()), 'plm_gemm_bf16')
  return out


# ---------------------------------------------------------------------------------------------- attention
def attn_fwd(qkv, out, lse, B, T, H, hd, seg_start=None, variant=None):
  """variant (diagnostics): None = the library's default kernel; 0..2 = plm_attn_fwd_variant; 'v1' = the round-1 kernel."""
  lib = _lib.load()
  args = (_ptr(qkv, bf16, 'qkv'), _ptr(seg_start, torch.int32, 'seg_start'), _ptr(out, bf16, 'out'),
          _ptr(lse, f32, 'lse'), B, T, H, hd)
  if variant is None:
    check(lib.plm_attn_fwd(*args, _stream()), 'plm_attn_fwd')
  elif variant == 'v1':
    check(lib.plm_attn_fwd_v1(*args, _stream()), 'plm_attn_fwd_v1')
  else:
    check(lib.plm_attn_fwd_variant(*args, int(variant), _stream()), 'plm_attn_fwd_variant')
  return out, lse


def attn_bwd(qkv, out, dout, lse, dqkv, delta, dq_acc, B, T, H, hd, seg_start=None, rope_table=None, variant=None):
  """variant (diagnostics): None = the library's default kernel, else plm_attn_bwd_variant."""
  lib = _lib.load()
  args = (_ptr(qkv, bf16, 'qkv'), _ptr(out, bf16, 'out'), _ptr(dout, bf16, 'dout'), _ptr(lse, f32, 'lse'),
          _ptr(seg_start, torch.int32, 'seg_start'), _ptr(rope_table, f32, 'rope_table'), _ptr(dqkv, bf16, 'dqkv'),
          _ptr(delta, f32, 'delta'), _ptr(dq_acc, f32, 'dq_acc'), B, T, H, hd)
  if variant is None:
    check(lib.plm_attn_bwd(*args, _stream()), 'plm_attn_bwd')
  else:
    check(lib.plm_attn_bwd_variant(*args, int(variant), _stream()), 'plm_attn_bwd_variant')
  return dqkv


def rope_qk_(qkv, rope_table, T, H, hd, inverse=False):
  lib = _lib.load()
  rows = qkv.numel() // (3 * H * hd)
  check(lib.plm_rope_qk(_ptr(qkv, bf16, 'qkv'), _ptr(rope_table, f32, 'rope_table'), rows, T, H, hd,
                        -1 if inverse else 1, _stream()), 'plm_rope_qk')
  return qkv


# ---------------------------------------------------------------------------------------------- RMSNorm
def rmsnorm_fwd(x, w, y, rstd, eps):
  lib = _lib.load()
  d = x.shape[-1]
  rows = x.numel() // d
  check(lib.plm_rmsnorm_fwd(_ptr(x, f32, 'x'), _ptr(w, f32, 'w'), _ptr(y, bf16, 'y'), _ptr(rstd, f32, 'rstd'), rows, d,
                            float(eps), _stream()), 'plm_rmsnorm_fwd')
  return y, rstd


def rmsnorm_bwd_blocks(rows):
  return _lib.load().plm_rmsnorm_bwd_blocks(rows)


def rmsnorm_bwd(dy, x, w, rstd, dx_in, dx_out, dx_out_bf16, dw_partial):
  lib = _lib.load()
  d = x.shape[-1]
  rows = x.numel() // d
  check(lib.plm_rmsnorm_bwd(_ptr(dy, bf16, 'dy'), _ptr(x, f32, 'x'), _ptr(w, f32, 'w'), _ptr(rstd, f32, 'rstd'),
                            _ptr(dx_in, f32, 'dx_in'), _ptr(dx_out, f32, 'dx_out'),
                            _ptr(dx_out_bf16, bf16, 'dx_out_bf16'), _ptr(dw_partial, f32, 'dw_partial'), rows, d,
                            _stream()), 'plm_rmsnorm_bwd')
  return dx_out


def colsum_accum(partial, dw, nblocks):
  lib = _lib.load()
  check(lib.plm_colsum_accum(_ptr(partial, f32, 'partial'), _ptr(dw, f32, 'dw'), nblocks, dw.numel(), _stream()),
        'plm_colsum_accum')
  return dw


def colsum_accum_batched(partial, dw, nblocks, d, batch):
  lib = _lib.load()
  check(lib.plm_colsum_accum_batched(_ptr(partial, f32, 'partial'), _ptr(dw, f32, 'dw'), nblocks, d, batch, _stream()),
        'plm_colsum_accum_batched')
  return dw


# ---------------------------------------------------------------------------------------------- SwiGLU
def swiglu_fwd(u, h):
  lib = _lib.load()
  F = h.shape[-1]
  rows = h.numel() // F
  check(lib.plm_swiglu_fwd(_ptr(u, bf16, 'u'), _ptr(h, bf16, 'h'), rows, F, _stream()), 'plm_swiglu_fwd')
  return h


def act_fwd(u, h, kind):
  """h = act(u): kind = _lib.ACT_SILU (MLP) or _lib.ACT_RELU2 (MLPReluSquared); bf16, same shape."""
  lib = _lib.load()
  if u.shape != h.shape:
    raise ValueError('plainlm_b200.act_fwd: shape mismatch')
  check(lib.plm_act_fwd(_ptr(u, bf16, 'u'), _ptr(h, bf16, 'h'), u.numel(), kind, _stream()), 'plm_act_fwd')
  return h


def act_bwd(dh, u, du, kind):
  lib = _lib.load()
  if not (dh.shape == u.shape == du.shape):
    raise ValueError('plainlm_b200.act_bwd: shape mismatch')
  check(lib.plm_act_bwd(_ptr(dh, bf16, 'dh'), _ptr(u, bf16, 'u'), _ptr(du, bf16, 'du'), u.numel(), kind, _stream()),
        'plm_act_bwd')
  return du


def swiglu_bwd(dh, u, du):
  lib = _lib.load()
  F = dh.shape[-1]
  rows = dh.numel() // F
  check(lib.plm_swiglu_bwd(_ptr(dh, bf16, 'dh'), _ptr(u, bf16, 'u'), _ptr(du, bf16, 'du'), rows, F, _stream()),
        'plm_swiglu_bwd')
  return du


# ---------------------------------------------------------------------------------------------- embedding
def embed_fwd(ids, W, x):
  lib = _lib.load()
  vocab, d = W.shape
  check(lib.plm_embed_fwd(_ptr(ids, torch.int64, 'ids'), _ptr(W, f32, 'W'), _ptr(x, f32, 'x'), ids.numel(), d, vocab,
                          _stream()), 'plm_embed_fwd')
  return x


def embed_bwd(ids, dx, dW):
  lib = _lib.load()
  vocab, d = dW.shape
  check(lib.plm_embed_bwd(_ptr(ids, torch.int64, 'ids'), _ptr(dx, f32, 'dx'), _ptr(dW, f32, 'dW'), ids.numel(), d,
                          vocab, _stream()), 'plm_embed_bwd')
  return dW


# ---------------------------------------------------------------------------------------------- cross-entropy
def ce_fwd_bwd(logits, targets, row_loss, row_lse, stats, V, grad_scale=1.0, write_grad=True):
  """logits: bf16 [rows, ld>=V]; overwritten with dlogits when write_grad. stats: fp32[4] -> [sum, n_valid, mean]."""
  lib = _lib.load()
  rows = logits.shape[0]
  check(lib.plm_ce_fwd_bwd(_ptr(logits, bf16, 'logits'), _ptr(targets, torch.int64, 'targets'),
                           _ptr(row_loss, f32, 'row_loss'), _ptr(row_lse, f32, 'row_lse'), _ptr(stats, f32, 'stats'),
                           rows, V, _rowmajor_ld(logits, 'logits'), float(grad_scale), int(write_grad), _stream()),
        'plm_ce_fwd_bwd')
  return stats


def lmhead_ce_tiles(V):
  return int(_lib.load().plm_lmhead_ce_tiles(V))


def lmhead_ce_fwd(h, w, targets, logits, partial, tgt_logit, row_loss, row_lse, stats, V):
  """LM head fused with the cross-entropy forward (plm_lmhead_ce_fwd): h bf16 [rows, d] x w bf16 [V, d]^T; the loss
  statistics come out of the GEMM epilogue.  logits: bf16 [rows, ld >= V] or None (loss-only forward: nothing of size
  [rows, V] is written).  partial: fp32 [2 * lmhead_ce_tiles(V) * rows]; tgt_logit/row_loss/row_lse: fp32 [rows];
  stats: fp32[4] -> [sum, n_valid, mean]."""
  lib = _lib.load()
  rows, d = h.shape
  if w.shape[0] != V or w.shape[1] != d:
    raise ValueError(f'plainlm_b200.lmhead_ce_fwd: weight {tuple(w.shape)} does not match V={V}, d={d}')
  if partial.numel() < 2 * lmhead_ce_tiles(V) * rows or min(tgt_logit.numel(), row_loss.numel(), row_lse.numel()) < rows:
    raise ValueError('plainlm_b200.lmhead_ce_fwd: workspace too small')
  if logits is not None and (logits.shape[0] != rows or logits.shape[1] < V):
    raise ValueError('plainlm_b200.lmhead_ce_fwd: logits must be [rows, >= V]')
  check(lib.plm_lmhead_ce_fwd(_ptr(h, bf16, 'h'), _ptr(w, bf16, 'w'), _ptr(targets, torch.int64, 'targets'),
                              _ptr(logits, bf16, 'logits'), _rowmajor_ld(logits, 'logits') if logits is not None else 0,
                              _ptr(partial, f32, 'partial'), _ptr(tgt_logit, f32, 'tgt_logit'),
                              _ptr(row_loss, f32, 'row_loss'), _ptr(row_lse, f32, 'row_lse'), _ptr(stats, f32, 'stats'),
                              rows, d, V, _rowmajor_ld(h, 'h'), _rowmajor_ld(w, 'w'), _stream()), 'plm_lmhead_ce_fwd')
  return stats


def ce_grad(logits, targets, row_lse, stats, V, grad_scale=1.0):
  """logits (bf16 [rows, ld >= V]) -> dlogits = (softmax - onehot) * grad_scale / n_valid, in place."""
  lib = _lib.load()
  check(lib.plm_ce_grad(_ptr(logits, bf16, 'logits'), _ptr(targets, torch.int64, 'targets'), _ptr(row_lse, f32, 'row_lse'),
                        _ptr(stats, f32, 'stats'), logits.shape[0], V, _rowmajor_ld(logits, 'logits'), float(grad_scale),
                        _stream()), 'plm_ce_grad')
  return logits


# ---------------------------------------------------------------------------------------------- optimizer path
def sumsq(g, workspace, out, accumulate=False):
  lib = _lib.load()
  check(lib.plm_sumsq(_ptr(g, f32, 'g'), g.numel(), _ptr(workspace, f32, 'workspace'), _ptr(out, f32, 'out'),
                      int(accumulate), _stream()), 'plm_sumsq')
  return out


def unpack_sumsq(src, dst, workspace, out, scale=1.0):
  """dst (fp32) = src (bf16) * scale over the whole buffer and out[0] = ||dst||^2, one pass (plm_unpack_sumsq)."""
  lib = _lib.load()
  check(lib.plm_unpack_sumsq(_ptr(src, bf16, 'src'), _ptr(dst, f32, 'dst'), src.numel(), float(scale),
                             _ptr(workspace, f32, 'workspace'), _ptr(out, f32, 'out'), _stream()), 'plm_unpack_sumsq')
  return out


def adamw_step(p, g, m, v, p_bf16, lr, beta1, beta2, eps, weight_decay, step, gnorm_sq=None, max_norm=0.0):
  lib = _lib.load()
  bc1 = 1.0 - beta1 ** step
  bc2 = 1.0 - beta2 ** step
  check(lib.plm_adamw_step(_ptr(p, f32, 'p'), _ptr(g, f32, 'g'), _ptr(m, f32, 'm'), _ptr(v, f32, 'v'),
                           _ptr(p_bf16, bf16, 'p_bf16'), p.numel(), lr, beta1, beta2, eps, weight_decay, bc1, bc2,
                           _ptr(gnorm_sq, f32, 'gnorm_sq'), float(max_norm or 0.0), _stream()), 'plm_adamw_step')


def signsgd_step(p, g, m, p_bf16, lr, momentum, dampening, weight_decay, first_step, gnorm_sq=None, max_norm=0.0):
  lib = _lib.load()
  check(lib.plm_signsgd_step(_ptr(p, f32, 'p'), _ptr(g, f32, 'g'), _ptr(m, f32, 'm'), _ptr(p_bf16, bf16, 'p_bf16'),
                             p.numel(), lr, momentum, dampening, weight_decay, int(first_step),
                             _ptr(gnorm_sq, f32, 'gnorm_sq'), float(max_norm or 0.0), _stream()), 'plm_signsgd_step')


def nadamw_step(p, g, m, v, p_bf16, lr, beta1, beta2, eps, weight_decay, bc2, c_grad, c_mom, gnorm_sq=None,
                max_norm=0.0):
  lib = _lib.load()
  check(lib.plm_nadamw_step(_ptr(p, f32, 'p'), _ptr(g, f32, 'g'), _ptr(m, f32, 'm'), _ptr(v, f32, 'v'),
                            _ptr(p_bf16, bf16, 'p_bf16'), p.numel(), lr, beta1, beta2, eps, weight_decay, bc2, c_grad,
                            c_mom, _ptr(gnorm_sq, f32, 'gnorm_sq'), float(max_norm or 0.0), _stream()), 'plm_nadamw_step')


def sgd_step(p, g, buf, p_bf16, lr, momentum, dampening, weight_decay, first_step, gnorm_sq=None, max_norm=0.0):
  lib = _lib.load()
  check(lib.plm_sgd_step(_ptr(p, f32, 'p'), _ptr(g, f32, 'g'), _ptr(buf, f32, 'buf'), _ptr(p_bf16, bf16, 'p_bf16'),
                         p.numel(), lr, momentum, dampening, weight_decay, int(first_step),
                         _ptr(gnorm_sq, f32, 'gnorm_sq'), float(max_norm or 0.0), _stream()), 'plm_sgd_step')


def cast_f32_bf16(src, dst, scale=1.0):
  lib = _lib.load()
  check(lib.plm_cast_f32_bf16(_ptr(src, f32, 'src'), _ptr(dst, bf16, 'dst'), src.numel(), float(scale), _stream()),
        'plm_cast_f32_bf16')
  return dst


def cast_bf16_f32(src, dst, scale=1.0):
  lib = _lib.load()
  check(lib.plm_cast_bf16_f32(_ptr(src, bf16, 'src'), _ptr(dst, f32, 'dst'), src.numel(), float(scale), _stream()),
        'plm_cast_bf16_f32')
  return dst


def seg_start_from_lengths(lengths, offsets, seg_start, B, T):
  lib = _lib.load()
  check(lib.plm_seg_start_from_lengths(_ptr(lengths, torch.int32, 'lengths'), _ptr(offsets, torch.int32, 'offsets'),
                                       _ptr(seg_start, torch.int32, 'seg_start'), B, T, _stream()),
        'plm_seg_start_from_lengths')
  return seg_start


# ---------------------------------------------------------------------------------------------- instrumentation
# Launch accounting and optional per-call CUDA-event timing, used by bench.py (roofline of the dominant kernel,
# `gpu_launches`).  Counting is always on (an integer add per call); event timing only when a Profiler is installed.
KERNELS_PER_CALL = {
  'gemm': 1, 'attn_fwd': 1, 'attn_bwd': 3, 'rope_qk_': 1, 'rmsnorm_fwd': 1, 'rmsnorm_bwd': 1, 'colsum_accum': 1, 'colsum_accum_batched': 1,
  'swiglu_fwd': 1, 'swiglu_bwd': 1, 'act_fwd': 1, 'act_bwd': 1, 'embed_fwd': 1, 'embed_bwd': 1, 'ce_fwd_bwd': 3, 'lmhead_ce_fwd': 3, 'ce_grad': 1, 'sumsq': 2, 'unpack_sumsq': 2, 'adamw_step': 1, 'nadamw_step': 1, 'sgd_step': 1,
  'signsgd_step': 1, 'cast_f32_bf16': 1, 'cast_bf16_f32': 1, 'seg_start_from_lengths': 1,
}  # fmt: skip
LAUNCHES = 0


class Profiler:
  """Records (op name, tag, cuda start/end events) for every op call while installed."""

  def __init__(self, sync=False):
    self.records = []
    self.sync = sync  # drain the device before every op: each interval then holds exactly one op (plus launch latency)

  def summary(self):
    torch.cuda.synchronize()
    per = {}
    for name, tag, e0, e1 in self.records:
      per.setdefault((name, tag), []).append(e0.elapsed_time(e1))
    # total = median x calls: a host hiccup (GC, a clock-sampling subprocess) while the device queue is empty lands in
    # whichever interval happens to be open and would otherwise be charged to that kernel
    out = {}
    for key, ts in per.items():
      ts.sort()
      out[key] = [len(ts), ts[len(ts) // 2] * len(ts)]
    return out


_profiler = None


def _describe(name, args, kwargs):
  """Shape tag used to group calls of the same kernel and to compute algorithmic work."""
  if name == 'gemm':
    a, b = args[0], args[1]
    ak, bk = kwargs.get('a_kmajor', True), kwargs.get('b_kmajor', True)
    M, K = (a.shape[0], a.shape[1]) if ak else (a.shape[1], a.shape[0])
    N = b.shape[0] if bk else b.shape[1]
    return (M, N, K, int(ak), int(bk), kwargs.get('epilogue', 0))
  if name == 'lmhead_ce_fwd':
    return (args[0].shape[0], args[1].shape[0], args[0].shape[1], args[3] is not None)
  if name in ('attn_fwd', 'attn_bwd'):
    idx = 3 if name == 'attn_fwd' else 7
    return tuple(args[idx : idx + 4]) + (kwargs.get('seg_start') is not None,)
  sizes = tuple(int(t.numel()) for t in args if torch.is_tensor(t))
  return sizes[:2]


def _instrument(name, fn):
  k = KERNELS_PER_CALL[name]

  def wrapped(*args, **kwargs):
    global LAUNCHES
    LAUNCHES += k
    prof = _profiler
    if prof is None:
      return fn(*args, **kwargs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if prof.sync:
      torch.cuda.synchronize()
    e0.record()
    out = fn(*args, **kwargs)
    e1.record()
    prof.records.append((name, _describe(name, args, kwargs), e0, e1))
    return out

  wrapped.__name__ = name
  wrapped.__doc__ = fn.__doc__
  return wrapped


for _name in KERNELS_PER_CALL:
  globals()[_name] = _instrument(_name, globals()[_name])


def set_profiler(prof):
  global _profiler
  _profiler = prof

"""autograd.Function wrappers over the CUDA kernels, one per block of the reference model.

These give the stand-alone modules (RMSNorm, GLU, Attention, Transformer.forward -> logits) working forward AND
backward under torch autograd.  The fused training step (plainlm_b200.models.runtime) calls the same kernels through
a hand-scheduled forward/backward instead, so that nothing but kernel launches sits on the hot path.

Dtype flow follows what torch.autocast(bf16) does to the reference: residual stream fp32, every
nn.Linear operand/output bf16, gradients of parameters fp32.
"""

import torch

from .. import _lib, ops

bf16, f32 = torch.bfloat16, torch.float32


def _flat2d(x):
  return x.reshape(-1, x.shape[-1])


def weight_bf16(w):
  """bf16 operand for a GEMM: the persistent shadow if the owning model attached one, else a fresh cast."""
  shadow = getattr(w, '_plm_shadow', None)
  if shadow is not None:
    return shadow
  out = torch.empty(w.shape, device=w.device, dtype=bf16)
  ops.cast_f32_bf16(w.detach().contiguous(), out)
  return out


# ------------------------------------------------------------------------------------------------- RMSNorm
class _RMSNorm(torch.autograd.Function):
  @staticmethod
  def forward(ctx, x, w, eps):
    x2 = _flat2d(x).contiguous()
    y = torch.empty(x2.shape, device=x.device, dtype=bf16)
    rstd = torch.empty(x2.shape[0], device=x.device, dtype=f32)
    ops.rmsnorm_fwd(x2, w.detach(), y, rstd, eps)
    ctx.save_for_backward(x2, w, rstd)
    ctx.shape = x.shape
    return y.view(x.shape)

  @staticmethod
  def backward(ctx, dy):
    x2, w, rstd = ctx.saved_tensors
    rows, d = x2.shape
    dy2 = _flat2d(dy).to(bf16).contiguous()
    nb = ops.rmsnorm_bwd_blocks(rows)
    part = torch.empty(nb, d, device=x2.device, dtype=f32)
    dx = torch.empty_like(x2)
    ops.rmsnorm_bwd(dy2, x2, w.detach(), rstd, None, dx, None, part)
    dw = torch.zeros(d, device=x2.device, dtype=f32)
    ops.colsum_accum(part, dw, nb)
    return dx.view(ctx.shape), dw, None


def rmsnorm(x, w, eps=1e-6):
  """x fp32 [..., d] -> bf16 [..., d]  (reference: models/components.py:22-28 followed by autocast's cast)."""
  return _RMSNorm.apply(x.float(), w, eps)


# ------------------------------------------------------------------------------------------------- Linear
class _Linear(torch.autograd.Function):
  """y = x W^T with optional fused epilogue: RoPE on the first rope_cols columns, or fp32 residual add."""

  @staticmethod
  def forward(ctx, x, w, residual, rope):
    x2 = _flat2d(x).to(bf16).contiguous()
    wb = weight_bf16(w)
    M, N = x2.shape[0], w.shape[0]
    if residual is not None:
      out = torch.empty(M, N, device=x.device, dtype=f32)
      ops.gemm(x2, wb, out, epilogue=_lib.EPI_RESID_F32, residual=_flat2d(residual).contiguous())
    elif rope is not None:
      table, rope_cols, T, hd = rope
      out = torch.empty(M, N, device=x.device, dtype=bf16)
      ops.gemm(x2, wb, out, epilogue=_lib.EPI_BF16_ROPE, rope_table=table, rope_cols=rope_cols, rope_T=T, head_dim=hd)
    else:
      out = torch.empty(M, N, device=x.device, dtype=bf16)
      ops.gemm(x2, wb, out)
    ctx.save_for_backward(x2, w)
    ctx.wb = wb
    ctx.rope = rope
    ctx.has_resid = residual is not None
    ctx.x_shape = x.shape
    return out.view(*x.shape[:-1], N)

  @staticmethod
  def backward(ctx, dy):
    x2, w = ctx.saved_tensors
    dy2 = _flat2d(dy)
    d_resid = dy if ctx.has_resid else None
    dyb = dy2.to(bf16).contiguous()
    if ctx.rope is not None:  # undo the rotation: gradient w.r.t. the un-rotated GEMM output
      table, rope_cols, T, hd = ctx.rope
      H = rope_cols // (2 * hd)
      dyb = dyb.clone()
      ops.rope_qk_(dyb, table, T, H, hd, inverse=True)
    M, K = x2.shape
    N = w.shape[0]
    dx = torch.empty(M, K, device=dy.device, dtype=bf16)
    ops.gemm(dyb, ctx.wb, dx, a_kmajor=True, b_kmajor=False)  # dx = dy W
    dw = torch.zeros(N, K, device=dy.device, dtype=f32)
    ops.gemm(dyb, x2, dw, a_kmajor=False, b_kmajor=False, epilogue=_lib.EPI_ATOMIC_F32, splits=0)  # dW = dy^T x
    return dx.view(ctx.x_shape), dw, d_resid, None


def linear(x, w):
  return _Linear.apply(x, w, None, None)


def linear_residual(x, w, residual):
  """residual (fp32) + x W^T, fused in the GEMM epilogue (reference: models/transformer.py:81-82)."""
  return _Linear.apply(x, w, residual, None)


def linear_rope(x, w, table, rope_cols, T, head_dim):
  """x W^T with RoPE applied to columns [0, rope_cols) in the epilogue (reference: transformer.py:42-47)."""
  return _Linear.apply(x, w, None, (table, rope_cols, T, head_dim))


# ------------------------------------------------------------------------------------------------- SwiGLU
class _SwiGLU(torch.autograd.Function):
  @staticmethod
  def forward(ctx, u):
    u2 = _flat2d(u).to(bf16).contiguous()
    F = u2.shape[1] // 2
    h = torch.empty(u2.shape[0], F, device=u.device, dtype=bf16)
    ops.swiglu_fwd(u2, h)
    ctx.save_for_backward(u2)
    ctx.shape = u.shape
    return h.view(*u.shape[:-1], F)

  @staticmethod
  def backward(ctx, dh):
    (u2,) = ctx.saved_tensors
    dh2 = _flat2d(dh).to(bf16).contiguous()
    du = torch.empty_like(u2)
    ops.swiglu_bwd(dh2, u2, du)
    return du.view(ctx.shape)


def swiglu(u):
  """u = [a | z] -> silu(a) * z  (reference: models/components.py:55-56)."""
  return _SwiGLU.apply(u)


class _Activation(torch.autograd.Function):
  @staticmethod
  def forward(ctx, u, kind):
    u2 = _flat2d(u).to(bf16).contiguous()
    h = torch.empty_like(u2)
    ops.act_fwd(u2, h, kind)
    ctx.save_for_backward(u2)
    ctx.shape, ctx.kind = u.shape, kind
    return h.view(u.shape)

  @staticmethod
  def backward(ctx, dh):
    (u2,) = ctx.saved_tensors
    du = torch.empty_like(u2)
    ops.act_bwd(_flat2d(dh).to(bf16).contiguous(), u2, du, ctx.kind)
    return du.view(ctx.shape), None


def activation(u, kind):
  """silu(u) (MLP, components.py:40) or relu(u)^2 (MLPReluSquared, components.py:70); kind = _lib.ACT_*."""
  return _Activation.apply(u, kind)


# ------------------------------------------------------------------------------------------------- attention
class _FlashAttention(torch.autograd.Function):
  @staticmethod
  def forward(ctx, qkv, seg_start, B, T, H, hd):
    qkv2 = qkv.reshape(B * T, 3 * H * hd).contiguous()
    out = torch.empty(B * T, H * hd, device=qkv.device, dtype=bf16)
    lse = torch.empty(B, H, T, device=qkv.device, dtype=f32)
    ops.attn_fwd(qkv2, out, lse, B, T, H, hd, seg_start=seg_start)
    ctx.save_for_backward(qkv2, out, lse)
    ctx.seg = seg_start
    ctx.dims = (B, T, H, hd)
    return out.view(B, T, H * hd)

  @staticmethod
  def backward(ctx, dout):
    qkv2, out, lse = ctx.saved_tensors
    B, T, H, hd = ctx.dims
    d = H * hd
    do2 = dout.reshape(B * T, d).to(bf16).contiguous()
    dqkv = torch.empty_like(qkv2)
    delta = torch.empty(B, H, T, device=dout.device, dtype=f32)
    dq_acc = torch.empty(B * T, d, device=dout.device, dtype=f32)
    ops.attn_bwd(qkv2, out, do2, lse, dqkv, delta, dq_acc, B, T, H, hd, seg_start=ctx.seg, rope_table=None)
    return dqkv.view(B, T, 3 * d), None, None, None, None, None


def flash_attention(qkv, B, T, H, hd, seg_start=None):
  """qkv bf16 [B,T,3*H*hd] (q,k already rotated) -> [B,T,H*hd]; causal, or document-masked through seg_start."""
  return _FlashAttention.apply(qkv, seg_start, B, T, H, hd)


def seg_start_from_mask(attn_mask):
  """Dense bool mask [B,T,T] (the reference's engine.py:19-23 format) -> int32 [B*T] segment starts.
  For a block-diagonal causal mask the first allowed key of row i is the start of i's document."""
  first = attn_mask.to(torch.int8).argmax(dim=-1)
  return first.to(torch.int32).reshape(-1).contiguous()


# ------------------------------------------------------------------------------------------------- embedding
class _Embedding(torch.autograd.Function):
  @staticmethod
  def forward(ctx, ids, W):
    ids1 = ids.reshape(-1).contiguous()
    x = torch.empty(ids1.numel(), W.shape[1], device=W.device, dtype=f32)
    ops.embed_fwd(ids1, W.detach(), x)
    ctx.save_for_backward(ids1)
    ctx.wshape = W.shape
    return x.view(*ids.shape, W.shape[1])

  @staticmethod
  def backward(ctx, dx):
    (ids1,) = ctx.saved_tensors
    dW = torch.zeros(ctx.wshape, device=dx.device, dtype=f32)
    ops.embed_bwd(ids1, dx.reshape(-1, dx.shape[-1]).float().contiguous(), dW)
    return None, dW


def embedding(ids, W):
  return _Embedding.apply(ids, W)


# ------------------------------------------------------------------------------------------------- cross-entropy
class _CrossEntropy(torch.autograd.Function):
  @staticmethod
  def forward(ctx, logits, targets):
    V = logits.shape[-1]
    lg = logits.reshape(-1, V).to(bf16).clone()  # the kernel writes dlogits in place
    tg = targets.reshape(-1).contiguous()
    rows = lg.shape[0]
    row_loss = torch.empty(rows, device=lg.device, dtype=f32)
    row_lse = torch.empty(rows, device=lg.device, dtype=f32)
    stats = torch.zeros(4, device=lg.device, dtype=f32)
    ops.ce_fwd_bwd(lg, tg, row_loss, row_lse, stats, V, grad_scale=1.0, write_grad=True)
    ctx.save_for_backward(lg)
    ctx.shape = logits.shape
    return stats[2].clone()

  @staticmethod
  def backward(ctx, dloss):
    (dlg,) = ctx.saved_tensors
    return (dlg.float() * dloss).to(bf16).view(ctx.shape), None


def cross_entropy(logits, targets):
  """Mean cross-entropy over non-ignored targets (reference: engine/engine.py:81,111)."""
  return _CrossEntropy.apply(logits, targets)

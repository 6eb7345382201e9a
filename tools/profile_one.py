"""Run a few launches of one kernel family at the 420M shapes, for `ncu --set full -k regex:...` captures."""

import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from plainlm_b200 import ops, _lib  # noqa: E402

which = sys.argv[1]
dev = 'cuda'
B, T, H, hd, d, F, V = 8, 2048, 16, 64, 1024, 2816, 50280
M = B * T
bf = torch.bfloat16
if which == 'attn':
  qkv = torch.randn(M, 3 * d, device=dev).to(bf)
  out = torch.empty(M, d, device=dev, dtype=bf)
  lse = torch.empty(B, H, T, device=dev)
  dout = torch.randn(M, d, device=dev).to(bf)
  dqkv = torch.empty(M, 3 * d, device=dev, dtype=bf)
  delta = torch.empty(B, H, T, device=dev)
  dq_acc = torch.empty(M, d, device=dev)
  variants = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [None]
  for _ in range(2):
    for v in variants:
      ops.attn_fwd(qkv, out, lse, B, T, H, hd, variant=v)
    ops.attn_bwd(qkv, out, dout, lse, dqkv, delta, dq_acc, B, T, H, hd)
elif which == 'gemm':
  x = torch.randn(M, d, device=dev).to(bf)
  w1 = torch.randn(2 * F, d, device=dev).to(bf)
  u = torch.empty(M, 2 * F, device=dev, dtype=bf)
  du = torch.randn(M, 2 * F, device=dev).to(bf)
  dw = torch.zeros(2 * F, d, device=dev)
  dx = torch.empty(M, d, device=dev, dtype=bf)
  g = torch.empty(M, F, device=dev, dtype=bf)
  res = torch.randn(M, d, device=dev)
  xo = torch.empty(M, d, device=dev)
  w2 = torch.randn(d, F, device=dev).to(bf)
  for _ in range(2):
    ops.gemm(x, w1, u, epilogue=_lib.EPI_BF16_SWIGLU, out2=g)                               # fc1 fwd + GLU gate
    ops.gemm(g, w2, xo, epilogue=_lib.EPI_RESID_F32, residual=res)                          # fc2 fwd + residual
    ops.gemm(du, w1, dx, a_kmajor=True, b_kmajor=False)                                     # dgrad (K,MN)
    ops.gemm(du, x, dw, a_kmajor=False, b_kmajor=False, epilogue=_lib.EPI_ATOMIC_F32, splits=0)  # wgrad (MN,MN)
elif which == 'glu_bwd':
  dy = torch.randn(M, d, device=dev).to(bf)
  w2 = (torch.randn(d, F, device=dev) * 0.03).to(bf)
  u = torch.randn(M, 2 * F, device=dev).to(bf)
  dg = torch.empty(M, F, device=dev, dtype=bf)
  du = torch.empty(M, 2 * F, device=dev, dtype=bf)
  for _ in range(2):
    ops.gemm(dy, w2, du, a_kmajor=True, b_kmajor=False, epilogue=_lib.EPI_BF16_GLU_BWD, out2=u)  # fused
    ops.gemm(dy, w2, dg, a_kmajor=True, b_kmajor=False)                                          # the plain dgrad GEMM
    ops.swiglu_bwd(dg, u, du)                                                                    # ... and its GLU backward
elif which == 'lmhead':
  V = 50280
  h = torch.randn(M, d, device=dev).to(bf)
  w = (torch.randn(V, d, device=dev) * 0.05).to(bf)
  tg = torch.randint(0, V, (M,), device=dev)
  logits = torch.empty(M, V, device=dev, dtype=bf)
  partial = torch.empty(2 * ops.lmhead_ce_tiles(V) * M, device=dev)
  tgl, rl, rlse = (torch.empty(M, device=dev) for _ in range(3))
  stats = torch.zeros(4, device=dev)
  for _ in range(2):
    ops.lmhead_ce_fwd(h, w, tg, logits, partial, tgl, rl, rlse, stats, V)   # LM head + cross-entropy statistics
    ops.ce_grad(logits, tg, rlse, stats, V, 1.0)                             # dlogits in place
    ops.gemm(h, w, logits)                                                   # the plain LM-head GEMM, for comparison
elif which == 'norm':
  x = torch.randn(M, d, device=dev)
  w = torch.ones(d, device=dev)
  y = torch.empty(M, d, device=dev, dtype=bf)
  rstd = torch.empty(M, device=dev)
  dy = torch.randn(M, d, device=dev).to(bf)
  nb = ops.rmsnorm_bwd_blocks(M)
  part = torch.empty(nb, d, device=dev)
  dx = torch.empty(M, d, device=dev)
  dxb = torch.empty(M, d, device=dev, dtype=bf)
  for _ in range(2):
    ops.rmsnorm_fwd(x, w, y, rstd, 1e-6)
    ops.rmsnorm_bwd(dy, x, w, rstd, dx, dx, dxb, part)
torch.cuda.synchronize()

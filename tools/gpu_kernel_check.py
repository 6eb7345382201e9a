"""Bring-up diagnostics for the CUDA kernels (run on a B200 through gpurun).

Each case runs in its own subprocess under a timeout so that a hung kernel (mbarrier deadlock) cannot take the
whole call down.  References here are plain torch ops on the GPU in fp32 — this is a debugging aid, not the parity
suite (tests/ compares against the oracle).  Results go to gpurun_out/kernel_check.json.

  python tools/gpu_kernel_check.py            # run every case
  python tools/gpu_kernel_check.py --case X   # run one case in-process
"""

import argparse
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _err(name, got, ref, extra=None):
  import torch

  got, ref = got.float(), ref.float()
  diff = (got - ref).abs()
  denom = ref.abs().max().item() + 1e-12
  out = {
    'case': name,
    'max_abs': diff.max().item(),
    'ref_absmax': denom,
    'rel_to_max': diff.max().item() / denom,
    'mean_abs': diff.mean().item(),
    'nan': bool(torch.isnan(got).any().item()),
  }
  if extra:
    out.update(extra)
  return out


def _tile_report(got, ref, bm=128, bn=128, limit=6):
  """Per-tile max error, to localise descriptor / swizzle mistakes."""
  got, ref = got.float(), ref.float()
  M, N = got.shape
  rows = []
  for i in range(0, M, bm):
    for j in range(0, N, bn):
      e = (got[i : i + bm, j : j + bn] - ref[i : i + bm, j : j + bn]).abs().max().item()
      rows.append((e, i, j))
  rows.sort(reverse=True)
  return [{'err': e, 'row0': i, 'col0': j} for e, i, j in rows[:limit]]


def case_gemm(M, N, K, a_k, b_k, epi='bf16', bn=0, splits=1):
  import torch
  from plainlm_b200 import ops, _lib

  if bn:
    _lib.gemm_tuning(bn=bn)
  torch.manual_seed(0)
  dev = 'cuda'
  A = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)  # logical [M,K]
  Bm = (torch.randn(N, K, device=dev) * 0.5).to(torch.bfloat16)  # logical [N,K]
  a_st = A if a_k else A.t().contiguous()
  b_st = Bm if b_k else Bm.t().contiguous()
  ref = A.float() @ Bm.float().t()
  name = f'gemm M{M} N{N} K{K} a_k{int(a_k)} b_k{int(b_k)} {epi} bn{bn} s{splits}'
  if epi == 'bf16':
    out = torch.full((M, N), float('nan'), device=dev, dtype=torch.bfloat16)
    ops.gemm(a_st, b_st, out, a_kmajor=a_k, b_kmajor=b_k, epilogue=_lib.EPI_BF16)
  elif epi == 'f32':
    out = torch.full((M, N), float('nan'), device=dev, dtype=torch.float32)
    ops.gemm(a_st, b_st, out, a_kmajor=a_k, b_kmajor=b_k, epilogue=_lib.EPI_F32)
  elif epi == 'resid':
    R = torch.randn(M, N, device=dev)
    out = torch.full((M, N), float('nan'), device=dev, dtype=torch.float32)
    ops.gemm(a_st, b_st, out, a_kmajor=a_k, b_kmajor=b_k, epilogue=_lib.EPI_RESID_F32, residual=R)
    ref = ref + R
  elif epi == 'atomic':
    C0 = torch.randn(M, N, device=dev)
    out = C0.clone()
    ops.gemm(a_st, b_st, out, a_kmajor=a_k, b_kmajor=b_k, epilogue=_lib.EPI_ATOMIC_F32, splits=splits)
    ref = ref + C0
  elif epi == 'rope':
    hd, T = 64, 128
    tab = _rope_table(hd, T).to(dev)
    out = torch.full((M, N), float('nan'), device=dev, dtype=torch.bfloat16)
    rope_cols = (2 * N // 3) // hd * hd
    ops.gemm(a_st, b_st, out, a_kmajor=a_k, b_kmajor=b_k, epilogue=_lib.EPI_BF16_ROPE, rope_table=tab,
             rope_cols=rope_cols, rope_T=T, head_dim=hd)
    ref = _rope_ref(ref, tab, rope_cols, T, hd)
  torch.cuda.synchronize()
  res = _err(name, out, ref)
  if res['rel_to_max'] > 2e-2 or res['nan']:
    res['tiles'] = _tile_report(out, ref)
    res['sample_got'] = out[:2, :8].float().tolist()
    res['sample_ref'] = ref[:2, :8].float().tolist()
  return res


def _rope_table(hd, T, theta=500000.0):
  import torch

  inv = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
  t = torch.arange(T, dtype=torch.float32)
  fr = torch.outer(t, inv)
  return torch.stack([torch.cos(fr), torch.sin(fr)], dim=-1).contiguous()  # [T, hd/2, 2]


def _rope_ref(x, tab, rope_cols, T, hd, inverse=False):
  import torch

  rows = x.shape[0]
  pos = torch.arange(rows, device=x.device) % T
  cs = tab[pos]  # [rows, hd/2, 2]
  xr = x[:, :rope_cols].float().reshape(rows, rope_cols // hd, hd // 2, 2)
  cos, sin = cs[:, None, :, 0], cs[:, None, :, 1]
  if inverse:
    sin = -sin
  o0 = xr[..., 0] * cos - xr[..., 1] * sin
  o1 = xr[..., 1] * cos + xr[..., 0] * sin
  rot = torch.stack([o0, o1], dim=-1).reshape(rows, rope_cols)
  return torch.cat([rot, x[:, rope_cols:].float()], dim=1)


def _attn_ref(qkv, B, T, H, hd, seg=None):
  import torch

  d = H * hd
  q, k, v = qkv.float().view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4)  # [B,H,T,hd]
  s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
  i = torch.arange(T, device=qkv.device)
  mask = i[None, :] <= i[:, None]
  mask = mask[None, None].expand(B, 1, T, T)
  if seg is not None:
    sg = seg.view(B, T)
    mask = mask & (i[None, None, None, :] >= sg[:, None, :, None])
  s = s.masked_fill(~mask, float('-inf'))
  lse = torch.logsumexp(s, dim=-1)
  p = torch.softmax(s, dim=-1)
  o = (p @ v).permute(0, 2, 1, 3).reshape(B * T, d)
  return o, lse


def _make_seg(B, T, seed=0):
  import random
  import torch

  rng = random.Random(seed)
  seg = torch.zeros(B, T, dtype=torch.int32)
  for b in range(B):
    t = 0
    while t < T:
      ln = rng.randint(1, max(2, T // 2))
      seg[b, t : t + ln] = t
      t += ln
  return seg.reshape(-1)


def case_attn_fwd(B, T, H, doc=False, variant=None, scale=1.0):
  import torch
  from plainlm_b200 import ops

  hd, dev = 64, 'cuda'
  torch.manual_seed(1)
  qkv = (torch.randn(B * T, 3 * H * hd, device=dev) * scale).to(torch.bfloat16)
  seg = _make_seg(B, T).to(dev) if doc else None
  out = torch.full((B * T, H * hd), float('nan'), device=dev, dtype=torch.bfloat16)
  lse = torch.full((B, H, T), float('nan'), device=dev)
  ops.attn_fwd(qkv, out, lse, B, T, H, hd, seg_start=seg, variant=variant)
  torch.cuda.synchronize()
  o_ref, lse_ref = _attn_ref(qkv, B, T, H, hd, seg)
  r = _err(f'attn_fwd B{B} T{T} H{H} doc{int(doc)} variant {variant} scale {scale} out', out, o_ref)
  r2 = _err('lse', lse, lse_ref)
  r['lse_max_abs'] = r2['max_abs']
  r['lse_nan'] = r2['nan']
  if r['rel_to_max'] > 3e-2 or r['nan']:
    r['tiles'] = _tile_report(out, o_ref, 128, 64)
  return r


def case_attn_bwd(B, T, H, doc=False, rope=False, variant=None):
  import torch
  from plainlm_b200 import ops

  hd, dev = 64, 'cuda'
  d = H * hd
  torch.manual_seed(2)
  qkv = torch.randn(B * T, 3 * d, device=dev).to(torch.bfloat16)
  dout = (torch.randn(B * T, d, device=dev) * 0.5).to(torch.bfloat16)
  seg = _make_seg(B, T).to(dev) if doc else None
  tab = _rope_table(hd, T).to(dev) if rope else None
  # reference through autograd on the fp32 math
  qkv_f = qkv.float().requires_grad_(True)
  o_ref, lse_ref = _attn_ref(qkv_f, B, T, H, hd, seg)
  o_ref.backward(dout.float())
  dqkv_ref = qkv_f.grad
  if rope:
    dqkv_ref = _rope_ref(dqkv_ref, tab, 2 * d, T, hd, inverse=True)
  out = torch.empty(B * T, d, device=dev, dtype=torch.bfloat16)
  lse = torch.empty(B, H, T, device=dev)
  ops.attn_fwd(qkv, out, lse, B, T, H, hd, seg_start=seg)
  dqkv = torch.full((B * T, 3 * d), float('nan'), device=dev, dtype=torch.bfloat16)
  delta = torch.empty(B, H, T, device=dev)
  dq_acc = torch.empty(B * T, d, device=dev)
  ops.attn_bwd(qkv, out, dout, lse, dqkv, delta, dq_acc, B, T, H, hd, seg_start=seg, rope_table=tab, variant=variant)
  torch.cuda.synchronize()
  name = f'attn_bwd B{B} T{T} H{H} doc{int(doc)} rope{int(rope)} variant {variant}'
  r = _err(name, dqkv, dqkv_ref)
  for nm, sl in (('dq', slice(0, d)), ('dk', slice(d, 2 * d)), ('dv', slice(2 * d, 3 * d))):
    e = _err(nm, dqkv[:, sl], dqkv_ref[:, sl])
    r[nm + '_rel'] = e['rel_to_max']
    r[nm + '_nan'] = e['nan']
  return r


def case_bandwidth():
  """RMSNorm / SwiGLU / embedding / CE / optimizers / casts against torch on the GPU."""
  import torch
  import torch.nn.functional as F
  from plainlm_b200 import ops, _lib

  dev = 'cuda'
  torch.manual_seed(3)
  res = []
  rows, d = 1000, 1024
  x = torch.randn(rows, d, device=dev)
  w = torch.rand(d, device=dev) + 0.5
  y = torch.empty(rows, d, device=dev, dtype=torch.bfloat16)
  rstd = torch.empty(rows, device=dev)
  ops.rmsnorm_fwd(x, w, y, rstd, 1e-6)
  r_ref = torch.rsqrt(x.pow(2).mean(-1) + 1e-6)
  res.append(_err('rmsnorm_fwd', y, x * r_ref[:, None] * w))
  res.append(_err('rmsnorm_rstd', rstd, r_ref))
  dy = torch.randn(rows, d, device=dev).to(torch.bfloat16)
  dx_in = torch.randn(rows, d, device=dev)
  xg = x.clone().requires_grad_(True)
  wg = w.clone().requires_grad_(True)
  yy = (xg * torch.rsqrt(xg.pow(2).mean(-1, keepdim=True) + 1e-6)) * wg
  yy.backward(dy.float())
  nb = ops.rmsnorm_bwd_blocks(rows)
  dwp = torch.empty(nb, d, device=dev)
  dx = torch.empty(rows, d, device=dev)
  dxb = torch.empty(rows, d, device=dev, dtype=torch.bfloat16)
  ops.rmsnorm_bwd(dy, x, w, rstd, dx_in, dx, dxb, dwp)
  dw = torch.zeros(d, device=dev)
  ops.colsum_accum(dwp, dw, nb)
  res.append(_err('rmsnorm_bwd_dx', dx, xg.grad + dx_in))
  res.append(_err('rmsnorm_bwd_dx_bf16', dxb, xg.grad + dx_in))
  res.append(_err('rmsnorm_bwd_dw', dw, wg.grad))

  Fh = 2816
  u = torch.randn(rows, 2 * Fh, device=dev).to(torch.bfloat16)
  h = torch.empty(rows, Fh, device=dev, dtype=torch.bfloat16)
  ops.swiglu_fwd(u, h)
  uf = u.float().requires_grad_(True)
  href = F.silu(uf[:, :Fh]) * uf[:, Fh:]
  res.append(_err('swiglu_fwd', h, href))
  dh = torch.randn(rows, Fh, device=dev).to(torch.bfloat16)
  href.backward(dh.float())
  du = torch.empty(rows, 2 * Fh, device=dev, dtype=torch.bfloat16)
  ops.swiglu_bwd(dh, u, du)
  res.append(_err('swiglu_bwd', du, uf.grad))

  V = 5000
  W = torch.randn(V, d, device=dev)
  ids = torch.randint(0, V, (rows,), device=dev)
  xe = torch.empty(rows, d, device=dev)
  ops.embed_fwd(ids, W, xe)
  res.append(_err('embed_fwd', xe, W[ids]))
  dW = torch.zeros(V, d, device=dev)
  ops.embed_bwd(ids, x, dW)
  dW_ref = torch.zeros(V, d, device=dev).index_add_(0, ids, x)
  res.append(_err('embed_bwd', dW, dW_ref))

  Vc = 50280
  rows_c = 256
  logits = (torch.randn(rows_c, Vc, device=dev) * 2).to(torch.bfloat16)
  tg = torch.randint(0, Vc, (rows_c,), device=dev)
  tg[5] = -100
  lf = logits.float().requires_grad_(True)
  loss_ref = F.cross_entropy(lf, tg)
  (loss_ref * 0.25).backward()
  rl = torch.empty(rows_c, device=dev)
  rlse = torch.empty(rows_c, device=dev)
  stats = torch.zeros(4, device=dev)
  lg = logits.clone()
  ops.ce_fwd_bwd(lg, tg, rl, rlse, stats, Vc, grad_scale=0.25)
  res.append(_err('ce_loss', stats[2:3], loss_ref.detach().reshape(1)))
  res.append(_err('ce_grad', lg, lf.grad))

  n = 1_000_003
  g = torch.randn(n, device=dev)
  ws = torch.empty(_lib.SUMSQ_WORKSPACE, device=dev)
  out = torch.zeros(1, device=dev)
  ops.sumsq(g, ws, out)
  res.append(_err('sumsq', out, g.double().pow(2).sum().float().reshape(1)))

  n = 1 << 20
  p = torch.randn(n, device=dev)
  g = torch.randn(n, device=dev) * 3
  p_ref = p.clone().requires_grad_(True)
  opt = torch.optim.AdamW([p_ref], lr=3e-3, betas=(0.9, 0.95), weight_decay=0.1, eps=1e-8, fused=True)
  m = torch.zeros(n, device=dev)
  v = torch.zeros(n, device=dev)
  pb = torch.empty(n, device=dev, dtype=torch.bfloat16)
  gn = torch.zeros(1, device=dev)
  for step in (1, 2, 3):
    p_ref.grad = g.clone()
    torch.nn.utils.clip_grad_norm_([p_ref], 1.0)
    opt.step()
    ops.sumsq(g, ws, gn)
    ops.adamw_step(p, g, m, v, pb, 3e-3, 0.9, 0.95, 1e-8, 0.1, step, gnorm_sq=gn, max_norm=1.0)
  res.append(_err('adamw_p', p, p_ref.detach()))
  res.append(_err('adamw_m', m, opt.state[p_ref]['exp_avg']))
  res.append(_err('adamw_v', v, opt.state[p_ref]['exp_avg_sq']))
  res.append(_err('adamw_bf16', pb, p_ref.detach()))

  p = torch.randn(n, device=dev)
  ps = p.clone()
  ms = torch.zeros(n, device=dev)
  mref = None
  for step in (1, 2):
    ops.signsgd_step(ps, g, ms, None, 1e-3, 0.9, 0.0, 0.1, step == 1)
    p.mul_(1 - 1e-3 * 0.1)
    if mref is None:
      mref = g.clone()
    mref.mul_(0.9).add_(g, alpha=1.0)
    p.add_(torch.sign(mref), alpha=-1e-3)
  res.append(_err('signsgd_p', ps, p))
  res.append(_err('signsgd_m', ms, mref))

  src = torch.randn(12345, device=dev)
  dst = torch.empty(12345, device=dev, dtype=torch.bfloat16)
  ops.cast_f32_bf16(src, dst, 0.5)
  res.append(_err('cast_f32_bf16', dst, src * 0.5))
  back = torch.empty(12345, device=dev)
  ops.cast_bf16_f32(dst, back, 2.0)
  res.append(_err('cast_bf16_f32', back, dst.float() * 2.0))

  hd, T, H = 64, 128, 4
  tab = _rope_table(hd, T).to(dev)
  qkv = torch.randn(2 * T, 3 * H * hd, device=dev).to(torch.bfloat16)
  ref = _rope_ref(qkv, tab, 2 * H * hd, T, hd)
  ops.rope_qk_(qkv, tab, T, H, hd)
  res.append(_err('rope_qk', qkv, ref))
  torch.cuda.synchronize()
  return res


def _time(fn, iters=10):
  import torch

  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters


def case_gemm_perf():
  """Every GEMM of the 420M micro-step (fwd / dgrad / wgrad), swept over tile width and rasterisation, vs cuBLASLt."""
  import torch
  from plainlm_b200 import ops, _lib

  dev = 'cuda'
  out = []
  M = 16384
  layers = [('qkv', 3072, 1024), ('out', 1024, 1024), ('fc1', 5632, 1024), ('fc2', 1024, 2816), ('lm_head', 50280, 1024)]
  for name, n, k in layers:
    x = torch.randn(M, k, device=dev).to(torch.bfloat16)
    w = torch.randn(n, k, device=dev).to(torch.bfloat16)
    y = torch.empty(M, n, device=dev, dtype=torch.bfloat16)
    dy = torch.randn(M, n, device=dev).to(torch.bfloat16)
    dx = torch.empty(M, k, device=dev, dtype=torch.bfloat16)
    dw = torch.zeros(n, k, device=dev)
    flops = 2.0 * M * n * k
    variants = {
      'fwd': lambda: ops.gemm(x, w, y),
      'dgrad': lambda: ops.gemm(dy, w, dx, a_kmajor=True, b_kmajor=False),
      'wgrad': lambda: ops.gemm(dy, x, dw, a_kmajor=False, b_kmajor=False, epilogue=_lib.EPI_ATOMIC_F32, splits=0),
    }
    cublas = {
      'fwd': lambda: torch.matmul(x, w.t(), out=y),
      'dgrad': lambda: torch.matmul(dy, w, out=dx),
      'wgrad': lambda: torch.matmul(dy.t(), x),
    }
    for vname, fn in variants.items():
      rec = {'case': f'{name} {vname} M{M} N{n} K{k}'}
      for cl in (1, 2):
        for raster in (0, 1):
          _lib.gemm_tuning(cluster=cl, raster=raster)
          rec[f'cl{cl}_r{raster}'] = round(flops / _time(fn, 6) / 1e9, 0)
      _lib.gemm_tuning()
      rec['auto'] = round(flops / _time(fn, 6) / 1e9, 0)
      rec['cublas'] = round(flops / _time(cublas[vname], 6) / 1e9, 0)
      out.append(rec)
    del x, w, y, dy, dx, dw
  return out


def case_gemm_sustained():
  """Seconds-long back-to-back runs (power-capped regime): sustained TFLOP/s, SM clock and power, ours vs cuBLASLt."""
  import subprocess as sp
  import statistics
  import torch
  from plainlm_b200 import ops, _lib

  dev = 'cuda'
  M, n, k = 16384, 5632, 1024
  x = torch.randn(M, k, device=dev).to(torch.bfloat16)
  w = torch.randn(n, k, device=dev).to(torch.bfloat16)
  y = torch.empty(M, n, device=dev, dtype=torch.bfloat16)
  flops = 2.0 * M * n * k
  out = []
  variants = [('ours_cl2', lambda: ops.gemm(x, w, y), {'cluster': 2}),
              ('ours_cl1', lambda: ops.gemm(x, w, y), {'cluster': 1}),
              ('cublas', lambda: torch.matmul(x, w.t(), out=y), {})]
  for name, fn, env in variants:
    _lib.gemm_tuning(**env)
    for _ in range(20):
      fn()
    torch.cuda.synchronize()
    mon = sp.Popen(['nvidia-smi', '--query-gpu=clocks.sm,power.draw', '--format=csv,noheader,nounits', '-lms', '100'],
                   stdout=sp.PIPE, text=True)
    iters = 12000
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
      fn()
    e1.record()
    torch.cuda.synchronize()
    mon.terminate()
    lines = mon.communicate()[0].strip().splitlines()
    clk = [float(l.split(',')[0]) for l in lines[3:] if ',' in l]
    pw = [float(l.split(',')[1]) for l in lines[3:] if ',' in l]
    ms = e0.elapsed_time(e1) / iters
    out.append({'case': f'sustained fc1 fwd {name}', 'tflops': round(flops / ms / 1e9, 0), 'secs': round(ms * iters / 1e3, 2),
                'sm_mhz_median': statistics.median(clk) if clk else None, 'power_w_median': statistics.median(pw) if pw else None})
    _lib.gemm_tuning()
  return out


def case_gemm_epi_perf():
  """Forward GEMMs with their fused epilogues at the 420M shapes (RoPE on q|k; fp32 residual add)."""
  import torch
  from plainlm_b200 import ops, _lib

  dev = 'cuda'
  M, d, F, T = 16384, 1024, 2816, 2048
  bf = torch.bfloat16
  x = torch.randn(M, d, device=dev).to(bf)
  g = torch.randn(M, F, device=dev).to(bf)
  wqkv = torch.randn(3 * d, d, device=dev).to(bf)
  wout = torch.randn(d, d, device=dev).to(bf)
  w2 = torch.randn(d, F, device=dev).to(bf)
  w1 = torch.randn(2 * F, d, device=dev).to(bf)
  u = torch.empty(M, 2 * F, device=dev, dtype=bf)
  qkv = torch.empty(M, 3 * d, device=dev, dtype=bf)
  res = torch.randn(M, d, device=dev)
  out = torch.empty(M, d, device=dev)
  tab = _rope_table(64, T).to(dev)
  cases = [
    ('qkv + rope', lambda: ops.gemm(x, wqkv, qkv, epilogue=_lib.EPI_BF16_ROPE, rope_table=tab, rope_cols=2 * d, rope_T=T, head_dim=64), 2.0 * M * 3 * d * d),
    ('qkv plain', lambda: ops.gemm(x, wqkv, qkv), 2.0 * M * 3 * d * d),
    ('out + resid', lambda: ops.gemm(x, wout, out, epilogue=_lib.EPI_RESID_F32, residual=res), 2.0 * M * d * d),
    ('fc2 + resid', lambda: ops.gemm(g, w2, out, epilogue=_lib.EPI_RESID_F32, residual=res), 2.0 * M * d * F),
    ('fc1 + swiglu', lambda: ops.gemm(x, w1, u, epilogue=_lib.EPI_BF16_SWIGLU, out2=g), 2.0 * M * d * 2 * F),
    ('fc1 plain', lambda: ops.gemm(x, w1, u), 2.0 * M * d * 2 * F),
    ('swiglu alone', lambda: ops.swiglu_fwd(u, g), 2.0 * M * d * 2 * F),
  ]
  results = []
  for dbg in (0, 1, 2, 3, 16):  # masks other than 0 only act in a -DPLM_GEMM_DEBUG build of the library
    _lib.gemm_tuning(debug=dbg)
    for n, fn, fl in cases:
      ms = _time(fn, 10)
      results.append({'case': f'{n} dbg{dbg}', 'ms': round(ms, 4), 'tflops': round(fl / ms / 1e9, 0)})
  _lib.gemm_tuning()
  return results


def case_gemm_feed_probe():
  """Is the main loop bound by operand delivery?  Store-less runs with the A and/or B tile loads removed."""
  import torch
  from plainlm_b200 import ops, _lib

  dev = 'cuda'
  M, d, F = 16384, 1024, 2816
  bf = torch.bfloat16
  x = torch.randn(M, d, device=dev).to(bf)
  w1 = torch.randn(2 * F, d, device=dev).to(bf)
  u = torch.empty(M, 2 * F, device=dev, dtype=bf)
  du = torch.randn(M, 2 * F, device=dev).to(bf)
  dx = torch.empty(M, d, device=dev, dtype=bf)
  cases = [
    ('fc1 fwd', lambda: ops.gemm(x, w1, u), 2.0 * M * 2 * F * d),
    ('fc1 dgrad', lambda: ops.gemm(du, w1, dx, a_kmajor=True, b_kmajor=False), 2.0 * M * 2 * F * d),
  ]
  results = []
  for cl in (2, 1):
    for dbg in (0, 2, 6, 10, 14):
      _lib.gemm_tuning(cluster=cl, debug=dbg)
      for n, fn, fl in cases:
        ms = _time(fn, 10)
        results.append({'case': f'{n} cl{cl} dbg{dbg}', 'ms': round(ms, 4), 'tflops': round(fl / ms / 1e9, 0)})
  _lib.gemm_tuning()
  return results


def case_gemm_n1024_probe():
  """N = 1024 outputs make 512 tiles of 128x256 = 3.46 waves on 148 SMs: does BN = 128 (6.9 waves) do better?"""
  import torch
  from plainlm_b200 import ops, _lib

  dev = 'cuda'
  M, d, F = 16384, 1024, 2816
  bf = torch.bfloat16
  res = torch.randn(M, d, device=dev)
  outf = torch.empty(M, d, device=dev)
  outb = torch.empty(M, d, device=dev, dtype=bf)
  shapes = []
  for K, bk, name in ((1024, True, 'out fwd+resid'), (2816, True, 'fc2 fwd+resid'), (1024, False, 'out dgrad'),
                      (3072, False, 'qkv dgrad'), (5632, False, 'fc1 dgrad')):
    a = torch.randn(M, K, device=dev).to(bf)
    b = (torch.randn(d, K, device=dev) if bk else torch.randn(K, d, device=dev)).to(bf)
    if 'resid' in name:
      fn = (lambda a=a, b=b: ops.gemm(a, b, outf, epilogue=_lib.EPI_RESID_F32, residual=res))
    else:
      fn = (lambda a=a, b=b: ops.gemm(a, b, outb, a_kmajor=True, b_kmajor=False))
    shapes.append((name, fn, 2.0 * M * d * K))
  results = []
  for bn in (256, 128):
    _lib.gemm_tuning(bn=bn)
    for n, fn, fl in shapes:
      ms = _time(fn, 10)
      results.append({'case': f'{n} bn{bn}', 'ms': round(ms, 4), 'tflops': round(fl / ms / 1e9, 0)})
  _lib.gemm_tuning()
  return results


def case_bw_perf():
  """Achieved GB/s of the bandwidth kernels at the 420M shapes (algorithmic bytes / CUDA-event time)."""
  import torch
  from plainlm_b200 import ops, _lib

  dev = 'cuda'
  M, d, F, V = 16384, 1024, 2816, 50280
  bf = torch.bfloat16
  x = torch.randn(M, d, device=dev)
  w = torch.ones(d, device=dev)
  y = torch.empty(M, d, device=dev, dtype=bf)
  rstd = torch.empty(M, device=dev)
  dy = torch.randn(M, d, device=dev).to(bf)
  nb = ops.rmsnorm_bwd_blocks(M)
  part = torch.empty(nb, d, device=dev)
  dx = torch.randn(M, d, device=dev)
  dxb = torch.empty(M, d, device=dev, dtype=bf)
  dw = torch.zeros(d, device=dev)
  u = torch.randn(M, 2 * F, device=dev).to(bf)
  h = torch.empty(M, F, device=dev, dtype=bf)
  dh = torch.randn(M, F, device=dev).to(bf)
  du = torch.empty(M, 2 * F, device=dev, dtype=bf)
  logits = torch.randn(M, V, device=dev).to(bf)
  tg = torch.randint(0, V, (M,), device=dev)
  rl, rlse, stats = torch.empty(M, device=dev), torch.empty(M, device=dev), torch.zeros(4, device=dev)
  n = 411_304_960
  p = torch.randn(n, device=dev)
  g = torch.randn(n, device=dev)
  m = torch.zeros(n, device=dev)
  v = torch.zeros(n, device=dev)
  pb = torch.empty(n, device=dev, dtype=bf)
  ws = torch.empty(_lib.SUMSQ_WORKSPACE, device=dev)
  gs = torch.zeros(1, device=dev)
  ops.rmsnorm_fwd(x, w, y, rstd, 1e-6)
  cases = [
    ('rmsnorm_fwd', lambda: ops.rmsnorm_fwd(x, w, y, rstd, 1e-6), M * d * 6),
    ('rmsnorm_bwd', lambda: ops.rmsnorm_bwd(dy, x, w, rstd, dx, dx, dxb, part), M * d * 16),
    ('colsum_accum', lambda: ops.colsum_accum(part, dw, nb), nb * d * 4),
    ('swiglu_fwd', lambda: ops.swiglu_fwd(u, h), M * F * 6),
    ('swiglu_bwd', lambda: ops.swiglu_bwd(dh, u, du), M * F * 10),
    ('ce_fwd_bwd', lambda: ops.ce_fwd_bwd(logits, tg, rl, rlse, stats, V, 1.0, True), M * V * 6),
    ('sumsq', lambda: ops.sumsq(g, ws, gs), n * 4),
    ('adamw', lambda: ops.adamw_step(p, g, m, v, pb, 1e-3, 0.9, 0.95, 1e-8, 0.1, 3, gnorm_sq=gs, max_norm=1.0), n * 30),
    ('cast_f32_bf16', lambda: ops.cast_f32_bf16(p, pb), n * 6),
  ]
  out = []
  for name, fn, nbytes in cases:
    ms = _time(fn, 10)
    out.append({'case': name, 'ms': round(ms, 4), 'GBps': round(nbytes / ms / 1e6, 0), 'frac_of_6544': round(nbytes / ms / 1e6 / 6543.7, 3)})
  return out


def case_lmhead_ce(rows=300, V=1000, d=128):
  """Small fused LM-head + cross-entropy instance (also the compute-sanitizer case of that epilogue)."""
  import torch
  from plainlm_b200 import ops

  dev = 'cuda'
  g = torch.Generator().manual_seed(5)
  h = torch.randn(rows, d, generator=g).to(torch.bfloat16).to(dev)
  w = (torch.randn(V, d, generator=g) * 0.2).to(torch.bfloat16).to(dev)
  tg = torch.randint(0, V, (rows,), generator=g)
  tg[3] = -100
  tg = tg.to(dev)
  tiles = ops.lmhead_ce_tiles(V)
  logits = torch.empty(rows, V, device=dev, dtype=torch.bfloat16)
  partial = torch.empty(2 * tiles * rows, device=dev)
  tgl, rl, rlse = (torch.empty(rows, device=dev) for _ in range(3))
  stats = torch.zeros(4, device=dev)
  ops.lmhead_ce_fwd(h, w, tg, logits, partial, tgl, rl, rlse, stats, V)
  ref_logits = (h.float() @ w.float().t()).to(torch.bfloat16).float()
  ref = torch.nn.functional.cross_entropy(ref_logits, tg)
  ops.ce_grad(logits, tg, rlse, stats, V, 0.5)
  torch.cuda.synchronize()
  return {'case': f'lmhead_ce rows{rows} V{V} d{d}', 'loss': stats[2].item(), 'ref': ref.item(),
          'rel_to_max': abs(stats[2].item() - ref.item()) / ref.item(), 'nan': bool(torch.isnan(logits).any())}


def case_lmhead_perf():
  """420M LM head: unfused (GEMM + plm_ce_fwd_bwd) vs fused (plm_lmhead_ce_fwd + plm_ce_grad), and the loss-only form."""
  import torch
  from plainlm_b200 import ops

  dev = 'cuda'
  M, V, d = 16384, 50280, 1024
  h = torch.randn(M, d, device=dev).to(torch.bfloat16)
  w = (torch.randn(V, d, device=dev) * 0.05).to(torch.bfloat16)
  tg = torch.randint(0, V, (M,), device=dev)
  logits = torch.empty(M, V, device=dev, dtype=torch.bfloat16)
  tiles = ops.lmhead_ce_tiles(V)
  partial = torch.empty(2 * tiles * M, device=dev)
  tgl, rl, rlse = (torch.empty(M, device=dev) for _ in range(3))
  stats = torch.zeros(4, device=dev)
  fl = 2.0 * M * V * d

  def unfused():
    ops.gemm(h, w, logits)
    ops.ce_fwd_bwd(logits, tg, rl, rlse, stats, V, 1.0, True)

  def fused():
    ops.lmhead_ce_fwd(h, w, tg, logits, partial, tgl, rl, rlse, stats, V)
    ops.ce_grad(logits, tg, rlse, stats, V, 1.0)

  out = []
  for name, fn in (('gemm only', lambda: ops.gemm(h, w, logits)),
                   ('unfused: gemm + ce_fwd_bwd', unfused),
                   ('fused fwd only (stores logits)', lambda: ops.lmhead_ce_fwd(h, w, tg, logits, partial, tgl, rl, rlse, stats, V)),
                   ('fused: lmhead_ce_fwd + ce_grad', fused),
                   ('fused loss-only (no [M,V] store)', lambda: ops.lmhead_ce_fwd(h, w, tg, None, partial, tgl, rl, rlse, stats, V))):
    ms = _time(fn, 10)
    out.append({'case': name, 'ms': round(ms, 4), 'gemm_tflops': round(fl / ms / 1e9, 1)})
  return out


def case_glu_bwd_perf():
  """fc2 input-gradient GEMM + GLU backward at the 420M shape: unfused (GEMM -> plm_swiglu_bwd) vs fused epilogue."""
  import torch
  from plainlm_b200 import ops, _lib

  dev = 'cuda'
  M, d, F = 16384, 1024, 2816
  bf = torch.bfloat16
  dy = torch.randn(M, d, device=dev).to(bf)
  w2 = (torch.randn(d, F, device=dev) * 0.03).to(bf)
  u = torch.randn(M, 2 * F, device=dev).to(bf)
  dg = torch.empty(M, F, device=dev, dtype=bf)
  du = torch.empty(M, 2 * F, device=dev, dtype=bf)
  du2 = torch.empty(M, 2 * F, device=dev, dtype=bf)
  fl = 2.0 * M * d * F

  def unfused():
    ops.gemm(dy, w2, dg, a_kmajor=True, b_kmajor=False)
    ops.swiglu_bwd(dg, u, du)

  fused = lambda: ops.gemm(dy, w2, du2, a_kmajor=True, b_kmajor=False, epilogue=_lib.EPI_BF16_GLU_BWD, out2=u)  # noqa: E731
  unfused()
  fused()
  torch.cuda.synchronize()
  err = float((du.float() - du2.float()).abs().max() / du.float().abs().max())
  out = [{'case': 'fused vs unfused rel err', 'rel_to_max': err, 'nan': bool(torch.isnan(du2).any())}]
  for name, fn in (('dgrad gemm only', lambda: ops.gemm(dy, w2, dg, a_kmajor=True, b_kmajor=False)),
                   ('swiglu_bwd only', lambda: ops.swiglu_bwd(dg, u, du)), ('unfused', unfused), ('fused', fused)):
    ms = _time(fn, 20)
    out.append({'case': name, 'us': round(ms * 1e3, 1), 'gemm_tflops': round(fl / ms / 1e9, 0)})
  return out


def case_attn_perf():
  import torch
  from plainlm_b200 import ops

  dev = 'cuda'
  B, T, H, hd = 8, 2048, 16, 64
  d = H * hd
  qkv = torch.randn(B * T, 3 * d, device=dev).to(torch.bfloat16)
  out = torch.empty(B * T, d, device=dev, dtype=torch.bfloat16)
  lse = torch.empty(B, H, T, device=dev)
  dout = torch.randn(B * T, d, device=dev).to(torch.bfloat16)
  dqkv = torch.empty(B * T, 3 * d, device=dev, dtype=torch.bfloat16)
  delta = torch.empty(B, H, T, device=dev)
  dq_acc = torch.empty(B * T, d, device=dev)
  flops_fwd = 4.0 * B * H * T * T * hd / 2
  res = []
  for label, fn, fl in (
    ('attn_fwd', lambda: ops.attn_fwd(qkv, out, lse, B, T, H, hd), flops_fwd),
    ('attn_fwd v1 (round 1)', lambda: ops.attn_fwd(qkv, out, lse, B, T, H, hd, variant='v1'), flops_fwd),
    ('attn_fwd variant 0', lambda: ops.attn_fwd(qkv, out, lse, B, T, H, hd, variant=0), flops_fwd),
    ('attn_fwd variant 1', lambda: ops.attn_fwd(qkv, out, lse, B, T, H, hd, variant=1), flops_fwd),
    ('attn_fwd variant 2', lambda: ops.attn_fwd(qkv, out, lse, B, T, H, hd, variant=2), flops_fwd),
    ('attn_fwd variant 10', lambda: ops.attn_fwd(qkv, out, lse, B, T, H, hd, variant=10), flops_fwd),
    ('attn_fwd variant 11', lambda: ops.attn_fwd(qkv, out, lse, B, T, H, hd, variant=11), flops_fwd),
    ('attn_fwd variant 12', lambda: ops.attn_fwd(qkv, out, lse, B, T, H, hd, variant=12), flops_fwd),
    ('attn_bwd', lambda: ops.attn_bwd(qkv, out, dout, lse, dqkv, delta, dq_acc, B, T, H, hd), 2.5 * flops_fwd),
    ('attn_bwd variant 0', lambda: ops.attn_bwd(qkv, out, dout, lse, dqkv, delta, dq_acc, B, T, H, hd, variant=0), 2.5 * flops_fwd),
    ('attn_bwd variant 1', lambda: ops.attn_bwd(qkv, out, dout, lse, dqkv, delta, dq_acc, B, T, H, hd, variant=1), 2.5 * flops_fwd),
  ):
    for _ in range(3):
      fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
      fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    res.append({'case': f'{label} perf B{B} T{T} H{H}', 'ms': ms, 'tflops_causal': fl / ms / 1e9})
  # torch SDPA for comparison
  q, k, v = qkv.view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4)
  import torch.nn.functional as F

  for _ in range(3):
    F.scaled_dot_product_attention(q, k, v, is_causal=True)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(10):
    F.scaled_dot_product_attention(q, k, v, is_causal=True)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / 10
  res.append({'case': 'torch sdpa fwd', 'ms': ms, 'tflops_causal': flops_fwd / ms / 1e9})
  return res


CASES = {}
for _a_k, _b_k in ((True, True), (True, False), (False, False), (False, True)):
  tag = f'{"k" if _a_k else "m"}{"k" if _b_k else "m"}'
  CASES[f'gemm_{tag}_1tile_bn128'] = lambda a=_a_k, b=_b_k: case_gemm(128, 128, 64, a, b, 'f32', 128)
  CASES[f'gemm_{tag}_1tile_bn256'] = lambda a=_a_k, b=_b_k: case_gemm(128, 256, 128, a, b, 'f32', 256)
  CASES[f'gemm_{tag}_multi'] = lambda a=_a_k, b=_b_k: case_gemm(512, 768, 512, a, b, 'bf16', 0)
  CASES[f'gemm_{tag}_tails'] = lambda a=_a_k, b=_b_k: case_gemm(328, 264, 200, a, b, 'f32', 0)
CASES['gemm_kk_big'] = lambda: case_gemm(4096, 3072, 1024, True, True, 'bf16')
CASES['gemm_kk_vocab_tail'] = lambda: case_gemm(1024, 50280, 256, True, True, 'bf16')
CASES['gemm_resid'] = lambda: case_gemm(512, 1024, 512, True, True, 'resid')
CASES['gemm_rope'] = lambda: case_gemm(256, 768, 256, True, True, 'rope')
CASES['gemm_wgrad_atomic_split'] = lambda: case_gemm(1024, 512, 4096, False, False, 'atomic', 0, 0)
CASES['gemm_wgrad_atomic_1'] = lambda: case_gemm(512, 512, 1024, False, False, 'atomic', 0, 1)
CASES['attn_fwd_1tile'] = lambda: case_attn_fwd(1, 128, 1)
CASES['attn_fwd_2tile'] = lambda: case_attn_fwd(1, 256, 1)
CASES['attn_fwd_3tile'] = lambda: case_attn_fwd(1, 384, 2)      # odd number of query tiles: the last pair is half empty
CASES['attn_fwd_ragged'] = lambda: case_attn_fwd(2, 200, 2)     # T not a multiple of 64
CASES['attn_fwd_multi'] = lambda: case_attn_fwd(2, 512, 3)
CASES['attn_fwd_multi_v1'] = lambda: case_attn_fwd(2, 512, 3, variant=1)
CASES['attn_fwd_multi_v2'] = lambda: case_attn_fwd(2, 512, 3, variant=2)
CASES['attn_fwd_multi_old'] = lambda: case_attn_fwd(2, 512, 3, variant='v1')
CASES['attn_fwd_peaky'] = lambda: case_attn_fwd(2, 512, 2, scale=4.0)   # large logits: exercises the lazy O rescale
CASES['attn_fwd_peaky_v2'] = lambda: case_attn_fwd(2, 512, 2, scale=4.0, variant=2)
CASES['attn_fwd_doc'] = lambda: case_attn_fwd(2, 512, 2, doc=True)
CASES['attn_fwd_doc_v2'] = lambda: case_attn_fwd(2, 512, 2, doc=True, variant=2)
for _v in (10, 11, 12):
  CASES[f'attn_fwd3_1tile_v{_v}'] = lambda v=_v: case_attn_fwd(1, 128, 1, variant=v)
  CASES[f'attn_fwd3_2tile_v{_v}'] = lambda v=_v: case_attn_fwd(1, 256, 1, variant=v)
  CASES[f'attn_fwd3_4tile_v{_v}'] = lambda v=_v: case_attn_fwd(1, 512, 2, variant=v)
  CASES[f'attn_fwd3_ragged_v{_v}'] = lambda v=_v: case_attn_fwd(2, 200, 2, variant=v)
  CASES[f'attn_fwd3_multi_v{_v}'] = lambda v=_v: case_attn_fwd(2, 640, 3, variant=v)
  CASES[f'attn_fwd3_peaky_v{_v}'] = lambda v=_v: case_attn_fwd(2, 512, 2, scale=4.0, variant=v)
  CASES[f'attn_fwd3_doc_v{_v}'] = lambda v=_v: case_attn_fwd(2, 512, 2, doc=True, variant=v)
  CASES[f'attn_fwd3_many_items_v{_v}'] = lambda v=_v: case_attn_fwd(3, 1024, 40, variant=v)
  CASES[f'attn_fwd3_long_v{_v}'] = lambda v=_v: case_attn_fwd(1, 4096, 2, variant=v)
CASES['attn_fwd_many_items'] = lambda: case_attn_fwd(3, 1024, 40)  # 480 items: several rounds of the persistent schedule
CASES['attn_fwd_long'] = lambda: case_attn_fwd(1, 4096, 2)
CASES['attn_bwd_1tile'] = lambda: case_attn_bwd(1, 128, 1)
CASES['attn_bwd_2tile'] = lambda: case_attn_bwd(1, 256, 1)
CASES['attn_bwd_multi'] = lambda: case_attn_bwd(2, 512, 3)
CASES['attn_bwd_doc_rope'] = lambda: case_attn_bwd(2, 512, 2, doc=True, rope=True)
CASES['attn_bwd_multi_persist'] = lambda: case_attn_bwd(2, 512, 3, variant=1)
CASES['attn_bwd_doc_rope_persist'] = lambda: case_attn_bwd(2, 512, 2, doc=True, rope=True, variant=1)
CASES['attn_bwd_long_persist'] = lambda: case_attn_bwd(1, 2048, 2, variant=1)
CASES['attn_bwd_1tile_persist'] = lambda: case_attn_bwd(1, 128, 1, variant=1)
CASES['attn_bwd_ragged_persist'] = lambda: case_attn_bwd(2, 200, 2, variant=1)
CASES['attn_bwd_many_items_persist'] = lambda: case_attn_bwd(3, 1024, 40, doc=True, rope=True, variant=1)
CASES['bandwidth'] = case_bandwidth
CASES['gemm_perf'] = case_gemm_perf
CASES['lmhead_ce'] = case_lmhead_ce
CASES['lmhead_ce_ragged'] = lambda: case_lmhead_ce(333, 50280, 64)
CASES['lmhead_perf'] = case_lmhead_perf
CASES['glu_bwd_perf'] = case_glu_bwd_perf
CASES['attn_perf'] = case_attn_perf
CASES['bw_perf'] = case_bw_perf
CASES['gemm_epi_perf'] = case_gemm_epi_perf
CASES['gemm_feed_probe'] = case_gemm_feed_probe
CASES['gemm_n1024_probe'] = case_gemm_n1024_probe
CASES['gemm_sustained'] = case_gemm_sustained


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--case', default=None)
  ap.add_argument('--cases', default=None, help='comma-separated cases run in THIS process (compute-sanitizer runs)')
  ap.add_argument('--only', default=None, help='substring filter when running all')
  ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'kernel_check.json'))
  ap.add_argument('--timeout', type=int, default=150)
  args = ap.parse_args()
  if args.case:
    res = CASES[args.case]()
    print('RESULT ' + json.dumps(res))
    return
  if args.cases:
    for name in args.cases.split(','):
      res = CASES[name]()
      print(f'RESULT {name} ' + json.dumps(res), flush=True)
    return
  os.makedirs(os.path.dirname(args.out), exist_ok=True)
  allres = {}
  for name in CASES:
    if args.only and args.only not in name:
      continue
    t0 = time.time()
    try:
      p = subprocess.run([sys.executable, os.path.abspath(__file__), '--case', name], capture_output=True, text=True,
                         timeout=args.timeout)
      line = [ln for ln in p.stdout.splitlines() if ln.startswith('RESULT ')]
      if line:
        allres[name] = json.loads(line[-1][7:])
      else:
        allres[name] = {'error': 'no result', 'rc': p.returncode, 'stderr': p.stderr[-1500:]}
    except subprocess.TimeoutExpired:
      allres[name] = {'error': f'timeout after {args.timeout}s (kernel hang?)'}
    allres[name + '__secs'] = round(time.time() - t0, 1)
    print(name, json.dumps(allres[name])[:600], flush=True)
    with open(args.out, 'w') as f:
      json.dump(allres, f, indent=1)


if __name__ == '__main__':
  main()

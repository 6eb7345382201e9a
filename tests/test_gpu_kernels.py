"""GPU parity tests, kernel level: the CUDA path (through the C ABI) against the oracle on the same seeded inputs and
against the fixtures the reference produced.  Tolerances (BASELINE.json north_star): bf16 outputs rtol 2e-2 of the
tensor scale, fp32-accumulated reductions 1e-3, integer work bit-exact."""

import math
import os

import pytest
import torch

from conftest import assert_close
from oracle import plainlm_oracle as orc

pytestmark = pytest.mark.gpu
BF16_RTOL = 2e-2
F32_RTOL = 1e-3
DEV = 'cuda'
bf16 = torch.bfloat16


@pytest.fixture(scope='module')
def comp(golden_dir):
  return torch.load(os.path.join(golden_dir, 'components.pt'))


def _ops():
  from plainlm_b200 import ops, _lib

  return ops, _lib


# ------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize('a_k,b_k', [(True, True), (True, False), (False, False), (False, True)])
@pytest.mark.parametrize('M,N,K', [(16, 64, 64), (128, 256, 128), (328, 264, 200), (1024, 1024, 1024)])
def test_gemm_operand_majors(a_k, b_k, M, N, K):
  ops, _lib = _ops()
  g = torch.Generator().manual_seed(M + N + K)
  A = (torch.randn(M, K, generator=g) * 0.5).to(bf16)
  B = (torch.randn(N, K, generator=g) * 0.5).to(bf16)
  ref = A.float() @ B.float().t()  # oracle: fp32 contraction of the same bf16 operands (what F.linear does)
  a_st = (A if a_k else A.t().contiguous()).to(DEV)
  b_st = (B if b_k else B.t().contiguous()).to(DEV)
  out = torch.full((M, N), float('nan'), device=DEV)
  ops.gemm(a_st, b_st, out, a_kmajor=a_k, b_kmajor=b_k, epilogue=_lib.EPI_F32)
  assert_close(out, ref, F32_RTOL, what='gemm f32')
  outb = torch.full((M, N), float('nan'), device=DEV, dtype=bf16)
  ops.gemm(a_st, b_st, outb, a_kmajor=a_k, b_kmajor=b_k)
  assert_close(outb, ref, BF16_RTOL, what='gemm bf16')


@pytest.mark.parametrize('bn', [128, 256])
def test_gemm_epilogues(bn, request):
  ops, _lib = _ops()
  _lib.gemm_tuning(bn=bn)  # force the tile width (diagnostic override; restored below)
  request.addfinalizer(_lib.gemm_tuning)
  g = torch.Generator().manual_seed(5)
  M, N, K, T, hd = 256, 768, 256, 128, 64
  A = (torch.randn(M, K, generator=g) * 0.5).to(bf16)
  B = (torch.randn(N, K, generator=g) * 0.5).to(bf16)
  R = torch.randn(M, N, generator=g)
  ref = A.float() @ B.float().t()
  a, b = A.to(DEV), B.to(DEV)
  out = torch.empty(M, N, device=DEV)
  ops.gemm(a, b, out, epilogue=_lib.EPI_RESID_F32, residual=R.to(DEV))
  assert_close(out, ref + R, F32_RTOL, what='residual epilogue')
  acc = R.to(DEV).clone()
  ops.gemm(a.t().contiguous(), b.t().contiguous(), acc, a_kmajor=False, b_kmajor=False, epilogue=_lib.EPI_ATOMIC_F32,
           splits=0)
  assert_close(acc, ref + R, F32_RTOL, what='atomic accumulate epilogue')
  # fused RoPE on the q|k columns == oracle apply_rope on the GEMM output (models/embeddings.py:15-30)
  table = orc.rope_table(hd, T)
  outb = torch.empty(M, N, device=DEV, dtype=bf16)
  ops.gemm(a, b, outb, epilogue=_lib.EPI_BF16_ROPE, rope_table=table.to(DEV), rope_cols=512, rope_T=T, head_dim=hd)
  qk = orc.apply_rope(ref[:, :512].reshape(M // T, T, 8, hd), table).reshape(M, 512)
  assert_close(outb[:, :512], qk, BF16_RTOL, what='rope epilogue q|k')
  assert_close(outb[:, 512:], ref[:, 512:], BF16_RTOL, what='rope epilogue v untouched')


@pytest.mark.parametrize('M,F,K', [(128, 128, 64), (256, 256, 128), (328, 384, 200), (2048, 2816, 1024)])
def test_gemm_swiglu_epilogue(M, F, K):
  """fc1 with the GLU gate applied in the epilogue (models/components.py:55-56): u must equal the plain GEMM bit for
  bit, h must equal the stand-alone SwiGLU kernel on that u bit for bit, and both must match the oracle."""
  ops, _lib = _ops()
  g = torch.Generator().manual_seed(M + F)
  A = (torch.randn(M, K, generator=g) * 0.5).to(bf16)
  B = (torch.randn(2 * F, K, generator=g) * 0.5).to(bf16)
  a, b = A.to(DEV), B.to(DEV)
  u_plain = torch.empty(M, 2 * F, device=DEV, dtype=bf16)
  ops.gemm(a, b, u_plain)
  h_plain = torch.empty(M, F, device=DEV, dtype=bf16)
  ops.swiglu_fwd(u_plain, h_plain)
  u = torch.full((M, 2 * F), float('nan'), device=DEV, dtype=bf16)
  h = torch.full((M, F), float('nan'), device=DEV, dtype=bf16)
  ops.gemm(a, b, u, epilogue=_lib.EPI_BF16_SWIGLU, out2=h)
  assert torch.equal(u, u_plain)
  assert torch.equal(h, h_plain)
  ref_u = A.float() @ B.float().t()
  assert_close(u, ref_u, BF16_RTOL, what='u')
  ub = u.float().cpu()
  assert_close(h, torch.nn.functional.silu(ub[:, :F]) * ub[:, F:], BF16_RTOL, what='h')


@pytest.mark.parametrize('M,F,K', [(128, 256, 64), (328, 512, 200), (2048, 2816, 1024)])
def test_gemm_glu_backward_epilogue(M, F, K):
  """fc2's input-gradient GEMM with the GLU backward fused into its epilogue (PLM_EPI_BF16_GLU_BWD) against autograd of
  silu(a) * z through the fp32 product of the same bf16 operands, and against the unfused pair GEMM -> plm_swiglu_bwd."""
  ops, _lib = _ops()
  g = torch.Generator().manual_seed(M + F)
  dy = (torch.randn(M, K, generator=g) * 0.5).to(bf16)          # gradient arriving at fc2's output  [M, d]
  w2 = (torch.randn(K, F, generator=g) * (1.0 / K ** 0.5)).to(bf16)  # fc2.weight [d, F]: read in place as an MN-major B
  u = (torch.randn(M, 2 * F, generator=g) * 1.5).to(bf16)        # saved fc1 output [a | z]
  uf = u.float().requires_grad_(True)
  a, z = uf[:, :F], uf[:, F:]
  h = torch.nn.functional.silu(a) * z
  dg_ref = dy.float() @ w2.float()
  h.backward(dg_ref)
  du = torch.full((M, 2 * F), float('nan'), device=DEV, dtype=bf16)
  ops.gemm(dy.to(DEV), w2.to(DEV), du, a_kmajor=True, b_kmajor=False, epilogue=_lib.EPI_BF16_GLU_BWD, out2=u.to(DEV))
  assert_close(du, uf.grad, BF16_RTOL, what='fused GLU backward')
  dg = torch.empty(M, F, device=DEV, dtype=bf16)
  du2 = torch.empty(M, 2 * F, device=DEV, dtype=bf16)
  ops.gemm(dy.to(DEV), w2.to(DEV), dg, a_kmajor=True, b_kmajor=False)
  ops.swiglu_bwd(dg, u.to(DEV), du2)
  assert_close(du, du2, BF16_RTOL, what='fused vs unfused GLU backward')
  with pytest.raises(Exception):  # K-major B is not instantiated for this kind
    ops.gemm(dy.to(DEV), w2.t().contiguous().to(DEV), du, epilogue=_lib.EPI_BF16_GLU_BWD, out2=u.to(DEV))


def test_gemm_lm_head_shape_sampled():
  """Full LM-head shape of the 420M config (16384 x 50280 x 1024): spot-check entries against fp64 dot products and
  the ragged vocabulary tail (50280 = 196*256 + 104)."""
  ops, _lib = _ops()
  M, N, K = 16384, 50280, 1024
  g = torch.Generator(device=DEV).manual_seed(7)
  a = (torch.randn(M, K, device=DEV, generator=g) * 0.5).to(bf16)
  b = (torch.randn(N, K, device=DEV, generator=g) * 0.5).to(bf16)
  out = torch.empty(M, N, device=DEV, dtype=bf16)
  ops.gemm(a, b, out)
  rows = torch.tensor([0, 1, 127, 128, 4095, 9000, 16383])
  cols = torch.tensor([0, 255, 256, 25000, 50175, 50176, 50200, 50279])
  ref = a[rows.to(DEV)].double() @ b[cols.to(DEV)].double().t()
  assert_close(out[rows.to(DEV)][:, cols.to(DEV)], ref, BF16_RTOL, what='lm_head samples')
  assert not torch.isnan(out[:, -104:].float()).any()


def test_gemm_rejects_bad_arguments():
  ops, _lib = _ops()
  from plainlm_b200._lib import PlmError

  a = torch.zeros(128, 64, device=DEV, dtype=bf16)
  b = torch.zeros(100, 64, device=DEV, dtype=bf16)  # N = 100 is not a multiple of 8
  with pytest.raises(PlmError, match='multiples of 8'):
    ops.gemm(a, b, torch.zeros(128, 100, device=DEV, dtype=bf16))
  with pytest.raises(PlmError, match='split-K'):
    ops.gemm(a, a, torch.zeros(128, 128, device=DEV, dtype=bf16), splits=2)
  with pytest.raises(PlmError, match='SwiGLU'):  # N/2 = 64 is not a multiple of 128
    ops.gemm(a, a, torch.zeros(128, 128, device=DEV, dtype=bf16), epilogue=_lib.EPI_BF16_SWIGLU,
             out2=torch.zeros(128, 64, device=DEV, dtype=bf16))


# ------------------------------------------------------------------------------------------- reference modules
def test_rmsnorm_against_reference_module(comp):
  from plainlm_b200.models import functional as PF

  f = comp['rmsnorm']
  x = f['x'].to(DEV).requires_grad_(True)
  w = f['w'].to(DEV).requires_grad_(True)
  y = PF.rmsnorm(x, w, 1e-6)
  assert y.dtype == bf16
  assert_close(y, f['y'], BF16_RTOL, what='y')
  y.backward(f['dy'].to(DEV).to(bf16))
  assert_close(x.grad, f['dx'], BF16_RTOL, what='dx')   # dy is rounded to bf16 on the way in
  assert_close(w.grad, f['dw'], BF16_RTOL, what='dw')


@pytest.mark.parametrize('rows,d', [(1, 128), (777, 384), (4096, 1024), (300, 2048)])
def test_rmsnorm_against_oracle(rows, d):
  ops, _ = _ops()
  g = torch.Generator().manual_seed(rows)
  x = torch.randn(rows, d, generator=g) * 3
  w = torch.rand(d, generator=g) + 0.5
  dy = torch.randn(rows, d, generator=g).to(bf16)
  dx_in = torch.randn(rows, d, generator=g)
  xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
  y_ref = orc.rmsnorm(xr, wr)
  y_ref.backward(dy.float())
  y = torch.empty(rows, d, device=DEV, dtype=bf16)
  rstd = torch.empty(rows, device=DEV)
  ops.rmsnorm_fwd(x.to(DEV), w.to(DEV), y, rstd, 1e-6)
  assert_close(y, y_ref, BF16_RTOL, what='y')
  assert_close(rstd, torch.rsqrt(x.pow(2).mean(-1) + 1e-6), F32_RTOL, what='rstd')
  nb = ops.rmsnorm_bwd_blocks(rows)
  part = torch.empty(nb, d, device=DEV)
  dx = torch.empty(rows, d, device=DEV)
  dxb = torch.empty(rows, d, device=DEV, dtype=bf16)
  ops.rmsnorm_bwd(dy.to(DEV), x.to(DEV), w.to(DEV), rstd, dx_in.to(DEV), dx, dxb, part)
  dw = torch.zeros(d, device=DEV)
  ops.colsum_accum(part, dw, nb)
  assert_close(dx, xr.grad + dx_in, F32_RTOL, what='dx')
  assert_close(dxb, xr.grad + dx_in, BF16_RTOL, what='dx bf16 copy')
  assert_close(dw, wr.grad, F32_RTOL, what='dw')


def test_glu_against_reference_module(comp):
  from plainlm_b200.models.components import GLU

  f = comp['glu']
  glu = GLU(64, int(8 / 3 * 64)).to(DEV)
  assert glu.hidden_dim == f['hidden']
  with torch.no_grad():
    glu.fc1.weight.copy_(f['w1'])
    glu.fc2.weight.copy_(f['w2'])
  x = f['x'].to(DEV).requires_grad_(True)
  y = glu(x.to(bf16))
  assert_close(y, f['y'], BF16_RTOL, what='y')
  y.backward(f['dy'].to(DEV).to(bf16))
  assert_close(x.grad, f['dx'], BF16_RTOL, what='dx')
  assert_close(glu.fc1.weight.grad, f['dw1'], BF16_RTOL, what='dw1')
  assert_close(glu.fc2.weight.grad, f['dw2'], BF16_RTOL, what='dw2')


def test_attention_against_reference_module(comp):
  from plainlm_b200.models.transformer import Attention, ModelConfig
  from plainlm_b200.models.embeddings import precompute_freqs_cis

  f = comp['attn']
  mc = ModelConfig(vocab_size=16, seq_len=32, dim=128, expand=8 / 3, n_layers=1, n_heads=2, mlp='glu')
  att = Attention(mc).to(DEV)
  with torch.no_grad():
    att.w_qkv.weight.copy_(f['w_qkv'])
    att.w_out.weight.copy_(f['w_out'])
  table = precompute_freqs_cis(64, 32, 500000)
  assert torch.equal(table, comp['rope']['table'])  # same table as the reference, bit for bit
  x = f['x'].to(DEV).requires_grad_(True)
  y = att(x.to(bf16), table, None)
  assert_close(y, f['y'], BF16_RTOL, what='y')
  y.backward(f['dy'].to(DEV).to(bf16))
  assert_close(x.grad, f['dx'], BF16_RTOL, what='dx')
  assert_close(att.w_qkv.weight.grad, f['dw_qkv'], BF16_RTOL, what='dw_qkv')
  assert_close(att.w_out.weight.grad, f['dw_out'], BF16_RTOL, what='dw_out')
  # dense bool mask in the reference's (bsz, L, L) format
  fm = comp['attn_masked']
  mask = torch.stack([orc.mask_from_segment_starts(orc.doc_segment_starts(dl, 32)) for dl in fm['docs_lengths']])
  with torch.no_grad():
    ym = att(x.detach().to(bf16), table, mask.to(DEV))
  assert_close(ym, fm['y'], BF16_RTOL, what='masked y')


# ------------------------------------------------------------------------------------------- attention vs oracle
def _rand_docs(B, T, seed):
  g = torch.Generator().manual_seed(seed)
  out = []
  for _ in range(B):
    cuts = sorted(set(torch.randint(1, T + 1, (5,), generator=g).tolist()))
    edges = [0] + cuts + [T + 1]
    out.append([edges[i + 1] - edges[i] for i in range(len(edges) - 1)])
  return out


@pytest.mark.parametrize('B,T,H,doc', [(1, 128, 1, False), (2, 384, 3, False), (2, 200, 2, False), (2, 512, 2, True),
                                       (1, 96, 2, True), (1, 2048, 2, True)])
def test_flash_attention_fwd_bwd_against_oracle(B, T, H, doc):
  ops, _ = _ops()
  from plainlm_b200.data_utils import seg_start_from_docs_lengths

  hd, d = 64, H * 64
  g = torch.Generator().manual_seed(T + H)
  qkv = torch.randn(B * T, 3 * d, generator=g).to(bf16)
  dout = (torch.randn(B * T, d, generator=g) * 0.5).to(bf16)
  seg, mask = None, None
  if doc:
    docs = _rand_docs(B, T, seed=T)
    seg = seg_start_from_docs_lengths(docs, T)
    mask = torch.stack([orc.mask_from_segment_starts(s) for s in seg]).unsqueeze(1)
  x = qkv.float().requires_grad_(True)
  q, k, v = x.view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4)
  o_ref = orc.sdpa(q, k, v, mask).transpose(1, 2).reshape(B * T, d)
  o_ref.backward(dout.float())
  out = torch.full((B * T, d), float('nan'), device=DEV, dtype=bf16)
  lse = torch.empty(B, H, T, device=DEV)
  segd = None if seg is None else seg.reshape(-1).to(DEV)
  ops.attn_fwd(qkv.to(DEV), out, lse, B, T, H, hd, seg_start=segd)
  assert_close(out, o_ref, BF16_RTOL, what='out')
  dqkv = torch.full((B * T, 3 * d), float('nan'), device=DEV, dtype=bf16)
  delta = torch.empty(B, H, T, device=DEV)
  dq_acc = torch.empty(B * T, d, device=DEV)
  ops.attn_bwd(qkv.to(DEV), out, dout.to(DEV), lse, dqkv, delta, dq_acc, B, T, H, hd, seg_start=segd)
  for name, sl in (('dq', slice(0, d)), ('dk', slice(d, 2 * d)), ('dv', slice(2 * d, 3 * d))):
    assert_close(dqkv[:, sl], x.grad[:, sl], BF16_RTOL, what=name)


def test_full_document_mask_equals_causal():
  """A single document per row must reproduce plain causal attention bit for bit (reference: SURVEY App. B)."""
  ops, _ = _ops()
  B, T, H, hd = 2, 256, 2, 64
  qkv = torch.randn(B * T, 3 * H * hd, device=DEV).to(bf16)
  o1 = torch.empty(B * T, H * hd, device=DEV, dtype=bf16)
  o2 = torch.empty_like(o1)
  lse = torch.empty(B, H, T, device=DEV)
  ops.attn_fwd(qkv, o1, lse, B, T, H, hd)
  ops.attn_fwd(qkv, o2, lse, B, T, H, hd, seg_start=torch.zeros(B * T, dtype=torch.int32, device=DEV))
  assert torch.equal(o1, o2)


# ------------------------------------------------------------------------------------------- small bandwidth kernels
def test_swiglu_embed_ce_casts_against_oracle():
  import torch.nn.functional as F

  ops, _lib = _ops()
  g = torch.Generator().manual_seed(3)
  rows, Fh = 300, 768
  u = torch.randn(rows, 2 * Fh, generator=g).to(bf16)
  dh = torch.randn(rows, Fh, generator=g).to(bf16)
  uf = u.float().requires_grad_(True)
  a, z = uf.split(Fh, dim=1)
  href = F.silu(a) * z  # models/components.py:55-56
  href.backward(dh.float())
  h = torch.empty(rows, Fh, device=DEV, dtype=bf16)
  ops.swiglu_fwd(u.to(DEV), h)
  du = torch.empty(rows, 2 * Fh, device=DEV, dtype=bf16)
  ops.swiglu_bwd(dh.to(DEV), u.to(DEV), du)
  assert_close(h, href, BF16_RTOL, what='swiglu fwd')
  assert_close(du, uf.grad, BF16_RTOL, what='swiglu bwd')

  V, d = 1000, 256
  W = torch.randn(V, d, generator=g)
  ids = torch.randint(0, V, (rows,), generator=g)
  x = torch.empty(rows, d, device=DEV)
  ops.embed_fwd(ids.to(DEV), W.to(DEV), x)
  assert torch.equal(x.cpu(), W[ids])  # gather is bit-exact
  dx = torch.randn(rows, d, generator=g)
  dW = torch.zeros(V, d, device=DEV)
  ops.embed_bwd(ids.to(DEV), dx.to(DEV), dW)
  assert_close(dW, torch.zeros(V, d).index_add_(0, ids, dx), F32_RTOL, what='embed bwd')

  Vc, rc = 50280, 64
  logits = (torch.randn(rc, Vc, generator=g) * 2).to(bf16)
  tg = torch.randint(0, Vc, (rc,), generator=g)
  tg[3] = -100  # ignore_index
  lf = logits.float().requires_grad_(True)
  loss_ref = orc.loss_fn(lf, tg)
  (loss_ref / 4).backward()
  lg = logits.to(DEV).clone()
  rl, rlse, stats = torch.empty(rc, device=DEV), torch.empty(rc, device=DEV), torch.zeros(4, device=DEV)
  ops.ce_fwd_bwd(lg, tg.to(DEV), rl, rlse, stats, Vc, grad_scale=0.25)
  assert abs(stats[2].item() - loss_ref.item()) <= F32_RTOL * loss_ref.item()
  assert stats[1].item() == rc - 1
  assert_close(lg, lf.grad, BF16_RTOL, what='dlogits')
  assert torch.count_nonzero(lg[3]).item() == 0

  src = torch.randn(12345, generator=g)
  dst = torch.empty(12345, device=DEV, dtype=bf16)
  ops.cast_f32_bf16(src.to(DEV), dst, 0.5)
  assert torch.equal(dst.cpu(), (src * 0.5).to(bf16))  # round-to-nearest-even, bit-exact
  back = torch.empty(12345, device=DEV)
  ops.cast_bf16_f32(dst, back, 2.0)
  assert torch.equal(back.cpu(), dst.cpu().float() * 2.0)

  # data-parallel tail: bf16 wire -> fp32 gradients + squared norm in one pass, deterministic
  wire = torch.randn(1_000_003, generator=g).to(bf16)
  gdst = torch.full((1_000_003,), float('nan'), device=DEV)
  wsu = torch.empty(_lib.SUMSQ_WORKSPACE, device=DEV)
  n1, n2 = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
  ops.unpack_sumsq(wire.to(DEV), gdst, wsu, n1)
  assert torch.equal(gdst.cpu(), wire.float())
  ops.sumsq(gdst, wsu, n2)
  assert torch.equal(n1, n2)  # same fixed-order reduction as plm_sumsq over the unpacked values

  gbuf = torch.randn(1_000_003, generator=g)
  ws = torch.empty(_lib.SUMSQ_WORKSPACE, device=DEV)
  o1, o2 = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
  ops.sumsq(gbuf.to(DEV), ws, o1)
  ops.sumsq(gbuf.to(DEV), ws, o2)
  assert torch.equal(o1, o2)  # deterministic
  assert abs(o1.item() - gbuf.double().pow(2).sum().item()) <= F32_RTOL * o1.item()


@pytest.mark.parametrize('rows,V,d,store', [(128, 256, 64, True), (200, 1000, 128, True), (333, 50280, 256, True),
                                             (200, 1000, 128, False), (1024, 2056, 1024, True)])
def test_lmhead_fused_cross_entropy_against_oracle(rows, V, d, store):
  """plm_lmhead_ce_fwd (LM-head GEMM whose epilogue reduces the cross-entropy statistics) + plm_ce_grad against the
  oracle's loss on the same bf16 operands: ragged last column tile (V % 256 != 0, V % 32 != 0), rows % 128 != 0,
  ignore_index rows, targets in the first / last column, and the loss-only form that never writes [rows, V]."""
  ops, _ = _ops()
  g = torch.Generator().manual_seed(rows + V)
  h = (torch.randn(rows, d, generator=g) * 1.5).to(bf16)
  w = (torch.randn(V, d, generator=g) * (2.0 / d ** 0.5)).to(bf16)
  tg = torch.randint(0, V, (rows,), generator=g)
  tg[1], tg[2], tg[5], tg[rows - 1] = 0, V - 1, -100, -100
  logits_ref = (h.float() @ w.float().t()).to(bf16)  # autocast: the Linear rounds to bf16, CrossEntropyLoss upcasts
  lf = logits_ref.float().requires_grad_(True)
  loss_ref = orc.loss_fn(lf, tg)
  (loss_ref * 0.25).backward()
  n_valid = rows - 2

  tiles = ops.lmhead_ce_tiles(V)
  assert tiles == (V + 255) // 256
  ld = (V + 7) // 8 * 8
  lg = torch.full((rows, ld), float('nan'), device=DEV, dtype=bf16)[:, :V] if store else None
  partial = torch.empty(2 * tiles * rows, device=DEV)
  tgl, rl, rlse = (torch.empty(rows, device=DEV) for _ in range(3))
  stats = torch.zeros(4, device=DEV)
  ops.lmhead_ce_fwd(h.to(DEV), w.to(DEV), tg.to(DEV), lg, partial, tgl, rl, rlse, stats, V)
  assert stats[1].item() == n_valid
  assert abs(stats[2].item() - loss_ref.item()) <= F32_RTOL * loss_ref.item(), (stats[2].item(), loss_ref.item())
  lse_ref = torch.logsumexp(logits_ref.float(), dim=1)
  assert_close(rlse, lse_ref, F32_RTOL, what='row lse')
  assert rl[5].item() == 0.0 and rl[rows - 1].item() == 0.0  # ignored rows contribute nothing
  if not store:
    return
  # the stored tile is the GEMM's bf16 rounding of the same fp32 accumulation: equal up to accumulation order
  assert_close(lg, logits_ref, BF16_RTOL, what='logits')
  # the loss the epilogue reduced must be the loss OF THE STORED (rounded) logits, to fp32 accuracy
  loss_of_stored = orc.loss_fn(lg.float().cpu(), tg)
  assert abs(stats[2].item() - loss_of_stored.item()) <= 2e-5 * loss_of_stored.item()
  ops.ce_grad(lg, tg.to(DEV), rlse, stats, V, grad_scale=0.25)
  assert_close(lg, lf.grad, BF16_RTOL, what='dlogits')
  assert torch.count_nonzero(lg[5]).item() == 0


def test_rope_kernel_against_reference_fixture(comp):
  ops, _ = _ops()
  r = comp['rope']
  B, T, H, hd = r['q'].shape
  q, k = r['q'].to(bf16), r['k'].to(bf16)
  qkv = torch.cat([q.reshape(B * T, -1), k.reshape(B * T, -1), torch.zeros(B * T, H * hd, dtype=bf16)], 1).to(DEV)
  table = r['table'].reshape(T, hd // 2, 2).contiguous().to(DEV)
  rot = qkv.clone()
  ops.rope_qk_(rot, table, T, H, hd)
  exp_q = orc.apply_rope(q, orc.rope_table(hd, T)).reshape(B * T, -1)
  assert_close(rot[:, : H * hd], exp_q, 1e-2, what='rope q')  # same bf16 inputs, one bf16 rounding
  ops.rope_qk_(rot, table, T, H, hd, inverse=True)
  assert_close(rot, qkv, 2e-2, what='rope inverse round trip')


def test_seg_start_kernel_bit_exact():
  ops, _ = _ops()
  from plainlm_b200.data_utils import seg_start_from_docs_lengths

  T = 300
  docs = _rand_docs(4, T, seed=1) + [[T + 1], [1] * (T + 1), [T, 1], [1, T]]
  host = seg_start_from_docs_lengths(docs, T)
  lengths = torch.tensor([n for dl in docs for n in dl], dtype=torch.int32, device=DEV)
  offsets = torch.tensor([0] + list(torch.tensor([len(dl) for dl in docs]).cumsum(0)), dtype=torch.int32, device=DEV)
  seg = torch.empty(len(docs) * T, dtype=torch.int32, device=DEV)
  ops.seg_start_from_lengths(lengths, offsets, seg, len(docs), T)
  assert torch.equal(seg.cpu().view(len(docs), T), host)


# ------------------------------------------------------------------------------------------- optimizers
def test_optimizers_against_reference_fixture(golden_dir):
  from collections import namedtuple
  from plainlm_b200.optim import intialize_optimizer
  from plainlm_b200.optim.flat import GradClip
  from plainlm_b200 import ops, _lib

  fx = torch.load(os.path.join(golden_dir, 'optim.pt'))
  for name in ('adamw', 'signsgd'):
    f = fx[name]
    cfgd = dict(f['cfg'])
    p = torch.nn.Parameter(f['p0'].to(DEV))
    n = torch.nn.Parameter(f['n0'].to(DEV))
    opt = intialize_optimizer([{'params': [p], 'weight_decay': 0.1}, {'params': [n], 'weight_decay': 0.0}],
                              namedtuple('Cfg', cfgd.keys())(**cfgd))
    ws = torch.empty(_lib.SUMSQ_WORKSPACE, device=DEV)
    gsq = torch.zeros(1, device=DEV)
    for i, (gp, gn) in enumerate(f['grads']):
      for grp in opt.param_groups:
        grp['lr'] = cfgd['lr'] * (i + 1) / 3
      p.grad, n.grad = gp.to(DEV), gn.to(DEV)
      ops.sumsq(p.grad, ws, gsq, accumulate=False)
      ops.sumsq(n.grad, ws, gsq, accumulate=True)
      assert abs(math.sqrt(gsq.item()) - f['snaps'][i]['norm']) <= F32_RTOL * f['snaps'][i]['norm']
      opt.step(grad_clip=GradClip(gsq, 1.0))
      assert_close(p, f['snaps'][i]['p'], 1e-5, what=f'{name} p step {i}')
      assert_close(n, f['snaps'][i]['n'], 1e-5, what=f'{name} n step {i}')
    assert sorted(opt.state[p].keys()) == f['state_keys']
    sd = opt.state_dict()
    assert sorted(sd['state'][0].keys()) == f['state_keys']


def test_sgd_nadamw_against_torch_fixture(golden_dir):
  """the flat SGD / NAdamW kernels against torch.optim.SGD / NAdam(decoupled_weight_decay=True)."""
  from collections import namedtuple
  from plainlm_b200.optim import intialize_optimizer
  from plainlm_b200.optim.flat import GradClip
  from plainlm_b200 import ops, _lib

  fx = torch.load(os.path.join(golden_dir, 'variants.pt'))['optim']
  for name in ('sgd', 'nadamw'):
    f = fx[name]
    cfgd = dict(f['cfg'])
    p = torch.nn.Parameter(f['p0'].to(DEV))
    n = torch.nn.Parameter(f['n0'].to(DEV))
    opt = intialize_optimizer([{'params': [p], 'weight_decay': 0.1}, {'params': [n], 'weight_decay': 0.0}],
                              namedtuple('Cfg', cfgd.keys())(**cfgd))
    ws = torch.empty(_lib.SUMSQ_WORKSPACE, device=DEV)
    gsq = torch.zeros(1, device=DEV)
    for i, (gp, gn) in enumerate(f['grads']):
      for grp in opt.param_groups:
        grp['lr'] = cfgd['lr'] * (i + 1) / 4
      p.grad, n.grad = gp.to(DEV), gn.to(DEV)
      ops.sumsq(p.grad, ws, gsq, accumulate=False)
      ops.sumsq(n.grad, ws, gsq, accumulate=True)
      opt.step(grad_clip=GradClip(gsq, 1.0))
      assert_close(p, f['snaps'][i]['p'], 1e-5, what=f'{name} p step {i}')
      assert_close(n, f['snaps'][i]['n'], 1e-5, what=f'{name} n step {i}')
    assert sorted(opt.state[p].keys()) == f['state_keys']
    sd = opt.state_dict()
    assert sorted(sd['state'][0].keys()) == f['state_keys']
    opt.load_state_dict(sd)  # checkpoint round trip keeps the host-side scalars usable
    p.grad, n.grad = f['grads'][0][0].to(DEV), f['grads'][0][1].to(DEV)
    opt.step()


def test_activation_kernels_against_torch():
  """plm_act_fwd / plm_act_bwd (MLP: silu, MLPReluSquared: relu^2) against torch on the same bf16 inputs."""
  ops, _lib = _ops()
  g = torch.Generator().manual_seed(3)
  u = (torch.randn(64, 512, generator=g) * 2).to(bf16)
  dh = torch.randn(64, 512, generator=g).to(bf16)
  for kind, fn in ((_lib.ACT_SILU, torch.nn.functional.silu), (_lib.ACT_RELU2, lambda x: torch.relu(x).pow(2))):
    uf = u.float().requires_grad_(True)
    ref = fn(uf)
    ref.backward(dh.float())
    h = torch.empty(64, 512, device=DEV, dtype=bf16)
    du = torch.empty(64, 512, device=DEV, dtype=bf16)
    ops.act_fwd(u.to(DEV), h, kind)
    ops.act_bwd(dh.to(DEV), u.to(DEV), du, kind)
    assert_close(h, ref.detach(), BF16_RTOL, what=f'act {kind} fwd')
    assert_close(du, uf.grad, BF16_RTOL, what=f'act {kind} bwd')


def test_first_call_from_autograd_thread_in_fresh_process():
  """Regression: the autograd worker thread may have no current CUDA context when a backward kernel is the first thing
  it runs; the C ABI must bind the context that owns its pointers (and never default to device 0)."""
  import subprocess
  import sys

  code = (
    'import torch, sys\n'
    f'sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})\n'
    'from plainlm_b200.models import functional as PF\n'
    'x = torch.randn(64, 128, device="cuda", dtype=torch.bfloat16, requires_grad=True)\n'
    'w = torch.randn(256, 128, device="cuda", requires_grad=True)\n'
    'y = PF.linear(x, w)\n'
    'y.backward(torch.ones_like(y))\n'
    'torch.cuda.synchronize()\n'
    'ref = torch.ones(64, 256, device="cuda") @ w.detach().bfloat16().float()\n'
    'assert (x.grad.float() - ref).abs().max() <= 2e-2 * ref.abs().max()\n'
    'print("OK")\n'
  )
  res = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
  assert res.returncode == 0 and 'OK' in res.stdout, res.stderr[-2000:]

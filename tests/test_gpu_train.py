"""GPU parity tests, model / engine level: plainlm_b200 against the reference's recorded outputs (tests/golden) and
against the oracle run on the same seeded inputs.  Loss-curve tolerance: 1 % (BASELINE.json north_star)."""

import json
import os
from collections import namedtuple

import pytest
import torch

from conftest import assert_close
from oracle import plainlm_oracle as orc

pytestmark = pytest.mark.gpu
DEV = 'cuda'
BF16_RTOL = 2e-2
TINY = dict(vocab_size=256, d_model=128, n_layers=2, n_heads=2, seq_len=32, expand='8/3', mlp_class='glu',
            tie_embeddings=False, model='transformer')


def _cfg(**kw):
  return namedtuple('Cfg', kw.keys())(**kw)


def _model(cfg=TINY, seed=7):
  from plainlm_b200.models import construct_model

  model, _ = construct_model(_cfg(**cfg))
  params = orc.init_params(cfg['vocab_size'], cfg['d_model'], cfg['n_layers'], cfg['n_heads'], seed=seed)
  model.load_state_dict(params, strict=True)
  return model.to(DEV), params


def _engine_cfg(**over):
  base = dict(seq_len=32, grad_accumulation_steps=2, grad_clip=1.0, dtype='bfloat16', intra_doc_masking=False,
              resume=False, torch_compile=False, weight_decay=0.1, optim='adamw', lr=3e-3, beta1=0.9, beta2=0.95,
              fused_optim=True, scheduler='warmup_cosine', warmup_steps=0.1, cooldown_steps=None, lr_start=0.0,
              lr_end=1e-5, lr_end_pct=None, steps_budget=20, dampening=0.0)
  base.update(over)
  return base


@pytest.fixture(scope='module')
def tiny(golden_dir):
  return torch.load(os.path.join(golden_dir, 'model_tiny.pt'))


def _check_grads(named_grads, ref_grads, rtol):
  for k, g in ref_grads.items():
    got = named_grads[k].detach().float().cpu()
    assert abs(float(got.double().norm()) - g['norm']) <= rtol * g['norm'] + 1e-9, (k, float(got.norm()), g['norm'])
    assert_close(got.flatten()[:32], g['head'], rtol, atol=rtol * g['norm'] / got.numel() ** 0.5, what=k)


def test_modular_forward_backward_vs_reference(tiny):
  """model(x, None) -> logits, autograd backward: the reference's own calling convention (engine.py:109-120)."""
  model, _ = _model()
  T, V = 32, 256
  ids = tiny['ids'].to(DEV)
  inputs, targets = ids[:, :T], ids[:, 1 : T + 1].contiguous()
  logits = model(inputs, None)
  assert logits.dtype == torch.bfloat16 and logits.shape == (2, T, V)
  loss = torch.nn.CrossEntropyLoss()(logits.float().view(-1, V), targets.view(-1))
  model.runtime().flat.zero_grads()
  loss.backward()
  for tag in ('bf16', 'fp32'):
    assert abs(loss.item() - tiny[tag]['loss']) <= 5e-3 * tiny[tag]['loss']
    assert_close(logits[:, :2, :], tiny[tag]['logits_head'], BF16_RTOL, what='logits vs ' + tag)
  _check_grads({k: p.grad for k, p in model.named_parameters()}, tiny['bf16']['grads'], 3e-2)


def test_fused_step_vs_reference(tiny):
  """runtime.loss_and_backward (the engine's path): loss and every gradient against the reference's."""
  model, _ = _model()
  rt = model.runtime()
  T = 32
  ids = tiny['ids'].to(DEV)
  inputs, targets = ids[:, :T].contiguous(), ids[:, 1 : T + 1].contiguous()
  rt.flat.zero_grads()
  loss = rt.loss_and_backward(inputs, targets, None, grad_scale=1.0)
  assert abs(loss.item() - tiny['fp32']['loss']) <= 5e-3 * tiny['fp32']['loss']
  _check_grads({k: p.grad for k, p in model.named_parameters()}, tiny['bf16']['grads'], 3e-2)
  _check_grads({k: p.grad for k, p in model.named_parameters()}, tiny['fp32']['grads'], 3e-2)
  # a second accumulation doubles the gradients (fp32 accumulation in the wgrad epilogue)
  g1 = rt.flat.grads.clone()
  rt.loss_and_backward(inputs, targets, None, grad_scale=1.0)
  assert_close(rt.flat.grads, 2 * g1, 1e-3, what='grad accumulation')


@pytest.mark.parametrize('mlp_class', ['mlp', 'mlp_relu_sq'])
def test_mlp_variants_vs_reference(golden_dir, mlp_class):
  """MLP / MLPReluSquared models — fused runtime step and the modular autograd path against the
  reference model's loss, logits and gradients (tests/golden/variants.pt)."""
  from plainlm_b200.models import construct_model

  fx = torch.load(os.path.join(golden_dir, 'variants.pt'))
  ref = fx['mlp'][mlp_class]
  model, _ = construct_model(_cfg(**dict(TINY, mlp_class=mlp_class)))
  model.load_state_dict(orc.init_params(256, 128, 2, 2, seed=ref['param_seed'], mlp_class=mlp_class), strict=True)
  model = model.to(DEV)
  rt = model.runtime()
  T, V = 32, 256
  ids = fx['ids'].to(DEV)
  inputs, targets = ids[:, :T].contiguous(), ids[:, 1 : T + 1].contiguous()
  rt.flat.zero_grads()
  loss = rt.loss_and_backward(inputs, targets, None, grad_scale=1.0)
  assert abs(loss.item() - ref['fp32']['loss']) <= 5e-3 * ref['fp32']['loss']
  _check_grads({k: p.grad for k, p in model.named_parameters()}, ref['bf16']['grads'], 3e-2)
  _check_grads({k: p.grad for k, p in model.named_parameters()}, ref['fp32']['grads'], 3e-2)
  fused = rt.flat.grads.clone()
  rt.flat.zero_grads()
  logits = model(inputs, None)
  assert_close(logits[:, :2, :], ref['bf16']['logits_head'], BF16_RTOL, what='logits')
  torch.nn.CrossEntropyLoss()(logits.float().view(-1, V), targets.view(-1)).backward()
  assert_close(rt.flat.grads, fused, 3e-2, atol=3e-2 * float(fused.abs().max()), what='modular vs fused grads')


def test_fused_step_doc_masked_vs_reference(tiny):
  from plainlm_b200.data_utils import seg_start_from_docs_lengths

  model, _ = _model()
  rt = model.runtime()
  T = 32
  ref = tiny['doc_fp32']
  ids = tiny['ids'].to(DEV)
  seg = seg_start_from_docs_lengths(ref['docs_lengths'], T).reshape(-1).to(DEV)
  rt.flat.zero_grads()
  loss = rt.loss_and_backward(ids[:, :T].contiguous(), ids[:, 1 : T + 1].contiguous(), seg, grad_scale=1.0)
  assert abs(loss.item() - ref['loss']) <= 5e-3 * ref['loss']
  _check_grads({k: p.grad for k, p in model.named_parameters()}, ref['grads'], 3e-2)
  # the dense-mask calling convention of the reference gives the same logits
  mask = torch.stack([orc.mask_from_segment_starts(orc.doc_segment_starts(dl, T)) for dl in ref['docs_lengths']])
  with torch.no_grad():
    logits = model(ids[:, :T], mask.to(DEV))
  assert_close(logits[:, :2, :], ref['logits_head'], BF16_RTOL, what='masked logits')


@pytest.mark.parametrize('run', ['adamw', 'signsgd', 'adamw_doc', 'adamw_noclip_nosched'])
def test_engine_loss_curve_vs_reference(golden_dir, run):
  """TorchEngine.step on the GPU against the loss curve the reference's TorchEngine produced (CPU, fp32):
  every micro-step within 1 %, learning rates identical, final weights within bf16 tolerance."""
  from plainlm_b200.engine import TorchEngine

  d = json.load(open(os.path.join(golden_dir, 'engine_curves.json')))[run]
  cfgd = dict(d['cfg'])
  cfgd['dtype'] = 'bfloat16'
  model, _ = _model()
  eng = TorchEngine(model, _cfg(**cfgd), DEV, None, None)
  data = torch.tensor(d['data'])
  losses = []
  for i in range(len(d['losses'])):
    batch = {'input_ids': data[i : i + 1]}
    if cfgd['intra_doc_masking']:
      batch['docs_lengths'] = [d['docs_lengths'][i]]
    loss = eng.step(batch)
    assert loss.dim() == 0 and loss.is_cuda
    losses.append(loss)
    assert abs(eng.optimizer.param_groups[0]['lr'] - d['lrs'][i]) <= 1e-12
  losses = torch.stack(losses).cpu().tolist()
  eng.check_nan(wait=True)
  worst = max(abs(a - b) / abs(b) for a, b in zip(losses, d['losses']))
  assert worst <= 1e-2, (worst, losses[-3:], d['losses'][-3:])
  assert losses[-1] < 0.9 * losses[0]  # it actually trains
  for k, f in d['final_params'].items():
    got = dict(model.named_parameters())[k]
    assert abs(float(got.double().norm()) - f['norm']) <= 2e-2 * f['norm'], k


def test_200_step_loss_curve_vs_oracle():
  """First 200 optimizer steps on a low-entropy stream: GPU engine vs the oracle's bf16-autocast restatement run on
  this box's CPU with the same weights and tokens — every step within 1 % (BASELINE.json north_star)."""
  from plainlm_b200.engine import TorchEngine

  cfgd = _engine_cfg(grad_accumulation_steps=1, steps_budget=200, lr=2e-3)
  model, params = _model()
  eng = TorchEngine(model, _cfg(**cfgd), DEV, None, None)
  ocfg = dict(cfgd, n_heads=2)
  tr = orc.OracleTrainer(params, ocfg, 'bf16')
  g = torch.Generator().manual_seed(1234)
  trans = torch.softmax(torch.randn(32, 32, generator=g) * 3, dim=-1)
  B, T = 4, 32
  worst, first, last = 0.0, None, None
  for step in range(200):
    seq = torch.empty(B, T + 1, dtype=torch.int64)
    seq[:, 0] = torch.randint(0, 32, (B,), generator=g)
    for t in range(T):
      seq[:, t + 1] = torch.multinomial(trans[seq[:, t]], 1, generator=g).squeeze(1)
    ref = float(tr.step({'input_ids': seq}))
    got = eng.step({'input_ids': seq}).item()
    worst = max(worst, abs(got - ref) / abs(ref))
    first = got if first is None else first
    last = got
  assert worst <= 1e-2, worst
  assert last < 0.7 * first


def test_checkpoint_roundtrip_and_reference_state_dict(tmp_path):
  """state_dict / optimizer state keep the reference's names and dtypes through flat buffers (checkpoint_utils.py)."""
  from plainlm_b200.engine import TorchEngine

  cfgd = _engine_cfg(grad_accumulation_steps=1)
  model, _ = _model()
  eng = TorchEngine(model, _cfg(**cfgd), DEV, None, None)
  g = torch.Generator().manual_seed(0)
  batches = [{'input_ids': torch.randint(0, 256, (2, 33), generator=g)} for _ in range(6)]
  for b in batches[:3]:
    eng.step(b)
  ckpt = {'step': 3, 'state_dict': model.state_dict(), 'optimizer': eng.optimizer.state_dict(),
          'scheduler': eng.scheduler.state_dict(), 'scaler': eng.scaler.state_dict()}
  path = str(tmp_path / 'ckpt_step_3.pth')
  torch.save(ckpt, path)
  assert sorted(ckpt['optimizer']['state'][0].keys()) == ['exp_avg', 'exp_avg_sq', 'step']
  assert list(ckpt['state_dict'].keys()) == orc.param_names(2)
  ref_losses = [eng.step(b).item() for b in batches[3:]]
  # resume in a fresh engine
  loaded = torch.load(path, map_location='cpu')
  model2, _ = _model(seed=99)
  eng2 = TorchEngine(model2, _cfg(**dict(cfgd, resume=True)), DEV, None, loaded)
  assert eng2.micro_steps == 3
  got = [eng2.step(b).item() for b in batches[3:]]
  for a, b in zip(got, ref_losses):
    assert abs(a - b) <= 1e-3 * abs(b), (got, ref_losses)


def test_nan_loss_raises():
  from plainlm_b200.engine import TorchEngine

  model, _ = _model()
  eng = TorchEngine(model, _cfg(**_engine_cfg(grad_accumulation_steps=1)), DEV, None, None)
  eng.step({'input_ids': torch.zeros(1, 33, dtype=torch.int64)})  # a healthy step first
  before = {k: v.detach().clone() for k, v in model.state_dict().items()}
  m_before = eng.optimizer.state_dict()['state'][0]['exp_avg'].clone()
  with torch.no_grad():
    model.embed_tokens.weight[0, 0] = float('nan')  # token 0 is the whole batch below: every activation goes NaN
  # reference: engine.py:116-117 raises BEFORE backward / optimizer.step; here the accumulation boundary checks every
  # micro-step's loss before the update, so step() itself raises ...
  with pytest.raises(ValueError, match='Train loss is nan'):
    eng.step({'input_ids': torch.zeros(1, 33, dtype=torch.int64)})
  # ... and nothing was updated: weights (except the cell we poisoned), moments and bf16 shadows are untouched
  after = model.state_dict()
  for k, v in before.items():
    if k == 'embed_tokens.weight':
      assert torch.equal(after[k].flatten()[1:], v.flatten()[1:])
    else:
      assert torch.equal(after[k], v), k
  assert torch.equal(eng.optimizer.state_dict()['state'][0]['exp_avg'], m_before)


def test_nonfinite_gradients_never_reach_the_weights():
  """Defence in depth: even if a caller swallowed the exception, the update kernels skip themselves when
  the gradient norm they are handed is not finite."""
  from plainlm_b200 import ops, _lib

  n = 4096
  p = torch.randn(n, device=DEV)
  g = torch.randn(n, device=DEV)
  g[7] = float('nan')
  m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
  pb = torch.zeros(n, device=DEV, dtype=torch.bfloat16)
  ws = torch.empty(_lib.SUMSQ_WORKSPACE, device=DEV)
  gn = torch.zeros(1, device=DEV)
  ops.sumsq(g, ws, gn)
  assert not torch.isfinite(gn).item()
  p0 = p.clone()
  for max_norm in (1.0, 0.0):  # clipping on and off
    ops.adamw_step(p, g, m, v, pb, 1e-3, 0.9, 0.95, 1e-8, 0.1, 1, gnorm_sq=gn, max_norm=max_norm)
    ops.signsgd_step(p, g, m, pb, 1e-3, 0.9, 0.0, 0.1, True, gnorm_sq=gn, max_norm=max_norm)
    ops.sgd_step(p, g, m, pb, 1e-3, 0.9, 0.0, 0.1, True, gnorm_sq=gn, max_norm=max_norm)
  assert torch.equal(p, p0) and not m.any() and not v.any() and not pb.any()


def test_eval_loop():
  from plainlm_b200.engine import TorchEngine

  model, params = _model()
  eng = TorchEngine(model, _cfg(**_engine_cfg()), DEV, None, None)
  g = torch.Generator().manual_seed(1)
  batches = [{'input_ids': torch.randint(0, 256, (2, 33), generator=g)} for _ in range(3)]
  got = eng.eval(batches)
  p = {k: v.clone() for k, v in params.items()}
  ref = sum(float(orc.loss_fn(orc.forward(p, b['input_ids'][:, :32], 2, None, 'bf16'), b['input_ids'][:, 1:33]))
            for b in batches) / 3
  assert abs(got - ref) <= 5e-3 * ref


def test_420m_micro_step_properties():
  """BASELINE config (2) at full width (420M, T=2048): size-independent properties of one micro-step."""
  from plainlm_b200.models import construct_model
  import math

  cfg = dict(vocab_size=50280, d_model=1024, n_layers=24, n_heads=16, seq_len=2048, expand='8/3', mlp_class='glu',
             tie_embeddings=False, model='transformer')
  torch.manual_seed(100)
  model, _ = construct_model(_cfg(**cfg))
  assert model.count_params(False) == 411_304_960
  model = model.to(DEV)
  rt = model.runtime()
  g = torch.Generator().manual_seed(1234)
  ids = torch.randint(0, 50280, (2, 2049), generator=g).to(DEV)
  x, y = ids[:, :2048].contiguous(), ids[:, 1:].contiguous()
  rt.flat.zero_grads()
  l1 = rt.loss_and_backward(x, y, None)
  gn = rt.flat.grads.double().norm().item()
  # random init on uniform tokens: logits ~ N(0, d * 0.02^2) => loss ~ ln V + d * 0.02^2 / 2 = 10.825 + 0.205
  assert abs(l1.item() - (math.log(50280) + 1024 * 0.02**2 / 2)) < 0.1
  assert math.isfinite(gn) and gn > 0
  # forward is deterministic; a single full-length document == plain causal; gradients are linear in grad_scale
  l2 = rt.loss_and_backward(x, y, torch.zeros(2 * 2048, dtype=torch.int32, device=DEV), backward=False)
  assert l1.item() == l2.item()
  rt.flat.zero_grads()
  rt.loss_and_backward(x, y, None, grad_scale=0.5)
  gn_half = rt.flat.grads.double().norm().item()
  assert abs(gn_half - 0.5 * gn) <= 2e-2 * gn

"""The oracle (oracle/plainlm_oracle.py) against fixtures produced by RUNNING the reference (tests/golden/make_golden.py).
CPU only.  This is what pins the oracle; the GPU parity tests then compare the CUDA path with the oracle."""

import json
import os

import pytest
import torch

from conftest import assert_close
from oracle import plainlm_oracle as orc


@pytest.fixture(scope='module')
def comp(golden_dir):
  return torch.load(os.path.join(golden_dir, 'components.pt'))


@pytest.fixture(scope='module')
def tiny(golden_dir):
  return torch.load(os.path.join(golden_dir, 'model_tiny.pt'))


def test_rope_table_bit_exact(comp):
  table = orc.rope_table(64, 32)
  ref = comp['rope']['table']  # [1, T, 1, hd/2, 2]
  assert torch.equal(table, ref.reshape(32, 32, 2))


def test_rope_apply_bit_exact(comp):
  r = comp['rope']
  table = orc.rope_table(64, 32)
  assert torch.equal(orc.apply_rope(r['q'], table), r['rq'])
  assert torch.equal(orc.apply_rope(r['k'], table), r['rk'])


def test_rmsnorm_fwd_bwd(comp):
  f = comp['rmsnorm']
  x = f['x'].clone().requires_grad_(True)
  w = f['w'].clone().requires_grad_(True)
  y = orc.rmsnorm(x, w)
  assert torch.equal(y, f['y'])
  y.backward(f['dy'])
  assert_close(x.grad, f['dx'], 1e-6, what='dx')
  assert_close(w.grad, f['dw'], 1e-6, what='dw')


def test_glu_fwd_bwd(comp):
  f = comp['glu']
  x = f['x'].clone().requires_grad_(True)
  w1 = f['w1'].clone().requires_grad_(True)
  w2 = f['w2'].clone().requires_grad_(True)
  y = orc.glu(x, w1, w2)
  assert_close(y, f['y'], 1e-6, what='y')
  y.backward(f['dy'])
  assert_close(x.grad, f['dx'], 1e-5, what='dx')
  assert_close(w1.grad, f['dw1'], 1e-5, what='dw1')
  assert_close(w2.grad, f['dw2'], 1e-5, what='dw2')
  assert f['hidden'] == orc.glu_hidden(64)


def test_attention_fwd_bwd(comp):
  f = comp['attn']
  x = f['x'].clone().requires_grad_(True)
  wq = f['w_qkv'].clone().requires_grad_(True)
  wo = f['w_out'].clone().requires_grad_(True)
  table = orc.rope_table(64, 32)
  y = orc.attention(x, wq, wo, table, 2)
  assert_close(y, f['y'], 1e-5, what='y')
  y.backward(f['dy'])
  assert_close(x.grad, f['dx'], 1e-4, what='dx')
  assert_close(wq.grad, f['dw_qkv'], 1e-4, what='dw_qkv')
  assert_close(wo.grad, f['dw_out'], 1e-4, what='dw_out')


def test_attention_doc_masked(comp):
  f, fm = comp['attn'], comp['attn_masked']
  table = orc.rope_table(64, 32)
  mask = torch.stack([orc.mask_from_segment_starts(orc.doc_segment_starts(dl, 32)) for dl in fm['docs_lengths']])
  y = orc.attention(f['x'], f['w_qkv'], f['w_out'], table, 2, mask.unsqueeze(1))
  assert_close(y, fm['y'], 1e-5, what='masked y')


def test_doc_mask_bit_exact(golden_dir):
  d = json.load(open(os.path.join(golden_dir, 'docmask.json')))
  T = d['T']
  for case in d['cases']:
    dl = case['docs_lengths']
    ref = torch.tensor([[(row >> (T - 1 - j)) & 1 for j in range(T)] for row in case['mask_rows']], dtype=torch.bool)
    seg = orc.doc_segment_starts(dl, T)
    assert torch.equal(orc.mask_from_segment_starts(seg), ref), dl
    assert torch.equal(orc.intra_doc_causal_mask(dl, T + 1)[:T, :T], ref), dl
  assert d['bad_sum_raises'] == 'Sum of doc_boundaries does not match max_seq_length.'
  with pytest.raises(ValueError, match='Sum of doc_boundaries'):
    orc.doc_segment_starts([3, 3], T)


def test_docs_boundaries_reference_docstring_example(golden_dir):
  """The only golden vector the reference itself carries (data_prep_utils.py:37-42): it defines the `docs_lengths`
  format (per chunk, lengths summing to the chunk size) that doc_segment_starts consumes."""
  d = json.load(open(os.path.join(golden_dir, 'docmask.json')))
  ex = d['docs_boundaries_example']
  assert ex['args'] == [[10, 20, 40], 2, 30] and ex['result'] == [[10, 20], [30]]
  for rec in [ex] + d['docs_boundaries_more']:
    _, n_chunks, size = rec['args']
    for chunk in rec['result'][:n_chunks]:
      if sum(chunk) == size:  # full chunks are valid inputs of the mask builder
        seg = orc.doc_segment_starts(chunk, size - 1)
        assert seg.numel() == size - 1 and int(seg[0]) == 0


@pytest.mark.parametrize('tag,rtol', [('fp32', 2e-5), ('bf16', 2e-2)])
def test_model_forward_backward(tiny, tag, rtol):
  cfg = tiny['cfg']
  params = orc.init_params(cfg['vocab_size'], cfg['d_model'], cfg['n_layers'], cfg['n_heads'], seed=tiny['param_seed'])
  chk = float(sum(v.double().abs().sum() for v in params.values()))
  assert abs(chk - tiny['param_checksum']) < 1e-6 * tiny['param_checksum']
  assert list(params.keys()) == tiny['state_dict_keys'] == orc.param_names(cfg['n_layers'])
  p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
  T = cfg['seq_len']
  inputs, targets = orc.split_batch(tiny['ids'], T)
  logits = orc.forward(p, inputs, cfg['n_heads'], None, tag)
  loss = orc.loss_fn(logits, targets)
  ref = tiny[tag]
  assert str(logits.dtype) == ref['logits_dtype']
  assert abs(float(loss) - ref['loss']) <= (1e-5 if tag == 'fp32' else 2e-3) * ref['loss']
  assert_close(logits[:, :2, :], ref['logits_head'], rtol, what='logits')
  loss.backward()
  for k, g in ref['grads'].items():
    got = p[k].grad
    assert abs(float(got.double().norm()) - g['norm']) <= rtol * 2 * g['norm'] + 1e-9, k
    assert_close(got.flatten()[:32], g['head'], rtol * 2, atol=rtol * g['norm'] / got.numel() ** 0.5, what=k)


def test_model_doc_masked(tiny):
  cfg = tiny['cfg']
  params = orc.init_params(cfg['vocab_size'], cfg['d_model'], cfg['n_layers'], cfg['n_heads'], seed=tiny['param_seed'])
  p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
  T = cfg['seq_len']
  ref = tiny['doc_fp32']
  inputs, targets = orc.split_batch(tiny['ids'], T)
  mask = torch.stack([orc.mask_from_segment_starts(orc.doc_segment_starts(dl, T)) for dl in ref['docs_lengths']])
  logits = orc.forward(p, inputs, cfg['n_heads'], mask, 'fp32')
  loss = orc.loss_fn(logits, targets)
  assert abs(float(loss) - ref['loss']) <= 1e-5 * ref['loss']
  loss.backward()
  for k, g in ref['grads'].items():
    assert abs(float(p[k].grad.double().norm()) - g['norm']) <= 1e-4 * g['norm'] + 1e-9, k


def test_param_groups_and_decay_rule(tiny):
  decay, no_decay = tiny['param_groups']
  assert decay['weight_decay'] == 0.1 and no_decay['weight_decay'] == 0.0
  for n in decay['names']:
    assert not orc.no_decay(n)
  for n in no_decay['names']:
    assert orc.no_decay(n)
  assert 'embed_tokens.weight' in decay['names'] and 'lm_head.weight' in decay['names']


def test_optimizers_against_torch_and_reference(golden_dir):
  fx = torch.load(os.path.join(golden_dir, 'optim.pt'))
  for name in ('adamw', 'signsgd'):
    f = fx[name]
    cfg = f['cfg']
    p, n = f['p0'].clone(), f['n0'].clone()
    st = {'p': {}, 'n': {}}
    for i, (gp, gn) in enumerate(f['grads']):
      lr = cfg['lr'] * (i + 1) / 3
      gp, gn = gp.clone(), gn.clone()
      norm = orc.clip_grad_norm_([gp, gn], 1.0)
      assert abs(float(norm) - f['snaps'][i]['norm']) < 1e-5 * f['snaps'][i]['norm']
      for key, t, g, wd in (('p', p, gp, 0.1), ('n', n, gn, 0.0)):
        s = st[key]
        if name == 'adamw':
          if not s:
            s['m'], s['v'] = torch.zeros_like(t), torch.zeros_like(t)
          orc.adamw_step(t, g, s['m'], s['v'], i + 1, lr, cfg['beta1'], cfg['beta2'], 1e-8, wd)
        else:
          first = not s
          if first:
            s['m'] = torch.zeros_like(t)
          orc.signsgd_step(t, g, s['m'], first, lr, cfg['beta1'], cfg['dampening'], wd)
      assert_close(p, f['snaps'][i]['p'], 1e-6, what=f'{name} p step {i}')
      assert_close(n, f['snaps'][i]['n'], 1e-6, what=f'{name} n step {i}')
  assert fx['adamw']['state_keys'] == ['exp_avg', 'exp_avg_sq', 'step'] and fx['signsgd']['state_keys'] == ['m']


@pytest.mark.parametrize('mlp_class', ['mlp', 'mlp_relu_sq'])
@pytest.mark.parametrize('tag,rtol', [('fp32', 2e-5), ('bf16', 2e-2)])
def test_mlp_variants_forward_backward(golden_dir, mlp_class, tag, rtol):
  """MLP / MLPReluSquared (models/components.py:31-40, 59-70) through the reference model."""
  fx = torch.load(os.path.join(golden_dir, 'variants.pt'))
  ref = fx['mlp'][mlp_class]
  params = orc.init_params(256, 128, 2, 2, seed=ref['param_seed'], mlp_class=mlp_class)
  chk = float(sum(v.double().abs().sum() for v in params.values()))
  assert abs(chk - ref['param_checksum']) < 1e-6 * ref['param_checksum']
  assert {k: list(v.shape) for k, v in params.items()} == ref['state_dict_shapes']
  p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
  inputs, targets = orc.split_batch(fx['ids'], 32)
  logits = orc.forward(p, inputs, 2, None, tag, mlp_class=mlp_class)
  loss = orc.loss_fn(logits, targets)
  assert abs(float(loss) - ref[tag]['loss']) <= (1e-5 if tag == 'fp32' else 2e-3) * ref[tag]['loss']
  assert_close(logits[:, :2, :], ref[tag]['logits_head'], rtol, what='logits')
  loss.backward()
  for k, g in ref[tag]['grads'].items():
    got = p[k].grad
    assert abs(float(got.double().norm()) - g['norm']) <= rtol * 2 * g['norm'] + 1e-9, k
    assert_close(got.flatten()[:32], g['head'], rtol * 2, atol=rtol * g['norm'] / got.numel() ** 0.5, what=k)


def test_sgd_and_nadamw_against_torch(golden_dir):
  """optim/init_optim.py:23-41.  The fixture records that the reference's own 'nadamw' branch cannot
  be built on this torch (it passes fused= to NAdam), so that optimizer is pinned on torch.optim.NAdam directly."""
  fx = torch.load(os.path.join(golden_dir, 'variants.pt'))
  assert 'fused' in fx['nadamw_factory_error']
  for name in ('sgd', 'nadamw'):
    f = fx['optim'][name]
    cfg = f['cfg']
    p, n = f['p0'].clone(), f['n0'].clone()
    st = {'p': {}, 'n': {}}
    for i, (gp, gn) in enumerate(f['grads']):
      lr = cfg['lr'] * (i + 1) / 4
      gp, gn = gp.clone(), gn.clone()
      orc.clip_grad_norm_([gp, gn], 1.0)
      for key, t, g, wd in (('p', p, gp, 0.1), ('n', n, gn, 0.0)):
        s = st[key]
        if name == 'sgd':
          first = not s
          if first:
            s['buf'] = torch.zeros_like(t)
          orc.sgd_step(t, g, s['buf'], first, lr, cfg['beta1'], cfg['dampening'], wd)
        else:
          if not s:
            s['m'], s['v'], s['st'] = torch.zeros_like(t), torch.zeros_like(t), {}
          orc.nadamw_step(t, g, s['m'], s['v'], s['st'], lr, cfg['beta1'], cfg['beta2'], 1e-8, wd)
      assert_close(p, f['snaps'][i]['p'], 2e-6, what=f'{name} p step {i}')
      assert_close(n, f['snaps'][i]['n'], 2e-6, what=f'{name} n step {i}')
  assert fx['optim']['sgd']['state_keys'] == ['momentum_buffer']
  assert fx['optim']['nadamw']['state_keys'] == ['exp_avg', 'exp_avg_sq', 'mu_product', 'step']


def test_schedule_and_sampler(golden_dir):
  d = json.load(open(os.path.join(golden_dir, 'misc.json')))
  lrs = [orc.warmup_cosine_lr(t, 0.0, 3e-3, 1e-5, 2, 20) for t in range(23)]
  assert lrs == d['warmup_cosine']
  for key, parts in d['sampler'].items():
    n, w = map(int, key.split('_'))
    for r in range(w):
      assert orc.sampler_partition(n, w, r) == parts[r]


@pytest.mark.parametrize('run', ['adamw', 'signsgd', 'adamw_doc', 'adamw_noclip_nosched'])
def test_engine_loss_curves(golden_dir, run):
  """OracleTrainer.step == the reference's TorchEngine.step on CPU (fp32), micro-step by micro-step."""
  d = json.load(open(os.path.join(golden_dir, 'engine_curves.json')))[run]
  cfg = dict(d['cfg'])
  cfg['n_heads'] = 2
  params = orc.init_params(256, 128, 2, 2, seed=7)
  tr = orc.OracleTrainer(params, cfg, 'fp32')
  data = torch.tensor(d['data'])
  for i, ref_loss in enumerate(d['losses']):
    batch = {'input_ids': data[i : i + 1]}
    if cfg['intra_doc_masking']:
      batch['docs_lengths'] = [d['docs_lengths'][i]]
    loss = float(tr.step(batch))
    assert abs(loss - ref_loss) <= 2e-4 * abs(ref_loss), (i, loss, ref_loss)
    assert abs(tr.lr - d['lrs'][i]) <= 1e-12
  for k, f in d['final_params'].items():
    assert abs(float(tr.p[k].double().norm()) - f['norm']) <= 1e-4 * f['norm'], k

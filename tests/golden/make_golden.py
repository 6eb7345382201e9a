"""Generate the golden fixtures in this directory by RUNNING THE REAL REFERENCE (Niccolo-Ajroldi/plainLM).

Run in the build container only (the GPU box has no /root/reference):

    TORCHDYNAMO_DISABLE=1 PYTHONPATH=/root/reference:/root/repo python tests/golden/make_golden.py

The reference is imported unmodified from /root/reference; nothing is copied from it.  Parameters are NOT stored:
both sides rebuild them from oracle.init_params(seed) (plain torch.randn on a seeded generator) and the reference
model loads them through load_state_dict; a checksum guards against RNG drift.  Fixtures hold inputs that are cheap
to store and the reference's OUTPUTS (losses, logits slices, gradient norms/slices, masks, optimizer results).
"""

import json
import os
import sys
from collections import namedtuple

os.environ.setdefault('TORCHDYNAMO_DISABLE', '1')
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import plainlm_oracle as orc  # noqa: E402  (only for init_params: seeded randn, no model code)

torch.set_num_threads(4)

TINY = dict(vocab_size=256, d_model=128, n_layers=2, n_heads=2, seq_len=32, expand='8/3', mlp_class='glu',
            tie_embeddings=False, model='transformer')


def ref_model(cfg_dict, seed=7):
  from models import construct_model

  Cfg = namedtuple('Cfg', cfg_dict.keys())
  model, _ = construct_model(Cfg(**cfg_dict))
  params = orc.init_params(cfg_dict['vocab_size'], cfg_dict['d_model'], cfg_dict['n_layers'], cfg_dict['n_heads'],
                           seed=seed)
  model.load_state_dict(params, strict=True)
  return model, params


def checksum(params):
  return float(sum(v.double().abs().sum() for v in params.values()))


def grad_summary(named_grads):
  out = {}
  for k, g in named_grads.items():
    out[k] = {'norm': float(g.double().norm()), 'head': g.flatten()[:32].clone(), 'sum': float(g.double().sum())}
  return out


def gen_model_fixture():
  torch.manual_seed(0)
  model, params = ref_model(TINY)
  B, T, V = 2, TINY['seq_len'], TINY['vocab_size']
  ids = torch.randint(0, V, (B, T + 1), generator=torch.Generator().manual_seed(11))
  inputs, targets = ids[:, :T], ids[:, 1 : T + 1].contiguous()
  fix = {'cfg': TINY, 'param_seed': 7, 'param_checksum': checksum(params), 'ids': ids,
         'state_dict_keys': list(model.state_dict().keys()),
         'state_dict_shapes': {k: list(v.shape) for k, v in model.state_dict().items()},
         'count_params': [model.count_params(False), model.count_params(True)]}
  for tag, ctx in (('fp32', None), ('bf16', torch.autocast('cpu', dtype=torch.bfloat16))):
    model.zero_grad(set_to_none=True)
    if ctx is None:
      logits = model(inputs, None)
      loss = torch.nn.CrossEntropyLoss()(logits.view(-1, V), targets.view(-1))
    else:
      with ctx:
        logits = model(inputs, None)
        loss = torch.nn.CrossEntropyLoss()(logits.view(-1, V), targets.view(-1))
    loss.backward()
    fix[tag] = {'loss': float(loss), 'logits_dtype': str(logits.dtype), 'logits_head': logits[:, :2, :].float().clone(),
                'logits_sum': float(logits.double().sum()),
                'grads': grad_summary({k: p.grad for k, p in model.named_parameters()})}
  # document-masked forward/backward (fp32) through the reference's own mask builder
  from data.datasets.data_prep_utils import intra_doc_causal_mask

  docs = [[5, 7, 21], [33]]
  masks = torch.stack([intra_doc_causal_mask(dl, T + 1) for dl in docs])[:, :T, :T].contiguous()
  model.zero_grad(set_to_none=True)
  logits = model(inputs, masks)
  loss = torch.nn.CrossEntropyLoss()(logits.view(-1, V), targets.view(-1))
  loss.backward()
  fix['doc_fp32'] = {'docs_lengths': docs, 'loss': float(loss), 'logits_head': logits[:, :2, :].float().clone(),
                     'logits_sum': float(logits.double().sum()),
                     'grads': grad_summary({k: p.grad for k, p in model.named_parameters()})}
  # param groups (models/construct.py:47-75)
  from models import get_param_groups

  groups = get_param_groups(model, 0.1)
  ids_to_name = {id(p): n for n, p in model.named_parameters()}
  fix['param_groups'] = [{'weight_decay': g['weight_decay'], 'names': [ids_to_name[id(p)] for p in g['params']]}
                         for g in groups]
  torch.save(fix, os.path.join(HERE, 'model_tiny.pt'))
  print('model_tiny.pt', {k: fix[k]['loss'] for k in ('fp32', 'bf16', 'doc_fp32')})


def gen_components_fixture():
  """Per-block inputs/outputs/grads from the reference modules."""
  from models.components import RMSNorm, GLU
  from models.embeddings import precompute_freqs_cis, apply_rotary_emb_complex_like
  from models.transformer import Attention, ModelConfig

  g = torch.Generator().manual_seed(3)
  fix = {}
  # RMSNorm
  x = torch.randn(6, 128, generator=g, requires_grad=True)
  norm = RMSNorm(128, 1e-6)
  with torch.no_grad():
    norm.weight.copy_(torch.rand(128, generator=g) + 0.5)
  y = norm(x)
  dy = torch.randn(6, 128, generator=g)
  y.backward(dy)
  fix['rmsnorm'] = {'x': x.detach().clone(), 'w': norm.weight.detach().clone(), 'dy': dy, 'y': y.detach().clone(),
                    'dx': x.grad.clone(), 'dw': norm.weight.grad.clone()}
  # RoPE
  table = precompute_freqs_cis(64, 32, 500000)[0:32]
  q = torch.randn(2, 32, 2, 64, generator=g)
  k = torch.randn(2, 32, 2, 64, generator=g)
  rq, rk = apply_rotary_emb_complex_like(q, k, freqs_cis=table)
  fix['rope'] = {'table': table.clone(), 'q': q, 'k': k, 'rq': rq.clone(), 'rk': rk.clone()}
  # GLU
  glu = GLU(64, int(8 / 3 * 64))
  xg = torch.randn(2, 8, 64, generator=g, requires_grad=True)
  yg = glu(xg)
  dyg = torch.randn(2, 8, 64, generator=g)
  yg.backward(dyg)
  fix['glu'] = {'x': xg.detach().clone(), 'w1': glu.fc1.weight.detach().clone(), 'w2': glu.fc2.weight.detach().clone(),
                'dy': dyg, 'y': yg.detach().clone(), 'dx': xg.grad.clone(), 'dw1': glu.fc1.weight.grad.clone(),
                'dw2': glu.fc2.weight.grad.clone(), 'hidden': glu.hidden_dim}
  # Attention (causal and masked)
  mc = ModelConfig(vocab_size=16, seq_len=32, dim=128, expand=8 / 3, n_layers=1, n_heads=2, mlp='glu')
  att = Attention(mc)
  xa = torch.randn(2, 32, 128, generator=g, requires_grad=True)
  ya = att(xa, table, None)
  dya = torch.randn(2, 32, 128, generator=g)
  ya.backward(dya)
  fix['attn'] = {'x': xa.detach().clone(), 'w_qkv': att.w_qkv.weight.detach().clone(),
                 'w_out': att.w_out.weight.detach().clone(), 'dy': dya, 'y': ya.detach().clone(),
                 'dx': xa.grad.clone(), 'dw_qkv': att.w_qkv.weight.grad.clone(),
                 'dw_out': att.w_out.weight.grad.clone()}
  from data.datasets.data_prep_utils import intra_doc_causal_mask

  docs = [[10, 3, 20], [1, 31, 1]]
  mask = torch.stack([intra_doc_causal_mask(dl, 33) for dl in docs])[:, :32, :32].contiguous()
  with torch.no_grad():
    ym = att(xa, table, mask)
  fix['attn_masked'] = {'docs_lengths': docs, 'y': ym.clone()}
  torch.save(fix, os.path.join(HERE, 'components.pt'))
  print('components.pt ok')


def gen_docmask_fixture():
  """data/datasets/data_prep_utils.py:7-23 on the edge cases of SURVEY.md §7.3, cropped as engine.py:23."""
  from data.datasets.data_prep_utils import intra_doc_causal_mask, _get_docs_boundaries

  T = 16
  cases = [[5, 7, 5], [17], [16, 1], [1, 16], [1] * 17, [2, 15], [8, 8, 1]]
  out = {'T': T, 'cases': []}
  for dl in cases:
    m = intra_doc_causal_mask(dl, T + 1)[:T, :T].contiguous()
    out['cases'].append({'docs_lengths': dl, 'mask_rows': [int(''.join('1' if v else '0' for v in row), 2)
                                                           for row in m.tolist()]})
  try:
    intra_doc_causal_mask([3, 3], T + 1)
    out['bad_sum_raises'] = False
  except ValueError as e:
    out['bad_sum_raises'] = str(e)
  # the reference's only golden vector (docstring, data_prep_utils.py:37-42)
  out['docs_boundaries_example'] = {'args': [[10, 20, 40], 2, 30], 'result': _get_docs_boundaries([10, 20, 40], 2, 30)}
  out['docs_boundaries_more'] = [{'args': [dl, n, m], 'result': _get_docs_boundaries(dl, n, m)}
                                 for dl, n, m in ([[3, 3, 3, 3], 2, 6], [[50], 3, 16], [[1, 2, 3, 4, 5, 6], 4, 5])]
  with open(os.path.join(HERE, 'docmask.json'), 'w') as f:
    json.dump(out, f)
  print('docmask.json ok')


def make_cfg(**over):
  base = dict(
    seq_len=32, grad_accumulation_steps=2, grad_clip=1.0, dtype='float32', intra_doc_masking=False, resume=False,
    torch_compile=False, weight_decay=0.1, optim='adamw', lr=3e-3, beta1=0.9, beta2=0.95, fused_optim=False,
    scheduler='warmup_cosine', warmup_steps=0.1, cooldown_steps=None, lr_start=0.0, lr_end=1e-5, lr_end_pct=None,
    steps_budget=20, dampening=0.0,
  )  # fmt: skip
  base.update(over)
  return base


def gen_engine_fixture():
  """Loss curves from the reference's TorchEngine.step on CPU (fp32; micro_batch 1)."""
  from engine import TorchEngine

  out = {}
  V, T = TINY['vocab_size'], TINY['seq_len']
  runs = {
    'adamw': make_cfg(),
    'signsgd': make_cfg(optim='signSGD', lr=1e-3, dampening=0.0, steps_budget=10),
    'adamw_doc': make_cfg(intra_doc_masking=True, steps_budget=6),
    'adamw_noclip_nosched': make_cfg(grad_clip=None, scheduler=None, steps_budget=5, lr=1e-3),
  }
  for name, cfgd in runs.items():
    model, params = ref_model(TINY)
    Cfg = namedtuple('Cfg', cfgd.keys())
    eng = TorchEngine(model, Cfg(**cfgd), 'cpu', None, None)
    n_micro = cfgd['steps_budget'] * cfgd['grad_accumulation_steps']
    gen = torch.Generator().manual_seed(1234)
    # low-entropy stream so the loss actually descends: order-1 Markov chain on 16 symbols
    trans = torch.softmax(torch.randn(16, 16, generator=gen) * 3, dim=-1)
    rows = []
    for _ in range(n_micro):
      seq = [int(torch.randint(0, 16, (1,), generator=gen))]
      for _t in range(T):
        seq.append(int(torch.multinomial(trans[seq[-1]], 1, generator=gen)))
      rows.append(seq)
    data = torch.tensor(rows, dtype=torch.int64)
    rng_docs = torch.Generator().manual_seed(5)
    losses, lrs, docs_all = [], [], []
    for i in range(n_micro):
      batch = {'input_ids': data[i : i + 1]}
      if cfgd['intra_doc_masking']:
        cut = sorted(set(torch.randint(1, T + 1, (2,), generator=rng_docs).tolist()))
        edges = [0] + cut + [T + 1]
        dl = [edges[j + 1] - edges[j] for j in range(len(edges) - 1)]
        batch['docs_lengths'] = [dl]
        docs_all.append(dl)
      losses.append(float(eng.step(batch)))
      lrs.append(float(eng.optimizer.param_groups[0]['lr']))
    final = {k: {'norm': float(v.double().norm()), 'head': v.detach().flatten()[:8].tolist()}
             for k, v in model.state_dict().items()}
    out[name] = {'cfg': cfgd, 'data': data.tolist(), 'docs_lengths': docs_all, 'losses': losses, 'lrs': lrs,
                 'final_params': final}
    print(name, losses[0], '->', losses[-1])
  with open(os.path.join(HERE, 'engine_curves.json'), 'w') as f:
    json.dump(out, f)


def gen_optim_fixture():
  """torch.optim.AdamW as built by optim/init_optim.py and the reference's signSGD, 3 steps with clipping."""
  from optim import intialize_optimizer

  g = torch.Generator().manual_seed(9)
  out = {}
  for name, cfgd in (('adamw', make_cfg()), ('signsgd', make_cfg(optim='signSGD', lr=1e-3, dampening=0.1))):
    p0 = torch.randn(300, generator=g)
    n0 = torch.rand(20, generator=g) + 0.5
    grads = [(torch.randn(300, generator=g) * 2, torch.randn(20, generator=g)) for _ in range(3)]
    p = torch.nn.Parameter(p0.clone())
    n = torch.nn.Parameter(n0.clone())
    Cfg = namedtuple('Cfg', cfgd.keys())
    opt = intialize_optimizer([{'params': [p], 'weight_decay': 0.1}, {'params': [n], 'weight_decay': 0.0}],
                              Cfg(**cfgd))
    snaps = []
    for i, (gp, gn) in enumerate(grads):
      for grp in opt.param_groups:
        grp['lr'] = cfgd['lr'] * (i + 1) / 3
      p.grad, n.grad = gp.clone(), gn.clone()
      norm = torch.nn.utils.clip_grad_norm_([p, n], 1.0)
      opt.step()
      snaps.append({'p': p.detach().clone(), 'n': n.detach().clone(), 'norm': float(norm)})
    state_keys = sorted(opt.state[p].keys())
    out[name] = {'cfg': cfgd, 'p0': p0, 'n0': n0, 'grads': grads, 'snaps': snaps, 'state_keys': state_keys}
  torch.save(out, os.path.join(HERE, 'optim.pt'))
  print('optim.pt ok', out['adamw']['state_keys'], out['signsgd']['state_keys'])


def gen_init_fixture():
  """Weights the reference draws for seed 100 (config/*.yaml `seed: 100`): same constructors + same registration
  order must give the same tensors in plainlm_b200.models.Transformer."""
  from models import construct_model

  Cfg = namedtuple('Cfg', TINY.keys())
  torch.manual_seed(100)
  model, _ = construct_model(Cfg(**TINY))
  sd = model.state_dict()
  out = {k: {'sum': float(v.double().sum()), 'abs': float(v.double().abs().sum()), 'head': v.flatten()[:8].tolist()}
         for k, v in sd.items()}
  with open(os.path.join(HERE, 'init_seed100.json'), 'w') as f:
    json.dump(out, f)
  print('init_seed100.json ok')


def gen_misc_fixture():
  from optim.lr_schedule import WarmupCosine
  from torch.utils.data import DistributedSampler

  class _Opt:
    param_groups = [{'lr': None}]

  sch = WarmupCosine(_Opt(), lr_start=0.0, lr_max=3e-3, lr_end=1e-5, warmup_steps=2, T=20)
  lrs = [_Opt.param_groups[0]['lr']]
  for _ in range(22):
    sch.step()
    lrs.append(_Opt.param_groups[0]['lr'])
  parts = {}
  for n, w in ((21, 4), (16, 2), (7, 8), (100, 8)):
    ds = list(range(n))
    parts[f'{n}_{w}'] = [list(DistributedSampler(ds, num_replicas=w, rank=r, shuffle=False, drop_last=True))
                         for r in range(w)]
  with open(os.path.join(HERE, 'misc.json'), 'w') as f:
    json.dump({'warmup_cosine': lrs, 'sampler': parts}, f)
  print('misc.json ok')


def gen_variants_fixture():
  """the other MLP classes (models/components.py:31-40, 59-70) through the reference model, and the
  other optimizers of optim/init_optim.py (sgd through the reference factory; nadamw directly through
  torch.optim.NAdam(decoupled_weight_decay=True) with the factory's arguments, because the factory also passes
  `fused=`, which torch 2.11's NAdam rejects — the reference's 'nadamw' branch raises TypeError on this torch)."""
  from optim import intialize_optimizer

  out = {'mlp': {}, 'optim': {}}
  B, T, V = 2, TINY['seq_len'], TINY['vocab_size']
  ids = torch.randint(0, V, (B, T + 1), generator=torch.Generator().manual_seed(12))
  inputs, targets = ids[:, :T], ids[:, 1 : T + 1].contiguous()
  out['ids'] = ids
  for mlp_class in ('mlp', 'mlp_relu_sq'):
    from models import construct_model

    cfgd = dict(TINY, mlp_class=mlp_class)
    Cfg = namedtuple('Cfg', cfgd.keys())
    model, _ = construct_model(Cfg(**cfgd))
    params = orc.init_params(V, cfgd['d_model'], cfgd['n_layers'], cfgd['n_heads'], seed=8, mlp_class=mlp_class)
    model.load_state_dict(params, strict=True)
    rec = {'param_seed': 8, 'param_checksum': checksum(params),
           'state_dict_shapes': {k: list(v.shape) for k, v in model.state_dict().items()}}
    for tag in ('fp32', 'bf16'):
      model.zero_grad(set_to_none=True)
      if tag == 'fp32':
        logits = model(inputs, None)
        loss = torch.nn.CrossEntropyLoss()(logits.view(-1, V), targets.view(-1))
      else:
        with torch.autocast('cpu', dtype=torch.bfloat16):
          logits = model(inputs, None)
          loss = torch.nn.CrossEntropyLoss()(logits.view(-1, V), targets.view(-1))
      loss.backward()
      rec[tag] = {'loss': float(loss), 'logits_head': logits[:, :2, :].float().clone(),
                  'logits_sum': float(logits.double().sum()),
                  'grads': grad_summary({k: p.grad for k, p in model.named_parameters()})}
    out['mlp'][mlp_class] = rec
  g = torch.Generator().manual_seed(10)
  for name, cfgd in (('sgd', make_cfg(optim='sgd', lr=1e-2, dampening=0.1)), ('nadamw', make_cfg(optim='nadamw'))):
    p0 = torch.randn(300, generator=g)
    n0 = torch.rand(20, generator=g) + 0.5
    grads = [(torch.randn(300, generator=g) * 2, torch.randn(20, generator=g)) for _ in range(4)]
    p = torch.nn.Parameter(p0.clone())
    n = torch.nn.Parameter(n0.clone())
    groups = [{'params': [p], 'weight_decay': 0.1}, {'params': [n], 'weight_decay': 0.0}]
    Cfg = namedtuple('Cfg', cfgd.keys())
    if name == 'sgd':
      opt = intialize_optimizer(groups, Cfg(**cfgd))
    else:
      try:
        intialize_optimizer([{'params': [torch.nn.Parameter(torch.zeros(1))], 'weight_decay': 0.0}], Cfg(**cfgd))
        out['nadamw_factory_error'] = None
      except TypeError as e:
        out['nadamw_factory_error'] = str(e)
      opt = torch.optim.NAdam(groups, lr=cfgd['lr'], betas=[cfgd['beta1'], cfgd['beta2']],
                              weight_decay=cfgd['weight_decay'], decoupled_weight_decay=True, eps=1e-8)
    snaps = []
    for i, (gp, gn) in enumerate(grads):
      for grp in opt.param_groups:
        grp['lr'] = cfgd['lr'] * (i + 1) / 4
      p.grad, n.grad = gp.clone(), gn.clone()
      norm = torch.nn.utils.clip_grad_norm_([p, n], 1.0)
      opt.step()
      snaps.append({'p': p.detach().clone(), 'n': n.detach().clone(), 'norm': float(norm)})
    out['optim'][name] = {'cfg': cfgd, 'p0': p0, 'n0': n0, 'grads': grads, 'snaps': snaps,
                          'state_keys': sorted(opt.state[p].keys())}
  torch.save(out, os.path.join(HERE, 'variants.pt'))
  print('variants.pt', {k: (v['fp32']['loss'], v['bf16']['loss']) for k, v in out['mlp'].items()},
        {k: v['state_keys'] for k, v in out['optim'].items()}, out['nadamw_factory_error'])


MICRO = dict(vocab_size=64, d_model=128, n_layers=1, n_heads=2, seq_len=32, expand='1', mlp_class='glu',
             tie_embeddings=False, model='transformer')


def _markov_rows(n_rows, T, seed, symbols=16):
  gen = torch.Generator().manual_seed(seed)
  trans = torch.softmax(torch.randn(symbols, symbols, generator=gen) * 3, dim=-1)
  rows = []
  for _ in range(n_rows):
    seq = [int(torch.randint(0, symbols, (1,), generator=gen))]
    for _t in range(T):
      seq.append(int(torch.multinomial(trans[seq[-1]], 1, generator=gen)))
    rows.append(seq)
  return torch.tensor(rows, dtype=torch.int64)


def gen_resume_fixture():
  """(1) A checkpoint WRITTEN BY THE REFERENCE: its TorchEngine trains a micro model for 3 optimizer steps on the CPU
  and its own checkpoint_utils.save_checkpoint writes ckpt_step_3.pth (fused AdamW, so `step` is a tensor exactly as on
  the GPU); the reference then keeps training and the next losses are recorded — SURVEY §8(f) N3 "cross-loading
  reference checkpoints".  (2) The reference's loss curve with tie_embeddings=True (N4)."""
  import shutil
  import tempfile

  from absl import flags

  flags.DEFINE_integer('job_idx', None, 'stand-in for train.py:15 (utils.get_exp_dir_path reads it)')
  flags.FLAGS(['make_golden'])
  import checkpoint_utils
  from engine import TorchEngine
  from models import construct_model

  T = MICRO['seq_len']
  out = {'cfg': MICRO}
  tmp = tempfile.mkdtemp()
  cfgd = make_cfg(grad_accumulation_steps=1, fused_optim=True, steps_budget=12, out_dir=tmp, exp_name='golden',
                  save_optim=True, save_scheduler=True, save_scaler=True)
  Cfg = namedtuple('Cfg', cfgd.keys())
  torch.manual_seed(31)
  model, _ = construct_model(namedtuple('M', MICRO.keys())(**MICRO))
  out['init_state_dict'] = {k: v.detach().clone() for k, v in model.state_dict().items()}
  eng = TorchEngine(model, Cfg(**cfgd), 'cpu', None, None)
  data = _markov_rows(8, T, seed=77)
  losses = []
  for i in range(3):
    losses.append(float(eng.step({'input_ids': data[i : i + 1]})))
  os.makedirs(os.path.join(tmp, 'golden'), exist_ok=True)
  checkpoint_utils.save_checkpoint(3, model, eng, Cfg(**cfgd), {'train/loss': losses})
  shutil.copy(os.path.join(tmp, 'golden', 'ckpt_step_3.pth'), os.path.join(HERE, 'ref_ckpt_step_3.pth'))
  after = [float(eng.step({'input_ids': data[i : i + 1]})) for i in range(3, 8)]
  # the reference resuming from its own file reproduces `after` (sanity of the fixture itself)
  ck = torch.load(os.path.join(HERE, 'ref_ckpt_step_3.pth'), map_location='cpu')
  model2, _ = construct_model(namedtuple('M', MICRO.keys())(**MICRO))
  eng2 = TorchEngine(model2, Cfg(**dict(cfgd, resume=True)), 'cpu', None, ck)
  again = [float(eng2.step({'input_ids': data[i : i + 1]})) for i in range(3, 8)]
  assert max(abs(a - b) for a, b in zip(after, again)) < 1e-5, (after, again)
  out['resume'] = {'engine_cfg': {k: v for k, v in cfgd.items() if k not in ('out_dir', 'exp_name')},
                   'data': data.tolist(), 'losses_before': losses, 'losses_after': after,
                   'optimizer_state_keys': sorted(ck['optimizer']['state'][0].keys()),
                   'step_dtype': str(ck['optimizer']['state'][0]['step'].dtype),
                   'ckpt_keys': sorted(ck.keys())}
  shutil.rmtree(tmp)

  tied_cfg = dict(MICRO, tie_embeddings=True)
  torch.manual_seed(32)
  modelt, _ = construct_model(namedtuple('M', tied_cfg.keys())(**tied_cfg))
  assert modelt.lm_head.weight is modelt.embed_tokens.weight
  cfgt = make_cfg(grad_accumulation_steps=2, steps_budget=8)
  init_t = {k: v.detach().clone() for k, v in modelt.state_dict().items()}
  engt = TorchEngine(modelt, namedtuple('Cfg', cfgt.keys())(**cfgt), 'cpu', None, None)
  datat = _markov_rows(16, T, seed=78)
  lt = [float(engt.step({'input_ids': datat[i : i + 1]})) for i in range(16)]
  out['tied'] = {'cfg': tied_cfg, 'engine_cfg': cfgt, 'init_state_dict': init_t, 'data': datat.tolist(), 'losses': lt,
                 'final_embed_norm': float(modelt.embed_tokens.weight.double().norm())}
  torch.save(out, os.path.join(HERE, 'resume_tied.pt'))
  print('resume', losses, after, '| tied', lt[0], '->', lt[-1], out['resume']['optimizer_state_keys'],
        out['resume']['step_dtype'])


if __name__ == '__main__':
  if len(sys.argv) > 1 and sys.argv[1] == 'variants':
    gen_variants_fixture()
    sys.exit(0)
  if len(sys.argv) > 1 and sys.argv[1] == 'resume':
    gen_resume_fixture()
    sys.exit(0)
  gen_model_fixture()
  gen_components_fixture()
  gen_docmask_fixture()
  gen_engine_fixture()
  gen_optim_fixture()
  gen_misc_fixture()
  gen_init_fixture()
  gen_variants_fixture()
  gen_resume_fixture()

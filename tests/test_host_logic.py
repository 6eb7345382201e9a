"""CPU-only checks of the host-side logic and of the drop-in boundary (no kernel launches)."""

import ctypes
import json
import os
import re
from collections import namedtuple

import pytest
import torch

from oracle import plainlm_oracle as orc
from plainlm_b200 import _lib, data_utils, ops
from plainlm_b200.models import construct_model, get_param_groups
from plainlm_b200.models.runtime import _bucket_param_names
from plainlm_b200.optim import initialize_scheduler, intialize_optimizer
from plainlm_b200.optim import lr_schedule

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TINY = dict(vocab_size=256, d_model=128, n_layers=2, n_heads=2, seq_len=32, expand='8/3', mlp_class='glu',
            tie_embeddings=False, model='transformer')


def _cfg(**kw):
  return namedtuple('Cfg', kw.keys())(**kw)


# ------------------------------------------------------------------------------------------- C ABI
def _header_functions():
  src = open(os.path.join(ROOT, 'include', 'plainlm_b200.h')).read()
  src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
  return sorted(set(re.findall(r'\b(plm_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
  names = _header_functions()
  assert len(names) >= 20
  lib = ctypes.CDLL(_lib.LIB_PATH)
  for n in names:
    assert hasattr(lib, n), f'{n} declared in include/plainlm_b200.h but not exported'
  assert sorted(_lib.SIGNATURES) == names, 'python binding and header disagree on the function set'
  assert _lib.load().plm_abi_version() == 5


def test_graft_entry_build_passes():
  """The driver's "does it build" check: incremental make + ABI version + every symbol (no GPU needed)."""
  import __graft_entry__ as g

  g.build()


def test_no_torch_types_in_abi():
  src = open(os.path.join(ROOT, 'include', 'plainlm_b200.h')).read()
  code = re.sub(r'/\*.*?\*/', '', src, flags=re.S)  # comments may mention PyTorch; declarations may not
  assert 'torch' not in code.lower() and 'at::' not in code and 'Tensor' not in code
  assert re.findall(r'#include\s*[<"]([^>"]+)[>"]', code) == ['stdint.h']


def test_missing_library_fails_loudly(monkeypatch):
  monkeypatch.setattr(_lib, '_lib', None)
  monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libplainlm_b200.so')
  with pytest.raises(RuntimeError, match='no fallback'):
    _lib.load()


def test_cpu_tensors_are_rejected():
  x = torch.randn(4, 128)
  with pytest.raises(RuntimeError, match='no CPU path'):
    ops.rmsnorm_fwd(x, torch.ones(128), torch.empty(4, 128, dtype=torch.bfloat16), torch.empty(4), 1e-6)
  model, _ = construct_model(_cfg(**TINY))
  with pytest.raises(RuntimeError, match='no CPU path'):
    model(torch.zeros(1, 32, dtype=torch.long), None)


def test_product_never_imports_the_oracle():
  bad = []
  for base in ('plainlm_b200', 'dropin'):
    for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
      for f in files:
        if f.endswith(('.py', '.cu', '.cuh', '.h')):
          txt = open(os.path.join(dirpath, f)).read()
          if re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M) or 'plainlm_oracle' in txt:
            bad.append(os.path.join(dirpath, f))
  assert not bad, bad


# ------------------------------------------------------------------------------------------- integers: bit-exact
def test_seg_start_matches_reference_masks(golden_dir):
  d = json.load(open(os.path.join(golden_dir, 'docmask.json')))
  T = d['T']
  for case in d['cases']:
    ref = torch.tensor([[(row >> (T - 1 - j)) & 1 for j in range(T)] for row in case['mask_rows']], dtype=torch.bool)
    seg = data_utils.seg_start_from_docs_lengths([case['docs_lengths']], T)[0]
    assert seg.dtype == torch.int32
    i = torch.arange(T)
    mask = (i[None, :] <= i[:, None]) & (i[None, :] >= seg.long()[:, None])
    assert torch.equal(mask, ref), case['docs_lengths']
    assert torch.equal(seg, orc.doc_segment_starts(case['docs_lengths'], T))
  with pytest.raises(ValueError, match='Sum of doc_boundaries does not match max_seq_length'):
    data_utils.seg_start_from_docs_lengths([[3, 3]], T)


def test_split_inputs_targets():
  ids = torch.arange(2 * 40).reshape(2, 40)
  x, y = data_utils.split_inputs_targets(ids, 32)
  assert torch.equal(x, ids[:, :32]) and torch.equal(y, ids[:, 1:33])


def test_rank_partition(golden_dir):
  d = json.load(open(os.path.join(golden_dir, 'misc.json')))
  for key, parts in d['sampler'].items():
    n, w = map(int, key.split('_'))
    for r in range(w):
      assert data_utils.rank_partition(n, w, r) == parts[r]
  # W ranks x accum micro-steps see exactly the rows 1 rank x (W*accum) micro-steps sees
  n, w, B, accum = 64, 4, 2, 2
  per_rank = [data_utils.rank_partition(n, w, r) for r in range(w)]
  dp_rows = sorted(row for r in range(w) for row in per_rank[r][: B * accum])
  assert dp_rows == list(range(w * B * accum))


# ------------------------------------------------------------------------------------------- model boundary
def test_state_dict_and_counts(golden_dir):
  fx = torch.load(os.path.join(golden_dir, 'model_tiny.pt'))
  model, mcfg = construct_model(_cfg(**TINY))
  sd = model.state_dict()
  assert list(sd.keys()) == fx['state_dict_keys']
  assert {k: list(v.shape) for k, v in sd.items()} == fx['state_dict_shapes']
  assert all(v.dtype == torch.float32 for v in sd.values())
  assert [model.count_params(False), model.count_params(True)] == fx['count_params']
  assert mcfg.dim == 128 and abs(mcfg.expand - 8 / 3) < 1e-12
  assert model.freqs_cis.shape == (1, 32, 1, 32, 2) and 'freqs_cis' not in sd
  assert model.layers[0].mlp.hidden_dim == 512


def test_init_matches_reference_rng_stream(golden_dir):
  ref = json.load(open(os.path.join(golden_dir, 'init_seed100.json')))
  torch.manual_seed(100)
  model, _ = construct_model(_cfg(**TINY))
  for k, v in model.state_dict().items():
    assert v.flatten()[:8].tolist() == ref[k]['head'], k
    assert abs(float(v.double().sum()) - ref[k]['sum']) <= 1e-9 * max(1.0, ref[k]['abs']), k


def test_param_groups_match_reference(golden_dir):
  fx = torch.load(os.path.join(golden_dir, 'model_tiny.pt'))
  model, _ = construct_model(_cfg(**TINY))
  names = {id(p): n for n, p in model.named_parameters()}
  groups = get_param_groups(model, 0.1)
  for g, ref in zip(groups, fx['param_groups']):
    assert g['weight_decay'] == ref['weight_decay']
    assert [names[id(p)] for p in g['params']] == ref['names']


def test_tied_embeddings_share_storage():
  cfg = dict(TINY, tie_embeddings=True)
  model, _ = construct_model(_cfg(**cfg))
  assert model.lm_head.weight is model.embed_tokens.weight
  assert model.count_params(True) == model.count_params(False) - model.embed_tokens.weight.numel()
  assert ['lm_head.weight'] not in _bucket_param_names(model)


def test_unsupported_variants_say_so():
  from plainlm_b200.optim import intialize_optimizer

  with pytest.raises(NotImplementedError):
    construct_model(_cfg(**dict(TINY, model='pythia-160m')))
  p = torch.nn.Parameter(torch.zeros(8))
  with pytest.raises(NotImplementedError, match='schedulefree'):
    intialize_optimizer([{'params': [p], 'weight_decay': 0.0}],
                        _cfg(optim='sfo_adamw', lr=1e-3, beta1=0.9, beta2=0.95, weight_decay=0.1))


def test_mlp_variants_keep_reference_shapes(golden_dir):
  """MLP / MLPReluSquared construct with the reference's parameter names and shapes."""
  fx = torch.load(os.path.join(golden_dir, 'variants.pt'))
  for mlp_class in ('mlp', 'mlp_relu_sq'):
    model, _ = construct_model(_cfg(**dict(TINY, mlp_class=mlp_class)))
    assert {k: list(v.shape) for k, v in model.state_dict().items()} == fx['mlp'][mlp_class]['state_dict_shapes']


def test_optimizer_factory_variants():
  from plainlm_b200.optim import intialize_optimizer

  p = torch.nn.Parameter(torch.zeros(8))
  base = dict(lr=1e-3, beta1=0.9, beta2=0.95, weight_decay=0.1, dampening=0.1)
  sgd = intialize_optimizer([{'params': [p], 'weight_decay': 0.0}], _cfg(optim='sgd', **base))
  assert sorted(sgd.param_groups[0]) == sorted(torch.optim.SGD([p], lr=1e-3, momentum=0.9).param_groups[0].keys() &
                                               sgd.param_groups[0].keys())
  assert sgd.param_groups[0]['momentum'] == 0.9 and sgd.param_groups[0]['dampening'] == 0.1
  nad = intialize_optimizer([{'params': [p], 'weight_decay': 0.0}], _cfg(optim='nadamw', **base))
  g = nad.param_groups[0]
  assert g['betas'] == (0.9, 0.95) and g['momentum_decay'] == 4e-3 and g['decoupled_weight_decay'] is True


def test_bucket_order_is_backward_completion_order():
  model, _ = construct_model(_cfg(**TINY))
  b = _bucket_param_names(model)
  assert b[0] == ['lm_head.weight']
  assert b[1][0].startswith('layers.1.') and b[2][0].startswith('layers.0.')
  assert b[3] == ['embed_tokens.weight']
  assert all('norm' in n for n in b[4]) and len(b[4]) == 2 * 2 + 1
  flat = [n for names in b for n in names]
  assert sorted(flat) == sorted(n for n, _ in model.named_parameters())


# ------------------------------------------------------------------------------------------- optimizer / schedule
def _train_cfg(**over):
  base = dict(optim='adamw', lr=3e-3, beta1=0.9, beta2=0.95, weight_decay=0.1, fused_optim=True, scheduler='warmup_cosine',
              warmup_steps=0.1, cooldown_steps=None, lr_start=0.0, lr_end=1e-5, lr_end_pct=None, steps_budget=20,
              dampening=0.0, resume_step=None)
  base.update(over)
  return _cfg(**base)


def test_optimizer_factory_and_schedule_coupling(golden_dir):
  p = torch.nn.Parameter(torch.zeros(4))
  opt = intialize_optimizer([{'params': [p], 'weight_decay': 0.1}], _train_cfg())
  assert isinstance(opt, torch.optim.Optimizer) and opt.param_groups[0]['betas'] == (0.9, 0.95)
  assert opt.param_groups[0]['eps'] == 1e-8 and opt.param_groups[0]['weight_decay'] == 0.1
  sch = initialize_scheduler(opt, _train_cfg())
  assert opt.param_groups[0]['lr'] == 0.0  # optimizer step 1 runs at lr_start (SURVEY Q-a)
  ref = json.load(open(os.path.join(golden_dir, 'misc.json')))['warmup_cosine']
  lrs = [opt.param_groups[0]['lr']]
  for _ in range(22):
    sch.step()
    lrs.append(opt.param_groups[0]['lr'])
  assert lrs == ref
  assert sorted(sch.state_dict()) == ['T', 'iter', 'lr_end', 'lr_max', 'lr_start', 'warmup_steps']
  opt2 = intialize_optimizer([{'params': [p], 'weight_decay': 0.1}], _train_cfg(optim='signSGD', dampening=0.1))
  assert opt2.param_groups[0]['momentum'] == 0.9 and opt2.param_groups[0]['dampening'] == 0.1
  with pytest.raises(NotImplementedError, match='N4'):
    intialize_optimizer([{'params': [p]}], _train_cfg(optim='sfo_adamw'))
  assert initialize_scheduler(opt, _train_cfg(scheduler=None)) is None


def test_other_schedules():
  class O:
    def __init__(self):
      self.param_groups = [{'lr': None}]

  wsd = lr_schedule.WSD(O(), 0.0, 1.0, 0.1, warmup_steps=2, cooldown_start_step=6, cooldown_steps=4)
  vals = [wsd.get_lr(t) for t in range(11)]
  assert vals[:3] == [0.0, 0.5, 1.0] and vals[6] == 1.0 and abs(vals[10] - 0.1) < 1e-12
  wc = lr_schedule.WarmupConstant(O(), 0.0, 2.0, 4)
  assert [wc.get_lr(t) for t in (0, 2, 4, 9)] == [0.0, 1.0, 2.0, 2.0]
  o = O()
  lc = lr_schedule.LinearCooldown(o, 1.0, 0.0, cooldown_start_step=10, cooldown_steps=5)
  assert o.param_groups[0]['lr'] is None and lc.cooldown_steps == 5
  assert lc.get_lr(10) == 1.0 and abs(lc.get_lr(15)) < 1e-12
  lc.load_state_dict({'iter': 7, 'lr_max': 99})
  assert lc.iter == 7 and lc.lr_max == 1.0


def test_prefetch_loader_keeps_order_length_and_errors():
  """PrefetchLoader (N1): same batches in the same order as the wrapped loader, `len` preserved, bounded look-ahead,
  loader exceptions re-raised in the consumer.  The device copy is injected (identity) so this runs without a GPU."""
  import threading

  from plainlm_b200.data_utils import PrefetchLoader

  class Ready:
    def wait(self):
      pass

  pulled = []

  class Loader:
    def __init__(self, n, fail_at=None):
      self.n, self.fail_at = n, fail_at

    def __len__(self):
      return self.n

    def __iter__(self):
      for i in range(self.n):
        if i == self.fail_at:
          raise RuntimeError('loader broke')
        pulled.append(i)
        yield {'input_ids': torch.full((2, 5), i), 'docs_lengths': [[5], [2, 3]]}

  pl = PrefetchLoader(Loader(7), 'cpu', depth=2, to_device=lambda t: (t, Ready()))
  assert len(pl) == 7
  seen = []
  for k, b in enumerate(pl):
    seen.append(int(b['input_ids'][0, 0]))
    assert b['docs_lengths'] == [[5], [2, 3]]
    assert len(pulled) <= k + 1 + 2 + 1  # never more than depth (+1 in flight) ahead of the consumer
  assert seen == list(range(7))
  with pytest.raises(RuntimeError, match='loader broke'):
    for _ in PrefetchLoader(Loader(5, fail_at=3), 'cpu', to_device=lambda t: (t, Ready())):
      pass
  # abandoning the iteration early stops the worker thread
  it = iter(PrefetchLoader(Loader(100), 'cpu', to_device=lambda t: (t, Ready())))
  next(it)
  it.close()
  import time
  time.sleep(0.5)
  assert not any(t.name == 'plm-prefetch' and t.is_alive() for t in threading.enumerate())


def test_binding_constants_match_the_header():
  """The ctypes side mirrors the header by hand: epilogue kinds, ABI version, activation kinds and the gemm-args struct
  (field order and count) must agree with include/plainlm_b200.h."""
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  hdr = open(os.path.join(root, 'include', 'plainlm_b200.h')).read()
  defines = {m.group(1): int(m.group(2)) for m in re.finditer(r'#define\s+(PLM_[A-Z0-9_]+)\s+\(?(-?\d+)\)?', hdr)}
  for name in ('BF16', 'BF16_ROPE', 'F32', 'RESID_F32', 'ATOMIC_F32', 'BF16_SWIGLU', 'BF16_CE', 'BF16_GLU_BWD'):
    assert getattr(_lib, 'EPI_' + name) == defines['PLM_EPI_' + name], name
  assert _lib.load().plm_abi_version() == defines['PLM_ABI_VERSION']
  assert _lib.SUMSQ_WORKSPACE == defines['PLM_SUMSQ_WORKSPACE']
  assert (_lib.ACT_SILU, _lib.ACT_RELU2) == (defines['PLM_ACT_SILU'], defines['PLM_ACT_RELU2'])
  body = re.search(r'typedef struct plm_gemm_args \{(.*?)\} plm_gemm_args;', hdr, re.S).group(1)
  body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
  fields = []
  for decl in body.split(';'):
    decl = decl.strip()
    if decl:
      fields += [f.strip().lstrip('*').strip() for f in decl.split(',')]
  fields = [f.split()[-1].lstrip('*') for f in fields]
  assert fields == [n for n, _ in _lib.GemmArgs._fields_], (fields, [n for n, _ in _lib.GemmArgs._fields_])

"""Model factory and weight-decay grouping behind the reference's names (reference: models/construct.py).

`construct_model(cfg) -> (model, model_cfg)` and `get_param_groups(model, weight_decay)` are what train.py and
engine/engine.py import; config keys keep their meaning.
"""

from fractions import Fraction

# config key of config/*.yaml -> (ModelConfig field, conversion)
_CFG_FIELDS = {
  'vocab_size': ('vocab_size', int),
  'd_model': ('dim', int),
  'expand': ('expand', lambda v: float(Fraction(v))),  # "8/3" style strings are allowed (reference :15)
  'n_layers': ('n_layers', int),
  'n_heads': ('n_heads', int),
  'mlp_class': ('mlp', str),
  'seq_len': ('seq_len', int),
  'tie_embeddings': ('tie_embeddings', bool),
}


def _report_size(model):
  total, body = model.count_params(non_embedding=False), model.count_params(non_embedding=True)
  print(f'Number of parameters: {total:_}')
  print(f'Number of non-embedding parameters: {body:_}')
  try:  # the reference logs the same two numbers to an active wandb run
    import wandb
  except ImportError:
    return
  if wandb.run is not None:
    wandb.log({'n_params': total, 'n_params_no_embed': body})


def construct_model(cfg):
  """reference: models/construct.py:5-44.  `cfg.model == 'transformer'` is the family the B200 path implements."""
  family = cfg.model
  if family != 'transformer':
    why = 'the HF model zoo is outside the B200 hot path; use the reference for it' if family.startswith('pythia') \
        else f'Not implemented model: {family}.'
    raise NotImplementedError(why)
  from .transformer import ModelConfig, Transformer

  fields = {dst: conv(getattr(cfg, src)) for src, (dst, conv) in _CFG_FIELDS.items()}
  model_cfg = ModelConfig(rmsnorm_eps=1e-6, **fields)
  model = Transformer(model_cfg)
  _report_size(model)
  return model, model_cfg


def _decays(name, param):
  """reference: models/construct.py:54-61 — biases, norm weights and tensors flagged `_no_weight_decay` are exempt;
  everything else (embed_tokens and lm_head included) decays."""
  return not ('bias' in name or 'norm' in name or getattr(param, '_no_weight_decay', False))


def get_param_groups(model, weight_decay):
  """Two groups in the reference's order: [decayed, exempt] (reference: models/construct.py:47-75)."""
  decayed, exempt = [], []
  for name, param in model.named_parameters():
    if param.requires_grad:
      (decayed if _decays(name, param) else exempt).append(param)
  return [dict(params=decayed, weight_decay=weight_decay), dict(params=exempt, weight_decay=0.0)]

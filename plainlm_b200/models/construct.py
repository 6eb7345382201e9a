"""Model factory and weight-decay grouping (reference: models/construct.py), same signatures and config keys."""

from fractions import Fraction


def construct_model(cfg):
  """reference: models/construct.py:5-44.  Only cfg.model == 'transformer' is on the B200 hot path."""
  if cfg.model == 'transformer':
    from .transformer import Transformer, ModelConfig

    model_cfg = ModelConfig(
      vocab_size=cfg.vocab_size,
      dim=cfg.d_model,
      expand=float(Fraction(cfg.expand)),
      n_layers=cfg.n_layers,
      n_heads=cfg.n_heads,
      rmsnorm_eps=1e-6,
      mlp=cfg.mlp_class,
      seq_len=cfg.seq_len,
      tie_embeddings=cfg.tie_embeddings,
    )
    model = Transformer(model_cfg)
  elif cfg.model.startswith('pythia'):
    raise NotImplementedError('pythia (HF model zoo) is outside the B200 hot path; use the reference for it')
  else:
    raise NotImplementedError(f'Not implemented model: {cfg.model}.')

  n_params = model.count_params(non_embedding=False)
  n_params_no_embed = model.count_params(non_embedding=True)
  print(f'Number of parameters: {n_params:_}')
  print(f'Number of non-embedding parameters: {n_params_no_embed:_}')
  try:  # same side effect as the reference when a wandb run is active
    import wandb

    if wandb.run is not None:
      wandb.log({'n_params': n_params, 'n_params_no_embed': n_params_no_embed})
  except ImportError:
    pass
  return model, model_cfg


def get_param_groups(model, weight_decay):
  """reference: models/construct.py:47-75 — no decay for names containing 'bias' or 'norm' or flagged
  `_no_weight_decay`; everything else (including embed_tokens and lm_head) decays."""
  named = {n: p for n, p in model.named_parameters() if p.requires_grad}
  decay_names = [n for n, p in model.named_parameters() if not getattr(p, '_no_weight_decay', False)]
  decay_names = [n for n in decay_names if 'bias' not in n and 'norm' not in n]
  decay_set = set(decay_names)
  decay_params = [p for n, p in named.items() if n in decay_set]
  no_decay_params = [p for n, p in named.items() if n not in decay_set]
  return [
    {'params': decay_params, 'weight_decay': weight_decay},
    {'params': no_decay_params, 'weight_decay': 0.0},
  ]

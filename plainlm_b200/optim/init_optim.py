"""Optimizer and scheduler factories behind the reference's names (reference: optim/init_optim.py), including the
`intialize_optimizer` spelling that train.py / engine.py import.  Every optimizer built here runs on the flat-buffer
CUDA kernels of this package; the per-group `weight_decay` of `param_groups` overrides the default passed to the
constructor, as in the reference."""

from .lr_schedule import LinearCooldown, WarmupConstant, WarmupCosine, WSD


def _adam_family(cls_path, cfg, **extra):
  module, name = cls_path
  cls = getattr(__import__(f'{__package__}.{module}', fromlist=[name]), name)
  return lambda groups: cls(groups, lr=cfg.lr, betas=[cfg.beta1, cfg.beta2], weight_decay=cfg.weight_decay,
                            fused=getattr(cfg, 'fused_optim', True), eps=getattr(cfg, 'eps', 1e-8), **extra)


def _momentum_family(cls_path, cfg):
  module, name = cls_path
  cls = getattr(__import__(f'{__package__}.{module}', fromlist=[name]), name)
  return lambda groups: cls(groups, lr=cfg.lr, momentum=cfg.beta1, dampening=cfg.dampening,
                            weight_decay=cfg.weight_decay)


def intialize_optimizer(param_groups, cfg):
  """cfg.optim -> optimizer (reference: init_optim.py:13-66): adamw (:13-21), nadamw (:23-32), sgd (:34-41),
  signSGD (:43-52)."""
  builders = {
    'adamw': lambda: _adam_family(('adamw', 'AdamW'), cfg),
    'nadamw': lambda: _adam_family(('nadamw', 'NAdamW'), cfg, decoupled_weight_decay=True),
    'sgd': lambda: _momentum_family(('sgd', 'SGD'), cfg),
    'signSGD': lambda: _momentum_family(('signSGD', 'signSGD'), cfg),
  }
  if cfg.optim == 'sfo_adamw':
    raise NotImplementedError(
      "optim 'sfo_adamw' needs the third-party schedulefree package (not a kernel of this path; SURVEY.md §8(f) N4); "
      f'supported: {sorted(builders)}'
    )
  if cfg.optim not in builders:
    raise NotImplementedError(f'Not implemented optim: {cfg.optim}.')
  return builders[cfg.optim]()(param_groups)


def _resolve_steps(value, budget):
  """An int is a number of steps, a float a fraction of `steps_budget` (reference: init_optim.py:79-88)."""
  if value is None:
    return None
  return value if isinstance(value, int) else int(value * budget)


def initialize_scheduler(optimizer, cfg):
  """cfg.scheduler -> schedule object or None (reference: optim/init_optim.py:73-137)."""
  kind = cfg.scheduler
  if kind is None:
    return None
  budget = cfg.steps_budget
  warmup = _resolve_steps(getattr(cfg, 'warmup_steps', None), budget)
  cooldown = _resolve_steps(getattr(cfg, 'cooldown_steps', None), budget)
  lr_end = getattr(cfg, 'lr_end', None)
  if lr_end is None and getattr(cfg, 'lr_end_pct', None) is not None:
    lr_end = cfg.lr_end_pct * cfg.lr

  if kind == 'warmup_cosine':
    return WarmupCosine(optimizer, lr_start=cfg.lr_start, lr_max=cfg.lr, lr_end=lr_end, warmup_steps=warmup, T=budget)
  if kind == 'wsd':
    return WSD(optimizer, lr_start=cfg.lr_start, lr_max=cfg.lr, lr_end=lr_end, warmup_steps=warmup,
               cooldown_start_step=budget - cooldown, cooldown_steps=cooldown)
  if kind == 'warmup_constant':
    return WarmupConstant(optimizer, lr_start=cfg.lr_start, lr_max=cfg.lr, warmup_steps=warmup)
  if kind == 'linear_cooldown':
    return LinearCooldown(optimizer, lr_max=cfg.lr, lr_end=lr_end, cooldown_start_step=cfg.resume_step,
                          cooldown_steps=cooldown)
  raise NotImplementedError(f'Not implemented scheduler: {kind}.')

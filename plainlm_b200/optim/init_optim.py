"""Optimizer / scheduler factories with the reference's names, signatures and config keys
(reference: optim/init_optim.py — including the `intialize_optimizer` spelling that train.py/engine.py import)."""

from .lr_schedule import WarmupCosine, WSD, WarmupConstant, LinearCooldown


def intialize_optimizer(param_groups, cfg):
  """cfg.optim in {'adamw', 'nadamw', 'sgd', 'signSGD'} run on the native flat-buffer kernels (reference: init_optim.py:13-21,43-52).
  The per-group `weight_decay` of `param_groups` overrides the default passed here, as in the reference."""
  if cfg.optim == 'adamw':
    from .adamw import AdamW

    return AdamW(
      param_groups,
      lr=cfg.lr,
      betas=[cfg.beta1, cfg.beta2],
      weight_decay=cfg.weight_decay,
      fused=getattr(cfg, 'fused_optim', True),
      eps=getattr(cfg, 'eps', 1e-8),
    )
  if cfg.optim == 'signSGD':
    from .signSGD import signSGD

    return signSGD(
      param_groups,
      lr=cfg.lr,
      momentum=cfg.beta1,
      dampening=cfg.dampening,
      weight_decay=cfg.weight_decay,
    )
  if cfg.optim == 'nadamw':
    from .nadamw import NAdamW

    return NAdamW(
      param_groups,
      lr=cfg.lr,
      betas=[cfg.beta1, cfg.beta2],
      weight_decay=cfg.weight_decay,
      decoupled_weight_decay=True,
      fused=getattr(cfg, 'fused_optim', True),
      eps=getattr(cfg, 'eps', 1e-8),
    )
  if cfg.optim == 'sgd':
    from .sgd import SGD

    return SGD(
      param_groups,
      lr=cfg.lr,
      momentum=cfg.beta1,
      dampening=cfg.dampening,
      weight_decay=cfg.weight_decay,
    )
  if cfg.optim == 'sfo_adamw':
    raise NotImplementedError(
      "optim 'sfo_adamw' needs the third-party schedulefree package (not a kernel of this path; SURVEY.md §8(f) N4); "
      "supported: 'adamw', 'nadamw', 'sgd', 'signSGD'"
    )
  raise NotImplementedError(f'Not implemented optim: {cfg.optim}.')


def _resolve_steps(value, budget):
  """int = absolute number of steps, float = fraction of steps_budget (reference: init_optim.py:79-88)."""
  return value if isinstance(value, int) else int(value * budget)


def initialize_scheduler(optimizer, cfg):
  """reference: optim/init_optim.py:73-137."""
  if cfg.scheduler is None:
    return None
  warmup_steps = cooldown_steps = lr_end = None
  if getattr(cfg, 'warmup_steps', None) is not None:
    warmup_steps = _resolve_steps(cfg.warmup_steps, cfg.steps_budget)
  if getattr(cfg, 'cooldown_steps', None) is not None:
    cooldown_steps = _resolve_steps(cfg.cooldown_steps, cfg.steps_budget)
  if getattr(cfg, 'lr_end', None) is not None or getattr(cfg, 'lr_end_pct', None) is not None:
    lr_end = cfg.lr_end if cfg.lr_end is not None else cfg.lr_end_pct * cfg.lr

  if cfg.scheduler == 'warmup_cosine':
    return WarmupCosine(optimizer, lr_start=cfg.lr_start, lr_max=cfg.lr, lr_end=lr_end, warmup_steps=warmup_steps,
                        T=cfg.steps_budget)
  if cfg.scheduler == 'wsd':
    return WSD(optimizer, lr_start=cfg.lr_start, lr_max=cfg.lr, lr_end=lr_end, warmup_steps=warmup_steps,
               cooldown_start_step=cfg.steps_budget - cooldown_steps, cooldown_steps=cooldown_steps)
  if cfg.scheduler == 'warmup_constant':
    return WarmupConstant(optimizer, lr_start=cfg.lr_start, lr_max=cfg.lr, warmup_steps=warmup_steps)
  if cfg.scheduler == 'linear_cooldown':
    return LinearCooldown(optimizer, lr_max=cfg.lr, lr_end=lr_end, cooldown_start_step=cfg.resume_step,
                          cooldown_steps=cooldown_steps)
  raise NotImplementedError(f'Not implemented scheduler: {cfg.scheduler}.')

"""Learning-rate schedules for the native optimizers: host-side scalar math that writes `group['lr']` before each step.

Drop-in for the reference's optim/lr_schedule.py: identical class names, constructor keywords, attribute names (they
are what `state_dict()` checkpoints) and lr(t) values.  Here every schedule is one piecewise function assembled from two
segment shapes — a straight line between two (step, lr) knots and a half cosine — evaluated by a shared base class.
"""

import math

_HYPER = ('lr_start', 'lr_max', 'lr_end', 'warmup_steps', 'T', 'cooldown_start_step', 'cooldown_steps')


def _line(t, t0, y0, t1, y1):
  """Value at t of the straight line through (t0, y0) and (t1, y1)."""
  return y0 + (y1 - y0) / (t1 - t0) * (t - t0)


def _half_cosine(t, t0, y0, t1, y1):
  """Value at t of the half cosine that falls from y0 at t0 to y1 at t1."""
  frac = (t - t0) / (t1 - t0)  # evaluated first: keeps lr(t) bit-identical to the reference's float arithmetic
  return y1 + 0.5 * (y0 - y1) * (1 + math.cos(math.pi * frac))


class CustomLRSchedule:
  """Base: subclasses implement get_lr(t).  step() advances `iter` and pushes lr(iter) into every param group."""

  push_initial_lr = True  # the constructor writes lr_start, so optimizer step 1 runs at lr_start (reference :40)

  def __init__(self, optimizer, **hyper):
    self.optimizer = optimizer
    for key in _HYPER:  # fixed attribute order; only the schedule's own hyper-parameters become attributes
      if key in hyper:
        setattr(self, key, hyper[key])
    self.iter = 0
    if self.push_initial_lr:
      self.set_optim_lr(self.lr_start)

  def set_optim_lr(self, lr):
    for group in self.optimizer.param_groups:
      group['lr'] = lr

  def get_lr(self, t):
    raise NotImplementedError

  def step(self):
    self.iter += 1
    self.set_optim_lr(self.get_lr(self.iter))

  def state_dict(self):
    return {k: v for k, v in vars(self).items() if k != 'optimizer'}

  def load_state_dict(self, state_dict):
    vars(self).update(state_dict)

  # shared segments -------------------------------------------------------------------------------------------------
  def _warmup(self, t):
    return _line(t, 0, self.lr_start, self.warmup_steps, self.lr_max)

  def _cooldown(self, t):
    t0 = self.cooldown_start_step
    return _line(t, t0, self.lr_max, t0 + self.cooldown_steps, self.lr_end)


class WarmupCosine(CustomLRSchedule):
  """line lr_start -> lr_max over warmup_steps, half cosine down to lr_end at step T, lr_end from then on."""

  def __init__(self, optimizer, lr_start, lr_max, lr_end, warmup_steps, T):
    super().__init__(optimizer, lr_start=lr_start, lr_max=lr_max, lr_end=lr_end, warmup_steps=warmup_steps, T=T)

  def get_lr(self, t):
    if t <= self.warmup_steps:
      return self._warmup(t)
    return _half_cosine(t, self.warmup_steps, self.lr_max, self.T, self.lr_end) if t <= self.T else self.lr_end


class WSD(CustomLRSchedule):
  """warmup, stable at lr_max until cooldown_start_step, then a line towards lr_end over cooldown_steps."""

  def __init__(self, optimizer, lr_start, lr_max, lr_end, warmup_steps, cooldown_start_step, cooldown_steps):
    super().__init__(optimizer, lr_start=lr_start, lr_max=lr_max, lr_end=lr_end, warmup_steps=warmup_steps,
                     cooldown_start_step=cooldown_start_step, cooldown_steps=cooldown_steps)

  def get_lr(self, t):
    if t <= self.warmup_steps:
      return self._warmup(t)
    return self.lr_max if t <= self.cooldown_start_step else self._cooldown(t)


class WarmupConstant(CustomLRSchedule):
  def __init__(self, optimizer, lr_start, lr_max, warmup_steps):
    super().__init__(optimizer, lr_start=lr_start, lr_max=lr_max, warmup_steps=warmup_steps)

  def get_lr(self, t):
    return self._warmup(t) if t <= self.warmup_steps else self.lr_max


class LinearCooldown(CustomLRSchedule):
  """lr_max until cooldown_start_step, then a line towards lr_end.  As in the reference it neither sets the lr at
  construction nor restores anything but `iter` from a checkpoint."""

  push_initial_lr = False

  def __init__(self, optimizer, lr_max, lr_end, cooldown_start_step, cooldown_steps):
    super().__init__(optimizer, lr_max=lr_max, lr_end=lr_end, cooldown_start_step=cooldown_start_step,
                     cooldown_steps=cooldown_steps)

  def get_lr(self, t):
    return self.lr_max if t <= self.cooldown_start_step else self._cooldown(t)

  def load_state_dict(self, state_dict):
    self.iter = state_dict.get('iter', 0)

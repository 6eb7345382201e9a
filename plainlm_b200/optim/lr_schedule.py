"""Learning-rate schedules that write `group['lr']` before every optimizer step (host-side scalar math).

Same class names, constructor arguments, attribute names (they are what `state_dict()` checkpoints) and lr(t)
values as the reference's optim/lr_schedule.py; the native optimizers read `group['lr']` on every step.
"""

import math


class CustomLRSchedule:
  """lr(t) = self.get_lr(t); `step()` advances t and pushes lr(t) into every param group."""

  def __init__(self, optimizer):
    self.optimizer = optimizer

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


def _linear_ramp(t, y0, y1, steps):
  return y0 + (y1 - y0) / steps * t


class WarmupCosine(CustomLRSchedule):
  """linear warmup lr_start -> lr_max over warmup_steps, cosine to lr_end at step T, lr_end afterwards."""

  def __init__(self, optimizer, lr_start, lr_max, lr_end, warmup_steps, T):
    super().__init__(optimizer)
    self.lr_start, self.lr_max, self.lr_end = lr_start, lr_max, lr_end
    self.warmup_steps, self.T = warmup_steps, T
    self.iter = 0
    self.set_optim_lr(lr_start)  # the first optimizer step runs at lr_start (reference: lr_schedule.py:40)

  def get_lr(self, t):
    if t <= self.warmup_steps:
      return _linear_ramp(t, self.lr_start, self.lr_max, self.warmup_steps)
    if t > self.T:
      return self.lr_end
    frac = (t - self.warmup_steps) / (self.T - self.warmup_steps)
    return self.lr_end + 0.5 * (self.lr_max - self.lr_end) * (1 + math.cos(math.pi * frac))


class WSD(CustomLRSchedule):
  """warmup, stable at lr_max until cooldown_start_step, then linear decay towards lr_end."""

  def __init__(self, optimizer, lr_start, lr_max, lr_end, warmup_steps, cooldown_start_step, cooldown_steps):
    super().__init__(optimizer)
    self.lr_start, self.lr_max, self.lr_end = lr_start, lr_max, lr_end
    self.warmup_steps = warmup_steps
    self.cooldown_start_step, self.cooldown_steps = cooldown_start_step, cooldown_steps
    self.iter = 0
    self.set_optim_lr(lr_start)

  def get_lr(self, t):
    if t <= self.warmup_steps:
      return _linear_ramp(t, self.lr_start, self.lr_max, self.warmup_steps)
    if t <= self.cooldown_start_step:
      return self.lr_max
    return _linear_ramp(t - self.cooldown_start_step, self.lr_max, self.lr_end, self.cooldown_steps)


class WarmupConstant(CustomLRSchedule):
  def __init__(self, optimizer, lr_start, lr_max, warmup_steps):
    super().__init__(optimizer)
    self.lr_start, self.lr_max, self.warmup_steps = lr_start, lr_max, warmup_steps
    self.iter = 0
    self.set_optim_lr(lr_start)

  def get_lr(self, t):
    if t <= self.warmup_steps:
      return _linear_ramp(t, self.lr_start, self.lr_max, self.warmup_steps)
    return self.lr_max


class LinearCooldown(CustomLRSchedule):
  """lr_max until cooldown_start_step, then linear towards lr_end. Does NOT set the lr at construction and restores
  only `iter` from a checkpoint (both as in the reference)."""

  def __init__(self, optimizer, lr_max, lr_end, cooldown_start_step, cooldown_steps):
    super().__init__(optimizer)
    self.lr_max, self.lr_end = lr_max, lr_end
    self.cooldown_start_step, self.cooldown_steps = cooldown_start_step, cooldown_steps
    self.iter = 0

  def get_lr(self, t):
    if t <= self.cooldown_start_step:
      return self.lr_max
    return _linear_ramp(t - self.cooldown_start_step, self.lr_max, self.lr_end, self.cooldown_steps)

  def load_state_dict(self, state_dict):
    self.iter = state_dict.get('iter', 0)

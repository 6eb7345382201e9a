"""signSGD / signum on flat buffers (reference: optim/signSGD.py), one kernel launch per contiguous parameter run.

Same constructor, param_groups keys and state key (`m`) as the reference; same update, including its first-step quirk:
the momentum buffer starts as a clone of the gradient and is THEN decayed and accumulated, so m_1 = (mu + 1 - damp) g_1.
"""

import torch
from torch.optim import Optimizer

from .. import ops
from .flat import plan_runs


class signSGD(Optimizer):
  STATE_KEY = 'm'  # reference: optim/signSGD.py state key

  def __init__(self, params, lr, momentum=0.0, dampening=0.0, weight_decay=0.1):
    if not 0.0 <= lr:
      raise ValueError(f'Invaid learing rate: {lr}')
    if not 0.0 <= momentum <= 1.0:
      raise ValueError(f'Invaid momentum: {momentum}')
    if not 0.0 <= dampening <= 1.0:
      raise ValueError(f'Invaid dampening: {dampening}')
    if not 0.0 <= weight_decay:
      raise ValueError(f'Invaid weight decay: {weight_decay}')
    super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay))
    self._plans = {}

  def _plan(self, gi, group):
    key = tuple(p.data_ptr() for p in group['params'])
    plan = self._plans.get(gi)
    if plan is None or plan[0] != key:
      runs, loose = plan_runs([p for p in group['params'] if p.requires_grad])
      bufs = []
      for flat, a, b, ps in runs:
        m = torch.zeros(b - a, device=flat.params.device, dtype=torch.float32)
        key_ = self.STATE_KEY
        have = [key_ in self.state[p] for p in ps]
        if any(have) and not all(have):
          raise RuntimeError('signSGD: partially initialised momentum state inside one flat run')
        for p in ps:
          o, k = p._plm_flat[1] - a, p._plm_flat[2]
          if key_ in self.state[p]:
            m[o : o + k].view(p.shape).copy_(self.state[p][key_])
        bufs.append([flat, a, b, ps, m, not all(have) if have else True])
      plan = (key, bufs, loose)
      self._plans[gi] = plan
    return plan

  def _kernel(self, p, g, buf, shadow, group, first, gsq, mx):
    ops.signsgd_step(p, g, buf, shadow, float(group['lr']), group['momentum'], group['dampening'],
                     group['weight_decay'], first, gnorm_sq=gsq, max_norm=mx)

  @torch.no_grad()
  def step(self, closure=None, grad_clip=None):
    loss = closure() if closure is not None else None
    gsq = grad_clip.gnorm_sq if grad_clip is not None else None
    mx = grad_clip.max_norm if grad_clip is not None else 0.0
    for gi, group in enumerate(self.param_groups):
      _, bufs, loose = self._plan(gi, group)
      key_ = self.STATE_KEY
      for rec in bufs:
        flat, a, b, ps, m, first = rec
        self._kernel(flat.params[a:b], flat.grads[a:b], m, flat.shadow[a:b], group, first, gsq, mx)
        if first:
          for p in ps:
            o, k = p._plm_flat[1] - a, p._plm_flat[2]
            self.state[p][key_] = m[o : o + k].view(p.shape)
          rec[5] = False
      for p in loose:
        if p.grad is None:
          continue
        st = self.state[p]
        first = key_ not in st
        if first:
          st[key_] = torch.zeros_like(p, memory_format=torch.preserve_format)
        self._kernel(p.data, p.grad.contiguous(), st[key_], getattr(p, '_plm_shadow', None), group, first, gsq, mx)
    return loss

  def zero_grad(self, set_to_none=True):
    flats = {}
    for group in self.param_groups:
      for p in group['params']:
        meta = getattr(p, '_plm_flat', None)
        if meta is not None:
          flats[id(meta[0])] = meta[0]
        elif p.grad is not None:
          if set_to_none:
            p.grad = None
          else:
            p.grad.zero_()
    for flat in flats.values():
      flat.zero_grads()
      flat.restore_grad_views()

  def load_state_dict(self, state_dict):
    super().load_state_dict(state_dict)
    self._plans = {}

"""AdamW on flat buffers, one kernel launch per contiguous parameter run.

Drop-in for `torch.optim.AdamW(param_groups, lr, betas, weight_decay, fused, eps)` as built by the reference at
optim/init_optim.py:13-21: same param_groups keys, same per-parameter state keys (`step`, `exp_avg`, `exp_avg_sq`),
same update (decoupled weight decay, bias-corrected) — verified against torch's fused kernel in tests.
The kernel also applies the gradient-clip coefficient and writes the bf16 weight shadow used by the GEMMs.
"""

import torch
from torch.optim import Optimizer

from .. import ops
from .flat import plan_runs


class AdamW(Optimizer):
  def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, fused=None):
    if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or weight_decay < 0.0:
      raise ValueError('AdamW: invalid hyper-parameter')
    super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
    self._plans = {}

  def _plan(self, gi, group):
    key = tuple(p.data_ptr() for p in group['params'])
    plan = self._plans.get(gi)
    if plan is None or plan[0] != key:
      runs, loose = plan_runs([p for p in group['params'] if p.requires_grad])
      bufs = []
      for flat, a, b, ps in runs:
        m = torch.zeros(b - a, device=flat.params.device, dtype=torch.float32)
        v = torch.zeros_like(m)
        for p in ps:  # carry over existing state (checkpoint resume) and expose per-parameter views
          o, k = p._plm_flat[1] - a, p._plm_flat[2]
          st = self.state[p]
          mv, vv = m[o : o + k].view(p.shape), v[o : o + k].view(p.shape)
          if 'exp_avg' in st:
            mv.copy_(st['exp_avg'])
            vv.copy_(st['exp_avg_sq'])
          st['exp_avg'], st['exp_avg_sq'] = mv, vv
          st.setdefault('step', torch.tensor(0.0, dtype=torch.float32))
        bufs.append((flat, a, b, ps, m, v))
      for p in loose:
        st = self.state[p]
        if 'exp_avg' not in st:
          st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
          st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
        st.setdefault('step', torch.tensor(0.0, dtype=torch.float32))
      plan = (key, bufs, loose)
      self._plans[gi] = plan
    return plan

  @torch.no_grad()
  def step(self, closure=None, grad_clip=None):
    """grad_clip: optional plainlm_b200.optim.flat.GradClip (device-side clipping fused into the update)."""
    loss = closure() if closure is not None else None
    gsq = grad_clip.gnorm_sq if grad_clip is not None else None
    mx = grad_clip.max_norm if grad_clip is not None else 0.0
    for gi, group in enumerate(self.param_groups):
      _, bufs, loose = self._plan(gi, group)
      b1, b2 = group['betas']
      lr, eps, wd = float(group['lr']), group['eps'], group['weight_decay']
      for flat, a, b, ps, m, v in bufs:
        st0 = self.state[ps[0]]
        t = int(st0['step'].item()) + 1  # CPU scalar tensor: no device sync
        ops.adamw_step(flat.params[a:b], flat.grads[a:b], m, v, flat.shadow[a:b], lr, b1, b2, eps, wd, t,
                       gnorm_sq=gsq, max_norm=mx)
        for p in ps:
          self.state[p]['step'].fill_(float(t))
      for p in loose:
        if p.grad is None:
          continue
        st = self.state[p]
        t = int(st['step'].item()) + 1
        g = p.grad.contiguous()
        ops.adamw_step(p.data, g, st['exp_avg'], st['exp_avg_sq'], getattr(p, '_plm_shadow', None), lr, b1, b2, eps,
                       wd, t, gnorm_sq=gsq, max_norm=mx)
        st['step'].fill_(float(t))
    return loss

  def zero_grad(self, set_to_none=True):
    """Flat gradient buffers are zeroed in place (their views stay attached); loose tensors follow torch."""
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
    for st in self.state.values():  # torch casts `step` to the param's device; keep it a host scalar
      if 'step' in st and torch.is_tensor(st['step']):
        st['step'] = st['step'].detach().to('cpu', torch.float32).reshape(())
    self._plans = {}

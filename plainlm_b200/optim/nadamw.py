"""NAdam with decoupled weight decay ("nadamw") on flat buffers, one kernel launch per contiguous parameter run.

Drop-in for `torch.optim.NAdam(param_groups, lr, betas, weight_decay, decoupled_weight_decay=True, eps)` as the
reference builds it at optim/init_optim.py:23-32: same param_groups keys (incl. `momentum_decay`), same per-parameter
state keys (`step`, `mu_product`, `exp_avg`, `exp_avg_sq`), same update (torch/optim/nadam.py, _single_tensor_nadam).
The momentum-schedule scalars are computed on the host in double precision and handed to the kernel.
"""

import torch
from torch.optim import Optimizer

from .. import ops
from .adamw import AdamW


class NAdamW(AdamW):
  def __init__(self, params, lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, momentum_decay=4e-3,
               decoupled_weight_decay=True, fused=None):
    if not decoupled_weight_decay:
      raise NotImplementedError('NAdamW: only decoupled weight decay (what optim/init_optim.py:23-32 builds)')
    if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or weight_decay < 0.0:
      raise ValueError('NAdamW: invalid hyper-parameter')
    Optimizer.__init__(self, params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay,
                                          momentum_decay=momentum_decay, decoupled_weight_decay=True))
    self._plans = {}

  def _scalars(self, st, group):
    """Advance (step, mu_product) of one state entry; return the kernel's host-side scalars."""
    b1, b2 = group['betas']
    lr, psi = float(group['lr']), group['momentum_decay']
    t = int(st['step'].item()) + 1
    mu = b1 * (1.0 - 0.5 * (0.96 ** (t * psi)))
    mu_next = b1 * (1.0 - 0.5 * (0.96 ** ((t + 1) * psi)))
    prod = float(st['mu_product'].item()) * mu
    c_grad = -lr * (1.0 - mu) / (1.0 - prod)
    c_mom = -lr * mu_next / (1.0 - prod * mu_next)
    return t, prod, 1.0 - b2 ** t, c_grad, c_mom

  @torch.no_grad()
  def step(self, closure=None, grad_clip=None):
    loss = closure() if closure is not None else None
    gsq = grad_clip.gnorm_sq if grad_clip is not None else None
    mx = grad_clip.max_norm if grad_clip is not None else 0.0
    for gi, group in enumerate(self.param_groups):
      _, bufs, loose = self._plan(gi, group)
      b1, b2 = group['betas']
      lr, eps, wd = float(group['lr']), group['eps'], group['weight_decay']
      for flat, a, b, ps, m, v in bufs:
        for p in ps:
          self.state[p].setdefault('mu_product', torch.tensor(1.0, dtype=torch.float32))
        t, prod, bc2, c_grad, c_mom = self._scalars(self.state[ps[0]], group)
        ops.nadamw_step(flat.params[a:b], flat.grads[a:b], m, v, flat.shadow[a:b], lr, b1, b2, eps, wd, bc2, c_grad,
                        c_mom, gnorm_sq=gsq, max_norm=mx)
        for p in ps:
          self.state[p]['step'].fill_(float(t))
          self.state[p]['mu_product'].fill_(prod)
      for p in loose:
        if p.grad is None:
          continue
        st = self.state[p]
        st.setdefault('mu_product', torch.tensor(1.0, dtype=torch.float32))
        t, prod, bc2, c_grad, c_mom = self._scalars(st, group)
        ops.nadamw_step(p.data, p.grad.contiguous(), st['exp_avg'], st['exp_avg_sq'], getattr(p, '_plm_shadow', None),
                        lr, b1, b2, eps, wd, bc2, c_grad, c_mom, gnorm_sq=gsq, max_norm=mx)
        st['step'].fill_(float(t))
        st['mu_product'].fill_(prod)
    return loss

  def load_state_dict(self, state_dict):
    super().load_state_dict(state_dict)
    for st in self.state.values():
      if 'mu_product' in st and torch.is_tensor(st['mu_product']):
        st['mu_product'] = st['mu_product'].detach().to('cpu', torch.float32).reshape(())

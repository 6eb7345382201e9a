"""SGD with momentum on flat buffers, one kernel launch per contiguous parameter run.

Drop-in for `torch.optim.SGD(param_groups, lr, momentum, dampening, weight_decay)` as the reference builds it at
optim/init_optim.py:34-41: same param_groups keys, same state key (`momentum_buffer`), same update (coupled L2 weight
decay; the buffer starts as a clone of the first gradient; no Nesterov).
"""

import torch

from .. import ops
from .signSGD import signSGD


class SGD(signSGD):
  STATE_KEY = 'momentum_buffer'

  def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
    if nesterov:
      raise NotImplementedError('SGD: nesterov is not part of the reference path (optim/init_optim.py:34-41)')
    if lr < 0.0 or momentum < 0.0 or weight_decay < 0.0:
      raise ValueError('SGD: invalid hyper-parameter')
    torch.optim.Optimizer.__init__(self, params, dict(lr=lr, momentum=momentum, dampening=dampening,
                                                      weight_decay=weight_decay, nesterov=False))
    self._plans = {}

  def _kernel(self, p, g, buf, shadow, group, first, gsq, mx):
    ops.sgd_step(p, g, buf, shadow, float(group['lr']), group['momentum'], group['dampening'], group['weight_decay'],
                 first, gnorm_sq=gsq, max_norm=mx)

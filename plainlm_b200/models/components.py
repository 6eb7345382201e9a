"""Parameter containers mirroring the reference's models/components.py (same constructors, same state_dict names).

The modules own parameters only; the arithmetic is done by plainlm_b200.models.runtime on the CUDA kernels.
Each `forward` here is the stand-alone entry used when a block is called outside the fused train step.
"""

import torch
from torch import nn

from .. import _lib
from . import functional as PF


class RMSNorm(nn.Module):
  """reference: models/components.py:16-28. Output is bf16 (what autocast feeds the next nn.Linear)."""

  def __init__(self, dim: int, eps: float = 1e-6):
    super().__init__()
    self.eps = eps
    self.weight = nn.Parameter(torch.ones(dim))

  def forward(self, x):
    return PF.rmsnorm(x, self.weight, self.eps)


class GLU(nn.Module):
  """reference: models/components.py:43-56 ("fused GLU": fc1 produces [gate | value])."""

  def __init__(self, dim: int, hidden_dim: int, multiple_of: int = 256):
    super().__init__()
    hidden_dim = multiple_of * ((hidden_dim + multiple_of - 1) // multiple_of)
    self.hidden_dim = hidden_dim
    self.fc1 = nn.Linear(dim, 2 * hidden_dim, bias=False)
    self.fc2 = nn.Linear(hidden_dim, dim, bias=False)

  def forward(self, x):
    # x: (bsz, T, dim) bf16
    u = PF.linear(x, self.fc1.weight)
    return PF.linear(PF.swiglu(u), self.fc2.weight)


class _PlainMLP(nn.Module):
  """fc2(act(fc1 x)) with a single-branch activation (kernel: plm_act_fwd / plm_act_bwd)."""

  act_kind = None  # _lib.ACT_*

  def __init__(self, dim: int, hidden_dim: int, multiple_of: int = 256):
    super().__init__()
    hidden_dim = multiple_of * ((hidden_dim + multiple_of - 1) // multiple_of)
    self.hidden_dim = hidden_dim
    self.fc1 = nn.Linear(dim, hidden_dim, bias=False)
    self.fc2 = nn.Linear(hidden_dim, dim, bias=False)

  def forward(self, x):
    return PF.linear(PF.activation(PF.linear(x, self.fc1.weight), self.act_kind), self.fc2.weight)


class MLP(_PlainMLP):
  """reference: models/components.py:31-40: fc2(silu(fc1 x))."""

  act_kind = _lib.ACT_SILU


class MLPReluSquared(_PlainMLP):
  """reference: models/components.py:59-70: fc2(relu(fc1 x)^2)."""

  act_kind = _lib.ACT_RELU2

"""Parameter containers mirroring the reference's models/components.py (same constructors, same state_dict names).

The modules own parameters only; the arithmetic is done by plainlm_b200.models.runtime on the CUDA kernels.
Each `forward` here is the stand-alone entry used when a block is called outside the fused train step.
"""

import torch
from torch import nn

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


class MLP(nn.Module):
  """reference: models/components.py:31-40. Constructor kept for config compatibility; SURVEY.md §8(f) N4 ("next")."""

  def __init__(self, dim: int, hidden_dim: int, multiple_of: int = 256):
    super().__init__()
    raise NotImplementedError("mlp_class 'mlp' is outside the B200 hot path (use 'glu'); SURVEY.md §8(f) N4")


class MLPReluSquared(nn.Module):
  """reference: models/components.py:59-70. See MLP."""

  def __init__(self, dim: int, hidden_dim: int, multiple_of: int = 256):
    super().__init__()
    raise NotImplementedError("mlp_class 'mlp_relu_sq' is outside the B200 hot path (use 'glu'); SURVEY.md §8(f) N4")

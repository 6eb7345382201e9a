"""Shared machinery of the native optimizers: group parameters into contiguous runs of the model's flat buffer so
that one kernel launch updates a whole run (instead of one launch — or five — per tensor)."""

import torch

from .. import _lib, ops


def plan_runs(params):
  """[(flat, start, end, [params])...] for parameters that are views of a FlatParams buffer, merged when adjacent;
  parameters that are not (stand-alone tensors) come back as single-tensor runs with flat=None."""
  tagged, loose = [], []
  for p in params:
    meta = getattr(p, '_plm_flat', None)
    if meta is not None and p.data_ptr() == meta[0].params.data_ptr() + 4 * meta[1]:
      tagged.append((meta[0], meta[1], meta[2], p))
    else:
      loose.append(p)
  tagged.sort(key=lambda t: (id(t[0]), t[1]))
  runs = []
  for flat, off, numel, p in tagged:
    padded_end = off + (numel + 63) // 64 * 64
    if runs and runs[-1][0] is flat and runs[-1][2] == off:
      runs[-1][2] = padded_end
      runs[-1][3].append(p)
    else:
      runs.append([flat, off, padded_end, [p]])
  return [tuple(r) for r in runs], loose


class GradClip:
  """Device-side clip state handed from the engine to optimizer.step(): sum of squared gradients (fp32[1] on the
  device, produced by plm_sumsq) and the max norm.  The coefficient is computed inside the update kernel, so clipping
  costs one 4 B/param read pass and no host synchronisation."""

  def __init__(self, gnorm_sq, max_norm):
    self.gnorm_sq = gnorm_sq
    self.max_norm = float(max_norm)


def sumsq_workspace(device):
  return torch.empty(_lib.SUMSQ_WORKSPACE, device=device, dtype=torch.float32)


def grad_sumsq(flat, workspace, out):
  """out[0] = ||g||^2 over the whole flat gradient buffer (padding elements are always zero)."""
  return ops.sumsq(flat.grads, workspace, out, accumulate=False)

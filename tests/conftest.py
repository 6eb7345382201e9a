import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
os.environ.setdefault('TORCHDYNAMO_DISABLE', '1')

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu under gpurun)')


@pytest.fixture(scope='session')
def golden_dir():
  return GOLDEN


def assert_close(got, ref, rtol, atol=0.0, what=''):
  """|got - ref| <= atol + rtol * max|ref|  (tolerance relative to the tensor's scale, as for bf16 kernels)."""
  import torch

  got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
  assert got.shape == ref.shape, f'{what}: shape {tuple(got.shape)} vs {tuple(ref.shape)}'
  assert not torch.isnan(got).any(), f'{what}: NaN in result'
  scale = ref.abs().max().item()
  err = (got - ref).abs().max().item()
  assert err <= atol + rtol * scale, f'{what}: max err {err:.3e} > {atol:.1e} + {rtol:.1e} * {scale:.3e}'


def assert_close_elementwise(got, ref, rtol, what='', outliers=1e-4, cap=3.0):
  """Elementwise bound next to the scale-relative one: bound_i = rtol * (|ref_i| + rms(ref)).  The rms term is the
  noise floor of a bf16 result whose operands were rounded to bf16 (an element that cancels to ~0 still carries rounding
  noise proportional to the typical magnitude); it is 4-6x tighter than rtol * max|ref| for the small elements the
  scale-relative check cannot see (masked-tile edges, dQ tails).  Rounding noise is random, so over 10^5 elements a
  handful land in its far tail: at most `outliers` of the elements may exceed bound_i, none may exceed cap * bound_i."""
  import torch

  got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
  assert got.shape == ref.shape, f'{what}: shape {tuple(got.shape)} vs {tuple(ref.shape)}'
  assert not torch.isnan(got).any(), f'{what}: NaN in result'
  rms = ref.pow(2).mean().sqrt().item()
  ratio = (got - ref).abs() / (rtol * (ref.abs() + rms) + 1e-300)
  n_out = int((ratio > 1).sum())
  worst = ratio.max().item()
  if worst > cap or n_out > max(1, int(outliers * ratio.numel())):
    idx = int(ratio.argmax())
    raise AssertionError(f'{what}: {n_out} of {ratio.numel()} elements beyond rtol*(|ref|+rms), worst {worst:.2f}x at '
                         f'{idx}: got {got.flatten()[idx]:.6e} ref {ref.flatten()[idx]:.6e} (rms {rms:.3e}, rtol {rtol:.1e})')

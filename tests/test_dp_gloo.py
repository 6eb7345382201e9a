"""Data-parallel gradient exchange on CPU: world_size 2 over gloo, exercising the bucket logic of plainlm_b200.dp
(the CUDA pack/unpack kernels are replaced by torch casts here — test doubles; the GPU tests cover the kernels)."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _FakeFlat:
  def __init__(self, total, buckets, seed):
    g = torch.Generator().manual_seed(seed)
    self.total = total
    self.buckets = buckets
    self.params = torch.randn(total, generator=g)
    self.grads = torch.randn(total, generator=g)
    self.shadow = torch.zeros(total, dtype=torch.bfloat16)

  def refresh_shadow(self):
    self.shadow.copy_(self.params)


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _worker(rank, world, port, wire, out):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from plainlm_b200.dp import GradReducer, broadcast_parameters

  total, buckets = 1000, [(0, 256), (256, 640), (640, 960), (960, 1000)]
  flat = _FakeFlat(total, buckets, seed=100 + rank)
  local = flat.grads.clone()
  broadcast_parameters(flat, src=0)
  red = GradReducer(
    flat,
    wire_dtype=torch.bfloat16 if wire == 'bf16' else torch.float32,
    pack=lambda s, d, sc: d.copy_((s * sc).to(d.dtype)),
    unpack=lambda s, d, sc: d.copy_(s.float() * sc),
  )
  for i in range(len(buckets)):
    red.bucket_ready(i)
  red.finish()
  gathered = [torch.zeros(total) for _ in range(world)]
  dist.all_gather(gathered, local)
  pg = [torch.zeros(total) for _ in range(world)]
  dist.all_gather(pg, flat.params)
  if rank == 0:
    torch.save({'reduced': flat.grads, 'locals': gathered, 'params': pg, 'launched': red.launched}, out)
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.parametrize('wire', ['bf16', 'fp32'])
def test_bucketed_allreduce_mean_world2(tmp_path, wire):
  out = str(tmp_path / 'res.pt')
  mp.spawn(_worker, args=(2, _free_port(), wire, out), nprocs=2, join=True)
  r = torch.load(out)
  mean = (r['locals'][0] + r['locals'][1]) / 2
  assert r['launched'] == 4
  if wire == 'fp32':
    assert torch.allclose(r['reduced'], mean, rtol=0, atol=1e-7)
  else:
    expect = ((r['locals'][0] / 2).bfloat16().float() + (r['locals'][1] / 2).bfloat16().float()).bfloat16().float()
    assert torch.equal(r['reduced'], expect)  # exactly: pre-scale, round to bf16, sum in bf16, widen
    assert (r['reduced'] - mean).abs().max() <= 2e-2 * mean.abs().max()
  assert torch.equal(r['params'][0], r['params'][1])  # rank-0 broadcast (engine.py:65 DDP constructor semantics)

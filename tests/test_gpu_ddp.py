"""Data-parallel equivalence on real GPUs (needs >= 2): W ranks x accum micro-steps == 1 rank x W*accum micro-steps
on the same rows, through TorchEngine + NCCL.  Run with `gpurun --gpus 2 -- pytest -m gpu tests/test_gpu_ddp.py`."""

import os
import socket
import time
from collections import namedtuple

import pytest
import torch
import torch.multiprocessing as mp

from oracle import plainlm_oracle as orc

pytestmark = pytest.mark.gpu
TINY = dict(vocab_size=256, d_model=128, n_layers=2, n_heads=2, seq_len=32, expand='8/3', mlp_class='glu',
            tie_embeddings=False, model='transformer')


def _cfg(**kw):
  return namedtuple('Cfg', kw.keys())(**kw)


def _engine_cfg(**over):
  base = dict(seq_len=32, grad_accumulation_steps=1, grad_clip=1.0, dtype='bfloat16', intra_doc_masking=False,
              resume=False, torch_compile=False, weight_decay=0.1, optim='adamw', lr=1e-3, beta1=0.9, beta2=0.95,
              fused_optim=True, scheduler=None, dampening=0.0, steps_budget=4)
  base.update(over)
  return base


def _rows():
  return torch.randint(0, 256, (8, 33), generator=torch.Generator().manual_seed(5))


def _worker(rank, world, port, fp32_wire, out):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                    LOCAL_RANK=str(rank))
  import torch.distributed as dist
  from plainlm_b200.data_utils import rank_partition
  from plainlm_b200.engine import TorchEngine
  from plainlm_b200.models import construct_model

  torch.cuda.set_device(rank)
  dist.init_process_group('nccl', device_id=torch.device(f'cuda:{rank}'))
  torch.manual_seed(100 + rank)  # ranks initialise differently; the engine must broadcast rank 0's weights
  model, _ = construct_model(_cfg(**TINY))
  if rank == 0:
    model.load_state_dict(orc.init_params(256, 128, 2, 2, seed=7))
  eng = TorchEngine(model, _cfg(**_engine_cfg(ddp_fp32_allreduce=fp32_wire)), f'cuda:{rank}', rank, None)
  rows = _rows()
  mine = rows[rank_partition(rows.shape[0], world, rank)]
  losses = []
  for s in range(2):
    losses.append(eng.step({'input_ids': mine[2 * s : 2 * s + 2]}).item())
  sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
  gathered = [None] * world
  dist.all_gather_object(gathered, float(sum(v.double().sum() for v in sd.values())))
  if rank == 0:
    torch.save({'sd': sd, 'losses': losses, 'checksums': gathered}, out)
  from plainlm_b200.torch_utils import destroy_ddp

  destroy_ddp()  # releases the captured graphs (they hold NCCL collectives) before the process group goes


@pytest.mark.parametrize('fp32_wire', [True, False])
def test_two_ranks_equal_one_rank_with_doubled_accumulation(tmp_path, fp32_wire):
  if torch.cuda.device_count() < 2:
    pytest.skip('needs 2 GPUs')
  from plainlm_b200.engine import TorchEngine
  from plainlm_b200.models import construct_model

  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  out = str(tmp_path / 'dp.pt')
  ctx = mp.spawn(_worker, args=(2, port, fp32_wire, out), nprocs=2, join=False)
  deadline = time.time() + 240
  while not ctx.join(timeout=5):
    if time.time() > deadline:
      for p in ctx.processes:
        p.kill()
      pytest.fail('data-parallel workers did not finish within 240 s')
  dp = torch.load(out)
  assert dp['checksums'][0] == dp['checksums'][1]  # replicas stay bit-identical

  init = orc.init_params(256, 128, 2, 2, seed=7)
  model, _ = construct_model(_cfg(**TINY))
  model.load_state_dict(init)
  eng = TorchEngine(model, _cfg(**_engine_cfg(grad_accumulation_steps=2)), 'cuda:0', None, None)
  rows = _rows()
  # DP step s consumed rows {4s..4s+3}: rank r took rows r, r+2 of them in one micro-batch of 2
  for s_ in range(2):
    blk = rows[4 * s_ : 4 * s_ + 4]
    eng.step({'input_ids': blk[0::2]})
    eng.step({'input_ids': blk[1::2]})
  tol = 2e-3 if fp32_wire else 0.15
  for k, v in model.state_dict().items():
    upd = (v.cpu() - init[k]).double().norm().item() + 1e-12
    diff = (v.cpu() - dp['sd'][k]).double().norm().item()
    assert diff <= tol * upd, (k, diff, upd)

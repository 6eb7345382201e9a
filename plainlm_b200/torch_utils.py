"""Process-group bring-up and teardown (reference: torch_utils.py), one process per GPU under torchrun."""

import os
import random

import numpy as np
import torch
from torch.distributed import destroy_process_group, init_process_group


def pytorch_setup(cfg):
  """reference: torch_utils.py:11-57 -> (local_rank, world_size, device, master_process).
  Differences: the CUDA device is actually selected (the reference builds a `torch.cuda.device` context it never
  enters, :21), and without a GPU this raises instead of falling back to the CPU."""
  ddp = int(os.environ.get('RANK', -1)) != -1
  if not torch.cuda.is_available():
    raise RuntimeError('plainlm_b200 needs a CUDA device (B200, sm_100a); there is no CPU path')
  if ddp:
    rank = int(os.environ['RANK'])
    local_rank = int(os.environ['LOCAL_RANK'])
    world_size = int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(local_rank)
    init_process_group(backend='nccl', device_id=torch.device(f'cuda:{local_rank}'))
    device = f'cuda:{local_rank}'
    master_process = rank == 0
    seed_offset = rank
  else:
    master_process, seed_offset, local_rank, world_size, device = True, 0, None, 1, 'cuda'

  random.seed(cfg.seed + seed_offset)
  np.random.seed(cfg.seed + seed_offset)
  torch.manual_seed(cfg.seed + seed_offset)

  torch.backends.cuda.matmul.allow_tf32 = getattr(cfg, 'cuda_matmul_allow_tf32', False)
  torch.backends.cudnn.allow_tf32 = getattr(cfg, 'cudnn_allow_tf32', True)
  if hasattr(cfg, 'set_memory_fraction'):
    torch.cuda.set_per_process_memory_fraction(cfg.set_memory_fraction, device=device)
  return local_rank, world_size, device, master_process


def destroy_ddp():
  if torch.distributed.is_initialized():
    torch.distributed.barrier()
    destroy_process_group()

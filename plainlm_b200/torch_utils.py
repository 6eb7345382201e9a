"""Process bring-up / teardown for one process per GPU under torchrun (reference: torch_utils.py).

`pytorch_setup(cfg) -> (local_rank, world_size, device, master_process)` and `destroy_ddp()` keep the reference's
contract.  Two deliberate differences: the CUDA device of the rank is actually made current (the reference constructs
a `torch.cuda.device` context at :21 and never enters it), and a machine without a GPU raises — there is no CPU path.
"""

import os
import random

import numpy as np
import torch
import torch.distributed as dist


def _torchrun_env():
  """(rank, local_rank, world_size) from the launcher's environment, or None for a single un-launched process."""
  if 'RANK' not in os.environ or int(os.environ['RANK']) < 0:
    return None
  return tuple(int(os.environ[k]) for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE'))


def _seed_everything(seed):
  for seeder in (random.seed, np.random.seed, torch.manual_seed):
    seeder(seed)


def pytorch_setup(cfg):
  if not torch.cuda.is_available():
    raise RuntimeError('plainlm_b200 needs a CUDA device (B200, sm_100a); there is no CPU path')
  env = _torchrun_env()
  if env is None:
    rank, local_rank, world_size, device = 0, None, 1, 'cuda'
  else:
    rank, local_rank, world_size = env
    device = f'cuda:{local_rank}'
    torch.cuda.set_device(local_rank)
    dist.init_process_group(backend='nccl', device_id=torch.device(device))
  _seed_everything(cfg.seed + rank)  # ranks draw different streams; rank 0's weights are broadcast by the engine

  torch.backends.cuda.matmul.allow_tf32 = getattr(cfg, 'cuda_matmul_allow_tf32', False)
  torch.backends.cudnn.allow_tf32 = getattr(cfg, 'cudnn_allow_tf32', True)
  fraction = getattr(cfg, 'set_memory_fraction', None)
  if fraction is not None:
    torch.cuda.set_per_process_memory_fraction(fraction, device=device)
  return local_rank, world_size, device, rank == 0


def destroy_ddp():
  """reference: torch_utils.py:62-65.  CUDA graphs that captured NCCL collectives (the data-parallel last micro-step,
  models/runtime.py) pin the communicator, and destroy_process_group() would wait for them forever: they go first."""
  if dist.is_initialized():
    from .models.runtime import release_all_graphs

    torch.cuda.synchronize()
    release_all_graphs()
    dist.barrier()
    dist.destroy_process_group()

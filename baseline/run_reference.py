"""Times the UNMODIFIED reference (baseline/_ref, installed by baseline/install_ref.sh) on this box.

  python baseline/run_reference.py --device cpu  --config 420m --steps K --warmup W [--micro-batch 1]
  python baseline/run_reference.py --device cuda --config 420m --steps K --warmup W

It drives the reference's own public API exactly as its train.py does (train.py:41-44,66-73):
`models.construct_model(cfg)` -> `engine.TorchEngine(model, cfg, device, local_rank, ckpt)` -> `engine.step(batch)` per
micro-batch, eager (`torch_compile: False`, TORCHDYNAMO_DISABLE=1: the RoPE function is unconditionally
`@torch.compile`d, models/embeddings.py:15, and inductor's default host compiler is broken in this image).  Nothing of
this repo's kernels, models or engine is imported here.  Prints ONE JSON line.

  device cpu : the reference's CPU path (fp32: engine/engine.py:73-75 gives the CPU a nullcontext), micro_batch 1
               (engine/engine.py:111 `targets.view(-1)` raises for B > 1 on the CPU: non-contiguous slice), all host
               threads.  One "step" = one micro-batch (fwd + bwd + clip + AdamW, grad_accumulation_steps = 1).
  device cuda: the reference's eager CUDA path under autocast(bf16) with the config's own micro-batch and accumulation
               (cuBLASLt, torch SDPA, ATen, torch fused AdamW) — the GPU-side point BASELINE.md §5 asks for.
"""

import argparse
import json
import os
import sys
import time
from collections import namedtuple

os.environ.setdefault('TORCHDYNAMO_DISABLE', '1')
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, '_ref')

CONFIGS = {
  '420m': dict(vocab_size=50280, d_model=1024, n_layers=24, n_heads=16, seq_len=2048, micro_batch_size=8,
               grad_accumulation_steps=4, intra_doc_masking=False),
  '124m_doc': dict(vocab_size=50280, d_model=768, n_layers=12, n_heads=12, seq_len=2048, micro_batch_size=8,
                   grad_accumulation_steps=2, intra_doc_masking=True),
  '1p5b': dict(vocab_size=50280, d_model=2048, n_layers=24, n_heads=16, seq_len=4096, micro_batch_size=4,
               grad_accumulation_steps=4, intra_doc_masking=False),
  'reduced': dict(vocab_size=50280, d_model=384, n_layers=6, n_heads=6, seq_len=512, micro_batch_size=4,
                  grad_accumulation_steps=2, intra_doc_masking=False),
}


def make_cfg(c, B, accum, dtype, optim, steps_budget):
  d = dict(model='transformer', vocab_size=c['vocab_size'], d_model=c['d_model'], n_layers=c['n_layers'],
           n_heads=c['n_heads'], seq_len=c['seq_len'], expand='8/3', mlp_class='glu', tie_embeddings=False, rms_norm=True,
           micro_batch_size=B, grad_accumulation_steps=accum, grad_clip=1.0, dtype=dtype,
           intra_doc_masking=c['intra_doc_masking'], resume=False, torch_compile=False, weight_decay=0.1, optim=optim,
           lr=3e-3, beta1=0.9, beta2=0.95, fused_optim=True, scheduler='warmup_cosine', warmup_steps=0.1,
           cooldown_steps=None, lr_start=0.0, lr_end=1e-5, lr_end_pct=None, steps_budget=max(steps_budget, 10),
           dampening=0.0, seed=100)
  return namedtuple('Config', d.keys())(**d)


def synth_docs(n_rows, T, seed=7):
  import random

  rng = random.Random(seed)
  out = []
  for _ in range(n_rows):
    left, dl = T + 1, []
    while left > 0:
      n = min(left, max(1, int(rng.lognormvariate(6.0, 0.8))))
      dl.append(n)
      left -= n
    out.append(dl)
  return out


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--device', default='cpu', choices=['cpu', 'cuda'])
  ap.add_argument('--config', default='420m', choices=sorted(CONFIGS))
  ap.add_argument('--steps', type=int, default=3)
  ap.add_argument('--warmup', type=int, default=1)
  ap.add_argument('--micro-batch', type=int, default=None)
  ap.add_argument('--optim', default='adamw')
  args = ap.parse_args()

  if not os.path.isdir(os.path.join(REF, 'engine')):
    print(json.dumps({'unavailable': 'baseline/_ref missing: run baseline/install_ref.sh where /root/reference exists'}))
    return
  sys.path.insert(0, REF)
  import torch

  if args.device == 'cpu':
    # torchrun exports OMP_NUM_THREADS=1 to its children: the CPU arm must use every host core
    torch.set_num_threads(os.cpu_count() or 1)
  import engine as ref_engine  # noqa: E402  (the reference's packages, from baseline/_ref)
  import models as ref_models  # noqa: E402

  assert os.path.realpath(ref_engine.__file__).startswith(os.path.realpath(REF)), ref_engine.__file__
  c = CONFIGS[args.config]
  T = c['seq_len']
  if args.device == 'cpu':
    B, accum, dtype, device = args.micro_batch or 1, 1, 'float32', 'cpu'
  else:
    B, accum, dtype, device = args.micro_batch or c['micro_batch_size'], c['grad_accumulation_steps'], 'bfloat16', 'cuda:0'
    torch.cuda.set_device(0)
  K, W = args.steps, args.warmup
  cfg = make_cfg(c, B, accum, dtype, args.optim, 2 * (K + W) + 4)
  torch.manual_seed(cfg.seed)
  model, _ = ref_models.construct_model(cfg)
  eng = ref_engine.TorchEngine(model, cfg, device, None, None)

  n_micro = (K + W) * accum
  rows = torch.randint(0, c['vocab_size'], (n_micro * B, T + 1), generator=torch.Generator().manual_seed(1234))
  docs = synth_docs(rows.shape[0], T) if c['intra_doc_masking'] else None

  def batch(i):
    b = {'input_ids': rows[i * B : (i + 1) * B]}
    if docs:
      b['docs_lengths'] = docs[i * B : (i + 1) * B]
    return b

  def sync():
    if device != 'cpu':
      torch.cuda.synchronize()

  times, loss = [], None
  for s in range(K + W):
    sync()
    t0 = time.perf_counter()
    for m in range(accum):
      loss = eng.step(batch(s * accum + m))
    lv = float(loss.item())  # device->host read of the step's result, as utils.log does (utils.py:171)
    sync()
    if s >= W:
      times.append(time.perf_counter() - t0)
  tot = sum(times)
  tokens = B * T * accum * len(times)
  out = {
    'value': round(tokens / tot, 1), 'unit': 'tokens/s', 'device': args.device, 'ms_per_step': round(1e3 * tot / len(times), 2),
    'steps': len(times), 'warmup': W, 'micro_batch': B, 'accum': accum, 'dtype': dtype, 'loss': round(lv, 4),
    'threads': torch.get_num_threads() if args.device == 'cpu' else None, 'host_cpus': os.cpu_count(),
    'torch': torch.__version__, 'impl': 'unmodified reference (baseline/_ref): construct_model + TorchEngine.step, eager',
    'seconds': round(tot, 2),
  }
  if device != 'cpu':
    out['gpu'] = torch.cuda.get_device_name(0)
    out['max_mem_gb'] = round(torch.cuda.max_memory_allocated() / 2**30, 1)
  print(json.dumps(out))


if __name__ == '__main__':
  main()

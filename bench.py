"""bench.py — tokens/sec & MFU of the plainLM 420M train step on B200 (BASELINE.json metric), plus the CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 420m|124m_doc|1p5b|reduced]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W

One "step" = one optimizer step of config/tr_420M_x8gpu.yaml on every rank: grad_accumulation_steps (4) micro-batches
of micro_batch_size (8) x seq_len (2048) tokens -> forward, backward, bucketed bf16 all-reduce (N>1), clip, AdamW.
Weak scaling: per-GPU work is fixed, `value` is the whole-job tokens/s.  Two timed regions of K steps each:
  value : micro-batches already resident in HBM when the clock starts
  e2e   : TorchEngine.step(batch) with HOST (pinned) batches — H2D of ids every micro-step and a D2H read of the loss
          every optimizer step inside the timed region
Rank 0 prints ONE JSON line.  `--impl reference` times the reference's own CPU path — the UNMODIFIED reference installed
into baseline/_ref by baseline/install_ref.sh (git-ignored, travels with the gpurun snapshot) — on the host cores for the
same config and metric; the bench line of our arm also carries `gpu_eager_reference`: that same reference run eagerly
on the same B200 (torch library kernels), in its own process before our engine allocates anything.
"""

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time
from collections import namedtuple

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

CONFIGS = {
  # config/tr_420M_x8gpu.yaml
  '420m': dict(vocab_size=50280, d_model=1024, n_layers=24, n_heads=16, seq_len=2048, micro_batch_size=8,
               grad_accumulation_steps=4, intra_doc_masking=False, optim='adamw'),
  # config/config_doc_mask.yaml (124M, document-masked)
  '124m_doc': dict(vocab_size=50280, d_model=768, n_layers=12, n_heads=12, seq_len=2048, micro_batch_size=8,
                   grad_accumulation_steps=2, intra_doc_masking=True, optim='adamw'),
  # BASELINE.json configs[4]: ~1.5B scale-up (H=32 keeps head_dim 64)
  '1p5b': dict(vocab_size=50280, d_model=2048, n_layers=24, n_heads=32, seq_len=4096, micro_batch_size=4,
               grad_accumulation_steps=4, intra_doc_masking=False, optim='adamw'),
  # BASELINE.json configs[0]: reduced CPU-runnable case
  'reduced': dict(vocab_size=50280, d_model=384, n_layers=6, n_heads=6, seq_len=512, micro_batch_size=4,
                  grad_accumulation_steps=2, intra_doc_masking=False, optim='adamw'),
}


def glu_hidden(d):
  return 256 * ((int(8 / 3 * d) + 255) // 256)


def flops_per_token(c, pairs_per_token=None):
  """F_tok = 6 N_mm + 6 L d T for causal attention (flash recompute not counted).  Document-masked:
  the attention term becomes 12 L d x (allowed (query, key) pairs per token) = 12 L d sum_docs len (len + 1) / 2 / tokens;
  full causal has (T + 1) / 2 ~ T / 2 pairs per token, which gives the 6 L d T above."""
  d, L, T, V = c['d_model'], c['n_layers'], c['seq_len'], c['vocab_size']
  n_mm = L * (4 * d * d + 3 * d * glu_hidden(d)) + d * V
  if pairs_per_token is None:
    return 6 * n_mm + 6 * L * d * T
  return 6 * n_mm + 12 * L * d * pairs_per_token


def doc_pairs_per_token(docs, T):
  """Allowed (query, key) pairs per token under intra-document causal masking: positions 0..T-1 of a row whose
  documents have the given lengths (summing to T + 1; the last position is only ever a target)."""
  pairs = tokens = 0
  for lengths in docs:
    left = T
    for n in lengths:
      n = min(n, left)
      pairs += n * (n + 1) // 2
      left -= n
      if left <= 0:
        break
    tokens += T
  return pairs / tokens


def make_cfgs(c, steps_budget):
  model_cfg = dict(vocab_size=c['vocab_size'], d_model=c['d_model'], n_layers=c['n_layers'], n_heads=c['n_heads'],
                   seq_len=c['seq_len'], expand='8/3', mlp_class='glu', tie_embeddings=False, model='transformer')
  train_cfg = dict(seq_len=c['seq_len'], grad_accumulation_steps=c['grad_accumulation_steps'], grad_clip=1.0,
                   dtype='bfloat16', intra_doc_masking=c['intra_doc_masking'], resume=False, torch_compile=False,
                   weight_decay=0.1, optim=c['optim'], lr=1e-4 if c['optim'] == 'signSGD' else 3e-3, beta1=0.9, beta2=0.95, fused_optim=True,
                   scheduler='warmup_cosine', warmup_steps=0.1, cooldown_steps=None, lr_start=0.0, lr_end=1e-5,
                   lr_end_pct=None, steps_budget=max(steps_budget, 10), dampening=0.0)
  nt = lambda d: namedtuple('Cfg', d.keys())(**d)  # noqa: E731
  return nt(model_cfg), nt(train_cfg), train_cfg


def synth_rows(n_rows, T, vocab):
  import torch

  return torch.randint(0, vocab, (n_rows, T + 1), generator=torch.Generator().manual_seed(1234))


def synth_docs(n_rows, T):
  """Document lengths per row summing to T+1 (log-normal, mean ~600 tokens), seeded."""
  import random

  rng = random.Random(7)
  out = []
  for _ in range(n_rows):
    left, dl = T + 1, []
    while left > 0:
      n = min(left, max(1, int(rng.lognormvariate(6.0, 0.8))))
      dl.append(n)
      left -= n
    out.append(dl)
  return out


class ClockSampler:
  FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
            'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
            'clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.path = f'/tmp/plm_clocks_{os.getpid()}.csv'
    self.index = index
    self.proc = None

  def start(self):
    try:
      self.f = open(self.path, 'w')
      self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.FIELDS}',
                                    '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.f,
                                   stderr=subprocess.DEVNULL)
    except OSError:
      self.proc = None

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    self.proc.terminate()
    self.proc.wait()
    self.f.close()
    sm, mx, reasons, power = [], [], set(), []
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for line in open(self.path):
      parts = [p.strip() for p in line.split(',')]
      if len(parts) < 7:
        continue
      try:
        sm.append(float(parts[0]))
        mx.append(float(parts[1]))
        power.append(float(parts[2]))
      except ValueError:
        continue
      for nm, val in zip(names, parts[3:7]):
        if val.lower().startswith('active'):
          reasons.add(nm)
    os.unlink(self.path)
    return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
            'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


def load_peaks():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    d = json.load(open(p))
    return {'hbm_gbs': d['hbm_gbs'], 'tf_burst': d['bf16_tflops'], 'tf_sustained': d['bf16_tflops_sustained'],
            'source': 'measured'}
  return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'source': 'fallback'}


def op_work(name, tag, c, pairs_per_token=None):
  """Algorithmic work of one op call: ('tensor', flops) or ('hbm', bytes) — DESIGN.md §kernels."""
  if name in ('gemm', 'lmhead_ce_fwd'):  # lmhead_ce_fwd = the LM-head GEMM with the cross-entropy forward in its epilogue
    M, N, K = tag[0], tag[1], tag[2]
    return 'tensor', 2.0 * M * N * K
  if name in ('attn_fwd', 'attn_bwd'):
    B, T, H, hd = tag[:4]
    pairs = (T + 1) / 2 if (pairs_per_token is None or not tag[4]) else pairs_per_token
    fwd = 4.0 * B * H * hd * T * pairs  # allowed pairs x (QK^T + PV) x 2 flop
    return 'tensor', fwd if name == 'attn_fwd' else 2.0 * fwd
  n = tag[0] if tag else 0
  # bytes per element of the op's FIRST tensor argument: swiglu_fwd's is u [M, 2F] (2 B read + 1 B written per u
  # element = 6 B per hidden element); swiglu_bwd's is dh [M, F] (dh 2 + u 4 + du 4 = 10 B per hidden element)
  per_elt = {'rmsnorm_fwd': 6, 'rmsnorm_bwd': 16, 'swiglu_fwd': 3, 'swiglu_bwd': 10, 'embed_fwd': 8, 'embed_bwd': 12,
             'ce_fwd_bwd': 6, 'ce_grad': 4, 'sumsq': 4, 'adamw_step': 30, 'signsgd_step': 22, 'cast_f32_bf16': 6,
             'cast_bf16_f32': 6, 'colsum_accum': 4}.get(name, 4)
  return 'hbm', float(per_elt) * n


def run_ours(args):
  import torch
  import torch.distributed as dist

  from plainlm_b200 import ops
  from plainlm_b200.data_utils import rank_partition, seg_start_from_docs_lengths
  from plainlm_b200.engine import TorchEngine
  from plainlm_b200.models import construct_model

  c = CONFIGS[args.config]
  rank = int(os.environ.get('RANK', 0))
  world = int(os.environ.get('WORLD_SIZE', 1))
  local_rank = int(os.environ.get('LOCAL_RANK', 0))
  if world != args.gpus:
    if world == 1 and args.gpus > 1:
      raise SystemExit('bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)')
  torch.cuda.set_device(local_rank)
  device = f'cuda:{local_rank}'
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device(device))

  K, W = args.steps, max(args.warmup, 3)
  accum, B, T = c['grad_accumulation_steps'], c['micro_batch_size'], c['seq_len']
  # The unmodified reference, eager, on this same GPU — in its own process, before this one allocates:
  # torch's library kernels (cuBLASLt, SDPA, ATen, fused AdamW), same config, same synthetic tokens.  Reported beside
  # our number; not part of any timed region.
  gpu_ref = None
  if rank == 0 and world == 1 and not args.no_gpu_reference:
    gpu_ref = reference_subprocess(args.config, 'cuda', steps=3, warmup=2, optim=c['optim'])
  mcfg, tcfg, _ = make_cfgs(c, 2 * (K + W) + 4)
  torch.manual_seed(100 + rank)  # reference: torch_utils.py:35-37 (rank 0's weights are broadcast by the engine)
  model, _ = construct_model(mcfg) if rank == 0 or True else (None, None)
  engine = TorchEngine(model, tcfg, device, local_rank if world > 1 else None, None)

  # synthetic tokens, partitioned across ranks like DistributedSampler(shuffle=False, drop_last=True)
  micro_total = (W + K) * accum
  rows = synth_rows(micro_total * B * world, T, c['vocab_size'])
  mine = rows[rank_partition(rows.shape[0], world, rank)]
  docs = synth_docs(rows.shape[0], T) if c['intra_doc_masking'] else None
  my_docs = [docs[i] for i in rank_partition(rows.shape[0], world, rank)] if docs else None

  def host_batch(i):
    b = {'input_ids': mine[i * B : (i + 1) * B]}
    if my_docs:
      b['docs_lengths'] = my_docs[i * B : (i + 1) * B]
    return b

  # device-resident copies for the `value` region
  dev_in = mine[:, :T].contiguous().to(device)
  dev_tg = mine[:, 1 : T + 1].contiguous().to(device)
  dev_seg = None
  if my_docs:
    dev_seg = seg_start_from_docs_lengths(my_docs, T).to(device)

  def dev_step(i):
    seg = dev_seg[i * B : (i + 1) * B].reshape(-1) if dev_seg is not None else None
    return engine.step_device(dev_in[i * B : (i + 1) * B], dev_tg[i * B : (i + 1) * B], seg)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def max_over_ranks(ms):
    if world == 1:
      return ms
    t = torch.tensor([ms], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()

  # ---------------------------------------------------------------- region 1: device-resident inputs (`value`)
  DP_CHECK_STEPS = 2  # optimizer steps after which the data-parallel state is snapshotted for the equivalence check
  dp_snap = None
  for s in range(W):
    for m in range(accum):
      dev_step(s * accum + m)
    if world > 1 and rank == 0 and s + 1 == DP_CHECK_STEPS:
      dp_snap = engine.rt.flat.params.clone()
  barrier()
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  launches0 = ops.LAUNCHES
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  e0.record()
  for s in range(W, W + K):
    for m in range(accum):
      loss = dev_step(s * accum + m)
  e1.record()
  barrier()
  ms_value = max_over_ranks(e0.elapsed_time(e1))
  launches = ops.LAUNCHES - launches0
  clocks = sampler.stop() if rank == 0 else None
  last_loss = loss.item()
  engine.check_nan(wait=True)

  # ---------------------------------------------------------------- region 2: host batches through the public API
  for s in range(min(W, 2)):
    for m in range(accum):
      engine.step(host_batch(s * accum + m))
  barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for s in range(W, W + K):
    for m in range(accum):
      loss = engine.step(host_batch(s * accum + m))
    _ = loss.item()  # device->host read of the step's result
  e1.record()
  barrier()
  ms_e2e = max_over_ranks(e0.elapsed_time(e1))

  # ---------------------------------------------------------------- one instrumented step: per-kernel CUDA events
  prof = ops.Profiler(sync=bool(os.environ.get('PLM_BENCH_SYNC_PROFILE')))
  ops.set_profiler(prof)
  engine.use_cuda_graphs = False  # per-kernel events need eager launches
  os.environ['PLM_NO_SIDE_STREAM'] = '1'  # ... and no overlap: weight-gradient GEMMs back on the main stream
  for m in range(accum):
    dev_step(m)
  summ = prof.summary()
  ops.set_profiler(None)
  os.environ.pop('PLM_NO_SIDE_STREAM', None)

  # ---------------------------------------------------------------- data-parallel equivalence (N > 1)
  # (1) every replica holds bit-identical weights after the run; (2) N ranks x accum micro-batches == ONE rank running
  # N x accum micro-batches on the same rows (reference: DDP averages the per-rank gradients, engine.py:64-65,104-105).
  dp_equiv = None
  if world > 1:
    flat = engine.rt.flat.params
    chk = flat.view(torch.int32).to(torch.int64).sum().reshape(1)
    allchk = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allchk, chk)
    identical = all(int(c) == int(allchk[0]) for c in allchk)
    if rank == 0:
      _, _, tdict1 = make_cfgs(c, 2 * (K + W) + 4)
      cfg1 = dict(tdict1, grad_accumulation_steps=accum * world, data_parallel=False, cuda_graphs=True)
      torch.manual_seed(100)
      model1, _ = construct_model(mcfg)
      eng1 = TorchEngine(model1, namedtuple('Cfg', cfg1.keys())(**cfg1), device, None, None)
      init1 = eng1.rt.flat.params.clone()
      for s in range(DP_CHECK_STEPS):
        for m in range(accum):
          i = s * accum + m  # global micro-step i: rank r consumed rows {(i B + b) W + r}
          for r in range(world):
            blk = rows[[(i * B + b) * world + r for b in range(B)]]
            seg1 = None
            if docs:
              seg1 = seg_start_from_docs_lengths([docs[(i * B + b) * world + r] for b in range(B)], T).to(device).reshape(-1)
            eng1.step_device(blk[:, :T].contiguous().to(device), blk[:, 1 : T + 1].contiguous().to(device), seg1)
      torch.cuda.synchronize()
      upd = (eng1.rt.flat.params - init1).double().norm().item()
      diff = (eng1.rt.flat.params - dp_snap).double().norm().item()
      dp_equiv = {'replicas_bit_identical': identical, 'optimizer_steps': DP_CHECK_STEPS,
                  'diff_over_update_norm': round(diff / max(upd, 1e-30), 5),
                  'what': f'{world} ranks x {accum} micro-batches vs 1 rank x {accum * world} micro-batches, same rows and init; '
                          'bf16 gradient wire (pre-scaled 1/world): tolerance 0.15 of the update norm'}
      del eng1, model1, init1, dp_snap
      torch.cuda.empty_cache()
      assert identical, 'data-parallel replicas diverged'
      assert dp_equiv['diff_over_update_norm'] <= 0.15, dp_equiv
    dist.barrier()

  tokens_per_step = B * T * accum * world
  value = tokens_per_step * K / (ms_value / 1e3)
  e2e_value = tokens_per_step * K / (ms_e2e / 1e3)
  ppt = doc_pairs_per_token(docs, T) if docs else None  # document masking: count only the allowed (query, key) pairs
  # document masking sends only the lengths (int32) and B + 1 offsets per micro-batch; the segment map is built on device
  doc_h2d_bytes = 0
  if my_docs:
    n_micro_timed = K * accum
    used = my_docs[W * accum * B : (W + K) * accum * B]
    doc_h2d_bytes = (4 * sum(len(dl) for dl in used) + 4 * (B + 1) * n_micro_timed) / K
  ftok = flops_per_token(c, ppt)
  peaks = load_peaks()

  if rank == 0 and os.environ.get('PLM_BENCH_DETAIL'):  # per-(kernel, shape) table of the instrumented step
    rows = []
    for (name, tag), (cnt, ms) in sorted(summ.items(), key=lambda kv: -kv[1][1]):
      kind, work = op_work(name, tag, c, ppt)
      rate = work * cnt / (ms / 1e3) / (1e12 if kind == 'tensor' else 1e9)
      rows.append(f'{name:22s} {str(tag):44s} calls {cnt:4d}  total {ms:8.3f} ms  avg {ms / cnt * 1e3:8.1f} us  '
                  f'{rate:8.1f} {"TF/s" if kind == "tensor" else "GB/s"}')
    torch.cuda.synchronize()
    per = {}
    for name, tag, e0, e1 in prof.records:
      per.setdefault(name, []).append(e0.elapsed_time(e1) * 1e3)
    for name, ts in per.items():
      srt = sorted(ts)
      rows.append(f'{name:22s} per-call us: min {srt[0]:.1f} median {srt[len(srt) // 2]:.1f} max {srt[-1]:.1f}; '
                  f'largest: {[round(x) for x in srt[-6:]]}; first calls: {[round(x) for x in ts[:4]]}')
    with open(os.environ['PLM_BENCH_DETAIL'], 'w') as f:
      f.write('\n'.join(rows) + '\n')

  if rank == 0:
    total_ms = sum(v[1] for v in summ.values())
    by_kernel = {}
    for (name, tag), (cnt, ms) in summ.items():
      kind, work = op_work(name, tag, c, ppt)
      rec = by_kernel.setdefault(name, {'ms': 0.0, 'work': 0.0, 'calls': 0, 'kind': kind})
      rec['ms'] += ms
      rec['work'] += work * cnt
      rec['calls'] += cnt
    top = max(by_kernel.items(), key=lambda kv: kv[1]['ms'])
    name, rec = top
    if rec['kind'] == 'tensor':
      achieved = rec['work'] / (rec['ms'] / 1e3) / 1e12
      peak, unit, bound = peaks['tf_sustained'], 'TFLOP/s', 'tensor'
    else:
      achieved = rec['work'] / (rec['ms'] / 1e3) / 1e9
      peak, unit, bound = peaks['hbm_gbs'], 'GB/s', 'hbm'
    traffic = None  # DRAM bytes per launch of the dominant kernel from the committed ncu pass (profiles/traffic.json)
    try:
      with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'traffic.json')) as f:
        tj = json.load(f)
      if args.config == '420m' and tj.get(name, {}).get('dram_bytes_per_call'):
        traffic = round(tj[name]['dram_bytes_per_call'])
    except (OSError, ValueError):
      pass
    roofline = {'bound': bound, 'kernel': name, 'achieved': round(achieved, 1), 'peak': peak, 'unit': unit,
                'frac': round(achieved / peak, 4), 'traffic': traffic, 'peak_source': peaks['source'] + ' (sustained)',
                'share_of_step': round(rec['ms'] / total_ms, 3), 'avg_launch_ms': round(rec['ms'] / rec['calls'], 4),
                'by_kernel_ms': {k: round(v['ms'], 2) for k, v in sorted(by_kernel.items(), key=lambda kv: -kv[1]['ms'])},
                'by_kernel_frac': {k: round((v['work'] / (v['ms'] / 1e3) / (1e12 if v['kind'] == 'tensor' else 1e9)) /
                                            (peaks['tf_sustained'] if v['kind'] == 'tensor' else peaks['hbm_gbs']), 3)
                                   for k, v in by_kernel.items() if v['ms'] > 0}}
    cpu = cpu_baseline(args.config, c, steps=3, warmup=1) if (world == 1 and not args.no_cpu_baseline) else None
    out = {
      'metric': 'tokens/sec (420M GLU/RoPE LM train step)' if args.config == '420m' else f'tokens/sec ({args.config} train step)',
      'value': round(value, 1), 'unit': 'tokens/s', 'n_gpus': world, 'steps': K, 'warmup': W,
      'ms_per_step': round(ms_value / K, 3), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
      'dtype': 'bf16', 'data': 'synthetic',
      'config': {'workload': f'plainLM {args.config}: d{c["d_model"]} L{c["n_layers"]} H{c["n_heads"]} T{T} V{c["vocab_size"]}, '
                             f'micro_batch {B} x accum {accum} per GPU, {c["optim"]} + clip 1.0, '
                             f'{"document-masked" if c["intra_doc_masking"] else "causal"} attention',
                 'global_batch_tokens': tokens_per_step, 'seq_len': T, 'parallelism': f'dp{world}',
                 'l2': 'working set (>=14 GB of activations + 5 GB of optimizer state per step) >> 126 MB L2: no flush needed'},
      'mfu': {'flops_per_token': int(ftok), 'attention_pairs_per_token': round(ppt if ppt is not None else (T + 1) / 2, 1),
              'model_tflops': round(value * ftok / world / 1e12, 1),
              'vs_nominal_2250': round(value * ftok / world / 2250e12, 4),
              'vs_measured_sustained': round(value * ftok / world / (peaks['tf_sustained'] * 1e12), 4)},
      'e2e': {'value': round(e2e_value, 1), 'unit': 'tokens/s', 'ms_per_step': round(ms_e2e / K, 3),
              'h2d_bytes_per_step': int(2 * B * T * 8 * accum + doc_h2d_bytes),
              'd2h_bytes_per_step': int(4 * accum + 4)},
      'gpu_launches': int(launches),
      'clocks': clocks,
      'roofline': roofline,
      'final_loss': round(last_loss, 4),
    }
    if cpu is not None:
      out['cpu_baseline'] = cpu
    if gpu_ref is not None:
      out['gpu_eager_reference'] = gpu_ref
    if dp_equiv is not None:
      out['dp_equiv'] = dp_equiv
    print(json.dumps(out), flush=True)
  if world > 1:
    # teardown: graphs that captured NCCL collectives pin the communicator (destroy_process_group would wait for them),
    # so they are released first; a watchdog keeps a wedged teardown from outliving the measurement
    import threading

    from plainlm_b200.torch_utils import destroy_ddp

    del engine
    t = threading.Thread(target=destroy_ddp, daemon=True)
    t.start()
    t.join(60)
    sys.stdout.flush()
    if t.is_alive():
      os._exit(0)


def reference_subprocess(config, device, steps, warmup, optim='adamw', timeout=1500):
  """Runs baseline/run_reference.py (the unmodified reference from baseline/_ref) in its own process; returns its JSON
  dict, or {'unavailable': why}."""
  script = os.path.join(ROOT, 'baseline', 'run_reference.py')
  cmd = [sys.executable, script, '--device', device, '--config', config, '--steps', str(steps), '--warmup', str(warmup),
         '--optim', optim]
  env = dict(os.environ, TORCHDYNAMO_DISABLE='1')
  for k in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'RANK', 'LOCAL_RANK', 'WORLD_SIZE', 'MASTER_ADDR', 'MASTER_PORT'):
    env.pop(k, None)  # torchrun pins OMP_NUM_THREADS=1 for its children; the reference process is single, un-distributed
  try:
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
  except subprocess.TimeoutExpired:
    return {'unavailable': f'reference run exceeded {timeout} s'}
  for line in reversed(p.stdout.splitlines()):
    if line.startswith('{'):
      return json.loads(line)
  return {'unavailable': f'reference run failed (rc {p.returncode}): {p.stderr.strip()[-300:]}'}


def cpu_baseline(config, c, steps, warmup):
  """The reference's CPU path timed on this box's host cores on a bounded sample: `steps` micro-batches of ONE sequence
  (fwd + bwd + clip + AdamW each) after `warmup` untimed ones.  kind "reference": the unmodified reference from
  baseline/_ref (construct_model + TorchEngine.step, fp32 as engine.py:73-75 dictates for the CPU); kind "port": the
  oracle restatement, only when baseline/_ref is not on the box."""
  T = c['seq_len']
  res = reference_subprocess(config, 'cpu', steps, warmup, optim=c['optim'])
  if 'unavailable' not in res:
    return {'value': res['value'], 'unit': 'tokens/s', 'cores': res['threads'], 'host_cpus': res['host_cpus'],
            'kind': 'reference',
            'sample': f"{res['steps']} micro-batch(es) of 1 x {T} tokens after {res['warmup']} warm-up, fp32, "
                      f"fwd+bwd+clip+AdamW through the reference's TorchEngine.step, {res['seconds']:.1f} s",
            'ms_per_step': res['ms_per_step'], 'loss': res['loss']}
  import torch

  from oracle import plainlm_oracle as orc

  torch.set_num_threads(os.cpu_count() or 1)
  _, _, tdict = make_cfgs(c, 10)
  cfg = dict(tdict, grad_accumulation_steps=1, n_heads=c['n_heads'], dtype='float32', intra_doc_masking=False)
  params = orc.init_params(c['vocab_size'], c['d_model'], c['n_layers'], c['n_heads'], seed=100)
  tr = orc.OracleTrainer(params, cfg, 'fp32')
  rows = synth_rows(steps + warmup, T, c['vocab_size'])
  times = []
  for i in range(steps + warmup):
    t0 = time.perf_counter()
    loss = float(tr.step({'input_ids': rows[i : i + 1]}))
    dt = time.perf_counter() - t0
    if i >= warmup:
      times.append(dt)
  tot = sum(times)
  return {'value': round(T * len(times) / tot, 1), 'unit': 'tokens/s', 'cores': torch.get_num_threads(),
          'host_cpus': os.cpu_count(), 'kind': 'port', 'why_port': res['unavailable'],
          'sample': f'{len(times)} micro-batch(es) of 1 x {T} tokens after {warmup} warm-up, fp32, fwd+bwd+clip+AdamW, {tot:.1f} s',
          'ms_per_step': round(1e3 * tot / len(times), 1), 'loss': round(loss, 4)}


def run_reference(args):
  """--impl reference: the reference's own CPU implementation of the path — the unmodified reference from baseline/_ref
  (the oracle port only if that is missing) — on all host threads.  Under torchrun only rank 0 works."""
  rank = int(os.environ.get('RANK', 0))
  if rank != 0:
    return
  c = CONFIGS[args.config]
  K, W = args.steps, args.warmup
  res = cpu_baseline(args.config, c, steps=K, warmup=W)
  T = c['seq_len']
  out = {
    'impl': 'reference',
    'metric': 'tokens/sec (420M GLU/RoPE LM train step)' if args.config == '420m' else f'tokens/sec ({args.config} train step)',
    'value': res['value'], 'unit': 'tokens/s', 'n_gpus': args.gpus, 'steps': K, 'warmup': W,
    'ms_per_step': res['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
    'dtype': 'f32', 'data': 'synthetic',
    'config': {'workload': f'plainLM {args.config}: d{c["d_model"]} L{c["n_layers"]} H{c["n_heads"]} T{T} V{c["vocab_size"]}; '
                           'each step = a bounded sample of the workload: 1 micro-batch of 1 sequence on the host CPU',
               'seq_len': T, 'parallelism': 'cpu'},
    'cpu_baseline': res,
    'e2e': {'value': res['value'], 'unit': 'tokens/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    'gpu_launches': 0,
  }
  print(json.dumps(out))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=None)
  ap.add_argument('--warmup', type=int, default=None)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--config', default='420m', choices=sorted(CONFIGS))
  ap.add_argument('--optim', default=None, choices=['adamw', 'signSGD', 'sgd', 'nadamw'], help='override the config optimizer')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-gpu-reference', action='store_true')
  args = ap.parse_args()
  if args.optim:
    CONFIGS[args.config] = dict(CONFIGS[args.config], optim=args.optim)
  if args.impl == 'reference':
    args.steps = 3 if args.steps is None else args.steps
    args.warmup = 1 if args.warmup is None else args.warmup
    run_reference(args)
  else:
    args.steps = 10 if args.steps is None else args.steps
    args.warmup = 3 if args.warmup is None else args.warmup
    run_ours(args)


if __name__ == '__main__':
  main()

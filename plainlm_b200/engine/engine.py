"""TorchEngine: one training micro-step on the B200 kernels (reference: engine/engine.py).

Same constructor, attributes (.model .optimizer .scheduler .scaler .micro_steps) and `step(batch)` / `eval(loader)`
contract as the reference, so the reference's train.py, utils.log and checkpoint_utils.py drive it unchanged.
What changes underneath (SURVEY.md §3.2 -> here):
  * fwd + loss + bwd is ONE hand-scheduled kernel sequence (models/runtime.py): no autocast, no autograd graph,
    no dense (B,T,T) mask — document masking travels as int32 segment starts;
  * DDP is replaced by a bucketed bf16 all-reduce overlapped with backward (dp.py), only on the last micro-step
    of an accumulation cycle (reference: engine.py:104-105);
  * clip_grad_norm_ + optimizer.step() is one reduction pass + one flat update kernel, clip coefficient computed on
    the device (no host sync);
  * the per-micro-step `isnan` host sync (reference: engine.py:116) is kept in meaning but moved: every micro-step's
    loss is copied to its own pinned slot asynchronously, and ALL slots of the accumulation cycle are checked (one host
    sync) at the accumulation boundary BEFORE the optimizer update — a NaN loss raises before any weight, moment or
    bf16 shadow changes, exactly as in the reference.  Independently the update kernels skip themselves on the device
    when the gradient norm is not finite, so a poisoned gradient cannot reach the master weights even if the caller
    swallows the exception.
"""

import os

import torch
from torch import distributed as dist

from .. import ops
from ..data_utils import pack_docs_lengths
from ..dp import GradReducer, broadcast_parameters
from ..models import get_param_groups
from ..optim import intialize_optimizer, initialize_scheduler
from ..optim.flat import GradClip, grad_sumsq, sumsq_workspace


class _HostStaging:
  """Double-buffered pinned staging for the per-step host->device copies (reference: engine.py:27-30 pins and copies
  synchronously on the main thread every micro-step)."""

  def __init__(self, device, slots=2):
    self.device = device
    self.slots = slots
    self.k = 0
    self.bufs = {}

  def put(self, name, cpu_tensor):
    key = (name, self.k, tuple(cpu_tensor.shape), cpu_tensor.dtype)
    rec = self.bufs.get(key)
    if rec is None:
      pinned = torch.empty(cpu_tensor.shape, dtype=cpu_tensor.dtype, pin_memory=True)
      dev = torch.empty(cpu_tensor.shape, dtype=cpu_tensor.dtype, device=self.device)
      rec = [pinned, dev, None]
      self.bufs[key] = rec
    pinned, dev, ev = rec
    if ev is not None:
      ev.synchronize()  # the copy that last used this slot finished long ago; cheap guard
    pinned.copy_(cpu_tensor)
    dev.copy_(pinned, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    rec[2] = ev
    return dev

  def put_prefix(self, name, cpu_tensor, capacity):
    """Like put() for a 1-D tensor whose length varies from step to step: one pinned / device pair of `capacity`
    elements per slot, only the used prefix is copied."""
    key = (name, self.k, capacity, cpu_tensor.dtype)
    rec = self.bufs.get(key)
    if rec is None:
      rec = [torch.empty(capacity, dtype=cpu_tensor.dtype, pin_memory=True),
             torch.empty(capacity, dtype=cpu_tensor.dtype, device=self.device), None]
      self.bufs[key] = rec
    pinned, dev, ev = rec
    if ev is not None:
      ev.synchronize()
    n = cpu_tensor.numel()
    pinned[:n].copy_(cpu_tensor)
    dev[:n].copy_(pinned[:n], non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    rec[2] = ev
    return dev[:n]

  def advance(self):
    self.k = (self.k + 1) % self.slots


class TorchEngine(torch.nn.Module):
  """A module containing model, optimizer, scheduler, grad scaler; wraps a training step with grad accumulation."""

  def __init__(self, model, cfg, device, local_rank, ckpt):
    super().__init__()
    self.micro_steps = 0
    self.accumulated_samples = 0
    self.seq_len = cfg.seq_len
    self.accumulation_steps = cfg.grad_accumulation_steps
    self.grad_clip = cfg.grad_clip
    self.dtype = cfg.dtype
    self.intra_doc_masking = getattr(cfg, 'intra_doc_masking', False)
    self.use_cuda_graphs = getattr(cfg, 'cuda_graphs', True)  # optional key: replay the micro-step from a CUDA graph
    # optional key: also capture the data-parallel last micro-step (bucket packs + NCCL all-reduces) in a CUDA graph
    self.dp_graph = getattr(cfg, 'dp_cuda_graph', os.environ.get('PLM_DP_GRAPH', '1') != '0')
    self.device = device
    if 'cuda' not in str(device):
      raise RuntimeError('plainlm_b200.TorchEngine needs a CUDA device (B200); there is no CPU path')
    if self.dtype != 'bfloat16':
      raise NotImplementedError(
        f"dtype '{self.dtype}': the B200 kernels implement the bfloat16 policy (bf16 GEMM operands, fp32 accumulation, "
        'fp32 master weights); float16+GradScaler / float32 are SURVEY.md §8(f) N4'
      )

    if cfg.resume:
      model.load_state_dict(ckpt['state_dict'])
      self.micro_steps = ckpt['step'] * cfg.grad_accumulation_steps

    self.model = model.to(device)
    self.rt = self.model.runtime()
    self.reducer = None
    # optional cfg key `data_parallel: False`: build a replica-less engine inside a distributed job (bench.py uses it
    # for the N ranks == 1 rank x N*accum equivalence check)
    if dist.is_initialized() and getattr(cfg, 'data_parallel', True):
      broadcast_parameters(self.rt.flat)
      wire = torch.float32 if getattr(cfg, 'ddp_fp32_allreduce', False) else torch.bfloat16
      self.reducer = GradReducer(self.rt.flat, wire_dtype=wire)

    # bf16 needs no loss scaling; a disabled scaler keeps checkpoint_utils.py's 'scaler' entry round-tripping
    self.scaler = torch.amp.GradScaler(enabled=False)

    param_groups = get_param_groups(model, cfg.weight_decay)
    self.optimizer = intialize_optimizer(param_groups, cfg)
    self.scheduler = initialize_scheduler(self.optimizer, cfg)
    if cfg.resume:
      self.optimizer.load_state_dict(ckpt['optimizer'])
      self.scheduler.load_state_dict(ckpt['scheduler'])
      self.scaler.load_state_dict(ckpt['scaler'])

    dev = self.rt.flat.params.device
    self._staging = _HostStaging(dev)
    self._seg_ring = {}
    self._sumsq_ws = sumsq_workspace(dev)
    self._gnorm_sq = torch.zeros(1, device=dev, dtype=torch.float32)
    n_slots = max(int(self.accumulation_steps), 1)  # one pinned slot per micro-step of an accumulation cycle
    self._loss_host = torch.zeros(n_slots, dtype=torch.float32, pin_memory=True)
    self._loss_events = [None] * n_slots

  # ------------------------------------------------------------------------------------------ batch staging
  def _move_to_device(self, batch):
    """reference: engine.py:13-34.  Slicing is done on the host (bit-exact integers).  The dense (B, T, T) mask the
    reference builds on the device with a Python loop per document (engine.py:19-23) is replaced by int32 segment
    starts: only the document lengths cross PCIe (a few hundred bytes) and plm_seg_start_from_lengths expands them on
    the device into seg_start[B*T]."""
    ids = batch['input_ids']
    T = self.seq_len
    B = ids.shape[0]
    seg = None
    if ids.is_cuda:
      inputs, targets = ids[:, :T].contiguous(), ids[:, 1 : T + 1].contiguous()
    else:
      inputs = self._staging.put('inputs', ids[:, :T])
      targets = self._staging.put('targets', ids[:, 1 : T + 1])
    if self.intra_doc_masking:
      lengths, offsets = pack_docs_lengths(batch['docs_lengths'], T)
      d_len = self._staging.put_prefix('doc_lengths', lengths, B * (T + 1))
      d_off = self._staging.put_prefix('doc_offsets', offsets, B + 1)
      seg = self._seg_buf(B, T)
      ops.seg_start_from_lengths(d_len, d_off, seg, B, T)
    self._staging.advance()
    return inputs, targets, seg

  def _seg_buf(self, B, T):
    """Segment maps live in a small ring of device buffers (a micro-step's map must outlive the next staging)."""
    key = (B, T)
    ring = self._seg_ring.get(key)
    if ring is None:
      ring = [[torch.empty(B * T, dtype=torch.int32, device=self.rt.flat.params.device) for _ in range(3)], 0]
      self._seg_ring[key] = ring
    ring[1] = (ring[1] + 1) % 3
    return ring[0][ring[1]]

  # ------------------------------------------------------------------------------------------ NaN guard
  def _record_loss(self, loss, k):
    """Queue the device->host copy of micro-step k's loss (k = index inside the accumulation cycle).  A slot is only
    reused after check_nan(wait=True) has consumed it at the accumulation boundary."""
    if self._loss_events[k] is not None:  # defensive: never overwrite an unchecked loss
      self._loss_events[k].synchronize()
      self._consume_loss(k)
    self._loss_host[k : k + 1].copy_(loss.reshape(1), non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    self._loss_events[k] = ev

  def _consume_loss(self, k):
    self._loss_events[k] = None
    if torch.isnan(self._loss_host[k]):
      raise ValueError('Train loss is nan')

  def check_nan(self, wait=False):
    """Raise ValueError('Train loss is nan') (reference: engine.py:116-117) for any completed micro-step; with
    wait=True, for every micro-step launched so far (one host sync)."""
    for k, ev in enumerate(self._loss_events):
      if ev is None:
        continue
      if wait:
        ev.synchronize()
      if ev.query():
        self._consume_loss(k)

  # ------------------------------------------------------------------------------------------ train step
  def step(self, batch):
    """Wraps a fwd pass, bwd pass, and (at the accumulation boundary) an optimization step."""
    inputs, targets, seg = self._move_to_device(batch)
    return self.step_device(inputs, targets, seg)

  def step_device(self, inputs, targets, seg_start=None):
    """Same as step() for a micro-batch already on the device (int64 [B,T] inputs / targets, int32 [B*T] segments)."""
    self.model.train()
    self.check_nan()
    self.micro_steps += 1
    self.accumulated_samples += 1
    if self.accumulated_samples == 1:
      self.rt.flat.zero_grads()
    last = self.accumulated_samples == self.accumulation_steps
    on_bucket = self.reducer.bucket_ready if (self.reducer is not None and last) else None

    dp_last = self.reducer is not None and last
    if self.use_cuda_graphs and (not dp_last or self.dp_graph):
      # the data-parallel last micro-step is captured with its bucket hooks (pack + NCCL all-reduce on the comm stream)
      loss_val = self.rt.graphed_loss_and_backward(inputs, targets, seg_start,
                                                   grad_scale=1.0 / self.accumulation_steps,
                                                   reducer=self.reducer if dp_last else None)
    else:  # eager launches; the data-parallel micro-step interleaves NCCL buckets with backward
      loss_val = self.rt.loss_and_backward(inputs, targets, seg_start, grad_scale=1.0 / self.accumulation_steps,
                                           backward=True, on_bucket=on_bucket)
      if dp_last:
        self.reducer.join()
    self._record_loss(loss_val, self.accumulated_samples - 1)

    if last:
      self.accumulated_samples = 0
      # the gradient norm is always computed: it clips when grad_clip is set (max_norm > 0) and, clip or not, gates the
      # update kernels on the device (non-finite norm -> no update).  Data-parallel with the bf16 wire: the same pass
      # that writes the averaged gradients back to fp32 reduces the norm (plm_unpack_sumsq).
      norm_done = False
      if self.reducer is not None:
        norm_done = self.reducer.finish(self._sumsq_ws, self._gnorm_sq, joined=True)
      if not norm_done:
        grad_sumsq(self.rt.flat, self._sumsq_ws, self._gnorm_sq)
      clip = GradClip(self._gnorm_sq, self.grad_clip or 0.0)
      self.check_nan(wait=True)  # reference: raise before optimizer.step (engine.py:116-117 precedes :131)
      self.optimizer.step(grad_clip=clip)
      if self.scheduler:
        self.scheduler.step()
    return loss_val

  def grad_norm(self):
    """Total gradient L2 norm of the last optimizer step (device tensor; what clip_grad_norm_ returns)."""
    return self._gnorm_sq.sqrt()

  # ------------------------------------------------------------------------------------------ eval
  @torch.no_grad()
  def eval(self, dataloader):
    """reference: engine.py:143-177, with the loss-only fused forward (and without the reference's `.item()` crash on
    a float, engine.py:175)."""
    self.model.eval()
    total = torch.zeros(1, device=self.rt.flat.params.device, dtype=torch.float32)
    num_batches = 0
    for batch in dataloader:
      inputs, targets, seg = self._move_to_device(batch)
      loss = self.rt.loss_and_backward(inputs, targets, seg, backward=False)
      total += loss
      num_batches += 1
    count = torch.tensor([float(num_batches)], device=total.device)
    if dist.is_initialized():
      dist.all_reduce(total, op=dist.ReduceOp.SUM)
      dist.all_reduce(count, op=dist.ReduceOp.SUM)
    if torch.isnan(total).item():
      raise ValueError('Validation loss is nan')
    return (total / count).item()

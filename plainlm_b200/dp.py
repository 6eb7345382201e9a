"""Data-parallel gradient exchange: bucketed bf16 all-reduce over NCCL, overlapped with backward.

Replaces torch DDP as wrapped at engine/engine.py:64-65 of the reference (fp32 all-reduce of ~25 MiB buckets in
reverse registration order, divide by world size, sync only on the last micro-step, engine.py:104-105).  Here the
buckets are contiguous ranges of the model's flat fp32 gradient buffer, in the order backward finalises them
(lm_head, layer L-1 .. layer 0, embedding, norm weights).  When the runtime reports bucket i final, the comm stream
  1. waits for the compute stream's event,
  2. packs the range to bf16 pre-scaled by 1/world  (plm_cast_f32_bf16),
  3. all-reduces it (NCCL SUM over NVLink/NVSwitch, torch.distributed is only the plumbing),
while the compute stream keeps running the rest of backward.  `finish()` joins the streams and then makes ONE pass over
the whole wire buffer that writes the averaged fp32 gradients back AND reduces their squared norm for the clip
(plm_unpack_sumsq: 6 B/param at full HBM speed) — instead of one unpack kernel per bucket squeezed in beside the
backward GEMMs plus a separate norm pass.  With the fp32 wire (or injected pack/unpack functions: the CPU/gloo tests)
every bucket is unpacked right after its all-reduce, as DDP's bucket copy-back does.
The whole sequence is stream-ordered and free of host synchronisation, so the last micro-step can be captured — NCCL
calls included — in the runtime's CUDA graph (models/runtime.py).
"""

import torch
import torch.distributed as dist

from . import ops


def broadcast_parameters(flat, src=0, group=None):
  """Make every rank start from rank `src`'s weights (what the DDP constructor does, engine.py:65): ranks seed
  their RNG with seed+rank (torch_utils.py:35-37), so their initialisations differ."""
  dist.broadcast(flat.params, src=src, group=group)
  flat.refresh_shadow()


class GradReducer:
  def __init__(self, flat, group=None, wire_dtype=torch.bfloat16, pack=None, unpack=None):
    """pack(src_f32, dst_wire, scale) / unpack(src_wire, dst_f32, scale) default to the CUDA cast kernels; the CPU
    (gloo) unit tests inject torch casts to exercise the bucket logic without a GPU."""
    self.flat = flat
    self.group = group
    self.world = dist.get_world_size(group) if dist.is_initialized() else 1
    self.buckets = list(flat.buckets)
    self.wire = torch.empty(flat.total, device=flat.grads.device, dtype=wire_dtype)
    self.wire_dtype = wire_dtype
    self.on_cuda = flat.grads.is_cuda
    if wire_dtype == torch.float32:  # cfg key `ddp_fp32_allreduce: True`: the reference DDP's fp32 wire format
      default_pack = lambda s, d, sc: d.copy_(s).mul_(sc)  # noqa: E731
      default_unpack = lambda s, d, sc: d.copy_(s)  # noqa: E731
    else:  # BASELINE north_star: bf16 on the wire (pre-scaled by 1/world before rounding)
      default_pack = lambda s, d, sc: ops.cast_f32_bf16(s, d, sc)  # noqa: E731
      default_unpack = lambda s, d, sc: ops.cast_bf16_f32(s, d, sc)  # noqa: E731
    self.pack = pack or default_pack  # caller-injected functions (CPU/gloo tests) always win
    self.unpack = unpack or default_unpack
    # bf16 wire with the library's own kernels: no per-bucket unpack, one fused unpack + norm pass in finish()
    self.fused_tail = self.on_cuda and wire_dtype == torch.bfloat16 and pack is None and unpack is None
    self.comm_stream = torch.cuda.Stream(device=flat.grads.device) if self.on_cuda else None
    self._pending = []
    self.launched = 0

  def bucket_ready(self, i):
    """Called by the runtime right after the launches that finalise bucket i (last micro-step only)."""
    if self.world == 1:
      return
    a, b = self.buckets[i]
    g = self.flat.grads[a:b]
    w = self.wire[a:b]
    if self.on_cuda:
      ev = torch.cuda.Event()
      ev.record(torch.cuda.current_stream())
      with torch.cuda.stream(self.comm_stream):
        self.comm_stream.wait_event(ev)
        self.pack(g, w, 1.0 / self.world)
        dist.all_reduce(w, op=dist.ReduceOp.SUM, group=self.group)
        if not self.fused_tail:
          self.unpack(w, g, 1.0)
    else:
      self.pack(g, w, 1.0 / self.world)
      dist.all_reduce(w, op=dist.ReduceOp.SUM, group=self.group)
      self.unpack(w, g, 1.0)
    self.launched += 1

  def join(self):
    """Later work on the compute stream sees everything the comm stream has done (stream-ordered, capturable)."""
    if self.world == 1 or not self.on_cuda:
      return
    ev = torch.cuda.Event()
    ev.record(self.comm_stream)
    torch.cuda.current_stream().wait_event(ev)

  def finish(self, sumsq_workspace=None, gnorm_sq=None, joined=False):
    """Join, then (bf16 wire) write the averaged gradients back to the fp32 buffer.  With a workspace and an output
    scalar the same pass also leaves ||g||^2 in gnorm_sq[0]; returns True when it did (the caller then skips its own
    norm pass)."""
    if self.world == 1:
      return False
    if not joined:
      self.join()
    if not self.fused_tail:
      return False
    if sumsq_workspace is not None and gnorm_sq is not None:
      ops.unpack_sumsq(self.wire, self.flat.grads, sumsq_workspace, gnorm_sq)
      return True
    ops.cast_bf16_f32(self.wire, self.flat.grads, 1.0)
    return False

"""Hand-scheduled forward/backward of the whole Transformer on the CUDA kernels — the fused training path.

Replaces, for one micro-batch, everything between engine/engine.py:108 and :120 of the reference:
  model(inputs, attn_mask) -> CrossEntropyLoss -> loss / accum -> backward()
with a fixed sequence of kernel launches over pre-allocated activation buffers.  No autograd graph, no allocator
traffic, no dtype-cast kernels: bf16 weight shadows are persistent (refreshed by the optimizer kernel), gradients are
accumulated straight into the flat fp32 .grad buffer by the wgrad GEMM epilogue, and the dense attention mask of the
reference is replaced by int32 segment starts.

Layout in HBM (M = B*T tokens, d model dim, F GLU hidden, V vocab):
  params / grads  flat fp32 in data-parallel bucket order [lm_head | layer L-1 .. layer 0 | embed | all norm weights]
  shadows         flat bf16, same order (GEMM B operands)
  per layer saved x_in fp32[M,d], h1 bf16[M,d], rstd1[M], qkv bf16[M,3d], attn bf16[M,d], lse[B,H,T],
                  x_mid fp32[M,d], h2 bf16[M,d], rstd2[M], u bf16[M,2F], g bf16[M,F]
"""

import gc
import os
import weakref

import torch
from .. import _lib, ops

bf16, f32 = torch.bfloat16, torch.float32
_ALIGN = 64  # elements; keeps every parameter view 256-byte aligned
_RUNTIMES = weakref.WeakSet()


def release_all_graphs():
  """Destroy every captured CUDA graph of every live runtime.  A graph that captured NCCL collectives (the data-parallel
  last micro-step) keeps the communicator alive: destroy_process_group() blocks until such graphs are gone, so
  torch_utils.destroy_ddp() and bench.py call this first."""
  for rt in list(_RUNTIMES):
    rt.release_graphs()


def _bucket_param_names(model):
  """Flat order = order in which gradients become final during backward (DDP reverse-registration order analogue)."""
  L = model.n_layers
  tied = model.lm_head.weight is model.embed_tokens.weight
  buckets = [] if tied else [['lm_head.weight']]  # tied: the shared matrix is final only after the embedding backward
  for i in reversed(range(L)):
    pre = f'layers.{i}.'
    buckets.append([pre + 'mlp.fc2.weight', pre + 'mlp.fc1.weight', pre + 'attn.w_out.weight', pre + 'attn.w_qkv.weight'])
  buckets.append(['embed_tokens.weight'])
  norms = ['out_norm.weight']
  for i in reversed(range(L)):
    norms += [f'layers.{i}.mlp_norm.weight', f'layers.{i}.attn_norm.weight']
  buckets.append(norms)
  return buckets


class FlatParams:
  """Re-homes every parameter of the model as a view into one flat fp32 buffer (plus flat grads and bf16 shadows).

  state_dict names, shapes and dtypes are unchanged (checkpoint_utils.py of the reference keeps working); the flat
  buffers and shadows are not registered as buffers and never appear in a state_dict.
  """

  def __init__(self, model, device):
    named = dict(model.named_parameters(remove_duplicate=False))
    seen = {}
    self.entries = []  # (name, param, offset, numel)
    self.buckets = []  # (start, end) element ranges
    off = 0
    for names in _bucket_param_names(model):
      start = off
      for n in names:
        p = named[n]
        if id(p) in seen:  # tied weights: one storage, one gradient
          continue
        seen[id(p)] = n
        self.entries.append((n, p, off, p.numel()))
        off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
      if off > start:
        self.buckets.append((start, off))
    self.norm_start = self.buckets[-1][0]
    self.total = off
    missing = [n for n, p in named.items() if id(p) not in seen]
    if missing:
      raise RuntimeError(f'plainlm_b200: parameters without a flat slot: {missing}')
    self.params = torch.zeros(off, device=device, dtype=f32)
    self.grads = torch.zeros(off, device=device, dtype=f32)
    self.shadow = torch.zeros(off, device=device, dtype=bf16)
    for n, p, o, k in self.entries:
      view = self.params[o : o + k].view(p.shape)
      view.copy_(p.data.to(device=device, dtype=f32))
      p.data = view
      p.grad = self.grads[o : o + k].view(p.shape)
      p._plm_shadow = self.shadow[o : o + k].view(p.shape)
      p._plm_flat = (self, o, k)
    self.refresh_shadow()
    self._versions = self._version_sum()

  def _version_sum(self):
    return sum(p._version for _, p, _, _ in self.entries)

  def refresh_shadow(self):
    ops.cast_f32_bf16(self.params, self.shadow)

  def refresh_if_stale(self):
    """Parameters changed behind our back (load_state_dict, a foreign optimizer): re-cast the bf16 shadows.
    Our own optimizer kernels write the shadows themselves and do not bump autograd version counters."""
    v = self._version_sum()
    if v != self._versions:
      self.refresh_shadow()
      self._versions = v

  def restore_grad_views(self):
    for n, p, o, k in self.entries:
      if p.grad is None or p.grad.data_ptr() != self.grads.data_ptr() + 4 * o:
        p.grad = self.grads[o : o + k].view(p.shape)

  def zero_grads(self):
    self.grads.zero_()


class _Workspace:
  """Activation buffers of one (B, T) micro-batch shape.  train=True: everything backward needs is kept per layer.
  train=False (eval): forward-only — the per-layer lists alias a handful of buffers, there are no backward temporaries
  and no [M, V] logits buffer (the loss comes out of the LM-head GEMM's epilogue)."""

  def __init__(self, rt, B, T, train=True):
    m = rt.model
    dev = rt.flat.params.device
    M, d, F, V, H, L = B * T, m.dim, m.hidden_dim, m.vocab_size, m.n_heads, m.n_layers
    e = lambda *shape, dtype=bf16: torch.empty(*shape, device=dev, dtype=dtype)  # noqa: E731
    self.B, self.T, self.M, self.train = B, T, M, train
    U = rt.fc1_out  # 2F for the GLU (u = [a | z]), F for the single-branch MLPs

    def per_layer(n, *shape, dtype=bf16):
      if train:
        return [e(*shape, dtype=dtype) for _ in range(n)]
      one = e(*shape, dtype=dtype)
      return [one] * n

    if train:
      self.x = [e(M, d, dtype=f32) for _ in range(L + 1)]    # x[l] = input of layer l; x[L] = final stream
    else:
      two = [e(M, d, dtype=f32), e(M, d, dtype=f32)]           # layer l reads x[l], writes x[l + 1]: ping-pong
      self.x = [two[l & 1] for l in range(L + 1)]
    self.x_mid = per_layer(L, M, d, dtype=f32)
    self.h1 = per_layer(L, M, d)
    self.h2 = per_layer(L, M, d)
    self.rstd1 = per_layer(L, M, dtype=f32)
    self.rstd2 = per_layer(L, M, dtype=f32)
    self.qkv = per_layer(L, M, 3 * d)
    self.attn = per_layer(L, M, d)
    self.lse = per_layer(L, B, H, T, dtype=f32)
    self.u = per_layer(L, M, U)
    self.g = per_layer(L, M, F)
    self.hf = e(M, d)
    self.rstd_f = e(M, dtype=f32)
    # LM head + cross-entropy: per-(256-column tile, row) softmax statistics written by the GEMM epilogue, row results
    self.ce_partial = e(2 * ops.lmhead_ce_tiles(V) * M, dtype=f32)
    self.tgt_logit = e(M, dtype=f32)
    self.row_loss = e(M, dtype=f32)
    self.row_lse = e(M, dtype=f32)
    self.stats = torch.zeros(4, device=dev, dtype=f32)
    self.graphs = {}
    if not train:
      return
    # bf16 logits of the micro-batch: written once by the LM-head GEMM, turned into dlogits in place, read by the two
    # LM-head gradient GEMMs (the only [M, V] buffer; DESIGN.md §3 explains why recomputing it instead loses)
    self.logits = e(M, V)
    # backward temporaries
    self.dx = e(M, d, dtype=f32)
    self.dx_b2 = [e(M, d), e(M, d)]  # ping-pong: the side-stream wgrad of one branch reads it while the next is written
    self.dh = e(M, d)
    self.dg = e(M, F)
    self.du = e(M, U)
    self.dattn = e(M, d)
    self.dqkv = e(M, 3 * d)
    self.delta = e(B, H, T, dtype=f32)
    self.dq_acc = e(M, d, dtype=f32)
    self.nblk = ops.rmsnorm_bwd_blocks(M)
    self.dw_part = e(2 * L + 1, self.nblk, d, dtype=f32)  # per-norm partials, folded once per micro-step
    # static inputs + captured CUDA graphs of the micro-step (keyed by (masked, grad_scale))
    self.ids = torch.zeros(B, T, device=dev, dtype=torch.int64)
    self.targets = torch.zeros(B, T, device=dev, dtype=torch.int64)
    self.seg = torch.zeros(B * T, device=dev, dtype=torch.int32)


class TrainRuntime:
  """Owns the flat parameter storage and activation workspaces of one Transformer and runs its train step."""

  def __init__(self, model, device):
    self.model = model
    self.flat = FlatParams(model, device)
    self.device = device
    self._ws = {}
    self.rope = model.rope_table(device)
    p = dict(model.named_parameters(remove_duplicate=False))
    L = model.n_layers
    sh = lambda n: p[n]._plm_shadow  # noqa: E731
    gr = lambda n: p[n].grad  # noqa: E731
    self.names = p
    self.tied = model.lm_head.weight is model.embed_tokens.weight
    mlp0 = model.layers[0].mlp if L else None
    self.act_kind = getattr(mlp0, 'act_kind', None)  # None: GLU (fc1 -> [a | z]); else single-branch activation
    self.fc1_out = (2 if self.act_kind is None else 1) * model.hidden_dim
    ns = self.flat.norm_start
    self.norm_grads = self.flat.grads[ns : ns + (2 * L + 1) * model.dim]
    self.wstream = torch.cuda.Stream(device=device)  # weight-gradient GEMMs run here, filling the tails of the main stream
    self._readers = {}
    self.W = {n: sh(n) for n in p}   # bf16 GEMM operands
    self.G = {n: gr(n) for n in p}   # fp32 grad views
    self.P = {n: p[n].data for n in p}
    self.L = L
    # GLU: fuse d(silu(a) z) into fc2's input-gradient GEMM (PLM_FUSE_GLU_BWD=0: GEMM + plm_swiglu_bwd, for A/B runs)
    self.fuse_glu_bwd = os.environ.get('PLM_FUSE_GLU_BWD', '1') != '0' and model.hidden_dim % 256 == 0
    _RUNTIMES.add(self)

  def release_graphs(self):
    """Drop the captured micro-step graphs (they are re-captured on the next use)."""
    had = False
    for ws in self._ws.values():
      had = had or bool(ws.graphs)
      ws.graphs.clear()
    if had:
      torch.cuda.synchronize()
      gc.collect()

  def workspace(self, B, T, train=True):
    """At most two workspaces stay alive (normally the training shape + a forward-only eval one).  A training
    workspace also serves an eval batch of the same shape.  Eviction is least-recently-used and never takes a workspace
    that holds captured CUDA graphs while another candidate exists."""
    key = (B, T, True)
    if not train and key not in self._ws:
      key = (B, T, False)
    ws = self._ws.pop(key, None)
    if ws is None:
      if len(self._ws) >= 2:
        victims = [k for k, w in self._ws.items() if not w.graphs] or list(self._ws)
        self._ws.pop(victims[0])
      ws = _Workspace(self, B, T, train=key[2])
    self._ws[key] = ws  # re-insert: dict order = recency
    return ws

  # ------------------------------------------------------------------------------------------ forward
  def forward_hidden(self, ids, seg_start, ws):
    """Embedding + all blocks + final norm. Fills ws (saved activations); returns ws.hf (bf16 [M,d])."""
    m = self.model
    B, T, d, H, hd, L = ws.B, ws.T, m.dim, m.n_heads, m.head_dim, self.L
    W, P = self.W, self.P
    ops.embed_fwd(ids.reshape(-1), P['embed_tokens.weight'], ws.x[0])
    for l in range(L):
      pre = f'layers.{l}.'
      ops.rmsnorm_fwd(ws.x[l], P[pre + 'attn_norm.weight'], ws.h1[l], ws.rstd1[l], m.eps)
      ops.gemm(ws.h1[l], W[pre + 'attn.w_qkv.weight'], ws.qkv[l], epilogue=_lib.EPI_BF16_ROPE, rope_table=self.rope,
               rope_cols=2 * d, rope_T=T, head_dim=hd)
      ops.attn_fwd(ws.qkv[l], ws.attn[l], ws.lse[l], B, T, H, hd, seg_start=seg_start)
      ops.gemm(ws.attn[l], W[pre + 'attn.w_out.weight'], ws.x_mid[l], epilogue=_lib.EPI_RESID_F32, residual=ws.x[l])
      ops.rmsnorm_fwd(ws.x_mid[l], P[pre + 'mlp_norm.weight'], ws.h2[l], ws.rstd2[l], m.eps)
      if self.act_kind is None:
        # fc1 with the GLU gate applied while the tile is on chip: writes u = [a | z] (saved for backward), g = silu(a) z
        ops.gemm(ws.h2[l], W[pre + 'mlp.fc1.weight'], ws.u[l], epilogue=_lib.EPI_BF16_SWIGLU, out2=ws.g[l])
      else:  # MLP / MLPReluSquared (components.py:31-40, 59-70)
        ops.gemm(ws.h2[l], W[pre + 'mlp.fc1.weight'], ws.u[l])
        ops.act_fwd(ws.u[l], ws.g[l], self.act_kind)
      ops.gemm(ws.g[l], W[pre + 'mlp.fc2.weight'], ws.x[l + 1], epilogue=_lib.EPI_RESID_F32, residual=ws.x_mid[l])
    ops.rmsnorm_fwd(ws.x[L], P['out_norm.weight'], ws.hf, ws.rstd_f, m.eps)
    return ws.hf

  def forward_logits(self, ids, seg_start, ws):
    """Transformer.forward (models/transformer.py:108-114): plain LM-head GEMM, bf16 logits [M, V] (training workspace)."""
    self.forward_hidden(ids, seg_start, ws)
    ops.gemm(ws.hf, self.W['lm_head.weight'], ws.logits)
    return ws.logits

  def forward_loss(self, ids, targets, seg_start, ws, keep_logits):
    """Forward + mean cross-entropy with the LM head FUSED with the loss (transformer.py:114 + engine.py:110-112): the
    softmax statistics and the target logit are reduced in the GEMM epilogue while each logits tile is on chip.  With
    keep_logits the bf16 tile is also stored (backward turns it into dlogits in place); without, nothing of size
    [M, V] is written at all.  Leaves [sum, n_valid, mean] in ws.stats."""
    self.forward_hidden(ids, seg_start, ws)
    ops.lmhead_ce_fwd(ws.hf, self.W['lm_head.weight'], targets.reshape(-1), ws.logits if keep_logits else None,
                      ws.ce_partial, ws.tgt_logit, ws.row_loss, ws.row_lse, ws.stats, self.model.vocab_size)

  # ------------------------------------------------------------------------------------------ loss + backward
  def loss_and_backward(self, ids, targets, seg_start=None, grad_scale=1.0, backward=True, on_bucket=None):
    """One micro-batch: returns the mean CE loss (0-dim device tensor, un-scaled); when `backward`, accumulates
    d(loss * grad_scale)/dparams into the flat grad buffer.  `on_bucket(i)` is called right after the launches that
    finalise gradient bucket i (data-parallel overlap hook)."""
    self.flat.refresh_if_stale()
    B, T = ids.shape
    ws = self.workspace(B, T, train=backward)
    self.forward_loss(ids, targets, seg_start, ws, keep_logits=backward)
    loss = ws.stats[2].clone()
    if backward:
      ops.ce_grad(ws.logits, targets.reshape(-1), ws.row_lse, ws.stats, self.model.vocab_size, grad_scale=grad_scale)
      self.backward_from_dlogits(ids, seg_start, ws, on_bucket)
    return loss

  def graphed_loss_and_backward(self, ids, targets, seg_start=None, grad_scale=1.0, reducer=None):
    """Same work as loss_and_backward(backward=True), replayed from a CUDA graph: the ~590 launches of a micro-step are
    captured once per (shape, masked, grad_scale, data-parallel) and then cost one cudaGraphLaunch — no per-kernel
    host work, no launch gaps.  Inputs are copied into static device buffers (stream-ordered, so the previous replay
    has finished reading them).  With a `reducer` (the last micro-step of a data-parallel step) the bucket hooks — pack,
    NCCL all-reduce on the comm stream — are captured in the same graph and joined before it ends."""
    self.flat.refresh_if_stale()
    B, T = ids.shape
    ws = self.workspace(B, T)
    ws.ids.copy_(ids, non_blocking=True)
    ws.targets.copy_(targets, non_blocking=True)
    masked = seg_start is not None
    if masked:
      ws.seg.copy_(seg_start.reshape(-1), non_blocking=True)
    key = (masked, float(grad_scale), reducer is not None)
    rec = ws.graphs.get(key)
    seg = ws.seg if masked else None
    if rec is None:
      cur = torch.cuda.current_stream()
      side = torch.cuda.Stream()
      side.wait_stream(cur)
      with torch.cuda.stream(side):
        saved = self.flat.grads.clone()  # the warm-up / capture runs must not leak into the accumulated gradients
        self._micro_step(ws, seg, grad_scale, reducer)  # warm-up: lazy one-time initialisation happens outside the capture
        graph = torch.cuda.CUDAGraph()
        n0 = ops.LAUNCHES
        try:
          # thread_local: only THIS thread's illegal calls invalidate the capture.  The reference's DataLoader runs
          # with pin_memory=True (data/dataloaders.py), whose pin-memory thread calls cudaHostAlloc at any time; under
          # the default 'global' mode that would abort the capture (or make the pin thread fail).
          with torch.cuda.graph(graph, stream=side, capture_error_mode='thread_local'):
            self._micro_step(ws, seg, grad_scale, reducer)
          rec = (graph, ops.LAUNCHES - n0)
        except RuntimeError as e:  # capture refused: run this (shape, mask, scale) eagerly from now on
          print(f'plainlm_b200: CUDA graph capture failed ({str(e).splitlines()[0]}); falling back to eager launches')
          self._readers.clear()
          rec = 'eager'
        self.flat.grads.copy_(saved)
        del saved
      cur.wait_stream(side)
      ws.graphs[key] = rec
    if rec == 'eager':
      self._micro_step(ws, seg, grad_scale, reducer)
      return ws.stats[2].clone()
    graph, launches = rec
    graph.replay()
    ops.LAUNCHES += launches
    return ws.stats[2].clone()

  def _micro_step(self, ws, seg, grad_scale, reducer=None):
    self.forward_loss(ws.ids, ws.targets, seg, ws, keep_logits=True)
    ops.ce_grad(ws.logits, ws.targets.reshape(-1), ws.row_lse, ws.stats, self.model.vocab_size, grad_scale=grad_scale)
    self.backward_from_dlogits(ws.ids, seg, ws, reducer.bucket_ready if reducer is not None else None)
    if reducer is not None:
      reducer.join()  # the comm stream's work is ordered before whatever follows the micro-step (and closes the capture)

  def _wgrad(self, dy, x, name, guard):
    """grad[name] += dy^T x  (contraction over tokens, both operands read in place as MN-major), on the side stream:
    nothing downstream in backward needs it, so it overlaps the main stream and fills the idle SMs of its kernel tails.
    `guard` names the dy buffer: whoever overwrites it later first waits for this GEMM (_release)."""
    if os.environ.get('PLM_NO_SIDE_STREAM'):  # profiling aid: serialise so per-kernel events are not overlapped
      ops.gemm(dy, x, self.G[name], a_kmajor=False, b_kmajor=False, epilogue=_lib.EPI_ATOMIC_F32, splits=0)
      return
    main = torch.cuda.current_stream()
    ready = torch.cuda.Event()
    ready.record(main)
    self.wstream.wait_event(ready)
    with torch.cuda.stream(self.wstream):
      ops.gemm(dy, x, self.G[name], a_kmajor=False, b_kmajor=False, epilogue=_lib.EPI_ATOMIC_F32, splits=0)
      done = torch.cuda.Event()
      done.record(self.wstream)
    self._readers[guard] = done

  def _release(self, guard):
    """Make the current stream wait until the side-stream GEMM that reads buffer `guard` has finished."""
    ev = self._readers.pop(guard, None)
    if ev is not None:
      torch.cuda.current_stream().wait_event(ev)

  def _dgrad(self, dy, name, out):
    """out = dy W[name]  (W read in place as an MN-major B operand)."""
    ops.gemm(dy, self.W[name], out, a_kmajor=True, b_kmajor=False)

  def backward_from_dlogits(self, ids, seg_start, ws, on_bucket=None):
    m = self.model
    B, T, H, hd, L = ws.B, ws.T, m.n_heads, m.head_dim, self.L
    P, G = self.P, self.G
    bucket = 0
    flip = 0

    def done(on_side=False):
      nonlocal bucket
      if on_bucket is not None:
        if on_side and not os.environ.get('PLM_NO_SIDE_STREAM'):
          # the bucket is final when the side-stream wgrads are: let the reducer record its event there
          with torch.cuda.stream(self.wstream):
            on_bucket(bucket)
        else:
          on_bucket(bucket)
      bucket += 1

    def next_dxb():
      nonlocal flip
      flip ^= 1
      self._release(f'dx_b{flip}')
      return ws.dx_b2[flip]

    dlogits = ws.logits  # overwritten in place by plm_ce_grad
    self._dgrad(dlogits, 'lm_head.weight', ws.dh)
    self._wgrad(dlogits, ws.hf, 'lm_head.weight', 'logits')
    if not self.tied:
      done(on_side=True)
    dx_b = next_dxb()
    norm_i = 0  # norms are visited in the order their gradients sit in the flat buffer's last bucket
    ops.rmsnorm_bwd(ws.dh, ws.x[L], P['out_norm.weight'], ws.rstd_f, None, ws.dx, dx_b, ws.dw_part[norm_i])
    for l in reversed(range(L)):
      pre = f'layers.{l}.'
      # ---- MLP branch: x[l+1] = x_mid + fc2(silu(a) * z)
      if self.act_kind is None and self.fuse_glu_bwd:
        # fc2's input-gradient GEMM with the GLU backward in its epilogue: du = [da | dz] straight from the accumulator
        # tile and the saved u = [a | z]; dg is never materialised
        self._release('du')
        ops.gemm(dx_b, self.W[pre + 'mlp.fc2.weight'], ws.du, a_kmajor=True, b_kmajor=False,
                 epilogue=_lib.EPI_BF16_GLU_BWD, out2=ws.u[l])
        self._wgrad(dx_b, ws.g[l], pre + 'mlp.fc2.weight', f'dx_b{flip}')
      else:
        self._dgrad(dx_b, pre + 'mlp.fc2.weight', ws.dg)
        self._wgrad(dx_b, ws.g[l], pre + 'mlp.fc2.weight', f'dx_b{flip}')
        self._release('du')
        if self.act_kind is None:
          ops.swiglu_bwd(ws.dg, ws.u[l], ws.du)
        else:
          ops.act_bwd(ws.dg, ws.u[l], ws.du, self.act_kind)
      self._dgrad(ws.du, pre + 'mlp.fc1.weight', ws.dh)
      self._wgrad(ws.du, ws.h2[l], pre + 'mlp.fc1.weight', 'du')
      dx_b = next_dxb()
      norm_i += 1
      ops.rmsnorm_bwd(ws.dh, ws.x_mid[l], P[pre + 'mlp_norm.weight'], ws.rstd2[l], ws.dx, ws.dx, dx_b,
                      ws.dw_part[norm_i])
      # ---- attention branch: x_mid = x[l] + w_out(attn(rope(w_qkv(h1))))
      self._dgrad(dx_b, pre + 'attn.w_out.weight', ws.dattn)
      self._wgrad(dx_b, ws.attn[l], pre + 'attn.w_out.weight', f'dx_b{flip}')
      self._release('dqkv')
      ops.attn_bwd(ws.qkv[l], ws.attn[l], ws.dattn, ws.lse[l], ws.dqkv, ws.delta, ws.dq_acc, B, T, H, hd,
                   seg_start=seg_start, rope_table=self.rope)
      self._dgrad(ws.dqkv, pre + 'attn.w_qkv.weight', ws.dh)
      self._wgrad(ws.dqkv, ws.h1[l], pre + 'attn.w_qkv.weight', 'dqkv')
      done(on_side=True)
      dx_b = next_dxb()
      norm_i += 1
      ops.rmsnorm_bwd(ws.dh, ws.x[l], P[pre + 'attn_norm.weight'], ws.rstd1[l], ws.dx, ws.dx, dx_b, ws.dw_part[norm_i])
    # join: every side-stream GEMM is complete before anything after backward (next forward, reducer, optimizer)
    for guard in list(self._readers):
      self._release(guard)
    ops.embed_bwd(ids.reshape(-1), ws.dx, G['embed_tokens.weight'])
    done()  # embedding bucket
    # all 2L+1 norm-weight gradients in one launch: their slots are contiguous (stride d) in the last flat bucket
    ops.colsum_accum_batched(ws.dw_part, self.norm_grads, ws.nblk, m.dim, 2 * L + 1)
    done()  # norm-weight bucket

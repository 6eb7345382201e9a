"""CPU oracle for the plainLM training step.  TEST INFRASTRUCTURE ONLY.

This file restates, function by function, the arithmetic of the reference's hot path
(Niccolo-Ajroldi/plainLM: engine/engine.py:93-141 -> models/transformer.py -> optim) as explicit torch-CPU code, so
that the CUDA path in plainlm_b200/ can be checked against it on a machine where /root/reference does not exist.
It is imported ONLY by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; nothing
under plainlm_b200/ may import it (tests/test_host_logic.py::test_product_never_imports_the_oracle enforces that).

Parity status: PINNED.  tests/golden/make_golden.py imports the real reference from /root/reference in the build
container, runs it (TORCHDYNAMO_DISABLE=1, CPU, fp32 and bf16-autocast) and commits its outputs as fixtures under
tests/golden/; tests/test_oracle_golden.py checks every function below against those fixtures.  The reference itself
ships no tests or golden vectors, and its arithmetic lives in torch (pyproject.toml:16 `torch>=2.6.0`,
unpinned; fixtures generated with torch 2.11.0+cu128).

precision='fp32' is the reference's CPU path (engine.py:73-75 gives CPU a nullcontext).  precision='bf16' restates what
torch.autocast(bf16) does to this model on the GPU (probed dtype flow): nn.Linear inputs/weights/outputs
bf16, RMSNorm / RoPE / residual stream / loss in fp32, SDPA on bf16 tensors.
"""

import math

import torch
import torch.nn.functional as F

ROPE_THETA = 500000.0  # models/transformer.py:99 (hard-coded)
RMS_EPS = 1e-6  # models/construct.py:18


# ----------------------------------------------------------------------------------------------- config helpers
def glu_hidden(dim, expand=8 / 3, multiple_of=256):
  """models/transformer.py:75 + models/components.py:48: int(expand*dim) rounded up to a multiple of 256."""
  hidden = int(expand * dim)
  return multiple_of * ((hidden + multiple_of - 1) // multiple_of)


def param_names(n_layers):
  """state_dict key order of models/transformer.py Transformer (probed)."""
  names = ['embed_tokens.weight']
  for i in range(n_layers):
    names += [
      f'layers.{i}.attn.w_qkv.weight',
      f'layers.{i}.attn.w_out.weight',
      f'layers.{i}.attn_norm.weight',
      f'layers.{i}.mlp.fc1.weight',
      f'layers.{i}.mlp.fc2.weight',
      f'layers.{i}.mlp_norm.weight',
    ]
  names += ['out_norm.weight', 'lm_head.weight']
  return names


def init_params(vocab, dim, n_layers, n_heads, seed=0, expand=8 / 3, mlp_class='glu'):
  """Random parameters with the reference's init statistics (transformer.py:116-129): N(0, 0.02), residual output
  projections N(0, 0.02/sqrt(2L)), norm weights 1.  (Not the reference's RNG stream: parity tests share a state_dict.)"""
  g = torch.Generator().manual_seed(seed)
  Fh = glu_hidden(dim, expand)
  std_out = 0.02 / math.sqrt(2 * n_layers)
  p = {'embed_tokens.weight': torch.randn(vocab, dim, generator=g) * 0.02}
  for i in range(n_layers):
    p[f'layers.{i}.attn.w_qkv.weight'] = torch.randn(3 * dim, dim, generator=g) * 0.02
    p[f'layers.{i}.attn.w_out.weight'] = torch.randn(dim, dim, generator=g) * std_out
    p[f'layers.{i}.attn_norm.weight'] = torch.ones(dim)
    fc1_rows = 2 * Fh if mlp_class == 'glu' else Fh  # components.py:50 vs :35,66
    p[f'layers.{i}.mlp.fc1.weight'] = torch.randn(fc1_rows, dim, generator=g) * 0.02
    p[f'layers.{i}.mlp.fc2.weight'] = torch.randn(dim, Fh, generator=g) * std_out
    p[f'layers.{i}.mlp_norm.weight'] = torch.ones(dim)
  p['out_norm.weight'] = torch.ones(dim)
  p['lm_head.weight'] = torch.randn(vocab, dim, generator=g) * 0.02
  return p


# ----------------------------------------------------------------------------------------------- RoPE
def rope_table(head_dim, seq_len, theta=ROPE_THETA):
  """models/embeddings.py:8-12 precompute_freqs_cis, returned as [T, head_dim/2, 2] = (cos, sin), fp32."""
  inv_freqs = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
  t = torch.arange(seq_len, dtype=torch.float32)
  freqs = torch.outer(t, inv_freqs).float()
  return torch.stack([torch.cos(freqs), torch.sin(freqs)], dim=-1)


def apply_rope(x, table):
  """models/embeddings.py:15-30 on one tensor x [B, T, H, hd]: interleaved pairs (x[2i], x[2i+1]) rotated by t*theta_i
  in fp32, result cast back to x.dtype."""
  B, T, H, hd = x.shape
  xr = x.float().reshape(B, T, H, hd // 2, 2)
  cos = table[:T, :, 0][None, :, None, :]
  sin = table[:T, :, 1][None, :, None, :]
  o0 = xr[..., 0] * cos - xr[..., 1] * sin
  o1 = xr[..., 1] * cos + xr[..., 0] * sin
  return torch.stack([o0, o1], dim=-1).flatten(3).type_as(x)


# ----------------------------------------------------------------------------------------------- blocks
def rmsnorm(x, w, eps=RMS_EPS):
  """models/components.py:22-28: (x.float() * rsqrt(mean(x^2) + eps)).type_as(x) * w."""
  xf = x.float()
  return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).type_as(x) * w


def _linear(x, w, precision):
  """nn.Linear(bias=False). Under autocast both operands are cast to bf16 and the result is bf16."""
  if precision == 'bf16':
    return F.linear(x.to(torch.bfloat16), w.to(torch.bfloat16))
  return F.linear(x, w)


def doc_segment_starts(docs_lengths, seq_len):
  """Segment map equivalent to data/datasets/data_prep_utils.py:7-23 cropped as in engine/engine.py:23:
  lengths must sum to seq_len+1; returns int32 [seq_len] with the start position of each position's document."""
  if sum(int(n) for n in docs_lengths) != seq_len + 1:
    raise ValueError('Sum of doc_boundaries does not match max_seq_length.')  # data_prep_utils.py:10-11
  starts = []
  pos = 0
  for n in docs_lengths:
    n = int(n)
    starts += [pos] * n
    pos += n
  return torch.tensor(starts[:seq_len], dtype=torch.int32)


def mask_from_segment_starts(seg_start):
  """bool [T, T]: allowed(i, j) <=> seg_start[i] <= j <= i  (block-diagonal causal mask)."""
  T = seg_start.numel()
  i = torch.arange(T)
  return (i[None, :] <= i[:, None]) & (i[None, :] >= seg_start.long()[:, None])


def intra_doc_causal_mask(docs_lengths, max_seq_length):
  """Literal restatement of data/datasets/data_prep_utils.py:7-23 (tril blocks + block_diag), for pinning."""
  if sum(docs_lengths) != max_seq_length:
    raise ValueError('Sum of doc_boundaries does not match max_seq_length.')
  blocks = [torch.tril(torch.ones((n, n), dtype=torch.bool)) for n in docs_lengths]
  return torch.block_diag(*blocks)


def sdpa(q, k, v, mask=None):
  """F.scaled_dot_product_attention (transformer.py:61,63) written out: softmax(q k^T / sqrt(hd) + mask) v with fp32
  accumulation; q,k,v [B,H,T,hd]; mask bool [B,1,T,T] or None (= causal).  Output in q.dtype."""
  B, H, T, hd = q.shape
  s = (q.float() @ k.float().transpose(-1, -2)) / math.sqrt(hd)
  if mask is None:
    i = torch.arange(T)
    mask = (i[None, :] <= i[:, None])[None, None]
  s = s.masked_fill(~mask, float('-inf'))
  p = torch.softmax(s, dim=-1)
  return (p @ v.float()).to(q.dtype)


def attention(x, w_qkv, w_out, table, n_heads, mask=None, precision='fp32'):
  """models/transformer.py:39-67."""
  B, T, d = x.shape
  hd = d // n_heads
  q, k, v = _linear(x, w_qkv, precision).split(d, dim=2)
  q = apply_rope(q.view(B, T, n_heads, hd), table)
  k = apply_rope(k.view(B, T, n_heads, hd), table)
  v = v.view(B, T, n_heads, hd)
  out = sdpa(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), mask)
  out = out.transpose(1, 2).contiguous().view(B, T, d)
  return _linear(out, w_out, precision)


def glu(x, w1, w2, precision='fp32'):
  """models/components.py:53-56: a, z = fc1(x).split(hidden); fc2(silu(a) * z)."""
  hidden = w2.shape[1]
  a, z = _linear(x, w1, precision).split(hidden, dim=2)
  return _linear(F.silu(a) * z, w2, precision)


def mlp_plain(x, w1, w2, mlp_class, precision='fp32'):
  """models/components.py:38-40 (MLP: fc2(silu(fc1 x))) and :68-70 (MLPReluSquared: fc2(relu(fc1 x)^2))."""
  u = _linear(x, w1, precision)
  h = F.silu(u) if mlp_class == 'mlp' else F.relu(u).pow(2)
  return _linear(h, w2, precision)


def forward(params, ids, n_heads, mask=None, precision='fp32', mlp_class='glu'):
  """models/transformer.py:108-114 (+ Block.forward :79-83). ids int64 [B,T]; mask bool [B,T,T] or None.
  Returns logits [B,T,V] (bf16 under precision='bf16', as autocast produces)."""
  n_layers = sum(1 for k in params if k.endswith('attn_norm.weight'))
  x = F.embedding(ids, params['embed_tokens.weight'])  # fp32 residual stream
  d = x.shape[-1]
  table = rope_table(d // n_heads, ids.shape[1])
  m = None if mask is None else mask.unsqueeze(1)
  for i in range(n_layers):
    pre = f'layers.{i}.'
    h = rmsnorm(x, params[pre + 'attn_norm.weight'])
    x = x + attention(h, params[pre + 'attn.w_qkv.weight'], params[pre + 'attn.w_out.weight'], table, n_heads, m,
                      precision)
    h = rmsnorm(x, params[pre + 'mlp_norm.weight'])
    if mlp_class == 'glu':
      x = x + glu(h, params[pre + 'mlp.fc1.weight'], params[pre + 'mlp.fc2.weight'], precision)
    else:
      x = x + mlp_plain(h, params[pre + 'mlp.fc1.weight'], params[pre + 'mlp.fc2.weight'], mlp_class, precision)
  return _linear(rmsnorm(x, params['out_norm.weight']), params['lm_head.weight'], precision)


def loss_fn(logits, targets):
  """engine/engine.py:81,111: CrossEntropyLoss() (mean, ignore_index=-100) on fp32 logits (autocast upcasts)."""
  return F.cross_entropy(logits.float().view(-1, logits.size(-1)), targets.reshape(-1))


def split_batch(input_ids, seq_len):
  """engine/engine.py:16-17."""
  return input_ids[:, :seq_len], input_ids[:, 1 : seq_len + 1]


# ----------------------------------------------------------------------------------------------- optimizer path
def clip_grad_norm_(grads, max_norm):
  """torch.nn.utils.clip_grad_norm_ as called at engine/engine.py:128 (L2, error_if_nonfinite=False)."""
  total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
  coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
  for g in grads:
    g.mul_(coef)
  return total


def adamw_step(p, g, m, v, step, lr, beta1, beta2, eps, weight_decay):
  """torch.optim.AdamW single-tensor update (torch/optim/adam.py, decoupled weight decay), in place."""
  p.mul_(1 - lr * weight_decay)
  m.lerp_(g, 1 - beta1)
  v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
  bc1 = 1 - beta1**step
  bc2 = 1 - beta2**step
  denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
  p.addcdiv_(m, denom, value=-(lr / bc1))


def signsgd_step(p, g, m, first, lr, momentum, dampening, weight_decay):
  """optim/signSGD.py:29-46 (note :38-39: m starts as a clone of g, THEN is decayed and accumulated)."""
  p.mul_(1 - lr * weight_decay)
  if first:
    m.copy_(g)
  m.mul_(momentum).add_(g, alpha=1.0 - dampening)
  p.add_(torch.sign(m), alpha=-lr)


def nadamw_step(p, g, m, v, state, lr, beta1, beta2, eps, weight_decay, momentum_decay=4e-3):
  """torch.optim.NAdam(decoupled_weight_decay=True) single-tensor update (torch/optim/nadam.py _single_tensor_nadam),
  what optim/init_optim.py:23-32 builds for cfg.optim == 'nadamw'.  state: {'step': int, 'mu_product': float}."""
  state['step'] = step = state.get('step', 0) + 1
  bc2 = 1 - beta2**step
  p.mul_(1 - lr * weight_decay)
  mu = beta1 * (1.0 - 0.5 * (0.96 ** (step * momentum_decay)))
  mu_next = beta1 * (1.0 - 0.5 * (0.96 ** ((step + 1) * momentum_decay)))
  state['mu_product'] = prod = state.get('mu_product', 1.0) * mu
  m.lerp_(g, 1 - beta1)
  v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
  denom = v.div(bc2).sqrt().add_(eps)
  p.addcdiv_(g, denom, value=-lr * (1.0 - mu) / (1.0 - prod))
  p.addcdiv_(m, denom, value=-lr * mu_next / (1.0 - prod * mu_next))


def sgd_step(p, g, buf, first, lr, momentum, dampening, weight_decay):
  """torch.optim.SGD single-tensor update (torch/optim/sgd.py), as built at optim/init_optim.py:34-41: coupled L2
  weight decay, momentum buffer initialised to a clone of the first (decayed) gradient, no Nesterov."""
  g = g.add(p, alpha=weight_decay) if weight_decay != 0 else g
  if momentum != 0:
    if first:
      buf.copy_(g)
    else:
      buf.mul_(momentum).add_(g, alpha=1 - dampening)
    g = buf
  p.add_(g, alpha=-lr)


def warmup_cosine_lr(t, lr_start, lr_max, lr_end, warmup_steps, T):
  """optim/lr_schedule.py:40-48."""
  if t <= warmup_steps:
    return lr_start + (lr_max - lr_start) / warmup_steps * t
  if t <= T:
    progress = (t - warmup_steps) / (T - warmup_steps)
    return lr_end + 0.5 * (lr_max - lr_end) * (1 + math.cos(math.pi * progress))
  return lr_end


def no_decay(name):
  """models/construct.py:54-61: names containing 'bias' or 'norm' get weight_decay 0."""
  return 'bias' in name or 'norm' in name


def sampler_partition(n_rows, world, rank):
  """DistributedSampler(shuffle=False, drop_last=True) as built at data/dataloaders.py:91."""
  per = n_rows // world
  return list(range(rank, per * world, world))


class OracleTrainer:
  """engine/engine.py TorchEngine.step restated on CPU over a plain dict of fp32 tensors.

  cfg keys used (same names as config/*.yaml): seq_len, grad_accumulation_steps, grad_clip, optim ('adamw'|'signSGD'),
  lr, beta1, beta2, weight_decay, eps, dampening, scheduler (None|'warmup_cosine'), warmup_steps, lr_start, lr_end,
  steps_budget, n_heads, intra_doc_masking.
  """

  def __init__(self, params, cfg, precision='fp32'):
    self.p = {k: v.detach().clone().float().requires_grad_(True) for k, v in params.items()}
    self.cfg = dict(cfg)
    self.precision = precision
    self.accum = self.cfg.get('grad_accumulation_steps', 1)
    self.accumulated = 0
    self.opt_step = 0
    self.state = {k: {} for k in self.p}
    self.sched_iter = 0
    self.lr = self.cfg['lr']
    if self.cfg.get('scheduler') == 'warmup_cosine':
      self.lr = self.cfg['lr_start']  # lr_schedule.py:40 — optimizer step 1 runs at lr_start
    self._warmup = self.cfg.get('warmup_steps')
    if isinstance(self._warmup, float):
      self._warmup = int(self._warmup * self.cfg['steps_budget'])  # init_optim.py:80

  def loss(self, input_ids, docs_lengths=None):
    T = self.cfg['seq_len']
    inputs, targets = split_batch(input_ids, T)
    mask = None
    if self.cfg.get('intra_doc_masking', False):
      mask = torch.stack([mask_from_segment_starts(doc_segment_starts(dl, T)) for dl in docs_lengths], dim=0)
    logits = forward(self.p, inputs, self.cfg['n_heads'], mask, self.precision)
    return loss_fn(logits, targets)

  def step(self, batch):
    """One micro-step; returns the un-scaled loss like engine.py:115,141."""
    self.accumulated += 1
    loss = self.loss(batch['input_ids'], batch.get('docs_lengths'))
    (loss / self.accum).backward()
    loss_val = loss.detach()
    if torch.isnan(loss_val):
      raise ValueError('Train loss is nan')
    if self.accumulated == self.accum:
      self.accumulated = 0
      self._optimizer_step()
    return loss_val

  @torch.no_grad()
  def _optimizer_step(self):
    c = self.cfg
    names = list(self.p)
    grads = [self.p[k].grad for k in names]
    if c.get('grad_clip'):
      self.last_grad_norm = clip_grad_norm_(grads, c['grad_clip'])
    self.opt_step += 1
    for k in names:
      p, g, st = self.p[k], self.p[k].grad, self.state[k]
      wd = 0.0 if no_decay(k) else c['weight_decay']
      if c['optim'] == 'adamw':
        if not st:
          st['exp_avg'] = torch.zeros_like(p)
          st['exp_avg_sq'] = torch.zeros_like(p)
        adamw_step(p, g, st['exp_avg'], st['exp_avg_sq'], self.opt_step, self.lr, c['beta1'], c['beta2'],
                   c.get('eps', 1e-8), wd)
      elif c['optim'] == 'signSGD':
        first = 'm' not in st
        if first:
          st['m'] = torch.zeros_like(p)
        signsgd_step(p, g, st['m'], first, self.lr, c['beta1'], c['dampening'], wd)
      else:
        raise NotImplementedError(c['optim'])
      self.p[k].grad = None
    if c.get('scheduler') == 'warmup_cosine':
      self.sched_iter += 1
      self.lr = warmup_cosine_lr(self.sched_iter, c['lr_start'], c['lr'], c['lr_end'], self._warmup,
                                 c['steps_budget'])

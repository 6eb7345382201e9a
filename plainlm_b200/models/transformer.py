"""Transformer++ (RMSNorm, RoPE, GLU) with the reference's constructors, forward signature and state_dict names
(reference: models/transformer.py), executed by hand-written sm_100a kernels.

Two ways in:
  * `model(x, attn_mask) -> logits`          autograd-capable composition of plainlm_b200.models.functional ops;
                                             what the reference's own engine / eval loop calls.
  * `model.runtime().loss_and_backward(...)`  the fused train step used by plainlm_b200.engine.TorchEngine (no autograd
                                             graph, logits never leave the workspace).
"""

import math
from dataclasses import dataclass

import torch
from torch import nn

from . import functional as PF
from .components import RMSNorm, MLP, GLU, MLPReluSquared
from .embeddings import precompute_freqs_cis, rope_table_2d


@dataclass
class ModelConfig:
  vocab_size: int
  seq_len: int
  dim: int
  expand: float
  n_layers: int
  n_heads: int
  mlp: str = 'mlp'
  rmsnorm_eps: float = 1e-6
  tie_embeddings: bool = False


MLP_CLASSES = {'mlp': MLP, 'glu': GLU, 'mlp_relu_sq': MLPReluSquared}


class Attention(nn.Module):
  """reference: models/transformer.py:29-67."""

  def __init__(self, cfg: ModelConfig):
    super().__init__()
    assert cfg.dim % cfg.n_heads == 0
    self.n_heads = cfg.n_heads
    self.head_dim = cfg.dim // cfg.n_heads
    self.w_qkv = nn.Linear(cfg.dim, 3 * cfg.dim, bias=False)
    self.w_out = nn.Linear(cfg.dim, cfg.dim, bias=False)

  def forward(self, x, freqs_cis, attn_mask=None, residual=None):
    """x: bf16 (bsz, seqlen, d) normed input. attn_mask: None (causal), a dense bool (bsz, L, L) mask in the
    reference's format, or int32 segment starts (bsz*L,).  With `residual` (fp32) the output projection adds it in its
    epilogue and returns fp32."""
    bsz, seqlen, d = x.shape
    table = rope_table_2d(freqs_cis[:, :seqlen]).to(x.device) if freqs_cis.dim() == 5 else freqs_cis
    seg = None
    if attn_mask is not None:
      seg = attn_mask if attn_mask.dtype == torch.int32 else PF.seg_start_from_mask(attn_mask)
    qkv = PF.linear_rope(x, self.w_qkv.weight, table, 2 * d, seqlen, self.head_dim)
    out = PF.flash_attention(qkv, bsz, seqlen, self.n_heads, self.head_dim, seg)
    if residual is not None:
      return PF.linear_residual(out, self.w_out.weight, residual)
    return PF.linear(out, self.w_out.weight)


class Block(nn.Module):
  """reference: models/transformer.py:70-83."""

  def __init__(self, layer_id: int, cfg: ModelConfig):
    super().__init__()
    self.attn = Attention(cfg)
    self.attn_norm = RMSNorm(cfg.dim, cfg.rmsnorm_eps)
    self.mlp = MLP_CLASSES[cfg.mlp](dim=cfg.dim, hidden_dim=int(cfg.expand * cfg.dim))
    self.mlp_norm = RMSNorm(cfg.dim, cfg.rmsnorm_eps)
    self.layer_id = layer_id

  def forward(self, x, freqs_cis, attn_mask):
    # x: fp32 residual stream (bsz, seqlen, dim); both residual adds are fused into the producing GEMM's epilogue
    x = self.attn(self.attn_norm(x), freqs_cis, attn_mask, residual=x)
    u = PF.linear(self.mlp_norm(x), self.mlp.fc1.weight)
    kind = getattr(self.mlp, 'act_kind', None)  # None: GLU; else MLP (silu) / MLPReluSquared (relu^2)
    h = PF.swiglu(u) if kind is None else PF.activation(u, kind)
    return PF.linear_residual(h, self.mlp.fc2.weight, x)


class Transformer(nn.Module):
  """reference: models/transformer.py:86-140."""

  def __init__(self, cfg):
    super().__init__()
    self.n_layers = cfg.n_layers
    head_dim = cfg.dim // cfg.n_heads
    if cfg.dim % cfg.n_heads != 0:
      raise ValueError('dim must be divisible by n_heads')
    self.dim, self.n_heads, self.head_dim = cfg.dim, cfg.n_heads, head_dim
    self.vocab_size, self.seq_len, self.eps = cfg.vocab_size, cfg.seq_len, cfg.rmsnorm_eps

    # same registration order as the reference so that a given seed yields the same weights
    self.embed_tokens = nn.Embedding(cfg.vocab_size, cfg.dim)
    self.layers = nn.ModuleList([Block(idx, cfg) for idx in range(cfg.n_layers)])
    self.out_norm = RMSNorm(cfg.dim, cfg.rmsnorm_eps)
    self.lm_head = nn.Linear(cfg.dim, cfg.vocab_size, bias=False)
    self.hidden_dim = self.layers[0].mlp.hidden_dim if cfg.n_layers else 0

    self.freqs_cis = precompute_freqs_cis(head_dim, cfg.seq_len, 500000)[0 : cfg.seq_len]

    self.apply(self._init_weights)
    self._scale_residual_branches()
    if cfg.tie_embeddings:
      self.tie_weights()
    self._runtime = None
    self._rope_dev = {}

  # ---------------------------------------------------------------------------- reference API
  def forward(self, x, attn_mask):
    """x: int64 (bsz, seqlen) -> logits bf16 (bsz, seqlen, vocab).  attn_mask as in Attention.forward."""
    if not x.is_cuda:
      raise RuntimeError('plainlm_b200.Transformer runs on CUDA (sm_100a) only: there is no CPU path')
    self.runtime()  # parameters live in the flat buffers; keeps bf16 shadows fresh
    self._runtime.flat.refresh_if_stale()
    seg = None
    if attn_mask is not None:
      seg = attn_mask if attn_mask.dtype == torch.int32 else PF.seg_start_from_mask(attn_mask)
    table = self.rope_table(x.device)[: x.shape[1]]
    h = PF.embedding(x, self.embed_tokens.weight)
    for layer in self.layers:
      h = layer(h, table, seg)
    return PF.linear(self.out_norm(h), self.lm_head.weight)

  # Initialisation draws from the global torch RNG in exactly the reference's order (transformer.py:102-103,116-129):
  # first every Linear / Embedding weight in module-traversal order at std 0.02, then the two residual-branch output
  # projections of each block again at 0.02 / sqrt(2 L) — so a given seed yields the reference's weights bit for bit.
  BASE_STD = 0.02

  def _init_weights(self, module):
    if isinstance(module, (nn.Linear, nn.Embedding)):
      module.weight.data.normal_(mean=0.0, std=self.BASE_STD)
      bias = getattr(module, 'bias', None)
      if bias is not None:
        bias.data.zero_()

  def _scale_residual_branches(self):
    branch_std = self.BASE_STD / math.sqrt(2 * self.n_layers)
    for name, param in self.named_parameters():
      for suffix in ('fc2.weight', 'w_out.weight'):
        if name.endswith(suffix):
          param.data.normal_(mean=0.0, std=branch_std)

  def tie_weights(self):
    self.lm_head.weight = self.embed_tokens.weight

  def count_params(self, non_embedding=True):
    """Number of parameters; `non_embedding` leaves out the token embedding and an untied LM head (reference :134-140)."""
    skip = set()
    if non_embedding:
      skip = {id(self.embed_tokens.weight), id(self.lm_head.weight)}
    return sum(p.numel() for p in self.parameters() if id(p) not in skip)

  # ---------------------------------------------------------------------------- B200 runtime
  def rope_table(self, device):
    """[seq_len, head_dim/2, 2] fp32 (cos, sin) on `device`; non-persistent (not in state_dict, like freqs_cis)."""
    key = str(device)
    if key not in self._rope_dev:
      self._rope_dev[key] = rope_table_2d(self.freqs_cis).to(device)
    return self._rope_dev[key]

  def runtime(self):
    """Flatten parameters on their current CUDA device (once) and return the fused train-step runtime."""
    dev = self.embed_tokens.weight.device
    if dev.type != 'cuda':
      raise RuntimeError('plainlm_b200: move the model to a CUDA device first (no CPU path)')
    if self._runtime is None or self._runtime.device != dev:
      from .runtime import TrainRuntime

      self._runtime = TrainRuntime(self, dev)
    return self._runtime

  def _apply(self, fn, *args, **kwargs):
    # .to()/.cuda() may re-create parameter storage: if the flat views were replaced, rebuild the runtime lazily
    out = super()._apply(fn, *args, **kwargs)
    rt = self.__dict__.get('_runtime')
    if rt is not None:
      base = rt.flat.params.data_ptr()
      if not all(p.data_ptr() == base + 4 * o for _, p, o, _ in rt.flat.entries):
        self._runtime = None
    return out

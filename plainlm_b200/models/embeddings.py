"""RoPE table and stand-alone application (reference: models/embeddings.py)."""

import torch

from .. import ops


def precompute_freqs_cis(dim: int, end: int, theta: float = 10000.0, condense_ratio: int = 1):
  """Same table and layout as the reference (models/embeddings.py:8-12): fp32 [1, end, 1, dim/2, 2] = (cos, sin),
  computed on the CPU in fp32 so that it is bit-identical to the reference's."""
  inv_freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32, device=torch.device('cpu')) / dim))
  t = torch.arange(end, dtype=torch.float32, device=inv_freqs.device) / condense_ratio
  freqs = torch.outer(t, inv_freqs).float()
  return torch.stack([torch.cos(freqs)[None, :, None, :], torch.sin(freqs)[None, :, None, :]], dim=4)


def rope_table_2d(freqs_cis):
  """[1, T, 1, hd/2, 2] -> contiguous [T, hd/2, 2] (the layout the kernels index)."""
  return freqs_cis.reshape(freqs_cis.shape[1], freqs_cis.shape[3], 2).contiguous()


def apply_rotary_emb_complex_like(q, k, freqs_cis):
  """reference: models/embeddings.py:15-30. q, k: bf16 [B, T, H, hd] CUDA tensors; returns rotated copies.
  The training path never calls this (rotation is fused into the QKV GEMM epilogue); it exists for API parity."""
  B, T, H, hd = q.shape
  table = rope_table_2d(freqs_cis).to(q.device)
  v = torch.zeros_like(q)
  qkv = torch.cat([q.reshape(B * T, H * hd), k.reshape(B * T, H * hd), v.reshape(B * T, H * hd)], dim=1)
  qkv = qkv.to(torch.bfloat16).contiguous()
  ops.rope_qk_(qkv, table, T, H, hd)
  d = H * hd
  return qkv[:, :d].reshape(B, T, H, hd).to(q.dtype), qkv[:, d : 2 * d].reshape(B, T, H, hd).to(q.dtype)

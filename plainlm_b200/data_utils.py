"""Host-side integer bookkeeping of the train step: batch slicing, document segments, rank partition.

All of it is bit-exact by construction and checked against fixtures produced by the reference
(tests/golden/docmask.json, misc.json)."""

import numpy as np
import torch


def split_inputs_targets(input_ids, seq_len):
  """reference: engine/engine.py:16-17."""
  return input_ids[:, :seq_len], input_ids[:, 1 : seq_len + 1]


def seg_start_from_docs_lengths(docs_lengths, seq_len):
  """docs_lengths: per example, document lengths summing to seq_len+1 (the `docs_lengths` column written by the
  reference's concat_chunck, data/datasets/data_prep_utils.py:86-110).  Returns int32 [B, seq_len]:
  seg_start[b, t] = first position of t's document.  allowed(i, j) <=> seg_start[i] <= j <= i reproduces
  intra_doc_causal_mask(...)[:T, :T] (data_prep_utils.py:7-23, engine.py:19-23) without the O(T^2) mask."""
  out = np.empty((len(docs_lengths), seq_len), dtype=np.int32)
  for b, lengths in enumerate(docs_lengths):
    lengths = np.asarray([int(n) for n in lengths], dtype=np.int64)
    if lengths.sum() != seq_len + 1:
      raise ValueError('Sum of doc_boundaries does not match max_seq_length.')  # same check as data_prep_utils.py:10
    starts = np.concatenate([[0], np.cumsum(lengths)[:-1]])
    out[b] = np.repeat(starts, lengths)[:seq_len]
  return torch.from_numpy(out)


def rank_partition(n_rows, world, rank):
  """Rows seen by `rank`: DistributedSampler(shuffle=False, drop_last=True) as built at data/dataloaders.py:91."""
  per = n_rows // world
  return list(range(rank, per * world, world))


def pack_docs_lengths(docs_lengths, seq_len):
  """Flattens `docs_lengths` for the on-device segment-map kernel (plm_seg_start_from_lengths): returns
  (lengths int32 [n_docs_total], offsets int32 [B + 1]).  Keeps the reference's validation
  (data/datasets/data_prep_utils.py:10-11: every row's lengths must sum to seq_len + 1)."""
  flat, offsets = [], [0]
  for lengths in docs_lengths:
    row = [int(n) for n in lengths]
    if sum(row) != seq_len + 1:
      raise ValueError('Sum of doc_boundaries does not match max_seq_length.')
    flat.extend(row)
    offsets.append(len(flat))
  return torch.tensor(flat, dtype=torch.int32), torch.tensor(offsets, dtype=torch.int32)

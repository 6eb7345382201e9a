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


class PrefetchLoader:
  """Loader-side prefetch for the reference's DataLoader (data/dataloaders.py:11-67).

  `for batch in PrefetchLoader(trainloader, device): engine.step(batch)` behaves like iterating the loader itself
  (same batches, same order, same `len`), but a background thread stays `depth` batches ahead: it pulls the next batch
  from the loader (which unpickles / collates it), pins `input_ids` if the loader did not, and starts its host->device
  copy on a side stream.  What reaches `TorchEngine.step` is a device tensor whose copy the consumer's stream waits for
  through an event, so the step only enqueues work.  `docs_lengths` stays a host list (the engine ships the lengths and
  expands them on the device).  Exceptions of the worker (loader errors) are re-raised in the consuming thread.

  `to_device(ids) -> (device_tensor, ready)` is injectable for CPU tests; `ready` needs a `.wait()`.
  """

  _END = object()

  def __init__(self, loader, device, depth=2, to_device=None):
    self.loader = loader
    self.device = device
    self.depth = max(int(depth), 1)
    self._to_device = to_device or self._cuda_copy
    self._stream = None

  def __len__(self):
    return len(self.loader)

  def _cuda_copy(self, ids):
    if self._stream is None:
      self._stream = torch.cuda.Stream(device=self.device)
    if not ids.is_pinned():
      ids = ids.pin_memory()
    with torch.cuda.stream(self._stream):
      dev = ids.to(self.device, non_blocking=True)
      ready = torch.cuda.Event()
      ready.record(self._stream)
    return dev, ready

  def __iter__(self):
    import queue
    import threading

    q = queue.Queue(maxsize=self.depth)
    stop = threading.Event()

    def put(item):
      while not stop.is_set():
        try:
          q.put(item, timeout=0.1)
          return True
        except queue.Full:
          continue
      return False

    def worker():
      try:
        for batch in self.loader:
          out = dict(batch) if isinstance(batch, dict) else {'input_ids': batch}
          dev, ready = self._to_device(out['input_ids'])
          out['input_ids'] = dev
          if not put((out, ready)):
            return
        put(self._END)
      except BaseException as e:  # noqa: BLE001 — handed to the consumer
        put(e)

    t = threading.Thread(target=worker, name='plm-prefetch', daemon=True)
    t.start()
    try:
      while True:
        item = q.get()
        if item is self._END:
          break
        if isinstance(item, BaseException):
          raise item
        batch, ready = item
        ready.wait()  # CUDA: the consumer's current stream waits for the copy (no host sync)
        ids = batch['input_ids']
        if ids.is_cuda:
          ids.record_stream(torch.cuda.current_stream())  # allocated on the side stream, consumed on this one
        yield batch
    finally:
      stop.set()

"""Host -> device input pipeline of the fine-tune step.

The reference moves every batch with `t.cuda(device=device, non_blocking=True)` inside ForwardModelsTrain
(volta/task_utils.py:L308-311): the copies are asynchronous for the host but serialised with the compute on the
one CUDA stream.  At B200 step times the 80 MB of fp32 region features per 256-sample batch (3 ms of PCIe) would
then cost a fifth of the step.  `DevicePrefetcher` keeps the copy of batch i+1 in flight on a side stream while
step i computes; batches it yields are already on the device, so ForwardModelsTrain's own `.cuda()` is a no-op.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Tuple

import torch


class LabelPriorTable:
    """The label-prior distance rows of a batch from a table that lives on the device.

    The reference builds `semantic_distance[b, j] = semantic_dict[(j, t_b)]` (0 for j == t_b) with a Python double loop over
    B x 1842 dictionary look-ups per batch and ships the fp32 [B, 1842] matrix to the GPU with every batch
    (datasets/gqa_dataset_semantic_code_mix.py:L371-381, L433).  The table is 1842 x 1842 fp32 = 13.6 MB: keep it resident
    and gather the B rows by label (`clg_rows_gather`); the batch then only carries the B labels.
    """

    def __init__(self, table: torch.Tensor, device):
        """table[t, j] = distance of answer j from target answer t (diagonal 0), shape [A, A]"""
        if table.dim() != 2 or table.shape[0] != table.shape[1]:
            raise ValueError("expected a square [num_labels, num_labels] table")
        self.table = table.to(device=device, dtype=torch.float32).contiguous()

    @classmethod
    def from_semantic_dict(cls, semantic_dict, num_labels: int, device):
        """semantic_dict[(j, t)] as loaded by the reference's dataset (missing pairs and the diagonal are 0)"""
        t = torch.zeros(num_labels, num_labels, dtype=torch.float32)
        for (j, tt), v in semantic_dict.items():
            if j != tt and 0 <= j < num_labels and 0 <= tt < num_labels:
                t[tt, j] = float(v)
        return cls(t, device)

    def rows(self, labels: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """[B, A] distances for the B target labels (int64 / int32, on the table's device or on the host)"""
        from . import ops
        idx = labels.reshape(-1).to(device=self.table.device, dtype=torch.int32, non_blocking=True)
        if out is None:
            out = torch.empty(idx.numel(), self.table.shape[1], dtype=torch.float32, device=self.table.device)
        ops.rows_gather(self.table, idx, out)
        return out


class DevicePrefetcher:
    """Iterates over an iterable of tuples of pinned host tensors, yielding tuples of device tensors.

    Two static device slots (no allocator traffic in the loop): while the consumer computes on slot i % 2, the side
    stream fills slot (i + 1) % 2 - after an event that says the consumer's kernels on that slot (batch i - 1) have run.
    A yielded batch stays valid until the next-but-one `next()`.
    """

    def __init__(self, batches: Iterable[Tuple[torch.Tensor, ...]], device, reuse: "DevicePrefetcher" = None):
        """`reuse`: a finished prefetcher whose device slots and copy stream are adopted (one per epoch in a training
        loop: no allocation, no stream creation after the first epoch)."""
        self.it: Iterator = iter(batches)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("DevicePrefetcher moves batches to a CUDA device; got %s" % (device,))
        if reuse is not None and reuse.device == self.device:
            torch.cuda.current_stream(self.device).wait_stream(reuse.stream)
            self.stream, self.slots = reuse.stream, reuse.slots
        else:
            self.stream = torch.cuda.Stream(device=self.device)
            self.slots = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        # Both slots may still be read by kernels already enqueued on the consumer's stream (the last batches of the
        # adopted prefetcher; a freshly allocated slot is ordered on the allocating stream): the first copies into them
        # wait for everything the consumer has enqueued so far.
        seed = torch.cuda.Event()
        seed.record(torch.cuda.current_stream(self.device))
        self.free = [seed, seed]
        self.k = 0                 # index of the next batch to be yielded
        self.loaded = -1           # index of the last batch whose copy has been issued
        self.exhausted = False
        self._preload()

    def _preload(self):
        if self.exhausted:
            return
        try:
            host = next(self.it)
        except StopIteration:
            self.exhausted = True
            return
        j = (self.loaded + 1) % 2
        slot = self.slots[j]
        if slot is None or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(slot, host)):
            slot = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host)
            self.slots[j] = slot
            if self.slots[1 - j] is None:           # both slots at once: no allocation later in the loop
                self.slots[1 - j] = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host)
        with torch.cuda.stream(self.stream):
            if self.free[j] is not None:
                self.stream.wait_event(self.free[j])
            for d, h in zip(slot, host):
                d.copy_(h, non_blocking=True)
            self.ready[j].record(self.stream)
        self.loaded += 1

    def __iter__(self):
        return self

    def __next__(self):
        if self.k > self.loaded:
            raise StopIteration
        cur = torch.cuda.current_stream(self.device)
        if self.k >= 1:                             # everything enqueued so far has consumed batch k-1: its slot may be refilled
            ev = torch.cuda.Event()
            ev.record(cur)
            self.free[(self.k - 1) % 2] = ev
        j = self.k % 2
        cur.wait_event(self.ready[j])               # batch k has landed
        batch = self.slots[j]
        self.k += 1
        self._preload()                             # batch k+1 starts moving into the other slot
        return batch

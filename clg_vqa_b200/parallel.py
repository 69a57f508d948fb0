"""Data-parallel gradient exchange (one process per GPU, torch.distributed / NCCL over NVLink).

Replaces apex DistributedDataParallel(delay_allreduce=True) as the reference uses it
(train_task.py:L288-297; volta/apex/apex/parallel/distributed.py:L491-510, L425-475): there every
micro-step flattens ALL gradients into one fp32 tensor (1.13 GB for UC2), all-reduces it on the main
stream after backward, scales by 1/world and copies back.

Here
  * gradients already live in one flat arena (engine.g_flat) - no flatten / unflatten copies;
  * the dense part is all-reduced IN PLACE in buckets of four layers that are launched as soon as the last layer's
    backward has been enqueued, so NCCL overlaps the rest of the backward;
  * the 192 M-parameter word-embedding gradient is row-sparse (<= B*40 rows): ranks all-gather the
    per-token gradient rows + token ids (31 MB per rank at B=256) instead of all-reducing 768 MB, and every rank
    reduces the SAME gathered list in the SAME order (sorted by token id, one warp per id: `reduce_rows`), so the
    replicas' gradients, clip coefficients and parameters stay bit-identical, as they do under apex's all-reduce;
  * the 1/world average is folded into the optimiser kernel's gradient read (AdamW.grad_scale);
  * with gradient accumulation, `no_sync()` skips the exchange on all but the last micro-step.
The class only needs `torch.distributed` collectives, so its logic is exercised with gloo on CPU.
"""
from __future__ import annotations

import contextlib
import os
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class GradExchange:
    def __init__(self, world_size: Optional[int] = None, group=None):
        self.group = group
        self.world = world_size if world_size is not None else dist.get_world_size(group)
        self.pending = []            # async work handles of the current backward
        self.enabled = True
        self.done_ranges: List[Tuple[int, int]] = []
        self.row_ids: List[torch.Tensor] = []      # word-gradient rows of the micro-steps since the last exchange
        self.row_vals: List[torch.Tensor] = []
        self.bucket_layers = int(os.environ.get("CLG_DP_BUCKET_LAYERS", "4"))
        self.scale_after_exchange = 1.0   # 1/world when the optimiser cannot fold the average into its gradient read
        self.trace = None                 # list -> (name, CUDA event) marks of the exchange phases (bench.py --phases)
        self.plan = None                  # forward-time part of the row exchange (plan_rows)
        self.row_work = None              # (async all-gather handle, gathered rows, plan) launched by rows_ready
        self.plan_enabled = os.environ.get("CLG_DP_PLAN", "1") != "0"
        self.compact_rows = os.environ.get("CLG_DP_COMPACT", "1") != "0"
        self.compact_granule = 256        # rows sent per rank are rounded up to this (few distinct all-gather sizes)
        self._cnt_host = None

    def _mark(self, name: str):
        if self.trace is not None and torch.cuda.is_available():
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.trace.append((name, ev))

    # -- generic pieces (device-agnostic; tested with gloo) ---------------------------------------
    def reduce_range(self, flat: torch.Tensor, start: int, end: int):
        """launch an in-place sum-all-reduce of flat[start:end] (asynchronous)"""
        if not self.enabled or self.world == 1 or end <= start:
            return
        w = dist.all_reduce(flat[start:end], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.pending.append(w)
        self.done_ranges.append((start, end))

    def add_rows(self, ids: torch.Tensor, rows: torch.Tensor):
        """word-table gradient rows of one backward (kept until the next exchange; micro-steps inside no_sync() pile up)"""
        ids, rows = ids.reshape(-1), rows.reshape(ids.numel(), -1)
        if not self.enabled:                       # accumulation micro-step: the engine reuses these buffers before the exchange
            ids, rows = ids.clone(), rows.clone()
        self.row_ids.append(ids)
        self.row_vals.append(rows)

    def exchange_rows_deterministic(self, reduce_rows: Callable[[torch.Tensor, torch.Tensor], None],
                                    prereduce: Optional[Callable[[torch.Tensor, torch.Tensor], torch.Tensor]] = None):
        """all-gather every stashed (id, row) and hand the rank-ordered concatenation to reduce_rows(ids, rows), which must
        reduce it in an order that does not depend on the calling rank"""
        if not self.row_ids:
            return
        ids = torch.cat(self.row_ids) if len(self.row_ids) > 1 else self.row_ids[0]
        rows = torch.cat(self.row_vals) if len(self.row_vals) > 1 else self.row_vals[0]
        self.row_ids, self.row_vals = [], []
        if prereduce is not None and self.world > 1:
            if not rows.is_contiguous():
                rows = rows.contiguous()
            ids = prereduce(ids, rows)            # in place on rows; duplicates renamed to the padding id
            self._mark("rows_prereduced")
        if self.world > 1:
            all_ids = torch.empty(self.world * ids.numel(), dtype=ids.dtype, device=ids.device)
            all_rows = torch.empty((self.world * rows.shape[0], rows.shape[1]), dtype=rows.dtype, device=rows.device)
            dist.all_gather_into_tensor(all_ids, ids, group=self.group)
            dist.all_gather_into_tensor(all_rows.view(-1), rows.reshape(-1), group=self.group)
            ids, rows = all_ids, all_rows
            self._mark("rows_gathered")
        reduce_rows(ids, rows)
        self._mark("rows_reduced")

    # -- the id-only half of the row exchange, done at FORWARD time -------------------------------------------------
    def plan_rows(self, ids: torch.Tensor, pad_id: int):
        """Everything of the word-row exchange that depends on the token ids alone - the stable sort of this rank's ids,
        which copy of a duplicated id carries the summed row, the all-gather of the ids and the sort of the gathered list -
        is computed when the forward starts, so that the exposed tail after the backward is one row reduction, one
        all-gather of the rows and one more row reduction.  Identical results: the same lists in the same order."""
        self.plan = None
        if not (self.plan_enabled and self.enabled and self.world > 1) or self.row_ids:
            return
        flat = ids.reshape(-1)
        n = flat.numel()
        sorted_ids, order = torch.sort(flat, stable=True)
        first = torch.ones(n, dtype=torch.bool, device=flat.device)
        first[1:] = sorted_ids[1:] != sorted_ids[:-1]
        send = torch.empty_like(flat)           # the run's first member keeps the id, the others (and padding) are renamed
        send.scatter_(0, order, torch.where(first, sorted_ids, torch.full_like(sorted_ids, pad_id)))
        all_ids = torch.empty(self.world * n, dtype=flat.dtype, device=flat.device)
        dist.all_gather_into_tensor(all_ids, send, group=self.group)
        # Compaction: of the n = B x 40 rows a rank holds, only the carriers of distinct real tokens have to travel (text
        # padding and duplicates are renamed to the padding id: ~45 % of the rows on GQA-length questions).  comp[r] lists
        # rank r's positions with the carriers first, in position order (same relative order as the uncompacted list, so
        # the reduction below adds in exactly the same order); the row count every rank sends is the maximum over the ranks,
        # read by the host from pinned memory - the copy is issued here, at the start of the forward, and consumed after the
        # forward's own host read of the packed row count, so nobody waits for it.
        keep = all_ids.view(self.world, n) != pad_id
        comp = torch.sort((~keep).to(torch.uint8), dim=1, stable=True).indices
        if self._cnt_host is None:
            self._cnt_host = torch.zeros(1, dtype=torch.int64).pin_memory() if flat.is_cuda else torch.zeros(1, dtype=torch.int64)
        self._cnt_host.copy_(keep.sum(1).max().reshape(1), non_blocking=True)
        ev = None
        if flat.is_cuda:
            ev = torch.cuda.Event()
            ev.record()
        self.plan = dict(n=n, pad=int(pad_id), sorted=sorted_ids, order=order, scratch=torch.empty_like(flat), all_ids=all_ids,
                         comp=comp, ev=ev, K=None)

    def finalize_plan(self):
        """second half of plan_rows, once the number of rows to send is known on the host (still during the forward)"""
        plan = self.plan
        if plan is None or plan["K"] is not None:
            return
        if plan["ev"] is not None:
            plan["ev"].synchronize()
        n = plan["n"]
        g = self.compact_granule
        K = min(n, (int(self._cnt_host[0]) + g - 1) // g * g) if self.compact_rows else n
        K = max(K, 1)
        comp = plan["comp"][:, :K]
        ids_c = plan["all_ids"].view(self.world, n).gather(1, comp).reshape(-1)            # [world * K], padding id past a rank's count
        plan["all_sorted"], plan["all_order"] = torch.sort(ids_c, stable=True)
        plan["ids_c"] = ids_c
        plan["mine"] = comp[dist.get_rank(self.group)].to(torch.int32).contiguous()
        plan["K"] = K

    def rows_ready(self, engine):
        """backward: the word-gradient rows of this step exist (add_rows was just called).  With a forward-time plan the local
        pre-reduction, the compaction and the all-gather of the rows start NOW, overlapping the rest of the backward (the
        region half of the embedding backward); after_backward only waits for it."""
        plan = self.plan
        if plan is None or not self.enabled or self.world == 1 or len(self.row_ids) != 1 or self.row_ids[0].numel() != plan["n"]:
            self.plan = None
            return
        from . import ops
        self.finalize_plan()
        self.plan = None
        rows = self.row_vals[0]
        if not rows.is_contiguous():
            rows = rows.contiguous()
        self.row_ids, self.row_vals = [], []
        ops.segment_sum_rows(rows, plan["order"], plan["sorted"], None, plan["pad"], ids_out=plan["scratch"])
        K = plan["K"]
        # the gathered id list is in compacted order (carriers first) even when every row travels: the rows must follow it
        send = torch.empty((K, rows.shape[1]), dtype=rows.dtype, device=rows.device)
        ops.rows_gather(rows, plan["mine"], send, n=K)
        self._mark("rows_prereduced")
        all_rows = torch.empty((self.world * K, rows.shape[1]), dtype=rows.dtype, device=rows.device)
        h = dist.all_gather_into_tensor(all_rows.view(-1), send.reshape(-1), group=self.group, async_op=True)
        self.row_work = (h, all_rows, send, plan)

    def finish_rows(self, engine) -> bool:
        if self.row_work is None:
            return False
        from . import ops
        h, all_rows, _rows, plan = self.row_work
        self.row_work = None
        h.wait()
        self._mark("rows_gathered")
        ops.segment_sum_rows(all_rows, plan["all_order"], plan["all_sorted"], engine.grad_view(engine.word_table_param()), plan["pad"])
        engine.mark_word_rows(plan["ids_c"])
        self._mark("rows_reduced")
        return True

    def rows_mode(self) -> bool:
        """True: the engine must not scatter its word-table gradient itself but hand the rows to add_rows()"""
        return self.world > 1

    def finish(self):
        for w in self.pending:
            w.wait()
        self.pending = []
        self.done_ranges = []

    @contextlib.contextmanager
    def no_sync(self):
        """gradient-accumulation micro-steps that must not communicate"""
        old = self.enabled
        self.enabled = False
        try:
            yield
        finally:
            self.enabled = old

    # -- engine integration ---------------------------------------------------------------------
    def bind(self, engine):
        self.engine = engine
        engine.dp = self

    def _layer_range(self, engine, layer: int) -> Tuple[int, int]:
        """arena range holding the GEMM weights and biases of encoder layer `layer` (contiguous by
        construction: Engine._gemm_weight_modules order)"""
        first = 4 + 8 * layer           # slots: [W_img, b_img, W_loc, b_loc] then 8 per layer
        last = first + 7
        s0, s1 = engine.slots[first], engine.slots[last]
        from .engine import _pad
        return s0.start, s1.start + _pad(s1.numel)

    def layer_done(self, engine, layer: int):
        """backward of encoder layer `layer` is enqueued (layers finish in descending order): all-reduce the finished
        layers in buckets of CLG_DP_BUCKET_LAYERS (default 4: measured 13.5 vs 13.8 ms per step at N = 2 and 13.9 vs 14.0 at
        N = 4 against per-layer buckets - fewer NCCL kernels taking SMs from the persistent GEMM grids; 0 = one exchange
        after the backward, no overlap: 13.6 at N = 2)"""
        if not self.enabled or self.world == 1:
            return
        k = self.bucket_layers
        if k <= 0:
            return
        if layer % k == 0:
            a, _ = self._layer_range(engine, layer)
            _, b = self._layer_range(engine, min(layer + k, engine.L) - 1)
            self.reduce_range(engine.g_flat, a, b)

    def after_backward(self, engine):
        """everything that was not sent per layer, except the word-embedding table, which is exchanged
        as rows; then wait (stream-wise) for all of it"""
        if not self.enabled or self.world == 1:
            return
        self._mark("backward_done")
        from .engine import _pad
        word = engine.word_table_param()
        wslot = engine._slot_of_param[id(word)]
        skip = [(wslot.start, wslot.start + _pad(wslot.numel))] + list(self.done_ranges)
        skip.sort()
        cur = 0
        for a, b in skip:
            if a > cur:
                self.reduce_range(engine.g_flat, cur, a)
            cur = max(cur, b)
        if cur < engine.total:
            self.reduce_range(engine.g_flat, cur, engine.total)
        if not self.finish_rows(engine):
            self.exchange_rows_deterministic(engine.reduce_word_rows, engine.prereduce_word_rows)
        self.finish()
        self._mark("allreduce_done")
        if self.scale_after_exchange != 1.0:
            engine.g_flat.mul_(self.scale_after_exchange)


def attach(model, optimizer=None, world_size: Optional[int] = None, group=None) -> GradExchange:
    """Make `model` (BertForVLTasks / M3PForVLTasks) data-parallel: parameters are broadcast from rank 0
    once (apex DDP does the same at construction, distributed.py:L253), gradients are exchanged after
    every backward, and the optimiser averages them through its fused gradient scale."""
    ex = GradExchange(world_size, group)
    eng = model.engine
    dev = next(model.parameters()).device
    if eng.p_flat is None:
        eng.adopt(dev)
    if ex.world > 1:
        dist.broadcast(eng.p_flat, src=0, group=group)
        eng._shadow_versions = None
        # independent dropout masks per rank (the reference's processes have independent RNG streams)
        eng.set_rank(dist.get_rank(group))
    ex.bind(eng)
    if ex.world > 1 and dev.type == "cuda":
        # leave a few SMs to NCCL: its kernels overlap the backward GEMMs, whose persistent CTAs (one per SM, static tile
        # schedule) would otherwise queue behind them (CLG_DP_GEMM_SMS overrides; 0 = use every SM)
        import os
        from . import ops
        ops.set_gemm_sm_limit(int(os.environ.get("CLG_DP_GEMM_SMS", "0")))
    eng.grad_scale = 1.0 / ex.world                      # clip_grad_norm_ reports / clips the AVERAGED gradient's norm
    if optimizer is not None:
        if hasattr(optimizer, "grad_scale"):
            optimizer.grad_scale = 1.0 / ex.world        # our AdamW: fused kernel and per-parameter path both honour it
        else:
            # foreign optimiser: the arena is averaged right after the exchange, so what clip_grad_norm_ sees IS the average
            ex.scale_after_exchange = 1.0 / ex.world
            eng.grad_scale = 1.0
    return ex

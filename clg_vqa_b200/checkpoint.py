"""Checkpoint I/O off the critical path.

Mirrors `volta/train_utils.py`: `save` (L351-372), `save_sft` (L375-416: masked weights folded back as
weight_orig * weight_mask plus the `unmasked` twin) and `resume` (L488-510) - same signatures, same file names
(`pytorch_model_<epoch>.bin`, `pytorch_model_best.bin`, `pytorch_model_unmasked<epoch>.bin`,
`pytorch_model_unmasked_best.bin`, `pytorch_ckpt_latest.tar`) and the same dictionary layout, so a checkpoint written
here is loaded by the reference's `resume` and vice versa.

The reference serialises 1.1 GB of parameters and 2.3 GB of AdamW moments with `torch.save` straight from device
tensors inside the training loop: the step stalls for the whole device-to-host copy and the file write.  Here `save`
returns after enqueueing work:

  1. on the compute stream, one device-to-device copy of the parameter arena and of the two moment arenas into
     snapshot buffers (3.4 GB at HBM speed, about 1.5 ms) - the snapshot is consistent whatever the next step does;
  2. on a side stream, the snapshot goes to pinned host memory while training continues;
  3. a writer thread waits for the copy event, rebuilds the state dicts as views of the host buffers and writes the
     files (to a temporary name, then an atomic rename).

Models or optimisers that do not live in an engine arena (CPU tests, foreign optimisers) take the same route with
per-tensor host copies.  `wait()` blocks until every pending checkpoint is on disk; a second `save` while one is still
being written waits for it first (one set of snapshot buffers).
"""
from __future__ import annotations

import os
import threading
from collections import OrderedDict
from typing import Any, Callable, Dict, List, Optional

import torch


class _Arena:
    """A flat device tensor snapshotted as a whole; tensors that are views into it are re-created on the host copy."""

    def __init__(self, flat: torch.Tensor):
        self.flat = flat
        self.lo = flat.data_ptr()
        self.hi = self.lo + flat.numel() * flat.element_size()
        self.snap: Optional[torch.Tensor] = None
        self.host: Optional[torch.Tensor] = None

    def holds(self, t: torch.Tensor) -> bool:
        return (t.is_cuda and t.dtype == self.flat.dtype and t.is_contiguous() and t.numel() > 0 and
                self.lo <= t.data_ptr() and t.data_ptr() + t.numel() * t.element_size() <= self.hi)

    def host_view(self, t: torch.Tensor) -> torch.Tensor:
        off = (t.data_ptr() - self.lo) // self.flat.element_size()
        return self.host[off:off + t.numel()].view(t.shape)


class AsyncCheckpointer:
    """Snapshot now, write later.  One instance per training process; `save` / `save_sft` below use a module-level one."""

    def __init__(self):
        self._thread: Optional[threading.Thread] = None
        self._error: Optional[BaseException] = None
        self._arenas: Dict[int, _Arena] = {}
        self._loose_host: Dict[int, torch.Tensor] = {}
        self._stream = None

    # -- snapshot ------------------------------------------------------------------------------------
    def _arena_for(self, flat: torch.Tensor) -> _Arena:
        a = self._arenas.get(flat.data_ptr())
        if a is None or a.flat.numel() != flat.numel():
            a = _Arena(flat)
            a.snap = torch.empty_like(flat)
            a.host = torch.empty(flat.shape, dtype=flat.dtype, device="cpu", pin_memory=True)
            self._arenas[flat.data_ptr()] = a
        return a

    def _snapshot(self, trees: List[Any], flats: List[torch.Tensor]) -> Callable[[], List[Any]]:
        """Enqueue the copies of every tensor reachable from `trees`; returns a function that (after the copies have
        completed) gives the same trees with host tensors."""
        arenas = [self._arena_for(f) for f in flats if f is not None and f.is_cuda]
        loose: Dict[int, torch.Tensor] = {}          # id(device tensor) -> host copy
        cuda_dev = None

        def visit(x):
            nonlocal cuda_dev
            if torch.is_tensor(x):
                if x.is_cuda:
                    cuda_dev = x.device
                    if not any(a.holds(x) for a in arenas) and id(x) not in loose:
                        loose[id(x)] = x
            elif isinstance(x, dict):
                for v in x.values():
                    visit(v)
            elif isinstance(x, (list, tuple)):
                for v in x:
                    visit(v)

        for t in trees:
            visit(t)
        done = None
        if cuda_dev is not None:
            cur = torch.cuda.current_stream(cuda_dev)
            for a in arenas:                                   # consistent device-side snapshot, on the compute stream
                a.snap.copy_(a.flat, non_blocking=True)
            staged = {k: v.detach().clone() for k, v in loose.items()}
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=cuda_dev)
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                for a in arenas:
                    a.host.copy_(a.snap, non_blocking=True)
                hosts = {}
                for i, (k, v) in enumerate(staged.items()):
                    h = self._loose_host.get(i)                # pinned buffers are kept: allocating them stalls the host
                    if h is None or h.shape != v.shape or h.dtype != v.dtype:
                        h = torch.empty(v.shape, dtype=v.dtype, device="cpu", pin_memory=True)
                        self._loose_host[i] = h
                    h.copy_(v, non_blocking=True)
                    v.record_stream(self._stream)
                    hosts[k] = h
                done = torch.cuda.Event()
                done.record(self._stream)
        else:
            hosts = {}

        def rebuild(x):
            if torch.is_tensor(x):
                if x.is_cuda:
                    for a in arenas:
                        if a.holds(x):
                            return a.host_view(x)
                    return hosts[id(x)]
                return x                                       # host leaves were copied when the snapshot was taken
            if isinstance(x, OrderedDict):
                return OrderedDict((k, rebuild(v)) for k, v in x.items())
            if isinstance(x, dict):
                return {k: rebuild(v) for k, v in x.items()}
            if isinstance(x, list):
                return [rebuild(v) for v in x]
            if isinstance(x, tuple):
                return tuple(rebuild(v) for v in x)
            return x

        # host-resident leaves are copied now (cheap), device leaves are resolved after the event
        host_now = [_clone_host_leaves(t) for t in trees]

        def finish():
            if done is not None:
                done.synchronize()
            return [rebuild(t) for t in host_now]
        return finish

    def prepare(self, model, optimizer=None):
        """Allocate the snapshot and pinned host buffers now (a few seconds for 3.4 GB) instead of inside the first
        `save` of the run.  The optimiser's moment arenas exist after its first step; call again then, or leave
        them to the first save."""
        m = model.module if hasattr(model, "module") else model
        for f in _flats(m, optimizer):
            if f is not None and f.is_cuda:
                self._arena_for(f)

    # -- writer --------------------------------------------------------------------------------------
    def submit(self, trees: List[Any], flats: List[torch.Tensor], write: Callable[[List[Any]], None]):
        self.wait()                                   # one set of snapshot buffers
        finish = self._snapshot(trees, flats)

        def run():
            try:
                write(finish())
            except BaseException as e:                # surfaced by the next wait()
                self._error = e
        self._thread = threading.Thread(target=run, name="clg-checkpoint-writer", daemon=False)
        self._thread.start()

    def wait(self):
        t = self._thread
        if t is not None:
            t.join()
            self._thread = None
        if self._error is not None:
            e, self._error = self._error, None
            raise e


def _clone_host_leaves(x):
    """Host tensors are cloned immediately; device tensors stay (by identity) for the deferred rebuild."""
    if torch.is_tensor(x):
        return x if x.is_cuda else x.detach().clone()
    if isinstance(x, OrderedDict):
        return OrderedDict((k, _clone_host_leaves(v)) for k, v in x.items())
    if isinstance(x, dict):
        return {k: _clone_host_leaves(v) for k, v in x.items()}
    if isinstance(x, list):
        return [_clone_host_leaves(v) for v in x]
    if isinstance(x, tuple):
        return tuple(_clone_host_leaves(v) for v in x)
    return x


_DEFAULT = AsyncCheckpointer()


def prepare(model, optimizer=None):
    """Pre-allocate the default checkpointer's buffers (see AsyncCheckpointer.prepare)."""
    _DEFAULT.prepare(model, optimizer)


def wait():
    """Block until every checkpoint handed to `save` / `save_sft` is on disk (call before exiting or evaluating a file)."""
    _DEFAULT.wait()


def _atomic_save(obj, path):
    tmp = path + ".tmp"
    torch.save(obj, tmp)
    os.replace(tmp, path)


def _flats(model, optimizer):
    flats = []
    m = model.module if hasattr(model, "module") else model
    eng = m.__dict__.get("_engine", None)          # encoders.py keeps the (lazily built) engine there
    if eng is not None and getattr(eng, "p_flat", None) is not None:
        flats.append(eng.p_flat)
    fs = getattr(optimizer, "_fused", None)
    if fs:
        flats += [fs["m"], fs["v"]]
    return flats


def fold_state_dict(state_dict):
    """train_utils.py:L383-395: (masked, unmasked) plain state dicts from one with `*_orig` / `*_mask` entries."""
    masked, unmasked = {}, {}
    for k in state_dict.keys():
        if "_orig" in k:
            masked[k.replace("_orig", "")] = state_dict[k] * state_dict[k.replace("_orig", "_mask")]
            unmasked[k.replace("_orig", "")] = state_dict[k]
        elif "_mask" in k:
            continue
        else:
            masked[k] = state_dict[k]
            unmasked[k] = state_dict[k]
    return masked, unmasked


def _save_impl(path, logger, epoch_id, model, optimizer, scheduler, global_step, tb_logger, default_gpu, score, is_best, sft,
               checkpointer):
    if not default_gpu:
        return
    if logger is not None:
        logger.info("** ** * Saving model * ** ** ")
    ck = checkpointer or _DEFAULT
    model_to_save = model.module if hasattr(model, "module") else model
    trees = [model_to_save.state_dict(), optimizer.state_dict()]
    meta = {"scheduler_state_dict": scheduler.state_dict() if scheduler is not None else None, "global_step": global_step,
            "epoch_id": epoch_id, "tb_logger": tb_logger, "score": score}
    eng = model_to_save.__dict__.get("_engine")
    if eng is not None:       # extra key (ignored by the reference's resume): the dropout stream continues after a resume
        meta["clg_engine_rng"] = eng.rng_state()

    def write(host_trees):
        sd, osd = host_trees
        if sft:
            sd, unmasked = fold_state_dict(sd)
            _atomic_save(unmasked, os.path.join(path, "pytorch_model_unmasked" + str(epoch_id) + ".bin"))
            if is_best:
                _atomic_save(unmasked, os.path.join(path, "pytorch_model_unmasked_best.bin"))
        _atomic_save(sd, os.path.join(path, "pytorch_model_" + str(epoch_id) + ".bin"))
        if is_best:
            _atomic_save(sd, os.path.join(path, "pytorch_model_best.bin"))
        _atomic_save({"model_state_dict": sd, "optimizer_state_dict": osd, **meta}, os.path.join(path, "pytorch_ckpt_latest.tar"))

    ck.submit(trees, _flats(model_to_save, optimizer), write)


def save(path, logger, epoch_id, model, optimizer, scheduler, global_step, tb_logger, default_gpu, score, is_best=False,
         checkpointer: Optional[AsyncCheckpointer] = None):
    """train_utils.save (L351-372), asynchronous: returns once the snapshot copies are enqueued."""
    _save_impl(path, logger, epoch_id, model, optimizer, scheduler, global_step, tb_logger, default_gpu, score, is_best, False,
               checkpointer)


def save_sft(path, logger, epoch_id, model, optimizer, scheduler, global_step, tb_logger, default_gpu, score, is_best=False,
             checkpointer: Optional[AsyncCheckpointer] = None):
    """train_utils.save_sft (L375-416), asynchronous; the mask fold-back runs on the host copy in the writer thread."""
    _save_impl(path, logger, epoch_id, model, optimizer, scheduler, global_step, tb_logger, default_gpu, score, is_best, True,
               checkpointer)


def resume(path, model, optimizer, scheduler, tb_logger):
    """train_utils.resume (L488-510)."""
    wait()
    start_iter_id, global_step, start_epoch, best_score = 0, 0, 0, float("-inf")
    if path != "" and os.path.exists(path):
        checkpoint = torch.load(path, map_location="cpu", weights_only=False)
        new_dict = {}
        for attr in checkpoint["model_state_dict"]:
            if attr.startswith("module."):
                new_dict[attr.replace("module.", "", 1)] = checkpoint["model_state_dict"][attr]
            else:
                new_dict[attr] = checkpoint["model_state_dict"][attr]
        model.load_state_dict(new_dict)
        if scheduler is not None and checkpoint.get("scheduler_state_dict") is not None:
            scheduler.load_state_dict(checkpoint.get("scheduler_state_dict"))
        optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
        global_step = checkpoint["global_step"]
        start_epoch = int(checkpoint["epoch_id"]) + 1
        tb_logger = checkpoint["tb_logger"]
        best_score = checkpoint.get("score", float("-inf"))
        rng = checkpoint.get("clg_engine_rng")
        target = model.module if hasattr(model, "module") else model
        if rng is not None and hasattr(target, "engine"):
            target.engine.set_rng_state(rng)
        del checkpoint
    return start_iter_id, global_step, start_epoch, tb_logger, best_score

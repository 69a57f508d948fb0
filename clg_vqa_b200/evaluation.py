"""xGQA evaluation over language shards.

The reference evaluates one language per `eval_task.py` invocation (eval_task.py:L178-210: a loop over the split's
batches through `EvaluatingModel`, then `<split>_result.json` / `<split>_others.json`); the eight xGQA languages are
eight runs of the same checkpoint.  The languages are independent, so with N GPUs they are sharded across the ranks -
no collective on the data path, one object gather of the per-language scores at the end:

  * `shard_languages` assigns whole languages to ranks (longest-processing-time first on the shard sizes, so 8
    languages on 8, 4 or 2 GPUs finish together);
  * `evaluate_split` is the reference's loop for one language: device argmax and ONE device-to-host copy per batch
    (`task_utils.EvaluatingModel`), host batches double-buffered to the device (`data.DevicePrefetcher`);
  * `evaluate_languages` runs this rank's languages, writes the reference's two JSON files per language and returns the
    accuracy table gathered from every rank (`torch.distributed.all_gather_object` when a process group exists).
"""
from __future__ import annotations

import json
import os
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from . import task_utils as TU


def shard_languages(languages: Sequence[str], rank: int, world_size: int, sizes: Optional[Dict[str, int]] = None) -> List[str]:
    """Languages owned by `rank`.  Deterministic on every rank: greedy longest-first onto the least loaded rank (ties by
    rank index); without sizes every language weighs the same, which gives a round-robin."""
    if world_size <= 1:
        return list(languages)
    order = sorted(range(len(languages)), key=lambda i: (-(sizes or {}).get(languages[i], 1), i))
    load = [0] * world_size
    owner = {}
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += (sizes or {}).get(languages[i], 1)
    return [languages[i] for i in range(len(languages)) if owner[i] == rank]


def evaluate_split(config, task_cfg, device, task_id, dataloader, model, criterion=None, prefetch: bool = True,
                   step_fn: Optional[Callable] = None) -> Tuple[List[dict], List[dict], Optional[float]]:
    """eval_task.py:L178-189 for one split.  Returns (results, others, accuracy); accuracy is computed on the device from
    the batch targets when the batches carry them (synthetic / validation splits), else None."""
    results: List[dict] = []
    others: List[dict] = []
    step = step_fn or TU.EvaluatingModel
    batches: Iterable = dataloader
    if prefetch and step_fn is None and torch.cuda.is_available() and str(device) != "cpu":
        from .data import DevicePrefetcher
        batches = DevicePrefetcher(dataloader, device)
    correct = None
    seen = 0
    for batch in batches:
        n_before = len(results)
        _, _, batch_size, results, others = step(config, task_cfg, device, task_id, batch, model, dataloader, criterion, results,
                                                  others)
        seen += batch_size
        target = batch[4] if len(batch) > 4 and torch.is_tensor(batch[4]) and batch[4].dim() == 2 else None
        l2a = getattr(getattr(dataloader, "dataset", None), "label2ans", None)
        if target is not None and step_fn is None and l2a is None:
            # accuracy accumulates on the device (no further host sync per batch; one read at the end of the split)
            target = target.to(device, non_blocking=True)
            pred = torch.tensor([r["prediction"] for r in results[n_before:]], dtype=torch.int64).to(target.device, non_blocking=True)
            hit = ((pred == target.argmax(1)) & (target.sum(1) > 0)).sum()
            correct = hit if correct is None else correct + hit
    acc = None if correct is None or seen == 0 else int(correct) / seen
    return results, others, acc


def evaluate_languages(config, task_cfg, device, task_id, loaders: Dict[str, object], model, save_path: Optional[str] = None,
                       rank: Optional[int] = None, world_size: Optional[int] = None, criterion=None,
                       step_fn: Optional[Callable] = None) -> Dict[str, dict]:
    """Runs this rank's share of `loaders` ({language: dataloader}) and returns {language: {"n": ..., "accuracy": ...,
    "rank": ...}} for ALL languages (gathered).  Files: `<save_path>/<language>_result.json`, `..._others.json`."""
    dist_on = torch.distributed.is_available() and torch.distributed.is_initialized()
    if rank is None:
        rank = torch.distributed.get_rank() if dist_on else 0
    if world_size is None:
        world_size = torch.distributed.get_world_size() if dist_on else 1
    languages = list(loaders.keys())
    sizes = {}
    for lang, dl in loaders.items():
        try:
            sizes[lang] = len(dl)
        except TypeError:
            sizes[lang] = 1
    mine = shard_languages(languages, rank, world_size, sizes)
    table: Dict[str, dict] = {}
    for lang in mine:
        results, others, acc = evaluate_split(config, task_cfg, device, task_id, loaders[lang], model, criterion, step_fn=step_fn)
        if save_path is not None:
            os.makedirs(save_path, exist_ok=True)
            json.dump(results, open(os.path.join(save_path, lang + "_result.json"), "w"))
            json.dump(others, open(os.path.join(save_path, lang + "_others.json"), "w"))
        table[lang] = {"n": len(results), "accuracy": acc, "rank": rank}
    if dist_on and world_size > 1:
        parts: List[Optional[dict]] = [None] * world_size
        torch.distributed.all_gather_object(parts, table)
        table = {}
        for part in parts:
            table.update(part or {})
    return {lang: table[lang] for lang in languages if lang in table}

"""Task glue of the hot path with the call signatures of the reference's volta/volta/task_utils.py:
`ForwardModelsTrain` (L308-463, GQA branch L413-428), `ForwardModelsVal` (L195-304, GQA L265-269),
`EvaluatingModel` (L716-907, GQA L831-841) and `compute_score_with_logits` (L706-711).

The loss + prior term + score + d(loss)/d(logits) are ONE kernel (`clg_gqa_prior_loss`); eval does a
device-side argmax and a single D2H copy per batch instead of 2*B `.item()` syncs.
"""
from __future__ import annotations

import torch

from . import ops


class _GqaPriorLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, distances, lam, topk, score_mean):
        B, A = logits.shape
        logits = logits.contiguous().float()
        target = target.contiguous().float()
        dist = distances.contiguous().float() if distances is not None else None
        need_grad = logits.requires_grad or ctx.needs_input_grad[0]
        dl = torch.empty_like(logits) if need_grad else None
        row_buf = torch.empty(B, 4, dtype=torch.float32, device=logits.device)
        out = torch.empty(2, dtype=torch.float32, device=logits.device)
        ops.gqa_prior_loss(logits, target, dist, dlogits=dl, row_buf=row_buf, out=out, topk=topk, lam=lam,
                           gscale=1.0, score_mean=score_mean)
        ctx.dl = dl
        loss, score = out[0], out[1]
        ctx.mark_non_differentiable(score)
        return loss, score

    @staticmethod
    def backward(ctx, gloss, gscore):
        return ctx.dl * gloss, None, None, None, None, None


def gqa_prior_loss(logits, target, distances, semantic_lambda=10.0, topk=10):
    """(loss, batch_score) of task_utils.py:L413-428 as 0-dim device tensors (no host sync)."""
    return _GqaPriorLossFn.apply(logits, target, distances, float(semantic_lambda), int(topk), True)


def gqa_val_loss(logits, target):
    """task_utils.py:L265-269: CE only; the score is the un-normalised sum."""
    with torch.no_grad():
        return _GqaPriorLossFn.apply(logits, target, None, 0.0, 0, False)


def compute_score_with_logits(logits, labels):
    """task_utils.py:L706-711 (kept for callers that want the per-sample matrix)."""
    idx = torch.max(logits, 1)[1]
    one_hots = torch.zeros_like(labels)
    one_hots.scatter_(1, idx.view(-1, 1), 1)
    return one_hots * labels


def _to_device(batch, device):
    return tuple(t.cuda(device=device, non_blocking=True) if torch.is_tensor(t) else t for t in batch)


def ForwardModelsTrain(config, task_cfg, device, task_id, batch, model, criterion=None):
    """Same contract as the reference for `type: VL-classifier-GQA`, `process: normal`."""
    batch = _to_device(batch, device)
    features, spatials, image_mask, question, target, input_mask, segment_ids, question_id, ixs, distances = batch
    ttype = task_cfg[task_id]["type"]
    vil_prediction, _, _, _ = model(question, features, spatials, task_id, segment_ids, input_mask, image_mask)
    if ttype == "VL-classifier-GQA":
        loss, batch_score = gqa_prior_loss(vil_prediction, target, distances, task_cfg[task_id]["semantic_lambda"])
    else:
        raise ValueError("task type %s is outside the GQA hot path" % ttype)
    return loss, batch_score


def ForwardModelsVal(config, task_cfg, device, task_id, batch, model, criterion=None):
    batch = _to_device(batch, device)
    features, spatials, image_mask, question, target, input_mask, segment_ids, question_id = batch[:8]
    batch_size = features.size(0)
    with torch.no_grad():
        vil_prediction, _, _, _ = model(question, features, spatials, task_id, segment_ids, input_mask, image_mask)
        loss, batch_score = gqa_val_loss(vil_prediction, target)
    return float(loss), float(batch_score), batch_size


def EvaluatingModel(config, task_cfg, device, task_id, batch, model, dataloader, criterion, results, others):
    """GQA branch of task_utils.py:L716-907: argmax on the device, ONE D2H copy per batch."""
    batch = _to_device(batch, device)
    features, spatials, image_mask, question, target, input_mask, segment_ids, question_id = batch[:8]
    batch_size = features.size(0)
    with torch.no_grad():
        vil_prediction, _, _, _ = model(question, features, spatials, task_id, segment_ids, input_mask, image_mask)
        pred = torch.max(vil_prediction, 1)[1]
        both = torch.stack([pred, question_id.to(pred.dtype)], 0).cpu()      # the only host sync of the batch
    label2ans = getattr(getattr(dataloader, "dataset", None), "label2ans", None)
    for i in range(batch_size):
        p = int(both[0, i])
        results.append({"questionId": str(int(both[1, i])), "prediction": label2ans[p] if label2ans is not None else p})
    return 0.0, 0.0, batch_size, results, others

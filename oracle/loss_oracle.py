"""Oracle: GQA loss with the label-prior term, score and eval argmax.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows volta/volta/task_utils.py:L413-428 (training loss), L265-269 (validation loss, CE only),
L706-711 (compute_score_with_logits) and L831-841 (eval argmax).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def gqa_prior_loss(logits, target, distances, semantic_lambda=10.0, topk=10):
    """Returns (loss, batch_score) exactly as ForwardModelsTrain does for VL-classifier-GQA.

    loss = (CE(z, argmax(target.long())) + lambda * mean_b sum_{j in top10(softmax z)} p_j D_bj) * A
    """
    p_top_k, idx_top_k = torch.topk(F.softmax(logits, dim=-1), k=topk)          # L418
    sem = p_top_k * distances[torch.arange(distances.size(0)).unsqueeze(1), idx_top_k]  # L419
    sem = torch.mean(torch.sum(sem, dim=-1), dim=0)                               # L420
    loss = F.cross_entropy(logits, torch.argmax(target.long(), dim=1))            # L423
    loss = loss.mean() * target.size(1)                                           # L424
    loss = loss + (semantic_lambda * sem.mean()) * target.size(1)                 # L425
    score = compute_score_with_logits(logits, target).sum() / float(logits.size(0))   # L428
    return loss, score


def gqa_val_loss(logits, target):
    """task_utils.py:L265-269: CE only, score is the un-normalised sum."""
    loss = F.cross_entropy(logits, torch.argmax(target.long(), dim=1)).mean() * target.size(1)
    return loss, compute_score_with_logits(logits, target).sum()


def compute_score_with_logits(logits, labels):
    idx = torch.max(logits, 1)[1]
    one_hots = torch.zeros_like(labels)
    one_hots.scatter_(1, idx.view(-1, 1), 1)
    return one_hots * labels


def gqa_prior_loss_grad(logits, target, distances, semantic_lambda=10.0, topk=10):
    """Closed-form d loss / d logits, used to cross-check autograd."""
    B, A = logits.shape
    p = torch.softmax(logits, dim=-1)
    _, idx = torch.topk(p, k=topk)
    inK = torch.zeros_like(p)
    inK.scatter_(1, idx, 1.0)
    s = (p * distances * inK).sum(-1, keepdim=True)
    y = torch.argmax(target.long(), dim=1)
    onehot = torch.zeros_like(p)
    onehot[torch.arange(B), y] = 1.0
    return (A / B) * (p - onehot + semantic_lambda * p * (distances * inK - s))

"""Oracle: optimiser step of the fine-tune loop.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the algorithm is `AdamW` / `WarmupLinearSchedule` of the un-vendored pip
dependency pytorch-transformers (>=1.1.0, volta/requirements.txt:L51), absent from
/root/reference.  Restated from its published algorithm and anchored on
the reference's call sites: train_task.py:L249-276 (groups, AdamW ctor, schedule) and L326-338
(clip_grad_norm_(1.0) -> optimizer.step() -> scheduler.step() -> zero_grad()).
Cross-checked in tests against torch.optim.Adam on the case where they coincide
(weight_decay = 0, eps -> 0 difference bounded).
"""
from __future__ import annotations

import math

import torch


def warmup_linear_lr(base_lr: float, step: int, warmup_steps: float, t_total: float) -> float:
    """lr after `step` scheduler steps (LambdaLR multiplies base_lr by this factor)."""
    if step < warmup_steps:
        f = float(step) / float(max(1, warmup_steps))
    else:
        f = max(0.0, float(t_total - step) / float(max(1.0, t_total - warmup_steps)))
    return base_lr * f


def clip_grad_norm(grads, max_norm: float):
    """torch.nn.utils.clip_grad_norm_ (L2): g *= min(1, max_norm / (total_norm + 1e-6)).
    Returns (total_norm, clip_coef)."""
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads if g is not None))
    coef = min(1.0, max_norm / (total + 1e-6))
    if coef < 1.0:
        for g in grads:
            if g is not None:
                g.mul_(coef)
    return total, coef


def adamw_step(p, g, m, v, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-6,
               weight_decay=0.0, correct_bias=True):
    """One pytorch-transformers AdamW update, in place on (p, m, v); `step` is 1-based.

    m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ; denom = sqrt(v) + eps   (eps BEFORE bias corr.)
    step_size = lr * sqrt(1-b2^t)/(1-b1^t) ; p <- p - step_size * m/denom
    then decoupled decay on the UPDATED p: p <- p - lr*wd*p
    """
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    denom = v.sqrt().add_(eps)
    step_size = lr
    if correct_bias:
        step_size = step_size * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    p.addcdiv_(m, denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)
    return p


NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight")


def param_group_table(named_parameters, base_lr: float, weight_decay: float, sft: bool = False):
    """(name, lr, wd) per parameter, one group each, as train_task.py:L249-260 builds them
    (train_task_sft.py:L494 adds 'LayerNorm.weight_orig' to the no-decay list)."""
    no_decay = NO_DECAY + (("LayerNorm.weight_orig",) if sft else ())
    out = []
    for name, p in named_parameters:
        if not p.requires_grad:
            continue
        lr = 1e-4 if "vil_" in name else base_lr
        wd = 0.0 if any(nd in name for nd in no_decay) else weight_decay
        out.append((name, lr, wd))
    return out

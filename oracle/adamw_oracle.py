"""Oracle: optimiser step of the fine-tune loop.  TEST INFRASTRUCTURE ONLY.

The fine-tune loop takes `AdamW` / `WarmupLinearSchedule` from the un-vendored pip dependency
pytorch-transformers (>=1.1.0, volta/requirements.txt:L51; call sites train_task.py:L249-276 groups /
ctor / schedule, L326-338 clip_grad_norm_(1.0) -> optimizer.step() -> scheduler.step() -> zero_grad()).
PINNED (round 2): the reference tree vendors that optimiser's step body verbatim at
features_extraction/mmf/mmf/modules/optimizers.py:L15-84 (the HuggingFace AdamW = pytorch-transformers
1.x algorithm: eps before the bias correction, decoupled decay after the update); oracle/make_golden_adamw.py
executes it unmodified and tests/test_oracle_golden.py::test_adamw_oracle_and_module_match_reference_tree_step
requires bit-equal trajectories (tests/golden/adamw_ref_step.npz).  The schedule is pinned against
transformers.get_linear_schedule_with_warmup (the same lambda, installed package).
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

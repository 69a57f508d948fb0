"""Optimiser, schedule and clipping of the fine-tune loop with the semantics the reference gets from
pytorch-transformers (un-vendored pip dependency; call sites train_task.py:L23, L263-276, L326-338;
algorithm restated in SURVEY.md Appendix A.4):

    AdamW(params, lr, betas, eps, weight_decay, correct_bias)   - eps added BEFORE bias correction,
                                                                  decoupled decay AFTER the update
    WarmupLinearSchedule / WarmupConstantSchedule               - LambdaLR factors
    clip_grad_norm_(parameters, max_norm)                       - torch semantics

When every parameter lives in an `Engine` arena the whole step (global-norm clip, AdamW, decay, bf16
shadow refresh with the SFT mask) is ONE multi-tensor kernel over the flat arenas plus a two-pass
norm reduction (`clg_adamw_step`, `clg_sumsq`); no per-parameter launches, no host sync.
"""
from __future__ import annotations

import inspect
import math
import weakref

import os

import torch
from torch.optim import Optimizer
from torch.optim.lr_scheduler import LambdaLR

from . import ops

_ENGINE_OF_PARAM = weakref.WeakValueDictionary()   # id(param) -> Engine (filled by register_engine)


def register_engine(engine):
    for s in engine.slots:
        for p in s.params:
            _ENGINE_OF_PARAM[id(p)] = engine


def _engine_of_parameters_generator(parameters):
    """`clip_grad_norm_(model.parameters(), ...)` is called every step; materialising that generator walks the whole module
    tree (0.4 ms of host time).  When the argument IS the un-started `Module.parameters()` generator of a model whose
    parameters all live in one engine, the engine is returned without consuming it."""
    try:
        if not inspect.isgenerator(parameters) or parameters.gi_code.co_name != "parameters" or parameters.gi_frame is None:
            return None
        if inspect.getgeneratorstate(parameters) != "GEN_CREATED":
            return None
        loc = parameters.gi_frame.f_locals
        owner = loc.get("self")
        if not isinstance(owner, torch.nn.Module) or not loc.get("recurse", True):
            return None
        owner = getattr(owner, "module", owner)
        eng = getattr(owner, "__dict__", {}).get("_engine")
        if eng is None or eng.p_flat is None or eng.model is not owner:
            return None
        return eng
    except Exception:
        return None


def _engine_for(params):
    eng = None
    for p in params:
        e = _ENGINE_OF_PARAM.get(id(p))
        if e is None or e.p_flat is None or e._slot_of_param.get(id(p)) is None:
            return None
        if eng is None:
            eng = e
        elif eng is not e:
            return None
    return eng


def _owner_engine(opt):
    """the engine that owns ALL of `opt`'s parameters, or None (cached per optimiser; re-derived when its groups change)"""
    key = (len(opt.param_groups), id(opt.param_groups[0]["params"][0]) if opt.param_groups and opt.param_groups[0]["params"] else 0)
    hint = opt.__dict__.get("_eng_hint")
    if hint is not None and hint[0] == key:
        eng = hint[1]() if hint[1] is not None else None
        if eng is not None and eng.p_flat is not None:
            return eng
    eng = _engine_for([p for g in opt.param_groups for p in g["params"]])
    opt.__dict__["_eng_hint"] = (key, weakref.ref(eng) if eng is not None else None)
    return eng


def _clip_consumer(eng) -> bool:
    """True when a live AdamW of this module steps the engine's parameters: its step (fused or per-parameter) applies the
    clip coefficient that clip_grad_norm_ deferred"""
    return any(_owner_engine(o) is eng for o in list(_OPTIMIZERS))


def _mark_prescaled(eng):
    """the arena was just multiplied by eng.grad_scale (1/world): the engine's AdamW instances must not apply it again"""
    for o in list(_OPTIMIZERS):
        if getattr(o, "grad_scale", 1.0) != 1.0 and _owner_engine(o) is eng:
            o._grads_prescaled = True


_OPTIMIZERS = weakref.WeakSet()      # live AdamW instances (clip_grad_norm_'s torch fallback tells them it pre-scaled)
_FUSED_ZERO = os.environ.get("CLG_FUSED_ZERO", "1") != "0"    # the fused step clears the gradients it consumes
_PENDING_CLIP = {}   # id(engine) -> max_norm recorded by clip_grad_norm_ for the next fused step


def clip_grad_norm_(parameters, max_norm, norm_type=2.0):
    """torch.nn.utils.clip_grad_norm_ for engine-owned parameters: computes the global L2 norm of the
    gradient arena on the device (no host sync) and defers the scaling into the fused AdamW kernel.
    Falls back to torch's implementation for parameters that are not engine-owned."""
    eng = _engine_of_parameters_generator(parameters) if float(norm_type) == 2.0 else None
    if eng is not None:
        params = None                      # `model.parameters()` of an engine-owned model: no need to walk the module tree
    else:
        params = [p for p in (parameters if not isinstance(parameters, torch.Tensor) else [parameters])]
        eng = _engine_for(params)
    if eng is None or float(norm_type) != 2.0:
        scale = float(getattr(eng, "grad_scale", 1.0)) if eng is not None else 1.0
        if scale != 1.0:        # data parallel: the arena holds the SUM over ranks; clip the average, like the reference
            eng.g_flat.mul_(scale)
            _mark_prescaled(eng)
        return torch.nn.utils.clip_grad_norm_(params, max_norm, norm_type)
    eng._sumsq.zero_()
    # the norm pass also marks which 256-element blocks of the gradient arena hold a non-zero (eng._hot); the fused
    # AdamW of THIS gradient then skips the m / v traffic of blocks that have never received one
    sparse = (eng._table_blocks[0], eng._table_blocks[1], eng._touched) if eng._touch_valid else None
    ops.sumsq(eng.g_flat, eng._partials, eng._sumsq, hot=eng._hot, sparse=sparse)
    eng._hot_valid = True
    scale = float(getattr(eng, "grad_scale", 1.0))
    # the norm of the AVERAGED gradient (what the reference logs and clips after apex's 1/world)
    norm = eng._sumsq.sqrt() * scale
    if _clip_consumer(eng):
        _PENDING_CLIP[id(eng)] = float(max_norm)       # applied inside the next AdamW.step of this engine's parameters
    else:
        # no AdamW of this module owns the engine's parameters (a foreign optimiser steps them): nobody would pick the
        # deferred coefficient up, so the gradients are scaled here, like torch.nn.utils.clip_grad_norm_ (still no host sync)
        coef = (float(max_norm) / (norm + 1e-6)).clamp(max=1.0)
        eng.g_flat.mul_(coef * scale if scale != 1.0 else coef)
        if scale != 1.0:
            _mark_prescaled(eng)
    return norm


class AdamW(Optimizer):
    """pytorch-transformers 1.x AdamW (reference call site train_task.py:L263-267)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter: {}".format(betas))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(eps))
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, correct_bias=correct_bias)
        super().__init__(params, defaults)
        self._fused = None
        self.grad_scale = 1.0          # e.g. 1/world_size when gradients were summed across ranks
        self._grads_prescaled = False  # the torch clip fallback already multiplied the arena by grad_scale
        _OPTIMIZERS.add(self)

    # -- fused path ------------------------------------------------------------------------------
    def _fused_state(self):
        params = [p for g in self.param_groups for p in g["params"]]
        eng = _engine_for(params)
        if eng is None:
            return None
        fs = self._fused
        if fs is None or fs["engine"] is not eng or fs["p_flat_ptr"] != eng.p_flat.data_ptr():
            betas = {g["betas"] for g in self.param_groups}
            epss = {g["eps"] for g in self.param_groups}
            cbs = {g["correct_bias"] for g in self.param_groups}
            if len(betas) != 1 or len(epss) != 1 or len(cbs) != 1:
                return None
            m = torch.zeros_like(eng.p_flat)
            v = torch.zeros_like(eng.p_flat)
            fs = dict(engine=eng, p_flat_ptr=eng.p_flat.data_ptr(), m=m, v=v, step=0)
            # fresh moments (m = v = 0): a block may take the cold shortcut iff it has not held a non-zero gradient since.
            # eng._hot (zeroed when the arenas were built, only ever set by the norm pass) is a superset of those blocks -
            # it already holds the marks of the gradient this very step consumes (clip_grad_norm_ ran before us) - so it is
            # kept as it is; marks left by an earlier optimiser only make some blocks take the dense (always correct) path
            # carry over per-parameter state created by an earlier load_state_dict
            for g in self.param_groups:
                for p in g["params"]:
                    st = self.state.get(p)
                    s = eng._slot_of_param[id(p)]
                    o = s.start + sum(q.numel() for q in s.params[:[id(q) for q in s.params].index(id(p))])
                    mv = m[o:o + p.numel()].view(p.shape)
                    vv = v[o:o + p.numel()].view(p.shape)
                    if st and "exp_avg" in st:
                        mv.copy_(st["exp_avg"]); vv.copy_(st["exp_avg_sq"])
                        eng._hot.fill_(1)     # restored moments may be non-zero anywhere: dense updates
                        fs["step"] = max(fs["step"], int(st.get("step", 0)))
                    self.state[p] = {"step": fs["step"], "exp_avg": mv, "exp_avg_sq": vv}
            self._fused = fs
        return fs

    def _slot_table(self, eng):
        """(lr, wd) per arena slot from the parameter groups; members of a fused q|k|v slot must agree."""
        table = [None] * len(eng.slots)
        for g in self.param_groups:
            for p in g["params"]:
                s = eng._slot_of_param[id(p)]
                val = (float(g["lr"]), float(g["weight_decay"]))
                if table[s.index] is not None and table[s.index] != val:
                    raise ValueError("parameters fused into one arena slot (%s) need identical lr / weight_decay" % s.names)
                table[s.index] = val
        return table

    def step(self, closure=None):
        loss = closure() if closure is not None else None
        fs = self._fused_state()
        if fs is None:
            self._step_unfused()
            return loss
        eng = fs["engine"]
        table = self._slot_table(eng)
        from . import _lib as L
        for i, s in enumerate(eng.slots):
            lr, wd = table[i] if table[i] is not None else (0.0, 0.0)
            eng.metas_host[i].lr = lr
            eng.metas_host[i].wd = wd
        # slots without an optimiser entry (or that never receive gradients) are skipped via chunk_tensor
        key = tuple(1 if (table[i] is not None and eng.slots[i].has_grad) else 0 for i in range(len(eng.slots)))
        if fs.get("active_key") != key:
            ct = eng.chunk_tensor.cpu().clone()
            for i, s in enumerate(eng.slots):
                c0 = s.start // L.CHUNK
                c1 = (s.start + (s.numel + L.CHUNK - 1) // L.CHUNK * L.CHUNK) // L.CHUNK
                ct[c0:c1] = i if key[i] else -1
            fs["chunk_tensor"] = ct.to(eng.device)
            fs["active_key"] = key
        tens = torch.frombuffer(bytearray(bytes(eng.metas_host)), dtype=torch.uint8).to(eng.device, non_blocking=True)
        g0 = self.param_groups[0]
        fs["step"] += 1
        t = fs["step"]
        b1, b2 = g0["betas"]
        c1 = 1.0 - b1 ** t if g0["correct_bias"] else 1.0
        c2 = 1.0 - b2 ** t if g0["correct_bias"] else 1.0
        max_norm = _PENDING_CLIP.pop(id(eng), None)
        gscale = 1.0 if self._grads_prescaled else self.grad_scale
        self._grads_prescaled = False
        if not eng._hot_valid:
            # no norm pass marked the blocks of THIS gradient: the update is dense, and it makes m / v non-zero wherever the
            # gradient is - keep every block hot from here on, or a later clipped step would skip their moment decay
            eng._hot.fill_(1)
        with ops.nvtx_range("clg.adamw"):
            ops.adamw_step(eng.p_flat, eng.g_flat, fs["m"], fs["v"], fs["chunk_tensor"], tens,
                           sumsq=eng._sumsq if max_norm is not None else None,
                           max_norm=max_norm if max_norm is not None else 0.0, gscale=gscale,
                           beta1=b1, beta2=b2, eps=g0["eps"], bias_c1=c1, bias_c2=c2,
                           shadow=eng.shadow if eng.precision == "bf16" else None,
                           mask_bits=eng.mask_bits if eng.has_masks else None,
                           hot=eng._hot if eng._hot_valid else None, zero_grad=_FUSED_ZERO)
        eng._hot_valid = False
        eng.grads_zeroed = _FUSED_ZERO        # the step cleared the arena as it read it: the next backward skips its fill
        for p in self.state:
            self.state[p]["step"] = t
        if eng.precision == "bf16":
            eng.mark_shadow_fresh()
        else:
            eng._shadow_versions = None
        eng.grads_consumed = True
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._fused = None          # rebuilt on the next step from the restored per-parameter moments (dense updates)

    def zero_grad(self, set_to_none: bool = True):
        fs = self._fused
        if fs is not None and fs["engine"].grads_consumed:
            return          # the arena is re-zeroed by the next backward; nothing to do
        super().zero_grad(set_to_none=set_to_none)

    # -- reference (per-parameter) path ------------------------------------------------------------
    @torch.no_grad()
    def _step_unfused(self):
        gscale = 1.0 if self._grads_prescaled else float(self.grad_scale)
        self._grads_prescaled = False
        eng = _engine_for([p for g in self.param_groups for p in g["params"]])
        clip = _PENDING_CLIP.pop(id(eng), None) if eng is not None else None
        if clip is not None:      # a fused-path clip_grad_norm_ recorded the norm but this step cannot run fused
            total = float(eng._sumsq.sqrt()) * gscale
            gscale = gscale * min(1.0, clip / (total + 1e-6))
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                grad = p.grad if gscale == 1.0 else p.grad * gscale
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p)
                    state["exp_avg_sq"] = torch.zeros_like(p)
                exp_avg, exp_avg_sq = state["exp_avg"], state["exp_avg_sq"]
                beta1, beta2 = group["betas"]
                state["step"] += 1
                exp_avg.mul_(beta1).add_(grad, alpha=1.0 - beta1)
                exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1.0 - beta2)
                denom = exp_avg_sq.sqrt().add_(group["eps"])
                step_size = group["lr"]
                if group["correct_bias"]:
                    step_size = step_size * math.sqrt(1.0 - beta2 ** state["step"]) / (1.0 - beta1 ** state["step"])
                p.addcdiv_(exp_avg, denom, value=-step_size)
                if group["weight_decay"] > 0.0:
                    p.add_(p, alpha=-group["lr"] * group["weight_decay"])


class WarmupLinearSchedule(LambdaLR):
    """Linear warm-up then linear decay to 0 (reference call site train_task.py:L271-274)."""

    def __init__(self, optimizer, warmup_steps, t_total, last_epoch=-1):
        self.warmup_steps = warmup_steps
        self.t_total = t_total
        super().__init__(optimizer, self.lr_lambda, last_epoch=last_epoch)

    def lr_lambda(self, step):
        if step < self.warmup_steps:
            return float(step) / float(max(1, self.warmup_steps))
        return max(0.0, float(self.t_total - step) / float(max(1.0, self.t_total - self.warmup_steps)))


class WarmupConstantSchedule(LambdaLR):
    def __init__(self, optimizer, warmup_steps, last_epoch=-1):
        self.warmup_steps = warmup_steps
        super().__init__(optimizer, self.lr_lambda, last_epoch=last_epoch)

    def lr_lambda(self, step):
        if step < self.warmup_steps:
            return float(step) / float(max(1.0, self.warmup_steps))
        return 1.0


def grouped_parameters(model, base_lr, weight_decay, sft=False):
    """One group per parameter, decay membership by name substring (train_task.py:L249-260;
    train_task_sft.py:L494 adds LayerNorm.weight_orig)."""
    no_decay = ["bias", "LayerNorm.bias", "LayerNorm.weight"] + (["LayerNorm.weight_orig"] if sft else [])
    groups = []
    for key, value in dict(model.named_parameters()).items():
        if not value.requires_grad:
            continue
        lr = 1e-4 if "vil_" in key else base_lr
        wd = 0.0 if any(nd in key for nd in no_decay) else weight_decay
        groups.append({"params": [value], "lr": lr, "weight_decay": wd})
    return groups

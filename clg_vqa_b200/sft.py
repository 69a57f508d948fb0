"""Sparse fine-tuning (BERT-Tickets) mask application with the reference's semantics
(volta/train_task_sft.py:L44-215 `pruning_model_custom[_m3p]`, L409-461 main; train_utils.py:L375-416 `save_sft`).

The reference keeps `weight_orig` (Parameter) + `weight_mask` (fp32 {0,1} buffer) on each masked sub-module via
torch.nn.utils.prune and recomputes `weight = weight_orig * weight_mask` in a forward-pre-hook on every call
(145 elementwise multiplies per forward).  Here the same re-parametrisation objects are installed -
so state_dict keys, optimiser parameter names and `save_sft` behave identically - but nothing is recomputed per
call: the engine packs the masks to 1 bit / weight, applies them when it writes the bf16 weight shadow (inside
the AdamW kernel) and in the wgrad GEMM epilogue.
"""
from __future__ import annotations

from typing import Dict, Iterable, List

import torch
import torch.nn.utils.prune as prune

from . import synthetic as S


def uc2_masked_module_names(num_layers: int = 12) -> List[str]:
    """the 73 modules of train_task_sft.py:L57-82 (shared modules are listed under their primary name)"""
    names = []
    for l in range(num_layers):
        a, f = 2 * l, 2 * l + 1
        names += ["bert.encoder.layer.%d.attention_self.query" % a, "bert.encoder.layer.%d.attention_self.key" % a,
                  "bert.encoder.layer.%d.attention_self.value" % a, "bert.encoder.layer.%d.attention_output.dense" % a,
                  "bert.encoder.layer.%d.intermediate.dense" % f, "bert.encoder.layer.%d.output.dense" % f]
    names.append("bert.t_pooler.dense")
    return names


def m3p_masked_module_names(num_layers: int = 12) -> List[str]:
    """the live subset of train_task_sft.py:L139-205 (attention + FFN); the never-executed modules the
    reference also masks (encoder_attn, latent / original transforms, ...) can be passed explicitly"""
    names = []
    for i in range(num_layers):
        names += ["bert.encoder.attentions.%d.%s" % (i, n) for n in ("q_lin", "k_lin", "v_lin", "out_lin")]
        names += ["bert.encoder.ffns.%d.lin1" % i, "bert.encoder.ffns.%d.lin2" % i]
    return names


def pruning_model_custom(model, mask_dict: Dict[str, torch.Tensor], module: str = "", names: Iterable[str] = None):
    """Install `prune.CustomFromMask` on every named sub-module that has a `<module><name>.weight_mask`
    entry in mask_dict (train_task_sft.py:L122-132)."""
    mods = dict(model.named_modules())
    if names is None:
        names = [k[len(module):-len(".weight_mask")] for k in mask_dict if k.startswith(module) and k.endswith(".weight_mask")]
    done = set()
    for name in names:
        m = mods.get(name)
        if m is None or id(m) in done:
            continue                      # aliases (v_query, ...) are the same module object
        key = "%s%s.weight_mask" % (module, name)
        if key not in mask_dict or "weight_orig" in m._parameters:
            continue
        w = m._parameters["weight"]
        prune.CustomFromMask.apply(m, "weight", mask=mask_dict[key].to(device=w.device, dtype=w.dtype))
        done.add(id(m))
    return model


def premask_state_dict(state_dict: Dict[str, torch.Tensor], mask_dict: Dict[str, torch.Tensor]):
    """train_task_sft.py:L432-453: multiply the masked weights of a plain state_dict by their masks."""
    out = dict(state_dict)
    for k, mask in mask_dict.items():
        if not k.endswith(".weight_mask"):
            continue
        wk = k[:-len("_mask")]
        if wk in out:
            out[wk] = out[wk] * mask.to(out[wk].dtype)
    return out


def synth_mask_dict(model, names: Iterable[str], density: float = 0.59, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic {0,1} masks (hash-generated, reproducible on any machine) for the given modules."""
    mods = dict(model.named_modules())
    out = {}
    for name in names:
        m = mods[name]
        w = m._parameters.get("weight", m._parameters.get("weight_orig"))
        n = w.numel()
        u = S.hash_uniform(S._name_seed(seed + 7919, name + ".weight_mask"), n).reshape(w.shape)
        out[name + ".weight_mask"] = (u < density).to(torch.float32)
    return out


def fold_masks(model) -> Dict[str, torch.Tensor]:
    """`save_sft` semantics (train_utils.py:L375-416): a plain state_dict whose masked weights are
    weight_orig * weight_mask, without the *_orig / *_mask keys."""
    sd = model.state_dict()
    out = {}
    for k, v in sd.items():
        if k.endswith("_mask"):
            continue
        if k.endswith("_orig"):
            base = k[:-len("_orig")]
            out[base] = v * sd[base + "_mask"]
        else:
            out[k] = v
    return out

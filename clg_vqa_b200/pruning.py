"""Iterative magnitude pruning (IMP / BERT-Tickets) helpers with the reference's names and semantics
(volta/train_task_prunning.py: `pruning_model_uc2` L45-90, `pruning_model_m3p`, `see_weight_rate_uc2` L92-177,
`rewind_uc2` L179-240, main loop L728-806).

The reference calls `prune.global_unstructured(parameters, L1Unstructured, amount=px)`: torch concatenates all |weight|
scores (342 MB for the 73 UC2 matrices), takes topk(k = px * #unpruned, largest=False) and stacks one more mask on every
module.  Here the k-th smallest magnitude among the unpruned entries is found on the device by a radix select
(`clg_prune_hist` x 4 + `clg_prune_apply`, csrc/prune.cu) and the `weight_mask` buffers are updated in place; the same
`torch.nn.utils.prune` re-parametrisation objects are installed (CustomFromMask), so `state_dict()` keys, `prune.remove`,
`save_sft` and the SFT loader see exactly what the reference produces.  The masks equal torch's whenever the k-th
magnitude is unique (ties are broken arbitrarily by both).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Tuple

import torch
import torch.nn.utils.prune as prune

from . import _lib as L
from . import ops
from . import sft


def uc2_prunable_names(num_layers: int = 12, embeddings=False, cls=False, task="TASK15") -> List[str]:
    names = []
    if embeddings:
        names += ["bert.embeddings.word_embeddings", "bert.embeddings.image_embeddings"]
    names += sft.uc2_masked_module_names(num_layers)
    if cls:
        names += ["clfs_dict.%s.logit_fc.%d" % (task, i) for i in (0, 2, 3)]
    return names


def _unique_modules(model, names: Iterable[str]):
    mods = dict(model.named_modules())
    out, seen = [], set()
    for n in names:
        m = mods.get(n)
        if m is None or id(m) in seen:
            continue
        seen.add(id(m))
        out.append((n, m))
    return out


def global_magnitude_prune(model, names: Iterable[str], amount: float) -> Tuple[int, int]:
    """prune.global_unstructured(L1Unstructured, amount) over `names` on the device.  Returns (#pruned now, #unpruned
    before).  Works on modules that are already pruned (amount is a fraction of the REMAINING weights) and on fresh ones."""
    mods = _unique_modules(model, names)
    if not mods:
        return 0, 0
    dev = None
    entries = []
    for n, m in mods:
        if "weight_orig" in m._parameters:
            w = m._parameters["weight_orig"].data
            mask = m._buffers["weight_mask"]
        else:
            w = m._parameters["weight"].data
            mask = None
        if not w.is_cuda:
            raise ops.ClgError("global_magnitude_prune runs on the GPU (module %s is on %s); there is no CPU fallback" % (n, w.device))
        dev = w.device
        entries.append((n, m, w.contiguous(), mask))
    arr = (L.PruneTensor * len(entries))()
    new_masks = []
    for i, (n, m, w, mask) in enumerate(entries):
        out = mask if mask is not None else torch.ones_like(w)
        new_masks.append(out)
        arr[i] = L.PruneTensor(w.data_ptr(), mask.data_ptr() if mask is not None else None, out.data_ptr(), w.numel())
    tens = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
    stream = torch.cuda.current_stream().cuda_stream
    lib = L.load()
    hist = torch.zeros(256, dtype=torch.int64, device=dev)

    def histogram(prefix, pmask, shift):
        hist.zero_()
        L.check(lib.clg_prune_hist(tens.data_ptr(), len(entries), prefix, pmask, shift, hist.data_ptr(), stream), "clg_prune_hist")
        return hist.cpu().tolist()

    h = histogram(0, 0, 24)
    n_unpruned = sum(h)
    k = int(round(float(amount) * n_unpruned))          # L1Unstructured: round(amount * t.nelement()) of the unpruned
    if k <= 0:
        return 0, n_unpruned
    if k >= n_unpruned:
        k = n_unpruned
    prefix, pmask, below = 0, 0, 0                      # `below`: entries strictly smaller than every key with this prefix
    for shift in (24, 16, 8, 0):
        if shift != 24:
            h = histogram(prefix, pmask, shift)
        run = below
        for d in range(256):
            if run + h[d] >= k:
                prefix |= d << shift
                below = run
                break
            run += h[d]
        pmask |= 0xFF << shift
    ties_to_prune = k - below                           # entries equal to the threshold that still have to go
    counter = torch.zeros(1, dtype=torch.int64, device=dev)
    L.check(lib.clg_prune_apply(tens.data_ptr(), len(entries), prefix, ties_to_prune, counter.data_ptr(), stream), "clg_prune_apply")
    for (n, m, w, mask), new in zip(entries, new_masks):
        if mask is None:
            prune.CustomFromMask.apply(m, "weight", mask=new)
        else:
            # the kernel rewrote the existing `weight_mask` through its raw pointer: tell torch (version counter), so that
            # Engine._refresh - which keys the packed mask bits and the masked bf16 shadow on (data_ptr, _version) -
            # repacks them before the next forward
            mask.mul_(1)
    eng = getattr(model, "module", model).__dict__.get("_engine")
    if eng is not None:
        eng.invalidate_masks()
    return k, n_unpruned


def pruning_model_uc2(model, px, embeddings=False, global_pruning=False, cls=False, bias=False):
    """train_task_prunning.py:L45-90.  global_pruning=True uses the device radix select; the per-tensor variant and the
    bias option defer to torch.nn.utils.prune (not on the hot path)."""
    n_layers = len(getattr(model.config, "t_ff_sublayers", range(12)))
    names = uc2_prunable_names(n_layers, embeddings=embeddings, cls=cls)
    if global_pruning and not bias:
        return global_magnitude_prune(model, names, px)
    for n, m in _unique_modules(model, names):
        prune.l1_unstructured(m, "weight", amount=px)
        if bias:
            prune.l1_unstructured(m, "bias", amount=px)
    return None


def pruning_model_m3p(model, px, global_pruning=True):
    n_layers = int(model.config.n_layers)
    names = sft.m3p_masked_module_names(n_layers)
    if global_pruning:
        return global_magnitude_prune(model, names, px)
    for n, m in _unique_modules(model, names):
        prune.l1_unstructured(m, "weight", amount=px)
    return None


def see_weight_rate(model, names: Iterable[str]) -> float:
    """percentage of zeros in the masks of `names` (see_weight_rate_uc2 / _m3p)"""
    total, zeros = 0.0, 0.0
    for n, m in _unique_modules(model, names):
        mask = m._buffers.get("weight_mask")
        if mask is None:
            total += float(m._parameters["weight"].numel())
            continue
        total += float(mask.numel())
        zeros += float((mask == 0).sum())
    return 100.0 * zeros / max(total, 1.0)


def rewind(pre_weight: Dict[str, torch.Tensor], names: Iterable[str], prefix: str = "") -> Dict[str, torch.Tensor]:
    """rewind_uc2 / rewind_m3p: the `<name>.weight` entries of an unpruned state_dict re-keyed as `<name>.weight_orig`
    (loaded back into the pruned model before the next IMP round)."""
    out = {}
    for n in names:
        k = "%s%s.weight" % (prefix, n)
        if k in pre_weight:
            out[k + "_orig"] = pre_weight[k]
    return out

"""IMP global magnitude pruning (clg_vqa_b200/pruning.py, csrc/prune.cu) against torch.nn.utils.prune.global_unstructured
on the same weights: identical masks (no ties in random fp32 weights), also for the second round on the remaining weights."""
import copy

import pytest
import torch
import torch.nn.utils.prune as prune

pytestmark = pytest.mark.gpu

from clg_vqa_b200 import config as C  # noqa: E402
from clg_vqa_b200 import pruning, sft  # noqa: E402
from clg_vqa_b200.encoders import BertForVLTasks  # noqa: E402

DEV = "cuda"


def test_global_magnitude_prune_matches_torch_two_rounds():
    cfg = C.uc2_base(2, vocab_size=512)
    torch.manual_seed(0)
    ours = BertForVLTasks(cfg, C.gqa_task_cfg(), ["TASK15"]).to(DEV)
    ref = copy.deepcopy(ours)
    names = pruning.uc2_prunable_names(2)
    rmods = dict(ref.named_modules())
    seen, plist = set(), []
    for n in names:
        if id(rmods[n]) not in seen:
            seen.add(id(rmods[n])); plist.append((rmods[n], "weight"))
    for rnd, amount in enumerate((0.1, 0.25)):
        prune.global_unstructured(tuple(plist), pruning_method=prune.L1Unstructured, amount=amount)
        k, n_un = pruning.pruning_model_uc2(ours, amount, global_pruning=True)
        omods = dict(ours.named_modules())
        total = zeros = 0
        for n in names:
            mo, mr = omods[n].weight_mask, rmods[n].weight_mask
            assert torch.equal(mo, mr), (rnd, n)
            total += mo.numel(); zeros += int((mo == 0).sum())
        assert abs(pruning.see_weight_rate(ours, names) - 100.0 * zeros / total) < 1e-9
        assert k == int(round(amount * n_un))
    sd = ours.state_dict()
    assert "bert.encoder.layer.0.attention_self.query.weight_orig" in sd and "bert.encoder.layer.0.attention_self.v_query.weight_mask" in sd
    # the pruned model runs through the engine (packed masks) and masked gradients vanish
    assert 32.0 < pruning.see_weight_rate(ours, names) < 33.0          # 1 - 0.9 * 0.75


def test_prune_ties_and_zero_weights():
    """ties at the threshold: exactly k entries go, whichever they are"""
    lin = torch.nn.Linear(64, 64, bias=False).to(DEV)
    with torch.no_grad():
        lin.weight.copy_(torch.randint(0, 4, (64, 64)).float())       # heavy ties, many exact zeros
    class M(torch.nn.Module):
        def __init__(self): super().__init__(); self.a = lin
    m = M()
    k, n_un = pruning.global_magnitude_prune(m, ["a"], 0.3)
    assert n_un == 4096 and k == round(0.3 * 4096)
    mask = m.a.weight_mask
    assert int((mask == 0).sum()) == k
    w = m.a.weight_orig.abs()
    assert float(w[mask == 0].max()) <= float(w[mask == 1].min())


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_second_prune_round_reaches_the_engine(precision):
    """IMP (train_task_prunning.py:L728-806): prune, train, prune again.  The second round rewrites the existing
    `weight_mask` buffers in place; the engine must repack its mask bits and re-cast the masked weight shadow, i.e. the
    next forward uses the NEW masks: gradients vanish exactly at the newly pruned positions and the logits equal those
    of a fresh model loaded with the folded weights W_orig * M."""
    from clg_vqa_b200 import synthetic as S, task_utils as TU
    cfg = C.uc2_base(2, vocab_size=512)
    task_cfg = C.gqa_task_cfg()
    model = BertForVLTasks(cfg, task_cfg, ["TASK15"])
    model.load_state_dict(S.synth_state_dict(model, seed=2))
    model.set_precision(precision)
    model.to(DEV).eval()
    names = pruning.uc2_prunable_names(2)
    b = S.synth_batch(4, seed=5, ragged_regions=True, vocab_size=512)
    batch = S.batch_tuple(b)

    def step():
        model.zero_grad()
        loss, _ = TU.ForwardModelsTrain(cfg, task_cfg, DEV, "TASK15", batch, model)
        loss.backward()
        with torch.no_grad():
            return model(b["question"].to(DEV), b["features"].to(DEV), b["spatials"].to(DEV), "TASK15", b["segment_ids"].to(DEV),
                         b["input_mask"].to(DEV), b["image_mask"].to(DEV))[0].clone()

    pruning.pruning_model_uc2(model, 0.2, global_pruning=True)
    step()
    mods = dict(model.named_modules())
    qmod = mods["bert.encoder.layer.0.attention_self.query"]
    mask1 = qmod.weight_mask.clone()
    pruning.pruning_model_uc2(model, 0.5, global_pruning=True)        # second round, in place on the existing masks
    mask2 = qmod.weight_mask
    newly = (mask1 == 1) & (mask2 == 0)
    assert int(newly.sum()) > 0
    logits = step()
    for n in names:
        m = mods[n]
        assert float(m.weight_orig.grad[m.weight_mask == 0].abs().max()) == 0.0, n
    assert float(qmod.weight_orig.grad[mask2 == 1].abs().max()) > 0.0
    # fresh model with the folded weights (what save_sft writes): identical forward
    folded = {}
    for k, v in model.state_dict().items():
        if k.endswith(".weight_mask"):
            continue
        if k.endswith(".weight_orig"):
            folded[k[:-5]] = v * model.state_dict()[k[:-5] + "_mask"]
        else:
            folded[k] = v
    fresh = BertForVLTasks(cfg, task_cfg, ["TASK15"])
    fresh.load_state_dict(folded, strict=True)
    fresh.set_precision(precision)
    fresh.to(DEV).eval()
    with torch.no_grad():
        ref = fresh(b["question"].to(DEV), b["features"].to(DEV), b["spatials"].to(DEV), "TASK15", b["segment_ids"].to(DEV),
                    b["input_mask"].to(DEV), b["image_mask"].to(DEV))[0]
    assert torch.equal(logits, ref)

"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.
TEST INFRASTRUCTURE ONLY.  Run in the build container:  python -m oracle.make_golden

For every case: build the reference model (`volta.encoders.BertForVLTasks` / `M3PForVLTasks`),
load hash-generated weights (clg_vqa_b200.synthetic.synth_state_dict - reproducible anywhere), run
the reference's own `volta.task_utils.ForwardModelsTrain` (fp32, eval-mode dropout) forward + loss +
backward on a hash-generated batch and store: logits, loss, score, per-parameter gradient norms, a
few full gradients, plus the logits / loss of (a) the fp64 reference and (b) the reference under
the bf16 operand-rounding shim.  Weights and inputs are NOT stored - the
tests regenerate them from the seeds recorded in the fixture.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from clg_vqa_b200 import config as C  # noqa: E402
from clg_vqa_b200 import synthetic as S  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (family, layers, vocab, B, R, ragged, weight_seed, batch_seed)
    "uc2_L2_B4": ("uc2", 2, 4096, 4, 36, True, 0, 1234),
    "uc2_L12_B8": ("uc2", 12, 250002, 8, 36, False, 0, 1234),      # BASELINE.json configs[0]
    "m3p_L2_B4": ("m3p", 2, 4096, 4, 100, True, 0, 4321),
    "m3p_L12_B4": ("m3p", 12, 250002, 4, 100, False, 0, 4321),
}

KEEP_GRADS_UC2 = [
    "bert.embeddings.LayerNorm.weight", "bert.embeddings.image_location_embeddings.weight",
    "bert.embeddings.v_LayerNorm.bias", "bert.embeddings.new_token_type_embeddings.weight",
    "bert.encoder.layer.0.attention_self.query.bias", "bert.encoder.layer.0.attention_self.key.bias",
    "bert.encoder.layer.1.output.LayerNorm.weight", "bert.encoder.layer.1.intermediate.dense.bias",
    "bert.t_pooler.dense.bias", "clfs_dict.TASK15.logit_fc.2.weight", "clfs_dict.TASK15.logit_fc.3.bias",
]
KEEP_GRADS_M3P = [
    "bert.encoder.layer_norm_emb.weight", "bert.encoder.image_embeddings.image_location_embeddings.weight",
    "bert.encoder.attentions.0.q_lin.bias", "bert.encoder.layer_norm2.1.weight",
    "bert.encoder.ffns.0.lin1.bias", "bert.encoder.pooled_layer.dense.bias",
    "clfs_dict.TASK15.logit_fc.2.weight", "clfs_dict.TASK15.logit_fc.3.bias",
]


def ref_config(ref, family, layers, vocab):
    if family == "uc2":
        mine = C.uc2_base(layers, vocab_size=vocab)
        return ref.BertConfig.from_dict(mine.to_dict()), mine
    mine = C.m3p_base(layers, vocab_size=vocab)
    d = mine.to_dict()
    return ref.M3PConfig.from_dict(d), mine


def build_ref_model(ref, family, cfg, task_cfg, seed):
    cls = ref.BertForVLTasks if family == "uc2" else ref.M3PForVLTasks
    model = cls(cfg, task_cfg, ["TASK15"])
    sd = S.synth_state_dict(model, seed=seed)
    missing = model.load_state_dict(sd, strict=True)
    model.eval()
    return model


def batch_for(family, B, R, ragged, seed, vocab):
    if family == "uc2":
        return S.synth_batch(B, num_regions=R, num_locs=7, seed=seed, ragged_regions=ragged, vocab_size=vocab)
    # M3P: prefix mask semantic -> ragged handled through the masks only; loc rows L2-normalised
    return S.synth_batch(B, num_regions=R, num_locs=5, seed=seed, ragged_regions=ragged, vocab_size=vocab,
                         normalize=True)


PROBE_ROW, PROBE_COL = 7, 16        # sub-sampling of the per-layer hidden-state probe


def hidden_hooks(model, family):
    """forward hooks that record the joint hidden state after the embeddings and after every encoder layer of the
    UNMODIFIED reference (UC2: cat(text, regions) after each feed-forward sub-layer, encoders.py:L866-892; M3P: the
    output of layer_norm_emb / layer_norm2[i], m3p_transformer.py:L940-948)."""
    captured, hooks = [], []
    if family == "uc2":
        def cat_hook(_m, _inp, outp):
            captured.append(torch.cat([outp[0], outp[1]], dim=1).detach().clone())
        hooks.append(model.bert.embeddings.register_forward_hook(cat_hook))
        for i, layer in enumerate(model.bert.encoder.layer):
            if i % 2 == 1:
                hooks.append(layer.register_forward_hook(cat_hook))
    else:
        def one_hook(_m, _inp, outp):
            captured.append(outp.detach().clone())
        enc = model.bert.encoder
        hooks.append(enc.layer_norm_emb.register_forward_hook(one_hook))
        for i in range(enc.n_layers):
            hooks.append(enc.layer_norm2[i].register_forward_hook(one_hook))
    return hooks, captured


def run_reference(ref, model, cfg, task_cfg, batch, backward=True, forward_again=True):
    """the reference's own ForwardModelsTrain (loss, score); the logits are picked up by a forward hook on the model"""
    tu = ref.task_utils
    crit = tu.LossMap["CrossEntropyLoss"]
    model.zero_grad()
    got = []
    hk = model.register_forward_hook(lambda _m, _i, outp: got.append(outp[0].detach().clone()))
    try:
        with ref_shim.cpu_cuda_noop():
            if backward:
                loss, score = tu.ForwardModelsTrain(cfg, task_cfg, "cpu", "TASK15", S.batch_tuple(batch), model, crit)
                loss.backward()
            else:
                with torch.no_grad():
                    loss, score = tu.ForwardModelsTrain(cfg, task_cfg, "cpu", "TASK15", S.batch_tuple(batch), model, crit)
    finally:
        hk.remove()
    return got[0], float(loss), float(score)


def make_case(ref, name, spec):
    family, layers, vocab, B, R, ragged, wseed, bseed = spec
    t0 = time.time()
    task_cfg = C.gqa_task_cfg(max_region_num=R)
    cfg, _ = ref_config(ref, family, layers, vocab)
    model = build_ref_model(ref, family, cfg, task_cfg, wseed)
    batch = batch_for(family, B, R, ragged, bseed, vocab)
    out = {}
    logits, loss, score = run_reference(ref, model, cfg, task_cfg, batch, backward=True)
    out["logits_fp32"] = logits.numpy()
    out["loss_fp32"] = np.float64(loss)
    out["score_fp32"] = np.float64(score)
    names, norms = [], []
    seen = set()
    for n, p in model.named_parameters():
        if p.grad is None or id(p) in seen:
            continue
        seen.add(id(p))
        names.append(n)
        norms.append(float(p.grad.double().norm()))
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms, dtype=np.float64)
    keep = KEEP_GRADS_UC2 if family == "uc2" else KEEP_GRADS_M3P
    pd = dict(model.named_parameters())
    for k in keep:
        if k in pd and pd[k].grad is not None:
            out["grad::" + k] = pd[k].grad.numpy().copy()
    # word-embedding gradient: rows touched + their values for the first 64 touched rows
    wkey = "bert.embeddings.word_embeddings.weight" if family == "uc2" else "bert.encoder.embeddings.weight"
    wg = pd[wkey].grad
    rows = torch.nonzero(wg.abs().sum(1) > 0).flatten()
    out["wordgrad_rows"] = rows.numpy()
    out["wordgrad_vals"] = wg[rows[:64]].numpy().copy()
    # bf16 operand shim (fp32 accumulate) and fp64 reference: forward only; per-layer hidden states of the shim run are
    # sub-sampled into a probe that pins the oracle port's hidden states (the teacher-forced per-layer GPU test feeds them)
    hooks, captured = hidden_hooks(model, family)
    with ref_shim.bf16_operand_shim():
        lg, ls, sc = run_reference(ref, model, cfg, task_cfg, batch, backward=False, forward_again=False)
    for h in hooks:
        h.remove()
    out["logits_bf16shim"] = lg.numpy()
    out["loss_bf16shim"] = np.float64(ls)
    hid = torch.stack(captured)                                    # [layers + 1, B, S, d]
    if family == "m3p":                                            # rows past the prefix length are re-zeroed (L948)
        n = batch["input_mask"].sum(1) + batch["image_mask"].sum(1)
        valid = (torch.arange(hid.shape[2]).unsqueeze(0) < n.unsqueeze(1)).to(hid.dtype)
        hid[1:] = hid[1:] * valid[None, :, :, None]
    out["hidden_bf16shim_probe"] = hid[:, :, ::PROBE_ROW, ::PROBE_COL].contiguous().numpy()
    # bf16 TRAINING shim: operands rounded in the forward and in the backward contractions (oracle/bf16_autograd.py)
    with ref_shim.bf16_train_shim():
        lgb, lsb, scb = run_reference(ref, model, cfg, task_cfg, batch, backward=True, forward_again=False)
    assert float((lgb - lg).abs().max()) == 0.0, "train shim forward differs from the operand shim"
    pdb = dict(model.named_parameters())
    out["gradb_norms"] = np.array([float(pdb[n_].grad.double().norm()) if pdb[n_].grad is not None else 0.0 for n_ in names],
                                  dtype=np.float64)
    for k in keep:
        if k in pdb and pdb[k].grad is not None:
            out["gradb::" + k] = pdb[k].grad.numpy().copy()
    out["wordgradb_vals"] = pdb[wkey].grad[rows[:64]].numpy().copy()
    model.zero_grad()
    model.double()
    b64 = dict(batch)
    for k in ("features", "spatials", "target", "distances"):
        b64[k] = batch[k].double()
    lg, ls, sc = run_reference(ref, model, cfg, task_cfg, b64, backward=False)
    out["logits_fp64"] = lg.numpy()
    out["loss_fp64"] = np.float64(ls)
    meta = dict(case=name, family=family, layers=layers, vocab=vocab, B=B, R=R, ragged=ragged,
                weight_seed=wseed, batch_seed=bseed, torch=torch.__version__,
                input_checksum=float(batch["features"].double().sum() + batch["spatials"].double().sum()
                                     + batch["question"].double().sum()),
                reference="unmodified volta.task_utils.ForwardModelsTrain + volta.encoders (eval-mode dropout)")
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print("%s: loss %.6f score %.4f  (%.1fs)" % (name, loss, score, time.time() - t0), flush=True)


def make_sft_case(ref, name="uc2_L2_B4_sft", layers=2, vocab=4096, B=4, R=36, wseed=0, bseed=1234, density=0.59):
    """SFT (train_task_sft.py:L409-461): pre-multiply the masked weights, install prune.CustomFromMask on the named
    sub-modules of the UNMODIFIED reference model (what pruning_model_custom does at L122-132; its hard-coded
    24-sub-layer loop is applied to this model's sub-layers), then the standard forward / loss / backward."""
    import torch.nn.utils.prune as prune
    from clg_vqa_b200 import sft
    t0 = time.time()
    task_cfg = C.gqa_task_cfg(max_region_num=R)
    cfg, _ = ref_config(ref, "uc2", layers, vocab)
    model = build_ref_model(ref, "uc2", cfg, task_cfg, wseed)
    names = sft.uc2_masked_module_names(layers)
    masks = sft.synth_mask_dict(model, names, density=density, seed=wseed)
    model.load_state_dict(sft.premask_state_dict(model.state_dict(), masks))           # L432-453
    mods = dict(model.named_modules())
    for n in names:
        prune.CustomFromMask.apply(mods[n], "weight", mask=masks[n + ".weight_mask"])  # L122-132
    model.eval()
    batch = batch_for("uc2", B, R, True, bseed, vocab)
    logits, loss, score = run_reference(ref, model, cfg, task_cfg, batch, backward=True)
    out = {"logits_fp32": logits.numpy(), "loss_fp32": np.float64(loss), "score_fp32": np.float64(score)}
    gn, gv = [], []
    seen = set()
    for n, p in model.named_parameters():
        if p.grad is None or id(p) in seen:
            continue
        seen.add(id(p))
        gn.append(n); gv.append(float(p.grad.double().norm()))
    out["grad_names"] = np.array(gn); out["grad_norms"] = np.array(gv, dtype=np.float64)
    q = dict(model.named_parameters())["bert.encoder.layer.0.attention_self.query.weight_orig"]
    out["grad::bert.encoder.layer.0.attention_self.query.weight_orig[:8]"] = q.grad[:8].numpy().copy()
    mk = masks["bert.encoder.layer.0.attention_self.query.weight_mask"]
    assert float(q.grad[mk == 0].abs().max()) == 0.0
    out["state_keys"] = np.array(sorted(model.state_dict().keys()))
    meta = dict(case=name, family="uc2", layers=layers, vocab=vocab, B=B, R=R, ragged=True, weight_seed=wseed, batch_seed=bseed,
                mask_density=density, torch=torch.__version__,
                input_checksum=float(batch["features"].double().sum() + batch["spatials"].double().sum()
                                     + batch["question"].double().sum()),
                reference="unmodified reference + torch.nn.utils.prune.CustomFromMask on the train_task_sft.py module list")
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print("%s: loss %.6f  (%.1fs)" % (name, loss, time.time() - t0), flush=True)


def main(argv):
    ref = ref_shim.import_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    which = argv[1:] or (list(CASES) + ["uc2_L2_B4_sft"])
    for name in which:
        if name == "uc2_L2_B4_sft":
            make_sft_case(ref)
        else:
            make_case(ref, name, CASES[name])


if __name__ == "__main__":
    main(sys.argv)

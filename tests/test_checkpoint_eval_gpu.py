"""f4 / f3 on the device: asynchronous checkpoints of the engine arenas (parameters + fused AdamW moments) and the
language-sharded evaluation loop on the real model."""
import json

import pytest
import torch

pytestmark = pytest.mark.gpu

from clg_vqa_b200 import checkpoint, config as C, evaluation, optim, sft  # noqa: E402
from clg_vqa_b200 import synthetic as S  # noqa: E402
from clg_vqa_b200 import task_utils as TU  # noqa: E402
from clg_vqa_b200.encoders import BertForVLTasks  # noqa: E402

DEV = "cuda"
LAYERS, VOCAB, R = 2, 3000, 12


def _model(precision="bf16", masks=False, seed=5):
    cfg = C.uc2_base(LAYERS, vocab_size=VOCAB)
    task_cfg = C.gqa_task_cfg(max_region_num=R)
    model = BertForVLTasks(cfg, task_cfg, ["TASK15"])
    model.load_state_dict(S.synth_state_dict(model, seed=seed), strict=True)
    md = None
    if masks:
        names = sft.uc2_masked_module_names(LAYERS)
        md = sft.synth_mask_dict(model, names, density=0.6, seed=seed)
        model.load_state_dict(sft.premask_state_dict(model.state_dict(), md))
        sft.pruning_model_custom(model, md, names=names)
    model.set_precision(precision)
    model.to(DEV)
    return model, cfg, task_cfg, md


def _step(model, cfg, task_cfg, opt, seed):
    model.train()
    b = S.synth_batch(8, num_regions=R, seed=seed, vocab_size=VOCAB)
    loss, _ = TU.ForwardModelsTrain(cfg, task_cfg, DEV, "TASK15", S.batch_tuple(b), model)
    loss.backward()
    optim.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
    model.zero_grad()


def test_async_save_snapshots_the_arenas(tmp_path):
    model, cfg, task_cfg, _ = _model()
    opt = optim.AdamW(optim.grouped_parameters(model, 4e-5, 1e-4), lr=4e-5)
    sched = optim.WarmupLinearSchedule(opt, warmup_steps=2, t_total=10)
    sched.step()                                      # leave the lr = 0 start of the warm-up
    _step(model, cfg, task_cfg, opt, 1)
    want = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    osd = opt.state_dict()
    want_m = {i: st["exp_avg"].detach().cpu().clone() for i, st in osd["state"].items()}
    want_v = {i: st["exp_avg_sq"].detach().cpu().clone() for i, st in osd["state"].items()}
    ck = checkpoint.AsyncCheckpointer()
    checkpoint.save(str(tmp_path), None, 0, model, opt, sched, 1, None, True, 0.0, is_best=False, checkpointer=ck)
    _step(model, cfg, task_cfg, opt, 2)              # overwrites parameters and moments while the snapshot is in flight
    _step(model, cfg, task_cfg, opt, 3)
    ck.wait()
    assert len(ck._arenas) == 3                       # parameters, exp_avg, exp_avg_sq went as whole arenas
    sd = torch.load(tmp_path / "pytorch_model_0.bin")
    assert list(sd.keys()) == list(want.keys())
    moved = 0
    now = model.state_dict()
    for k in want:
        assert torch.equal(sd[k], want[k]), k
        moved += int(not torch.equal(now[k].cpu(), want[k]))
    assert moved > 10                                 # the later steps really changed the live parameters
    full = torch.load(tmp_path / "pytorch_ckpt_latest.tar", weights_only=False)
    for i in want_m:
        assert torch.equal(full["optimizer_state_dict"]["state"][i]["exp_avg"], want_m[i])
        assert torch.equal(full["optimizer_state_dict"]["state"][i]["exp_avg_sq"], want_v[i])
    # resume: a fresh model + optimiser continue from the snapshot exactly like a model loaded from `want`
    m2, _, _, _ = _model(seed=9)
    o2 = optim.AdamW(optim.grouped_parameters(m2, 4e-5, 1e-4), lr=4e-5)
    _, gs, ep, _, _ = checkpoint.resume(str(tmp_path / "pytorch_ckpt_latest.tar"), m2, o2, None, None)
    assert (gs, ep) == (1, 1)
    for k, v in m2.state_dict().items():
        assert torch.equal(v.cpu(), want[k]), k
    _step(m2, cfg, task_cfg, o2, 2)                   # the restored moments are used (dense update, no cold shortcut)
    st = o2.state_dict()["state"]
    assert all(int(s["step"]) == 2 for s in st.values())
    assert all(bool(torch.isfinite(p).all()) for p in m2.parameters())


def test_async_save_sft_twins(tmp_path):
    model, cfg, task_cfg, md = _model(masks=True)
    opt = optim.AdamW(optim.grouped_parameters(model, 4e-5, 1e-4), lr=4e-5)
    _step(model, cfg, task_cfg, opt, 1)
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    checkpoint.save_sft(str(tmp_path), None, 2, model, opt, None, 1, None, True, 0.0, is_best=True)
    _step(model, cfg, task_cfg, opt, 2)
    checkpoint.wait()
    masked = torch.load(tmp_path / "pytorch_model_2.bin")
    unmasked = torch.load(tmp_path / "pytorch_model_unmasked2.bin")
    assert not any(k.endswith("_orig") or k.endswith("_mask") for k in masked)
    assert masked.keys() == unmasked.keys()
    n_masked = 0
    for k in sd:
        if k.endswith("_orig"):
            base = k[:-len("_orig")]
            assert torch.equal(masked[base], sd[k] * sd[base + "_mask"]) and torch.equal(unmasked[base], sd[k])
            n_masked += 1
        elif not k.endswith("_mask"):
            assert torch.equal(masked[k], sd[k])
    assert n_masked >= 6 * LAYERS + 1                 # plus the aliased v_* entries of the shared modules
    assert (tmp_path / "pytorch_model_unmasked_best.bin").exists() and (tmp_path / "pytorch_model_best.bin").exists()
    # the folded file loads into an unmasked model of the same architecture (what eval_task.py does with it)
    plain, _, _, _ = _model(seed=11)
    plain.load_state_dict(masked, strict=True)


def test_evaluate_languages_on_device(tmp_path):
    model, cfg, task_cfg, _ = _model()
    model.eval()
    loaders, want = {}, {}
    for li, lang in enumerate(["en", "de", "zh"]):
        batches = []
        for bi in range(3):
            b = S.synth_batch(16 if bi < 2 else 5, num_regions=R, seed=100 * li + bi, vocab_size=VOCAB)      # ragged last batch
            batches.append(tuple(t.pin_memory() if torch.is_tensor(t) else t for t in S.batch_tuple(b)))
        loaders[lang] = batches
        hits = n = 0
        with torch.no_grad():
            for t in batches:
                d = tuple(x.to(DEV) for x in t)
                logits = model(d[3], d[0], d[1], "TASK15", d[6], d[5], d[2])[0]
                hits += int((logits.argmax(1) == d[4].argmax(1)).sum())
                n += d[0].shape[0]
        want[lang] = (n, hits / n)
    table = evaluation.evaluate_languages(cfg, task_cfg, torch.device(DEV), "TASK15", loaders, model, save_path=str(tmp_path))
    assert list(table) == ["en", "de", "zh"]
    for lang in table:
        assert table[lang]["n"] == want[lang][0] == 37
        assert abs(table[lang]["accuracy"] - want[lang][1]) < 1e-12
        res = json.load(open(tmp_path / (lang + "_result.json")))
        assert len(res) == 37 and set(res[0]) == {"questionId", "prediction"}
        ids = torch.cat([t[7] for t in loaders[lang]]).tolist()
        assert [r["questionId"] for r in res] == [str(int(i)) for i in ids]

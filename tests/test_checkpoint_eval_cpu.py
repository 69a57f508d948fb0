"""Host logic of the two 'next' rows built on top of the hot path, without a GPU:
f4 - checkpoint.save / save_sft / resume write what volta/train_utils.py:L351-416, L488-510 write (file names, dictionary
     layout, mask fold-back and the `unmasked` twin), asynchronously, from a snapshot taken at the time of the call;
f3 - evaluation.shard_languages / evaluate_languages: whole languages per rank, deterministic, gathered score table
     (world_size 2, gloo)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.utils.prune as prune

from clg_vqa_b200 import checkpoint, evaluation


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(6, 5)
        self.b = torch.nn.Linear(5, 3)

    def forward(self, x):
        return self.b(torch.relu(self.a(x)))


def _train_step(net, opt, seed):
    torch.manual_seed(seed)
    loss = net(torch.randn(4, 6)).pow(2).sum()
    opt.zero_grad()
    loss.backward()
    opt.step()


def test_save_is_a_snapshot_and_resume_round_trips(tmp_path):
    net = _Net()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-2)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    _train_step(net, opt, 0); sched.step()
    want = {k: v.clone() for k, v in net.state_dict().items()}
    want_m = opt.state_dict()["state"][0]["exp_avg"].clone()
    checkpoint.save(str(tmp_path), None, 3, net, opt, sched, 17, {"tb": 1}, True, 0.5, is_best=True)
    _train_step(net, opt, 1)                        # the step after the call must not leak into the files
    checkpoint.wait()
    for name in ("pytorch_model_3.bin", "pytorch_model_best.bin", "pytorch_ckpt_latest.tar"):
        assert (tmp_path / name).exists(), name
    assert not any(p.name.endswith(".tmp") for p in tmp_path.iterdir())
    sd = torch.load(tmp_path / "pytorch_model_3.bin")
    assert list(sd.keys()) == list(want.keys())
    for k in want:
        assert torch.equal(sd[k], want[k]), k
    ck = torch.load(tmp_path / "pytorch_ckpt_latest.tar", weights_only=False)
    assert set(ck.keys()) == {"model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "global_step", "epoch_id",
                              "tb_logger", "score"}
    assert ck["global_step"] == 17 and ck["epoch_id"] == 3 and ck["score"] == 0.5 and ck["tb_logger"] == {"tb": 1}
    assert torch.equal(ck["optimizer_state_dict"]["state"][0]["exp_avg"], want_m)
    # resume into fresh objects: same parameters, moments, schedule position, bookkeeping
    net2 = _Net()
    opt2 = torch.optim.AdamW(net2.parameters(), lr=1e-2)
    sched2 = torch.optim.lr_scheduler.LambdaLR(opt2, lambda s: 1.0 / (1 + s))
    it, gs, ep, tb, best = checkpoint.resume(str(tmp_path / "pytorch_ckpt_latest.tar"), net2, opt2, sched2, None)
    assert (it, gs, ep, tb, best) == (0, 17, 4, {"tb": 1}, 0.5)
    for k, v in net2.state_dict().items():
        assert torch.equal(v, want[k])
    assert sched2.last_epoch == 1
    # not the default GPU: nothing is written
    checkpoint.save(str(tmp_path / "none"), None, 0, net, opt, sched, 0, None, False, 0.0)
    checkpoint.wait()
    assert not (tmp_path / "none").exists()
    # a missing path leaves everything at its start value
    assert checkpoint.resume(str(tmp_path / "absent.tar"), net2, opt2, sched2, "tb") == (0, 0, 0, "tb", float("-inf"))


def test_save_sft_folds_masks_and_keeps_the_unmasked_twin(tmp_path):
    net = _Net()
    torch.manual_seed(3)
    mask = (torch.rand(5, 6) > 0.4).float()
    prune.CustomFromMask.apply(net.a, "weight", mask=mask)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-2)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0)
    _train_step(net, opt, 0)
    orig = net.a.weight_orig.detach().clone()
    checkpoint.save_sft(str(tmp_path), None, 1, net, opt, sched, 5, None, True, 0.25, is_best=True)
    _train_step(net, opt, 1)
    checkpoint.wait()
    masked = torch.load(tmp_path / "pytorch_model_1.bin")
    unmasked = torch.load(tmp_path / "pytorch_model_unmasked1.bin")
    assert set(masked.keys()) == set(unmasked.keys()) == {"a.weight", "a.bias", "b.weight", "b.bias"}
    assert torch.equal(masked["a.weight"], orig * mask) and torch.equal(unmasked["a.weight"], orig)
    assert torch.equal(masked["b.weight"], unmasked["b.weight"])
    for name in ("pytorch_model_best.bin", "pytorch_model_unmasked_best.bin"):
        assert (tmp_path / name).exists()
    ck = torch.load(tmp_path / "pytorch_ckpt_latest.tar", weights_only=False)
    assert torch.equal(ck["model_state_dict"]["a.weight"], orig * mask)        # the folded dict, as in the reference
    assert checkpoint.fold_state_dict(net.state_dict())[0].keys() == masked.keys()


def test_writer_errors_surface_on_wait(tmp_path):
    net = _Net()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-2)
    checkpoint.save(str(tmp_path / "does" / "not" / "exist"), None, 0, net, opt, None, 0, None, True, 0.0)
    try:
        checkpoint.wait()
    except (FileNotFoundError, RuntimeError, OSError):
        pass
    else:
        raise AssertionError("the failed write was swallowed")
    checkpoint.wait()                                # the error is reported once


def test_shard_languages_is_a_balanced_partition():
    langs = ["en", "de", "pt", "ru", "id", "bn", "ko", "zh"]
    for world in (1, 2, 3, 4, 8):
        parts = [evaluation.shard_languages(langs, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == sorted(langs)
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
        assert all(p == [l for l in langs if l in p] for p in parts)          # each rank keeps the caller's order
    sizes = {"en": 12, "de": 3, "pt": 3, "ru": 3, "id": 3}
    parts = [evaluation.shard_languages(list(sizes), r, 2, sizes) for r in range(2)]
    assert parts[0] == ["en"] and parts[1] == ["de", "pt", "ru", "id"]        # longest first onto the least loaded rank


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_step(config, task_cfg, device, task_id, batch, model, dataloader, criterion, results, others):
    for q in batch:
        results.append({"questionId": str(q), "prediction": q % 7})
    return 0.0, 0.0, len(batch), results, others


def _eval_worker(rank, world, port, save_path, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        langs = ["en", "de", "pt", "ru", "id"]
        loaders = {l: [list(range(100 * i + 10 * b, 100 * i + 10 * b + 10)) for b in range(3 + (i == 0) * 5)] for i, l in enumerate(langs)}
        table = evaluation.evaluate_languages(None, None, "cpu", "TASK15", loaders, None, save_path=save_path, step_fn=_fake_step)
        out[rank] = {l: (table[l]["n"], table[l]["rank"]) for l in table}
    finally:
        dist.destroy_process_group()


def test_evaluate_languages_world2_gloo(tmp_path):
    import json
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_eval_worker, args=(world, port, str(tmp_path), out), nprocs=world, join=True)
        out = dict(out)
    assert out[0] == out[1]                                                    # every rank holds the whole table
    assert out[0] == {"en": (80, 0), "de": (30, 1), "pt": (30, 1), "ru": (30, 0), "id": (30, 1)} or \
        sorted(n for n, _ in out[0].values()) == [30, 30, 30, 30, 80]
    assert {r for _, r in out[0].values()} == {0, 1}
    for lang in ("en", "de", "pt", "ru", "id"):
        res = json.load(open(tmp_path / (lang + "_result.json")))
        assert len(res) == out[0][lang][0] and set(res[0]) == {"questionId", "prediction"}
        assert json.load(open(tmp_path / (lang + "_others.json"))) == []

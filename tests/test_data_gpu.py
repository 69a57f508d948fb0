"""Host -> device prefetcher (clg_vqa_b200/data.py): batches arrive complete, in order, in static slots."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from clg_vqa_b200.data import DevicePrefetcher  # noqa: E402


def test_prefetcher_order_and_contents():
    host = [(torch.full((1 << 20,), float(i)).pin_memory(), torch.arange(i, i + 40).pin_memory()) for i in range(7)]
    seen = []
    for i, (a, b) in enumerate(DevicePrefetcher(iter(host), "cuda")):
        assert a.is_cuda and b.is_cuda
        # consume on the current stream with a kernel that takes a while, so refills must really wait for it
        s = (a * 2).sum() / (2 * a.numel())
        seen.append((float(s), b.clone()))
    assert len(seen) == 7
    for i, (s, b) in enumerate(seen):
        assert abs(s - i) < 1e-3
        assert torch.equal(b.cpu(), torch.arange(i, i + 40))


def test_prefetcher_empty_and_single():
    assert list(DevicePrefetcher(iter([]), "cuda")) == []
    out = list(DevicePrefetcher(iter([(torch.ones(8).pin_memory(),)]), "cuda"))
    assert len(out) == 1 and float(out[0][0].sum()) == 8.0


def test_label_prior_table_rows():
    """device-resident label-prior table: rows gathered by label equal the reference's per-batch dictionary look-ups"""
    from clg_vqa_b200.data import LabelPriorTable
    A = 1842
    g = torch.Generator().manual_seed(1)
    sem = {}
    for _ in range(5000):
        j, t = int(torch.randint(0, A, (1,), generator=g)), int(torch.randint(0, A, (1,), generator=g))
        sem[(j, t)] = float(torch.rand(1, generator=g)) * 1.2
    tab = LabelPriorTable.from_semantic_dict(sem, A, "cuda")
    labels = torch.randint(0, A, (64,), generator=g)
    rows = tab.rows(labels).cpu()
    for b in range(64):
        t = int(labels[b])
        for j in (0, 1, t, A - 1, int(torch.randint(0, A, (1,), generator=g))):
            want = 0.0 if j == t else sem.get((j, t), 0.0)
            assert abs(float(rows[b, j]) - want) < 1e-7


@pytest.mark.parametrize("family", ["uc2", "m3p"])
def test_bf16_feature_delivery_is_bit_identical(family):
    """region features delivered as bf16 (half the H2D bytes, no cast pass) are the GEMM operand as is - the
    rounding point does not move, so logits, loss and every gradient are bit-identical to fp32 delivery + on-device cast."""
    from clg_vqa_b200 import config as C, synthetic as S, task_utils as TU
    from clg_vqa_b200.encoders import BertForVLTasks, M3PForVLTasks
    if family == "uc2":
        cfg = C.uc2_base(2, vocab_size=2048)
        model = BertForVLTasks(cfg, C.gqa_task_cfg(36), ["TASK15"])
        b = S.synth_batch(6, seed=3, ragged_regions=True, vocab_size=2048)
    else:
        cfg = C.m3p_base(2, n_words=2048)
        model = M3PForVLTasks(cfg, C.gqa_task_cfg(100), ["TASK15"])
        b = S.synth_batch(6, num_regions=100, num_locs=5, seed=3, vocab_size=2048, normalize=True)
    tc = C.gqa_task_cfg(36 if family == "uc2" else 100)
    model.load_state_dict(S.synth_state_dict(model, seed=4))
    model.set_precision("bf16")
    model.to("cuda").eval()
    outs = []
    for dt in (torch.float32, torch.bfloat16):
        bb = dict(b)
        bb["features"] = b["features"].to(dt)
        model.zero_grad()
        loss, _ = TU.ForwardModelsTrain(cfg, tc, "cuda", "TASK15", S.batch_tuple(bb), model)
        loss.backward()
        outs.append((float(loss), model.engine.g_flat.clone()))
    assert outs[0][0] == outs[1][0]
    d = (outs[0][1] - outs[1][1]).abs().max()
    assert float(d) <= 1e-6 * float(outs[0][1].abs().max())        # split-K atomics order only

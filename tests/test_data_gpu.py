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

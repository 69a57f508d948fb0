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

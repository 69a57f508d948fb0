import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clg_vqa_b200 import synthetic as S
dev = torch.device("cuda", 0)
host = [tuple(t.pin_memory() for t in S.batch_tuple(S.synth_batch(256, num_regions=36, seed=i))) for i in range(2)]
slot = tuple(torch.empty(h.shape, dtype=h.dtype, device=dev) for h in host[0])
side = torch.cuda.Stream(device=dev)
x = torch.randn(8192, 8192, device=dev)
torch.cuda.synchronize()
for trial in range(3):
    # busy compute stream
    for _ in range(20): y = x @ x
    t0 = time.perf_counter()
    with torch.cuda.stream(side):
        for d, h in zip(slot, host[trial % 2]):
            d.copy_(h, non_blocking=True)
    t1 = time.perf_counter()
    print("copy_ issue under busy compute: %.3f ms" % ((t1 - t0) * 1e3))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(side):
        for d, h in zip(slot, host[trial % 2]):
            d.copy_(h, non_blocking=True)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("copy_ issue idle: %.3f ms ; total %.3f" % ((t1 - t0) * 1e3, (time.perf_counter() - t0) * 1e3))
    t0 = time.perf_counter()
    for d, h in zip(slot, host[trial % 2]):
        t2 = time.perf_counter(); d.copy_(h, non_blocking=True); print("   ", tuple(h.shape), "%.3f" % ((time.perf_counter() - t2) * 1e3))
    torch.cuda.synchronize()

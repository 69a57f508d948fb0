import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clg_vqa_b200 import ops
dev = "cuda"
torch.manual_seed(0)
for (M, N, K) in [(19456, 768, 128), (19456, 768, 768), (4096, 768, 128)]:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    res = torch.randn(M, N, device=dev)
    y = torch.full((M, N), float("nan"), device=dev)
    ops.gemm(a, w, y, residual=res)
    ref = a.float() @ w.float().t() + res
    bad = (y - ref).abs() > 1e-3
    print(M, N, K, "bad frac", float(bad.float().mean()), "nan", int(torch.isnan(y).sum()))
    if bad.any():
        bt = bad.view(M // 128, 128, N // 128, 128).any(3).any(1)       # per tile
        print(" bad tiles", int(bt.sum()), "of", bt.numel())
        idx = torch.nonzero(bt)[:12].tolist()
        print(" first bad tiles (m_blk, n_blk):", idx)
        mb, nb = idx[0]
        sub = bad[mb * 128:(mb + 1) * 128, nb * 128:(nb + 1) * 128]
        print(" within tile: rows bad", torch.nonzero(sub.any(1)).flatten().tolist()[:40])
        print(" within tile: col chunks(16) bad", torch.nonzero(sub.view(128, 8, 16).any(2).any(0)).flatten().tolist())
        # is the wrong value equal to lin + some other residual?
        lin = a.float() @ w.float().t()
        d = (y - lin)[mb * 128:(mb + 1) * 128, nb * 128:(nb + 1) * 128]
        r0, c0 = torch.nonzero(sub)[0].tolist()
        val = float(d[r0, c0])
        where = torch.nonzero((res - val).abs() < 1e-6)
        print(" y-lin at first bad elem", val, "true res", float(res[mb * 128 + r0, nb * 128 + c0]), "matches res at", where[:4].tolist(), "elem", (mb * 128 + r0, nb * 128 + c0))

"""Times the row kernels of the step at the cfg-2 shape (15 409 packed rows x 768): LayerNorm forward (fp32 + bf16 outputs,
statistics, dropout) and backward, and the bf16 column sum (QKV bias gradient).  `warm` = launches back to back on the same
buffers (the input is what the producing GEMM has just left in L2, as in the step); `cold` = a 256 MB write in between."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clg_vqa_b200 import _lib, ops
if os.environ.get("BENCH_LIB"):          # A/B against another build of the library
    _lib.LIB_PATH = os.path.abspath(os.environ["BENCH_LIB"])

dev = "cuda"
rows, N = int(os.environ.get("ROWS", "15409")), 768
x = torch.randn(rows, N, device=dev)
gam = torch.randn(N, device=dev); bet = torch.randn(N, device=dev)
y = torch.empty_like(x); yl = torch.empty(rows, N, device=dev, dtype=torch.bfloat16)
mean = torch.empty(rows, device=dev); rstd = torch.empty(rows, device=dev)
dy = torch.randn(rows, N, device=dev); dx = torch.empty_like(x); dxl = torch.empty_like(yl)
dg = torch.zeros(N, device=dev); db = torch.zeros(N, device=dev); dbias = torch.zeros(N, device=dev)
q = torch.randn(rows, 2304, device=dev).to(torch.bfloat16); qs = torch.zeros(2304, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
drop = ops.DropSite(0.1, 1234, 7)


def timed(fn, cold):
    ts = []
    for _ in range(30):
        if cold:
            flush.fill_(1.0)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


cases = {
    "ln_fwd": lambda: ops.layernorm_fwd(x, gam, bet, 1e-12, y=y, y_lowp=yl, mean=mean, rstd=rstd, dropout=drop),
    "ln_bwd": lambda: ops.layernorm_bwd(dy, x, gam, mean, rstd, dx=dx, dx_lowp=dxl, dgamma=dg, dbeta=db, dbias=dbias, dropout_out=drop),
    "colsum_bf16": lambda: ops.colsum(q, qs),
}
for name, fn in cases.items():
    fn(); torch.cuda.synchronize()
    print("%-12s rows=%d  warm %6.1f us   cold %6.1f us" % (name, rows, timed(fn, False), timed(fn, True)), flush=True)

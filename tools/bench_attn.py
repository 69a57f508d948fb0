import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clg_vqa_b200 import ops
from clg_vqa_b200.ops import DropSite
dev = "cuda"
B, S, H = int(os.environ.get("B", 256)), int(os.environ.get("S", 76)), 12
D = H * 64
qkv = torch.randn(B * S, 3 * D, device=dev).bfloat16()
kb = torch.zeros(B, S, device=dev)
ctx = torch.zeros(B * S, D, device=dev, dtype=torch.bfloat16); lse = torch.zeros(B, H, S, device=dev)
dctx = torch.randn(B * S, D, device=dev).bfloat16(); dqkv = torch.zeros_like(qkv)
site = DropSite(0.1, 5, 1 << 41)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2] * 1e3
for backend, name in ((2, "one CTA / head"), (0, "pipelined")):
    ops.set_attention_backend(backend)
    tf = timeit(lambda: ops.attention_fwd(qkv, kb, ctx, lse, B, S, H, 0.125, site))
    tb = timeit(lambda: ops.attention_bwd(qkv, kb, ctx, dctx, lse, dqkv, B, S, H, 0.125, site))
    print("%-16s fwd %7.1f us   bwd %7.1f us" % (name, tf, tb))
# packed rows (engine._forward with CLG_PACK): item b owns rows cu[b] .. cu[b+1], lengths like the synthetic GQA batch
g = torch.Generator(device="cpu").manual_seed(0)
lens = torch.randint(int(os.environ.get("SMIN", 44)), S + 1, (B,), generator=g)
cu = torch.zeros(B + 1, dtype=torch.int32); cu[1:] = torch.cumsum(lens, 0)
Mp = int(cu[-1]); cu = cu.to(dev)
qkv_p = qkv[:Mp].contiguous(); kb_p = torch.zeros(Mp, device=dev); ctx_p = torch.zeros(Mp, D, device=dev, dtype=torch.bfloat16)
dctx_p = dctx[:Mp].contiguous(); dqkv_p = torch.zeros_like(qkv_p)
ops.set_attention_backend(0)
tf = timeit(lambda: ops.attention_fwd(qkv_p, kb_p, ctx_p, lse, B, S, H, 0.125, site, cu_rows=cu, total_rows=Mp))
tb = timeit(lambda: ops.attention_bwd(qkv_p, kb_p, ctx_p, dctx_p, lse, dqkv_p, B, S, H, 0.125, site, cu_rows=cu, total_rows=Mp))
print("%-16s fwd %7.1f us   bwd %7.1f us   (%d packed rows, mean length %.1f)" % ("pipelined packed", tf, tb, Mp, Mp / B))

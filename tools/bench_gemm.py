"""Micro-benchmark of the tcgen05 GEMM on the shapes of the step (and a square reference shape)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clg_vqa_b200 import ops
dev = "cuda"
from clg_vqa_b200.ops import DropSite
def run(M, N, K, a_k=True, b_k=True, cdt=torch.bfloat16, res=False, bias=True, iters=20, name="", act=ops.ACT_NONE, aux=None, drop=False):
    a = torch.randn((M, K) if a_k else (K, M), device=dev).bfloat16()
    b = torch.randn((N, K) if b_k else (K, N), device=dev).bfloat16()
    c = torch.empty(M, N, device=dev, dtype=cdt)
    bi = torch.randn(N, device=dev) if bias else None
    r = torch.randn(M, N, device=dev) if res else None
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    kw = dict(a_kmajor=a_k, b_kmajor=b_k, bias=bi, residual=r, act=act)
    if aux == "out": kw["aux_out"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    if aux == "in": kw["aux_in"] = torch.randn(M, N, device=dev).bfloat16()
    if drop: kw["dropout"] = DropSite(0.1, seed=1, offset=1 << 20)
    for _ in range(3): ops.gemm(a, b, c, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(a, b, c, **kw); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); t = ts[len(ts)//2]
    print("%-28s M=%6d N=%5d K=%6d  %8.1f us  %7.1f TFLOP/s" % (name, M, N, K, t*1e3, 2.0*M*N*K/t/1e9))
    # cuBLAS reference (library) for context
    if a_k and b_k:
        for _ in range(3): torch.matmul(a, b.t())
        torch.cuda.synchronize(); ts=[]
        for _ in range(iters):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b.t()); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ts.sort(); t = ts[len(ts)//2]
        print("%-28s   cuBLAS (no epilogue)      %8.1f us  %7.1f TFLOP/s" % ("", t*1e3, 2.0*M*N*K/t/1e9))
ONLY = os.environ.get("ONLY", "")
_run = run
def run(*a, **k):
    if ONLY and ONLY not in k.get("name", ""): return
    _run(*a, **k)
run(8192, 8192, 8192, name="square", bias=False)
run(19456, 2304, 768, name="QKV fwd")
run(19456, 3072, 768, name="FFN1 (plain epi)")
run(19456, 3072, 768, name="FFN1 (gelu pair)", act=ops.ACT_GELU_GRAD, aux="out")
run(19456, 768, 3072, cdt=torch.float32, res=True, drop=True, name="FFN2 fwd (+drop+res fp32)")
run(19456, 3072, 768, b_k=False, bias=False, name="FFN2 dgrad (mul aux)", act=ops.ACT_MUL_AUX, aux="in")
run(19456, 768, 768, b_k=False, bias=False, name="out-proj dgrad (plain)")
run(19456, 768, 3072, cdt=torch.float32, res=True, name="FFN2 fwd (+res fp32)")
run(19456, 768, 3072, name="FFN2 shape, plain bf16")
run(19456, 768, 3072, cdt=torch.float32, bias=False, name="FFN2 shape, generic f32")
run(19456, 768, 768, cdt=torch.float32, res=True, name="out-proj (+res fp32)")
run(2304, 768, 19456, a_k=False, b_k=False, cdt=torch.float32, bias=False, name="QKV wgrad")
run(3072, 768, 19456, a_k=False, b_k=False, cdt=torch.float32, bias=False, name="FFN1 wgrad")
run(768, 3072, 19456, a_k=False, b_k=False, cdt=torch.float32, bias=False, name="FFN2 wgrad")
run(768, 768, 19456, a_k=False, b_k=False, cdt=torch.float32, bias=False, name="out-proj wgrad")
run(2304, 768, 15409, a_k=False, b_k=False, cdt=torch.float32, bias=False, name="QKV wgrad K=15409")
run(3072, 768, 15409, a_k=False, b_k=False, cdt=torch.float32, bias=False, name="FFN1 wgrad K=15409")
run(768, 768, 15409, a_k=False, b_k=False, cdt=torch.float32, bias=False, name="out-proj wgrad K=15409")
run(19456, 768, 2304, b_k=False, cdt=torch.float32, res=True, bias=False, name="QKV dgrad (+res)")
# answer head (M = batch): latency-bound, few output tiles
run(256, 1024, 768, cdt=torch.float32, name="head pooler fwd")
run(256, 1536, 1024, cdt=torch.float32, name="head clf0 fwd")
run(256, 1842, 1536, cdt=torch.float32, name="head logits fwd")
run(256, 1536, 1848, b_k=False, cdt=torch.float32, bias=False, name="head logits dgrad")
run(256, 1024, 1536, b_k=False, cdt=torch.float32, bias=False, name="head clf0 dgrad")
run(256, 768, 1024, b_k=False, cdt=torch.float32, bias=False, name="head pooler dgrad")
run(1842, 1536, 256, a_k=False, b_k=False, cdt=torch.float32, bias=False, name="head logits wgrad")
run(1536, 1024, 256, a_k=False, b_k=False, cdt=torch.float32, bias=False, name="head clf0 wgrad")
run(1024, 768, 256, a_k=False, b_k=False, cdt=torch.float32, bias=False, name="head pooler wgrad")
run(9216, 768, 2048, name="head image embedding fwd")
run(768, 2048, 9216, a_k=False, b_k=False, cdt=torch.float32, bias=False, name="head image embedding wgrad")

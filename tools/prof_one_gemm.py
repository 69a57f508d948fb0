"""one GEMM shape for ncu:  SHAPE=ffn2|outproj|qkv|ffn1|ffn1g|wgrad|dgrad python tools/prof_one_gemm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clg_vqa_b200 import ops
from clg_vqa_b200.ops import DropSite
dev = "cuda"
shape = os.environ.get("SHAPE", "ffn2")
M = 19456
def T(*s): return torch.randn(*s, device=dev)
if shape == "ffn2":
    a, b, c = T(M, 3072).bfloat16(), T(768, 3072).bfloat16(), torch.empty(M, 768, device=dev)
    kw = dict(bias=T(768), residual=T(M, 768), dropout=DropSite(0.1, seed=1, offset=1 << 20))
elif shape == "outproj":
    a, b, c = T(M, 768).bfloat16(), T(768, 768).bfloat16(), torch.empty(M, 768, device=dev)
    kw = dict(bias=T(768), residual=T(M, 768), dropout=DropSite(0.1, seed=1, offset=1 << 20))
elif shape == "qkv":
    a, b, c = T(M, 768).bfloat16(), T(2304, 768).bfloat16(), torch.empty(M, 2304, device=dev, dtype=torch.bfloat16)
    kw = dict(bias=T(2304))
elif shape == "ffn1g":
    a, b, c = T(M, 768).bfloat16(), T(3072, 768).bfloat16(), torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
    kw = dict(bias=T(3072), act=ops.ACT_GELU_GRAD, aux_out=torch.empty(M, 3072, device=dev, dtype=torch.bfloat16))
elif shape == "wgrad":
    a, b, c = T(M, 3072).bfloat16(), T(M, 768).bfloat16(), torch.zeros(3072, 768, device=dev)
    kw = dict(a_kmajor=False, b_kmajor=False, accumulate=True)
elif shape == "dgrad":
    a, b, c = T(M, 3072).bfloat16(), T(3072, 768).bfloat16(), torch.empty(M, 768, device=dev)
    kw = dict(b_kmajor=False, residual=T(M, 768))
for _ in range(3):
    ops.gemm(a, b, c, **kw)
torch.cuda.synchronize()
print("done", shape)

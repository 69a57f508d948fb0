"""Times the fused AdamW on an arena of the UC2 size (282 M parameters, 68 % of them in never-touched blocks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clg_vqa_b200 import ops, _lib as L
dev = "cuda"
n = 282 * 1024 * 1024 // L.CHUNK * L.CHUNK
nh = int(n * 0.32)
p = torch.randn(n, device=dev) * 0.02; m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
g = torch.zeros(n, device=dev); g[:nh] = torch.randn(nh, device=dev) * 1e-3
g0 = g.clone()
shadow = torch.zeros(n, device=dev, dtype=torch.bfloat16)
hot = torch.zeros(n // 256, device=dev, dtype=torch.uint8)
ct = torch.zeros(n // L.CHUNK, dtype=torch.int32, device=dev)
metas = (L.TensorMeta * 1)(); metas[0] = L.TensorMeta(4e-5, 1e-4, 0, -1, 0, n)
tens = torch.frombuffer(bytearray(bytes(metas)), dtype=torch.uint8).to(dev)
partials = torch.zeros(4096, device=dev); ss = torch.zeros(1, device=dev)
ops.sumsq(g, partials, ss, hot=hot)
for mode in ("plain", "zero", "plain+fill", "plain", "zero"):
    ts = []
    for it in range(8):
        g.copy_(g0)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.adamw_step(p, g, m, v, ct, tens, sumsq=ss, max_norm=1.0, eps=1e-6, bias_c1=0.1, bias_c2=0.001, shadow=shadow, hot=hot,
                       zero_grad=(mode == "zero"))
        if mode == "plain+fill": g.zero_()
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    print("%-12s %8.1f us" % (mode, ts[len(ts) // 2] * 1e3))

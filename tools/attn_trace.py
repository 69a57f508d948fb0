"""Timeline of the pipelined attention forward for 128 < S <= 160 (CTA 0, first 16 units) from the kernel's clock64 stamps
(clg_debug_attn_trace): where the per-item latency chain spends its cycles.   python tools/attn_trace.py [B] [S]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clg_vqa_b200 import _lib as L, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 140       # the instrumented instantiation is the S in (128, 160] forward
H, D = 12, 768
dev = "cuda"
torch.manual_seed(0)
qkv = (torch.randn(B * S, 3 * D, device=dev) * 0.5).to(torch.bfloat16)
kb = torch.zeros(B, S, device=dev)
ctx = torch.zeros(B * S, D, device=dev, dtype=torch.bfloat16)
lse = torch.zeros(B, H, S, device=dev)
for _ in range(3):
    ops.attention_fwd(qkv, kb, ctx, lse, B, S, H, 0.125)
torch.cuda.synchronize()
trace = torch.zeros(5 * 16 * 8, dtype=torch.int64, device=dev)
lib = L.load()
lib.clg_debug_attn_trace.argtypes = [ctypes.c_void_p]
lib.clg_debug_attn_trace(trace.data_ptr())
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
ops.attention_fwd(qkv, kb, ctx, lse, B, S, H, 0.125)
e1.record()
torch.cuda.synchronize()
lib.clg_debug_attn_trace(None)
print("kernel %.1f us, items per CTA %.1f" % (e0.elapsed_time(e1) * 1e3, B * H / 148))
t = trace.cpu().view(5, 16, 8)
t0 = int(t[t > 0].min())
names = {0: ["tma_issue"], 1: ["S_issued", "PV_issued"],
         2: (["top", "S_ready", "tmem_loaded", "-", "-", "P_written", "O_ready", "copied"] if S <= 128 else
             ["top", "S_ready", "max_done", "exp_done", "P_written", "O_ready", "copied"])}
for it in range(16):
    row = ["it%2d" % it]
    for role in range(5):
        ev = names[min(role, 2)]
        vals = [int(t[role, it, k]) for k in range(len(ev))]
        if not any(vals):
            continue
        row.append(("P" if role == 0 else "M" if role == 1 else "G%d" % (role - 2)) + " " +
                   " ".join("%s=%d" % (n, v - t0) for n, v in zip(ev, vals) if v))
    print(" | ".join(row))

"""Diagnostic: per-stage error of the bf16 CUDA path against the bf16-operand oracle."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from clg_vqa_b200 import config as C, synthetic as S
from clg_vqa_b200.encoders import BertForVLTasks
from oracle import uc2_oracle

L = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = C.uc2_base(L, vocab_size=4096)
tc = C.gqa_task_cfg()
model = BertForVLTasks(cfg, tc, ["TASK15"])
sd = S.synth_state_dict(model, seed=0)
model.load_state_dict(sd)
b = S.synth_batch(4, seed=1234, ragged_regions=True, vocab_size=4096)
def rel(a, r):
    a = a.double().cpu(); r = r.double().cpu()
    return float((a - r).abs().max() / r.abs().max())
for mode, mm in (("fp32", None), ("bf16", torch.bfloat16)):
    ref, hid = uc2_oracle.uc2_forward(sd, cfg, b["question"], b["features"], b["spatials"], b["segment_ids"], b["input_mask"], b["image_mask"], mm_dtype=mm, return_hidden=True)
    model.set_precision(mode); model.to("cuda").eval()
    with torch.no_grad():
        out = model(b["question"].cuda(), b["features"].cuda(), b["spatials"].cuda(), "TASK15", b["segment_ids"].cuda(), b["input_mask"].cuda(), b["image_mask"].cuda())[0]
    eng = model.engine
    ws = list(eng._ws.values())[0]
    B, S_ = 4, 76
    print(mode, "x0", rel(ws["x0"].view(B, S_, -1), hid[0]))
    lw = ws["layers"][0]
    print(mode, " (eval keeps one slot: last layer) x2", rel(lw["x2"].view(B, S_, -1), hid[-1]))
    print(mode, "logits", rel(out, ref))
    # first-layer pieces via oracle internals

# ---- head sensitivity: feed OUR final hidden state through the oracle head
import torch.nn.functional as F
def head(x0rows, mm):
    p = torch.relu(uc2_oracle.linear(x0rows, sd["bert.t_pooler.dense.weight"], sd["bert.t_pooler.dense.bias"], mm))
    c = "clfs_dict.TASK15.logit_fc."
    hc = uc2_oracle.gelu(uc2_oracle.linear(p, sd[c+"0.weight"], sd[c+"0.bias"], mm))
    hl = uc2_oracle.layer_norm(hc, sd[c+"2.weight"], sd[c+"2.bias"], cfg.layer_norm_eps)
    return p, hc, hl, uc2_oracle.linear(hl, sd[c+"3.weight"], sd[c+"3.bias"], mm)
mine_x = lw["x2"].view(B, S_, -1)[:, 0].float().cpu()
ref_x = hid[-1][:, 0]
print("token0 hidden err", rel(mine_x, ref_x))
pm, hcm, hlm, lgm = head(mine_x, torch.bfloat16)
pr, hcr, hlr, lgr = head(ref_x, torch.bfloat16)
print("oracle head on our hidden vs oracle: pooled %g hc %g hln %g logits %g" % (rel(pm, pr), rel(hcm, hcr), rel(hlm, hlr), rel(lgm, lgr)))
hw = ws["heads"]["TASK15"]
print("our head vs oracle head on OUR hidden: pact %g hc %g hln %g logits %g" % (rel(ws["pact"], pm), rel(hw["hc"], hcm), rel(hw["hln"].float(), hlm.bfloat16().float()), rel(out, lgm)))
print("hc std/mean", float(hcr.std()), float(hcr.mean()), "logit max", float(lgr.abs().max()))

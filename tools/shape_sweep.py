"""A few non-benchmark shapes through the full step (packing, tile tails, SIMT fallbacks): finite loss, parameters move."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clg_vqa_b200 import config as C, optim, synthetic as S, task_utils as TU
from clg_vqa_b200.encoders import BertForVLTasks
dev = torch.device("cuda", 0)
for (B, R, L) in [(1, 36, 2), (2, 36, 2), (3, 10, 2), (37, 36, 3), (64, 60, 2), (200, 36, 2), (16, 100, 2)]:
    cfg = C.uc2_base(L, vocab_size=4096); tc = C.gqa_task_cfg(R)
    torch.manual_seed(0)
    model = BertForVLTasks(cfg, tc, ["TASK15"]); model.set_precision("bf16"); model.to(dev); model.train()
    opt = optim.AdamW(optim.grouped_parameters(model, 4e-5, 1e-4), lr=4e-5)
    w0 = model.bert.t_pooler.dense.weight.detach().clone()
    for it in range(2):
        b = S.synth_batch(B, num_regions=R, seed=it, vocab_size=4096, ragged_regions=True)
        loss, score = TU.ForwardModelsTrain(cfg, tc, dev, "TASK15", tuple(t.to(dev) for t in S.batch_tuple(b)), model)
        loss.backward(); optim.clip_grad_norm_(model.parameters(), 1.0); opt.step(); model.zero_grad()
    torch.cuda.synchronize()
    ok = bool(torch.isfinite(loss)) and not torch.equal(w0, model.bert.t_pooler.dense.weight.detach())
    fin = all(bool(torch.isfinite(p).all()) for p in model.parameters())
    print("B=%3d R=%3d L=%d loss %.3f ok=%s finite_params=%s" % (B, R, L, float(loss), ok, fin), flush=True)
    assert ok and fin
    model.eval()
    with torch.no_grad():
        out = model(*[t.to(dev) for t in (b["question"], b["features"], b["spatials"])], "TASK15", b["segment_ids"].to(dev),
                    b["input_mask"].to(dev), b["image_mask"].to(dev))[0]
    assert bool(torch.isfinite(out).all())
print("shape sweep ok")

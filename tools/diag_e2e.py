import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clg_vqa_b200 import config as C, optim, synthetic as S, task_utils as TU
from clg_vqa_b200.encoders import BertForVLTasks
from clg_vqa_b200.data import DevicePrefetcher
dev = torch.device("cuda", 0)
cfg = C.uc2_base(12); tc = C.gqa_task_cfg(36)
torch.manual_seed(0)
model = BertForVLTasks(cfg, tc, ["TASK15"]); model.set_precision("bf16"); model.to(dev); model.train()
opt = optim.AdamW(optim.grouped_parameters(model, 4e-5, 1e-4), lr=4e-5)
host = [tuple(t.pin_memory() for t in S.batch_tuple(S.synth_batch(256, num_regions=36, seed=i))) for i in range(2)]
def step(batch):
    loss, score = TU.ForwardModelsTrain(cfg, tc, dev, "TASK15", batch, model)
    loss.backward(); optim.clip_grad_norm_(model.parameters(), 1.0); opt.step(); model.zero_grad()
    return loss
res = [tuple(t.to(dev) for t in hb) for hb in host]
for i in range(3): step(res[i % 2])
torch.cuda.synchronize()
def run(name, gen, n=6):
    torch.cuda.synchronize(); t0 = time.perf_counter(); ts = []
    for b in gen:
        t1 = time.perf_counter(); l = float(step(b)); ts.append((time.perf_counter() - t1) * 1e3)
    torch.cuda.synchronize()
    print("%-28s total %.1f ms/step ; per-step(step+sync) %s" % (name, (time.perf_counter() - t0) * 1e3 / len(ts), ["%.1f" % x for x in ts]))
run("resident", (res[i % 2] for i in range(6)))
run("in-stream H2D", (host[i % 2] for i in range(6)))
run("prefetcher", DevicePrefetcher((host[i % 2] for i in range(6)), dev))
# copy alone
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(6):
    d = tuple(t.to(dev, non_blocking=True) for t in host[i % 2])
torch.cuda.synchronize(); print("copy alone %.2f ms/batch" % ((time.perf_counter() - t0) * 1e3 / 6))
print([ (tuple(t.shape), t.dtype, t.is_pinned()) for t in host[0]])

"""One profiled fine-tune step (after warm-up) bracketed by cudaProfilerStart/Stop, for
   ncu --profile-from-start off ...   (launch list / --set full captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clg_vqa_b200 import config as C, optim, synthetic as S, task_utils as TU
from clg_vqa_b200.encoders import BertForVLTasks

B = int(os.environ.get("B", "256")); R = int(os.environ.get("R", "36")); LAYERS = int(os.environ.get("LAYERS", "12"))
dev = torch.device("cuda", 0)
cfg = C.uc2_base(LAYERS); tc = C.gqa_task_cfg(R)
torch.manual_seed(0)
model = BertForVLTasks(cfg, tc, ["TASK15"]); model.set_precision(os.environ.get("PREC", "bf16")); model.to(dev); model.train()
opt = optim.AdamW(optim.grouped_parameters(model, 4e-5, 1e-4), lr=4e-5)
batch = tuple(t.to(dev) for t in S.batch_tuple(S.synth_batch(B, num_regions=R, seed=1)))

def step():
    loss, score = TU.ForwardModelsTrain(cfg, tc, dev, "TASK15", batch, model)
    loss.backward()
    optim.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
    model.zero_grad()

for _ in range(int(os.environ.get("WARM", "2"))):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled step done")

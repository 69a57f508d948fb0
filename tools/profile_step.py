"""One profiled step (after warm-up) of a bench.py configuration bracketed by cudaProfilerStart/Stop, for
   ncu --profile-from-start off ...   (launch list / --set full captures).
   env: CONFIG = uc2_train (default) | m3p_train | uc2_sft41 | uc2_sft70 | eval_uc2 | eval_m3p;  B, R, LAYERS, PREC, WARM"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from clg_vqa_b200 import optim, synthetic as S, task_utils as TU

spec = dict(bench.CONFIGS[os.environ.get("CONFIG", "uc2_train")])
B = int(os.environ.get("B", str(spec["B"])))
if "R" in os.environ:
    spec["R"] = int(os.environ["R"])
LAYERS = int(os.environ.get("LAYERS", "12"))
dev = torch.device("cuda", 0)
cfg, tc, model = bench.build_model(spec, LAYERS, dev, os.environ.get("PREC", "bf16"))
train = spec["mode"] == "train"
model.train() if train else model.eval()
opt = optim.AdamW(optim.grouped_parameters(model, 4e-5, 1e-4, sft=spec["masks"] is not None), lr=4e-5) if train else None
tup = S.batch_tuple(bench.host_batch(spec, B, 1))
batch = tuple(t.to(dev) for t in (tup if train else tup[:8]))


class _DS:
    label2ans = None


class _DL:
    dataset = _DS()


def step():
    if train:
        loss, score = TU.ForwardModelsTrain(cfg, tc, dev, "TASK15", batch, model)
        loss.backward()
        optim.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        model.zero_grad()
    else:
        TU.EvaluatingModel(cfg, tc, dev, "TASK15", batch, model, _DL(), None, [], [])


for _ in range(int(os.environ.get("WARM", "2"))):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled step done")

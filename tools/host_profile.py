"""cProfile of the host side of the training step at a small per-GPU batch (where the host, not the GPU, bounds the step):
   python tools/host_profile.py [B]"""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from clg_vqa_b200 import optim, synthetic as S, task_utils as TU

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
spec = dict(bench.CONFIGS["uc2_train"])
dev = torch.device("cuda", 0)
cfg, tc, model = bench.build_model(spec, 12, dev, "bf16")
model.train()
opt = optim.AdamW(optim.grouped_parameters(model, 4e-5, 1e-4), lr=4e-5)
sched = optim.WarmupLinearSchedule(opt, warmup_steps=100, t_total=10000)
batch = tuple(t.to(dev) for t in S.batch_tuple(bench.host_batch(spec, B, 1)))


def step():
    loss, _ = TU.ForwardModelsTrain(cfg, tc, dev, "TASK15", batch, model)
    loss.backward()
    optim.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
    sched.step()
    model.zero_grad()


for _ in range(5):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)

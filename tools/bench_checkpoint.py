"""Row f4: what a checkpoint costs the training loop.  UC2-base 12L, B = 256: (a) the reference's way - torch.save of the
device state dicts inside the loop (train_utils.save); (b) checkpoint.save - snapshot + background writer - measured as
the stall of the call itself and the slowdown of the steps that run while the files are written."""
import os, sys, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clg_vqa_b200 import checkpoint, config as C, optim, synthetic as S, task_utils as TU
from clg_vqa_b200.encoders import BertForVLTasks

dev = torch.device("cuda", 0)
cfg = C.uc2_base(12); tc = C.gqa_task_cfg(36)
torch.manual_seed(0)
model = BertForVLTasks(cfg, tc, ["TASK15"]); model.set_precision("bf16"); model.to(dev); model.train()
opt = optim.AdamW(optim.grouped_parameters(model, 4e-5, 1e-4), lr=4e-5)
sched = optim.WarmupLinearSchedule(opt, warmup_steps=10, t_total=1000)
batch = tuple(t.to(dev) for t in S.batch_tuple(S.synth_batch(256, num_regions=36, seed=1)))

def step():
    loss, _ = TU.ForwardModelsTrain(cfg, tc, dev, "TASK15", batch, model)
    loss.backward(); optim.clip_grad_norm_(model.parameters(), 1.0); opt.step(); sched.step(); model.zero_grad()

def steps(n):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): step()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

for _ in range(5): step()
base = steps(20)
out = tempfile.mkdtemp(prefix="clg_ckpt_")
try:
    # (a) reference: synchronous torch.save of device tensors (train_utils.py:L351-372)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    torch.save(model.state_dict(), os.path.join(out, "ref_model.bin"))
    torch.save({"model_state_dict": model.state_dict(), "optimizer_state_dict": opt.state_dict(),
                "scheduler_state_dict": sched.state_dict(), "global_step": 1, "epoch_id": 0, "tb_logger": None, "score": 0.0},
               os.path.join(out, "ref_ckpt.tar"))
    t_ref = (time.perf_counter() - t0) * 1e3
    # (b) asynchronous; the first call also allocates the snapshot and pinned buffers
    for label in ("first call (allocates 3.4 GB snapshot + pinned buffers)", "steady state"):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        checkpoint.save(out, None, 0, model, opt, sched, 1, None, True, 0.0)
        t_call = (time.perf_counter() - t0) * 1e3
        during = steps(20)
        t0 = time.perf_counter(); checkpoint.wait(); t_wait = (time.perf_counter() - t0) * 1e3
        print("async %-58s call %7.1f ms   20 steps meanwhile %6.2f ms/step (plain %6.2f)   writer still busy after them %7.1f ms"
              % (label, t_call, during, base, t_wait))
    size = sum(os.path.getsize(os.path.join(out, f)) for f in ("pytorch_model_0.bin", "pytorch_ckpt_latest.tar")) / 1e9
    print("reference-style synchronous save: %.1f ms stall (%.2f GB written per checkpoint)" % (t_ref, size))
finally:
    shutil.rmtree(out, ignore_errors=True)

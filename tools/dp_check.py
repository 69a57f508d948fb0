"""torchrun --nproc-per-node N tools/dp_check.py : a few data-parallel fine-tune steps of UC2 and M3P with rank-specific
batches; after every step all ranks must hold bit-identical parameters, and the N-rank result must equal a single-rank
run on the concatenated batch up to summation order (checked on rank 0 for UC2 with dropout off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from clg_vqa_b200 import config as C, optim, parallel, synthetic as S, task_utils as TU
from clg_vqa_b200.encoders import BertForVLTasks, M3PForVLTasks

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr_ = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr_); dev = torch.device("cuda", lr_)
dist.init_process_group("nccl", device_id=dev)

def run(family, layers, B, R, steps, train_mode):
    torch.manual_seed(0)
    if family == "uc2":
        cfg = C.uc2_base(layers, vocab_size=4096); tc = C.gqa_task_cfg(R); model = BertForVLTasks(cfg, tc, ["TASK15"]); nl = 7
    else:
        cfg = C.m3p_base(layers, n_words=4096); tc = C.gqa_task_cfg(R); model = M3PForVLTasks(cfg, tc, ["TASK15"]); nl = 5
    model.load_state_dict(S.synth_state_dict(model, seed=1)); model.set_precision("bf16"); model.to(dev)
    model.train() if train_mode else model.eval()
    opt = optim.AdamW(optim.grouped_parameters(model, 4e-5, 1e-4), lr=4e-5)
    parallel.attach(model, opt, world)
    for it in range(steps):
        b = S.synth_batch(B, num_regions=R, num_locs=nl, seed=100 + 7 * it + rank, vocab_size=4096, normalize=(family == "m3p"),
                          ragged_regions=(family == "uc2"))
        loss, _ = TU.ForwardModelsTrain(cfg, tc, dev, "TASK15", tuple(t.to(dev) for t in S.batch_tuple(b)), model)
        loss.backward(); optim.clip_grad_norm_(model.parameters(), 1.0); opt.step(); model.zero_grad()
        flat = model.engine.p_flat
        chk = torch.stack([flat.double().sum(), flat.double().abs().sum(), flat[::997].double().pow(2).sum()])
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        same = all(torch.equal(allc[0], c) for c in allc)
        if rank == 0:
            print("%s step %d loss %.4f params identical across %d ranks: %s" % (family, it, float(loss), world, same), flush=True)
        assert same and torch.isfinite(loss)
    return model

m = run("uc2", 2, 16, 36, 3, True)
m = run("m3p", 2, 8, 100, 2, True)
if rank == 0:
    print("dp_check ok")
dist.destroy_process_group()

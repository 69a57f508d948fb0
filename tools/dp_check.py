"""Data-parallel equivalence on real GPUs (apex DistributedDataParallel semantics,
apex/parallel/distributed.py:L491-510: gradients are summed over ranks and divided by the world size).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dp_check.py

Run by tests/test_dp_gpu.py (skipped below 2 GPUs).  For UC2 and M3P, bf16, through the engine + parallel.attach:
  1. N ranks == 1 rank on the concatenated batch (dropout off): after the exchange every rank's gradient arena x 1/N equals
     the gradient of a single-process model fed cat(batch_0 .. batch_{N-1}) up to fp32 summation order (the per-sample
     arithmetic is identical; only the token sums of the weight gradients are split differently), the clipped norm
     agrees, and so do the loss and the gradients of the following step (i.e. the parameter update was the same).
  2. Replicas stay bit-identical in training mode (dropout on, independent masks per rank) after every step.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from clg_vqa_b200 import config as C, optim, parallel, synthetic as S, task_utils as TU
from clg_vqa_b200.encoders import BertForVLTasks, M3PForVLTasks

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def build(family, layers, R):
    if family == "uc2":
        cfg = C.uc2_base(layers, vocab_size=4096)
        model = BertForVLTasks(cfg, C.gqa_task_cfg(R), ["TASK15"])
    else:
        cfg = C.m3p_base(layers, n_words=4096)
        model = M3PForVLTasks(cfg, C.gqa_task_cfg(R), ["TASK15"])
    model.load_state_dict(S.synth_state_dict(model, seed=1))
    model.set_precision("bf16")
    model.to(dev)
    opt = optim.AdamW(optim.grouped_parameters(model, 4e-5, 1e-4), lr=4e-5)
    return cfg, C.gqa_task_cfg(R), model, opt


def batch(family, B, R, seed):
    return S.synth_batch(B, num_regions=R, num_locs=7 if family == "uc2" else 5, seed=seed, vocab_size=4096,
                         normalize=(family == "m3p"), ragged_regions=(family == "uc2"))


def cat_batches(bs):
    return {k: torch.cat([b[k] for b in bs], dim=0) for k in bs[0]}


def to_dev(b):
    return tuple(t.to(dev) for t in S.batch_tuple(b))


def slot_errors(eng_a, scale_a, eng_b):
    """worst max-norm relative difference per arena slot (floored at 1e-3 of the largest gradient entry)"""
    ga, gb = eng_a.g_flat * scale_a, eng_b.g_flat
    gmax = float(gb.abs().max())
    worst, where = 0.0, None
    for s in eng_b.slots:
        a = ga[s.start:s.start + s.numel]
        b = gb[s.start:s.start + s.numel]
        e = float((a - b).abs().max()) / max(float(b.abs().max()), 1e-3 * gmax)
        if e > worst:
            worst, where = e, s.names[0]
    return worst, where


def n_vs_1(family, layers, B, R):
    cfg, tc, dp_model, dp_opt = build(family, layers, R)
    _, _, solo, solo_opt = build(family, layers, R)
    dp_model.eval()
    solo.eval()                                   # dropout off: value parity
    parallel.attach(dp_model, dp_opt, world)
    for it in range(2):
        mine = batch(family, B, R, 100 + 10 * it + rank)
        everyone = cat_batches([batch(family, B, R, 100 + 10 * it + r) for r in range(world)])
        loss, _ = TU.ForwardModelsTrain(cfg, tc, dev, "TASK15", to_dev(mine), dp_model)
        loss.backward()
        sloss, _ = TU.ForwardModelsTrain(cfg, tc, dev, "TASK15", to_dev(everyone), solo)
        sloss.backward()
        lsum = loss.detach().clone()
        dist.all_reduce(lsum)
        lmean = float(lsum) / world                # mean-reduced loss, equal shard sizes: the mean of the rank losses
        e_loss = abs(lmean - float(sloss)) / float(sloss)
        e_grad, where = slot_errors(dp_model.engine, 1.0 / world, solo.engine)
        n_dp = float(optim.clip_grad_norm_(dp_model.parameters(), 1.0))
        n_solo = float(optim.clip_grad_norm_(solo.parameters(), 1.0))
        e_norm = abs(n_dp - n_solo) / n_solo
        dp_opt.step()
        solo_opt.step()
        dp_model.zero_grad()
        solo.zero_grad()
        if rank == 0:
            print("%s N=%d vs 1, step %d: loss %.3e  gradients %.3e (%s)  clip norm %.3e" %
                  (family, world, it, e_loss, e_grad, where, e_norm), flush=True)
        # step 0: same parameters on both sides -> summation-order noise only; step 1: after one (equivalent) update
        tol_g = 2e-4 if it == 0 else 3e-2
        assert e_loss <= (1e-5 if it == 0 else 1e-3), e_loss
        assert e_grad <= tol_g, (e_grad, where)
        assert e_norm <= (1e-4 if it == 0 else 1e-2), e_norm
    dp = float((dp_model.engine.p_flat - solo.engine.p_flat).abs().max())
    assert dp <= 4 * 1e-4 * 2, dp                  # two AdamW steps move an entry by at most ~lr each (lr <= 1e-4)


def replicas_identical(family, layers, B, R, steps):
    cfg, tc, model, opt = build(family, layers, R)
    model.train()
    parallel.attach(model, opt, world)
    seeds = [torch.tensor([model.engine.base_seed], dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(seeds, seeds[rank].clone())
    assert len({int(s) for s in seeds}) == world, "ranks must draw different dropout masks"
    for it in range(steps):
        b = batch(family, B, R, 200 + 7 * it + rank)
        loss, _ = TU.ForwardModelsTrain(cfg, tc, dev, "TASK15", to_dev(b), model)
        loss.backward()
        optim.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        model.zero_grad()
        flat = model.engine.p_flat
        chk = torch.stack([flat.double().sum(), flat.double().abs().sum(), flat[::997].double().pow(2).sum()])
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        same = all(torch.equal(allc[0], c) for c in allc)
        if rank == 0:
            print("%s step %d loss %.4f params identical across %d ranks: %s" % (family, it, float(loss), world, same), flush=True)
        assert same and torch.isfinite(loss)


n_vs_1("uc2", 2, 8, 36)
n_vs_1("m3p", 2, 4, 100)
replicas_identical("uc2", 2, 16, 36, 3)
replicas_identical("m3p", 2, 8, 100, 2)
dist.barrier()
if rank == 0:
    print("dp_check ok", flush=True)
dist.destroy_process_group()

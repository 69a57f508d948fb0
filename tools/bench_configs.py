"""Throughput of the other BASELINE.json configs on one B200 (the parity cases of the suite at their full sizes):
  cfg 3  M3P (XLM-R 250k vocab) bf16 fine-tune step, B = 256, 100 regions          (per-GPU part of the 8-GPU DP run)
  cfg 4  UC2 sparse fine-tuning (BERT-Tickets masks, ~41 % / ~70 % zeros), B = 256, 36 regions
  cfg 5  xGQA zero-shot eval: forward + argmax, B = 1024, 100 regions (UC2 and M3P)
Prints one JSON line per case.   python tools/bench_configs.py [case ...]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clg_vqa_b200 import config as C, optim, sft, synthetic as S, task_utils as TU
from clg_vqa_b200.encoders import BertForVLTasks, M3PForVLTasks

dev = torch.device("cuda", 0)


def timed(fn, steps, warmup=3):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def build(family, R, masks_density=None):
    torch.manual_seed(0)
    if family == "uc2":
        cfg = C.uc2_base(12); tc = C.gqa_task_cfg(R); model = BertForVLTasks(cfg, tc, ["TASK15"])
    else:
        cfg = C.m3p_base(12); tc = C.gqa_task_cfg(R); model = M3PForVLTasks(cfg, tc, ["TASK15"])
    if masks_density is not None:
        names = sft.uc2_masked_module_names(12)
        masks = sft.synth_mask_dict(model, names, density=masks_density, seed=0)
        model.load_state_dict(sft.premask_state_dict(model.state_dict(), masks))
        sft.pruning_model_custom(model, masks, names=names)
    model.set_precision("bf16"); model.to(dev)
    return cfg, tc, model


def train_case(name, family, B, R, masks_density=None, steps=8):
    cfg, tc, model = build(family, R, masks_density)
    model.train()
    opt = optim.AdamW(optim.grouped_parameters(model, 4e-5, 1e-4, sft=masks_density is not None), lr=4e-5)
    nl = 7 if family == "uc2" else 5
    batches = [tuple(t.to(dev) for t in S.batch_tuple(S.synth_batch(B, num_regions=R, num_locs=nl, seed=i, normalize=(family == "m3p"))))
               for i in range(2)]
    it = [0]
    def step():
        loss, _ = TU.ForwardModelsTrain(cfg, tc, dev, "TASK15", batches[it[0] % 2], model)
        loss.backward(); optim.clip_grad_norm_(model.parameters(), 1.0); opt.step(); model.zero_grad(); it[0] += 1
        return loss
    ms = timed(step, steps)
    out = {"case": name, "samples_per_s": B / ms * 1e3, "ms_per_step": ms, "B": B, "R": R, "family": family, "loss": float(step())}
    if masks_density is not None:
        out["mask_zero_fraction"] = 1.0 - masks_density
    print(json.dumps(out), flush=True)
    del model, opt
    torch.cuda.empty_cache()


def eval_case(name, family, B, R, steps=6):
    cfg, tc, model = build(family, R)
    model.eval()
    nl = 7 if family == "uc2" else 5
    b = S.batch_tuple(S.synth_batch(B, num_regions=R, num_locs=nl, seed=3, normalize=(family == "m3p")))
    b = tuple(t.to(dev) for t in b)
    class DS: label2ans = None
    class DL: dataset = DS()
    def step():
        res = []
        TU.EvaluatingModel(cfg, tc, dev, "TASK15", b, model, DL(), None, res, [])
        return res
    for _ in range(2): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(steps): r = step()
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3 / steps        # includes the D2H of the predictions
    print(json.dumps({"case": name, "samples_per_s": B / ms * 1e3, "ms_per_batch": ms, "B": B, "R": R, "family": family,
                      "predictions": len(r)}), flush=True)
    del model
    torch.cuda.empty_cache()


CASES = {
    "cfg3_m3p_train": lambda: train_case("cfg3_m3p_train_b256_r100", "m3p", 256, 100),
    "cfg4_sft41": lambda: train_case("cfg4_uc2_sft_41pct_zeros", "uc2", 256, 36, masks_density=0.59),
    "cfg4_sft70": lambda: train_case("cfg4_uc2_sft_70pct_zeros", "uc2", 256, 36, masks_density=0.30),
    "cfg5_uc2_eval": lambda: eval_case("cfg5_uc2_eval_b1024_r100", "uc2", 1024, 100),
    "cfg5_m3p_eval": lambda: eval_case("cfg5_m3p_eval_b1024_r100", "m3p", 1024, 100),
}
for c in (sys.argv[1:] or list(CASES)):
    CASES[c]()

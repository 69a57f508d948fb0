#!/usr/bin/env python
"""Benchmark of the hot path.  Default: one UC2-base (12L, 768-d) GQA-shape fine-tune step = forward, label-prior
loss, backward, global-norm clip, AdamW - BASELINE.json's metric "fine-tune samples/sec (UC2-base, GQA
shape)" on configs[1] (batch 256 per GPU, 36 regions, bf16, WordNet/embedding-distance prior loss).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|torch_gpu]
                    [--config uc2_train|m3p_train|uc2_sft41|uc2_sft70|eval_uc2|eval_m3p] [--scaling weak|strong] [--phases]

--config selects the other BASELINE.json configs (same JSON schema: value, e2e, roofline, clocks):
    uc2_train  configs[1]  UC2 fine-tune step, B = 256 / GPU, 36 regions                      (the default, the headline)
    m3p_train  configs[2]  M3P (XLM-R 250k vocabulary) fine-tune step, B = 256 / GPU, 100 regions, data parallel
    uc2_sft41, uc2_sft70   configs[3]  UC2 sparse fine-tuning, BERT-Tickets masks with 41 % / 70 % zeros
    eval_uc2, eval_m3p     configs[4]  xGQA zero-shot eval: forward + argmax, B = 1024, 100 regions (one language shard per rank)
Prints ONE JSON line.  Under torchrun (N > 1) every rank runs its own per-GPU batch (weak scaling; --scaling strong divides
the global batch of 256 by N, the reference's own multi-GPU mode, task_utils.py:L473-482), gradients are exchanged with
NCCL, time is the max over ranks.
--impl reference   the CPU oracle port of the reference step on the host cores (rank 0 only);
--impl torch_gpu   the same step in stock PyTorch on the SAME GPU (oracle/torch_gpu_baseline.py: cuBLAS + SDPA + fused
                   AdamW under torch.autocast(bf16)) - what plain PyTorch gets on this B200, a baseline only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "fine-tune samples/sec (UC2-base, GQA shape)"
UNIT = "samples/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: an NVML polling thread (nvidia_ml_py) with a 5 ms
    period, so even a 150 ms region gets a few dozen samples; `nvidia-smi -lms` is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []
        self.samples = []            # (sm_mhz, reasons bitmask)
        self.max_mhz = None
        self._stop = threading.Event()
        self.thread = None
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES remapping: look the device up by its PCI bus id
            bus = getattr(torch.cuda.get_device_properties(self.index), "pci_bus_id", None)
            h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(hi).bus == bus:
                        h = hi
                        break
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nvml = (pynvml, h)

            def poll():
                get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                    getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
                while not self._stop.is_set():
                    try:
                        self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), int(get_reasons(h))))
                    except Exception:
                        pass
                    time.sleep(0.005)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=1)
            pynvml = self.nvml[0]
            def bit(ev, thr, dflt):
                return getattr(pynvml, ev, getattr(pynvml, thr, dflt))
            bits = {"hw_slowdown": bit("nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": bit("nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": bit("nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": bit("nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            sm = sorted(s for s, _ in self.samples)
            reasons = sorted(n for n, b in bits.items() if any(r & b for _, r in self.samples))
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# --------------------------------------------------------------------------------------------------
# configurations (BASELINE.json configs[1..4])
# --------------------------------------------------------------------------------------------------
CONFIGS = {
    "uc2_train": dict(family="uc2", mode="train", B=256, R=36, masks=None, which="configs[1]",
                      metric=METRIC, what="UC2-base (12L, 768-d) fine-tune step (fwd + label-prior loss + bwd + clip + AdamW, dropout 0.1)"),
    "m3p_train": dict(family="m3p", mode="train", B=256, R=100, masks=None, which="configs[2]",
                      metric="fine-tune samples/sec (M3P, GQA shape)",
                      what="M3P (XLM-R 250k vocabulary, 12L) fine-tune step on code-mixed ids (full multilingual id range)"),
    "uc2_sft41": dict(family="uc2", mode="train", B=256, R=36, masks=0.59, which="configs[3]",
                      metric="sparse fine-tune samples/sec (UC2-base, 41% zeros)",
                      what="UC2 sparse fine-tuning step, BERT-Tickets masks (41 % zeros) fused into the GEMM weight shadow / wgrad epilogue"),
    "uc2_sft70": dict(family="uc2", mode="train", B=256, R=36, masks=0.30, which="configs[3]",
                      metric="sparse fine-tune samples/sec (UC2-base, 70% zeros)",
                      what="UC2 sparse fine-tuning step, BERT-Tickets masks (70 % zeros) fused into the GEMM weight shadow / wgrad epilogue"),
    "eval_uc2": dict(family="uc2", mode="eval", B=1024, R=100, masks=None, which="configs[4]",
                     metric="zero-shot eval samples/sec (UC2-base, xGQA shape)",
                     what="xGQA zero-shot eval: forward + answer argmax + one D2H of the predictions per batch, X101 features"),
    "eval_m3p": dict(family="m3p", mode="eval", B=1024, R=100, masks=None, which="configs[4]",
                     metric="zero-shot eval samples/sec (M3P, xGQA shape)",
                     what="xGQA zero-shot eval: forward + answer argmax + one D2H of the predictions per batch, X101 features"),
}


def flops_per_sample(family="uc2", R=36, T=40, d=768, L=12, A=1842):
    """(training, forward) FLOP per padded sample; multiply-add = 2 FLOP."""
    C, locs = (768, 7) if family == "uc2" else (1536, 5)
    S = T + R
    fwd = L * (24 * S * d * d + 4 * S * S * d) + R * 2 * d * (2048 + locs) + 2 * d * d + 2 * d * C + 2 * C * A
    return 3 * fwd - R * 2 * d * (2048 + locs), fwd


def train_flops_per_sample(R=36, T=40, d=768, L=12, C=768, A=1842, locs=7):
    return flops_per_sample("uc2", R, T, d, L, A)


def build_model(spec, layers, dev=None, precision="bf16"):
    from clg_vqa_b200 import config as C, sft
    from clg_vqa_b200.encoders import BertForVLTasks, M3PForVLTasks
    R = spec["R"]
    task_cfg = C.gqa_task_cfg(R)
    torch.manual_seed(0)
    if spec["family"] == "uc2":
        cfg = C.uc2_base(layers)
        model = BertForVLTasks(cfg, task_cfg, ["TASK15"])
    else:
        cfg = C.m3p_base(layers)
        model = M3PForVLTasks(cfg, task_cfg, ["TASK15"])
    with torch.no_grad():
        for n, p in model.named_parameters():           # non-vacuous LN / bias paths
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    if spec["masks"] is not None:
        names = sft.uc2_masked_module_names(layers)
        masks = sft.synth_mask_dict(model, names, density=spec["masks"], seed=0)
        model.load_state_dict(sft.premask_state_dict(model.state_dict(), masks))
        sft.pruning_model_custom(model, masks, names=names)
    if dev is not None:
        model.set_precision(precision)
        model.to(dev)
    return cfg, task_cfg, model


def host_batch(spec, B, seed):
    from clg_vqa_b200 import synthetic as S
    fam = spec["family"]
    b = S.synth_batch(B, num_regions=spec["R"], num_locs=7 if fam == "uc2" else 5, seed=seed, normalize=(fam == "m3p"))
    return b


# --------------------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference step (also the --impl reference arm)
# --------------------------------------------------------------------------------------------------
def cpu_reference_step_rate(B=8, R=36, steps=2, warmup=1, layers=12, with_optimizer=True, family="uc2", mode="train"):
    from clg_vqa_b200 import synthetic as S
    from oracle import adamw_oracle, loss_oracle, m3p_oracle, uc2_oracle
    torch.set_num_threads(os.cpu_count() or 1)
    spec = dict(family=family, R=R, masks=None)
    cfg, _, model = build_model(spec, layers)
    fwd = uc2_oracle.uc2_forward if family == "uc2" else m3p_oracle.m3p_forward
    train = mode == "train"
    sd = {k: v.detach().clone().requires_grad_(train) for k, v in model.state_dict().items()
          if S.canonical_key(k) == k}
    all_keys = list(model.state_dict().keys())
    full = lambda: {k: sd[S.canonical_key(k)] for k in all_keys}   # noqa: E731
    del model
    b = host_batch(spec, B, 1234)
    st = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in sd.items()} if (with_optimizer and train) else None
    times = []
    loss = torch.zeros(())
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        with torch.set_grad_enabled(train):
            logits = fwd(full(), cfg, b["question"], b["features"], b["spatials"], b["segment_ids"],
                         b["input_mask"], b["image_mask"])
        if train:
            loss, score = loss_oracle.gqa_prior_loss(logits, b["target"], b["distances"], 10.0)
            loss.backward()
            if with_optimizer:
                with torch.no_grad():
                    grads = [v.grad for v in sd.values() if v.grad is not None]
                    adamw_oracle.clip_grad_norm(grads, 1.0)
                    for k, v in sd.items():
                        if v.grad is None:
                            continue
                        wd = 0.0 if any(nd in k for nd in adamw_oracle.NO_DECAY) else 1e-4
                        adamw_oracle.adamw_step(v, v.grad, st[k][0], st[k][1], it + 1, 4e-5, weight_decay=wd)
                        v.grad = None
        else:
            logits.argmax(1).tolist()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return B / mean, mean, float(loss)


def run_reference_arm(args, spec):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = 8
    rate, sec, _ = cpu_reference_step_rate(B=B, R=spec["R"], steps=max(1, args.steps), warmup=max(1, min(args.warmup, 1)),
                                           family=spec["family"], mode=spec["mode"], layers=args.layers)
    cores = os.cpu_count() or 1
    what = "full step (fwd + prior loss + bwd + clip + AdamW)" if spec["mode"] == "train" else "forward + argmax"
    sample = ("oracle PORT of the reference (plain-BERT restatement, not the reference's own gated bi-stream module graph), %s %dL %s on "
              "a B=%d sample of the batch, fp32, %d host threads" % (spec["family"].upper(), args.layers, what, B, cores))
    line = {
        "impl": "reference", "metric": spec["metric"], "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": max(1, args.steps),
        "warmup": max(1, min(args.warmup, 1)), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE.json %s: %s, batch %d/GPU, %d regions (CPU arm: bounded B=8 sample per step)"
                               % (spec["which"], spec["what"], spec["B"], spec["R"]), "regions": spec["R"], "name": args.config},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# stock PyTorch on the same GPU (baseline only)
# --------------------------------------------------------------------------------------------------
def run_torch_gpu(args, spec):
    from clg_vqa_b200 import synthetic as S
    from oracle import loss_oracle, torch_gpu_baseline as tb
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    B = args.batch or spec["B"]
    cfg, task_cfg, model = build_model(dict(spec, masks=None), args.layers)
    train = spec["mode"] == "train"
    sd = {}
    canon = {}
    for k, v in model.state_dict().items():
        ck = S.canonical_key(k)
        if ck not in canon:
            canon[ck] = v.detach().to(dev).clone().requires_grad_(train)
        sd[k] = canon[ck]
    del model
    fam = spec["family"]
    fp = tb.FusedParams(sd, fam, args.layers)
    fwd = tb.uc2_forward if fam == "uc2" else tb.m3p_forward
    params = fp.parameters(fam)
    opt = torch.optim.AdamW(params, lr=4e-5, eps=1e-6, weight_decay=1e-4, fused=True) if train else None
    NB = 2
    host = [tuple(t.pin_memory() for t in S.batch_tuple(host_batch(spec, B, 1234 + i))) for i in range(NB)]
    resident = [tuple(t.to(dev) for t in hb) for hb in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])

    def step(batch):
        features, spatials, image_mask, question, target, input_mask, segment_ids, qid, _, distances = batch
        with torch.autocast("cuda", dtype=torch.bfloat16), torch.set_grad_enabled(train):
            logits = fwd(fp, cfg, question, features, spatials, segment_ids, input_mask, image_mask, train=train)
        if not train:
            return logits.argmax(1).cpu()
        loss, _ = loss_oracle.gqa_prior_loss(logits, target, distances, 10.0)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def timed(fn, n):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    for i in range(max(3, args.warmup)):
        step(resident[i % NB])
    clocks = ClockSampler(dev.index)
    clocks.start()
    ms = timed(lambda i: step(resident[i % NB]), args.steps)
    clk = clocks.stop()

    def e2e_step(i):
        out = step(tuple(t.to(dev, non_blocking=True) for t in host[i % NB]))
        if train:
            float(out.detach())
    e2e_step(0)
    ms_e2e = timed(e2e_step, args.steps)
    value = B * args.steps / (ms * 1e-3)
    train_f, fwd_f = flops_per_sample(fam, spec["R"], L=args.layers)
    peaks = measured_peaks()
    tfl = value * (train_f if train else fwd_f) / 1e12
    line = {
        "impl": "torch_gpu", "metric": spec["metric"], "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "BASELINE.json %s: %s, batch %d, text 40 + %d regions; STOCK PyTorch on this GPU: torch.autocast(bf16), cuBLAS "
                               "linears with a fused QKV weight, F.scaled_dot_product_attention, F.layer_norm / F.gelu, torch clip_grad_norm_, "
                               "torch.optim.AdamW(fused=True); padded rows (no packing)" % (spec["which"], spec["what"], B, spec["R"]),
                   "name": args.config, "global_batch": B, "regions": spec["R"]},
        "e2e": {"value": B * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4 if train else 8 * B,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": 0, "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "whole step (library kernels)", "achieved": tfl, "peak": peaks["bf16_tflops_sustained"],
                     "unit": "TFLOP/s", "frac": tfl / peaks["bf16_tflops_sustained"], "traffic": None,
                     "note": "reference-equivalent FLOPs (SURVEY 8d, padded sample) x samples/s over the whole step"},
        "cpu_baseline": None,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_ours(args, spec):
    import torch.distributed as dist
    from clg_vqa_b200 import ops, optim, synthetic as S, task_utils as TU

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")                 # these levels print "NCCL version ..." on stdout: keep it to the one JSON line
        dist.init_process_group(backend="nccl", device_id=dev)
    ops.device_check(local_rank)

    train = spec["mode"] == "train"
    fam = spec["family"]
    R, T = spec["R"], 40
    B = args.batch or spec["B"]
    if args.scaling == "strong":
        if B % world:
            raise SystemExit("--scaling strong: the global batch %d does not divide by %d ranks" % (B, world))
        B = B // world                                 # the reference's multi-GPU mode: batch_size / world per process
    cfg, task_cfg, model = build_model(spec, args.layers, dev, args.precision)
    model.train() if train else model.eval()
    opt = sched = None
    if train:
        groups = optim.grouped_parameters(model, base_lr=4e-5, weight_decay=1e-4, sft=spec["masks"] is not None)
        opt = optim.AdamW(groups, lr=4e-5, betas=(0.9, 0.999), eps=1e-6, correct_bias=True)
        sched = optim.WarmupLinearSchedule(opt, warmup_steps=100, t_total=10000)
    ex = None
    if world > 1 and train:
        from clg_vqa_b200 import parallel
        ex = parallel.attach(model, opt, world)

    # host batches (pinned), distinct per step modulo NB, rank-seeded
    NB = 2
    host = []
    for i in range(NB):
        tup = S.batch_tuple(host_batch(spec, B, 1234 + 17 * rank + i))
        if not train:
            tup = tup[:8]
        if args.features == "bf16":                      # features delivered as bf16 (the GEMM operand dtype)
            tup = (tup[0].to(torch.bfloat16),) + tuple(tup[1:])
        host.append(tuple(t.pin_memory() for t in tup))
    resident = [tuple(t.to(dev) for t in hb) for hb in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])

    class _DS:
        label2ans = None

    class _DL:
        dataset = _DS()

    def step(batch):
        if train:
            loss, score = TU.ForwardModelsTrain(cfg, task_cfg, dev, "TASK15", batch, model)
            loss.backward()
            optim.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            sched.step()
            model.zero_grad()
            return loss, score
        res = []
        TU.EvaluatingModel(cfg, task_cfg, dev, "TASK15", batch, model, _DL(), None, res, [])    # incl. the D2H of the predictions
        return res, None

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(loop_fn, steps):
        sync_all()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        loop_fn(steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        sync_all()
        return ms

    def loop_resident(n):
        for i in range(n):
            step(resident[i % NB])

    losses = []
    pf_state = {}

    def loop_e2e(n):
        # every step's inputs come from pinned HOST memory: DevicePrefetcher issues the H2D copy of batch i+1 on a side
        # stream while step i computes (all n copies happen inside the timed region), and every step's result (the loss /
        # the predictions) is read back
        from clg_vqa_b200.data import DevicePrefetcher
        pf = DevicePrefetcher((host[i % NB] for i in range(n)), dev, reuse=pf_state.get("pf"))
        pf_state["pf"] = pf
        # every step's loss is copied to pinned host memory on the compute stream and read by the host ONE step later
        # (the logger's float(loss) of train_task.py:L340, without stalling the launch queue); the last one before the end
        pin = pf_state.setdefault("pin", torch.empty(4096, dtype=torch.float32).pin_memory())
        prev = None
        for i, batch in enumerate(pf):
            out, _ = step(batch)
            if train:
                pin[i % 4096].copy_(out.detach(), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                if prev is not None:
                    prev[0].synchronize()
                    losses.append(float(pin[prev[1] % 4096]))          # D2H read of the previous step's loss
                prev = (ev, i)
        if prev is not None:
            prev[0].synchronize()
            losses.append(float(pin[prev[1] % 4096]))

    # warm-up (also builds arenas / workspaces)
    loop_resident(max(3, args.warmup))
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = ops.launch_count()
    ms = timed(loop_resident, args.steps)
    launches = ops.launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    loop_e2e(2)                                   # untimed: device slots / copy stream of the input pipeline exist
    losses.clear()
    ms_e2e = timed(loop_e2e, args.steps)

    phases = None
    if args.phases and ex is not None:
        phases = measure_phases(model, ex, lambda: step(resident[0]), dist, dev, world)

    # dominant kernel (tcgen05 GEMM) timed live with CUDA events on the launching stream
    gemm_rec = []
    ops.set_gemm_timing(gemm_rec)
    loop_resident(2)
    torch.cuda.synchronize()
    ops.set_gemm_timing(None)
    g_ms = sum(a.elapsed_time(b) for (a, b, f) in gemm_rec)
    g_flops = sum(f for (a, b, f) in gemm_rec)          # 2*M*N*K of the GEMMs actually launched (packed rows)
    peaks = measured_peaks()
    achieved = g_flops / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    # rows the encoder layers really process (masked positions are dropped, engine._forward) vs the padded B*S
    packed_rows = getattr(model.engine, "last_rows", None)
    traffic = None
    for name in ("r02_gemm_traffic.json", "r01_gemm_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
            break
        except Exception:
            traffic = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = B * world * args.steps / (ms * 1e-3)
    e2e = B * world * args.steps / (ms_e2e * 1e-3)
    train_f, fwd_f = flops_per_sample(fam, R, L=args.layers)
    step_tflops = value / world * (train_f if train else fwd_f) / 1e12
    cpu = None
    if world == 1 and not args.no_cpu_baseline and args.config == "uc2_train":
        rate, sec, _ = cpu_reference_step_rate(B=8, R=R, steps=16, warmup=1, layers=args.layers)    # about 10 s on 16 cores
        cpu = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "oracle PORT of the reference (plain-BERT restatement), UC2-base %dL full step on a B=8 sample x16 (fp32, all host "
                         "threads, %.1f s of CPU work)" % (args.layers, 16 * sec)}
    d2h = 4 if train else 16 * B
    line = {
        "metric": spec["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "BASELINE.json %s: %s, batch %d per GPU, text 40 + %d regions x 2048-d, 1842 answers, random init"
                               % (spec["which"], spec["what"], B, R),
                   "name": args.config, "global_batch": B * world, "regions": R, "parallelism": "dp%d" % world,
                   "features": args.features,
                   "packing": ("masked positions (text padding: lengths U{6..40}) are dropped before the encoder layers - exact, "
                               "see DESIGN.md section 5; rows per step %s of %d" % (packed_rows, B * (T + R))),
                   "l2_policy": "no explicit flush: one step streams several GB of activations/weights/optimizer state (>> 126 MB L2)"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05, all instantiations)", "achieved": achieved,
                     "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops_sustained"],
                     "peak_source": peaks["source"] + " (sustained bf16 cuBLAS figure: the kernel is timed inside a long step)",
                     "traffic": traffic, "gemm_ms_per_step": g_ms / 2, "gemm_launches_per_step": len(gemm_rec) // 2,
                     "gemm_flops_per_step_executed": g_flops / 2,
                     "step_executed_tflops": (g_flops / 2) / (ms / args.steps * 1e-3) / 1e12,
                     "step_executed_frac_of_peak": (g_flops / 2) / (ms / args.steps * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"],
                     "step_reference_equiv_tflops": step_tflops,
                     "step_reference_equiv_frac_of_burst_peak": step_tflops / peaks["bf16_tflops"],
                     "note": "achieved / frac: executed GEMM FLOPs (2MNK of the launches, packed rows) over their summed CUDA-event time; "
                             "step_executed_*: the same FLOPs over the whole step; step_reference_equiv_tflops: SURVEY 8d's FLOPs "
                             "per padded sample x samples/s (what the reference would have to execute for this throughput)"},
        "cpu_baseline": cpu,
        "loss_first_last": [losses[0], losses[-1]] if losses else None,
    }
    if phases is not None:
        line["phases"] = phases
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_phases(model, ex, one_step, dist, dev, world, reps=5):
    """per-phase CUDA-event breakdown of a data-parallel step (max over ranks, mean over `reps` steps): forward + backward
    compute, then the exposed tail of the exchange and the optimiser - see parallel.GradExchange.trace"""
    acc = {}
    for _ in range(reps):
        ex.trace = []
        torch.cuda.synchronize()
        dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        one_step()
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        torch.cuda.synchronize()
        marks = [("start", e0)] + list(ex.trace) + [("step_end", e1)]
        ex.trace = None
        for (n0, a), (n1, b) in zip(marks[:-1], marks[1:]):
            acc.setdefault(n1, []).append(a.elapsed_time(b))
    out = {}
    for k, v in acc.items():
        t = torch.tensor([sum(v) / len(v)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["ms_until_" + k] = float(t)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"])
    ap.add_argument("--config", default="uc2_train", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--phases", action="store_true", help="N > 1: add a per-phase CUDA-event breakdown of the step")
    ap.add_argument("--features", default="fp32", choices=["fp32", "bf16"], help="dtype of the region features in the host batches")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the config's)")
    ap.add_argument("--regions", type=int, default=0)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    spec = dict(CONFIGS[args.config])
    if args.regions:
        spec["R"] = args.regions
    if args.impl == "reference":
        run_reference_arm(args, spec)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device visible and there is no CPU fallback for the product path")
    if args.impl == "torch_gpu":
        run_torch_gpu(args, spec)
        return
    run_ours(args, spec)


if __name__ == "__main__":
    main()

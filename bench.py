#!/usr/bin/env python
"""Benchmark of the hot path: one UC2-base (12L, 768-d) GQA-shape fine-tune step = forward, label-prior
loss, backward, global-norm clip, AdamW - BASELINE.json's metric "fine-tune samples/sec (UC2-base, GQA
shape)" on configs[1] (batch 256 per GPU, 36 regions, bf16, WordNet/embedding-distance prior loss).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (see the keys below).  Under torchrun (N > 1) every rank runs the same per-GPU
batch (weak scaling), gradients are summed with NCCL, time is the max over ranks.
`--impl reference` times the CPU oracle port of the reference step on the host cores (rank 0 only).
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


def train_flops_per_sample(R=36, T=40, d=768, L=12, C=768, A=1842, locs=7):
    """SURVEY.md section 8d / BASELINE.md section 3: multiply-add = 2 FLOP."""
    S = T + R
    fwd = L * (24 * S * d * d + 4 * S * S * d) + R * 2 * d * (2048 + locs) + 2 * d * d + 2 * d * C + 2 * C * A
    return 3 * fwd - R * 2 * d * (2048 + locs), fwd


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
# CPU arm: oracle port of the reference step (also the --impl reference arm)
# --------------------------------------------------------------------------------------------------
def cpu_reference_step_rate(B=8, R=36, steps=2, warmup=1, layers=12, with_optimizer=True):
    from clg_vqa_b200 import config as C, synthetic as S
    from clg_vqa_b200.encoders import BertForVLTasks
    from oracle import adamw_oracle, loss_oracle, uc2_oracle
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = C.uc2_base(layers)
    model = BertForVLTasks(cfg, C.gqa_task_cfg(R), ["TASK15"])
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()
          if S.canonical_key(k) == k}
    all_keys = list(model.state_dict().keys())
    full = lambda: {k: sd[S.canonical_key(k)] for k in all_keys}   # noqa: E731
    del model
    b = S.synth_batch(B, num_regions=R, seed=1234)
    st = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in sd.items()} if with_optimizer else None
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        logits = uc2_oracle.uc2_forward(full(), cfg, b["question"], b["features"], b["spatials"], b["segment_ids"],
                                        b["input_mask"], b["image_mask"])
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
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return B / mean, mean, float(loss)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = 8
    rate, sec, _ = cpu_reference_step_rate(B=B, R=args.regions, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 1)))
    cores = os.cpu_count() or 1
    sample = "UC2-base 12L full step (fwd + prior loss + bwd + clip + AdamW) on a B=%d sample of the batch, fp32, %d host threads" % (B, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": max(1, args.steps),
        "warmup": max(1, min(args.warmup, 1)), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[1]: UC2-base fine-tune step, batch 256/GPU, 36 regions, label-prior loss "
                               "(CPU arm: bounded B=8 sample per step)", "regions": args.regions},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from clg_vqa_b200 import config as C, ops, optim, synthetic as S, task_utils as TU
    from clg_vqa_b200.encoders import BertForVLTasks

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

    B, R, T = args.batch, args.regions, 40
    cfg = C.uc2_base(args.layers)
    task_cfg = C.gqa_task_cfg(R)
    torch.manual_seed(0)
    model = BertForVLTasks(cfg, task_cfg, ["TASK15"])
    with torch.no_grad():
        for n, p in model.named_parameters():           # non-vacuous LN / bias paths
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    model.set_precision(args.precision)
    model.to(dev)
    model.train()
    groups = optim.grouped_parameters(model, base_lr=4e-5, weight_decay=1e-4)
    opt = optim.AdamW(groups, lr=4e-5, betas=(0.9, 0.999), eps=1e-6, correct_bias=True)
    sched = optim.WarmupLinearSchedule(opt, warmup_steps=100, t_total=10000)
    if world > 1:
        from clg_vqa_b200 import parallel
        parallel.attach(model, opt, world)

    # host batches (pinned), distinct per step modulo NB, rank-seeded
    NB = 2
    host = []
    for i in range(NB):
        b = S.synth_batch(B, num_regions=R, seed=1234 + 17 * rank + i)
        tup = S.batch_tuple(b)
        host.append(tuple(t.pin_memory() for t in tup))
    resident = [tuple(t.to(dev) for t in hb) for hb in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])

    def step(batch):
        loss, score = TU.ForwardModelsTrain(cfg, task_cfg, dev, "TASK15", batch, model)
        loss.backward()
        optim.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        sched.step()
        model.zero_grad()
        return loss, score

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
        # stream while step i computes (all n copies happen inside the timed region), and every step's loss is read back
        from clg_vqa_b200.data import DevicePrefetcher
        pf = DevicePrefetcher((host[i % NB] for i in range(n)), dev, reuse=pf_state.get("pf"))
        pf_state["pf"] = pf
        for batch in pf:
            loss, _ = step(batch)
            losses.append(float(loss.detach()))   # D2H read of the step's loss

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
    sv_ws = model.engine._ws if hasattr(model, "engine") else {}
    packed_rows = None
    try:
        for w_ in sv_ws.values():
            if "cu" in w_:
                packed_rows = int(w_["cu"][B])
    except Exception:
        packed_rows = None
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        traffic = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = B * world * args.steps / (ms * 1e-3)
    e2e = B * world * args.steps / (ms_e2e * 1e-3)
    train_f, _ = train_flops_per_sample(R=R, L=args.layers)
    step_tflops = value / world * train_f / 1e12
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, sec, _ = cpu_reference_step_rate(B=8, R=R, steps=16, warmup=1, layers=args.layers)    # about 10 s on 16 cores
        cpu = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "oracle port, UC2-base %dL full step on a B=8 sample x16 (fp32, all host threads, %.1f s of CPU work)" % (args.layers, 16 * sec)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[1]: UC2-base (%dL, 768-d) fine-tune step (fwd + label-prior loss + bwd + "
                               "clip + AdamW, dropout 0.1), batch %d per GPU, text 40 + %d regions x 2048-d, 1842 answers, random init"
                               % (args.layers, B, R),
                   "global_batch": B * world, "regions": R, "parallelism": "dp%d" % world,
                   "packing": ("masked positions (text padding: lengths U{6..40}) are dropped before the encoder layers - exact, "
                               "see DESIGN.md section 5; rows per step %s of %d" % (packed_rows, B * (T + R))),
                   "l2_policy": "no explicit flush: one step streams ~8 GB of activations/weights/optimizer state (>> 126 MB L2)"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
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
                     "note": "achieved / frac: executed GEMM FLOPs (2MNK of the launches, packed rows) over their summed CUDA-event time; "
                             "step_executed_*: the same FLOPs over the whole step; step_reference_equiv_tflops: SURVEY 8d's 39.61 GFLOP "
                             "per padded sample x samples/s (what the reference would have to execute for this throughput)"},
        "cpu_baseline": cpu,
        "loss_first_last": [losses[0], losses[-1]] if losses else None,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--regions", type=int, default=36)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device visible and there is no CPU fallback for the product path")
    run_ours(args)


if __name__ == "__main__":
    main()

#!/bin/bash
# 8-GPU experiments of the data-parallel step (run under gpurun --gpus 8)
out=gpurun_out; mkdir -p $out
N=${N:-8}
python -m pytest tests/test_ops2_gpu.py -m gpu -q -k "segment or sumsq or adamw" 2>&1 | tail -2
run() {  # $1 = tag, rest = env assignments / bench args
  tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --phases $EXTRA > $out/bench_n${N}_$tag.log 2>&1
  python - $out/bench_n${N}_$tag.log $tag <<'PY'
import json, sys
try:
    l = [x for x in open(sys.argv[1]).read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l)
    print(sys.argv[2], "value %.0f ms %.3f e2e %.0f gemm_frac %.3f scaling %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["scaling"]))
    print("   phases", {k: round(v, 3) for k, v in (d.get("phases") or {}).items()})
except Exception as e:
    print(sys.argv[2], "FAILED", e); print(open(sys.argv[1]).read()[-1500:])
PY
}
run default A=1
run nobucket CLG_DP_BUCKET_LAYERS=0
run bucket2 CLG_DP_BUCKET_LAYERS=2
run maxctas8 NCCL_MAX_CTAS=8
EXTRA="--scaling strong" run strong A=1
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 2 --warmup 3 2>&1 | grep -i -E "NVLS|Channel|algo|Connected" | sort | uniq -c | sort -rn | head -12 > $out/nccl_info_n$N.txt; head -12 $out/nccl_info_n$N.txt

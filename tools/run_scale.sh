#!/bin/bash
# weak-scaling points of the default bench (the driver's launch line) on the GPUs of this box: N = $@
out=gpurun_out; mkdir -p $out
for N in "$@"; do
  if [ $N = 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $out/r02_scale_n$N.json 2> $out/r02_scale_n$N.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --steps 20 --warmup 5 > $out/r02_scale_n$N.json 2> $out/r02_scale_n$N.err; fi
done
python tools/show_bench.py $out/r02_scale_n*.json

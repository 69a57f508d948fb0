#!/bin/bash
# usage: tools/gpu_retry_n.sh <ngpus> <logfile> <timeout_s> <command...>
n=$1; shift; log=$1; shift; to=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --gpus $n --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|rc=3\|busy" $log && ! grep -q "status=ok" $log; then sleep 120; continue; fi
  break
done

#!/bin/bash
# usage: tools/gpu_retry.sh <logfile> <timeout_s> <command...>   (retries while the pod answers "transient")
log=$1; shift; to=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient" $log; then sleep 90; continue; fi
  break
done

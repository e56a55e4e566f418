#!/bin/bash
# usage: scripts/gpu_retry.sh <outfile> <timeout> <command...>   -- retries while the pod has no free GPU slot
out=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  gpurun --timeout $to -- "$@" > $out 2>&1
  if grep -q "status=transient" $out; then sleep 45; continue; fi
  break
done

#!/bin/bash
# gpurun with retries while the pod answers "transient" (no box free): tools/gpurun_retry.sh <timeout> [--gpus N] -- '<command>'
t=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" "$@" 2>&1)
  echo "$out" | tail -4
  if echo "$out" | grep -q "status=transient\|status=busy"; then sleep 90; continue; fi
  break
done

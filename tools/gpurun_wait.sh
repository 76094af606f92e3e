#!/bin/bash
# gpurun that stays in the pod's queue: re-submits at once while the pod answers "transient"/"busy" (nothing is
# charged for those): tools/gpurun_wait.sh <timeout> [--gpus N] -- '<command>'
t=$1; shift
for i in $(seq 1 80); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy"; then sleep 5; continue; fi
  echo "$out" | tail -60
  break
done

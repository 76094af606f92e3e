#!/bin/bash
# A/B of (samples per thread, threads per CTA) of the rollout kernel at shard-sized K (environment overrides).
for shape in "131072 100" "262144 100" "524288 100" "16384 50"; do
  set -- $shape
  for ns in 1 2; do for thr in 128 64 32; do
    echo "K=$1 T=$2 NS=$ns threads=$thr $(MPPI_NS=$ns MPPI_ROLL_THREADS=$thr python tools/profile_step.py --K $1 --T $2 --steps 12 --timing 2>&1 | grep -o "'rollout': [0-9.]*")"
  done; done
done

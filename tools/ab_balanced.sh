#!/bin/bash
# A/B of the balanced rollout kernel (MPPI_NS=0: one wave, warp-samples dealt evenly to the warps) against the
# wave-by-wave kernels (MPPI_NS=1 / 2) at shard sizes and on the batched shapes.  rho must agree in every row.
for shape in "131072 100" "262144 100" "524288 100" "1048576 100" "16384 50" "65536 100"; do
  set -- $shape
  for v in "1 128" "2 128" "0 128" "0 64" "0 32"; do
    set -- $shape $v
    echo "K=$1 T=$2 NS=$3 threads=$4 $(MPPI_NS=$3 MPPI_ROLL_THREADS=$4 python tools/profile_step.py --K $1 --T $2 --steps 12 --timing 2>&1 | grep -o "'rollout': [0-9.]*\|^ok.*" | tr '\n' ' ')"
  done
done
for envs in 128 256 512 1024; do
  for v in "1 128" "2 128" "0 128" "0 64" "0 32"; do
    set -- $v
    echo "C5 envs=$envs NS=$1 threads=$2 $(MPPI_NS=$1 MPPI_ROLL_THREADS=$2 python tools/profile_batched.py --envs $envs --steps 12 --timing 2>&1 | grep -o "'rollout': [0-9.]*\|^ok.*" | tr '\n' ' ')"
  done
done

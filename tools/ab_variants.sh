#!/bin/bash
# A/B the rollout-kernel build variants in build/variants/*.so (tools/build_variants.py) at the bench shape
# (K=2^20, T=100, tracking state), the 8-GPU shard size and the latency shape.  SHAPES overrides the list.
for lib in build/variants/*.so; do
  for shape in ${SHAPES:-"1048576,100" "131072,100" "16384,50"}; do
    K=${shape%,*}; T=${shape#*,}
    echo "== $lib K=$K T=$T $(MPPI_B200_LIB=$PWD/$lib python tools/profile_step.py --K $K --T $T --steps 12 --timing 2>&1 | grep -o "'rollout': [0-9.]*")"
  done
done
for lib in ${NS1_LIBS:-}; do
  echo "== $lib NS=1 K=1048576 $(MPPI_NS=1 MPPI_B200_LIB=$PWD/$lib python tools/profile_step.py --K 1048576 --T 100 --steps 12 --timing 2>&1 | grep -o "'rollout': [0-9.]*")"
done

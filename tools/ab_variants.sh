#!/bin/bash
# A/B the rollout-kernel build variants in build/variants/*.so at the bench shape (K=2^20, T=100, tracking state).
for lib in build/variants/*.so; do
  echo "== $lib"
  MPPI_B200_LIB=$PWD/$lib python tools/profile_step.py --K 1048576 --T 100 --steps 12 --timing "$@" 2>&1 | tail -3 | head -1
done
for lib in build/variants/v0_default.so build/variants/v2_cconst_mb5.so; do
  echo "== $lib NS=1"
  MPPI_NS=1 MPPI_B200_LIB=$PWD/$lib python tools/profile_step.py --K 1048576 --T 100 --steps 12 --timing "$@" 2>&1 | tail -3 | head -1
done

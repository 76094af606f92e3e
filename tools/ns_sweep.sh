#!/bin/bash
# Rollout-kernel time by launch layout at shard sizes: default (pick_layout), one sample per thread (MPPI_NS=1),
# two samples per thread without the mixed last wave (MPPI_NO_MIXED=1).
for K in ${KS:-65536 98304 131072 196608 262144 524288 1048576}; do
for mode in "MPPI_UNUSED=1" "MPPI_NS=1" "MPPI_NO_MIXED=1"; do
echo "K=$K $mode $(env $mode python tools/profile_step.py --K $K --T 100 --steps 12 --timing 2>&1 | grep -o "'rollout': [0-9.]*\|^ok.*" | tr '\n' ' ')"
done; done
for E in ${ES:-128 256 512 1024}; do
for mode in "MPPI_UNUSED=1" "MPPI_NS=1" "MPPI_NO_MIXED=1"; do
echo "C5 envs=$E $mode $(env $mode python tools/profile_batched.py --envs $E --steps 12 --timing 2>&1 | grep -o "'rollout': [0-9.]*")"
done; done

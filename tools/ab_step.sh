#!/bin/bash
# Per-kernel times of one step (CUDA events between the kernels, plain launches; "softmin" = the fused weight-sum
# kernel, "finalize" = combine / filter / update / optimal trajectory as its own kernel) for the library variants in
# build/variants/*.so and the in-tree build, at the bench shape, the 8-GPU shard size and the latency shape; then the
# phase clocks of the debug build (-DMPPI_PHASE_PRINT), if there is one.
for shape in "1048576 100" "131072 100" "16384 50"; do
  set -- $shape
  for lib in in-tree build/variants/*.so; do
    case $lib in *phase.so) continue;; esac
    if [ "$lib" = in-tree ]; then pre="MPPI_UNUSED=1"; else pre="MPPI_B200_LIB=$PWD/$lib"; fi
    echo "K=$1 T=$2 lib=$lib $(env $pre python tools/profile_step.py --K $1 --T $2 --steps 24 --timing 2>&1 | grep "^{'steps'\|^ok" | tr '\n' ' ')"
  done
done
if [ -f build/variants/phase.so ]; then
  for K in 1048576 131072 16384; do
    echo "== phase clocks K=$K"; MPPI_B200_LIB=$PWD/build/variants/phase.so python tools/profile_step.py --K $K --T 100 --steps 3 2>&1 | grep "wsum last block\|finalize:" | tail -2
  done
fi

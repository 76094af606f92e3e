#!/bin/bash
# One GPU-box visit for an A/B round: GPU parity suite on the in-tree build, then the rollout-kernel time of the in-tree
# build and of every library in build/variants/ (tools/build_variants.py) at the shapes in SHAPES.
tag=${1:-ab}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_gputests.log
{
for shape in ${SHAPES:-"1048576,100" "131072,100" "16384,50"}; do
  K=${shape%,*}; T=${shape#*,}
  echo "== in-tree K=$K T=$T $(python tools/profile_step.py --K $K --T $T --steps 12 --timing 2>&1 | grep -o "'rollout': [0-9.]*\|^ok.*" | tr '\n' ' ')"
done
SHAPES="${SHAPES:-}" tools/ab_variants.sh
echo "== C5 in-tree $(python tools/profile_batched.py --steps 12 --timing 2>&1 | grep -o "'rollout': [0-9.]*")"
} > gpurun_out/${tag}_ab.txt 2>&1
tail -3 gpurun_out/${tag}_gputests.log; cat gpurun_out/${tag}_ab.txt

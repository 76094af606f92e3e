#!/bin/bash
# Round 2, visit a: parity tests of the new certified lookups + kernel timings of the shapes that matter.
tag=${1:-r2a}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_$tag.log
{
for s in certified full; do
  echo "== K=2^20 T=100 tracking $s"; python tools/profile_step.py --K 1048576 --T 100 --steps 12 --timing --search $s
  echo "== K=131072 T=100 tracking $s"; python tools/profile_step.py --K 131072 --T 100 --steps 12 --timing --search $s
  echo "== K=16384 T=50 tracking $s"; python tools/profile_step.py --K 16384 --T 50 --steps 12 --timing --search $s
done
echo "== K=2^20 T=100 rest certified"; python tools/profile_step.py --K 1048576 --T 100 --steps 12 --timing --state rest
} > gpurun_out/timing_$tag.log 2>&1
python bench.py --steps 20 --warmup 3 --no-injected --no-cpu > gpurun_out/bench_$tag.log 2>&1; echo rc=$? >> gpurun_out/bench_$tag.log
true

#!/bin/bash
# One GPU-box visit: parity tests, lookup-mode A/B, bench line, ncu capture of the rollout kernel, launch list.
tag=${1:-r2}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_$tag.log
bash tools/ab_search.sh > gpurun_out/ab_search_$tag.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.log 2>&1; echo rc=$? >> gpurun_out/bench_$tag.log
P="python tools/profile_step.py --K 1048576 --T 100 --steps 3"
$P > gpurun_out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mppi_rollout -s 2 -c 1 -f -o gpurun_out/prof_rollout_$tag $P > gpurun_out/ncu_$tag.log 2>&1
B="python bench.py --steps 2 --warmup 3 --no-latency --no-cpu --no-injected --no-batched --no-search"
$B > gpurun_out/plain_bench_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_bench_$tag.log 2>&1
true

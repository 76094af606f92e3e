#!/bin/bash
# One GPU-box visit for the profiles/ evidence of a build: every command runs plain first (exit 0), then under ncu.
tag=${1:-r2}
P="python tools/profile_step.py --K 1048576 --T 100 --steps 3"
$P > gpurun_out/${tag}_plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mppi_rollout -s 2 -c 1 -f -o gpurun_out/${tag}_prof_rollout_c4 $P > gpurun_out/${tag}_ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mppi_softmin_wsum -s 2 -c 1 -f -o gpurun_out/${tag}_prof_wsum_c4 $P > gpurun_out/${tag}_ncu_wsum_c4.log 2>&1
Q="python tools/profile_batched.py --steps 3"
$Q > gpurun_out/${tag}_plain_c5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mppi_rollout -s 2 -c 1 -f -o gpurun_out/${tag}_prof_rollout_c5 $Q > gpurun_out/${tag}_ncu_c5.log 2>&1
B="python bench.py --steps 2 --warmup 3 --no-latency --no-cpu --no-injected --no-batched --no-search"
$B > gpurun_out/${tag}_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_bench.log 2>&1
true

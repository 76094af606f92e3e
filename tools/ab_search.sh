#!/bin/bash
# A/B of the nearest-waypoint lookup modes at the bench shape (K=2^20, T=100) and the latency shape.
for st in tracking rest; do for se in certified full; do
  echo "== K=1048576 T=100 state=$st search=$se"
  python tools/profile_step.py --K 1048576 --T 100 --steps 12 --timing --state $st --search $se 2>&1 | tail -3
done; done
for se in certified full; do
  echo "== K=16384 T=50 state=tracking search=$se"
  python tools/profile_step.py --K 16384 --T 50 --steps 50 --timing --state tracking --search $se 2>&1 | tail -3
  echo "== K=131072 T=100 state=tracking search=$se (8-GPU shard size)"
  python tools/profile_step.py --K 131072 --T 100 --steps 20 --timing --state tracking --search $se 2>&1 | tail -3
done

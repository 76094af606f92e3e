"""A few steps of BASELINE.json configs[4] (1024 environments x K=1024, T=64, resident controller state) — the command
profiled under ncu for the batched rollout kernel (first run plain, as the profiling recipe requires).
    python tools/profile_batched.py [--envs 1024] [--steps 3]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=bench.C5_ENVS)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--timing", action="store_true", help="CUDA events between the kernels (plain launches)")
    a = ap.parse_args()
    from mppi_robotarm_b200.batched import BatchedMPPIController
    B, K, T = a.envs, bench.C5_K, bench.C5_T
    ref = bench.synthetic_ref_path()
    bat = BatchedMPPIController(B, **bench.run_py_kwargs(ref, K, T), seed=11, search_stats=True, use_graph=False)
    rows = (np.arange(B) * (1900 // B + 1)) % 1900
    th = 2 * np.pi * rows / (ref.shape[0] - 1)
    x, y = 0.8 + 0.6 * np.cos(th), 0.8 + 0.6 * np.sin(th)
    q2 = -np.arccos(np.clip((x * x + y * y - 2.0) / 2.0, -1, 1))
    q1 = np.arctan2(y, x) - np.arctan2(np.sin(q2), 1.0 + np.cos(q2))
    X = np.stack([q1, q2, np.zeros(B), np.zeros(B)], axis=1)
    bat.prev_waypoints_idx = rows.astype(np.int64)
    if a.timing:                                   # first launches carry module load: keep them out
        for _ in range(3):
            bat.calc_control_input(X)
        bat.engine.search_stats(reset=True)
        bat.engine.set_timing(True)
    for _ in range(a.steps):
        u0, _, _ = bat.calc_control_input(X)
    if a.timing:
        print(bat.engine.get_timing())
    print("search", bat.engine.search_stats())
    print("ok", float(np.abs(u0).max()))
    bat.close()


if __name__ == "__main__":
    main()

"""Many independent arm instances stepped together (BASELINE.json configs[4]): one launch of each kernel per control
step for all environments, controller state resident on the device, plant on the host (utils.Arm_Dynamic) or — with
--device-loop — on the GPU as well.

    python tools/export_ref_paths.py      # once
    python examples/run_batched.py [--envs 256] [--K 1024] [--T 64] [--ticks 200] [--device-loop]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mppi_robotarm_b200.batched import BatchedMPPIController  # noqa: E402
from utils import Arm_Dynamic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--K", type=int, default=1024)
    ap.add_argument("--T", type=int, default=64)
    ap.add_argument("--ticks", type=int, default=200)
    ap.add_argument("--device-loop", action="store_true")
    a = ap.parse_args()
    path = os.path.join(ROOT, "xydq_circle.txt")
    if not os.path.isfile(path):
        sys.exit("xydq_circle.txt not found: run `python tools/export_ref_paths.py` first")
    ref = np.loadtxt(path)[:, 0:4]
    dt = 0.003
    # every environment starts on the path, at its own waypoint, with the joint angles that reach it
    rows = (np.arange(a.envs) * 7) % 1500 + 100
    x, y = ref[rows, 0], ref[rows, 1]
    q2 = -np.arccos(np.clip((x * x + y * y - 2.0) / 2.0, -1, 1))
    q1 = np.arctan2(y, x) - np.arctan2(np.sin(q2), 1.0 + np.cos(q2))
    X = np.stack([q1, q2, np.zeros(a.envs), np.zeros(a.envs)], axis=1)
    bat = BatchedMPPIController(a.envs, delta_t=2 * dt, ref_path=ref, horizon_step_T=a.T, number_of_samples_K=a.K,
                                param_lambda=100.0, param_alpha=0.98, sigma=np.array([[20.0, 0.0], [0.0, 20.0]]),
                                stage_cost_weight=np.array([0.5, 0.5, 5.0, 5.0]),
                                terminal_cost_weight=np.array([5.0, 5.0, 50.0, 50.0]), seed=3)
    bat.prev_waypoints_idx = rows
    t0 = time.perf_counter()
    if a.device_loop:
        log, stop = bat.run_closed_loop(X, a.ticks, dt)
        X, ticks = log[-1, :, 0:4], int(min(a.ticks, stop.min()))
        idx = log[-1, :, 6].astype(np.int64)
    else:
        for _ in range(a.ticks):
            u0, _, _ = bat.calc_control_input(X)              # only X in, (u0, waypoint index) out
            for e in range(a.envs):                           # run.py:53-55 per environment
                X[e, 2:4] += dt * Arm_Dynamic(X[e, 0:2], X[e, 2:4], u0[e])
                X[e, 0:2] += dt * X[e, 2:4]
        ticks, idx = a.ticks, bat.last_waypoint_idx()
    wall = time.perf_counter() - t0
    ex = np.cos(X[:, 0]) + np.cos(X[:, 0] + X[:, 1]) - ref[idx, 0]
    ey = np.sin(X[:, 0]) + np.sin(X[:, 0] + X[:, 1]) - ref[idx, 1]
    print(f"{a.envs} environments x K={a.K}, T={a.T}: {ticks} ticks in {wall:.2f} s ({1e3 * wall / max(ticks, 1):.3f} ms per tick), "
          f"waypoints advanced by {np.mean(idx - rows):.0f} on average, mean tracking error {np.mean(np.hypot(ex, ey)):.4f} m")
    bat.close()


if __name__ == "__main__":
    main()

"""Headless closed loop: the loop of the reference's run.py (lines 48-59: controller, plant, state update)
around this repository's drop-in `control.MPPIControllerForPathTracking`, without the matplotlib part.
The reference's own run.py works unchanged as well (same module names, same data files, same call).

    python tools/export_ref_paths.py      # once: writes xydq_circle.txt etc. next to control.py
    python examples/run_closed_loop.py [--ticks 1500] [--K 100] [--T 30] [--device-loop]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from control import MPPIControllerForPathTracking  # noqa: E402
from utils import Arm_Dynamic, Forward_Kinemetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ticks", type=int, default=1500)
    ap.add_argument("--K", type=int, default=100)
    ap.add_argument("--T", type=int, default=30)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--device-loop", action="store_true", help="run controller AND plant on the GPU (no host round trip per tick)")
    a = ap.parse_args()
    path = os.path.join(ROOT, "xydq_circle.txt")
    if not os.path.isfile(path):
        sys.exit("xydq_circle.txt not found: run `python tools/export_ref_paths.py` first")
    ref_path = np.loadtxt(path)[:, 0:4]                                    # run.py:18-19
    dt = 0.003
    q = np.array([1.152198236517471885e+00, -1.266101672070702344e+00])    # run.py:14
    dq = np.array([0.0, 0.0])
    mppi = MPPIControllerForPathTracking(                                  # run.py:25-37
        delta_t=dt * 2, ref_path=ref_path, horizon_step_T=a.T, number_of_samples_K=a.K, param_exploration=0.0,
        param_lambda=100.0, param_alpha=0.98, sigma=np.array([[20.0, 0.0], [0.0, 20.0]]),
        stage_cost_weight=np.array([0.50, 0.50, 5.0, 5.0]), terminal_cost_weight=np.array([5.0, 5.0, 50.0, 50.0]),
        seed=a.seed, verbose=False)
    t0 = time.perf_counter()
    err = []
    try:
        if a.device_loop:
            out = mppi.run_closed_loop(np.concatenate([q, dq]), a.ticks, dt)
            n = out["ticks"]
            for s, i in zip(out["state"], out["waypoint_idx"]):
                _, _, x2, y2 = Forward_Kinemetic(s[0:2])
                err.append(np.hypot(x2 - ref_path[i, 0], y2 - ref_path[i, 1]))
        else:
            state = np.concatenate([q, dq])
            for n in range(1, a.ticks + 1):
                u, _, _, _ = mppi.calc_control_input(observed_x=state)     # run.py:49-51
                dq += dt * Arm_Dynamic(q, dq, u)                           # run.py:53-55
                q += dt * dq
                _, _, x2, y2 = Forward_Kinemetic(q)
                state = np.concatenate((q, dq))
                i = mppi.prev_waypoints_idx
                err.append(np.hypot(x2 - ref_path[i, 0], y2 - ref_path[i, 1]))
    except IndexError:
        pass                                                               # end of the path (control.py:76-78)
    wall = time.perf_counter() - t0
    print(f"{len(err)} ticks in {wall:.2f} s ({1e3 * wall / max(len(err), 1):.3f} ms per tick), waypoint "
          f"{mppi.prev_waypoints_idx} of {ref_path.shape[0]}, mean tracking error {np.mean(err):.4f} m")
    mppi.close()


if __name__ == "__main__":
    main()

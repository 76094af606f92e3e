"""Randomised sweep of the certified lookups on the GPU: for many (path file, window start, arm state, horizon, noise
scale) combinations the per-sample costs with search="certified" must be the very floats of search="full".
    python tools/stress_lookup.py [--cases 300] [--seed 0]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.golden import cases  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=300)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    from mppi_robotarm_b200 import MppiEngine
    from mppi_robotarm_b200.arm_params import SYS_PARAMS
    rng = np.random.default_rng(a.seed)
    paths = cases.load_paths()
    refs = {"xydq_circle": cases.ref_path_for(paths, "xydq_circle.txt"), "xydq": cases.ref_path_for(paths, "xydq.txt"),
            "trajectory(xydq layout)": cases.ref_path_for(paths, "trajectory.txt", "xydq"),
            "trajectory1(verbatim)": cases.ref_path_for(paths, "trajectory1.txt")}
    traj1 = paths["trajectory1"]
    totals = {"lookups": 0, "certified": 0, "triples": 0}
    bad = 0
    for name, ref in refs.items():
        for T in (20, 64, 100):
            K = 4096
            eng = {m: MppiEngine(K=K, T=T, delta_t=0.006, param_lambda=100.0, param_gamma=2.0, sigma=np.eye(2) * 20.0,
                                 stage_cost_weight=[0.5, 0.5, 5, 5], terminal_cost_weight=[5, 5, 50, 50],
                                 arm_params=SYS_PARAMS(), ref_path=ref, seed=7, search=m, search_stats=True,
                                 optimal_traj=False) for m in ("certified", "full")}
            for _ in range(max(1, a.cases // 12)):
                p = int(rng.integers(0, ref.shape[0] - 2))
                # an arm near waypoint p (joint angles of the recorded run where the path is the circle, else IK), perturbed
                x, y = ref[p, 0], ref[p, 1]
                r2 = min(x * x + y * y, 3.99)
                q2 = -np.arccos(np.clip((r2 - 2.0) / 2.0, -1, 1))
                q1 = np.arctan2(y, x) - np.arctan2(np.sin(q2), 1.0 + np.cos(q2))
                x0 = np.array([q1, q2, 0, 0]) + rng.normal(0, 10.0 ** rng.uniform(-4, -0.5), 4) * np.array([1, 1, 10, 10])
                u = rng.normal(0, 10.0 ** rng.uniform(-1, 1.3), (T, 2)) + np.array([10.0, -2.0])
                prev = max(0, p - int(rng.integers(0, 25)))
                S = {}
                for m, e in eng.items():
                    e.step_counter = 0
                    e.step(x0, u, prev, None)
                    S[m] = e.last_costs()[0].cpu().numpy().copy()
                if not np.array_equal(S["certified"], S["full"]):
                    bad += 1
                    print("MISMATCH", name, T, p, prev, x0)
            st = eng["certified"].search_stats()
            for k in totals:
                totals[k] += st[k]
            for e in eng.values():
                e.close()
    n = totals["lookups"]
    print(f"cases {12 * max(1, a.cases // 12)}, mismatches {bad}; warp-lookups {n}: end tests {totals['certified'] / n:.3f}, "
          f"triples {totals['triples'] / n:.3f}, searched {(n - totals['certified'] - totals['triples']) / n:.4f}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()

"""Run a few MPPI steps of a given shape — the command that is profiled under ncu (and first run
plain, as the profiling recipe requires).   python tools/profile_step.py --K 131072 --T 100 --steps 3"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=131072)
    ap.add_argument("--T", type=int, default=100)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--noise", default="philox", choices=["philox", "injected"])
    ap.add_argument("--timing", action="store_true")
    ap.add_argument("--lam", type=float, default=100.0, help="param_lambda (1e9: every weight non-zero)")
    ap.add_argument("--search", default="certified", choices=["certified", "full"])
    ap.add_argument("--state", default="tracking", choices=["tracking", "rest"],
                    help="tracking: bench.bench_workload() (tick 500 of the reference's closed loop); "
                         "rest: arm at rest at the start of the synthetic path")
    a = ap.parse_args()
    from mppi_robotarm_b200 import MPPIControllerForPathTracking
    if a.state == "tracking":
        ref, x0, u, p0, _ = bench.bench_workload(a.T)
    else:
        ref, x0, u, p0 = bench.synthetic_ref_path(), bench.X0, np.tile([10.0, -2.0], (a.T, 1)), 0
    kw = bench.run_py_kwargs(ref, a.K, a.T)
    kw["param_lambda"] = a.lam
    ctrl = MPPIControllerForPathTracking(**kw, noise="philox", seed=1, verbose=False, use_graph=False,
                                         search=a.search, search_stats=True)
    eng = ctrl._engine()
    eps = eng.philox_noise(step=0) if a.noise == "injected" else None
    if a.timing:                                   # first launches carry module load and graph set-up: keep them out
        for _ in range(3):
            eng.step(x0, u, p0, eps)
        eng.set_timing(True)                       # (the timed path launches the final stage as its own kernel: warm that too)
        for _ in range(2):
            eng.step(x0, u, p0, eps)
        eng.search_stats(reset=True)
    eng.set_timing(a.timing)
    for _ in range(a.steps):
        eng.step(x0, u, p0, eps)
    if a.timing:
        print(eng.get_timing())
    print("search", eng.search_stats())
    print("ok", float(eng.out_rho[0]), float(np.abs(eng.out_u_new).max()))
    ctrl.close()


if __name__ == "__main__":
    main()

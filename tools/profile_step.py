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
    a = ap.parse_args()
    from mppi_robotarm_b200 import MPPIControllerForPathTracking
    kw = bench.run_py_kwargs(bench.synthetic_ref_path(), a.K, a.T)
    kw["param_lambda"] = a.lam
    ctrl = MPPIControllerForPathTracking(**kw, noise="philox",
                                         seed=1, verbose=False, use_graph=False)
    eng = ctrl._engine()
    eps = eng.philox_noise(step=0) if a.noise == "injected" else None
    eng.set_timing(a.timing)
    u = ctrl.u_prev.copy()
    for _ in range(a.steps):
        eng.step(bench.X0, u, 0, eps)
    if a.timing:
        print(eng.get_timing())
    print("ok", float(eng.out_rho[0]), float(np.abs(eng.out_u_new).max()))
    ctrl.close()


if __name__ == "__main__":
    main()

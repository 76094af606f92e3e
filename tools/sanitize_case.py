"""Smallest run that touches every kernel once — the target of `compute-sanitizer --tool memcheck`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mppi_robotarm_b200 import MPPIControllerForPathTracking  # noqa: E402
from mppi_robotarm_b200.batched import BatchedMPPIController  # noqa: E402

ref = bench.synthetic_ref_path()
for K, T in ((1000, 30), (333, 7)):
    kw = bench.run_py_kwargs(ref, K, T)
    c = MPPIControllerForPathTracking(**kw, seed=1, verbose=False, visualze_sampled_trajs=True)
    c.calc_control_input(bench.X0)                       # philox: prepare, rollout, fused weight-sum, finalize, sampled traj
    eps = c._engine().philox_noise(step=0)
    c._engine().step(bench.X0, c.u_prev, 0, eps)         # injected: softmin, wsum_injected, reduce
    c.run_closed_loop(bench.X0, 3, 0.003)                # plant tick
    c.close()
kw = bench.run_py_kwargs(ref, 200, 16)
b = BatchedMPPIController(3, **kw, visualize_optimal_traj=True, seed=2)
b.calc_control_input(np.tile(bench.X0, (3, 1)))
b.close()
print("SANITIZE_CASE_OK")

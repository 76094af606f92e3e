"""CPU test of bench.py's reference arm (`--impl reference`): it must run without a GPU, time the FP64
CPU restatement on a bounded sample and print one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_a_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                          "--steps", "2", "--warmup", "3"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mppi_sample_steps_per_s" and d["unit"] == "sample-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2
    assert d["value"] > 1e6                                   # the C restatement does >1e7/s per core
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "sample-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "K=1048576" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_bench_workload_is_the_reference_closed_loop_tick():
    """bench.py times the C4 step on a state of the reference's own closed loop (committed fixtures); the
    nominal sequence is that tick's shifted sequence extended to the bench horizon."""
    import numpy as np
    import bench
    ref, x0, u, p, desc = bench.bench_workload()
    assert ref.shape == (2000, 4) and x0.shape == (4,) and u.shape == (bench.T_HORIZON, 2)
    assert p == 528 and "tick 500" in desc
    with np.load(bench.ROOT + "/tests/golden/closed_loop_c1.npz") as z:
        np.testing.assert_array_equal(x0, z["state"][500])
        np.testing.assert_array_equal(u[:29], z["u_new"][499][1:])
        np.testing.assert_array_equal(u[29:], np.repeat(z["u_new"][499][-1:], bench.T_HORIZON - 29, axis=0))
    # the end effector of that state sits next to the waypoint the window starts at
    x, y = np.cos(x0[0]) + np.cos(x0[0] + x0[1]), np.sin(x0[0]) + np.sin(x0[0] + x0[1])
    assert np.hypot(x - ref[p, 0], y - ref[p, 1]) < 0.05
    u30 = bench.bench_workload(30)[2]
    assert u30.shape == (30, 2)

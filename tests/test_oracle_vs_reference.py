"""Pins the oracle against the live, unmodified reference (only where /root/reference exists — the
build container).  On the GPU box these tests skip and the committed goldens carry the pin."""
import numpy as np
import pytest

from oracle import mppi_oracle as mo
from oracle import ref_harness as rh
from tests.golden import cases

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference checkout not present")


def test_committed_data_equals_reference_files(paths):
    for name, arr in paths.items():
        np.testing.assert_array_equal(arr, rh.load_reference_data(name + ".txt"))


@pytest.mark.parametrize("seed", [31, 32])
def test_fresh_random_step_matches_reference(paths, seed):
    rng = np.random.default_rng(seed)
    K, T = int(rng.integers(20, 90)), int(rng.integers(5, 40))
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    kw = cases.run_py_kwargs(ref, K, T, param_exploration=float(rng.choice([0.0, 0.2])),
                             visualze_sampled_trajs=True)
    probe = rh.ReferenceProbe(**kw)
    c = mo.OracleMPPI(**kw)
    p0 = int(rng.integers(0, 1900))
    probe.ctrl.prev_waypoints_idx = p0
    c.prev_waypoints_idx = p0
    x = np.array(cases.X0) + rng.normal(0, 0.05, 4)
    for s in range(2):
        eps = mo.injected_noise(seed * 10 + s, K, T, kw["sigma"]).astype(np.float64)
        r = probe.step(list(x), eps)
        o = mo.step_vectorized(c, x, eps)
        for k in ["S", "w", "w_eps_raw", "w_eps_filt", "u_new", "u0", "optimal_traj", "sampled_traj"]:
            a, b = np.asarray(r[k]), np.asarray(o[k])
            assert np.max(np.abs(a - b)) <= 1e-12 * max(1.0, np.max(np.abs(a))), k
        assert r["prev_idx_after"] == o["prev_idx_after"]
        np.testing.assert_array_equal(probe.ctrl.u_prev, c.u_prev)


@pytest.mark.parametrize("seed", [41, 42])
def test_alternative_model_and_smoother_match_reference(paths, seed):
    """The reference's dead alternatives, live: its own _F1 (control.py:265-295) routed into every self._F call and
    its own _moving_average_filter (control.py:329-344) routed into the smoother call, on fresh random steps."""
    rng = np.random.default_rng(seed)
    K, T = int(rng.integers(20, 70)), int(rng.integers(12, 36))
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    kw = cases.run_py_kwargs(ref, K, T, param_lambda=3.0e4, visualze_sampled_trajs=True)
    probe = rh.ReferenceProbe(**kw)
    probe.ctrl._F = probe.ctrl._F1
    ref_avg = probe.ctrl._moving_average_filter

    def avg(xx, window_size):
        probe.last["w_eps_raw"] = np.array(xx, copy=True)
        out = ref_avg(xx, window_size)
        probe.last["w_eps_filt"] = np.array(out, copy=True)
        return out
    probe.ctrl._moving_median_filter = avg
    c = mo.OracleMPPI(**kw, dynamics="F1", smoother="average")
    p0 = int(rng.integers(0, 1900))
    probe.ctrl.prev_waypoints_idx = p0
    c.prev_waypoints_idx = p0
    x = np.array(cases.X0) + rng.normal(0, 0.05, 4)
    for s in range(2):
        eps = mo.injected_noise(seed * 10 + s, K, T, kw["sigma"]).astype(np.float64)
        r = probe.step(list(x), eps)
        o = mo.step_vectorized(c, x, eps)
        for k in ["S", "w", "w_eps_raw", "w_eps_filt", "u_new", "u0", "optimal_traj", "sampled_traj"]:
            a, b = np.asarray(r[k]), np.asarray(o[k])
            assert np.max(np.abs(a - b)) <= 1e-12 * max(1.0, np.max(np.abs(a))), k
        assert r["prev_idx_after"] == o["prev_idx_after"]


def test_plant_twin_matches_reference_utils():
    _, utils = rh.import_reference()
    rng = np.random.default_rng(5)
    for _ in range(20):
        q, dq, u = rng.normal(0, 1, 2), rng.normal(0, 1, 2), rng.normal(0, 10, 2)
        a = utils.Arm_Dynamic(q, dq, u)
        b = mo.arm_accel(q[0], q[1], dq[0], dq[1], u[0], u[1], mo.default_arm_params())
        np.testing.assert_allclose(a, b, rtol=1e-13, atol=1e-13)


def test_reference_run_py_drives_our_control_module_unchanged(tmp_path, monkeypatch):
    """The reference's own run.py, byte for byte, with THIS repo's control.py / utils.py / sys_params.py
    ahead of it on sys.path (and a no-op matplotlib): it must import, load its data file, construct the
    controller with its keywords and reach the first calc_control_input call.  Without a GPU that call
    raises NativeLibraryError (no CPU fallback) — which is the point where the drop-in boundary ends; on a
    GPU box the same loop is covered by tests/test_gpu_features.py."""
    import os
    import runpy
    import shutil
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-side check of the drop-in boundary")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shutil.copy(os.path.join(rh.REFERENCE_DIR, "xydq_circle.txt"), tmp_path / "xydq_circle.txt")
    monkeypatch.chdir(tmp_path)
    rh._install_stubs()
    for name in ("control", "utils", "sys_params"):
        monkeypatch.delitem(sys.modules, name, raising=False)
    monkeypatch.syspath_prepend(root)
    from mppi_robotarm_b200 import _cabi
    with pytest.raises(_cabi.NativeLibraryError):
        runpy.run_path(os.path.join(rh.REFERENCE_DIR, "run.py"), run_name="__main__")
    import control
    assert control.MPPIControllerForPathTracking.__module__ == "mppi_robotarm_b200.controller"


def test_exported_data_files_are_byte_identical(tmp_path):
    """tools/export_ref_paths.py (run by build()) writes the four .txt files a fresh checkout needs for
    `np.loadtxt('xydq_circle.txt')` (run.py:18): the same bytes as the reference's own files."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import export_ref_paths
    for path in export_ref_paths.export(str(tmp_path)):
        with open(path, "rb") as a, open(os.path.join(rh.REFERENCE_DIR, os.path.basename(path)), "rb") as b:
            assert a.read() == b.read(), path


def test_remaining_utils_helpers_match_the_reference():
    """`from utils import *` (run.py:5) also brings Inverse_Kinemetic, Feedback_linearization and Controller
    (utils.py:41-93; never called by run.py): same names, arguments and values here."""
    import utils
    _, ref_utils = rh.import_reference()
    rng = np.random.default_rng(3)
    for th in list(rng.uniform(0, 6.0, 20)) + [6.1, 6.2, 6.3, 6.48, 6.6, 7.0]:
        a, b = utils.Inverse_Kinemetic(th), ref_utils.Inverse_Kinemetic(th)
        np.testing.assert_allclose(a[0], b[0], rtol=0, atol=1e-13)
        assert a[1] == b[1] and a[2] == b[2]
    for _ in range(20):
        q, dq, v, r, dr = (rng.normal(0, 1, 2) for _ in range(5))
        np.testing.assert_allclose(utils.Feedback_linearization(q, dq, v), ref_utils.Feedback_linearization(q, dq, v),
                                   rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(utils.Controller(q, dq, r, dr, v), ref_utils.Controller(q, dq, r, dr, v),
                                   rtol=1e-13, atol=1e-13)

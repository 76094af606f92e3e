"""CPU tests of the host-side logic around the CUDA step."""
import numpy as np
import pytest

from mppi_robotarm_b200 import load_ref_path
from mppi_robotarm_b200.engine import ShardSpec, exploit_count
from oracle import mppi_oracle as mo
from tests.golden import cases


def test_constructor_mirrors_reference_attributes(paths):
    from control import MPPIControllerForPathTracking
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    c = MPPIControllerForPathTracking(**cases.run_py_kwargs(ref, 100, 30), visualze_sampled_trajs=True)
    assert (c.dim_x, c.dim_u, c.T, c.K) == (4, 2, 30, 100)
    assert c.param_gamma == pytest.approx(100.0 * (1 - 0.98))
    assert c.l1 == 1 and c.l2 == 1 and c.prev_waypoints_idx == 0
    np.testing.assert_array_equal(c.u_prev, np.tile([10.0, -2.0], (30, 1)))
    assert c.visualze_sampled_trajs is True and c.visualize_optimal_traj is True
    assert c.ref_path is ref and c.delta_t == 0.006
    # defaults of control.py:21-35
    d = MPPIControllerForPathTracking()
    assert (d.T, d.K, d.param_lambda, d.param_alpha, d.delta_t) == (20, 500, 50.0, 1.0, 0.01)
    np.testing.assert_array_equal(d.Sigma, [[10.0, 10.0], [100.0, 100.0]])


def test_calc_epsilon_seam_and_sigma_check(paths, capsys):
    from control import MPPIControllerForPathTracking
    c = MPPIControllerForPathTracking(**cases.run_py_kwargs(cases.ref_path_for(paths, "xydq_circle.txt"), 16, 5))
    np.random.seed(0)
    e = c._calc_epsilon(c.Sigma, 16, 5, 2)
    np.random.seed(0)
    np.testing.assert_array_equal(e, np.random.multivariate_normal(np.zeros(2), c.Sigma, (16, 5)))
    with pytest.raises(ValueError):
        c._calc_epsilon(np.eye(3), 16, 5, 2)
    assert "sigma must be a square matrix" in capsys.readouterr().out


def test_host_nearest_waypoint_helper_matches_oracle(paths):
    from control import MPPIControllerForPathTracking
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    c = MPPIControllerForPathTracking(**cases.run_py_kwargs(ref, 16, 5), verbose=False)
    o = mo.OracleMPPI(**cases.run_py_kwargs(ref, 16, 5))
    for p, q in ((0, cases.X0[:2]), (700, [0.9, 0.6]), (1990, [1.15, -1.26])):
        c.prev_waypoints_idx = o.prev_waypoints_idx = p
        idx, rx, ry, r1, r2 = c._get_nearest_waypoint(q[0], q[1])
        assert idx == p + int(mo.nearest_in_window(mo.window_of(ref, p), *mo.end_effector(q[0], q[1])))
        np.testing.assert_array_equal([rx, ry, r1, r2], ref[idx])
        assert c.prev_waypoints_idx == p
        c._get_nearest_waypoint(q[0], q[1], update_prev_idx=True)
        assert c.prev_waypoints_idx == mo.update_waypoint(o, q[0], q[1])


def test_exploit_count_is_the_python_comparison():
    for K in (1, 7, 50, 64, 100, 4096, 1 << 20):
        for ex in (0.0, 0.1, 0.25, 0.33, 0.5, 0.999, 1.0):
            ks = np.arange(K)
            assert exploit_count(K, ex) == int(np.count_nonzero(ks < (1.0 - ex) * K))


def test_shard_bounds_partition_the_samples():
    for K in (7, 100, 4096, 1 << 20, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [ShardSpec(r, world).bounds(K) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == K
            for (o1, n1), (o2, _) in zip(spans, spans[1:]):
                assert o1 + n1 == o2
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1


def test_load_ref_path_layouts(tmp_path, paths):
    p = tmp_path / "xydq_circle.txt"
    np.savetxt(p, paths["xydq_circle"], fmt="%.18e")
    np.testing.assert_array_equal(load_ref_path(p), paths["xydq_circle"][:, 0:4])     # run.py:18-19
    t = tmp_path / "trajectory.txt"
    np.savetxt(t, paths["trajectory"], fmt="%.18e")
    np.testing.assert_array_equal(load_ref_path(t, "verbatim"), paths["trajectory"][:, 0:4])
    conv = load_ref_path(t)                     # auto -> (x, y, dq1, dq2)
    np.testing.assert_array_equal(conv[:, 0:2], paths["trajectory"][:, 2:4])
    np.testing.assert_allclose(conv, cases.relayout_qxy(paths["trajectory"]))
    with pytest.raises(ValueError):
        load_ref_path(t, "nonsense")


def test_plant_helpers_match_oracle_twin():
    import utils
    rng = np.random.default_rng(2)
    for _ in range(10):
        q, dq, u = rng.normal(0, 1, 2), rng.normal(0, 1, 2), rng.normal(0, 10, 2)
        np.testing.assert_allclose(utils.Arm_Dynamic(q, dq, u),
                                   mo.arm_accel(q[0], q[1], dq[0], dq[1], u[0], u[1], mo.default_arm_params()),
                                   rtol=1e-13, atol=1e-13)
        x1, y1, x2, y2 = utils.Forward_Kinemetic(q)
        np.testing.assert_allclose((x2, y2), mo.end_effector(q[0], q[1]), rtol=1e-14)
        np.testing.assert_allclose((x1, y1), (np.cos(q[0]), np.sin(q[0])), rtol=1e-14)
    import sys_params
    assert sys_params.SYS_PARAMS() == mo.default_arm_params()


def test_refgen_reproduces_trajectory_txt(paths):
    from mppi_robotarm_b200 import refgen
    gen = refgen.circle_joint_reference(3000)
    np.testing.assert_allclose(gen, paths["trajectory"], rtol=0, atol=2e-12)      # the reference's own file
    # forward kinematics of the generated joints lands on the generated end-effector
    x, y = mo.end_effector(gen[:, 0], gen[:, 1])
    np.testing.assert_allclose(np.stack([x, y], 1), gen[:, 2:4], atol=1e-12)


def test_refgen_tracking_run_has_the_shape_of_xydq_circle(paths):
    from mppi_robotarm_b200 import refgen
    rec = refgen.record_tracking_run(2000)
    ref = paths["xydq_circle"]
    assert rec.shape == ref.shape
    rad = np.hypot(rec[:, 0] - 0.8, rec[:, 1] - 0.8)
    assert 0.57 <= rad.min() and rad.max() <= 0.63
    np.testing.assert_allclose(rec[0, 0:2], [1.4, 0.8], atol=5e-3)
    # same order of magnitude of joint rates and torques as the recorded file
    assert 0.3 <= np.abs(rec[:, 2:4]).max() / np.abs(ref[:, 2:4]).max() <= 3.0
    assert 0.3 <= np.abs(rec[:, 4:6]).max() / np.abs(ref[:, 4:6]).max() <= 3.0


def test_noise_factor_handles_definite_semidefinite_and_rejects_asymmetric():
    """Factor of the in-kernel draw eps = L z (engine.noise_factor): Cholesky for a positive definite Sigma, an
    eigen-factor made lower triangular for a semi-definite one (np.random.multivariate_normal, control.py:163,
    accepts those), LinAlgError for an asymmetric or indefinite matrix."""
    from mppi_robotarm_b200.engine import noise_factor
    for sig in (np.array([[20.0, 0.0], [0.0, 20.0]]), np.array([[20.0, 6.0], [6.0, 10.0]]),
                np.array([[4.0, 2.0], [2.0, 1.0]]), np.array([[0.0, 0.0], [0.0, 9.0]])):
        L = noise_factor(sig)
        np.testing.assert_allclose(L @ L.T, sig, atol=1e-12)
        assert L[0, 1] == 0.0 or abs(L[0, 1]) < 1e-12
    with pytest.raises(np.linalg.LinAlgError):
        noise_factor(np.array([[10.0, 10.0], [100.0, 100.0]]))       # the reference's (unusable) default Sigma: asymmetric
    with pytest.raises(np.linalg.LinAlgError):
        noise_factor(np.array([[1.0, 2.0], [2.0, 1.0]]))             # indefinite


def test_median_network_of_the_final_stage_sorts_ten_inputs():
    """The final stage takes the rank-5 element of each 10-sample window (scipy.ndimage.median_filter(size=10),
    control.py:319-327) from a sorting network written out in the kernel source: parse the exchanges from that source
    and check them by the 0-1 principle (a network that sorts every 0/1 input sorts every input)."""
    import itertools
    import os
    import re
    src = open(os.path.join(os.path.dirname(__file__), "..", "mppi_robotarm_b200", "csrc", "mppi_kernels.cuh")).read()
    body = src[src.index("#define MPPI_CE(a, b)"):src.index("#undef MPPI_CE")]
    ces = [(int(a), int(b)) for a, b in re.findall(r"MPPI_CE\((\d+), (\d+)\)", body)]
    assert len(ces) == 29 and all(0 <= a < b < 10 for a, b in ces)
    for bits in itertools.product((0, 1), repeat=10):
        v = list(bits)
        for a, b in ces:
            if v[a] > v[b]:
                v[a], v[b] = v[b], v[a]
        assert v == sorted(v)
    assert "med = win[kFilter / 2];" in src

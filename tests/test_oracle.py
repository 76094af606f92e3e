"""CPU tests: the FP64 oracle (oracle/mppi_oracle.py) against the golden vectors produced by the
unmodified reference (tests/golden/make_golden.py), plus the oracle's building blocks."""
import numpy as np
import pytest

from oracle import mppi_oracle as mo
from tests.golden import cases

GOLD = cases.load_golden("single_steps.npz")
KEYS = ["S", "w", "w_eps_raw", "w_eps_filt", "u_new", "u0", "optimal_traj"]


def _run(case, paths, fn):
    kw = cases.ctor_kwargs(case, paths)
    c = mo.OracleMPPI(**kw, smoother=case.get("smoother", "median"), dynamics=case.get("dynamics", "F"))
    if "prev_idx" in case:
        c.prev_waypoints_idx = case["prev_idx"]
    if "u_prev" in case:
        c.u_prev = np.array(case["u_prev"], dtype=np.float64)
    g = GOLD[case["name"]]
    for s in range(case.get("steps", 1)):
        eps = mo.injected_noise(case["seed"] + s, case["K"], case["T"], kw["sigma"])
        e64 = eps.astype(np.float64)
        # the generator stream must be the one the goldens were made with
        np.testing.assert_allclose([e64.sum(), np.abs(e64).sum()], g[f"eps_sum.{s}"], rtol=1e-12)
        np.testing.assert_array_equal(c.u_prev, g[f"u_prev_before.{s}"]) if s == 0 else None
        out = fn(c, g[f"x0.{s}"], e64)
        assert [out["prev_idx_before"], out["prev_idx_after"]] == list(g[f"prev_idx.{s}"])
        for k in KEYS:
            ref = g[f"{k}.{s}"]
            scale = max(np.max(np.abs(ref)), 1e-300)
            err = np.max(np.abs(np.asarray(out[k]) - ref)) / scale
            assert err <= 1e-12, (case["name"], s, k, err)
        if kw.get("visualze_sampled_trajs"):
            np.testing.assert_allclose(out["sampled_traj"], g[f"sampled_traj.{s}"], rtol=0, atol=1e-12)
        # quirk Q2: returned u0 is the first row of the *shifted* sequence
        np.testing.assert_array_equal(out["u0"], c.u_prev[0])
        assert out["u_seq_returned"] is c.u_prev          # quirk Q1 (alias)


@pytest.mark.parametrize("name", sorted(GOLD))
def test_vectorized_oracle_matches_reference_goldens(name, paths):
    case = {c["name"]: c for c in cases.single_cases(paths)}[name]
    _run(case, paths, mo.step_vectorized)


@pytest.mark.parametrize("name", ["c1_seed1", "explore_odd", "end_of_path_2", "full_sigma", "c1_viz"])
def test_loop_oracle_matches_reference_goldens(name, paths):
    case = {c["name"]: c for c in cases.single_cases(paths)}[name]
    _run(case, paths, mo.step_loops)


def test_c2_costs_match_reference(paths):
    """Config 2 (K=4096, T=50, trajectory.txt): per-sample costs and the update."""
    gold = cases.load_golden("c2_steps.npz")
    for case in cases.c2_cases():
        kw = cases.ctor_kwargs(case, paths)
        c = mo.OracleMPPI(**kw)
        g = gold[case["name"]]
        eps = mo.injected_noise(case["seed"], case["K"], case["T"], kw["sigma"]).astype(np.float64)
        out = mo.step_vectorized(c, g["x0.0"], eps)
        for k in ["S", "w_eps_raw", "w_eps_filt", "u_new", "u0", "optimal_traj"]:
            ref = g[f"{k}.0"]
            err = np.max(np.abs(np.asarray(out[k]) - ref)) / np.max(np.abs(ref))
            assert err <= 1e-11, (case["name"], k, err)


def test_closed_loop_replay_teacher_forced(paths):
    """Replay the reference's 1500-step closed loop step by step (teacher forced: every step starts
    from the reference's recorded state / u_prev / prev_idx) and compare the update."""
    with np.load(cases.HERE + "/closed_loop_c1.npz") as z:
        cl = {k: z[k] for k in z.files}
    K, T, seed0, _ = (int(v) for v in cl["meta"])
    kw = cases.run_py_kwargs(cases.ref_path_for(paths, "xydq_circle.txt"), K, T,
                             visualize_optimal_traj=False)
    n = cl["state"].shape[0]
    assert n == 1500                      # the reference completes run.py's loop without IndexError
    c = mo.OracleMPPI(**kw)
    worst = 0.0
    for s in list(range(0, 40)) + list(range(40, n, 37)):
        if s > 0:
            prev = cl["u_new"][s - 1]
            c.u_prev = np.concatenate([prev[1:], prev[-1:]], axis=0)
        c.prev_waypoints_idx = int(cl["prev_idx"][s, 0])
        eps = mo.injected_noise(seed0 + s, K, T, kw["sigma"]).astype(np.float64)
        out = mo.step_vectorized(c, cl["state"][s], eps)
        assert out["prev_idx_after"] == cl["prev_idx"][s, 1]
        worst = max(worst, np.max(np.abs(out["u_new"] - cl["u_new"][s])) / np.max(np.abs(cl["u_new"][s])))
        np.testing.assert_allclose(out["u0"], cl["u0"][s], rtol=1e-9, atol=1e-9)
    assert worst <= 1e-9, worst


def test_median_filter_matches_scipy():
    from scipy.ndimage import median_filter
    rng = np.random.default_rng(3)
    for n in [1, 2, 3, 7, 9, 10, 11, 30, 50, 64, 100]:
        x = rng.standard_normal(n)
        np.testing.assert_array_equal(median_filter(x, size=10, mode="reflect"),
                                      mo.median_filter_reflect(x, 10))
    # ties and constant input
    x = np.repeat(np.arange(5.0), 6)
    np.testing.assert_array_equal(median_filter(x, size=10, mode="reflect"), mo.median_filter_reflect(x, 10))


def test_first_argmin_tie_break():
    win = np.array([[0.0, 0.0, 1, 1], [2.0, 0.0, 2, 2], [0.0, 0.0, 3, 3]])
    assert int(mo.nearest_in_window(win, 0.0, 0.0)) == 0        # duplicates -> first
    assert int(mo.nearest_in_window(win, 1.0, 0.0)) == 0        # exact tie between 0 and 1 -> first


def test_window_truncates_at_end_of_path(paths):
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    assert mo.window_of(ref, 1985).shape[0] == 15
    assert mo.window_of(ref, 0).shape[0] == 30


def test_end_of_path_raises_index_error(paths, capsys):
    kw = cases.run_py_kwargs(cases.ref_path_for(paths, "xydq_circle.txt"), 8, 5)
    c = mo.OracleMPPI(**kw)
    c.prev_waypoints_idx = 1999
    with pytest.raises(IndexError):
        mo.step_vectorized(c, cases.X0, np.zeros((8, 5, 2)))
    assert "Reached the end of the reference path" in capsys.readouterr().out


def test_bad_sigma_raises_value_error(paths):
    kw = cases.run_py_kwargs(cases.ref_path_for(paths, "xydq_circle.txt"), 8, 5, sigma=np.eye(3))
    c = mo.OracleMPPI(**kw)
    with pytest.raises(ValueError):
        mo.step_vectorized(c, cases.X0, np.zeros((8, 5, 2)))


def test_exploit_count_matches_python_comparison():
    for K in (1, 7, 50, 64, 100, 4096):
        for ex in (0.0, 0.25, 0.33, 0.5, 0.999, 1.0):
            assert mo.exploit_count(K, ex) == sum(1 for k in range(K) if k < (1.0 - ex) * K)


def test_joint_limit_cost_python_and_c_oracle_agree(paths):
    """The joint-limit stage cost is an extension (weight 0 = the reference): both restatements define it
    the same way, it vanishes inside the limits, and it grows quadratically with the violation."""
    from oracle import c_oracle
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    K, T = 256, 30
    kw = cases.run_py_kwargs(ref, K, T)
    eps = mo.injected_noise(5, K, T, kw["sigma"]).astype(np.float64)
    base = mo.rollout_costs(mo.OracleMPPI(**kw), np.array(cases.X0), eps, prev_idx=0)
    lim = dict(joint_limit_lo=(1.10, -1.30), joint_limit_hi=(1.20, -1.20), joint_limit_weight=3.0)
    c = mo.OracleMPPI(**kw, **lim)
    S_py = mo.rollout_costs(c, np.array(cases.X0), eps, prev_idx=0)
    S_c = c_oracle.rollout_costs(c, np.array(cases.X0), eps, 0)
    np.testing.assert_allclose(S_c, S_py, rtol=1e-12)
    assert np.all(S_py >= base) and np.mean(S_py > base * (1 + 1e-9)) > 0.9      # these arms fall out of the 0.1 rad box
    wide = mo.OracleMPPI(**kw, joint_limit_lo=(-50, -50), joint_limit_hi=(50, 50), joint_limit_weight=3.0)
    np.testing.assert_array_equal(mo.rollout_costs(wide, np.array(cases.X0), eps, prev_idx=0), base)
    assert mo.joint_limit_cost(1.5, 0.0, (0, -1), (1, 1), 2.0) == 2.0 * 0.25 * 1e4
    assert mo.joint_limit_cost(-0.5, 2.0, (0, -1), (1, 1), 2.0) == 2.0 * (0.25 + 1.0) * 1e4
    # the step as a whole carries the term too (loops and vectorised forms agree)
    a = mo.step_loops(mo.OracleMPPI(**cases.run_py_kwargs(ref, 24, 8), **lim), cases.X0, eps[:24, :8])
    b = mo.step_vectorized(mo.OracleMPPI(**cases.run_py_kwargs(ref, 24, 8), **lim), cases.X0, eps[:24, :8])
    np.testing.assert_allclose(a["S"], b["S"], rtol=1e-12)

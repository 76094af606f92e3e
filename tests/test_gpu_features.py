"""GPU tests of everything around the injected-noise parity: Philox mode, sharding over ranks,
batched environments, error behaviour, closed loop."""
import contextlib
import io

import numpy as np
import pytest

from oracle import mppi_oracle as mo
from tests import helpers as H
from tests.golden import cases

pytestmark = pytest.mark.gpu

TOL_U = 1e-4
TOL_S = 2e-6


def _engine(paths, K, T, **kw):
    from mppi_robotarm_b200 import MppiEngine
    from mppi_robotarm_b200.arm_params import SYS_PARAMS
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    base = dict(K=K, T=T, delta_t=0.006, param_lambda=100.0, param_gamma=2.0, sigma=np.eye(2) * 20.0,
                stage_cost_weight=[0.5, 0.5, 5, 5], terminal_cost_weight=[5, 5, 50, 50], arm_params=SYS_PARAMS(),
                ref_path=ref, seed=99)
    base.update(kw)
    return MppiEngine(**base)


def _u0(T):
    return np.tile([10.0, -2.0], (T, 1))


# ---------------------------------------------------------------------------------------------
# Philox mode
# ---------------------------------------------------------------------------------------------
def test_philox_step_equals_oracle_on_the_exported_noise(paths):
    """The in-kernel draw can be exported; the oracle on that tensor must reproduce the step, and
    the injected-noise kernels fed the same tensor must give bit-identical costs."""
    K, T = 4096, 50
    eng = _engine(paths, K, T)
    eps = eng.philox_noise(step=0)                                  # [1, K, T, 2] on the device
    eng.step(cases.X0, _u0(T), 0, None)
    S_ph = eng.last_costs()[0][0].clone()
    u_ph = eng.out_u_new[0].copy()
    raw_ph = eng.out_w_eps_raw[0].copy()
    c = mo.OracleMPPI(**cases.run_py_kwargs(cases.ref_path_for(paths, "xydq_circle.txt"), K, T))
    o = mo.step_vectorized(c, cases.X0, eps[0].cpu().numpy().astype(np.float64))
    assert H.rel_err(S_ph.cpu().numpy().astype(np.float64), o["S"]) <= TOL_S
    assert H.rel_err(u_ph, o["u_new"]) <= TOL_U
    inj = _engine(paths, K, T)
    inj.step(cases.X0, _u0(T), 0, eps)
    assert bool((inj.last_costs()[0][0] == S_ph).all()), "rollout arithmetic must not depend on the noise source"
    np.testing.assert_allclose(inj.out_w_eps_raw[0], raw_ph, rtol=0, atol=1e-5 * np.max(np.abs(raw_ph)) + 1e-12)
    eng.close(); inj.close()


def test_batched_philox_step_with_several_blocks_per_environment(paths):
    """n_env > 1 with K large enough for the fused weight-sum kernel to run several blocks per environment
    (K > 512): every environment's normaliser and update must be its own.  Checked per environment against
    the unfused kernels fed the exported noise, and against the oracle (round-1 bug: the eta partials of
    neighbouring environments overlapped)."""
    K, T, E = 3000, 20, 3
    lam = 5.0e4                                    # soft weights: every block of an environment contributes to eta
    x0 = np.array([cases.X0, [0.9, 0.6, 0.3, -0.2], [1.0, -1.0, 0.1, 0.0]])
    u = np.stack([_u0(T), 0.5 * _u0(T), 2.0 * _u0(T)])
    p = [0, 700, 300]
    eng = _engine(paths, K, T, n_env=E, param_lambda=lam)
    eps = eng.philox_noise(step=0)                 # [E, K, T, 2]
    eng.step(x0, u, p, None)
    inj = _engine(paths, K, T, n_env=E, param_lambda=lam)
    inj.step(x0, u, p, eps)
    assert bool((inj.last_costs()[0] == eng.last_costs()[0]).all())
    np.testing.assert_allclose(eng.out_eta, inj.out_eta, rtol=1e-6)
    np.testing.assert_allclose(eng.out_rho, inj.out_rho, rtol=0, atol=0)
    np.testing.assert_allclose(eng.out_w_eps_raw, inj.out_w_eps_raw, rtol=0, atol=1e-5 * np.max(np.abs(inj.out_w_eps_raw)))
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    for e in range(E):
        c = mo.OracleMPPI(**cases.run_py_kwargs(ref, K, T, param_lambda=lam, param_alpha=1.0 - 2.0 / lam))
        c.u_prev = u[e].copy(); c.prev_waypoints_idx = p[e]
        o = mo.step_vectorized(c, x0[e], eps[e].cpu().numpy().astype(np.float64))
        assert H.rel_err(eng.out_u_new[e], o["u_new"]) <= TOL_U, e
        assert abs(eng.out_eta[e] - o["eta"]) <= 1e-4 * o["eta"], (e, eng.out_eta[e], o["eta"])
    assert float(np.max(eng.out_eta)) > 2.0       # the weights really are spread over many samples
    eng.close(); inj.close()


def test_philox_noise_statistics_and_streams(paths):
    from scipy import stats
    K, T = 8192, 50
    sig = np.array([[20.0, 6.0], [6.0, 10.0]])
    eng = _engine(paths, K, T, sigma=sig)
    e0 = eng.philox_noise(step=0)[0].cpu().numpy().astype(np.float64)
    e1 = eng.philox_noise(step=1)[0].cpu().numpy().astype(np.float64)
    flat = e0.reshape(-1, 2)
    assert np.all(np.abs(flat.mean(0)) < 4 * np.sqrt(np.diag(sig) / flat.shape[0]) + 1e-3)
    np.testing.assert_allclose(np.cov(flat.T), sig, atol=0.15)
    z = np.linalg.solve(np.linalg.cholesky(sig), flat.T).T           # whitened -> N(0, I)
    for col in range(2):
        assert stats.kstest(z[::7, col], "norm").pvalue > 1e-3
    assert abs(stats.kurtosis(z[:, 0])) < 0.05 and abs(stats.skew(z[:, 1])) < 0.02
    # different control steps / neighbouring samples / neighbouring horizon steps are uncorrelated
    for a, b in ((e0[..., 0].ravel(), e1[..., 0].ravel()), (e0[:-1, :, 0].ravel(), e0[1:, :, 0].ravel()),
                 (e0[:, :-1, 1].ravel(), e0[:, 1:, 1].ravel())):
        assert abs(np.corrcoef(a, b)[0, 1]) < 0.01
    assert not np.array_equal(e0, e1)
    # reproducible, and independent of how the samples are sharded (keyed on the global index)
    np.testing.assert_array_equal(eng.philox_noise(step=0)[0].cpu().numpy(), e0.astype(np.float32))
    from mppi_robotarm_b200 import ShardSpec
    half = _engine(paths, K, T, sigma=sig, shard=ShardSpec(1, 2))
    np.testing.assert_array_equal(half.philox_noise(step=0)[0].cpu().numpy(), e0[K // 2:].astype(np.float32))
    other = _engine(paths, K, T, sigma=sig, seed=100)
    assert abs(np.corrcoef(other.philox_noise(step=0)[0].cpu().numpy()[..., 0].ravel(), e0[..., 0].ravel())[0, 1]) < 0.01
    eng.close(); half.close(); other.close()


def test_philox_odd_horizon_and_graph_replay(paths):
    """T odd (last Philox pair half used) and the CUDA-graph replay give the same result as the
    plain launches."""
    for T in (7, 30):
        a = _engine(paths, 1000, T, use_graph=True)
        b = _engine(paths, 1000, T, use_graph=False)
        for s in range(3):
            a.step(cases.X0, _u0(T), 0, None)
            b.step(cases.X0, _u0(T), 0, None)
            np.testing.assert_array_equal(a.out_u_new, b.out_u_new)
            np.testing.assert_array_equal(a.out_opt_traj, b.out_opt_traj)
        eps = a.philox_noise(step=2)
        inj = _engine(paths, 1000, T)
        inj.step(cases.X0, _u0(T), 0, eps)
        np.testing.assert_allclose(inj.out_u_new, a.out_u_new, rtol=1e-6, atol=1e-6)
        a.close(); b.close(); inj.close()


# ---------------------------------------------------------------------------------------------
# sharded step: rank-partials + combine == single-GPU step
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world,K,T", [(2, 4096, 50), (4, 1000, 30), (8, 1003, 9)])
def test_sharded_partials_combine_to_the_single_gpu_result(paths, world, K, T):
    """Emulates `world` ranks one after the other on this GPU (no inter-kernel waiting): each shard
    runs mppi_step_local, the partials are concatenated as an all-gather would, and every 'rank'
    runs mppi_step_combine."""
    import torch
    from mppi_robotarm_b200 import ShardSpec, _cabi
    lam = 2000.0 if K == 1000 else 100.0                              # one case with many non-zero weights
    single = _engine(paths, K, T, param_lambda=lam, param_gamma=lam * 0.02)
    for mode in ("philox", "injected"):
        eps_full = single.philox_noise(step=0) if mode == "injected" else None
        single.step_counter = 0
        single.step(cases.X0, _u0(T), 0, eps_full)
        ref_u, ref_raw = single.out_u_new[0].copy(), single.out_w_eps_raw[0].copy()
        ref_rho, ref_eta = single.out_rho[0], single.out_eta[0]
        shards = [_engine(paths, K, T, param_lambda=lam, param_gamma=lam * 0.02, shard=ShardSpec(r, world))
                  for r in range(world)]
        parts = []
        for e in shards:
            e.write_inputs(cases.X0, _u0(T), 0)
            if mode == "injected":
                ptr = e._stage_eps(eps_full[0, e.k_offset:e.k_offset + e.K_local]).data_ptr()
                part = e.launch_local(_cabi.NOISE_INJECTED, ptr)
            else:
                part = e.launch_local(_cabi.NOISE_PHILOX, None)
            e.stream.synchronize()                                    # the partial is written on the engine's stream
            parts.append(part.clone())
        torch.cuda.synchronize()
        gathered = torch.stack(parts, 0).contiguous()                 # [world, n_env, 2 + 2T]
        torch.cuda.synchronize()
        for e in shards:
            e.launch_combine(gathered, world)
            e.wait()
            assert e.out_rho[0] == ref_rho
            np.testing.assert_allclose(e.out_eta[0], ref_eta, rtol=1e-6)
            scale = np.max(np.abs(ref_raw)) + 1e-12
            assert np.max(np.abs(e.out_w_eps_raw[0] - ref_raw)) <= 2e-6 * scale
            assert H.rel_err(e.out_u_new[0], ref_u) <= 2e-6
            np.testing.assert_array_equal(e.out_u_new, shards[0].out_u_new)   # bit-identical on all ranks
            e.close()
    single.close()


# ---------------------------------------------------------------------------------------------
# batched environments
# ---------------------------------------------------------------------------------------------
def test_batched_environments_equal_independent_controllers(paths):
    from mppi_robotarm_b200.batched import BatchedMPPIController
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    traj1 = paths["trajectory1"]
    B, K, T = 5, 512, 24
    rows = [0, 300, 900, 1500, 1975]
    X = np.array([[traj1[r, 0], traj1[r, 1], 0.05 * i, -0.03 * i] for i, r in enumerate(rows)])
    kw = cases.run_py_kwargs(ref, K, T)
    bat = BatchedMPPIController(B, **{k: v for k, v in kw.items()}, visualize_optimal_traj=True, seed=5,
                                return_sequences=True)
    bat.prev_waypoints_idx = np.array(rows)
    eps = np.stack([mo.injected_noise(50 + b, K, T, kw["sigma"]) for b in range(B)])
    u0, useq, opt = bat.calc_control_input(X, eps=eps)
    for b in range(B):
        c = mo.OracleMPPI(**kw)
        c.prev_waypoints_idx = rows[b]
        o = mo.step_vectorized(c, X[b], eps[b].astype(np.float64))
        assert bat.prev_waypoints_idx[b] == o["prev_idx_after"]
        assert H.rel_err(bat.engine.out_u_new[b], o["u_new"]) <= TOL_U
        assert np.max(np.abs(u0[b] - o["u0"])) <= TOL_U * np.max(np.abs(o["u_new"]))
        assert H.rel_err(useq[b], c.u_prev) <= TOL_U
        np.testing.assert_allclose(opt[b], o["optimal_traj"], rtol=0, atol=2e-5)
    # Philox: every environment draws its own stream
    bat.calc_control_input(X)
    e = bat.engine.philox_noise(step=1).cpu().numpy()
    assert abs(np.corrcoef(e[0, ..., 0].ravel(), e[1, ..., 0].ravel())[0, 1]) < 0.03
    bat.close()


def test_resident_controller_state_equals_host_driven_steps(paths):
    """BatchedMPPIController keeps sequences, waypoint indices and the step counter on the device
    (MPPI_FLAG_RESIDENT_STATE: only x0 in, (new_idx, u0) out per step).  Six steps, one environment reaching the end
    of its path on the way (it must freeze, control.py:76-78): controls, indices and the fetched sequences are
    the very floats of an engine whose state is shifted on the host (control.py:126, 148-149) every step —
    also for a block too large for zero-copy (320 environments: DMA path, compact read-back)."""
    from mppi_robotarm_b200.batched import BatchedMPPIController
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    traj1 = paths["trajectory1"]
    for B, K, T in ((4, 256, 16), (320, 64, 24)):
        rows = np.array(([0, 400, 1000, 1999] * (B // 4))[:B])       # (1999: the last waypoint — that environment is at the end)
        X = np.array([[traj1[r, 0], traj1[r, 1], 0.0, 0.0] for r in rows])
        kw = cases.run_py_kwargs(ref, K, T)
        bat = BatchedMPPIController(B, **kw, seed=21, visualize_optimal_traj=True)
        bat.prev_waypoints_idx = rows
        eng = _engine(paths, K, T, n_env=B, seed=21)
        assert (eng.layout.bytes > 65536) == (B == 320)
        u = np.tile([10.0, -2.0], (B, T, 1))
        p = rows.astype(np.int64).copy()
        finished = np.zeros(B, bool)
        for step in range(6):
            x = X + 0.001 * step
            u0, useq, opt = bat.calc_control_input(x)
            assert useq is None and opt is None
            eng.step(x, u, p, None)
            p = eng.out_new_idx.astype(np.int64)
            ended = p >= ref.shape[0] - 1
            live = ~ended
            finished |= ended
            u_ret = u[:, 0].copy()
            u[live] = eng.out_u_new[live]
            u[live, :-1] = u[live, 1:]
            u_ret[live] = u[live, 0]
            np.testing.assert_array_equal(bat.last_waypoint_idx(), p)
            np.testing.assert_array_equal(u0, u_ret)
            np.testing.assert_array_equal(bat.finished, finished)
            if step in (2, 5):                      # reading the attribute fetches the device's sequences
                np.testing.assert_array_equal(bat.u_prev, u)
                np.testing.assert_array_equal(bat.prev_waypoints_idx, p)
                np.testing.assert_array_equal(bat.engine.out_opt_traj, eng.out_opt_traj)
        assert finished.any() and not finished.all()
        # a host-side change of the state is pushed before the next step
        bat.u_prev[...] = 0.5 * u
        u0, _, _ = bat.calc_control_input(X)
        eng.step(X, 0.5 * u, p, None)
        live = eng.out_new_idx < ref.shape[0] - 1
        np.testing.assert_array_equal(u0[live], eng.out_u_new[live, 1])
        bat.close(); eng.close()


# ---------------------------------------------------------------------------------------------
# error behaviour and API details of the drop-in class
# ---------------------------------------------------------------------------------------------
def test_end_of_path_raises_index_error_and_keeps_sequence(paths, capsys):
    from control import MPPIControllerForPathTracking
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    c = MPPIControllerForPathTracking(**cases.run_py_kwargs(ref, 64, 10), verbose=False)
    c.prev_waypoints_idx = 1995
    before = c.u_prev.copy()
    with pytest.raises(IndexError):
        c.calc_control_input([1.15, -1.26, 0.0, 0.0])          # nearest of rows 1995..1999 is the last
    assert "[ERROR] Reached the end of the reference path." in capsys.readouterr().out
    assert c.prev_waypoints_idx == 1999                          # control.py:230 ran before the check
    np.testing.assert_array_equal(c.u_prev, before)
    c.close()


def test_sigma_errors_match_reference_types(paths):
    from control import MPPIControllerForPathTracking
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    bad = MPPIControllerForPathTracking(**cases.run_py_kwargs(ref, 8, 5, sigma=np.eye(3)), verbose=False)
    with pytest.raises(ValueError):
        bad.calc_control_input(cases.X0)
    kw = cases.run_py_kwargs(ref, 8, 5)
    kw.pop("sigma")
    singular = MPPIControllerForPathTracking(**kw, verbose=False)          # default Sigma of control.py:30
    with pytest.raises(np.linalg.LinAlgError):
        singular.calc_control_input(cases.X0)


def test_prints_three_lines_per_step_like_the_reference(paths):
    from control import MPPIControllerForPathTracking
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    c = MPPIControllerForPathTracking(**cases.run_py_kwargs(ref, 64, 10), seed=3)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        c.calc_control_input(list(cases.X0))
    lines = buf.getvalue().splitlines()
    assert lines[0] == "0     prev_idx = 0" and lines[1].startswith("0     nearest_idx = ")
    assert lines[2] == "======================updated======================="
    c.close()


def test_user_can_set_controller_state_between_steps(paths):
    """u_prev and prev_waypoints_idx are plain host attributes, read every step (checkpoint/resume)."""
    from control import MPPIControllerForPathTracking
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    a = MPPIControllerForPathTracking(**cases.run_py_kwargs(ref, 256, 20), seed=11, verbose=False)
    b = MPPIControllerForPathTracking(**cases.run_py_kwargs(ref, 256, 20), seed=11, verbose=False)
    x = np.array(cases.X0)
    for _ in range(3):
        a.calc_control_input(x)
    # resume b from a's state after 2 steps
    c2 = MPPIControllerForPathTracking(**cases.run_py_kwargs(ref, 256, 20), seed=11, verbose=False)
    for _ in range(2):
        c2.calc_control_input(x)
    b.u_prev = c2.u_prev.copy()
    b.prev_waypoints_idx = c2.prev_waypoints_idx
    b._engine().step_counter = 2
    b.calc_control_input(x)
    np.testing.assert_array_equal(a.u_prev, b.u_prev)
    for c in (a, b, c2):
        c.close()


# ---------------------------------------------------------------------------------------------
# closed loop, free running, exactly as run.py drives the controller (run.py:8-59)
# ---------------------------------------------------------------------------------------------
def _run_py_loop(mppi, ref, steps=1500, on_step=None):
    """The loop of run.py:48-59 with the reference's plant helpers (this repo's utils.py)."""
    from utils import Arm_Dynamic, Forward_Kinemetic
    dt = 0.003
    q = np.array(cases.X0[0:2]); dq = np.array([0.0, 0.0])
    state = [q[0], q[1], dq[0], dq[1]]
    states, err = [], []
    for k in range(1, steps + 1):
        states.append(np.array(state, dtype=np.float64))
        if on_step:
            on_step(k - 1)
        u, seq, opt, samp = mppi.calc_control_input(observed_x=state)
        dq += dt * Arm_Dynamic(q, dq, u)
        q += dt * dq
        _, _, x2, y2 = Forward_Kinemetic(q)
        state = np.concatenate((q, dq))
        p = mppi.prev_waypoints_idx
        err.append(np.hypot(x2 - ref[p, 0], y2 - ref[p, 1]))
    return np.array(states), np.array(err), samp


def _run_py_controller(ref, **extra):
    from control import MPPIControllerForPathTracking
    dt = 0.003
    return MPPIControllerForPathTracking(
        delta_t=dt * 2, ref_path=ref, horizon_step_T=30, number_of_samples_K=100, param_exploration=0.0,
        param_lambda=100.0, param_alpha=0.98, sigma=np.array([[20.0, 0.0], [0.0, 20.0]]),
        stage_cost_weight=np.array([0.50, 0.50, 5.0, 5.0]), terminal_cost_weight=np.array([5.0, 5.0, 50.0, 50.0]),
        visualze_sampled_trajs=True, verbose=False, **extra)


def test_reference_run_py_runs_unchanged_on_the_gpu(tmp_path, monkeypatch, capsys):
    """The reference's own run.py, byte for byte (staged by oracle/make_ref.py; skipped where no copy exists),
    executed with THIS repository's control.py / utils.py / sys_params.py on the path, the exported data file in
    the working directory and matplotlib stubbed: 50 ticks of its loop (run.py:48-59) on the GPU."""
    import os
    import runpy
    import sys
    from oracle import make_ref, ref_harness as rh
    src = make_ref.staged_dir()
    if src is None:
        pytest.skip("no staged copy of the reference (oracle/_ref/) on this machine")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import export_ref_paths
    export_ref_paths.export(str(tmp_path))
    monkeypatch.chdir(tmp_path)
    rh._install_stubs()
    for name in ("control", "utils", "sys_params"):
        monkeypatch.delitem(sys.modules, name, raising=False)
    monkeypatch.syspath_prepend(root)
    import control

    class _Enough(Exception):
        pass
    seen = {"ticks": 0, "ctrl": None, "u": []}
    real = control.MPPIControllerForPathTracking.calc_control_input

    def counted(self, observed_x):
        if seen["ticks"] == 50:
            raise _Enough
        out = real(self, observed_x)
        seen["ticks"] += 1
        seen["ctrl"] = self
        seen["u"].append(np.array(out[0], copy=True))
        assert out[3].shape == (100, 30, 4)                   # run.py asks for the sampled trajectories
        return out
    monkeypatch.setattr(control.MPPIControllerForPathTracking, "calc_control_input", counted)
    # run.py leaves the noise unseeded, and about one stream in twenty keeps the arm at waypoint 0 for 50 ticks (the
    # oracle loop does the same on those streams): fix the Philox key to one on which the FP64 oracle loop is at
    # waypoint 20 after 20 ticks and at 62 after 50
    real_init = control.MPPIControllerForPathTracking.__init__

    def seeded(self, *a, **k):
        k.setdefault("seed", 4)
        real_init(self, *a, **k)
    monkeypatch.setattr(control.MPPIControllerForPathTracking, "__init__", seeded)
    with pytest.raises(_Enough):
        runpy.run_path(os.path.join(src, "run.py"), run_name="__main__")
    text = capsys.readouterr().out
    c = seen["ctrl"]
    assert seen["ticks"] == 50 and type(c).__module__ == "mppi_robotarm_b200.controller"
    assert (c.K, c.T, c.visualze_sampled_trajs) == (100, 30, True)
    assert "50     state = " in text and "======================updated=======================" in text
    assert np.all(np.isfinite(seen["u"])) and 0 < c.prev_waypoints_idx < 200
    c.close()


def test_free_running_same_noise_follows_the_reference_run(paths):
    """Same seeded noise as the reference's recorded 1500-step run, free running (errors feed back).
    Stated bound (closed loop is chaotic w.r.t. rounding even in FP64): joints within
    1e-3 rad of the reference for the first 30 steps and within 0.5 rad over the whole run; the run
    completes without IndexError and ends at a comparable waypoint."""
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    with np.load(cases.HERE + "/closed_loop_c1.npz") as z:
        cl = {k: z[k] for k in z.files}
    K, T, seed0, _ = (int(v) for v in cl["meta"])
    mppi = _run_py_controller(ref, noise="numpy")
    sig = mppi.Sigma
    states, err, _ = _run_py_loop(mppi, ref, on_step=lambda s: H.inject(mppi, mo.injected_noise(seed0 + s, K, T, sig)))
    dev = np.max(np.abs(states[:, 0:2] - cl["state"][:, 0:2]), axis=1)
    first = int(np.argmax(dev > 1e-3)) if np.any(dev > 1e-3) else len(dev)
    print(f"same-noise closed loop: joints within 1e-3 rad for {first} steps; max deviation {dev.max():.3f} rad")
    assert dev[:30].max() <= 1e-3
    assert dev.max() <= 0.5
    assert abs(mppi.prev_waypoints_idx - int(cl["prev_idx"][-1, 1])) <= 120
    mppi.close()


def test_free_running_philox_closed_loop_tracks_like_the_reference(paths):
    """run.py's loop with in-kernel Philox noise: statistical agreement with the reference's runs
    (its own seeds spread over mean error 0.011-0.028 m, max 0.024-0.078 m, final index 1720-1809)."""
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    finals, means, maxs = [], [], []
    for seed in (2024, 7, 99):
        mppi = _run_py_controller(ref, seed=seed)
        _, err, samp = _run_py_loop(mppi, ref)
        assert samp.shape == (100, 30, 4) and np.all(np.isfinite(samp))
        finals.append(mppi.prev_waypoints_idx); means.append(err[50:].mean()); maxs.append(err[50:].max())
        mppi.close()
    print("philox closed loops: mean err", means, "max err", maxs, "final idx", finals)
    assert np.median(means) <= 0.04 and max(maxs) <= 0.15
    assert all(1600 <= f <= 1900 for f in finals), finals


def test_two_handles_in_flight_do_not_share_window_state(paths):
    """Single-environment handles stage their window in one constant-bank table per device; steps of
    different handles launched back to back on different streams must still see their own window."""
    T = 20
    a = _engine(paths, 20000, T, seed=1)
    b = _engine(paths, 20000, T, seed=2)
    xa, xb = np.array(cases.X0), np.array([paths["trajectory1"][900, 0], paths["trajectory1"][900, 1], 0.1, -0.1])
    # sequential reference results
    a.step(xa, _u0(T), 0, None); ua = a.out_u_new.copy(); ia = int(a.out_new_idx[0])
    b.step(xb, _u0(T), 890, None); ub = b.out_u_new.copy(); ib = int(b.out_new_idx[0])
    assert ia != ib
    for _ in range(5):
        a.step_counter = 0; b.step_counter = 0
        a.write_inputs(xa, _u0(T), 0); b.write_inputs(xb, _u0(T), 890)
        a.launch(None); b.launch(None)           # both in flight
        b.wait(); a.wait()
        np.testing.assert_array_equal(a.out_u_new, ua)
        np.testing.assert_array_equal(b.out_u_new, ub)
    a.close(); b.close()


def test_peer_exchange_timeout_skips_the_update(paths):
    """A peer whose partial never arrives (here: rank 1 of 2 does not exist — both exchange buffers are local
    allocations of this GPU and nobody raises rank 1's flag): after the configured timeout the step must NOT combine
    the stale slot.  It skips the update (u_new = u_prev, zero update, NaN rho / eta), reports it in the status
    word and through mppi_exchange_status, and leaves the handle usable: the next step reports again instead of
    staying poisoned, and after re-configuring the exchange for a world of one the step is the single-GPU step."""
    import ctypes as C
    import torch
    from mppi_robotarm_b200 import ShardSpec, _cabi
    K, T = 2048, 20
    eng = _engine(paths, K, T, shard=ShardSpec(0, 2), exchange="nccl")
    lib = eng.lib
    nbytes = int(lib.mppi_exchange_bytes(C.byref(eng.cfg), 2))
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=eng.device) for _ in range(2)]
    table = (C.c_void_p * 2)(*[b.data_ptr() for b in bufs])
    _cabi.check(lib.mppi_set_peer_exchange(eng.handle, 0, 2, table), eng.handle)
    _cabi.check(lib.mppi_set_exchange_timeout(eng.handle, 5.0), eng.handle)
    assert lib.mppi_set_exchange_timeout(eng.handle, 0.0) != 0
    u = _u0(T)
    for _ in range(2):
        eng.write_inputs(cases.X0, u, 0)
        _cabi.check(lib.mppi_step_sharded(eng.handle, _cabi.NOISE_PHILOX, None, eng.stream.cuda_stream), eng.handle)
        _cabi.check(lib.mppi_wait(eng.handle), eng.handle)
        eng.step_counter += 1
        assert eng.out_status[0] == 1 and lib.mppi_exchange_status(eng.handle) == 1
        np.testing.assert_array_equal(eng.out_u_new[0], u)
        assert not np.any(eng.out_w_eps_filt) and np.isnan(eng.out_rho[0]) and np.isnan(eng.out_eta[0])
    # a world of one (rank 0 alone, same buffers): its own put satisfies the wait
    eng.close()
    one = _engine(paths, K, T, seed=99)
    solo = _engine(paths, K, T, seed=99)
    table1 = (C.c_void_p * 1)(bufs[0].zero_().data_ptr())
    assert int(lib.mppi_exchange_bytes(C.byref(solo.cfg), 1)) <= nbytes
    _cabi.check(lib.mppi_set_peer_exchange(solo.handle, 0, 1, table1), solo.handle)
    solo.write_inputs(cases.X0, u, 0)
    _cabi.check(lib.mppi_step_sharded(solo.handle, _cabi.NOISE_PHILOX, None, solo.stream.cuda_stream), solo.handle)
    _cabi.check(lib.mppi_wait(solo.handle), solo.handle)
    one.step(cases.X0, u, 0, None)
    assert solo.out_status[0] == 0 and lib.mppi_exchange_status(solo.handle) == 0
    np.testing.assert_array_equal(solo.out_u_new, one.out_u_new)
    one.close(); solo.close()


def test_nccl_sharded_step_matches_single_gpu():
    """Real multi-process NCCL run (one rank per GPU); skipped on single-GPU boxes."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "run_dist_gpu.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and f"DIST_OK world={n}" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


def test_device_closed_loop_equals_host_driven_loop(paths):
    """mppi_closed_loop (plant + loop on the GPU) against the same controller driven tick by tick from
    the host with this repo's utils.Arm_Dynamic plant: same Philox stream, so the trajectories agree to
    FP64 plant rounding until chaos amplifies it."""
    from control import MPPIControllerForPathTracking
    from utils import Arm_Dynamic
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    kw = cases.run_py_kwargs(ref, 512, 30)
    dev = MPPIControllerForPathTracking(**kw, seed=31, verbose=False)
    host = MPPIControllerForPathTracking(**kw, seed=31, verbose=False)
    n, dt = 400, 0.003
    out = dev.run_closed_loop(cases.X0, n, dt)
    assert out["ticks"] == n and out["state"].shape == (n, 4)
    q, dq = np.array(cases.X0[0:2]), np.array(cases.X0[2:4])
    state = np.concatenate([q, dq])
    hs, hu, hi = [], [], []
    for _ in range(n):
        u, _, _, _ = host.calc_control_input(state)
        dq = dq + dt * Arm_Dynamic(q, dq, u)
        q = q + dt * dq
        state = np.concatenate([q, dq])
        hs.append(state.copy()); hu.append(u.copy()); hi.append(host.prev_waypoints_idx)
    hs, hu = np.array(hs), np.array(hu)
    dev_err = np.max(np.abs(out["state"] - hs), axis=1)
    np.testing.assert_allclose(out["state"][:25], hs[:25], rtol=0, atol=1e-9)
    np.testing.assert_allclose(out["u"][:25], hu[:25], rtol=0, atol=1e-7)
    assert list(out["waypoint_idx"][:25]) == hi[:25]
    assert dev_err.max() <= 0.5, dev_err.max()
    assert abs(dev.prev_waypoints_idx - host.prev_waypoints_idx) <= 60
    assert dev._engine().step_counter == n
    # the loop can be resumed from the controller state it leaves behind
    out2 = dev.run_closed_loop(out["final_state"], 50, dt)
    assert out2["ticks"] == 50 and np.all(np.isfinite(out2["state"]))
    dev.close(); host.close()


def test_device_closed_loop_ticks_against_the_oracle_loop(paths):
    """The ticks of the device-resident loop themselves against the FP64 oracle: the noise the kernels will
    draw at ticks 0..24 is exported first, then the oracle controller (restated control.py:67-152) and the oracle
    plant (utils.py:14-29, run.py:53-55) run the same 25 ticks on the CPU with that noise, each loop on its OWN
    state (free running).  Joints within 2e-4 rad over the 25 ticks and equal waypoint indices; the applied controls
    agree to the north_star bound (1e-4 relative) on the typical tick — a single tick may differ more, because at
    lambda = 100 the weights are winner-take-all and two near-tied samples can swap once the two loops' states
    differ by 1e-5 rad; the teacher-forced tests hold every step to 1e-4."""
    from control import MPPIControllerForPathTracking
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    K, T, n, dt = 512, 30, 25, 0.003
    kw = cases.run_py_kwargs(ref, K, T)
    dev = MPPIControllerForPathTracking(**kw, seed=77, verbose=False)
    eng = dev._engine()
    eps = [eng.philox_noise(step=t)[0].cpu().numpy().astype(np.float64) for t in range(n)]
    # start from tick 200 of the reference's recorded run (a tracking state, not the rest at the path's first rows)
    with np.load(cases.HERE + "/closed_loop_c1.npz") as z:
        x0, p0, prev = z["state"][200].copy(), int(z["prev_idx"][200, 0]), z["u_new"][199].copy()
    u_start = np.concatenate([prev[1:], prev[-1:]], axis=0)
    dev.u_prev[...] = u_start
    dev.prev_waypoints_idx = p0
    out = dev.run_closed_loop(x0, n, dt)
    c = mo.OracleMPPI(**kw)
    c.u_prev = u_start.copy()
    c.prev_waypoints_idx = p0
    q, dq = x0[0:2].copy(), x0[2:4].copy()
    err_u, worst_q = [], 0.0
    for t in range(n):
        o = mo.step_vectorized(c, np.concatenate([q, dq]), eps[t])
        u = o["u0"]                                        # the control run.py applies (post-shift, quirk Q2)
        q, dq = mo.plant_step(q, dq, u, dt)
        err_u.append(float(np.max(np.abs(out["u"][t] - u)) / np.max(np.abs(o["u_new"]))))
        worst_q = max(worst_q, float(np.max(np.abs(out["state"][t, 0:2] - q))))
        assert int(out["waypoint_idx"][t]) == o["prev_idx_after"], t
    print(f"device loop vs oracle loop, {n} ticks: control error median {np.median(err_u):.2e} / worst {max(err_u):.2e} "
          f"(relative), worst joint error {worst_q:.2e} rad")
    assert err_u[0] <= 1e-4 and np.median(err_u) <= 1e-4 and worst_q <= 2e-4, (err_u, worst_q)
    dev.close()


def test_device_closed_loop_stops_at_the_end_of_the_path(paths, capsys):
    from control import MPPIControllerForPathTracking
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    c = MPPIControllerForPathTracking(**cases.run_py_kwargs(ref, 256, 20), seed=5, verbose=False)
    c.prev_waypoints_idx = 1960
    x0 = [paths["trajectory1"][1960, 0], paths["trajectory1"][1960, 1], 0.0, 0.0]
    with pytest.raises(IndexError):
        c.run_closed_loop(x0, 400, 0.003)
    assert "[ERROR] Reached the end of the reference path." in capsys.readouterr().out
    assert 0 < c.last_loop["ticks"] < 400
    assert c.prev_waypoints_idx >= ref.shape[0] - 1
    assert np.all(np.diff(c.last_loop["waypoint_idx"]) >= 0)
    c.close()


def test_top_n_sampled_trajectories_are_the_best_rows_of_the_full_set(paths):
    """sampled_traj_top_n: the n lowest-cost samples in np.argsort(S) order, identical to the
    corresponding rows of the full (K, T, 4) output of control.py:137-145."""
    case = dict(name="topn", file="xydq_circle.txt", K=300, T=16)
    full, kw = H.make_controller(case, paths, visualze_sampled_trajs=True)
    top, _ = H.make_controller(case, paths, visualze_sampled_trajs=True, sampled_traj_top_n=7)
    eps = mo.injected_noise(3, 300, 16, kw["sigma"])
    for c in (full, top):
        H.inject(c, eps)
    _, _, _, all_traj = H.quiet_step(full, cases.X0)
    _, _, _, best = H.quiet_step(top, cases.X0)
    S = full._engine().last_costs()[0][0].cpu().numpy()
    order = np.argsort(S, kind="stable")[:7]
    assert best.shape == (7, 16, 4)
    np.testing.assert_array_equal(top.last["sampled_idx"], order)
    np.testing.assert_array_equal(best, all_traj[order])
    o = mo.step_vectorized(mo.OracleMPPI(**{**kw, "visualze_sampled_trajs": True}), cases.X0, eps.astype(np.float64))
    np.testing.assert_allclose(best, o["sampled_traj"][np.argsort(o["S"])[:7]], rtol=0, atol=2e-5)
    full.close(); top.close()


def test_reassigning_ref_path_is_followed(paths):
    """The reference reads self.ref_path on every call; a re-assigned path (same length) must be used."""
    from control import MPPIControllerForPathTracking
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    kw = cases.run_py_kwargs(ref, 256, 16)
    eps = mo.injected_noise(8, 256, 16, kw["sigma"])
    c = MPPIControllerForPathTracking(**kw, noise="numpy", verbose=False)
    H.inject(c, eps)
    H.quiet_step(c, cases.X0)
    shifted = ref.copy()
    shifted[:, 0] += 0.05
    c.ref_path = shifted
    c.u_prev[...] = [10.0, -2.0]
    c.prev_waypoints_idx = 0
    H.quiet_step(c, cases.X0)
    o = mo.step_vectorized(mo.OracleMPPI(**{**kw, "ref_path": shifted}), cases.X0, eps.astype(np.float64))
    assert H.rel_err(c._engine().out_u_new[0], o["u_new"]) <= TOL_U
    c.close()


def test_batched_device_closed_loop_equals_single_environment_loops(paths):
    """mppi_closed_loop with n_env > 1: every environment evolves exactly like a batched controller
    stepped from the host with the same plant (same Philox streams: same seed, env index, step)."""
    from mppi_robotarm_b200.batched import BatchedMPPIController
    from utils import Arm_Dynamic
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    traj1 = paths["trajectory1"]
    B, K, T, n, dt = 4, 256, 20, 60, 0.003
    rows = [0, 400, 1000, 1600]
    X0 = np.array([[traj1[r, 0], traj1[r, 1], 0.0, 0.0] for r in rows])
    kw = cases.run_py_kwargs(ref, K, T)
    dev = BatchedMPPIController(B, **kw, seed=9)
    host = BatchedMPPIController(B, **kw, seed=9)
    for c in (dev, host):
        c.prev_waypoints_idx = np.array(rows)
    log, stop = dev.run_closed_loop(X0, n, dt)
    assert log.shape == (n, B, 8) and np.all(stop >= n)
    X = X0.copy()
    for t in range(n):
        u0, _, _ = host.calc_control_input(X)
        for b in range(B):
            q, dq = X[b, 0:2], X[b, 2:4]
            dq = dq + dt * Arm_Dynamic(q, dq, u0[b])
            q = q + dt * dq
            X[b] = np.concatenate([q, dq])
        if t < 20:
            np.testing.assert_allclose(log[t, :, 0:4], X, rtol=0, atol=1e-9)
            np.testing.assert_allclose(log[t, :, 4:6], u0, rtol=0, atol=1e-7)
            np.testing.assert_array_equal(log[t, :, 6].astype(np.int64), host.last_waypoint_idx())
    assert np.max(np.abs(log[-1, :, 0:4] - X)) <= 0.3
    np.testing.assert_array_equal(dev.prev_waypoints_idx >= np.array(rows), True)
    dev.close(); host.close()


# ---------------------------------------------------------------------------------------------
# certified nearest-waypoint lookups (search="certified", the default)
# ---------------------------------------------------------------------------------------------
def _tracking_state(cl, s, T):
    prev = cl["u_new"][s - 1]
    u = np.concatenate([prev[1:], np.repeat(prev[-1:], max(T - 29, 1), axis=0)], axis=0)[:T]
    return cl["state"][s], u, int(cl["prev_idx"][s, 0])


@pytest.mark.parametrize("K,T,s,noise", [
    (32768, 100, 500, "philox"),       # 2 samples per thread; full: constant-bank window
    (32768, 100, 1000, "injected"),    # same kernel family, noise read from HBM
    (1000, 30, 100, "injected"),       # 1 sample per thread, partial last warp (1000 = 31 warps + 8 lanes); full: register window
    (4096, 50, 1499, "philox"),
    (64, 20, -1, "injected"),          # window truncated by the end of the path
    (16384, 50, 0, "philox"),          # arm at rest at the start of the path: every sample falls behind row 0
])
def test_certified_search_is_bit_identical_to_the_full_search(paths, K, T, s, noise):
    with np.load(cases.HERE + "/closed_loop_c1.npz") as z:
        cl = {k: z[k] for k in z.files}
    if s > 0:
        x0, u, p = _tracking_state(cl, s, T)
    elif s == 0:
        x0, u, p = cases.X0, _u0(T), 0
    else:
        x0, u, p = paths["trajectory1"][1987, 0:2].tolist() + [0.01, 0.01], _u0(T), 1985
    out = {}
    for mode in ("certified", "full"):
        eng = _engine(paths, K, T, search=mode, search_stats=True)
        eps = eng.philox_noise(step=0) if noise == "injected" else None
        eng.step(x0, u, p, eps)
        S, w = (t[0].cpu().numpy().copy() for t in eng.last_costs())
        out[mode] = (S, w, eng.out_u_new[0].copy(), eng.out_opt_traj[0].copy(), int(eng.out_new_idx[0]),
                     eng.search_stats())
        eng.close()
    a, b = out["certified"], out["full"]
    for i in range(4):
        assert np.array_equal(a[i], b[i]), f"output {i} differs between search modes"
    assert a[4] == b[4]
    assert b[5]["certified"] == 0 and b[5]["triples"] == 0 and b[5]["lookups"] == a[5]["lookups"] > 0
    if s > 0:
        assert a[5]["searched_fraction"] < 0.02, a[5]      # end tests + certified triples answer a tracking state
    if T >= 50 and s > 0:
        assert a[5]["fraction"] > 0.6, a[5]      # most of a long horizon lies beyond the 30-row window


@pytest.mark.parametrize("K,T,n_env,noise", [
    (200003, 12, 1, "philox"),      # 1563 units of 128 samples: one full wave of two-sample CTAs + 379 one-sample CTAs, partial last unit
    (113000, 20, 1, "injected"),    # 883 units: a single mixed wave (291 two-sample + 301 one-sample CTAs)
    (32768, 10, 6, "philox"),       # batched, 1536 units over 6 environments: the flat layout crosses environments
])
def test_rollout_launch_layouts_give_the_same_costs(paths, monkeypatch, K, T, n_env, noise):
    """The launch layout of the rollout kernel (mppi_cabi.cu::pick_layout: one sample per thread / two samples per
    thread / the mixed last wave) only decides which thread computes which sample: costs, weights and the update are
    the same floats in all three."""
    with np.load(cases.HERE + "/closed_loop_c1.npz") as z:
        cl = {k: z[k] for k in z.files}
    x0, u, p = _tracking_state(cl, 500, T)
    out = {}
    for name, env in (("default", {}), ("one", {"MPPI_NS": "1"}), ("two", {"MPPI_NO_MIXED": "1"})):
        for k in ("MPPI_NS", "MPPI_NO_MIXED"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng = _engine(paths, K, T, n_env=n_env)
        eps = eng.philox_noise(step=0) if noise == "injected" else None
        eng.step(np.tile(x0, (n_env, 1)) if n_env > 1 else x0, np.tile(u, (n_env, 1, 1)) if n_env > 1 else u,
                 np.full(n_env, p) if n_env > 1 else p, eps)
        S, w = (t.cpu().numpy().copy() for t in eng.last_costs())
        out[name] = (S, w, eng.out_u_new.copy(), eng.out_rho.copy(), eng.out_eta.copy())
        eng.close()
    assert np.all(np.isfinite(out["default"][0])) and out["default"][0].shape[-1] == K
    for name in ("one", "two"):
        for i in range(5):
            assert np.array_equal(out["default"][i], out[name][i]), (name, i)


def test_certified_search_batched_environments(paths):
    """n_env > 1: every environment has its own window and certificate (step block in shared memory)."""
    from mppi_robotarm_b200.batched import BatchedMPPIController
    with np.load(cases.HERE + "/closed_loop_c1.npz") as z:
        cl = {k: z[k] for k in z.files}
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    T, K, steps = 64, 512, (0, 100, 500, 1000, 1499)
    kw = cases.run_py_kwargs(ref, K, T)
    kw.pop("visualize_optimal_traj", None); kw.pop("visualze_sampled_trajs", None)
    res = {}
    for mode in ("certified", "full"):
        b = BatchedMPPIController(len(steps), **kw, seed=4, search=mode, search_stats=True)
        xs, us, ps = [], [], []
        for s in steps:
            x0, u, p = _tracking_state(cl, s, T) if s else (cases.X0, _u0(T), 0)
            xs.append(x0); us.append(u); ps.append(p)
        b.u_prev[...] = np.array(us)
        b.prev_waypoints_idx[...] = ps
        b.calc_control_input(np.array(xs))
        b.engine.download_state()          # (the sequences stay on the device unless asked for)
        res[mode] = (b.engine.last_costs()[0].cpu().numpy().copy(), b.engine.out_u_new.copy(), b.engine.search_stats())
        b.close()
    assert np.array_equal(res["certified"][0], res["full"][0])
    assert np.array_equal(res["certified"][1], res["full"][1])
    assert res["certified"][2]["fraction"] > 0.4 and res["full"][2]["certified"] == 0
    assert res["certified"][2]["searched_fraction"] < 0.3, res["certified"][2]     # (one of the five starts at rest at the path's first rows)


def test_window_table_equals_on_the_spot_construction(paths):
    """mppi_set_ref_path() builds the window part of the step block (search tables, certificate, wedges) for every
    window start of the path; the prepare kernel copies entry p.  Paths too long for a table (max_ref_rows > 131072)
    build the window in every step instead: same bytes, same results."""
    with np.load(cases.HERE + "/closed_loop_c1.npz") as z:
        cl = {k: z[k] for k in z.files}
    for s in (0, 100, 1000, -1):
        if s > 0:
            x0, u, p = _tracking_state(cl, s, 40)
        elif s == 0:
            x0, u, p = cases.X0, _u0(40), 0
        else:
            x0, u, p = paths["trajectory1"][1987, 0:2].tolist() + [0.01, 0.01], _u0(40), 1985
        res = []
        for max_rows in (None, 131073):
            eng = _engine(paths, 2048, 40, max_ref_rows=max_rows)
            eng.step(x0, u, p, None)
            res.append((eng.step_block(0).copy(), eng.last_costs()[0].cpu().numpy().copy(), eng.out_u_new.copy(),
                        int(eng.out_new_idx[0])))
            eng.close()
        for a, b in zip(*res):
            np.testing.assert_array_equal(a, b)


def test_device_built_certificates_are_sound(paths, emul):
    """The certificates the prepare kernel builds (one lane per row, FP64 with fast reciprocals) are read back
    from the step block and probed on the host against the exact FP32 search: no certified query may disagree,
    also for queries hugging every threshold.  On the reference's own files they must also be (nearly) the
    certificates of the serial construction the CPU tests cover."""
    import ctypes as C
    rng = np.random.default_rng(11)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))        # noqa: E731
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))          # noqa: E731
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))       # noqa: E731
    off = 64 + 32 * 32 + 16 * 16
    n_armed = n_certified = 0
    from tests.test_emul_cpu import _synthetic_window_path, cert_boundary_queries, CERT_FLOATS, ADVERSARIAL_KINDS
    todo = [(name, np.ascontiguousarray(paths[name][:, 0:4], dtype=np.float64), True)
            for name in ("xydq_circle", "trajectory1", "xydq")]
    # shapes the reference files do not contain (duplicates, zigzag, U-turn, spiral, micrometre spacing, a window
    # far from the base): soundness only — at their decision thresholds the two constructions may arm differently
    todo += [(kind, _synthetic_window_path(kind, rng, n=64), False) for kind in ADVERSARIAL_KINDS]
    for name, ref, check_serial in todo:
        n = ref.shape[0]
        eng = _engine(paths, 64, 8, ref_path=ref)
        for p in list(rng.integers(0, n - 31, 10 if check_serial else 4)) + [n - 31, n - 12, n - 3, n - 2]:
            p = int(p)
            # an arm state whose end effector sits on waypoint p (so the window starts exactly there)
            x, y = ref[p, 0], ref[p, 1]
            q2 = -np.arccos(np.clip((x * x + y * y - 2.0) / 2.0, -1, 1))
            q1 = np.arctan2(y, x) - np.arctan2(np.sin(q2), 1.0 + np.cos(q2))
            eng.step([q1, q2, 0.0, 0.0], _u0(8), p, None)
            blk = eng.step_block(0)
            start = int(blk[24:28].view(np.int32)[0])
            cert = blk[off:off + 4 * CERT_FLOATS].view(np.float32).copy()       # 64 B certificate + 32 row records + wedges
            nv = min(30, n - start)
            assert cert[13:14].view(np.int32)[0] == nv - 1
            N = 20000
            base = ref[start + rng.integers(0, nv, N), 0:2] - ref[start, 0:2]
            q = base + rng.standard_normal((N, 2)) * (10.0 ** rng.uniform(-5, 0.3, N))[:, None]
            hug = cert_boundary_queries(cert, rng, 200)
            n_armed += int(hug.shape[0] > 0)
            xy = np.ascontiguousarray(np.concatenate([q, hug]).astype(np.float32))
            pick, full = np.zeros(len(xy), np.int32), np.zeros(len(xy), np.int32)
            emul.emul_cert_probe_given(dp(ref), n, start, fp(cert), fp(xy), len(xy), ip(pick), ip(full))
            m = pick >= 0
            assert np.array_equal(pick[m], full[m]), (name, p)
            n_certified += int(m.sum())
            if not check_serial:
                continue
            serial = np.zeros(CERT_FLOATS, np.float32)
            scan = np.zeros(1, np.int32)
            emul.emul_cert_probe(dp(ref), n, start, C.c_double(2.0), fp(xy), 1, ip(pick), ip(full), ip(scan), fp(serial))
            # same roles armed, same thresholds (the device uses fast reciprocals: 1e-9 relative on the range)
            dev_rec, ser_rec = cert[16:272].reshape(32, 8), serial[16:272].reshape(32, 8)
            assert np.array_equal(np.abs(dev_rec[:, 6:8]) > 1e30, np.abs(ser_rec[:, 6:8]) > 1e30), (name, p)
            np.testing.assert_allclose(dev_rec[:, 0:3], ser_rec[:, 0:3], rtol=0, atol=0)
            np.testing.assert_allclose(dev_rec[:, 4:8], ser_rec[:, 4:8], rtol=2e-5, atol=2e-6)
            np.testing.assert_allclose(cert[0:4], serial[0:4], rtol=1e-4, atol=1e-6)       # normal and lateral range
            assert cert[11] == serial[11] and cert[12] == serial[12]                        # jhi, dom
            dw, sw = cert[272:284], serial[272:284]                                          # far-field wedges
            assert np.array_equal(np.abs(dw) > 1e30, np.abs(sw) > 1e30), (name, p)
            fin = np.abs(sw) < 1e30
            np.testing.assert_allclose(dw[fin], sw[fin], rtol=2e-4, atol=2e-6)
        eng.close()
    assert n_armed > 40 and n_certified > 500000


"""TEST INFRASTRUCTURE ONLY — FP64 CPU restatement of the reference MPPI step (the parity oracle).

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``mppi_robotarm_b200``) never does and fails loudly when its CUDA library is missing.

Pinning: the reference ships no golden vectors, so this oracle is pinned against outputs of the
*reference itself* executed in the build container (``oracle/ref_harness.py``); the vectors live in
``tests/golden/*.npz`` with the script that made them (``tests/golden/make_golden.py``).

Two restatements of the same algorithm:

* :func:`step_loops`      — scalar, follows the reference's loop nest one to one (small cases only).
* :func:`step_vectorized` — NumPy-vectorised over the K samples (scales to K = 16384 and beyond).

Reference lines restated (all in ``/root/reference``):
  control.py:21-65   constructor defaults, gamma = lambda*(1-alpha), cost-side l1=l2=1, u_prev init
  control.py:67-152  the step: waypoint update, noise, rollouts, weights, weighted sum, filter,
                     in-place update, visualisation rollouts, shift, return (quirks Q1-Q5)
  control.py:174-198 stage / terminal cost
  control.py:200-232 nearest waypoint in a 30-point forward window, first arg-min
  control.py:234-263 arm dynamics + semi-implicit Euler (mass matrix uses link lengths, Q7)
  control.py:265-295 the alternative rollout model _F1 (``dynamics="F1"``)
  control.py:297-314 soft-min weights
  control.py:319-327 scipy.ndimage.median_filter(size=10, mode='reflect') per column
  control.py:329-344 the alternative smoother _moving_average_filter (``smoother="average"``)
  sys_params.py:1-13 arm constants;  utils.py:14-38 plant dynamics / forward kinematics twins
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

SEARCH_IDX_LEN = 30      # control.py:203
FILTER_WINDOW = 10       # control.py:122
COST_SCALE = 10000.0     # control.py:185,198


def default_arm_params() -> dict:
    """sys_params.py:3-10 (``Ts`` is never read by the controller)."""
    return dict(Ts=0.0025, m1=1, m2=1, l1=1, l2=1, lc1=0.5, lc2=0.5, g=9.81)


@dataclass
class OracleMPPI:
    """Mutable controller state + hyper-parameters (mirror of control.py:21-65)."""
    delta_t: float
    ref_path: np.ndarray
    horizon_step_T: int
    number_of_samples_K: int
    param_exploration: float
    param_lambda: float
    param_alpha: float
    sigma: np.ndarray
    stage_cost_weight: np.ndarray
    terminal_cost_weight: np.ndarray
    visualize_optimal_traj: bool = True
    visualze_sampled_trajs: bool = False
    smoother: str = "median"          # "median" = control.py:122; "average" = control.py:329-344; "none"
    dynamics: str = "F"               # "F" = control.py:234-263; "F1" = control.py:265-295 (feedback-linearised)
    arm: dict = field(default_factory=default_arm_params)
    cost_l1: float = 1.0      # control.py:55
    cost_l2: float = 1.0      # control.py:56
    # joint-limit stage cost (BASELINE.json north_star item 1; NOT in the reference, whose only limits are the
    # commented-out clamps of _g, control.py:166-172): weight 0 = the reference's step, bit for bit
    joint_limit_lo: tuple = (-np.inf, -np.inf)
    joint_limit_hi: tuple = (np.inf, np.inf)
    joint_limit_weight: float = 0.0

    def __post_init__(self):
        self.T = int(self.horizon_step_T)
        self.K = int(self.number_of_samples_K)
        self.param_gamma = self.param_lambda * (1.0 - self.param_alpha)        # control.py:45
        self.u_prev = np.tile(np.array([10.0, -2.0]), (self.T, 1))             # control.py:59
        self.prev_waypoints_idx = 0                                            # control.py:65
        self.ref_path = np.asarray(self.ref_path, dtype=np.float64)
        self.sigma = np.asarray(self.sigma, dtype=np.float64)
        self.stage_cost_weight = np.asarray(self.stage_cost_weight, dtype=np.float64)
        self.terminal_cost_weight = np.asarray(self.terminal_cost_weight, dtype=np.float64)


# ----------------------------------------------------------------------------------------------
# building blocks (each works on scalars or on arrays broadcast over samples)
# ----------------------------------------------------------------------------------------------
def end_effector(q1, q2, l1=1.0, l2=1.0):
    """control.py:178-179 / 206-207 / utils.py:35-36."""
    return l1 * np.cos(q1) + l2 * np.cos(q1 + q2), l1 * np.sin(q1) + l2 * np.sin(q1 + q2)


def window_of(ref_path: np.ndarray, prev_idx: int) -> np.ndarray:
    """The slice control.py:208-209 searches; Python slicing truncates it at the end of the path."""
    return ref_path[prev_idx:prev_idx + SEARCH_IDX_LEN]


def nearest_in_window(win: np.ndarray, x, y):
    """First arg-min of ((x-rx)^2 + (y-ry)^2)*100 over the window (control.py:208-215).

    ``x``/``y`` may be scalars or [K] arrays; returns offsets into the window."""
    dx = np.asarray(x)[..., None] - win[:, 0]
    dy = np.asarray(y)[..., None] - win[:, 1]
    d = (dx ** 2 + dy ** 2) * 100
    return np.argmin(d, axis=-1)          # np.argmin returns the first minimum, like list.index(min)


def arm_accel(q1, q2, d1, d2, v1, v2, arm):
    """ddq = M^-1 (v - C dq - G) with the reference's matrices (control.py:241-252, utils.py:15-27).

    np.linalg.inv of a 2x2 is restated in closed form (adjugate / determinant); the difference to
    LAPACK is FP64 rounding (<=1e-15 relative, checked in tests)."""
    m1, m2, l1, l2, lc1, lc2, g = (arm[k] for k in ("m1", "m2", "l1", "l2", "lc1", "lc2", "g"))
    c2 = np.cos(q2)
    M11 = m1 * lc1 ** 2 + l1 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * c2) + l2
    M22 = m2 * lc2 ** 2 + l2
    M12 = m2 * l1 * lc2 * c2 + m2 * lc2 ** 2 + l2
    h = m2 * l1 * lc2 * np.sin(q2)
    g1 = m1 * lc1 * g * np.cos(q1) + m2 * g * (lc2 * np.cos(q1 + q2) + l1 * np.cos(q1))
    g2 = m2 * lc2 * g * np.cos(q1 + q2)
    cd1 = (-h * d2) * d1 + (-h * d1 - h * d2) * d2
    cd2 = (h * d1) * d1
    b1 = v1 - cd1 - g1
    b2 = v2 - cd2 - g2
    det = M11 * M22 - M12 * M12
    return (M22 * b1 - M12 * b2) / det, (M11 * b2 - M12 * b1) / det


def arm_accel_f1(q1, q2, d1, d2, v1, v2, arm):
    """control.py:265-290 (``_F1``): the input is turned into a torque u = M v + C dq + G with G = 0
    (control.py:281-284) and pushed through the same ddq = M^-1 (u - C dq - G) — i.e. ddq = v up to FP64
    rounding.  Restated operation by operation so the rounding is the reference's."""
    m1, m2, l1, l2, lc1, lc2 = (arm[k] for k in ("m1", "m2", "l1", "l2", "lc1", "lc2"))
    c2 = np.cos(q2)
    M11 = m1 * lc1 ** 2 + l1 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * c2) + l2
    M22 = m2 * lc2 ** 2 + l2
    M12 = m2 * l1 * lc2 * c2 + m2 * lc2 ** 2 + l2
    h = m2 * l1 * lc2 * np.sin(q2)
    cd1 = (-h * d2) * d1 + (-h * d1 - h * d2) * d2            # C.dot(dq)
    cd2 = (h * d1) * d1 + 0.0 * d2
    u1 = (M11 * v1 + M12 * v2) + cd1 + 0.0
    u2 = (M12 * v1 + M22 * v2) + cd2 + 0.0
    b1 = u1 - cd1 - 0.0
    b2 = u2 - cd2 - 0.0
    det = M11 * M22 - M12 * M12
    return (M22 * b1 - M12 * b2) / det, (M11 * b2 - M12 * b1) / det


def arm_step(q1, q2, d1, d2, v1, v2, arm, dt, dynamics="F"):
    """control.py:253-259: dq += ddq*dt, then q += (new dq)*dt."""
    a1, a2 = (arm_accel if dynamics == "F" else arm_accel_f1)(q1, q2, d1, d2, v1, v2, arm)
    d1 = d1 + a1 * dt
    d2 = d2 + a2 * dt
    return q1 + d1 * dt, q2 + d2 * dt, d1, d2


def tracking_cost(q1, q2, d1, d2, win, weights, l1, l2):
    """control.py:174-198 (same formula for stage and terminal cost, different weights)."""
    x, y = end_effector(q1, q2, l1, l2)
    j = nearest_in_window(win, x, y)
    r = win[j]
    c = (weights[0] * (x - r[..., 0]) ** 2 + weights[1] * (y - r[..., 1]) ** 2
         + weights[2] * (d1 - r[..., 2]) ** 2 + weights[3] * (d2 - r[..., 3]) ** 2)
    return c * COST_SCALE


def joint_limit_cost(q1, q2, lo, hi, weight):
    """Extension (north_star item 1), zero by default: weight * (viol(q1)^2 + viol(q2)^2) * 1e4 with
    viol(q) = max(q - hi, lo - q, 0), added to the stage cost of every horizon step."""
    v1 = np.maximum(np.maximum(q1 - hi[0], lo[0] - q1), 0.0)
    v2 = np.maximum(np.maximum(q2 - hi[1], lo[1] - q2), 0.0)
    return weight * (v1 * v1 + v2 * v2) * COST_SCALE


def _stage_extra(c, s):
    if c.joint_limit_weight == 0.0:
        return 0.0
    return joint_limit_cost(s[0], s[1], c.joint_limit_lo, c.joint_limit_hi, c.joint_limit_weight)


def softmin_weights(S, lam):
    """control.py:297-314."""
    rho = S.min()
    e = np.exp((-1.0 / lam) * (S - rho))
    return e / e.sum(), rho, e.sum()


def median_filter_reflect(x: np.ndarray, size: int = FILTER_WINDOW) -> np.ndarray:
    """1-D ``scipy.ndimage.median_filter(x, size, mode='reflect')`` restated (control.py:325).

    Semantics (checked against SciPy in tests/test_oracle.py): the window covers
    offsets -(size//2) .. size-1-size//2, the output is the element of rank size//2 in the sorted
    window (upper median for an even size), and out-of-range indices are mirrored about the array
    edges with the edge sample repeated (d c b a | a b c d | d c b a), repeatedly if needed."""
    n = x.shape[0]
    back = size // 2
    idx = np.arange(n)[:, None] + (np.arange(size) - back)[None, :]
    period = 2 * n
    idx = np.mod(idx, period)
    idx = np.where(idx >= n, period - 1 - idx, idx)
    return np.sort(x[idx], axis=1)[:, size // 2]


def filter_columns(xx: np.ndarray, size: int = FILTER_WINDOW) -> np.ndarray:
    return np.stack([median_filter_reflect(xx[:, d], size) for d in range(xx.shape[1])], axis=1)


def moving_average_columns(xx: np.ndarray, size: int = FILTER_WINDOW) -> np.ndarray:
    """control.py:329-344 (`_moving_average_filter`, unused by the reference's step but kept as an
    option): np.convolve(x, ones/size, 'same') with the edge rows rescaled to proper means, i.e. the mean
    of x over [n - size//2, n + (size-1)//2] clipped to the array.  Needs len(x) >= size."""
    T = xx.shape[0]
    back, fwd = size // 2, (size - 1) // 2
    out = np.zeros_like(xx)
    n_conv = -(-size // 2)
    for n in range(T):
        lo, hi = max(0, n - back), min(T - 1, n + fwd)
        acc = np.zeros(xx.shape[1])
        for k in range(lo, hi + 1):
            acc = acc + xx[k] * (1.0 / size)
        out[n] = acc
    out[0] *= size / n_conv
    for i in range(1, n_conv):
        out[i] *= size / (i + n_conv)
        out[-i] *= size / (i + n_conv - (size % 2))
    return out


def exploit_count(K: int, exploration: float) -> int:
    """Number of leading samples with ``k < (1-exploration)*K`` (control.py:98)."""
    thr = (1.0 - exploration) * K
    return int(sum(1 for k in range(K) if k < thr)) if K <= 4096 else int(np.count_nonzero(np.arange(K) < thr))


def check_sigma(sigma, dim_u=2):
    """control.py:157-159."""
    if sigma.shape[0] != sigma.shape[1] or sigma.shape[0] != dim_u or dim_u < 1:
        print("[ERROR] sigma must be a square matrix with the size of size_dim_u.")
        raise ValueError


def update_waypoint(c: OracleMPPI, q1, q2) -> int:
    """control.py:75 + 200-232 with update_prev_idx=True (prints omitted)."""
    x, y = end_effector(q1, q2, c.cost_l1, c.cost_l2)
    j = int(nearest_in_window(window_of(c.ref_path, c.prev_waypoints_idx), x, y))
    c.prev_waypoints_idx += j
    return c.prev_waypoints_idx


def _visual_rollouts(c, x0, u, v):
    """control.py:129-145 including the index wrap (t=0 uses the *last* control, quirks Q3/Q4)."""
    T, K = c.T, c.K
    opt = np.zeros((T, 4))
    if c.visualize_optimal_traj:
        s = tuple(float(a) for a in x0)
        for t in range(T):
            s = arm_step(*s, u[t - 1, 0], u[t - 1, 1], c.arm, c.delta_t, c.dynamics)
            opt[t] = s
    if c.visualze_sampled_trajs:
        samp = np.zeros((K, T, 4))
        s = tuple(np.full(K, float(a)) for a in x0)
        for t in range(T):
            s = arm_step(*s, v[:, t - 1, 0], v[:, t - 1, 1], c.arm, c.delta_t, c.dynamics)
            samp[:, t, :] = np.stack(s, axis=1)
    else:
        samp = np.broadcast_to(0.0, (K, T, 4))      # reference allocates zeros (control.py:137)
    return opt, samp


def _finish(c, x0, eps, v, S, out):
    w, rho, eta = softmin_weights(S, c.param_lambda)
    raw = np.einsum("k,ktm->tm", w, eps) if eps.shape[0] > 512 else _weighted_sum_loops(w, eps)
    filt = {"median": filter_columns, "average": moving_average_columns, "none": lambda a: a.copy()}[c.smoother](raw)
    u = c.u_prev                                   # alias (Q1)
    u += filt                                      # control.py:126
    u_new = u.copy()
    opt, samp = _visual_rollouts(c, x0, u, v)
    c.u_prev[:-1] = u[1:]                          # control.py:148
    c.u_prev[-1] = u[-1]                           # control.py:149
    out.update(S=S, w=w, rho=rho, eta=eta, w_eps_raw=raw, w_eps_filt=filt, u_new=u_new,
               u0=u[0].copy(), u_seq_returned=u, optimal_traj=opt, sampled_traj=samp,
               prev_idx_after=c.prev_waypoints_idx)
    return out


def _weighted_sum_loops(w, eps):
    """control.py:115-118 in the reference's summation order (k ascending for each t)."""
    K, T, m = eps.shape
    acc = np.zeros((T, m))
    for k in range(K):
        acc += w[k] * eps[k]
    return acc


def _enter(c: OracleMPPI, observed_x, eps):
    x0 = np.asarray(observed_x, dtype=np.float64)
    out = {"u_prev_before": c.u_prev.copy(), "prev_idx_before": c.prev_waypoints_idx}
    update_waypoint(c, x0[0], x0[1])
    if c.prev_waypoints_idx >= c.ref_path.shape[0] - 1:              # control.py:76-78
        print("[ERROR] Reached the end of the reference path.")
        raise IndexError
    check_sigma(c.sigma)
    eps = np.asarray(eps, dtype=np.float64)
    assert eps.shape == (c.K, c.T, 2)
    return x0, eps, out


def step_loops(c: OracleMPPI, observed_x, eps) -> dict:
    """Scalar restatement of control.py:67-152 — one Python iteration per (k, t)."""
    x0, eps, out = _enter(c, observed_x, eps)
    K, T = c.K, c.T
    u = c.u_prev
    sig_inv = np.linalg.inv(c.sigma)
    win = window_of(c.ref_path, c.prev_waypoints_idx)
    S = np.zeros(K)
    v = np.zeros((K, T, 2))
    for k in range(K):
        s = tuple(float(a) for a in x0)
        for t in range(T):
            if k < (1.0 - c.param_exploration) * K:
                v[k, t] = u[t] + eps[k, t]
            else:
                v[k, t] = eps[k, t]
            s = arm_step(*s, v[k, t, 0], v[k, t, 1], c.arm, c.delta_t, c.dynamics)
            S[k] += float(tracking_cost(*s, win, c.stage_cost_weight, c.cost_l1, c.cost_l2)) \
                + c.param_gamma * u[t] @ sig_inv @ v[k, t] + float(_stage_extra(c, s))
        S[k] += float(tracking_cost(*s, win, c.terminal_cost_weight, c.cost_l1, c.cost_l2))
    return _finish(c, x0, eps, v, S, out)


def step_vectorized(c: OracleMPPI, observed_x, eps) -> dict:
    """Same step with the K loop vectorised (t stays sequential: it is a recurrence)."""
    x0, eps, out = _enter(c, observed_x, eps)
    K, T = c.K, c.T
    u = c.u_prev
    sig_inv = np.linalg.inv(c.sigma)
    win = window_of(c.ref_path, c.prev_waypoints_idx)
    n_exploit = exploit_count(K, c.param_exploration)
    v = eps.copy()
    v[:n_exploit] += u[None, :, :]
    s = tuple(np.full(K, float(a)) for a in x0)
    S = np.zeros(K)
    for t in range(T):
        s = arm_step(*s, v[:, t, 0], v[:, t, 1], c.arm, c.delta_t, c.dynamics)
        ctrl = c.param_gamma * ((u[t] @ sig_inv) @ v[:, t, :].T)
        S += tracking_cost(*s, win, c.stage_cost_weight, c.cost_l1, c.cost_l2) + ctrl + _stage_extra(c, s)
    S += tracking_cost(*s, win, c.terminal_cost_weight, c.cost_l1, c.cost_l2)
    return _finish(c, x0, eps, v, S, out)


def rollout_costs(c: OracleMPPI, x0, eps, prev_idx=None, u=None) -> np.ndarray:
    """Costs S[K] only (no state mutation) — for large-K checks where only S is compared."""
    u = c.u_prev if u is None else np.asarray(u, dtype=np.float64)
    p = c.prev_waypoints_idx if prev_idx is None else prev_idx
    eps = np.asarray(eps, dtype=np.float64)
    K, T = eps.shape[0], eps.shape[1]
    sig_inv = np.linalg.inv(c.sigma)
    win = window_of(c.ref_path, p)
    n_exploit = exploit_count(K, c.param_exploration)
    v = eps.copy()
    v[:n_exploit] += u[None, :, :]
    s = tuple(np.full(K, float(a)) for a in x0)
    S = np.zeros(K)
    for t in range(T):
        s = arm_step(*s, v[:, t, 0], v[:, t, 1], c.arm, c.delta_t, c.dynamics)
        S += tracking_cost(*s, win, c.stage_cost_weight, c.cost_l1, c.cost_l2) \
            + c.param_gamma * ((u[t] @ sig_inv) @ v[:, t, :].T) + _stage_extra(c, s)
    S += tracking_cost(*s, win, c.terminal_cost_weight, c.cost_l1, c.cost_l2)
    return S


# ----------------------------------------------------------------------------------------------
# plant used by run.py:53-59 (for closed-loop tests) — utils.py:14-38
# ----------------------------------------------------------------------------------------------
def plant_step(q, dq, u, dt, arm=None):
    arm = arm or default_arm_params()
    a1, a2 = arm_accel(q[0], q[1], dq[0], dq[1], u[0], u[1], arm)
    dq = dq + dt * np.array([a1, a2])
    q = q + dt * dq
    return q, dq


def run_py_settings(ref_path, K=100, T=30, **over) -> dict:
    """Constructor keywords of run.py:25-37."""
    kw = dict(delta_t=0.003 * 2, ref_path=ref_path, horizon_step_T=T, number_of_samples_K=K,
              param_exploration=0.0, param_lambda=100.0, param_alpha=0.98,
              sigma=np.array([[20.0, 0.0], [0.0, 20.0]]),
              stage_cost_weight=np.array([0.50, 0.50, 5.0, 5.0]),
              terminal_cost_weight=np.array([5.0, 5.0, 50.0, 50.0]))
    kw.update(over)
    return kw


RUN_PY_X0 = (1.152198236517471885e+00, -1.266101672070702344e+00, 0.0, 0.0)   # run.py:14-15


def injected_noise(seed: int, K: int, T: int, sigma) -> np.ndarray:
    """Deterministic FP32 noise tensor [K,T,2] ~ N(0, sigma) used by the golden vectors and parity
    tests (the reference's own draw, control.py:163, is unseeded).  Returned as float32 so the CUDA
    path and the FP64 oracle (after ``astype(float64)``) consume bit-identical values."""
    L = np.linalg.cholesky(np.asarray(sigma, dtype=np.float64)).astype(np.float32)
    z = np.random.default_rng(seed).standard_normal((K, T, 2)).astype(np.float32)
    return (z @ L.T).astype(np.float32)

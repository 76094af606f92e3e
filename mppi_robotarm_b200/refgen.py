"""Reference-trajectory generation — the producers of the reference's data files,
restated from utils.py:41-93 so benchmarks and users can synthesise references of any length in the
same formats (App. D).  Host-side NumPy: this runs once, before the control loop.

* ``circle_joint_reference``  -> rows (q1, q2, x, y): the ideal circle of ``trajectory.txt``
  (inverse kinematics of utils.py:41-62 over Theta_k = 2*pi*k/(n-1), with its snap zone).
* ``record_tracking_run``     -> rows (x, y, dq1, dq2, u1, u2): a computed-torque + PD tracking run
  of that circle (utils.py:65-93) recorded at Ts, the shape of ``xydq_circle.txt``.
"""
from __future__ import annotations

import numpy as np

from .arm_params import SYS_PARAMS


def circle_point(theta):
    """End-effector target of utils.py:45-52: circle of radius 0.6 about (0.8, 0.8); within 0.2 rad of a
    full turn it snaps to the start point (1.4, 0.8); beyond, to the stretched-out pose (2, 0)."""
    theta = np.asarray(theta, dtype=np.float64)
    xe = 0.8 + 0.6 * np.cos(theta)
    ye = 0.8 + 0.6 * np.sin(theta)
    snap = (theta >= 2 * np.pi - 0.2) & (theta <= 2 * np.pi + 0.2)
    far = theta > 2 * np.pi + 0.2
    xe = np.where(snap, 1.4, np.where(far, 2.0, xe))
    ye = np.where(snap, 0.8, np.where(far, 0.0, ye))
    return xe, ye


def inverse_kinematics(xe, ye, l1=1.0, l2=1.0):
    """Elbow-down joint angles (q1, q2) of the planar 2-link arm reaching (xe, ye) — the closed form of
    utils.py:54-61 written with the law of cosines: with R^2 = xe^2 + ye^2,
    s = sqrt((2 l1 l2)^2 - (R^2 - l1^2 - l2^2)^2), a_{1,2} = 2 atan((2 ye l1 +- s) / ((xe + l1)^2 + ye^2 - l2^2)),
    q1 = a_1, q2 = a_2 - a_1."""
    xe, ye = np.asarray(xe, dtype=np.float64), np.asarray(ye, dtype=np.float64)
    r2 = xe * xe + ye * ye
    s = np.sqrt(np.maximum((2 * l1 * l2) ** 2 - (r2 - l1 * l1 - l2 * l2) ** 2, 0.0))
    den = (xe + l1) ** 2 + ye * ye - l2 * l2
    a1 = 2 * np.arctan((2 * ye * l1 + s) / den)
    a2 = 2 * np.arctan((2 * ye * l1 - s) / den)
    return a1, a2 - a1


def circle_joint_reference(n: int = 3000, snap: bool = True, turn: float = 1.0) -> np.ndarray:
    """[n, 4] rows (q1, q2, x, y) — the layout of trajectory.txt (snap=True, one full turn).
    snap=False keeps the pure circle (no jump to the start point near the end of the turn)."""
    theta = turn * 2 * np.pi * np.arange(n) / (n - 1)
    if snap:
        xe, ye = circle_point(theta)
    else:
        xe, ye = 0.8 + 0.6 * np.cos(theta), 0.8 + 0.6 * np.sin(theta)
    q1, q2 = inverse_kinematics(xe, ye)
    return np.stack([q1, q2, xe, ye], axis=1)


def computed_torque(q, dq, v, params: dict | None = None) -> np.ndarray:
    """u = M(q) v + C(q, dq) dq + G(q): the feedback-linearising torque of utils.py:65-84."""
    p = params or SYS_PARAMS()
    m1, m2, l1, l2, lc1, lc2, g = (p[k] for k in ("m1", "m2", "l1", "l2", "lc1", "lc2", "g"))
    q, dq, v = (np.asarray(a, dtype=np.float64) for a in (q, dq, v))
    c2, s2 = np.cos(q[1]), np.sin(q[1])
    m12 = m2 * l1 * lc2 * c2 + m2 * lc2 ** 2 + l2
    M = np.array([[m1 * lc1 ** 2 + l1 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * c2) + l2, m12],
                  [m12, m2 * lc2 ** 2 + l2]])
    h = m2 * l1 * lc2 * s2
    C = np.array([[-h * dq[1], -h * dq[0] - h * dq[1]], [h * dq[0], 0.0]])
    G = np.array([m1 * lc1 * g * np.cos(q[0]) + m2 * g * (lc2 * np.cos(q[0] + q[1]) + l1 * np.cos(q[0])),
                  m2 * lc2 * g * np.cos(q[0] + q[1])])
    return M @ v + C @ dq + G


def pd_outer_loop(q, dq, r, dr, ddr, kp: float = 100.0, kd: float = 20.0) -> np.ndarray:
    """v = ddr - kd (dq - dr) - kp (q - r) (utils.py:87-93)."""
    q, dq, r, dr, ddr = (np.asarray(a, dtype=np.float64) for a in (q, dq, r, dr, ddr))
    return ddr - kd * (dq - dr) - kp * (q - r)


def record_tracking_run(n: int = 2000, Ts: float = 0.0025, kp: float = 100.0, kd: float = 20.0,
                        params: dict | None = None) -> np.ndarray:
    """[n, 6] rows (x, y, dq1, dq2, u1, u2): the arm tracks the circle under the reference's computed-torque
    law u = M v + C dq + G (utils.py:65-84) with v = ddr - kd (dq - dr) - kp (q - r) (utils.py:87-93),
    integrated with the semi-implicit Euler step of run.py:53-55 at Ts."""
    from utils import Arm_Dynamic          # this repo's plant twin (FP64)
    p = params or SYS_PARAMS()
    l1, l2 = p["l1"], p["l2"]
    # the recorded file covers theta in [0, 6.2546]: a smooth circle, no snap zone
    tgt = circle_joint_reference(n + 2, snap=False, turn=0.9955)[:, 0:2]
    r = tgt[:n]
    dr = np.gradient(tgt, Ts, axis=0)[:n]
    ddr = np.gradient(np.gradient(tgt, Ts, axis=0), Ts, axis=0)[:n]
    q, dq = r[0].copy(), np.zeros(2)
    out = np.zeros((n, 6))
    for k in range(n):
        u = computed_torque(q, dq, pd_outer_loop(q, dq, r[k], dr[k], ddr[k], kp, kd), p)
        dq = dq + Ts * Arm_Dynamic(q, dq, u)
        q = q + Ts * dq
        out[k] = (l1 * np.cos(q[0]) + l2 * np.cos(q[0] + q[1]), l1 * np.sin(q[0]) + l2 * np.sin(q[0] + q[1]),
                  dq[0], dq[1], u[0], u[1])
    return out

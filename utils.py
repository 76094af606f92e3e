"""Plant-side helpers with the names run.py imports via `from utils import *` (reference:
utils.py:14-38).  These run on the host in FP64 — they are the simulated *plant* of run.py:53-57,
not part of the MPPI step."""
import numpy as np

from sys_params import SYS_PARAMS

_P = SYS_PARAMS()


def Arm_Dynamic(q, dq, u):
    """Joint accelerations M(q)^-1 (u - C(q,dq) dq - G(q)) of the 2-link arm (utils.py:14-29);
    the 2x2 solve is written out with the adjugate."""
    m1, m2, l1, l2, lc1, lc2, g = (_P[k] for k in ("m1", "m2", "l1", "l2", "lc1", "lc2", "g"))
    c2, s2 = np.cos(q[1]), np.sin(q[1])
    a = m1 * lc1 ** 2 + l1 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * c2) + l2
    d = m2 * lc2 ** 2 + l2
    b = m2 * l1 * lc2 * c2 + d
    h = m2 * l1 * lc2 * s2
    grav = np.array([m1 * lc1 * g * np.cos(q[0]) + m2 * g * (lc2 * np.cos(q[0] + q[1]) + l1 * np.cos(q[0])),
                     m2 * lc2 * g * np.cos(q[0] + q[1])])
    cor = np.array([-h * dq[1] * dq[0] + (-h * dq[0] - h * dq[1]) * dq[1], h * dq[0] * dq[0]])
    rhs = np.asarray(u, dtype=float) - cor - grav
    det = a * d - b * b
    return np.array([d * rhs[0] - b * rhs[1], a * rhs[1] - b * rhs[0]]) / det


def Forward_Kinemetic(q):
    """Elbow and end-effector positions (utils.py:32-38)."""
    l1, l2 = _P["l1"], _P["l2"]
    x1, y1 = l1 * np.cos(q[0]), l1 * np.sin(q[0])
    return x1, y1, x1 + l2 * np.cos(q[0] + q[1]), y1 + l2 * np.sin(q[0] + q[1])


# The three helpers run.py pulls in with `from utils import *` but never calls (utils.py:41-93): the producers of
# the reference's data files.  Same names, arguments and return values; the arithmetic lives in refgen.py.
def Inverse_Kinemetic(Theta):
    """(r, XE, YE): joint target r = [q1, q2] reaching the circle point of angle Theta (utils.py:41-62)."""
    from mppi_robotarm_b200 import refgen
    xe, ye = refgen.circle_point(Theta)
    q1, q2 = refgen.inverse_kinematics(xe, ye)
    return np.array([float(q1), float(q2)]), float(xe), float(ye)


def Feedback_linearization(q, dq, v):
    """Computed-torque input u = M(q) v + C(q, dq) dq + G(q) (utils.py:65-84)."""
    from mppi_robotarm_b200 import refgen
    return refgen.computed_torque(q, dq, v, _P)


def Controller(q, dq, r, dr, ddr):
    """PD outer loop v = ddr - 20 (dq - dr) - 100 (q - r) (utils.py:87-93)."""
    from mppi_robotarm_b200 import refgen
    return refgen.pd_outer_loop(q, dq, r, dr, ddr)

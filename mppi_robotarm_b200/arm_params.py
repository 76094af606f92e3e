"""Arm constants of the 2-link planar arm (reference: sys_params.py:1-13).

The drop-in ``control.py`` prefers the user's own ``sys_params.SYS_PARAMS()`` when one is importable
(that is how the reference reads them, control.py:11-18); this table is the stand-in with the same
keys and values for running this repository on its own.
"""

_KEYS = ("Ts", "m1", "m2", "l1", "l2", "lc1", "lc2", "g")
_VALUES = (0.0025, 1, 1, 1, 1, 0.5, 0.5, 9.81)


def SYS_PARAMS() -> dict:
    """A fresh dict per call, like the reference: sample time (unused by the controller), link
    masses, link lengths, centre-of-mass offsets and gravity."""
    return dict(zip(_KEYS, _VALUES))


def arm_vector(params: dict) -> list:
    """(m1, m2, l1, l2, lc1, lc2, g) in the order the C ABI expects (include/mppi_b200.h)."""
    return [float(params[k]) for k in ("m1", "m2", "l1", "l2", "lc1", "lc2", "g")]

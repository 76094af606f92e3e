"""TEST INFRASTRUCTURE ONLY — harness that imports the *unmodified* reference controller.

Only usable where ``/root/reference`` exists (this build container, not the GPU box).
It is used by ``tests/golden/make_golden.py`` to produce the committed golden vectors and by
``tests/test_oracle_vs_reference.py`` to pin ``oracle/mppi_oracle.py`` against the real thing.
Nothing in the product package imports this module.

How it works: the reference's ``control.py:3-7`` imports matplotlib and
IPython at module top although the MPPI step never touches them; neither is installed here, so we
register empty stand-in modules before importing.  Noise is injected by replacing the bound method
``_calc_epsilon`` (its only call site is ``control.py:84``), and intermediates are captured by
wrapping ``_compute_weights`` (``control.py:112``) and ``_moving_median_filter`` (``control.py:122``).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

REFERENCE_DIR = os.environ.get("MPPI_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "control.py"))


def _install_stubs() -> None:
    names = ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.animation",
             "IPython", "IPython.display"]
    for n in names:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
    if not hasattr(sys.modules["matplotlib.animation"], "ArtistAnimation"):
        sys.modules["matplotlib.animation"].ArtistAnimation = object
    sys.modules["IPython"].display = sys.modules["IPython.display"]


def import_reference():
    """Return the reference's ``control`` and ``utils`` modules, loaded under private names so they
    never shadow this repo's own drop-in ``control.py`` / ``utils.py``."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_DIR}")
    import importlib.util
    _install_stubs()
    cached = sys.modules.get("_mppi_reference_control")
    if cached is not None:
        return cached, sys.modules["_mppi_reference_utils"]
    # The reference does `from sys_params import SYS_PARAMS`: give it its own sys_params module.
    saved = {k: sys.modules.get(k) for k in ("sys_params",)}
    spec = importlib.util.spec_from_file_location("sys_params", os.path.join(REFERENCE_DIR, "sys_params.py"))
    sp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sp)
    sys.modules["sys_params"] = sp
    try:
        mods = []
        for name in ("control", "utils"):
            spec = importlib.util.spec_from_file_location(f"_mppi_reference_{name}",
                                                          os.path.join(REFERENCE_DIR, f"{name}.py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules[f"_mppi_reference_{name}"] = m
            spec.loader.exec_module(m)
            mods.append(m)
    finally:
        if saved["sys_params"] is None:
            sys.modules.pop("sys_params", None)
        else:
            sys.modules["sys_params"] = saved["sys_params"]
    return mods[0], mods[1]


def load_reference_data(name: str) -> np.ndarray:
    return np.loadtxt(os.path.join(REFERENCE_DIR, name))


class ReferenceProbe:
    """Owns one reference controller and records the intermediates of each step."""

    def __init__(self, **ctor_kwargs):
        control, _ = import_reference()
        self.ctrl = control.MPPIControllerForPathTracking(**ctor_kwargs)
        self._eps = None
        self.last = {}
        orig_w = self.ctrl._compute_weights
        orig_f = self.ctrl._moving_median_filter

        def cap_w(S):
            self.last["S"] = np.array(S, copy=True)
            w = orig_w(S)
            self.last["w"] = np.array(w, copy=True)
            return w

        def cap_f(xx, window_size):
            self.last["w_eps_raw"] = np.array(xx, copy=True)
            out = orig_f(xx=xx, window_size=window_size)
            self.last["w_eps_filt"] = np.array(out, copy=True)
            return out

        self.ctrl._compute_weights = cap_w
        self.ctrl._moving_median_filter = cap_f

    def step(self, observed_x, eps=None):
        """One reference step.  ``eps`` ([K,T,2] float64) is injected when given; otherwise the
        reference's own unseeded ``np.random.multivariate_normal`` draw is used."""
        c = self.ctrl
        if eps is not None:
            eps = np.asarray(eps, dtype=np.float64)
            c._calc_epsilon = lambda *a, **k: eps
        self.last = {"u_prev_before": np.array(c.u_prev, copy=True),
                     "prev_idx_before": int(c.prev_waypoints_idx)}
        with contextlib.redirect_stdout(io.StringIO()):
            u0, useq, opt, samp = c.calc_control_input(observed_x)
        self.last.update(u0=np.array(u0, copy=True), u_seq_returned=np.array(useq, copy=True),
                         optimal_traj=np.array(opt, copy=True), sampled_traj=samp,
                         prev_idx_after=int(c.prev_waypoints_idx),
                         returned_is_alias=bool(useq is c.u_prev))
        # u before the shift (control.py:126) = u_prev_before + filtered update
        self.last["u_new"] = self.last["u_prev_before"] + self.last["w_eps_filt"]
        return self.last

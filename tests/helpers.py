"""Shared helpers of the parity tests (test infrastructure; imports the oracle, never shipped)."""
import contextlib
import io

import numpy as np

from oracle import mppi_oracle as mo
from tests.golden import cases


def make_controller(case, paths, **extra):
    """Our drop-in controller configured like a golden case, with the reference's injection seam."""
    from control import MPPIControllerForPathTracking
    kw = cases.ctor_kwargs(case, paths)
    if "smoother" in case:
        extra = dict(extra, smoother=case["smoother"])
    if "dynamics" in case:
        extra = dict(extra, dynamics=case["dynamics"])
    ctrl = MPPIControllerForPathTracking(**kw, noise="numpy", verbose=False, **extra)
    if "prev_idx" in case:
        ctrl.prev_waypoints_idx = case["prev_idx"]
    if "u_prev" in case:
        ctrl.u_prev = np.array(case["u_prev"], dtype=np.float64)
    return ctrl, kw


def inject(ctrl, eps32):
    """Replace the bound _calc_epsilon exactly like the oracle harness does for the reference."""
    e64 = np.asarray(eps32).astype(np.float64)
    ctrl._calc_epsilon = lambda *a, **k: e64


def quiet_step(ctrl, x):
    with contextlib.redirect_stdout(io.StringIO()):
        return ctrl.calc_control_input(x)


def rel_err(a, b):
    b = np.asarray(b)
    return float(np.max(np.abs(np.asarray(a) - b)) / max(np.max(np.abs(b)), 1e-300))


def combine_partials(parts, lam):
    """NumPy mirror of the device combine (mppi_finalize_sm100a) for the CPU-side distributed test:
    parts = [(rho_g, eta_g, V_g[T,2]), ...] -> (rho, eta, V/eta)."""
    rho = min(p[0] for p in parts)
    eta, V = 0.0, 0.0
    for r, e, v in parts:
        s = np.exp(-(r - rho) / lam)
        eta += s * e
        V = V + s * np.asarray(v)
    return rho, eta, V / eta


def oracle_partial(c: mo.OracleMPPI, x0, eps, k0, k1, prev_idx):
    """(rho_g, eta_g, V_g) of samples [k0, k1) computed with the FP64 oracle."""
    S = mo.rollout_costs(c, np.asarray(x0, dtype=np.float64), eps, prev_idx=prev_idx)[k0:k1]
    rho = S.min()
    e = np.exp(-(S - rho) / c.param_lambda)
    return rho, e.sum(), np.einsum("k,ktm->tm", e, eps[k0:k1])

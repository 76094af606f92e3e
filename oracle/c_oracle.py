"""TEST INFRASTRUCTURE ONLY — ctypes wrapper of oracle/mppi_oracle.c (plain-C FP64 restatement,
OpenMP over samples).  Used by tests (cross-check against mppi_oracle.py) and by bench.py's
cpu_baseline / --impl reference legs.  Never imported by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libmppi_oracle.so")


class OracleCfg(C.Structure):
    _fields_ = [("dt", C.c_double), ("lam", C.c_double), ("gamma", C.c_double), ("sig_inv", C.c_double * 4),
                ("ws", C.c_double * 4), ("wt", C.c_double * 4),
                ("m1", C.c_double), ("m2", C.c_double), ("l1", C.c_double), ("l2", C.c_double),
                ("lc1", C.c_double), ("lc2", C.c_double), ("g", C.c_double), ("cl1", C.c_double), ("cl2", C.c_double),
                ("jl_lo", C.c_double * 2), ("jl_hi", C.c_double * 2), ("jl_w", C.c_double)]


def build(force=False):
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "mppi_oracle.c")):
        subprocess.run(["make", "-C", HERE, "-B"], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_weighted_sum.restype = C.c_double
    return _lib


def make_cfg(c) -> OracleCfg:
    """c: oracle.mppi_oracle.OracleMPPI"""
    o = OracleCfg()
    o.dt, o.lam, o.gamma = c.delta_t, c.param_lambda, c.param_gamma
    o.sig_inv[:] = np.linalg.inv(c.sigma).reshape(-1).tolist()
    o.ws[:] = list(c.stage_cost_weight)
    o.wt[:] = list(c.terminal_cost_weight)
    for k in ("m1", "m2", "l1", "l2", "lc1", "lc2", "g"):
        setattr(o, k, float(c.arm[k]))
    o.cl1, o.cl2 = float(c.cost_l1), float(c.cost_l2)
    o.jl_lo[:] = [float(v) for v in getattr(c, "joint_limit_lo", (-np.inf, -np.inf))]
    o.jl_hi[:] = [float(v) for v in getattr(c, "joint_limit_hi", (np.inf, np.inf))]
    o.jl_w = float(getattr(c, "joint_limit_weight", 0.0))
    return o


def rollout_costs(c, x0, eps, prev_idx, n_exploit=None):
    """FP64 costs S[K] of the samples eps [K,T,2] for window start prev_idx (already updated)."""
    if getattr(c, "dynamics", "F") != "F":
        raise ValueError("mppi_oracle.c restates the default rollout model _F only (use mppi_oracle.py for _F1)")
    eps = np.ascontiguousarray(eps, dtype=np.float64)
    K, T = eps.shape[0], eps.shape[1]
    win = np.ascontiguousarray(c.ref_path[prev_idx:prev_idx + 30, 0:4])
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    u = np.ascontiguousarray(c.u_prev, dtype=np.float64)
    S = np.zeros(K)
    cfg = make_cfg(c)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))   # noqa: E731
    lib().oracle_rollout_costs(C.byref(cfg), dp(win), win.shape[0], dp(x0), dp(u), dp(eps), K, T,
                               K if n_exploit is None else n_exploit, dp(S))
    return S


def weighted_sum(S, eps, lam):
    eps = np.ascontiguousarray(eps, dtype=np.float64)
    K, T = eps.shape[0], eps.shape[1]
    w = np.zeros(K)
    out = np.zeros((T, 2))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))   # noqa: E731
    rho = lib().oracle_weighted_sum(dp(np.ascontiguousarray(S)), dp(eps), K, T, C.c_double(lam), dp(w), dp(out))
    return rho, w, out


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP team size of the next calls (bench.py sets it from the affinity mask: torchrun exports
    OMP_NUM_THREADS=1 to its workers)."""
    lib().oracle_set_num_threads(int(n))
    return num_threads()

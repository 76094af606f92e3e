"""Reference-trajectory files: whitespace text as written by np.savetxt('%.18e').

``xydq*.txt``       columns (x, y, dq1_ref, dq2_ref[, u1, u2]) — what the controller consumes
                    (run.py:18-19 takes [:, 0:4]).
``trajectory*.txt`` columns (q1, q2, x, y).  Read verbatim they are a valid but meaningless input;
                    ``layout="xydq"`` re-lays them out as (x, y, dq1/Ts, dq2/Ts) with finite
                    differences at the sample time Ts of sys_params.py:3.
"""
from __future__ import annotations

import os

import numpy as np


def relayout_qxy(traj: np.ndarray, Ts: float = 0.0025) -> np.ndarray:
    dq = np.gradient(traj[:, 0:2], Ts, axis=0)
    return np.concatenate([traj[:, 2:4], dq], axis=1)


def load_ref_path(path, layout: str = "auto", Ts: float = 0.0025) -> np.ndarray:
    """Load a reference path as float64 [N, 4] = (x, y, dq1_ref, dq2_ref).

    layout: "verbatim" = first four columns as they are (exactly run.py:18-19);
            "xydq"     = file holds (q1, q2, x, y): convert;
            "auto"     = "xydq" for files named trajectory*, else "verbatim".
    """
    arr = np.load(path) if str(path).endswith(".npy") else np.loadtxt(path)
    if layout == "auto":
        layout = "xydq" if os.path.basename(str(path)).startswith("trajectory") else "verbatim"
    if layout == "verbatim":
        return np.ascontiguousarray(arr[:, 0:4], dtype=np.float64)
    if layout == "xydq":
        return relayout_qxy(np.asarray(arr, dtype=np.float64), Ts)
    raise ValueError("layout must be 'auto', 'verbatim' or 'xydq'")

"""mppi_robotarm_b200 — B200-native (sm_100a) MPPI step for the 2-link arm of junofficial/mppi_RobotArm.

Only what the hot path needs lives here: the CUDA sources and C ABI (``csrc/``, ``include/``), the
ctypes binding, the engine that owns a handle, and the host mirrors of the reference's interface.
"""
from .controller import MPPIControllerForPathTracking     # noqa: F401
from .engine import MppiEngine, ShardSpec                 # noqa: F401
from .refpath import load_ref_path                        # noqa: F401

__all__ = ["MPPIControllerForPathTracking", "MppiEngine", "ShardSpec", "load_ref_path"]

"""ctypes binding of libmppi_b200.so (declarations: include/mppi_b200.h).

This is the only place the Python host touches native code.  There is no fallback: if the library
is missing or was not built, importing the controller raises with the build instruction.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

ABI_VERSION = 3
MAX_T = 256

NOISE_PHILOX = 0
NOISE_INJECTED = 1
FLAG_OPTIMAL_TRAJ = 1
FLAG_DEVICE_GRAPH = 2
FLAG_SMOOTH_AVERAGE = 4
FLAG_SMOOTH_NONE = 8
FLAG_FULL_SEARCH = 16
FLAG_DYNAMICS_F1 = 32
FLAG_SEARCH_STATS = 64
FLAG_RESIDENT_STATE = 128

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_WORKSPACE = 0, -1, -2, -3, -4


class MppiConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("n_env", C.c_int32),
        ("K_total", C.c_int32), ("K_local", C.c_int32), ("k_offset", C.c_int32),
        ("T", C.c_int32), ("n_exploit", C.c_int32), ("flags", C.c_int32), ("max_ref_rows", C.c_int32),
        ("delta_t", C.c_double), ("param_lambda", C.c_double), ("param_gamma", C.c_double),
        ("sigma_chol", C.c_double * 4), ("sigma_inv", C.c_double * 4),
        ("stage_cost_weight", C.c_double * 4), ("terminal_cost_weight", C.c_double * 4),
        ("arm", C.c_double * 7), ("cost_l1", C.c_double), ("cost_l2", C.c_double),
        ("seed", C.c_uint64),
        ("joint_limit_lo", C.c_double * 2), ("joint_limit_hi", C.c_double * 2), ("joint_limit_weight", C.c_double),
    ]


class MppiIoLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in (
        "bytes", "off_x0", "off_u_prev", "off_prev_idx", "off_step", "off_new_idx", "off_status", "off_rho", "off_eta",
        "off_u0", "off_w_eps_raw", "off_w_eps_filt", "off_u_new", "off_opt_traj")]


# name -> (restype, argtypes); must list every symbol include/mppi_b200.h declares
SYMBOLS = {
    "mppi_abi_version": (C.c_int, []),
    "mppi_device_count": (C.c_int, []),
    "mppi_workspace_bytes": (C.c_size_t, [C.POINTER(MppiConfig)]),
    "mppi_io_layout": (C.c_int, [C.POINTER(MppiConfig), C.POINTER(MppiIoLayout)]),
    "mppi_create": (C.c_int, [C.POINTER(MppiConfig), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                              C.POINTER(C.c_void_p)]),
    "mppi_destroy": (None, [C.c_void_p]),
    "mppi_last_error": (C.c_char_p, [C.c_void_p]),
    "mppi_set_ref_path": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "mppi_step": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "mppi_step_local": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mppi_step_combine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mppi_wait": (C.c_int, [C.c_void_p]),
    "mppi_upload_state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mppi_download_state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mppi_closed_loop": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mppi_exchange_bytes": (C.c_size_t, [C.POINTER(MppiConfig), C.c_int32]),
    "mppi_set_peer_exchange": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "mppi_step_sharded": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "mppi_exchange_status": (C.c_int, [C.c_void_p]),
    "mppi_set_exchange_timeout": (C.c_int, [C.c_void_p, C.c_double]),
    "mppi_set_capture_mode": (C.c_int, [C.c_void_p, C.c_int32]),
    "mppi_replay_begin": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mppi_replay_end": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mppi_last_costs": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "mppi_step_block": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "mppi_sampled_trajectories": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mppi_sampled_trajectories_subset": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                                   C.c_void_p, C.c_void_p]),
    "mppi_philox_noise": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "mppi_launch_count": (C.c_uint64, [C.c_void_p]),
    "mppi_search_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_int32, C.c_void_p]),
    "mppi_set_timing": (C.c_int, [C.c_void_p, C.c_int32]),
    "mppi_get_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int32]),
    "mppi_probe_fp32": (C.c_int, [C.c_int32, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def library_path() -> str:
    return os.environ.get("MPPI_B200_LIB", _build.LIB_PATH)


def load() -> C.CDLL:
    """Load libmppi_b200.so and bind every declared symbol.  Raises if it is not there — the MPPI
    step has no Python/NumPy implementation in this package."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.isfile(path) and "MPPI_B200_LIB" not in os.environ:
        try:                                   # fresh checkout: compile the CUDA library once (nvcc, a few seconds)
            _build.build_library()
        except Exception as ex:                # noqa: BLE001
            raise NativeLibraryError(
                f"{path} not found and building it failed ({ex}). Build it with `python -m mppi_robotarm_b200.build` "
                "(needs nvcc; the library is sm_100a CUDA and there is no CPU fallback).") from ex
    if not os.path.isfile(path):
        raise NativeLibraryError(
            f"{path} not found. Build it with `python -m mppi_robotarm_b200.build` (needs nvcc; the "
            "library is sm_100a CUDA and there is no CPU fallback).")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError = stale library: rebuild
        fn.restype = res
        fn.argtypes = args
    if lib.mppi_abi_version() != ABI_VERSION:
        raise NativeLibraryError(f"{path}: ABI {lib.mppi_abi_version()} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def last_error(handle=None) -> str:
    msg = load().mppi_last_error(handle)
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, handle=None, what: str = "") -> None:
    if rc == OK:
        return
    msg = f"libmppi_b200 {what} failed (code {rc}): {last_error(handle if rc != ERR_NO_DEVICE or handle else None)}"
    if rc == ERR_NO_DEVICE:
        raise NativeLibraryError(msg)
    if rc in (ERR_INVALID, ERR_WORKSPACE):
        raise ValueError(msg)
    raise RuntimeError(msg)

"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, sizes its buffers, and — with no GPU — refuses to run instead of falling back to a CPU
path.  No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from mppi_robotarm_b200 import _cabi, build
from tests.golden import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cuda():
    import torch
    return torch.cuda.is_available()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "mppi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mppi_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_in_tree():
    assert os.path.isfile(build.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.dirname(build.LIB_PATH) == os.path.join(ROOT, "mppi_robotarm_b200")


def test_every_declared_symbol_is_exported_and_bound():
    lib = _cabi.load()
    names = header_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mppi_b200.h but not exported"
    assert sorted(_cabi.SYMBOLS) == names, "python binding table and header disagree"
    assert lib.mppi_abi_version() == _cabi.ABI_VERSION


def _cfg(K=4096, T=50, n_env=1, K_local=None, k_offset=0):
    c = _cabi.MppiConfig()
    c.abi_version = _cabi.ABI_VERSION
    c.device, c.n_env, c.K_total, c.K_local, c.k_offset, c.T = 0, n_env, K, K_local or K, k_offset, T
    c.n_exploit, c.flags, c.max_ref_rows = K, 1, 2000
    c.delta_t, c.param_lambda, c.param_gamma = 0.006, 100.0, 2.0
    c.sigma_chol[:] = [4.47, 0, 0, 4.47]
    c.sigma_inv[:] = [0.05, 0, 0, 0.05]
    return c


def test_layout_and_workspace_sizing():
    lib = _cabi.load()
    lay = _cabi.MppiIoLayout()
    c = _cfg()
    assert lib.mppi_io_layout(C.byref(c), C.byref(lay)) == 0
    offs = [lay.off_x0, lay.off_u_prev, lay.off_prev_idx, lay.off_step, lay.off_new_idx, lay.off_status, lay.off_rho,
            lay.off_eta, lay.off_u0, lay.off_w_eps_raw, lay.off_w_eps_filt, lay.off_u_new, lay.off_opt_traj, lay.bytes]
    assert offs == sorted(offs) and len(set(offs)) == len(offs) and all(o % 8 == 0 for o in offs)
    # the compact results of a resident handle are the contiguous head of the outputs
    assert lay.off_w_eps_raw - lay.off_new_idx <= 256 and lay.off_u0 + 16 <= lay.off_w_eps_raw
    assert lay.off_u_prev - lay.off_x0 >= 4 * 8 and lay.bytes - lay.off_opt_traj >= 50 * 4 * 8
    ws = lib.mppi_workspace_bytes(C.byref(c))
    assert ws >= 2 * 4096 * 4 + 2000 * 32
    big = lib.mppi_workspace_bytes(C.byref(_cfg(K=1 << 20, T=100)))
    assert 8 * (1 << 20) <= big <= 64 * (1 << 20)       # Philox mode: S and w only, no K*T tensor


@pytest.mark.parametrize("mutate, why", [
    (lambda c: setattr(c, "T", 0), "T"), (lambda c: setattr(c, "T", _cabi.MAX_T + 1), "T"),
    (lambda c: setattr(c, "K_local", 5000), "shard"), (lambda c: setattr(c, "abi_version", 99), "abi"),
    (lambda c: setattr(c, "param_lambda", 0.0), "lambda"), (lambda c: setattr(c, "n_env", 0), "n_env"),
    (lambda c: setattr(c, "joint_limit_weight", -1.0), "joint_limit_weight"),
    (lambda c: (setattr(c, "joint_limit_weight", 1.0), c.joint_limit_lo.__setitem__(0, 2.0), c.joint_limit_hi.__setitem__(0, 1.0)), "lo <= hi"),
    (lambda c: (setattr(c, "joint_limit_weight", 1.0), setattr(c, "flags", c.flags | _cabi.FLAG_DYNAMICS_F1)), "_F only")])
def test_invalid_configs_are_rejected(mutate, why):
    lib = _cabi.load()
    c = _cfg()
    mutate(c)
    assert lib.mppi_workspace_bytes(C.byref(c)) == 0
    assert why.lower() in _cabi.last_error().lower()


@pytest.mark.skipif(_cuda(), reason="checks the no-GPU behaviour")
def test_no_gpu_means_loud_failure_not_cpu_fallback(paths):
    lib = _cabi.load()
    assert lib.mppi_device_count() == 0
    h = C.c_void_p()
    buf = (C.c_char * 4096)()
    rc = lib.mppi_create(C.byref(_cfg()), buf, 4096, buf, 4096, C.byref(h))
    assert rc == _cabi.ERR_NO_DEVICE and not h.value
    assert "no CPU fallback" in _cabi.last_error()
    from control import MPPIControllerForPathTracking
    ctrl = MPPIControllerForPathTracking(**cases.run_py_kwargs(cases.ref_path_for(paths, "xydq_circle.txt"), 8, 4),
                                         verbose=False)
    with pytest.raises(_cabi.NativeLibraryError):
        ctrl.calc_control_input(cases.X0)
    np.testing.assert_array_equal(ctrl.u_prev, np.tile([10.0, -2.0], (4, 1)))   # nothing was computed


def test_missing_library_raises(monkeypatch, tmp_path):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setenv("MPPI_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.NativeLibraryError, match="no CPU fallback"):
        _cabi.load()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mppi_robotarm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "oracle/" not in txt or f.endswith(".cuh"), f
    for f in ("control.py", "utils.py", "sys_params.py"):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(ROOT, f)).read(), flags=re.M)

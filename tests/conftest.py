import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _cuda_ok() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_ok():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def paths():
    from tests.golden import cases
    return cases.load_paths()


@pytest.fixture(scope="session")
def emul():
    """tests/emul: mppi_math.cuh compiled for the host (test infrastructure, rebuilt when stale)."""
    import ctypes as C
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    so, src = os.path.join(here, "emul", "_emul.so"), os.path.join(here, "emul", "emul.cpp")
    hdr = os.path.join(ROOT, "mppi_robotarm_b200", "csrc", "mppi_math.cuh")
    if not os.path.isfile(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src], check=True)
    lib = C.CDLL(so)
    lib.emul_rollout_costs.restype = C.c_int
    return lib

"""CPU tests of the FP32 arithmetic the kernels execute (mppi_math.cuh compiled for the host by
tests/emul): the stated FP32 tolerances hold against the FP64 oracle without needing a GPU, and the
Philox generator matches the published known-answer vectors."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import mppi_oracle as mo
from tests.golden import cases

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "emul", "_emul.so")
SRC = os.path.join(HERE, "emul", "emul.cpp")
HDR = os.path.join(os.path.dirname(HERE), "mppi_robotarm_b200", "csrc", "mppi_math.cuh")


@pytest.fixture(scope="module")
def emul():
    if not os.path.isfile(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", SO, SRC], check=True)
    lib = C.CDLL(SO)
    lib.emul_rollout_costs.restype = C.c_int
    return lib


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def emul_costs(lib, c, x0, eps32, prev_idx, u=None):
    ref = np.ascontiguousarray(c.ref_path)
    u = np.ascontiguousarray(c.u_prev if u is None else u)
    K, T = eps32.shape[:2]
    S = np.zeros(K, np.float32)
    arm = np.array([c.arm[k] for k in ("m1", "m2", "l1", "l2", "lc1", "lc2", "g")], dtype=np.float64)
    sinv = np.ascontiguousarray(np.linalg.inv(c.sigma))
    x0 = np.ascontiguousarray(np.asarray(x0, dtype=np.float64))
    p = lib.emul_rollout_costs(dp(ref), ref.shape[0], prev_idx, dp(x0), dp(u), K, T,
                               mo.exploit_count(K, c.param_exploration), C.c_double(c.delta_t),
                               C.c_double(c.param_gamma), dp(sinv), dp(np.ascontiguousarray(c.stage_cost_weight)),
                               dp(np.ascontiguousarray(c.terminal_cost_weight)), dp(arm), C.c_double(c.cost_l1),
                               C.c_double(c.cost_l2), fp(np.ascontiguousarray(eps32)), fp(S))
    return S, p


def test_fp32_costs_within_stated_tolerance(emul, paths):
    for case in cases.single_cases(paths)[:6] + cases.c2_cases()[:1]:
        kw = cases.ctor_kwargs(case, paths)
        c = mo.OracleMPPI(**kw)
        if "u_prev" in case:
            c.u_prev = np.array(case["u_prev"])
        eps = mo.injected_noise(case["seed"], case["K"], case["T"], kw["sigma"])
        S32, p = emul_costs(emul, c, case["x0"], eps, case.get("prev_idx", 0))
        S64 = mo.rollout_costs(c, np.array(case["x0"]), eps.astype(np.float64), prev_idx=p)
        assert np.max(np.abs(S32 - S64)) <= 2e-6 * np.max(S64), case["name"]


def test_fp32_update_within_1e4_teacher_forced(emul, paths):
    with np.load(cases.HERE + "/closed_loop_c1.npz") as z:
        cl = {k: z[k] for k in z.files}
    K, T, seed0, _ = (int(v) for v in cl["meta"])
    kw = cases.run_py_kwargs(cases.ref_path_for(paths, "xydq_circle.txt"), K, T)
    c = mo.OracleMPPI(**kw)
    worst = 0.0
    for s in range(1, cl["state"].shape[0], 7):
        prev = cl["u_new"][s - 1]
        u = np.concatenate([prev[1:], prev[-1:]], axis=0)
        eps = mo.injected_noise(seed0 + s, K, T, kw["sigma"])
        S32, p = emul_costs(emul, c, cl["state"][s], eps, int(cl["prev_idx"][s, 0]), u=u)
        assert p == cl["prev_idx"][s, 1]
        w, _, _ = mo.softmin_weights(S32.astype(np.float64), c.param_lambda)
        un = u + mo.filter_columns(np.einsum("k,ktm->tm", w, eps.astype(np.float64)))
        worst = max(worst, np.max(np.abs(un - cl["u_new"][s])) / np.max(np.abs(cl["u_new"][s])))
    assert worst <= 1e-4, worst


def test_sincos_accuracy(emul):
    x = np.random.default_rng(0).uniform(-40, 40, 200000).astype(np.float32)
    s, c = np.zeros_like(x), np.zeros_like(x)
    emul.emul_sincos(fp(x), x.size, fp(s), fp(c))
    assert np.max(np.abs(s - np.sin(x.astype(np.float64)))) <= 1.0e-7
    assert np.max(np.abs(c - np.cos(x.astype(np.float64)))) <= 1.0e-7


def test_philox4x32_10_known_answers(emul):
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    out = (C.c_uint32 * 4)()
    for ctr, key, exp in kat:
        emul.emul_philox(*[C.c_uint32(v) for v in ctr], *[C.c_uint32(v) for v in key], out)
        assert tuple(out) == exp


def test_host_noise_moments(emul):
    K, T = 4096, 50
    chol = np.linalg.cholesky(np.array([[20.0, 6.0], [6.0, 10.0]]))
    eps = np.zeros((K, T, 2), np.float32)
    emul.emul_noise(1, 2, 3, dp(np.ascontiguousarray(chol)), 0, 0, K, T, fp(eps))
    flat = eps.reshape(-1, 2).astype(np.float64)
    assert np.all(np.abs(flat.mean(0)) < 0.05)
    np.testing.assert_allclose(np.cov(flat.T), [[20.0, 6.0], [6.0, 10.0]], atol=0.25)

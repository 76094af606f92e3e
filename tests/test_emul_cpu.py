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


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def emul_costs(lib, c, x0, eps32, prev_idx, u=None, use_cert=True, hits=None):
    ref = np.ascontiguousarray(c.ref_path)
    u = np.ascontiguousarray(c.u_prev if u is None else u)
    K, T = eps32.shape[:2]
    S = np.zeros(K, np.float32)
    arm = np.array([c.arm[k] for k in ("m1", "m2", "l1", "l2", "lc1", "lc2", "g")], dtype=np.float64)
    sinv = np.ascontiguousarray(np.linalg.inv(c.sigma))
    x0 = np.ascontiguousarray(np.asarray(x0, dtype=np.float64))
    h = (C.c_longlong * 2)(0, 0)
    p = lib.emul_rollout_costs(dp(ref), ref.shape[0], prev_idx, dp(x0), dp(u), K, T,
                               mo.exploit_count(K, c.param_exploration), C.c_double(c.delta_t),
                               C.c_double(c.param_gamma), dp(sinv), dp(np.ascontiguousarray(c.stage_cost_weight)),
                               dp(np.ascontiguousarray(c.terminal_cost_weight)), dp(arm), C.c_double(c.cost_l1),
                               C.c_double(c.cost_l2), fp(np.ascontiguousarray(eps32)), fp(S), int(use_cert), h,
                               1 if getattr(c, "dynamics", "F") == "F1" else 0)
    if hits is not None:
        hits.append((h[0], h[1]))
    return S, p


def test_fp32_costs_within_stated_tolerance(emul, paths):
    single = cases.single_cases(paths)
    for case in single[:6] + [c for c in single if "dynamics" in c] + cases.c2_cases()[:1]:
        kw = cases.ctor_kwargs(case, paths)
        c = mo.OracleMPPI(**kw, dynamics=case.get("dynamics", "F"))
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


def test_fixed_point_sincos_accuracy_and_bound(emul):
    """What the rollouts evaluate: sin / cos of an angle held in units of 2 pi / 2^32 (mppi_math.cuh::sincos_fix).
    The result is bounded by 1 + 2^-23 for every argument — the certified lookups rely on that (their box test on
    the end-effector coordinates is implied by it)."""
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-40, 40, 400000), rng.uniform(-1e6, 1e6, 100000),
                        np.pi / 2 * np.arange(-40, 41) + rng.uniform(-1e-6, 1e-6, 81),
                        np.pi / 2 * np.arange(-40, 41)])
    s, c = np.zeros(x.size, np.float32), np.zeros(x.size, np.float32)
    emul.emul_sincos_fix(dp(x), x.size, fp(s), fp(c))
    # the fixed-point image of x is within 2 pi / 2^33 = 7.3e-10 rad of x
    assert np.max(np.abs(s - np.sin(x))) <= 1.5e-7
    assert np.max(np.abs(c - np.cos(x))) <= 2.5e-7
    assert np.max(np.abs(s)) <= 1.0 + 2.0 ** -23 and np.max(np.abs(c)) <= 1.0 + 2.0 ** -23


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


# ---------------------------------------------------------------------------------------------
# certified nearest-waypoint lookups (mppi_math.cuh: WinCert / RowRec)
# ---------------------------------------------------------------------------------------------
CERT_FLOATS = 16 + 32 * 8 + 16      # 64 certificate bytes + 32 row records of 32 bytes + 64 bytes of end wedges


def _probe(emul, ref, p, xy):
    n = xy.shape[0]
    pick, full, scan = (np.zeros(n, np.int32) for _ in range(3))
    cert = np.zeros(CERT_FLOATS, np.float32)
    xy = np.ascontiguousarray(xy.astype(np.float32))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))          # noqa: E731
    emul.emul_cert_probe(dp(ref), ref.shape[0], int(p), C.c_double(2.0), fp(xy), n, ip(pick), ip(full), ip(scan), fp(cert))
    assert np.array_equal(scan, full)          # the in-memory search is the register tournament
    return pick, full, cert


def cert_boundary_queries(cert, rng, N):
    """Queries hugging every threshold of a certificate: the tangent line of each row (both sides, from 1 nm
    to 1 mm off it) at lateral positions spread over, and right at, the certified lateral range."""
    c = cert.astype(np.float64)
    nx, ny, blo, bhi = c[0:4]
    rec = c[16:16 + 256].reshape(32, 8)
    out = []
    for o in (272, 278):                         # apex and edges of the two far-field wedges
        mx, my, k = (c[o + 2 * i:o + 2 * i + 2] for i in range(3))
        A = np.array([[mx[0], my[0]], [mx[1], my[1]]])
        if np.any(np.abs(k) > 1e30) or abs(np.linalg.det(A)) < 1e-9:
            continue
        z = np.linalg.solve(A, -k)
        M = 4 * N
        t = (10.0 ** rng.uniform(-7, 0.5, M) * rng.choice([-1, 1], M))[:, None]
        which = rng.integers(0, 3, M)[:, None]
        out.append(z[None, :] + rng.standard_normal((M, 2)) * (10.0 ** rng.uniform(-9, -5, M))[:, None]
                   + np.where(which == 0, t * np.array([-my[0], mx[0]]), 0.0)
                   + np.where(which == 1, t * np.array([-my[1], mx[1]]), 0.0))
    if not (blo < bhi):
        return np.concatenate(out) if out else np.zeros((0, 2))
    for a in range(32):
        tx, ty = rec[a, 4:6]
        for k in rec[a, 6:8]:
            if abs(k) > 1e30 or abs(tx * ny - ty * nx) < 1e-6:
                continue
            beta = np.concatenate([rng.uniform(max(blo, -4.0), min(bhi, 4.0), N),
                                   blo + rng.normal(0, 1e-6, N // 4), bhi + rng.normal(0, 1e-6, N // 4)])
            off = 10.0 ** rng.uniform(-9, -3, beta.size) * rng.choice([-1, 1], beta.size)
            A = np.array([[tx, ty], [nx, ny]])
            out.append(np.linalg.solve(A, np.stack([-k + off, beta])).T)
    return np.concatenate(out) if out else np.zeros((0, 2))


def test_certificate_never_disagrees_with_the_full_search(emul, paths):
    """Whenever the certificate names a row, the exact FP32 30-candidate search returns the same row:
    random queries from 10 um to 2 m around windows of all four reference files (noisy recorded paths
    included), windows truncated by the end of the path, and queries placed on every threshold."""
    rng = np.random.default_rng(5)
    certified = 0
    for name in ("xydq_circle", "xydq", "trajectory", "trajectory1"):
        ref = np.ascontiguousarray(paths[name][:, 0:4], dtype=np.float64)
        n = ref.shape[0]
        for p in list(rng.integers(0, n - 31, 12)) + [n - 31, n - 30, n - 12, n - 3, n - 2, n - 1, 0]:
            nv = min(30, n - int(p))
            N = 6000
            base = ref[int(p) + rng.integers(0, nv, N), 0:2] - ref[int(p), 0:2]
            q = base + rng.standard_normal((N, 2)) * (10.0 ** rng.uniform(-5, 0.3, N))[:, None]
            pick, full, cert = _probe(emul, ref, p, q)
            m = pick >= 0
            assert np.array_equal(pick[m], full[m]), (name, int(p))
            certified += int(m.sum())
            q = cert_boundary_queries(cert, rng, 120)
            if q.shape[0]:
                pick, full, _ = _probe(emul, ref, p, q)
                m = pick >= 0
                assert np.array_equal(pick[m], full[m]), (name, int(p), "boundary")
                certified += int(m.sum())
    assert certified > 500000         # the test exercised the certificate, not only its refusals


def test_certified_rollout_costs_are_bit_identical(emul, paths):
    """Rollout costs with certified lookups, with the in-memory search only, and with the register tournament
    of the kernels built without the certificate are the same floats; on a tracking state nearly every
    lookup is certified (window ends ~15 steps ahead of the arm: end rows; before that: triples)."""
    with np.load(cases.HERE + "/closed_loop_c1.npz") as z:
        cl = {k: z[k] for k in z.files}
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    for s, T in ((100, 30), (500, 100), (1000, 64), (1499, 40)):
        K = 256
        kw = cases.run_py_kwargs(ref, K, T)
        c = mo.OracleMPPI(**kw)
        prev = cl["u_new"][s - 1]
        u = np.concatenate([prev[1:], np.repeat(prev[-1:], max(T - 29, 1), axis=0)], axis=0)[:T]
        eps = mo.injected_noise(1000 + s, K, T, kw["sigma"])
        hits = []
        S_on, p_on = emul_costs(emul, c, cl["state"][s], eps, int(cl["prev_idx"][s, 0]), u=u, use_cert=1, hits=hits)
        S_off, p_off = emul_costs(emul, c, cl["state"][s], eps, int(cl["prev_idx"][s, 0]), u=u, use_cert=0, hits=hits)
        S_reg, _ = emul_costs(emul, c, cl["state"][s], eps, int(cl["prev_idx"][s, 0]), u=u, use_cert=2, hits=hits)
        assert p_on == p_off and np.array_equal(S_on, S_off) and np.array_equal(S_on, S_reg)
        # two samples per call, in lockstep like the throughput kernels
        emul.emul_set_ns(2)
        try:
            S_two, _ = emul_costs(emul, c, cl["state"][s], eps, int(cl["prev_idx"][s, 0]), u=u, use_cert=1)
        finally:
            emul.emul_set_ns(1)
        assert np.array_equal(S_on, S_two)
        assert hits[1] == (0, 0)
        assert hits[0][0] + hits[0][1] > 0.995 * K * T, (s, T, hits)
        if T >= 64:
            assert hits[0][0] > 0.6 * K * T, (s, T, hits)
    # a window cut short by the end of the path, and the arm at rest at the start of the path
    for x0, p, T in ((cases.X0, 0, 50), (paths["trajectory1"][1987, 0:2].tolist() + [0.01, 0.01], 1985, 20)):
        kw = cases.run_py_kwargs(ref, 128, T)
        c = mo.OracleMPPI(**kw)
        eps = mo.injected_noise(3, 128, T, kw["sigma"])
        S_on, _ = emul_costs(emul, c, x0, eps, p, use_cert=1)
        S_off, _ = emul_costs(emul, c, x0, eps, p, use_cert=0)
        assert np.array_equal(S_on, S_off)


def _synthetic_window_path(kind, rng, n=40):
    t = np.arange(n)
    if kind == "line":
        xy = np.stack([0.5 + 0.002 * t, 0.3 + 0.001 * t], 1)
    elif kind == "line_noise":
        xy = np.stack([0.5 + 0.002 * t, 0.3 + 0.001 * t], 1) + rng.normal(0, 10.0 ** rng.uniform(-9, -3.3), (n, 2))
    elif kind == "walk":
        xy = np.cumsum(rng.normal(0, 10.0 ** rng.uniform(-4, -2), (n, 2)), 0) + rng.uniform(-1, 1, 2)
    elif kind == "arc":
        r, th = 10.0 ** rng.uniform(-2, 1), rng.uniform(0, 6.28) + t * 10.0 ** rng.uniform(-4, -1.5)
        xy = np.stack([r * np.cos(th), r * np.sin(th)], 1) + rng.uniform(-1, 1, 2)
    elif kind == "dups":                                   # every waypoint three times
        xy = np.stack([0.5 + 0.002 * (t // 3), 0.3 + 0.0 * t], 1)
    elif kind == "neardup":                                # pairs of waypoints 1 nm apart
        xy = np.stack([0.5 + 0.002 * t, 0.3 + 0.0 * t], 1)
        xy[1::7] = xy[0::7][:len(xy[1::7])] + 1e-9
    elif kind == "zigzag":
        xy = np.stack([0.5 + 0.002 * t, 0.3 + 0.002 * (t % 2)], 1)
    elif kind == "uturn":
        th = np.linspace(0, np.pi * rng.uniform(0.5, 1.5), n)
        xy = np.stack([0.02 * np.cos(th), 0.02 * np.sin(th)], 1) + 0.7
    elif kind == "far":                                    # window far from the arm's base: large local coordinates
        xy = np.stack([50 + 0.002 * t, -30 + 0.001 * t], 1)
    elif kind == "tiny":                                   # micrometre spacing
        xy = np.stack([0.5 + 1e-6 * t, 0.3 + 2e-6 * t], 1)
    elif kind == "accel":                                  # spacing growing 30x along the window (a start from rest)
        xy = np.stack([0.5 + 2e-5 * t * t, 0.3 + 1e-5 * t * t], 1)
    else:                                                  # spiral
        th, r = t * 0.3, 0.001 * t + 0.001
        xy = np.stack([r * np.cos(th), r * np.sin(th)], 1) + 0.4
    return np.ascontiguousarray(np.concatenate([xy, np.zeros((n, 2))], 1))


ADVERSARIAL_KINDS = ["line", "line_noise", "walk", "arc", "dups", "neardup", "zigzag", "uturn", "far", "tiny", "accel",
                     "spiral"]
ARMED_KINDS = {"line", "line_noise", "arc", "far", "accel"}          # smooth enough for most rows to get a role


@pytest.mark.parametrize("kind", ADVERSARIAL_KINDS)
def test_certificate_is_sound_on_adversarial_paths(emul, kind):
    """Paths the reference files do not contain: exact and near duplicates (a tie must go to the lower index, so
    those rows have to be refused), zigzags, U-turns and spirals (direction spread), micrometre spacing, noise
    from 1 nm to half a spacing, a window 58 m from the base, strongly non-uniform spacing.  Wherever a
    certificate is issued it agrees with the exact FP32 search, also for queries hugging every threshold."""
    rng = np.random.default_rng(sum(kind.encode()))
    certified = 0
    for _ in range(25):
        ref = _synthetic_window_path(kind, rng)
        n = ref.shape[0]
        p = int(rng.integers(0, n - 1))
        nv = min(30, n - p)
        N = 3000
        base = ref[p + rng.integers(0, nv, N), 0:2] - ref[p, 0:2]
        span = max(np.ptp(ref[p:p + nv, 0]), np.ptp(ref[p:p + nv, 1]), 1e-9)
        qs = [base + rng.standard_normal((N, 2)) * (span * 10.0 ** rng.uniform(-4, 2, N))[:, None]]
        _, _, cert = _probe(emul, ref, p, qs[0])
        qs.append(cert_boundary_queries(cert, rng, 60))
        pick, full, _ = _probe(emul, ref, p, np.concatenate(qs))
        m = pick >= 0
        assert np.array_equal(pick[m], full[m]), kind
        certified += int(m.sum())
    if kind in ARMED_KINDS:
        assert certified > 10000, (kind, certified)


def test_cost_sum_compensation_is_not_what_holds_the_tolerance(emul, paths, tmp_path):
    """Numerics study kept as a test: without the Kahan term of the cost accumulator
    (MPPI_KAHAN_MASK=3, the default) the updated sequence stays within the same bound over the reference's
    closed loop; without the compensation of the joint rates (mask 2) it does not keep the margin.  (The
    angles are integrated in fixed point by the rollouts and need no compensation term.)"""
    worst = {}
    with np.load(cases.HERE + "/closed_loop_c1.npz") as z:
        cl = {k: z[k] for k in z.files}
    K, T, seed0, _ = (int(v) for v in cl["meta"])
    kw = cases.run_py_kwargs(cases.ref_path_for(paths, "xydq_circle.txt"), K, T)
    c = mo.OracleMPPI(**kw)
    for mask in (3, 2):
        so = str(tmp_path / f"emul_mask{mask}.so")
        subprocess.run(["g++", "-O2", "-ffp-contract=off", f"-DMPPI_KAHAN_MASK={mask}", "-shared", "-fPIC", "-o", so, SRC],
                       check=True)
        lib = C.CDLL(so)
        lib.emul_rollout_costs.restype = C.c_int
        w_err = 0.0
        for s in range(1, cl["state"].shape[0], 6):
            prev = cl["u_new"][s - 1]
            u = np.concatenate([prev[1:], prev[-1:]], axis=0)
            eps = mo.injected_noise(seed0 + s, K, T, kw["sigma"])
            S32, _ = emul_costs(lib, c, cl["state"][s], eps, int(cl["prev_idx"][s, 0]), u=u)
            w, _, _ = mo.softmin_weights(S32.astype(np.float64), c.param_lambda)
            un = u + mo.filter_columns(np.einsum("k,ktm->tm", w, eps.astype(np.float64)))
            w_err = max(w_err, np.max(np.abs(un - cl["u_new"][s])) / np.max(np.abs(cl["u_new"][s])))
        worst[mask] = w_err
    assert worst[3] <= 5e-5, worst
    assert worst[2] > worst[3], worst


@pytest.mark.parametrize("dynamics", ["F", "F1"])
def test_latency_form_of_the_optimal_trajectory_rollout(emul, paths, dynamics):
    """The final stage rolls the optimal trajectory out with ONE thread (control.py:129-134) in the latency form of
    the step (mppi_math.cuh::arm_step_serial: sin / cos advanced by rotating through the step's increment).  Against
    the FP64 oracle on the reference's own closed-loop states and sequences, and on fast motions (increments
    beyond 0.25 rad take the exact fallback): within the stated 2e-5 of the trajectory tolerance."""
    with np.load(cases.HERE + "/closed_loop_c1.npz") as z:
        cl = {k: z[k] for k in z.files}
    arm_d = mo.default_arm_params()
    arm = np.array([arm_d[k] for k in ("m1", "m2", "l1", "l2", "lc1", "lc2", "g")], dtype=np.float64)
    rng = np.random.default_rng(5)
    worst = 0.0
    todo = [(cl["state"][s], cl["u_new"][s], 0.006) for s in (1, 200, 700, 1400)]
    todo += [(np.array([0.3, -0.5, 1.0, -2.0]), rng.normal(0, 20, (100, 2)), 0.006),
             (np.array([2.0, 1.0, 25.0, -30.0]), rng.normal(0, 40, (64, 2)), 0.01),      # increments up to ~0.5 rad
             (np.array([-3.0, 0.2, -45.0, 20.0]), rng.normal(0, 5, (50, 2)), 0.006)]
    for x0, u, dt in todo:
        T = u.shape[0]
        s = tuple(float(a) for a in x0)
        ref = np.zeros((T, 4))
        for t in range(T):
            s = mo.arm_step(*s, u[t - 1, 0], u[t - 1, 1], arm_d, dt, dynamics)
            ref[t] = s
        out = np.zeros((T, 4))
        emul.emul_optimal_traj(dp(np.ascontiguousarray(x0, dtype=np.float64)), dp(np.ascontiguousarray(u, dtype=np.float64)),
                               T, C.c_double(dt), dp(arm), C.c_double(1.0), C.c_double(1.0), 1 if dynamics == "F1" else 0, dp(out))
        scale = np.maximum(1.0, np.abs(ref))
        worst = max(worst, float(np.max(np.abs(out - ref) / scale)))
    assert worst <= 2e-5, worst

// mppi_math.cuh — FP32 arithmetic of one MPPI sample-step, shared by every kernel.
//
// Everything here is `__host__ __device__` with explicitly rounded operations (no compiler FMA
// contraction decisions): the CUDA kernels and the CPU emulation used by the parity *tests*
// (tests/emul) execute the same operation sequence.  The only device-specific pieces are the MUFU
// approximations (reciprocal, and log2/sqrt/sin/cos inside the Gaussian generator).
//
// What is restated from the reference (file:line in /root/reference):
//   arm dynamics + semi-implicit Euler ........ control.py:234-263 (twin: utils.py:14-29)
//   forward kinematics on the cost side ....... control.py:178-179, 190-191, 206-207
//   nearest waypoint, first arg-min of 30 ..... control.py:200-215
//   stage / terminal tracking cost ............ control.py:174-198
//   control cost gamma*u^T Sigma^-1 v ......... control.py:106
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define MPPI_HD __host__ __device__ __forceinline__
#else
#define MPPI_HD inline
#endif

namespace mppi {

constexpr int kWindow = 30;       // SEARCH_IDX_LEN, control.py:203
constexpr int kWindowPad = 32;    // table rows (two never-selected sentinels)
constexpr int kFilter = 10;       // control.py:122
constexpr float kSentinel = 3.0e38f;
#ifndef MPPI_UNROLL_T
#define MPPI_UNROLL_T 2
#endif
constexpr int kUnrollT = MPPI_UNROLL_T;     // unroll factor of the horizon loop of the rollouts

// Which accumulators of a rollout carry a Kahan compensation term: bit 0 joint rates, bit 1 joint angles,
// bit 2 the cost sum S.  Default: rates and angles.  (Study kept as tests/test_emul_cpu.py::
// test_cost_sum_compensation_is_not_what_holds_the_tolerance: compensating S changes neither the update nor S
// beyond 1e-7 max S; the angle term is the one that holds the 1e-4 bound, the rate term second.)
#ifndef MPPI_KAHAN_MASK
#ifdef MPPI_NO_KAHAN
#define MPPI_KAHAN_MASK 0
#else
#define MPPI_KAHAN_MASK 3
#endif
#endif

// ---- explicitly rounded primitives ---------------------------------------------------------
#if defined(__CUDA_ARCH__)
MPPI_HD float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
MPPI_HD float mul_(float a, float b) { return __fmul_rn(a, b); }
MPPI_HD float add_(float a, float b) { return __fadd_rn(a, b); }
MPPI_HD float sub_(float a, float b) { return __fsub_rn(a, b); }
MPPI_HD float rcp_(float a) {             // MUFU.RCP + one Newton step (<1 ulp)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    float e = __fmaf_rn(-a, r, 1.0f);
    return __fmaf_rn(r, e, r);
}
MPPI_HD int f2i(float a) { return __float_as_int(a); }
MPPI_HD float i2f(int a) { return __int_as_float(a); }
MPPI_HD int32_t f2i_rn_(float a) { return __float2int_rn(a); }     // F2I: nearest-even, saturating, NaN -> 0
#else
MPPI_HD float fma_(float a, float b, float c) { return fmaf(a, b, c); }
MPPI_HD float mul_(float a, float b) { return a * b; }
MPPI_HD float add_(float a, float b) { return a + b; }
MPPI_HD float sub_(float a, float b) { return a - b; }
MPPI_HD float rcp_(float a) { return 1.0f / a; }
MPPI_HD int f2i(float a) { union { float f; int i; } u; u.f = a; return u.i; }
MPPI_HD float i2f(int a) { union { float f; int i; } u; u.i = a; return u.f; }
MPPI_HD int32_t f2i_rn_(float a) {                                  // what cvt.rni.s32.f32 returns
    if (!(a == a)) return 0;
    if (a >= 2147483648.0f) return 2147483647;
    if (a <= -2147483648.0f) return (int32_t)(-2147483647 - 1);
    return (int32_t)nearbyintf(a);                                  // (default rounding mode: nearest-even)
}
#endif

// FP64 reciprocal / reciprocal square root.  Device: FP32 MUFU seed + two Newton steps (~1e-14 relative) —
// the certificate construction is a chain of divisions on the latency path of every control step and
// its results only feed quantities that carry >= 1e-9 of deliberate slack.
#if defined(__CUDA_ARCH__)
MPPI_HD double rcp64_(double x) {
    double r = (double)__frcp_rn((float)x);
    r = fma(r, fma(-x, r, 1.0), r);
    return fma(r, fma(-x, r, 1.0), r);
}
MPPI_HD double rsqrt64_(double x) {
    double r = (double)rsqrtf((float)x);
    r = r * fma(-0.5 * x * r, r, 1.5);
    return r * fma(-0.5 * x * r, r, 1.5);
}
// FP32-accurate reciprocal (relative error < 2e-7) for the certificate's pair loop, whose results are widened by 1e-6
MPPI_HD double rcp32_(double x) { return (double)__frcp_rn((float)x); }
#else
MPPI_HD double rcp64_(double x) { return 1.0 / x; }
MPPI_HD double rsqrt64_(double x) { return 1.0 / sqrt(x); }
MPPI_HD double rcp32_(double x) { return 1.0 / x; }
#endif

// ---- sin & cos of one angle, ~1 ulp, no slow path --------------------------------------------
// Cody-Waite reduction by pi/2 in two FMA steps (the third term of pi/2, 5.4e-15 per quadrant, is below
// 1e-11 for |x| < 3000 rad: a thousandth of an ulp of the result), then degree-7 / degree-8 minimax
// polynomials on [-pi/4, pi/4].  A diverged rollout (huge angle, Inf, NaN) yields a garbage-but-finite or
// NaN cost that the soft-min kernel maps to weight 0.
MPPI_HD void sincos_(float x, float& s, float& c) {
    const float kMagic = 12582912.0f;                      // 1.5 * 2^23: round-to-nearest trick
    float kf = fma_(x, 0.636619772367581343f, kMagic);
    int q = f2i(kf);                                       // low bits hold the quadrant
    kf = sub_(kf, kMagic);
    float r = fma_(kf, -1.57079601287841796875f, x);
    r = fma_(kf, -3.1391647326017846e-07f, r);
#ifdef MPPI_CW3
    r = fma_(kf, -5.3903025299577648e-15f, r);
#endif
    float r2 = mul_(r, r);
    // sin(r) = r + r*r2*(S1 + r2*(S2 + r2*S3))
    float ps = fma_(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fma_(ps, r2, -1.6666654611e-1f);
    float sr = fma_(mul_(ps, r2), r, r);
    // cos(r) = 1 + r2*(C0 + r2*(C1 + r2*(C2 + r2*C3)))
    float pc = fma_(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fma_(pc, r2, 4.166664568298827e-2f);
    pc = fma_(pc, r2, -0.5f);
    float cr = fma_(pc, r2, 1.0f);
    const bool odd = (q & 1) != 0;
    float ss = odd ? cr : sr;
    float cc = odd ? sr : cr;
    s = (q & 2) != 0 ? -ss : ss;
    c = ((q + 1) & 2) != 0 ? -cc : cc;
}

// ---- joint angles in fixed point ----------------------------------------------------------------
// The dynamics only ever use sin/cos of q1 and q1 + q2, so the rollouts carry those two angles as 32-bit
// integers in units of 2 pi / 2^32 (one turn = 2^32: the integer wraps where the angle does).  The
// integration q <- q + dq dt (control.py:258-259) becomes one F2I of dq * dt in those units and an
// integer add: resolution 1.46e-9 rad per step, ~40x finer than an FP32 q near 1 rad — what the Kahan
// term of a float angle buys — and the sum q1 + q2 is exact.  The range reduction of the sincos is a
// shift and a conversion, the sign of the half turn one XOR (no quadrant selects).
#ifndef MPPI_ANGLE_FIX
#define MPPI_ANGLE_FIX 1
#endif
constexpr double kFixPerRad = 4294967296.0 / 6.283185307179586476925;     // 2^32 / (2 pi)
MPPI_HD uint32_t angle_fix(double q) {                  // FP64 angle -> fixed point (prepare kernel, host)
    if (!(fabs(q) < 1.0e15)) return 0u;
    double t = q * kFixPerRad;
    t -= 4294967296.0 * floor(t * (1.0 / 4294967296.0));                   // [0, 2^32]
    return (uint32_t)(unsigned long long)rint(t);                          // (2^32 wraps to 0)
}
// sin & cos of a fixed-point angle.  k = round(a / 2^31) half turns are removed by the shift (the remainder
// 2 (a - k 2^31) IS the signed value of a << 1), r = remainder * pi / 2^32 lies in [-pi/2, pi/2];
// sin(a) = (-1)^k sin r, cos(a) = (-1)^k cos r with k's parity = bit 31 of a + 2^30.  Minimax polynomials
// of degree 9 / 10 on [-pi/2, pi/2] (4.6e-9, 2.4e-10); measured max abs error of the FP32 evaluation
// over all arguments: 1.4e-7 (sin), 2.3e-7 (cos); |s|, |c| <= 1 + 2^-23 for EVERY argument (there is
// no argument for which the result is not finite).
MPPI_HD float flip_sign_(float v, uint32_t m) {       // v with its sign flipped where bit 31 of m is set: ONE LOP3
#if defined(__CUDA_ARCH__)
    uint32_t o;
    asm("lop3.b32 %0, %1, %2, %3, 0x6A;" : "=r"(o) : "r"(m), "r"(0x80000000u), "r"(__float_as_uint(v)));   // (m & mask) ^ v
    return __uint_as_float(o);
#else
    return i2f((int)((uint32_t)f2i(v) ^ (m & 0x80000000u)));
#endif
}
MPPI_HD void sincos_fix(uint32_t a, float& s, float& c) {
    const int32_t ri = (int32_t)(a << 1);
    const float r = mul_((float)ri, 7.3145906e-10f);                   // pi / 2^32
    const uint32_t flip_src = a + 0x40000000u;                         // bit 31: odd number of half turns
    const float r2 = mul_(r, r);
    float ps = fma_(r2, 2.60005481e-06f, -1.98066147e-04f);
    ps = fma_(ps, r2, 8.33301712e-03f);
    ps = fma_(ps, r2, -1.66666567e-01f);
    const float sr = fma_(mul_(ps, r2), r, r);
    float pc = fma_(r2, -2.6077106e-07f, 2.47618864e-05f);
    pc = fma_(pc, r2, -1.38884038e-03f);
    pc = fma_(pc, r2, 4.16666418e-02f);
    pc = fma_(pc, r2, -0.5f);
    const float cr = fma_(pc, r2, 1.0f);
    s = flip_sign_(sr, flip_src);
    c = flip_sign_(cr, flip_src);
}

// ---- per-controller constants (derived once on the host in FP64, rounded to FP32) -----------
struct ArmF {
    float A0, A1;        // M11 = A0 + A1*cos q2         (control.py:241-242)
    float M22, B1;       // M12 = M22 + B1*cos q2; h = B1*sin q2   (control.py:243-244, 247)
    float G1a, G1b;      // g1 = G1a*cos q1 + G1b*cos q12; g2 = G1b*cos q12   (control.py:248-249)
    float dt;            // controller integration step (control.py:240)
    float L1, L2;        // cost-side link lengths self.l1 / self.l2 (control.py:55-56)
    float dtfix;         // dt in fixed-point angle units per rad/s: dt * 2^32 / (2 pi)
};
MPPI_HD float arm_dtfix(double dt) { return (float)(dt * kFixPerRad); }

struct CostW {           // weights already multiplied by 1e4 (control.py:185, 198)
    float s0, s1, s2, s3;
    float t0, t1, t2, t3;
    // joint-limit stage cost (north_star item 1; not in the reference — its only limits are the commented-out
    // clamps of _g, control.py:166-172): jw * (viol(q1)^2 + viol(q2)^2), viol(q) = max(q - hi, lo - q, 0)
    float jw, lo1, hi1, lo2, hi2;
    // square roots of s0..s3: the stage cost is evaluated on residuals that are already scaled (stage_cost below)
    float r0, r1, r2, r3;
};
MPPI_HD void cost_roots(CostW& W) {
    W.r0 = (float)sqrt((double)W.s0); W.r1 = (float)sqrt((double)W.s1);
    W.r2 = (float)sqrt((double)W.s2); W.r3 = (float)sqrt((double)W.s3);
}
MPPI_HD float joint_limit_cost(const CostW& W, float q1, float q2) {
    const float v1 = fmaxf(fmaxf(sub_(q1, W.hi1), sub_(W.lo1, q1)), 0.0f);
    const float v2 = fmaxf(fmaxf(sub_(q2, W.hi2), sub_(W.lo2, q2)), 0.0f);
    return mul_(W.jw, fma_(v1, v1, mul_(v2, v2)));
}

struct WinEntry { float a, b, c, pad; };     // d_j - |p'|^2 = c + a*x' + b*y'   (local coordinates)
struct RefRow { float rx, ry, rd1, rd2; };   // waypoint in local coordinates + reference joint rates
struct StepCtl { float u1, u2, g1, g2; };    // nominal control and gamma*(u^T Sigma^-1)

// acc += y with the rounding error left in `comp` (y already has the old comp subtracted)
MPPI_HD void kahan_(float& acc, float& comp, float y) {
    float t = add_(acc, y);
    comp = sub_(sub_(t, acc), y);
    acc = t;
}

// Arm state carried through the horizon.  sin/cos of q1 and q1+q2 are kept from the previous step
// (they were needed for its forward kinematics) so each step evaluates two sincos, not eight cos/sin.
struct ArmState {
    float q1, q2, d1, d2;
    float s1, c1, s12, c12;
    float kq1, kq2, kd1, kd2;      // Kahan compensation terms of the four integrators
    uint32_t a1, a12;              // q1 and q1 + q2 in fixed point (what the sincos are taken of)
};

// (a1, a12) = angle_fix of the FP64 q1 and q1 + q2 (the step header carries them)
MPPI_HD void arm_init(ArmState& st, float q1, float q2, float d1, float d2, uint32_t a1, uint32_t a12) {
    st.q1 = q1; st.q2 = q2; st.d1 = d1; st.d2 = d2;
    st.kq1 = 0.f; st.kq2 = 0.f; st.kd1 = 0.f; st.kd2 = 0.f;
    st.a1 = a1; st.a12 = a12;
#if MPPI_ANGLE_FIX
    sincos_fix(a1, st.s1, st.c1);
    sincos_fix(a12, st.s12, st.c12);
#else
    sincos_(q1, st.s1, st.c1);
    sincos_(add_(q1, q2), st.s12, st.c12);
#endif
}

// q <- q + dq dt (control.py:258-259, with the NEW rates) and the sin/cos of the new angles.  TRACKQ: the
// float angles are kept as well (compensated), for the kernels that output them or charge a joint-limit cost;
// the rollouts proper only need the fixed-point pair.
template <bool TRACKQ>
MPPI_HD void arm_advance_angles(ArmState& st, const ArmF& A) {
#if MPPI_ANGLE_FIX
    const int32_t i1 = f2i_rn_(mul_(st.d1, A.dtfix));
    const int32_t i2 = f2i_rn_(mul_(st.d2, A.dtfix));
    st.a1 += (uint32_t)i1;
    st.a12 += (uint32_t)i1 + (uint32_t)i2;
    if (TRACKQ) {
#endif
#if (MPPI_KAHAN_MASK & 2)
        kahan_(st.q1, st.kq1, fma_(st.d1, A.dt, -st.kq1));
        kahan_(st.q2, st.kq2, fma_(st.d2, A.dt, -st.kq2));
#else
        st.q1 = fma_(st.d1, A.dt, st.q1);
        st.q2 = fma_(st.d2, A.dt, st.q2);
#endif
#if MPPI_ANGLE_FIX
    }
    sincos_fix(st.a1, st.s1, st.c1);
    sincos_fix(st.a12, st.s12, st.c12);
#else
    sincos_(st.q1, st.s1, st.c1);
    sincos_(add_(st.q1, st.q2), st.s12, st.c12);
#endif
}

// One integration step (control.py:241-259) under control (v1, v2).
// DYN = 0: the arm model _F.  DYN = 1: the reference's other rollout model _F1 (control.py:265-295),
// which forms u = M v + C dq (gravity dropped, control.py:281-284) and solves ddq = M^-1 (u - C dq):
// ddq = v up to FP64 rounding, so the input is applied as the joint acceleration.
// the joint rates of the step: d <- d + ddq dt
template <int DYN>
MPPI_HD void arm_rates(ArmState& st, const ArmF& A, float v1, float v2) {
    if (DYN == 1) {
#if (MPPI_KAHAN_MASK & 1)
        kahan_(st.d1, st.kd1, fma_(v1, A.dt, -st.kd1));
        kahan_(st.d2, st.kd2, fma_(v2, A.dt, -st.kd2));
#else
        st.d1 = fma_(v1, A.dt, st.d1); st.d2 = fma_(v2, A.dt, st.d2);
#endif
        return;
    }
    // cos/sin of q2 = (q1+q2) - q1 by the angle-difference identity
    float c2 = fma_(st.c12, st.c1, mul_(st.s12, st.s1));
    float s2 = fma_(st.s12, st.c1, -mul_(st.c12, st.s1));
    float M11 = fma_(A.A1, c2, A.A0);
    float M12 = fma_(A.B1, c2, A.M22);
    float h = mul_(A.B1, s2);
    // v - G: g2 = G1b cos q12, g1 = G1a cos q1 + g2 (control.py:248-249), subtracted by FMAs
    float w2 = fma_(-A.G1b, st.c12, v2);
    float w1 = fma_(-A.G1a, st.c1, fma_(-A.G1b, st.c12, v1));
    // v - C dq - G with C dq = [-h d2 (2 d1 + d2), h d1^2]
    float tt = fma_(2.0f, st.d1, st.d2);
    float b1 = fma_(mul_(h, st.d2), tt, w1);
    float b2 = fma_(-mul_(h, st.d1), st.d1, w2);
    float det = fma_(M11, A.M22, -mul_(M12, M12));
    float idt = mul_(rcp_(det), A.dt);
    float n1 = fma_(A.M22, b1, -mul_(M12, b2));
    float n2 = fma_(M11, b2, -mul_(M12, b1));
    // compensated (Kahan) integration: the rounding error of each accumulator is carried, so the
    // state error stays ~1 ulp instead of growing like sqrt(T) ulp over the horizon
#if (MPPI_KAHAN_MASK & 1)
    kahan_(st.d1, st.kd1, fma_(n1, idt, -st.kd1));
    kahan_(st.d2, st.kd2, fma_(n2, idt, -st.kd2));
#else
    st.d1 = fma_(n1, idt, st.d1);
    st.d2 = fma_(n2, idt, st.d2);
#endif
    // (the rate's own compensation term times dt, ~1e-8 * dt, is far below one ulp of q and is dropped)
}

template <int DYN = 0, bool TRACKQ = true>
MPPI_HD void arm_step(ArmState& st, const ArmF& A, float v1, float v2) {
    arm_rates<DYN>(st, A, v1, v2);
    arm_advance_angles<TRACKQ>(st, A);
}

// The same step for ONE trajectory computed by ONE thread (the optimal-trajectory rollout of the final stage,
// control.py:129-134), where what counts is the length of the dependent chain, not the instruction count: sin / cos
// of the new angles follow from the old ones by rotating through the step's increment e = dq dt (|e| <= 0.25 rad:
// Taylor terms to e^5 / e^6, truncation < 2e-8) instead of F2I -> add -> I2FP -> polynomial.  The angles themselves
// are still integrated (compensated floats: for the output, and for the exact fallback of a larger step).
// Each rotation adds one rounding (~6e-8) to the unit vector: ~6e-7 after 100 steps, against a stated tolerance
// of 2e-5 on the trajectory.
MPPI_HD void rotate_small_(float& s, float& c, float e) {
    const float e2 = mul_(e, e);
    const float se = mul_(e, fma_(e2, fma_(e2, 8.3333333e-3f, -1.6666667e-1f), 1.0f));
    const float ce = fma_(e2, fma_(e2, fma_(e2, -1.3888889e-3f, 4.1666668e-2f), -0.5f), 1.0f);
    const float sn = fma_(s, ce, mul_(c, se));
    c = fma_(c, ce, -mul_(s, se));
    s = sn;
}
template <int DYN = 0>
MPPI_HD void arm_step_serial(ArmState& st, const ArmF& A, float v1, float v2) {
    arm_rates<DYN>(st, A, v1, v2);
    const float e1 = mul_(st.d1, A.dt), e12 = mul_(add_(st.d1, st.d2), A.dt);
    kahan_(st.q1, st.kq1, fma_(st.d1, A.dt, -st.kq1));
    kahan_(st.q2, st.kq2, fma_(st.d2, A.dt, -st.kq2));
    rotate_small_(st.s1, st.c1, e1);
    rotate_small_(st.s12, st.c12, e12);
    if (!(fmaxf(fabsf(e1), fabsf(e12)) <= 0.25f)) {        // rare (also NaN): the exact values
        // of the compensated float angles (value = q - kq, held to ~1e-8 relative), reduced in FP64: the serial loop does
        // not carry the fixed-point pair (7 instructions per step on a chain that is all latency)
        const double q1 = (double)st.q1 - (double)st.kq1, q2 = (double)st.q2 - (double)st.kq2;
        st.a1 = angle_fix(q1); st.a12 = angle_fix(q1 + q2);
        sincos_fix(st.a1, st.s1, st.c1);
        sincos_fix(st.a12, st.s12, st.c12);
    }
}

// End-effector in window-local coordinates: (x - ox, y - oy), origin = first row of the window.
MPPI_HD void fk_local(const ArmState& st, const ArmF& A, float ox, float oy, float& xl, float& yl) {
    xl = fma_(A.L2, st.c12, fma_(A.L1, st.c1, -ox));
    yl = fma_(A.L2, st.s12, fma_(A.L1, st.s1, -oy));
}

// Weighted squared residuals of (x, y, dq1, dq2) against one waypoint row (control.py:183-185).
MPPI_HD void residuals(const ArmState& st, float xl, float yl, const RefRow& r,
                       float& ex, float& ey, float& e1, float& e2) {
    ex = sub_(xl, r.rx); ey = sub_(yl, r.ry); e1 = sub_(st.d1, r.rd1); e2 = sub_(st.d2, r.rd2);
}
MPPI_HD float wsq(float w0, float w1, float w2, float w3, float ex, float ey, float e1, float e2) {
    float c = mul_(mul_(w0, ex), ex);
    c = fma_(mul_(w1, ey), ey, c);
    c = fma_(mul_(w2, e1), e1, c);
    c = fma_(mul_(w3, e2), e2, c);
    return c;
}

// Stage cost (control.py:183-185) on pre-scaled rows: with r_i = sqrt(weight_i) the table holds
// -r_i * (waypoint component) (formed in FP64, rounded once), so each weighted residual is ONE FMA
// r_i * state + row and the cost four squares: 8 instructions instead of 12.  The terminal cost
// (other weights, once per sample) still goes through residuals() / wsq() on the plain rows.
#ifndef MPPI_STAGE_FOLD
#define MPPI_STAGE_FOLD 1
#endif
MPPI_HD RefRow stage_row(const CostW& W, double rx, double ry, double rd1, double rd2) {
    RefRow r;
    r.rx = (float)(-(double)W.r0 * rx); r.ry = (float)(-(double)W.r1 * ry);
    r.rd1 = (float)(-(double)W.r2 * rd1); r.rd2 = (float)(-(double)W.r3 * rd2);
    return r;
}
MPPI_HD float stage_cost(const CostW& W, const ArmState& st, float xl, float yl, const RefRow& sr) {
    const float ex = fma_(W.r0, xl, sr.rx), ey = fma_(W.r1, yl, sr.ry);
    const float e1 = fma_(W.r2, st.d1, sr.rd1), e2 = fma_(W.r3, st.d2, sr.rd2);
    return fma_(e2, e2, fma_(e1, e1, fma_(ey, ey, mul_(ex, ex))));
}

// ---- Philox4x32-10 counter-based generator (Salmon et al., SC'11; same constants as cuRAND) ----
struct U4 { uint32_t x, y, z, w; };

MPPI_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b; hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * b; lo = (uint32_t)p; hi = (uint32_t)(p >> 32);
#endif
}

// The ten round keys (k0 + i*0x9E3779B9, k1 + i*0xBB67AE85) depend only on the seed: they are
// expanded once on the host and read as constant-bank operands instead of being re-derived by
// every thread for every call.
struct PhiloxKeys { uint32_t k0[10], k1[10]; };

MPPI_HD PhiloxKeys philox_expand_key(uint32_t k0, uint32_t k1) {
    PhiloxKeys k;
    for (int i = 0; i < 10; ++i) { k.k0[i] = k0; k.k1[i] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    return k;
}

MPPI_HD U4 philox4x32_10(U4 ctr, const PhiloxKeys& key) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 10; ++i) {
        uint32_t h0, l0, h1, l1;
        mulhilo(0xD2511F53u, ctr.x, h0, l0);
        mulhilo(0xCD9E8D57u, ctr.z, h1, l1);
        U4 n = { h1 ^ ctr.y ^ key.k0[i], l1, h0 ^ ctr.w ^ key.k1[i], l0 };
        ctr = n;
    }
    return ctr;
}

// Counter layout: x = horizon pair index (t/2), y = global sample index, z = control-step counter,
// w = environment index.  Keyed on the GLOBAL sample index so results do not depend on how samples
// are sharded over GPUs.  One call yields the noise of two consecutive horizon steps.
struct NoiseCfg { PhiloxKeys key; uint32_t step; float L11, L21, L22; };

MPPI_HD float u01_(uint32_t x) {          // (0, 1]
    return fma_((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}

// Box-Muller: two uniforms -> two independent N(0,1)
MPPI_HD float u02pi_(uint32_t x) {       // 2 pi * (0, 1], one FMA
    return fma_((float)x, 1.4629180792671596e-09f, 7.3145903963357980e-10f);
}
MPPI_HD void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
    float u = u01_(a);
#if defined(__CUDA_ARCH__)
    float r;                                                          // sqrt(-2 ln u), MUFU.SQRT (no slow path)
    float lg;                                                         // u >= 2^-33: never denormal, so .ftz changes nothing
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(mul_(-1.3862943611198906f, lg)));
    float sn, cs;
    __sincosf(u02pi_(b), &sn, &cs);
#else
    float r = sqrtf(mul_(-1.3862943611198906f, log2f(u)));
    float th = u02pi_(b);
    float sn = sinf(th), cs = cosf(th);
#endif
    z0 = mul_(r, cs); z1 = mul_(r, sn);
}

// eps for horizon steps 2*pair and 2*pair+1 of global sample k (each a 2-vector ~ N(0, L L^T))
MPPI_HD void noise_pair(const NoiseCfg& nc, uint32_t env, uint32_t k, uint32_t pair,
                        float& e0a, float& e0b, float& e1a, float& e1b) {
    U4 ctr = { pair, k, nc.step, env };
    U4 r = philox4x32_10(ctr, nc.key);
    float z0, z1, z2, z3;
    box_muller(r.x, r.y, z0, z1);
    box_muller(r.z, r.w, z2, z3);
    e0a = mul_(nc.L11, z0); e0b = fma_(nc.L21, z0, mul_(nc.L22, z1));
    e1a = mul_(nc.L11, z2); e1b = fma_(nc.L21, z2, mul_(nc.L22, z3));
}

}  // namespace mppi

// =================================================================================================
// One sample's rollout: T integration steps with stage costs, then the terminal cost
// (control.py:91-109).  `Noise` provides eps for step t; tables may live in registers / shared
// memory (device) or plain arrays (tests/emul on the host).
// =================================================================================================
namespace mppi {

struct StepHeader {            // first 64 bytes of a step block (one per environment and control step)
    float q1, q2, d1, d2;      // observed state rounded to FP32
    float ox, oy;              // window origin = first row of the window (FP32 of the FP64 row)
    int32_t win_start;         // updated prev_waypoints_idx (control.py:230)
    int32_t n_valid;           // rows of the window that exist (control.py:208-209 truncation)
    int32_t status;            // bit0: reached the end of the path (control.py:76)
    uint32_t a1, a12;          // angle_fix of the FP64 q1 and q1 + q2
    int32_t pad[5];
};
static_assert(sizeof(StepHeader) == 64, "header is 64 bytes");

// ---- certified nearest-waypoint lookup ------------------------------------------------------------
// The 30-candidate search (control.py:208-215) costs more than the arm dynamics.  The window is fixed
// for the whole horizon while the rollouts move along a smooth path, so almost every lookup can be
// answered from a handful of half-plane tests whose validity the prepare kernel PROVES in FP64 for the
// window at hand.  Results are bit-identical with and without the shortcut (tests: emulation + GPU).
//
// Per window row a the prepare kernel stores a direction tau_a (the local path tangent, rounded to
// FP32) and the offset k_a = -tau_a.r_a, and for the whole window a lateral coordinate b = nu.p (nu =
// normal of the window chord, rounded to FP32) with a range [blo, bhi].  It guarantees:
//   L(a): tau_a.p + k_a >= 0 ("p is ahead of row a"), blo <= b <= bhi, |x'|,|y'| <= dom
//         ==> the FP32 search value D_a is strictly below D_j for every j < a;
//   U(a): tau_a.p + k_a <= 0 ("p is behind row a"), same box  ==>  D_a < D_j for every j > a.
// Derivation: with g = r_a - r_j, d_j - d_a = |g|^2 + 2 g.(p - r_a) is linear in p.  The region of
// L(a) is the half-strip  V0 + beta v1 + alpha d'  (alpha >= 0, beta in [blo, bhi]) where (tau; nu) V0 =
// (-k_a; 0), (tau; nu) v1 = (0; 1), (tau; nu) d' = (1; 0).  A linear function is bounded below on it iff
// g.d' >= 0 and then attains its minimum at beta = blo or bhi, alpha = 0: per pair (a, j) that is one
// sign test plus one bound on blo or bhi, and [blo, bhi] is the intersection over all pairs.  `margin`
// covers twice the rounding of both FP32 distances, every threshold is moved by the rounding of its own
// FP32 evaluation (cert_delta), and roles whose bounds would not even admit 1 % of the reach around
// the path are switched off instead (k = -/+ huge: the test can never pass).
// The rollouts use it twice (nearest_wp below): first for the two window ends — L(last) or U(0) for a
// whole warp: the answer is that row without loading anything; else each lane guesses j0 from a
// fitted estimate (centre of curvature + quadratic), and L(j0-1) and U(j0+1) certify that the arg-min is
// one of j0-1, j0, j0+1, which are then compared exactly like the full search would.  The estimate needs
// no proof: a wrong guess fails the tests and the warp runs the exact search.
struct RowRec {                // 32 bytes per window row: search coefficients + certificate of the row
    float a, b, c, pad;        // d_j - |p'|^2 = c + a x' + b y'
    float tx, ty, kL, kU;      // L: tx x' + ty y' + kL >= 0,  U: tx x' + ty y' + kU <= 0
};
struct alignas(16) WinCert {   // 64 bytes of a step block
    float nx, ny;              // lateral coordinate b = nx x' + ny y'
    float blo, bhi;            // certified lateral range (blo > bhi: nothing is certified)
    float sx, sy, s0;          // s = sx x' + sy y' + s0: along-chord coordinate relative to the fitted centre
    float bc;                  // lateral coordinate of the fitted centre: w = s / (bc - b)
    float c0, c1, c2;          // index estimate c0 + w (c1 + w c2)
    float jhi;                 // j0 is clamped to [1, jhi]; jhi < 1 (fewer than 3 rows): no triples
    float dom;                 // the margins hold for |x'|, |y'| <= dom
    int32_t last;              // index of the last valid row (n_valid - 1)
    int32_t pad[2];
};
static_assert(sizeof(WinCert) == 64 && sizeof(RowRec) == 32, "certificate layout");

constexpr float kCertHuge = 3.0e38f;
constexpr double kCertU = 5.9604644775390625e-08;      // 2^-24
constexpr double kCertMinLateral = 0.01;               // a role must admit +-1 % of the reach around its row

MPPI_HD double cert_margin(double amax, double bmax, double cmax, double dom) {
    // |D_j - d_j| <= 3u(|a x| + |b y| + c) for D = fma(a, x, fma(b, y, c)) with rounded a, b, c; the
    // test needs d_j - d_t > 2 * that; factor 2 of slack on top
    return 16.0 * kCertU * (amax * dom + bmax * dom + cmax);
}
// rounding of fma(tx, x, fma(ty, y, k)) incl. the conversion of k to float, for |x|, |y| <= dom
MPPI_HD double cert_delta(double tx, double ty, double k, double dom) {
    return 8.0 * kCertU * ((fabs(tx) + fabs(ty)) * dom + fabs(k)) + 1e-30;
}
MPPI_HD void cert_disable(WinCert& c, int n_valid) {
    c.nx = c.ny = 0.f; c.blo = kCertHuge; c.bhi = -kCertHuge;
    c.sx = c.sy = c.s0 = 0.f; c.bc = 1.f; c.c0 = c.c1 = c.c2 = 0.f; c.jhi = 0.f;
    c.dom = 0.f; c.last = n_valid - 1; c.pad[0] = c.pad[1] = 0;
}
MPPI_HD void cert_row_disable(RowRec& r) { r.tx = 0.f; r.ty = 0.f; r.kL = -kCertHuge; r.kU = kCertHuge; }

// What one row contributes (FP64): its tangent and offset, and for each of its two roles (a) the bounds the role
// puts on [blo, bhi] when its threshold sits at the row itself, (b) how far the threshold has to be pushed
// away from the row ("ahead by f" / "behind by f") for the role to hold on b_a +- kCertOffsetLateral * reach —
// the fallback where the rows are closer together than the FP32 margin (a path that starts from rest).
struct RowRole { double lo, hi, fneed; bool ok; };
struct RowGeom { double tx, ty, k, b; RowRole L, U; };
constexpr double kCertOffsetLateral = 0.25;            // lateral half-width of a pushed role, in units of the reach
constexpr double kCertMaxOffset = 0.25;                // give up when the threshold would be > 25 cm from the row

// Row a of `n` local rows against every other row.  nu = window normal (already rounded to FP32 and widened
// back); `row(j, x, y)` fetches row j; wt = lateral half-width of the pushed form.  PUSH = false computes only
// the bounds of the threshold-at-the-row form (the common case; the caller asks again with PUSH = true for rows
// whose roles that form cannot serve).  Bounds carry the 2e-7 relative error of rcp32_: the caller widens.
template <bool PUSH, class RowFn>
MPPI_HD RowGeom cert_row_geom(RowFn row, int a, int n, double nux, double nuy, double margin, double wt) {
    RowGeom g;
    double ax, ay, px, py, qx, qy;
    row(a, ax, ay);
    row(a > 0 ? a - 1 : a, px, py);
    row(a + 1 < n ? a + 1 : a, qx, qy);
    double tx = qx - px, ty = qy - py;
    const double tn2 = tx * tx + ty * ty;
    g.L.ok = g.U.ok = tn2 > 0.0;
    g.L.lo = g.U.lo = -1e300; g.L.hi = g.U.hi = 1e300; g.L.fneed = g.U.fneed = 0.0;
    const double itn = tn2 > 0.0 ? rsqrt64_(tn2) : 0.0;
    tx *= itn; ty *= itn;
    tx = (double)(float)tx; ty = (double)(float)ty;          // the direction the FP32 test will use
    g.tx = tx; g.ty = ty; g.k = -(tx * ax + ty * ay); g.b = nux * ax + nuy * ay;
    const double det = tx * nuy - ty * nux;
    if (!(fabs(det) > 0.5)) { g.L.ok = g.U.ok = false; return g; }     // tangent more than 60 degrees off the chord
    const double idet = rcp64_(det);
    const double v0x = -g.k * nuy * idet, v0y = g.k * nux * idet;      // tau.V0 = -k, nu.V0 = 0
    const double v1x = -ty * idet, v1y = tx * idet;                    // tau.v1 = 0,  nu.v1 = 1
    const double dx = nuy * idet, dy = -nux * idet;                    // tau.d' = 1,  nu.d' = 0
    for (int j = 0; j < n; ++j) {
        if (j == a) continue;
        double jx, jy;
        row(j, jx, jy);
        const double gx = ax - jx, gy = ay - jy;
        RowRole& r = j < a ? g.L : g.U;
        const double along = j < a ? gx * dx + gy * dy : -(gx * dx + gy * dy);   // gain per unit of push
        if (!(along > 0.0)) { r.ok = false; continue; }
        const double f0 = gx * gx + gy * gy + 2.0 * (gx * (v0x - ax) + gy * (v0y - ay));
        const double sl = 2.0 * (gx * v1x + gy * v1y);
        if (!PUSH) {
            const double bnd = (margin - f0) * rcp32_(sl);
            if (sl > 0.0) r.lo = bnd > r.lo ? bnd : r.lo;
            else if (sl < 0.0) r.hi = bnd < r.hi ? bnd : r.hi;
            else if (!(f0 >= margin)) r.lo = 1e300;
        } else {
            const double need = (margin - (f0 + sl * g.b - fabs(sl) * wt)) * 0.5 * rcp32_(along);
            r.fneed = need > r.fneed ? need : r.fneed;
        }
    }
    return g;
}
// Decide the form of one role: threshold at the row (push = 0, the role's own lateral bounds), pushed by
// `push` (bounds b +- wt), or off.  Returns false when the role is off.
MPPI_HD bool cert_role_plain(const RowRole& r, double b, double wmin) {
    return r.ok && r.lo <= b - wmin && r.hi >= b + wmin;
}
// `plain` = the role as computed with PUSH = false, `pushed` = with PUSH = true (only read when plain fails)
MPPI_HD bool cert_role_form(const RowRole& plain, const RowRole& pushed, double b, double wmin, double wt,
                            double& push, double& lo, double& hi) {
    push = 0.0; lo = plain.lo; hi = plain.hi;
    if (!plain.ok) return false;
    if (cert_role_plain(plain, b, wmin)) return true;
    if (!(pushed.fneed <= kCertMaxOffset)) return false;
    push = pushed.fneed * (1.0 + 1e-6) + 1e-12; lo = b - wt; hi = b + wt;
    return true;
}

// [blo, bhi] as the FP32 test will see it: moved inwards by the rounding of b = fma(nx, x, ny * y), of the
// float conversion, and by 1e-6 relative for the FP32-accurate reciprocals of the device construction.
MPPI_HD void cert_store_range(WinCert& c, double nux, double nuy, double blo, double bhi, double bmax) {
    const double eb = (8.0 * kCertU + 1e-6) * bmax + 1e-30;
    c.nx = (float)nux; c.ny = (float)nuy;
    c.blo = (float)(blo + eb); c.bhi = (float)(bhi - eb);
}

// Index estimate (no proof needed): circle through the rows (Kasa fit on centred chord coordinates), then
// a quadratic least-squares fit of the row index on w = (s - sc) / (bc - b), which is the tangent of the
// angle about the centre and therefore linear in the index on an arc traversed at constant speed.
MPPI_HD void cert_fit_estimate(const double* s, const double* b, int n, WinCert& c, double sgx, double sgy) {
    c.jhi = 0.f;
    if (n < 3) return;
    double ms = 0, mb = 0;
    for (int j = 0; j < n; ++j) { ms += s[j]; mb += b[j]; }
    ms /= n; mb /= n;
    double suu = 0, svv = 0, suv = 0, suz = 0, svz = 0;
    for (int j = 0; j < n; ++j) {
        const double u = s[j] - ms, v = b[j] - mb, z = u * u + v * v;
        suu += u * u; svv += v * v; suv += u * v; suz += u * z; svz += v * z;
    }
    const double det = suu * svv - suv * suv;
    double sc = ms, bc = mb + 1.0e6;                                   // straight window: centre far away
    if (fabs(det) > 1e-12 * suu * suu) {
        const double uc = 0.5 * (suz * svv - svz * suv) / det, vc = 0.5 * (svz * suu - suz * suv) / det;
        if (fabs(vc) > 0.0 && fabs(vc) < 1.0e6 && fabs(uc) < 1.0e6) { sc = ms + uc; bc = mb + vc; }
    }
    double wmax = 0;
    for (int j = 0; j < n; ++j) {
        const double w = (s[j] - sc) / (bc - b[j]);
        wmax = fabs(w) > wmax ? fabs(w) : wmax;
    }
    if (!(wmax > 0.0) || !(wmax < 1e30)) return;
    double m[5] = { 0, 0, 0, 0, 0 }, r[3] = { 0, 0, 0 };                // moments of W = w / wmax
    for (int j = 0; j < n; ++j) {
        const double W = (s[j] - sc) / (bc - b[j]) / wmax;
        double p = 1.0;
        for (int e = 0; e < 5; ++e) { m[e] += p; if (e < 3) r[e] += p * j; p *= W; }
    }
    // solve [m0 m1 m2; m1 m2 m3; m2 m3 m4] q = r (Cramer)
    const double D = m[0] * (m[2] * m[4] - m[3] * m[3]) - m[1] * (m[1] * m[4] - m[3] * m[2]) + m[2] * (m[1] * m[3] - m[2] * m[2]);
    if (!(fabs(D) > 1e-300)) return;
    const double q0 = (r[0] * (m[2] * m[4] - m[3] * m[3]) - m[1] * (r[1] * m[4] - m[3] * r[2]) + m[2] * (r[1] * m[3] - m[2] * r[2])) / D;
    const double q1 = (m[0] * (r[1] * m[4] - r[2] * m[3]) - r[0] * (m[1] * m[4] - m[3] * m[2]) + m[2] * (m[1] * r[2] - m[2] * r[1])) / D;
    const double q2 = (m[0] * (m[2] * r[2] - m[3] * r[1]) - m[1] * (m[1] * r[2] - m[2] * r[1]) + r[0] * (m[1] * m[3] - m[2] * m[2])) / D;
    c.sx = (float)sgx; c.sy = (float)sgy; c.s0 = (float)(-sc); c.bc = (float)bc;
    c.c0 = (float)q0; c.c1 = (float)(q1 / wmax); c.c2 = (float)(q2 / (wmax * wmax));
    const bool fin = fabs(q0) < 1e6 && fabs(q1 / wmax) < 1e30 && fabs(q2 / (wmax * wmax)) < 1e30;
    c.jhi = fin ? (float)(n - 2) : 0.f;
}

// ---- far-field complement of the two end tests -----------------------------------------------------
// A sample that has flown far off the path (an arm falling from rest) is still nearest to an end row, but it
// may lie outside the lateral box.  For those the prepare kernel also derives, for the last valid row L (and
// for row 0), two half-planes  m_i.p' + k_i >= 0  whose intersection (a wedge) lies inside the Voronoi cell
// of the row: with g_j = r_L - r_j,  d_j - d_L = |g_j|^2 + 2 g_j.(p - r_L).  All g_j lie in the cone spanned
// by the two extreme directions m_0, m_1; for p = z + s with s.m_0 >= 0 and s.m_1 >= 0 every s.g_j >= 0,
// hence d_j - d_L >= |g_j|^2 + 2 tau n.g_j at the apex z = r_L + tau n, and tau is chosen so that this is
// >= the rounding margin for every j.  Tried after the certified triples, read from the staged table.
struct alignas(16) EndWedges {  // 64 bytes of a step block
    float lx[2], ly[2], lk[2]; // wedge inside the cell of the last valid row
    float fx[2], fy[2], fk[2]; // wedge inside the cell of row 0
    float pad[4];
};
static_assert(sizeof(EndWedges) == 64, "wedge block is 64 bytes");
MPPI_HD void wedge_disable(float (&mx)[2], float (&my)[2], float (&k)[2]) {
    mx[0] = mx[1] = 0.f; my[0] = my[1] = 0.f; k[0] = k[1] = -kCertHuge;
}
constexpr double kCertMaxTau = 0.25;                   // give up when the apex would be > 25 cm past the row
MPPI_HD void wedge_finish(double zx, double zy, const double (&mx)[2], const double (&my)[2], double dom,
                          float (&ox)[2], float (&oy)[2], float (&ok)[2]) {
    for (int i = 0; i < 2; ++i) {
        const double k = -(mx[i] * zx + my[i] * zy);
        ox[i] = (float)mx[i]; oy[i] = (float)my[i]; ok[i] = (float)(k - cert_delta(mx[i], my[i], k, dom));
    }
}
// One wedge, serial form (host tests and the reference for the warp-parallel version in the prepare kernel):
// rows[j] = (x, y) of window row j in local coordinates (FP64), n rows valid, `target` = 0 or n-1.
MPPI_HD void make_wedge(const double (*rows)[2], int n, int target, double margin, double dom,
                        float (&ox)[2], float (&oy)[2], float (&ok)[2]) {
    wedge_disable(ox, oy, ok);
    if (n < 2) return;
    const int other = target == 0 ? n - 1 : 0;
    double n0x = rows[target][0] - rows[other][0], n0y = rows[target][1] - rows[other][1];
    const double n0 = sqrt(n0x * n0x + n0y * n0y);
    if (!(n0 > 0.0)) return;
    n0x /= n0; n0y /= n0;
    double smin = 1e300, smax = -1e300;
    for (int j = 0; j < n; ++j) {
        if (j == target) continue;
        const double gx = rows[target][0] - rows[j][0], gy = rows[target][1] - rows[j][1];
        const double along = gx * n0x + gy * n0y, across = n0x * gy - n0y * gx;
        if (!(along > 0.05 * fabs(across)) || !(along > 0.0)) return;   // direction spread too wide (or duplicate rows)
        const double sl = across / along;
        smin = sl < smin ? sl : smin; smax = sl > smax ? sl : smax;
    }
    smin -= 1e-7 * (1.0 + smin * smin); smax += 1e-7 * (1.0 + smax * smax);     // widen the cone by ~1e-7 rad
    double mx[2], my[2];
    const double s2[2] = { smin, smax };
    for (int i = 0; i < 2; ++i) {
        const double vx = n0x - s2[i] * n0y, vy = n0y + s2[i] * n0x, vn = sqrt(vx * vx + vy * vy);
        mx[i] = vx / vn; my[i] = vy / vn;
    }
    double bx = mx[0] + mx[1], by = my[0] + my[1];
    const double bn = sqrt(bx * bx + by * by);
    if (!(bn > 1e-3)) return;
    bx /= bn; by /= bn;
    double tau = 0.0;
    for (int j = 0; j < n; ++j) {
        if (j == target) continue;
        const double gx = rows[target][0] - rows[j][0], gy = rows[target][1] - rows[j][1];
        const double g2 = gx * gx + gy * gy, ng = bx * gx + by * gy;
        if (!(ng > 0.0)) return;
        const double t = (margin - g2) / (2.0 * ng);
        tau = t > tau ? t : tau;
    }
    if (!(tau <= kCertMaxTau)) return;
    wedge_finish(rows[target][0] + tau * bx, rows[target][1] + tau * by, mx, my, dom, ox, oy, ok);
}
MPPI_HD void make_end_wedges(const double (*rows)[2], int n_valid, double margin, double domw, bool enabled, EndWedges& w) {
    wedge_disable(w.lx, w.ly, w.lk); wedge_disable(w.fx, w.fy, w.fk);
    w.pad[0] = w.pad[1] = w.pad[2] = w.pad[3] = 0.f;
    if (!enabled || n_valid < 2) return;
    make_wedge(rows, n_valid, n_valid - 1, margin, domw, w.lx, w.ly, w.lk);
    make_wedge(rows, n_valid, 0, margin, domw, w.fx, w.fy, w.fk);
}
// wl >= 0: inside the wedge of the last row; wf >= 0: inside the wedge of row 0 (both need |x'|, |y'| <= dom)
MPPI_HD void wedge_test(const EndWedges& c, float xl, float yl, float& wl, float& wf) {
    wl = fminf(fma_(c.lx[0], xl, fma_(c.ly[0], yl, c.lk[0])), fma_(c.lx[1], xl, fma_(c.ly[1], yl, c.lk[1])));
    wf = fminf(fma_(c.fx[0], xl, fma_(c.fy[0], yl, c.fk[0])), fma_(c.fx[1], xl, fma_(c.fy[1], yl, c.fk[1])));
}

// The certificate of one window, serial form (host tests; the prepare kernel runs the same steps with one
// lane per row).  rows: the n_valid local rows (FP64); reach = L1 + L2 of the cost-side kinematics;
// (ox, oy) = window origin; tab[j].{tx, ty, kL, kU} are filled for all kWindowPad rows.
MPPI_HD void make_win_cert(const double (*rows)[2], int n_valid, double reach, double ox, double oy,
                           bool enabled, WinCert& c, RowRec* tab, EndWedges& wed) {
    cert_disable(c, n_valid);
    for (int j = 0; j < kWindowPad; ++j) cert_row_disable(tab[j]);
    make_end_wedges(rows, 0, 0.0, 0.0, false, wed);
    if (!enabled || n_valid < 1) return;
    const double aox = fabs(ox), aoy = fabs(oy);
    const double dom = 1.01 * reach + (aox > aoy ? aox : aoy) + 0.01, domw = 1.0001 * dom;
    c.dom = (float)dom;
    if (n_valid == 1) {                                     // a one-row window: the search can only return row 0
        tab[0].kL = kCertHuge; tab[0].kU = -kCertHuge; c.blo = -kCertHuge; c.bhi = kCertHuge;
        return;
    }
    double chx = rows[n_valid - 1][0] - rows[0][0], chy = rows[n_valid - 1][1] - rows[0][1];
    const double chn = sqrt(chx * chx + chy * chy);
    if (!(chn > 0.0)) return;
    chx /= chn; chy /= chn;
    const double nux = (double)(float)(-chy), nuy = (double)(float)chx;
    double cmax = 0;
    for (int j = 0; j < n_valid; ++j) {
        const double cc = rows[j][0] * rows[j][0] + rows[j][1] * rows[j][1];
        cmax = cc > cmax ? cc : cmax;
    }
    const double ab = 2.0000001 * sqrt(cmax);               // |a_j|, |b_j| <= 2 sqrt(cmax)
    const double margin = cert_margin(ab, ab, cmax, domw);
    make_end_wedges(rows, n_valid, margin, domw, true, wed);
    const double wmin = kCertMinLateral * reach, wt = kCertOffsetLateral * reach;
    auto row = [&](int j, double& x, double& y) { x = rows[j][0]; y = rows[j][1]; };
    const double bmax = (fabs(nux) + fabs(nuy)) * domw;
    double blo = -bmax, bhi = bmax;
    double sj[kWindow], bj[kWindow];
    for (int a = 0; a < n_valid; ++a) {
        const RowGeom g = cert_row_geom<false>(row, a, n_valid, nux, nuy, margin, wt);
        RowGeom gp = g;
        if ((a > 0 && !cert_role_plain(g.L, g.b, wmin)) || (a < n_valid - 1 && !cert_role_plain(g.U, g.b, wmin)))
            gp = cert_row_geom<true>(row, a, n_valid, nux, nuy, margin, wt);
        RowRec& r = tab[a];
        r.tx = (float)g.tx; r.ty = (float)g.ty;
        double push, lo, hi;
        if (a == 0) r.kL = kCertHuge;                        // no rows below row 0
        else if (cert_role_form(g.L, gp.L, g.b, wmin, wt, push, lo, hi)) {
            r.kL = (float)(g.k - push - cert_delta(g.tx, g.ty, g.k - push, domw));
            blo = lo > blo ? lo : blo; bhi = hi < bhi ? hi : bhi;
        }
        if (a == n_valid - 1) r.kU = -kCertHuge;             // no rows above the last one
        else if (cert_role_form(g.U, gp.U, g.b, wmin, wt, push, lo, hi)) {
            r.kU = (float)(g.k + push + cert_delta(g.tx, g.ty, g.k + push, domw));
            blo = lo > blo ? lo : blo; bhi = hi < bhi ? hi : bhi;
        }
        sj[a] = chx * rows[a][0] + chy * rows[a][1]; bj[a] = g.b;
    }
    // rounding of b = fma(nx, x, ny * y) and of the conversion of the bounds
    cert_store_range(c, nux, nuy, blo, bhi, bmax);
    cert_fit_estimate(sj, bj, n_valid, c, chx, chy);
}

// ---- the FP32 side ---------------------------------------------------------------------------------
// Exact search over a table in memory: first arg-min of D_j = fma(a, x, fma(b, y, c)), strict `<` in index
// order like list.index(min(d)) (the same D_j, hence the same answer, as the register tournament below).
MPPI_HD int nearest_scan(const RowRec* tab, float xl, float yl) {
    int best = 0;
    float bd = fma_(tab[0].a, xl, fma_(tab[0].b, yl, tab[0].c));
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int j0 = 1; j0 < kWindow; j0 += 5) {          // 29 = 5 x 5 + 4 candidates: blocks of five loads in flight
        float d[5];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 5; ++i) {
            const RowRec& r = tab[j0 + i < kWindowPad ? j0 + i : kWindowPad - 1];
            d[i] = fma_(r.a, xl, fma_(r.b, yl, r.c));
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 5; ++i)
            if (j0 + i < kWindow && d[i] < bd) { bd = d[i]; best = j0 + i; }
    }
    return best;
}

struct CertEnds {              // what every lookup needs, kept in registers by the rollouts:
    float lx, ly, lk;          // L(last)
    float fx, fy, fk;          // U(0)
    float nx, ny, blo, bhi, dom;   // the box the certificate was proven on
    int last;
};
MPPI_HD CertEnds cert_ends(const WinCert& c, const RowRec* tab) {
    CertEnds e;
    const int last = c.last < 0 ? 0 : c.last;
    e.lx = tab[last].tx; e.ly = tab[last].ty; e.lk = tab[last].kL;
    e.fx = tab[0].tx; e.fy = tab[0].ty; e.fk = tab[0].kU;
    e.nx = c.nx; e.ny = c.ny; e.blo = c.blo; e.bhi = c.bhi; e.dom = c.dom; e.last = c.last;
    return e;
}
struct LookupStats { int tri, scan; };     // lookups answered by a certified triple / by the exact search (the rest: end tests)
struct LookupMode { bool far; };           // per sample slot, warp-uniform: the last lookup was answered by a far-field wedge

// (bitwise & on purpose: every operand is evaluated, the compiler chains predicates instead of branching)
MPPI_HD bool cert_in_box(const CertEnds& c, float xl, float yl, float& b) {
    b = fma_(c.nx, xl, mul_(c.ny, yl));
    return (fmaxf(fabsf(xl), fabsf(yl)) <= c.dom) & (b >= c.blo) & (b <= c.bhi);      // false for NaN
}
// The rollouts' form.  Their query is fk_local() of sin / cos values that sincos_fix bounds by 1 + 2^-23 for
// every argument, so |x'| <= (L1 + L2)(1 + 2^-23) + |ox| < dom = 1.01 (L1 + L2) + max(|ox|, |oy|) + 0.01 (and the
// same for y') holds by construction and is never NaN: the domain test of the certificate is implied.
#ifndef MPPI_CERT_DOM_TEST
#define MPPI_CERT_DOM_TEST (!MPPI_ANGLE_FIX)
#endif
MPPI_HD bool cert_in_box_fk(const CertEnds& c, float xl, float yl, float& b) {
#if MPPI_CERT_DOM_TEST
    return cert_in_box(c, xl, yl, b);
#else
    b = fma_(c.nx, xl, mul_(c.ny, yl));
    return (b >= c.blo) & (b <= c.bhi);
#endif
}
MPPI_HD bool cert_dom_fk(const CertEnds& c, float xl, float yl) {
#if MPPI_CERT_DOM_TEST
    return fmaxf(fabsf(xl), fabsf(yl)) <= c.dom;
#else
    return true;
#endif
}
MPPI_HD int cert_guess(const WinCert& c, float xl, float yl, float b) {
    const float s = fma_(c.sx, xl, fma_(c.sy, yl, c.s0));
#if defined(__CUDA_ARCH__)
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(sub_(c.bc, b)));
#else
    const float inv = 1.0f / sub_(c.bc, b);
#endif
    const float w = mul_(s, inv);
    const float est = fma_(w, fma_(w, c.c2, c.c1), c.c0);
    const float cl = fminf(fmaxf(est, 1.0f), c.jhi);          // NaN -> 1
    return f2i(add_(cl, 12582912.0f)) - 0x4B400000;            // round to nearest, exact for 0 <= cl < 2^22
}
// j0-1, j0, j0+1 compared exactly like the full search compares them
MPPI_HD int cert_triple_pick(const RowRec* t, int j0, float xl, float yl) {
    const float d0 = fma_(t[0].a, xl, fma_(t[0].b, yl, t[0].c));
    const float d1 = fma_(t[1].a, xl, fma_(t[1].b, yl, t[1].c));
    const float d2 = fma_(t[2].a, xl, fma_(t[2].b, yl, t[2].c));
    int j = j0 - 1;
    float d = d0;
    if (d1 < d) { d = d1; j = j0; }
    if (d2 < d) j = j0 + 1;
    return j;
}
MPPI_HD bool cert_triple_ok(const RowRec* t, float xl, float yl) {
    return (fma_(t[0].tx, xl, fma_(t[0].ty, yl, t[0].kL)) >= 0.0f) & (fma_(t[2].tx, xl, fma_(t[2].ty, yl, t[2].kU)) <= 0.0f);
}
// Certified answer for one query, or -1 (tests; the rollouts vote per warp in nearest_wp)
MPPI_HD int cert_pick(const WinCert& c, const RowRec* tab, const EndWedges& wed, float xl, float yl) {
    float b;
    const CertEnds e = cert_ends(c, tab);
    if (cert_in_box(e, xl, yl, b)) {
        if (fma_(e.lx, xl, fma_(e.ly, yl, e.lk)) >= 0.0f) return c.last;
        if (fma_(e.fx, xl, fma_(e.fy, yl, e.fk)) <= 0.0f) return 0;
        if (c.jhi >= 1.0f) {
            const int j0 = cert_guess(c, xl, yl, b);
            const RowRec* t = tab + (j0 - 1);
            if (cert_triple_ok(t, xl, yl)) return cert_triple_pick(t, j0, xl, yl);
        }
    }
    float wl, wf;
    wedge_test(wed, xl, yl, wl, wf);
    if (!(fmaxf(fabsf(xl), fabsf(yl)) <= e.dom)) return -1;
    return wl >= 0.0f ? c.last : (wf >= 0.0f ? 0 : -1);
}

// Where the 30 x (a, b, c) window coefficients live.  WinRegs: per-thread registers (any number of
// environments).  The kernels add WinConst: the constant bank, read as immediate FFMA operands.
struct WinRegs {
    float wa[kWindow], wb[kWindow], wc[kWindow];
    MPPI_HD float a(int j) const { return wa[j]; }
    MPPI_HD float b(int j) const { return wb[j]; }
    MPPI_HD float c(int j) const { return wc[j]; }
    MPPI_HD void load(const WinEntry* tab) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < kWindow; ++j) { const WinEntry w = tab[j]; wa[j] = w.a; wb[j] = w.b; wc[j] = w.c; }
    }
};

// Nearest-waypoint search (control.py:208-215): first arg-min over the 30 window candidates.
// Exact FP32 comparisons on d_j - |p'|^2 = c_j + a_j x' + b_j y', as a tournament tree of depth 5;
// `<` is strict and the right operand always carries the larger index, so ties keep the first
// candidate like list.index(min(d)) does.
// d[j] = c_j + a_j x' + b_j y' for the 30 candidates; a window policy may provide its own
// `distances` (the constant-bank policy of the kernels uses packed FFMA2 there)
template <class Win>
MPPI_HD auto window_distances(const Win& win, float xl, float yl, float (&d)[kWindowPad], int)
    -> decltype(win.distances(xl, yl, d), void()) { win.distances(xl, yl, d); }
template <class Win>
MPPI_HD void window_distances(const Win& win, float xl, float yl, float (&d)[kWindowPad], long) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < kWindow; ++j) d[j] = fma_(win.a(j), xl, fma_(win.b(j), yl, win.c(j)));
}
template <class Win>
MPPI_HD void window_distances(const Win& win, float xl, float yl, float (&d)[kWindowPad]) {
    window_distances(win, xl, yl, d, 0);
}

template <class Win>
MPPI_HD int nearest_candidate(const Win& win, float xl, float yl) {
    float d[kWindowPad];
    float id[kWindowPad / 2];            // indices as small exact floats
    window_distances(win, xl, yl, d);
    d[30] = kSentinel; d[31] = kSentinel;
    // level 1: adjacent pairs; the index is 2i + [d(2i+1) < d(2i)]
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < kWindowPad / 2; ++i) {
        const float lt = d[2 * i + 1] < d[2 * i] ? 1.0f : 0.0f;
        id[i] = add_(lt, (float)(2 * i));
        d[i] = fminf(d[2 * i + 1], d[2 * i]);     // d[i] is only overwritten after d[2i], d[2i+1] were read
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int n = kWindowPad / 4; n >= 1; n /= 2) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < n; ++i) {
            const bool lt = d[2 * i + 1] < d[2 * i];
            id[i] = lt ? id[2 * i + 1] : id[2 * i];
            d[i] = lt ? d[2 * i + 1] : d[2 * i];
        }
    }
    return (int)id[0];
}

// Window policy of the certified kernels: the table stays in memory (shared memory on the device), only
// the two end tests live in registers.
struct WinTable {
    const RowRec* tab; const EndWedges* wed; CertEnds ends;
    MPPI_HD void load(const WinCert& c, const RowRec* t, const EndWedges* w) { tab = t; wed = w; ends = cert_ends(c, t); }
};

// The lookup of the rollouts (control.py:200-215 via _c / _phi).  Every decision is taken per WARP (one
// vote, no divergence): an end row when all lanes pass that end test; else the certified triples when all
// lanes pass theirs; else the exact search.  All three return what the exact search returns.
#if defined(__CUDA_ARCH__)
#define MPPI_ALL_LANES(p) __all_sync(0xffffffffu, (p))
#else
#define MPPI_ALL_LANES(p) (p)
#endif
MPPI_HD int nearest_wp(const WinTable& win, const WinCert& c, float xl, float yl, LookupStats& st, LookupMode& md) {
    const CertEnds& e = win.ends;
    if (md.far) {                              // a sample that left the lateral box usually stays out: wedges first
        float wl, wf;
        wedge_test(*win.wed, xl, yl, wl, wf);
        const bool dom_ok = cert_dom_fk(e, xl, yl);
        if (MPPI_ALL_LANES(dom_ok & (wl >= 0.0f))) return e.last;
        if (MPPI_ALL_LANES(dom_ok & (wf >= 0.0f))) return 0;
        md.far = false;
    }
    float b;
    const bool in = cert_in_box_fk(e, xl, yl, b);
    if (MPPI_ALL_LANES(in & (fma_(e.lx, xl, fma_(e.ly, yl, e.lk)) >= 0.0f))) return e.last;
    if (MPPI_ALL_LANES(in & (fma_(e.fx, xl, fma_(e.fy, yl, e.fk)) <= 0.0f))) return 0;
    if (c.jhi >= 1.0f) {
        const int j0 = cert_guess(c, xl, yl, b);
        const RowRec* t = win.tab + (j0 - 1);
        if (MPPI_ALL_LANES(in & cert_triple_ok(t, xl, yl))) { ++st.tri; return cert_triple_pick(t, j0, xl, yl); }
    }
    // far field: the wedges of the two end rows (samples that left the lateral box)
    float wl, wf;
    wedge_test(*win.wed, xl, yl, wl, wf);
    const bool dom_ok = cert_dom_fk(e, xl, yl);
    if (MPPI_ALL_LANES(dom_ok & (wl >= 0.0f))) { md.far = true; return e.last; }
    if (MPPI_ALL_LANES(dom_ok & (wf >= 0.0f))) { md.far = true; return 0; }
    ++st.scan;
    return nearest_scan(win.tab, xl, yl);
}
// the same for a window policy that holds the coefficients itself (kernels without the certificate)
template <class Win>
MPPI_HD int nearest_wp(const Win& win, const WinCert&, float xl, float yl, LookupStats& st, LookupMode&) {
    ++st.scan;
    return nearest_candidate(win, xl, yl);
}

// NS samples advance in lockstep inside one thread: they share the window registers, the per-step
// constants and the loop overhead, and give the scheduler NS independent instruction streams.
// Win = WinTable: certified lookups; any other window policy: plain searches (MPPI_FLAG_FULL_SEARCH: no test, no vote)
// JL: the stage cost carries the joint-limit term (a template flag: the default kernels do not contain it)
template <int NS, int DYN = 0, bool JL = false, class Win, class Noise>
MPPI_HD void rollout_cost_n(const StepHeader& hd, const ArmF& A, const CostW& W,
                            const Win& win, const WinCert& cert, const RefRow* rows, const RefRow* srows, const StepCtl* ctl,
                            int T, const float (&um)[NS], Noise (&noise)[NS], float (&S_out)[NS], LookupStats& hits) {
    ArmState st[NS];
    LookupMode md[NS];
    float S[NS], kS[NS], xl[NS], yl[NS];
    int j[NS];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int s = 0; s < NS; ++s) {
        arm_init(st[s], hd.q1, hd.q2, hd.d1, hd.d2, hd.a1, hd.a12);
        S[s] = 0.f; kS[s] = 0.f; xl[s] = 0.f; yl[s] = 0.f; j[s] = 0;
        md[s].far = false;
    }
#if defined(__CUDA_ARCH__)
#pragma unroll kUnrollT
#endif
    for (int t = 0; t < T; ++t) {
        const StepCtl c = ctl[t];
        float v1[NS], v2[NS];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int s = 0; s < NS; ++s) {
            float n1, n2;
            noise[s](t, n1, n2);
            v1[s] = fma_(um[s], c.u1, n1);             // control.py:98-101 (um = 0 for exploration samples)
            v2[s] = fma_(um[s], c.u2, n2);
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int s = 0; s < NS; ++s) arm_step<DYN, JL>(st[s], A, v1[s], v2[s]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int s = 0; s < NS; ++s) {
            fk_local(st[s], A, hd.ox, hd.oy, xl[s], yl[s]);
            j[s] = nearest_wp(win, cert, xl[s], yl[s], hits, md[s]);
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int s = 0; s < NS; ++s) {
#if MPPI_STAGE_FOLD
            float cst = stage_cost(W, st[s], xl[s], yl[s], srows[j[s]]);
#else
            float ex, ey, e1, e2;
            residuals(st[s], xl[s], yl[s], rows[j[s]], ex, ey, e1, e2);
            float cst = wsq(W.s0, W.s1, W.s2, W.s3, ex, ey, e1, e2);
#endif
            cst = fma_(c.g1, v1[s], fma_(c.g2, v2[s], cst));   // + gamma * u^T Sigma^-1 v  (control.py:106)
            if (JL) cst = add_(cst, joint_limit_cost(W, st[s].q1, st[s].q2));
#if (MPPI_KAHAN_MASK & 4)
            kahan_(S[s], kS[s], sub_(cst, kS[s]));
#else
            S[s] = add_(S[s], cst);
#endif
        }
    }
    // terminal cost on the same final state and the same nearest waypoint (control.py:109, Q5)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int s = 0; s < NS; ++s) {
        float ex = 0.f, ey = 0.f, e1 = 0.f, e2 = 0.f;
        if (T > 0) residuals(st[s], xl[s], yl[s], rows[j[s]], ex, ey, e1, e2);
        S_out[s] = add_(S[s], sub_(wsq(W.t0, W.t1, W.t2, W.t3, ex, ey, e1, e2), kS[s]));
    }
}

template <int DYN = 0, bool JL = false, class Win, class Noise>
MPPI_HD float rollout_cost(const StepHeader& hd, const ArmF& A, const CostW& W,
                           const Win& win, const WinCert& cert, const RefRow* rows, const RefRow* srows, const StepCtl* ctl,
                           int T, float um, Noise& noise, LookupStats& hits) {
    const float ums[1] = { um };
    float out[1];
    Noise (&nz)[1] = reinterpret_cast<Noise (&)[1]>(noise);
    rollout_cost_n<1, DYN, JL>(hd, A, W, win, cert, rows, srows, ctl, T, ums, nz, out, hits);
    return out[0];
}

}  // namespace mppi

// =================================================================================================
// Step-block construction (FP64 in, FP32 tables out) — the per-entry pieces of the prepare kernel.
// =================================================================================================
namespace mppi {

// Squared distance *100 exactly as control.py:210-212 computes it (FP64).
MPPI_HD double waypoint_d(const double* ref, int row, double x, double y) {
    double dx = x - ref[4 * row + 0], dy = y - ref[4 * row + 1];
    return (dx * dx + dy * dy) * 100;
}

// Row j of the window starting at waypoint p, in coordinates local to row p.
MPPI_HD void make_window_row(const double* ref, int n_rows, int p, int j, WinEntry& w, RefRow& r) {
    const int row = p + j;
    if (j < kWindow && row < n_rows) {
        double rx = ref[4 * row + 0] - ref[4 * p + 0];
        double ry = ref[4 * row + 1] - ref[4 * p + 1];
        w.a = (float)(-2.0 * rx); w.b = (float)(-2.0 * ry); w.c = (float)(rx * rx + ry * ry); w.pad = 0.f;
        r.rx = (float)rx; r.ry = (float)ry; r.rd1 = (float)ref[4 * row + 2]; r.rd2 = (float)ref[4 * row + 3];
    } else {                       // beyond the end of the path (control.py:208-209) or table padding
        w.a = 0.f; w.b = 0.f; w.c = kSentinel; w.pad = 0.f;
        r.rx = 0.f; r.ry = 0.f; r.rd1 = 0.f; r.rd2 = 0.f;
    }
}

// The same plus the pre-scaled row of the stage cost.
MPPI_HD void make_window_row(const double* ref, int n_rows, int p, int j, const CostW& W, WinEntry& w, RefRow& r, RefRow& sr) {
    make_window_row(ref, n_rows, p, j, w, r);
    const int row = p + j;
    if (j < kWindow && row < n_rows)
        sr = stage_row(W, ref[4 * row + 0] - ref[4 * p + 0], ref[4 * row + 1] - ref[4 * p + 1], ref[4 * row + 2], ref[4 * row + 3]);
    else sr = r;
}

// Nominal control of horizon step t and the row vector gamma * u_t^T Sigma^-1 (control.py:106).
MPPI_HD void make_step_ctl(const double* u_t, double gamma, const double* sig_inv, StepCtl& c) {
    c.u1 = (float)u_t[0]; c.u2 = (float)u_t[1];
    c.g1 = (float)(gamma * (u_t[0] * sig_inv[0] + u_t[1] * sig_inv[2]));
    c.g2 = (float)(gamma * (u_t[0] * sig_inv[1] + u_t[1] * sig_inv[3]));
}

}  // namespace mppi

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

// Which accumulators of a rollout carry a Kahan compensation term: bit 0 joint rates, bit 1 joint angles,
// bit 2 the cost sum S.  Default: all.  (CPU study, DESIGN.md section 8: bit 2 buys nothing measurable.)
#ifndef MPPI_KAHAN_MASK
#ifdef MPPI_NO_KAHAN
#define MPPI_KAHAN_MASK 0
#else
#define MPPI_KAHAN_MASK 7
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
#else
MPPI_HD float fma_(float a, float b, float c) { return fmaf(a, b, c); }
MPPI_HD float mul_(float a, float b) { return a * b; }
MPPI_HD float add_(float a, float b) { return a + b; }
MPPI_HD float sub_(float a, float b) { return a - b; }
MPPI_HD float rcp_(float a) { return 1.0f / a; }
MPPI_HD int f2i(float a) { union { float f; int i; } u; u.f = a; return u.i; }
MPPI_HD float i2f(int a) { union { float f; int i; } u; u.i = a; return u.f; }
#endif

// ---- sin & cos of one angle, ~1 ulp, no slow path --------------------------------------------
// Cody-Waite reduction by pi/2 in three FMA steps, then degree-7 / degree-8 minimax polynomials on
// [-pi/4, pi/4].  Valid for |x| < ~1e5 rad; a diverged rollout (larger angle, Inf, NaN) yields a
// garbage-but-finite or NaN cost that the soft-min kernel maps to weight 0.
MPPI_HD void sincos_(float x, float& s, float& c) {
    const float kMagic = 12582912.0f;                      // 1.5 * 2^23: round-to-nearest trick
    float kf = fma_(x, 0.636619772367581343f, kMagic);
    int q = f2i(kf);                                       // low bits hold the quadrant
    kf = sub_(kf, kMagic);
    float r = fma_(kf, -1.57079601287841796875f, x);
    r = fma_(kf, -3.1391647326017846e-07f, r);
    r = fma_(kf, -5.3903025299577648e-15f, r);
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
    float ss = (q & 1) ? cr : sr;
    float cc = (q & 1) ? sr : cr;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
}

// ---- per-controller constants (derived once on the host in FP64, rounded to FP32) -----------
struct ArmF {
    float A0, A1;        // M11 = A0 + A1*cos q2         (control.py:241-242)
    float M22, B1;       // M12 = M22 + B1*cos q2; h = B1*sin q2   (control.py:243-244, 247)
    float G1a, G1b;      // g1 = G1a*cos q1 + G1b*cos q12; g2 = G1b*cos q12   (control.py:248-249)
    float dt;            // controller integration step (control.py:240)
    float L1, L2;        // cost-side link lengths self.l1 / self.l2 (control.py:55-56)
};

struct CostW {           // weights already multiplied by 1e4 (control.py:185, 198)
    float s0, s1, s2, s3;
    float t0, t1, t2, t3;
};

struct WinEntry { float a, b, c, pad; };     // d_j - |p'|^2 = c + a*x' + b*y'   (local coordinates)
struct RefRow { float rx, ry, rd1, rd2; };   // waypoint in local coordinates + reference joint rates
struct StepCtl { float u1, u2, g1, g2; };    // nominal control and gamma*(u^T Sigma^-1)

// Arm state carried through the horizon.  sin/cos of q1 and q1+q2 are kept from the previous step
// (they were needed for its forward kinematics) so each step evaluates two sincos, not eight cos/sin.
struct ArmState {
    float q1, q2, d1, d2;
    float s1, c1, s12, c12;
    float kq1, kq2, kd1, kd2;      // Kahan compensation terms of the four integrators
};

MPPI_HD void arm_init(ArmState& st, float q1, float q2, float d1, float d2) {
    st.q1 = q1; st.q2 = q2; st.d1 = d1; st.d2 = d2;
    st.kq1 = 0.f; st.kq2 = 0.f; st.kd1 = 0.f; st.kd2 = 0.f;
    sincos_(q1, st.s1, st.c1);
    sincos_(add_(q1, q2), st.s12, st.c12);
}

// acc += y with the rounding error left in `comp` (y already has the old comp subtracted)
MPPI_HD void kahan_(float& acc, float& comp, float y) {
    float t = add_(acc, y);
    comp = sub_(sub_(t, acc), y);
    acc = t;
}

// One integration step (control.py:241-259) under control (v1, v2).
// DYN = 0: the arm model _F.  DYN = 1: the reference's other rollout model _F1 (control.py:265-295),
// which forms u = M v + C dq (gravity dropped, control.py:281-284) and solves ddq = M^-1 (u - C dq):
// ddq = v up to FP64 rounding, so the input is applied as the joint acceleration.
template <int DYN = 0>
MPPI_HD void arm_step(ArmState& st, const ArmF& A, float v1, float v2) {
    if (DYN == 1) {
#if (MPPI_KAHAN_MASK & 1)
        kahan_(st.d1, st.kd1, fma_(v1, A.dt, -st.kd1));
        kahan_(st.d2, st.kd2, fma_(v2, A.dt, -st.kd2));
#else
        st.d1 = fma_(v1, A.dt, st.d1); st.d2 = fma_(v2, A.dt, st.d2);
#endif
#if (MPPI_KAHAN_MASK & 2)
        kahan_(st.q1, st.kq1, fma_(st.d1, A.dt, -st.kq1));
        kahan_(st.q2, st.kq2, fma_(st.d2, A.dt, -st.kq2));
#else
        st.q1 = fma_(st.d1, A.dt, st.q1); st.q2 = fma_(st.d2, A.dt, st.q2);
#endif
        sincos_(st.q1, st.s1, st.c1);
        sincos_(add_(st.q1, st.q2), st.s12, st.c12);
        return;
    }
    // cos/sin of q2 = (q1+q2) - q1 by the angle-difference identity
    float c2 = fma_(st.c12, st.c1, mul_(st.s12, st.s1));
    float s2 = fma_(st.s12, st.c1, -mul_(st.c12, st.s1));
    float M11 = fma_(A.A1, c2, A.A0);
    float M12 = fma_(A.B1, c2, A.M22);
    float h = mul_(A.B1, s2);
    float g2 = mul_(A.G1b, st.c12);
    float g1 = fma_(A.G1a, st.c1, g2);
    // v - C dq - G with C dq = [-h d2 (2 d1 + d2), h d1^2]
    float tt = fma_(2.0f, st.d1, st.d2);
    float b1 = fma_(mul_(h, st.d2), tt, sub_(v1, g1));
    float b2 = fma_(-mul_(h, st.d1), st.d1, sub_(v2, g2));
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
#if (MPPI_KAHAN_MASK & 2)
    kahan_(st.q1, st.kq1, fma_(st.d1, A.dt, -st.kq1));
    kahan_(st.q2, st.kq2, fma_(st.d2, A.dt, -st.kq2));
#else
    st.q1 = fma_(st.d1, A.dt, st.q1);
    st.q2 = fma_(st.d2, A.dt, st.q2);
#endif
    sincos_(st.q1, st.s1, st.c1);
    sincos_(add_(st.q1, st.q2), st.s12, st.c12);
}

// End-effector in window-local coordinates: (x - ox, y - oy), origin = first row of the window.
MPPI_HD void fk_local(const ArmState& st, const ArmF& A, float ox, float oy, float& xl, float& yl) {
    xl = fma_(A.L2, st.c12, fma_(A.L1, st.c1, -ox));
    yl = fma_(A.L2, st.s12, fma_(A.L1, st.s1, -oy));
}

// Candidate key: distance (minus the common |p'|^2) with the candidate index in the 5 low mantissa
// bits, so a plain float min returns value and arg-min together (FMNMX3 on sm_100a).
MPPI_HD float cand_key(float a, float b, float c, float xl, float yl, int j) {
    float d = fma_(a, xl, fma_(b, yl, c));
    return i2f((f2i(d) & ~31) | j);
}

MPPI_HD float min3_(float a, float b, float c) { return fminf(fminf(a, b), c); }

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
MPPI_HD void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
    float u = u01_(a), v = u01_(b);
#if defined(__CUDA_ARCH__)
    float r;                                                          // sqrt(-2 ln u), MUFU.SQRT (no slow path)
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(mul_(-1.3862943611198906f, __log2f(u))));
    float sn, cs;
    __sincosf(mul_(6.2831853071795865f, v), &sn, &cs);
#else
    float r = sqrtf(mul_(-1.3862943611198906f, log2f(u)));
    float th = mul_(6.2831853071795865f, v);
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
    int32_t pad[7];
};
static_assert(sizeof(StepHeader) == 64, "header is 64 bytes");

// ---- end-of-window certificate ------------------------------------------------------------------
// The window (control.py:208-209) is fixed for the whole horizon while the rollouts move along the
// path at about one waypoint per step, so after ~10-20 horizon steps every sample of a warp lies
// beyond the last window row (or, for an arm that falls back, before the first) and the 30-candidate
// search can only return that row.  The prepare kernel derives, in FP64, two half-planes whose
// intersection (a wedge) lies inside the Voronoi cell of the last row — and two for row 0 — with a
// margin that covers every FP32 rounding of the search AND of the test itself: if
//     mx_i x' + my_i y' + k_i >= 0  (i = 0, 1)   and   |x'|, |y'| <= dom
// then the full FP32 search is guaranteed to return that row, so it is skipped when a whole warp is
// certified.  Results are bit-identical with and without the shortcut (tests: emulation + GPU).
// Derivation: with g_j = r_L - r_j,  d_j - d_L = |g_j|^2 + 2 g_j.(p - r_L).  All g_j lie in the cone
// spanned by the two extreme directions m_0, m_1; for p = z + s with s.m_0 >= 0 and s.m_1 >= 0 every
// s.g_j >= 0, hence d_j - d_L >= |g_j|^2 + 2 tau n.g_j at the apex z = r_L + tau n, and tau is chosen
// so that this is >= the rounding margin for every j.
struct alignas(16) EndCert {   // 64 bytes of a step block
    float lx[2], ly[2], lk[2]; // wedge inside the cell of the last valid row
    float fx[2], fy[2], fk[2]; // wedge inside the cell of row 0
    float dom;                 // the margins hold for |x'|, |y'| <= dom
    int32_t last;              // index of the last valid row (n_valid - 1)
    int32_t pad[2];
};
static_assert(sizeof(EndCert) == 64, "certificate block is 64 bytes");

// Certificate test for (x', y'): wl >= 0 certifies the last row, wf >= 0 certifies row 0 (the two
// wedges are disjoint), both only inside the domain the margins were derived for.
struct CertTest { float wl, wf; bool in_dom; };
MPPI_HD CertTest cert_test(const EndCert& c, float xl, float yl) {
    CertTest t;
#if defined(__CUDA_ARCH__) && !defined(MPPI_CERT_SCALAR)
    // the two half-planes of a wedge as one packed fma.rn.f32x2 chain (same roundings as the scalar form)
    const unsigned long long x2 = ((unsigned long long)__float_as_uint(xl) << 32) | __float_as_uint(xl);
    const unsigned long long y2 = ((unsigned long long)__float_as_uint(yl) << 32) | __float_as_uint(yl);
    const unsigned long long* q = reinterpret_cast<const unsigned long long*>(&c);   // lx, ly, lk, fx, fy, fk pairs
    unsigned long long a, b;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(a) : "l"(q[1]), "l"(y2), "l"(q[2]));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(a) : "l"(q[0]), "l"(x2), "l"(a));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(b) : "l"(q[4]), "l"(y2), "l"(q[5]));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(b) : "l"(q[3]), "l"(x2), "l"(b));
    t.wl = fminf(__uint_as_float((unsigned)a), __uint_as_float((unsigned)(a >> 32)));
    t.wf = fminf(__uint_as_float((unsigned)b), __uint_as_float((unsigned)(b >> 32)));
#else
    t.wl = fminf(fma_(c.lx[0], xl, fma_(c.ly[0], yl, c.lk[0])), fma_(c.lx[1], xl, fma_(c.ly[1], yl, c.lk[1])));
    t.wf = fminf(fma_(c.fx[0], xl, fma_(c.fy[0], yl, c.fk[0])), fma_(c.fx[1], xl, fma_(c.fy[1], yl, c.fk[1])));
#endif
    t.in_dom = fmaxf(fabsf(xl), fabsf(yl)) <= c.dom;               // false for NaN
    return t;
}
MPPI_HD bool cert_ok(const CertTest& t) { return t.in_dom && fmaxf(t.wl, t.wf) >= 0.0f; }
MPPI_HD int cert_row(const EndCert& c, const CertTest& t) { return t.wl >= 0.0f ? c.last : 0; }
// index the search is certain to return for (x', y'), or -1 when nothing is certified
MPPI_HD int cert_pick(const EndCert& c, float xl, float yl) {
    const CertTest t = cert_test(c, xl, yl);
    return cert_ok(t) ? cert_row(c, t) : -1;
}

MPPI_HD void cert_disable(float (&mx)[2], float (&my)[2], float (&k)[2]) {
    mx[0] = mx[1] = 0.f; my[0] = my[1] = 0.f; k[0] = k[1] = -INFINITY;
}

// One wedge: rows[j] = (x, y) of window row j in local coordinates (FP64), n rows valid, `target`
// = 0 or n-1.  amax/bmax/cmax bound |a_j|, |b_j|, c_j of the FP32 table.  Serial form (host tests and
// the reference for the warp-parallel version in the prepare kernel).
constexpr double kCertU = 5.9604644775390625e-08;      // 2^-24
constexpr double kCertMaxTau = 0.25;                   // give up when the apex would be > 25 cm past the row
MPPI_HD double cert_margin(double amax, double bmax, double cmax, double dom) {
    // |D_j - d_j| <= 3u(|a x| + |b y| + c) for D = fma(a, x, fma(b, y, c)) with rounded a, b, c; the
    // test needs d_j - d_t > 2 * that; factor 2 of slack on top
    return 16.0 * kCertU * (amax * dom + bmax * dom + cmax);
}
MPPI_HD void cert_finish(double zx, double zy, const double (&mx)[2], const double (&my)[2], double dom,
                         float (&ox)[2], float (&oy)[2], float (&ok)[2]) {
    for (int i = 0; i < 2; ++i) {
        const double k = -(mx[i] * zx + my[i] * zy);
        // the FP32 value of mx x + my y + k differs from the exact one by <= 3u(|x| + |y| + |k|)
        const double delta = 8.0 * kCertU * (2.0 * dom + fabs(k)) + 1e-30;
        ox[i] = (float)mx[i]; oy[i] = (float)my[i]; ok[i] = (float)(k - delta);
    }
}
MPPI_HD void make_wedge(const double (*rows)[2], int n, int target, double margin, double dom,
                        float (&ox)[2], float (&oy)[2], float (&ok)[2]) {
    cert_disable(ox, oy, ok);
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
        const double s = across / along;
        smin = s < smin ? s : smin; smax = s > smax ? s : smax;
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
    cert_finish(rows[target][0] + tau * bx, rows[target][1] + tau * by, mx, my, dom, ox, oy, ok);
}
// rows: the n_valid local rows (FP64); reach = L1 + L2 of the cost-side kinematics; (ox, oy) = window origin
MPPI_HD void make_end_cert(const double (*rows)[2], int n_valid, double reach, double ox, double oy,
                           bool enabled, EndCert& c) {
    const double aox = fabs(ox), aoy = fabs(oy);
    const double dom = 1.01 * reach + (aox > aoy ? aox : aoy) + 0.01;
    c.dom = (float)dom; c.last = n_valid - 1; c.pad[0] = c.pad[1] = 0;
    cert_disable(c.lx, c.ly, c.lk); cert_disable(c.fx, c.fy, c.fk);
    if (!enabled || n_valid < 1) return;
    if (n_valid == 1) { c.fk[0] = c.fk[1] = 1.0f; return; }       // a one-row window: the search can only return row 0
    double cmax = 0;
    for (int j = 0; j < n_valid; ++j) {
        const double cc = rows[j][0] * rows[j][0] + rows[j][1] * rows[j][1];
        cmax = cc > cmax ? cc : cmax;
    }
    const double ab = 2.0000001 * sqrt(cmax);                 // |a_j|, |b_j| <= 2 sqrt(cmax)
    const double margin = cert_margin(ab, ab, cmax, 1.0001 * dom);
    make_wedge(rows, n_valid, n_valid - 1, margin, 1.0001 * dom, c.lx, c.ly, c.lk);
    make_wedge(rows, n_valid, 0, margin, 1.0001 * dom, c.fx, c.fy, c.fk);
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

template <class Win>
MPPI_HD int nearest_wp(const Win& win, float xl, float yl) { return nearest_candidate(win, xl, yl); }

// The lookup of the rollouts: the certified row when the whole warp is certified (one vote, no
// divergence), else the full search.  `hits` counts the skipped searches (per warp on the device).
template <class Win>
MPPI_HD int nearest_wp(const Win& win, const EndCert& cert, float xl, float yl, int& hits) {
    const CertTest t = cert_test(cert, xl, yl);
#if defined(__CUDA_ARCH__)
    if (__all_sync(0xffffffffu, cert_ok(t))) { ++hits; return cert_row(cert, t); }
#else
    if (cert_ok(t)) { ++hits; return cert_row(cert, t); }
#endif
    return nearest_candidate(win, xl, yl);
}

// NS samples advance in lockstep inside one thread: they share the window registers, the per-step
// constants and the loop overhead, and give the scheduler NS independent instruction streams.
// CERT = false compiles the lookups as plain searches (the kernel shape for MPPI_FLAG_FULL_SEARCH: no test, no vote)
template <int NS, int DYN = 0, bool CERT = true, class Win, class Noise>
MPPI_HD void rollout_cost_n(const StepHeader& hd, const ArmF& A, const CostW& W,
                            const Win& win, const EndCert& cert, const RefRow* rows, const StepCtl* ctl,
                            int T, const float (&um)[NS], Noise (&noise)[NS], float (&S_out)[NS], int& hits) {
    ArmState st[NS];
    float S[NS], kS[NS], ex[NS], ey[NS], e1[NS], e2[NS];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int s = 0; s < NS; ++s) {
        arm_init(st[s], hd.q1, hd.q2, hd.d1, hd.d2);
        S[s] = 0.f; kS[s] = 0.f; ex[s] = 0.f; ey[s] = 0.f; e1[s] = 0.f; e2[s] = 0.f;
    }
    for (int t = 0; t < T; ++t) {
        const StepCtl c = ctl[t];
        float v1[NS], v2[NS], xl[NS], yl[NS];
        int j[NS];
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
        for (int s = 0; s < NS; ++s) arm_step<DYN>(st[s], A, v1[s], v2[s]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int s = 0; s < NS; ++s) {
            fk_local(st[s], A, hd.ox, hd.oy, xl[s], yl[s]);
            j[s] = CERT ? nearest_wp(win, cert, xl[s], yl[s], hits) : nearest_candidate(win, xl[s], yl[s]);
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int s = 0; s < NS; ++s) {
            const RefRow r = rows[j[s]];
            residuals(st[s], xl[s], yl[s], r, ex[s], ey[s], e1[s], e2[s]);
            float cst = wsq(W.s0, W.s1, W.s2, W.s3, ex[s], ey[s], e1[s], e2[s]);
            cst = fma_(c.g1, v1[s], fma_(c.g2, v2[s], cst));   // + gamma * u^T Sigma^-1 v  (control.py:106)
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
    for (int s = 0; s < NS; ++s)
        S_out[s] = add_(S[s], sub_(wsq(W.t0, W.t1, W.t2, W.t3, ex[s], ey[s], e1[s], e2[s]), kS[s]));
}

template <int DYN = 0, class Win, class Noise>
MPPI_HD float rollout_cost(const StepHeader& hd, const ArmF& A, const CostW& W,
                           const Win& win, const EndCert& cert, const RefRow* rows, const StepCtl* ctl,
                           int T, float um, Noise& noise, int& hits) {
    const float ums[1] = { um };
    float out[1];
    Noise (&nz)[1] = reinterpret_cast<Noise (&)[1]>(noise);
    rollout_cost_n<1, DYN>(hd, A, W, win, cert, rows, ctl, T, ums, nz, out, hits);
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

// Nominal control of horizon step t and the row vector gamma * u_t^T Sigma^-1 (control.py:106).
MPPI_HD void make_step_ctl(const double* u_t, double gamma, const double* sig_inv, StepCtl& c) {
    c.u1 = (float)u_t[0]; c.u2 = (float)u_t[1];
    c.g1 = (float)(gamma * (u_t[0] * sig_inv[0] + u_t[1] * sig_inv[2]));
    c.g2 = (float)(gamma * (u_t[0] * sig_inv[1] + u_t[1] * sig_inv[3]));
}

}  // namespace mppi

// TEST INFRASTRUCTURE ONLY — CPU emulation of the FP32 rollout arithmetic in
// mppi_robotarm_b200/csrc/mppi_math.cuh, compiled with g++ (-ffp-contract=off).  Used by CPU tests
// to study FP32-vs-FP64 parity without a GPU; never loaded by the product package.
#include "../../mppi_robotarm_b200/csrc/mppi_math.cuh"
#include <cmath>
#include <cstring>
#include <vector>
using namespace mppi;

// the tables the prepare kernel builds for the window starting at row p (serial form)
struct Tables {
    WinRegs regs; WinTable win; WinCert cert; EndWedges wed;
    RefRow rows[kWindowPad]; RefRow srows[kWindowPad]; WinEntry tab[kWindowPad]; RowRec rec[kWindowPad];
};
static void build_tables(const double* ref, int n_rows, int p, double reach, bool use_cert, Tables& tb, const CostW& W = CostW{}) {
    for (int j = 0; j < kWindowPad; ++j) make_window_row(ref, n_rows, p, j, W, tb.tab[j], tb.rows[j], tb.srows[j]);
    tb.regs.load(tb.tab);
    double lrows[kWindow][2];
    int n_valid = 0;
    for (int j = 0; j < kWindow && p + j < n_rows; ++j, ++n_valid) {
        lrows[j][0] = ref[4 * (p + j)] - ref[4 * p]; lrows[j][1] = ref[4 * (p + j) + 1] - ref[4 * p + 1];
    }
    make_win_cert(lrows, n_valid, reach, ref[4 * p], ref[4 * p + 1], use_cert, tb.cert, tb.rec, tb.wed);
    for (int j = 0; j < kWindowPad; ++j) { tb.rec[j].a = tb.tab[j].a; tb.rec[j].b = tb.tab[j].b; tb.rec[j].c = tb.tab[j].c; tb.rec[j].pad = 0.f; }
    tb.win.load(tb.cert, tb.rec, &tb.wed);
}

struct EpsArray {
    const float* e; int T;
    void operator()(int t, float& a, float& b) const { a = e[2 * t]; b = e[2 * t + 1]; }
};

static int g_ns = 1;      // samples per rollout_cost_n call (emul_set_ns)

extern "C" {

void emul_set_ns(int ns) { g_ns = ns == 2 ? 2 : 1; }

// S[K] for injected noise eps[K][T][2]; mirrors prepare + rollout kernels.  Returns new window start.
int emul_rollout_costs(const double* ref, int n_rows, int prev_idx, const double* x0, const double* u_prev,
                       int K, int T, int n_exploit, double dt, double gamma, const double* sig_inv,
                       const double* ws, const double* wt, const double* arm, double cl1, double cl2,
                       const float* eps, float* S_out, int use_cert, long long* hits_out, int dynamics_f1) {
    // waypoint update, FP64 (control.py:75, 200-232)
    double x = cl1 * cos(x0[0]) + cl2 * cos(x0[0] + x0[1]);
    double y = cl1 * sin(x0[0]) + cl2 * sin(x0[0] + x0[1]);
    int best = 0; double bd = 1e300;
    for (int j = 0; j < kWindow && prev_idx + j < n_rows; ++j) {
        double d = waypoint_d(ref, prev_idx + j, x, y);
        if (d < bd) { bd = d; best = j; }
    }
    int p = prev_idx + best;
    StepHeader hd{};
    hd.q1 = (float)x0[0]; hd.q2 = (float)x0[1]; hd.d1 = (float)x0[2]; hd.d2 = (float)x0[3];
    hd.ox = (float)ref[4 * p]; hd.oy = (float)ref[4 * p + 1]; hd.win_start = p;
    hd.a1 = angle_fix(x0[0]); hd.a12 = angle_fix(x0[0] + x0[1]);
    CostW W{ (float)(ws[0] * 1e4), (float)(ws[1] * 1e4), (float)(ws[2] * 1e4), (float)(ws[3] * 1e4),
             (float)(wt[0] * 1e4), (float)(wt[1] * 1e4), (float)(wt[2] * 1e4), (float)(wt[3] * 1e4) };
    cost_roots(W);
    Tables tb; build_tables(ref, n_rows, p, cl1 + cl2, use_cert != 0, tb, W);
    const RefRow* rows = tb.rows;
    long long hits_total = 0;
    std::vector<StepCtl> ctl(T);
    for (int t = 0; t < T; ++t) make_step_ctl(u_prev + 2 * t, gamma, sig_inv, ctl[t]);
    const double m1 = arm[0], m2 = arm[1], l1 = arm[2], l2 = arm[3], lc1 = arm[4], lc2 = arm[5], g = arm[6];
    ArmF A;
    A.A0 = (float)(m1 * lc1 * lc1 + l1 + m2 * (l1 * l1 + lc2 * lc2) + l2);
    A.A1 = (float)(2 * m2 * l1 * lc2);
    A.M22 = (float)(m2 * lc2 * lc2 + l2); A.B1 = (float)(m2 * l1 * lc2);
    A.G1a = (float)((m1 * lc1 + m2 * l1) * g); A.G1b = (float)(m2 * lc2 * g);
    A.dt = (float)dt; A.dtfix = arm_dtfix(dt); A.L1 = (float)cl1; A.L2 = (float)cl2;
    long long tri_total = 0;
    for (int k = 0; k < K; ++k) {
        LookupStats hits{0, 0};
        if (g_ns == 2 && use_cert != 2 && !dynamics_f1 && k + 1 < K) {     // two samples per "thread", like the throughput kernels
            EpsArray n2[2] = { { eps + (size_t)k * T * 2, T }, { eps + (size_t)(k + 1) * T * 2, T } };
            const float um2[2] = { k < n_exploit ? 1.f : 0.f, k + 1 < n_exploit ? 1.f : 0.f };
            float out2[2];
            rollout_cost_n<2, 0, false>(hd, A, W, tb.win, tb.cert, rows, tb.srows, ctl.data(), T, um2, n2, out2, hits);
            S_out[k] = out2[0]; S_out[k + 1] = out2[1];
            hits_total += 2 * T - hits.tri - hits.scan; tri_total += hits.tri;
            ++k;
            continue;
        }
        EpsArray n{ eps + (size_t)k * T * 2, T };
        const float um = k < n_exploit ? 1.f : 0.f;
        if (use_cert == 2)      // the kernels without the certificate: register tournament
            S_out[k] = dynamics_f1 ? rollout_cost<1>(hd, A, W, tb.regs, tb.cert, rows, tb.srows, ctl.data(), T, um, n, hits)
                                   : rollout_cost<0>(hd, A, W, tb.regs, tb.cert, rows, tb.srows, ctl.data(), T, um, n, hits);
        else
            S_out[k] = dynamics_f1 ? rollout_cost<1>(hd, A, W, tb.win, tb.cert, rows, tb.srows, ctl.data(), T, um, n, hits)
                                   : rollout_cost<0>(hd, A, W, tb.win, tb.cert, rows, tb.srows, ctl.data(), T, um, n, hits);
        hits_total += T - hits.tri - hits.scan; tri_total += hits.tri;
    }
    if (hits_out) { hits_out[0] = hits_total; hits_out[1] = tri_total; }
    return p;
}

// Soundness probe of the lookup certificate: for n queries (x', y') in the local coordinates of the window
// starting at row p, pick[i] = certified row or -1, full[i] = result of the exact FP32 search (register
// tournament), scan[i] = the in-memory form of the same search.  cert_out (optional): 64 certificate bytes,
// 32 x 32 bytes of row records, 64 bytes of end wedges (the layout of a step block).
void emul_cert_probe(const double* ref, int n_rows, int p, double reach, const float* xy, int n, int* pick, int* full,
                     int* scan, float* cert_out) {
    Tables tb; build_tables(ref, n_rows, p, reach, true, tb);
    if (cert_out) { memcpy(cert_out, &tb.cert, sizeof(tb.cert)); memcpy(cert_out + 16, tb.rec, sizeof(tb.rec)); memcpy(cert_out + 16 + 256, &tb.wed, sizeof(tb.wed)); }
    for (int i = 0; i < n; ++i) {
        pick[i] = cert_pick(tb.cert, tb.rec, tb.wed, xy[2 * i], xy[2 * i + 1]);
        full[i] = nearest_candidate(tb.regs, xy[2 * i], xy[2 * i + 1]);
        if (scan) scan[i] = nearest_scan(tb.rec, xy[2 * i], xy[2 * i + 1]);
    }
}

// The same probe for a certificate that was built elsewhere (the prepare kernel on the GPU): cert_in =
// the 64 certificate bytes + the 1024 bytes of row records + the 64 bytes of end wedges of a step block.
void emul_cert_probe_given(const double* ref, int n_rows, int p, const float* cert_in, const float* xy, int n,
                           int* pick, int* full) {
    Tables tb; build_tables(ref, n_rows, p, 2.0, false, tb);
    WinCert cert; RowRec rec[kWindowPad]; EndWedges wed;
    memcpy(&cert, cert_in, sizeof(cert)); memcpy(rec, cert_in + 16, sizeof(rec)); memcpy(&wed, cert_in + 16 + 256, sizeof(wed));
    for (int i = 0; i < n; ++i) {
        pick[i] = cert_pick(cert, rec, wed, xy[2 * i], xy[2 * i + 1]);
        full[i] = nearest_candidate(tb.regs, xy[2 * i], xy[2 * i + 1]);
    }
}

// The optimal-trajectory rollout of the final stage (control.py:129-134): x <- F(x, u[t-1]) with the t = 0 wrap,
// one thread, the latency form of the step (arm_step_serial).  out[T][4] = (q1, q2, dq1, dq2) as the kernel stores
// them (value - compensation in FP64).
void emul_optimal_traj(const double* x0, const double* u, int T, double dt, const double* arm, double cl1, double cl2,
                       int dynamics_f1, double* out) {
    const double m1 = arm[0], m2 = arm[1], l1 = arm[2], l2 = arm[3], lc1 = arm[4], lc2 = arm[5], g = arm[6];
    ArmF A;
    A.A0 = (float)(m1 * lc1 * lc1 + l1 + m2 * (l1 * l1 + lc2 * lc2) + l2);
    A.A1 = (float)(2 * m2 * l1 * lc2);
    A.M22 = (float)(m2 * lc2 * lc2 + l2); A.B1 = (float)(m2 * l1 * lc2);
    A.G1a = (float)((m1 * lc1 + m2 * l1) * g); A.G1b = (float)(m2 * lc2 * g);
    A.dt = (float)dt; A.dtfix = arm_dtfix(dt); A.L1 = (float)cl1; A.L2 = (float)cl2;
    ArmState st;
    arm_init(st, (float)x0[0], (float)x0[1], (float)x0[2], (float)x0[3], angle_fix(x0[0]), angle_fix(x0[0] + x0[1]));
    for (int t = 0; t < T; ++t) {
        const int tc = t == 0 ? T - 1 : t - 1;
        if (dynamics_f1) arm_step_serial<1>(st, A, (float)u[2 * tc], (float)u[2 * tc + 1]);
        else arm_step_serial<0>(st, A, (float)u[2 * tc], (float)u[2 * tc + 1]);
        out[4 * t + 0] = (double)st.q1 - (double)st.kq1; out[4 * t + 1] = (double)st.q2 - (double)st.kq2;
        out[4 * t + 2] = (double)st.d1 - (double)st.kd1; out[4 * t + 3] = (double)st.d2 - (double)st.kd2;
    }
}

void emul_sincos(const float* x, int n, float* s, float* c) { for (int i = 0; i < n; ++i) sincos_(x[i], s[i], c[i]); }
// sin / cos as the rollouts take them: of the fixed-point image of an FP64 angle
void emul_sincos_fix(const double* x, int n, float* s, float* c) { for (int i = 0; i < n; ++i) sincos_fix(angle_fix(x[i]), s[i], c[i]); }

void emul_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
    U4 r = philox4x32_10(U4{c0, c1, c2, c3}, philox_expand_key(k0, k1)); out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

void emul_noise(uint32_t seed_lo, uint32_t seed_hi, uint32_t step, const double* chol, uint32_t env,
                uint32_t k0, int K, int T, float* eps) {
    NoiseCfg nc{ philox_expand_key(seed_lo, seed_hi), step, (float)chol[0], (float)chol[2], (float)chol[3] };
    for (int k = 0; k < K; ++k)
        for (int p = 0; 2 * p < T; ++p) {
            float a, b, c, d; noise_pair(nc, env, k0 + k, p, a, b, c, d);
            float* e = eps + ((size_t)k * T + 2 * p) * 2;
            e[0] = a; e[1] = b; if (2 * p + 1 < T) { e[2] = c; e[3] = d; }
        }
}
}

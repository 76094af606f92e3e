/* TEST INFRASTRUCTURE ONLY — plain-C FP64 restatement of the reference's rollout + weighting
 * (the part of the MPPI step that is K-sized), used as the multi-threaded CPU baseline of
 * bench.py (`--impl reference`, `cpu_baseline`) and cross-checked against oracle/mppi_oracle.py
 * in tests/test_oracle.py.  Never linked into or loaded by the product package.
 *
 * Restates, in /root/reference:
 *   control.py:91-109   K x T rollout loop with stage, control and terminal cost
 *   control.py:174-198  tracking cost;  control.py:200-215 nearest waypoint (first arg-min of 30)
 *   control.py:234-263  arm dynamics + semi-implicit Euler
 *   control.py:297-314  soft-min weights;  control.py:115-118 weighted noise sum
 * OpenMP over the K samples is the only liberty taken (the reference is a single Python thread).
 */
#include <math.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define WINDOW 30

typedef struct {
    double dt, lambda, gamma;
    double sig_inv[4];
    double ws[4], wt[4];
    double m1, m2, l1, l2, lc1, lc2, g;
    double cl1, cl2;
    double jl_lo[2], jl_hi[2], jl_w;   /* joint-limit stage cost (extension, weight 0 = the reference) */
} oracle_cfg;

static void arm_step(const oracle_cfg* c, double* q1, double* q2, double* d1, double* d2, double v1, double v2) {
    double c2 = cos(*q2);
    double M11 = c->m1 * c->lc1 * c->lc1 + c->l1 + c->m2 * (c->l1 * c->l1 + c->lc2 * c->lc2 + 2 * c->l1 * c->lc2 * c2) + c->l2;
    double M22 = c->m2 * c->lc2 * c->lc2 + c->l2;
    double M12 = c->m2 * c->l1 * c->lc2 * c2 + c->m2 * c->lc2 * c->lc2 + c->l2;
    double h = c->m2 * c->l1 * c->lc2 * sin(*q2);
    double g1 = c->m1 * c->lc1 * c->g * cos(*q1) + c->m2 * c->g * (c->lc2 * cos(*q1 + *q2) + c->l1 * cos(*q1));
    double g2 = c->m2 * c->lc2 * c->g * cos(*q1 + *q2);
    double cd1 = (-h * *d2) * *d1 + (-h * *d1 - h * *d2) * *d2;
    double cd2 = (h * *d1) * *d1;
    double b1 = v1 - cd1 - g1, b2 = v2 - cd2 - g2;
    double det = M11 * M22 - M12 * M12;
    *d1 += (M22 * b1 - M12 * b2) / det * c->dt;
    *d2 += (M11 * b2 - M12 * b1) / det * c->dt;
    *q1 += *d1 * c->dt;
    *q2 += *d2 * c->dt;
}

static double track_cost(const oracle_cfg* c, const double* win, int n_win, const double* w,
                         double q1, double q2, double d1, double d2) {
    double x = c->cl1 * cos(q1) + c->cl2 * cos(q1 + q2);
    double y = c->cl1 * sin(q1) + c->cl2 * sin(q1 + q2);
    int best = 0; double bd = INFINITY;
    for (int j = 0; j < n_win; ++j) {
        double dx = x - win[4 * j], dy = y - win[4 * j + 1];
        double d = (dx * dx + dy * dy) * 100;
        if (d < bd) { bd = d; best = j; }
    }
    const double* r = win + 4 * best;
    double cost = w[0] * (x - r[0]) * (x - r[0]) + w[1] * (y - r[1]) * (y - r[1])
                + w[2] * (d1 - r[2]) * (d1 - r[2]) + w[3] * (d2 - r[3]) * (d2 - r[3]);
    return cost * 10000;
}

static double joint_limit_cost(const oracle_cfg* c, double q1, double q2) {
    if (c->jl_w == 0.0) return 0.0;
    double v1 = fmax(fmax(q1 - c->jl_hi[0], c->jl_lo[0] - q1), 0.0);
    double v2 = fmax(fmax(q2 - c->jl_hi[1], c->jl_lo[1] - q2), 0.0);
    return c->jl_w * (v1 * v1 + v2 * v2) * 10000;
}

/* S[k] for k in [0, K): eps is [K][T][2] (double), u is [T][2], window = ref rows [p, p+n_win) */
void oracle_rollout_costs(const oracle_cfg* c, const double* win, int n_win, const double* x0,
                          const double* u, const double* eps, int K, int T, int n_exploit, double* S) {
#pragma omp parallel for schedule(static)
    for (int k = 0; k < K; ++k) {
        double q1 = x0[0], q2 = x0[1], d1 = x0[2], d2 = x0[3], s = 0.0;
        const double* e = eps + (size_t)k * T * 2;
        for (int t = 0; t < T; ++t) {
            double v1 = e[2 * t], v2 = e[2 * t + 1];
            if (k < n_exploit) { v1 += u[2 * t]; v2 += u[2 * t + 1]; }
            arm_step(c, &q1, &q2, &d1, &d2, v1, v2);
            double ui0 = u[2 * t] * c->sig_inv[0] + u[2 * t + 1] * c->sig_inv[2];
            double ui1 = u[2 * t] * c->sig_inv[1] + u[2 * t + 1] * c->sig_inv[3];
            s += track_cost(c, win, n_win, c->ws, q1, q2, d1, d2) + c->gamma * (ui0 * v1 + ui1 * v2)
               + joint_limit_cost(c, q1, q2);
        }
        s += track_cost(c, win, n_win, c->wt, q1, q2, d1, d2);
        S[k] = s;
    }
}

/* weights (control.py:297-314) and weighted noise sum (control.py:115-118); returns rho */
double oracle_weighted_sum(const double* S, const double* eps, int K, int T, double lambda, double* w, double* w_eps) {
    double rho = S[0];
    for (int k = 1; k < K; ++k) if (S[k] < rho) rho = S[k];
    double eta = 0.0;
    for (int k = 0; k < K; ++k) { w[k] = exp((-1.0 / lambda) * (S[k] - rho)); eta += w[k]; }
    for (int i = 0; i < 2 * T; ++i) w_eps[i] = 0.0;
    for (int k = 0; k < K; ++k) {
        w[k] /= eta;
        const double* e = eps + (size_t)k * T * 2;
        for (int i = 0; i < 2 * T; ++i) w_eps[i] += w[k] * e[i];
    }
    return rho;
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// mppi_kernels.cuh — sm_100a kernels of the MPPI step (included by mppi_cabi.cu only).
//
//   mppi_window_table_sm100a  window tables + lookup certificate of every window start (once per path)   control.py:203-215
//   mppi_prepare_sm100a    waypoint update in FP64, header, window copy, step controls   control.py:75, 200-232
//   mppi_rollout_sm100a    K fused rollouts, costs only                       control.py:84-109
//   mppi_softmin_sm100a    min / exp / partial normaliser                     control.py:297-314
//   mppi_wsum_injected_sm100a   weighted noise sum, K x (T*2) reduction       control.py:115-118
//   mppi_softmin_wsum_philox_sm100a  the two above + reduce (+ exchange, combine, filter, update) fused (Philox)
//                                                                              control.py:297-314, 115-134
//   mppi_reduce_sm100a     this GPU's partial (rho_g, eta_g, V_g)             (sharding)
//   mppi_finalize_sm100a   combine, median filter, update, optimal rollout    control.py:122-134
//   mppi_plant_sm100a      one tick of the device-resident closed loop        run.py:53-59, utils.py:14-29
//   mppi_sampled_traj_sm100a  trajectories of all samples                     control.py:137-145
//   mppi_philox_export_sm100a the noise tensor the kernels draw               control.py:154-164
//
// Data layout in HBM (per environment e): step block = 64 B header | 32 x WinEntry | 32 x RefRow | 16 x pair |
// WinCert | 32 x RowRec | EndWedges | T x StepCtl (contiguous, 16-byte aligned, moved into shared memory with ONE
// 1-D TMA bulk copy);
// costs S[e][K_local] and weights w[e][K_local] float32; injected noise eps[e][K_local][T][2]
// float32 (the reference's own [K,T,2] layout, control.py:84).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "mppi_math.cuh"

namespace mppi {

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     " selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

// Programmatic dependent launch (PTX griddepcontrol): `wait` blocks until the kernel this one depends on has
// completed and its memory is visible (a no-op for a normal launch); `launch_dependents` lets the next kernel's
// CTAs be scheduled early (they still wait before touching our results).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

// Order-preserving map float -> uint32 (radix-sort key): the rollout CTAs fold their minima into ONE word per
// environment with atomicMin (min is order-independent, so the result is deterministic); 0xffffffff = "no finite cost".
__device__ __forceinline__ unsigned int min_key(float v) {
    const unsigned int b = __float_as_uint(v);
    return b ^ ((unsigned int)((int)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float min_key_decode(unsigned int k) {
    if (k == 0xffffffffu) return INFINITY;
    return __uint_as_float(k ^ (((k >> 31) - 1u) | 0x80000000u));
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ bool finite_(float x) { return fabsf(x) <= 3.4028234e38f; }   // false for NaN/Inf

// ------------------------------------------------------------------------------------------------
// constants of one controller (kernel parameter, lives in the constant bank)
// ------------------------------------------------------------------------------------------------
struct DevCfg {
    ArmF arm;
    CostW cost;
    NoiseCfg noise;            // .step is filled per launch from the input block
    int K_local, K_total, k_offset, T, n_env, n_exploit, n_ref_rows, flags;
    int step_block_bytes;      // kStepBlockFixed + 16*T
    int g_roll, g_soft, g_wsum;   // blocks per environment of the three K-sized kernels
    double gamma, lambda, inv_lambda;
    double sig_inv[4];
    double cost_l1, cost_l2;
    double arm64[7];           // m1, m2, l1, l2, lc1, lc2, g for the FP64 plant of the device closed loop
};

// parameters of one mppi_closed_loop() call, kept in device memory so that the captured per-tick
// graph is identical for every call
struct LoopParams {
    double plant_dt;
    double* log;               // [n_steps][n_env][8]
    int32_t* stop;             // [n_env] first tick at which the end of the path was reached
    int32_t tick, n_steps;
};

struct StepBlockView {          // pointers into one environment's step block (global or shared)
    StepHeader* hd; WinEntry* win; RefRow* rows; float4* pairs; WinCert* cert; RowRec* rec; EndWedges* wed; RefRow* srows; StepCtl* ctl;
};
// header + win + rows + pairs + certificate + row records + end wedges + pre-scaled rows of the stage cost
constexpr int kStepBlockFixed = 64 + 32 * kWindowPad + 16 * (kWindowPad / 2) + 64 + 32 * kWindowPad + 64 + 16 * kWindowPad;
__host__ __device__ __forceinline__ StepBlockView view_step_block(void* base) {
    char* b = (char*)base;
    StepBlockView v;
    v.hd = (StepHeader*)b;
    v.win = (WinEntry*)(b + 64);
    v.rows = (RefRow*)(b + 64 + 16 * kWindowPad);
    v.pairs = (float4*)(b + 64 + 32 * kWindowPad);          // (a_2i, a_2i+1, b_2i, b_2i+1): operands of packed FFMA2
    v.cert = (WinCert*)(b + 64 + 32 * kWindowPad + 16 * (kWindowPad / 2));   // lookup certificate of the window
    v.rec = (RowRec*)(b + 64 + 32 * kWindowPad + 16 * (kWindowPad / 2) + 64);   // (a, b, c) + certificate of each row
    v.wed = (EndWedges*)(b + 64 + 32 * kWindowPad + 16 * (kWindowPad / 2) + 64 + 32 * kWindowPad);   // far-field wedges of the end rows
    v.srows = (RefRow*)(b + kStepBlockFixed - 16 * kWindowPad);   // -sqrt(weight) * waypoint: stage_cost() of mppi_math.cuh
    v.ctl = (StepCtl*)(b + kStepBlockFixed);
    return v;
}

// ------------------------------------------------------------------------------------------------
// Peer-memory exchange of the sharded step (replaces the NCCL all-gather): every rank owns an
// exchange buffer that all ranks of the node have mapped over NVLink.  Layout of one buffer:
//   double slot[2][world][n_env][2 + 2T]   partial triples, double-buffered by step parity
//   uint64 flag[2][world][n_env]           sequence number of the step whose triple is in the slot
// The last block of the weight-sum kernel PUTS this rank's triple into slot[parity][rank] of every
// peer (plain stores to peer addresses), fences system-wide and raises the flags; the finalize
// kernel of each rank waits for its `world` flags and then reads only local memory.  Double
// buffering is enough: a rank can only be one step ahead of a peer, because its finalize waits
// for that peer's flag of the same step.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 16;
struct PeerExchange {
    int rank, world;              // world == 0: exchange disabled
    char* buf[kMaxPeers];         // exchange buffer of rank r as mapped on THIS GPU (buf[rank] is local)
    size_t slot_bytes;            // world * n_env * (2 + 2T) * 8
    size_t flags_off;             // 2 * slot_bytes rounded up to 256
    const unsigned long long* seq;   // step sequence number of this handle (device memory, set by prepare)
    unsigned long long timeout_ns;   // how long the combine waits for a peer's partial (mppi_set_exchange_timeout)
};
__device__ __forceinline__ double* px_slot(const PeerExchange& px, int on_rank, int parity, int from_rank, int n_env,
                                           int e, int n) {
    return (double*)(px.buf[on_rank] + (size_t)parity * px.slot_bytes) + ((size_t)from_rank * n_env + e) * n;
}
__device__ __forceinline__ unsigned long long* px_flag(const PeerExchange& px, int on_rank, int parity, int from_rank,
                                                       int n_env, int e) {
    return (unsigned long long*)(px.buf[on_rank] + px.flags_off) + ((size_t)parity * px.world + from_rank) * n_env + e;
}
// called by all threads of the block that just wrote `triple` (this rank's partial of environment e)
__device__ __forceinline__ void px_put(const PeerExchange& px, const double* triple, int n_env, int e, int n) {
    __syncthreads();                                            // the triple is complete
    const unsigned long long seq = *px.seq;
    const int par = (int)(seq & 1ull);
    for (int r = 0; r < px.world; ++r) {
        double* dst = px_slot(px, r, par, px.rank, n_env, e, n);
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldcg(triple + i);
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < px.world) {
        volatile unsigned long long* f = px_flag(px, threadIdx.x, par, px.rank, n_env, e);
        *f = seq;
    }
}

// input block on the device (mirror of the caller's pinned block, see MppiIoLayout)
struct DevIo {
    const double* x0;        // [n_env][4]
    const double* u_prev;    // [n_env][T][2]
    const int32_t* prev_idx; // [n_env]
    const uint64_t* step;    // [1]
    int32_t* new_idx;        // [n_env]
    int32_t* status;         // [n_env] bit 0: the peer exchange timed out (the update was skipped)
    double* rho; double* eta;            // [n_env]
    double* u0;              // [n_env][2] first row of the shifted sequence (what calc_control_input returns)
    double* w_eps_raw; double* w_eps_filt; double* u_new;   // [n_env][T][2]
    double* opt_traj;                    // [n_env][T][4]
    // zero-copy mode: the caller's pinned block is read / written directly by the kernels
    const char* host_in;     // device-visible address of the pinned block's inputs, or null
    ptrdiff_t in_delta;      // (device input mirror) - (pinned block): same field offsets in both
    ptrdiff_t out_delta;     // (pinned block outputs) - (device output mirror), 0 = do not mirror to the host
};

// store an output to the device mirror and, in zero-copy mode, to the same field of the pinned block
template <class TT>
__device__ __forceinline__ void out_store(const DevIo& io, TT* dev_ptr, TT v) {
    *dev_ptr = v;
    if (io.out_delta != 0) *(TT*)((char*)dev_ptr + io.out_delta) = v;
}

__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// make_wedge() of mppi_math.cuh for BOTH targets at once (w = 0: last row, w = 1: row 0; the two
// dependency chains interleave), one lane per window row: (rx, ry) = this lane's local row, valid for
// lane < n, n >= 2.  Every lane returns the same coefficients.
__device__ __forceinline__ void warp_wedges(double rx, double ry, int lane, int n, double margin, double dom, EndWedges& c) {
    const int target[2] = { n - 1, 0 };
    double tx[2], ty[2], n0x[2], n0y[2], gx[2], gy[2], sl[2];
    bool ok[2], mine[2];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        tx[w] = __shfl_sync(0xffffffffu, rx, target[w]); ty[w] = __shfl_sync(0xffffffffu, ry, target[w]);
    }
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        n0x[w] = tx[w] - tx[1 - w]; n0y[w] = ty[w] - ty[1 - w];          // target minus the other end (not normalised:
        mine[w] = lane < n && lane != target[w];                         //  the slopes below are ratios)
        gx[w] = tx[w] - rx; gy[w] = ty[w] - ry;
        const double along = gx[w] * n0x[w] + gy[w] * n0y[w], across = n0x[w] * gy[w] - n0y[w] * gx[w];
        const bool bad = mine[w] && (!(along > 0.05 * fabs(across)) || !(along > 0.0));
        ok[w] = !__any_sync(0xffffffffu, bad) && (n0x[w] * n0x[w] + n0y[w] * n0y[w] > 0.0);
        sl[w] = mine[w] && along > 0.0 ? across * rcp64_(along) : 0.0;
    }
    double smin[2], smax[2];
#pragma unroll
    for (int w = 0; w < 2; ++w) { smin[w] = mine[w] ? sl[w] : 1e300; smax[w] = mine[w] ? sl[w] : -1e300; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            smin[w] = fmin(smin[w], __shfl_xor_sync(0xffffffffu, smin[w], o));
            smax[w] = fmax(smax[w], __shfl_xor_sync(0xffffffffu, smax[w], o));
        }
    }
    double mx[2][2], my[2][2], bx[2], by[2], tau[2];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        smin[w] -= 2e-7 * (1.0 + smin[w] * smin[w]); smax[w] += 2e-7 * (1.0 + smax[w] * smax[w]);   // widen the cone
        const double s2[2] = { smin[w], smax[w] };
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const double vx = n0x[w] - s2[i] * n0y[w], vy = n0y[w] + s2[i] * n0x[w];
            const double inv = rsqrt64_(vx * vx + vy * vy);
            mx[w][i] = vx * inv; my[w][i] = vy * inv;
        }
        bx[w] = mx[w][0] + mx[w][1]; by[w] = my[w][0] + my[w][1];       // bisector, not normalised: z = r_t + tau * b
        const double bn2 = bx[w] * bx[w] + by[w] * by[w];
        ok[w] = ok[w] && bn2 > 1e-6;
        const double g2 = gx[w] * gx[w] + gy[w] * gy[w], ng = bx[w] * gx[w] + by[w] * gy[w];
        ok[w] = ok[w] && !__any_sync(0xffffffffu, mine[w] && !(ng > 0.0));
        tau[w] = mine[w] && ng > 0.0 ? fmax(0.5 * (margin - g2) * rcp64_(ng), 0.0) * 1.000001 : 0.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int w = 0; w < 2; ++w) tau[w] = fmax(tau[w], __shfl_xor_sync(0xffffffffu, tau[w], o));
    }
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        const double bn2 = bx[w] * bx[w] + by[w] * by[w];
        ok[w] = ok[w] && (tau[w] * tau[w] * bn2 <= kCertMaxTau * kCertMaxTau);
    }
    if (ok[0]) wedge_finish(tx[0] + tau[0] * bx[0], ty[0] + tau[0] * by[0], mx[0], my[0], dom, c.lx, c.ly, c.lk);
    if (ok[1]) wedge_finish(tx[1] + tau[1] * bx[1], ty[1] + tau[1] * by[1], mx[1], my[1], dom, c.fx, c.fy, c.fk);
}

// Lookup certificate of one window (make_win_cert of mppi_math.cuh), one lane per row: srow = the local rows
// in shared memory (valid for j < n), lane `a` derives row a's tangent and the bounds its two roles put on
// the lateral range; the range is their intersection over the warp.  Every lane returns the same certificate
// and its own row record (tx, ty, kL, kU).
__device__ __forceinline__ void warp_win_cert(const double (*srow)[2], int lane, int n, double reach, double ox,
                                              double oy, bool enabled, WinCert& c, RowRec& rec, EndWedges& wed) {
    cert_disable(c, n);
    cert_row_disable(rec);
    wedge_disable(wed.lx, wed.ly, wed.lk); wedge_disable(wed.fx, wed.fy, wed.fk);
    wed.pad[0] = wed.pad[1] = wed.pad[2] = wed.pad[3] = 0.f;
    if (!enabled || n < 1) return;
    const double dom = 1.01 * reach + fmax(fabs(ox), fabs(oy)) + 0.01, domw = 1.0001 * dom;
    c.dom = (float)dom;
    if (n == 1) {
        if (lane == 0) { rec.kL = kCertHuge; rec.kU = -kCertHuge; }
        c.blo = -kCertHuge; c.bhi = kCertHuge;
        return;
    }
    double chx = srow[n - 1][0] - srow[0][0], chy = srow[n - 1][1] - srow[0][1];
    const double ch2 = chx * chx + chy * chy;
    if (!(ch2 > 0.0)) return;
    const double ich = rsqrt64_(ch2);
    chx *= ich; chy *= ich;
    const double nux = (double)(float)(-chy), nuy = (double)(float)chx;
    const bool valid = lane < n;
    const double rx = valid ? srow[lane][0] : 0.0, ry = valid ? srow[lane][1] : 0.0;
    const double cmax = warp_max_d(rx * rx + ry * ry);
    const double ab = 2.0000001 * cmax * rsqrt64_(fmax(cmax, 1e-300));       // |a_j|, |b_j| <= 2 sqrt(cmax)
    const double margin = cert_margin(ab, ab, cmax, domw);
    warp_wedges(rx, ry, lane, n, margin, domw, wed);
    const double wmin = kCertMinLateral * reach, wt = kCertOffsetLateral * reach;
    const double bmax = (fabs(nux) + fabs(nuy)) * domw;
    double lo = -bmax, hi = bmax;
    {
        auto row = [&](int j, double& x, double& y) { x = srow[j][0]; y = srow[j][1]; };
        RowGeom g, gp;
        bool again = false;
        if (valid) {
            g = cert_row_geom<false>(row, lane, n, nux, nuy, margin, wt);
            gp = g;
            again = (lane > 0 && !cert_role_plain(g.L, g.b, wmin)) || (lane < n - 1 && !cert_role_plain(g.U, g.b, wmin));
        }
        // rows closer together than the FP32 margin (a path that starts from rest): thresholds pushed off the row
        if (__any_sync(0xffffffffu, again) && again) gp = cert_row_geom<true>(row, lane, n, nux, nuy, margin, wt);
        if (valid) {
            rec.tx = (float)g.tx; rec.ty = (float)g.ty;
            double push, rlo, rhi;
            if (lane == 0) rec.kL = kCertHuge;
            else if (cert_role_form(g.L, gp.L, g.b, wmin, wt, push, rlo, rhi)) {
                rec.kL = (float)(g.k - push - cert_delta(g.tx, g.ty, g.k - push, domw)); lo = fmax(lo, rlo); hi = fmin(hi, rhi);
            }
            if (lane == n - 1) rec.kU = -kCertHuge;
            else if (cert_role_form(g.U, gp.U, g.b, wmin, wt, push, rlo, rhi)) {
                rec.kU = (float)(g.k + push + cert_delta(g.tx, g.ty, g.k + push, domw)); lo = fmax(lo, rlo); hi = fmin(hi, rhi);
            }
        }
    }
    // ---- index estimate: Kasa circle fit on centred chord coordinates, quadratic fit of the index on w ----
    const double sj = chx * rx + chy * ry, bj = nux * rx + nuy * ry;
    double ms = valid ? sj : 0.0, mb = valid ? bj : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmax(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmin(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        ms += __shfl_xor_sync(0xffffffffu, ms, o); mb += __shfl_xor_sync(0xffffffffu, mb, o);
    }
    cert_store_range(c, nux, nuy, lo, hi, bmax);
    if (n < 3) return;
    const double inv_n = rcp64_((double)n);
    ms *= inv_n; mb *= inv_n;
    const double u = valid ? sj - ms : 0.0, v = valid ? bj - mb : 0.0, z = u * u + v * v;
    double suu = u * u, svv = v * v, suv = u * v, suz = u * z, svz = v * z;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        suu += __shfl_xor_sync(0xffffffffu, suu, o); svv += __shfl_xor_sync(0xffffffffu, svv, o);
        suv += __shfl_xor_sync(0xffffffffu, suv, o); suz += __shfl_xor_sync(0xffffffffu, suz, o);
        svz += __shfl_xor_sync(0xffffffffu, svz, o);
    }
    const double det = suu * svv - suv * suv;
    double sc = ms, bc = mb + 1.0e6;                                   // straight window: centre far away
    if (fabs(det) > 1e-12 * suu * suu) {
        const double idet = rcp64_(det);
        const double uc = 0.5 * (suz * svv - svz * suv) * idet, vc = 0.5 * (svz * suu - suz * suv) * idet;
        if (fabs(vc) > 0.0 && fabs(vc) < 1.0e6 && fabs(uc) < 1.0e6) { sc = ms + uc; bc = mb + vc; }
    }
    const double w = valid ? (sj - sc) * rcp64_(bc - bj) : 0.0;
    const double wmax = warp_max_d(fabs(w));
    if (!(wmax > 0.0) || !(wmax < 1e30)) return;
    const double iw = rcp64_(wmax), W = w * iw, jd = (double)lane;
    double m0 = valid ? 1.0 : 0.0, m1 = m0 * W, m2 = m1 * W, m3 = m2 * W, m4 = m3 * W;
    double r0 = m0 * jd, r1 = m1 * jd, r2 = m2 * jd;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m0 += __shfl_xor_sync(0xffffffffu, m0, o); m1 += __shfl_xor_sync(0xffffffffu, m1, o);
        m2 += __shfl_xor_sync(0xffffffffu, m2, o); m3 += __shfl_xor_sync(0xffffffffu, m3, o);
        m4 += __shfl_xor_sync(0xffffffffu, m4, o); r0 += __shfl_xor_sync(0xffffffffu, r0, o);
        r1 += __shfl_xor_sync(0xffffffffu, r1, o); r2 += __shfl_xor_sync(0xffffffffu, r2, o);
    }
    const double D = m0 * (m2 * m4 - m3 * m3) - m1 * (m1 * m4 - m3 * m2) + m2 * (m1 * m3 - m2 * m2);
    if (!(fabs(D) > 1e-300)) return;
    const double iD = rcp64_(D);
    const double q0 = (r0 * (m2 * m4 - m3 * m3) - m1 * (r1 * m4 - m3 * r2) + m2 * (r1 * m3 - m2 * r2)) * iD;
    const double q1 = (m0 * (r1 * m4 - r2 * m3) - r0 * (m1 * m4 - m3 * m2) + m2 * (m1 * r2 - m2 * r1)) * iD * iw;
    const double q2 = (m0 * (m2 * r2 - m3 * r1) - m1 * (m1 * r2 - m2 * r1) + r0 * (m1 * m3 - m2 * m2)) * iD * iw * iw;
    c.sx = (float)chx; c.sy = (float)chy; c.s0 = (float)(-sc); c.bc = (float)bc;
    c.c0 = (float)q0; c.c1 = (float)q1; c.c2 = (float)q2;
    c.jhi = (fabs(q0) < 1e6 && fabs(q1) < 1e30 && fabs(q2) < 1e30) ? (float)(n - 2) : 0.f;
}

// Bytes of a step block that depend only on the path and the window start p (everything between the header
// and the step controls): built once per path for every p (mppi_window_table_sm100a) and copied by the prepare
// kernel, or built on the spot when the path is too long for a table.
constexpr int kWinBytes = kStepBlockFixed - 64;
static_assert(kWinBytes % 16 == 0, "window part is copied in 16-byte words");

// One warp: `row` = this lane's path row p + lane (x, y, dq1, dq2; anything for lanes beyond the path end),
// (ox, oy) = row p.  Writes the window tables, the lookup certificate and the far-field wedges through sb.
__device__ __forceinline__ void write_window_tables(const double4& row, double ox, double oy, int lane, int p, int n,
                                                    const DevCfg& cfg, double (*srow)[2], const StepBlockView& sb) {
    // window row `lane` in coordinates local to row p (make_window_row on the preloaded rows)
    WinEntry w; RefRow r;
    if (lane < kWindow && p + lane < n) {
        const double rx = row.x - ox, ry = row.y - oy;
        w.a = (float)(-2.0 * rx); w.b = (float)(-2.0 * ry); w.c = (float)(rx * rx + ry * ry); w.pad = 0.f;
        r.rx = (float)rx; r.ry = (float)ry; r.rd1 = (float)row.z; r.rd2 = (float)row.w;
    } else {                   // beyond the end of the path (control.py:208-209) or table padding
        w.a = 0.f; w.b = 0.f; w.c = kSentinel; w.pad = 0.f;
        r.rx = 0.f; r.ry = 0.f; r.rd1 = 0.f; r.rd2 = 0.f;
    }
    // the rollouts subtract the FP32 origin from the FP32 end-effector; rows are relative to
    // the FP64 origin — the difference (<= 6e-8) is common to all candidates of a lookup
    sb.win[lane] = w; sb.rows[lane] = r;
    sb.srows[lane] = (lane < kWindow && p + lane < n) ? stage_row(cfg.cost, row.x - ox, row.y - oy, row.z, row.w) : r;
    {   // lookup certificate of the window (make_win_cert of mppi_math.cuh, one lane per row)
        const int nv = min(kWindow, n - p);
        if (lane < kWindow) { srow[lane][0] = row.x - ox; srow[lane][1] = row.y - oy; }
        __syncwarp();
        WinCert c; RowRec rec; EndWedges wed;
        warp_win_cert(srow, lane, nv, cfg.cost_l1 + cfg.cost_l2, ox, oy, !(cfg.flags & 16), c, rec, wed);   // 16: MPPI_FLAG_FULL_SEARCH
        rec.a = w.a; rec.b = w.b; rec.c = w.c; rec.pad = 0.f;
        sb.rec[lane] = rec;
        if (lane == 0) { *sb.cert = c; *sb.wed = wed; }
    }
    // the same a/b coefficients once more, laid out as candidate pairs
    const float a1 = __shfl_down_sync(0xffffffffu, w.a, 1), b1 = __shfl_down_sync(0xffffffffu, w.b, 1);
    if ((lane & 1) == 0) sb.pairs[lane >> 1] = make_float4(w.a, a1, w.b, b1);
}

// Window part of the step block for EVERY window start p of the path (one warp per p), run by
// mppi_set_ref_path(): table[p] = bytes [64, kStepBlockFixed) of the step block of a controller at waypoint p.
__global__ void __launch_bounds__(32) mppi_window_table_sm100a(DevCfg cfg, const double* __restrict__ ref, int n,
                                                               char* __restrict__ table) {
    __shared__ double srow[kWindowPad][2];
    const int p = blockIdx.x, lane = threadIdx.x;
    const double4* ref4 = (const double4*)ref;
    const double4 row = ref4[min(p + lane, n - 1)];
    const double ox = __shfl_sync(0xffffffffu, row.x, 0), oy = __shfl_sync(0xffffffffu, row.y, 0);
    const StepBlockView sb = view_step_block(table + (size_t)p * kWinBytes - 64);    // (header and controls not touched)
    write_window_tables(row, ox, oy, lane, p, n, cfg, srow, sb);
}

// ================================================================================================
// 1. prepare: one warp per environment
// ================================================================================================
// One warp per environment.  The kernel is a chain of memory latencies, so every load is issued as
// early as its address is known: (1) all inputs at once — from the caller's pinned block over PCIe
// (zero-copy) or from the device mirror; (2) the 60 reference rows the new window can touch, while the
// FP64 forward kinematics run; the rows of the final window are then picked by shuffles, not reloaded.
// pull: which inputs are read from the caller's pinned block (zero-copy): bit 0 the observed state, bit 1 the
// controller state (sequence, waypoint index, step counter).  MPPI_FLAG_RESIDENT_STATE (128) keeps the controller
// state on the device between steps: the step counter then advances here.
__global__ void __launch_bounds__(32) mppi_prepare_sm100a(DevCfg cfg, DevIo io, const double* __restrict__ ref,
                                                          char* __restrict__ step_blocks, int pull,
                                                          unsigned long long* __restrict__ seq,
                                                          const char* __restrict__ win_table) {
    __shared__ double srow[kWindowPad][2];        // local window rows for the certificate construction
    pdl_launch_dependents();                      // the rollout CTAs may be scheduled now (they wait before reading)
    const int e = blockIdx.x, lane = threadIdx.x, T = cfg.T;
    if (e == 0 && lane == 0) *seq += 1ull;        // step sequence number, read by every later kernel of the step
    if (lane == 0) ((unsigned int*)(seq + 2))[e] = 0xffffffffu;   // this step's minimum cost: no finite cost seen yet
    const bool zx = io.host_in != nullptr && (pull & 1), zc = io.host_in != nullptr && (pull & 2);
    const ptrdiff_t back = zc ? io.in_delta : 0;  // zero-copy: read the pinned block instead of the mirror
    const double* sx = (const double*)((const char*)(io.x0 + 4 * e) - (zx ? io.in_delta : 0));
    const double2* su = (const double2*)((const char*)(io.u_prev + (size_t)e * T * 2) - back);
    const int32_t* sp = (const int32_t*)((const char*)(io.prev_idx + e) - back);
    if ((cfg.flags & 128) && e == 0 && lane == 0) *const_cast<uint64_t*>(io.step) += 1;   // resident state: next control step
    // ---- (1) inputs --------------------------------------------------------------------------------
    constexpr int kTS = (MPPI_MAX_T_INTERNAL + 31) / 32;
    double2 uu[kTS];
#pragma unroll
    for (int i = 0; i < kTS; ++i) { const int t = lane + 32 * i; uu[i] = t < T ? su[t] : make_double2(0.0, 0.0); }
    const double q1 = sx[0], q2 = sx[1], dq1 = sx[2], dq2 = sx[3];
    int p = *sp;
    const uint64_t stp = (zc && lane == 0 && e == 0) ? *(const uint64_t*)((const char*)io.step - back) : 0;
    const int n = cfg.n_ref_rows;
    p = max(0, min(p, n - 1));
    if (win_table != nullptr) {
        // the new window starts at p .. p+29, almost always within a few rows of p: pull the tables of p .. p+3 into
        // L2 now (76 lines of 128 B), so that the copy below — at the end of the kernel's latency chain — hits
        for (int ln = lane; ln < (4 * kWinBytes + 127) / 128; ln += 32) {
            const char* a = win_table + (size_t)p * kWinBytes + (size_t)ln * 128;
            if (a < win_table + (size_t)n * kWinBytes) asm volatile("prefetch.global.L2 [%0];" :: "l"(a));
        }
    }
    // ---- (2) reference rows p .. p+63 (clamped), two per lane --------------------------------------
    const double4* ref4 = (const double4*)ref;
    const double4 r_lo = ref4[min(p + lane, n - 1)];
    const double4 r_hi = ref4[min(p + 32 + lane, n - 1)];
    if (zx && lane == 0) {                        // fill the device mirror every later kernel reads
        double* dx = const_cast<double*>(io.x0) + 4 * e;
        dx[0] = q1; dx[1] = q2; dx[2] = dq1; dx[3] = dq2;
    }
    if (zc) {
        double2* du = (double2*)(const_cast<double*>(io.u_prev) + (size_t)e * T * 2);
        if (lane == 0) {
            const_cast<int32_t*>(io.prev_idx)[e] = *sp;
            if (e == 0) *const_cast<uint64_t*>(io.step) = stp;
        }
#pragma unroll
        for (int i = 0; i < kTS; ++i) { const int t = lane + 32 * i; if (t < T) du[t] = uu[i]; }
    }
    // control.py:206-215 in FP64: end-effector, distances to the forward window, first arg-min
    // (the two FP64 sincos are the longest arithmetic chain of this kernel: even lanes evaluate q1, odd lanes q1 + q2,
    //  and neighbours swap — the same function on the same arguments, half the latency)
    double s1, c1, s12, c12;
    {
        const bool odd = (lane & 1) != 0;
        double sa, ca;
        sincos(odd ? q1 + q2 : q1, &sa, &ca);
        const double sb = __shfl_xor_sync(0xffffffffu, sa, 1), cb = __shfl_xor_sync(0xffffffffu, ca, 1);
        s1 = odd ? sb : sa; c1 = odd ? cb : ca;
        s12 = odd ? sa : sb; c12 = odd ? ca : cb;
    }
    const double x = cfg.cost_l1 * c1 + cfg.cost_l2 * c12;
    const double y = cfg.cost_l1 * s1 + cfg.cost_l2 * s12;
    double d = 1.0e300;
    int j = lane;
    if (lane < kWindow && p + lane < n) {
        const double dx_ = x - r_lo.x, dy_ = y - r_lo.y;
        d = (dx_ * dx_ + dy_ * dy_) * 100;                   // control.py:210-212
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double d2 = __shfl_xor_sync(0xffffffffu, d, o);
        const int j2 = __shfl_xor_sync(0xffffffffu, j, o);
        if (d2 < d || (d2 == d && j2 < j)) { d = d2; j = j2; }
    }
    p += j;                                                  // control.py:230
    // row p + lane of the path = preloaded row (j + lane) relative to p_old
    const int src = j + lane;                                // 0 .. 60
    const bool from_hi = src >= 32;
    double4 row;
    {
        const int sl = src & 31;
        const double ax = __shfl_sync(0xffffffffu, r_lo.x, sl), bx = __shfl_sync(0xffffffffu, r_hi.x, sl);
        const double ay = __shfl_sync(0xffffffffu, r_lo.y, sl), by = __shfl_sync(0xffffffffu, r_hi.y, sl);
        const double az = __shfl_sync(0xffffffffu, r_lo.z, sl), bz = __shfl_sync(0xffffffffu, r_hi.z, sl);
        const double aw = __shfl_sync(0xffffffffu, r_lo.w, sl), bw = __shfl_sync(0xffffffffu, r_hi.w, sl);
        row = from_hi ? make_double4(bx, by, bz, bw) : make_double4(ax, ay, az, aw);
    }
    const double ox = __shfl_sync(0xffffffffu, row.x, 0), oy = __shfl_sync(0xffffffffu, row.y, 0);   // row p
    StepBlockView sb = view_step_block(step_blocks + (size_t)e * cfg.step_block_bytes);
    if (lane == 0) {
        StepHeader h;
        h.q1 = (float)q1; h.q2 = (float)q2; h.d1 = (float)dq1; h.d2 = (float)dq2;
        h.ox = (float)ox; h.oy = (float)oy;
        h.win_start = p; h.n_valid = min(kWindow, n - p);
        h.status = (p >= n - 1) ? 1 : 0;                     // control.py:76
        h.a1 = angle_fix(q1); h.a12 = angle_fix(q1 + q2);
        for (int i = 0; i < 5; ++i) h.pad[i] = 0;
        *sb.hd = h;
        out_store(io, io.new_idx + e, p);
    }
    if (win_table != nullptr) {
        // everything of the step block that depends only on the window start was built when the path was set:
        // copy the 2.4 KB of window p (five independent 16-byte loads per lane)
        const uint4* src = (const uint4*)(win_table + (size_t)p * kWinBytes);
        uint4* dst = (uint4*)((char*)sb.hd + 64);
#pragma unroll
        for (int i = 0; i < (kWinBytes / 16 + 31) / 32; ++i) {
            const int k = lane + 32 * i;
            if (k < kWinBytes / 16) dst[k] = __ldg(src + k);
        }
    } else {
        write_window_tables(row, ox, oy, lane, p, n, cfg, srow, sb);
    }
#pragma unroll
    for (int i = 0; i < kTS; ++i) {
        const int t = lane + 32 * i;
        if (t < T) {
            const double ut[2] = { uu[i].x, uu[i].y };
            StepCtl c; make_step_ctl(ut, cfg.gamma, cfg.sig_inv, c);
            sb.ctl[t] = c;
        }
    }
}

// ================================================================================================
// 2. fused rollout kernel: one thread per sample, step block staged in shared memory by TMA,
//    window coefficients held in registers, only S[k] and one block-min reach HBM.
// ================================================================================================
constexpr int kRollThreads = 128;

// MPPI_NOISE_AHEAD: the draw for horizon steps t+2, t+3 is issued at step t, next to the dynamics of step t instead
// of at the head of the dependent chain of step t+2 (Philox rounds -> Box-Muller -> input -> rates): the same calls
// on the same counters, so the same floats.  One call past the end of the horizon is drawn and dropped.
#ifndef MPPI_NOISE_AHEAD
#define MPPI_NOISE_AHEAD 1
#endif
struct PhiloxNoise {            // eps drawn in-kernel: one Philox call serves two horizon steps
    NoiseCfg nc; uint32_t env, k;
    float e1a, e1b;
#if MPPI_NOISE_AHEAD
    float n0a, n0b, n1a, n1b;   // the pair after the current one
    __device__ __forceinline__ void prime() { noise_pair(nc, env, k, 0u, n0a, n0b, n1a, n1b); }
    __device__ __forceinline__ void operator()(int t, float& a, float& b) {
        if ((t & 1) == 0) {
            a = n0a; b = n0b; e1a = n1a; e1b = n1b;
            noise_pair(nc, env, k, ((uint32_t)t >> 1) + 1u, n0a, n0b, n1a, n1b);
        } else { a = e1a; b = e1b; }
    }
#else
    __device__ __forceinline__ void prime() {}
    __device__ __forceinline__ void operator()(int t, float& a, float& b) {
        if ((t & 1) == 0) noise_pair(nc, env, k, (uint32_t)t >> 1, a, b, e1a, e1b);
        else { a = e1a; b = e1b; }
    }
#endif
};
struct InjectedNoise {          // eps read from the caller's [K,T,2] tensor
    const float2* row;
    __device__ __forceinline__ void operator()(int t, float& a, float& b) const {
        const float2 v = __ldg(row + t); a = v.x; b = v.y;
    }
};

// samples per thread in the rollout kernel: 2 for throughput (A/B on B200: profiles/r1_variants.md),
// 1 when there are too few samples to fill the GPU twice over (latency runs)

// Single-environment fast path: the window coefficients of the current step are copied (device to
// device, in stream order after the prepare kernel) into this constant-bank table, and the search's
// FFMAs read them as immediate constant operands: no registers, no shared-memory loads, and two
// register operands per FFMA instead of three.
struct ConstWindow {                 // same layout as the bytes of a step block that follow its header: ONE copy fills it
    WinEntry win[kWindowPad];
    RefRow rows[kWindowPad];         // (not read from here: the winning row is fetched from shared memory)
    float4 pairs[kWindowPad / 2];
};
__constant__ ConstWindow c_window;
#ifndef MPPI_FFMA2_SEARCH
#define MPPI_FFMA2_SEARCH 1      // packed fma.rn.f32x2 for the candidate distances (-4.6 % kernel time on B200)
#endif
#ifndef MPPI_CONST_C_REG
#define MPPI_CONST_C_REG 1     // keep c_j in registers so that each FFMA of the search has ONE constant operand
#endif
struct WinConst {
#if MPPI_CONST_C_REG
    float wc[kWindow];
    __device__ __forceinline__ float c(int j) const { return wc[j]; }
    __device__ __forceinline__ void load(const WinEntry* tab) {
#pragma unroll
        for (int j = 0; j < kWindow; ++j) wc[j] = tab[j].c;
    }
#else
    __device__ __forceinline__ float c(int j) const { return c_window.win[j].c; }
    __device__ __forceinline__ void load(const WinEntry*) {}
#endif
    __device__ __forceinline__ float a(int j) const { return c_window.win[j].a; }
    __device__ __forceinline__ float b(int j) const { return c_window.win[j].b; }
#if MPPI_FFMA2_SEARCH
    // distances of two adjacent candidates per packed FFMA2 (fma.rn.f32x2): same two roundings per
    // candidate as the scalar form, half the issue slots; the a/b pairs come from the constant bank
    // through uniform registers, the c pair from registers
    __device__ __forceinline__ void distances(float xl, float yl, float (&d)[kWindowPad]) const {
        const unsigned long long x2 = pack2(xl, xl), y2 = pack2(yl, yl);
#pragma unroll
        for (int i = 0; i < kWindow / 2; ++i) {
            const float4 ab = c_window.pairs[i];
            unsigned long long t, r;
            asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t) : "l"(pack2(ab.z, ab.w)), "l"(y2), "l"(pack2(wc[2 * i], wc[2 * i + 1])));
            asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pack2(ab.x, ab.y)), "l"(x2), "l"(t));
            d[2 * i] = __uint_as_float((unsigned)(r & 0xffffffffull));
            d[2 * i + 1] = __uint_as_float((unsigned)(r >> 32));
        }
    }
    static __device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
        return ((unsigned long long)__float_as_uint(hi) << 32) | (unsigned long long)__float_as_uint(lo);
    }
#endif
};
#ifndef MPPI_ROLL_MIN_BLOCKS_CONST
#define MPPI_ROLL_MIN_BLOCKS_CONST 4
#endif

#ifndef MPPI_ROLL_MIN_BLOCKS_CERT
#define MPPI_ROLL_MIN_BLOCKS_CERT 4      // <= 128 registers, no spills, per-step constants stay in registers (A/B on B200: profiles/r2s3_variants.md; 5 CTAs/SM was the choice while the float angles and their compensation terms were live)
#endif
#ifndef MPPI_ROLL_MIN_BLOCKS_CERT_NS1
#define MPPI_ROLL_MIN_BLOCKS_CERT_NS1 5  // the one-sample-per-thread kernel (small shards, latency runs)
#endif

#ifndef MPPI_NS_WIDE
#define MPPI_NS_WIDE 2                   // samples per thread of the throughput kernels
#endif
constexpr int kNsWide = MPPI_NS_WIDE;

// Window policy per kernel family.  CERT: certified lookups, the table stays in shared memory (any number of
// environments).  Otherwise plain searches with the coefficients in the constant bank (CONSTWIN) or in registers.
template <bool CONSTWIN, bool CERT> struct WinPolicy { typedef WinRegs type; };
template <> struct WinPolicy<true, false> { typedef WinConst type; };
template <bool CONSTWIN> struct WinPolicy<CONSTWIN, true> { typedef WinTable type; };
__device__ __forceinline__ void win_load(WinTable& w, const StepBlockView& sb) { w.load(*sb.cert, sb.rec, sb.wed); }
__device__ __forceinline__ void win_load(WinRegs& w, const StepBlockView& sb) { w.load(sb.win); }
__device__ __forceinline__ void win_load(WinConst& w, const StepBlockView& sb) { w.load(sb.win); }

// NS samples of one thread (kl0 + s * stride, s < NS) of environment e: noise source, rollout, cost store.
// Returns the smallest finite cost of the thread's samples.
template <int NS, int NOISE, int DYN, bool JL, class Win>
__device__ __forceinline__ float roll_samples(const DevCfg& cfg, const StepHeader& hd, const Win& win, const WinCert& cert,
                                              const StepBlockView& sb, const uint64_t* __restrict__ step_ctr,
                                              const float* __restrict__ eps, float* __restrict__ S_out, int e, int kl0,
                                              int stride, LookupStats& hits) {
    float um[NS], S[NS];
    int kl[NS];
    const int T = cfg.T;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        // a padding sample past the end recomputes the last one (its result is not stored)
        kl[s] = min(kl0 + s * stride, cfg.K_local - 1);
        um[s] = (cfg.k_offset + kl[s]) < cfg.n_exploit ? 1.0f : 0.0f;
        asm volatile("" : "+f"(um[s]));            // keep it in a register: not re-derived in every horizon step
    }
    if (NOISE == 0) {
        PhiloxNoise nz[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            nz[s].nc = cfg.noise; nz[s].nc.step = (uint32_t)(*step_ctr); nz[s].env = (uint32_t)e;
            nz[s].k = (uint32_t)(cfg.k_offset + kl[s]);
            nz[s].prime();
        }
        rollout_cost_n<NS, DYN, JL>(hd, cfg.arm, cfg.cost, win, cert, sb.rows, sb.srows, sb.ctl, T, um, nz, S, hits);
    } else {
        InjectedNoise nz[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) nz[s].row = (const float2*)eps + ((size_t)e * cfg.K_local + kl[s]) * T;
        rollout_cost_n<NS, DYN, JL>(hd, cfg.arm, cfg.cost, win, cert, sb.rows, sb.srows, sb.ctl, T, um, nz, S, hits);
    }
    float tmin = INFINITY;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        if (kl0 + s * stride < cfg.K_local) {
            S_out[(size_t)e * cfg.K_local + kl[s]] = S[s];
            if (finite_(S[s])) tmin = fminf(tmin, S[s]);
        }
    }
    return tmin;
}

// Two launch layouts.
//  * Uniform (kNS = 1, and the kernels without certified lookups): grid (blocks per environment, environments), every
//    thread takes kNS samples, grid-stride over the environment's samples.
//  * Flat, mixed (certified kernels with kNS > 1): a 1-D grid over "units" of 128 consecutive samples, numbered through
//    all environments (cfg.g_roll units each).  CTAs [0, n_wide) take kNS units (kNS samples per thread), the rest ONE
//    unit (one sample per thread).  The host picks n_wide so that the LAST wave of CTAs is full: a shard of 1.0 to
//    2.0 waves' worth of samples (131072 samples: an 8-GPU run, or 128 batched environments) runs as one wave of 432
//    two-sample and 160 one-sample CTAs instead of 1.4 waves of one-sample CTAs.
template <int NOISE, bool CONSTWIN, int kNS, int DYN = 0, bool CERT = true, bool JL = false>
__global__ void __launch_bounds__(kRollThreads, CERT ? (kNS == 1 ? MPPI_ROLL_MIN_BLOCKS_CERT_NS1 : MPPI_ROLL_MIN_BLOCKS_CERT)
                                                     : (CONSTWIN ? MPPI_ROLL_MIN_BLOCKS_CONST : (kNS == 1 ? 3 : 2)))
mppi_rollout_sm100a(DevCfg cfg, const uint64_t* __restrict__ step_ctr, const char* __restrict__ step_blocks,
                    const float* __restrict__ eps, float* __restrict__ S_out, float* __restrict__ block_min,
                    unsigned long long* __restrict__ search_stats, unsigned int* __restrict__ rho_key, int n_wide) {
    extern __shared__ __align__(128) unsigned char smem_roll[];
    unsigned char* smem = smem_roll;
    __shared__ uint64_t bar;
    __shared__ float red[kRollThreads / 32];
    constexpr bool kFlat = CERT && kNS > 1;
    const int tid = threadIdx.x;
    int e = blockIdx.y, unit = 0, units = 1;          // flat layout: first unit of this CTA within its environment, and how many
    if (kFlat) {
        const int b = blockIdx.x;
        const int first = b < n_wide ? b * kNS : n_wide * kNS + (b - n_wide);
        e = first / cfg.g_roll; unit = first - e * cfg.g_roll;
        units = b < n_wide ? kNS : 1;
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, (uint32_t)cfg.step_block_bytes);
    }
    pdl_wait();                                   // the prepare kernel's step block (and step counter) are complete
    if (tid == 0)
        tma_load_1d(smem, step_blocks + (size_t)e * cfg.step_block_bytes, (uint32_t)cfg.step_block_bytes, &bar);
    __syncthreads();
    mbar_wait(&bar, 0);
    const StepBlockView sb = view_step_block(smem);
    const StepHeader hd = *sb.hd;
    typename WinPolicy<CONSTWIN, CERT>::type win;
    win_load(win, sb);
    const WinCert& cert = *sb.cert;                   // (staged step block; read by the certified lookups only)
    LookupStats hits = { 0, 0 };
    int lookups = 0;

    float tmin = INFINITY;
    const int T = cfg.T;
    // thread handles samples kl0 + s*blockDim.x (s < NS): consecutive lanes -> consecutive samples
    const int nthr = blockDim.x;
    if (kFlat) {
        const int kl0 = unit * kRollThreads + tid;
        if (units == kNS) {
            lookups = kNS * T;
            tmin = roll_samples<kNS, NOISE, DYN, JL>(cfg, hd, win, cert, sb, step_ctr, eps, S_out, e, kl0, kRollThreads, hits);
        } else {
            lookups = T;
            tmin = roll_samples<1, NOISE, DYN, JL>(cfg, hd, win, cert, sb, step_ctr, eps, S_out, e, kl0, kRollThreads, hits);
        }
    } else {
        // The trip count is decided per WARP (its first lane), because the lookups vote across the warp.
        for (int kw0 = blockIdx.x * (nthr * kNS) + (tid & ~31); kw0 < cfg.K_local; kw0 += gridDim.x * nthr * kNS) {
            lookups += kNS * T;
            tmin = fminf(tmin, roll_samples<kNS, NOISE, DYN, JL>(cfg, hd, win, cert, sb, step_ctr, eps, S_out, e,
                                                                 kw0 + (tid & 31), nthr, hits));
        }
    }
    tmin = warp_min(tmin);
    if ((tid & 31) == 0) red[tid >> 5] = tmin;
    if ((cfg.flags & 64) && (tid & 31) == 0) {              // MPPI_FLAG_SEARCH_STATS: warp-lookups certified / done
        atomicAdd(search_stats, (unsigned long long)(lookups - hits.tri - hits.scan));
        atomicAdd(search_stats + 1, (unsigned long long)lookups);
        atomicAdd(search_stats + 2, (unsigned long long)hits.tri);
    }
    __syncthreads();
    if (tid == 0) {
        float m = red[0];
#pragma unroll
        for (int i = 1; i < nthr / 32; ++i) m = fminf(m, red[i]);
        if (kFlat) {                                        // one slot per unit (read by the unfused soft-min kernel)
            for (int i = 0; i < units; ++i)
                if (unit + i < cfg.g_roll) block_min[(size_t)e * cfg.g_roll + unit + i] = m;
        } else {
            block_min[(size_t)e * gridDim.x + blockIdx.x] = m;
        }
        if (finite_(m)) atomicMin(rho_key + e, min_key(m));     // what the fused weight-sum kernel reads (one word, not g_roll)
    }
}

// ================================================================================================
// 3. soft-min: rho = min S (from the rollout's block minima), w~_k = exp(-(S_k - rho)/lambda),
//    per-block partial sums of w~ (control.py:297-314).  Non-finite costs get weight 0.
// ================================================================================================
constexpr int kSoftThreads = 256;

__global__ void __launch_bounds__(kSoftThreads)
mppi_softmin_sm100a(DevCfg cfg, const float* __restrict__ S, const float* __restrict__ block_min,
                    float* __restrict__ w, double* __restrict__ eta_part, float* __restrict__ rho_out) {
    __shared__ float redf[kSoftThreads / 32];
    __shared__ double redd[kSoftThreads / 32];
    __shared__ float rho_s;
    const int e = blockIdx.y, tid = threadIdx.x;
    float m = INFINITY;
    for (int i = tid; i < cfg.g_roll; i += kSoftThreads) m = fminf(m, block_min[(size_t)e * cfg.g_roll + i]);
    m = warp_min(m);
    if ((tid & 31) == 0) redf[tid >> 5] = m;
    __syncthreads();
    if (tid == 0) {
        float r = redf[0];
#pragma unroll
        for (int i = 1; i < kSoftThreads / 32; ++i) r = fminf(r, redf[i]);
        rho_s = r;
        if (blockIdx.x == 0) rho_out[e] = r;
    }
    __syncthreads();
    const float rho = rho_s;
    const float nil = (float)(-cfg.inv_lambda);
    const float* Se = S + (size_t)e * cfg.K_local;
    float* we = w + (size_t)e * cfg.K_local;
    double acc = 0.0;
    for (int k = blockIdx.x * kSoftThreads + tid; k < cfg.K_local; k += gridDim.x * kSoftThreads) {
        const float s = Se[k];
        const float x = expf((s - rho) * nil);
        const float wk = finite_(s) ? x : 0.0f;
        we[k] = wk;
        acc += (double)wk;
    }
    acc = warp_sum(acc);
    if ((tid & 31) == 0) redd[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double r = 0.0;
#pragma unroll
        for (int i = 0; i < kSoftThreads / 32; ++i) r += redd[i];
        eta_part[(size_t)e * gridDim.x + blockIdx.x] = r;
    }
}

// ================================================================================================
// 4a. weighted noise sum, injected noise: V[t,m] = sum_k w~_k eps[k,t,m]  (control.py:115-118).
//     The K x (2T) matrix is read as a flat stream of VEC-wide words: thread (r, c) owns column c
//     of every (R*gridDim.x)-th row, so consecutive threads read consecutive 16-byte words.
//     Rows whose weight is exactly 0 are not read at all.
// ================================================================================================
constexpr int kWsumThreads = 256;

template <typename VEC> struct VecOps;
template <> struct VecOps<float4> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void fmadd(float4& a, float w, const float4& v) {
        a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
    }
    static __device__ __forceinline__ void add(float4& a, const float4& v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
    static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ float4 ld(const float4* p) {
        float4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
        return v;
    }
};
template <> struct VecOps<float2> {
    static constexpr int N = 2;
    static __device__ __forceinline__ void fmadd(float2& a, float w, const float2& v) {
        a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y);
    }
    static __device__ __forceinline__ void add(float2& a, const float2& v) { a.x += v.x; a.y += v.y; }
    static __device__ __forceinline__ float2 zero() { return make_float2(0.f, 0.f); }
    static __device__ __forceinline__ float2 ld(const float2* p) { return __ldg(p); }
};

constexpr int kWsumRows = 8;          // rows (samples) in flight per thread

template <typename VEC>
__global__ void __launch_bounds__(kWsumThreads)
mppi_wsum_injected_sm100a(DevCfg cfg, const float* __restrict__ w, const float* __restrict__ eps,
                          float* __restrict__ v_part) {
    extern __shared__ __align__(16) unsigned char smem_wsum[];
    VEC* sh = (VEC*)smem_wsum;
    const int e = blockIdx.y, tid = threadIdx.x;
    const int C = (2 * cfg.T) / VecOps<VEC>::N;               // words per row
    const int R = kWsumThreads / C;                           // rows per pass of one block
    const int r = tid / C, c = tid - r * C;
    const bool active = r < R;
    const float* we = w + (size_t)e * cfg.K_local;
    const VEC* ee = (const VEC*)eps + (size_t)e * cfg.K_local * C;
    VEC acc = VecOps<VEC>::zero();
    if (active) {
        const int stride = R * gridDim.x;
        const int K = cfg.K_local;
        // kWsumRows independent weight loads, then kWsumRows independent 16-byte row loads per thread:
        // ~32 KB of loads in flight per CTA keeps HBM busy with a few CTAs per SM
        for (int k = blockIdx.x * R + r; k < K; k += kWsumRows * stride) {
            float wk[kWsumRows]; VEC v[kWsumRows];
#pragma unroll
            for (int i = 0; i < kWsumRows; ++i) {
                const int ki = k + i * stride;
                wk[i] = ki < K ? __ldg(we + ki) : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < kWsumRows; ++i)
                v[i] = wk[i] != 0.0f ? VecOps<VEC>::ld(ee + (size_t)(k + i * stride) * C + c) : VecOps<VEC>::zero();
#pragma unroll
            for (int i = 0; i < kWsumRows; ++i) VecOps<VEC>::fmadd(acc, wk[i], v[i]);
        }
        sh[r * C + c] = acc;
    }
    __syncthreads();
    if (tid < C) {                                            // fixed-order sum over the R row slots
        VEC s = sh[tid];
        for (int i = 1; i < R; ++i) VecOps<VEC>::add(s, sh[i * C + tid]);
        ((VEC*)(v_part + ((size_t)e * cfg.g_wsum + blockIdx.x) * 2 * cfg.T))[tid] = s;
    }
}

// ================================================================================================
// 6. finalize: combine the gathered partials of all ranks (identically on every rank), normalise,
//    median-filter, update the sequence, roll the optimal trajectory out (control.py:122-134).
// ================================================================================================
__device__ __forceinline__ int reflect_idx(int i, int n) {          // scipy 'reflect': d c b a | a b c d | d c b a
    if (n >= kFilter) {                     // |overhang| < n: one mirror about either end, no division
        i = i < 0 ? -1 - i : i;
        return i >= n ? 2 * n - 1 - i : i;
    }
    const int period = 2 * n;
    i %= period; if (i < 0) i += period;
    return i >= n ? period - 1 - i : i;
}

struct FinalizeSmem {
    double raw[2 * MPPI_MAX_T_INTERNAL];
    double unew[2 * MPPI_MAX_T_INTERNAL];
    __align__(16) float tr[8 * MPPI_MAX_T_INTERNAL];    // optimal trajectory: (value, compensation) per state and step
    __align__(8) float vf[2 * MPPI_MAX_T_INTERNAL];     // the updated sequence in FP32, in the order the trajectory applies it
    double scale[64];
    double eta_s;
    int timed_out;
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Everything after this GPU's partial triple, for environment e, by all threads of the calling block:
// wait for the peers' triples (peer exchange only), combine, normalise, filter, update, optimal trajectory.
// `gathered` = double [world][n_env][2 + 2T] (ignored with the peer exchange: its slots are read instead).
__device__ __forceinline__ void finalize_env(const DevCfg& cfg, const DevIo& io, int e, const double* gathered, int world,
                                             const PeerExchange& px, FinalizeSmem& sm) {
    const int tid = threadIdx.x, T = cfg.T;
    // the nominal control this thread updates below: loaded now, so that the L2 round trip is over by then
    const double u_first = tid < 2 * T ? io.u_prev[(size_t)e * 2 * T + tid] : 0.0;
#ifdef MPPI_PHASE_PRINT
    unsigned long long fp[8]; fp[0] = global_ns();
#define MPPI_FPHASE(i) fp[i] = global_ns()
#else
#define MPPI_FPHASE(i)
#endif
    if (tid == 0) sm.timed_out = 0;
    __syncthreads();
    if (px.world > 0) {
        // wait until every rank's triple of THIS step has landed in the local exchange buffer
        const unsigned long long seq = *px.seq;
        const int par = (int)(seq & 1ull);
        if (tid < px.world) {
            volatile const unsigned long long* f = px_flag(px, px.rank, par, tid, cfg.n_env, e);
            const unsigned long long t0 = global_ns();
            while (*f < seq) {
                if (global_ns() - t0 > px.timeout_ns) { sm.timed_out = 1; break; }   // a peer died or is far behind
                __nanosleep(64);
            }
        }
        __threadfence_system();
        __syncthreads();
        gathered = (const double*)(px.buf[px.rank] + (size_t)par * px.slot_bytes);
        world = px.world;
    }
    MPPI_FPHASE(1);
    // A partial that did not arrive must not be combined (its slot holds an older step): the update is
    // skipped (u_new = u_prev, zero update, NaN rho / eta) and bit 0 of the status word tells the caller.
    const bool skip = sm.timed_out != 0;
    const size_t stride_rank = (size_t)cfg.n_env * (2 + 2 * T);
    const double* g0 = gathered + (size_t)e * (2 + 2 * T);
#ifndef MPPI_COMBINE_GENERAL
    if (world == 1) {
        // one GPU: rho = rho_0, the rescaling factor is exp(-0) = 1 and eta = 1 * eta_0 — the same values as the
        // general form below without its FP64 exp and minimum on the latency path
        if (tid == 0) {
            const double rho = g0[0], d = rho - rho;          // 0, or NaN for a shard without a finite cost
            const double sc = d == 0.0 ? 1.0 : d, eta = sc * g0[1];
            sm.scale[0] = sc;
            sm.eta_s = eta;
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            out_store(io, io.rho + e, skip ? nan : rho); out_store(io, io.eta + e, skip ? nan : eta);
            out_store(io, io.status + e, skip ? 1 : 0);
        }
    } else
#endif
    if (tid < 32) {
        // one lane per rank: its minimum, and its rescaling factor (an FP64 exp each — side by side, not one after the
        // other); lane 0 then adds the weight sums in rank order
        double rho = __longlong_as_double(0x7ff0000000000000ll);
        for (int g = tid; g < world; g += 32) rho = fmin(rho, g0[g * stride_rank]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rho = fmin(rho, __shfl_xor_sync(0xffffffffu, rho, o));
        // a shard with no finite cost reports rho_g = +inf and eta_g = 0
        for (int g = tid; g < world; g += 32) sm.scale[g] = exp(-(g0[g * stride_rank] - rho) * cfg.inv_lambda);
        __syncwarp();
        if (tid == 0) {
            double eta = 0.0;
            for (int g = 0; g < world; ++g) eta += sm.scale[g] * g0[g * stride_rank + 1];
            sm.eta_s = eta;
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            out_store(io, io.rho + e, skip ? nan : rho); out_store(io, io.eta + e, skip ? nan : eta);
            out_store(io, io.status + e, skip ? 1 : 0);
        }
    }
    __syncthreads();
    for (int c = tid; c < 2 * T; c += blockDim.x) {
        double v = 0.0;
        for (int g = 0; g < world; ++g) {
            const double sg = sm.scale[g];
            if (sg != 0.0) v += sg * g0[g * stride_rank + 2 + c];
        }
        v = skip ? 0.0 : v / sm.eta_s;
        sm.raw[c] = v;
        out_store(io, io.w_eps_raw + (size_t)e * 2 * T + c, v);
    }
    __syncthreads();
    MPPI_FPHASE(2);
    // scipy.ndimage.median_filter(size=10, mode='reflect') per column (control.py:319-327):
    // window offsets -5..+4, output = element of rank 5 of the sorted window
    for (int c = tid; c < 2 * T; c += blockDim.x) {
        const int t = c >> 1, m = c & 1;
        double med;
        if (cfg.flags & 8) {                                  // MPPI_FLAG_SMOOTH_NONE
            med = sm.raw[c];
        } else if (cfg.flags & 4) {                           // MPPI_FLAG_SMOOTH_AVERAGE, control.py:329-344
            // np.convolve(x, ones/10, 'same') = sum of x/10 over [t-5, t+4] clipped to the array, then the
            // first and last ceil(10/2)-1 rows (and row 0) rescaled by 10/(number of terms)
            const int lo = max(0, t - kFilter / 2), hi = min(T - 1, t + (kFilter - 1) / 2);
            double acc = 0.0;
            for (int k = lo; k <= hi; ++k) acc += sm.raw[2 * k + m] * (1.0 / kFilter);
            const int n_conv = (kFilter + 1) / 2;
            if (t == 0) acc *= (double)kFilter / n_conv;
            else if (t < n_conv) acc *= (double)kFilter / (t + n_conv);
            if (t > T - n_conv && t != 0) acc *= (double)kFilter / ((T - t) + n_conv - (kFilter % 2));
            med = acc;
        } else {
            double win[kFilter];
#pragma unroll
            for (int o = 0; o < kFilter; ++o) win[o] = sm.raw[2 * reflect_idx(t + o - kFilter / 2, T) + m];
#ifdef MPPI_MEDIAN_BY_RANK
            med = win[0];
#pragma unroll
            for (int a = 0; a < kFilter; ++a) {
                int rank = 0;
#pragma unroll
                for (int b = 0; b < kFilter; ++b) rank += (win[b] < win[a]) || (win[b] == win[a] && b < a);
                if (rank == kFilter / 2) med = win[a];
            }
#else
            // element of rank 5 by a 29-exchange sorting network for ten inputs (verified by the 0-1 principle,
            // tests/test_host_cpu.py); the VALUE of that rank does not depend on how ties are ordered.  The
            // exchanges that cannot reach output 5 are removed by the compiler.  (58 FP64 min / max for ~250
            // compares of the rank count.)
            static_assert(kFilter == 10, "the network below sorts ten inputs");
#define MPPI_CE(a, b) { const double lo_ = fmin(win[a], win[b]), hi_ = fmax(win[a], win[b]); win[a] = lo_; win[b] = hi_; }
            MPPI_CE(0, 8) MPPI_CE(1, 9) MPPI_CE(2, 7) MPPI_CE(3, 5) MPPI_CE(4, 6)
            MPPI_CE(0, 2) MPPI_CE(1, 4) MPPI_CE(5, 8) MPPI_CE(7, 9)
            MPPI_CE(0, 3) MPPI_CE(2, 4) MPPI_CE(5, 7) MPPI_CE(6, 9)
            MPPI_CE(0, 1) MPPI_CE(3, 6) MPPI_CE(8, 9)
            MPPI_CE(1, 5) MPPI_CE(2, 3) MPPI_CE(4, 8) MPPI_CE(6, 7)
            MPPI_CE(1, 2) MPPI_CE(3, 5) MPPI_CE(4, 6) MPPI_CE(7, 8)
            MPPI_CE(2, 3) MPPI_CE(4, 5) MPPI_CE(6, 7)
            MPPI_CE(3, 4) MPPI_CE(5, 6)
#undef MPPI_CE
            med = win[kFilter / 2];
#endif
        }
        const double u = (c == tid ? u_first : io.u_prev[(size_t)e * 2 * T + c]) + med;        // control.py:126
        sm.unew[c] = u;
        out_store(io, io.w_eps_filt + (size_t)e * 2 * T + c, med);
        out_store(io, io.u_new + (size_t)e * 2 * T + c, u);
    }
    __syncthreads();
    MPPI_FPHASE(3);
    // What calc_control_input returns as the control (control.py:148-152, quirk Q2: the first row AFTER the
    // shift), and — MPPI_FLAG_RESIDENT_STATE — the controller state of the next step, kept on the device:
    // shifted sequence (control.py:148-149) and waypoint index (control.py:230).  An environment at the end
    // of its path (control.py:76-78: the reference raises there) is frozen.
    {
        const bool ended = io.new_idx[e] >= cfg.n_ref_rows - 1;
        const int t1 = (T > 1 && !ended) ? 1 : 0;
        if (tid < 2) out_store(io, io.u0 + 2 * e + tid, ended ? io.u_prev[(size_t)e * 2 * T + tid] : sm.unew[2 * t1 + tid]);
        if (cfg.flags & 128) {
            double* up = const_cast<double*>(io.u_prev) + (size_t)e * 2 * T;
            if (!ended)
                for (int c = tid; c < 2 * T; c += blockDim.x) up[c] = sm.unew[2 * min((c >> 1) + 1, T - 1) + (c & 1)];
            if (tid == 0) const_cast<int32_t*>(io.prev_idx)[e] = io.new_idx[e];
        }
    }
    // control.py:129-134: x <- F(x, u[t-1]) for t = 0..T-1 (t = 0 wraps to the last control, Q3).
    // The recurrence is serial: one thread runs it and parks (value, compensation) pairs in shared
    // memory; all threads then convert and store the trajectory (keeps global / PCIe stores off the chain).
    MPPI_FPHASE(4);
    if (cfg.flags & 1) {
        // controls in the order of use (t = 0 wraps to the last one, Q3), converted once by all threads: the serial
        // loop then only waits for arithmetic — its loads run one step ahead of their use
        for (int t = tid; t < T; t += blockDim.x) {
            const int tc = t == 0 ? T - 1 : t - 1;
            ((float2*)sm.vf)[t] = make_float2((float)sm.unew[2 * tc], (float)sm.unew[2 * tc + 1]);
        }
        __syncthreads();
        if (tid == 0) {
            const double* x0 = io.x0 + 4 * e;
            ArmState st;
            arm_init(st, (float)x0[0], (float)x0[1], (float)x0[2], (float)x0[3], angle_fix(x0[0]), angle_fix(x0[0] + x0[1]));
            const bool f1 = (cfg.flags & 32) != 0;                       // MPPI_FLAG_DYNAMICS_F1
            // the arm constants in registers: left to itself the compiler re-loads four of them from the constant bank
            // (LDC, a memory-path load) at the head of every iteration of this latency-bound loop
            // (a one-lane shuffle: the only form ptxas does not see through and turn back into LDC)
            ArmF A = cfg.arm;
#ifndef MPPI_TRAJ_LDC
            A.A0 = __shfl_sync(1u, A.A0, 0); A.A1 = __shfl_sync(1u, A.A1, 0);
            A.M22 = __shfl_sync(1u, A.M22, 0); A.B1 = __shfl_sync(1u, A.B1, 0);
#endif
            float2 v = ((const float2*)sm.vf)[0];
            for (int t = 0; t < T; ++t) {
                const float2 vn = ((const float2*)sm.vf)[t + 1 < T ? t + 1 : t];
                if (f1) arm_step_serial<1>(st, A, v.x, v.y);
                else arm_step_serial<0>(st, A, v.x, v.y);
                float4* o = (float4*)(sm.tr + 8 * t);
                o[0] = make_float4(st.q1, st.q2, st.d1, st.d2);
                o[1] = make_float4(st.kq1, st.kq2, st.kd1, st.kd2);
                v = vn;
            }
        }
        __syncthreads();
        MPPI_FPHASE(5);
        double* o = io.opt_traj + (size_t)e * 4 * T;
        for (int c = tid; c < 4 * T; c += blockDim.x) {
            const int t = c >> 2, k = c & 3;
            out_store(io, o + c, (double)sm.tr[8 * t + k] - (double)sm.tr[8 * t + 4 + k]);
        }
    } else {
        MPPI_FPHASE(5);
        for (int c = tid; c < 4 * T; c += blockDim.x) out_store(io, io.opt_traj + (size_t)e * 4 * T + c, 0.0);
    }
#ifdef MPPI_PHASE_PRINT
    MPPI_FPHASE(6);
    if (tid == 0 && e == 0)
        printf("finalize: wait %llu, combine %llu, filter+update %llu, u0/resident %llu, trajectory %llu, stores %llu ns\n",
               fp[1] - fp[0], fp[2] - fp[1], fp[3] - fp[2], fp[4] - fp[3], fp[5] - fp[4], fp[6] - fp[5]);
#endif
#undef MPPI_FPHASE
}

constexpr int kPairSlots = MPPI_MAX_T_INTERNAL / 2 / 32;      // Philox calls per lane and sample (weight-sum kernel)
#ifndef MPPI_WSUM_SAMPLES_PER_BLOCK
#define MPPI_WSUM_SAMPLES_PER_BLOCK 2048                      // two batches of four cost loads per thread; 512 blocks at K = 2^20 are one wave (64 registers: 4 blocks per SM)
#endif
constexpr int kWsumSamplesPerBlock = MPPI_WSUM_SAMPLES_PER_BLOCK;
constexpr int kWsumMaxBlocks = 4 * MPPI_MAX_T_INTERNAL;                        // blocks per environment of the fused weight-sum kernel (host: <= 8 per SM)

// ================================================================================================
// 4b. Philox mode, fused: soft-min weights + weighted noise sum + this GPU's partial triple in ONE
//     launch (three launches less per control step, which is what bounds the 1 kHz loop and the
//     8-GPU strong-scaling run).  Same arithmetic as kernels 3, 4a and 5 with eps regenerated
//     (bit-identical to the rollout kernel's draw: same noise_pair() on the same counters): every
//     block derives rho from the rollout's block minima, weights its slice of samples, regenerates
//     eps for the non-zero weights, writes its partials; the block that arrives last (atomic ticket)
//     adds the partials of all blocks in block order, so the result does not depend on arrival order.
// ================================================================================================
__global__ void __launch_bounds__(kWsumThreads)
mppi_softmin_wsum_philox_sm100a(DevCfg cfg, const uint64_t* __restrict__ step_ctr, const float* __restrict__ S,
                                const unsigned int* __restrict__ rho_key, float* __restrict__ w,
                                double* __restrict__ eta_part, float* __restrict__ v_part,
                                unsigned int* __restrict__ tickets, float* __restrict__ rho_out,
                                double* __restrict__ partial, PeerExchange px, DevIo io, int fuse_finalize) {
    __shared__ FinalizeSmem fin;                              // (used by the last block only)
    extern __shared__ __align__(16) unsigned char smem_wsum[];
    float4* sh = (float4*)smem_wsum;                          // [warps][pairs]
    __shared__ double redd[kWsumThreads / 32];
    __shared__ bool is_last;
    __shared__ int n_live;
    // scratch of the last block's cross-block sum, laid over arrays the final stage only fills afterwards:
    // eta partials of this environment's blocks, and the blocks with a non-zero weight in block order
    static_assert(offsetof(FinalizeSmem, unew) == offsetof(FinalizeSmem, raw) + sizeof(fin.raw), "raw and unew are adjacent");
    static_assert(kWsumMaxBlocks * sizeof(double) <= sizeof(fin.raw) + sizeof(fin.unew), "eta partials fit");
    static_assert(kWsumMaxBlocks * sizeof(unsigned short) <= sizeof(fin.tr), "row list fits");
    double* eta_s = fin.raw;
    unsigned short* live_rows = (unsigned short*)fin.tr;
    const int e = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_pairs = (cfg.T + 1) >> 1;
#ifdef MPPI_PHASE_PRINT
    unsigned long long ph[8]; ph[0] = global_ns();
#define MPPI_PHASE(i) ph[i] = global_ns()
#else
#define MPPI_PHASE(i)
#endif
    pdl_wait();                                   // the rollout kernel's costs and block minima are complete
    MPPI_PHASE(1);
    // rho = min over the rollout kernel's block minima
    // rho = min S: the rollout CTAs folded their minima into one key per environment
    const float rho = min_key_decode(__ldcg(rho_key + e));
    if (tid == 0 && blockIdx.x == 0) rho_out[e] = rho;
    const float nil = (float)(-cfg.inv_lambda);
    NoiseCfg nc = cfg.noise; nc.step = (uint32_t)(*step_ctr);
    const float* Se = S + (size_t)e * cfg.K_local;
    float* we = w + (size_t)e * cfg.K_local;
    float4 acc[kPairSlots];
#pragma unroll
    for (int i = 0; i < kPairSlots; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    double eta = 0.0;
    const int warps_total = gridDim.x * (kWsumThreads / 32);
    const int chunk_stride = warps_total * 32;
    constexpr int kAhead = 4;                                 // cost loads in flight per lane (the scan is latency bound)
    for (int kb = (blockIdx.x * (kWsumThreads / 32) + warp) * 32; kb < cfg.K_local; kb += kAhead * chunk_stride) {
        float sv[kAhead];
#pragma unroll
        for (int u = 0; u < kAhead; ++u) {
            const int k = kb + u * chunk_stride + lane;
            sv[u] = k < cfg.K_local ? Se[k] : INFINITY;
        }
#pragma unroll
        for (int u = 0; u < kAhead; ++u) {
            const int k0 = kb + u * chunk_stride;
            if (k0 >= cfg.K_local) break;
            const int k = k0 + lane;
            float wk = 0.0f;
            if (k < cfg.K_local) {
                const float x = expf((sv[u] - rho) * nil);
                wk = finite_(sv[u]) ? x : 0.0f;
                we[k] = wk;
                eta += (double)wk;
            }
            unsigned mask = __ballot_sync(0xffffffffu, wk != 0.0f);
            while (mask) {
                const int b = __ffs(mask) - 1;
                mask &= mask - 1;
                const float wb = __shfl_sync(0xffffffffu, wk, b);
                const uint32_t kg = (uint32_t)(cfg.k_offset + k0 + b);
#pragma unroll
                for (int i = 0; i < kPairSlots; ++i) {
                    const int pr = lane + 32 * i;
                    if (pr < n_pairs) {
                        float a0, a1, b0, b1;
                        noise_pair(nc, (uint32_t)e, kg, (uint32_t)pr, a0, a1, b0, b1);
                        acc[i].x = fmaf(wb, a0, acc[i].x); acc[i].y = fmaf(wb, a1, acc[i].y);
                        acc[i].z = fmaf(wb, b0, acc[i].z); acc[i].w = fmaf(wb, b1, acc[i].w);
                    }
                }
            }
        }
    }
    MPPI_PHASE(2);
    eta = warp_sum(eta);
    if (lane == 0) redd[warp] = eta;
#pragma unroll
    for (int i = 0; i < kPairSlots; ++i) {
        const int pr = lane + 32 * i;
        if (pr < n_pairs) sh[warp * n_pairs + pr] = acc[i];
    }
    __syncthreads();
    if (tid == 0) {
        double r = 0.0;
#pragma unroll
        for (int i = 0; i < kWsumThreads / 32; ++i) r += redd[i];
        eta_part[(size_t)e * gridDim.x + blockIdx.x] = r;       // [n_env][gridDim.x] (sized for g_wsum >= gridDim.x)
    }
    float* vrow = v_part + ((size_t)e * cfg.g_wsum + blockIdx.x) * 2 * cfg.T;
    for (int pr = tid; pr < n_pairs; pr += kWsumThreads) {
        float4 s4 = sh[pr];
        for (int wv = 1; wv < kWsumThreads / 32; ++wv) {
            const float4 v = sh[wv * n_pairs + pr];
            s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
        }
        vrow[4 * pr] = s4.x; vrow[4 * pr + 1] = s4.y;
        if (2 * pr + 1 < cfg.T) { vrow[4 * pr + 2] = s4.z; vrow[4 * pr + 3] = s4.w; }
    }
    // ---- last block: this GPU's partial (rho_g, eta_g, V_g), fixed summation order ----------------
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = (atomicAdd(&tickets[e], 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    MPPI_PHASE(3);
    __threadfence();
    double* out = partial + (size_t)e * (2 + 2 * cfg.T);
    const int G = gridDim.x;
    // Rows of blocks without a single non-zero weight (eta partial exactly 0: weights are >= 0) hold +0.0 in every
    // column; adding them changes nothing, so only the other rows are read — in block order, like before.  With the
    // reference's lambda the weights are winner-take-all: a handful of the G rows.
    for (int i = tid; i < G; i += kWsumThreads) eta_s[i] = __ldcg(eta_part + (size_t)e * G + i);   // (all loads in flight at once)
    __syncthreads();
    if (warp == 0) {
        double a = 0.0;
        int n = 0;
        for (int base = 0; base < G; base += 32) {
            const int i = base + lane;
            const double ep = i < G ? eta_s[i] : 0.0;
            a += ep;
            const unsigned mk = __ballot_sync(0xffffffffu, ep != 0.0);
            if (ep != 0.0) live_rows[n + __popc(mk & ((1u << lane) - 1u))] = (unsigned short)i;
            n += __popc(mk);
        }
        a = warp_sum(a);
        if (lane == 0) { out[0] = (double)rho; out[1] = a; tickets[e] = 0u; n_live = n; }
    }
    __syncthreads();
    // (a row costs ~37 ns through ONE block — bound by the number of requests a single SM keeps in flight: 9.5 us for
    //  all G = 256 rows at K = 2^20; unrolled / batched / warp-split forms of this loop measured the same or worse,
    //  profiles/r2_variants.md)
    {
        const int nl = n_live;
        for (int c = tid; c < 2 * cfg.T; c += kWsumThreads) {
            double a = 0.0;
            const float* src = v_part + (size_t)e * cfg.g_wsum * 2 * cfg.T + c;
            for (int b = 0; b < nl; ++b) a += (double)__ldcg(src + (size_t)live_rows[b] * 2 * cfg.T);
            out[2 + c] = a;
        }
    }
    MPPI_PHASE(4);
    if (px.world > 0) px_put(px, out, cfg.n_env, e, 2 + 2 * cfg.T);   // fused exchange: triple -> every peer
    MPPI_PHASE(5);
    // fused combine / filter / update / optimal trajectory: one kernel boundary less per control step.  With
    // the peer exchange this block then waits for the other ranks' triples (their last blocks put them the
    // same way; every rank's kernel is resident on its own GPU, so the wait cannot deadlock).
    if (fuse_finalize) {
        __threadfence();
        __syncthreads();
        finalize_env(cfg, io, e, partial, 1, px, fin);
    }
#ifdef MPPI_PHASE_PRINT
    MPPI_PHASE(6);
    if (tid == 0 && e == 0)
        printf("wsum last block: wait %llu, scan %llu, partials+ticket %llu, reduce %llu, put %llu, finalize %llu ns (G=%d)\n",
               ph[1] - ph[0], ph[2] - ph[1], ph[3] - ph[2], ph[4] - ph[3], ph[5] - ph[4], ph[6] - ph[5], (int)gridDim.x);
#endif
#undef MPPI_PHASE
}

// ================================================================================================
// 5. reduce: this GPU's partial triple per environment, FP64, fixed summation order.
//    partial[e] = { rho_g, eta_g, V_g[2T] }
// ================================================================================================
constexpr int kReduceThreads = 1024;

__global__ void __launch_bounds__(kReduceThreads)
mppi_reduce_sm100a(DevCfg cfg, int n_wsum_blocks, const float* __restrict__ rho, const double* __restrict__ eta_part,
                   const float* __restrict__ v_part, double* __restrict__ partial, PeerExchange px) {
    __shared__ double red[kReduceThreads / 32];
    __shared__ double colsum[kReduceThreads];
    const int e = blockIdx.x, tid = threadIdx.x;
    double* out = partial + (size_t)e * (2 + 2 * cfg.T);
    double a = 0.0;
    for (int i = tid; i < cfg.g_soft; i += kReduceThreads) a += eta_part[(size_t)e * cfg.g_soft + i];
    a = warp_sum(a);
    if ((tid & 31) == 0) red[tid >> 5] = a;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int i = 0; i < kReduceThreads / 32; ++i) s += red[i];
        out[0] = (double)rho[e]; out[1] = s;
    }
    // V_g[c] = sum over the weighted-sum kernel's blocks: thread (slice, c) adds every n_slice-th
    // block, then the slices are added in a fixed order
    const int C = 2 * cfg.T;
    const int n_slice = kReduceThreads / C;                 // >= 2 because C <= 512
    const int slice = tid / C, c = tid - slice * C;
    double s = 0.0;
    if (slice < n_slice) {
        const float* src = v_part + (size_t)e * cfg.g_wsum * C + c;
        for (int b = slice; b < n_wsum_blocks; b += n_slice) s += (double)src[(size_t)b * C];
        colsum[slice * C + c] = s;
    }
    __syncthreads();
    if (tid < C) {
        double t = colsum[tid];
        for (int i = 1; i < n_slice; ++i) t += colsum[i * C + tid];
        out[2 + tid] = t;
    }
    if (px.world > 0) px_put(px, out, cfg.n_env, e, 2 + 2 * cfg.T);
}

// ================================================================================================
// 6b. finalize as a kernel of its own (injected-noise mode, NCCL exchange, timed runs)
// ================================================================================================
__global__ void __launch_bounds__(256)
mppi_finalize_sm100a(DevCfg cfg, DevIo io, const double* gathered, int world, PeerExchange px) {
    __shared__ FinalizeSmem sm;
    finalize_env(cfg, io, blockIdx.x, gathered, world, px, sm);
}

// ================================================================================================
// 7. sampled trajectories (control.py:137-145): x <- F(x, v[k, t-1]) with the t = 0 wrap (Q4)
// ================================================================================================
template <int NOISE>
__global__ void __launch_bounds__(128)
mppi_sampled_traj_sm100a(DevCfg cfg, const uint64_t* __restrict__ step_ctr, const char* __restrict__ step_blocks,
                         const float* __restrict__ eps, const int32_t* __restrict__ subset, int n_rows,
                         float* __restrict__ traj) {
    // row i of the output is sample i (subset == nullptr, n_rows = K_local) or sample subset[e][i]
    const int e = blockIdx.y;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const int kl = subset ? subset[(size_t)e * n_rows + row] : row;
    if (kl < 0 || kl >= cfg.K_local) return;
    const StepBlockView sb = view_step_block((void*)(step_blocks + (size_t)e * cfg.step_block_bytes));
    const StepHeader hd = *sb.hd;
    const int kg = cfg.k_offset + kl, T = cfg.T;
    const float um = kg < cfg.n_exploit ? 1.0f : 0.0f;
    NoiseCfg nc = cfg.noise; nc.step = (uint32_t)(*step_ctr);
    ArmState st; arm_init(st, hd.q1, hd.q2, hd.d1, hd.d2, hd.a1, hd.a12);
    float4* out = (float4*)traj + ((size_t)e * n_rows + row) * T;
    for (int t = 0; t < T; ++t) {
        const int tc = t == 0 ? T - 1 : t - 1;
        float n1, n2;
        if (NOISE == 0) {
            float a0, a1, b0, b1;
            noise_pair(nc, (uint32_t)e, (uint32_t)kg, (uint32_t)tc >> 1, a0, a1, b0, b1);
            n1 = (tc & 1) ? b0 : a0; n2 = (tc & 1) ? b1 : a1;
        } else {
            const float2 v = __ldg((const float2*)eps + ((size_t)e * cfg.K_local + kl) * T + tc);
            n1 = v.x; n2 = v.y;
        }
        const StepCtl c = sb.ctl[tc];
        if (cfg.flags & 32) arm_step<1>(st, cfg.arm, fma_(um, c.u1, n1), fma_(um, c.u2, n2));
        else arm_step<0>(st, cfg.arm, fma_(um, c.u1, n1), fma_(um, c.u2, n2));
        out[t] = make_float4(st.q1, st.q2, st.d1, st.d2);
    }
}

// ================================================================================================
// 7b. plant tick of the device-resident closed loop (run.py:53-59 with utils.py:14-29, FP64):
//     u = first row of the shifted sequence (the reference's return value, quirk Q2),
//     dq += dt * Arm_Dynamic(q, dq, u); q += dt * dq; then the controller state for the next tick:
//     observed state, shifted sequence (control.py:148-149), waypoint index, step counter.
// ================================================================================================
__global__ void __launch_bounds__(32)
mppi_plant_sm100a(DevCfg cfg, DevIo io, const char* __restrict__ step_blocks, LoopParams* lp_ptr) {
    const int e = blockIdx.x, lane = threadIdx.x, T = cfg.T;
    const LoopParams lp = *lp_ptr;
    double* x0 = const_cast<double*>(io.x0) + 4 * e;
    double* u_prev = const_cast<double*>(io.u_prev) + (size_t)e * 2 * T;
    const double* u_new = io.u_new + (size_t)e * 2 * T;
    const StepHeader* hd = (const StepHeader*)(step_blocks + (size_t)e * cfg.step_block_bytes);
    const bool ended = (hd->status & 1) != 0;                   // control.py:76-78: the reference raises here
    const int t1 = T > 1 ? 1 : 0;
    const double u1 = u_new[2 * t1], u2 = u_new[2 * t1 + 1];
    if (!ended) {
        // shifted sequence: u_prev[:-1] = u[1:], u_prev[-1] = u[-1]
        for (int c = lane; c < 2 * T; c += 32) {
            const int t = c >> 1;
            u_prev[c] = u_new[2 * (t + 1 < T ? t + 1 : T - 1) + (c & 1)];
        }
    }
    if (lane == 0) {
        double q1 = x0[0], q2 = x0[1], d1 = x0[2], d2 = x0[3];
        if (!ended) {
            const double m1 = cfg.arm64[0], m2 = cfg.arm64[1], l1 = cfg.arm64[2], l2 = cfg.arm64[3],
                         lc1 = cfg.arm64[4], lc2 = cfg.arm64[5], g = cfg.arm64[6];
            const double c2 = cos(q2);
            const double M11 = m1 * lc1 * lc1 + l1 + m2 * (l1 * l1 + lc2 * lc2 + 2 * l1 * lc2 * c2) + l2;
            const double M22 = m2 * lc2 * lc2 + l2;
            const double M12 = m2 * l1 * lc2 * c2 + m2 * lc2 * lc2 + l2;
            const double hh = m2 * l1 * lc2 * sin(q2);
            const double g1 = m1 * lc1 * g * cos(q1) + m2 * g * (lc2 * cos(q1 + q2) + l1 * cos(q1));
            const double g2 = m2 * lc2 * g * cos(q1 + q2);
            const double b1 = u1 - ((-hh * d2) * d1 + (-hh * d1 - hh * d2) * d2) - g1;
            const double b2 = u2 - (hh * d1) * d1 - g2;
            const double det = M11 * M22 - M12 * M12;
            d1 += lp.plant_dt * ((M22 * b1 - M12 * b2) / det);
            d2 += lp.plant_dt * ((M11 * b2 - M12 * b1) / det);
            q1 += lp.plant_dt * d1;
            q2 += lp.plant_dt * d2;
            x0[0] = q1; x0[1] = q2; x0[2] = d1; x0[3] = d2;
        } else {
            atomicMin(lp.stop + e, lp.tick);
        }
        const_cast<int32_t*>(io.prev_idx)[e] = io.new_idx[e];   // control.py:230 (also when the path ended)
        if (lp.tick < lp.n_steps) {
            double* row = lp.log + ((size_t)lp.tick * cfg.n_env + e) * 8;
            row[0] = q1; row[1] = q2; row[2] = d1; row[3] = d2; row[4] = u1; row[5] = u2;
            row[6] = (double)io.new_idx[e]; row[7] = io.rho[e];
        }
        if (e == 0) {
            if (!(cfg.flags & 128)) *const_cast<uint64_t*>(io.step) += 1;   // next tick draws fresh Philox noise (resident state: prepare does it)
            lp_ptr->tick = lp.tick + 1;
        }
    }
}

// ================================================================================================
// 8. export of the Philox noise tensor (tests, and users who want to log the draw)
// ================================================================================================
__global__ void mppi_philox_export_sm100a(DevCfg cfg, uint32_t step, float* __restrict__ eps) {
    const int e = blockIdx.y;
    const int n_pairs = (cfg.T + 1) >> 1;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)cfg.K_local * n_pairs) return;
    const int kl = (int)(idx / n_pairs), pr = (int)(idx - (long long)kl * n_pairs);
    NoiseCfg nc = cfg.noise; nc.step = step;
    float a0, a1, b0, b1;
    noise_pair(nc, (uint32_t)e, (uint32_t)(cfg.k_offset + kl), (uint32_t)pr, a0, a1, b0, b1);
    float* dst = eps + (((size_t)e * cfg.K_local + kl) * cfg.T + 2 * pr) * 2;
    dst[0] = a0; dst[1] = a1;
    if (2 * pr + 1 < cfg.T) { dst[2] = b0; dst[3] = b1; }
}

// ================================================================================================
// 9. roofline probes: register-resident dependent chains, 8 independent chains per thread
// ================================================================================================
__global__ void __launch_bounds__(256) mppi_probe_fma_sm100a(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void __launch_bounds__(256) mppi_probe_mufu_sm100a(float* out, int iters) {
    float x0 = 1.0f + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x0));
            asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x1));
            asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x2));
            asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x3));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}

}  // namespace mppi

// mppi_cabi.cu — host side of libmppi_b200.so: the C ABI declared in include/mppi_b200.h.
// Orchestrates the sm_100a kernels of mppi_kernels.cuh for one control step
// (reference: /root/reference/control.py:67-152).  No torch types, no exceptions, no CPU fallback.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <new>
#include <mutex>
#include <nvtx3/nvToolsExt.h>        // header-only: ranges show up in Nsight Systems / ncu --nvtx, cost nothing otherwise

#include "../../include/mppi_b200.h"
#define MPPI_MAX_T_INTERNAL MPPI_MAX_T
#include "mppi_kernels.cuh"

using namespace mppi;

namespace {

constexpr int kNumTimers = 6;        // prepare, rollout, softmin, wsum, reduce, finalize
constexpr int kWinTableMaxRows = 131072;   // 320 MB of window tables at most
char g_create_error[512] = "";

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// NVTX range around the host-side enqueue of one stage of the step (the kernels of a graph replay inherit the
// names of their nodes; these ranges name the stages of plain launches and of the capture)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct Workspace {
    size_t bytes;
    size_t off_ref, off_step_blocks, off_in, off_out, off_S, off_w, off_block_min, off_eta_part,
        off_rho, off_v_part, off_partial, off_loop, off_eta_fused, off_tickets, off_seq, off_stats, off_win_table;
};

}  // namespace

// launch layout of the rollout kernel (pick_layout): samples per thread, and for the flat mixed layout the CTA counts
struct RollLayout { int ns; int n_wide, n_narrow; bool flat; };

struct MppiHandle {
    MppiConfig cfg;
    DevCfg dc;
    MppiIoLayout io;
    Workspace ws;
    char* dev;                 // workspace base
    char* host;                // pinned io block
    DevIo dio;                 // device mirrors of the io block fields
    size_t in_bytes, out_off, out_bytes;
    int sm_count;
    int n_ref_rows;
    size_t roll_smem;
    cudaEvent_t done;
    cudaEvent_t tev[kNumTimers + 1];
    bool timing, timing_pending;
    double t_acc[kNumTimers];
    int t_steps;
    uint64_t launches;
    cudaGraphExec_t graph_exec;
    cudaGraphExec_t tick_exec;   // one tick of the device closed loop
    void* tick_stream;
    uint64_t tick_kernels;
    void* graph_stream;
    uint64_t graph_kernels;
    bool have_step;            // a step has run (step blocks valid)
    bool const_window;         // single environment: window coefficients go through the constant bank
    int ns;                    // samples per thread of the rollout kernel
    int roll_threads;          // threads per CTA of the rollout kernel
    RollLayout layout;         // launch layout of the rollout kernel
    bool zero_copy;            // kernels read / write the caller's pinned block directly (no memcpy nodes)
    DevIo dio_dev;             // same as dio but never touching the pinned block (device closed loop)
    PeerExchange px;           // peer-memory exchange (world == 0: not configured)
    cudaGraphExec_t sharded_exec;
    void* sharded_stream;
    uint64_t sharded_kernels;
    bool capture_mode;         // caller is capturing: enqueue capturable work only
    uint64_t capture_kernels;  // kernels enqueued while capture mode was on (= per replay)
    cudaEvent_t const_ev;      // recorded after this handle's last reader of the constant-bank window
    uint64_t step_scratch;     // source of the step-counter upload of a resident handle
    bool pdl;                  // programmatic dependent launch of the rollout and weight-sum kernels
    char err[512];
};

namespace {

int fail(MppiHandle* h, int code, const char* fmt, const char* detail) {
    char* dst = h ? h->err : g_create_error;
    snprintf(dst, 512, fmt, detail);
    return code;
}
#define CU(h, call)                                                                        \
    do {                                                                                   \
        cudaError_t _e = (call);                                                           \
        if (_e != cudaSuccess) {                                                           \
            char _b[384];                                                                  \
            snprintf(_b, sizeof(_b), "%s -> %s", #call, cudaGetErrorString(_e));           \
            return fail(h, MPPI_ERR_CUDA, "%s", _b);                                       \
        }                                                                                  \
    } while (0)

bool valid_cfg(const MppiConfig* c, const char** why) {
    if (!c) { *why = "null config"; return false; }
    if (c->abi_version != MPPI_ABI_VERSION) { *why = "abi_version mismatch"; return false; }
    if (c->n_env < 1) { *why = "n_env < 1"; return false; }
    if (c->T < 1 || c->T > MPPI_MAX_T) { *why = "T out of range [1, MPPI_MAX_T]"; return false; }
    if (c->K_total < 1 || c->K_local < 1 || c->k_offset < 0 || c->k_offset + c->K_local > c->K_total) {
        *why = "sample shard [k_offset, k_offset+K_local) not inside [0, K_total)"; return false;
    }
    if (c->max_ref_rows < 2) { *why = "max_ref_rows < 2"; return false; }
    if (!(c->param_lambda > 0.0)) { *why = "param_lambda must be > 0"; return false; }
    if (!(c->joint_limit_weight >= 0.0)) { *why = "joint_limit_weight must be >= 0"; return false; }
    for (int i = 0; i < 4; ++i)        // the stage cost is evaluated on residuals scaled by sqrt(weight)
        if (!(c->stage_cost_weight[i] >= 0.0) || !(c->stage_cost_weight[i] < 1e30)) { *why = "stage_cost_weight entries must be finite and >= 0"; return false; }
    if (c->joint_limit_weight > 0.0) {
        if (!(c->joint_limit_lo[0] <= c->joint_limit_hi[0]) || !(c->joint_limit_lo[1] <= c->joint_limit_hi[1])) {
            *why = "joint limits need lo <= hi"; return false;
        }
        if (c->flags & MPPI_FLAG_DYNAMICS_F1) { *why = "the joint-limit cost is built for the rollout model _F only"; return false; }
    }
    return true;
}

// samples per thread of the rollout kernel: 2 when there is enough work to fill the GPU that way
bool certified_kernels(const MppiConfig* c) {
    // certified lookups (window table in shared memory) unless the caller asked for plain searches; the _F1
    // model and the joint-limit cost exist in the certified shape only (with MPPI_FLAG_FULL_SEARCH the certificate is never armed)
    return !(c->flags & MPPI_FLAG_FULL_SEARCH) || (c->flags & MPPI_FLAG_DYNAMICS_F1) || c->joint_limit_weight > 0.0;
}
bool pick_const_window(const MppiConfig* c) {
    // constant-bank window of the plain-search kernels: single environment and enough work to pay for the copy node
    return !certified_kernels(c) && (c->n_env == 1) && ((long long)c->K_local * c->T >= (1ll << 21)) &&
           getenv("MPPI_NO_CONST_WINDOW") == nullptr;
}

// Layout of the rollout launch (kernels: mppi_kernels.cuh, section 2).
//  * Kernels without certified lookups: uniform layouts; one or kNsWide samples per thread by a wave model — CTAs
//    of 4 warps (one per sub-partition), `per_sm` resident CTAs per SM; full waves cost per_sm warp-times each, the
//    last partial wave ceil(rest / SMs) warp-times; a warp-time is proportional to ns (x 1.03 for ns = 1).
//  * Certified kernels: work is counted in units of 128 samples (U over all environments).  U <= one wave of
//    one-sample CTAs: the one-sample kernel (a lone one-sample warp is the shortest pass there is).  Otherwise the
//    flat mixed layout of the two-sample kernel: full waves of two-sample CTAs, and a LAST wave that is exactly full —
//    R units left for S slots: R <= S one-sample CTAs, else (R - S) two-sample + (2S - R) one-sample CTAs.
//    Measured on B200 (profiles/r2s3_layout_sweep.txt): 98304 samples 82.8 -> 74.3 us, 196608 samples 137.4 -> 128.7 us.
int pick_ns_uniform(const MppiConfig* c, int sm) {
    const bool cw = pick_const_window(c);
    double best_cost = 0.0; int best = 1;
    for (int ns = 1; ns <= kNsWide; ns += kNsWide - 1) {
        const int per_sm = certified_kernels(c) ? (ns == 1 ? MPPI_ROLL_MIN_BLOCKS_CERT_NS1 : MPPI_ROLL_MIN_BLOCKS_CERT)
                                                : (cw ? MPPI_ROLL_MIN_BLOCKS_CONST : (ns == 1 ? 3 : 2));
        const long long ctas = (((long long)c->K_local + 128 * ns - 1) / (128 * ns)) * c->n_env;
        const long long slots = (long long)sm * per_sm;
        const long long full = ctas / slots, rest = ctas % slots;
        const double warp_time = ns == 1 ? 1.03 : (double)ns;
        const double cost = (double)(full * per_sm + (rest + sm - 1) / sm) * warp_time;
        if (ns == 1 || cost < best_cost) { best_cost = cost; best = ns; }
    }
    return best;
}
RollLayout pick_layout(const MppiConfig* c, int sm) {
    RollLayout L = { 1, 0, 0, false };
    int forced = 0;
    if (const char* f = getenv("MPPI_NS")) { const int v = atoi(f); if (v == 1 || v == kNsWide) forced = v; }
    if (!certified_kernels(c) || kNsWide != 2) {
        L.ns = forced ? forced : pick_ns_uniform(c, sm);
        if (L.ns != 1 && certified_kernels(c)) {       // (the certified wide kernel only knows the flat layout)
            const long long upe = ((long long)c->K_local + 127) / 128, U = upe * c->n_env;
            if (c->n_env == 1 || c->K_local % (128 * kNsWide) == 0) { L.flat = true; L.n_wide = (int)(U / kNsWide); L.n_narrow = (int)(U % kNsWide); }
            else L.ns = 1;
        }
        return L;
    }
    const long long upe = ((long long)c->K_local + 127) / 128, U = upe * c->n_env;
    const bool flat_ok = (c->n_env == 1 || c->K_local % 256 == 0) && U < (1ll << 30);
    const long long S = (long long)sm * MPPI_ROLL_MIN_BLOCKS_CERT, Sn = (long long)sm * MPPI_ROLL_MIN_BLOCKS_CERT_NS1;
    // one wave of one-sample CTAs, or (measured, 131072 samples: 92.9 us against 95.2 mixed) a shard whose mixed single
    // wave would be mostly two-sample CTAs: the one-sample kernel
    if (forced == 1 || !flat_ok || (!forced && (U <= Sn || (U <= 2 * S && 2 * U > 3 * S)))) return L;
    L.ns = 2; L.flat = true;
    if (forced == 2 || getenv("MPPI_NO_MIXED")) { L.n_wide = (int)(U / 2); L.n_narrow = (int)(U % 2); return L; }
    const long long full = U / (2 * S), R = U % (2 * S);
    if (R == 0) { L.n_wide = (int)(U / 2); L.n_narrow = 0; }
    else if (R <= S) { L.n_wide = (int)(full * S); L.n_narrow = (int)R; }
    else { L.n_wide = (int)(full * S + (R - S)); L.n_narrow = (int)(2 * S - R); }
    return L;
}

// Threads per CTA of the rollout kernel: 128, or fewer for small shards, where finer CTAs spread more evenly over
// the SMs (MPPI_ROLL_THREADS overrides; A/B in profiles/r2_variants.md).
int pick_roll_threads(const MppiConfig* c, int sm) {
    if (const char* f = getenv("MPPI_ROLL_THREADS")) { const int v = atoi(f); if (v == 32 || v == 64 || v == 128) return v; }
    (void)c; (void)sm;
    return kRollThreads;
}

void grid_sizes(const MppiConfig* c, int sm, int* g_roll, int* g_soft, int* g_wsum) {
    const int K = c->K_local;
    const RollLayout L = pick_layout(c, sm);
    const int thr = pick_roll_threads(c, sm);
    // uniform layout: blocks per environment; flat layout: units of 128 samples per environment
    int gr = L.flat ? (K + 127) / 128 : (K + thr * L.ns - 1) / (thr * L.ns);
    if (!L.flat && gr > 32768) gr = 32768;
    *g_roll = gr;
    int gs = (K + kSoftThreads * 4 - 1) / (kSoftThreads * 4);
    const int cap = (4 * sm + c->n_env - 1) / c->n_env;
    if (gs > cap) gs = cap;
    if (gs < 1) gs = 1;
    *g_soft = gs;
    // weighted-sum grid (upper bound; the Philox variant launches fewer blocks): enough CTAs to keep
    // ~8 CTAs per SM streaming the injected noise tensor
    int gw = (K + 63) / 64;
    const int capw = (8 * sm + c->n_env - 1) / c->n_env;
    if (gw > capw) gw = capw;
    if (gw < 1) gw = 1;
    *g_wsum = gw;
}

void carve(const MppiConfig* c, int sm, Workspace* w) {
    int g_roll, g_soft, g_wsum;
    grid_sizes(c, sm, &g_roll, &g_soft, &g_wsum);
    const size_t E = c->n_env, K = c->K_local, T = c->T;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    w->off_ref = take((size_t)c->max_ref_rows * 4 * sizeof(double));
    w->off_step_blocks = take(E * (kStepBlockFixed + 16 * T));
    MppiIoLayout io; mppi_io_layout(c, &io);
    w->off_in = take(io.off_new_idx);                 // inputs occupy [0, off_new_idx)
    w->off_out = take(io.bytes - io.off_new_idx);
    w->off_S = take(E * K * sizeof(float));
    w->off_w = take(E * K * sizeof(float));
    w->off_block_min = take(E * g_roll * sizeof(float));
    w->off_eta_part = take(E * g_soft * sizeof(double));
    w->off_rho = take(E * sizeof(float));
    w->off_v_part = take(E * g_wsum * 2 * T * sizeof(float));
    w->off_partial = take(E * (2 + 2 * T) * sizeof(double));
    w->off_loop = take(sizeof(LoopParams));
    w->off_eta_fused = take(E * g_wsum * sizeof(double));
    w->off_tickets = take(E * sizeof(unsigned int));
    // [0] step sequence number, [1] exchange status, then one 32-bit minimum-cost key per environment
    w->off_seq = take(2 * sizeof(unsigned long long) + E * sizeof(unsigned int));
    w->off_stats = take(4 * sizeof(unsigned long long));     // warp-lookups: [0] answered by an end test, [1] all, [2] by a certified triple
    // window tables + certificates of every window start of the path (2.4 KB per waypoint), built by
    // mppi_set_ref_path(); paths beyond kWinTableMaxRows build the window on the spot in every step instead
    w->off_win_table = take(c->max_ref_rows <= kWinTableMaxRows ? (size_t)c->max_ref_rows * kWinBytes : 0);
    w->bytes = off;
}

int sm_count_or_default(int device) {
    int sm = 0;
    if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sm <= 0) {
        cudaGetLastError();
        sm = 148;                                     // B200; only used for sizing without a device
    }
    return sm;
}

void fill_dev_cfg(MppiHandle* h) {
    const MppiConfig& c = h->cfg;
    DevCfg& d = h->dc;
    const double m1 = c.arm[0], m2 = c.arm[1], l1 = c.arm[2], l2 = c.arm[3], lc1 = c.arm[4], lc2 = c.arm[5],
                 g = c.arm[6];
    // control.py:241-249 with the constant sub-expressions folded in FP64
    d.arm.A0 = (float)(m1 * lc1 * lc1 + l1 + m2 * (l1 * l1 + lc2 * lc2) + l2);
    d.arm.A1 = (float)(2 * m2 * l1 * lc2);
    d.arm.M22 = (float)(m2 * lc2 * lc2 + l2);
    d.arm.B1 = (float)(m2 * l1 * lc2);
    d.arm.G1a = (float)((m1 * lc1 + m2 * l1) * g);
    d.arm.G1b = (float)(m2 * lc2 * g);
    d.arm.dt = (float)c.delta_t;
    d.arm.dtfix = arm_dtfix(c.delta_t);
    d.arm.L1 = (float)c.cost_l1; d.arm.L2 = (float)c.cost_l2;
    d.cost.s0 = (float)(c.stage_cost_weight[0] * 1e4); d.cost.s1 = (float)(c.stage_cost_weight[1] * 1e4);
    d.cost.s2 = (float)(c.stage_cost_weight[2] * 1e4); d.cost.s3 = (float)(c.stage_cost_weight[3] * 1e4);
    d.cost.t0 = (float)(c.terminal_cost_weight[0] * 1e4); d.cost.t1 = (float)(c.terminal_cost_weight[1] * 1e4);
    d.cost.t2 = (float)(c.terminal_cost_weight[2] * 1e4); d.cost.t3 = (float)(c.terminal_cost_weight[3] * 1e4);
    cost_roots(d.cost);
    {   // joint limits: +-inf would turn into NaN in (q - hi)^2 * 0; clamp the bounds to a huge finite value
        auto lim = [](double v) { return (float)(v > 1e30 ? 1e30 : (v < -1e30 ? -1e30 : v)); };
        d.cost.jw = (float)(c.joint_limit_weight * 1e4);
        d.cost.lo1 = lim(c.joint_limit_lo[0]); d.cost.hi1 = lim(c.joint_limit_hi[0]);
        d.cost.lo2 = lim(c.joint_limit_lo[1]); d.cost.hi2 = lim(c.joint_limit_hi[1]);
    }
    d.noise.key = philox_expand_key((uint32_t)(c.seed & 0xffffffffu), (uint32_t)(c.seed >> 32));
    d.noise.step = 0;
    d.noise.L11 = (float)c.sigma_chol[0]; d.noise.L21 = (float)c.sigma_chol[2]; d.noise.L22 = (float)c.sigma_chol[3];
    d.K_local = c.K_local; d.K_total = c.K_total; d.k_offset = c.k_offset; d.T = c.T; d.n_env = c.n_env;
    d.n_exploit = c.n_exploit; d.n_ref_rows = 0; d.flags = c.flags;
    d.step_block_bytes = kStepBlockFixed + 16 * c.T;
    grid_sizes(&c, h->sm_count, &d.g_roll, &d.g_soft, &d.g_wsum);
    d.gamma = c.param_gamma; d.lambda = c.param_lambda; d.inv_lambda = 1.0 / c.param_lambda;
    for (int i = 0; i < 4; ++i) d.sig_inv[i] = c.sigma_inv[i];
    d.cost_l1 = c.cost_l1; d.cost_l2 = c.cost_l2;
    for (int i = 0; i < 7; ++i) d.arm64[i] = c.arm[i];
}

// The constant-bank window table is one symbol per device, shared by every handle of the process.
// Work of one handle is stream-ordered (copy -> rollout); work of *different* handles is ordered on
// the device by making the newcomer's stream wait for the previous owner's last reader.
struct ConstOwner { MppiHandle* h; cudaEvent_t ev; };
ConstOwner g_const_owner[64];
std::mutex g_const_mu;

int const_acquire(MppiHandle* h, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_const_mu);
    ConstOwner& o = g_const_owner[h->cfg.device & 63];
    if (o.h && o.h != h) CU(h, cudaStreamWaitEvent(s, o.ev, 0));
    return MPPI_OK;
}
int const_release(MppiHandle* h, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_const_mu);
    CU(h, cudaEventRecord(h->const_ev, s));
    g_const_owner[h->cfg.device & 63] = ConstOwner{h, h->const_ev};
    return MPPI_OK;
}
void const_forget(MppiHandle* h) {
    std::lock_guard<std::mutex> lk(g_const_mu);
    ConstOwner& o = g_const_owner[h->cfg.device & 63];
    if (o.h == h) { cudaEventSynchronize(o.ev); o.h = nullptr; }
}

// How one step is enqueued.
struct StepOpts {
    bool timed = false;        // record the per-kernel timing events
    bool capturing = false;    // inside a stream capture: no event records, no cross-handle waits
    bool host_io = true;       // read inputs from / deliver outputs to the caller's pinned block
                               // (false: device mirror only — the ticks of the device closed loop)
    bool use_px = false;       // exchange the partial triple through peer memory (sharded step)
    bool record_done = true;   // record the completion event after the step
    bool fuse_finalize = false;   // combine / filter / update run in the last block of the fused weight-sum kernel
};

// enqueue everything up to this shard's partial triple
int enqueue_local(MppiHandle* h, int noise_mode, const float* eps_dev, double* partial_dev, cudaStream_t s,
                  const StepOpts& o) {
    const bool timed = o.timed, capturing = o.capturing, copy_inputs = o.host_io, use_px = o.use_px;
    const DevCfg& dc = h->dc;
    char* ws = h->dev;
    if (noise_mode != MPPI_NOISE_PHILOX && noise_mode != MPPI_NOISE_INJECTED)
        return fail(h, MPPI_ERR_INVALID, "%s", "unknown noise mode");
    if (noise_mode == MPPI_NOISE_INJECTED && !eps_dev)
        return fail(h, MPPI_ERR_INVALID, "%s", "injected noise mode needs eps_dev");
    if (h->n_ref_rows < 2) return fail(h, MPPI_ERR_INVALID, "%s", "mppi_set_ref_path() has not been called");
    const uint64_t* step_ctr = (const uint64_t*)(ws + h->ws.off_in + h->io.off_step);
    char* step_blocks = ws + h->ws.off_step_blocks;
    float* S = (float*)(ws + h->ws.off_S);
    float* w = (float*)(ws + h->ws.off_w);
    float* bmin = (float*)(ws + h->ws.off_block_min);
    double* eta_part = (double*)(ws + h->ws.off_eta_part);
    float* rho = (float*)(ws + h->ws.off_rho);
    float* v_part = (float*)(ws + h->ws.off_v_part);
    const double* ref = (const double*)(ws + h->ws.off_ref);

    // host_io = false: the step works on the device mirror only (ticks of the device closed loop)
    const bool resident = (dc.flags & MPPI_FLAG_RESIDENT_STATE) != 0;
    const int pull = (h->zero_copy && copy_inputs) ? (resident ? 1 : 3) : 0;
    const DevIo& dio = copy_inputs ? h->dio : h->dio_dev;
    if (copy_inputs && !h->zero_copy) {
        if (resident)        // only the observed states travel; the controller state lives on the device
            CU(h, cudaMemcpyAsync(ws + h->ws.off_in + h->io.off_x0, h->host + h->io.off_x0,
                                  (size_t)dc.n_env * 4 * sizeof(double), cudaMemcpyHostToDevice, s));
        else
            CU(h, cudaMemcpyAsync(ws + h->ws.off_in, h->host, h->in_bytes, cudaMemcpyHostToDevice, s));
    }
    if (timed) CU(h, cudaEventRecord(h->tev[0], s));
    {
        NvtxRange r("mppi.prepare");
        mppi_prepare_sm100a<<<dc.n_env, 32, 0, s>>>(dc, dio, ref, step_blocks, pull,
                                                    (unsigned long long*)(ws + h->ws.off_seq),
                                                    h->cfg.max_ref_rows <= kWinTableMaxRows ? ws + h->ws.off_win_table : nullptr);
    }
    if (timed) CU(h, cudaEventRecord(h->tev[1], s));
    {
        NvtxRange r("mppi.rollout");
        const RollLayout& L = h->layout;
        dim3 grid(dc.g_roll, dc.n_env);
        if (L.flat) grid = dim3(L.n_wide + L.n_narrow, 1);
        const int n_wide = L.n_wide;
        const bool ph = noise_mode == MPPI_NOISE_PHILOX;
        if (h->const_window) {
            // single environment, large K: stage this step's window coefficients in the constant bank
            if (!capturing) { int rc = const_acquire(h, s); if (rc != MPPI_OK) return rc; }
            // (window + rows + pair table are adjacent in the step block and in the constant bank: one copy)
            CU(h, cudaMemcpyToSymbolAsync(c_window, step_blocks + 64, sizeof(ConstWindow), 0,
                                          cudaMemcpyDeviceToDevice, s));
        }
        // kernel specialisations: noise source x samples per thread x { certified lookups x rollout model,
        // plain searches x window policy }.  MPPI_FLAG_FULL_SEARCH selects the kernels compiled without the
        // certificate; the _F1 model exists in the certified shape only (prepare then never arms the certificate).
        const bool ns2 = h->ns != 1, f1 = (dc.flags & MPPI_FLAG_DYNAMICS_F1) != 0;
        const bool cert = certified_kernels(&h->cfg), jl = h->cfg.joint_limit_weight > 0.0;
        unsigned long long* stats = (unsigned long long*)(ws + h->ws.off_stats);
        unsigned int* rho_key = (unsigned int*)(ws + h->ws.off_seq + 2 * sizeof(unsigned long long));
        const float* eps_arg = ph ? nullptr : eps_dev;
        // Programmatic dependent launch: the rollout CTAs are scheduled while the prepare kernel still runs and wait
        // (griddepcontrol.wait) just before they read the step block — the launch latency leaves the critical path.
        // (Not with the constant-bank window: the copy node sits between the two kernels.)
        cudaLaunchAttribute pdl_attr[1];
        pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t lc = {};
        lc.gridDim = grid; lc.blockDim = dim3(L.flat ? kRollThreads : h->roll_threads); lc.dynamicSmemBytes = h->roll_smem; lc.stream = s;
        lc.attrs = pdl_attr; lc.numAttrs = (h->pdl && !h->const_window && !timed) ? 1 : 0;
#define MPPI_ROLL(NOISE, CW, NS_, DYN, CERT) do { \
        if (CERT && DYN == 0 && jl) CU(h, cudaLaunchKernelEx(&lc, mppi_rollout_sm100a<NOISE, false, NS_, 0, true, true>, dc, step_ctr, (const char*)step_blocks, eps_arg, S, bmin, stats, rho_key, n_wide)); \
        else CU(h, cudaLaunchKernelEx(&lc, mppi_rollout_sm100a<NOISE, CW, NS_, DYN, CERT>, dc, step_ctr, (const char*)step_blocks, eps_arg, S, bmin, stats, rho_key, n_wide)); } while (0)
#define MPPI_ROLL_NS(NOISE, CW, DYN, CERT) do { if (ns2) MPPI_ROLL(NOISE, CW, kNsWide, DYN, CERT); else MPPI_ROLL(NOISE, CW, 1, DYN, CERT); } while (0)
#define MPPI_ROLL_NOISE(CW, DYN, CERT) do { if (ph) MPPI_ROLL_NS(0, CW, DYN, CERT); else MPPI_ROLL_NS(1, CW, DYN, CERT); } while (0)
        if (f1) MPPI_ROLL_NOISE(false, 1, true);
        else if (cert) MPPI_ROLL_NOISE(false, 0, true);
        else if (h->const_window) MPPI_ROLL_NOISE(true, 0, false);
        else MPPI_ROLL_NOISE(false, 0, false);
        if (h->const_window && !capturing) { int rc = const_release(h, s); if (rc != MPPI_OK) return rc; }
#undef MPPI_ROLL_NOISE
#undef MPPI_ROLL_NS
#undef MPPI_ROLL
    }
    if (timed) CU(h, cudaEventRecord(h->tev[2], s));
    NvtxRange r_w(use_px ? "mppi.weights+sum+exchange" : "mppi.weights+sum");
    if (noise_mode == MPPI_NOISE_PHILOX) {
        // fused: soft-min weights + weighted sum + this GPU's partial triple
        // 2048 samples per block for large K (rows of blocks without a non-zero weight are never read back, so many
        // blocks cost nothing); small K is latency bound: up to 64 blocks of >= 512 samples
        int g = (dc.K_local + kWsumSamplesPerBlock - 1) / kWsumSamplesPerBlock;
        const int g_small = (dc.K_local + 511) / 512 < 64 ? (dc.K_local + 511) / 512 : 64;
        if (g < g_small) g = g_small;
        if (g > dc.g_wsum) g = dc.g_wsum;
        if (g > kWsumMaxBlocks) g = kWsumMaxBlocks;
        const size_t sm = (size_t)(kWsumThreads / 32) * ((dc.T + 1) / 2) * sizeof(float4);
        cudaLaunchAttribute pdl_attr[1];
        pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(g, dc.n_env); lc.blockDim = dim3(kWsumThreads); lc.dynamicSmemBytes = sm; lc.stream = s;
        lc.attrs = pdl_attr; lc.numAttrs = (h->pdl && !timed) ? 1 : 0;      // (timed: an event record sits in between)
        CU(h, cudaLaunchKernelEx(&lc, mppi_softmin_wsum_philox_sm100a, dc, step_ctr, (const float*)S,
                                 (const unsigned int*)(ws + h->ws.off_seq + 2 * sizeof(unsigned long long)), w,
                                 (double*)(ws + h->ws.off_eta_fused), v_part, (unsigned int*)(ws + h->ws.off_tickets), rho,
                                 partial_dev, use_px ? h->px : PeerExchange{}, dio, o.fuse_finalize ? 1 : 0));
        if (timed) { CU(h, cudaEventRecord(h->tev[3], s)); CU(h, cudaEventRecord(h->tev[4], s)); }
        h->launches += 3;
    } else {
        mppi_softmin_sm100a<<<dim3(dc.g_soft, dc.n_env), kSoftThreads, 0, s>>>(dc, S, bmin, w, eta_part, rho);
        if (timed) CU(h, cudaEventRecord(h->tev[3], s));
        dim3 grid(dc.g_wsum, dc.n_env);
        if ((dc.T & 1) == 0 && (((uintptr_t)eps_dev) & 15) == 0)
            mppi_wsum_injected_sm100a<float4><<<grid, kWsumThreads, kWsumThreads * sizeof(float4), s>>>(dc, w, eps_dev, v_part);
        else
            mppi_wsum_injected_sm100a<float2><<<grid, kWsumThreads, kWsumThreads * sizeof(float2), s>>>(dc, w, eps_dev, v_part);
        if (timed) CU(h, cudaEventRecord(h->tev[4], s));
        mppi_reduce_sm100a<<<dc.n_env, kReduceThreads, 0, s>>>(dc, dc.g_wsum, rho, eta_part, v_part, partial_dev,
                                                               use_px ? h->px : PeerExchange{});
        h->launches += 5;
    }
    if (timed) CU(h, cudaEventRecord(h->tev[5], s));
    CU(h, cudaGetLastError());
    h->have_step = true;
    return MPPI_OK;
}

int enqueue_combine(MppiHandle* h, const double* gathered_dev, int world, cudaStream_t s, const StepOpts& o) {
    const bool timed = o.timed, record_done = o.record_done && !o.capturing, copy_outputs = o.host_io, use_px = o.use_px;
    if (world < 1 || world > 64) return fail(h, MPPI_ERR_INVALID, "%s", "world must be in [1, 64]");
    NvtxRange r("mppi.combine+update");
    if (!o.fuse_finalize) {      // (fused: the last block of the weight-sum kernel has done this already)
        mppi_finalize_sm100a<<<h->dc.n_env, 256, 0, s>>>(h->dc, copy_outputs ? h->dio : h->dio_dev, gathered_dev, world,
                                                         use_px ? h->px : PeerExchange{});
        h->launches += 1;
    }
    if (timed) CU(h, cudaEventRecord(h->tev[6], s));
    CU(h, cudaGetLastError());
    if (copy_outputs && !h->zero_copy) {
        const size_t n = (h->dc.flags & MPPI_FLAG_RESIDENT_STATE) ? h->io.off_w_eps_raw - h->io.off_new_idx : h->out_bytes;
        CU(h, cudaMemcpyAsync(h->host + h->out_off, h->dev + h->ws.off_out, n, cudaMemcpyDeviceToHost, s));
    }
    if (record_done) CU(h, cudaEventRecord(h->done, s));      // not inside a stream capture
    return MPPI_OK;
}

}  // namespace

extern "C" {

int mppi_abi_version(void) { return MPPI_ABI_VERSION; }

int mppi_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

int mppi_io_layout(const MppiConfig* c, MppiIoLayout* o) {
    const char* why = nullptr;
    if (!o || !valid_cfg(c, &why)) return fail(nullptr, MPPI_ERR_INVALID, "%s", why ? why : "null layout");
    const size_t E = c->n_env, T = c->T;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t r = off; off = align_up(off + bytes, 16); return r; };
    o->off_x0 = take(E * 4 * sizeof(double));
    o->off_u_prev = take(E * T * 2 * sizeof(double));
    o->off_prev_idx = take(E * sizeof(int32_t));
    o->off_step = take(sizeof(uint64_t));
    off = align_up(off, 256);
    o->off_new_idx = take(E * sizeof(int32_t));
    o->off_status = take(E * sizeof(int32_t));
    o->off_rho = take(E * sizeof(double));
    o->off_eta = take(E * sizeof(double));
    o->off_u0 = take(E * 2 * sizeof(double));
    o->off_w_eps_raw = take(E * T * 2 * sizeof(double));
    o->off_w_eps_filt = take(E * T * 2 * sizeof(double));
    o->off_u_new = take(E * T * 2 * sizeof(double));
    o->off_opt_traj = take(E * T * 4 * sizeof(double));
    o->bytes = align_up(off, 256);
    return MPPI_OK;
}

size_t mppi_workspace_bytes(const MppiConfig* c) {
    const char* why = nullptr;
    if (!valid_cfg(c, &why)) { fail(nullptr, MPPI_ERR_INVALID, "%s", why); return 0; }
    Workspace w;
    carve(c, sm_count_or_default(c->device), &w);
    return w.bytes;
}

int mppi_create(const MppiConfig* c, void* workspace, size_t workspace_bytes, void* io_host, size_t io_bytes,
                MppiHandle** out) {
    const char* why = nullptr;
    if (!out) return fail(nullptr, MPPI_ERR_INVALID, "%s", "null out");
    *out = nullptr;
    if (!valid_cfg(c, &why)) return fail(nullptr, MPPI_ERR_INVALID, "%s", why);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || c->device < 0 || c->device >= ndev) {
        cudaGetLastError();
        return fail(nullptr, MPPI_ERR_NO_DEVICE, "%s",
                    "no usable CUDA device: libmppi_b200 is sm_100a only and has no CPU fallback");
    }
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, c->device);
    if (major != 10)
        return fail(nullptr, MPPI_ERR_NO_DEVICE, "%s", "device is not compute capability 10.x (B200, sm_100a)");
    MppiHandle* h = new (std::nothrow) MppiHandle();
    if (!h) return fail(nullptr, MPPI_ERR_INVALID, "%s", "out of host memory");
    memset(h, 0, sizeof(*h));
    h->cfg = *c;
    CU(nullptr, cudaSetDevice(c->device));
    h->sm_count = sm_count_or_default(c->device);
    mppi_io_layout(c, &h->io);
    carve(c, h->sm_count, &h->ws);
    if (!workspace || workspace_bytes < h->ws.bytes || ((uintptr_t)workspace & 255) != 0) {
        delete h;
        return fail(nullptr, MPPI_ERR_WORKSPACE, "%s", "workspace missing, too small or not 256-byte aligned");
    }
    if (!io_host || io_bytes < h->io.bytes) {
        delete h;
        return fail(nullptr, MPPI_ERR_WORKSPACE, "%s", "io block missing or too small");
    }
    h->dev = (char*)workspace;
    h->host = (char*)io_host;
    h->in_bytes = h->io.off_new_idx;
    h->out_off = h->io.off_new_idx;
    h->out_bytes = h->io.bytes - h->io.off_new_idx;
    fill_dev_cfg(h);
    char* din = h->dev + h->ws.off_in;
    char* dout = h->dev + h->ws.off_out - h->io.off_new_idx;      // same offsets as the host block
    h->dio.x0 = (const double*)(din + h->io.off_x0);
    h->dio.u_prev = (const double*)(din + h->io.off_u_prev);
    h->dio.prev_idx = (const int32_t*)(din + h->io.off_prev_idx);
    h->dio.step = (const uint64_t*)(din + h->io.off_step);
    h->dio.new_idx = (int32_t*)(dout + h->io.off_new_idx);
    h->dio.status = (int32_t*)(dout + h->io.off_status);
    h->dio.rho = (double*)(dout + h->io.off_rho);
    h->dio.eta = (double*)(dout + h->io.off_eta);
    h->dio.u0 = (double*)(dout + h->io.off_u0);
    h->dio.w_eps_raw = (double*)(dout + h->io.off_w_eps_raw);
    h->dio.w_eps_filt = (double*)(dout + h->io.off_w_eps_filt);
    h->dio.u_new = (double*)(dout + h->io.off_u_new);
    h->dio.opt_traj = (double*)(dout + h->io.off_opt_traj);
    h->roll_smem = (size_t)h->dc.step_block_bytes;
    h->layout = pick_layout(c, h->sm_count);
    h->ns = h->layout.ns;
    h->roll_threads = pick_roll_threads(c, h->sm_count);
    h->const_window = pick_const_window(c);
    h->dio_dev = h->dio;
    h->dio_dev.host_in = nullptr; h->dio_dev.in_delta = 0; h->dio_dev.out_delta = 0;
    h->dio.host_in = nullptr; h->dio.in_delta = 0; h->dio.out_delta = 0;
    {   // zero-copy io: let the kernels touch the caller's pinned block directly (UVA-mapped)
        void* dptr = nullptr;
        // (only while the block is small: for many environments a DMA copy beats PCIe loads/stores)
        if (getenv("MPPI_NO_ZERO_COPY") == nullptr && h->io.bytes <= 65536 &&
            cudaHostGetDevicePointer(&dptr, io_host, 0) == cudaSuccess && dptr != nullptr) {
            h->zero_copy = true;
            h->dio.host_in = (const char*)dptr;
            h->dio.in_delta = (char*)din - (char*)dptr;
            h->dio.out_delta = (char*)dptr - dout;
        } else {
            cudaGetLastError();
        }
    }
    h->pdl = getenv("MPPI_NO_PDL") == nullptr;
    h->px.world = 0;
    h->px.timeout_ns = 3000000000ull;
    h->px.seq = (const unsigned long long*)(h->dev + h->ws.off_seq);
    if (cudaMemset(h->dev + h->ws.off_seq, 0, 2 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(h->dev + h->ws.off_stats, 0, 4 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(h->dev + h->ws.off_tickets, 0, sizeof(unsigned int) * c->n_env) != cudaSuccess) {
        snprintf(g_create_error, sizeof(g_create_error), "cudaMemset(tickets) failed");
        delete h;
        return MPPI_ERR_CUDA;
    }
    cudaError_t e = cudaEventCreateWithFlags(&h->done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->const_ev, cudaEventDisableTiming);
    for (int i = 0; i <= kNumTimers && e == cudaSuccess; ++i) e = cudaEventCreate(&h->tev[i]);
    if (e != cudaSuccess) {
        snprintf(g_create_error, sizeof(g_create_error), "cudaEventCreate -> %s", cudaGetErrorString(e));
        delete h;
        return MPPI_ERR_CUDA;
    }
    *out = h;
    return MPPI_OK;
}

void mppi_destroy(MppiHandle* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    const_forget(h);
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    if (h->tick_exec) cudaGraphExecDestroy(h->tick_exec);
    if (h->sharded_exec) cudaGraphExecDestroy(h->sharded_exec);
    if (h->done) cudaEventDestroy(h->done);
    if (h->const_ev) cudaEventDestroy(h->const_ev);
    for (int i = 0; i <= kNumTimers; ++i)
        if (h->tev[i]) cudaEventDestroy(h->tev[i]);
    delete h;
}

const char* mppi_last_error(const MppiHandle* h) { return h ? h->err : g_create_error; }

int mppi_set_ref_path(MppiHandle* h, const double* ref, int32_t n_rows) {
    if (!h) return MPPI_ERR_INVALID;
    if (!ref || n_rows < 2 || n_rows > h->cfg.max_ref_rows)
        return fail(h, MPPI_ERR_INVALID, "%s", "ref path needs 2..max_ref_rows rows of (x, y, dq1, dq2)");
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaMemcpy(h->dev + h->ws.off_ref, ref, (size_t)n_rows * 4 * sizeof(double), cudaMemcpyHostToDevice));
    h->n_ref_rows = n_rows;
    h->dc.n_ref_rows = n_rows;
    if (h->cfg.max_ref_rows <= kWinTableMaxRows) {
        // everything of a step block that depends only on (path, window start): built here, once, for every start
        mppi_window_table_sm100a<<<n_rows, 32>>>(h->dc, (const double*)(h->dev + h->ws.off_ref), n_rows,
                                                  h->dev + h->ws.off_win_table);
        CU(h, cudaGetLastError());
        CU(h, cudaDeviceSynchronize());
        h->launches += 1;
    }
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
    if (h->tick_exec) { cudaGraphExecDestroy(h->tick_exec); h->tick_exec = nullptr; }
    if (h->sharded_exec) { cudaGraphExecDestroy(h->sharded_exec); h->sharded_exec = nullptr; }
    return MPPI_OK;
}

int mppi_step_local(MppiHandle* h, int32_t noise_mode, const float* eps_dev, double* partial_dev, void* stream) {
    if (!h) return MPPI_ERR_INVALID;
    if (!partial_dev) return fail(h, MPPI_ERR_INVALID, "%s", "null partial_dev");
    h->timing_pending = false;
    const uint64_t before = h->launches;
    StepOpts o; o.capturing = h->capture_mode;
    int rc = enqueue_local(h, noise_mode, eps_dev, partial_dev, (cudaStream_t)stream, o);
    if (h->capture_mode) { h->capture_kernels += h->launches - before; h->launches = before; }
    return rc;
}

int mppi_step_combine(MppiHandle* h, const double* gathered_dev, int32_t world, void* stream) {
    if (!h) return MPPI_ERR_INVALID;
    if (!gathered_dev) return fail(h, MPPI_ERR_INVALID, "%s", "null gathered_dev");
    const uint64_t before = h->launches;
    StepOpts o; o.capturing = h->capture_mode;
    int rc = enqueue_combine(h, gathered_dev, world, (cudaStream_t)stream, o);
    if (h->capture_mode) { h->capture_kernels += h->launches - before; h->launches = before; }
    return rc;
}

int mppi_set_capture_mode(MppiHandle* h, int32_t on) {
    if (!h) return MPPI_ERR_INVALID;
    if (on && !h->capture_mode) h->capture_kernels = 0;
    h->capture_mode = on != 0;
    return MPPI_OK;
}

int mppi_replay_begin(MppiHandle* h, void* stream) {
    if (!h) return MPPI_ERR_INVALID;
    if (h->const_window) return const_acquire(h, (cudaStream_t)stream);
    return MPPI_OK;
}

int mppi_replay_end(MppiHandle* h, void* stream) {
    if (!h) return MPPI_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    if (h->const_window) { int rc = const_release(h, s); if (rc != MPPI_OK) return rc; }
    CU(h, cudaEventRecord(h->done, s));
    h->launches += h->capture_kernels;
    h->have_step = true;
    h->timing_pending = false;
    return MPPI_OK;
}

int mppi_step(MppiHandle* h, int32_t noise_mode, const float* eps_dev, void* stream) {
    if (!h) return MPPI_ERR_INVALID;
    if (h->cfg.K_local != h->cfg.K_total)
        return fail(h, MPPI_ERR_INVALID, "%s", "mppi_step needs the whole sample set on this handle; use mppi_step_local/combine");
    cudaStream_t s = (cudaStream_t)stream;
    double* partial = (double*)(h->dev + h->ws.off_partial);
    const bool use_graph = (h->cfg.flags & MPPI_FLAG_DEVICE_GRAPH) && noise_mode == MPPI_NOISE_PHILOX &&
                           !h->timing && s != nullptr;
    if (use_graph) {
        if (!h->graph_exec || h->graph_stream != stream) {
            if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
            const uint64_t before = h->launches;
            cudaGraph_t g = nullptr;
            CU(h, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            StepOpts o; o.capturing = true; o.fuse_finalize = true;
            int rc = enqueue_local(h, noise_mode, nullptr, partial, s, o);
            if (rc == MPPI_OK) rc = enqueue_combine(h, partial, 1, s, o);
            cudaError_t ce = cudaStreamEndCapture(s, &g);
            if (rc != MPPI_OK) { if (g) cudaGraphDestroy(g); return rc; }
            CU(h, ce);
            CU(h, cudaGraphInstantiate(&h->graph_exec, g, 0));
            cudaGraphDestroy(g);
            h->graph_kernels = h->launches - before;
            h->launches = before;
            h->graph_stream = stream;
        }
        if (h->const_window) { int rc = const_acquire(h, s); if (rc != MPPI_OK) return rc; }
        CU(h, cudaGraphLaunch(h->graph_exec, s));
        if (h->const_window) { int rc = const_release(h, s); if (rc != MPPI_OK) return rc; }
        CU(h, cudaEventRecord(h->done, s));
        h->launches += h->graph_kernels;
        h->timing_pending = false;
        return MPPI_OK;
    }
    StepOpts o; o.timed = h->timing;
    o.fuse_finalize = noise_mode == MPPI_NOISE_PHILOX && !h->timing;      // (timed: finalize as its own kernel, own timer)
    int rc = enqueue_local(h, noise_mode, eps_dev, partial, s, o);
    if (rc != MPPI_OK) return rc;
    rc = enqueue_combine(h, partial, 1, s, o);
    h->timing_pending = h->timing && rc == MPPI_OK;
    return rc;
}

size_t mppi_exchange_bytes(const MppiConfig* c, int32_t world) {
    const char* why = nullptr;
    if (!valid_cfg(c, &why) || world < 1 || world > kMaxPeers) { fail(nullptr, MPPI_ERR_INVALID, "%s", why ? why : "world must be 1..16"); return 0; }
    const size_t slot = (size_t)world * c->n_env * (2 + 2 * c->T) * sizeof(double);
    return align_up(2 * slot, 256) + align_up(2 * (size_t)world * c->n_env * sizeof(unsigned long long), 256);
}

int mppi_set_peer_exchange(MppiHandle* h, int32_t rank, int32_t world, void* const* peer_bufs) {
    if (!h) return MPPI_ERR_INVALID;
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || !peer_bufs)
        return fail(h, MPPI_ERR_INVALID, "%s", "peer exchange needs 1 <= world <= 16, 0 <= rank < world and the peer buffer table");
    for (int r = 0; r < world; ++r)
        if (!peer_bufs[r] || ((uintptr_t)peer_bufs[r] & 15)) return fail(h, MPPI_ERR_INVALID, "%s", "null or misaligned peer buffer");
    h->px.rank = rank; h->px.world = world;   // (timeout_ns keeps its value)
    for (int r = 0; r < kMaxPeers; ++r) h->px.buf[r] = r < world ? (char*)peer_bufs[r] : nullptr;
    h->px.slot_bytes = (size_t)world * h->cfg.n_env * (2 + 2 * h->cfg.T) * sizeof(double);
    h->px.flags_off = align_up(2 * h->px.slot_bytes, 256);
    if (h->sharded_exec) { cudaGraphExecDestroy(h->sharded_exec); h->sharded_exec = nullptr; }
    return MPPI_OK;
}

int mppi_step_sharded(MppiHandle* h, int32_t noise_mode, const float* eps_dev, void* stream) {
    if (!h) return MPPI_ERR_INVALID;
    if (h->px.world < 1) return fail(h, MPPI_ERR_INVALID, "%s", "mppi_set_peer_exchange() has not been called");
    cudaStream_t s = (cudaStream_t)stream;
    double* partial = (double*)(h->dev + h->ws.off_partial);
    const double* local_slots = (const double*)h->px.buf[h->px.rank];
    const bool use_graph = (h->cfg.flags & MPPI_FLAG_DEVICE_GRAPH) && noise_mode == MPPI_NOISE_PHILOX && s != nullptr;
    h->timing_pending = false;
    if (use_graph) {
        if (!h->sharded_exec || h->sharded_stream != stream) {
            if (h->sharded_exec) { cudaGraphExecDestroy(h->sharded_exec); h->sharded_exec = nullptr; }
            const uint64_t before = h->launches;
            cudaGraph_t g = nullptr;
            CU(h, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            StepOpts o; o.capturing = true; o.use_px = true; o.fuse_finalize = true;
            int rc = enqueue_local(h, noise_mode, nullptr, partial, s, o);
            if (rc == MPPI_OK) rc = enqueue_combine(h, local_slots, h->px.world, s, o);
            cudaError_t ce = cudaStreamEndCapture(s, &g);
            if (rc != MPPI_OK) { if (g) cudaGraphDestroy(g); return rc; }
            CU(h, ce);
            CU(h, cudaGraphInstantiate(&h->sharded_exec, g, 0));
            cudaGraphDestroy(g);
            h->sharded_kernels = h->launches - before;
            h->launches = before;
            h->sharded_stream = stream;
        }
        if (h->const_window) { int rc = const_acquire(h, s); if (rc != MPPI_OK) return rc; }
        CU(h, cudaGraphLaunch(h->sharded_exec, s));
        if (h->const_window) { int rc = const_release(h, s); if (rc != MPPI_OK) return rc; }
        CU(h, cudaEventRecord(h->done, s));
        h->launches += h->sharded_kernels;
        h->have_step = true;
        return MPPI_OK;
    }
    StepOpts o; o.use_px = true; o.fuse_finalize = noise_mode == MPPI_NOISE_PHILOX;
    int rc = enqueue_local(h, noise_mode, eps_dev, partial, s, o);
    if (rc != MPPI_OK) return rc;
    return enqueue_combine(h, local_slots, h->px.world, s, o);
}

int mppi_exchange_status(MppiHandle* h) {
    if (!h) return MPPI_ERR_INVALID;
    // the status words of the last step were delivered to io_host with its other results (valid after mppi_wait)
    const int32_t* st = (const int32_t*)(h->host + h->io.off_status);
    int any = 0;
    for (int e = 0; e < h->cfg.n_env; ++e) any |= st[e];
    return any & 1;
}

int mppi_set_exchange_timeout(MppiHandle* h, double milliseconds) {
    if (!h) return MPPI_ERR_INVALID;
    if (!(milliseconds > 0.0)) return fail(h, MPPI_ERR_INVALID, "%s", "the exchange timeout must be > 0 ms");
    h->px.timeout_ns = (unsigned long long)(milliseconds * 1.0e6);
    if (h->sharded_exec) { cudaGraphExecDestroy(h->sharded_exec); h->sharded_exec = nullptr; }   // (a kernel argument)
    return MPPI_OK;
}

int mppi_upload_state(MppiHandle* h, void* stream) {
    if (!h) return MPPI_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    CU(h, cudaMemcpyAsync(h->dev + h->ws.off_in, h->host, h->in_bytes, cudaMemcpyHostToDevice, s));
    if (h->cfg.flags & MPPI_FLAG_RESIDENT_STATE) {
        // the prepare kernel of a resident handle advances the step counter BEFORE using it
        h->step_scratch = *(const uint64_t*)(h->host + h->io.off_step) - 1ull;
        CU(h, cudaMemcpyAsync(h->dev + h->ws.off_in + h->io.off_step, &h->step_scratch, sizeof(uint64_t),
                              cudaMemcpyHostToDevice, s));
    }
    CU(h, cudaEventRecord(h->done, s));
    return MPPI_OK;
}

int mppi_download_state(MppiHandle* h, void* stream) {
    if (!h) return MPPI_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    // u_prev, prev_idx and the step counter as the device holds them (x0 is the caller's own)
    CU(h, cudaMemcpyAsync(h->host + h->io.off_u_prev, h->dev + h->ws.off_in + h->io.off_u_prev,
                          h->in_bytes - h->io.off_u_prev, cudaMemcpyDeviceToHost, s));
    CU(h, cudaMemcpyAsync(h->host + h->out_off, h->dev + h->ws.off_out, h->out_bytes, cudaMemcpyDeviceToHost, s));
    CU(h, cudaEventRecord(h->done, s));
    return MPPI_OK;
}

int mppi_closed_loop(MppiHandle* h, int32_t n_steps, double plant_dt, double* log_dev, int32_t* stop_dev,
                     void* stream) {
    if (!h) return MPPI_ERR_INVALID;
    if (h->cfg.K_local != h->cfg.K_total)
        return fail(h, MPPI_ERR_INVALID, "%s", "the device closed loop needs the whole sample set on this handle");
    if (n_steps < 1 || !log_dev || !stop_dev || !(plant_dt > 0.0))
        return fail(h, MPPI_ERR_INVALID, "%s", "closed loop needs n_steps >= 1, plant_dt > 0, log_dev and stop_dev");
    if (h->n_ref_rows < 2) return fail(h, MPPI_ERR_INVALID, "%s", "mppi_set_ref_path() has not been called");
    cudaStream_t s = (cudaStream_t)stream;
    if (!s) return fail(h, MPPI_ERR_INVALID, "%s", "the device closed loop needs a non-default stream");
    char* ws = h->dev;
    LoopParams* lp_dev = (LoopParams*)(ws + h->ws.off_loop);
    double* partial = (double*)(ws + h->ws.off_partial);
    LoopParams lp;
    lp.plant_dt = plant_dt; lp.log = log_dev; lp.stop = stop_dev; lp.tick = 0; lp.n_steps = n_steps;
    // pageable source: the runtime stages it before returning, so the stack copy is safe
    CU(h, cudaMemcpyAsync(lp_dev, &lp, sizeof(lp), cudaMemcpyHostToDevice, s));
    CU(h, cudaMemsetAsync(stop_dev, 0x7f, sizeof(int32_t) * h->cfg.n_env, s));
    CU(h, cudaMemcpyAsync(ws + h->ws.off_in, h->host, h->in_bytes, cudaMemcpyHostToDevice, s));
    if (h->cfg.flags & MPPI_FLAG_RESIDENT_STATE) {      // (the prepare kernel advances the counter before using it)
        h->step_scratch = *(const uint64_t*)(h->host + h->io.off_step) - 1ull;
        CU(h, cudaMemcpyAsync(ws + h->ws.off_in + h->io.off_step, &h->step_scratch, sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    }
    if (!h->tick_exec || h->tick_stream != stream) {
        if (h->tick_exec) { cudaGraphExecDestroy(h->tick_exec); h->tick_exec = nullptr; }
        const uint64_t before = h->launches;
        cudaGraph_t g = nullptr;
        CU(h, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        StepOpts o; o.capturing = true; o.host_io = false; o.fuse_finalize = true;
        int rc = enqueue_local(h, MPPI_NOISE_PHILOX, nullptr, partial, s, o);
        if (rc == MPPI_OK) rc = enqueue_combine(h, partial, 1, s, o);
        if (rc == MPPI_OK) {
            mppi_plant_sm100a<<<h->dc.n_env, 32, 0, s>>>(h->dc, h->dio, ws + h->ws.off_step_blocks, lp_dev);
            h->launches += 1;
        }
        cudaError_t ce = cudaStreamEndCapture(s, &g);
        if (rc != MPPI_OK) { if (g) cudaGraphDestroy(g); return rc; }
        CU(h, ce);
        CU(h, cudaGraphInstantiate(&h->tick_exec, g, 0));
        cudaGraphDestroy(g);
        h->tick_kernels = h->launches - before;
        h->launches = before;
        h->tick_stream = stream;
    }
    if (h->const_window) { int rc = const_acquire(h, s); if (rc != MPPI_OK) return rc; }
    for (int i = 0; i < n_steps; ++i) CU(h, cudaGraphLaunch(h->tick_exec, s));
    if (h->const_window) { int rc = const_release(h, s); if (rc != MPPI_OK) return rc; }
    h->launches += h->tick_kernels * (uint64_t)n_steps;
    // final controller state back into the caller's input fields, last tick's results into the outputs
    CU(h, cudaMemcpyAsync(h->host, ws + h->ws.off_in, h->in_bytes, cudaMemcpyDeviceToHost, s));
    CU(h, cudaMemcpyAsync(h->host + h->out_off, h->dev + h->ws.off_out, h->out_bytes, cudaMemcpyDeviceToHost, s));
    CU(h, cudaEventRecord(h->done, s));
    h->have_step = true;
    h->timing_pending = false;
    return MPPI_OK;
}

int mppi_wait(MppiHandle* h) {
    if (!h) return MPPI_ERR_INVALID;
    CU(h, cudaEventSynchronize(h->done));
    if (h->timing_pending) {
        for (int i = 0; i < kNumTimers; ++i) {
            float ms = 0.f;
            CU(h, cudaEventElapsedTime(&ms, h->tev[i], h->tev[i + 1]));
            h->t_acc[i] += (double)ms * 1e3;
        }
        h->t_steps += 1;
        h->timing_pending = false;
    }
    return MPPI_OK;
}

int mppi_last_costs(MppiHandle* h, const float** S_dev, const float** w_dev) {
    if (!h) return MPPI_ERR_INVALID;
    if (!h->have_step) return fail(h, MPPI_ERR_INVALID, "%s", "no step has run yet");
    if (S_dev) *S_dev = (const float*)(h->dev + h->ws.off_S);
    if (w_dev) *w_dev = (const float*)(h->dev + h->ws.off_w);
    return MPPI_OK;
}

int mppi_step_block(MppiHandle* h, int32_t env, const void** dev_ptr, size_t* bytes) {
    if (!h || !dev_ptr || !bytes) return MPPI_ERR_INVALID;
    if (!h->have_step) return fail(h, MPPI_ERR_INVALID, "%s", "no step has run yet");
    if (env < 0 || env >= h->cfg.n_env) return fail(h, MPPI_ERR_INVALID, "%s", "environment index out of range");
    *dev_ptr = h->dev + h->ws.off_step_blocks + (size_t)env * h->dc.step_block_bytes;
    *bytes = (size_t)h->dc.step_block_bytes;
    return MPPI_OK;
}

int mppi_sampled_trajectories_subset(MppiHandle* h, int32_t noise_mode, const float* eps_dev, const int32_t* subset_dev,
                                     int32_t n_subset, float* traj_dev, void* stream) {
    if (!h) return MPPI_ERR_INVALID;
    if (!h->have_step) return fail(h, MPPI_ERR_INVALID, "%s", "no step has run yet");
    if (!traj_dev) return fail(h, MPPI_ERR_INVALID, "%s", "null traj_dev");
    if (noise_mode == MPPI_NOISE_INJECTED && !eps_dev) return fail(h, MPPI_ERR_INVALID, "%s", "injected noise mode needs eps_dev");
    if (subset_dev && n_subset < 1) return fail(h, MPPI_ERR_INVALID, "%s", "empty subset");
    cudaStream_t s = (cudaStream_t)stream;
    const uint64_t* step_ctr = (const uint64_t*)(h->dev + h->ws.off_in + h->io.off_step);
    const char* step_blocks = h->dev + h->ws.off_step_blocks;
    const int n_rows = subset_dev ? n_subset : h->dc.K_local;
    dim3 grid((n_rows + 127) / 128, h->dc.n_env);
    if (noise_mode == MPPI_NOISE_PHILOX)
        mppi_sampled_traj_sm100a<0><<<grid, 128, 0, s>>>(h->dc, step_ctr, step_blocks, nullptr, subset_dev, n_rows, traj_dev);
    else
        mppi_sampled_traj_sm100a<1><<<grid, 128, 0, s>>>(h->dc, step_ctr, step_blocks, eps_dev, subset_dev, n_rows, traj_dev);
    CU(h, cudaGetLastError());
    h->launches += 1;
    return MPPI_OK;
}

int mppi_sampled_trajectories(MppiHandle* h, int32_t noise_mode, const float* eps_dev, float* traj_dev, void* stream) {
    return mppi_sampled_trajectories_subset(h, noise_mode, eps_dev, nullptr, 0, traj_dev, stream);
}

int mppi_philox_noise(MppiHandle* h, uint64_t step, float* eps_dev, void* stream) {
    if (!h) return MPPI_ERR_INVALID;
    if (!eps_dev) return fail(h, MPPI_ERR_INVALID, "%s", "null eps_dev");
    const long long n = (long long)h->dc.K_local * ((h->dc.T + 1) / 2);
    dim3 grid((unsigned)((n + 255) / 256), h->dc.n_env);
    mppi_philox_export_sm100a<<<grid, 256, 0, (cudaStream_t)stream>>>(h->dc, (uint32_t)step, eps_dev);
    CU(h, cudaGetLastError());
    h->launches += 1;
    return MPPI_OK;
}

uint64_t mppi_launch_count(const MppiHandle* h) { return h ? h->launches : 0; }

int mppi_search_stats(MppiHandle* h, uint64_t* out3, int32_t reset, void* stream) {
    if (!h || !out3) return MPPI_ERR_INVALID;
    if (!(h->cfg.flags & MPPI_FLAG_SEARCH_STATS))
        return fail(h, MPPI_ERR_INVALID, "%s", "the handle was created without MPPI_FLAG_SEARCH_STATS");
    CU(h, cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long v[4] = { 0, 0, 0, 0 };
    CU(h, cudaMemcpyAsync(v, h->dev + h->ws.off_stats, sizeof(v), cudaMemcpyDeviceToHost, s));
    if (reset) CU(h, cudaMemsetAsync(h->dev + h->ws.off_stats, 0, sizeof(v), s));
    CU(h, cudaStreamSynchronize(s));                    // only the stream the steps were enqueued on
    out3[0] = v[0]; out3[1] = v[1]; out3[2] = v[2];
    return MPPI_OK;
}

int mppi_set_timing(MppiHandle* h, int32_t enable) {
    if (!h) return MPPI_ERR_INVALID;
    h->timing = enable != 0;
    h->timing_pending = false;
    for (int i = 0; i < kNumTimers; ++i) h->t_acc[i] = 0.0;
    h->t_steps = 0;
    return MPPI_OK;
}

int mppi_get_timing(MppiHandle* h, double* out_us, int32_t n) {
    if (!h || !out_us) return MPPI_ERR_INVALID;
    for (int i = 0; i < n && i < kNumTimers; ++i) out_us[i] = h->t_steps ? h->t_acc[i] / h->t_steps : 0.0;
    return h->t_steps;
}

int mppi_probe_fp32(int32_t device, double ms, double* fma_flops, double* mufu_ops) {
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return MPPI_ERR_NO_DEVICE; }
    int sm = sm_count_or_default(device);
    float* out = nullptr;
    const int blocks = sm * 8, threads = 256;
    if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(float)) != cudaSuccess) return MPPI_ERR_CUDA;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    int rc = MPPI_OK;
    for (int which = 0; which < 2; ++which) {
        int iters = 256;
        double best = 0.0, spent = 0.0;
        for (int rep = 0; rep < 64 && spent < ms; ++rep) {
            cudaEventRecord(a);
            if (which == 0) mppi_probe_fma_sm100a<<<blocks, threads>>>(out, iters, 0.999f, 1e-3f);
            else mppi_probe_mufu_sm100a<<<blocks, threads>>>(out, iters);
            cudaEventRecord(b);
            if (cudaEventSynchronize(b) != cudaSuccess) { rc = MPPI_ERR_CUDA; break; }
            float t = 0.f; cudaEventElapsedTime(&t, a, b);
            spent += t;
            const double ops = (double)blocks * threads * iters * 16.0 * (which == 0 ? 8.0 * 2.0 : 4.0);
            const double rate = ops / (t * 1e-3);
            if (rep > 0 && rate > best) best = rate;
            if (t < 5.0f) iters *= 2;
        }
        if (which == 0 && fma_flops) *fma_flops = best;
        if (which == 1 && mufu_ops) *mufu_ops = best;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(out);
    return rc;
}

}  // extern "C"

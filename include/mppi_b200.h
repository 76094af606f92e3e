/* mppi_b200.h — C ABI of libmppi_b200.so: the B200 (sm_100a) MPPI step.
 *
 * This is the drop-in boundary below the Python class `MPPIControllerForPathTracking`
 * (reference: /root/reference/control.py:20-152).  The reference has no FFI of its own (it is pure
 * NumPy); each entry point names the reference lines whose work it takes over.  Plain C types only:
 * no C++ exceptions cross this boundary, every function returns 0 on success or a negative
 * MPPI_ERR_* code, and mppi_last_error() describes the last failure of a handle.
 *
 * Memory and threading contract
 *   - The caller owns all memory.  `workspace` is device memory of at least mppi_workspace_bytes()
 *     bytes (256-byte aligned); `io_host` is PINNED host memory of at least mppi_io_bytes() bytes.
 *     Nothing is allocated or freed inside mppi_step*().
 *   - All device work is enqueued on the caller's stream (a cudaStream_t passed as void*); the
 *     mppi_step*() calls return without synchronising.  mppi_wait() blocks until the results of the
 *     last step are visible in `io_host`.
 *   - One handle drives one GPU and one control loop; a handle is not thread-safe, different
 *     handles are independent.
 *
 * Sample sharding (multi-GPU): a handle rolls out the global samples
 * [k_offset, k_offset + K_local) of K_total.  mppi_step_local() leaves this shard's partial result
 * (rho_g, eta_g, V_g[T*2]) per environment in device memory; the caller all-gathers the partials of
 * all ranks (NCCL, 8*(2+2T) bytes per rank and environment) and hands them to mppi_step_combine(),
 * which every rank evaluates identically.
 */
#ifndef MPPI_B200_H_
#define MPPI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPPI_ABI_VERSION 3

enum {
    MPPI_OK = 0,
    MPPI_ERR_INVALID = -1,      /* bad argument / configuration                                  */
    MPPI_ERR_CUDA = -2,         /* a CUDA runtime call failed (see mppi_last_error)              */
    MPPI_ERR_NO_DEVICE = -3,    /* no sm_100 device: there is deliberately no CPU fallback       */
    MPPI_ERR_WORKSPACE = -4     /* workspace / io block too small or misaligned                  */
};

enum {
    MPPI_NOISE_PHILOX = 0,      /* eps drawn in-kernel, Philox4x32-10 keyed on (seed, step, env, k, t) */
    MPPI_NOISE_INJECTED = 1     /* eps read from a caller tensor [K_local, T, 2] float32         */
};

enum {
    MPPI_FLAG_OPTIMAL_TRAJ = 1, /* control.py:129-134: roll the updated sequence out             */
    MPPI_FLAG_DEVICE_GRAPH = 2, /* replay the step as one CUDA graph (Philox mode only)          */
    MPPI_FLAG_SMOOTH_AVERAGE = 4, /* smooth the update with control.py:329-344 instead of the median filter (needs T >= 10) */
    MPPI_FLAG_SMOOTH_NONE = 8,  /* no smoothing of the weighted noise sum                        */
    MPPI_FLAG_FULL_SEARCH = 16, /* control.py:208-215: always run the 30-candidate search (kernels compiled without the
                                   certified lookups; results are bit-identical either way) */
    MPPI_FLAG_DYNAMICS_F1 = 32, /* roll out with control.py:265-295 (_F1, feedback-linearised) instead of _F */
    MPPI_FLAG_SEARCH_STATS = 64, /* count certified / total warp-lookups of the rollouts (mppi_search_stats) */
    MPPI_FLAG_RESIDENT_STATE = 128 /* the controller state (u_prev, prev_idx, step counter: control.py:59, 65) stays on the
                                   device between steps: each step reads only x0 from io_host, applies the shift of
                                   control.py:148-149 and the index update of control.py:230 itself, and delivers only
                                   the compact results (new_idx, status, rho, eta, u0).  mppi_upload_state() /
                                   mppi_download_state() move the rest on demand (many environments: 36 B instead of
                                   5 KB per environment and step over PCIe at T = 64) */
};

/* Hyper-parameters: control.py:21-65 + sys_params.py:3-10, fixed for the life of a handle. */
typedef struct MppiConfig {
    int32_t abi_version;        /* MPPI_ABI_VERSION                                              */
    int32_t device;             /* CUDA ordinal                                                  */
    int32_t n_env;              /* independent arm instances stepped together (1 = the reference)*/
    int32_t K_total;            /* number_of_samples_K (control.py:26)                           */
    int32_t K_local;            /* samples rolled out by this handle                             */
    int32_t k_offset;           /* global index of this handle's first sample                    */
    int32_t T;                  /* horizon_step_T (control.py:25), 1..MPPI_MAX_T                 */
    int32_t n_exploit;          /* #samples with k < (1-param_exploration)*K (control.py:98)     */
    int32_t flags;              /* MPPI_FLAG_*                                                   */
    int32_t max_ref_rows;       /* capacity reserved in the workspace for mppi_set_ref_path()    */
    double delta_t;             /* control.py:24                                                 */
    double param_lambda;        /* control.py:28                                                 */
    double param_gamma;         /* lambda*(1-alpha), control.py:45                               */
    double sigma_chol[4];       /* lower Cholesky factor of Sigma, row-major (noise draw)        */
    double sigma_inv[4];        /* inverse of Sigma, row-major (control.py:106)                  */
    double stage_cost_weight[4];    /* control.py:31                                             */
    double terminal_cost_weight[4]; /* control.py:32                                             */
    double arm[7];              /* m1, m2, l1, l2, lc1, lc2, g  (sys_params.py:4-10)             */
    double cost_l1, cost_l2;    /* self.l1, self.l2 of the cost-side kinematics (control.py:55-56)*/
    uint64_t seed;              /* Philox key                                                    */
    /* Joint-limit stage cost (BASELINE north_star item 1).  NOT in the reference, whose only limits are the
     * commented-out clamps of _g (control.py:166-172): every horizon step adds
     * joint_limit_weight * 1e4 * (viol(q1)^2 + viol(q2)^2), viol(q) = max(q - hi, lo - q, 0).
     * joint_limit_weight = 0 (default) gives the reference's step, and launches kernels without the term. */
    double joint_limit_lo[2], joint_limit_hi[2];
    double joint_limit_weight;
} MppiConfig;

#define MPPI_MAX_T 256

/* Layout of the pinned host block shared with the library (all offsets in doubles / ints below are
 * per environment e; arrays are [n_env][...] contiguous).  Obtain the offsets with mppi_io_layout(). */
typedef struct MppiIoLayout {
    size_t bytes;               /* total size of the block                                       */
    /* inputs, written by the caller before mppi_step*()                                         */
    size_t off_x0;              /* double [n_env][4]      observed state (control.py:72)         */
    size_t off_u_prev;          /* double [n_env][T][2]   nominal sequence (control.py:70)       */
    size_t off_prev_idx;        /* int32  [n_env]         prev_waypoints_idx (control.py:204)    */
    size_t off_step;            /* uint64 [1]             control-step counter (Philox)          */
    /* outputs, valid after mppi_wait()                                                          */
    size_t off_new_idx;         /* int32  [n_env]         updated waypoint index (control.py:230)*/
    size_t off_status;          /* int32  [n_env]         bit 0: peer exchange timed out, update skipped */
    size_t off_rho;             /* double [n_env]         min cost (control.py:303)              */
    size_t off_eta;             /* double [n_env]         normaliser (control.py:306-308)        */
    size_t off_u0;              /* double [n_env][2]      the control calc_control_input returns (control.py:152:
                                                          first row of the shifted sequence)     */
    /* [off_new_idx, off_w_eps_raw) = the compact results of MPPI_FLAG_RESIDENT_STATE            */
    size_t off_w_eps_raw;       /* double [n_env][T][2]   weighted noise sum (control.py:115-118)*/
    size_t off_w_eps_filt;      /* double [n_env][T][2]   after the median filter (control.py:122)*/
    size_t off_u_new;           /* double [n_env][T][2]   u + filtered update (control.py:126)   */
    size_t off_opt_traj;        /* double [n_env][T][4]   control.py:129-134 (zeros if flag off) */
} MppiIoLayout;

typedef struct MppiHandle MppiHandle;

/* --- life cycle -------------------------------------------------------------------------------- */
int mppi_abi_version(void);
/* Number of CUDA devices that can run the library (compute capability 10.x); 0 if none. */
int mppi_device_count(void);
size_t mppi_workspace_bytes(const MppiConfig* cfg);
int mppi_io_layout(const MppiConfig* cfg, MppiIoLayout* out);
/* replaces MPPIControllerForPathTracking.__init__ (control.py:21-65) */
int mppi_create(const MppiConfig* cfg, void* workspace, size_t workspace_bytes,
                void* io_host, size_t io_bytes, MppiHandle** out);
void mppi_destroy(MppiHandle* h);
const char* mppi_last_error(const MppiHandle* h);   /* h may be NULL: last create() failure */

/* Upload the reference path [n_rows][4] = (x, y, dq1_ref, dq2_ref), float64 host memory
 * (self.ref_path, control.py:54; run.py:18-19).  Synchronous. */
int mppi_set_ref_path(MppiHandle* h, const double* ref_xydq, int32_t n_rows);

/* --- the step ---------------------------------------------------------------------------------- */
/* One full MPPI step on one GPU (K_local == K_total): replaces control.py:75-134 —
 * waypoint update, noise, K rollouts with costs, soft-min weights, weighted noise sum, median
 * filter, sequence update and the optimal-trajectory rollout.  `eps_dev` is NULL in Philox mode,
 * else device float32 [n_env][K_local][T][2].  Inputs are read from / results written to io_host. */
int mppi_step(MppiHandle* h, int32_t noise_mode, const float* eps_dev, void* stream);

/* Sharded step, first half: everything up to this shard's partial (rho_g, eta_g, V_g).
 * `partial_dev` receives double [n_env][2 + 2T]. */
int mppi_step_local(MppiHandle* h, int32_t noise_mode, const float* eps_dev,
                    double* partial_dev, void* stream);
/* Sharded step, second half: combine `world` gathered partials, double [world][n_env][2 + 2T]
 * (device), then filter / update / optimal trajectory and the copy to io_host. */
int mppi_step_combine(MppiHandle* h, const double* gathered_dev, int32_t world, void* stream);

int mppi_wait(MppiHandle* h);

/* MPPI_FLAG_RESIDENT_STATE: push the controller state in io_host (u_prev, prev_idx, step counter; x0 too) to the
 * device / fetch the device's controller state and the full outputs of the last step into io_host.  Both are
 * enqueued on `stream`; mppi_wait() (or a stream synchronise) completes them.  A resident handle must be
 * uploaded once before its first step and after every host-side change of the state. */
int mppi_upload_state(MppiHandle* h, void* stream);
int mppi_download_state(MppiHandle* h, void* stream);

/* Device-resident closed loop — n_steps ticks of run.py:48-59 without a host round trip (Philox noise,
 * whole sample set on this handle).  Per tick: the MPPI step above; u = first row of the shifted
 * sequence (what calc_control_input returns, control.py:152); plant dq += dt*Arm_Dynamic(q,dq,u),
 * q += dt*dq in FP64 (utils.py:14-29, run.py:53-55); sequence shift (control.py:148-149).
 * Starts from the inputs in io_host; afterwards the input fields of io_host hold the final controller
 * state (x0, u_prev, prev_idx, step) and the output fields the last tick's results.
 * log_dev:  device double [n_steps][n_env][8] = (q1, q2, dq1, dq2, u1, u2, waypoint index, rho) after each tick.
 * stop_dev: device int32 [n_env] = first tick at which the end of the path was reached (control.py:76-78;
 *           the environment is frozen from there), or >= n_steps if it never was. */
int mppi_closed_loop(MppiHandle* h, int32_t n_steps, double plant_dt, double* log_dev, int32_t* stop_dev,
                     void* stream);

/* Sharded step with the exchange fused into the kernels (no NCCL call): every rank owns an exchange
 * buffer of mppi_exchange_bytes() bytes, ZERO-INITIALISED, that all ranks of the node have mapped over
 * NVLink (peer / symmetric memory).  peer_bufs[r] = address of rank r's buffer as mapped on THIS GPU.
 * The last block of the weight-sum kernel stores this shard's (rho_g, eta_g, V_g) into every peer's
 * buffer and raises a flag, then waits for its `world` flags (bounded: mppi_set_exchange_timeout) and
 * combines.  mppi_step_sharded() = the whole step. */
size_t mppi_exchange_bytes(const MppiConfig* cfg, int32_t world);
int mppi_set_peer_exchange(MppiHandle* h, int32_t rank, int32_t world, void* const* peer_bufs);
int mppi_step_sharded(MppiHandle* h, int32_t noise_mode, const float* eps_dev, void* stream);
/* != 0 when the last finished step (after mppi_wait) skipped its update because a peer's partial did not arrive
 * within the timeout (default 3000 ms): u_new = u_prev, rho = eta = NaN.  The next step starts clean. */
int mppi_exchange_status(MppiHandle* h);
int mppi_set_exchange_timeout(MppiHandle* h, double milliseconds);

/* Caller-side CUDA-graph capture of the sharded step (mppi_step_local + the caller's collective +
 * mppi_step_combine on one capturing stream): while capture mode is on, the library enqueues only
 * capturable work (no event records).  Each replay of the caller's graph is bracketed by
 * mppi_replay_begin() / mppi_replay_end() on the replay stream; mppi_wait() then works as usual. */
int mppi_set_capture_mode(MppiHandle* h, int32_t on);
int mppi_replay_begin(MppiHandle* h, void* stream);
int mppi_replay_end(MppiHandle* h, void* stream);

/* --- auxiliary outputs ------------------------------------------------------------------------- */
/* Per-sample costs S (control.py:91-109) and un-normalised weights exp(-(S-rho_g)/lambda)
 * (control.py:297-314) of the last step: device float32 [n_env][K_local] each. */
int mppi_last_costs(MppiHandle* h, const float** S_dev, const float** w_dev);
/* The tables the prepare kernel built for environment `env` in the last step (tests / debugging), device
 * memory: 64 B header (state, window origin and start) | 32 x 16 B window coefficients | 32 x 16 B reference
 * rows | 16 x 16 B coefficient pairs | 64 B lookup certificate | 32 x 32 B row records | T x 16 B step controls. */
int mppi_step_block(MppiHandle* h, int32_t env, const void** dev_ptr, size_t* bytes);
/* control.py:137-145 — trajectories of all samples under v[k, t-1] (index wrap included), for the
 * state / sequence / window of the last step.  traj_dev: float32 [n_env][K_local][T][4]. */
int mppi_sampled_trajectories(MppiHandle* h, int32_t noise_mode, const float* eps_dev,
                              float* traj_dev, void* stream);
/* The same for a chosen subset of samples (e.g. the n lowest-cost ones, which is the order in which
 * control.py:138-145 walks them): subset_dev int32 [n_env][n_subset] local sample indices,
 * traj_dev float32 [n_env][n_subset][T][4]. */
int mppi_sampled_trajectories_subset(MppiHandle* h, int32_t noise_mode, const float* eps_dev,
                                     const int32_t* subset_dev, int32_t n_subset, float* traj_dev, void* stream);
/* The Philox noise tensor the kernels draw for control step `step`: float32 [n_env][K_local][T][2]. */
int mppi_philox_noise(MppiHandle* h, uint64_t step, float* eps_dev, void* stream);
/* Number of kernels launched by this handle so far (graph replays count their kernel nodes). */
uint64_t mppi_launch_count(const MppiHandle* h);
/* Nearest-waypoint lookups of the rollouts (control.py:200-215 via _c / _phi) since the last reset, counted
 * per warp (32 samples): out3[0] = lookups answered by a certified end-of-window test, out3[1] = all lookups,
 * out3[2] = lookups answered by a certified three-row comparison; the rest ran the 30-candidate search.
 * Needs MPPI_FLAG_SEARCH_STATS; synchronises `stream` (the stream the steps were enqueued on). */
int mppi_search_stats(MppiHandle* h, uint64_t* out3, int32_t reset, void* stream);
/* Mean device time of each kernel family over the steps run with timing enabled, in microseconds:
 * out[0..5] = prepare, rollout, softmin, weighted-sum, reduce, finalize.  Returns #timed steps. */
int mppi_set_timing(MppiHandle* h, int32_t enable);
int mppi_get_timing(MppiHandle* h, double* out_us, int32_t n);

/* --- roofline probes (bench.py) ---------------------------------------------------------------- */
/* Sustained FP32 FMA rate (FLOP/s, FMA = 2) and MUFU rate (ops/s) of `device`, measured with
 * register-resident dependent-chain kernels for about `ms` milliseconds. */
int mppi_probe_fp32(int32_t device, double ms, double* fma_flops, double* mufu_ops);

#ifdef __cplusplus
}
#endif
#endif /* MPPI_B200_H_ */

#!/usr/bin/env python
"""bench.py — MPPI sample-steps/s (and p50 control-step latency) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[3], the one the metric is quoted on; it fits one GPU): the 2-link
arm tracking xydq_circle-shaped synthetic reference, K = 1,048,576 rollouts, T = 100, in-kernel
Philox noise, run.py's hyper-parameters.  With N > 1 the K samples are sharded over the ranks
(strong scaling, total work fixed) and combined by one NCCL all-gather of 8*(2+2T) bytes per step.
A "step" is one full MPPI step: waypoint update, K rollouts, soft-min, weighted noise sum, median
filter, sequence update, optimal-trajectory rollout.

`value`   device-timed (CUDA events on the engine's stream around every step, steps enqueued back to
          back, one host sync at the end); nothing large to pre-load in Philox mode.
`e2e`     the same steps through MPPIControllerForPathTracking.calc_control_input with host buffers
          in and out (pinned H2D of state+sequence, D2H of the result, host sync) every step.
`roofline` the rollout kernel against the FP32 FMA peak measured live on this GPU.
`cpu_baseline` / `--impl reference`: the FP64 CPU restatement of the reference (oracle/mppi_oracle.c,
          OpenMP over all host threads) on a bounded sample; the reference's own pure-Python loop rate
          is reported beside it.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_TOTAL = 1 << 20
T_HORIZON = 100
FLOPS_PER_SAMPLE_STEP = 254.0          # 70 + 6*30 + 4
NCU_ROLLOUT_DRAM_BYTES = 21504         # dram__bytes_read+write of the rollout kernel, one ncu --set full capture
NCU_WSUM_DRAM_BYTES = 843073280 + 3861504   # same for mppi_wsum_injected_sm100a (algorithmic: K*T*8 = 838,860,800)
LAT_K, LAT_T = 16384, 50               # BASELINE.json configs[2]


def synthetic_ref_path(n=2000):
    """xydq_circle-shaped reference: circle r=0.6 about (0.8, 0.8) traversed once, with the joint
    rates of the arm following it (finite differences of the inverse kinematics at Ts = 0.0025)."""
    th = np.linspace(0.0, 2 * np.pi, n)
    x, y = 0.8 + 0.6 * np.cos(th), 0.8 + 0.6 * np.sin(th)
    r2 = x * x + y * y
    q2 = -np.arccos(np.clip((r2 - 2.0) / 2.0, -1, 1))
    q1 = np.arctan2(y, x) - np.arctan2(np.sin(q2), 1.0 + np.cos(q2))
    dq = np.gradient(np.stack([np.unwrap(q1), q2], axis=1), 0.0025, axis=0)
    return np.stack([x, y, dq[:, 0], dq[:, 1]], axis=1)


def run_py_kwargs(ref, K, T):
    return dict(delta_t=0.006, ref_path=ref, horizon_step_T=T, number_of_samples_K=K, param_exploration=0.0,
                param_lambda=100.0, param_alpha=0.98, sigma=np.array([[20.0, 0.0], [0.0, 20.0]]),
                stage_cost_weight=np.array([0.5, 0.5, 5.0, 5.0]),
                terminal_cost_weight=np.array([5.0, 5.0, 50.0, 50.0]))


X0 = np.array([1.152198236517471885, -1.266101672070702344, 0.0, 0.0])


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML) — runs during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """NVML clocks / throttle reasons of one GPU during the timed region.  NVML queries go through the
    driver's global lock, so they are kept sparse (first sample 5 ms into the region, then every 200 ms,
    like `nvidia-smi -lms 200`) and only rank 0 samples: eight ranks polling at 20 Hz measurably slowed
    the 8-GPU run (steps of 0.3 ms)."""

    def __init__(self, index, enabled=True, threaded=True):
        self.threaded = threaded
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self.nv = None
        if not enabled:
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown"}
        if self._stop.wait(0.005):
            return
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.2)

    def sample_now(self):
        """One synchronous sample from the calling thread (used while queued GPU work is executing)."""
        if not self.nv:
            return
        nv = self.nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown"}
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            self.reasons.update(name for bit, name in names.items() if r & bit)
        except Exception:
            pass

    def __enter__(self):
        if self.nv and self.threaded:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# CPU baseline (oracle port) — the only place bench.py executes anything under oracle/
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(T, seconds=10.0, K_sample=65536):
    from oracle import c_oracle, mppi_oracle as mo
    ref = synthetic_ref_path()
    c = mo.OracleMPPI(**run_py_kwargs(ref, K_sample, T))
    rng = np.random.default_rng(0)
    eps = rng.standard_normal((K_sample, T, 2)) * np.sqrt(20.0)
    c_oracle.rollout_costs(c, X0, eps[:1024], 0)          # build + warm
    t0, n = time.perf_counter(), 0
    per_step = []
    while True:
        t1 = time.perf_counter()
        S = c_oracle.rollout_costs(c, X0, eps, 0)
        c_oracle.weighted_sum(S, eps, c.param_lambda)
        per_step.append(time.perf_counter() - t1)
        n += 1
        if time.perf_counter() - t0 > seconds or n >= 64:
            break
    rate = K_sample * T / float(np.median(per_step))
    # the reference's own loop nest, restated in pure Python (tiny sample: it runs at ~1e4/s)
    c2 = mo.OracleMPPI(**run_py_kwargs(ref, 32, 20))
    e2 = rng.standard_normal((32, 20, 2)) * np.sqrt(20.0)
    t1 = time.perf_counter()
    mo.step_loops(c2, X0, e2)
    py_rate = 32 * 20 / (time.perf_counter() - t1)
    return dict(value=rate, unit="sample-steps/s", cores=c_oracle.num_threads(), kind="port",
                sample=f"K={K_sample} T={T} rollouts+weights, {n} passes, median; FP64 C restatement (OpenMP)",
                python_loop_rate=py_rate), per_step


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K_sample = 131072
    from oracle import c_oracle, mppi_oracle as mo
    ref = synthetic_ref_path()
    c = mo.OracleMPPI(**run_py_kwargs(ref, K_sample, T_HORIZON))
    rng = np.random.default_rng(0)
    eps = rng.standard_normal((K_sample, T_HORIZON, 2)) * np.sqrt(20.0)
    times = []
    for i in range(args.warmup + args.steps):
        t1 = time.perf_counter()
        S = c_oracle.rollout_costs(c, X0, eps, 0)
        c_oracle.weighted_sum(S, eps, c.param_lambda)
        if i >= args.warmup:
            times.append(time.perf_counter() - t1)
    ms = 1e3 * float(np.mean(times))
    value = K_sample * T_HORIZON / (ms * 1e-3)
    sample = (f"each step = K={K_sample} of the {K_TOTAL} rollouts (T={T_HORIZON}) + weights; FP64 C restatement of "
              f"control.py:91-118 with OpenMP over all host threads (the reference itself is one Python thread)")
    line = {"impl": "reference", "metric": "mppi_sample_steps_per_s", "value": value, "unit": "sample-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "sample-steps/s", "cores": c_oracle.num_threads(),
                             "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "sample-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def injected_mode_numbers(ref, torch, K=K_TOTAL, T=T_HORIZON, steps=8):
    """Correctness-run mode: eps is a [K,T,2] float32 tensor in HBM (0.84 GB at K=2^20, T=100; larger than
    L2).  The rollout reads 8 B per sample-step; the weighted sum reads 8 B per sample-step for every sample
    with a non-zero weight, so it is timed at a temperature where all weights are non-zero."""
    from mppi_robotarm_b200 import MPPIControllerForPathTracking
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    out = {"K": K, "T": T, "tensor_bytes": K * T * 8, "hbm_peak_gbs": hbm_peak,
           "hbm_peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"}
    for label, lam in (("run_py_lambda", 100.0), ("all_weights_nonzero", 1.0e9)):
        kw = run_py_kwargs(ref, K, T)
        kw["param_lambda"] = lam
        c = MPPIControllerForPathTracking(**kw, noise="philox", seed=5, verbose=False, use_graph=False)
        eng = c._engine()
        eps = eng.philox_noise(step=0)
        u = c.u_prev.copy()
        for _ in range(3):
            eng.step(X0, u, 0, eps)
        eng.set_timing(True)
        for _ in range(steps):
            eng.step(X0, u, 0, eps)
        t = eng.get_timing()
        eng.set_timing(False)
        nz = int((eng.last_costs()[1] != 0).sum().item())
        out[label] = {"rollout_us": t["rollout"], "rollout_gbs": K * T * 8 / (t["rollout"] * 1e-6) / 1e9,
                      "wsum_us": t["wsum"], "nonzero_weights": nz,
                      "wsum_gbs": nz * T * 8 / (t["wsum"] * 1e-6) / 1e9,
                      "wsum_frac_of_hbm_peak": nz * T * 8 / (t["wsum"] * 1e-6) / 1e9 / hbm_peak,
                      "sample_steps_per_s": K * T / (sum(v for k, v in t.items() if k != "steps") * 1e-6)}
        del eps
        c.close()
    return out


def batched_numbers(ref, torch, B=1024, K=1024, T=64, steps=20):
    """BASELINE.json configs[4]: B independent arm instances x K rollouts, T = 64, stepped by one launch
    of each kernel (grid.y = environment).  On N GPUs each rank takes B/N environments; this is the
    single-GPU figure for all B."""
    from mppi_robotarm_b200.batched import BatchedMPPIController
    kw = run_py_kwargs(ref, K, T)
    bat = BatchedMPPIController(B, **kw, visualize_optimal_traj=False, seed=11)
    rows = (np.arange(B) * (1900 // B + 1)) % 1900
    th = 2 * np.pi * rows / (ref.shape[0] - 1)
    x, y = 0.8 + 0.6 * np.cos(th), 0.8 + 0.6 * np.sin(th)
    q2 = -np.arccos(np.clip((x * x + y * y - 2.0) / 2.0, -1, 1))
    q1 = np.arctan2(y, x) - np.arctan2(np.sin(q2), 1.0 + np.cos(q2))
    X = np.stack([q1, q2, np.zeros(B), np.zeros(B)], axis=1)          # on-path starts
    bat.prev_waypoints_idx = rows.astype(np.int64)
    for _ in range(3):
        bat.calc_control_input(X)
        bat.prev_waypoints_idx = rows.astype(np.int64)
    bat.engine.set_timing(True)
    t0 = time.perf_counter()
    for _ in range(steps):
        bat.calc_control_input(X)
        bat.prev_waypoints_idx = rows.astype(np.int64)
    wall = (time.perf_counter() - t0) / steps
    t = bat.engine.get_timing()
    dev_us = sum(v for k, v in t.items() if k != "steps")
    out = {"workload": f"C5: {B} environments x K={K}, T={T}, Philox", "device_us_per_step": dev_us,
           "sample_steps_per_s": B * K * T / (dev_us * 1e-6), "e2e_ms_per_step": wall * 1e3,
           "e2e_sample_steps_per_s": B * K * T / wall, "kernel_us": {k: v for k, v in t.items() if k != "steps"},
           "finished_envs": int(bat.finished.sum())}
    bat.close()
    return out


def workload_config(n_gpus):
    return {"workload": f"C4: 2-link arm, K={K_TOTAL} rollouts, T={T_HORIZON}, Philox noise in-kernel, run.py "
                        f"hyper-parameters, synthetic xydq_circle-shaped reference (2000 waypoints)",
            "K": K_TOTAL, "T": T_HORIZON, "noise": "philox4x32-10", "parallelism": f"samples sharded over {n_gpus} GPU(s)",
            "l2": "256 MiB buffer rewritten between timed steps (L2 flush); the step has no reused inputs"}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-latency", action="store_true", help="skip the config-3 latency loop")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-injected", action="store_true", help="skip the injected-noise (HBM) leg")
    ap.add_argument("--no-batched", action="store_true", help="skip the config-5 batched-environments leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if os.environ.get("BENCH_HANG_DUMP"):              # debugging aid: dump all Python stacks if we stall
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["BENCH_HANG_DUMP"]), repeat=False, exit=True)
    if args.impl == "reference":
        return main_reference(args)

    import torch
    import torch.distributed as dist
    from mppi_robotarm_b200 import MPPIControllerForPathTracking

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    ref = synthetic_ref_path()
    # sharded run: partial triples are exchanged by the kernels themselves over NVLink peer memory
    # ("p2p"); BENCH_EXCHANGE=nccl selects the NCCL all-gather path instead.  If the peer mapping cannot be
    # set up on this box the NCCL path is used (both are GPU paths; the choice is recorded in `config`).
    exchange = os.environ.get("BENCH_EXCHANGE", "p2p") if distributed else "nccl"

    def make_ctrl(exch):
        c = MPPIControllerForPathTracking(**run_py_kwargs(ref, K_TOTAL, T_HORIZON), noise="philox", seed=1234,
                                          verbose=False, distributed=distributed, use_graph=True, exchange=exch)
        return c, c._engine()
    try:
        ctrl, eng = make_ctrl(exchange)
        ok = 1
    except Exception as ex:                             # noqa: BLE001
        print(f"[bench] rank {rank}: p2p exchange unavailable ({type(ex).__name__}: {ex})", file=sys.stderr)
        ok = 0
    if distributed:
        flag = torch.tensor([ok], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:                       # any rank failed: everyone falls back together
            exchange = "nccl"
            ctrl, eng = make_ctrl(exchange)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    u_nom = ctrl.u_prev.copy()

    def device_step():
        eng.write_inputs(X0, u_nom, 0)
        eng.launch(None)

    # ---- device-timed throughput ------------------------------------------------------------
    for _ in range(args.warmup):
        device_step()
        eng.wait()
    eng.set_timing(not distributed)      # per-kernel events (single-GPU path) for the roofline
    for _ in range(2):
        device_step(); eng.wait()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = eng.launch_count()
    # nvmlInit() takes milliseconds: construct the sampler BEFORE the barrier, or rank 0 starts late and
    # every other rank's first bracket contains the wait
    clock_sampler = ClockSampler(local_rank, enabled=(rank == 0), threaded=not distributed)
    barrier()
    # Device throughput: the K steps are enqueued back to back (each one bracketed by its own pair of
    # events, with the L2 flush between brackets) and the host synchronises once at the end — the
    # synthetic batches are independent, so nothing forces a host round trip per step here.  The
    # host-synchronised, closed-loop-style number is `e2e` below.
    # single GPU: background sampler thread.  Sharded run: the steps are all queued first and rank 0 then
    # samples from the main thread while the GPUs work through the queue — an NVML query takes the
    # driver's global lock for milliseconds, and one late rank makes every other rank wait in the
    # all-gather (measured: +90 us per step averaged over 50 steps of 0.23 ms).
    with clock_sampler as clocks:
        for a, b in ev:
            with torch.cuda.stream(eng.stream):
                flush.zero_()
                a.record()
            device_step()
            with torch.cuda.stream(eng.stream):
                b.record()
            if not distributed:
                eng.wait()       # single GPU: the library's per-kernel timing events are read per step
        if distributed:
            clocks.sample_now()
        eng.wait()
        barrier()
    launches = eng.launch_count() - launches0
    per_step = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(per_step)
    timing = eng.get_timing() if not distributed else None
    eng.set_timing(False)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = K_TOTAL * T_HORIZON / (ms_per_step * 1e-3)

    # ---- end-to-end through the drop-in class -------------------------------------------------
    ctrl.u_prev[...] = u_nom
    for _ in range(args.warmup):
        ctrl.prev_waypoints_idx = 0
        ctrl.calc_control_input(X0)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctrl.prev_waypoints_idx = 0
        ctrl.calc_control_input(X0)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = K_TOTAL * T_HORIZON * args.steps / float(t.item())
    lay = eng.layout
    e2e = {"value": e2e_value, "unit": "sample-steps/s", "h2d_bytes_per_step": int(lay.off_new_idx),
           "d2h_bytes_per_step": int(lay.bytes - lay.off_new_idx), "ms_per_step": 1e3 * float(t.item()) / args.steps}

    line = {"metric": "mppi_sample_steps_per_s", "value": value, "unit": "sample-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(world), exchange=(exchange if distributed else "none")),
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches),
            "ms_per_step_stats_rank0": {"min": min(per_step), "median": float(np.median(per_step)), "max": max(per_step)}}

    if rank == 0:
        # ---- roofline of the dominant kernel (rollout) against the measured FP32 peak ----------
        import ctypes as C
        fma, mufu = C.c_double(0), C.c_double(0)
        eng.lib.mppi_probe_fp32(local_rank, 200.0, C.byref(fma), C.byref(mufu))
        peak = fma.value / 1e12 if fma.value > 0 else 74.4
        if timing and timing["steps"] > 0:
            roll_us = timing["rollout"]
            ach = FLOPS_PER_SAMPLE_STEP * eng.K_local * T_HORIZON / (roll_us * 1e-6) / 1e12
            line["roofline"] = {"bound": "fp32", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                                "traffic": NCU_ROLLOUT_DRAM_BYTES,
                                "traffic_source": "ncu --set full, profiles/r1d_rollout_ncu_summary.txt "
                                                  "(dram read 21,504 B + write 0 B per launch at K=2^20, T=100)",
                                "kernel": "mppi_rollout_sm100a<philox>",
                                "kernel_us": roll_us, "peak_source": "mppi_probe_fp32 FMA chain, this GPU, this run",
                                "mufu_peak_gops": mufu.value / 1e9,
                                "flops_per_sample_step": FLOPS_PER_SAMPLE_STEP,
                                "kernel_us_all": {k: v for k, v in timing.items() if k != "steps"}}
        else:
            ach = FLOPS_PER_SAMPLE_STEP * value / world / 1e12
            line["roofline"] = {"bound": "fp32", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                                "traffic": None, "kernel": "whole step per GPU (sharded run: no per-kernel events)",
                                "peak_source": "mppi_probe_fp32 FMA chain, this GPU, this run"}
    ctrl.close()

    # ---- config 3: closed-loop latency, K=16384, T=50, one GPU ------------------------------------
    if rank == 0 and world == 1 and not args.no_latency:
        from utils import Arm_Dynamic
        c3 = MPPIControllerForPathTracking(**run_py_kwargs(ref, LAT_K, LAT_T), noise="philox", seed=7, verbose=False)
        q, dq = X0[0:2].copy(), X0[2:4].copy()
        state = np.concatenate([q, dq])
        lat = []
        n_lat, n_warm = 1000, 100
        try:
            for i in range(n_warm + n_lat):
                t1 = time.perf_counter()
                u, _, _, _ = c3.calc_control_input(state)
                dt_s = time.perf_counter() - t1
                if i >= n_warm:
                    lat.append(dt_s)
                dq = dq + 0.003 * Arm_Dynamic(q, dq, u)         # run.py:53-59 plant
                q = q + 0.003 * dq
                state = np.concatenate([q, dq])
        except IndexError:
            pass
        lat = np.array(lat) * 1e3
        line["latency"] = {"workload": f"C3: K={LAT_K}, T={LAT_T}, Philox, closed loop with the run.py plant on the host",
                           "p50_ms": float(np.percentile(lat, 50)), "p90_ms": float(np.percentile(lat, 90)),
                           "p99_ms": float(np.percentile(lat, 99)), "steps": int(lat.size),
                           "sample_steps_per_s": LAT_K * LAT_T / (float(np.percentile(lat, 50)) * 1e-3)}
        c3.close()

    # ---- injected-noise mode: achieved HBM GB/s of the two kernels that read the K x T x 2 tensor ----
    if rank == 0 and world == 1 and not args.no_injected:
        inj = injected_mode_numbers(ref, torch)
        line["injected_noise"] = inj
        a = inj["all_weights_nonzero"]
        # second roofline: the HBM-bound kernel of the noise-read mode (north_star: achieved HBM GB/s)
        line["roofline_hbm"] = {"bound": "hbm", "kernel": "mppi_wsum_injected_sm100a<float4>", "achieved": a["wsum_gbs"],
                                "peak": inj["hbm_peak_gbs"], "unit": "GB/s", "frac": a["wsum_frac_of_hbm_peak"],
                                "traffic": NCU_WSUM_DRAM_BYTES, "peak_source": inj["hbm_peak_source"],
                                "traffic_source": "ncu --set full, profiles/r1c_wsum_injected_ncu_summary.txt",
                                "algorithmic_bytes": inj["tensor_bytes"], "kernel_us": a["wsum_us"]}

    # ---- config 5: batched multi-environment sweep (environments shard over GPUs, no collective) ----
    if rank == 0 and world == 1 and not args.no_batched:
        line["batched_envs"] = batched_numbers(ref, torch)

    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"], _ = cpu_reference_rate(T_HORIZON)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if distributed:
        del eng, ctrl
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

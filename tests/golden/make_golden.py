"""Generate the golden fixtures in this directory by running the UNMODIFIED reference controller.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py [--quick]

Outputs (committed):
  ref_paths.npz        the four reference-trajectory data files as float64 arrays (inputs, App. D)
  single_steps.npz     one-step input/output vectors for a set of edge cases (K<=256)
  c2_steps.npz         config-2 vectors: K=4096, T=50, trajectory.txt in both column layouts
  closed_loop_c1.npz   a 1500-step closed loop of the reference at run.py settings with seeded noise
Noise is never stored: it is regenerated from (seed, K, T, Sigma) by oracle.mppi_oracle.injected_noise
and its checksum is stored to detect a change of NumPy's generator stream.
"""
import argparse
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh          # noqa: E402
from oracle import mppi_oracle as mo          # noqa: E402
from tests.golden import cases                # noqa: E402

PATHS = {n.replace(".txt", ""): rh.load_reference_data(n)
         for n in ("xydq_circle.txt", "xydq.txt", "trajectory.txt", "trajectory1.txt")}

KEYS = ["S", "w", "w_eps_raw", "w_eps_filt", "u_new", "u0", "optimal_traj", "u_prev_before"]


def run_case(case):
    kw = cases.ctor_kwargs(case, PATHS)
    probe = rh.ReferenceProbe(**kw)
    if case.get("smoother") == "average":
        # the step calls self._moving_median_filter(xx=..., window_size=10) (control.py:122): route that call
        # to the reference's own _moving_average_filter, keeping the probe's capture of raw / filtered
        ref_avg = probe.ctrl._moving_average_filter

        def avg(xx, window_size):
            probe.last["w_eps_raw"] = np.array(xx, copy=True)
            out = ref_avg(xx, window_size)
            probe.last["w_eps_filt"] = np.array(out, copy=True)
            return out
        probe.ctrl._moving_median_filter = avg
    if case.get("dynamics") == "F1":
        # every rollout of the step calls self._F (control.py:104, 133, 143): route it to the reference's own _F1
        probe.ctrl._F = probe.ctrl._F1
    if "prev_idx" in case:
        probe.ctrl.prev_waypoints_idx = case["prev_idx"]
    if "u_prev" in case:
        probe.ctrl.u_prev = np.array(case["u_prev"], dtype=np.float64)
    out = {}
    x = np.array(case["x0"], dtype=np.float64)
    for s in range(case.get("steps", 1)):
        eps = mo.injected_noise(case["seed"] + s, case["K"], case["T"], kw["sigma"])
        r = probe.step(list(x), eps.astype(np.float64))
        for k in KEYS:
            out[f"{k}.{s}"] = np.asarray(r[k])
        out[f"prev_idx.{s}"] = np.array([r["prev_idx_before"], r["prev_idx_after"]])
        out[f"eps_sum.{s}"] = np.array([eps.astype(np.float64).sum(), np.abs(eps.astype(np.float64)).sum()])
        out[f"x0.{s}"] = x.copy()
        if kw.get("visualze_sampled_trajs"):
            out[f"sampled_traj.{s}"] = np.asarray(r["sampled_traj"])
        # advance the plant like run.py:53-59 so multi-step cases see changing states
        q, dq = mo.plant_step(x[0:2], x[2:4], r["u0"], 0.003)
        x = np.concatenate([q, dq])
    return case["name"], out


X0 = list(cases.X0)


def closed_loop(steps, K=100, T=30, seed0=5000):
    kw = cases.run_py_kwargs(cases.ref_path_for(PATHS, "xydq_circle.txt"), K, T)
    probe = rh.ReferenceProbe(**kw)
    x = np.array(X0)
    rec = dict(state=[], u0=[], prev_idx=[], u_new=[], rho=[], gap=[], ess=[])
    t0 = time.time()
    for s in range(steps):
        eps = mo.injected_noise(seed0 + s, K, T, kw["sigma"]).astype(np.float64)
        try:
            r = probe.step(list(x), eps)
        except IndexError:
            break
        Ss = np.sort(r["S"])
        rec["state"].append(x.copy()); rec["u0"].append(r["u0"]); rec["u_new"].append(r["u_new"])
        rec["prev_idx"].append([r["prev_idx_before"], r["prev_idx_after"]])
        rec["rho"].append(Ss[0]); rec["gap"].append(Ss[1] - Ss[0]); rec["ess"].append(1.0 / np.sum(r["w"] ** 2))
        q, dq = mo.plant_step(x[0:2], x[2:4], r["u0"], 0.003)
        x = np.concatenate([q, dq])
        if s % 100 == 0:
            print(f"closed loop step {s} {time.time()-t0:.0f}s", flush=True)
    out = {k: np.array(v) for k, v in rec.items()}
    out["final_state"] = x
    out["meta"] = np.array([K, T, seed0, steps])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="skip the K=4096 and closed-loop fixtures")
    ap.add_argument("--closed-loop-steps", type=int, default=1500)
    ap.add_argument("--only", nargs="*", default=None,
                    help="(re)generate only these single-step cases and merge them into single_steps.npz")
    a = ap.parse_args()
    if a.only:
        wanted = [c for c in cases.single_cases(PATHS) if c["name"] in a.only]
        with np.load(os.path.join(HERE, "single_steps.npz")) as z:
            single = {k: z[k] for k in z.files if k.split("/", 1)[0] not in a.only}
        for c in wanted:
            name, out = run_case(c)
            single.update({f"{name}/{k}": v for k, v in out.items()})
            print("generated", name, flush=True)
        np.savez_compressed(os.path.join(HERE, "single_steps.npz"), **single)
        return

    np.savez_compressed(os.path.join(HERE, "ref_paths.npz"), **PATHS)
    SINGLE, C2 = cases.single_cases(PATHS), cases.c2_cases()

    with Pool(8) as pool:
        jobs = [pool.apply_async(run_case, (c,)) for c in SINGLE]
        c2jobs = [] if a.quick else [pool.apply_async(run_case, (c,)) for c in C2]
        cl = None if a.quick else pool.apply_async(closed_loop, (a.closed_loop_steps,))
        single = {}
        for j in jobs:
            name, out = j.get()
            single.update({f"{name}/{k}": v for k, v in out.items()})
        np.savez_compressed(os.path.join(HERE, "single_steps.npz"), **single)
        print("single_steps done", flush=True)
        if not a.quick:
            c2 = {}
            for j in c2jobs:
                name, out = j.get()
                c2.update({f"{name}/{k}": v for k, v in out.items() if not k.startswith("w.")})
            np.savez_compressed(os.path.join(HERE, "c2_steps.npz"), **c2)
            print("c2_steps done", flush=True)
            np.savez_compressed(os.path.join(HERE, "closed_loop_c1.npz"), **cl.get())
            print("closed_loop done", flush=True)


if __name__ == "__main__":
    main()

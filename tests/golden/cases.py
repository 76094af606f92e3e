"""Case table shared by make_golden.py (which runs the reference) and the parity tests (which run the
oracle and the CUDA path on the same inputs).  Pure data + tiny helpers; imports nothing from the
reference."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# run.py:14-15 — initial joint state of the reference's closed loop
X0 = [1.152198236517471885e+00, -1.266101672070702344e+00, 0.0, 0.0]


def load_paths() -> dict:
    """The four reference-trajectory files as float64 arrays (keys: file name without .txt)."""
    with np.load(os.path.join(HERE, "ref_paths.npz")) as z:
        return {k: z[k] for k in z.files}


def relayout_qxy(traj, Ts=0.0025):
    """trajectory*.txt columns are (q1,q2,x,y); the controller reads (x,y,dq1_ref,dq2_ref)
    (control.py:220-223).  Meaningful layout: finite-difference rates at Ts."""
    dq = np.gradient(traj[:, 0:2], Ts, axis=0)
    return np.concatenate([traj[:, 2:4], dq], axis=1)


def ref_path_for(paths: dict, file: str, layout: str = "verbatim") -> np.ndarray:
    a = paths[file.replace(".txt", "")]
    return np.ascontiguousarray(a[:, 0:4]) if layout == "verbatim" else relayout_qxy(a)


def single_cases(paths: dict) -> list:
    traj1 = paths["trajectory1"]
    return [
        dict(name="c1_seed0", file="xydq_circle.txt", K=100, T=30, seed=0, x0=X0, steps=3),
        dict(name="c1_seed1", file="xydq_circle.txt", K=100, T=30, seed=100, x0=X0),
        dict(name="c1_viz", file="xydq_circle.txt", K=32, T=12, seed=7, x0=X0,
             ctor=dict(visualze_sampled_trajs=True)),
        dict(name="explore", file="xydq_circle.txt", K=64, T=20, seed=11, x0=X0,
             ctor=dict(param_exploration=0.25)),
        dict(name="explore_odd", file="xydq_circle.txt", K=50, T=9, seed=12, x0=X0,
             ctor=dict(param_exploration=0.33)),
        # window truncated by the end of the path (control.py:208-209 slices past N): start joints
        # are rows of trajectory1.txt, whose end-effector lies within 1e-3 of that row of xydq_circle
        dict(name="end_of_path", file="xydq_circle.txt", K=64, T=20, seed=13, prev_idx=1985,
             x0=traj1[1987, 0:2].tolist() + [0.01, 0.01]),
        dict(name="end_of_path_2", file="xydq_circle.txt", K=40, T=16, seed=14, prev_idx=1990,
             x0=traj1[1992, 0:2].tolist() + [0.0, 0.0]),
        dict(name="full_sigma", file="xydq_circle.txt", K=80, T=25, seed=15, x0=X0,
             ctor=dict(sigma=np.array([[20.0, 6.0], [6.0, 10.0]]), param_lambda=5000.0,
                       param_alpha=0.9)),
        dict(name="soft_weights", file="xydq_circle.txt", K=256, T=30, seed=16, x0=X0, steps=2,
             ctor=dict(param_lambda=2.0e5, param_alpha=0.5)),
        dict(name="short_T", file="xydq_circle.txt", K=48, T=7, seed=17, x0=X0),
        dict(name="line_path", file="xydq.txt", K=64, T=30, seed=18, x0=[0.05, -0.1, 0.0, 0.0]),
        dict(name="no_opt_traj", file="xydq_circle.txt", K=32, T=10, seed=19, x0=X0,
             ctor=dict(visualize_optimal_traj=False)),
        # the reference's other smoother (control.py:329-344) swapped in for the median filter
        dict(name="avg_filter", file="xydq_circle.txt", K=96, T=30, seed=21, x0=X0, steps=2, smoother="average",
             ctor=dict(param_lambda=4.0e4)),
        # the reference's other rollout model (_F1, control.py:265-295) swapped in for _F
        dict(name="f1_dynamics", file="xydq_circle.txt", K=96, T=30, seed=23, x0=X0, steps=2, dynamics="F1",
             ctor=dict(param_lambda=2.0e4)),
        dict(name="f1_viz", file="xydq_circle.txt", K=32, T=12, seed=24, x0=X0, dynamics="F1",
             ctor=dict(visualze_sampled_trajs=True)),
        dict(name="mid_path", file="xydq_circle.txt", K=128, T=30, seed=20, prev_idx=700,
             x0=[0.9, 0.6, 0.3, -0.2],
             u_prev=(np.array([[3.0, 1.0]]) * np.linspace(1, 2, 30)[:, None]).tolist()),
    ]


def c2_cases() -> list:
    """BASELINE.json config 2: K=4096, T=50, trajectory.txt, both column layouts, two seeds."""
    return [dict(name=f"c2_{lay}_s{seed}", file="trajectory.txt", layout=lay, K=4096, T=50,
                 seed=seed, x0=X0)
            for lay in ("verbatim", "xydq") for seed in (0, 1)]


def run_py_kwargs(ref_path, K, T, **over) -> dict:
    """Constructor keywords of run.py:25-37."""
    kw = dict(delta_t=0.003 * 2, ref_path=ref_path, horizon_step_T=T, number_of_samples_K=K,
              param_exploration=0.0, param_lambda=100.0, param_alpha=0.98,
              sigma=np.array([[20.0, 0.0], [0.0, 20.0]]),
              stage_cost_weight=np.array([0.50, 0.50, 5.0, 5.0]),
              terminal_cost_weight=np.array([5.0, 5.0, 50.0, 50.0]))
    kw.update(over)
    return kw


def ctor_kwargs(case: dict, paths: dict) -> dict:
    ref = ref_path_for(paths, case["file"], case.get("layout", "verbatim"))
    kw = run_py_kwargs(ref, case["K"], case["T"])
    kw.update(case.get("ctor", {}))
    return kw


def load_golden(name: str) -> dict:
    """Return {case_name: {key.step: array}} from one of the committed .npz fixtures."""
    out = {}
    with np.load(os.path.join(HERE, name)) as z:
        for full in z.files:
            case, key = full.split("/", 1)
            out.setdefault(case, {})[key] = z[full]
    return out

"""Drop-in host mirror of the reference controller class.

``MPPIControllerForPathTracking`` keeps the reference's constructor keywords (including the
misspelt ``visualze_sampled_trajs``), public attributes (``u_prev``, ``prev_waypoints_idx``, ``K``,
``T``, ``Sigma`` ...), the ``calc_control_input(observed_x)`` call and its 4-tuple return, the
aliasing quirks of the returned sequence, the printed lines and the exception types
(reference: /root/reference/control.py:20-152, caller: run.py:25-51).  The step itself runs on the
GPU through libmppi_b200.so; this class only moves a few hundred bytes in and out per step.
"""
from __future__ import annotations

import numpy as np

from .engine import MppiEngine, ShardSpec

SEARCH_IDX_LEN = 30   # control.py:203 (fixed in the kernels as mppi::kWindow)


def _arm_params():
    """The reference reads sys_params.SYS_PARAMS() at import (control.py:10-18); prefer the user's
    module when this class is dropped into their checkout."""
    try:
        from sys_params import SYS_PARAMS        # the user's (or this repo's root-level) module
    except ImportError:
        from .arm_params import SYS_PARAMS
    return SYS_PARAMS()


class MPPIControllerForPathTracking:
    def __init__(
            self,
            delta_t: float = 0.01,
            ref_path=0,
            horizon_step_T: int = 20,
            number_of_samples_K: int = 500,
            param_exploration: float = 0.0,
            param_lambda: float = 50.0,
            param_alpha: float = 1.0,
            sigma: np.ndarray = np.array([[10.0, 10.0], [100.0, 100.0]]),
            stage_cost_weight: np.ndarray = np.array([10.0, 10.0, 10.0, 10.0]),
            terminal_cost_weight: np.ndarray = np.array([10.0, 10.0, 10.0, 10.0]),
            visualize_optimal_traj=True,
            visualze_sampled_trajs=False,
            *,
            noise: str = "philox",      # "philox": in-kernel draw; "numpy": self._calc_epsilon() injected
            seed=None,                  # Philox key; None -> drawn once from OS entropy (rank 0's draw in sharded runs)
            device=None,                # CUDA device (default: current)
            verbose: bool = True,       # print the reference's three lines per step (control.py:227-229)
            use_graph: bool = True,     # replay the step as a CUDA graph
            distributed: bool = False,  # shard the K samples over torch.distributed ranks
            process_group=None,
            exchange: str = "auto",     # sharded runs: "p2p" (exchange fused into the kernels, NVLink peer memory),
                                        #  "nccl" (all-gather) or "auto" = p2p where the peer mapping works, else nccl
            sampled_traj_top_n=None,    # with visualze_sampled_trajs: return only the n best, best first
            smoother: str = "median",  # "median" (control.py:122), "average" (control.py:329-344) or "none"
            dynamics: str = "F",        # rollout model: "F" (control.py:234-263) or "F1" (control.py:265-295)
            search: str = "certified",  # nearest-waypoint lookups: "certified" shortcut or always "full" (same results)
            search_stats: bool = False,  # count how many lookups the shortcut answered (engine.search_stats())
            joint_limit_lo=None,        # joint-limit stage cost (not in the reference: its clamps are commented out,
            joint_limit_hi=None,        #  control.py:166-172): (q1, q2) bounds in rad, None = unbounded
            joint_limit_weight: float = 0.0,   # weight of 1e4 * (violation of q1)^2 + (violation of q2)^2 per horizon step; 0 = off
    ) -> None:
        # same attributes as control.py:37-65
        self.dim_x = 4
        self.dim_u = 2
        self.T = horizon_step_T
        self.K = number_of_samples_K
        self.param_exploration = param_exploration
        self.param_lambda = param_lambda
        self.param_alpha = param_alpha
        self.param_gamma = self.param_lambda * (1.0 - (self.param_alpha))
        self.Sigma = sigma
        self.stage_cost_weight = stage_cost_weight
        self.terminal_cost_weight = terminal_cost_weight
        self.visualize_optimal_traj = visualize_optimal_traj
        self.visualze_sampled_trajs = visualze_sampled_trajs
        self.delta_t = delta_t
        self.ref_path = ref_path
        self.l1 = 1
        self.l2 = 1
        self.u_prev = np.array([[10.0, -2.0] for _ in range(self.T)])
        self.prev_waypoints_idx = 0

        if noise not in ("philox", "numpy"):
            raise ValueError("noise must be 'philox' or 'numpy'")
        self.noise = noise
        # (a private generator: the reference never touches numpy's global stream at construction)
        self.seed = int(np.random.default_rng().integers(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self._seed_is_drawn = seed is None
        self.verbose = verbose
        self._device = device
        self._use_graph = use_graph
        self._distributed = distributed
        self._group = process_group
        self._exchange = exchange
        self.sampled_traj_top_n = sampled_traj_top_n
        self.smoother = smoother
        self.dynamics = dynamics
        self._search = search
        self._search_stats = search_stats
        self.joint_limit_lo, self.joint_limit_hi = joint_limit_lo, joint_limit_hi
        self.joint_limit_weight = joint_limit_weight
        self._engine_obj = None
        self._engine_ref_path = None
        self._sigma_checked = None
        self._zero_view = None
        self.last = {}               # intermediates of the last step (rho, eta, raw / filtered update)

    # ---- engine life cycle ---------------------------------------------------------------------
    def _shard(self):
        if not self._distributed:
            return ShardSpec()
        import torch.distributed as dist
        return ShardSpec(rank=dist.get_rank(self._group), world=dist.get_world_size(self._group))

    def _engine(self) -> MppiEngine:
        if self._engine_obj is None:
            if self._distributed and self._seed_is_drawn:
                # every rank must key Philox on the SAME seed (samples are keyed on their global index, so the
                # result does not depend on the sharding): take rank 0's draw
                import torch
                import torch.distributed as dist
                dev = "cuda" if dist.get_backend(self._group) == "nccl" else "cpu"
                t = torch.tensor([self.seed], dtype=torch.int64, device=dev)
                dist.broadcast(t, src=dist.get_global_rank(self._group, 0) if self._group is not None else 0, group=self._group)
                self.seed = int(t.item())
                self._seed_is_drawn = False
            self._engine_obj = MppiEngine(
                K=self.K, T=self.T, delta_t=self.delta_t, param_lambda=self.param_lambda,
                param_gamma=self.param_gamma, sigma=self.Sigma, stage_cost_weight=self.stage_cost_weight,
                terminal_cost_weight=self.terminal_cost_weight, arm_params=_arm_params(),
                ref_path=self.ref_path, param_exploration=self.param_exploration,
                cost_l1=self.l1, cost_l2=self.l2, n_env=1, seed=self.seed, device=self._device,
                optimal_traj=bool(self.visualize_optimal_traj), use_graph=self._use_graph,
                smoother=self.smoother, shard=self._shard(), process_group=self._group,
                exchange=self._exchange, search=self._search, search_stats=self._search_stats,
                dynamics=self.dynamics, joint_limit_lo=self.joint_limit_lo, joint_limit_hi=self.joint_limit_hi,
                joint_limit_weight=self.joint_limit_weight)
            self._engine_ref_path = self.ref_path
        elif self.ref_path is not self._engine_ref_path:
            # the reference reads self.ref_path on every call (control.py:208); follow a re-assignment
            self._engine_obj.set_ref_path(self.ref_path)
            self._engine_ref_path = self.ref_path
        return self._engine_obj

    def close(self):
        if self._engine_obj is not None:
            self._engine_obj.close()
            self._engine_obj = None

    # ---- the step (control.py:67-152) ------------------------------------------------------------
    def calc_control_input(self, observed_x):
        u = self.u_prev                                   # alias on purpose (control.py:70)
        x0 = observed_x
        eng = self._engine()

        eps = None
        if self.noise == "numpy":
            eps = self._calc_epsilon(self.Sigma, self.K, self.T, self.dim_u)     # control.py:84
        elif self.Sigma is not self._sigma_checked:       # (the reference validates Sigma in every call, control.py:157-159)
            self._check_sigma(self.Sigma, self.dim_u)
            self._sigma_checked = self.Sigma

        prev_idx = self.prev_waypoints_idx
        eng.step(x0, u, prev_idx, eps)
        nearest_idx = int(eng.out_new_idx[0])
        if self.verbose:                                  # control.py:227-229
            print(f"0     prev_idx = {prev_idx}")
            print(f"0     nearest_idx = {nearest_idx}")
            print("======================updated=======================")
        self.prev_waypoints_idx = nearest_idx             # control.py:230
        if self.prev_waypoints_idx >= self.ref_path.shape[0] - 1:               # control.py:76-78
            print("[ERROR] Reached the end of the reference path.")
            raise IndexError

        w_epsilon = eng.out_w_eps_filt[0]
        self.last = dict(rho=float(eng.out_rho[0]), eta=float(eng.out_eta[0]),
                         w_eps_raw=eng.out_w_eps_raw[0].copy(), w_eps_filt=w_epsilon.copy())
        u += w_epsilon                                    # control.py:126 (in place: mutates u_prev)

        optimal_traj = eng.out_opt_traj[0].copy()         # zeros when visualize_optimal_traj is off

        if self.visualze_sampled_trajs:                   # control.py:137-145
            sampled_traj_list = self._gather_sampled(eng)
        elif self.K * self.T * self.dim_x <= (1 << 17):
            sampled_traj_list = np.zeros((self.K, self.T, self.dim_x))
        else:   # the reference allocates K*T*4 float64 zeros every call (26 MB at K=16384, T=50): return a
                # read-only zero view of the same shape and dtype instead
            if self._zero_view is None or self._zero_view.shape != (self.K, self.T, self.dim_x):
                self._zero_view = np.broadcast_to(np.zeros(()), (self.K, self.T, self.dim_x))
            sampled_traj_list = self._zero_view

        self.u_prev[:-1] = u[1:]                          # control.py:148
        self.u_prev[-1] = u[-1]                           # control.py:149
        return u[0], u, optimal_traj, sampled_traj_list   # control.py:152 (u[0] is post-shift, Q2)

    def run_closed_loop(self, observed_x, n_steps: int, plant_dt: float):
        """The loop of run.py:48-59 on the GPU: n_steps x { calc_control_input; dq += dt*Arm_Dynamic(q, dq, u);
        q += dt*dq } with no host round trip per tick (in-kernel Philox noise).

        Returns a dict of float64 arrays over the executed ticks: ``state`` [n, 4] (after each tick),
        ``u`` [n, 2] (the control applied), ``waypoint_idx`` [n], ``rho`` [n].  The controller's
        ``u_prev`` / ``prev_waypoints_idx`` end up as after the last tick.  Raises IndexError like the
        reference when the end of the path is reached (after returning state up to that tick in
        ``self.last_loop``)."""
        if self.noise != "philox":
            raise ValueError("run_closed_loop draws its noise in-kernel: construct with noise='philox'")
        self._check_sigma(self.Sigma, self.dim_u)
        eng = self._engine()
        x0 = np.asarray(observed_x, dtype=np.float64).reshape(4)
        log, stop = eng.closed_loop(x0, self.u_prev, self.prev_waypoints_idx, n_steps, plant_dt)
        n = int(min(n_steps, stop[0]))
        self.u_prev[...] = eng.in_u_prev[0]
        self.prev_waypoints_idx = int(eng.in_prev_idx[0])
        out = dict(state=log[:n, 0, 0:4].copy(), u=log[:n, 0, 4:6].copy(),
                   waypoint_idx=log[:n, 0, 6].astype(np.int64), rho=log[:n, 0, 7].copy(),
                   final_state=eng.in_x0[0].copy(), ticks=n)
        self.last_loop = out
        if n < n_steps:
            print("[ERROR] Reached the end of the reference path.")
            raise IndexError
        return out

    def _gather_sampled(self, eng):
        if self.sampled_traj_top_n and eng.shard.world == 1:
            # opt-in deviation from the reference's (K, T, 4): only the n lowest-cost samples, ordered
            # like np.argsort(S) (control.py:138); their sample indices are kept in self.last
            idx, traj = eng.best_sampled_trajectories(self.sampled_traj_top_n)
            self.last["sampled_idx"] = idx[0]
            return traj[0].cpu().numpy().astype(np.float64)
        local = eng.sampled_trajectories()[0].cpu().numpy().astype(np.float64)
        if eng.shard.world == 1:
            return local
        import torch
        import torch.distributed as dist
        parts = [None] * eng.shard.world
        dist.all_gather_object(parts, local, group=self._group)
        del torch
        return np.concatenate(parts, axis=0)

    # ---- helpers kept for API compatibility -------------------------------------------------------
    @staticmethod
    def _check_sigma(sigma, size_dim_u):
        sigma = np.asarray(sigma)
        if sigma.ndim != 2 or sigma.shape[0] != sigma.shape[1] or sigma.shape[0] != size_dim_u or size_dim_u < 1:
            print("[ERROR] sigma must be a square matrix with the size of size_dim_u.")
            raise ValueError

    def _calc_epsilon(self, sigma, size_sample, size_time_step, size_dim_u):
        """control.py:154-164 — the noise-injection seam: replace this bound method to feed a fixed
        [K, T, 2] tensor (that is how the reference itself is driven in the parity tests)."""
        self._check_sigma(sigma, size_dim_u)
        mu = np.zeros(size_dim_u)
        return np.random.multivariate_normal(mu, sigma, (size_sample, size_time_step))

    def _get_nearest_waypoint(self, q1, q2, update_prev_idx=False):
        """control.py:200-232 for callers that use it directly (the step does this on the GPU)."""
        p = self.prev_waypoints_idx
        x = self.l1 * np.cos(q1) + self.l2 * np.cos(q1 + q2)
        y = self.l1 * np.sin(q1) + self.l2 * np.sin(q1 + q2)
        win = self.ref_path[p:p + SEARCH_IDX_LEN]
        d = ((x - win[:, 0]) ** 2 + (y - win[:, 1]) ** 2) * 100
        nearest_idx = int(np.argmin(d)) + p
        row = self.ref_path[nearest_idx]
        if update_prev_idx:
            if self.verbose:
                print(f"0     prev_idx = {p}")
                print(f"0     nearest_idx = {nearest_idx}")
                print("======================updated=======================")
            self.prev_waypoints_idx = nearest_idx
        return nearest_idx, row[0], row[1], row[2], row[3]

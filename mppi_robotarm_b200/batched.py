"""Batched multi-environment MPPI (BASELINE.json configs[4]: many independent arm instances).

Every environment is one instance of the reference controller (control.py:20-152) with its own
observed state, nominal sequence and waypoint index; all of them are stepped by ONE launch of each
kernel (grid.y = environment).  Environments are independent, so multi-GPU runs shard the
*environments* over ranks and need no collective in the step.
"""
from __future__ import annotations

import numpy as np

from .controller import MPPIControllerForPathTracking, _arm_params
from .engine import MppiEngine


class BatchedMPPIController:
    def __init__(self, n_env: int, *, delta_t, ref_path, horizon_step_T, number_of_samples_K,
                 param_exploration=0.0, param_lambda=50.0, param_alpha=1.0, sigma=None,
                 stage_cost_weight=None, terminal_cost_weight=None, visualize_optimal_traj=False,
                 seed=0, device=None, use_graph=True, env_offset=0, search="certified", search_stats=False,
                 dynamics="F"):
        self.n_env = int(n_env)
        self.T, self.K = int(horizon_step_T), int(number_of_samples_K)
        self.param_lambda, self.param_alpha = param_lambda, param_alpha
        self.param_gamma = param_lambda * (1.0 - param_alpha)
        self.Sigma = np.asarray(sigma, dtype=np.float64)
        MPPIControllerForPathTracking._check_sigma(self.Sigma, 2)
        self.ref_path = ref_path
        self.u_prev = np.tile(np.array([10.0, -2.0]), (self.n_env, self.T, 1))       # control.py:59 per env
        self.prev_waypoints_idx = np.zeros(self.n_env, dtype=np.int64)               # control.py:65 per env
        self.finished = np.zeros(self.n_env, dtype=bool)
        self._want_opt = bool(visualize_optimal_traj)
        # env_offset decorrelates the Philox streams of environment shards living on different ranks
        self.engine = MppiEngine(
            K=self.K, T=self.T, delta_t=delta_t, param_lambda=param_lambda, param_gamma=self.param_gamma,
            sigma=self.Sigma, stage_cost_weight=stage_cost_weight, terminal_cost_weight=terminal_cost_weight,
            arm_params=_arm_params(), ref_path=ref_path, param_exploration=param_exploration, n_env=self.n_env,
            seed=int(seed) + 0x9E3779B97F4A7C15 * int(env_offset), device=device,
            optimal_traj=bool(visualize_optimal_traj), use_graph=use_graph, search=search,
            search_stats=search_stats, dynamics=dynamics)

    def calc_control_input(self, observed_x, eps=None, strict=False):
        """observed_x [n_env, 4] -> (u0 [n_env, 2], u_seq [n_env, T, 2], optimal_traj [n_env, T, 4] or None
        when visualize_optimal_traj is off).

        Per environment this is control.py:67-152 including the post-shift return value (Q2).
        Environments that reached the end of the path (control.py:76-78) are frozen and flagged in
        ``self.finished``; with strict=True the reference's IndexError is raised instead."""
        eng = self.engine
        x = np.asarray(observed_x, dtype=np.float64).reshape(self.n_env, 4)
        eng.step(x, self.u_prev, self.prev_waypoints_idx, eps)
        new_idx = eng.out_new_idx.astype(np.int64)
        ended = new_idx >= np.asarray(self.ref_path).shape[0] - 1
        self.prev_waypoints_idx = new_idx
        if strict and ended.any():
            print("[ERROR] Reached the end of the reference path.")
            raise IndexError
        live = ~ended & ~self.finished
        self.finished |= ended
        u = self.u_prev
        if live.all():                                   # common case: plain slices, no fancy indexing
            u += eng.out_w_eps_filt
            u[:, :-1] = u[:, 1:]
        else:
            u[live] += eng.out_w_eps_filt[live]
            u[live, :-1] = u[live, 1:]
        opt = eng.out_opt_traj.copy() if self._want_opt else None
        return u[:, 0].copy(), u, opt

    def run_closed_loop(self, observed_x, n_steps: int, plant_dt: float):
        """n_steps ticks of the run.py loop (controller + FP64 plant of utils.py:14-29) for all environments
        on the device, no host round trip per tick.  Returns (log, stop): log float64 [n_steps, n_env, 8] =
        (q1, q2, dq1, dq2, u1, u2, waypoint index, rho) after each tick; stop int32 [n_env] = first tick at
        which an environment reached the end of the path (it is frozen from there), >= n_steps if never."""
        eng = self.engine
        x = np.asarray(observed_x, dtype=np.float64).reshape(self.n_env, 4)
        log, stop = eng.closed_loop(x, self.u_prev, self.prev_waypoints_idx, n_steps, plant_dt)
        self.u_prev[...] = eng.in_u_prev
        self.prev_waypoints_idx = eng.in_prev_idx.astype(np.int64)
        self.finished |= stop < n_steps
        return log, stop

    def close(self):
        self.engine.close()

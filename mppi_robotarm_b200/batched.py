"""Batched multi-environment MPPI (BASELINE.json configs[4]: many independent arm instances).

Every environment is one instance of the reference controller (control.py:20-152) with its own
observed state, nominal sequence and waypoint index; all of them are stepped by ONE launch of each
kernel (grid.y = environment).  Environments are independent, so multi-GPU runs shard the
*environments* over ranks and need no collective in the step.

The controller state (nominal sequences, waypoint indices, step counter) lives on the device between
steps (MPPI_FLAG_RESIDENT_STATE): a step moves only the observed states in (32 B per environment) and
the controls + waypoint indices out (36 B per environment) — the sequences are shifted by the kernels
(control.py:148-149) and fetched only when somebody reads ``u_prev``.
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
                 dynamics="F", joint_limit_lo=None, joint_limit_hi=None, joint_limit_weight=0.0,
                 return_sequences=False):
        self.n_env = int(n_env)
        self.T, self.K = int(horizon_step_T), int(number_of_samples_K)
        self.param_lambda, self.param_alpha = param_lambda, param_alpha
        self.param_gamma = param_lambda * (1.0 - param_alpha)
        self.Sigma = np.asarray(sigma, dtype=np.float64)
        MPPIControllerForPathTracking._check_sigma(self.Sigma, 2)
        self.ref_path = ref_path
        self._u_prev = np.tile(np.array([10.0, -2.0]), (self.n_env, self.T, 1))      # control.py:59 per env
        self._prev_idx = np.zeros(self.n_env, dtype=np.int64)                        # control.py:65 per env
        self._host_fresh = True        # the host copies above are the truth (not yet / no longer on the device)
        self._dev_fresh = False        # the device holds the current controller state
        self.finished = np.zeros(self.n_env, dtype=bool)
        self._want_opt = bool(visualize_optimal_traj)
        self.return_sequences = bool(return_sequences)
        # env_offset decorrelates the Philox streams of environment shards living on different ranks
        self.engine = MppiEngine(
            K=self.K, T=self.T, delta_t=delta_t, param_lambda=param_lambda, param_gamma=self.param_gamma,
            sigma=self.Sigma, stage_cost_weight=stage_cost_weight, terminal_cost_weight=terminal_cost_weight,
            arm_params=_arm_params(), ref_path=ref_path, param_exploration=param_exploration, n_env=self.n_env,
            seed=int(seed) + 0x9E3779B97F4A7C15 * int(env_offset), device=device,
            optimal_traj=bool(visualize_optimal_traj), use_graph=use_graph, search=search,
            search_stats=search_stats, dynamics=dynamics, joint_limit_lo=joint_limit_lo,
            joint_limit_hi=joint_limit_hi, joint_limit_weight=joint_limit_weight, resident_state=True)

    # ---- controller state: host attributes like the reference's, backed by the device ---------------------
    def _pull(self):
        """Make the host copies current (one device -> host transfer of the sequences, on demand)."""
        if not self._host_fresh:
            eng = self.engine
            eng.download_state()
            self._u_prev[...] = eng.in_u_prev
            self._prev_idx[...] = eng.in_prev_idx
            self._host_fresh = True

    @property
    def u_prev(self):
        """Nominal sequences [n_env, T, 2].  Reading fetches them from the device; the returned array may be
        modified in place (it is pushed back before the next step)."""
        self._pull()
        self._dev_fresh = False
        return self._u_prev

    @u_prev.setter
    def u_prev(self, v):
        self._pull()
        self._u_prev[...] = np.asarray(v, dtype=np.float64).reshape(self.n_env, self.T, 2)
        self._dev_fresh = False

    @property
    def prev_waypoints_idx(self):
        self._pull()
        self._dev_fresh = False
        return self._prev_idx

    @prev_waypoints_idx.setter
    def prev_waypoints_idx(self, v):
        self._pull()
        self._prev_idx[...] = np.asarray(v, dtype=np.int64).reshape(self.n_env)
        self._dev_fresh = False

    def _push(self):
        if not self._dev_fresh:
            eng = self.engine
            eng.in_u_prev[...] = self._u_prev
            eng.in_prev_idx[...] = self._prev_idx
            eng.upload_state()
            self._dev_fresh = True

    # ---- the step -------------------------------------------------------------------------------------
    def calc_control_input(self, observed_x, eps=None, strict=False):
        """observed_x [n_env, 4] -> (u0 [n_env, 2], u_seq, optimal_traj).

        Per environment this is control.py:67-152 including the post-shift return value (Q2).  u_seq
        [n_env, T, 2] and optimal_traj [n_env, T, 4] are returned only by controllers built with
        return_sequences=True (they cost a 5 KB transfer per environment at T = 64); otherwise they are None
        and ``self.u_prev`` fetches the sequences on demand.
        Environments that reached the end of the path (control.py:76-78) are frozen and flagged in
        ``self.finished``; with strict=True the reference's IndexError is raised instead."""
        eng = self.engine
        self._push()
        eng.step_resident(observed_x, eps)
        self._host_fresh = False
        new_idx = eng.out_new_idx.astype(np.int64)
        ended = new_idx >= np.asarray(self.ref_path).shape[0] - 1
        if strict and ended.any():
            print("[ERROR] Reached the end of the reference path.")
            raise IndexError
        self.finished |= ended
        u0 = eng.out_u0.copy()
        if not self.return_sequences:
            return u0, None, None
        self._pull()
        return u0, self._u_prev, (eng.out_opt_traj.copy() if self._want_opt else None)

    def last_waypoint_idx(self):
        """Waypoint indices after the last step (control.py:230), without fetching the sequences."""
        return self.engine.out_new_idx.astype(np.int64)

    def run_closed_loop(self, observed_x, n_steps: int, plant_dt: float):
        """n_steps ticks of the run.py loop (controller + FP64 plant of utils.py:14-29) for all environments
        on the device, no host round trip per tick.  Returns (log, stop): log float64 [n_steps, n_env, 8] =
        (q1, q2, dq1, dq2, u1, u2, waypoint index, rho) after each tick; stop int32 [n_env] = first tick at
        which an environment reached the end of the path (it is frozen from there), >= n_steps if never."""
        eng = self.engine
        self._pull()
        x = np.asarray(observed_x, dtype=np.float64).reshape(self.n_env, 4)
        log, stop = eng.closed_loop(x, self._u_prev, self._prev_idx, n_steps, plant_dt)
        self._u_prev[...] = eng.in_u_prev
        self._prev_idx[...] = eng.in_prev_idx
        self._host_fresh, self._dev_fresh = True, False
        self.finished |= stop < n_steps
        return log, stop

    def close(self):
        self.engine.close()

"""MppiEngine — owns one libmppi_b200 handle plus the memory it works in.

PyTorch is used here for plumbing only: device memory (the workspace and the optional injected-noise
tensor), pinned host memory (the io block), the CUDA stream, and ``torch.distributed`` for the one
small all-gather of the sharded step.  All arithmetic of the MPPI step happens inside the CUDA
library; nothing here (or anywhere in this package) computes rollouts on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _cabi
from .arm_params import arm_vector


@dataclass
class ShardSpec:
    """Which global samples [k_offset, k_offset + K_local) this rank rolls out."""
    rank: int = 0
    world: int = 1

    def bounds(self, K_total: int):
        base, rem = divmod(K_total, self.world)
        K_local = base + (1 if self.rank < rem else 0)
        k_offset = self.rank * base + min(self.rank, rem)
        return k_offset, K_local


def noise_factor(sigma: np.ndarray) -> np.ndarray:
    """Lower-triangular L with L L^T = Sigma for the in-kernel draw eps = L z.  The reference draws with
    np.random.multivariate_normal (control.py:163), which accepts any symmetric positive SEMI-definite Sigma
    (SVD factor) and only warns otherwise: a positive definite Sigma gets its Cholesky factor, a semi-definite
    one a factor from the eigen-decomposition; an asymmetric or indefinite Sigma is an error here, because the
    noise covariance would not be the Sigma whose inverse enters the control cost (control.py:106)."""
    sigma = np.asarray(sigma, dtype=np.float64)
    if not np.allclose(sigma, sigma.T, rtol=1e-12, atol=1e-12 * max(1.0, float(np.max(np.abs(sigma))))):
        raise np.linalg.LinAlgError("Sigma must be symmetric (noise covariance and control-cost weight must agree)")
    try:
        return np.linalg.cholesky(sigma)
    except np.linalg.LinAlgError:
        w, v = np.linalg.eigh(sigma)
        if w.min() < -1e-10 * max(1.0, w.max()):
            raise
        a = v * np.sqrt(np.clip(w, 0.0, None))              # A A^T = Sigma; make it lower triangular by QR
        q, r = np.linalg.qr(a.T)
        L = r.T
        return L * np.sign(np.where(np.diag(L) == 0, 1.0, np.diag(L)))[None, :]


def exploit_count(K: int, exploration: float) -> int:
    """#samples with ``k < (1 - param_exploration) * K`` — the Python float comparison of
    control.py:98 evaluated once on the host (an integer threshold on the global sample index)."""
    thr = (1.0 - float(exploration)) * K
    n = int(np.ceil(thr))
    n = max(0, min(K, n))
    # make the boundary exact under the same float comparison
    while n > 0 and not ((n - 1) < thr):
        n -= 1
    while n < K and (n < thr):
        n += 1
    return n


class MppiEngine:
    def __init__(self, *, K, T, delta_t, param_lambda, param_gamma, sigma, stage_cost_weight,
                 terminal_cost_weight, arm_params, ref_path, param_exploration=0.0, cost_l1=1.0, cost_l2=1.0,
                 n_env=1, seed=0, device=None, optimal_traj=True, use_graph=True, smoother="median",
                 shard: ShardSpec | None = None, process_group=None, max_ref_rows=None, exchange="auto",
                 search="certified", search_stats=False, dynamics="F", joint_limit_lo=None, joint_limit_hi=None,
                 joint_limit_weight=0.0, resident_state=False, exchange_timeout_ms=None):
        import torch
        self.torch = torch
        self.lib = _cabi.load()
        if not torch.cuda.is_available() or self.lib.mppi_device_count() < 1:
            raise _cabi.NativeLibraryError(
                "no CUDA device of compute capability 10.x visible: the MPPI step runs only as sm_100a "
                "CUDA (libmppi_b200.so); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else
                                   (device if isinstance(device, int) else torch.device(device).index or 0))
        self.shard = shard or ShardSpec()
        self.group = process_group
        self.K, self.T, self.n_env = int(K), int(T), int(n_env)
        k_offset, K_local = self.shard.bounds(self.K)
        if K_local < 1:
            raise ValueError(f"rank {self.shard.rank} of {self.shard.world} would get no samples (K={K})")
        self.k_offset, self.K_local = k_offset, K_local
        sigma = np.asarray(sigma, dtype=np.float64)
        sig_inv = np.linalg.inv(sigma)                       # LinAlgError on a singular Sigma (control.py:106)
        chol = noise_factor(sigma)                           # factor of the in-kernel draw (unused with injected noise)
        ref = np.ascontiguousarray(np.asarray(ref_path, dtype=np.float64)[:, 0:4])
        cfg = _cabi.MppiConfig()
        cfg.abi_version = _cabi.ABI_VERSION
        cfg.device = self.device.index
        cfg.n_env = self.n_env
        cfg.K_total, cfg.K_local, cfg.k_offset = self.K, K_local, k_offset
        cfg.T = self.T
        cfg.n_exploit = exploit_count(self.K, param_exploration)
        if smoother not in ("median", "average", "none"):
            raise ValueError("smoother must be 'median', 'average' or 'none'")
        if smoother == "average" and self.T < 10:
            raise ValueError("the moving-average smoother needs horizon_step_T >= 10 (np.convolve 'same', control.py:338)")
        if search not in ("certified", "full"):
            raise ValueError("search must be 'certified' (lookups answered by per-window certificates, bit-identical results) or 'full'")
        if dynamics not in ("F", "F1"):
            raise ValueError("dynamics must be 'F' (control.py:234-263) or 'F1' (control.py:265-295)")
        cfg.flags = ((_cabi.FLAG_OPTIMAL_TRAJ if optimal_traj else 0) | (_cabi.FLAG_DEVICE_GRAPH if use_graph else 0)
                     | {"median": 0, "average": _cabi.FLAG_SMOOTH_AVERAGE, "none": _cabi.FLAG_SMOOTH_NONE}[smoother]
                     | (_cabi.FLAG_FULL_SEARCH if search == "full" else 0)
                     | (_cabi.FLAG_SEARCH_STATS if search_stats else 0)
                     | (_cabi.FLAG_DYNAMICS_F1 if dynamics == "F1" else 0)
                     | (_cabi.FLAG_RESIDENT_STATE if resident_state else 0))
        self.resident_state = bool(resident_state)
        if self.resident_state and self.shard.world != 1:
            raise ValueError("resident_state keeps the whole controller state on one GPU (unsharded handles only)")
        cfg.max_ref_rows = int(max_ref_rows or ref.shape[0])
        cfg.delta_t, cfg.param_lambda, cfg.param_gamma = float(delta_t), float(param_lambda), float(param_gamma)
        cfg.sigma_chol[:] = chol.reshape(-1).tolist()
        cfg.sigma_inv[:] = sig_inv.reshape(-1).tolist()
        cfg.stage_cost_weight[:] = np.asarray(stage_cost_weight, dtype=np.float64).reshape(-1)[:4].tolist()
        cfg.terminal_cost_weight[:] = np.asarray(terminal_cost_weight, dtype=np.float64).reshape(-1)[:4].tolist()
        cfg.arm[:] = arm_vector(arm_params)
        cfg.cost_l1, cfg.cost_l2 = float(cost_l1), float(cost_l2)
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        # joint-limit stage cost (extension; weight 0 = the reference's step, and kernels without the term)
        lo = (-np.inf, -np.inf) if joint_limit_lo is None else tuple(float(v) for v in joint_limit_lo)
        hi = (np.inf, np.inf) if joint_limit_hi is None else tuple(float(v) for v in joint_limit_hi)
        if len(lo) != 2 or len(hi) != 2:
            raise ValueError("joint_limit_lo / joint_limit_hi need one value per joint (q1, q2)")
        cfg.joint_limit_lo[:] = lo
        cfg.joint_limit_hi[:] = hi
        cfg.joint_limit_weight = float(joint_limit_weight)
        self.cfg = cfg

        lay = _cabi.MppiIoLayout()
        _cabi.check(self.lib.mppi_io_layout(C.byref(cfg), C.byref(lay)), None, "mppi_io_layout")
        self.layout = lay
        ws_bytes = self.lib.mppi_workspace_bytes(C.byref(cfg))
        if ws_bytes == 0:
            raise ValueError("libmppi_b200 rejected the configuration: " + _cabi.last_error())
        with torch.cuda.device(self.device):
            self._ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=self.device)
            self._io = torch.zeros(lay.bytes, dtype=torch.uint8).pin_memory()
            self.stream = torch.cuda.Stream(device=self.device)
        ws_ptr = (self._ws.data_ptr() + 255) // 256 * 256
        h = C.c_void_p()
        _cabi.check(self.lib.mppi_create(C.byref(cfg), ws_ptr, ws_bytes, self._io.data_ptr(), lay.bytes, C.byref(h)),
                    None, "mppi_create")
        self.handle = h
        io = self._io.numpy()
        E, T = self.n_env, self.T

        def view(off, dtype, shape):
            n = int(np.prod(shape)) * np.dtype(dtype).itemsize
            return io[off:off + n].view(dtype).reshape(shape)

        self.in_x0 = view(lay.off_x0, np.float64, (E, 4))
        self.in_u_prev = view(lay.off_u_prev, np.float64, (E, T, 2))
        self.in_prev_idx = view(lay.off_prev_idx, np.int32, (E,))
        self.in_step = view(lay.off_step, np.uint64, (1,))
        self.out_new_idx = view(lay.off_new_idx, np.int32, (E,))
        self.out_status = view(lay.off_status, np.int32, (E,))
        self.out_rho = view(lay.off_rho, np.float64, (E,))
        self.out_eta = view(lay.off_eta, np.float64, (E,))
        self.out_u0 = view(lay.off_u0, np.float64, (E, 2))
        self.out_w_eps_raw = view(lay.off_w_eps_raw, np.float64, (E, T, 2))
        self.out_w_eps_filt = view(lay.off_w_eps_filt, np.float64, (E, T, 2))
        self.out_u_new = view(lay.off_u_new, np.float64, (E, T, 2))
        self.out_opt_traj = view(lay.off_opt_traj, np.float64, (E, T, 4))
        self.step_counter = 0
        self._eps_dev = None
        self._eps_pin = None
        self._partial = None
        self._gathered = None
        self._dist_graph = None
        self.use_graph = bool(use_graph)
        self._symm = None
        if exchange not in ("auto", "nccl", "p2p"):
            raise ValueError("exchange must be 'auto' (p2p where the peer mapping works, else nccl), 'nccl' or 'p2p'")
        self.exchange = exchange if self.shard.world > 1 else "none"
        self.set_ref_path(ref)
        if self.exchange == "auto":
            import torch.distributed as dist
            # (a shard without a process group — partials gathered by the caller — has nobody to map buffers with)
            live = dist.is_available() and dist.is_initialized()
            self.exchange = "p2p" if live and self._try_peer_exchange() else "nccl"
        elif self.exchange == "p2p":
            self._setup_peer_exchange()
        if self.exchange == "p2p":
            if exchange_timeout_ms is not None:
                _cabi.check(self.lib.mppi_set_exchange_timeout(self.handle, float(exchange_timeout_ms)), self.handle,
                            "mppi_set_exchange_timeout")

    def _try_peer_exchange(self) -> bool:
        """exchange="auto": map the peer buffers if every rank can; all ranks take the same decision."""
        torch = self.torch
        import torch.distributed as dist
        try:
            self._setup_peer_exchange()
            ok = 1
        except Exception:                                   # noqa: BLE001 (no symmetric memory / no peer access)
            ok = 0
        flag = torch.tensor([ok], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            self._symm = None
            return False
        return True

    def _setup_peer_exchange(self):
        """Map one exchange buffer per rank into every rank's address space (torch symmetric memory over
        NVLink) and hand the table of peer addresses to the library: the kernels then exchange the
        shard partials themselves (mppi_step_sharded) — no NCCL call in the control step."""
        torch = self.torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        world, rank = self.shard.world, self.shard.rank
        nbytes = int(self.lib.mppi_exchange_bytes(C.byref(self.cfg), world))
        if nbytes == 0:
            raise ValueError("libmppi_b200: " + _cabi.last_error())
        group = self.group if self.group is not None else dist.group.WORLD
        with torch.cuda.device(self.device):
            buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, group)
            torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)                    # every buffer is zeroed before anyone may write to it
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        assert len(ptrs) == world and ptrs[rank] == buf.data_ptr()
        table = (C.c_void_p * world)(*ptrs)
        _cabi.check(self.lib.mppi_set_peer_exchange(self.handle, rank, world, table), self.handle, "mppi_set_peer_exchange")
        self._symm = (buf, hdl)

    # ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "handle", None):
            self.torch.cuda.synchronize(self.device)
            # a captured graph that contains NCCL kernels must be gone before the process group is
            # destroyed (otherwise destroy_process_group() blocks)
            self._dist_graph = None
            self._gathered = self._partial = self._gathered_keepalive = None
            self.lib.mppi_destroy(self.handle)
            self.handle = None
            self._symm = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_ref_path(self, ref):
        ref = np.ascontiguousarray(np.asarray(ref, dtype=np.float64)[:, 0:4])
        if ref.shape[0] > self.cfg.max_ref_rows:
            raise ValueError(f"reference path has {ref.shape[0]} rows; this engine was sized for {self.cfg.max_ref_rows} "
                             "(pass max_ref_rows= at construction)")
        self.torch.cuda.synchronize(self.device)
        self._dist_graph = None          # captured graphs carry the old row count in their kernel arguments
        self.n_ref_rows = ref.shape[0]
        _cabi.check(self.lib.mppi_set_ref_path(self.handle, ref.ctypes.data, ref.shape[0]), self.handle,
                    "mppi_set_ref_path")

    # ------------------------------------------------------------------------------------------
    def _stage_eps(self, eps):
        """Host noise [n_env?, K_local, T, 2] -> device float32 (injected-noise mode)."""
        torch = self.torch
        shape = (self.n_env, self.K_local, self.T, 2)
        if torch.is_tensor(eps) and eps.is_cuda:
            e = eps.to(torch.float32).reshape(shape).contiguous()
            self._eps_dev = e
            return e
        eps = np.asarray(eps)
        if eps.shape[0] == self.K and self.K_local != self.K and eps.ndim == 3:
            eps = eps[self.k_offset:self.k_offset + self.K_local]        # this rank's shard
        if self._eps_dev is None or tuple(self._eps_dev.shape) != shape or not self._eps_dev.is_cuda:
            self._eps_dev = torch.empty(shape, dtype=torch.float32, device=self.device)
            self._eps_pin = torch.empty(shape, dtype=torch.float32).pin_memory()
        self._eps_pin.numpy()[...] = eps.reshape(shape)                  # float64 -> float32 here
        with torch.cuda.stream(self.stream):
            self._eps_dev.copy_(self._eps_pin, non_blocking=True)
        return self._eps_dev

    def write_inputs(self, x0, u_prev, prev_idx):
        if self.n_env == 1:            # the drop-in class: plain assignments (numpy converts), no temporaries
            self.in_x0[0] = x0
            self.in_u_prev[0] = u_prev
            self.in_prev_idx[0] = prev_idx
        else:
            self.in_x0[...] = np.asarray(x0, dtype=np.float64).reshape(self.n_env, 4)
            self.in_u_prev[...] = np.asarray(u_prev, dtype=np.float64).reshape(self.n_env, self.T, 2)
            self.in_prev_idx[...] = np.asarray(prev_idx, dtype=np.int64).reshape(self.n_env)
        self.in_step[0] = self.step_counter

    def step(self, x0, u_prev, prev_idx, eps=None):
        """One MPPI step.  Results are then readable from the ``out_*`` views (valid until the next
        step).  ``eps`` = None draws Philox noise in-kernel; otherwise it is injected."""
        self.write_inputs(x0, u_prev, prev_idx)
        if self.resident_state:
            self.upload_state()
        self.launch(eps)
        self.wait()
        if self.resident_state:
            self.download_state()

    # ---- resident controller state (many environments: only x0 in, compact results out) ----------------
    def upload_state(self):
        """Push in_x0 / in_u_prev / in_prev_idx and the step counter to the device (resident_state engines: once
        before the first step and after every host-side change of the controller state)."""
        self.in_step[0] = self.step_counter
        _cabi.check(self.lib.mppi_upload_state(self.handle, self.stream.cuda_stream), self.handle, "mppi_upload_state")

    def download_state(self):
        """Fetch the device's controller state (in_u_prev, in_prev_idx) and the full outputs of the last step."""
        _cabi.check(self.lib.mppi_download_state(self.handle, self.stream.cuda_stream), self.handle, "mppi_download_state")
        _cabi.check(self.lib.mppi_wait(self.handle), self.handle, "mppi_wait")

    def step_resident(self, x0, eps=None):
        """One step of a resident_state engine: only the observed states go in; afterwards out_new_idx, out_status,
        out_rho, out_eta and out_u0 are valid (everything else stays on the device until download_state())."""
        self.in_x0[...] = np.asarray(x0, dtype=np.float64).reshape(self.n_env, 4)
        self.launch(eps)
        self.wait()

    def launch(self, eps=None):
        mode = _cabi.NOISE_PHILOX if eps is None else _cabi.NOISE_INJECTED
        eps_ptr = None if eps is None else self._stage_eps(eps).data_ptr()
        s = self.stream.cuda_stream
        if self.shard.world == 1:
            _cabi.check(self.lib.mppi_step(self.handle, mode, eps_ptr, s), self.handle, "mppi_step")
        else:
            self._launch_sharded(mode, eps_ptr, s)
        self.last_mode, self.last_eps_ptr = mode, eps_ptr
        self.step_counter += 1

    def _launch_sharded(self, mode, eps_ptr, s):
        """rollouts on this shard -> all-gather of (rho_g, eta_g, V_g) -> identical combine on every rank.

        In Philox mode with use_graph the whole sequence (our kernels + NCCL's all-gather) is captured
        once into a torch.cuda.CUDAGraph and replayed: one launch per control step instead of ten."""
        torch = self.torch
        import torch.distributed as dist
        if self.exchange == "p2p":
            _cabi.check(self.lib.mppi_step_sharded(self.handle, mode, eps_ptr, s), self.handle, "mppi_step_sharded")
            return
        if self.use_graph and mode == _cabi.NOISE_PHILOX:
            if self._dist_graph is None:
                self._sharded_eager(mode, eps_ptr, dist)          # NCCL must have run once eagerly
                self.wait()
                g = torch.cuda.CUDAGraph()
                _cabi.check(self.lib.mppi_set_capture_mode(self.handle, 1), self.handle, "capture on")
                try:
                    with torch.cuda.graph(g, stream=self.stream, capture_error_mode="thread_local"):
                        self._sharded_eager(mode, eps_ptr, dist)
                finally:
                    _cabi.check(self.lib.mppi_set_capture_mode(self.handle, 0), self.handle, "capture off")
                self._dist_graph = g
            _cabi.check(self.lib.mppi_replay_begin(self.handle, s), self.handle, "mppi_replay_begin")
            with torch.cuda.stream(self.stream):
                self._dist_graph.replay()
            _cabi.check(self.lib.mppi_replay_end(self.handle, s), self.handle, "mppi_replay_end")
            return
        self._sharded_eager(mode, eps_ptr, dist)

    def _sharded_eager(self, mode, eps_ptr, dist):
        torch = self.torch
        self.launch_local(mode, eps_ptr)
        if self._gathered is None:
            self._gathered = torch.zeros(self.shard.world * self._partial.numel(), dtype=torch.float64,
                                         device=self.device)
        with torch.cuda.stream(self.stream):
            dist.all_gather_into_tensor(self._gathered, self._partial, group=self.group)
        self.launch_combine(self._gathered, self.shard.world)

    def launch_local(self, mode, eps_ptr):
        """First half of the sharded step; returns this shard's partial, float64 [n_env, 2 + 2T] on
        the device: (rho_g, eta_g, V_g[T, 2])."""
        if self._partial is None:
            self._partial = self.torch.zeros((self.n_env, 2 + 2 * self.T), dtype=self.torch.float64,
                                             device=self.device)
        _cabi.check(self.lib.mppi_step_local(self.handle, mode, eps_ptr, self._partial.data_ptr(),
                                             self.stream.cuda_stream), self.handle, "mppi_step_local")
        return self._partial

    def launch_combine(self, gathered, world):
        """Second half: ``gathered`` is float64 [world, n_env, 2 + 2T] on the device."""
        assert gathered.is_cuda and gathered.dtype == self.torch.float64 and gathered.is_contiguous()
        assert gathered.numel() == world * self.n_env * (2 + 2 * self.T)
        self._gathered_keepalive = gathered
        _cabi.check(self.lib.mppi_step_combine(self.handle, gathered.data_ptr(), world, self.stream.cuda_stream),
                    self.handle, "mppi_step_combine")

    def wait(self):
        _cabi.check(self.lib.mppi_wait(self.handle), self.handle, "mppi_wait")
        if self.exchange == "p2p" and self.out_status.any():
            # the step skipped its update (u_new = u_prev) instead of combining a stale partial; the next step is clean
            raise RuntimeError("peer exchange timed out: a rank did not deliver its partial in time; this step's "
                               "update was skipped (u_new == u_prev)")

    def closed_loop(self, x0, u_prev, prev_idx, n_steps, plant_dt):
        """n_steps ticks of the run.py loop entirely on the device (Philox noise, FP64 plant).

        Returns (log, stop): log float64 [n_steps, n_env, 8] = (q1, q2, dq1, dq2, u1, u2, waypoint idx, rho)
        after each tick, stop int32 [n_env] = first tick that hit the end of the path (>= n_steps: none).
        Afterwards in_x0 / in_u_prev / in_prev_idx hold the final controller state."""
        torch = self.torch
        if self.shard.world != 1:
            raise ValueError("the device closed loop runs on one GPU (whole sample set on this handle)")
        self.write_inputs(x0, u_prev, prev_idx)
        log = torch.empty((int(n_steps), self.n_env, 8), dtype=torch.float64, device=self.device)
        stop = torch.empty((self.n_env,), dtype=torch.int32, device=self.device)
        _cabi.check(self.lib.mppi_closed_loop(self.handle, int(n_steps), float(plant_dt), log.data_ptr(),
                                              stop.data_ptr(), self.stream.cuda_stream), self.handle, "mppi_closed_loop")
        self.last_mode, self.last_eps_ptr = _cabi.NOISE_PHILOX, None
        self.wait()
        # (a resident engine's device counter holds the step last used, the host-driven one the next step)
        self.step_counter = int(self.in_step[0]) + (1 if self.resident_state else 0)
        with torch.cuda.stream(self.stream):
            out = log.cpu().numpy(), stop.cpu().numpy()
        return out

    # ------------------------------------------------------------------------------------------
    def last_costs(self):
        """(S, w~) of the last step as device tensors [n_env, K_local] (float32)."""
        torch = self.torch
        ps, pw = C.c_void_p(), C.c_void_p()
        _cabi.check(self.lib.mppi_last_costs(self.handle, C.byref(ps), C.byref(pw)), self.handle, "mppi_last_costs")
        off = self._ws.data_ptr()
        n = self.n_env * self.K_local

        def as_tensor(ptr):
            start = ptr - off
            return self._ws[start:start + 4 * n].view(torch.float32).reshape(self.n_env, self.K_local)
        return as_tensor(ps.value), as_tensor(pw.value)

    def step_block(self, env=0):
        """Raw bytes (uint8 array) of the tables the prepare kernel built for ``env`` in the last step."""
        ptr, n = C.c_void_p(), C.c_size_t()
        _cabi.check(self.lib.mppi_step_block(self.handle, int(env), C.byref(ptr), C.byref(n)), self.handle, "mppi_step_block")
        self.stream.synchronize()
        start = ptr.value - self._ws.data_ptr()
        return self._ws[start:start + n.value].cpu().numpy()

    def sampled_trajectories(self):
        """control.py:137-145 for the last step: device tensor [n_env, K_local, T, 4] float32."""
        torch = self.torch
        out = torch.empty((self.n_env, self.K_local, self.T, 4), dtype=torch.float32, device=self.device)
        _cabi.check(self.lib.mppi_sampled_trajectories(self.handle, self.last_mode, self.last_eps_ptr,
                                                       out.data_ptr(), self.stream.cuda_stream),
                    self.handle, "mppi_sampled_trajectories")
        self.stream.synchronize()
        return out

    def best_sampled_trajectories(self, n):
        """Trajectories of the n lowest-cost samples of the last step, best first (the order of
        np.argsort(S) in control.py:138): (indices int64 [n_env, n], traj float32 tensor [n_env, n, T, 4]).
        The ranking itself is host logic like in the reference (a partial sort of the K costs)."""
        torch = self.torch
        n = int(min(n, self.K_local))
        self.stream.synchronize()
        S = self.last_costs()[0].cpu().numpy()
        part = np.argpartition(S, n - 1, axis=1)[:, :n] if n < self.K_local else np.tile(np.arange(n), (self.n_env, 1))
        order = np.take_along_axis(part, np.argsort(np.take_along_axis(S, part, axis=1), axis=1, kind="stable"), axis=1)
        idx = torch.from_numpy(order.astype(np.int32)).to(self.device)
        out = torch.empty((self.n_env, n, self.T, 4), dtype=torch.float32, device=self.device)
        torch.cuda.current_stream(self.device).synchronize()
        _cabi.check(self.lib.mppi_sampled_trajectories_subset(self.handle, self.last_mode, self.last_eps_ptr,
                                                              idx.data_ptr(), n, out.data_ptr(), self.stream.cuda_stream),
                    self.handle, "mppi_sampled_trajectories_subset")
        self.stream.synchronize()
        return order.astype(np.int64), out

    def philox_noise(self, step=None):
        """The noise tensor the kernels draw at control step ``step``: [n_env, K_local, T, 2] float32."""
        torch = self.torch
        out = torch.empty((self.n_env, self.K_local, self.T, 2), dtype=torch.float32, device=self.device)
        step = self.step_counter if step is None else step
        _cabi.check(self.lib.mppi_philox_noise(self.handle, int(step), out.data_ptr(), self.stream.cuda_stream),
                    self.handle, "mppi_philox_noise")
        self.stream.synchronize()
        return out

    def search_stats(self, reset=True) -> dict:
        """Nearest-waypoint lookups of the rollouts since the last reset, per warp of 32 samples: how many a
        certified end-of-window test answered (``certified``), how many a certified three-row comparison
        (``triples``), and how many there were; the rest ran the 30-candidate search
        (engine built with search_stats=True)."""
        buf = (C.c_uint64 * 3)()
        _cabi.check(self.lib.mppi_search_stats(self.handle, buf, 1 if reset else 0, self.stream.cuda_stream),
                    self.handle, "mppi_search_stats")
        n = int(buf[1])
        return {"certified": int(buf[0]), "triples": int(buf[2]), "lookups": n,
                "fraction": (buf[0] / n) if n else 0.0, "triple_fraction": (buf[2] / n) if n else 0.0,
                "searched_fraction": ((n - buf[0] - buf[2]) / n) if n else 0.0}

    def launch_count(self) -> int:
        return int(self.lib.mppi_launch_count(self.handle))

    def set_timing(self, on: bool):
        _cabi.check(self.lib.mppi_set_timing(self.handle, 1 if on else 0), self.handle, "mppi_set_timing")

    def get_timing(self) -> dict:
        buf = (C.c_double * 6)()
        n = self.lib.mppi_get_timing(self.handle, buf, 6)
        names = ["prepare", "rollout", "softmin", "wsum", "reduce", "finalize"]
        return {"steps": n, **{k: buf[i] for i, k in enumerate(names)}}

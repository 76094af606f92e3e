"""world_size-2 (and 3) gloo test of the N>1 path's host logic, on CPU: each rank owns the
contiguous sample shard ShardSpec gives it, forms its partial triple (rho_g, eta_g, V_g) — here
with the FP64 oracle standing in for the rollout kernels — all-gathers the partials exactly as
MppiEngine._launch_sharded does, and combines them with the formula of mppi_finalize_sm100a.
The combined update must equal the single-process oracle step."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, K, T, seed, out_dir):
    sys.path.insert(0, ROOT)
    from mppi_robotarm_b200.engine import ShardSpec
    from oracle import mppi_oracle as mo
    from tests import helpers as H
    from tests.golden import cases
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    paths = cases.load_paths()
    kw = cases.run_py_kwargs(cases.ref_path_for(paths, "xydq_circle.txt"), K, T, param_lambda=3000.0)
    c = mo.OracleMPPI(**kw)
    eps = mo.injected_noise(seed, K, T, kw["sigma"]).astype(np.float64)
    k0, n = ShardSpec(rank, world).bounds(K)
    p = mo.update_waypoint(c, cases.X0[0], cases.X0[1])
    rho, eta, V = H.oracle_partial(c, cases.X0, eps, k0, k0 + n, p)
    part = torch.from_numpy(np.concatenate([[rho, eta], V.reshape(-1)]))          # [2 + 2T] per environment
    gathered = torch.zeros(world * part.numel(), dtype=torch.float64)
    dist.all_gather_into_tensor(gathered, part)
    g = gathered.reshape(world, -1).numpy()
    parts = [(g[r, 0], g[r, 1], g[r, 2:].reshape(T, 2)) for r in range(world)]
    rho_c, eta_c, w_eps = H.combine_partials(parts, c.param_lambda)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([[rho_c, eta_c], w_eps.reshape(-1)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,K", [(2, 200), (3, 101)])
def test_sharded_combine_equals_single_process(tmp_path, world, K):
    sys.path.insert(0, ROOT)
    from oracle import mppi_oracle as mo
    from tests.golden import cases
    T, seed = 12, 77
    mp.spawn(_worker, args=(world, _free_port(), K, T, seed, str(tmp_path)), nprocs=world, join=True)
    paths = cases.load_paths()
    kw = cases.run_py_kwargs(cases.ref_path_for(paths, "xydq_circle.txt"), K, T, param_lambda=3000.0)
    c = mo.OracleMPPI(**kw)
    o = mo.step_vectorized(c, cases.X0, mo.injected_noise(seed, K, T, kw["sigma"]).astype(np.float64))
    outs = [np.load(tmp_path / f"r{r}.npy") for r in range(world)]
    for r in range(1, world):
        np.testing.assert_array_equal(outs[0], outs[r])           # every rank ends with identical bits
    assert outs[0][0] == o["rho"]
    np.testing.assert_allclose(outs[0][1], o["eta"], rtol=1e-12)
    np.testing.assert_allclose(outs[0][2:].reshape(T, 2), o["w_eps_raw"], rtol=1e-10, atol=1e-12)

"""Multi-GPU check launched by torchrun (one rank per GPU): the sharded MPPI step over NCCL, eager and
CUDA-graph replayed, must reproduce the single-GPU step (run on rank 0 with the whole sample set).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_dist_gpu.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mppi_robotarm_b200 import MPPIControllerForPathTracking   # noqa: E402
from tests.golden import cases                                  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    paths = cases.load_paths()
    ref = cases.ref_path_for(paths, "xydq_circle.txt")
    K, T = 50001, 40                                            # not divisible by the world size
    kw = cases.run_py_kwargs(ref, K, T, param_lambda=3000.0)
    results = {}
    for label, graph, exch in (("eager", False, "nccl"), ("graph", True, "nccl"), ("p2p", True, "p2p"),
                               ("p2p_eager", False, "p2p"), ("auto", True, "auto")):
        c = MPPIControllerForPathTracking(**kw, seed=77, verbose=False, distributed=True, use_graph=graph, exchange=exch)
        x = np.array(cases.X0)
        seq = []
        for _ in range(4):
            u0, useq, opt, _ = c.calc_control_input(x)
            seq.append((u0.copy(), useq.copy(), opt.copy(), c.prev_waypoints_idx))
        results[label] = seq
        if label == "auto":
            assert c._engine().exchange in ("p2p", "nccl")
            if rank == 0:
                print(f"exchange='auto' resolved to {c._engine().exchange!r}")
        # every rank must hold bit-identical controller state
        mine = torch.from_numpy(c.u_prev.copy()).cuda()
        allu = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allu, mine)
        for other in allu:
            assert torch.equal(other, allu[0]), "ranks diverged"
        c.close()
    for other in ("graph", "p2p", "p2p_eager", "auto"):   # same partials, same combine: bit-identical
        for a, b in zip(results["eager"], results[other]):
            for xa, xb in zip(a[:3], b[:3]):
                np.testing.assert_array_equal(xa, xb)
    if rank == 0:
        single = MPPIControllerForPathTracking(**kw, seed=77, verbose=False, device=torch.cuda.current_device())
        x = np.array(cases.X0)
        for s in range(4):
            u0, useq, opt, _ = single.calc_control_input(x)
            r = results["graph"][s]
            scale = np.max(np.abs(useq))
            assert np.max(np.abs(useq - r[1])) <= 5e-6 * scale, (s, np.max(np.abs(useq - r[1])))
            assert np.max(np.abs(opt - r[2])) <= 1e-5
            assert single.prev_waypoints_idx == r[3]
        single.close()
        print(f"DIST_OK world={world}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

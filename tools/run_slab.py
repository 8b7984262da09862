"""Slab-decomposed C5 solve under torchrun (one process per GPU), one JSON line from rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/run_slab.py [--n3d 256] [--reps 1]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import econpdes_b200 as E  # noqa: E402
from econpdes_b200 import examples as EX  # noqa: E402
from econpdes_b200 import models as PM  # noqa: E402


def main():
    opt = {a.split("=")[0]: a.split("=")[1] for a in sys.argv[1:] if a.startswith("--") and "=" in a}
    n = int(opt.get("--n3d", 256))
    reps = int(opt.get("--reps", 1))
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    grp = E.SlabGroup.from_torch(dist, local)
    g, gu = EX.hjb3d_grid(n)
    first, count = E.slab_range(n, rank, world)
    guess = {"v": np.zeros((n, n, count))}
    out = None
    for _ in range(reps + 1):           # first solve is the warm-up (graph capture, allocations)
        dist.barrier()
        t = time.time()
        res = E.pdesolve_slab(PM.Hjb3dSynth(), g, guess, grp)
        wall = time.time() - t
        ms = torch.tensor([res.device_ms, wall * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out = (res, ms.tolist())
    res, (dev_ms, wall_ms) = out
    lo = torch.tensor([float(res.solution["v"].min())], dtype=torch.float64, device="cuda")
    hi = torch.tensor([float(res.solution["v"].max())], dtype=torch.float64, device="cuda")
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": f"C5 hjb3d {n}^3 slab x{world}", "n_gpus": world, "converged": res.converged,
                          "residual_norm": res.residual_norm, "n_outer": res.n_outer, "n_newton": res.n_newton,
                          "n_linear_iters": res.n_linear_iters, "device_ms_max": dev_ms, "wall_ms_max": wall_ms,
                          "newton_solves_per_sec": res.n_newton / (dev_ms * 1e-3),
                          "kernel_launches_rank0": res.kernel_launches, "nccl_calls_rank0": grp.collectives,
                          "p2p_halo_exchanges_rank0": grp.p2p_exchanges,
                          "log": [(r["iter"], "%.2e" % r["delta"], "%.2e" % r["residual"], r["status"],
                                   r["newton_iters"], r["linear_iters"]) for r in res.log],
                          "v_range": [lo.item(), hi.item()]}), flush=True)
    grp.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""world_size-2 gloo test of the multi-rank host logic of the batched sweep (config C4):
balanced / weak slicing, the max-over-ranks and sum-over-ranks reductions bench.py reports,
and that two ranks solving their slices reproduce the single-process sweep.  The per-rank
"solve" here is the plain-C oracle (test infrastructure): the GPU kernel is covered by the
-m gpu tests; what is exercised is the partition + reduction + gather path."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import econpdes_b200  # noqa: F401
    from econpdes_b200 import examples as EX
    from econpdes_b200.sharding import Reducer, shard_range
    from oracle import c_oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    red = Reducer(dist)
    first, last = shard_range(total, rank, world)
    params, grids, guess = EX.campbell_cochrane_sweep(first, last - first)
    sol, rn, no, nn = c_oracle.solve_batch("campbell_cochrane", params, grids, guess)
    newton_total = red.sum(float(nn.sum()))
    t_max = red.max(1.0 + rank)            # stand-in for a per-rank device time
    all_sol = red.gather_concat(np.pad(sol, ((0, (total + world - 1) // world - sol.shape[0]), (0, 0))))
    if rank == 0:
        q.put((first, last, newton_total, t_max, all_sol))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_and_balance():
    sys.path.insert(0, ROOT)
    import econpdes_b200  # noqa: F401
    from econpdes_b200.sharding import shard_range, weak_slice
    for total, world in ((16384, 8), (10, 3), (7, 8), (1, 2)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert weak_slice(16384, 3) == (49152, 65536)
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)


def test_two_rank_sweep_matches_single_process():
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import econpdes_b200  # noqa: F401
    from econpdes_b200 import examples as EX
    from oracle import c_oracle
    total, world = 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    first, last, newton_total, t_max, all_sol = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    params, grids, guess = EX.campbell_cochrane_sweep(0, total)
    sol, rn, no, nn = c_oracle.solve_batch("campbell_cochrane", params, grids, guess)
    assert (first, last) == (0, 6)
    assert newton_total == float(nn.sum())
    assert t_max == 2.0
    # rank 0 holds 6 problems, rank 1 holds 5 (+1 row of padding)
    np.testing.assert_array_equal(all_sol[:6], sol[:6])
    np.testing.assert_array_equal(all_sol[6:11], sol[6:11])


def _slab_worker(rank, world, port, S_last, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import econpdes_b200 as E
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    # the byte channel SlabGroup.from_torch uses for the NCCL unique id (the id itself needs a GPU)
    box = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    first, count = E.slab_range(S_last, rank, world)
    spans = [None] * world
    dist.all_gather_object(spans, (first, count))
    if rank == 0:
        q.put((box[0], spans))
    dist.barrier()
    dist.destroy_process_group()


def test_slab_partition_and_id_broadcast_two_ranks():
    """Host side of the slab-decomposed solve (config C5): every rank derives its plane range from
    the C ABI's econpdes_slab_range and receives rank 0's 128-byte communicator id."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, 2, port, 257, q)) for r in range(2)]
    for p in procs:
        p.start()
    uid, spans = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert uid == bytes(range(128))
    assert spans == [(0, 128), (128, 129)]


def test_slab_ranges_even_aligned_cover():
    sys.path.insert(0, ROOT)
    import econpdes_b200 as E
    for S, world in ((256, 8), (256, 2), (33, 3), (64, 4), (17, 8)):
        spans = [E.slab_range(S, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == S
        for (f, c), (f2, _) in zip(spans, spans[1:]):
            assert f + c == f2 and f % 2 == 0 and c >= 2 and c % 2 == 0
    with pytest.raises(E.EconPDEsCudaError):
        E.slab_range(6, 0, 4)

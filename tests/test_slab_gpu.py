"""Slab-decomposed solve (config C5) on 2 GPUs against the single-GPU solve of the
same problem: value function <= 1e-8 relative, final residual norm within 1e-10, same converged
flag, same accept/reject sequence and Newton iteration counts (BASELINE.json north_star).

Needs two CUDA devices (`gpurun --gpus 2 -- python -m pytest tests/test_slab_gpu.py -m gpu`);
skipped on a one-GPU box.  Each rank is its own process (one process per GPU); the NCCL unique
id travels through a multiprocessing queue, so the test does not depend on torch.
"""
import os
import sys

import numpy as np
import pytest

import econpdes_b200 as E
from econpdes_b200 import _lib as L
from econpdes_b200 import examples as EX
from econpdes_b200 import models as PM

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem(name):
    if name in ("golden24", "golden24_bound"):   # the oracle fixture's own grid and guess
        d = np.load(os.path.join(ROOT, "tests", "golden", "c5_hjb3d_24.npz"))
        return (PM.Hjb3dSynth(*d["params"]), {"x": d["grid_x"], "y": d["grid_y"], "z": d["grid_z"]}, {"v": d["guess_v"]})
    if name == "hjb3d":
        g, gu = EX.hjb3d_grid(40)
        return PM.Hjb3dSynth(), g, gu
    if name == "hjb3d_odd":      # odd size: last rank owns one more plane
        g, gu = EX.hjb3d_grid(33)
        return PM.Hjb3dSynth(), g, gu
    g, gu = EX.bansal_yaron_grid(96, 64)
    return PM.BansalYaronModel(), g, gu


def _worker(rank, world, idq, outq, name, replicate_below):
    sys.path.insert(0, ROOT)
    if replicate_below:
        os.environ["ECONPDES_REPLICATE_BELOW"] = str(replicate_below)
    import econpdes_b200 as E2
    from econpdes_b200 import _lib as L2
    L2.check(L2.load().econpdes_set_device(rank))
    uid = idq.get(timeout=120)
    grp = E2.SlabGroup(rank, world, uid)
    model, g, gu = _problem(name)
    kw = {}
    if name == "golden24_bound":
        # a (never binding) lower bound that is finite on the LAST rank's planes only: every rank must still take the
        # MCP residual and the masked norm, whose collectives all ranks have to call (the flag was local)
        lb = np.full(gu["v"].shape, -np.inf)
        lb[..., -3:] = -1e6
        kw["lower_bound"] = lb
    res = E2.pdesolve_slab(model, g, gu, grp, **kw)
    key = next(iter(res.solution))
    outq.put((rank, res.slab, res.solution[key], res.residual_norm, res.converged,
              [(r["status"], r["newton_iters"]) for r in res.log], grp.collectives))
    grp.close()


# replicate_below: multigrid levels with at most that many points run replicated on every rank;
# 0 keeps the default (the test grids then have ONE distributed level), 1500 forces several
# distributed levels so the distributed transfers and Galerkin passes are exercised too
@pytest.mark.parametrize("name,replicate_below", [("hjb3d", 0), ("hjb3d", 1500), ("hjb3d_odd", 0),
                                                  ("hjb3d_odd", 1500), ("bansal_yaron", 0), ("bansal_yaron", 200)])
def test_two_gpu_slab_matches_single_gpu(name, replicate_below):
    if L.load().econpdes_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    world = 2
    idqs = [ctx.Queue() for _ in range(world)]
    outq = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, idqs[r], outq, name, replicate_below)) for r in range(world)]
    for p in procs:
        p.start()
    uid = E.slab.new_unique_id()
    for q in idqs:
        q.put(uid)
    outs = sorted([outq.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    model, g, gu = _problem(name)
    ref = E.pdesolve(model, g, gu, verbose=False)
    key = next(iter(ref.solution))
    full = np.concatenate([o[2] for o in outs], axis=-1)
    assert [o[1] for o in outs] == [E.slab_range(full.shape[-1], r, world) for r in range(world)]
    np.testing.assert_allclose(full, ref.solution[key], rtol=1e-8)
    for o in outs:
        assert abs(o[3] - ref.residual_norm) <= 1e-10          # identical scalars on every rank
        assert o[3] == outs[0][3]
        assert o[4] == ref.converged
        assert o[5] == [(r["status"], r["newton_iters"]) for r in ref.log]
        assert o[6] > 0                                          # NCCL traffic actually happened


@pytest.mark.parametrize("name", ["golden24", "golden24_bound"])
def test_two_gpu_slab_matches_the_oracle_fixture(name):
    """The 2-GPU slab solve against the ORACLE fixture c5_hjb3d_24.npz (not only against the 1-GPU solve): values
    1e-8 relative, residual norm 1e-10, same accept sequence and Newton counts.  `golden24_bound` adds a lower bound
    that is finite on one rank's slab only and never binds: same solution, and the ranks must not dead-lock."""
    if L.load().econpdes_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    world = 2
    idqs = [ctx.Queue() for _ in range(world)]
    outq = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, idqs[r], outq, name, 0)) for r in range(world)]
    for p in procs:
        p.start()
    uid = E.slab.new_unique_id()
    for q in idqs:
        q.put(uid)
    outs = sorted([outq.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    d = np.load(os.path.join(ROOT, "tests", "golden", "c5_hjb3d_24.npz"))
    full = np.concatenate([o[2] for o in outs], axis=-1)
    np.testing.assert_allclose(full, d["sol_v"], rtol=1e-8)
    for o in outs:
        assert o[4] == bool(d["converged"])
        assert [s for s, _ in o[5]] == [int(r[3]) for r in d["log"]]
        assert [n for _, n in o[5]] == [int(r[4]) for r in d["log"]]
        if name == "golden24":
            assert abs(o[3] - float(d["residual_norm"])) <= 1e-10

"""Slab-decomposed `pdesolve` of ONE large grid across the GPUs of a node (config C5).

One process per GPU.  The slowest state dimension is cut into contiguous slabs
(`econpdes_slab_range`); every rank passes its slab of the initial guess and receives its slab
of the solution, while the scalars of the result (residual norm, converged flag, iteration log)
are identical on all ranks.  Halo planes travel over NVLink (NCCL send/recv inside
libeconpdes_cuda.so); this module only does the plumbing: creating the NCCL unique id on rank 0
and handing it to the other ranks through `torch.distributed` (or any byte channel).
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Optional, Tuple

import numpy as np

from . import _lib as L
from .models import RegisteredModel
from .pdesolve import SQRT_EPS, EconPDEResult, _as_ordered, _options

UNIQUE_ID_BYTES = 128


def slab_range(size_last: int, rank: int, world: int) -> Tuple[int, int]:
    """(first, count) of the planes of the slowest dimension owned by `rank` (host-only)."""
    lib = L.load()
    first, count = C.c_int64(), C.c_int64()
    L.check(lib.econpdes_slab_range(size_last, rank, world, C.byref(first), C.byref(count)))
    return first.value, count.value


def new_unique_id() -> bytes:
    lib = L.load()
    buf = (C.c_uint8 * UNIQUE_ID_BYTES)()
    L.check(lib.econpdes_dist_unique_id(buf))
    return bytes(buf)


class SlabGroup:
    """The ranks that share one grid: wraps the library's NCCL communicator (`econpdes_dist`)."""

    def __init__(self, rank: int, world: int, unique_id: bytes = b""):
        if world == 1 and not unique_id:
            unique_id = bytes(UNIQUE_ID_BYTES)   # one slab: no NCCL communicator is created
        if len(unique_id) != UNIQUE_ID_BYTES:
            raise ValueError("unique id must be %d bytes" % UNIQUE_ID_BYTES)
        self.rank, self.world = rank, world
        lib = L.load()
        self._h = C.c_void_p()
        buf = (C.c_uint8 * UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        L.check(lib.econpdes_dist_create(rank, world, buf, C.byref(self._h)))

    @classmethod
    def from_torch(cls, dist, device_index: Optional[int] = None) -> "SlabGroup":
        """Build the group over an initialised `torch.distributed` process group (NCCL or gloo):
        rank 0 creates the id and broadcasts it; every rank binds to `device_index` (default:
        LOCAL_RANK / rank) first."""
        import os
        rank, world = dist.get_rank(), dist.get_world_size()
        dev = device_index if device_index is not None else int(os.environ.get("LOCAL_RANK", rank))
        L.check(L.load().econpdes_set_device(dev))
        box = [new_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(rank, world, box[0])

    @property
    def collectives(self) -> int:
        return int(L.load().econpdes_dist_collectives(self._h))

    @property
    def p2p_exchanges(self) -> int:
        return int(L.load().econpdes_dist_p2p_exchanges(self._h))

    def close(self):
        if self._h:
            L.load().econpdes_dist_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def pdesolve_slab(pde, grid, guess, group: SlabGroup, *, abstol=SQRT_EPS, maxiters=100, lower_bound=None,
                  upper_bound=None, **kwargs) -> EconPDEResult:
    """`pdesolve` (src/pdesolve.jl:96-154) with the grid slab-decomposed over `group`.

    `grid` is the GLOBAL state grid; `guess` holds either global arrays (sliced here) or this
    rank's slab.  The returned `solution` is this rank's slab; `result.slab = (first, count)`.
    """
    if not isinstance(pde, RegisteredModel):
        raise TypeError("the device path needs a registered model functor (econpdes_b200.models)")
    grid = _as_ordered(grid, "the state grid")
    guess = _as_ordered(guess, "the initial guess")
    statenames, ynames = tuple(grid), tuple(guess)
    if statenames != tuple(pde.statenames) or ynames != tuple(pde.solnames):
        raise ValueError(f"model `{pde.name}` expects states {pde.statenames} and unknowns {pde.solnames}")
    if len(ynames) != 1 or len(statenames) < 2:
        raise ValueError("the slab-decomposed solve takes 2-D/3-D grids with one unknown")
    grids = [np.ascontiguousarray(np.asarray(g, dtype=np.float64)) for g in grid.values()]
    S = tuple(g.shape[0] for g in grids)
    first, count = slab_range(S[-1], group.rank, group.world)
    Sloc = S[:-1] + (count,)
    g0 = np.asarray(guess[ynames[0]], dtype=np.float64)
    if g0.shape == S:
        g0 = g0[..., first:first + count]
    elif g0.shape != Sloc:
        raise ValueError(f"the initial guess has size {g0.shape}; expected the grid {S} or this rank's slab {Sloc}")
    guess_vec = np.ascontiguousarray(g0.reshape(-1, order="F"))

    def bound(b, default):   # global array, this rank's slab, or a scalar -> this rank's slab, flattened
        if b is None:
            return None
        a = np.asarray(b, dtype=np.float64)
        if a.ndim == 0:
            a = np.full(Sloc, float(a))
        elif a.shape == S:
            a = a[..., first:first + count]
        elif a.shape != Sloc:
            raise ValueError(f"a bound has size {a.shape}; expected the grid {S} or this rank's slab {Sloc}")
        return np.ascontiguousarray(a.reshape(-1, order="F"))
    lb_vec, ub_vec = bound(lower_bound, -np.inf), bound(upper_bound, np.inf)
    delta = kwargs.pop("Δ", kwargs.pop("delta", 1.0))
    opts = _options(abstol, maxiters, kwargs)
    opts.delta = float(delta)
    if kwargs:
        raise TypeError(f"unknown keyword arguments {sorted(kwargs)}")

    lib = L.load()
    mid = L.check(lib.econpdes_model_id(pde.name.encode()))
    prm = pde.params()
    p = L.Problem()
    p.model_id, p.ndims, p.nfuncs, p.nparams = mid, len(S), 1, prm.shape[0]
    for d in range(len(S)):
        p.size[d] = S[d]
        p.grid[d] = L.dptr(grids[d])
    p.params = L.dptr(prm) if prm.shape[0] else None
    p.guess = L.dptr(guess_vec)
    p.lower_bound = L.dptr(lb_vec)
    p.upper_bound = L.dptr(ub_vec)
    sol = np.empty(guess_vec.shape[0])
    cap = max(int(maxiters), 1)
    log = (L.LogRow * cap)()
    r = L.Result()
    r.solution = L.dptr(sol)
    r.saved = None
    r.log = log
    r.log_capacity = cap
    L.check(lib.econpdes_solve_stationary_slab(group._h, C.byref(p), first, count, C.byref(opts), C.byref(r)))
    solution = OrderedDict([(ynames[0], sol.reshape(Sloc, order="F"))])
    rows = tuple(dict(iter=log[i].iter, status=log[i].status, newton_iters=log[i].newton_iters,
                      linear_iters=log[i].linear_iters, delta=log[i].delta, residual=log[i].residual)
                 for i in range(r.log_len))
    res = EconPDEResult(solution, r.residual_norm, OrderedDict(), abstol, r.exit_reason, rows, r.n_outer, r.n_newton,
                        r.n_linear_iters, r.kernel_launches, r.device_ms, 0, 0, 0.0, r.n_linear_solves,
                        r.n_linear_loose, r.lin_max_relres)
    res.slab = (first, count)
    return res

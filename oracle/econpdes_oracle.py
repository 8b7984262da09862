"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

A numpy/scipy restatement of the finite-scheme hot path of EconPDEs.jl v3.0.0
(reference tree mounted at /root/reference while this file was written; nothing
is read from it at run time).  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

What is restated, and from where (file:line relative to the reference root):

* grid spacing, upwind first differences, second differences with ghost-node
  boundary conditions, the three cross-derivative stencils
      -> src/utils/derivatives.jl:39-159                      (``differentiate``)
* the residual sweep over the grid and the (S..., F) column-major layout
      -> src/pdesolve.jl:111, 436-462                          (``pde_residual``)
* the boundary-condition array, including its one-slot-per-point quirk
      -> src/pdesolve.jl:312-335                               (``bc_array``)
* Jacobian sparsity pattern and closed-form distance-2 coloring
      -> src/utils/coloring.jl:18-35, 52-70     (``local_stencil_jacobian``,
                                                  ``stencil_colors``)
* implicit-Euler residual, MCP min-max residual, Newton seam
      -> src/finiteschemesolve.jl:125-147, 153-230   (``implicit_timestep``)
* SER time-step controller and the pseudo-transient loop
      -> src/finiteschemesolve.jl:236-275, 343-465   (``finiteschemesolve``)
* result type                -> src/EconPDEs.jl:36-48           (``EconPDEResult``)

Third-party pieces the reference calls but does not vendor (Project.toml:20-27,
Manifest.toml git-ignored, so versions are unpinned).  Their published algorithms
are restated here:

* FiniteDiff.jl 2.x ``finite_difference_jacobian!`` (forward, colorvec+sparsity):
  f(x) once; per color c:  eps_c = sqrt(eps)*max(1, sqrt(||x1 .* mask_c||_2));
  x1 += eps_c*mask_c ; f(x1) ; J[:, cols of c] = (f(x1)-f(x))/eps_c ;
  x1 -= eps_c*mask_c  (restoration by subtraction, so x1 drifts by <= 1 ulp).
* NLsolve.jl 4.x ``newton_`` with the default ``Static`` line search: undamped
  full steps, stop on ||F||_inf <= ftol, x == xold, iteration cap, or NaN.
* Julia ``\\`` on SparseMatrixCSC (UMFPACK LU): here scipy's SuperLU.

PARITY PINNING: the oracle is pinned end to end by the two iteration traces the
reference publishes (README.md:72-82, docs/src/solver.md:66-73), by the closed-form
answers in test/runtests.jl (:228-232, :276-281, :314-319, :386-394, :409-416,
:428-431, :447-450) and the structural facts (:500-502, :511-524, :305-312), and by
the "11 outer iterations" of the 1000-point growth model (examples/details.tex:173).
For Bansal-Yaron, consumption-saving and 3-D problems the reference ships no numeric
output: parity there is *unpinned* beyond those checks.
"""
from __future__ import annotations

import itertools
import math
import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

SQRT_EPS = math.sqrt(np.finfo(np.float64).eps)  # 1.4901161193847656e-8
EPS = float(np.finfo(np.float64).eps)


# --------------------------------------------------------------------------------------
# Grid spacing and stencils (src/utils/derivatives.jl:39-159), vectorised over the grid.
# --------------------------------------------------------------------------------------
def spacings(x: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """dm[i] = x[max(i,2)]-x[max(i-1,1)], dp[i] = x[min(i+1,S)]-x[min(i,S-1)],
    d = (dm+dp)/2   (1-based in the reference, derivatives.jl:47-49)."""
    x = np.asarray(x, dtype=np.float64)
    S = x.shape[0]
    i = np.arange(S)
    dm = x[np.maximum(i, 1)] - x[np.maximum(i - 1, 0)]
    dp = x[np.minimum(i + 1, S - 1)] - x[np.minimum(i, S - 2)]
    return dm, dp, (dm + dp) / 2


def _bshape(v: np.ndarray, d: int, N: int) -> np.ndarray:
    shp = [1] * N
    shp[d] = v.shape[0]
    return v.reshape(shp)


def _take_clamped(a: np.ndarray, shifts: Dict[int, int]) -> np.ndarray:
    """a[..., clamp(i_d + shift_d), ...]  (derivatives.jl:139-147)."""
    out = a
    for d, s in shifts.items():
        S = a.shape[d]
        idx = np.clip(np.arange(S) + s, 0, S - 1)
        out = np.take(out, idx, axis=d)
    return out


def differentiate(solnames: Sequence[str], statenames: Sequence[str],
                  grids: Sequence[np.ndarray], y: np.ndarray, bc: np.ndarray) -> Dict[str, np.ndarray]:
    """All finite-difference fields for every unknown at every grid point.

    ``y`` and ``bc`` have shape (S_1..S_N, F).  Field names follow
    docs/src/pde_function.md:53-62 (``v``, ``vx_up``, ``vx_down``, ``vxx``, ``vxy``,
    ``vxy_up``, ``vxy_down``).  Expression order mirrors derivatives.jl line by line.
    """
    N = len(grids)
    S = y.shape[:N]
    u: Dict[str, np.ndarray] = {}
    sp_ = [spacings(g) for g in grids]
    dm = [_bshape(s[0], d, N) for d, s in enumerate(sp_)]
    dp = [_bshape(s[1], d, N) for d, s in enumerate(sp_)]
    dd = [_bshape(s[2], d, N) for d, s in enumerate(sp_)]
    for k, name in enumerate(solnames):
        yk = y[..., k]
        bck = bc[..., k]
        u[name] = yk
        for d in range(N):
            sn = statenames[d]
            fwd = _take_clamped(yk, {d: +1})
            bwd = _take_clamped(yk, {d: -1})
            # bc[i1.., (1|end) at dim d, ..iN, k] keeping the *current* other indices
            bc_lo = np.take(bck, [0], axis=d)
            bc_hi = np.take(bck, [S[d] - 1], axis=d)
            idx = _bshape(np.arange(S[d]), d, N)
            with np.errstate(all="ignore"):
                up = np.where(idx < S[d] - 1, (fwd - yk) / dp[d], np.broadcast_to(bc_hi, yk.shape))
                down = np.where(idx > 0, (yk - bwd) / dm[d], np.broadcast_to(bc_lo, yk.shape))
            u[f"{name}{sn}_up"] = up
            u[f"{name}{sn}_down"] = down
        for d in range(N):
            sn = statenames[d]
            fwd = _take_clamped(yk, {d: +1})
            bwd = _take_clamped(yk, {d: -1})
            bc_lo = np.take(bck, [0], axis=d)
            bc_hi = np.take(bck, [S[d] - 1], axis=d)
            idx = _bshape(np.arange(S[d]), d, N)
            y_at_1 = np.take(yk, [0], axis=d)
            y_at_2 = np.take(yk, [1], axis=d)
            y_at_end = np.take(yk, [S[d] - 1], axis=d)
            y_at_endm1 = np.take(yk, [S[d] - 2], axis=d)
            interior = fwd / (dp[d] * dd[d]) + bwd / (dm[d] * dd[d]) - 2 * yk / (dp[d] * dm[d])
            left = (y_at_2 / (dp[d] * dd[d]) + (y_at_1 - bc_lo * dm[d]) / (dm[d] * dd[d])
                    - 2 * y_at_1 / (dp[d] * dm[d]))
            right = ((y_at_end + bc_hi * dp[d]) / (dp[d] * dd[d]) + y_at_endm1 / (dm[d] * dd[d])
                     - 2 * y_at_end / (dp[d] * dm[d]))
            u[f"{name}{sn}{sn}"] = np.where((idx > 0) & (idx < S[d] - 1), interior,
                                            np.where(idx == 0, left, right))
        for d1 in range(N):
            for d2 in range(d1 + 1, N):
                sn1, sn2 = statenames[d1], statenames[d2]
                y_00 = yk
                y_p0 = _take_clamped(yk, {d1: +1})
                y_m0 = _take_clamped(yk, {d1: -1})
                y_0p = _take_clamped(yk, {d2: +1})
                y_0m = _take_clamped(yk, {d2: -1})
                y_pp = _take_clamped(yk, {d1: +1, d2: +1})
                y_pm = _take_clamped(yk, {d1: +1, d2: -1})
                y_mp = _take_clamped(yk, {d1: -1, d2: +1})
                y_mm = _take_clamped(yk, {d1: -1, d2: -1})
                u[f"{name}{sn1}{sn2}"] = (y_pp - y_pm - y_mp + y_mm) / (4 * dd[d1] * dd[d2])
                u[f"{name}{sn1}{sn2}_up"] = ((y_pp - y_p0 - y_0p + y_00) / (2 * dp[d1] * dp[d2])
                                            + (y_00 - y_m0 - y_0m + y_mm) / (2 * dm[d1] * dm[d2]))
                u[f"{name}{sn1}{sn2}_down"] = (-(y_pm - y_p0 - y_0m + y_00) / (2 * dp[d1] * dm[d2])
                                              - (y_00 - y_m0 - y_0p + y_mp) / (2 * dm[d1] * dp[d2]))
    return u


def differentiate_point(solnames, statenames, grids, y, icar, bc) -> Dict[str, float]:
    """Scalar, literal transcription of the generated function body
    (derivatives.jl:34-167) for ONE grid point ``icar`` (0-based tuple).  Slow; used by
    the tests to cross-check the vectorised version on small grids."""
    N = len(grids)
    S = y.shape[:N]
    dm, dp, dd = [], [], []
    for d in range(N):
        g, i = grids[d], icar[d]
        m = g[max(i, 1)] - g[max(i - 1, 0)]
        p = g[min(i + 1, S[d] - 1)] - g[min(i, S[d] - 2)]
        dm.append(m); dp.append(p); dd.append((m + p) / 2)

    def yref(k, subs=None):
        idx = list(icar)
        for d, v in (subs or {}).items():
            idx[d] = v
        return float(y[tuple(idx) + (k,)])

    def bcref(k, d, hi):
        idx = list(icar)
        idx[d] = S[d] - 1 if hi else 0
        return float(bc[tuple(idx) + (k,)])

    out: Dict[str, float] = {}
    for k, name in enumerate(solnames):
        out[name] = yref(k)
        for d in range(N):
            i, sn = icar[d], statenames[d]
            out[f"{name}{sn}_up"] = ((yref(k, {d: i + 1}) - yref(k)) / dp[d]) if i < S[d] - 1 else bcref(k, d, True)
            out[f"{name}{sn}_down"] = ((yref(k) - yref(k, {d: i - 1})) / dm[d]) if i > 0 else bcref(k, d, False)
        for d in range(N):
            i, sn = icar[d], statenames[d]
            if 0 < i < S[d] - 1:
                val = (yref(k, {d: i + 1}) / (dp[d] * dd[d]) + yref(k, {d: i - 1}) / (dm[d] * dd[d])
                       - 2 * yref(k) / (dp[d] * dm[d]))
            elif i == 0:
                val = (yref(k, {d: 1}) / (dp[d] * dd[d])
                       + (yref(k, {d: 0}) - bcref(k, d, False) * dm[d]) / (dm[d] * dd[d])
                       - 2 * yref(k, {d: 0}) / (dp[d] * dm[d]))
            else:
                val = ((yref(k, {d: S[d] - 1}) + bcref(k, d, True) * dp[d]) / (dp[d] * dd[d])
                       + yref(k, {d: S[d] - 2}) / (dm[d] * dd[d])
                       - 2 * yref(k, {d: S[d] - 1}) / (dp[d] * dm[d]))
            out[f"{name}{sn}{sn}"] = val
        for d1 in range(N):
            for d2 in range(d1 + 1, N):
                i1, i2 = icar[d1], icar[d2]
                p1, m1 = min(i1 + 1, S[d1] - 1), max(i1 - 1, 0)
                p2, m2 = min(i2 + 1, S[d2] - 1), max(i2 - 1, 0)
                y00 = yref(k)
                yp0, ym0 = yref(k, {d1: p1}), yref(k, {d1: m1})
                y0p, y0m = yref(k, {d2: p2}), yref(k, {d2: m2})
                ypp, ypm = yref(k, {d1: p1, d2: p2}), yref(k, {d1: p1, d2: m2})
                ymp, ymm = yref(k, {d1: m1, d2: p2}), yref(k, {d1: m1, d2: m2})
                sn = f"{name}{statenames[d1]}{statenames[d2]}"
                out[sn] = (ypp - ypm - ymp + ymm) / (4 * dd[d1] * dd[d2])
                out[sn + "_up"] = ((ypp - yp0 - y0p + y00) / (2 * dp[d1] * dp[d2])
                                   + (y00 - ym0 - y0m + ymm) / (2 * dm[d1] * dm[d2]))
                out[sn + "_down"] = (-(ypm - yp0 - y0m + y00) / (2 * dp[d1] * dm[d2])
                                     - (y00 - ym0 - y0p + ymp) / (2 * dm[d1] * dp[d2]))
    return out


# --------------------------------------------------------------------------------------
# Sparsity pattern + coloring (src/utils/coloring.jl:18-35, 52-70).  1-based like Julia.
# --------------------------------------------------------------------------------------
def local_stencil_jacobian(S: Sequence[int], F: int) -> Tuple[np.ndarray, np.ndarray]:
    """CSC pattern (colptr, rowval), both 1-based int64, of the clipped +-1 box coupling
    all F x F unknowns; rows inside a column are sorted (unknown-major, then
    column-major grid point), as ``sparse(rows, cols, ...)`` produces."""
    S = tuple(int(s) for s in S)
    N = len(S)
    l = int(np.prod(S))
    strides = np.cumprod((1,) + S[:-1]).astype(np.int64)
    coords = np.stack(np.unravel_index(np.arange(l), S, order="F"), axis=1).astype(np.int64)  # (l, N)
    # neighbours of every point, offsets enumerated so that the linear index increases
    offs = np.array(list(itertools.product(*[(-1, 0, 1)] * N)), dtype=np.int64)[:, ::-1]  # dim 1 fastest
    nb = coords[:, None, :] + offs[None, :, :]                       # (l, K, N)
    valid = np.all((nb >= 0) & (nb < np.array(S)[None, None, :]), axis=2)
    nblin = (nb * strides[None, None, :]).sum(axis=2)                # (l, K) 0-based
    cnt = valid.sum(axis=1)                                          # neighbours per point
    # column (c, fin): rows = for fout: sorted nblin[c, valid] + fout*l
    percol = cnt * F
    colptr = np.empty(l * F + 1, dtype=np.int64)
    colptr[0] = 1
    colptr[1:] = 1 + np.cumsum(np.tile(percol, F))
    nnz = int(colptr[-1] - 1)
    rowval = np.empty(nnz, dtype=np.int64)
    rows_pt = nblin[valid]                                           # concatenated per point, ascending
    ptstart = np.concatenate([[0], np.cumsum(cnt)])
    pos = 0
    # build for fin = 0 then copy for the other fin (same rows)
    block = np.empty(int(percol.sum()), dtype=np.int64)
    # position of each (point, fout, neighbour)
    # vectorised: for each point, its rows repeated for fout with +fout*l
    rep = np.repeat(np.arange(l), cnt)
    within = np.arange(rows_pt.shape[0]) - ptstart[rep]
    colstart = np.concatenate([[0], np.cumsum(percol)])[:-1]
    for fout in range(F):
        block[colstart[rep] + fout * cnt[rep] + within] = rows_pt + fout * l
    for fin in range(F):
        rowval[pos:pos + block.shape[0]] = block + 1
        pos += block.shape[0]
    return colptr, rowval


def stencil_colors(S: Sequence[int], F: int) -> np.ndarray:
    """1-based colors, coloring.jl:52-70."""
    S = tuple(int(s) for s in S)
    l = int(np.prod(S))
    coords = np.unravel_index(np.arange(l), S, order="F")
    base = np.zeros(l, dtype=np.int64)
    stride = 1
    for d, sd in enumerate(S):
        base += (coords[d] % 3) * stride
        stride *= min(3, sd)
    ncg = stride
    colors = np.empty(l * F, dtype=np.int64)
    for f in range(F):
        colors[f * l:(f + 1) * l] = base + 1 + f * ncg
    return colors


# --------------------------------------------------------------------------------------
# Problem container + residual sweep (src/pdesolve.jl:96-154, 312-335, 436-462)
# --------------------------------------------------------------------------------------
def bc_array(bc: Optional[dict], solnames, statenames, S) -> np.ndarray:
    """src/pdesolve.jl:312-335 (slices written dimension by dimension; later dims win)."""
    out = np.zeros(tuple(S) + (len(solnames),))
    if bc is None:
        return out
    valid = [f"{yn}{s}" for yn in solnames for s in statenames]
    for key in bc:
        if key not in valid:
            raise ValueError(f"unknown `bc` entry `{key}`: valid entries are {valid}")
    for k, yn in enumerate(solnames):
        for d, s in enumerate(statenames):
            key = f"{yn}{s}"
            if key not in bc:
                continue
            lo, hi = bc[key]
            sl = [slice(None)] * len(S)
            sl[d] = 0
            out[tuple(sl) + (k,)] = lo
            sl[d] = S[d] - 1
            out[tuple(sl) + (k,)] = hi
    return out


@dataclass
class Problem:
    """Everything ``pde!`` closes over (src/pdesolve.jl:141): model, grid, names, bc."""
    pde: Callable  # pde(state: dict, u: dict) -> (dict vt) or (dict vt, dict saved); vectorised
    statenames: Tuple[str, ...]
    grids: Tuple[np.ndarray, ...]
    solnames: Tuple[str, ...]
    bc: np.ndarray = None

    def __post_init__(self):
        self.grids = tuple(np.asarray(g, dtype=np.float64) for g in self.grids)
        self.S = tuple(g.shape[0] for g in self.grids)
        self.N = len(self.S)
        self.F = len(self.solnames)
        self.l = int(np.prod(self.S))
        self.n = self.l * self.F
        if self.bc is None:
            self.bc = np.zeros(self.S + (self.F,))
        self._state = {sn: np.broadcast_to(_bshape(g, d, self.N), self.S)
                       for d, (sn, g) in enumerate(zip(self.statenames, self.grids))}
        self.nsweeps = 0

    def _call(self, yvec: np.ndarray):
        y = np.reshape(yvec, self.S + (self.F,), order="F")
        u = differentiate(self.solnames, self.statenames, self.grids, y, self.bc)
        self.nsweeps += 1
        with np.errstate(all="ignore"):
            out = self.pde(self._state, u)
        if isinstance(out, tuple):
            return out[0], out[1]
        return out, None

    def G(self, yvec: np.ndarray) -> np.ndarray:
        """``pde!`` (src/pdesolve.jl:436-462): ydot[i,k] = pde(state_i, u_i).<k>t."""
        res, _ = self._call(yvec)
        ydot = np.empty(self.S + (self.F,))
        for k, name in enumerate(self.solnames):
            ydot[..., k] = res[name + "t"]
        return ydot.reshape(-1, order="F")

    def saved(self, yvec: np.ndarray) -> Dict[str, np.ndarray]:
        """``_fill_saved!`` (src/pdesolve.jl:424-433)."""
        _, sv = self._call(yvec)
        if sv is None:
            return {}
        return {k: np.array(np.broadcast_to(v, self.S), dtype=np.float64) for k, v in sv.items()}


# --------------------------------------------------------------------------------------
# FiniteDiff.jl colored forward-difference Jacobian (restated; see module docstring)
# --------------------------------------------------------------------------------------
class JacobianCache:
    """``_sparse_fd_setup`` (src/finiteschemesolve.jl:64-74): pattern + colors, once."""

    def __init__(self, S, F, colptr=None, rowval=None, colors=None):
        if colptr is None:
            colptr, rowval = local_stencil_jacobian(S, F)
        self.colptr, self.rowval = colptr, rowval
        self.colors = stencil_colors(S, F) if colors is None else np.asarray(colors, dtype=np.int64)
        n = self.colors.shape[0]
        self.n = n
        self.ncolors = int(self.colors.max())
        self.indptr0 = (colptr - 1).astype(np.int64)
        self.indices0 = (rowval - 1).astype(np.int64)
        col_of_nz = np.repeat(np.arange(n), np.diff(self.indptr0))
        self.color_of_nz = self.colors[col_of_nz]
        self.masks = [(self.colors == c + 1) for c in range(self.ncolors)]
        self.nz_of_color = [np.nonzero(self.color_of_nz == c + 1)[0] for c in range(self.ncolors)]


def finite_difference_jacobian(f: Callable[[np.ndarray], np.ndarray], x: np.ndarray,
                               cache: JacobianCache, eps_policy: str = "finitediff_color_norm"):
    """Returns (nzval, eps_per_color).  C+1 evaluations of ``f``."""
    x1 = x.copy()
    fx = f(x)
    nzval = np.zeros(cache.indices0.shape[0])
    eps_used = np.empty(cache.ncolors)
    for c in range(cache.ncolors):
        m = cache.masks[c]
        x2 = x1 * m
        tmp = np.sqrt(np.dot(x2, x2))  # norm(x2)
        epsilon = max(SQRT_EPS * abs(math.sqrt(tmp)), SQRT_EPS)
        eps_used[c] = epsilon
        x1 = x1 + epsilon * m
        fx1 = f(x1)
        with np.errstate(all="ignore"):
            dfx = (fx1 - fx) / epsilon
        idx = cache.nz_of_color[c]
        nzval[idx] = dfx[cache.indices0[idx]]
        x1 = x1 - epsilon * m
    return nzval, eps_used


# --------------------------------------------------------------------------------------
# NLsolve.jl newton_ with Static line search (restated) + _newtonstep seam
# --------------------------------------------------------------------------------------
class SingularException(Exception):
    pass


def _lu_solve(n, cache: JacobianCache, nzval, b):
    A = sp.csc_matrix((nzval, cache.indices0, cache.indptr0), shape=(n, n))
    if not np.all(np.isfinite(nzval)):
        raise SingularException()
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("error", spla.MatrixRankWarning)
            lu = spla.splu(A)
        x = lu.solve(b)
    except (RuntimeError, spla.MatrixRankWarning):
        raise SingularException()
    if not np.all(np.isfinite(x)):
        raise SingularException()
    return x, A


@dataclass
class NewtonResult:
    zero: np.ndarray
    iterations: int
    f_converged: bool
    x_converged: bool
    residual_inf: float


# how often the Newton iteration met a singular Jacobian and took NLsolve's regularised-normal-equations retry
# (the device path takes the same direction, csrc/normaleq.cuh; tests/test_oracle.py checks that none of the five
# configs gets there, tests/test_gpu_parity.py compares the retry itself on exactly singular Jacobians)
STATS = {"singular_retries": 0}


def nlsolve_newton(R, Jfun, cache: JacobianCache, x0, ftol, iterations) -> NewtonResult:
    """NLsolve 4.x newton_ (Static line search).  One Jacobian per iteration; the
    Jacobian at the initial point is only built if an iteration is taken (the published
    code builds it eagerly -- identical results)."""
    n = x0.shape[0]
    x = x0.copy()
    fx = R(x)
    if not np.all(np.isfinite(fx)):
        raise FloatingPointError("IsFiniteException: initial residual not finite")
    it = 0
    f_conv = float(np.max(np.abs(fx))) <= ftol
    x_conv = False
    stopped = bool(np.any(np.isnan(x)) or np.any(np.isnan(fx)))
    while (not stopped) and (not (x_conv or f_conv)) and it < iterations:
        it += 1
        nzval = Jfun(x)
        try:
            p, A = _lu_solve(n, cache, nzval, fx)
            p = -p
        except SingularException:
            # NLsolve's fallback: regularised normal equations (newton.jl, Nocedal&Wright p.289)
            STATS["singular_retries"] += 1
            A = sp.csc_matrix((np.nan_to_num(nzval), cache.indices0, cache.indptr0), shape=(n, n))
            if not np.all(np.isfinite(nzval)):
                raise
            fjac2 = (A.T @ A).tocsc()
            lam = 1e6 * math.sqrt(n * EPS) * abs(fjac2).sum(axis=0).max()
            M = -(fjac2 + lam * sp.identity(n, format="csc"))
            try:
                p = spla.splu(M.tocsc()).solve(A.T @ fx)
            except RuntimeError:
                raise SingularException()
        xold = x
        x = xold + p
        fx = R(x)
        x_conv = bool(np.max(np.abs(x - xold)) <= 0.0) if not np.any(np.isnan(xold)) else False
        with np.errstate(invalid="ignore"):
            f_conv = bool(np.max(np.abs(fx)) <= ftol)
        stopped = bool(np.any(np.isnan(x)) or np.any(np.isnan(fx)))
    with np.errstate(invalid="ignore"):
        rinf = float(np.max(np.abs(fx)))
    return NewtonResult(x, it, f_conv, x_conv, rinf)


@dataclass
class StepInfo:
    newton_iters: int = 0
    singular: bool = False


def implicit_timestep(G, ypost, delta, cache: JacobianCache, *, is_algebraic=None,
                      lower_bound=None, upper_bound=None, abstol=SQRT_EPS, maxiters=100,
                      eps_policy="finitediff_color_norm", info: Optional[StepInfo] = None):
    """src/finiteschemesolve.jl:153-230 + _newtonstep (:125-147).
    ``delta`` is a scalar or a per-row vector (signed; +-Inf = undamped row)."""
    n = ypost.shape[0]
    notalg = np.ones(n) if is_algebraic is None else (~np.asarray(is_algebraic, dtype=bool)).astype(np.float64)
    has_bounds = (lower_bound is not None and np.any(lower_bound != -np.inf)) or \
                 (upper_bound is not None and np.any(upper_bound != np.inf))

    def G_helper(y):
        ydot = G(y)
        with np.errstate(all="ignore"):
            ydot = ydot - notalg * (ypost - y) / delta           # :160
        return ydot

    R = G_helper
    if has_bounds:
        lb = np.full(n, -np.inf) if lower_bound is None else lower_bound
        ub = np.full(n, np.inf) if upper_bound is None else upper_bound

        def mcp(y):                                              # :197-203
            ydot = G_helper(y)
            return np.minimum(np.maximum(ydot, y - ub), y - lb)
        R = mcp

    def Jfun(y):                                                 # :183-188 / :216-223
        nz, _ = finite_difference_jacobian(R, y, cache, eps_policy)
        return nz

    try:
        res = nlsolve_newton(R, Jfun, cache, ypost, abstol, maxiters)
    except SingularException:
        if info is not None:
            info.singular = True
        return ypost, math.inf                                   # :142-146
    if info is not None:
        info.newton_iters = res.iterations
    final = R(res.zero)                                          # :139-141
    with np.errstate(all="ignore"):
        nrm = float(np.sqrt(np.dot(final, final)) / n)
    return res.zero, nrm


# --------------------------------------------------------------------------------------
# SER controller + pseudo-transient loop (src/finiteschemesolve.jl:236-275, 343-465)
# --------------------------------------------------------------------------------------
class SERTimeStepController:
    def __init__(self, delta, scale, max_delta):
        self.delta = float(delta)
        self.scale = float(scale)
        self.max_delta = float(max_delta)
        self.acceleration = 1.0
        self.delta_cap = float(max_delta)

    def accept_step(self, old, new):                             # :248-266
        if new <= old:
            self.acceleration *= self.scale
        else:
            self.acceleration = 1.0
        with np.errstate(all="ignore"):
            proposed = np.float64(self.delta) * self.acceleration * old / np.float64(new)
        limit = min(self.delta_cap, self.max_delta)
        if proposed > limit:
            self.delta = limit
            self.acceleration = 1.0
        else:
            self.delta = float(proposed)
        self.delta_cap = min(self.delta_cap * self.scale, self.max_delta)
        return self.delta

    def reject_step(self):                                       # :270-275
        self.acceleration = 1.0
        self.delta_cap = self.delta / 2
        self.delta /= 10
        return self.delta


# per-step statuses / exit reasons (shared vocabulary with include/econpdes_cuda.h)
ACCEPTED, REJECTED_GROWTH, REJECTED_INNER, REJECTED_NAN, REJECTED_SINGULAR = 0, 1, 2, 3, 4
EXIT_CONVERGED, EXIT_ALREADY_ZERO, EXIT_MAXITERS, EXIT_BELOW_MIN_DELTA, EXIT_INF_STEP_FAILED = 0, 1, 2, 3, 4


@dataclass
class SolveLog:
    rows: List[Tuple[int, float, float, int, int]] = field(default_factory=list)  # iter, Δ, residual|nan, status, newton
    exit_reason: int = EXIT_CONVERGED
    n_outer: int = 0
    n_newton: int = 0

    def format(self) -> str:
        lines = ["Iter TimeStep Residual", "---- -------- --------"]
        for it, d, r, st, _ in self.rows:
            if st == ACCEPTED:
                lines.append("%4d %8.2e %8.2e" % (it, d, r))
            else:
                lines.append("%4d %8.2e        ✗" % (it, d))
        return "\n".join(lines)


def residual_norm(v: np.ndarray) -> float:
    """norm(v)/length(v)  (finiteschemesolve.jl:141,371,415)."""
    with np.errstate(all="ignore"):
        return float(np.sqrt(np.dot(v, v)) / v.shape[0])


def finiteschemesolve(G, y0, cache: JacobianCache, *, is_algebraic=None, lower_bound=None,
                      upper_bound=None, abstol=SQRT_EPS, maxiters=100, delta=1.0, scale=10.0,
                      min_delta=1e-9, max_delta=math.inf, max_residual_growth=10.0,
                      inner_maxiters=10, inner_abstol=SQRT_EPS,
                      eps_policy="finitediff_color_norm") -> Tuple[np.ndarray, float, SolveLog]:
    log = SolveLog()
    n = y0.shape[0]
    dvec = np.asarray(delta, dtype=np.float64)
    if dvec.ndim == 1 and dvec.shape[0] != n:
        raise ValueError("`Δ` must be a real number or have one entry per element of `y0`")
    if np.any(np.isnan(dvec)) or np.any(dvec == 0):
        raise ValueError("`Δ` entries must be nonzero and not NaN")
    has_bounds = (lower_bound is not None and np.any(lower_bound != -np.inf)) or \
                 (upper_bound is not None and np.any(upper_bound != np.inf))
    ypost = y0.copy()
    ydot = G(ypost)
    r = residual_norm(ydot)
    if math.isnan(r):
        raise ValueError("G! returns NaN with the initial value")      # :372
    if r <= abstol:
        log.exit_reason = EXIT_ALREADY_ZERO
        return ypost, r, log
    kw = dict(is_algebraic=is_algebraic, lower_bound=lower_bound, upper_bound=upper_bound,
              eps_policy=eps_policy)
    if np.all(np.isinf(dvec)):                                          # :377-389
        info = StepInfo()
        ypost, r = implicit_timestep(G, y0, math.inf, cache, abstol=abstol, maxiters=maxiters, info=info, **kw)
        log.n_newton = info.newton_iters
        log.exit_reason = EXIT_CONVERGED if r <= abstol else EXIT_INF_STEP_FAILED
        return ypost, r, log
    if dvec.ndim == 1:
        dbase = float(np.min(np.abs(dvec[np.isfinite(dvec)])))
        drel = dvec / dbase
    else:
        dbase = abs(float(dvec))
        drel = float(np.sign(dvec))
    stepper = SERTimeStepController(dbase, scale, max_delta)
    it = 0
    while it < maxiters and stepper.delta >= min_delta and r > abstol:   # :406
        it += 1
        info = StepInfo()
        y, nl = implicit_timestep(G, ypost, stepper.delta * drel, cache, abstol=inner_abstol,
                                  maxiters=inner_maxiters, info=info, **kw)
        ydot = G(y)
        old = r
        if has_bounds:                                                   # :410-413
            lb = np.full(n, -np.inf) if lower_bound is None else lower_bound
            ub = np.full(n, np.inf) if upper_bound is None else upper_bound
            mask = (lb + EPS <= y) & (y <= ub - EPS)
            with np.errstate(all="ignore"):
                r = float(np.linalg.norm(ydot * mask) / mask.sum())
        else:
            r = residual_norm(ydot)
        r = math.inf if math.isnan(r) else r
        log.n_newton += info.newton_iters
        if os.environ.get("ORACLE_PROGRESS"):                            # long fixture runs: one line per outer step
            print("oracle step", it, stepper.delta, r, nl, info.newton_iters, flush=True)
        if nl <= inner_abstol and r <= max_residual_growth * old:       # :418
            log.rows.append((it, stepper.delta, r, ACCEPTED, info.newton_iters))
            stepper.accept_step(old, r)
            ypost = y
        elif nl <= inner_abstol:                                         # :425
            log.rows.append((it, stepper.delta, math.nan, REJECTED_GROWTH, info.newton_iters))
            stepper.reject_step()
            r = old
        else:                                                            # :439
            st = REJECTED_NAN if math.isnan(nl) else (REJECTED_SINGULAR if math.isinf(nl) else REJECTED_INNER)
            log.rows.append((it, stepper.delta, math.nan, st, info.newton_iters))
            stepper.reject_step()
            r = old
    log.n_outer = it
    if r <= abstol:
        log.exit_reason = EXIT_CONVERGED
    elif stepper.delta < min_delta:
        log.exit_reason = EXIT_BELOW_MIN_DELTA
    else:
        log.exit_reason = EXIT_MAXITERS
    return ypost, r, log


# --------------------------------------------------------------------------------------
# pdesolve front + result type (src/pdesolve.jl:96-154, src/EconPDEs.jl:36-48)
# --------------------------------------------------------------------------------------
@dataclass
class EconPDEResult:
    solution: Dict[str, np.ndarray]
    residual_norm: float
    saved: Dict[str, np.ndarray]
    tolerance: float
    log: SolveLog = None

    @property
    def converged(self) -> bool:   # src/EconPDEs.jl:44-48: the worst residual of a time-dependent result
        return bool(np.all(np.asarray(self.residual_norm) <= self.tolerance))


def _expand(val, solnames, S, default, dtype):
    F = len(solnames)
    out = np.full(tuple(S) + (F,), default, dtype=dtype)
    if val is None:
        return out
    if isinstance(val, dict):
        for k, name in enumerate(solnames):
            if name in val:
                out[..., k] = val[name]
        return out
    out[...] = val
    return out


def pdesolve(pde, grid: Dict[str, np.ndarray], guess: Dict[str, np.ndarray], *, is_algebraic=None,
             bc=None, lower_bound=None, upper_bound=None, abstol=SQRT_EPS, maxiters=100,
             delta=1.0, **kw) -> EconPDEResult:
    """Stationary ``pdesolve`` (src/pdesolve.jl:96-154).  ``delta`` may be a scalar or a
    dict per unknown (expanded like ``_delta_array``, :376-394)."""
    statenames = tuple(grid.keys())
    solnames = tuple(guess.keys())
    grids = tuple(np.asarray(g, dtype=np.float64) for g in grid.values())
    S = tuple(g.shape[0] for g in grids)
    for name, v in guess.items():
        if tuple(np.shape(v)) != S:
            raise ValueError(f"the initial guess for `{name}` has size {np.shape(v)} but the state grid has size {S}")
    if any(s < 2 for s in S):
        raise ValueError("state grid entries must have at least 2 points")
    F = len(solnames)
    guess_array = np.stack([np.asarray(guess[k], dtype=np.float64) for k in solnames], axis=-1)
    prob = Problem(pde, statenames, grids, solnames, bc_array(bc, solnames, statenames, S))
    cache = JacobianCache(S, F)
    alg = _expand(is_algebraic, solnames, S, False, bool).reshape(-1, order="F") if is_algebraic is not None else None
    lb = _expand(lower_bound, solnames, S, -np.inf, np.float64).reshape(-1, order="F") if lower_bound is not None else None
    ub = _expand(upper_bound, solnames, S, np.inf, np.float64).reshape(-1, order="F") if upper_bound is not None else None
    if isinstance(delta, dict):
        delta = _expand(delta, solnames, S, 1.0, np.float64).reshape(-1, order="F")
    y0 = guess_array.reshape(-1, order="F")
    y, r, log = finiteschemesolve(prob.G, y0, cache, is_algebraic=alg, lower_bound=lb, upper_bound=ub,
                                  abstol=abstol, maxiters=maxiters, delta=delta, **kw)
    ya = y.reshape(S + (F,), order="F")
    solution = {name: np.array(ya[..., k]) for k, name in enumerate(solnames)}
    saved = prob.saved(y)
    if saved:
        saved = {**solution, **saved}
    return EconPDEResult(solution, r, saved, abstol, log)


def jacobian_diagnostics(J, delta, is_algebraic, S, F, tol=1e-6):
    """src/monotonicity.jl on an assembled sparse Jacobian J (scipy CSC) of the implicit step with scalar
    or per-row Δ: returns (nchecked[F], nwrong[F]) of _check_sign_convention! (:57-98) and the number and
    largest value of same-unknown spatial off-diagonal entries above tol (_check_monotonicity!, :100-133)."""
    n = J.shape[0]
    l = n // F
    dvec = np.full(n, float(delta)) if np.ndim(delta) == 0 else np.asarray(delta, dtype=np.float64)
    alg = np.zeros(n, dtype=bool) if is_algebraic is None else np.asarray(is_algebraic, dtype=bool).reshape(-1, order="F")
    diag = J.diagonal()
    nchecked, nwrong = np.zeros(F, dtype=int), np.zeros(F, dtype=int)
    for i in range(n):
        if alg[i] or not np.isfinite(dvec[i]):
            continue
        d = diag[i] - 1 / dvec[i]
        if not abs(d) > tol:
            continue
        k = i // l
        nchecked[k] += 1
        if np.sign(d) != np.sign(dvec[i]):
            nwrong[k] += 1
    coo = J.tocoo()
    cnt, mx = 0, 0.0
    for r, c, v in zip(coo.row, coo.col, coo.data):
        if not v > tol or r == c or r // l != c // l:
            continue
        ri = np.unravel_index(r % l, S, order="F")
        ci = np.unravel_index(c % l, S, order="F")
        if max(abs(a - b) for a, b in zip(ri, ci)) <= 1:
            cnt += 1
            mx = max(mx, v)
    return nchecked, nwrong, cnt, mx


def pdesolve_backward(pde, grid: Dict[str, np.ndarray], guess: Dict[str, np.ndarray], taus, *, abstol=SQRT_EPS,
                      maxiters=100, inner_maxiters=None, inner_abstol=None, bc=None,
                      has_time=False) -> EconPDEResult:
    """Time-dependent ``pdesolve(pde, grid, guess, τs)``: ``_solve_backward`` (src/pdesolve.jl:158-224).
    The guess is the terminal value at τs[end]; every earlier time is one ``implicit_timestep`` with
    Δ = τs[i] − τs[i−1] (:203), ``inner_maxiters = maxiters`` and ``inner_abstol = abstol`` by
    default (:158).  solution[name][..., i] is the value at τs[i]; residual_norm[i] is the Newton
    residual of the step that produced time i−1 (entry 0 stays 0).  ``has_time``: the PDE function has
    the three-argument form ``pde(state, u, t)`` (:166, :253-260) and is evaluated at t = τs[i] for the
    saved outputs of time i and for the step from τs[i] to τs[i−1] (:195)."""
    taus = np.asarray(taus, dtype=np.float64)
    inner_maxiters = maxiters if inner_maxiters is None else inner_maxiters
    inner_abstol = abstol if inner_abstol is None else inner_abstol
    statenames, solnames = tuple(grid.keys()), tuple(guess.keys())
    grids = tuple(np.asarray(g, dtype=np.float64) for g in grid.values())
    S = tuple(g.shape[0] for g in grids)
    F, nt = len(solnames), taus.shape[0]
    clock = {"t": float(taus[-1])}
    pde_now = (lambda state, u: pde(state, u, clock["t"])) if has_time else pde
    prob = Problem(pde_now, statenames, grids, solnames, bc_array(bc, solnames, statenames, S))
    cache = JacobianCache(S, F)
    y = np.stack([np.asarray(guess[k], dtype=np.float64) for k in solnames], axis=-1).reshape(-1, order="F")
    solution = {name: np.empty(S + (nt,)) for name in solnames}
    saved: Dict[str, np.ndarray] = {}
    rn = np.zeros(nt)
    for it in range(nt - 1, -1, -1):                                   # :191
        clock["t"] = float(taus[it])                                   # :195
        ya = y.reshape(S + (F,), order="F")
        for k, name in enumerate(solnames):                            # :192-194
            solution[name][..., it] = ya[..., k]
        sv = prob.saved(y)                                             # :196-201
        for name, v in (sv or {}).items():
            saved.setdefault(name, np.empty(S + (nt,)))[..., it] = v
        if it > 0:                                                     # :202-204
            y, rn[it] = implicit_timestep(prob.G, y, taus[it] - taus[it - 1], cache, abstol=inner_abstol,
                                          maxiters=inner_maxiters)
    if saved:
        saved = {**solution, **saved}
    return EconPDEResult(solution, rn, saved, abstol, None)

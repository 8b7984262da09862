"""Host side of the drop-in: `pdesolve(pde, grid, guess; ...) -> EconPDEResult`.

Mirrors the reference's input normalisation (src/pdesolve.jl:96-133, 293-394) and result
type (src/EconPDEs.jl:36-95), then makes ONE call into libeconpdes_cuda.so, which runs the
whole pseudo-transient Newton loop on the GPU.  Nothing here computes a residual or a
Jacobian: without the CUDA library the call raises.
"""
from __future__ import annotations

import ctypes as C
import math
import sys
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

from . import _lib as L
from .models import RegisteredModel

SQRT_EPS = math.sqrt(np.finfo(np.float64).eps)

_REMOVED = {  # src/finiteschemesolve.jl:18-32
    "maxdist": "`maxdist` was removed in EconPDEs 2.0; use `abstol` instead.",
    "innerdist": "`innerdist` was removed in EconPDEs 2.0; use `inner_abstol` instead.",
    "method": "`method` was removed in EconPDEs 2.0; use `alg = :newton`.",
    "autodiff": "`autodiff` was removed in EconPDEs 2.0; EconPDEs configures its Jacobian internally.",
    "algorithm": "`algorithm` was renamed in EconPDEs 2.0; use `alg` instead.",
    "iterations": "`iterations` was renamed in EconPDEs 2.0; use `maxiters` instead.",
    "inner_iterations": "`inner_iterations` was renamed in EconPDEs 2.0; use `inner_maxiters` instead.",
    "J0": "`J0` was renamed in EconPDEs 2.0; use `jac_prototype` instead.",
    "reformulation": "`reformulation` was removed in EconPDEs 2.0; bounded solves always use the minmax mixed-complementarity residual.",
    "autoscale": "`autoscale` was removed in EconPDEs 2.0.",
}
_LOWER_LEVEL = {  # src/pdesolve.jl:1-8
    "jac": "`jac` is only supported by `finiteschemesolve`; `pdesolve` constructs the discretized residual internally.",
    "jac_prototype": "`jac_prototype` is only supported by `finiteschemesolve`; `pdesolve` builds the stencil sparsity pattern internally.",
    "colorvec": "`colorvec` is only supported by `finiteschemesolve`; `pdesolve` computes the stencil coloring internally.",
    "fdcache": "`fdcache` is internal and cannot be passed to `pdesolve`.",
    "monotonicity_check": "`monotonicity_check` is internal; use `check_monotonicity` instead.",
}


@dataclass
class EconPDEResult:
    """src/EconPDEs.jl:36-48: solution, residual_norm, saved, tolerance; computed `converged`."""
    solution: "OrderedDict[str, np.ndarray]"
    residual_norm: float
    saved: "OrderedDict[str, np.ndarray]"
    tolerance: float
    # extras the reference prints or warns with
    exit_reason: int = 0
    log: tuple = ()
    n_outer: int = 0
    n_newton: int = 0
    n_linear_iters: int = 0
    kernel_launches: int = 0
    device_ms: float = 0.0
    # src/monotonicity.jl diagnostics of the first assembled Jacobian (see econpdes_result)
    sign_flags: int = 0
    nonmonotone_entries: int = 0
    nonmonotone_max: float = 0.0
    # quality of the Krylov solves that stand in for the reference's sparse LU (see econpdes_result)
    n_linear_solves: int = 0
    n_linear_loose: int = 0
    lin_max_relres: float = 0.0

    @property
    def converged(self) -> bool:
        # time-dependent results store one residual per time; convergence uses the worst one
        # (src/EconPDEs.jl:44-48); a NaN residual is not converged
        return bool(np.all(np.asarray(self.residual_norm) <= self.tolerance))

    @property
    def zero(self):  # src/EconPDEs.jl:49-51: deprecated 1.x alias
        import warnings
        warnings.warn("`result.zero` is deprecated; use `result.solution` instead.", DeprecationWarning, stacklevel=2)
        return self.solution

    @property
    def optional(self):  # src/EconPDEs.jl:52-54: deprecated 1.x alias
        import warnings
        warnings.warn("`result.optional` is deprecated; use `result.saved` instead.", DeprecationWarning, stacklevel=2)
        return self.saved

    def __iter__(self):  # tuple destructuring: solution, residual_norm, saved = result
        return iter((self.solution, self.residual_norm, self.saved))

    def __repr__(self):  # compact show (src/EconPDEs.jl:60-95): names and sizes, not arrays
        def fmt(d):
            return ", ".join(f"{k} ({'×'.join(str(s) for s in v.shape)})" for k, v in d.items())
        lines = ["EconPDEResult", f"  solution:      {fmt(self.solution)}"]
        if self.saved:
            lines.append(f"  saved:         {fmt(self.saved)}")
        rn = self.residual_norm
        if np.ndim(rn) == 0:
            lines.append(f"  residual_norm: {rn:.2e}")
        else:
            lines.append(f"  residual_norm: max {np.max(rn):.2e} over {len(rn)} times")
        lines.append(f"  converged:     {'true' if self.converged else 'false'} (tolerance {self.tolerance:.2e})")
        return "\n".join(lines)


def _as_ordered(x, what):
    if isinstance(x, OrderedDict):
        return x
    if isinstance(x, dict):
        # Python dicts keep insertion order (unlike Julia's Dict, src/pdesolve.jl:233-239)
        return OrderedDict(x)
    raise ValueError(f"pass {what} as a dict / OrderedDict of names to arrays")


def _check_generated_names(statenames, ynames):
    """src/pdesolve.jl:267-291."""
    seen = {}

    def reg(name, desc):
        if name in seen:
            raise ValueError(f"ambiguous names: the field `{name}` would mean both {seen[name]} and {desc}.")
        seen[name] = desc
    for s in ynames:
        reg(s, f"the value of unknown `{s}`")
        for x in statenames:
            reg(f"{s}{x}_up", f"the forward derivative of `{s}` in `{x}`")
            reg(f"{s}{x}_down", f"the backward derivative of `{s}` in `{x}`")
            reg(f"{s}{x}{x}", f"the second derivative of `{s}` in `{x}`")
        for i1, x1 in enumerate(statenames):
            for x2 in statenames[i1 + 1:]:
                reg(f"{s}{x1}{x2}", "the cross derivative")
                reg(f"{s}{x1}{x2}_up", "the directional cross derivative")
                reg(f"{s}{x1}{x2}_down", "the directional cross derivative")


def _bc_array(bc, ynames, statenames, S):
    """src/pdesolve.jl:312-335 (one slot per point; later dimensions overwrite earlier ones)."""
    out = np.zeros(tuple(S) + (len(ynames),), order="F")
    if bc is None:
        return None
    valid = [f"{y}{s}" for y in ynames for s in statenames]
    for key in bc:
        if key not in valid:
            raise ValueError(f"unknown `bc` entry `{key}`: valid entries are {valid}, i.e. unknown+state")
    for k, y in enumerate(ynames):
        for d, s in enumerate(statenames):
            key = f"{y}{s}"
            if key not in bc:
                continue
            lo, hi = bc[key]
            sl = [slice(None)] * len(S)
            sl[d] = 0
            out[tuple(sl) + (k,)] = lo
            sl[d] = S[d] - 1
            out[tuple(sl) + (k,)] = hi
    return out


def _per_unknown(val, ynames, S, default, dtype, keyword):
    """_bound_array / _fill_is_algebraic / _delta_array, src/pdesolve.jl:293-394."""
    F = len(ynames)
    if val is None:
        return None
    out = np.full(tuple(S) + (F,), default, dtype=dtype, order="F")
    if isinstance(val, dict):
        extra = set(val) - set(ynames)
        if extra:
            raise ValueError(f"`{keyword}` contains names that are not in the initial guess: {sorted(extra)}")
        if keyword != "Δ" and set(val) != set(ynames):
            raise ValueError(f"`{keyword}` must have the same names as the initial guess")
        for k, name in enumerate(ynames):
            if name in val:
                v = np.asarray(val[name])
                if v.ndim != 0 and v.shape != tuple(S):
                    raise ValueError(f"`{keyword}.{name}` has size {v.shape} but the state grid has size {tuple(S)}")
                if dtype is bool and v.dtype != np.bool_:
                    raise ValueError(f"`{keyword}.{name}` must be a Bool or an array of Bools")
                out[..., k] = v
        return out
    v = np.asarray(val)
    if v.ndim == 0:
        out[...] = v
    elif v.shape == tuple(S) + (F,):
        out[...] = v
    elif F == 1 and v.shape == tuple(S):
        out[..., 0] = v
    elif v.size == out.size:
        out[...] = v.reshape(out.shape, order="F")
    else:
        raise ValueError(f"`{keyword}` has size {v.shape}")
    return out


def _options(abstol, maxiters, kw) -> L.Options:
    lib = L.load()
    o = L.Options()
    lib.econpdes_default_options(C.byref(o))
    o.abstol = abstol
    o.maxiters = maxiters
    names = {"scale": "scale", "minΔ": "min_delta", "min_delta": "min_delta", "maxΔ": "max_delta",
             "max_delta": "max_delta", "max_residual_growth": "max_residual_growth",
             "inner_maxiters": "inner_maxiters", "inner_abstol": "inner_abstol", "eps_policy": "eps_policy",
             "lin_rtol": "lin_rtol", "lin_maxiters": "lin_maxiters", "lin_restart": "lin_restart",
             "linear_solver": "linear_solver", "monotonicity_tol": "monotonicity_tol"}
    for k in list(kw):
        if k in names:
            setattr(o, names[k], kw.pop(k))
    return o


def _problem_description(ynames, statenames, S) -> str:
    """src/pdesolve.jl:227-231, e.g. "2 unknowns (pA, pB) on grid x×y×z = 2×3×4"."""
    n = len(ynames)
    return "%d %s (%s) on grid %s = %s" % (n, "unknown" if n == 1 else "unknowns", ", ".join(ynames),
                                           "×".join(statenames), "×".join(str(s) for s in S))


def _print_log(res: EconPDEResult, description: str, file):
    """The verbose table of src/finiteschemesolve.jl:403-404, 421, 435, 445, 453-455."""
    print("Solving for " + description, file=file)
    if res.n_outer > 0 or res.log:
        print("Iter TimeStep Residual", file=file)
        print("---- -------- --------", file=file)
    for row in res.log:
        if row["status"] == L.STEP_ACCEPTED:
            print("%4d %8.2e %8.2e" % (row["iter"], row["delta"], row["residual"]), file=file)
        elif row["status"] == L.STEP_REJECTED_GROWTH:
            growth = row["residual"]      # the generic engine stores residual/old_residual on these rows (:433-435)
            text = "" if growth != growth else ("%.1f× " % growth if np.isfinite(growth) else "∞ ")
            print("%4d %8.2e        ✗ (residual grew %s> max_residual_growth)" % (row["iter"], row["delta"], text), file=file)
        else:
            reason = {L.STEP_REJECTED_NAN: " (model returned NaN)", L.STEP_REJECTED_SINGULAR: " (singular Jacobian)"}
            print("%4d %8.2e        ✗%s" % (row["iter"], row["delta"], reason.get(row["status"], "")), file=file)
    if res.converged and res.exit_reason == L.EXIT_CONVERGED:
        ms = res.device_ms
        t = "%.0fms" % ms if ms < 1000 else "%.1fs" % (ms / 1000)
        if res.n_outer == 0 and not res.log:      # one-shot solve with an infinite step (:383)
            print("Converged (%s)" % t, file=file)
        else:
            print("Converged after %d time steps (%s)" % (res.n_outer, t), file=file)


def _diagnostic_warnings(res: EconPDEResult, ynames) -> list:
    """@warn texts of src/monotonicity.jl:84-96 (sign convention, checked by default) and a summary of
    the opt-in stencil check (:100-133; the reference prints one warning per offending entry)."""
    out = []
    if res.sign_flags & L.SIGN_WHOLE_SYSTEM:
        out.append("The residual Jacobian has a negative diagonal at every grid point: the PDE function most likely "
                   "returns `RHS - ρv` instead of `-(RHS - ρv)`, which makes the false transient unstable (the solve "
                   "typically diverges from iteration 1). Flip the sign of the returned time derivative.")
    else:
        for k, name in enumerate(ynames):
            if res.sign_flags & (1 << k):
                out.append(f"The equation for `{name}` has a residual-Jacobian diagonal opposite in sign to its "
                           "pseudo-time step at every grid point, which makes its false transient unstable (the solve "
                           f"typically diverges from iteration 1). Flip the sign of `Δ.{name}` — positive Δ marches an "
                           "equation backward like a value function and needs a positive diagonal; negative Δ marches "
                           "it forward like a tâtonnement price update and needs a negative one — or flip the sign of "
                           "the returned time derivative.")
    if res.nonmonotone_entries > 0:
        out.append("Non-monotone finite-difference stencil detected: positive residual off-diagonal implies a "
                   f"negative generator weight ({res.nonmonotone_entries} same-variable spatial entries of the first "
                   f"Jacobian above the tolerance, largest {res.nonmonotone_max:.3g})")
    return out


def _warn_text(res: EconPDEResult, min_delta: float) -> Optional[str]:
    """@warn texts of src/finiteschemesolve.jl:387, 456-461."""
    if res.converged:
        return None
    r, tol = "%.2e" % res.residual_norm, "%.2e" % res.tolerance
    if res.exit_reason == L.EXIT_INF_STEP_FAILED:
        return (f"did not converge: residual {r} > tolerance {tol}. `Δ = Inf` disables pseudo-transient continuation "
                "— try a finite initial time step `Δ`, a better initial guess, or more `maxiters`.")
    if res.exit_reason == L.EXIT_BELOW_MIN_DELTA:
        return (f"did not converge after {res.n_outer} time steps: the pseudo-transient time step fell below "
                f"`minΔ` = {min_delta} (residual {r} > tolerance {tol}). This usually means a non-monotone "
                "finite-difference scheme (wrong upwind direction) or a poor initial guess — try "
                "`check_monotonicity = true` or a better initial guess.")
    return (f"did not converge after {res.n_outer} time steps: residual {r} > tolerance {tol}. If the residual was still "
            "falling, increase `maxiters`; otherwise try a better initial guess or `check_monotonicity = true`.")


def pdesolve(pde, grid, guess, taus=None, *, is_algebraic=None, bc=None, lower_bound=None, upper_bound=None,
             abstol=SQRT_EPS, verbose=True, warn=True, alg="newton", maxiters=100, check_monotonicity=False,
             file=None, **kwargs) -> EconPDEResult:
    """Stationary `pdesolve` on the GPU (src/pdesolve.jl:96-154).

    `pde` must be a registered model (econpdes_b200.models): the device cannot call back into
    a host closure per grid point.  `grid` / `guess` are dicts name -> array like the
    reference's NamedTuples; `Δ` (or `delta`) may be a scalar or a dict per unknown.
    """
    import warnings
    for k, msg in _REMOVED.items():
        if k in kwargs:
            raise ValueError(msg)
    for k, msg in _LOWER_LEVEL.items():
        if k in kwargs:
            raise ValueError(msg)
    if not isinstance(pde, RegisteredModel):
        raise TypeError("the device path needs a registered model functor (econpdes_b200.models); arbitrary host "
                        "closures `pde(state, u)` only run on the reference's CPU path")
    if taus is not None:   # _reject_time_dependent_stationary_keywords + _check_time_grid, src/pdesolve.jl:10-15, 241-251
        for k in ("Δ", "delta", "scale", "minΔ", "min_delta", "maxΔ", "max_delta", "max_residual_growth"):
            if k in kwargs:
                raise ValueError(f"`{k}` only applies to stationary pseudo-transient continuation; time-dependent "
                                 "`pdesolve(..., τs)` uses the spacing of `τs` as the time step.")
        taus = np.ascontiguousarray(np.asarray(taus, dtype=np.float64))
        if taus.ndim != 1 or taus.size == 0:
            raise ValueError("the time grid `τs` must contain at least one time")
        if not np.all(np.isfinite(taus)):
            raise ValueError("the time grid `τs` must contain finite real numbers")
        if np.any(np.diff(taus) <= 0):
            raise ValueError("the time grid `τs` must be strictly increasing")
        # src/pdesolve.jl:158: inner_maxiters = maxiters, inner_abstol = abstol unless given
        kwargs.setdefault("inner_maxiters", maxiters)
        kwargs.setdefault("inner_abstol", abstol)
    if alg not in ("newton", ":newton"):
        raise ValueError(f"NLsolve `alg` must be `:newton` on the device path; got `{alg!r}`")
    grid = _as_ordered(grid, "the state grid")
    guess = _as_ordered(guess, "the initial guess")
    statenames, ynames = tuple(grid), tuple(guess)
    _check_generated_names(list(statenames), list(ynames))
    if statenames != tuple(pde.statenames) or ynames != tuple(pde.solnames):
        raise ValueError(f"model `{pde.name}` expects states {pde.statenames} and unknowns {pde.solnames}; "
                         f"got {statenames} and {ynames}")
    grids = [np.ascontiguousarray(np.asarray(g, dtype=np.float64)) for g in grid.values()]
    if any(g.ndim != 1 for g in grids):
        raise ValueError("state grid entries must be vectors")
    S = tuple(g.shape[0] for g in grids)
    if any(s < 2 for s in S):
        raise ValueError("state grid entries must have at least 2 points")
    for name, v in guess.items():
        if tuple(np.shape(v)) != S:
            raise ValueError(f"the initial guess (e.g. terminal value) for `{name}` has size {np.shape(v)} "
                             f"but the state grid has size {S}")
    F = len(ynames)
    guess_array = np.asfortranarray(np.stack([np.asarray(guess[k], dtype=np.float64) for k in ynames], axis=-1))
    delta = kwargs.pop("Δ", kwargs.pop("delta", 1.0))
    delta_arr = None
    if isinstance(delta, dict):
        delta_arr = _per_unknown(delta, ynames, S, 1.0, np.float64, "Δ")
    bc_arr = _bc_array(bc, ynames, statenames, S)
    alg_arr = _per_unknown(is_algebraic, ynames, S, False, bool, "is_algebraic")
    lb_arr = _per_unknown(lower_bound, ynames, S, -np.inf, np.float64, "lower_bound")
    ub_arr = _per_unknown(upper_bound, ynames, S, np.inf, np.float64, "upper_bound")
    opts = _options(abstol, maxiters, kwargs)
    if delta_arr is None:
        opts.delta = float(delta)
    if kwargs:
        raise TypeError(f"unknown keyword arguments {sorted(kwargs)}")

    lib = L.load()
    mid = L.check(lib.econpdes_model_id(pde.name.encode()))
    prm = pde.params()
    p = L.Problem()
    p.model_id, p.ndims, p.nfuncs, p.nparams = mid, len(S), F, prm.shape[0]
    for d in range(len(S)):
        p.size[d] = S[d]
        p.grid[d] = L.dptr(grids[d])
    keep = [prm, guess_array]
    p.params = L.dptr(prm) if prm.shape[0] else None
    p.guess = L.dptr(guess_array)

    def fvec(a, dtype=np.float64):
        if a is None:
            return None
        v = np.ascontiguousarray(a.reshape(-1, order="F"), dtype=dtype)
        keep.append(v)
        return v
    v = fvec(bc_arr); p.bc = L.dptr(v)
    v = fvec(alg_arr, np.uint8); p.is_algebraic = None if v is None else v.ctypes.data_as(L.c_uint8_p)
    v = fvec(lb_arr); p.lower_bound = L.dptr(v)
    v = fvec(ub_arr); p.upper_bound = L.dptr(v)
    v = fvec(delta_arr); p.delta_vec = L.dptr(v)

    opts.check_monotonicity = 1 if check_monotonicity else 0
    if taus is not None:
        return _solve_backward(lib, p, opts, taus, pde, S, F, ynames, statenames, abstol, verbose, warn,
                               file or sys.stdout, keep)
    n = int(np.prod(S)) * F
    sol = np.empty(n)
    nsaved = len(pde.savednames)
    saved = np.empty(int(np.prod(S)) * nsaved) if nsaved else None
    cap = max(int(maxiters), 1)
    log = (L.LogRow * cap)()
    r = L.Result()
    r.solution = L.dptr(sol)
    r.saved = L.dptr(saved)
    r.log = log
    r.log_capacity = cap
    L.check(lib.econpdes_solve_stationary(C.byref(p), C.byref(opts), C.byref(r)))

    ya = sol.reshape(S + (F,), order="F")
    solution = OrderedDict((name, np.array(ya[..., k])) for k, name in enumerate(ynames))
    sv = OrderedDict()
    if nsaved:
        sa = saved.reshape(S + (nsaved,), order="F")
        sv = OrderedDict(solution)
        for k, name in enumerate(pde.savednames):
            sv[name] = np.array(sa[..., k])
    rows = tuple(dict(iter=log[i].iter, status=log[i].status, newton_iters=log[i].newton_iters,
                      linear_iters=log[i].linear_iters, delta=log[i].delta, residual=log[i].residual)
                 for i in range(r.log_len))
    res = EconPDEResult(solution, r.residual_norm, sv, abstol, r.exit_reason, rows, r.n_outer, r.n_newton,
                        r.n_linear_iters, r.kernel_launches, r.device_ms, r.sign_flags, r.nonmonotone_entries,
                        r.nonmonotone_max, r.n_linear_solves, r.n_linear_loose, r.lin_max_relres)
    for txt in _diagnostic_warnings(res, ynames):   # src/monotonicity.jl (emitted regardless of `warn`, like @warn there)
        warnings.warn(txt)
    if verbose:
        desc = _problem_description(ynames, statenames, S)
        _print_log(res, desc, file or sys.stdout)
        if res.exit_reason == L.EXIT_ALREADY_ZERO:
            warnings.warn("G! already returns zero with the initial value")
    if warn:
        txt = _warn_text(res, opts.min_delta)
        if txt:
            warnings.warn(txt)
    return res


def _solve_backward(lib, p, opts, taus, pde, S, F, ynames, statenames, abstol, verbose, warn, file, keep):
    """Time-dependent `pdesolve(pde, grid, guess, τs)` (src/pdesolve.jl:158-224): every unknown gains a
    trailing time dimension, `residual_norm` is one value per time."""
    import warnings
    nt = taus.shape[0]
    l = int(np.prod(S))
    nsaved = len(pde.savednames)
    sol = np.empty(l * F * nt)
    saved = np.empty(l * nsaved * nt) if nsaved else None
    rn = np.zeros(nt)
    nit = np.zeros(nt, dtype=np.int32)
    ms, launches = C.c_double(), C.c_int64()
    L.check(lib.econpdes_solve_backward(C.byref(p), C.byref(opts), L.dptr(taus), nt, L.dptr(sol), L.dptr(saved),
                                        L.dptr(rn), nit.ctypes.data_as(L.c_int32_p), C.byref(ms), C.byref(launches)))
    ya = sol.reshape(S + (F, nt), order="F")
    solution = OrderedDict((name, np.array(ya[..., k, :])) for k, name in enumerate(ynames))
    sv = OrderedDict()
    if nsaved:
        sa = saved.reshape(S + (nsaved, nt), order="F")
        sv = OrderedDict(solution)
        for k, name in enumerate(pde.savednames):
            sv[name] = np.array(sa[..., k, :])
    res = EconPDEResult(solution, rn, sv, abstol, L.EXIT_CONVERGED, (), nt - 1, int(nit.sum()), 0,
                        launches.value, ms.value)
    if verbose:
        desc = _problem_description(ynames, statenames, S)
        print("Solving for %s, backward from τ = %g to %g (%d steps)" % (desc, taus[-1], taus[0], nt - 1), file=file)
        print("    Time Residual", file=file)
        print("-------- --------", file=file)
        for i in range(nt - 1, 0, -1):
            print("%8g %8.2e" % (taus[i - 1], rn[i]), file=file)
        nfailed = int(np.sum(~(rn[1:] <= abstol)))
        t = "%.0fms" % ms.value if ms.value < 1000 else "%.1fs" % (ms.value / 1000)
        if nfailed == 0:
            print("Completed %d time steps (%s)" % (nt - 1, t), file=file)
        else:
            print("Completed %d time steps (%s): %d steps did not converge" % (nt - 1, t, nfailed), file=file)
    if warn:
        for i in range(nt - 1, 0, -1):
            if not (rn[i] <= abstol):
                warnings.warn("the implicit time step at τ = %g did not converge (residual norm: %.2e); the solution "
                              "at this and earlier times may be inaccurate — try a finer time grid or more "
                              "`inner_maxiters`" % (taus[i - 1], rn[i]))
    return res


@dataclass
class BatchResult:
    solution: np.ndarray       # [B, S]
    residual_norm: np.ndarray  # [B]
    converged: np.ndarray      # [B] bool
    exit_reason: np.ndarray
    n_outer: np.ndarray
    n_newton: np.ndarray
    device_ms: float
    kernel_launches: int
    log: Optional[list] = None


def pdesolve_batched(model_cls, params: np.ndarray, grids: np.ndarray, guess: np.ndarray, *, abstol=SQRT_EPS,
                     maxiters=100, log_capacity=0, **kwargs) -> BatchResult:
    """B independent 1-D problems (a loop of `pdesolve` calls in the reference) in one launch.

    params [B, nparams] (field order of `model_cls`), grids [B, S], guess [B, S].
    """
    lib = L.load()
    delta = kwargs.pop("Δ", kwargs.pop("delta", 1.0))
    opts = _options(abstol, maxiters, kwargs)
    opts.delta = float(delta)
    if kwargs:
        raise TypeError(f"unknown keyword arguments {sorted(kwargs)}")
    params = np.ascontiguousarray(params, dtype=np.float64)
    grids = np.ascontiguousarray(grids, dtype=np.float64)
    guess = np.ascontiguousarray(guess, dtype=np.float64)
    B, S = grids.shape
    if guess.shape != (B, S):
        raise ValueError(f"guess has shape {guess.shape}, expected {(B, S)}")
    mid = L.check(lib.econpdes_model_id(model_cls.name.encode()))
    h = C.c_void_p()
    L.check(lib.econpdes_batch_create(mid, B, S, C.byref(h)))
    try:
        L.check(lib.econpdes_batch_upload(h, L.dptr(params) if params.size else None, L.dptr(grids), L.dptr(guess)))
        if log_capacity:
            L.check(lib.econpdes_batch_set_log_capacity(h, log_capacity))
        ms = C.c_double()
        L.check(lib.econpdes_batch_solve(h, C.byref(opts), C.byref(ms)))
        sol = np.empty((B, S))
        rn = np.empty(B)
        er = np.empty(B, dtype=np.int32)
        no = np.empty(B, dtype=np.int32)
        nn = np.empty(B, dtype=np.int32)
        logbuf = (L.LogRow * (B * log_capacity))() if log_capacity else None
        loglen = np.zeros(B, dtype=np.int32)
        L.check(lib.econpdes_batch_download(h, L.dptr(sol), L.dptr(rn), er.ctypes.data_as(L.c_int32_p),
                                            no.ctypes.data_as(L.c_int32_p), nn.ctypes.data_as(L.c_int32_p),
                                            logbuf, log_capacity, loglen.ctypes.data_as(L.c_int32_p) if log_capacity else None))
        launches = lib.econpdes_batch_kernel_launches(h)
        logs = None
        if log_capacity:
            logs = [[dict(iter=logbuf[b * log_capacity + i].iter, status=logbuf[b * log_capacity + i].status,
                          newton_iters=logbuf[b * log_capacity + i].newton_iters,
                          delta=logbuf[b * log_capacity + i].delta, residual=logbuf[b * log_capacity + i].residual)
                     for i in range(loglen[b])] for b in range(B)]
    finally:
        lib.econpdes_batch_destroy(h)
    return BatchResult(sol, rn, rn <= abstol, er, no, nn, ms.value, launches, logs)


def last_batched_output_mode() -> int:
    """1: this thread's last one-call batched solve wrote its results straight into page-locked host memory
    (zero-copy); 0: it staged them through device buffers and copies; -1: there was none."""
    return int(L.load().econpdes_last_batched_output_mode())


def set_batch_points_per_thread(points: int) -> int:
    """Tuning knob of the batched engine (grid points owned by one thread; 0 = automatic).
    Returns the previous setting."""
    return L.check(L.load().econpdes_set_batch_points_per_thread(int(points)))


def solve_batched_one_call(model_cls, params, grids, guess, *, abstol=SQRT_EPS, maxiters=100, out=None, **kwargs):
    """`econpdes_solve_batched`: upload, solve and download in one C call that pipelines slices of
    the batch over three streams.  `out` = optional preallocated (solution, residual_norm, exit_reason, n_outer,
    n_newton) arrays: with a page-locked `solution` the kernel stores its results straight into it (zero-copy,
    `last_batched_output_mode() == 1`).  Returns (solution, residual_norm, exit_reason, n_outer, n_newton)."""
    lib = L.load()
    delta = kwargs.pop("Δ", kwargs.pop("delta", 1.0))
    opts = _options(abstol, maxiters, kwargs)
    opts.delta = float(delta)
    if kwargs:
        raise TypeError(f"unknown keyword arguments {sorted(kwargs)}")
    params = np.ascontiguousarray(params, dtype=np.float64)
    grids = np.ascontiguousarray(grids, dtype=np.float64)
    guess = np.ascontiguousarray(guess, dtype=np.float64)
    B, S = grids.shape
    if guess.shape != (B, S):
        raise ValueError(f"guess has shape {guess.shape}, expected {(B, S)}")
    mid = L.check(lib.econpdes_model_id(model_cls.name.encode()))
    if out is None:
        out = (np.empty((B, S)), np.empty(B), np.empty(B, dtype=np.int32), np.empty(B, dtype=np.int32),
               np.empty(B, dtype=np.int32))
    sol, rn, er, no, nn = out
    L.check(lib.econpdes_solve_batched(mid, B, S, L.dptr(params) if params.size else None, L.dptr(grids),
                                       L.dptr(guess), C.byref(opts), L.dptr(sol), L.dptr(rn),
                                       er.ctypes.data_as(L.c_int32_p), no.ctypes.data_as(L.c_int32_p),
                                       nn.ctypes.data_as(L.c_int32_p)))
    return sol, rn, er, no, nn


def solve_batched_compact(model_cls, params, range_lo, points, n_range: int, guess_fill: float = 1.0, *,
                          abstol=SQRT_EPS, maxiters=100, out=None, **kwargs):
    """`econpdes_solve_batched_inputs` with compact inputs (`econpdes_batch_inputs`): the first `n_range` points of
    problem b are Julia's range(range_lo[b], points[b, 0], length = n_range + 1)[1:n_range], built on the device,
    followed by the explicit `points[b]`; every unknown starts at `guess_fill`.  Same pipelined one-call path as
    `solve_batched_one_call`, without shipping B x S grids and guesses.  `out` = optional preallocated
    (solution, residual_norm, exit_reason, n_outer, n_newton) arrays (page-locked ones overlap the copies)."""
    lib = L.load()
    delta = kwargs.pop("Δ", kwargs.pop("delta", 1.0))
    opts = _options(abstol, maxiters, kwargs)
    opts.delta = float(delta)
    if kwargs:
        raise TypeError(f"unknown keyword arguments {sorted(kwargs)}")
    params = np.ascontiguousarray(params, dtype=np.float64)
    range_lo = np.ascontiguousarray(range_lo, dtype=np.float64)
    points = np.ascontiguousarray(points, dtype=np.float64)
    B, npts = points.shape
    S = int(n_range) + npts
    if range_lo.shape != (B,):
        raise ValueError(f"range_lo has shape {range_lo.shape}, expected {(B,)}")
    mid = L.check(lib.econpdes_model_id(model_cls.name.encode()))
    if out is None:
        out = (np.empty((B, S)), np.empty(B), np.empty(B, dtype=np.int32), np.empty(B, dtype=np.int32),
               np.empty(B, dtype=np.int32))
    sol, rn, er, no, nn = out
    inp = L.BatchInputs()
    inp.params = L.dptr(params) if params.size else None
    inp.grids = None
    inp.n_range = int(n_range)
    inp.range_lo = L.dptr(range_lo)
    inp.points = L.dptr(points)
    inp.guess = None
    inp.guess_fill = float(guess_fill)
    L.check(lib.econpdes_solve_batched_inputs(mid, B, S, C.byref(inp), C.byref(opts), L.dptr(sol), L.dptr(rn),
                                              er.ctypes.data_as(L.c_int32_p), no.ctypes.data_as(L.c_int32_p),
                                              nn.ctypes.data_as(L.c_int32_p)))
    return sol, rn, er, no, nn


class Handle:
    """Repeated residual / Jacobian / Newton-step evaluations on one device-resident problem
    (the `G!`, `finite_difference_jacobian!` and `implicit_timestep` seams of
    src/finiteschemesolve.jl:153-230)."""

    def __init__(self, pde: RegisteredModel, grids, guess_array: np.ndarray, bc: Optional[np.ndarray] = None):
        lib = L.load()
        self.lib = lib
        self.model = pde
        self.grids = [np.ascontiguousarray(g, dtype=np.float64) for g in grids]
        self.S = tuple(g.shape[0] for g in self.grids)
        self.F = pde.nfuncs
        self.n = int(np.prod(self.S)) * self.F
        self._prm = pde.params()
        self._guess = np.ascontiguousarray(np.asarray(guess_array, dtype=np.float64).reshape(-1, order="F"))
        self._bc = None if bc is None else np.ascontiguousarray(np.asarray(bc, dtype=np.float64).reshape(-1, order="F"))
        p = L.Problem()
        p.model_id = L.check(lib.econpdes_model_id(pde.name.encode()))
        p.ndims, p.nfuncs, p.nparams = len(self.S), self.F, self._prm.shape[0]
        for d, g in enumerate(self.grids):
            p.size[d] = self.S[d]
            p.grid[d] = L.dptr(g)
        p.params = L.dptr(self._prm) if self._prm.shape[0] else None
        p.guess = L.dptr(self._guess)
        p.bc = L.dptr(self._bc)
        self.h = C.c_void_p()
        L.check(lib.econpdes_create(C.byref(p), C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.econpdes_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _vec(self, a):
        return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1, order="F"))

    def residual(self, y):
        yv = self._vec(y)
        out = np.empty(self.n)
        L.check(self.lib.econpdes_residual(self.h, L.dptr(yv), L.dptr(out)))
        return out

    def saved(self, y):
        ns = len(self.model.savednames)
        yv = self._vec(y)
        out = np.empty(int(np.prod(self.S)) * max(ns, 1))
        L.check(self.lib.econpdes_saved(self.h, L.dptr(yv), L.dptr(out)))
        return out.reshape(self.S + (max(ns, 1),), order="F")

    def jacobian_csc(self, y, ypost=None, delta=math.inf, eps_policy=L.EPS_FINITEDIFF_COLOR_NORM):
        yv = self._vec(y)
        yp = None if ypost is None else self._vec(ypost)
        size = (C.c_int64 * len(self.S))(*self.S)
        nnz = self.lib.econpdes_stencil_nnz(len(self.S), size, self.F)
        nz = np.empty(nnz)
        ncol = self.F * int(np.prod([min(3, s) for s in self.S]))
        eps = np.empty(ncol)
        L.check(self.lib.econpdes_jacobian_csc(self.h, L.dptr(yv), L.dptr(yp), float(delta), eps_policy, L.dptr(nz),
                                               L.dptr(eps)))
        return nz, eps

    def bench_kernels(self, y, reps=5):
        """ms per launch: residual, implicit residual+norms, Jacobian, SpMV, V-cycle, MG set-up"""
        yv = self._vec(y)
        out = np.zeros(6)
        L.check(self.lib.econpdes_bench_kernels(self.h, L.dptr(yv), reps, L.dptr(out)))
        return dict(zip(("residual", "implicit_residual_norms", "jacobian", "spmv", "vcycle", "mg_setup"), out))

    def linear_solve(self, b, **kw):
        opts = _options(SQRT_EPS, 100, kw)
        bv = self._vec(b)
        x = np.empty(self.n)
        it = C.c_int32()
        rr = C.c_double()
        rc = L.check(self.lib.econpdes_linear_solve(self.h, C.byref(opts), L.dptr(bv), L.dptr(x), C.byref(it),
                                                    C.byref(rr)))
        return x, it.value, rr.value, rc

    def implicit_timestep(self, ypost, delta, maxiters=10, abstol=SQRT_EPS, **kw):
        opts = _options(SQRT_EPS, 100, kw)
        yp = self._vec(ypost)
        out = np.empty(self.n)
        nl = C.c_double()
        it = C.c_int32()
        L.check(self.lib.econpdes_implicit_timestep(self.h, C.byref(opts), L.dptr(yp), float(delta), maxiters, abstol,
                                                    L.dptr(out), C.byref(nl), C.byref(it)))
        return out, nl.value, it.value


def stencil_colors(S, F) -> np.ndarray:
    """stencil_colors(S, F), src/utils/coloring.jl:52-70 (host-only entry point)."""
    lib = L.load()
    S = tuple(int(s) for s in S)
    size = (C.c_int64 * len(S))(*S)
    out = np.empty(int(np.prod(S)) * F, dtype=np.int64)
    L.check(lib.econpdes_stencil_colors(len(S), size, F, out.ctypes.data_as(L.c_int64_p)))
    return out


def stencil_pattern(S, F):
    """local_stencil_jacobian(S, F) as 1-based CSC (colptr, rowval), src/utils/coloring.jl:18-35."""
    lib = L.load()
    S = tuple(int(s) for s in S)
    size = (C.c_int64 * len(S))(*S)
    nnz = L.check(lib.econpdes_stencil_nnz(len(S), size, F))
    colptr = np.empty(int(np.prod(S)) * F + 1, dtype=np.int64)
    rowval = np.empty(nnz, dtype=np.int64)
    L.check(lib.econpdes_stencil_pattern(len(S), size, F, colptr.ctypes.data_as(L.c_int64_p),
                                         rowval.ctypes.data_as(L.c_int64_p)))
    return colptr, rowval

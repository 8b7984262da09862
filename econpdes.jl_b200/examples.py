"""Grids and guesses of the example models named by the configs (host-side set-up only).

These mirror the few lines each reference example spends building its state grid and
initial guess; they use plain numpy (`linspace`) and are the synthetic-input generators of
bench.py.  Parity tests build their inputs with the oracle's Julia-exact `range` instead and
feed the *same arrays* to both sides, so nothing here is on a parity path.
"""
from __future__ import annotations

import math

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64_uniforms(seeds: np.ndarray, count: int) -> np.ndarray:
    """`count` uniforms per seed from splitmix64, vectorised over seeds."""
    state = np.asarray(seeds, dtype=np.uint64).copy()
    out = np.empty((state.shape[0], count))
    with np.errstate(over="ignore"):
        for i in range(count):
            state = state + np.uint64(0x9E3779B97F4A7C15)
            z = state.copy()
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            out[:, i] = (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    return out


def _two_prod(a, b):
    """a * b = p + e exactly (Dekker / Veltkamp splitting; numpy has no fma)."""
    p = a * b
    c = 134217729.0   # 2^27 + 1
    ta, tb = c * a, c * b
    ah = ta - (ta - a)
    bh = tb - (tb - b)
    al, bl = a - ah, b - bh
    e = ((ah * bh - p) + ah * bl + al * bh) + al * bl
    return p, e


def julia_range_rows(a, b, n: int) -> np.ndarray:
    """Row r = Julia's range(a[r], b[r], length = n), i.e. the correctly rounded a + i (b - a)/(n - 1): the same
    double-double evaluation as the device grid generator (range_point, csrc/stencil.cuh), vectorised, so a sweep
    whose grids are built on the device solves bit-identical problems to one that ships explicit grids."""
    a = np.asarray(a, dtype=np.float64)[:, None]
    b = np.asarray(b, dtype=np.float64)[:, None]
    i = np.arange(n, dtype=np.float64)[None, :]
    m1, m2, d = (n - 1) - i, i, float(n - 1)
    with np.errstate(invalid="ignore", over="ignore"):
        p1, e1 = _two_prod(a, m1)
        p2, e2 = _two_prod(b, m2)
        sh = p1 + p2
        bb = sh - p1
        sl = ((p1 - (sh - bb)) + (p2 - bb)) + (e1 + e2)
        h = sh + sl
        lo = sl - (h - sh)
        q1 = h / d
        ph, pl = _two_prod(q1, np.full_like(q1, d))
        r = ((h - ph) - pl) + lo
        out = q1 + r / d
    out[:, 0] = a[:, 0]
    out[:, -1] = b[:, 0]
    return out


def campbell_cochrane_sweep_compact(first: int, count: int, sn: int = 1000):
    """The same sweep as `campbell_cochrane_sweep` in the compact form of `econpdes_batch_inputs`:
    (params [B,6], range_lo [B], points [B, sn//10 - 1], n_range, guess_fill).  The grid of a problem is
    range(range_lo, points[0], length = n_range + 1)[1:n_range] followed by `points`
    (initialize_stategrid, examples/asset_pricing/campbell_cochrane.jl:40-48)."""
    params, shigh = _campbell_cochrane_sweep_head(first, count, sn)
    nh = sn // 10
    return (params, np.full(count, -300.0), np.ascontiguousarray(shigh[:, 1:]), sn - nh - 1, 1.0)


def _campbell_cochrane_sweep_head(first: int, count: int, sn: int):
    u = splitmix64_uniforms(np.uint64(0xEC0DE5) + np.arange(first, first + count, dtype=np.uint64), 5)
    mu = 0.015 + (0.023 - 0.015) * u[:, 0]
    sigma = 0.012 + (0.018 - 0.012) * u[:, 1]
    gamma = 1.5 + (2.5 - 1.5) * u[:, 2]
    rho = 0.09 + (0.14 - 0.09) * u[:, 3]
    kappa_s = 0.11 + (0.17 - 0.11) * u[:, 4]
    b = np.zeros(count)
    params = np.stack([mu, sigma, gamma, rho, kappa_s, b], axis=1)
    Sbar = sigma * np.sqrt(gamma / (kappa_s - b / gamma))
    sbar = np.log(Sbar)
    smax = sbar + 0.5 * (1 - Sbar * Sbar)
    nh = sn // 10
    with np.errstate(divide="ignore"):
        shigh = np.log(julia_range_rows(np.zeros(count), np.exp(smax), nh))
    return params, shigh


def campbell_cochrane_sweep(first: int, count: int, sn: int = 1000):
    """Problems first..first+count-1 of the C4 calibration sweep: (params [B,6], grids [B,S],
    guess [B,S]).  Parameters are i.i.d. uniform draws around the reference defaults
    (examples/asset_pricing/campbell_cochrane.jl:28-35); each problem has its own state grid
    (initialize_stategrid, :40-48) and the example's guess p = 1 (:50)."""
    params, shigh = _campbell_cochrane_sweep_head(first, count, sn)
    nh = sn // 10
    slow = julia_range_rows(np.full(count, -300.0), shigh[:, 1], sn - nh)
    grids = np.ascontiguousarray(np.concatenate([slow[:, :-1], shigh[:, 1:]], axis=1))
    return params, grids, np.ones_like(grids)


def bansal_yaron_grid(n_mu: int, n_v: int, m=None):
    """examples/asset_pricing/bansal_yaron.jl:43-50 (needs scipy for the two quantiles)."""
    from scipy import stats
    from .models import BansalYaronModel
    m = m or BansalYaronModel()
    sd = math.sqrt(m.nu_mu ** 2 * m.vbar / (2 * m.kappa_mu))
    mus = np.linspace(stats.norm.ppf(0.01, m.mubar, sd), stats.norm.ppf(0.99, m.mubar, sd), n_mu)
    shape, scale = 2 * m.kappa_v * m.vbar / m.nu_v ** 2, m.nu_v ** 2 / (2 * m.kappa_v)
    vs = np.linspace(0.0, stats.gamma.ppf(0.99, shape, scale=scale), n_v)
    return {"mu": mus, "v": vs}, {"p": np.ones((n_mu, n_v))}


def ahlm_diffusion_grid(n_y: int, n_a: int, m=None):
    """examples/consumption_saving/consumption_saving_diffusion_income.jl:45-58."""
    from scipy import stats
    from .models import AchdouHanLasryLionsMollModel_Diffusion
    m = m or AchdouHanLasryLionsMollModel_Diffusion()
    shape, scale = 2 * m.kappa_y * m.ybar / m.sigma_y ** 2, m.sigma_y ** 2 / (2 * m.kappa_y)
    ys = np.linspace(stats.gamma.ppf(0.001, shape, scale=scale), stats.gamma.ppf(0.999, shape, scale=scale), n_y)
    as_ = np.linspace(m.amin, m.amax, n_a)
    Y, A = np.meshgrid(ys, as_, indexing="ij")
    guess = (m.rho / m.gamma + (1 - 1 / m.gamma) * m.r) ** (-m.gamma) * (A + Y / m.r) ** (1 - m.gamma) / (1 - m.gamma)
    return {"y": ys, "a": as_}, {"v": guess}


def hjb3d_grid(n: int):
    """Config C5: [0,1]^3, n uniform points per axis, guess 0."""
    g = np.linspace(0.0, 1.0, n)
    return {"x": g, "y": g.copy(), "z": g.copy()}, {"v": np.zeros((n, n, n))}

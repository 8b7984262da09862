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


def campbell_cochrane_sweep(first: int, count: int, sn: int = 1000):
    """Problems first..first+count-1 of the C4 calibration sweep: (params [B,6], grids [B,S],
    guess [B,S]).  Parameters are i.i.d. uniform draws around the reference defaults
    (examples/asset_pricing/campbell_cochrane.jl:28-35); each problem has its own state grid
    (initialize_stategrid, :40-48) and the example's guess p = 1 (:50)."""
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
    t = np.linspace(0.0, 1.0, nh)[None, :]
    with np.errstate(divide="ignore"):
        shigh = np.log(t * np.exp(smax)[:, None])
    tl = np.linspace(0.0, 1.0, sn - nh)[None, :]
    slow = -300.0 + tl * (shigh[:, 1:2] + 300.0)
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

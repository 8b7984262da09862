"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see econpdes_oracle.py).

Vectorised numpy restatements of the model functors named by the configs, each citing the
reference body it follows, plus the grids/guesses the examples build.  Expression order is
kept exactly as written in the reference (Julia evaluates left to right and never
contracts a*b+c into an fma; ``x^2`` with a literal exponent is ``x*x``).
"""
from __future__ import annotations

import math
from fractions import Fraction
from typing import Dict, Tuple

import numpy as np


def julia_range(a: float, b: float, n: int) -> np.ndarray:
    """``range(a, b, length=n)``: Julia evaluates a + (i-1)*step in twice precision, i.e.
    (to the last bit in all but pathological cases) the correctly rounded exact value."""
    fa, fb = Fraction(float(a)), Fraction(float(b))
    step = (fb - fa) / (n - 1)
    out = np.array([float(fa + i * step) for i in range(n)], dtype=np.float64)
    out[-1] = float(b)
    return out


# --------------------------------------------------------------------------------------
# C1  README quickstart (README.md:48-67)
# --------------------------------------------------------------------------------------
def readme_linear(rho=0.05, kappa=0.2):
    def pde(state, u):
        x = state["x"]
        mu = -kappa * (x - 0.5)
        vt = -(x + mu * np.where(mu >= 0, u["vx_up"], u["vx_down"]) - rho * u["v"])
        return {"vt": vt}
    return pde


def readme_linear_problem(n=50):
    grid = {"x": julia_range(0.0, 1.0, n)}
    guess = {"v": np.zeros(n)}
    return grid, guess


# --------------------------------------------------------------------------------------
# C4  Campbell-Cochrane habit model (examples/asset_pricing/campbell_cochrane.jl:28-84)
# --------------------------------------------------------------------------------------
CC_DEFAULT = dict(mu=0.0189, sigma=0.015, gamma=2.0, rho=0.116, kappa_s=0.138, b=0.0)


def campbell_cochrane_grid(mu, sigma, gamma, rho, kappa_s, b, sn=1000) -> np.ndarray:
    """initialize_stategrid, campbell_cochrane.jl:40-48."""
    Sbar = sigma * math.sqrt(gamma / (kappa_s - b / gamma))
    sbar = math.log(Sbar)
    smax = sbar + 0.5 * (1 - Sbar * Sbar)
    nh = sn // 10
    with np.errstate(divide="ignore"):
        shigh = np.log(julia_range(0.0, math.exp(smax), nh))
    slow = julia_range(-300.0, shigh[1], sn - nh)
    return np.concatenate([slow[:-1], shigh[1:]])


def campbell_cochrane(mu, sigma, gamma, rho, kappa_s, b):
    def pde(state, u):
        s = state["s"]
        p, ps_up, ps_down, pss = u["p"], u["ps_up"], u["ps_down"], u["pss"]
        Sbar = sigma * math.sqrt(gamma / (kappa_s - b / gamma))
        sbar = math.log(Sbar)
        lam = 1 / Sbar * np.sqrt(1 - 2 * (s - sbar)) - 1
        mus = -kappa_s * (s - sbar)
        sigs = lam * sigma
        mus_eff = mus + sigs * (sigma - gamma * (sigma + sigs))
        ps = np.where(mus_eff >= 0, ps_up, ps_down)
        sigp = ps / p * sigs
        mup = ps / p * mus + 0.5 * pss / p * (sigs * sigs)
        kap = gamma * (sigma + sigs)
        r = rho + gamma * mu - (gamma * kappa_s - b) / 2 + b * (sbar - s)
        pt = -p * (1 / p + mu + mup + sigp * sigma - r - kap * (sigma + sigp))
        return {"pt": pt}
    return pde


def campbell_cochrane_problem(params=None, sn=1000):
    prm = dict(CC_DEFAULT if params is None else params)
    s = campbell_cochrane_grid(sn=sn, **prm)
    return {"s": s}, {"p": np.ones(s.shape[0])}, prm


def splitmix64_uniforms(seed: int, count: int) -> np.ndarray:
    """language-independent parameter draws for the C4 sweep."""
    M = (1 << 64) - 1
    state = seed & M
    out = np.empty(count)
    for i in range(count):
        state = (state + 0x9E3779B97F4A7C15) & M
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z = z ^ (z >> 31)
        out[i] = (z >> 11) * 2.0 ** -53
    return out


def campbell_cochrane_sweep_params(index: int) -> Dict[str, float]:
    """Problem ``index`` of the batched calibration sweep (config C4)."""
    u = splitmix64_uniforms(0xEC0DE5 + index, 5)
    return dict(mu=0.015 + (0.023 - 0.015) * u[0], sigma=0.012 + (0.018 - 0.012) * u[1],
                gamma=1.5 + (2.5 - 1.5) * u[2], rho=0.09 + (0.14 - 0.09) * u[3],
                kappa_s=0.11 + (0.17 - 0.11) * u[4], b=0.0)


# --------------------------------------------------------------------------------------
# C3  Bansal-Yaron long-run risk (examples/asset_pricing/bansal_yaron.jl:22-95)
# --------------------------------------------------------------------------------------
BY_DEFAULT = dict(mubar=0.018, vbar=0.00073, kappa_mu=0.252, nu_mu=0.528, kappa_v=0.156,
                  nu_v=0.00354, rho=0.024, gamma=7.5, psi=1.5)


def bansal_yaron(mubar, vbar, kappa_mu, nu_mu, kappa_v, nu_v, rho, gamma, psi):
    def pde(state, u):
        mu, v = state["mu"], state["v"]
        p = u["p"]
        muc = mu
        sigc = np.sqrt(v)
        mumu = kappa_mu * (mubar - mu)
        sigmu = nu_mu * np.sqrt(v)
        muv = kappa_v * (vbar - v)
        sigv = nu_v * np.sqrt(v)
        pmu = np.where(mumu >= 0, u["pmu_up"], u["pmu_down"])
        pv = np.where(muv >= 0, u["pv_up"], u["pv_down"])
        sigp_Zmu = pmu / p * sigmu
        sigp_Zv = pv / p * sigv
        sigp2 = sigp_Zmu * sigp_Zmu + sigp_Zv * sigp_Zv
        mup = (pmu / p * mumu + pv / p * muv + 0.5 * u["pmumu"] / p * (sigmu * sigmu)
               + 0.5 * u["pvv"] / p * (sigv * sigv))
        kap_Zc = gamma * sigc
        kap_Zmu = -(1 - gamma * psi) / (psi - 1) * sigp_Zmu
        kap_Zv = -(1 - gamma * psi) / (psi - 1) * sigp_Zv
        r = rho + muc / psi - (1 + 1 / psi) / 2 * gamma * (sigc * sigc) - (gamma * psi - 1) / (2 * (psi - 1)) * sigp2
        pt = -p * (1 / p + muc + mup - r - kap_Zc * sigc - kap_Zmu * sigp_Zmu - kap_Zv * sigp_Zv)
        return {"pt": pt}
    return pde


def bansal_yaron_problem(n_mu=30, n_v=30, params=None):
    from scipy import stats
    prm = dict(BY_DEFAULT if params is None else params)
    sd = math.sqrt(prm["nu_mu"] ** 2 * prm["vbar"] / (2 * prm["kappa_mu"]))
    mus = julia_range(stats.norm.ppf(0.01, prm["mubar"], sd), stats.norm.ppf(0.99, prm["mubar"], sd), n_mu)
    shape = 2 * prm["kappa_v"] * prm["vbar"] / prm["nu_v"] ** 2
    scale = prm["nu_v"] ** 2 / (2 * prm["kappa_v"])
    vs = julia_range(0.0, stats.gamma.ppf(0.99, shape, scale=scale), n_v)
    return {"mu": mus, "v": vs}, {"p": np.ones((n_mu, n_v))}, prm


# --------------------------------------------------------------------------------------
# C2  consumption-saving with diffusive income
#     (examples/consumption_saving/consumption_saving_diffusion_income.jl:25-106)
# --------------------------------------------------------------------------------------
AHLM_DEFAULT = dict(kappa_y=0.1, ybar=1.0, sigma_y=0.07, r=0.03, rho=0.05, gamma=2.0, amin=0.0, amax=100.0)


def ahlm_diffusion(kappa_y, ybar, sigma_y, r, rho, gamma, amin, amax):
    def pde(state, u):
        y, a = state["y"], state["a"]
        v, vy_up, vy_down, va_up, va_down, vyy = u["v"], u["vy_up"], u["vy_down"], u["va_up"], u["va_down"], u["vyy"]
        muy = kappa_y * (ybar - y)
        cmax = 100.0 * (y + r * np.maximum(a, 0.0))
        vy = np.where(muy >= 0, vy_up, vy_down)
        c_up = np.where(va_up > 0, np.minimum(va_up ** (-1 / gamma), cmax), cmax)
        mua_up = y + r * a - c_up
        c_down = np.where(va_down > 0, np.minimum(va_down ** (-1 / gamma), cmax), cmax)
        mua_down = y + r * a - c_down
        case1 = mua_up >= 0.0
        case2 = (~case1) & (mua_down <= 0.0) & (a > amin)
        c3 = y + r * a
        va3 = c3 ** (-gamma)
        c = np.where(case1, c_up, np.where(case2, c_down, c3))
        va = np.where(case1, va_up, np.where(case2, va_down, va3))
        mua = np.where(case1, mua_up, np.where(case2, mua_down, 0.0))
        vt = -(c ** (1 - gamma) / (1 - gamma) + mua * va + muy * vy + 0.5 * vyy * (sigma_y * sigma_y) - rho * v)
        return {"vt": vt}, {"mua": mua, "c": c}
    return pde


def ahlm_diffusion_problem(n_y=10, n_a=1000, params=None):
    from scipy import stats
    prm = dict(AHLM_DEFAULT if params is None else params)
    shape = 2 * prm["kappa_y"] * prm["ybar"] / prm["sigma_y"] ** 2
    scale = prm["sigma_y"] ** 2 / (2 * prm["kappa_y"])
    ys = julia_range(stats.gamma.ppf(0.001, shape, scale=scale), stats.gamma.ppf(0.999, shape, scale=scale), n_y)
    as_ = julia_range(prm["amin"], prm["amax"], n_a)
    g, rr, rho = prm["gamma"], prm["r"], prm["rho"]
    Y, A = np.meshgrid(ys, as_, indexing="ij")
    guess = (rho / g + (1 - 1 / g) * rr) ** (-g) * (A + Y / rr) ** (1 - g) / (1 - g)
    return {"y": ys, "a": as_}, {"v": guess}, prm


# --------------------------------------------------------------------------------------
# C5  synthetic 3-state HJB (this repo's own definition, shared
#     verbatim with the CUDA functor in econpdes.jl_b200/csrc/models.cuh)
# --------------------------------------------------------------------------------------
HJB3D_DEFAULT = dict(kx=0.2, ky=0.3, kz=0.4, sx=0.1, sy=0.1, sz=0.1, corr=0.5, rho=0.05, theta=1e-3)


def hjb3d_synth(kx, ky, kz, sx, sy, sz, corr, rho, theta):
    def pde(state, u):
        x, y, z = state["x"], state["y"], state["z"]
        v = u["v"]
        mx = -kx * (x - 0.5)
        my = -ky * (y - 0.5)
        mz = -kz * (z - 0.5)
        vx = np.where(mx >= 0, u["vx_up"], u["vx_down"])
        vy = np.where(my >= 0, u["vy_up"], u["vy_down"])
        vz = np.where(mz >= 0, u["vz_up"], u["vz_down"])
        drift = mx * vx + my * vy + mz * vz
        diff = 0.5 * (sx * sx) * u["vxx"] + 0.5 * (sy * sy) * u["vyy"] + 0.5 * (sz * sz) * u["vzz"]
        cross = corr * sx * sy * u["vxy_up"]
        vt = -(x + y + z + drift + diff + cross - rho * v - theta * (v * v))
        return {"vt": vt}
    return pde


HJB4D_DEFAULT = dict(kx=0.2, ky=0.3, kz=0.4, kw=0.25, sx=0.1, sy=0.12, sz=0.08, sw=0.1, cxw=0.5, cyz=-0.3, cxy=0.2,
                     rho=0.05, theta=1e-3)


def hjb4d_synth(kx, ky, kz, kw, sx, sy, sz, sw, cxw, cyz, cxy, rho, theta):
    """Synthetic 4-state HJB (csrc/models.cuh: Hjb4dSynth): the stencil work behind test/runtests.jl:294-320."""
    def pde(state, u):
        x, y, z, w = state["x"], state["y"], state["z"], state["w"]
        v = u["v"]
        drift = diff = None
        for nm, st, k, s in (("x", x, kx, sx), ("y", y, ky, sy), ("z", z, kz, sz), ("w", w, kw, sw)):
            m = -k * (st - 0.5)
            vd = np.where(m >= 0, u["v" + nm + "_up"], u["v" + nm + "_down"])
            drift = m * vd if drift is None else drift + m * vd
            dd = 0.5 * (s * s) * u["v" + nm + nm]
            diff = dd if diff is None else diff + dd
        cross = cxw * sx * sw * u["vxw_up"] + cyz * sy * sz * u["vyz_down"] + cxy * sx * sy * u["vxy"]
        vt = -(x + y + z + w + drift + diff + cross - rho * v - theta * (v * v))
        return {"vt": vt}
    return pde


def hjb4d_problem(n=(5, 4, 6, 3), params=None, warp=True):
    prm = dict(HJB4D_DEFAULT if params is None else params)
    g = {}
    for nm, k in zip("xyzw", n):
        a = julia_range(0.0, 1.0, k)
        g[nm] = a ** 1.3 if warp else a     # non-uniform spacings
    return g, {"v": np.zeros(tuple(n))}, prm


def hjb3d_problem(n=16, params=None):
    prm = dict(HJB3D_DEFAULT if params is None else params)
    g = julia_range(0.0, 1.0, n)
    return {"x": g, "y": g.copy(), "z": g.copy()}, {"v": np.zeros((n, n, n))}, prm


# --------------------------------------------------------------------------------------
# Neoclassical growth: example version (examples/neoclassical_growth.jl:25-84; the 11
# outer iterations of details.tex:173 pin the FiniteDiff epsilon rule) and the test-suite
# version (test/runtests.jl:17-44).
# --------------------------------------------------------------------------------------
def growth_example_params():
    gamma = 4.0
    delta, alpha, A = 0.05, 0.3, 0.5
    return dict(A=A, alpha=alpha, delta=delta, rho=delta * (alpha * gamma - 1), gamma=gamma)


def neoclassical_growth(A, alpha, delta, rho, gamma):
    def pde(state, u):
        k = state["k"]
        v, vk_up, vk_down = u["v"], u["vk_up"], u["vk_down"]
        cap = 10 * A * k ** alpha
        c_up = np.where(vk_up >= 0, np.minimum(vk_up ** (-1 / gamma), cap), cap)
        muk_up = A * k ** alpha - delta * k - c_up
        c_down = np.where(vk_down >= 0, np.minimum(vk_down ** (-1 / gamma), cap), cap)
        muk_down = A * k ** alpha - delta * k - c_down
        case1 = muk_up > 0
        case2 = (~case1) & (muk_down < 0)
        c3 = A * k ** alpha - delta * k
        c = np.where(case1, c_up, np.where(case2, c_down, c3))
        vk = np.where(case1, vk_up, np.where(case2, vk_down, c3 ** (-gamma)))
        muk = np.where(case1, muk_up, np.where(case2, muk_down, 0.0))
        vt = -(c ** (1 - gamma) / (1 - gamma) + muk * vk - rho * v)
        return {"vt": vt}, {"c": c, "muk": muk}
    return pde


def neoclassical_growth_problem(n=1000, params=None):
    prm = growth_example_params() if params is None else dict(params)
    A, alpha, delta, rho, gamma = (prm[k] for k in ("A", "alpha", "delta", "rho", "gamma"))
    kbar = (alpha * A / (rho + delta)) ** (1 / (1 - alpha))
    ks = julia_range(0.1 * kbar, 5 * kbar, n)
    guess = (A * ks ** alpha) ** (1 - gamma) / (1 - gamma) / rho
    return {"k": ks}, {"v": guess}, prm


GROWTH_TEST = dict(A=0.5, alpha=0.3, delta=0.05, rho=0.05, gamma=2.0)


def growth_test_hjb(A=0.5, alpha=0.3, delta=0.05, rho=0.05, gamma=2.0):
    """test/runtests.jl:28-44 (note: differs from the example in its branch structure)."""
    def pde(state, u):
        k = state["k"]
        cmax = 10 * A * k ** alpha
        c_up = np.where(u["vk_up"] > 0, np.minimum(u["vk_up"] ** (-1 / gamma), cmax), cmax)
        mu_up = A * k ** alpha - delta * k - c_up
        c_down = np.where(u["vk_down"] > 0, np.minimum(u["vk_down"] ** (-1 / gamma), cmax), cmax)
        mu_down = A * k ** alpha - delta * k - c_down
        case1 = mu_up > 0
        case2 = (~case1) & (mu_down < 0)
        c3 = A * k ** alpha - delta * k
        c = np.where(case1, c_up, np.where(case2, c_down, c3))
        vk = np.where(case1, u["vk_up"], np.where(case2, u["vk_down"], c3 ** (-gamma)))
        muk = np.where(case1, mu_up, np.where(case2, mu_down, 0.0))
        vt = -(c ** (1 - gamma) / (1 - gamma) + muk * vk - rho * u["v"])
        return {"vt": vt}, {"c": c, "muk": muk}
    return pde


def growth_test_problem(n=200):
    A, alpha, delta, rho, gamma = (GROWTH_TEST[k] for k in ("A", "alpha", "delta", "rho", "gamma"))
    kbar = (alpha * A / (rho + delta)) ** (1 / (1 - alpha))
    ks = julia_range(0.1 * kbar, 5.0 * kbar, n)
    guess = (A * ks ** alpha) ** (1 - gamma) / (1 - gamma) / rho
    return {"k": ks}, {"v": guess}, kbar


# --------------------------------------------------------------------------------------
# Closed-form test models (test/runtests.jl:276-281, 314-319, 379-389, 225-232)
# --------------------------------------------------------------------------------------
def linear_sum(statenames):
    """vt = -(v - x - z - ...)  -> v = sum of the coordinates."""
    def pde(state, u):
        acc = u["v"]
        for s in statenames:
            acc = acc - state[s]
        return {"vt": -acc}
    return pde


def forward_price():
    """test/runtests.jl:379-384: v = 2x, p = x."""
    def pde(state, u):
        vt = -(state["x"] + u["p"] - u["v"])
        pt = u["v"] / 2 - u["p"]
        return {"vt": vt, "pt": pt}
    return pde


HJB2D_REGIMES_DEFAULT = dict(kx1=0.2, kx2=0.35, ky=0.3, sx=0.1, sy=0.12, corr=0.3, rho=0.05, lam12=0.4, lam21=0.25,
                             theta=1e-3)


def hjb2d_regimes(kx1, kx2, ky, sx, sy, corr, rho, lam12, lam21, theta):
    """Two coupled HJB equations on (x, y) (regime switching); same definition as the device
    functor Hjb2dRegimes (csrc/models.cuh).  Unknowns v, w."""
    def pde(state, u):
        x, y = state["x"], state["y"]
        v, w = u["v"], u["w"]
        mx1 = -kx1 * (x - 0.5)
        mx2 = -kx2 * (x - 0.3)
        my = -ky * (y - 0.5)
        vx = np.where(mx1 >= 0, u["vx_up"], u["vx_down"])
        vy = np.where(my >= 0, u["vy_up"], u["vy_down"])
        wx = np.where(mx2 >= 0, u["wx_up"], u["wx_down"])
        wy = np.where(my >= 0, u["wy_up"], u["wy_down"])
        hx, hy, cxy = 0.5 * (sx * sx), 0.5 * (sy * sy), corr * sx * sy
        vt = -(x + y + mx1 * vx + my * vy + hx * u["vxx"] + hy * u["vyy"] + cxy * u["vxy"] - rho * v
               + lam12 * (w - v) - theta * (v * v))
        wt = -(2 * x - y + 1 + mx2 * wx + my * wy + hx * u["wxx"] + hy * u["wyy"] + cxy * u["wxy"] - rho * w
               + lam21 * (v - w) - theta * (w * w))
        return {"vt": vt, "wt": wt}
    return pde


def hjb2d_regimes_problem(nx=24, ny=20, params=None):
    prm = dict(HJB2D_REGIMES_DEFAULT if params is None else params)
    return ({"x": julia_range(0.0, 1.0, nx), "y": julia_range(0.0, 1.0, ny)},
            {"v": np.zeros((nx, ny)), "w": np.zeros((nx, ny))}, prm)


def tuckman_vila(c=0.06, r=0.09, rho=5.42, sigma=26.72, a=26.72, T=100.0):
    """examples/asset_pricing/tuckman_vila.jl:67-86: three-argument form pde(state, u, tau)."""
    def pde(state, u, tau):
        z = state["z"]
        F, Fz_up, Fz_down, Fzz = u["F"], u["Fz_up"], u["Fz_down"], u["Fzz"]
        Fz = np.where(z >= 0, Fz_up, Fz_down)
        with np.errstate(all="ignore"):
            phi = z * (1 + z ** 2) ** (-1 / 2)
            phiz = (1 + z ** 2) ** (-3 / 2)
            phizz = -3 * z * (1 + z ** 2) ** (-5 / 2)
            stau = c / r * (1 - np.exp(-r * (T - tau)))
            muL = rho * z - 0.5 * sigma ** 2 * phizz / phiz + np.float64(c) / stau * (phi - 1) / phiz
            iL = 1 / a * (muL / sigma ** 2 + Fz)
            muS = rho * z - 0.5 * sigma ** 2 * phizz / phiz + np.float64(c) / stau * (phi + 1) / phiz
            iS = 1 / a * (muS / sigma ** 2 + Fz)
            i = np.where(iL > 0, iL, np.where(iS < 0, iS, 0.0))
            mu = np.where(iL > 0, muL, np.where(iS < 0, muS, 0.0))
            Ft = -((mu + sigma ** 2 * Fz) * a * i - 0.5 * sigma ** 2 * (a * i) ** 2 - rho * z * Fz
                   + 0.5 * sigma ** 2 * (Fzz - Fz ** 2))
        return {"Ft": Ft}
    return pde


def tuckman_vila_problem(zn=200, nt=10, **prm):
    """Grid, terminal value and time grid of the example (tuckman_vila.jl:39-58)."""
    from scipy.stats import norm
    p = dict(c=0.06, r=0.09, rho=5.42, sigma=26.72, a=26.72, T=100.0)
    p.update(prm)
    sd = np.sqrt(p["sigma"] ** 2 / (2 * p["rho"]))
    z = julia_range(norm.ppf(0.00001, 0, sd), norm.ppf(0.99999, 0, sd), zn)
    return {"z": z}, {"F": np.zeros(zn)}, julia_range(0.0, p["T"], nt), p


DI_TELLA_DEFAULT = dict(gamma=5.0, psi=2.0, rho=0.0665, tau=1.15, A=54.0, delta=0.05,
                        B=(0.20 - 54.0 * (0.02 + 0.05) ** 2) / (0.02 + 0.05), sigma=0.0125, phi=0.2, nubar=0.25,
                        kappanu=1.38, signubar=-0.17)


def di_tella(gamma, psi, rho, tau, A, delta, B, sigma, phi, nubar, kappanu, signubar):
    """examples/asset_pricing/di_tella.jl:92-146 (states x, nu; unknowns pA, pB, p; p algebraic).
    The two-pass upwinding in x (`@goto start`) is evaluated for both passes and selected pointwise."""
    def pde(state, u):
        x, nu = state["x"], state["nu"]
        pA, pB, p = u["pA"], u["pB"], u["p"]
        with np.errstate(all="ignore"):
            q = (p - B) / (2 * A)
            g = q - delta
            i = A * q ** 2 + B * q
            munu = kappanu * (nubar - nu)
            signu = signubar * np.sqrt(nu)
            up = munu >= 0
            pAnu = np.where(up, u["pAnu_up"], u["pAnu_down"])
            pBnu = np.where(up, u["pBnu_up"], u["pBnu_down"])
            pnu = np.where(up, u["pnu_up"], u["pnu_down"])

            def block(pAx, pBx, px):
                sigX = x * (1 - x) * (1 - gamma) / (gamma * (psi - 1)) * (pAnu / pA - pBnu / pB) * signu / (
                    1 - x * (1 - x) * (1 - gamma) / (gamma * (psi - 1)) * (pAx / pA - pBx / pB))
                sigpA = pAx / pA * sigX + pAnu / pA * signu
                sigpB = pBx / pB * sigX + pBnu / pB * signu
                sigp = px / p * sigX + pnu / p * signu
                kappa = (sigp + sigma - (1 - gamma) / (gamma * (psi - 1)) * (x * sigpA + (1 - x) * sigpB)) / (1 / gamma)
                kidio = gamma * phi * nu / x
                sigA = kappa / gamma + (1 - gamma) / (gamma * (psi - 1)) * sigpA
                nuA = kidio / gamma
                sigB = kappa / gamma + (1 - gamma) / (gamma * (psi - 1)) * sigpB
                muX = x * (1 - x) * ((sigA * kappa + nuA * kidio - 1 / pA - tau)
                                     - (sigB * kappa - 1 / pB + tau * x / (1 - x)) - (sigA - sigB) * (sigma + sigp))
                return sigX, sigpA, sigpB, sigp, kappa, kidio, sigA, nuA, sigB, muX
            first = block(u["pAx_up"], u["pBx_up"], u["px_up"])
            second = block(u["pAx_down"], u["pBx_down"], u["px_down"])
            redo = first[-1] <= 0
            sigX, sigpA, sigpB, sigp, kappa, kidio, sigA, nuA, sigB, muX = (np.where(redo, b2, b1)
                                                                              for b1, b2 in zip(first, second))
            pAx = np.where(redo, u["pAx_down"], u["pAx_up"])
            pBx = np.where(redo, u["pBx_down"], u["pBx_up"])
            px = np.where(redo, u["px_down"], u["px_up"])
            cup = sigX * signu >= 0
            pAxnu = np.where(cup, u["pAxnu_up"], u["pAxnu_down"])
            pBxnu = np.where(cup, u["pBxnu_up"], u["pBxnu_down"])
            pxnu = np.where(cup, u["pxnu_up"], u["pxnu_down"])
            mupA = (pAx / pA * muX + pAnu / pA * munu + 0.5 * u["pAxx"] / pA * sigX ** 2 + 0.5 * u["pAnunu"] / pA * signu ** 2
                    + pAxnu / pA * sigX * signu)
            mupB = (pBx / pB * muX + pBnu / pB * munu + 0.5 * u["pBxx"] / pB * sigX ** 2 + 0.5 * u["pBnunu"] / pB * signu ** 2
                    + pBxnu / pB * sigX * signu)
            mup = (px / p * muX + pnu / p * munu + 0.5 * u["pxx"] / p * sigX ** 2 + 0.5 * u["pnunu"] / p * signu ** 2
                   + pxnu / p * sigX * signu)
            r = (1 - i) / p + g + mup + sigma * sigp - kappa * (sigma + sigp) - gamma / x * (phi * nu) ** 2
            pAt = -pA * (1 / pA + (psi - 1) * tau / (1 - gamma) * ((pA / pB) ** ((1 - gamma) / (1 - psi)) - 1) - psi * rho
                         + (psi - 1) * (r + kappa * sigA + kidio * nuA) + mupA - (psi - 1) * gamma / 2 * (sigA ** 2 + nuA ** 2)
                         + (2 - psi - gamma) / (2 * (psi - 1)) * sigpA ** 2 + (1 - gamma) * sigpA * sigA)
            pBt = -pB * (1 / pB - psi * rho + (psi - 1) * (r + kappa * sigB) + mupB - (psi - 1) * gamma / 2 * sigB ** 2
                         + (2 - psi - gamma) / (2 * (psi - 1)) * sigpB ** 2 + (1 - gamma) * sigpB * sigB)
            pt = -p * ((1 - i) / p - x / pA - (1 - x) / pB)
            Q = (pA / pB) ** (1 / (psi - 1))
        return ({"pAt": pAt, "pBt": pBt, "pt": pt},
                {"sigmax": sigX, "Q": Q, "r": r, "kappa": kappa, "sigmap": sigp, "growth": g, "investment": i})
    return pde


def di_tella_problem(xn=30, nun=30, params=None):
    """Grid and guess of the example (di_tella.jl:52-76)."""
    prm = dict(DI_TELLA_DEFAULT if params is None else params)
    xs = np.array(sorted(set(julia_range(0.01, 0.99, xn).tolist() + [0.05, 0.10, 0.20])))
    nus = np.array(sorted(set(julia_range(0.01, 1.00, nun).tolist() + [0.10, 0.25, 0.60])))
    shape = (len(xs), len(nus))
    return {"x": xs, "nu": nus}, {"pA": np.ones(shape), "pB": np.ones(shape), "p": np.ones(shape)}, prm


def ahlm_two_states(yl=0.5, yh=1.5, llh=0.2, lhl=0.2, r=0.03, rho=0.04, gamma=2.0, amin=None, amax=50.0):
    """examples/consumption_saving/consumption_saving_two_income.jl:68-112 (unknowns vl, vh on the asset grid)."""
    if amin is None:
        amin = -yl / r + 0.001

    def one(a, va_up, va_down, y):
        cmax = 100.0 * (y + r * np.maximum(a, 0.0))
        with np.errstate(all="ignore"):
            c_up = np.where(va_up > 0, np.minimum(np.abs(va_up) ** (-1 / gamma), cmax), cmax)
            mu_up = y + r * a - c_up
            c_down = np.where(va_down > 0, np.minimum(np.abs(va_down) ** (-1 / gamma), cmax), cmax)
            mu_down = y + r * a - c_down
            use_up = mu_up >= 0.0
            use_down = (~use_up) & (mu_down <= 0.0) & (a > amin)
            c3 = y + r * a
            c = np.where(use_up, c_up, np.where(use_down, c_down, c3))
            mu = np.where(use_up, mu_up, np.where(use_down, mu_down, 0.0))
            va = np.where(use_up, va_up, np.where(use_down, va_down, np.abs(c3) ** (-gamma)))
        return va, c, mu

    def pde(state, u):
        a = state["a"]
        vl, vh = u["vl"], u["vh"]
        vla, cl, mula = one(a, u["vla_up"], u["vla_down"], yl)
        vha, ch, muha = one(a, u["vha_up"], u["vha_down"], yh)
        with np.errstate(all="ignore"):
            vlt = -(cl ** (1 - gamma) / (1 - gamma) + mula * vla + llh * (vh - vl) - rho * vl)
            vht = -(ch ** (1 - gamma) / (1 - gamma) + muha * vha + lhl * (vl - vh) - rho * vh)
        return {"vlt": vlt, "vht": vht}, {"cl": cl, "ch": ch, "mula": mula, "muha": muha}
    return pde


def ahlm_two_states_problem(n=200, **over):
    """Grid and guess of the example (consumption_saving_two_income.jl:41-54)."""
    prm = dict(yl=0.5, yh=1.5, llh=0.2, lhl=0.2, r=0.03, rho=0.04, gamma=2.0, amax=50.0)
    prm.update(over)
    prm.setdefault("amin", -prm["yl"] / prm["r"] + 0.001)
    yl, yh, r, rho, gamma, amin, amax = (prm[k] for k in ("yl", "yh", "r", "rho", "gamma", "amin", "amax"))
    a = amin + julia_range(0.0, (amax - amin) ** (1 / 2), n) ** 2
    vl = (rho / gamma + (1 - 1 / gamma) * r) ** (-gamma) * (a + yl / r) ** (1 - gamma) / (1 - gamma)
    vh = (rho / gamma + (1 - gamma) * r) ** (-gamma) * (a + yh / r) ** (1 - gamma) / (1 - gamma)
    return {"a": a}, {"vl": vl, "vh": vh}, prm


def garleanu_panageas_params(gA=1.5, psA=0.7, gB=10.0, psB=0.05, rho=0.001, delta=0.02, nuA=0.01, mu=0.02, sigma=0.041,
                             B1=30.72, d1=0.0525, B2=-30.29, d2=0.0611, omega=0.92):
    """Struct of examples/asset_pricing/garleanu_panageas.jl:20-40; the constructor normalises B1, B2."""
    scale = delta / (delta + d1) * B1 + delta / (delta + d2) * B2
    return dict(gA=gA, psA=psA, gB=gB, psB=psB, rho=rho, delta=delta, nuA=nuA, mu=mu, sigma=sigma, B1=B1 / scale, d1=d1,
                B2=B2 / scale, d2=d2, omega=omega)


def garleanu_panageas(gA, psA, gB, psB, rho, delta, nuA, mu, sigma, B1, d1, B2, d2, omega):
    """examples/asset_pricing/garleanu_panageas.jl:74-117 (unknowns pA, pB, phi1, phi2 on x in [0, 1])."""
    def pde(state, u):
        x = state["x"]
        pA, pB, f1, f2 = u["pA"], u["pB"], u["phi1"], u["phi2"]

        def block(pAx, pBx, f1x, f2x):
            G = 1 / (x / gA + (1 - x) / gB)
            sigx = sigma * x * (G / gA - 1) / (1 + G * x * (1 - x) / (gA * gB) * (
                (1 - gB * psB) / (psB - 1) * (pBx / pB) - (1 - gA * psA) / (psA - 1) * (pAx / pA)))
            spA = pAx / pA * sigx
            spB = pBx / pB * sigx
            sf1 = f1x / f1 * sigx
            sf2 = f2x / f2 * sigx
            kappa = G * (sigma - x * (1 - gA * psA) / (gA * (psA - 1)) * spA - (1 - x) * (1 - gB * psB) / (gB * (psB - 1)) * spB)
            sCA = kappa / gA + (1 - gA * psA) / (gA * (psA - 1)) * spA
            sCB = kappa / gB + (1 - gB * psB) / (gB * (psB - 1)) * spB
            mcA = kappa ** 2 * (1 + psA) / (2 * gA) + (1 - psA * gA) / (gA * (psA - 1)) * kappa * spA - (1 - gA * psA) / (
                2 * gA * (psA - 1)) * spA ** 2
            mcB = kappa ** 2 * (1 + psB) / (2 * gB) + (1 - psB * gB) / (gB * (psB - 1)) * kappa * spB - (1 - gB * psB) / (
                2 * gB * (psB - 1)) * spB ** 2
            r = rho + 1 / (psA * x + psB * (1 - x)) * (mu - x * mcA - (1 - x) * mcB - delta * (
                (nuA / pA + (1 - nuA) / pB) * omega * (B1 * f1 + B2 * f2) - 1))
            muCA = psA * (r - rho) + mcA
            muCB = psB * (r - rho) + mcB
            mux = x * (muCA - mu) + delta * (nuA / pA * omega * (B1 * f1 + B2 * f2) - x) - sigma * sigx
            return sigx, spA, spB, sf1, sf2, kappa, sCA, sCB, r, muCA, muCB, mux
        with np.errstate(all="ignore"):
            first = block(u["pAx_up"], u["pBx_up"], u["phi1x_up"], u["phi2x_up"])
            second = block(u["pAx_down"], u["pBx_down"], u["phi1x_down"], u["phi2x_down"])
            redo = first[-1] <= 0
            sigx, spA, spB, sf1, sf2, kappa, sCA, sCB, r, muCA, muCB, mux = (np.where(redo, b2, b1)
                                                                              for b1, b2 in zip(first, second))
            pAx = np.where(redo, u["pAx_down"], u["pAx_up"])
            pBx = np.where(redo, u["pBx_down"], u["pBx_up"])
            f1x = np.where(redo, u["phi1x_down"], u["phi1x_up"])
            f2x = np.where(redo, u["phi2x_down"], u["phi2x_up"])
            mupA = pAx / pA * mux + 0.5 * u["pAxx"] / pA * sigx ** 2
            mupB = pBx / pB * mux + 0.5 * u["pBxx"] / pB * sigx ** 2
            muf1 = f1x / f1 * mux + 0.5 * u["phi1xx"] / f1 * sigx ** 2
            muf2 = f2x / f2 * mux + 0.5 * u["phi2xx"] / f2 * sigx ** 2
            pAt = -pA * (1 / pA + muCA + mupA + sCA * spA - r - delta - kappa * (spA + sCA))
            pBt = -pB * (1 / pB + muCB + mupB + sCB * spB - r - delta - kappa * (spB + sCB))
            f1t = -f1 * (1 / f1 + mu - delta - d1 + muf1 + sigma * sf1 - r - kappa * (sf1 + sigma))
            f2t = -f2 * (1 / f2 + mu - delta - d2 + muf2 + sigma * sf2 - r - kappa * (sf2 + sigma))
        return {"pAt": pAt, "pBt": pBt, "phi1t": f1t, "phi2t": f2t}, {"r": r, "kappa": kappa}
    return pde


def garleanu_panageas_problem(n=100, **over):
    prm = garleanu_panageas_params(**over)
    x = julia_range(0.0, 1.0, n)
    return {"x": x}, {k: np.ones(n) for k in ("pA", "pB", "phi1", "phi2")}, prm


def garleanu_panageas_lrr_params(gA=1.5, psA=0.7, gB=10.0, psB=0.05, rho=0.001, delta=0.02, nuA=0.01, mubar=0.018,
                                 vbar=0.00073, kmu=0.252, numu=0.528, kv=0.156, nuv=0.00354, B1=30.72, d1=0.0525,
                                 B2=-30.29, d2=0.0611, omega=0.92):
    """Struct of examples/asset_pricing/garleanu_panageas_long_run_risk.jl:41-72; the constructor normalises B1, B2."""
    scale = delta / (delta + d1) * B1 + delta / (delta + d2) * B2
    return dict(gA=gA, psA=psA, gB=gB, psB=psB, rho=rho, delta=delta, nuA=nuA, mubar=mubar, vbar=vbar, kmu=kmu,
                numu=numu, kv=kv, nuv=nuv, B1=B1 / scale, d1=d1, B2=B2 / scale, d2=d2, omega=omega)


GP_LRR_UNKNOWNS = ("pA", "pB", "phi1", "phi2", "r", "kC", "kM", "kV")
GP_LRR_SAVED = ("kappa", "r_implied", "kC_implied", "kM_implied", "kV_implied", "r_gap", "kC_gap", "kM_gap", "kV_gap",
                "pA_residual", "pB_residual", "phi1_residual", "phi2_residual", "mux", "sigmax", "sigmaxC", "sigmaxM",
                "sigmaxV")


def garleanu_panageas_lrr(gA, psA, gB, psB, rho, delta, nuA, mubar, vbar, kmu, numu, kv, nuv, B1, d1, B2, d2, omega):
    """examples/asset_pricing/garleanu_panageas_long_run_risk.jl:74-232: states (x, mu, v), unknowns pA, pB, phi1, phi2
    (valuation ratios, backward equations) and r, kC, kM, kV (market-clearing prices, forward equations).  The
    two-pass upwinding in x (`@goto start`, :117-165) is evaluated for both passes and selected pointwise."""
    def pde(state, u):
        x, mu, v = state["x"], state["mu"], state["v"]
        pA, pB, f1, f2 = u["pA"], u["pB"], u["phi1"], u["phi2"]
        r, kC, kM, kV = u["r"], u["kC"], u["kM"], u["kV"]
        with np.errstate(all="ignore"):
            sc = np.sqrt(v)
            mumu = kmu * (mubar - mu)
            smu = numu * np.sqrt(v)
            muv = kv * (vbar - v)
            sv = nuv * np.sqrt(v)
            upm, upv = mumu >= 0, muv >= 0
            pAmu = np.where(upm, u["pAmu_up"], u["pAmu_down"])
            pBmu = np.where(upm, u["pBmu_up"], u["pBmu_down"])
            f1mu = np.where(upm, u["phi1mu_up"], u["phi1mu_down"])
            f2mu = np.where(upm, u["phi2mu_up"], u["phi2mu_down"])
            pAv = np.where(upv, u["pAv_up"], u["pAv_down"])
            pBv = np.where(upv, u["pBv_up"], u["pBv_down"])
            f1v = np.where(upv, u["phi1v_up"], u["phi1v_down"])
            f2v = np.where(upv, u["phi2v_up"], u["phi2v_down"])
            G = 1 / (x / gA + (1 - x) / gB)
            aA = (1 - gA * psA) / (gA * (psA - 1))
            aB = (1 - gB * psB) / (gB * (psB - 1))
            pAmu_pA, pBmu_pB = pAmu / pA, pBmu / pB
            pAv_pA, pBv_pB = pAv / pA, pBv / pB
            k2 = kC ** 2 + kM ** 2 + kV ** 2
            human_capital = omega * (B1 * f1 + B2 * f2)

            def block(pAx, pBx):
                pAx_pA = pAx / pA
                pBx_pB = pBx / pB
                denom = 1 - x * aA * pAx_pA
                sxC = x * (kC / gA - sc) / denom
                sxM = x * (kM / gA + aA * pAmu_pA * smu) / denom
                sxV = x * (kV / gA + aA * pAv_pA * sv) / denom
                sx2 = sxC ** 2 + sxM ** 2 + sxV ** 2
                spA_C = pAx_pA * sxC
                spA_M = pAx_pA * sxM + pAmu_pA * smu
                spA_V = pAx_pA * sxV + pAv_pA * sv
                spB_C = pBx_pB * sxC
                spB_M = pBx_pB * sxM + pBmu_pB * smu
                spB_V = pBx_pB * sxV + pBv_pB * sv
                sCA_C = kC / gA + aA * spA_C
                sCA_M = kM / gA + aA * spA_M
                sCA_V = kV / gA + aA * spA_V
                sCB_C = kC / gB + aB * spB_C
                sCB_M = kM / gB + aB * spB_M
                sCB_V = kV / gB + aB * spB_V
                spA2 = spA_C ** 2 + spA_M ** 2 + spA_V ** 2
                spB2 = spB_C ** 2 + spB_M ** 2 + spB_V ** 2
                kspA = kC * spA_C + kM * spA_M + kV * spA_V
                kspB = kC * spB_C + kM * spB_M + kV * spB_V
                mcA = (1 + psA) / (2 * gA) * k2 + aA * kspA - 0.5 * aA * spA2
                mcB = (1 + psB) / (2 * gB) * k2 + aB * kspB - 0.5 * aB * spB2
                muCA = psA * (r - rho) + mcA
                muCB = psB * (r - rho) + mcB
                mux = x * (muCA - mu) + delta * (nuA / pA * human_capital - x) - sc * sxC
                return (pAx_pA, pBx_pB, sxC, sxM, sxV, sx2, spA_C, spA_M, spA_V, spB_C, spB_M, spB_V, sCA_C, sCA_M, sCA_V,
                        sCB_C, sCB_M, sCB_V, mcA, mcB, muCA, muCB, mux)
            first = block(u["pAx_up"], u["pBx_up"])
            second = block(u["pAx_down"], u["pBx_down"])
            redo = first[-1] <= 0
            (pAx_pA, pBx_pB, sxC, sxM, sxV, sx2, spA_C, spA_M, spA_V, spB_C, spB_M, spB_V, sCA_C, sCA_M, sCA_V, sCB_C, sCB_M,
             sCB_V, mcA, mcB, muCA, muCB, mux) = (np.where(redo, b2, b1) for b1, b2 in zip(first, second))
            f1x = np.where(redo, u["phi1x_down"], u["phi1x_up"])
            f2x = np.where(redo, u["phi2x_down"], u["phi2x_up"])
            cm, cv = sxM * smu >= 0, sxV * sv >= 0
            pAxmu = np.where(cm, u["pAxmu_up"], u["pAxmu_down"])
            pBxmu = np.where(cm, u["pBxmu_up"], u["pBxmu_down"])
            f1xmu = np.where(cm, u["phi1xmu_up"], u["phi1xmu_down"])
            f2xmu = np.where(cm, u["phi2xmu_up"], u["phi2xmu_down"])
            pAxv = np.where(cv, u["pAxv_up"], u["pAxv_down"])
            pBxv = np.where(cv, u["pBxv_up"], u["pBxv_down"])
            f1xv = np.where(cv, u["phi1xv_up"], u["phi1xv_down"])
            f2xv = np.where(cv, u["phi2xv_up"], u["phi2xv_down"])
            f1x_f1, f2x_f2 = f1x / f1, f2x / f2
            f1mu_f1, f2mu_f2 = f1mu / f1, f2mu / f2
            f1v_f1, f2v_f2 = f1v / f1, f2v / f2
            sf1_C = f1x_f1 * sxC
            sf1_M = f1x_f1 * sxM + f1mu_f1 * smu
            sf1_V = f1x_f1 * sxV + f1v_f1 * sv
            sf2_C = f2x_f2 * sxC
            sf2_M = f2x_f2 * sxM + f2mu_f2 * smu
            sf2_V = f2x_f2 * sxV + f2v_f2 * sv
            mupA = (pAx_pA * mux + pAmu_pA * mumu + pAv_pA * muv + 0.5 * u["pAxx"] / pA * sx2 + 0.5 * u["pAmumu"] / pA * smu ** 2
                    + 0.5 * u["pAvv"] / pA * sv ** 2 + pAxmu / pA * sxM * smu + pAxv / pA * sxV * sv)
            mupB = (pBx_pB * mux + pBmu_pB * mumu + pBv_pB * muv + 0.5 * u["pBxx"] / pB * sx2 + 0.5 * u["pBmumu"] / pB * smu ** 2
                    + 0.5 * u["pBvv"] / pB * sv ** 2 + pBxmu / pB * sxM * smu + pBxv / pB * sxV * sv)
            muf1 = (f1x_f1 * mux + f1mu_f1 * mumu + f1v_f1 * muv + 0.5 * u["phi1xx"] / f1 * sx2
                    + 0.5 * u["phi1mumu"] / f1 * smu ** 2 + 0.5 * u["phi1vv"] / f1 * sv ** 2 + f1xmu / f1 * sxM * smu
                    + f1xv / f1 * sxV * sv)
            muf2 = (f2x_f2 * mux + f2mu_f2 * mumu + f2v_f2 * muv + 0.5 * u["phi2xx"] / f2 * sx2
                    + 0.5 * u["phi2mumu"] / f2 * smu ** 2 + 0.5 * u["phi2vv"] / f2 * sv ** 2 + f2xmu / f2 * sxM * smu
                    + f2xv / f2 * sxV * sv)
            sCAspA = sCA_C * spA_C + sCA_M * spA_M + sCA_V * spA_V
            sCBspB = sCB_C * spB_C + sCB_M * spB_M + sCB_V * spB_V
            kspA_CA = kC * (spA_C + sCA_C) + kM * (spA_M + sCA_M) + kV * (spA_V + sCA_V)
            kspB_CB = kC * (spB_C + sCB_C) + kM * (spB_M + sCB_M) + kV * (spB_V + sCB_V)
            pAt = -pA * (1 / pA + muCA + mupA + sCAspA - r - delta - kspA_CA)
            pBt = -pB * (1 / pB + muCB + mupB + sCBspB - r - delta - kspB_CB)
            ksf1_Y = kC * (sf1_C + sc) + kM * sf1_M + kV * sf1_V
            ksf2_Y = kC * (sf2_C + sc) + kM * sf2_M + kV * sf2_V
            f1t = -f1 * (1 / f1 + mu - delta - d1 + muf1 + sc * sf1_C - r - ksf1_Y)
            f2t = -f2 * (1 / f2 + mu - delta - d2 + muf2 + sc * sf2_C - r - ksf2_Y)
            kC_imp = G * (sc - x * aA * spA_C - (1 - x) * aB * spB_C)
            kM_imp = G * (0.0 - x * aA * spA_M - (1 - x) * aB * spB_M)
            kV_imp = G * (0.0 - x * aA * spA_V - (1 - x) * aB * spB_V)
            r_imp = rho + (mu - x * mcA - (1 - x) * mcB - delta * ((nuA / pA + (1 - nuA) / pB) * human_capital - 1)) / (
                psA * x + psB * (1 - x))
            rt, kCt, kMt, kVt = r_imp - r, kC_imp - kC, kM_imp - kM, kV_imp - kV
            z = 0.0 * x   # saved fields keep the grid shape even where the expression is a scalar
        return ({"pAt": pAt, "pBt": pBt, "phi1t": f1t, "phi2t": f2t, "rt": rt, "kCt": kCt, "kMt": kMt, "kVt": kVt},
                {"kappa": np.sqrt(k2), "r_implied": r_imp, "kC_implied": kC_imp, "kM_implied": kM_imp, "kV_implied": kV_imp,
                 "r_gap": rt, "kC_gap": kCt, "kM_gap": kMt, "kV_gap": kVt, "pA_residual": pAt, "pB_residual": pBt,
                 "phi1_residual": f1t, "phi2_residual": f2t, "mux": mux + z, "sigmax": np.sqrt(sx2), "sigmaxC": sxC,
                 "sigmaxM": sxM, "sigmaxV": sxV})
    return pde


def garleanu_panageas_lrr_problem(xn=20, mun=5, vn=5, **over):
    """Grid, guess and the signed pseudo-time steps of the example's first stage
    (garleanu_panageas_long_run_risk.jl:235-276, 337-344): returns (grid, guess, params, delta_profile, options)."""
    from scipy import stats
    prm = garleanu_panageas_lrr_params(**over)
    musd = math.sqrt(prm["numu"] ** 2 * prm["vbar"] / (2 * prm["kmu"]))
    shape_v = 2 * prm["kv"] * prm["vbar"] / prm["nuv"] ** 2
    scale_v = prm["nuv"] ** 2 / (2 * prm["kv"])
    x = julia_range(0.0, 1.0, xn) ** 2
    mus = julia_range(stats.norm.ppf(0.10, prm["mubar"], musd), stats.norm.ppf(0.90, prm["mubar"], musd), mun)
    vs = julia_range(stats.gamma.ppf(0.10, shape_v, scale=scale_v), stats.gamma.ppf(0.90, shape_v, scale=scale_v), vn)
    S = (xn, mun, vn)
    X, MU, V = np.meshgrid(x, mus, vs, indexing="ij")
    sc = np.sqrt(V)
    Gam = 1 / (X / prm["gA"] + (1 - X) / prm["gB"])
    kC = Gam * sc
    k2 = kC ** 2
    mcA = (1 + prm["psA"]) / (2 * prm["gA"]) * k2
    mcB = (1 + prm["psB"]) / (2 * prm["gB"]) * k2
    human_capital = prm["omega"] * (prm["B1"] + prm["B2"])
    r = prm["rho"] + (MU - X * mcA - (1 - X) * mcB - prm["delta"] * (human_capital - 1)) / (
        prm["psA"] * X + prm["psB"] * (1 - X))
    guess = {"pA": np.ones(S), "pB": np.ones(S), "phi1": np.ones(S), "phi2": np.ones(S), "r": r, "kC": kC,
             "kM": np.zeros(S), "kV": np.zeros(S)}
    delta = {"pA": 100.0, "pB": 100.0, "phi1": 100.0, "phi2": 100.0, "r": -1.0, "kC": -1.0, "kM": -1.0, "kV": -1.0}
    return {"x": x, "mu": mus, "v": vs}, guess, prm, delta, dict(maxΔ=3.0, abstol=1e-5, maxiters=100)


def leland(r=0.02, mu=0.0, sigma=0.1, C=0.05, tau=0.2):
    """examples/corporate_finance/leland.jl:44-53 (equity value with the default option as a lower bound)."""
    def pde(state, u):
        d = state["delta"]
        mud, sd = mu * d, sigma * d
        Ed = np.where(mud >= 0, u["Edelta_up"], u["Edelta_down"])
        f = d - (1 - tau) * C
        return {"Et": -(f + Ed * mud + 0.5 * u["Edeltadelta"] * sd ** 2 - r * u["E"])}
    return pde


def leland_problem(**over):
    """Grid, guess, lower bound and boundary derivatives of the example (leland.jl:31-37, 68-78)."""
    prm = dict(r=0.02, mu=0.0, sigma=0.1, C=0.05, tau=0.2)
    prm.update(over)
    d = np.arange(0, 41) * 0.005          # range(0, 0.2, step = 0.005)
    d = np.array([julia_range(0.0, 0.2, 41)[i] for i in range(41)])
    guess = np.maximum(d / (prm["r"] - prm["mu"]) - (1 - prm["tau"]) * prm["C"] / prm["r"], 0.0)
    return {"delta": d}, {"E": guess}, prm, np.zeros(41), {"Edelta": (0.0, 1 / (prm["r"] - prm["mu"]))}


def leland_closed_form(d, r, mu, sigma, C, tau):
    """Closed-form equity value above the default threshold (leland.jl:88-103); returns (E, delta_star)."""
    a, b, c = 0.5 * sigma ** 2, mu - 0.5 * sigma ** 2, -r
    beta = min((-b + np.sqrt(b * b - 4 * a * c)) / (2 * a), (-b - np.sqrt(b * b - 4 * a * c)) / (2 * a))
    dstar = beta / (beta - 1) * (r - mu) * (1 - tau) * C / r
    slope = 1 / (r - mu)
    load = -slope * dstar ** (1 - beta) / beta
    with np.errstate(all="ignore"):
        E = np.where(d >= dstar, slope * d - (1 - tau) * C / r + load * d ** beta, 0.0)
    return E, dstar


def wang_wang_yang(mu=0.015, sigma=0.1, r=0.035, rho=0.04, gamma=3.0, psi=1.1, wmin=0.0, wmax=1000.0):
    """examples/consumption_saving/wang_wang_yang.jl:50-95 (unknown p on the wealth-income ratio w)."""
    eps = np.finfo(float).eps

    def pde(state, u):
        w = state["w"]
        p, pw_up, pw_down, pww = u["p"], u["pw_up"], u["pw_down"], u["pww"]
        mstar = r + psi * (rho - r)
        with np.errstate(all="ignore"):
            p_trial = np.maximum(p, eps)
            cmax = 100.0 * (1 + np.maximum((r - mu + sigma ** 2) * w, 0.0))
            c_up = np.where(pw_up > 0, np.minimum(mstar * p * np.abs(pw_up) ** (-psi), cmax), cmax)
            muw_up = (r - mu + sigma ** 2) * w + 1 - c_up
            c_down = np.where(pw_down > 0, np.minimum(mstar * p * np.abs(pw_down) ** (-psi), cmax), cmax)
            muw_down = (r - mu + sigma ** 2) * w + 1 - c_down
            use_up = muw_up >= 0
            use_down = (~use_up) & (muw_down <= 0) & (w > wmin)
            pw_a = np.where(pw_up > 0, pw_up, (c_up / (mstar * p_trial)) ** (-1 / psi))
            pw_b = np.where(pw_down > 0, pw_down, (c_down / (mstar * p_trial)) ** (-1 / psi))
            c3 = 1 + (r - mu + sigma ** 2) * w
            pw_c = (c3 / (mstar * p_trial)) ** (-1 / psi)
            pw = np.where(use_up, pw_a, np.where(use_down, pw_b, pw_c))
            c = np.where(use_up, c_up, np.where(use_down, c_down, c3))
            muw = np.where(use_up, muw_up, np.where(use_down, muw_down, 0.0))
            pt = -((((r + psi * (rho - r)) * pw ** (1 - psi) - psi * rho) / (psi - 1) + mu - gamma * sigma ** 2 / 2) * p
                   + ((r - mu + gamma * sigma ** 2) * w + 1) * pw + sigma ** 2 * w ** 2 / 2 * (pww - gamma * pw ** 2 / p))
        return {"pt": pt}, {"c": c, "muw": muw}
    return pde


def wang_wang_yang_problem(n=1001, **over):
    """Grid, guess and boundary derivatives of the example (wang_wang_yang.jl:33-41, 100)."""
    prm = dict(mu=0.015, sigma=0.1, r=0.035, rho=0.04, gamma=3.0, psi=1.1, wmin=0.0, wmax=1000.0)
    prm.update(over)
    w = julia_range(prm["wmin"], prm["wmax"], n)
    return {"w": w}, {"p": 1 + w}, prm, {"pw": (1.0, 1.0)}


def linear_upwind_test(s=-1.0, a=1.0, dir=1.0):
    """test/runtests.jl:326-372: Vt = s (1 + a V + Vx) with a forward (dir >= 0) or backward first difference."""
    def pde(state, u):
        vx = u["Vx_up"] if dir >= 0 else u["Vx_down"]
        return {"Vt": s * (1.0 + a * u["V"] + vx) + 0.0 * state["x"]}
    return pde


HK_SAVED = ("pS", "p", "r", "kappa", "sigmap", "muR", "sigmaR", "alphaI", "mux", "sigmax", "x", "px", "pxx", "sigmaI")


def he_krishnamurthy(m=4.0, lam=0.6, g=0.02, sigma=0.09, rho=0.04, gamma=2.0, l=1.84):
    """examples/asset_pricing/he_krishnamurthy.jl:45-92 (unknown pS on the specialists' wealth share x)."""
    def pde(state, u):
        x = state["x"]
        pS, pSxx = u["pS"], u["pSxx"]

        def block(pSx):
            den = x / pS + (1 - x) * rho      # Julia's x^3 is x*x*x (literal_pow), numpy's ** 3 is pow(): write it out
            p = (1 + l) / (x / pS + (1 - x) * rho)
            px = -(1 + l) / (x / pS + (1 - x) * rho) ** 2 * (1 / pS - rho - x / pS ** 2 * pSx)
            pxx = (-(1 + l) / (x / pS + (1 - x) * rho) ** 2 * (-1 / pS ** 2 * pSx - 1 / pS ** 2 * pSx - x / pS ** 2 * pSxx
                                                             + 2 * x / (pS * pS * pS) * pSx ** 2)
                   + 2 * (1 + l) / (den * den * den) * (1 / pS - rho - x / pS ** 2 * pSx) ** 2)
            aI = 1 / (1 - lam * (1 - x)) + 0.0 * pS
            aI = np.where(x * (1 + m) * aI < 1, 1 / (x * (1 + m)), aI)
            sigx = x * (aI - 1) * sigma / (1 - x * (aI - 1) * px / p)
            spS = pSx / pS * sigx
            sp = px / p * sigx
            sR = sigma + sp
            kappa = gamma * (aI * sR - spS)
            mux = x * (1 - x) * ((aI - (1 - aI * x) / (1 - x)) * kappa * sR - 1 / pS + rho
                                 - (aI - (1 - aI * x) / (1 - x)) * sR ** 2 - l / ((1 - x) * p))
            return p, px, pxx, aI, sigx, spS, sp, sR, kappa, mux
        with np.errstate(all="ignore"):
            first, second = block(u["pSx_up"]), block(u["pSx_down"])
            redo = first[-1] < 0
            p, px, pxx, aI, sigx, spS, sp, sR, kappa, mux = (np.where(redo, b2, b1) for b1, b2 in zip(first, second))
            pSx = np.where(redo, u["pSx_down"], u["pSx_up"])
            mupS = pSx / pS * mux + 0.5 * pSxx / pS * sigx ** 2
            mup = px / p * mux + 0.5 * pxx / p * sigx ** 2
            r = 1 / p + g + mup + sigma * sp - kappa * sR
            sCS = kappa / gamma
            muCS = (r - rho) / gamma + (1 + 1 / gamma) / (2 * gamma) * kappa ** 2
            pSt = -(1 / pS + muCS + mupS + sCS * spS - r - kappa * (sCS + spS))
        saved = dict(zip(HK_SAVED, (pS, p, r, kappa, sp, r + kappa * sR, sR, aI, mux, sigx, x + 0.0 * pS, px, pxx, aI * sR)))
        return {"pSt": pSt}, saved
    return pde


def he_krishnamurthy_problem(xn=100, **over):
    """Grid and guess of the example (he_krishnamurthy.jl:31-39)."""
    prm = dict(m=4.0, lam=0.6, g=0.02, sigma=0.09, rho=0.04, gamma=2.0, l=1.84)
    prm.update(over)
    x = julia_range(0.0, 1.0, xn + 2)[1:-1] ** 1.5
    return {"x": x}, {"pS": np.ones(xn)}, prm


def normal_mgf(t, mu, sigma):
    """Distributions.jl `mgf(d::Normal, t) = exp(t * mu + sigma^2 / 2 * t^2)` (third-party; parity unpinned)."""
    return float(np.exp(t * mu + sigma ** 2 / 2 * t ** 2))


def wachter(mu=0.025, sigma=0.02, lambar=0.0355, kappa_lambda=0.08, nu_lambda=0.067, muZ=-0.4, sigmaZ=0.25,
            rho=0.012, gamma=3.0, psi=1.1, phi=2.6):
    """examples/asset_pricing/wachter.jl:66-92 (unknown p on the disaster intensity lambda)."""
    m1, mg, m1g = normal_mgf(1.0, muZ, sigmaZ), normal_mgf(-gamma, muZ, sigmaZ), normal_mgf(1.0 - gamma, muZ, sigmaZ)

    def pde(state, u):
        lam = state["lambda"]
        p, pll = u["p"], u["plambdalambda"]
        mul = kappa_lambda * (lambar - lam)
        sigl = nu_lambda * np.sqrt(lam)
        pl = np.where(mul >= 0, u["plambda_up"], u["plambda_down"])
        mup = pl / p * mul + 0.5 * pll / p * sigl ** 2
        sp = pl / p * sigl
        kZc = gamma * sigma
        kZl = (gamma * psi - 1) / (psi - 1) * sp
        eta = lam * (m1 - 1 + mg - 1 - (m1g - 1))
        r = (rho + mu / psi - (1 + 1 / psi) / 2 * gamma * sigma ** 2 - lam * (mg - 1)
             + (1 / psi - gamma) / (1 - gamma) * lam * (m1g - 1) - (gamma * psi - 1) / (2 * (psi - 1)) * sp ** 2)
        return {"pt": -p * (1 / p + mu + mup + lam * (m1 - 1) - r - kZc * sigma - kZl * sp - eta)}
    return pde


def wachter_problem(n=30, **over):
    """Grid and guess of the example (wachter.jl:48-58)."""
    prm = dict(mu=0.025, sigma=0.02, lambar=0.0355, kappa_lambda=0.08, nu_lambda=0.067, muZ=-0.4, sigmaZ=0.25,
               rho=0.012, gamma=3.0, psi=1.1, phi=2.6)
    prm.update(over)
    return {"lambda": julia_range(0.0, 0.1, n)}, {"p": np.ones(n)}, prm


HADDAD_DEFAULT = dict(mubar=0.018, vbar=0.00052, kappa_mu=0.3, nu_mu=0.456, kappa_v=0.012, nu_v=0.00472,
                      alphabar=1.2, lam=0.018, rho=0.0132, gamma=10.0, psi=1.5)


def haddad(mubar, vbar, kappa_mu, nu_mu, kappa_v, nu_v, alphabar, lam, rho, gamma, psi):
    """examples/asset_pricing/haddad.jl:63-93 (long-run risk with an endogenous active-capital share)."""
    def pde(state, u):
        mu, v = state["mu"], state["v"]
        p = u["p"]
        muc = mu
        sigc = np.sqrt(v)
        mumu = kappa_mu * (mubar - mu)
        sigmu = nu_mu * np.sqrt(v)
        muv = kappa_v * (vbar - v)
        pmu = np.where(mumu >= 0, u["pmu_up"], u["pmu_down"])
        pv = np.where(muv >= 0, u["pv_up"], u["pv_down"])
        sigv = nu_v * np.sqrt(v)
        with np.errstate(all="ignore"):
            sZmu = pmu / p * sigmu
            sZv = pv / p * sigv
            sp2 = sZmu ** 2 + sZv ** 2
            mup = pmu / p * mumu + pv / p * muv + 0.5 * u["pmumu"] / p * sigmu ** 2 + 0.5 * u["pvv"] / p * sigv ** 2
            astar = alphabar - np.sqrt(lam * alphabar / (0.5 * gamma * (sigc ** 2 + sZmu ** 2 + sZv ** 2)))
            astar = np.where(astar > 1.0, 1.0, np.where(astar < 0.0, 0.0, astar))      # clamp(astar, 0.0, 1.0)
            pt = -p * (1 / p - rho + (1 - 1 / psi) * (muc - 0.5 * gamma * sigc ** 2 * (1 - (astar - 1) ** 2)) + mup
                       + (0.5 * (1 / psi - gamma) / (1 - 1 / psi) + 0.5 * gamma * (1 - 1 / psi) * (astar - 1) ** 2) * sp2)
        return {"pt": pt}
    return pde


def haddad_problem(n_mu=30, n_v=30, params=None):
    """Grid and guess of the example (haddad.jl:41-56)."""
    from scipy import stats
    prm = dict(HADDAD_DEFAULT if params is None else params)
    sd = math.sqrt(prm["nu_mu"] ** 2 * prm["vbar"] / (2 * prm["kappa_mu"]))
    mus = julia_range(stats.norm.ppf(0.001, prm["mubar"], sd), stats.norm.ppf(0.999, prm["mubar"], sd), n_mu)
    shape = 2 * prm["kappa_v"] * prm["vbar"] / prm["nu_v"] ** 2
    scale = prm["nu_v"] ** 2 / (2 * prm["kappa_v"])
    vs = julia_range(stats.gamma.ppf(0.001, shape, scale=scale), stats.gamma.ppf(0.999, shape, scale=scale), n_v)
    return {"mu": mus, "v": vs}, {"p": np.ones((n_mu, n_v))}, prm


def bolton_chen_wang(r=0.06, delta=0.10, A=0.18, sigma=0.09, theta=1.5, lam=0.01, l=0.9, gamma=0.06, phi=0.01):
    """examples/corporate_finance/bolton_chen_wang.jl:56-72 (firm value v on the cash-capital ratio w)."""
    def pde(state, u):
        w = state["w"]
        v, vww = u["v"], u["vww"]
        with np.errstate(all="ignore"):
            i_up = 1 / theta * (v / u["vw_up"] - w - 1)
            muw = (r - lam) * w + A - i_up - theta * i_up ** 2 / 2 - (i_up - delta) * w
            redo = muw <= 0
            vw = np.where(redo, u["vw_down"], u["vw_up"])
            i = np.where(redo, 1 / theta * (v / u["vw_down"] - w - 1), i_up)
            vt = -((i - delta) * (v - vw * w) + ((r - lam) * w + A - i - theta * i ** 2 / 2) * vw + sigma ** 2 / 2 * vww - r * v)
        return {"vt": vt}, {"v": v, "vw": vw, "vww": vww, "w": w + 0.0 * v}
    return pde


def bolton_chen_wang_problem(n=100, **over):
    """Grid, guess and the first boundary pair of the example (bolton_chen_wang.jl:36-42, 78)."""
    prm = dict(r=0.06, delta=0.10, A=0.18, sigma=0.09, theta=1.5, lam=0.01, l=0.9, gamma=0.06, phi=0.01)
    prm.update(over)
    w = julia_range(0.0, 0.3, n)
    return {"w": w}, {"v": w.copy()}, prm, {"vw": (1.5, 1.0)}


def bolton_chen_wang_boundary_residual(prm, saved):
    """The two smooth-pasting conditions the example's outer `nlsolve` drives to zero (bolton_chen_wang.jl:84-95)."""
    v, vw, w = saved["v"], saved["vw"], saved["w"]
    mi = int(np.argmin(np.abs(1 + prm["gamma"] - vw)))
    wi = int(np.argmin(np.abs(1 - vw)))
    i = 1 / prm["theta"] * (v[wi] - w[wi] - 1)
    return np.array([v[mi] - prm["phi"] - (1 + prm["gamma"]) * w[mi] - v[0],
                     prm["r"] * v[wi] - (i - prm["delta"]) * (v[wi] - w[wi])
                     - ((prm["r"] - prm["lam"]) * w[wi] + prm["A"] - i - prm["theta"] * i ** 2 / 2)])


GOMEZ_SAVED = ("x", "mux", "sigmax", "kappa", "r", "sigmaWH", "muWH", "sigmap", "p", "px", "sigmaR", "muR",
               "muRlambda", "sigmaRlambda", "muErelative", "sigmaErelative", "muHrelative", "sigmaHrelative",
               "muCE", "muxW", "sigmaxW", "pH", "muCH", "wedge", "xW", "sigmaWE", "sigmaCH", "alphaE", "pE", "mup")


def gomez(gamma=10.3, psi=0.05, rho=0.1, rhoE=0.022, alphaE=2.0, nu=0.1, eta=0.015, delta=0.025, phi=0.01, piE=0.09,
          g=0.02, sigma=0.04, lam=2.3, tau=0.0):
    """examples/asset_pricing/gomez.jl:58-118 (unknown pH on the entrepreneurs' consumption share x)."""
    def pde(state, u):
        x = state["x"]
        pH, pHxx = u["pH"], u["pHxx"]
        with np.errstate(all="ignore"):
            p = x / rhoE + (1 - x) * pH
            xW = x / rhoE / p
            aEeff = np.where(xW * alphaE >= 1, 1 / xW, alphaE)

            def block(pHx):
                px = 1 / rhoE - pH + (1 - x) * pHx
                sigx = x * (aEeff - 1) * sigma / (1 - x * aEeff * px / p)
                spH = pHx / pH * sigx
                sp = px / p * sigx
                sCE = aEeff * (sigma + sp)
                sCH = np.where(x < 1, (sigma - x * sCE) / (1 - x), 0.0)
                sWH = sCH + spH
                kappa = gamma * sCH + (gamma * psi - 1) / (psi - 1) * spH
                PhiE = aEeff * kappa * (sigma + sp)
                PhiH = (kappa ** 2 * (1 + psi) / (2 * gamma) + (1 - psi * gamma) / (gamma * (psi - 1)) * kappa * spH
                        - (1 - gamma * psi) / (2 * gamma * (psi - 1)) * spH ** 2)
                wealth = (eta + delta + phi + tau) / (eta + delta) * p
                wedge = (eta + delta) * ((rhoE * piE + (1 - piE) / pH) * wealth - 1)
                Psi = x + (1 - x) * psi
                P = (x * rhoE + (1 - x) * psi * rho) / Psi
                r = P + 1 / Psi * (g - x * PhiE - (1 - x) * PhiH - wedge) + tau
                muCE = r - tau - rhoE + PhiE
                muCH = psi * (r - tau - rho) + PhiH
                mux = x * (muCE - g) + (eta + delta) * (rhoE * piE * wealth - x) - sigma * sigx
                return px, sigx, spH, sp, sCE, sCH, sWH, kappa, wedge, r, muCE, muCH, mux
            first, second = block(u["pHx_up"]), block(u["pHx_down"])
            redo = first[-1] <= 0
            px, sigx, spH, sp, sCE, sCH, sWH, kappa, wedge, r, muCE, muCH, mux = (
                np.where(redo, b2, b1) for b1, b2 in zip(first, second))
            pHx = np.where(redo, u["pHx_down"], u["pHx_up"])
            mupH = pHx / pH * mux + 0.5 * pHxx / pH * sigx ** 2
            pxx = -2 * pHx + (1 - x) * pHxx
            mup = px / p * mux + 0.5 * pxx / p * sigx ** 2
            muWH = muCH + mupH + sCH * spH
            pHt = -pH * (1 / pH + muWH - r + tau - kappa * sWH)
            sR = sigma + sp
            muR = r + kappa * sR
            sErel = (aEeff - 1) * (sigma + sp)
            muErel = muCE - (g + mup + sigma * sp) - sErel * (sigma + sp)
            alpha = np.where(x < 1, (1 - xW * aEeff) / (1 - xW), 0.0)
            sHrel = (alpha - 1) * (sigma + sp)
            muHrel = muWH - (g + mup + sigma * sp) - sHrel * (sigma + sp)
            muxW = (xW * (muCE - (g + mup + sigma * sp) - sErel * (sigma + sp))
                    + (eta + delta) * (piE * (eta + delta + phi) / (eta + delta) - xW))
            sxW = xW * (aEeff - 1) * (sigma + sp)
            z = 0.0 * pH
            vals = (x + z, mux, sigx, kappa, r, sWH, muWH, sp, p, px, sR, muR, r + lam * (muR - r), lam * sR, muErel, sErel,
                    muHrel, sHrel, muCE, muxW, sxW, pH, muCH, wedge, xW + z, sCE, sCH, aEeff + z, 1 / rhoE + z, mup)
        return {"pHt": pHt}, dict(zip(GOMEZ_SAVED, vals))
    return pde


def gomez_problem(n=300, **over):
    """Grid and guess of the example (gomez.jl:46-52)."""
    prm = dict(gamma=10.3, psi=0.05, rho=0.1, rhoE=0.022, alphaE=2.0, nu=0.1, eta=0.015, delta=0.025, phi=0.01, piE=0.09,
               g=0.02, sigma=0.04, lam=2.3, tau=0.0)
    prm.update(over)
    return {"x": julia_range(0.0, 1.0, n) ** 2}, {"pH": np.ones(n)}, prm


_LIBM = None


def _cpow(x: float, y: float) -> float:
    """libm pow on scalars (inf/nan instead of Python's exceptions; Julia's `^` on Float64 has the same edge cases)."""
    global _LIBM
    if _LIBM is None:
        import ctypes
        import ctypes.util
        _LIBM = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        _LIBM.pow.restype = ctypes.c_double
        _LIBM.pow.argtypes = [ctypes.c_double, ctypes.c_double]
    return _LIBM.pow(x, y)


def pointwise(scalar_pde, statenames, fields, outname, savednames):
    """Wraps a scalar per-point body (a literal restatement of a branchy Julia body) as the vectorised callable
    `Problem` expects: loops over the grid points exactly like `pde!` (src/pdesolve.jl:436-462)."""
    def pde(state, u):
        shape = np.shape(u[fields[0]])
        st = [np.broadcast_to(state[k], shape).ravel() for k in statenames]
        uu = [np.broadcast_to(u[k], shape).ravel() for k in fields]
        n = uu[0].shape[0]
        out = np.empty(n)
        sv = np.empty((len(savednames), n))
        for i in range(n):
            out[i], sv[:, i] = scalar_pde(*[float(x[i]) for x in st], *[float(x[i]) for x in uu])
        return {outname: out.reshape(shape)}, {k: sv[j].reshape(shape) for j, k in enumerate(savednames)}
    return pde


RISKY_SAVED = ("v", "c", "k", "va", "vaa", "vy", "y", "a", "mua")
RISKY_DEFAULT = dict(kappa_y=0.1, ybar=1.0, sigma_y=0.07, r=0.03, muR=0.04, sigmaR=0.1, rho=0.05, gamma=2.0,
                     amin=0.0, amax=300.0)
_SQRT_EPS = math.sqrt(np.finfo(float).eps)


def ahlm_risky_asset(kappa_y, ybar, sigma_y, r, muR, sigmaR, rho, gamma, amin, amax):
    """examples/consumption_saving/consumption_saving_risky_asset.jl:84-192: consumption-saving with a risky asset
    (portfolio constraint 0 <= k <= a - amin); scalar restatement, statement by statement."""
    isfinite = math.isfinite

    def clamp(x, lo, hi):
        return hi if x > hi else (lo if x < lo else x)

    def body(y, a, v, vy_up, vy_down, va_up, va_down, vyy, vaa):
        muy = kappa_y * (ybar - y)
        cash = y + r * a
        prem = muR - r
        kmax = max(a - amin, 0.0)
        cmax = max(100.0 * (y + r * max(a, 0.0)), _SQRT_EPS)
        cmin = _SQRT_EPS
        if a == amax or abs(a - amax) <= _SQRT_EPS * max(abs(a), abs(amax)):      # a ≈ amax
            va_tail = va_down if (isfinite(va_down) and va_down > 0.0) else va_up
            vaa = -gamma * va_tail / a
        vy = vy_up if muy >= 0 else vy_down

        def policy_candidate(va_c):
            c_c = clamp(_cpow(va_c, -1 / gamma), cmin, cmax) if va_c > 0.0 else cmax
            H_kmax = prem * va_c * kmax + 0.5 * vaa * sigmaR ** 2 * kmax ** 2
            k_c = kmax if H_kmax > 0.0 else 0.0
            if va_c > 0.0 and vaa < 0.0:
                k_merton = clamp(-prem * va_c / (sigmaR ** 2 * vaa), 0.0, kmax)
                H_merton = prem * va_c * k_merton + 0.5 * vaa * sigmaR ** 2 * k_merton ** 2
                H_candidate = prem * va_c * k_c + 0.5 * vaa * sigmaR ** 2 * k_c ** 2
                k_c = k_merton if H_merton >= H_candidate else k_c
            return c_c, k_c, cash + prem * k_c - c_c

        c_up, k_up, mua_up = policy_candidate(va_up)
        if mua_up >= 0.0:
            va, c, k, mua = va_up, c_up, k_up, mua_up
        else:
            c_down, k_down, mua_down = policy_candidate(va_down)
            if (mua_down <= 0.0) and (a > amin):
                va, c, k, mua = va_down, c_down, k_down, mua_down
            else:
                k_hi = kmax
                if prem < 0.0:
                    k_hi = min(k_hi, max((cash - cmin) / (-prem), 0.0))
                if k_hi == 0.0 or not (isfinite(cash) and isfinite(prem) and isfinite(vaa) and isfinite(sigmaR)):
                    k = 0.0
                elif vaa < 0.0 and sigmaR > 0.0:
                    dh0 = prem * _cpow(max(cash, cmin), -gamma)
                    dh1 = prem * _cpow(max(cash + prem * k_hi, cmin), -gamma) + vaa * sigmaR ** 2 * k_hi
                    if isfinite(dh0) and isfinite(dh1):
                        if dh0 <= 0.0:
                            k = 0.0
                        elif dh1 >= 0.0:
                            k = k_hi
                        else:
                            lo, hi = 0.0, k_hi
                            for _ in range(40):
                                mid = (lo + hi) / 2.0
                                dh = prem * _cpow(max(cash + prem * mid, cmin), -gamma) + vaa * sigmaR ** 2 * mid
                                if dh > 0.0:
                                    lo = mid
                                else:
                                    hi = mid
                            k = (lo + hi) / 2.0
                    else:
                        k = 0.0
                else:
                    c0 = max(cash, cmin)
                    c1 = max(cash + prem * k_hi, cmin)
                    h0 = _cpow(c0, 1 - gamma) / (1 - gamma)
                    h1 = _cpow(c1, 1 - gamma) / (1 - gamma) + 0.5 * vaa * sigmaR ** 2 * k_hi ** 2
                    k = k_hi if h1 > h0 else 0.0
                c = max(cash + prem * k, cmin)
                va = _cpow(c, -gamma)
                mua = 0.0
        siga = k * sigmaR
        vt = -(_cpow(c, 1 - gamma) / (1 - gamma) + va * mua + 0.5 * vaa * siga ** 2 + vy * muy + 0.5 * vyy * sigma_y ** 2
               - rho * v)
        return vt, (v, c, k, va, vaa, vy, y, a, mua)
    return pointwise(body, ("y", "a"), ("v", "vy_up", "vy_down", "va_up", "va_down", "vyy", "vaa"), "vt", RISKY_SAVED)


def ahlm_risky_asset_problem(n_y=5, n_a=150, params=None):
    """Grid, guess and boundary derivatives of the example (consumption_saving_risky_asset.jl:41-68, 199-200)."""
    from scipy import stats
    prm = dict(RISKY_DEFAULT if params is None else params)
    shape = 2 * prm["kappa_y"] * prm["ybar"] / prm["sigma_y"] ** 2
    scale = prm["sigma_y"] ** 2 / (2 * prm["kappa_y"])
    ys = julia_range(stats.gamma.ppf(0.001, shape, scale=scale), stats.gamma.ppf(0.999, shape, scale=scale), n_y)
    as_ = prm["amin"] + (prm["amax"] - prm["amin"]) * julia_range(0.0, 1.0, n_a) ** 2
    g, rr = prm["gamma"], prm["r"]
    prem = prm["muR"] - rr
    cshare = (prm["rho"] - rr) / g + rr - (1 - g) * prem ** 2 / (2 * g ** 2 * prm["sigmaR"] ** 2)
    Y, A = np.meshgrid(ys, as_, indexing="ij")
    guess = cshare ** (-g) * (A + Y / rr) ** (1 - g) / (1 - g)
    bc = {"va": ((ys + rr * prm["amin"]) ** (-g), np.full(n_y, cshare ** (-g) * prm["amax"] ** (-g)))}
    return {"y": ys, "a": as_}, {"v": guess}, prm, bc


def pointwise_cubic(a=1.0, k=1.0):
    """test/runtests.jl:422-431: hump!(ydot, y) = y - y^3, at every grid point."""
    def pde(state, u):
        v = u["v"]
        return {"vt": a * v - k * (v * v * v) + 0.0 * state["x"]}
    return pde


def affine_pair(a1=1.0, b1=-2.0, a2=-1.0, b2=3.0):
    """test/runtests.jl:403-411 and :456-489: two uncoupled affine rows at every grid point."""
    def pde(state, u):
        z = 0.0 * state["x"]
        return {"vt": a1 * u["v"] + b1 + z, "pt": a2 * u["p"] + b2 + z}
    return pde


def v_minus_two():
    """test/runtests.jl:227: vt = v - 2."""
    def pde(state, u):
        return {"vt": u["v"] - 2.0 + 0.0 * state["x"]}
    return pde

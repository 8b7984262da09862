"""Registered models: the host-side mirror of the `__device__` functors in csrc/models.cuh.

In the reference a model is any callable `pde(state, u)` (README.md:58-67); the examples
use callable structs of Float64 fields (e.g. examples/asset_pricing/bansal_yaron.jl:22-32).
A GPU backend cannot run an opaque host closure per grid point, so -- as BASELINE.json's
north star prescribes -- each supported example model is a *registered* functor selected by
the type of `pde`.  The dataclasses below carry exactly the reference structs' fields
(same names transliterated to ASCII, same defaults) and know their parameter-vector
layout, state names and unknown names.
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import ClassVar, Tuple

import numpy as np


@dataclass(frozen=True)
class RegisteredModel:
    name: ClassVar[str] = ""
    statenames: ClassVar[Tuple[str, ...]] = ()
    solnames: ClassVar[Tuple[str, ...]] = ()
    savednames: ClassVar[Tuple[str, ...]] = ()
    derivednames: ClassVar[Tuple[str, ...]] = ()     # entries `params()` appends after the struct's fields

    def params(self) -> np.ndarray:
        return np.array([float(getattr(self, f.name)) for f in fields(self)], dtype=np.float64)

    @property
    def ndims(self) -> int:
        return len(self.statenames)

    @property
    def nfuncs(self) -> int:
        return len(self.solnames)


@dataclass(frozen=True)
class ReadmeLinear(RegisteredModel):
    """README.md:48-67."""
    rho: float = 0.05
    kappa: float = 0.2
    name: ClassVar[str] = "readme_linear"
    statenames: ClassVar = ("x",)
    solnames: ClassVar = ("v",)


@dataclass(frozen=True)
class CampbellCochraneModel(RegisteredModel):
    """examples/asset_pricing/campbell_cochrane.jl:28-35."""
    mu: float = 0.0189
    sigma: float = 0.015
    gamma: float = 2.0
    rho: float = 0.116
    kappa_s: float = 0.138
    b: float = 0.0
    name: ClassVar[str] = "campbell_cochrane"
    statenames: ClassVar = ("s",)
    solnames: ClassVar = ("p",)


@dataclass(frozen=True)
class BansalYaronModel(RegisteredModel):
    """examples/asset_pricing/bansal_yaron.jl:22-32."""
    mubar: float = 0.018
    vbar: float = 0.00073
    kappa_mu: float = 0.252
    nu_mu: float = 0.528
    kappa_v: float = 0.156
    nu_v: float = 0.00354
    rho: float = 0.024
    gamma: float = 7.5
    psi: float = 1.5
    name: ClassVar[str] = "bansal_yaron"
    statenames: ClassVar = ("mu", "v")
    solnames: ClassVar = ("p",)


@dataclass(frozen=True)
class AchdouHanLasryLionsMollModel_Diffusion(RegisteredModel):
    """examples/consumption_saving/consumption_saving_diffusion_income.jl:25-34."""
    kappa_y: float = 0.1
    ybar: float = 1.0
    sigma_y: float = 0.07
    r: float = 0.03
    rho: float = 0.05
    gamma: float = 2.0
    amin: float = 0.0
    amax: float = 100.0
    name: ClassVar[str] = "ahlm_diffusion"
    statenames: ClassVar = ("y", "a")
    solnames: ClassVar = ("v",)
    savednames: ClassVar = ("mua", "c")


@dataclass(frozen=True)
class Hjb3dSynth(RegisteredModel):
    """Synthetic 3-state HJB of config C5."""
    kx: float = 0.2
    ky: float = 0.3
    kz: float = 0.4
    sx: float = 0.1
    sy: float = 0.1
    sz: float = 0.1
    corr: float = 0.5
    rho: float = 0.05
    theta: float = 1e-3
    name: ClassVar[str] = "hjb3d_synth"
    statenames: ClassVar = ("x", "y", "z")
    solnames: ClassVar = ("v",)


@dataclass(frozen=True)
class Hjb2dRegimes(RegisteredModel):
    """Two coupled HJB equations on a 2-D grid (csrc/models.cuh: Hjb2dRegimes)."""
    kx1: float = 0.2
    kx2: float = 0.35
    ky: float = 0.3
    sx: float = 0.1
    sy: float = 0.12
    corr: float = 0.3
    rho: float = 0.05
    lam12: float = 0.4
    lam21: float = 0.25
    theta: float = 1e-3
    name: ClassVar[str] = "hjb2d_regimes"
    statenames: ClassVar = ("x", "y")
    solnames: ClassVar = ("v", "w")


@dataclass(frozen=True)
class TuckmanVilaModel(RegisteredModel):
    """examples/asset_pricing/tuckman_vila.jl:20-27; the body takes the time as a third argument, so the
    model is solved with a time grid: pdesolve(m, grid, guess, τs)."""
    c: float = 0.06
    r: float = 0.09
    rho: float = 5.42
    sigma: float = 26.72
    a: float = 26.72
    T: float = 100.0
    name: ClassVar[str] = "tuckman_vila"
    statenames: ClassVar = ("z",)
    solnames: ClassVar = ("F",)


@dataclass(frozen=True)
class DiTellaModel(RegisteredModel):
    """examples/asset_pricing/di_tella.jl:16-38; solve with is_algebraic = {"pA": False, "pB": False, "p": True}."""
    gamma: float = 5.0
    psi: float = 2.0
    rho: float = 0.0665
    tau: float = 1.15
    A: float = 54.0
    delta: float = 0.05
    B: float = (0.20 - 54.0 * (0.02 + 0.05) ** 2) / (0.02 + 0.05)
    sigma: float = 0.0125
    phi: float = 0.2
    nubar: float = 0.25
    kappanu: float = 1.38
    signubar: float = -0.17
    name: ClassVar[str] = "di_tella"
    statenames: ClassVar = ("x", "nu")
    solnames: ClassVar = ("pA", "pB", "p")
    savednames: ClassVar = ("sigmax", "Q", "r", "kappa", "sigmap", "growth", "investment")


@dataclass(frozen=True)
class AchdouHanLasryLionsMollModel_TwoStates(RegisteredModel):
    """examples/consumption_saving/consumption_saving_two_income.jl:22-32 (amin already nudged inside the
    natural borrowing limit, :44)."""
    yl: float = 0.5
    yh: float = 1.5
    llh: float = 0.2
    lhl: float = 0.2
    r: float = 0.03
    rho: float = 0.04
    gamma: float = 2.0
    amin: float = -0.5 / 0.03 + 0.001
    amax: float = 50.0
    name: ClassVar[str] = "ahlm_two_states"
    statenames: ClassVar = ("a",)
    solnames: ClassVar = ("vl", "vh")
    savednames: ClassVar = ("cl", "ch", "mula", "muha")


@dataclass(frozen=True)
class GarleanuPanageasModel(RegisteredModel):
    """examples/asset_pricing/garleanu_panageas.jl:20-40.  B1, B2 are the NORMALISED loadings the reference
    struct stores (its constructor divides by δ/(δ+δ1) B1 + δ/(δ+δ2) B2); `default()` builds them."""
    gA: float = 1.5
    psA: float = 0.7
    gB: float = 10.0
    psB: float = 0.05
    rho: float = 0.001
    delta: float = 0.02
    nuA: float = 0.01
    mu: float = 0.02
    sigma: float = 0.041
    B1: float = 30.72 / (0.02 / (0.02 + 0.0525) * 30.72 + 0.02 / (0.02 + 0.0611) * -30.29)
    d1: float = 0.0525
    B2: float = -30.29 / (0.02 / (0.02 + 0.0525) * 30.72 + 0.02 / (0.02 + 0.0611) * -30.29)
    d2: float = 0.0611
    omega: float = 0.92
    name: ClassVar[str] = "garleanu_panageas"
    statenames: ClassVar = ("x",)
    solnames: ClassVar = ("pA", "pB", "phi1", "phi2")
    savednames: ClassVar = ("r", "kappa")


_LRR_SCALE = 0.02 / (0.02 + 0.0525) * 30.72 + 0.02 / (0.02 + 0.0611) * -30.29


@dataclass(frozen=True)
class GarleanuPanageasLongRunRiskModel(RegisteredModel):
    """examples/asset_pricing/garleanu_panageas_long_run_risk.jl:41-72 (the reference's only 3-D example: 8 unknowns
    on an (x, mu, v) grid).  B1, B2 are the NORMALISED loadings the reference struct stores.  Solve its first stage
    with Δ = {"pA": 100, "pB": 100, "phi1": 100, "phi2": 100, "r": -1, "kC": -1, "kM": -1, "kV": -1}, maxΔ = 3,
    abstol = 1e-5 (:337-344)."""
    gA: float = 1.5
    psA: float = 0.7
    gB: float = 10.0
    psB: float = 0.05
    rho: float = 0.001
    delta: float = 0.02
    nuA: float = 0.01
    mubar: float = 0.018
    vbar: float = 0.00073
    kmu: float = 0.252
    numu: float = 0.528
    kv: float = 0.156
    nuv: float = 0.00354
    B1: float = 30.72 / _LRR_SCALE
    d1: float = 0.0525
    B2: float = -30.29 / _LRR_SCALE
    d2: float = 0.0611
    omega: float = 0.92
    name: ClassVar[str] = "garleanu_panageas_lrr"
    statenames: ClassVar = ("x", "mu", "v")
    solnames: ClassVar = ("pA", "pB", "phi1", "phi2", "r", "kC", "kM", "kV")
    savednames: ClassVar = ("kappa", "r_implied", "kC_implied", "kM_implied", "kV_implied", "r_gap", "kC_gap", "kM_gap",
                            "kV_gap", "pA_residual", "pB_residual", "phi1_residual", "phi2_residual", "mux", "sigmax",
                            "sigmaxC", "sigmaxM", "sigmaxV")


@dataclass(frozen=True)
class LelandModel(RegisteredModel):
    """examples/corporate_finance/leland.jl:22-28; solve with lower_bound = 0 and bc = {"Edelta": (0, 1/(r-mu))}."""
    r: float = 0.02
    mu: float = 0.0
    sigma: float = 0.1
    C: float = 0.05
    tau: float = 0.2
    name: ClassVar[str] = "leland"
    statenames: ClassVar = ("delta",)
    solnames: ClassVar = ("E",)


@dataclass(frozen=True)
class WangWangYangModel(RegisteredModel):
    """examples/consumption_saving/wang_wang_yang.jl:21-30; solve with bc = {"pw": (1.0, 1.0)}."""
    mu: float = 0.015
    sigma: float = 0.1
    r: float = 0.035
    rho: float = 0.04
    gamma: float = 3.0
    psi: float = 1.1
    wmin: float = 0.0
    wmax: float = 1000.0
    name: ClassVar[str] = "wang_wang_yang"
    statenames: ClassVar = ("w",)
    solnames: ClassVar = ("p",)
    savednames: ClassVar = ("c", "muw")


@dataclass(frozen=True)
class LinearUpwindTest(RegisteredModel):
    """test/runtests.jl:326-372: Vt = s (1 + a V + Vx), Vx forward (dir >= 0) or backward (dir < 0)."""
    s: float = -1.0
    a: float = 1.0
    dir: float = 1.0
    name: ClassVar[str] = "linear_upwind_test"
    statenames: ClassVar = ("x",)
    solnames: ClassVar = ("V",)


@dataclass(frozen=True)
class HeKrishnamurthyModel(RegisteredModel):
    """examples/asset_pricing/he_krishnamurthy.jl:21-30."""
    m: float = 4.0
    lam: float = 0.6
    g: float = 0.02
    sigma: float = 0.09
    rho: float = 0.04
    gamma: float = 2.0
    l: float = 1.84
    name: ClassVar[str] = "he_krishnamurthy"
    statenames: ClassVar = ("x",)
    solnames: ClassVar = ("pS",)
    savednames: ClassVar = ("pS", "p", "r", "kappa", "sigmap", "muR", "sigmaR", "alphaI", "mux", "sigmax", "x", "px", "pxx",
                            "sigmaI")


@dataclass(frozen=True)
class WachterModel(RegisteredModel):
    """examples/asset_pricing/wachter.jl:20-39; `ZDistribution = Normal(muZ, sigmaZ)`."""
    mu: float = 0.025
    sigma: float = 0.02
    lambar: float = 0.0355
    kappa_lambda: float = 0.08
    nu_lambda: float = 0.067
    muZ: float = -0.4
    sigmaZ: float = 0.25
    rho: float = 0.012
    gamma: float = 3.0
    psi: float = 1.1
    phi: float = 2.6
    name: ClassVar[str] = "wachter"
    statenames: ClassVar = ("lambda",)
    solnames: ClassVar = ("p",)
    derivednames: ClassVar = ("mgf(Z, 1)", "mgf(Z, -gamma)", "mgf(Z, 1 - gamma)")

    def mgf(self, t: float) -> float:
        """Moment generating function of Normal(muZ, sigmaZ) (Distributions.jl `mgf(::Normal, t)`)."""
        return float(np.exp(t * self.muZ + self.sigmaZ ** 2 / 2 * t ** 2))

    def params(self) -> np.ndarray:
        base = RegisteredModel.params(self)
        return np.concatenate([base, [self.mgf(1.0), self.mgf(-self.gamma), self.mgf(1.0 - self.gamma)]])


@dataclass(frozen=True)
class HaddadModel(RegisteredModel):
    """examples/asset_pricing/haddad.jl:20-36."""
    mubar: float = 0.018
    vbar: float = 0.00052
    kappa_mu: float = 0.3
    nu_mu: float = 0.456
    kappa_v: float = 0.012
    nu_v: float = 0.00472
    alphabar: float = 1.2
    lam: float = 0.018
    rho: float = 0.0132
    gamma: float = 10.0
    psi: float = 1.5
    name: ClassVar[str] = "haddad"
    statenames: ClassVar = ("mu", "v")
    solnames: ClassVar = ("p",)


@dataclass(frozen=True)
class BoltonChenWangModel(RegisteredModel):
    """examples/corporate_finance/bolton_chen_wang.jl:19-29."""
    r: float = 0.06
    delta: float = 0.10
    A: float = 0.18
    sigma: float = 0.09
    theta: float = 1.5
    lam: float = 0.01
    l: float = 0.9
    gamma: float = 0.06
    phi: float = 0.01
    name: ClassVar[str] = "bolton_chen_wang"
    statenames: ClassVar = ("w",)
    solnames: ClassVar = ("v",)
    savednames: ClassVar = ("v", "vw", "vww", "w")


@dataclass(frozen=True)
class GomezModel(RegisteredModel):
    """examples/asset_pricing/gomez.jl:18-39."""
    gamma: float = 10.3
    psi: float = 0.05
    rho: float = 0.1
    rhoE: float = 0.022
    alphaE: float = 2.0
    nu: float = 0.1
    eta: float = 0.015
    delta: float = 0.025
    phi: float = 0.01
    piE: float = 0.09
    g: float = 0.02
    sigma: float = 0.04
    lam: float = 2.3
    tau: float = 0.0
    name: ClassVar[str] = "gomez"
    statenames: ClassVar = ("x",)
    solnames: ClassVar = ("pH",)
    savednames: ClassVar = ("x", "mux", "sigmax", "kappa", "r", "sigmaWH", "muWH", "sigmap", "p", "px", "sigmaR", "muR",
                            "muRlambda", "sigmaRlambda", "muErelative", "sigmaErelative", "muHrelative", "sigmaHrelative",
                            "muCE", "muxW", "sigmaxW", "pH", "muCH", "wedge", "xW", "sigmaWE", "sigmaCH", "alphaE", "pE", "mup")


@dataclass(frozen=True)
class AchdouHanLasryLionsMollModel_RiskyAsset(RegisteredModel):
    """examples/consumption_saving/consumption_saving_risky_asset.jl:19-30."""
    kappa_y: float = 0.1
    ybar: float = 1.0
    sigma_y: float = 0.07
    r: float = 0.03
    muR: float = 0.04
    sigmaR: float = 0.1
    rho: float = 0.05
    gamma: float = 2.0
    amin: float = 0.0
    amax: float = 300.0
    name: ClassVar[str] = "ahlm_risky_asset"
    statenames: ClassVar = ("y", "a")
    solnames: ClassVar = ("v",)
    savednames: ClassVar = ("v", "c", "k", "va", "vaa", "vy", "y", "a", "mua")


@dataclass(frozen=True)
class PointwiseCubic(RegisteredModel):
    """test/runtests.jl:422-431: `hump!(ydot, y) = y - y^3` at every grid point (a = k = 1)."""
    a: float = 1.0
    k: float = 1.0
    name: ClassVar[str] = "pointwise_cubic"
    statenames: ClassVar = ("x",)
    solnames: ClassVar = ("v",)


@dataclass(frozen=True)
class AffinePair(RegisteredModel):
    """test/runtests.jl:403-411 (`mixed_residual!`: a1, b1, a2, b2 = 1, -2, -1, 3) and :456-489
    (`residual!`: 2, -4, 1, -3): vt = a1 v + b1, pt = a2 p + b2 at every grid point."""
    a1: float = 1.0
    b1: float = -2.0
    a2: float = -1.0
    b2: float = 3.0
    name: ClassVar[str] = "affine_pair"
    statenames: ClassVar = ("x",)
    solnames: ClassVar = ("v", "p")


@dataclass(frozen=True)
class NeoclassicalGrowthModel(RegisteredModel):
    """examples/neoclassical_growth.jl:25-31."""
    A: float = 0.5
    alpha: float = 0.3
    delta: float = 0.05
    rho: float = 0.05
    gamma: float = 2.0
    name: ClassVar[str] = "neoclassical_growth"
    statenames: ClassVar = ("k",)
    solnames: ClassVar = ("v",)
    savednames: ClassVar = ("c", "muk")


@dataclass(frozen=True)
class GrowthTest(RegisteredModel):
    """test/runtests.jl:17-44."""
    A: float = 0.5
    alpha: float = 0.3
    delta: float = 0.05
    rho: float = 0.05
    gamma: float = 2.0
    name: ClassVar[str] = "growth_test"
    statenames: ClassVar = ("k",)
    solnames: ClassVar = ("v",)
    savednames: ClassVar = ("c", "muk")


@dataclass(frozen=True)
class LinearSum1D(RegisteredModel):
    """One-state version of test/runtests.jl:276-279 (vt = x - v: known answer v = x)."""
    name: ClassVar[str] = "linear_sum_1d"
    statenames: ClassVar = ("x",)
    solnames: ClassVar = ("v",)


@dataclass(frozen=True)
class LinearSum2D(RegisteredModel):
    """test/runtests.jl:276-279."""
    name: ClassVar[str] = "linear_sum_2d"
    statenames: ClassVar = ("x", "z")
    solnames: ClassVar = ("v",)


@dataclass(frozen=True)
class LinearSum3D(RegisteredModel):
    """Three-state version of test/runtests.jl:276-279 and :314-319 (known answer v = x + z + q)."""
    name: ClassVar[str] = "linear_sum_3d"
    statenames: ClassVar = ("x", "z", "q")
    solnames: ClassVar = ("v",)


@dataclass(frozen=True)
class LinearSum4D(RegisteredModel):
    """test/runtests.jl:294-320: four state variables, known answer v = x + z + q + r."""
    name: ClassVar[str] = "linear_sum_4d"
    statenames: ClassVar = ("x", "z", "q", "r")
    solnames: ClassVar = ("v",)


@dataclass(frozen=True)
class Hjb4dSynth(RegisteredModel):
    """Synthetic 4-state HJB (csrc/models.cuh: Hjb4dSynth): upwinded drifts, second differences and the three cross
    forms on an 81-point box, the stencil work behind test/runtests.jl:294-320."""
    kx: float = 0.2
    ky: float = 0.3
    kz: float = 0.4
    kw: float = 0.25
    sx: float = 0.1
    sy: float = 0.12
    sz: float = 0.08
    sw: float = 0.1
    cxw: float = 0.5
    cyz: float = -0.3
    cxy: float = 0.2
    rho: float = 0.05
    theta: float = 1e-3
    name: ClassVar[str] = "hjb4d_synth"
    statenames: ClassVar = ("x", "y", "z", "w")
    solnames: ClassVar = ("v",)


@dataclass(frozen=True)
class ForwardPrice(RegisteredModel):
    """test/runtests.jl:379-384."""
    name: ClassVar[str] = "forward_price"
    statenames: ClassVar = ("x",)
    solnames: ClassVar = ("v", "p")


@dataclass(frozen=True)
class VMinusTwo(RegisteredModel):
    """test/runtests.jl:227."""
    name: ClassVar[str] = "v_minus_two"
    statenames: ClassVar = ("x",)
    solnames: ClassVar = ("v",)


ALL_MODELS = (ReadmeLinear, CampbellCochraneModel, BansalYaronModel, AchdouHanLasryLionsMollModel_Diffusion,
              Hjb3dSynth, NeoclassicalGrowthModel, GrowthTest, LinearSum1D, LinearSum2D, LinearSum3D,
              ForwardPrice, VMinusTwo, Hjb2dRegimes, TuckmanVilaModel, DiTellaModel,
              AchdouHanLasryLionsMollModel_TwoStates, GarleanuPanageasModel, LelandModel,
              WangWangYangModel, LinearUpwindTest, HeKrishnamurthyModel, WachterModel, HaddadModel,
              BoltonChenWangModel, GomezModel, AchdouHanLasryLionsMollModel_RiskyAsset, PointwiseCubic,
              AffinePair, GarleanuPanageasLongRunRiskModel, LinearSum4D, Hjb4dSynth)

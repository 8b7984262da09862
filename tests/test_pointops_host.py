"""No-GPU checks of the product's per-point device code (stencils, functors, fused
row-wise colored-FD Jacobian), compiled for the host by tests/host_emu, against the oracle."""
import math

import ctypes

import numpy as np
import pytest

import econpdes_b200  # noqa: F401
from econpdes_b200 import models as PM
from oracle import econpdes_oracle as eo
from oracle import models as OM
from tests import emu_util as EU


def _cases():
    out = []
    g, gu = OM.readme_linear_problem(50)
    out.append(("readme", PM.ReadmeLinear(), OM.readme_linear(), g, gu))
    g, gu, prm = OM.campbell_cochrane_problem(sn=200)
    out.append(("cc", PM.CampbellCochraneModel(**prm), OM.campbell_cochrane(**prm), g, gu))
    g, gu, prm = OM.bansal_yaron_problem(17, 13)
    out.append(("by", PM.BansalYaronModel(**prm), OM.bansal_yaron(**prm), g, gu))
    g, gu, prm = OM.ahlm_diffusion_problem(7, 40)
    out.append(("ahlm", PM.AchdouHanLasryLionsMollModel_Diffusion(**prm), OM.ahlm_diffusion(**prm), g, gu))
    g, gu, prm = OM.hjb3d_problem(6)
    gu = {"v": 20.0 + np.zeros((6, 6, 6))}
    out.append(("hjb3d", PM.Hjb3dSynth(**prm), OM.hjb3d_synth(**prm), g, gu))
    g, gu, prm = OM.neoclassical_growth_problem(60)
    out.append(("growth", PM.NeoclassicalGrowthModel(**prm), OM.neoclassical_growth(**prm), g, gu))
    g, gu, _ = OM.growth_test_problem(40)
    out.append(("growth_test", PM.GrowthTest(), OM.growth_test_hjb(), g, gu))
    g = {"x": OM.julia_range(0, 1, 4), "z": OM.julia_range(0, 1, 5)}
    out.append(("lin2d", PM.LinearSum2D(), OM.linear_sum(("x", "z")), g, {"v": np.zeros((4, 5))}))
    g4 = {k: OM.julia_range(0, 1, n) for k, n in (("x", 2), ("z", 3), ("q", 2), ("r", 4))}
    out.append(("lin4d", PM.LinearSum4D(), OM.linear_sum(tuple("xzqr")), g4, {"v": np.zeros((2, 3, 2, 4))}))
    g, gu, prm = OM.hjb4d_problem((5, 4, 3, 4))
    rng4 = np.random.default_rng(5)
    gu = {"v": 20.0 + rng4.uniform(-1, 1, (5, 4, 3, 4))}
    out.append(("hjb4d", PM.Hjb4dSynth(**prm), OM.hjb4d_synth(**prm), g, gu))
    g = {"x": OM.julia_range(0.1, 1, 5)}
    out.append(("fwdprice", PM.ForwardPrice(), OM.forward_price(), g, {"v": np.ones(5), "p": np.ones(5) * 0.5}))
    g, gu, prm = OM.hjb2d_regimes_problem(9, 7)
    gu = {"v": 15.0 + np.zeros((9, 7)), "w": 16.0 + np.zeros((9, 7))}
    out.append(("regimes2d", PM.Hjb2dRegimes(**prm), OM.hjb2d_regimes(**prm), g, gu))
    g, gu, prm = OM.di_tella_problem(7, 6)
    out.append(("ditella", PM.DiTellaModel(**prm), OM.di_tella(**prm), g, gu))
    g, gu, prm, _, _ = OM.garleanu_panageas_lrr_problem(7, 4, 5)
    rng = np.random.default_rng(11)
    gu = {k: v + (0.05 * rng.uniform(-1, 1, v.shape) if k in ("pA", "pB", "phi1", "phi2") else 0.002 * rng.uniform(-1, 1, v.shape))
          for k, v in gu.items()}
    out.append(("gp_lrr", PM.GarleanuPanageasLongRunRiskModel(**prm), OM.garleanu_panageas_lrr(**prm), g, gu))
    g, gu, prm = OM.ahlm_two_states_problem(60)
    out.append(("twoincome", PM.AchdouHanLasryLionsMollModel_TwoStates(**prm), OM.ahlm_two_states(**prm), g, gu))
    g, gu, prm = OM.garleanu_panageas_problem(40)
    gu = {k: v * (1.0 + 0.2 * np.linspace(0, 1, 40)) for k, v in gu.items()}
    out.append(("gp", PM.GarleanuPanageasModel(**prm), OM.garleanu_panageas(**prm), g, gu))
    g, gu, prm, _ = OM.wang_wang_yang_problem(80)
    out.append(("wwy", PM.WangWangYangModel(**prm), OM.wang_wang_yang(**prm), g, gu))
    g, gu, prm = OM.he_krishnamurthy_problem(40)
    gu = {"pS": 20.0 + 5.0 * np.linspace(0, 1, 40)}
    out.append(("hk", PM.HeKrishnamurthyModel(**prm), OM.he_krishnamurthy(**prm), g, gu))
    g, gu, prm = OM.wachter_problem(30)
    gu = {"p": 80.0 - 30.0 * g["lambda"]}
    out.append(("wachter", PM.WachterModel(**prm), OM.wachter(**prm), g, gu))
    g, gu, prm = OM.haddad_problem(13, 11)
    gu = {"p": 100.0 + 10.0 * np.linspace(0, 1, 13)[:, None] - 8.0 * np.linspace(0, 1, 11)[None, :]}
    out.append(("haddad", PM.HaddadModel(**prm), OM.haddad(**prm), g, gu))
    g, gu, prm, _ = OM.bolton_chen_wang_problem(40)
    gu = {"v": 100.0 + 120.0 * g["w"]}      # large level: the 1 % test perturbation then flips the upwind direction at some points
    out.append(("bcw", PM.BoltonChenWangModel(**prm), OM.bolton_chen_wang(**prm), g, gu))
    g, gu, prm = OM.gomez_problem(50)
    gu = {"pH": 6.0 + 25.0 * g["x"]}
    out.append(("gomez", PM.GomezModel(**prm), OM.gomez(**prm), g, gu))
    g, gu, prm, _ = OM.ahlm_risky_asset_problem(5, 24)
    out.append(("risky", PM.AchdouHanLasryLionsMollModel_RiskyAsset(**prm), OM.ahlm_risky_asset(**prm), g, gu))
    g = {"x": OM.julia_range(0, 1, 4)}
    out.append(("cubic", PM.PointwiseCubic(), OM.pointwise_cubic(), g, {"v": np.full(4, 0.6)}))
    out.append(("affine", PM.AffinePair(2.0, -4.0, 1.0, -3.0), OM.affine_pair(2.0, -4.0, 1.0, -3.0), g,
                {"v": np.ones(4), "p": 2.0 * np.ones(4)}))
    # three-argument body: the host emulation leaves the time slot at 0
    g, gu, _, prm = OM.tuckman_vila_problem(zn=50)
    tv = OM.tuckman_vila(**prm)
    out.append(("tuckman", PM.TuckmanVilaModel(**prm), lambda state, u: tv(state, u, 0.0), g,
                {"F": 100.0 * np.cos(np.linspace(0, 3, 50))}))
    return out


CASES = _cases()
POW_MODELS = ("ahlm", "growth", "growth_test", "tuckman", "ditella", "twoincome", "wwy")


def _perturbed(gu, seed=1):
    rng = np.random.default_rng(seed)
    return np.stack([v * (1 + 0.01 * rng.uniform(-1, 1, v.shape)) + 1e-3 * rng.uniform(-1, 1, v.shape)
                     for v in gu.values()], axis=-1)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_residual_matches_oracle(case):
    _, pm, om, g, gu = case
    grids = list(g.values())
    y = _perturbed(gu)
    prob = eo.Problem(om, tuple(g), grids, tuple(gu))
    want = prob.G(y.reshape(-1, order="F")).reshape(y.shape, order="F")
    got = EU.residual(pm, grids, y)
    if case[0] in POW_MODELS:
        # numpy's vectorised pow and libm's pow differ by an ulp on a few inputs
        np.testing.assert_allclose(got, want, rtol=0, atol=8 * np.finfo(float).eps * np.abs(want).max())
    else:
        # same operand order, IEEE +,-,*,/,sqrt only: bit-exact
        np.testing.assert_array_equal(got, want)


def test_residual_with_bc_matches_oracle():
    pm, om = PM.Hjb3dSynth(), OM.hjb3d_synth(**OM.HJB3D_DEFAULT)
    rng = np.random.default_rng(3)
    grids = [np.cumsum(rng.uniform(0.05, 0.2, s)) for s in (5, 4, 6)]
    y = rng.normal(size=(5, 4, 6, 1)) + 10
    bc = rng.normal(size=(5, 4, 6, 1))
    prob = eo.Problem(om, ("x", "y", "z"), grids, ("v",), bc)
    want = prob.G(y.reshape(-1, order="F")).reshape(y.shape, order="F")
    np.testing.assert_array_equal(EU.residual(pm, grids, y, bc=bc), want)


def test_risky_asset_with_its_boundary_rows_is_bit_exact():
    """consumption_saving_risky_asset.jl:199-200 sets bc for (v, a) only: the slots of the a-boundary rows are also
    what the y-direction stencils read at the grid corners (one slot per point, src/pdesolve.jl:312-335).  The host
    build of the functor and the scalar oracle call the same libm pow, so residual and saved outputs agree to the bit."""
    from econpdes_b200.pdesolve import _bc_array
    _, pm, om, g, gu = [c for c in CASES if c[0] == "risky"][0]
    _, _, _, bc = OM.ahlm_risky_asset_problem(5, 24)
    grids = list(g.values())
    S = tuple(len(x) for x in grids)
    bca = _bc_array(bc, ("v",), ("y", "a"), S)
    np.testing.assert_array_equal(bca, eo.bc_array(bc, ("v",), ("y", "a"), S))
    assert bca[0, 0, 0] != 0 and bca[-1, -1, 0] != 0 and (bca[:, 1:-1] == 0).all()
    for seed in (1, 2, 3):
        y = _perturbed(gu, seed=seed)
        prob = eo.Problem(om, tuple(g), grids, tuple(gu), bca)
        want = prob.G(y.reshape(-1, order="F")).reshape(y.shape, order="F")
        sv = prob.saved(y.reshape(-1, order="F"))
        got, gsv = EU.residual(pm, grids, y, bc=bca, saved=True)
        np.testing.assert_array_equal(got, want)
        for k, key in enumerate(pm.savednames):
            np.testing.assert_array_equal(gsv[..., k], sv[key], err_msg=key)
    # the three upwind cases and the zero-drift bisection are all exercised by the last perturbed field
    assert (sv["mua"] > 0).any() and (sv["mua"] < 0).any() and (sv["mua"] == 0).any()


def test_saved_outputs_match_oracle():
    _, pm, om, g, gu = [c for c in CASES if c[0] == "ahlm"][0]
    grids = list(g.values())
    y = _perturbed(gu)
    prob = eo.Problem(om, tuple(g), grids, tuple(gu))
    sv = prob.saved(y.reshape(-1, order="F"))
    _, got = EU.residual(pm, grids, y, saved=True)
    np.testing.assert_allclose(got[..., 0], sv["mua"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(got[..., 1], sv["c"], rtol=1e-14)


@pytest.mark.parametrize("name", ["hk", "bcw", "gomez"])
def test_saved_outputs_of_the_1d_examples_are_bit_exact(name):
    """Saved quantities of the He-Krishnamurthy (14), Bolton-Chen-Wang (4) and Gomez (30) examples, in the
    examples' order; +,-,*,/ only, so the host-compiled functor and the oracle agree to the bit."""
    _, pm, om, g, gu = [c for c in CASES if c[0] == name][0]
    grids = list(g.values())
    y = _perturbed(gu)
    prob = eo.Problem(om, tuple(g), grids, tuple(gu))
    sv = prob.saved(y.reshape(-1, order="F"))
    _, got = EU.residual(pm, grids, y, saved=True)
    assert tuple(sv) == pm.savednames
    for k, key in enumerate(pm.savednames):
        np.testing.assert_array_equal(got[..., k], sv[key], err_msg=key)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("delta", [1.0, math.inf])
def test_fused_jacobian_matches_colored_fd(case, delta):
    """The fused row-wise Jacobian equals FiniteDiff's C global sweeps entry for entry
    (up to the <=1 ulp x1 drift of the +eps/-eps restoration, SURVEY.md hard part 5)."""
    name, pm, om, g, gu = case
    grids = list(g.values())
    S = tuple(len(x) for x in grids)
    F = len(gu)
    y = _perturbed(gu)
    yvec = y.reshape(-1, order="F")
    ypost = _perturbed(gu, seed=2)
    ypvec = ypost.reshape(-1, order="F")
    prob = eo.Problem(om, tuple(g), grids, tuple(gu))
    cache = eo.JacobianCache(S, F)

    def R(v):
        with np.errstate(all="ignore"):
            return prob.G(v) - (ypvec - v) / delta
    nz_want, eps = eo.finite_difference_jacobian(R, yvec, cache)
    box = name in ("hjb3d", "regimes2d", "ditella", "gp_lrr", "hjb4d")
    planes, r0 = EU.jacobian_planes(pm, grids, y, ypost, delta, eps, box)
    if name in POW_MODELS:
        np.testing.assert_allclose(r0, R(yvec), rtol=0, atol=8 * np.finfo(float).eps * np.abs(r0).max())
    else:
        np.testing.assert_array_equal(r0, R(yvec))
    nz_got = EU.planes_to_csc(planes, S, F, box, cache.colptr, cache.rowval)
    scale = np.abs(nz_want).max()
    # drift of x1 by 1 ulp changes (f(x1)-f(x))/eps by ~ulp(y)*|J|/eps ~ 1e-8 relative
    np.testing.assert_allclose(nz_got, nz_want, rtol=0, atol=2e-6 * scale)
    if not box and pm.ndims > 1:
        # corners of the box are exact structural zeros for star models
        full, _ = EU.jacobian_planes(pm, grids, y, ypost, delta, eps, True)
        nz_full = EU.planes_to_csc(full, S, F, True, cache.colptr, cache.rowval)
        np.testing.assert_array_equal(nz_full, nz_got)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reciprocal_quotients_give_the_same_bits(case):
    """stencil.cuh forms the stencils' quotients either by IEEE division (RECIP 0) or from
    reciprocals of the grid-only divisors with Markstein's correction step (RECIP 1: per-axis,
    RECIP 2: also the cross-derivative divisors).  All modes must agree to the last bit, on
    uniform and non-uniform grids and with non-zero boundary derivatives."""
    _, pm, om, g, gu = case
    grids = list(g.values())
    # non-uniform variant of the grid: monotone warp of each axis
    warped = [x[0] + (x[-1] - x[0]) * ((x - x[0]) / (x[-1] - x[0])) ** 1.7 for x in grids]
    lib = EU.lib()
    for gs in (grids, warped):
        for seed in (1, 2, 3):
            y = _perturbed(gu, seed)
            bc = 0.3 * np.random.default_rng(seed).uniform(-1, 1, y.shape)
            outs = []
            for mode in (0, 1, 2):
                lib.emu_set_recip(mode)
                outs.append(EU.residual(pm, gs, y, bc=bc))
            lib.emu_set_recip(0)
            np.testing.assert_array_equal(outs[0], outs[1])
            np.testing.assert_array_equal(outs[0], outs[2])


def test_reciprocal_jacobian_planes_give_the_same_bits():
    g, gu, prm = OM.hjb3d_problem(7)
    pm = PM.Hjb3dSynth(**prm)
    grids = list(g.values())
    y = _perturbed({"v": 20.0 + np.zeros((7, 7, 7))})
    ypost = _perturbed({"v": 20.0 + np.zeros((7, 7, 7))}, seed=5)
    eps = np.full(27, 3e-8)
    lib = EU.lib()
    outs = []
    for mode in (0, 2):
        lib.emu_set_recip(mode)
        outs.append(EU.jacobian_planes(pm, grids, y, ypost, 1.7, eps, True))
    lib.emu_set_recip(0)
    for a, b in zip(outs[0], outs[1]):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("bad", [np.inf, -np.inf, np.nan, 1e200, -3e305])
def test_reciprocal_quotients_with_non_finite_and_huge_values(bad):
    """Inf / NaN / near-overflow iterates (a diverging Newton step) must come out of the
    reciprocal path exactly as IEEE division returns them: same Infs, same NaN positions."""
    g, gu, prm = OM.hjb3d_problem(6)
    pm = PM.Hjb3dSynth(**prm)
    grids = list(g.values())
    y = _perturbed({"v": 20.0 + np.zeros((6, 6, 6))})
    y[2, 3, 1, 0] = bad
    y[0, 0, 5, 0] = bad
    lib = EU.lib()
    outs = []
    for mode in (0, 1, 2):
        lib.emu_set_recip(mode)
        with np.errstate(all="ignore"):
            outs.append(EU.residual(pm, grids, y))
    lib.emu_set_recip(0)
    np.testing.assert_array_equal(outs[0], outs[1])
    np.testing.assert_array_equal(outs[0], outs[2])
    g1, gu1, prm1 = OM.campbell_cochrane_problem(sn=60)
    pm1 = PM.CampbellCochraneModel(**prm1)
    y1 = _perturbed(gu1)
    y1[17, 0] = bad
    outs = []
    for mode in (0, 1):
        lib.emu_set_recip(mode)
        with np.errstate(all="ignore"):
            outs.append(EU.residual(pm1, list(g1.values()), y1))
    lib.emu_set_recip(0)
    np.testing.assert_array_equal(outs[0], outs[1])


def test_shared_reciprocal_in_functors_matches_ieee_for_degenerate_values():
    """The asset-pricing functors divide by the unknown p through one reciprocal (mdiv); p = 0,
    Inf or NaN must give what the reference's IEEE divisions give (oracle: numpy `/`)."""
    g, gu, prm = OM.campbell_cochrane_problem(sn=40)
    pm, om = PM.CampbellCochraneModel(**prm), OM.campbell_cochrane(**prm)
    grids = list(g.values())
    for bad in (0.0, np.inf, -np.inf, np.nan, 1e308):
        y = _perturbed(gu)
        y[11, 0] = bad
        prob = eo.Problem(om, tuple(g), grids, tuple(gu))
        with np.errstate(all="ignore"):
            want = prob.G(y.reshape(-1, order="F")).reshape(y.shape, order="F")
            got = EU.residual(pm, grids, y)
        np.testing.assert_array_equal(got, want)


def test_fast_division_by_grid_sizes():
    """unravel divides the flat point index by the grid sizes through a multiply-high (FastDiv)."""
    lib = EU.lib()
    lib.emu_fastdiv_mismatches.restype = ctypes.c_longlong
    lib.emu_fastdiv_mismatches.argtypes = [ctypes.c_uint, ctypes.c_uint]
    for d in (1, 2, 3, 5, 7, 9, 50, 64, 255, 256, 257, 998, 1000, 2048, 4096, 65535, 65537, 1 << 20, (1 << 31) - 1):
        assert lib.emu_fastdiv_mismatches(d, 200000) == 0, d


# ----------------------------------------------------------------------------------------------
# property test: stencils on random non-uniform grids of every small shape, with boundary values
# ----------------------------------------------------------------------------------------------
from hypothesis import given, settings, strategies as st  # noqa: E402

_PROP_MODELS = {
    1: (PM.ReadmeLinear(), OM.readme_linear(), ("x",), ("v",)),
    2: (PM.Hjb2dRegimes(**OM.HJB2D_REGIMES_DEFAULT), OM.hjb2d_regimes(**OM.HJB2D_REGIMES_DEFAULT), ("x", "y"), ("v", "w")),
    3: (PM.Hjb3dSynth(**OM.HJB3D_DEFAULT), OM.hjb3d_synth(**OM.HJB3D_DEFAULT), ("x", "y", "z"), ("v",)),
}


@settings(max_examples=60, deadline=None)
@given(shape=st.lists(st.integers(min_value=2, max_value=5), min_size=1, max_size=3), seed=st.integers(0, 2 ** 31 - 1),
       with_bc=st.booleans())
def test_stencils_on_random_small_grids_are_bit_exact(shape, seed, with_bc):
    """derivatives.jl:39-159 on every small shape (2 points per dimension included: both neighbours are
    boundary ghosts there), random spacings, random fields, random boundary values: the host build of the
    device stencils + functor equals the oracle to the last bit (upwind and cross stencils are exercised by
    the 2-D pair and the 3-D model)."""
    pm, om, statenames, solnames = _PROP_MODELS[len(shape)]
    rng = np.random.default_rng(seed)
    grids = [np.cumsum(rng.uniform(0.02, 0.3, s)) for s in shape]
    F = len(solnames)
    y = rng.normal(size=tuple(shape) + (F,)) + 12.0
    bc = rng.normal(size=tuple(shape) + (F,)) if with_bc else None
    prob = eo.Problem(om, statenames, grids, solnames, bc)
    want = prob.G(y.reshape(-1, order="F")).reshape(y.shape, order="F")
    np.testing.assert_array_equal(EU.residual(pm, grids, y, bc=bc), want)


@settings(max_examples=40, deadline=None)
@given(shape=st.lists(st.integers(min_value=2, max_value=5), min_size=1, max_size=3), seed=st.integers(0, 2 ** 31 - 1),
       delta=st.sampled_from([1.0, 0.05, math.inf]))
def test_fused_jacobian_on_random_small_grids(shape, seed, delta):
    """The fused row-wise colored-FD Jacobian against FiniteDiff's global sweeps on every small shape: with 2 points
    in a dimension there are only two colors along it (coloring.jl:52-70) and both stencil neighbours are ghosts."""
    pm, om, statenames, solnames = _PROP_MODELS[len(shape)]
    rng = np.random.default_rng(seed)
    grids = [np.cumsum(rng.uniform(0.05, 0.3, s)) for s in shape]
    S, F = tuple(shape), len(solnames)
    y = rng.normal(size=S + (F,)) + 12.0
    ypost = y + 0.1 * rng.normal(size=y.shape)
    bc = rng.normal(size=S + (F,))
    prob = eo.Problem(om, statenames, grids, solnames, bc)
    cache = eo.JacobianCache(S, F)
    yvec, ypvec = y.reshape(-1, order="F"), ypost.reshape(-1, order="F")

    def R(v):
        with np.errstate(all="ignore"):
            return prob.G(v) - (ypvec - v) / delta
    nz_want, eps = eo.finite_difference_jacobian(R, yvec, cache)
    box = len(shape) > 1
    planes, r0 = EU.jacobian_planes(pm, grids, y, ypost, delta, eps, box, bc=bc)
    np.testing.assert_array_equal(r0, R(yvec))
    nz_got = EU.planes_to_csc(planes, S, F, box, cache.colptr, cache.rowval)
    np.testing.assert_allclose(nz_got, nz_want, rtol=0, atol=2e-6 * np.abs(nz_want).max())


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_row_terms_match_g_helper_and_minmax_mcp(mode):
    """`apply_row_terms` against `G_helper!` (finiteschemesolve.jl:158-162: R = G - .!alg .* (ypost - y) ./ Δ with a
    signed per-row Δ, ±Inf included) and the min-max MCP residual (:197-203: min(max(R, y - ub), y - lb), infinite
    bounds, NaN residuals propagating like Julia's max/min), for the dividing and the two reciprocal forms."""
    rng = np.random.default_rng(11)
    n = 4000
    g = rng.normal(size=n) * 10.0 ** rng.integers(-6, 6, n)
    y = rng.normal(size=n) * 10.0 ** rng.integers(-3, 3, n)
    ypost = y + rng.normal(size=n) * 10.0 ** rng.integers(-8, 2, n)
    delta = rng.choice([1.0, -1.0, 0.1, 1e-9, 3.7e5, -2.5e-3, math.inf, -math.inf], size=n) * rng.uniform(0.5, 2.0, n)
    alg = rng.random(n) < 0.2
    lb = np.where(rng.random(n) < 0.3, -np.inf, y - rng.uniform(-1, 2, n))
    ub = np.where(rng.random(n) < 0.3, np.inf, lb + rng.uniform(0, 3, n))
    g[::97] = np.nan
    g[5::191] = np.inf
    L = EU.lib()
    dp = ctypes.POINTER(ctypes.c_double)
    L.emu_row_terms.argtypes = [ctypes.c_longlong, dp, dp, dp, dp, ctypes.POINTER(ctypes.c_ubyte), dp, dp,
                                ctypes.c_int, ctypes.c_int, dp]
    L.emu_row_terms.restype = None
    a8 = np.ascontiguousarray(alg, dtype=np.uint8)

    def run(bounds):
        out = np.empty(n)
        L.emu_row_terms(n, EU._p(g), EU._p(y), EU._p(ypost), EU._p(delta), a8.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                        EU._p(lb), EU._p(ub), int(bounds), mode, EU._p(out))
        return out
    with np.errstate(all="ignore"):
        want = g - np.where(alg, 0.0, (ypost - y) / delta)
        np.testing.assert_array_equal(run(False), want)
        np.testing.assert_array_equal(run(True), np.minimum(np.maximum(want, y - ub), y - lb))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_wild_fields_give_the_same_bits_and_the_same_nans(case):
    """Newton's off-path iterates: fields of mixed sign and scale, exact zeros included.  The functor must produce a NaN
    or an infinity exactly where the reference body does (that decides "model returned NaN" rejections) and the same
    bits elsewhere (to a few ulp where `pow`/`log` enter).  This test caught numpy's `x ** 3` (pow) standing in for
    Julia's `x^3` (= x*x*x) in one oracle body."""
    name, pm, om, g, gu = case
    grids = list(g.values())
    shape = next(iter(gu.values())).shape
    for seed in range(8):
        rng = np.random.default_rng(100 + seed)
        y = np.stack([rng.normal(size=shape) * 10.0 ** rng.integers(-2, 3) for _ in gu], axis=-1)
        if seed % 2:
            y[tuple(rng.integers(0, s) for s in shape) + (0,)] = 0.0
        # every other seed also random boundary values for every (unknown, point) slot (src/pdesolve.jl:312-335)
        bc = rng.normal(size=y.shape) * 10.0 ** rng.integers(-2, 2) if seed >= 4 else None
        prob = eo.Problem(om, tuple(g), grids, tuple(gu), bc)
        with np.errstate(all="ignore"):
            want = prob.G(y.reshape(-1, order="F")).reshape(y.shape, order="F")
        got = EU.residual(pm, grids, y, bc=bc)
        np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
        fin = np.isfinite(want)
        np.testing.assert_array_equal(got[~fin], want[~fin])
        if name in POW_MODELS + ("risky",):
            scale = np.abs(want[fin]).max() if fin.any() else 1.0
            np.testing.assert_allclose(got[fin], want[fin], rtol=1e-9, atol=1e-9 * scale)
        else:
            np.testing.assert_array_equal(got[fin], want[fin])


@pytest.mark.parametrize("case", [c for c in CASES if c[1].savednames], ids=[c[0] for c in CASES if c[1].savednames])
def test_wild_fields_saved_outputs(case):
    """Same off-path fields for the saved outputs (`_fill_saved!`, src/pdesolve.jl:424-433): same NaN positions,
    same bits (a few ulp where `pow` enters)."""
    name, pm, om, g, gu = case
    grids = list(g.values())
    shape = next(iter(gu.values())).shape
    prob = eo.Problem(om, tuple(g), grids, tuple(gu))
    for seed in range(6):
        rng = np.random.default_rng(500 + seed)
        y = np.stack([rng.normal(size=shape) * 10.0 ** rng.integers(-2, 3) for _ in gu], axis=-1)
        with np.errstate(all="ignore"):
            sv = prob.saved(y.reshape(-1, order="F"))
        _, got = EU.residual(pm, grids, y, saved=True)
        for k, key in enumerate(pm.savednames):
            a, b = got[..., k], sv[key]
            np.testing.assert_array_equal(np.isnan(a), np.isnan(b), err_msg=key)
            fin = np.isfinite(b)
            np.testing.assert_array_equal(a[~fin], b[~fin], err_msg=key)
            if name in POW_MODELS + ("risky",):
                scale = np.abs(b[fin]).max() if fin.any() else 1.0
                np.testing.assert_allclose(a[fin], b[fin], rtol=1e-9, atol=1e-9 * scale, err_msg=key)
            else:
                np.testing.assert_array_equal(a[fin], b[fin], err_msg=key)


@pytest.mark.parametrize("n", [2, 3, 7, 100, 900, 901])
def test_device_grid_generator_matches_julia_range(n):
    """range_point (csrc/stencil.cuh), the grid generator of the compact batched inputs, against the oracle's
    exact-rational restatement of Julia's `range` (oracle/models.py:julia_range): same bits, including ranges
    that cross zero, tiny / huge endpoints and the Campbell-Cochrane endpoints (campbell_cochrane.jl:44-46);
    the vectorised numpy form the synthetic sweeps are built with (examples.julia_range_rows) agrees too."""
    from econpdes_b200 import examples as EX
    rng = np.random.default_rng(n)
    a = np.concatenate([rng.uniform(-400, 10, 60), rng.uniform(-1, 0, 20), [-300.0, 0.0, -1e-300, 1e200, -3.0]])
    b = np.concatenate([a[:60] + rng.uniform(0.1, 500, 60), rng.uniform(0, 1, 20),
                        [-2.6347180103190485, 0.08, 1e-300, 3e200, 7.0]])
    got = EU.range_rows(a, b, n)
    want = np.stack([OM.julia_range(x, y, n) for x, y in zip(a, b)])
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(EX.julia_range_rows(a, b, n), want)


def test_sweep_grids_equal_the_oracle_grids():
    """The synthetic C4 sweep (examples.campbell_cochrane_sweep, explicit or compact form) describes the grids the
    oracle builds for the same parameter draws (oracle/models.py:campbell_cochrane_grid), bit for bit."""
    from econpdes_b200 import examples as EX
    for i in (0, 3, 16383):
        prm = OM.campbell_cochrane_sweep_params(i)
        want = OM.campbell_cochrane_grid(sn=1000, **prm)
        params, grids, guess = EX.campbell_cochrane_sweep(i, 1)
        np.testing.assert_array_equal(grids[0], want)
        np.testing.assert_array_equal(params[0], [prm[k] for k in ("mu", "sigma", "gamma", "rho", "kappa_s", "b")])
        p2, lo, pts, n_range, fill = EX.campbell_cochrane_sweep_compact(i, 1)
        assert n_range == 899 and fill == 1.0 and pts.shape == (1, 99)
        rebuilt = np.concatenate([EU.range_rows(lo, pts[:, 0], n_range + 1)[0, :n_range], pts[0]])
        np.testing.assert_array_equal(rebuilt, want)

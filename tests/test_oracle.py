"""Pins the CPU oracle (oracle/) against everything the reference publishes for the hot
path: the two iteration traces, the closed-form answers and structural facts of
test/runtests.jl, and the 11-outer-iteration growth run of examples/details.tex:173."""
import math

import numpy as np
import pytest

from oracle import econpdes_oracle as eo
from oracle import models as M

INF = math.inf


def _trace(res):
    return [("%8.2e" % d, "%8.2e" % r) for (_, d, r, st, _) in res.log.rows if st == eo.ACCEPTED]


def test_readme_quickstart_trace():
    # README.md:72-82
    g, gu = M.readme_linear_problem()
    res = eo.pdesolve(M.readme_linear(), g, gu)
    tr = _trace(res)
    assert [t[0] for t in tr] == ["1.00e+00", "1.09e+01", "1.85e+03", "1.76e+08"]
    assert [t[1] for t in tr[:3]] == ["7.51e-02", "4.45e-02", "4.67e-04"]
    # last row is pure round-off of the linear solve: published 1.22e-10
    assert abs(res.residual_norm - 1.22e-10) < 0.03e-10
    assert res.converged and res.log.n_outer == 4 and res.tolerance == eo.SQRT_EPS
    assert "%.2e" % res.tolerance == "1.49e-08"


def test_campbell_cochrane_trace():
    # docs/src/solver.md:66-73, result block :96-101
    g, gu, prm = M.campbell_cochrane_problem()
    assert g["s"].shape == (998,)
    res = eo.pdesolve(M.campbell_cochrane(**prm), g, gu)
    assert _trace(res) == [("1.00e+00", "2.51e-02"), ("1.12e+01", "1.21e-02"),
                           ("2.33e+03", "8.01e-05"), ("3.50e+08", "3.45e-12")]
    assert res.converged and res.log.n_newton == 8


def test_growth_example_eleven_outer_iterations():
    # examples/details.tex:173 -- pins the FiniteDiff per-color epsilon rule (a per-column
    # epsilon takes 15 outer iterations)
    g, gu, prm = M.neoclassical_growth_problem(1000)
    res = eo.pdesolve(M.neoclassical_growth(**prm), g, gu)
    assert res.converged and res.log.n_outer == 11
    assert sum(1 for r in res.log.rows if r[3] != eo.ACCEPTED) == 3


def test_growth_testset_shape_properties():
    # test/runtests.jl:46-65
    g, gu, kbar = M.growth_test_problem(200)
    res = eo.pdesolve(M.growth_test_hjb(), g, gu)
    v, c, muk = res.solution["v"], res.saved["c"], res.saved["muk"]
    assert res.residual_norm <= 1e-6
    assert np.all(np.diff(v) > 0) and np.all(np.diff(np.diff(v)) < 1e-8) and np.all(np.diff(c) > 0)
    icross = int(np.argmax(muk < 0))
    assert abs(g["k"][icross] - kbar) < 0.05 * kbar
    assert np.array_equal(res.saved["v"], v)


def test_mixed_grid_known_answer_2d():
    # test/runtests.jl:269-281
    grid = {"x": M.julia_range(0, 1, 4), "z": M.julia_range(0, 1, 5)}
    res = eo.pdesolve(M.linear_sum(("x", "z")), grid, {"v": np.zeros((4, 5))}, delta=INF)
    assert res.residual_norm <= 1e-8
    np.testing.assert_allclose(res.solution["v"], grid["x"][:, None] + grid["z"][None, :], rtol=1e-12)
    assert res.saved == {}


def test_nd_coloring_and_4d_known_answer():
    # test/runtests.jl:294-320
    S = (2, 2, 2, 2)
    colptr, rowval = eo.local_stencil_jacobian(S, 1)
    colors = eo.stencil_colors(S, 1)
    assert colptr.shape[0] == 17 and colors.max() == 16
    _check_distance2(colptr, rowval, colors)
    g4 = {k: M.julia_range(0, 1, 2) for k in "xzqr"}
    res = eo.pdesolve(M.linear_sum(tuple("xzqr")), g4, {"v": np.zeros(S)}, delta=INF)
    x = g4["x"]
    want = x[:, None, None, None] + x[None, :, None, None] + x[None, None, :, None] + x[None, None, None, :]
    assert res.converged
    np.testing.assert_allclose(res.solution["v"], want, rtol=1e-12)


def test_signed_per_function_delta():
    # test/runtests.jl:375-394
    g = {"x": M.julia_range(0.1, 1.0, 5)}
    guess = {"v": np.zeros(5), "p": np.zeros(5)}
    res = eo.pdesolve(M.forward_price(), g, guess, delta={"v": 1.0, "p": -1.0})
    assert res.converged
    np.testing.assert_allclose(res.solution["v"], 2 * g["x"], atol=1e-6)
    np.testing.assert_allclose(res.solution["p"], g["x"], atol=1e-6)
    res = eo.pdesolve(M.forward_price(), g, guess, delta={"v": INF, "p": -1.0})
    assert res.converged
    np.testing.assert_allclose(res.solution["p"], g["x"], atol=1e-6)
    with pytest.raises(ValueError):
        eo.pdesolve(M.forward_price(), g, guess, delta={"v": 1.0, "p": 0.0})


def _scalar_cache(n):
    return eo.JacobianCache((n,), 1, colptr=np.arange(1, n + 2), rowval=np.arange(1, n + 1),
                            colors=np.ones(n, dtype=np.int64))


def test_vector_and_negative_delta_finiteschemesolve():
    # test/runtests.jl:404-416
    G = lambda y: np.array([y[0] - 2.0, 3.0 - y[1]])
    y, rn, _ = eo.finiteschemesolve(G, np.zeros(2), _scalar_cache(2), delta=np.array([1.0, -1.0]))
    np.testing.assert_allclose(y, [2.0, 3.0], atol=1e-6)
    assert rn <= 1e-6
    y, rn, _ = eo.finiteschemesolve(lambda y: 3.0 - y, np.zeros(1), _scalar_cache(1), delta=-1.0)
    np.testing.assert_allclose(y, [3.0], atol=1e-6)
    with pytest.raises(ValueError):
        eo.finiteschemesolve(G, np.zeros(2), _scalar_cache(2), delta=np.array([1.0]))
    with pytest.raises(ValueError):
        eo.finiteschemesolve(G, np.zeros(2), _scalar_cache(2), delta=0.0)


def test_residual_growth_guard():
    # test/runtests.jl:422-431
    hump = lambda y: y - y ** 3
    _, rn, _ = eo.finiteschemesolve(hump, np.array([0.6]), _scalar_cache(1), delta=0.1)
    assert rn <= eo.SQRT_EPS
    _, rn, log = eo.finiteschemesolve(hump, np.array([0.6]), _scalar_cache(1), delta=0.1, max_residual_growth=1.0)
    assert not (rn <= eo.SQRT_EPS) and log.exit_reason == eo.EXIT_BELOW_MIN_DELTA


def test_mcp_lower_bound_known_answer():
    # test/runtests.jl:225-232 and :440-450
    g = {"x": M.julia_range(0, 1, 5)}
    res = eo.pdesolve(M.v_minus_two(), g, {"v": np.zeros(5)}, delta=INF, lower_bound={"v": np.full(5, 3.0)})
    np.testing.assert_allclose(res.solution["v"], 3.0)
    y, rn, _ = eo.finiteschemesolve(lambda y: y - 2.0, np.array([0.5]), _scalar_cache(1), delta=INF,
                                    lower_bound=np.array([0.0]), upper_bound=np.array([1.0]))
    assert abs(y[0] - 1.0) < 1e-6 and rn <= 1e-6


def test_nan_at_guess_raises():
    # src/finiteschemesolve.jl:372
    with pytest.raises(ValueError, match="NaN"):
        eo.finiteschemesolve(lambda y: y * np.nan, np.ones(1), _scalar_cache(1))


def test_already_zero():
    # src/finiteschemesolve.jl:373-376
    y, rn, log = eo.finiteschemesolve(lambda y: y * 0.0, np.ones(1), _scalar_cache(1))
    assert log.exit_reason == eo.EXIT_ALREADY_ZERO and rn == 0.0


def test_maxiters_one_does_not_converge():
    # test/runtests.jl:162-192 (converged flag false with maxiters = 1)
    g, gu, _ = M.growth_test_problem(50)
    res = eo.pdesolve(M.growth_test_hjb(), g, gu, maxiters=1)
    assert not res.converged and res.log.exit_reason == eo.EXIT_MAXITERS


def _check_distance2(colptr, rowval, colors):
    n = colptr.shape[0] - 1
    col_of = np.repeat(np.arange(n), np.diff(colptr))
    order = np.argsort(rowval, kind="stable")
    rows_sorted, cols_sorted = rowval[order], col_of[order]
    bounds = np.flatnonzero(np.diff(rows_sorted)) + 1
    for cols in np.split(cols_sorted, bounds):
        cc = colors[cols]
        assert len(set(cc.tolist())) == len(cc)


def test_2d_multifunction_sparsity_520():
    # test/runtests.jl:491-502
    colptr, rowval = eo.local_stencil_jacobian((4, 5), 2)
    assert colptr.shape[0] - 1 == 40 and rowval.shape[0] == 520


@pytest.mark.parametrize("S,F", [((5,), 1), ((2,), 2), ((4, 5), 2), ((2, 5), 1), ((3, 4, 5), 2)])
def test_stencil_colors(S, F):
    # test/runtests.jl:509-524
    colptr, rowval = eo.local_stencil_jacobian(S, F)
    colors = eo.stencil_colors(S, F)
    _check_distance2(colptr, rowval, colors)
    assert colors.max() == F * int(np.prod([min(3, s) for s in S]))
    assert sorted(set(colors.tolist())) == list(range(1, colors.max() + 1))
    assert rowval.shape[0] == F * F * int(np.prod([3 * s - 2 for s in S]))
    # rows sorted inside every column, 1-based
    for j in range(colptr.shape[0] - 1):
        r = rowval[colptr[j] - 1:colptr[j + 1] - 1]
        assert np.all(np.diff(r) > 0) and r.min() >= 1


def test_pattern_matches_bruteforce():
    # literal transcription of coloring.jl:18-35 on a small 3-D case
    import itertools
    S, F = (3, 2, 4), 2
    l = int(np.prod(S))
    lin = lambda c: c[0] + S[0] * (c[1] + S[1] * c[2])
    ent = set()
    for fout in range(F):
        for ci in itertools.product(*[range(s) for s in S]):
            row = lin(ci) + fout * l
            for fin in range(F):
                for nb in itertools.product(*[range(max(c - 1, 0), min(c + 1, s - 1) + 1) for c, s in zip(ci, S)]):
                    ent.add((row, lin(nb) + fin * l))
    colptr, rowval = eo.local_stencil_jacobian(S, F)
    got = set()
    for j in range(l * F):
        for r in rowval[colptr[j] - 1:colptr[j + 1] - 1]:
            got.add((int(r) - 1, j))
    assert got == ent


@pytest.mark.parametrize("S,F", [((6,), 1), ((4, 5), 2), ((3, 4, 3), 1), ((2, 3), 1)])
def test_vectorised_stencils_match_pointwise(S, F):
    rng = np.random.default_rng(7)
    N = len(S)
    statenames = ["x", "z", "q"][:N]
    solnames = ["v", "w"][:F]
    grids = [np.cumsum(rng.uniform(0.1, 1.0, s)) for s in S]   # non-uniform
    y = rng.normal(size=S + (F,))
    bc = rng.normal(size=S + (F,))
    u = eo.differentiate(solnames, statenames, grids, y, bc)
    for icar in np.ndindex(*S):
        up = eo.differentiate_point(solnames, statenames, grids, y, icar, bc)
        assert set(up) == set(u)
        for key, val in up.items():
            assert u[key][icar] == val, (key, icar)
    nfields = 1 + 3 * N + 3 * N * (N - 1) // 2
    assert len(u) == F * nfields


def test_bc_array_quirk():
    # src/pdesolve.jl:312-335: one slot per (point, unknown); later dimensions overwrite
    out = eo.bc_array({"vx": (1.0, 2.0), "vz": (3.0, 4.0)}, ("v",), ("x", "z"), (3, 4))
    assert out[0, 1, 0] == 1.0 and out[2, 1, 0] == 2.0
    assert out[0, 0, 0] == 3.0 and out[2, 3, 0] == 4.0   # corners: z wins
    assert out[1, 1, 0] == 0.0
    with pytest.raises(ValueError):
        eo.bc_array({"vk": (0, 0)}, ("v",), ("x",), (3,))


def test_splitmix_sweep_params_are_in_range():
    p0 = M.campbell_cochrane_sweep_params(0)
    p1 = M.campbell_cochrane_sweep_params(1)
    assert p0 != p1
    assert 0.015 <= p0["mu"] <= 0.023 and 0.11 <= p0["kappa_s"] <= 0.17 and p0["b"] == 0.0
    # splitmix64 known answer: seed 0 -> first output 0xE220A8397B1DCDAF
    z = M.splitmix64_uniforms(0, 1)[0]
    assert z == (0xE220A8397B1DCDAF >> 11) * 2.0 ** -53


def test_sweep_problems_converge_like_the_default():
    # every sweep problem takes 4 outer steps / 8 Newton iterations
    for idx in (0, 1, 17):
        prm = M.campbell_cochrane_sweep_params(idx)
        g, gu, _ = M.campbell_cochrane_problem(prm)
        res = eo.pdesolve(M.campbell_cochrane(**prm), g, gu)
        assert res.converged and res.log.n_outer == 4 and res.log.n_newton == 8


def test_time_dependent_solve_reproduces_reference_test():
    """test/runtests.jl:100-148 ("Time-dependent solve"): growth model, 100 points, 10 times on
    [0, 100]; trailing time dimension, terminal condition kept bit for bit, every step converged,
    saved outputs per time."""
    from oracle import econpdes_oracle as eo2
    from oracle import models as OM2
    g, gu, _ = OM2.growth_test_problem(100)
    taus = OM2.julia_range(0, 100, 10)
    res = eo2.pdesolve_backward(OM2.growth_test_hjb(), g, gu, taus)
    assert res.solution["v"].shape == (100, 10)                              # :108
    np.testing.assert_array_equal(res.solution["v"][:, -1], gu["v"])         # :109
    assert res.residual_norm[1:].max() <= 1e-6                               # :110
    assert res.converged                                                     # :111
    assert res.saved["c"].shape == (100, 10)                                 # :112
    res2 = eo2.pdesolve_backward(OM2.growth_test_hjb(), g, gu, taus, inner_maxiters=20, inner_abstol=1e-8)
    assert res2.residual_norm[1:].max() <= 1e-6                              # :136-137
    # going far enough back in time the value approaches the stationary solution
    stat = eo2.pdesolve(OM2.growth_test_hjb(), g, gu)
    assert np.abs(res.solution["v"][:, 0] - stat.solution["v"]).max() < 0.05 * np.abs(stat.solution["v"]).max()


def _first_jacobian(pde, grid, guess, delta):
    """CSC Jacobian of the first implicit step at the guess (what the reference's checker receives)."""
    import scipy.sparse as sp
    names, sol = tuple(grid), tuple(guess)
    grids = [np.asarray(g, dtype=float) for g in grid.values()]
    S = tuple(len(g) for g in grids)
    prob = eo.Problem(pde, names, grids, sol)
    cache = eo.JacobianCache(S, len(sol))
    y = np.stack([np.asarray(guess[k], dtype=float) for k in sol], axis=-1).reshape(-1, order="F")

    def R(v):
        return prob.G(v) - (y - v) / delta
    nz, _ = eo.finite_difference_jacobian(R, y, cache)
    n = y.shape[0]
    return sp.csc_matrix((nz, cache.rowval - 1, cache.colptr - 1), shape=(n, n)), S


def test_diagnostics_reproduce_the_reference_testsets():
    """test/runtests.jl:322-372: the restated checks (oracle.jacobian_diagnostics) flag exactly the cases the
    reference's "Monotonicity diagnostics" and "Sign-convention diagnostic" testsets expect."""
    R = M.julia_range
    # wrong upwind direction, 2-point grid, one-shot solve -> a non-monotone entry
    J, S = _first_jacobian(M.linear_upwind_test(-1.0, 1.0, -1.0), {"x": R(0.0, 1.0, 2)}, {"V": R(1.0, 2.0, 2)}, math.inf)
    assert eo.jacobian_diagnostics(J, math.inf, None, S, 1)[2] > 0
    # correct upwind direction -> none
    J, S = _first_jacobian(M.linear_upwind_test(-1.0, 1.0, 1.0), {"x": R(0.0, 1.0, 3)}, {"V": R(1.0, 2.0, 3)}, math.inf)
    assert eo.jacobian_diagnostics(J, math.inf, None, S, 1)[2] == 0
    # flipped sign of the time derivative, Δ = 0.5 -> negative residual diagonal at every grid point
    J, S = _first_jacobian(M.linear_upwind_test(1.0, -1.0, 1.0), {"x": R(0.0, 1.0, 5)}, {"V": R(1.0, 2.0, 5)}, 0.5)
    nchk, nwrong, _, _ = eo.jacobian_diagnostics(J, 0.5, None, S, 1)
    assert nchk[0] == 5 and nwrong[0] == 5
    # correct convention, default Δ -> nothing
    J, S = _first_jacobian(M.linear_upwind_test(-1.0, -1.0, 1.0), {"x": R(0.0, 1.0, 5)}, {"V": R(1.0, 2.0, 5)}, 1.0)
    nchk, nwrong, cnt, _ = eo.jacobian_diagnostics(J, 1.0, None, S, 1)
    assert nchk[0] == 5 and nwrong[0] == 0 and cnt == 0


def test_example_model_fixtures_are_what_the_oracle_produces():
    """tests/golden/ex_*.npz (make_golden_examples.py): the committed outputs of the example models on the examples'
    own grids are reproduced by the current oracle and model restatements -- same iteration log, solution and
    saved outputs -- so a change in either shows up here before it reaches the device parity tests."""
    import os
    from tests.golden import make_golden_examples as G
    n = 0
    for name, pde, g, gu, prm, kw in G.cases():
        gold = np.load(os.path.join(G.HERE, name + ".npz"))
        log = gold["log"]
        if name == "ex_garleanu_panageas_lrr_20x5x5":
            # two minutes of SuperLU for the whole first stage: re-run its first three outer steps only (one rejected
            # through the singular-Jacobian retry, two accepted) and compare the log rows
            res = eo.pdesolve(pde, g, gu, **dict(kw, maxiters=3))
            assert [r[3] for r in res.log.rows] == [int(s) for s in log[:3, 3]]
            assert [r[4] for r in res.log.rows] == [int(s) for s in log[:3, 4]]
            np.testing.assert_allclose([r[2] for r in res.log.rows[1:]], log[1:3, 2], rtol=1e-9)
            n += 1
            continue
        res = eo.pdesolve(pde, g, gu, **kw)
        assert bool(gold["converged"]) == res.converged and res.converged, name
        assert [int(s) for s in log[:, 3]] == [r[3] for r in res.log.rows], name
        assert [int(s) for s in log[:, 4]] == [r[4] for r in res.log.rows], name
        for k, v in res.solution.items():
            np.testing.assert_allclose(v, gold["sol_" + k], rtol=1e-9, atol=1e-12, err_msg=name)
        for k, v in res.saved.items():
            if k not in res.solution:
                np.testing.assert_allclose(v, gold["saved_" + k], rtol=1e-7, atol=1e-9, err_msg=f"{name}.{k}")
        n += 1
    assert n == 11


def test_example_models_economic_sanity():
    """Qualitative facts the examples' texts state, checked on the oracle solutions (the reference pins no numbers)."""
    g, gu, prm = M.he_krishnamurthy_problem()
    res = eo.pdesolve(M.he_krishnamurthy(**prm), g, gu)
    kappa = res.saved["kappa"]
    assert kappa[0] > 10 * kappa[-1]                                    # price of risk is elevated when specialists are poor
    np.testing.assert_allclose(kappa[-1], prm["gamma"] * prm["sigma"], rtol=2e-2)
    g, gu, prm = M.wachter_problem()
    res = eo.pdesolve(M.wachter(**prm), g, gu)
    assert (np.diff(res.solution["p"]) < 0).all()                        # psi > 1: valuations fall with disaster risk
    g, gu, prm = M.gomez_problem()
    res = eo.pdesolve(M.gomez(**prm), g, gu)
    sR = res.saved["sigmaR"]
    np.testing.assert_allclose(sR[[0, -1]], prm["sigma"], rtol=1e-12)    # no amplification with one type of agent
    assert sR[1:-1].max() > 1.5 * prm["sigma"]                           # excess volatility in between
    g, gu, prm, bc = M.ahlm_risky_asset_problem()
    res = eo.pdesolve(M.ahlm_risky_asset(**prm), g, gu, bc=bc)
    k, a = res.saved["k"], res.saved["a"]
    assert (k >= 0).all() and (k <= np.maximum(a - prm["amin"], 0) + 1e-12).all() and (k[:, 0] == 0).all()
    # consumption rises with wealth; the y-edge rows are left out: their corners read the (v, a) boundary values in the
    # y-direction stencils too (one bc slot per point, src/pdesolve.jl:312-335)
    assert (np.diff(res.saved["c"][1:-1], axis=1) > 0).all()


def test_no_config_takes_the_singular_jacobian_retry():
    """NLsolve retries a singular Newton system with regularised normal equations (restated in
    econpdes_oracle.nlsolve_newton); the device path rejects such a step as "singular Jacobian" instead.  The two
    only differ if some solve gets there: none of the configs and example fixtures does (the full-size C2 64 x 4096,
    C3 256^2 and C5 24^3 oracle solves were checked the same way when their fixtures were generated: 0 retries)."""
    eo.STATS["singular_retries"] = 0
    g, gu = M.readme_linear_problem(50)
    eo.pdesolve(M.readme_linear(), g, gu)
    g, gu, prm = M.campbell_cochrane_problem()
    eo.pdesolve(M.campbell_cochrane(**prm), g, gu)
    g, gu, prm = M.bansal_yaron_problem(30, 30)
    eo.pdesolve(M.bansal_yaron(**prm), g, gu)
    g, gu, prm = M.ahlm_diffusion_problem(10, 100)
    eo.pdesolve(M.ahlm_diffusion(**prm), g, gu)
    g, gu, prm = M.hjb3d_problem(8)
    eo.pdesolve(M.hjb3d_synth(**prm), g, gu)
    g, gu, prm = M.neoclassical_growth_problem(200)
    eo.pdesolve(M.neoclassical_growth(**prm), g, gu)
    g, gu, prm = M.gomez_problem(100)
    eo.pdesolve(M.gomez(**prm), g, gu)
    assert eo.STATS["singular_retries"] == 0

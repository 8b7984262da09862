"""The plain-C oracle (oracle/oracle1d.c, the CPU baseline of bench.py) agrees with the numpy
oracle and reproduces the published traces' final residuals."""
import numpy as np

from oracle import c_oracle
from oracle import econpdes_oracle as eo
from oracle import models as M


def test_c_oracle_readme_quickstart():
    g, gu = M.readme_linear_problem()
    sol, rn, no, nn = c_oracle.solve_batch("readme_linear", np.array([[0.05, 0.2]]), g["x"][None], gu["v"][None])
    ora = eo.pdesolve(M.readme_linear(), g, gu)
    assert no[0] == 4 and nn[0] == 6
    assert "%.2e" % rn[0] == "1.22e-10"          # README.md:78, with LAPACK-style pivoted tridiagonal LU
    np.testing.assert_allclose(sol[0], ora.solution["v"], rtol=0, atol=1e-8)


def test_c_oracle_campbell_cochrane_and_sweep():
    g, gu, prm = M.campbell_cochrane_problem()
    P = np.array([[prm[k] for k in ("mu", "sigma", "gamma", "rho", "kappa_s", "b")]])
    sol, rn, no, nn = c_oracle.solve_batch("campbell_cochrane", P, g["s"][None], gu["p"][None])
    ora = eo.pdesolve(M.campbell_cochrane(**prm), g, gu)
    assert no[0] == 4 and nn[0] == 8 and "%.2e" % rn[0] == "3.45e-12"   # docs/src/solver.md:73
    np.testing.assert_allclose(sol[0], ora.solution["p"], rtol=1e-10)
    prms = [M.campbell_cochrane_sweep_params(i) for i in range(6)]
    grids = np.stack([M.campbell_cochrane_grid(**p) for p in prms])
    params = np.array([[p[k] for k in ("mu", "sigma", "gamma", "rho", "kappa_s", "b")] for p in prms])
    sol, rn, no, nn = c_oracle.solve_batch("campbell_cochrane", params, grids, np.ones_like(grids), threads=2)
    for b, p in enumerate(prms):
        ora = eo.pdesolve(M.campbell_cochrane(**p), {"s": grids[b]}, {"p": np.ones(grids.shape[1])})
        np.testing.assert_allclose(sol[b], ora.solution["p"], rtol=1e-9)
        assert no[b] == ora.log.n_outer and abs(rn[b] - ora.residual_norm) <= 1e-10


def test_c_oracle_he_krishnamurthy_and_wachter():
    """A third, independent restatement (plain C) of two of the example bodies, solved by the C driver, against the
    numpy oracle: same outer steps and Newton counts, solutions to round-off of the two linear solvers."""
    from econpdes_b200 import models as PM
    g, gu, prm = M.he_krishnamurthy_problem()
    P = np.array([[prm[k] for k in ("m", "lam", "g", "sigma", "rho", "gamma", "l")]])
    sol, rn, no, nn = c_oracle.solve_batch("he_krishnamurthy", P, g["x"][None], gu["pS"][None])
    ora = eo.pdesolve(M.he_krishnamurthy(**prm), g, gu)
    assert no[0] == ora.log.n_outer == 4 and nn[0] == ora.log.n_newton == 21
    np.testing.assert_allclose(sol[0], ora.solution["pS"], rtol=1e-9)
    assert abs(rn[0] - ora.residual_norm) <= 1e-10
    g, gu, prm = M.wachter_problem()
    P = PM.WachterModel(**prm).params()[None]
    sol, rn, no, nn = c_oracle.solve_batch("wachter", P, g["lambda"][None], gu["p"][None])
    ora = eo.pdesolve(M.wachter(**prm), g, gu)
    assert no[0] == ora.log.n_outer == 5 and nn[0] == ora.log.n_newton == 11
    np.testing.assert_allclose(sol[0], ora.solution["p"], rtol=1e-9)
    assert abs(rn[0] - ora.residual_norm) <= 1e-10

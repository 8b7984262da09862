"""Generates tests/golden/ex_*.npz: oracle outputs (solution, saved outputs, iteration log) of the example models
on the examples' own grids.  Same status as make_golden.py: outputs of the oracle restatement (the reference
cannot be executed here), committed so that a change of the oracle or of a model restatement shows up as a diff.

    python tests/golden/make_golden_examples.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import econpdes_oracle as eo  # noqa: E402
from oracle import models as M  # noqa: E402
from tests.golden.make_golden import pack  # noqa: E402


def cases():
    g, gu, prm = M.he_krishnamurthy_problem()
    yield "ex_he_krishnamurthy_100", M.he_krishnamurthy(**prm), g, gu, prm, {}
    g, gu, prm = M.wachter_problem()
    yield "ex_wachter_30", M.wachter(**prm), g, gu, prm, {}
    g, gu, prm = M.haddad_problem()
    yield "ex_haddad_30x30", M.haddad(**prm), g, gu, prm, {}
    g, gu, prm, bc = M.bolton_chen_wang_problem()
    yield "ex_bolton_chen_wang_100", M.bolton_chen_wang(**prm), g, gu, prm, dict(bc=bc)
    g, gu, prm = M.gomez_problem()
    yield "ex_gomez_300", M.gomez(**prm), g, gu, prm, {}
    g, gu, prm, bc = M.ahlm_risky_asset_problem()
    yield "ex_ahlm_risky_asset_5x150", M.ahlm_risky_asset(**prm), g, gu, prm, dict(bc=bc)
    g, gu, prm, bc = M.wang_wang_yang_problem()
    yield "ex_wang_wang_yang_1001", M.wang_wang_yang(**prm), g, gu, prm, dict(bc=bc)
    g, gu, prm, lb, bc = M.leland_problem()
    yield "ex_leland_41", M.leland(**prm), g, gu, prm, dict(bc=bc, lower_bound={"E": lb})
    g, gu, prm = M.ahlm_two_states_problem()
    yield "ex_ahlm_two_states_200", M.ahlm_two_states(**prm), g, gu, prm, {}
    g, gu, prm = M.garleanu_panageas_problem()
    yield "ex_garleanu_panageas_100", M.garleanu_panageas(**prm), g, gu, prm, {}
    # first stage of the 3-D long-run-risk example (garleanu_panageas_long_run_risk.jl:337-344), ~2 minutes of SuperLU
    g, gu, prm, delta, opts = M.garleanu_panageas_lrr_problem()
    yield ("ex_garleanu_panageas_lrr_20x5x5", M.garleanu_panageas_lrr(**prm), g, gu, prm,
           dict(delta=delta, abstol=opts["abstol"], maxiters=opts["maxiters"], max_delta=opts["maxΔ"]))


def main():
    only = sys.argv[1:]
    for name, pde, g, gu, prm, kw in cases():
        if only and name not in only:
            continue
        res = eo.pdesolve(pde, g, gu, **kw)
        d = pack(res, g, gu, prm)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "residual %.3e" % d["residual_norm"], "converged", bool(d["converged"]),
              "outer", d["log"].shape[0], "newton", int(d["n_newton"]))


if __name__ == "__main__":
    main()

"""Generates tests/golden/*.npz from the CPU oracle (oracle/), which is itself pinned on the
reference's published traces and closed-form tests (tests/test_oracle.py).

    python tests/golden/make_golden.py

The reference cannot be executed in the build container (no Julia), so these vectors are
outputs of the oracle restatement, not of the reference binary; see DESIGN.md "Oracle".
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import econpdes_oracle as eo  # noqa: E402
from oracle import models as M  # noqa: E402


def pack(res, grid, guess, prm, **extra):
    out = {f"grid_{k}": v for k, v in grid.items()}
    out.update({f"guess_{k}": v for k, v in guess.items()})
    out.update({f"sol_{k}": v for k, v in res.solution.items()})
    out.update({f"saved_{k}": v for k, v in res.saved.items() if k not in res.solution})
    out["params"] = np.array(list(prm.values()), dtype=np.float64) if prm else np.zeros(0)
    out["residual_norm"] = np.float64(res.residual_norm)
    out["converged"] = np.bool_(res.converged)
    out["exit_reason"] = np.int32(res.log.exit_reason)
    out["log"] = np.array([(r[0], r[1], r[2], r[3], r[4]) for r in res.log.rows], dtype=np.float64).reshape(-1, 5)
    out["n_newton"] = np.int32(res.log.n_newton)
    out.update(extra)
    return out


def main():
    cases = {}
    g, gu = M.readme_linear_problem(50)
    cases["c1_readme_50"] = pack(eo.pdesolve(M.readme_linear(), g, gu), g, gu, dict(rho=0.05, kappa=0.2))
    g, gu, prm = M.campbell_cochrane_problem()
    cases["c4_campbell_cochrane_998"] = pack(eo.pdesolve(M.campbell_cochrane(**prm), g, gu), g, gu, prm)
    for idx in (0, 1, 2, 3, 16383):
        prm = M.campbell_cochrane_sweep_params(idx)
        g, gu, _ = M.campbell_cochrane_problem(prm)
        cases[f"c4_sweep_{idx}"] = pack(eo.pdesolve(M.campbell_cochrane(**prm), g, gu), g, gu, prm)
    g, gu, prm = M.neoclassical_growth_problem(1000)
    cases["growth_1000"] = pack(eo.pdesolve(M.neoclassical_growth(**prm), g, gu), g, gu, prm)
    for n in (30, 64):
        g, gu, prm = M.bansal_yaron_problem(n, n)
        cases[f"c3_bansal_yaron_{n}"] = pack(eo.pdesolve(M.bansal_yaron(**prm), g, gu), g, gu, prm)
    for ny, na in ((10, 100), (16, 256)):
        g, gu, prm = M.ahlm_diffusion_problem(ny, na)
        cases[f"c2_ahlm_{ny}x{na}"] = pack(eo.pdesolve(M.ahlm_diffusion(**prm), g, gu), g, gu, prm)
    g, gu, prm = M.hjb3d_problem(12)
    cases["c5_hjb3d_12"] = pack(eo.pdesolve(M.hjb3d_synth(**prm), g, gu), g, gu, prm)
    for name, d in cases.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "residual %.3e" % d["residual_norm"], "converged", bool(d["converged"]),
              "outer", d["log"].shape[0], "newton", int(d["n_newton"]))


if __name__ == "__main__":
    main()

"""Stand-in for compute-sanitizer (closed on the GPU pool this repo is developed on): every device buffer of the
library gets canary regions (ECONPDES_GUARD=1, engine.cuh:GuardRegistry) and a cross-section of the solve paths
runs twice -- out-of-bounds writes show up as corrupted canaries, data races of the hand-rolled shared-memory
exchanges (PCR, chunked tridiagonal solves, block reductions, zebra line smoothers) as run-to-run differences."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, %(root)r)
import econpdes_b200 as E
from econpdes_b200 import _lib as L
from econpdes_b200 import examples as EX
from econpdes_b200 import models as PM
from oracle import models as OM

lib = L.load()
assert lib.econpdes_guard_violations() == 0      # the facility is on (-1 when off)


def twice(tag, fn):
    a, b = fn(), fn()
    for x, y in zip(a, b):
        assert np.array_equal(np.asarray(x), np.asarray(y), equal_nan=True), tag + ": two runs differ"
    v = lib.econpdes_guard_violations()
    assert v == 0, "%%s: %%d corrupted canary bytes" %% (tag, v)
    print("ok", tag)


def solve(pm, g, gu, **kw):
    r = E.pdesolve(pm, g, gu, verbose=False, warn=False, **kw)
    return [r.solution[k] for k in gu] + [np.float64(r.residual_norm)] + [np.array([row["newton_iters"] for row in r.log])]


# fused batched 1-D kernel: every points-per-thread variant, ragged sizes, compact inputs
for S in (2, 33, 257, 998, 1024):
    x = OM.julia_range(0.0, 1.0, S)
    twice("batched readme S=%%d" %% S, lambda: E.pdesolve_batched(
        PM.ReadmeLinear, np.tile([[0.05, 0.2]], (5, 1)), np.tile(x, (5, 1)), np.zeros((5, S))).solution)
params, grids, guess = EX.campbell_cochrane_sweep(0, 300)
twice("batched campbell-cochrane", lambda: E.solve_batched_one_call(PM.CampbellCochraneModel, params, grids, guess)[:2])
p2, lo, pts, n_range, fill = EX.campbell_cochrane_sweep_compact(0, 300)
twice("batched compact", lambda: E.solve_batched_compact(PM.CampbellCochraneModel, p2, lo, pts, n_range, fill)[:2])
# generic engine: 1-D PCR with 1, 2 and 4 unknowns
g, gu = OM.readme_linear_problem(50)
twice("readme", lambda: solve(PM.ReadmeLinear(), g, gu))
g, gu, prm = OM.ahlm_two_states_problem(200)
twice("two-income F=2", lambda: solve(PM.AchdouHanLasryLionsMollModel_TwoStates(**prm), g, gu))
g, gu, prm = OM.garleanu_panageas_problem(100)
twice("garleanu-panageas F=4", lambda: solve(PM.GarleanuPanageasModel(**prm), g, gu))
# 2-D / 3-D: marching sweeps, Krylov + multigrid, block line relaxation, boundary rows
g, gu, prm = OM.bansal_yaron_problem(30, 30)
twice("bansal-yaron 30x30", lambda: solve(PM.BansalYaronModel(**prm), g, gu))
twice("bansal-yaron 30x30 bicgstab", lambda: solve(PM.BansalYaronModel(**prm), g, gu, linear_solver=2))
g, gu, prm = OM.ahlm_diffusion_problem(10, 100)
twice("consumption-saving 10x100", lambda: solve(PM.AchdouHanLasryLionsMollModel_Diffusion(**prm), g, gu))
g, gu, prm, bc = OM.ahlm_risky_asset_problem()
twice("risky asset 5x150 with bc", lambda: solve(PM.AchdouHanLasryLionsMollModel_RiskyAsset(**prm), g, gu, bc=bc))
g, gu, prm = OM.hjb3d_problem(12)
twice("hjb3d 12^3", lambda: solve(PM.Hjb3dSynth(**prm), g, gu))
g, gu, prm = OM.hjb3d_problem(33)
twice("hjb3d 33^3 (odd sizes, several multigrid levels)", lambda: solve(PM.Hjb3dSynth(**prm), g, gu))
g, gu, prm = OM.hjb2d_regimes_problem(24, 20)
twice("two-regime 2-D pair F=2", lambda: solve(PM.Hjb2dRegimes(**prm), g, gu))
print("guard violations:", lib.econpdes_guard_violations())
"""


@pytest.mark.gpu
def test_no_out_of_bounds_writes_and_no_run_to_run_differences():
    env = dict(os.environ, ECONPDES_GUARD="1")
    out = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True,
                         timeout=900)
    sys.stdout.write(out.stdout[-3000:])
    assert out.returncode == 0, out.stderr[-3000:]
    assert "guard violations: 0" in out.stdout

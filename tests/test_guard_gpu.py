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
# round 2: four state variables (81-point box, line relaxation), the singular-Jacobian fallback (normal-equations CG over
# the stencil planes, ||J'J||_1 kernel), chunked smoother / 32-bit transfers / paired transposes on larger levels, the
# one-cluster coarse cycle, the 8-unknown 3-D block solver with cached reduction multipliers (three outer steps)
g, gu, prm = OM.hjb4d_problem((5, 4, 6, 3))
twice("hjb4d 5x4x6x3", lambda: solve(PM.Hjb4dSynth(**prm), g, gu))
x51 = OM.julia_range(0.0, 1.0, 51)
twice("singular jacobian 1-D (fallback)", lambda: solve(PM.ReadmeLinear(0.0, -0.2), {"x": x51}, {"v": 1 + x51 * x51}, Δ=float("inf"), maxiters=3))
g3 = {k: OM.julia_range(0, 1, n) for k, n in zip("xyz", (5, 7, 5))}
X, Y, Z = np.meshgrid(*g3.values(), indexing="ij")
prm3 = dict(kx=-0.2, ky=-0.3, kz=-0.4, sx=0.0, sy=0.0, sz=0.0, corr=0.5, rho=0.0, theta=0.0)
def fallback3d():
    h = E.Handle(PM.Hjb3dSynth(**prm3), list(g3.values()), (1 + X * Y + Z * Z)[..., None])
    out = h.implicit_timestep((1 + X * Y + Z * Z)[..., None], float("inf"), maxiters=1)
    h.close()
    return [out[0], np.float64(out[1])]
twice("singular jacobian 3-D (fallback)", fallback3d)
g, gu = EX.bansal_yaron_grid(300, 260)
twice("bansal-yaron 300x260 (chunked smoother, 5 levels)", lambda: solve(PM.BansalYaronModel(), g, gu, maxiters=2))
g, gu = EX.hjb3d_grid(136)
twice("hjb3d 136^3 (chunked 3-D smoother, masked planes)", lambda: solve(PM.Hjb3dSynth(), g, gu, maxiters=2))
import os
os.environ["ECONPDES_COARSE_CLUSTER"] = "1"; os.environ["ECONPDES_COARSE_MAX_POINTS"] = "70000"
g, gu = EX.hjb3d_grid(40)
twice("hjb3d 40^3, coarse levels in one cluster kernel", lambda: solve(PM.Hjb3dSynth(), g, gu))
del os.environ["ECONPDES_COARSE_CLUSTER"], os.environ["ECONPDES_COARSE_MAX_POINTS"]
g, gu, prm, delta, opts = OM.garleanu_panageas_lrr_problem()
twice("long-run risk 20x5x5 F=8, three outer steps", lambda: solve(PM.GarleanuPanageasLongRunRiskModel(**prm), g, gu, Δ=delta,
                                                                 maxΔ=opts["maxΔ"], abstol=opts["abstol"], maxiters=3))
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

"""Launches each grid-wide kernel of the generic engine a few times on the C3 (2048^2) shape and
one 3-D shape so that one `ncu --set full` capture covers residual, Jacobian, SpMV and the line
smoother (used to produce profiles/r01_grid_kernels_ncu_full.json)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import econpdes_b200 as E  # noqa: E402
from econpdes_b200 import examples as EX  # noqa: E402
from econpdes_b200 import models as PM  # noqa: E402

n2d = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n3d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
cases = []
if n2d > 0:
    cases.append((PM.BansalYaronModel(), EX.bansal_yaron_grid(n2d, n2d), 0.0))
if n3d > 0:
    cases.append((PM.Hjb3dSynth(), EX.hjb3d_grid(n3d), 20.0))
for model, (g, gu), shift in cases:
    guess = next(iter(gu.values())) + shift
    y = guess * (1 + 0.01 * np.random.default_rng(1).uniform(-1, 1, guess.shape))
    h = E.Handle(model, list(g.values()), y[..., None])
    print(model.name, h.bench_kernels(y[..., None], 1))
    h.close()

import sys, time
sys.path.insert(0, '.')
import numpy as np
import econpdes_b200 as E
from econpdes_b200 import models as PM
from oracle import models as OM
g, gu, prm, delta, opts = OM.garleanu_panageas_lrr_problem()
t = time.time()
res = E.pdesolve(PM.GarleanuPanageasLongRunRiskModel(**prm), g, gu, Δ=delta, maxΔ=opts["maxΔ"], abstol=opts["abstol"], maxiters=opts["maxiters"], verbose=True)
print("converged", res.converged, res.residual_norm, "outer", res.n_outer, "newton", res.n_newton, "lin iters", res.n_linear_iters, "loose", res.n_linear_loose, res.lin_max_relres, "ms", res.device_ms, "wall", time.time() - t, "launches", res.kernel_launches)
for r in res.log: print(r)
np.savez_compressed('gpurun_out/r2m/lrr_device.npz', **{("sol_" + k): v for k, v in res.solution.items()})

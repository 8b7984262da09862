"""Log rows of the first outer steps of C3 at 2048^2 on the device (to set beside the oracle's):
    python tools/run_c3_first_steps.py [steps=3]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import econpdes_b200 as E  # noqa: E402
from econpdes_b200 import models as PM  # noqa: E402
from oracle import models as OM  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g, gu, prm = OM.bansal_yaron_problem(2048, 2048)
res = E.pdesolve(PM.BansalYaronModel(**prm), g, gu, verbose=False, warn=False, maxiters=steps)
for r in res.log:
    print("device step", r["iter"], repr(r["delta"]), repr(r["residual"]), r["status"], r["newton_iters"], r["linear_iters"])
print("residual_norm", repr(res.residual_norm), "device_ms", res.device_ms)

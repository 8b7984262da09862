"""Device time of the fused batched kernel for smaller problems (the C4 sweep on coarser grids) and every
points-per-thread variant:  python tools/bench_batched_sizes.py [problems]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import econpdes_b200 as E  # noqa: E402
from econpdes_b200 import examples as EX  # noqa: E402
from econpdes_b200.models import CampbellCochraneModel  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
out = {}
for sn in (60, 100, 200, 300, 500):
    params, grids, guess = EX.campbell_cochrane_sweep(0, B, sn=sn)
    rec = {"points": int(grids.shape[1])}
    for P in (0, 1, 2, 3, 4):
        E.set_batch_points_per_thread(P)
        try:
            E.pdesolve_batched(CampbellCochraneModel, params, grids, guess)
            ms = min(E.pdesolve_batched(CampbellCochraneModel, params, grids, guess).device_ms for _ in range(7))
            rec["P%d" % P] = round(ms, 3)
        except Exception as e:   # a variant that does not fit the block size
            rec["P%d" % P] = str(e)[:40]
    E.set_batch_points_per_thread(0)
    out["sn=%d" % sn] = rec
    print(json.dumps({("sn=%d" % sn): rec}), flush=True)

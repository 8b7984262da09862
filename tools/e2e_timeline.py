"""Where the time of one end-to-end batched call goes (ECONPDES_BATCH_TIMELINE=1 makes the library print its own
host/device timeline per call):  python tools/e2e_timeline.py [problems ...]"""
import ctypes as C
import os
import sys
import time

import numpy as np

os.environ.setdefault("ECONPDES_BATCH_TIMELINE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import econpdes_b200 as E  # noqa: E402
from econpdes_b200 import _lib as L  # noqa: E402
from econpdes_b200 import examples as EX  # noqa: E402
from econpdes_b200.models import CampbellCochraneModel  # noqa: E402
from econpdes_b200.pdesolve import _options  # noqa: E402

lib = L.load()
mid = L.check(lib.econpdes_model_id(CampbellCochraneModel.name.encode()))
opts = _options(1.4901161193847656e-08, 100, {})
opts.delta = 1.0


def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
    t.numpy()[...] = a
    return t


for B in [int(a) for a in sys.argv[1:]] or [2048, 16384]:
    p2, lo, pts, n_range, fill = EX.campbell_cochrane_sweep_compact(0, B)
    S = n_range + pts.shape[1]
    tp, tlo, tpts = pinned(p2), pinned(lo), pinned(pts)
    tsol = torch.empty((B, S), dtype=torch.float64, pin_memory=True)
    trn = torch.empty(B, dtype=torch.float64, pin_memory=True)
    ex, no, nn = (np.empty(B, dtype=np.int32) for _ in range(3))
    inp = L.BatchInputs()
    inp.params, inp.grids, inp.n_range = L.dptr(tp.numpy()), None, int(n_range)
    inp.range_lo, inp.points, inp.guess, inp.guess_fill = L.dptr(tlo.numpy()), L.dptr(tpts.numpy()), None, float(fill)

    def step():
        L.check(lib.econpdes_solve_batched_inputs(mid, B, S, C.byref(inp), C.byref(opts), L.dptr(tsol.numpy()),
                                                  L.dptr(trn.numpy()), ex.ctypes.data_as(L.c_int32_p),
                                                  no.ctypes.data_as(L.c_int32_p), nn.ctypes.data_as(L.c_int32_p)))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 10
    for _ in range(K):
        step()
    torch.cuda.synchronize()
    print("B", B, "e2e ms per call", (time.perf_counter() - t0) * 1e3 / K, "mode", lib.econpdes_last_batched_output_mode(), flush=True)
    lib.econpdes_release_workspace()

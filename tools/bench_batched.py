"""Times the fused batched kernel (config C4) for every points-per-thread variant, and the
one-call pipelined path from pinned host buffers.

    python tools/bench_batched.py [--problems 16384] [--reps 10]
"""
import argparse
import ctypes as C
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problems", type=int, default=16384)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--variants", default="1,2,3,4,0")
    args = ap.parse_args()
    import torch
    import econpdes_b200 as E
    from econpdes_b200 import _lib as L
    from econpdes_b200 import examples as EX
    from econpdes_b200.models import CampbellCochraneModel
    from econpdes_b200.pdesolve import _options

    lib = L.load()
    B = args.problems
    params, grids, guess = EX.campbell_cochrane_sweep(0, B)
    S = grids.shape[1]
    opts = _options(math.sqrt(np.finfo(float).eps), 100, {})
    mid = L.check(lib.econpdes_model_id(CampbellCochraneModel.name.encode()))
    out = {}
    h = C.c_void_p()
    L.check(lib.econpdes_batch_create(mid, B, S, C.byref(h)))
    L.check(lib.econpdes_batch_upload(h, L.dptr(params), L.dptr(grids), L.dptr(guess)))
    ms = C.c_double()
    for P in [int(v) for v in args.variants.split(",")]:
        E.set_batch_points_per_thread(P)
        for _ in range(3):
            L.check(lib.econpdes_batch_solve(h, C.byref(opts), C.byref(ms)))
        tot = 0.0
        for _ in range(args.reps):
            L.check(lib.econpdes_batch_solve(h, C.byref(opts), C.byref(ms)))
            tot += ms.value
        nn = np.empty(B, dtype=np.int32)
        rn = np.empty(B)
        L.check(lib.econpdes_batch_download(h, None, L.dptr(rn), None, None, nn.ctypes.data_as(L.c_int32_p), None, 0, None))
        out[f"P{P}"] = {"ms": tot / args.reps, "newton": int(nn.sum()), "converged": int((rn <= opts.abstol).sum())}
    lib.econpdes_batch_destroy(h)
    E.set_batch_points_per_thread(0)

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = a
        return t
    tp, tg, tu = pinned(params), pinned(grids), pinned(guess)
    tsol = torch.empty((B, S), dtype=torch.float64, pin_memory=True)
    trn = torch.empty(B, dtype=torch.float64, pin_memory=True)
    i32 = [torch.empty(B, dtype=torch.int32, pin_memory=True) for _ in range(3)]

    def step():
        L.check(lib.econpdes_solve_batched(mid, B, S, L.dptr(tp.numpy()), L.dptr(tg.numpy()), L.dptr(tu.numpy()),
                                           C.byref(opts), L.dptr(tsol.numpy()), L.dptr(trn.numpy()),
                                           *[t.numpy().ctypes.data_as(L.c_int32_p) for t in i32]))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        step()
    out["one_call_pinned_ms"] = (time.perf_counter() - t0) * 1e3 / args.reps
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""ctypes loader of the plain-C oracle (oracle/oracle1d.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liboracle1d.so")
MODEL_IDS = {"readme_linear": 0, "campbell_cochrane": 1, "he_krishnamurthy": 2, "wachter": 3}
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "oracle1d.c")
        if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", HERE])
        _lib = C.CDLL(SO)
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def solve_batch(model: str, params, grids, guess, threads=1):
    """params [B,P], grids [B,n], guess [B,n] -> (solution [B,n], residual_norm, n_outer, n_newton)."""
    L = lib()
    params = np.ascontiguousarray(params, dtype=np.float64)
    grids = np.ascontiguousarray(grids, dtype=np.float64)
    guess = np.ascontiguousarray(guess, dtype=np.float64)
    B, n = grids.shape
    sol = np.empty((B, n))
    rn = np.empty(B)
    no = np.empty(B, dtype=np.int32)
    nn = np.empty(B, dtype=np.int32)
    rc = L.oracle1d_solve_batch(MODEL_IDS[model], params.shape[1], C.c_long(B), n, _p(params), _p(grids), _p(guess),
                                _p(sol), _p(rn), _p(no, C.c_int), _p(nn, C.c_int), int(threads))
    if rc != 0:
        raise ValueError("G! returns NaN with the initial value")
    return sol, rn, no, nn

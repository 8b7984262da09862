"""Loads tests/host_emu/libemu.so (the product's per-point device code compiled for the
host CPU) for the no-GPU unit tests; builds it on demand with g++."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "econpdes.jl_b200", "csrc")
SO = os.path.join(HERE, "host_emu", "libemu.so")
_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(HERE, "host_emu", "emu.cpp")
    deps = [src] + [os.path.join(CSRC, f) for f in ("stencil.cuh", "models.cuh", "pointops.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                               "-I", CSRC, src, "-o", SO])
    _lib = ctypes.CDLL(SO)
    _lib.emu_model_id.argtypes = [ctypes.c_char_p]
    return _lib


def _grid_args(grids):
    S = (ctypes.c_longlong * len(grids))(*[len(g) for g in grids])
    keep = [np.ascontiguousarray(g, dtype=np.float64) for g in grids]
    ptrs = (ctypes.POINTER(ctypes.c_double) * len(grids))(
        *[g.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) for g in keep])
    return S, ptrs, keep


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def residual(model, grids, y, bc=None, saved=False):
    """y: (S..., F) array -> ydot same shape [, saved (S..., NSAVED)]"""
    L = lib()
    mid = L.emu_model_id(model.name.encode())
    assert mid >= 0
    S, ptrs, keep = _grid_args(grids)
    yf = np.asfortranarray(y, dtype=np.float64).reshape(-1, order="F")
    bcf = None if bc is None else np.asfortranarray(bc, dtype=np.float64).reshape(-1, order="F")
    out = np.empty_like(yf)
    l = int(np.prod([len(g) for g in grids]))
    sv = np.empty(l * max(len(model.savednames), 1)) if saved else None
    prm = model.params()
    rc = L.emu_residual(mid, S, ptrs, _p(prm), _p(yf), _p(bcf), _p(out), _p(sv))
    assert rc == 0
    out = out.reshape(y.shape, order="F")
    if saved:
        return out, sv.reshape(tuple(len(g) for g in grids) + (-1,), order="F")
    return out


def jacobian_planes(model, grids, y, ypost, delta, eps_table, box, bc=None, eps_per_column=False):
    L = lib()
    mid = L.emu_model_id(model.name.encode())
    S, ptrs, keep = _grid_args(grids)
    N, F = len(grids), y.shape[-1]
    l = int(np.prod([len(g) for g in grids]))
    nk = 3 ** N if box else 1 + 2 * N
    yf = np.asfortranarray(y, dtype=np.float64).reshape(-1, order="F")
    ypf = None if ypost is None else np.asfortranarray(ypost, dtype=np.float64).reshape(-1, order="F")
    bcf = None if bc is None else np.asfortranarray(bc, dtype=np.float64).reshape(-1, order="F")
    planes = np.zeros(nk * F * F * l)
    r0 = np.empty(F * l)
    prm = model.params()
    et = np.ascontiguousarray(eps_table, dtype=np.float64)
    rc = L.emu_jacobian(mid, int(box), S, ptrs, _p(prm), _p(yf), _p(bcf), _p(ypf), ctypes.c_double(delta), _p(et),
                        int(eps_per_column), _p(planes), _p(r0))
    assert rc == 0
    return planes.reshape(nk, F, F, l), r0


def planes_to_csc(planes, S, F, box, colptr, rowval):
    """Gather stencil-coefficient planes [k][fin][fout][pt] into CSC nzval for the pattern
    of local_stencil_jacobian (1-based colptr/rowval)."""
    S = tuple(S)
    N = len(S)
    l = int(np.prod(S))
    n = l * F
    col_of = np.repeat(np.arange(n), np.diff(colptr))
    row = rowval - 1
    cpt, fin = col_of % l, col_of // l
    rpt, fout = row % l, row // l
    cc = np.stack(np.unravel_index(cpt, S, order="F"), axis=1)
    rc = np.stack(np.unravel_index(rpt, S, order="F"), axis=1)
    off = cc - rc  # column point = row point + off
    if box:
        k = np.zeros(len(row), dtype=np.int64)
        for d in range(N):
            k += (off[:, d] + 1) * 3 ** d
        return planes[k, fin, fout, rpt]
    nzo = np.count_nonzero(off, axis=1)
    k = np.zeros(len(row), dtype=np.int64)
    for d in range(N):
        k = np.where((off[:, d] != 0) & (nzo == 1), 1 + 2 * d + (off[:, d] > 0), k)
    vals = planes[k, fin, fout, rpt]
    return np.where(nzo <= 1, vals, 0.0)


def range_rows(a, b, n):
    """Rows of Julia ranges by the device grid generator (range_point, csrc/stencil.cuh) on the host."""
    L = lib()
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    out = np.empty((a.shape[0], n))
    L.emu_range_rows(_p(a), _p(b), ctypes.c_int(a.shape[0]), ctypes.c_int(n), _p(out))
    return out

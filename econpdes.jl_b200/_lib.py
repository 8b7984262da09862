"""ctypes binding of libeconpdes_cuda.so (include/econpdes_cuda.h).

The library is the product: if it is missing or no CUDA device is usable, every compute
entry point raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libeconpdes_cuda.so")

MAX_DIMS = 4
ABI_VERSION = 4   # ECONPDES_ABI_VERSION of include/econpdes_cuda.h these mirrors were written against

OK, E_INVALID, E_NAN_GUESS, E_NO_DEVICE, E_CUDA, E_UNSUPPORTED, E_NCCL = 0, -1, -2, -3, -4, -5, -6
EXIT_CONVERGED, EXIT_ALREADY_ZERO, EXIT_MAXITERS, EXIT_BELOW_MIN_DELTA, EXIT_INF_STEP_FAILED = range(5)
STEP_ACCEPTED, STEP_REJECTED_GROWTH, STEP_REJECTED_INNER, STEP_REJECTED_NAN, STEP_REJECTED_SINGULAR = range(5)
EPS_FINITEDIFF_COLOR_NORM, EPS_PER_COLUMN = 0, 1

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_uint8_p = C.POINTER(C.c_uint8)


class Problem(C.Structure):
    _fields_ = [("model_id", C.c_int32), ("ndims", C.c_int32), ("nfuncs", C.c_int32), ("nparams", C.c_int32),
                ("size", C.c_int64 * MAX_DIMS), ("grid", c_double_p * MAX_DIMS), ("params", c_double_p),
                ("guess", c_double_p), ("bc", c_double_p), ("is_algebraic", c_uint8_p),
                ("lower_bound", c_double_p), ("upper_bound", c_double_p), ("delta_vec", c_double_p)]


class Options(C.Structure):
    _fields_ = [("abstol", C.c_double), ("inner_abstol", C.c_double), ("delta", C.c_double), ("scale", C.c_double),
                ("min_delta", C.c_double), ("max_delta", C.c_double), ("max_residual_growth", C.c_double),
                ("maxiters", C.c_int32), ("inner_maxiters", C.c_int32), ("eps_policy", C.c_int32),
                ("linear_solver", C.c_int32), ("lin_rtol", C.c_double), ("lin_maxiters", C.c_int32),
                ("lin_restart", C.c_int32), ("check_monotonicity", C.c_int32), ("monotonicity_tol", C.c_double)]


class LogRow(C.Structure):
    _fields_ = [("iter", C.c_int32), ("status", C.c_int32), ("newton_iters", C.c_int32), ("linear_iters", C.c_int32),
                ("delta", C.c_double), ("residual", C.c_double)]


class Result(C.Structure):
    _fields_ = [("solution", c_double_p), ("saved", c_double_p), ("log", C.POINTER(LogRow)),
                ("log_capacity", C.c_int32), ("log_len", C.c_int32), ("residual_norm", C.c_double),
                ("converged", C.c_int32), ("exit_reason", C.c_int32), ("n_outer", C.c_int32), ("n_newton", C.c_int32),
                ("n_linear_iters", C.c_int64), ("kernel_launches", C.c_int64), ("device_ms", C.c_double),
                ("sign_flags", C.c_int32), ("nonmonotone_entries", C.c_int32), ("nonmonotone_max", C.c_double),
                ("n_linear_solves", C.c_int32), ("n_linear_loose", C.c_int32), ("lin_max_relres", C.c_double)]


class BatchInputs(C.Structure):
    _fields_ = [("params", c_double_p), ("grids", c_double_p), ("n_range", C.c_int32), ("range_lo", c_double_p),
                ("points", c_double_p), ("guess", c_double_p), ("guess_fill", C.c_double)]


SIGN_WHOLE_SYSTEM = 1 << 30


# every symbol include/econpdes_cuda.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "econpdes_abi_version": (C.c_int, []),
    "econpdes_last_error": (C.c_char_p, []),
    "econpdes_model_count": (C.c_int, []),
    "econpdes_model_id": (C.c_int, [C.c_char_p]),
    "econpdes_model_name": (C.c_char_p, [C.c_int]),
    "econpdes_model_info": (C.c_int, [C.c_int, c_int32_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p]),
    "econpdes_stencil_colors": (C.c_int, [C.c_int, c_int64_p, C.c_int, c_int64_p]),
    "econpdes_stencil_nnz": (C.c_int64, [C.c_int, c_int64_p, C.c_int]),
    "econpdes_stencil_pattern": (C.c_int, [C.c_int, c_int64_p, C.c_int, c_int64_p, c_int64_p]),
    "econpdes_default_options": (None, [C.POINTER(Options)]),
    "econpdes_solve_stationary": (C.c_int, [C.POINTER(Problem), C.POINTER(Options), C.POINTER(Result)]),
    "econpdes_solve_backward": (C.c_int, [C.POINTER(Problem), C.POINTER(Options), c_double_p, C.c_int32, c_double_p,
                                          c_double_p, c_double_p, c_int32_p, c_double_p, c_int64_p]),
    "econpdes_create": (C.c_int, [C.POINTER(Problem), C.POINTER(C.c_void_p)]),
    "econpdes_destroy": (None, [C.c_void_p]),
    "econpdes_residual": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "econpdes_saved": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "econpdes_jacobian_csc": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_double, C.c_int32, c_double_p,
                                        c_double_p]),
    "econpdes_linear_solve": (C.c_int, [C.c_void_p, C.POINTER(Options), c_double_p, c_double_p, c_int32_p,
                                        c_double_p]),
    "econpdes_implicit_timestep": (C.c_int, [C.c_void_p, C.POINTER(Options), c_double_p, C.c_double, C.c_int32,
                                             C.c_double, c_double_p, c_double_p, c_int32_p]),
    "econpdes_bench_kernels": (C.c_int, [C.c_void_p, c_double_p, C.c_int, c_double_p]),
    "econpdes_solve": (C.c_int, [C.c_void_p, C.POINTER(Options), C.POINTER(Result)]),
    "econpdes_batch_create": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]),
    "econpdes_batch_destroy": (None, [C.c_void_p]),
    "econpdes_batch_upload": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p]),
    "econpdes_batch_set_log_capacity": (C.c_int, [C.c_void_p, C.c_int32]),
    "econpdes_batch_solve": (C.c_int, [C.c_void_p, C.POINTER(Options), c_double_p]),
    "econpdes_batch_download": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_int32_p, c_int32_p, c_int32_p,
                                          C.POINTER(LogRow), C.c_int32, c_int32_p]),
    "econpdes_batch_kernel_launches": (C.c_int64, [C.c_void_p]),
    "econpdes_solve_batched": (C.c_int, [C.c_int, C.c_int64, C.c_int64, c_double_p, c_double_p, c_double_p,
                                         C.POINTER(Options), c_double_p, c_double_p, c_int32_p, c_int32_p,
                                         c_int32_p]),
    "econpdes_batch_upload_inputs": (C.c_int, [C.c_void_p, C.POINTER(BatchInputs)]),
    "econpdes_solve_batched_inputs": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.POINTER(BatchInputs),
                                                C.POINTER(Options), c_double_p, c_double_p, c_int32_p, c_int32_p,
                                                c_int32_p]),
    "econpdes_release_workspace": (None, []),
    "econpdes_last_batched_output_mode": (C.c_int, []),
    "econpdes_set_batch_points_per_thread": (C.c_int, [C.c_int]),
    "econpdes_selftest_reciprocal": (C.c_int, [C.c_uint64, C.c_int64, C.POINTER(C.c_int64)]),
    "econpdes_device_count": (C.c_int, []),
    "econpdes_set_device": (C.c_int, [C.c_int]),
    "econpdes_total_kernel_launches": (C.c_int64, []),
    "econpdes_guard_violations": (C.c_int64, []),
    "econpdes_dist_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "econpdes_dist_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_void_p)]),
    "econpdes_dist_destroy": (None, [C.c_void_p]),
    "econpdes_dist_collectives": (C.c_int64, [C.c_void_p]),
    "econpdes_dist_p2p_exchanges": (C.c_int64, [C.c_void_p]),
    "econpdes_slab_range": (C.c_int, [C.c_int64, C.c_int, C.c_int, c_int64_p, c_int64_p]),
    "econpdes_solve_stationary_slab": (C.c_int, [C.c_void_p, C.POINTER(Problem), C.c_int64, C.c_int64,
                                                 C.POINTER(Options), C.POINTER(Result)]),
}

_lib = None


class EconPDEsCudaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libeconpdes_cuda error {code}: {msg}")
        self.code = code


def load():
    """Load the shared library (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python econpdes.jl_b200/build.py` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.econpdes_abi_version() != ABI_VERSION:
        raise ImportError("libeconpdes_cuda ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int):
    if rc < 0:
        msg = load().econpdes_last_error()
        raise EconPDEsCudaError(rc, msg.decode() if msg else "")
    return rc


def dptr(a):
    return None if a is None else a.ctypes.data_as(c_double_p)

"""econpdes_b200: B200-native finite-scheme hot path of EconPDEs.jl behind its own API.

    from econpdes_b200 import pdesolve, models
    result = pdesolve(models.BansalYaronModel(), grid, guess)      # -> EconPDEResult

The compute lives in libeconpdes_cuda.so (hand-written FP64 sm_100a kernels, C ABI in
include/econpdes_cuda.h); this package is the thin host mirror of the reference's
`pdesolve(pde, grid, guess; ...)` front end.  No CPU fallback.
"""
from . import models  # noqa: F401
from ._lib import EconPDEsCudaError, LIB_PATH  # noqa: F401
from .pdesolve import (BatchResult, EconPDEResult, Handle, pdesolve, pdesolve_batched,  # noqa: F401
                       set_batch_points_per_thread, last_batched_output_mode, solve_batched_one_call, solve_batched_compact, stencil_colors, stencil_pattern)
from .slab import SlabGroup, pdesolve_slab, slab_range  # noqa: F401

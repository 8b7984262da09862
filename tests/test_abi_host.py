"""No-GPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports
every symbol include/econpdes_cuda.h declares, its host-only entry points (pattern, coloring,
model registry, default options) are bit-exact with the oracle, and compute entry points
fail loudly without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import econpdes_b200 as E
from econpdes_b200 import _lib as L
from econpdes_b200 import models as PM
from oracle import econpdes_oracle as eo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ep_build", os.path.join(ROOT, "econpdes.jl_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    return L.load()


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "econpdes_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(econpdes_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.econpdes_abi_version() == L.ABI_VERSION == 4


def test_model_registry_matches_python_mirror(lib):
    assert lib.econpdes_model_count() == len(PM.ALL_MODELS)
    for cls in PM.ALL_MODELS:
        mid = lib.econpdes_model_id(cls.name.encode())
        assert mid >= 0, cls.name
        nd, nf, npar, ns, box = (C.c_int32() for _ in range(5))
        assert lib.econpdes_model_info(mid, nd, nf, npar, ns, box) == 0
        m = cls()
        assert (nd.value, nf.value, npar.value, ns.value) == (m.ndims, m.nfuncs, m.params().shape[0], len(m.savednames))
        assert lib.econpdes_model_name(mid).decode() == cls.name
    assert lib.econpdes_model_id(b"no_such_model") < 0
    assert b"no_such_model" in lib.econpdes_last_error()


def test_models_md_is_current():
    """MODELS.md (parameter-vector layouts, saved-plane order: both part of the C ABI) is generated from the mirror."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_gen_models_md", os.path.join(ROOT, "tools", "gen_models_md.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert open(os.path.join(ROOT, "MODELS.md"), encoding="utf-8").read() == mod.render(), \
        "run `python tools/gen_models_md.py`"


def test_default_options_are_the_reference_keywords(lib):
    # src/finiteschemesolve.jl:343, docs/src/solver.md:286-298
    o = L.Options()
    lib.econpdes_default_options(C.byref(o))
    assert o.abstol == o.inner_abstol == eo.SQRT_EPS
    assert (o.delta, o.scale, o.min_delta, o.max_residual_growth) == (1.0, 10.0, 1e-9, 10.0)
    assert o.max_delta == float("inf") and o.maxiters == 100 and o.inner_maxiters == 10
    assert o.eps_policy == L.EPS_FINITEDIFF_COLOR_NORM


@pytest.mark.parametrize("S,F", [((5,), 1), ((2,), 2), ((4, 5), 2), ((2, 5), 1), ((3, 4, 5), 2), ((2, 2, 2, 2), 1),
                                 ((64, 128), 1)])
def test_pattern_and_colors_bit_exact(lib, S, F):
    # test/runtests.jl:294-312, 491-502, 509-524
    colors = E.stencil_colors(S, F)
    np.testing.assert_array_equal(colors, eo.stencil_colors(S, F))
    colptr, rowval = E.stencil_pattern(S, F)
    cp, rv = eo.local_stencil_jacobian(S, F)
    np.testing.assert_array_equal(colptr, cp)
    np.testing.assert_array_equal(rowval, rv)
    assert rowval.shape[0] == F * F * int(np.prod([3 * s - 2 for s in S]))
    assert colors.max() == F * int(np.prod([min(3, s) for s in S]))


def test_pattern_520_and_16_colors(lib):
    colptr, rowval = E.stencil_pattern((4, 5), 2)
    assert colptr.shape[0] - 1 == 40 and rowval.shape[0] == 520
    assert E.stencil_colors((2, 2, 2, 2), 1).max() == 16
    with pytest.raises(E.EconPDEsCudaError):
        E.stencil_colors((1,), 1)          # at least 2 points per dimension (stategrid.jl:16)


def test_compute_fails_loudly_without_device(lib):
    if lib.econpdes_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(E.EconPDEsCudaError) as ei:
        E.pdesolve(PM.ReadmeLinear(), {"x": np.linspace(0, 1, 50)}, {"v": np.zeros(50)}, verbose=False)
    assert ei.value.code == L.E_NO_DEVICE and "no CPU fallback" in str(ei.value)
    with pytest.raises(E.EconPDEsCudaError) as ei:
        E.pdesolve_batched(PM.ReadmeLinear, np.array([[0.05, 0.2]]), np.linspace(0, 1, 50)[None], np.zeros((1, 50)))
    assert ei.value.code == L.E_NO_DEVICE


def test_host_side_normalisation_errors():
    # src/pdesolve.jl:96-133 argument checks that need no device
    g = {"k": np.linspace(1, 2, 10)}
    with pytest.raises(TypeError):
        E.pdesolve(lambda s, u: None, g, {"v": np.zeros(10)}, verbose=False)
    with pytest.raises(ValueError, match="size"):
        E.pdesolve(PM.GrowthTest(), g, {"v": np.zeros(3)}, verbose=False)
    with pytest.raises(ValueError, match="at least 2 points"):
        E.pdesolve(PM.GrowthTest(), {"k": np.array([1.0])}, {"v": np.zeros(1)}, verbose=False)
    with pytest.raises(ValueError, match="unknown `bc` entry"):
        E.pdesolve(PM.GrowthTest(), g, {"v": np.zeros(10)}, bc={"vx": (0, 0)}, verbose=False)
    with pytest.raises(ValueError, match="removed"):
        E.pdesolve(PM.GrowthTest(), g, {"v": np.zeros(10)}, maxdist=1e-6, verbose=False)
    with pytest.raises(ValueError, match="only supported by `finiteschemesolve`"):
        E.pdesolve(PM.GrowthTest(), g, {"v": np.zeros(10)}, colorvec=np.ones(10), verbose=False)
    with pytest.raises(ValueError, match="ambiguous names"):
        from econpdes_b200.pdesolve import _check_generated_names
        _check_generated_names(["k", "a"], ["v", "vk"])


def test_result_object_surface():
    """src/EconPDEs.jl:36-95 and test/runtests.jl:66-85: computed `converged`, the deprecated 1.x aliases,
    tuple destructuring and the compact `show` (names and sizes, not the arrays)."""
    from collections import OrderedDict
    from econpdes_b200.pdesolve import EconPDEResult
    sol = OrderedDict(v=np.arange(200.0))
    saved = OrderedDict(v=sol["v"], c=np.ones(200))
    res = EconPDEResult(sol, 1e-9, saved, eo.SQRT_EPS)
    assert res.converged and not EconPDEResult(sol, float("nan"), saved, eo.SQRT_EPS).converged
    assert not EconPDEResult(sol, np.array([0.0, 1e-3]), saved, eo.SQRT_EPS).converged      # one residual per time
    with pytest.warns(DeprecationWarning, match=r"result\.zero"):
        assert res.zero is res.solution
    with pytest.warns(DeprecationWarning, match=r"result\.optional"):
        assert res.optional is res.saved
    solution, residual_norm, sv = res
    assert solution is res.solution and residual_norm == res.residual_norm and sv is res.saved
    text = repr(res)
    assert "EconPDEResult" in text and "v (200)" in text and len(text) < 500
    assert "converged:     true" in text
    assert "converged:     false" in repr(EconPDEResult(sol, 1.0, saved, eo.SQRT_EPS))
    # test/runtests.jl:179-183
    from econpdes_b200.pdesolve import _problem_description
    assert _problem_description(("v",), ("k",), (3,)) == "1 unknown (v) on grid k = 3"
    assert _problem_description(("pA", "pB"), ("x", "y", "z"), (2, 3, 4)) == "2 unknowns (pA, pB) on grid x×y×z = 2×3×4"


def test_verbose_table_and_warning_texts():
    """The printer the host runs over the returned iteration log: the formats of src/finiteschemesolve.jl:403-404
    (header), :421 (accepted row), :433-435 (growth rejection with its factor, "∞" when the residual blew up),
    :441-445 (inner failure, NaN, singular Jacobian), :383 / :455 (closing line) and the @warn texts of :387, :459, :461."""
    import io
    from collections import OrderedDict
    from econpdes_b200.pdesolve import EconPDEResult, _print_log, _warn_text
    sol = OrderedDict(v=np.zeros(50))

    def row(i, st, d, r, nit=3):
        return dict(iter=i, status=st, newton_iters=nit, linear_iters=0, delta=d, residual=r)
    rows = (row(1, L.STEP_ACCEPTED, 1.0, 7.51e-2), row(2, L.STEP_REJECTED_GROWTH, 10.9, 12.34),
            row(3, L.STEP_REJECTED_GROWTH, 1.09, float("inf")), row(4, L.STEP_REJECTED_GROWTH, 0.109, float("nan")),
            row(5, L.STEP_REJECTED_INNER, 1.09e-2, float("nan")), row(6, L.STEP_REJECTED_NAN, 1.09e-3, float("nan")),
            row(7, L.STEP_REJECTED_SINGULAR, 1.09e-4, float("nan")), row(8, L.STEP_ACCEPTED, 1.09e-5, 1.22e-10))
    res = EconPDEResult(sol, 1.22e-10, OrderedDict(), eo.SQRT_EPS, L.EXIT_CONVERGED, rows, 8, 20, 0, 100, 63.4)
    buf = io.StringIO()
    _print_log(res, "1 unknown (v) on grid x = 50", buf)
    assert buf.getvalue().splitlines() == [
        "Solving for 1 unknown (v) on grid x = 50",
        "Iter TimeStep Residual",
        "---- -------- --------",
        "   1 1.00e+00 7.51e-02",
        "   2 1.09e+01        ✗ (residual grew 12.3× > max_residual_growth)",
        "   3 1.09e+00        ✗ (residual grew ∞ > max_residual_growth)",
        "   4 1.09e-01        ✗ (residual grew > max_residual_growth)",      # engines that do not report the factor
        "   5 1.09e-02        ✗",
        "   6 1.09e-03        ✗ (model returned NaN)",
        "   7 1.09e-04        ✗ (singular Jacobian)",
        "   8 1.09e-05 1.22e-10",
        "Converged after 8 time steps (63ms)"]
    one_shot = EconPDEResult(sol, 0.0, OrderedDict(), eo.SQRT_EPS, L.EXIT_CONVERGED, (), 0, 2, 0, 10, 1234.0)
    buf = io.StringIO()
    _print_log(one_shot, "1 unknown (v) on grid x = 50", buf)
    assert buf.getvalue().splitlines() == ["Solving for 1 unknown (v) on grid x = 50", "Converged (1.2s)"]
    assert _warn_text(res, 1e-9) is None
    bad = EconPDEResult(sol, 2.5e-3, OrderedDict(), eo.SQRT_EPS, L.EXIT_MAXITERS, rows, 8)
    assert _warn_text(bad, 1e-9) == ("did not converge after 8 time steps: residual 2.50e-03 > tolerance 1.49e-08. If the "
                                     "residual was still falling, increase `maxiters`; otherwise try a better initial guess "
                                     "or `check_monotonicity = true`.")
    bad.exit_reason = L.EXIT_BELOW_MIN_DELTA
    assert "fell below `minΔ` = 1e-09 (residual 2.50e-03 > tolerance 1.49e-08). This usually means" in _warn_text(bad, 1e-9)
    bad.exit_reason = L.EXIT_INF_STEP_FAILED
    assert _warn_text(bad, 1e-9).endswith("`Δ = Inf` disables pseudo-transient continuation — try a finite initial time "
                                          "step `Δ`, a better initial guess, or more `maxiters`.")


def test_graft_entry_build_check_passes():
    """The driver's "does it build" entry point (incremental: a no-op when the library is current)."""
    import __graft_entry__ as ge
    ge.build()

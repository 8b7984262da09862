"""Device-time of the grid-wide kernels (residual sweep, fused colored-FD Jacobian, stencil SpMV,
multigrid V-cycle) on the C3 (2048^2 Bansal-Yaron) and C5 (256^3 synthetic HJB) shapes, as
achieved algorithmic GB/s against the measured HBM peak.  L2 is flushed between launches."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(reps=5, n2d=2048, n3d=256):
    import econpdes_b200 as E
    from econpdes_b200 import examples as EX
    from econpdes_b200 import models as PM
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    out = {}
    for name, model, (g, gu), K in (
            (f"C3 bansal_yaron {n2d}x{n2d}", PM.BansalYaronModel(), EX.bansal_yaron_grid(n2d, n2d), 5),
            (f"C5 hjb3d_synth {n3d}^3", PM.Hjb3dSynth(), EX.hjb3d_grid(n3d), 27)):
        guess = next(iter(gu.values()))
        rng = np.random.default_rng(1)
        y = (guess if name.startswith("C3") else guess + 20.0) * (1 + 0.01 * rng.uniform(-1, 1, guess.shape))
        h = E.Handle(model, list(g.values()), y[..., None])
        ms = h.bench_kernels(y[..., None], reps)
        h.close()
        l = guess.size
        alg = {"residual": 16 * l, "implicit_residual_norms": 24 * l, "jacobian": (16 + 8 * K) * l,
               "spmv": (8 * K + 16) * l}
        rec = {"points": l, "stencil_planes": K}
        for k, b in alg.items():
            gbs = b / (ms[k] * 1e-3) / 1e9
            rec[k] = {"ms": ms[k], "algorithmic_bytes": b, "achieved_gbs": gbs, "frac_of_measured_hbm_peak": gbs / peak}
        rec["vcycle_ms"] = ms["vcycle"]
        rec["mg_setup_ms"] = ms["mg_setup"]
        out[name] = rec
    return out


if __name__ == "__main__":
    n2d = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    n3d = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    print(json.dumps(measure(n2d=n2d, n3d=n3d), indent=1))

"""Device-time of the grid-wide kernels (residual sweep, fused colored-FD Jacobian, stencil SpMV,
multigrid V-cycle) on the C3 (2048^2 Bansal-Yaron) and C5 (256^3 synthetic HJB) shapes, as
achieved algorithmic GB/s against the measured HBM peak.  L2 is flushed between launches."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _ncu(summary, kernel_substr):
    """Counters of one kernel from a committed `ncu --set full` summary under profiles/ (tools/ncu_summary.py):
    what bounds the kernel when that is not HBM.  Static evidence (the capture's file is named), not a live figure."""
    try:
        with open(os.path.join(ROOT, "profiles", summary)) as f:
            k = [k for k in json.load(f)["kernels"] if kernel_substr in k["kernel"]][0]
    except Exception:
        return None
    num = lambda key: float(k[key].split()[0]) if key in k else None   # noqa: E731
    return {"source": "profiles/" + summary, "kernel": k["kernel"].split("(")[0],
            "fp64_pipe_active_pct": num("fp64_pipe_active_pct"), "issue_active_pct": num("issue_active_pct"),
            "achieved_occupancy_pct": num("achieved_occupancy_pct"), "registers_per_thread": num("registers_per_thread"),
            "warp_instructions": num("warp_instructions"), "dram_throughput_pct": num("dram_throughput_pct")}


# planes of the 27-point box the synthetic 3-D model's residual reads (Hjb3dSynth::slot_used): the others are exact
# zeros of the Jacobian, in the reference's too, and are neither written by the Jacobian kernel nor read by the SpMV
HJB3D_PLANES_USED = 9


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
        is3d = name.startswith("C5")
        prof = "r02_grid_kernels_256cube_ncu_full.json" if is3d else "r02_grid_kernels_2048sq_ncu_full.json"
        tag = "Hjb3dSynth" if is3d else "BansalYaron"
        ncu = {"residual": _ncu(prof, "k_residual_march<%s, 0" % tag),
               "implicit_residual_norms": _ncu(prof, "k_residual_march<%s, 1" % tag),
               "jacobian": _ncu(prof, "k_jacobian_march<%s" % tag), "spmv": _ncu(prof, "k_spmv<")}
        for k, b in alg.items():
            gbs = b / (ms[k] * 1e-3) / 1e9
            rec[k] = {"ms": ms[k], "algorithmic_bytes": b, "achieved_gbs": gbs, "frac_of_measured_hbm_peak": gbs / peak}
            if ncu.get(k):
                # FP64 / issue roofline next to the HBM one: share of the FP64 pipe's issue slots the kernel used
                # (B200: 64 FP64 lanes per SM = one warp instruction every 2 cycles per scheduler), from the capture
                rec[k]["ncu"] = ncu[k]
                if ncu[k]["warp_instructions"]:
                    rec[k]["warp_instructions_per_point"] = ncu[k]["warp_instructions"] * 32 / l
        if is3d:
            # bytes the kernels actually move: the model reads 9 of the 27 planes (see HJB3D_PLANES_USED)
            for k, b in (("spmv", (8 * HJB3D_PLANES_USED + 16) * l), ("jacobian", (16 + 8 * HJB3D_PLANES_USED) * l)):
                rec[k]["planes_nonzero"] = HJB3D_PLANES_USED
                rec[k]["bytes_on_nonzero_planes"] = b
                rec[k]["frac_of_measured_hbm_peak_on_nonzero_planes"] = b / (ms[k] * 1e-3) / 1e9 / peak
        rec["vcycle_ms"] = ms["vcycle"]
        rec["mg_setup_ms"] = ms["mg_setup"]
        out[name] = rec
    return out


if __name__ == "__main__":
    n2d = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    n3d = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    print(json.dumps(measure(n2d=n2d, n3d=n3d), indent=1))

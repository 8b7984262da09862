"""Runs the BASELINE.json configs at full size on one GPU and prints one JSON line each.

    python tools/run_configs.py [c1 c2 c3 c5 ...] [--n2048 N] [--n3d N]
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import econpdes_b200 as E  # noqa: E402
from econpdes_b200 import examples as EX  # noqa: E402
from econpdes_b200 import models as PM  # noqa: E402


def summary(name, res, wall, extra=None):
    out = {"config": name, "converged": res.converged, "residual_norm": res.residual_norm,
           "exit_reason": res.exit_reason, "n_outer": res.n_outer, "n_newton": res.n_newton,
           "n_linear_iters": res.n_linear_iters, "device_ms": res.device_ms, "wall_ms": wall * 1e3,
           "kernel_launches": res.kernel_launches,
           "log": [(r["iter"], "%.2e" % r["delta"], "%.2e" % r["residual"], r["status"], r["newton_iters"],
                    r["linear_iters"]) for r in res.log]}
    out.update(extra or {})
    print(json.dumps(out), flush=True)


def main():
    args = sys.argv[1:]
    which = [a for a in args if not a.startswith("--")] or ["c1", "c2", "c3", "c5"]
    opt = {a.split("=")[0]: a.split("=")[1] for a in args if a.startswith("--") and "=" in a}
    kw = {}
    if "--maxiters" in opt:
        kw["maxiters"] = int(opt["--maxiters"])
    if "--eps_policy" in opt:
        kw["eps_policy"] = int(opt["--eps_policy"])
    if "c1" in which:
        g = {"x": np.linspace(0, 1, 50)}
        t = time.time(); res = E.pdesolve(PM.ReadmeLinear(), g, {"v": np.zeros(50)}, verbose=False)
        summary("C1 readme 50", res, time.time() - t)
    if "c2" in which:
        g, gu = EX.ahlm_diffusion_grid(64, 4096)
        t = time.time(); res = E.pdesolve(PM.AchdouHanLasryLionsMollModel_Diffusion(), g, gu, verbose=False, **kw)
        summary("C2 consumption-saving 64x4096", res, time.time() - t,
                {"v_range": [float(res.solution["v"].min()), float(res.solution["v"].max())]})
    if "c3" in which:
        n = int(opt.get("--n2d", 2048))
        g, gu = EX.bansal_yaron_grid(n, n)
        t = time.time(); res = E.pdesolve(PM.BansalYaronModel(), g, gu, verbose=False, warn=False, **kw)
        summary(f"C3 bansal-yaron {n}x{n}", res, time.time() - t,
                {"p_range": [float(res.solution["p"].min()), float(res.solution["p"].max())]})
    if "c5" in which:
        n = int(opt.get("--n3d", 256))
        g, gu = EX.hjb3d_grid(n)
        t = time.time(); res = E.pdesolve(PM.Hjb3dSynth(), g, gu, verbose=False, warn=False, **kw)
        summary(f"C5 hjb3d {n}^3", res, time.time() - t,
                {"v_range": [float(res.solution["v"].min()), float(res.solution["v"].max())]})


if __name__ == "__main__":
    main()

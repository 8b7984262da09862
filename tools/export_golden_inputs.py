"""Companion of tools/make_reference_goldens.jl.

    python tools/export_golden_inputs.py            writes tests/golden/reference_io/<case>/{params,grid_i,guess_i}.csv
                                                    (17 significant digits) for the oracle fixtures in tests/golden
    python tools/export_golden_inputs.py --compare  after the Julia script ran: compares the reference's solutions,
                                                    residual norms, converged flags and accepted-step tables with the
                                                    committed oracle fixtures (north-star tolerances: 1e-8 relative on
                                                    values, 1e-10 on the residual norm, same flag)
The reference cannot run in the build container, so `--compare` has never been executed there: it is the pinning
procedure for the configs the reference publishes no numbers for (C2, C3, C5).
"""
import glob
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
IO = os.path.join(GOLD, "reference_io")
PREFIXES = ("c1_readme", "c2_ahlm", "c3_bansal_yaron", "c4_", "c5_hjb3d")


def cases():
    for path in sorted(glob.glob(os.path.join(GOLD, "*.npz"))):
        name = os.path.basename(path)[:-4]
        if name.startswith(PREFIXES):
            yield name, np.load(path)


def export():
    for name, d in cases():
        out = os.path.join(IO, name)
        os.makedirs(out, exist_ok=True)
        np.savetxt(os.path.join(out, "params.csv"), np.atleast_1d(d["params"])[None, :], delimiter=",", fmt="%.17g")
        grids = [k for k in d.files if k.startswith("grid_")]
        guesses = [k for k in d.files if k.startswith("guess_")]
        for i, k in enumerate(grids, 1):
            np.savetxt(os.path.join(out, f"grid_{i}.csv"), d[k], fmt="%.17g")
        for i, k in enumerate(guesses, 1):
            np.savetxt(os.path.join(out, f"guess_{i}.csv"), d[k].reshape(-1, order="F"), fmt="%.17g")
        print("exported", name, [d[k].shape for k in grids])


def compare():
    bad = 0
    for name, d in cases():
        out = os.path.join(IO, name)
        summ = os.path.join(out, "ref_summary.csv")
        if not os.path.exists(summ):
            print(name, "no reference output")
            continue
        rn, conv, wall = np.loadtxt(summ, delimiter=",")
        sols = [k for k in d.files if k.startswith("sol_")]
        worst = 0.0
        for i, k in enumerate(sols, 1):
            ref = np.loadtxt(os.path.join(out, f"ref_solution_{i}.csv")).reshape(d[k].shape, order="F")
            worst = max(worst, float(np.max(np.abs(ref - d[k]) / np.maximum(np.abs(ref), 1e-300))))
        # accepted rows of the printed table: "   1 1.00e+00 3.51e-03"
        rows = [ln.split() for ln in open(os.path.join(out, "ref_log.txt")) if re.match(r"^\s*\d+\s+\d", ln)]
        ref_status = [0 if len(r) == 3 and "✗" not in r[2] else 1 for r in rows]
        ora_status = [int(s != 0) for s in d["log"][:, 3]]
        ok = (worst <= 1e-8 and abs(rn - float(d["residual_norm"])) <= 1e-10 and bool(conv) == bool(d["converged"])
              and ref_status == ora_status)
        bad += not ok
        print(f"{name}: {'OK' if ok else 'MISMATCH'}  max rel diff {worst:.2e}  residual {rn:.3e} vs {float(d['residual_norm']):.3e}  "
              f"converged {bool(conv)} vs {bool(d['converged'])}  steps {len(ref_status)} vs {len(ora_status)}  reference wall {wall:.2f} s")
    return bad


if __name__ == "__main__":
    sys.exit(compare() if "--compare" in sys.argv else export())

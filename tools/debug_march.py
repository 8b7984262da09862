"""Cross-check of the marching sweeps (sweep.cuh) against the one-thread-per-point kernels: run once with
ECONPDES_MARCH=0 and once with ECONPDES_MARCH=1 (the switch is read once per process), each writing an .npz;
`compare` then reports the first differing quantity bit for bit.

    ECONPDES_MARCH=0 python tools/debug_march.py dump /tmp/a.npz
    ECONPDES_MARCH=1 python tools/debug_march.py dump /tmp/b.npz
    python tools/debug_march.py compare /tmp/a.npz /tmp/b.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cases():
    from econpdes_b200 import models as PM
    from oracle import models as OM
    out = []
    g, gu, prm, bc = OM.ahlm_risky_asset_problem()
    out.append(("risky", PM.AchdouHanLasryLionsMollModel_RiskyAsset(**prm), g, gu, bc))
    g, gu, prm = OM.bansal_yaron_problem(24, 20)
    out.append(("by", PM.BansalYaronModel(**prm), g, gu, None))
    g, gu, prm = OM.ahlm_diffusion_problem(10, 100)
    out.append(("ahlm", PM.AchdouHanLasryLionsMollModel_Diffusion(**prm), g, gu, None))
    g, gu, prm = OM.hjb3d_problem(12)
    out.append(("hjb3d", PM.Hjb3dSynth(**prm), g, gu, None))
    g, gu, prm = OM.haddad_problem(12, 10)
    out.append(("haddad", PM.HaddadModel(**prm), g, gu, None))
    return out


def dump(path):
    import econpdes_b200 as E
    from econpdes_b200.pdesolve import _bc_array
    res = {}
    for name, pm, g, gu, bc in cases():
        grids = list(g.values())
        S = tuple(len(x) for x in grids)
        y = np.stack([v * (1 + 0.01 * np.random.default_rng(5).uniform(-1, 1, v.shape)) for v in gu.values()], axis=-1)
        bca = _bc_array(bc, list(gu), list(g), S) if bc is not None else None
        h = E.Handle(pm, grids, y, bc=bca)
        res[name + "_G"] = h.residual(y)
        nz, eps = h.jacobian_csc(y)
        res[name + "_J"] = nz
        nz, eps = h.jacobian_csc(y, ypost=y * 1.001, delta=0.7)
        res[name + "_Jimpl"] = nz
        ynew, nl, it = h.implicit_timestep(y, 1.0)
        res[name + "_step"] = ynew
        res[name + "_nl"] = np.array([nl, it])
        h.close()
        r = E.pdesolve(pm, g, gu, bc=bc, verbose=False)
        res[name + "_sol"] = np.stack([r.solution[k] for k in gu], axis=-1)
        res[name + "_log"] = np.array([[row["status"], row["newton_iters"], row["delta"]] for row in r.log])
    np.savez(path, **res)


def compare(a, b):
    A, B = np.load(a), np.load(b)
    bad = 0
    for k in A.files:
        x, y = A[k], B[k]
        if x.shape != y.shape:
            print(k, "SHAPE", x.shape, y.shape)
            bad += 1
            continue
        same = np.array_equal(x, y, equal_nan=True)
        if not same:
            d = np.abs(x - y)
            with np.errstate(all="ignore"):
                rel = np.nanmax(d / np.maximum(np.abs(x), 1e-300))
            idx = np.unravel_index(np.nanargmax(d), d.shape)
            print(f"{k}: DIFFERS  max abs {np.nanmax(d):.3e} rel {rel:.3e} at {idx}  ({x[idx]!r} vs {y[idx]!r}), "
                  f"{np.count_nonzero(~(x == y))} of {x.size} entries")
            bad += 1
        else:
            print(f"{k}: identical")
    return bad


if __name__ == "__main__":
    if sys.argv[1] == "dump":
        dump(sys.argv[2])
    else:
        sys.exit(1 if compare(sys.argv[2], sys.argv[3]) else 0)

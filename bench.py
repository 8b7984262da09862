#!/usr/bin/env python
"""Benchmark of the finite-scheme hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[3], the batched calibration sweep -- 16 384
independent Campbell-Cochrane habit problems of 998 grid points PER GPU (weak scaling: each
rank solves its own contiguous slice of the sweep, no data-path collective).  One "step" is
one full pass of the hot path over the batch: every problem is solved to convergence by the
pseudo-transient Newton loop (4 outer steps / 8 Newton solves each under reference
semantics).  metric = Newton solves per second, whole job.

  value     inputs already resident in HBM, device time (CUDA events on the launching stream)
  e2e       same metric through the C-ABI call with pinned HOST buffers (H2D + solve + D2H)
  roofline  fused batched kernel: algorithmic bytes / launch time vs measured HBM peak
  cpu_baseline  the plain-C CPU oracle (port of the reference algorithm) on a bounded sample, 1 core
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PROBLEMS_PER_GPU = 16384
POINTS = 998
# algorithmic bytes per grid point (SURVEY.md 8(d), N=1, F=1): Jacobian 16F+8KF^2 = 40,
# implicit residual 24F = 24, tridiagonal solve 3 coefficients + rhs + solution = 40
BYTES_PER_NEWTON_POINT = 40 + 24 + 40
BYTES_PER_SWEEP_POINT = 16


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic_bytes(summary, kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of one kernel from the committed
    `ncu --set full` summary under profiles/ (tools/ncu_summary.py); None if absent."""
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        with open(os.path.join(ROOT, "profiles", summary)) as f:
            ks = [k for k in json.load(f)["kernels"] if kernel_substr in k["kernel"]]
        tot = 0.0
        for key in ("dram_read", "dram_write"):
            v, u = ks[0][key].split()
            tot += float(v) * unit[u]
        return tot
    except Exception:
        return None


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def _dist():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def run_ours(args):
    import torch
    import econpdes_b200 as E
    from econpdes_b200 import _lib as L
    from econpdes_b200 import examples as EX
    from econpdes_b200.models import CampbellCochraneModel
    from econpdes_b200.pdesolve import _options

    rank, world, local = _dist()
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist = None
        torch.cuda.set_device(local)
    lib = L.load()
    L.check(lib.econpdes_set_device(local))
    from econpdes_b200.sharding import Reducer, weak_slice
    B, S = args.problems, POINTS
    first, _ = weak_slice(B, rank)
    params, grids, guess = EX.campbell_cochrane_sweep(first, B)
    red = Reducer(dist, "cuda")
    opts = _options(math.sqrt(np.finfo(float).eps), 100, {})
    mid = L.check(lib.econpdes_model_id(CampbellCochraneModel.name.encode()))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    allmax, allsum = red.max, red.sum

    # ---- device-resident leg (value) ----
    h = C.c_void_p()
    L.check(lib.econpdes_batch_create(mid, B, S, C.byref(h)))
    L.check(lib.econpdes_batch_upload(h, L.dptr(params), L.dptr(grids), L.dptr(guess)))
    ms = C.c_double()
    for _ in range(args.warmup):
        L.check(lib.econpdes_batch_solve(h, C.byref(opts), C.byref(ms)))
    launches0 = lib.econpdes_total_kernel_launches()
    barrier()
    dev_ms = 0.0
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            L.check(lib.econpdes_batch_solve(h, C.byref(opts), C.byref(ms)))
            dev_ms += ms.value
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
    launches = lib.econpdes_total_kernel_launches() - launches0
    nn = np.empty(B, dtype=np.int32)
    no = np.empty(B, dtype=np.int32)
    rn = np.empty(B)
    L.check(lib.econpdes_batch_download(h, None, L.dptr(rn), None, no.ctypes.data_as(L.c_int32_p),
                                        nn.ctypes.data_as(L.c_int32_p), None, 0, None))
    lib.econpdes_batch_destroy(h)
    newton_rank = float(nn.sum())
    converged_rank = float((rn <= opts.abstol).sum())
    dev_ms_max = allmax(dev_ms)
    wall_ms_max = allmax(wall_ms)
    newton_total = allsum(newton_rank)
    converged_total = allsum(converged_rank)
    ms_per_step = dev_ms_max / args.steps
    value = newton_total / (ms_per_step * 1e-3)

    # ---- end-to-end leg through the C-ABI with pinned host buffers ----
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = a
        return t
    tp, tg, tu = pinned(params), pinned(grids), pinned(guess)
    tsol = torch.empty((B, S), dtype=torch.float64, pin_memory=True)
    trn = torch.empty(B, dtype=torch.float64, pin_memory=True)
    ex = np.empty(B, dtype=np.int32)

    def e2e_step():
        L.check(lib.econpdes_solve_batched(mid, B, S, L.dptr(tp.numpy()), L.dptr(tg.numpy()), L.dptr(tu.numpy()),
                                           C.byref(opts), L.dptr(tsol.numpy()), L.dptr(trn.numpy()),
                                           ex.ctypes.data_as(L.c_int32_p), no.ctypes.data_as(L.c_int32_p),
                                           nn.ctypes.data_as(L.c_int32_p)))
    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = allmax((time.perf_counter() - t0) * 1e3) / args.steps
    e2e_value = newton_total / (e2e_ms * 1e-3)
    h2d = int(params.nbytes + grids.nbytes + guess.nbytes)
    d2h = int(tsol.numpy().nbytes + trn.numpy().nbytes + ex.nbytes + no.nbytes + nn.nbytes)

    # ---- config C5: ONE 256^3 grid, slab-decomposed over the same ranks (strong scaling) ----
    c5 = None
    if not args.no_c5:
        c5 = c5_slab(args, dist, rank, world, local, allmax)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = _peaks()
    alg_bytes = float(S) * (BYTES_PER_NEWTON_POINT * newton_rank + BYTES_PER_SWEEP_POINT * float(no.sum() + B))
    kernel_ms = dev_ms / args.steps
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = _ncu_traffic_bytes("r01_batched_kernel_P4_ncu_full.json", "k_batched_pdesolve_1d")
    line = {
        "metric": "newton_solves_per_sec", "value": value, "unit": "Newton solves/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C4 batched calibration sweep: independent Campbell-Cochrane habit 1-D problems, "
                               "full pdesolve per problem",
                   "problems_per_gpu": B, "points_per_problem": S, "problems_total": B * world,
                   "newton_solves_per_step": newton_total, "converged": converged_total,
                   "parallelism": f"batch-sharded x{world} (no collective)",
                   "l2": "inputs larger than L2 (grid+guess 262 MB per step, solution 131 MB)"},
        "wall_ms_per_step": wall_ms_max / args.steps,
        "e2e": {"value": e2e_value, "unit": "Newton solves/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "api": "econpdes_solve_batched (C ABI, pinned host buffers)"},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": {"bound": "hbm", "kernel": "k_batched_pdesolve_1d<CampbellCochrane, P=4>", "achieved": achieved,
                     "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes,
                     "note": "fused on-chip solve: algorithmic bytes never reach HBM; DRAM traffic is grid+guess in, solution out"},
        "pdesolves_per_sec": B * world / (ms_per_step * 1e-3),
    }
    if c5 is not None:
        line["c5_slab"] = c5
    if world == 1 and not args.no_kernels:
        # residual / Jacobian / SpMV / V-cycle of the generic engine on the C3 and C5 shapes
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_kernels
        line["kernels"] = bench_kernels.measure(reps=3)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.cpu_sample, 1)
    _emit(line)
    if dist is not None:
        dist.destroy_process_group()


def c5_slab(args, dist, rank, world, local, allmax):
    """BASELINE.json configs[4]: 3-state HJB on a 256^3 grid, slab-decomposed across the ranks
    (halo planes over NVLink, all-gathered Krylov dot products / norms).  Fixed total work:
    "scaling": "strong".  One warm-up solve (graph capture, NCCL connections), one timed."""
    import econpdes_b200 as E
    from econpdes_b200 import examples as EX
    from econpdes_b200 import models as PM
    n = args.c5_n
    if world > 1:
        grp = E.SlabGroup.from_torch(dist, local)
    else:
        grp = E.SlabGroup(0, 1)
    g, _ = EX.hjb3d_grid(n)
    first, count = E.slab_range(n, rank, world)
    guess = {"v": np.zeros((n, n, count))}
    res = None
    for _ in range(2):
        res = E.pdesolve_slab(PM.Hjb3dSynth(), g, guess, grp)
    ms = allmax(res.device_ms)
    out = {"workload": f"C5 synthetic 3-state HJB, {n}^3 grid, one pdesolve", "n_gpus": world, "scaling": "strong",
           "ms_per_pdesolve": ms, "newton_solves": res.n_newton, "newton_solves_per_sec": res.n_newton / (ms * 1e-3),
           "outer_steps": res.n_outer, "krylov_iterations": res.n_linear_iters, "converged": bool(res.converged),
           "residual_norm": res.residual_norm, "kernel_launches_rank0": res.kernel_launches,
           "nccl_calls_rank0": grp.collectives, "p2p_halo_exchanges_rank0": grp.p2p_exchanges,
           "decomposition": f"{world} slab(s) of {count} planes along the slowest dimension"}
    grp.close()
    return out


def _c_oracle_time(first, count, threads):
    """Times the plain-C oracle (oracle/oracle1d.c: the reference's serial algorithm, restated)
    on `count` problems of the sweep; grids are built before the clock starts, like the GPU leg."""
    from econpdes_b200 import examples as EX
    from oracle import c_oracle
    params, grids, guess = EX.campbell_cochrane_sweep(first, count)
    c_oracle.solve_batch("campbell_cochrane", params[:8], grids[:8], guess[:8], threads=1)   # warm
    t0 = time.perf_counter()
    _, rn, no, nn = c_oracle.solve_batch("campbell_cochrane", params, grids, guess, threads=threads)
    dt = time.perf_counter() - t0
    return float(nn.sum()), dt, bool((rn <= math.sqrt(np.finfo(float).eps)).all())


def cpu_baseline(sample, threads):
    newton, dt, ok = _c_oracle_time(0, sample, threads)
    return {"value": newton / dt, "unit": "Newton solves/s", "cores": threads, "kind": "port",
            "sample": f"{sample} of the {PROBLEMS_PER_GPU} sweep problems (998 points each), full pdesolve by the "
                      f"plain-C oracle (oracle/oracle1d.c), {dt:.1f} s, all converged: {ok}",
            "pdesolves_per_sec": sample / dt, "ms_per_pdesolve": dt / sample * 1e3,
            "reference_published_ms_per_pdesolve": 2.575}


def run_reference(args):
    """The reference algorithm on the box's host cores.  Julia is not installable here (DESIGN.md
    section 8), so this is the CPU oracle port (oracle/oracle1d.c) with one thread per core on a
    bounded sample of the same sweep."""
    rank, world, _ = _dist()
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = max(threads, args.ref_sample)
    for _ in range(min(args.warmup, 1)):
        _c_oracle_time(0, threads * 4, threads)
    tot_newton, tot_t = 0.0, 0.0
    for _ in range(args.steps):
        nwt, dt, _ = _c_oracle_time(0, sample, threads)
        tot_newton += nwt
        tot_t += dt
    value = tot_newton / tot_t
    line = {"impl": "reference", "metric": "newton_solves_per_sec", "value": value, "unit": "Newton solves/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_t / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C4 batched calibration sweep: independent Campbell-Cochrane habit 1-D problems, "
                                   "full pdesolve per problem", "points_per_problem": POINTS,
                       "sample_problems_per_step": sample},
            "cpu_baseline": {"value": value, "unit": "Newton solves/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} sweep problems per step, {threads} threads, plain-C oracle"},
            "e2e": {"value": value, "unit": "Newton solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


class _OneJsonLine:
    """The driver parses stdout as ONE JSON line, but libraries (NCCL's version banner, torch
    warnings) also write to fd 1.  Everything goes to stderr while the benchmark runs; only the
    final line is written to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.saved, (text + "\n").encode())

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


_OUT = None


def _emit(line):
    text = json.dumps(line)
    if _OUT is not None:
        _OUT.emit(text)
    else:
        print(text, flush=True)


def main():
    global _OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--problems", type=int, default=PROBLEMS_PER_GPU, help="problems per GPU")
    ap.add_argument("--cpu-sample", type=int, default=4096)
    ap.add_argument("--ref-sample", type=int, default=16384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernels", action="store_true", help="skip the grid-kernel roofline leg")
    ap.add_argument("--no-c5", action="store_true", help="skip the slab-decomposed 3-D solve")
    ap.add_argument("--c5-n", type=int, default=256, help="points per axis of the C5 grid")
    args = ap.parse_args()
    with _OneJsonLine() as out:
        _OUT = out
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    _OUT = None


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the finite-scheme hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[3] verbatim, the batched calibration sweep -- 16 384
independent Campbell-Cochrane habit problems of 998 grid points IN TOTAL, sharded across the GPUs
("scaling": "strong": rank r of N solves a contiguous slice of 16384/N problems, no data-path
collective).  `--scaling weak` gives every GPU 16 384 problems instead; at N > 1 the default run adds
that measurement as the extra key "weak".  One "step" is one full pass of the hot path over the batch:
every problem is solved to convergence by the pseudo-transient Newton loop (4 outer steps / 8 Newton
solves each under reference semantics).  metric = Newton solves per second, whole job.

  value     inputs already resident in HBM, device time (CUDA events on the launching stream)
  e2e       same metric through the C-ABI call with pinned HOST buffers (H2D + solve + D2H).  The sweep is
            described compactly (econpdes_batch_inputs: per problem six parameters, the lower end of the
            linear grid segment and the 99 log-spaced points; the 899-point range is built on the device,
            the guess is a scalar), the solution comes back in full
  roofline  fused batched kernel: algorithmic bytes / launch time vs measured HBM peak
  cpu_baseline  the plain-C CPU oracle (port of the reference algorithm) on a bounded sample, 1 core
  configs   (N = 1) full pdesolve of configs C2 (64 x 4096) and C3 (1024^2, 2048^2) with outcome and counts
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


def _issue_roofline(summary, kernel_substr, kernel_ms, launch_scale, clocks):
    """The bound the fused kernel really runs against: it keeps the whole solve on chip, so its ceiling is the SMs'
    instruction issue rate (one warp instruction per cycle per scheduler, FP64 ones every second cycle), not HBM.
    Executed warp instructions per launch come from the committed ncu capture (scaled to this launch's problem count),
    the launch time is measured live here.  None if the capture is absent."""
    try:
        with open(os.path.join(ROOT, "profiles", summary)) as f:
            k = [k for k in json.load(f)["kernels"] if kernel_substr in k["kernel"]][0]
        inst = float(k["warp_instructions"].split()[0]) * launch_scale
        mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
        peak = 148 * 4 * mhz * 1e6          # warp instructions per second, all schedulers issuing every cycle
        achieved = inst / (kernel_ms * 1e-3)
        return {"bound": "issue", "unit": "warp instructions/s", "achieved": achieved, "peak": peak, "frac": achieved / peak,
                "warp_instructions_per_launch": inst, "issue_bound_ms": inst / peak * 1e3,
                "ncu_fp64_pipe_active_pct": float(k["fp64_pipe_active_pct"].split()[0]),
                "ncu_issue_active_pct": float(k["issue_active_pct"].split()[0]),
                "ncu_achieved_occupancy_pct": float(k["achieved_occupancy_pct"].split()[0]),
                "source": "profiles/" + summary + " (instruction count) + live CUDA-event time"}
    except Exception:
        return None


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        if self.index < 0:   # not sampling
            return self
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


def _bind_to_gpu_numa(local):
    """Pin this process to the CPU cores next to its GPU (PCI root complex / NUMA node), so the pinned host
    buffers it allocates afterwards live there (first touch) and eight ranks do not share one node's memory
    controllers and PCIe root.  Best effort: a container may expose only part of the machine."""
    info = {"bound": False}
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        with open(base + "/numa_node") as f:
            info["numa_node"] = int(f.read().strip())
        with open(base + "/local_cpulist") as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        info["allowed_cpus"] = len(allowed)
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            info["bound"] = True
        info["cpus"] = len(use) if use else len(allowed)
    except Exception as e:   # no sysfs / no permission: run unbound
        info["error"] = type(e).__name__
    return info


def _copy_bandwidth(torch, nbytes=64 << 20, reps=5):
    """Host<->device copy rate of this rank from page-locked memory (GB/s), all ranks at the same time."""
    h = torch.empty(nbytes // 8, dtype=torch.float64, pin_memory=True)
    d = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
    out = {}
    for name, (src, dst) in (("h2d", (h, d)), ("d2h", (d, h))):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        out[name] = nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    return out


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
    torch.cuda.set_device(local)
    numa = _bind_to_gpu_numa(local)   # before any pinned allocation
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist = None
    lib = L.load()
    L.check(lib.econpdes_set_device(local))
    from econpdes_b200.sharding import Reducer, shard_range, weak_slice
    S = POINTS
    red = Reducer(dist, "cuda")
    opts = _options(math.sqrt(np.finfo(float).eps), 100, {})
    mid = L.check(lib.econpdes_model_id(CampbellCochraneModel.name.encode()))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    allmax, allsum = red.max, red.sum

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = a
        return t

    def measure_sweep(first, B, sample_clocks):
        """Problems [first, first + B) of the sweep on this rank: device-resident leg, then the end-to-end leg."""
        params, grids, guess = EX.campbell_cochrane_sweep(first, B)
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
        with ClockSampler(local if sample_clocks else -1) as clocks:
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
        out = {"B": B, "dev_ms": dev_ms, "newton_rank": float(nn.sum()), "outer_rank": float(no.sum()),
               "converged_rank": float((rn <= opts.abstol).sum()), "launches": int(launches),
               "clocks": clocks.summary() if sample_clocks else None}
        out["dev_ms_max"] = allmax(dev_ms)
        out["wall_ms_max"] = allmax(wall_ms)
        out["newton_total"] = allsum(out["newton_rank"])
        out["converged_total"] = allsum(out["converged_rank"])
        out["problems_total"] = allsum(float(B))
        out["ms_per_step"] = out["dev_ms_max"] / args.steps
        out["value"] = out["newton_total"] / (out["ms_per_step"] * 1e-3)

        # ---- end-to-end leg through the C ABI with pinned host buffers, compact inputs ----
        p2, lo, pts, n_range, fill = EX.campbell_cochrane_sweep_compact(first, B)
        tp, tlo, tpts = pinned(p2), pinned(lo), pinned(pts)
        tsol = torch.empty((B, S), dtype=torch.float64, pin_memory=True)
        trn = torch.empty(B, dtype=torch.float64, pin_memory=True)
        ex = np.empty(B, dtype=np.int32)
        inp = L.BatchInputs()
        inp.params, inp.grids, inp.n_range = L.dptr(tp.numpy()), None, int(n_range)
        inp.range_lo, inp.points, inp.guess, inp.guess_fill = L.dptr(tlo.numpy()), L.dptr(tpts.numpy()), None, float(fill)

        # the argument pointers of the C call, formed once like a C caller's
        e2e_args = (mid, B, S, C.byref(inp), C.byref(opts), L.dptr(tsol.numpy()), L.dptr(trn.numpy()),
                    ex.ctypes.data_as(L.c_int32_p), no.ctypes.data_as(L.c_int32_p), nn.ctypes.data_as(L.c_int32_p))

        def e2e_step():
            L.check(lib.econpdes_solve_batched_inputs(*e2e_args))
        for _ in range(max(1, args.warmup // 2)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        out["e2e_ms"] = allmax((time.perf_counter() - t0) * 1e3) / args.steps
        out["e2e_output_mode"] = int(lib.econpdes_last_batched_output_mode())
        out["e2e_value"] = out["newton_total"] / (out["e2e_ms"] * 1e-3)
        out["h2d"] = int(p2.nbytes + lo.nbytes + pts.nbytes)
        out["d2h"] = int(tsol.numpy().nbytes + trn.numpy().nbytes + ex.nbytes + no.nbytes + nn.nbytes)
        assert float((trn.numpy() <= opts.abstol).sum()) == out["converged_rank"]

        # the same call with explicit grids and guesses (what round 1 measured): B x S doubles twice over PCIe
        if args.e2e_explicit:
            tg, tu = pinned(grids), pinned(guess)

            def e2e_explicit_step():
                L.check(lib.econpdes_solve_batched(mid, B, S, L.dptr(tp.numpy()), L.dptr(tg.numpy()), L.dptr(tu.numpy()),
                                                   C.byref(opts), L.dptr(tsol.numpy()), L.dptr(trn.numpy()),
                                                   ex.ctypes.data_as(L.c_int32_p), no.ctypes.data_as(L.c_int32_p),
                                                   nn.ctypes.data_as(L.c_int32_p)))
            e2e_explicit_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                e2e_explicit_step()
            barrier()
            ems = allmax((time.perf_counter() - t0) * 1e3) / args.steps
            out["e2e_explicit"] = {"value": out["newton_total"] / (ems * 1e-3), "ms_per_step": ems,
                                   "h2d_bytes_per_step": int(params.nbytes + grids.nbytes + guess.nbytes),
                                   "d2h_bytes_per_step": out["d2h"]}
        lib.econpdes_release_workspace()
        return out

    total = args.problems
    if args.scaling == "strong":
        first, last = shard_range(total, rank, world)
    else:
        first, last = weak_slice(total, rank)
    main = measure_sweep(first, last - first, True)
    weak = None
    if args.scaling == "strong" and world > 1 and not args.no_weak:
        wf, wl = weak_slice(total, rank)
        weak = measure_sweep(wf, wl - wf, False)
    barrier()
    bw = _copy_bandwidth(torch)
    bw_min = {k: -allmax(-v) for k, v in bw.items()}
    bw_sum = {k: allsum(v) for k, v in bw.items()}

    # ---- config C5: ONE 256^3 grid, slab-decomposed over the same ranks (strong scaling) ----
    c5 = None
    if not args.no_c5:
        c5 = c5_slab(args, dist, rank, world, local, allmax)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = _peaks()
    B = main["B"]
    alg_bytes = float(S) * (BYTES_PER_NEWTON_POINT * main["newton_rank"] + BYTES_PER_SWEEP_POINT * (main["outer_rank"] + B))
    kernel_ms = main["dev_ms"] / args.steps
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    ncu_file = "r02_batched_kernel_P4_ncu_full.json"
    traffic = _ncu_traffic_bytes(ncu_file, "k_batched_pdesolve_1d")
    if traffic is not None:
        traffic *= B / float(PROBLEMS_PER_GPU)   # the capture is of a 16 384-problem launch
    per = "total" if args.scaling == "strong" else "per GPU"
    line = {
        "metric": "newton_solves_per_sec", "value": main["value"], "unit": "Newton solves/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C4 batched calibration sweep: independent Campbell-Cochrane habit 1-D problems, "
                               "full pdesolve per problem",
                   "problems": f"{total} {per}", "problems_per_gpu": B, "points_per_problem": S,
                   "problems_total": int(main["problems_total"]),
                   "newton_solves_per_step": main["newton_total"], "converged": main["converged_total"],
                   "parallelism": f"batch-sharded x{world} (no collective)",
                   "l2": "inputs larger than L2 between timed steps at N = 1 (grid + guess 262 MB, solution 131 MB per "
                         "16 384 problems); smaller shards are streamed once per step from HBM all the same: nothing is "
                         "re-used across steps on chip"},
        "wall_ms_per_step": main["wall_ms_max"] / args.steps,
        "e2e": {"value": main["e2e_value"], "unit": "Newton solves/s", "ms_per_step": main["e2e_ms"],
                "h2d_bytes_per_step": main["h2d"], "d2h_bytes_per_step": main["d2h"],
                "api": "econpdes_solve_batched_inputs (C ABI, pinned host buffers; compact inputs: parameters + "
                       "grid end points + 99 explicit grid points per problem, scalar guess; full solution back)",
                "results": ("zero-copy: every thread block stores its solution and scalars straight into the caller's "
                            "pinned host buffers over PCIe (the d2h bytes are those stores)"
                            if main.get("e2e_output_mode") == 1 else "staged: device buffers + one device->host copy per slice")},
        "gpu_launches": main["launches"],
        "clocks": main["clocks"],
        "roofline": {"bound": "hbm", "kernel": "k_batched_pdesolve_1d<CampbellCochrane, P=4>", "achieved": achieved,
                     "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes,
                     "note": "fused on-chip solve: algorithmic bytes never reach HBM; DRAM traffic is grid+guess in, solution out"},
        "roofline_fp64_issue": _issue_roofline(ncu_file, "k_batched_pdesolve_1d", kernel_ms, B / float(PROBLEMS_PER_GPU),
                                               main["clocks"]),
        "pdesolves_per_sec": main["problems_total"] / (main["ms_per_step"] * 1e-3),
        "host": {"numa_binding_rank0": numa, "pinned_copy_gbs_per_rank_min": bw_min, "pinned_copy_gbs_all_ranks": bw_sum},
    }
    if "e2e_explicit" in main:
        line["e2e_explicit_inputs"] = main["e2e_explicit"]
    if weak is not None:
        line["weak"] = {"scaling": "weak", "problems_per_gpu": weak["B"], "value": weak["value"],
                        "ms_per_step": weak["ms_per_step"], "e2e_value": weak["e2e_value"], "e2e_ms_per_step": weak["e2e_ms"],
                        "h2d_bytes_per_step": weak["h2d"], "d2h_bytes_per_step": weak["d2h"]}
    if c5 is not None:
        line["c5_slab"] = c5
    if world == 1 and not args.no_kernels:
        # residual / Jacobian / SpMV / V-cycle of the generic engine on the C3 and C5 shapes
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_kernels
        line["kernels"] = bench_kernels.measure(reps=3)
    if world == 1 and not args.no_configs:
        line["configs"] = full_configs()
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.cpu_sample, 1)
    _emit(line)
    if dist is not None:
        dist.destroy_process_group()


def full_configs():
    """Full `pdesolve` of BASELINE.json configs[1] (consumption-saving, 64 x 4096) and configs[2] (Bansal-Yaron at
    1024^2 and 2048^2) under reference semantics, plus the 2048^2 run with per-column finite-difference steps
    (documented robust option, not reference parity): outcome, counts, device time, quality of the Krylov solves."""
    import econpdes_b200 as E
    from econpdes_b200 import _lib as L
    from econpdes_b200 import examples as EX
    from econpdes_b200 import models as PM
    out = {}

    def run(tag, model, g, gu, **kw):
        t0 = time.perf_counter()
        r = E.pdesolve(model, g, gu, verbose=False, warn=False, **kw)
        wall = (time.perf_counter() - t0) * 1e3
        out[tag] = {"converged": bool(r.converged), "residual_norm": float(r.residual_norm), "exit_reason": int(r.exit_reason),
                    "outer_steps": int(r.n_outer), "rejected_steps": int(sum(1 for row in r.log if row["status"] != 0)),
                    "newton_solves": int(r.n_newton), "krylov_iterations": int(r.n_linear_iters),
                    "linear_solves": int(r.n_linear_solves), "linear_solves_above_1e-11": int(r.n_linear_loose),
                    "max_linear_relres": float(r.lin_max_relres), "device_ms": float(r.device_ms), "wall_ms": wall,
                    "kernel_launches": int(r.kernel_launches)}

    g, gu = EX.ahlm_diffusion_grid(64, 4096)
    run("C2 consumption-saving 64x4096", PM.AchdouHanLasryLionsMollModel_Diffusion(), g, gu)
    for n in (1024, 2048):
        g, gu = EX.bansal_yaron_grid(n, n)
        run(f"C3 bansal-yaron {n}x{n}", PM.BansalYaronModel(), g, gu)
    g, gu = EX.bansal_yaron_grid(2048, 2048)
    run("C3 bansal-yaron 2048x2048 eps_policy=per_column", PM.BansalYaronModel(), g, gu, eps_policy=L.EPS_PER_COLUMN)
    return out


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
                                   "full pdesolve per problem",
                       "problems": f"{args.problems} total", "problems_per_gpu": sample, "points_per_problem": POINTS,
                       "problems_total": sample,
                       "parallelism": f"{threads} host threads, one problem each at a time (no GPU)",
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
    ap.add_argument("--problems", type=int, default=PROBLEMS_PER_GPU, help="problems of the sweep (see --scaling)")
    ap.add_argument("--cpu-sample", type=int, default=4096)
    ap.add_argument("--ref-sample", type=int, default=16384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernels", action="store_true", help="skip the grid-kernel roofline leg")
    ap.add_argument("--no-c5", action="store_true", help="skip the slab-decomposed 3-D solve")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: --problems in total, sharded over the GPUs (BASELINE config 4); weak: --problems per GPU")
    ap.add_argument("--no-weak", action="store_true", help="N > 1, strong scaling: skip the extra weak-scaling measurement")
    ap.add_argument("--no-configs", action="store_true", help="skip the full C2 / C3 solves (N = 1)")
    ap.add_argument("--e2e-explicit", action="store_true",
                    help="also time the end-to-end call with explicit B x S grids and guesses")
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

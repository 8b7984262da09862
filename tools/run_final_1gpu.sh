#!/usr/bin/env bash
# End-of-round record on ONE GPU: smoke, the default bench line, the reference arm, the ncu launch list of a short run.
out=${1:-gpurun_out/final1}
mkdir -p "$out"
python __graft_entry__.py smoke > "$out/smoke.log" 2>&1; echo "smoke rc=$?"
python bench.py > "$out/bench.json" 2> "$out/bench.err"; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 3 > "$out/bench_reference.json" 2> "$out/bench_reference.err"; echo "reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches.csv" \
  python bench.py --steps 2 --warmup 3 --no-c5 --no-kernels --no-cpu-baseline --no-configs > "$out/launches.run.log" 2>&1; echo "launch list rc=$?"
python - "$out" <<'PY'
import json, sys
d = json.load(open(sys.argv[1] + "/bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"],
      d.get("roofline_fp64_issue", {}).get("frac"), d["cpu_baseline"]["value"], d["clocks"])
print(d.get("c5_slab", {}).get("ms_per_pdesolve"))
for k, v in d.get("configs", {}).items():
    print(k, v.get("converged"), v.get("outer_steps"), v.get("newton_solves"), v.get("krylov_iterations"), round(v.get("device_ms", 0), 1))
r = json.load(open(sys.argv[1] + "/bench_reference.json"))
print("reference", r["value"], r["unit"], r.get("cpu_baseline"))
PY

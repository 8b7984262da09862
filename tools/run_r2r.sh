#!/usr/bin/env bash
out=${1:-gpurun_out/r2r}
mkdir -p "$out"
for mp in 1200 5000 17000 40000; do
  ECONPDES_COARSE_MAX_POINTS=$mp python tools/bench_kernels.py > "$out/kernels_mp$mp.json" 2> "$out/kernels_mp$mp.err"
  python - "$out/kernels_mp$mp.json" $mp <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
for k, v in d.items():
    if isinstance(v, dict) and "vcycle_ms" in v:
        print("max_points", sys.argv[2], k, "vcycle_ms", round(v["vcycle_ms"], 4), "mg_setup_ms", round(v["mg_setup_ms"], 3))
PY
done

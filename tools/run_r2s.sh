#!/usr/bin/env bash
out=${1:-gpurun_out/r2s}
mkdir -p "$out"
t=/tmp/r2s; mkdir -p $t
ECONPDES_COARSE_MAX_POINTS=${2:-17000} timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_coarse_cycle -c 2 -o $t/coarse \
  python tools/profile_kernels.py 2048 0 > "$out/ncu_coarse.log" 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py $t/coarse.ncu-rep "$out/coarse_ncu_full.json" >> "$out/ncu_coarse.log" 2>&1
python tools/ncu_lines.py $t/coarse.ncu-rep 30 > "$out/coarse_ncu_lines.txt" 2>&1
python - "$out/coarse_ncu_full.json" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
for k in d["kernels"]:
    print({a: b for a, b in k.items() if a != "kernel"})
PY
head -34 "$out/coarse_ncu_lines.txt" | cut -c1-150

#!/usr/bin/env bash
out=${1:-gpurun_out/r3i}
mkdir -p "$out"
t=/tmp/r3i; mkdir -p $t
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_batched_pdesolve_1d -s 1 -c 1 -o $t/batched \
  python tools/bench_batched.py --problems 16384 --reps 1 --variants 0 > "$out/ncu_batched.log" 2>&1; echo "ncu batched rc=$?"
python tools/ncu_summary.py $t/batched.ncu-rep "$out/batched_ncu_full.json" >> "$out/ncu_batched.log" 2>&1
python tools/ncu_lines.py $t/batched.ncu-rep 70 > "$out/batched_ncu_lines.txt" 2>&1
python - "$out/batched_ncu_full.json" <<'PY'
import json, sys
k = json.load(open(sys.argv[1]))["kernels"][0]
print({a: b for a, b in k.items() if a != "kernel"})
PY
head -60 "$out/batched_ncu_lines.txt" | cut -c1-150

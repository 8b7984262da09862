#!/usr/bin/env bash
# A/B of the marching sweeps (sweep.cuh) against the one-thread-per-point kernels on the C3 / C5 shapes:
#   gpurun --timeout 900 -- 'bash tools/sweep_march.sh gpurun_out/march'
# One line per configuration: residual / implicit residual + norms / Jacobian ms at 2048^2 and 256^3.
set -uo pipefail
out=${1:-gpurun_out/march}
mkdir -p "$out"
: > "$out/summary.txt"
run() {   # run <tag> VAR=VAL ...
  local tag=$1
  shift
  env "$@" timeout 120 python tools/bench_kernels.py > "$out/$tag.json" 2> "$out/$tag.err"
  python - "$out/$tag.json" "$tag" >> "$out/summary.txt" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    parts = []
    for k, v in d.items():
        parts.append("%s: res %.4f impl %.4f jac %.4f spmv %.4f vc %.3f setup %.2f" % (k.split()[0], v["residual"]["ms"], v["implicit_residual_norms"]["ms"], v["jacobian"]["ms"], v["spmv"]["ms"], v["vcycle_ms"], v["mg_setup_ms"]))
    print(sys.argv[2], " | ".join(parts))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run old ECONPDES_MARCH=0
for v in 0 1 2 3; do
  run "v${v}" ECONPDES_MARCH=1 ECONPDES_MARCH_VARIANT=$v
done
run "v1_L16" ECONPDES_MARCH=1 ECONPDES_MARCH_VARIANT=1 ECONPDES_MARCH_L=16
run "v1_L32" ECONPDES_MARCH=1 ECONPDES_MARCH_VARIANT=1 ECONPDES_MARCH_L=32
run "v3_L16" ECONPDES_MARCH=1 ECONPDES_MARCH_VARIANT=3 ECONPDES_MARCH_L=16
run "v3_L32" ECONPDES_MARCH=1 ECONPDES_MARCH_VARIANT=3 ECONPDES_MARCH_L=32
run "v3_occ3" ECONPDES_MARCH=1 ECONPDES_MARCH_VARIANT=3 ECONPDES_MARCH_OCC=3
run "jocc" ECONPDES_MARCH=1 ECONPDES_MARCH_VARIANT=3 ECONPDES_MARCH_JOCC=1
run "jocc_L16" ECONPDES_MARCH=1 ECONPDES_MARCH_VARIANT=3 ECONPDES_MARCH_JOCC=1 ECONPDES_MARCH_LJ=16
cat "$out/summary.txt"

#!/usr/bin/env bash
# A/B of kernel variants on the C3 / C5 shapes (one line per configuration):
#   gpurun --timeout 900 -- 'bash tools/sweep_march.sh gpurun_out/march'
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
for v in 0 1 2 3; do
  run "v${v}" ECONPDES_MARCH_VARIANT=$v
done
run "v1_L16" ECONPDES_MARCH_VARIANT=1 ECONPDES_MARCH_L=16
run "v1_L32" ECONPDES_MARCH_VARIANT=1 ECONPDES_MARCH_L=32
run "warp256" ECONPDES_WARP_LINES_MAX=256 ECONPDES_CHUNK_LINES_3D=0
run "nochunk3d" ECONPDES_CHUNK_LINES_3D=0
cat "$out/summary.txt"

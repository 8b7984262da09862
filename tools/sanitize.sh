#!/usr/bin/env bash
# compute-sanitizer passes (memcheck, racecheck, synccheck, initcheck) over the small-grid device
# parity tests: the kernels with shared-memory hand-offs are the batched 1-D solve (block-wide PCR,
# block reductions), the (block-)tridiagonal PCR, the zebra line smoothers and the reduction tails of
# the residual / color-norm kernels.  The selection keeps every grid small: the sanitizer slows a
# kernel by 10-100x.
#
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh gpurun_out/sanitize'
#
# Writes <out>/<tool>.log (sanitizer report), <out>/<tool>.pytest.log and <out>/summary.txt
# (one line per tool: return code of the run; 0 = tests passed and no sanitizer error).
# NOT YET RUN: the round-1 GPU budget was spent before this script existed.
set -uo pipefail
out=${1:-gpurun_out/sanitize}
mkdir -p "$out"
: > "$out/summary.txt"
sel='readme_quickstart or two_income or garleanu or he_krishnamurthy or wachter or batched_small or batched_inf'
sel="$sel or signed_per_function or mcp_lower_bound or linear_sum_1d or pdesolve_2d_known or hump"
sel="$sel or (pdesolve_nd_against_golden and (bansal_yaron_30 or ahlm_10x100 or hjb3d_12))"
for tool in memcheck racecheck synccheck initcheck; do
  timeout ${SAN_LIMIT:-1200} compute-sanitizer --tool "$tool" --error-exitcode 9 --log-file "$out/$tool.log" \
    python -m pytest tests/test_gpu_parity.py -q -x -k "$sel" > "$out/$tool.pytest.log" 2>&1
  echo "$tool rc=$?" >> "$out/summary.txt"
done
cat "$out/summary.txt"

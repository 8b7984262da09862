#!/usr/bin/env bash
# Everything a round wants from its FIRST GPU call, in one command (1 GPU, about 10-15 minutes):
#
#   gpurun --timeout 1500 -- 'bash tools/first_call.sh gpurun_out/first'
#
#   1. the full device parity suite                              -> <out>/pytest_gpu.log
#   2. smoke()                                                   -> <out>/smoke.log
#   3. bench.py with its default flags, and the reference arm    -> <out>/bench.json, <out>/bench_reference.json
#   4. grid-kernel timings at the C3 / C5 shapes                 -> <out>/kernels.json
#   5. the ncu launch list of a short bench run (launch shares)  -> <out>/launches.csv
#   6. compute-sanitizer over the small-grid parity tests        -> <out>/sanitize/
#
# Each step has its own time limit and the script carries on after a failure; <out>/summary.txt holds one
# "step rc=N seconds=S" line per step.  Nothing here is a bench value if it ran under ncu or the sanitizer.
set -uo pipefail
out=${1:-gpurun_out/first}
mkdir -p "$out"
: > "$out/summary.txt"

step() {   # step <name> <limit-seconds> <command...>
  local name=$1 limit=$2
  shift 2
  local t0=$SECONDS
  timeout "$limit" "$@"
  echo "$name rc=$? seconds=$((SECONDS - t0))" >> "$out/summary.txt"
}

step pytest_gpu 600 bash -c "python -m pytest tests -x -q -m gpu > '$out/pytest_gpu.log' 2>&1"
step smoke 120 bash -c "python __graft_entry__.py smoke > '$out/smoke.log' 2>&1"
step bench 420 bash -c "python bench.py > '$out/bench.json' 2> '$out/bench.err'"
step bench_reference 300 bash -c "python bench.py --impl reference --steps 5 --warmup 3 > '$out/bench_reference.json' 2> '$out/bench_reference.err'"
step kernels 180 bash -c "python tools/bench_kernels.py > '$out/kernels.json' 2> '$out/kernels.err'"
step launches 300 bash -c "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file '$out/launches.csv' \
  python bench.py --steps 2 --warmup 1 --no-c5 --no-kernels --no-cpu-baseline > '$out/launches.run.log' 2>&1"
step sanitize 1300 bash tools/sanitize.sh "$out/sanitize"
cat "$out/summary.txt"

#!/usr/bin/env bash
# one GPU call: device parity suite, then the C4 bench with zero-copy results and with staged results
out=${1:-gpurun_out/r2n}
mkdir -p "$out"
python -m pytest tests -x -q -m gpu > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?"
python bench.py --no-c5 --no-kernels --no-configs --no-cpu-baseline --e2e-explicit > "$out/bench_direct.json" 2> "$out/bench_direct.err"; echo "bench rc=$?"
ECONPDES_DIRECT_OUTPUT=0 python bench.py --no-c5 --no-kernels --no-configs --no-cpu-baseline --e2e-explicit > "$out/bench_staged.json" 2> "$out/bench_staged.err"; echo "bench staged rc=$?"
tail -5 "$out/pytest_gpu.log"
python - "$out" <<'PY'
import json, sys
for f in ("direct", "staged"):
    try:
        d = json.load(open("%s/bench_%s.json" % (sys.argv[1], f)))
        print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "explicit",
              d.get("e2e_explicit_inputs", {}).get("ms_per_step"), "frac", d["roofline"]["frac"])
    except Exception as e:
        print(f, "failed", e)
PY

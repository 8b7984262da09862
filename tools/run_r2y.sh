#!/usr/bin/env bash
out=${1:-gpurun_out/r2y}
mkdir -p "$out"
python -m pytest tests -x -q -m gpu > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -3 "$out/pytest_gpu.log"
python tools/bench_kernels.py > "$out/kernels.json" 2> "$out/kernels.err"
python - "$out/kernels.json" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
for k, v in d.items():
    if isinstance(v, dict) and "vcycle_ms" in v:
        print(k, "vcycle_ms", round(v["vcycle_ms"], 4), "mg_setup_ms", round(v["mg_setup_ms"], 3))
PY
for rep in 1 2; do python tools/run_configs.py c2 c3 c5 --n2d=1024 > "$out/configs_$rep.log" 2>&1; done
python - "$out" <<'PY'
import json, sys, collections
best = collections.defaultdict(lambda: 1e30); its = {}
for rep in (1, 2):
    for line in open("%s/configs_%d.log" % (sys.argv[1], rep)):
        try:
            d = json.loads(line)
        except Exception:
            continue
        best[d["config"]] = min(best[d["config"]], d["device_ms"]); its[d["config"]] = (d["n_newton"], d["n_linear_iters"], d["converged"])
for k, v in sorted(best.items()):
    print(k, round(v, 1), its[k])
PY

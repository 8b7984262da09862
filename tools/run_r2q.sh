#!/usr/bin/env bash
# coarse-cycle cluster kernel: parity suite, then V-cycle / full-solve timings with and without it
out=${1:-gpurun_out/r2q}
mkdir -p "$out"
python -m pytest tests -x -q -m gpu > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -4 "$out/pytest_gpu.log"
for mode in 1 0; do
  ECONPDES_COARSE_CLUSTER=$mode python tools/bench_kernels.py > "$out/kernels_cc$mode.json" 2> "$out/kernels_cc$mode.err"; echo "kernels cc=$mode rc=$?"
  ECONPDES_COARSE_CLUSTER=$mode python tools/run_configs.py c2 c3 c5 --n2d=1024 > "$out/configs_cc$mode.log" 2>&1; echo "configs cc=$mode rc=$?"
done
python - "$out" <<'PY'
import json, sys
for m in (1, 0):
    d = json.load(open("%s/kernels_cc%d.json" % (sys.argv[1], m)))
    for k, v in d.items():
        if isinstance(v, dict) and "vcycle_ms" in v:
            print("cc", m, k, "vcycle_ms", v["vcycle_ms"], "mg_setup_ms", v["mg_setup_ms"])
PY
python - "$out" <<'PY'
import json, sys
for m in (1, 0):
    for line in open("%s/configs_cc%d.log" % (sys.argv[1], m)):
        try:
            d = json.loads(line)
        except Exception:
            print(line.strip()[:200]); continue
        print("cc", m, d["config"], d["converged"], "outer", d["n_outer"], "newton", d["n_newton"], "krylov", d["n_linear_iters"],
              "ms", round(d["device_ms"], 1), "launches", d["kernel_launches"], "resid", d["residual_norm"])
PY

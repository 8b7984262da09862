#!/usr/bin/env bash
out=${1:-gpurun_out/r2t}
mkdir -p "$out"
python -m pytest tests -x -q -m gpu > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -3 "$out/pytest_gpu.log"
for mode in 1 0; do
  ECONPDES_DEVICE_GS=$mode python tools/run_configs.py c2 c3 c5 --n2d=1024 > "$out/configs_gs$mode.log" 2>&1; echo "configs gs=$mode rc=$?"
done
python - "$out" <<'PY'
import json, sys
for m in (1, 0):
    for line in open("%s/configs_gs%d.log" % (sys.argv[1], m)):
        try:
            d = json.loads(line)
        except Exception:
            print(line.strip()[:200]); continue
        print("gs", m, d["config"], d["converged"], "outer", d["n_outer"], "newton", d["n_newton"], "krylov", d["n_linear_iters"],
              "ms", round(d["device_ms"], 1), "wall", round(d["wall_ms"], 1), "launches", d["kernel_launches"], "resid", d["residual_norm"])
PY
python tools/run_lrr.py 2>&1 | grep "^converged" | cut -c1-300

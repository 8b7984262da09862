#!/usr/bin/env bash
out=${1:-gpurun_out/r2u}
mkdir -p "$out"
python tools/run_configs.py c2 > /dev/null 2>&1   # warm the box
for rep in 1 2 3; do
for mode in 1 0; do
  env ${ABVAR:-ECONPDES_DEVICE_GS}=$mode python tools/run_configs.py c2 c3 c5 --n2d=1024 > "$out/configs_gs${mode}_$rep.log" 2>&1
done
done
python - "$out" <<'PY'
import json, sys, collections
best = collections.defaultdict(lambda: 1e30)
for m in (1, 0):
    for rep in (1, 2, 3):
        for line in open("%s/configs_gs%d_%d.log" % (sys.argv[1], m, rep)):
            try:
                d = json.loads(line)
            except Exception:
                continue
            best[(m, d["config"])] = min(best[(m, d["config"])], d["device_ms"])
for k, v in sorted(best.items()):
    print(k, round(v, 1))
PY

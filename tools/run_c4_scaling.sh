#!/usr/bin/env bash
# C4 strong-scaling line at N GPUs (torchrun), zero-copy results and staged results
n=${1:-8}; out=${2:-gpurun_out/r2o}
mkdir -p "$out"
for mode in 1 0; do
  ECONPDES_DIRECT_OUTPUT=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --no-c5 --no-kernels --no-configs --no-cpu-baseline --no-weak > "$out/bench_n${n}_direct${mode}.json" 2> "$out/bench_n${n}_direct${mode}.err"
  echo "n=$n direct=$mode rc=$?"
done
python - "$out" "$n" <<'PY'
import json, sys
for m in (1, 0):
    try:
        d = json.load(open("%s/bench_n%s_direct%d.json" % (sys.argv[1], sys.argv[2], m)))
        print("direct", m, "value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["host"])
    except Exception as e:
        print(m, "failed", e)
PY

#!/usr/bin/env bash
# End-of-round record on N GPUs: (N = 2: the device parity suite with the slab tests first), then the bench line under torchrun.
n=${1:-2}; out=${2:-gpurun_out/final$n}
mkdir -p "$out"
if [ "$n" = "2" ]; then
  python -m pytest tests -q -m gpu > "$out/pytest_gpu_2gpu.log" 2>&1; echo "pytest rc=$?"; tail -3 "$out/pytest_gpu_2gpu.log"
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $n > "$out/bench_n$n.json" 2> "$out/bench_n$n.err"; echo "bench n=$n rc=$?"
python - "$out/bench_n$n.json" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print({k: d[k] for k in ("n_gpus", "value", "ms_per_step", "scaling")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["host"]["pinned_copy_gbs_per_rank_min"])
print("c5", d.get("c5_slab", {}).get("ms_per_pdesolve"), d.get("c5_slab", {}).get("krylov_iterations"), d.get("weak", {}))
PY

#!/usr/bin/env bash
# Round-2 profile call (1 GPU): ncu --set full of the batched kernel and of the grid / multigrid kernels at the
# C3 (2048^2) and C5 (256^3) shapes, summarised on the box (the .ncu-rep files stay there), then kernel timings.
out=${1:-gpurun_out/r2p}
mkdir -p "$out"
t=/tmp/r2p; mkdir -p $t
python tools/bench_batched.py --problems 16384 --reps 2 --variants 0 > "$out/bench_batched.txt" 2>&1; echo "bench_batched rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_batched_pdesolve_1d -s 1 -c 1 -o $t/batched \
  python tools/bench_batched.py --problems 16384 --reps 1 --variants 0 > "$out/ncu_batched.log" 2>&1; echo "ncu batched rc=$?"
python tools/ncu_summary.py $t/batched.ncu-rep "$out/batched_ncu_full.json" >> "$out/ncu_batched.log" 2>&1
python tools/ncu_lines.py $t/batched.ncu-rep 60 > "$out/batched_ncu_lines.txt" 2>&1
python tools/profile_kernels.py 2048 256 > "$out/profile_kernels.txt" 2>&1; echo "profile_kernels rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'k_residual_march|k_jacobian_march|k_line_zebra_chunk|k_rap_1d|k_prolong_add|k_restrict|k_spmv' -c 70 -o $t/grid \
  python tools/profile_kernels.py 2048 256 > "$out/ncu_grid.log" 2>&1; echo "ncu grid rc=$?"
python tools/ncu_summary.py $t/grid.ncu-rep "$out/grid_ncu_full.json" >> "$out/ncu_grid.log" 2>&1
for i in 0 1 2 3; do python tools/ncu_lines.py $t/grid.ncu-rep 25 $i > "$out/grid_ncu_lines_$i.txt" 2>&1; done
ls -la $t
python tools/bench_kernels.py > "$out/kernels.json" 2> "$out/kernels.err"; echo "bench_kernels rc=$?"

#!/usr/bin/env bash
# ncu --set full of the 3-D grid / multigrid kernels at 256^3 (the 2-D capture is run_r2p.sh)
out=${1:-gpurun_out/r2w}
mkdir -p "$out"
t=/tmp/r2w; mkdir -p $t
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'k_residual_march|k_jacobian_march|k_line_zebra_chunk|k_rap_1d|k_prolong_add|k_restrict|k_spmv|k_transpose' -c 48 -o $t/grid3d \
  python tools/profile_kernels.py 0 256 > "$out/ncu_grid3d.log" 2>&1; echo "ncu grid3d rc=$?"
python tools/ncu_summary.py $t/grid3d.ncu-rep "$out/grid3d_ncu_full.json" >> "$out/ncu_grid3d.log" 2>&1
for i in 0 2 3; do python tools/ncu_lines.py $t/grid3d.ncu-rep 25 $i > "$out/grid3d_ncu_lines_$i.txt" 2>&1; done
ls -la $t

#!/usr/bin/env bash
# ncu --set full of the 3-D V-cycle kernels at 256^3 (line smoother, transfers, SpMV)
out=${1:-gpurun_out/r2x}
mkdir -p "$out"
t=/tmp/r2x; mkdir -p $t
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'k_line_zebra_chunk|k_prolong_add|k_restrict|k_spmv' -c 14 -o $t/vc3d \
  python tools/profile_kernels.py 0 256 > "$out/ncu_vc3d.log" 2>&1; echo "ncu vc3d rc=$?"
python tools/ncu_summary.py $t/vc3d.ncu-rep "$out/vcycle3d_ncu_full.json" >> "$out/ncu_vc3d.log" 2>&1
python tools/ncu_lines.py $t/vc3d.ncu-rep 25 1 > "$out/vcycle3d_ncu_lines_1.txt" 2>&1

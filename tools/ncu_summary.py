"""Summarise an .ncu-rep (ncu --set full) into a small JSON committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_x.json
"""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_per_block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_inst_issue_pct_of_peak",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__cycles_elapsed.avg": "sm_cycles_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefront_pct",
    "smsp__sass_inst_executed_op_local_ld.sum": "local_loads",
    "smsp__sass_inst_executed_op_local_st.sum": "local_stores",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
}
STALLS = ["wait", "long_scoreboard", "short_scoreboard", "math_pipe_throttle", "barrier", "mio_throttle",
          "not_selected", "branch_resolving", "lg_throttle", "dispatch_stall", "no_instruction"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    kernels = []
    for r in rows[2:]:
        k = {"kernel": r[col["Kernel Name"]]}
        for key, name in KEYS.items():
            if key in col:
                k[name] = f"{r[col[key]]} {units[col[key]]}".strip()
        st = {}
        for s in STALLS:
            key = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if key in col:
                st[s] = round(float(r[col[key]]), 3)
        k["stall_warps_per_issue"] = st
        kernels.append(k)
    json.dump({"report": rep, "kernels": kernels}, open(out, "w"), indent=1)
    print(f"{len(kernels)} kernel launches summarised -> {out}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

#!/usr/bin/env python
"""Prints the metrics we track from an ncu report (raw page).  usage: ncu_summary.py file.ncu-rep [row]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "inst_executed", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__issue_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "TriageCompute.l1tex__data_pipe_lsu_wavefronts", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active", "sm__inst_executed_pipe_fp64",
        "smsp__average_warp", "smsp__warp_issue_stalled", "smsp__average_warps_issue_stalled", "lts__t_sectors_op_red", "lts__t_sectors_op_atom", "sm__throughput.avg.pct",
        "l1tex__throughput.avg.pct", "lts__throughput.avg.pct", "smsp__inst_executed_pipe_lsu", "smsp__inst_executed_pipe_uniform", "smsp__inst_executed_pipe_fp64", "smsp__inst_executed_pipe_alu", "smsp__inst_executed_pipe_fma",
        "smsp__warps_eligible", "smsp__issue_inst0", "l1tex__lsuin_requests", "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_red", "smsp__inst_executed_op_shared",
        "smsp__inst_executed_op_global", "l1tex__t_requests_pipe_lsu_mem_global_op_red", "l1tex__t_sectors_pipe_lsu_mem_global_op_red", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld", "l1tex__t_requests_pipe_lsu_mem_global_op_ld",
        "sm__cycles_elapsed.avg ", "sm__cycles_active.avg"]


def main():
    rep = sys.argv[1]
    row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2 + row]
    for h, u, v in zip(hdr, units, data):
        if h in ("Kernel Name",) or any(k in h for k in KEYS):
            if "device__attribute" in h:
                continue
            print(f"{h} [{u}] = {v}")


if __name__ == "__main__":
    main()

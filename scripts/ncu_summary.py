#!/usr/bin/env python
"""Text summary of one `ncu --set full` report (the metrics DESIGN.md quotes + warp stall reasons).
Usage: python scripts/ncu_summary.py report.ncu-rep "header line" > profiles/<name>.txt"""
import csv
import subprocess
import sys

KEYS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main(path, header):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, val = rows[0], rows[1], rows[-1]
    print("# " + header)
    print("# kernel: " + val[hdr.index("Kernel Name")])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k} [{units[i]}] = {val[i]}")
    print("# warp stall reasons (issue-slot ratio per issued instruction)")
    stalls = []
    for i, c in enumerate(hdr):
        if c.startswith("smsp__average_warp") and "issue_stalled" in c and "not_issued" not in c and c.endswith("per_issue_active.ratio"):
            try:
                stalls.append((float(val[i]), c.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    for v, name in sorted(stalls, reverse=True)[:10]:
        print(f"stalled_{name} = {v:.6f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")

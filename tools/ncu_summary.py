#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU): per-kernel duration, DRAM bytes, occupancy, issue and stall mix.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/<name>.txt]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__cycles_active.avg",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    stall = [c for c in h if c.startswith("smsp__average_warps_issue_stalled") and c.endswith("_per_issue_active.ratio")] or \
            [c for c in h if "warp_issue_stalled" in c and c.endswith("per_warp_active.pct")]
    for r in rows[2:]:
        print("=" * 100)
        print(r[h.index("Kernel Name")][:90], " id", r[h.index("ID")])
        for w in WANT:
            if w in h:
                print(f"  {w:72s} {r[h.index(w)]:>16s} {units[h.index(w)]}")
        st = sorted(((float(r[h.index(c)].replace(',', '') or 0), c) for c in stall), reverse=True)[:8]
        print("  top stall reasons:")
        for v, c in st:
            print(f"    {v:10.3f}  {c}")


if __name__ == "__main__":
    main()

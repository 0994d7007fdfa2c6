#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (ncu --set full) as the small text files under profiles/:
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_kernel_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "sm__cycles_elapsed.max", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_inst_executed_op_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct")


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("== launch ==")
        print("Kernel Name [] =", d.get("Kernel Name"))
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")
                             and "not_issued" not in h):
                print(f"{h} [{u}] = {v}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Markdown summary of an ncu report: one column per captured kernel launch.
usage: ncu_summary.py <report.ncu-rep> > profiles/<name>.md"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM limit (shared memory)"),
    ("launch__occupancy_limit_registers", "CTAs/SM limit (registers)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data pipe % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "LSU wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "  of which shared memory"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "  shared-memory bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("gpc__cycles_elapsed.avg.per_second", "SM clock"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "barrier", "mio_throttle", "lg_throttle",
          "math_pipe_throttle", "wait", "not_selected", "branch_resolving", "no_instruction", "selected"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    names = [r[col["Kernel Name"]].replace("void ", "").split("(")[0] for r in data]
    print("| metric | unit | " + " | ".join(names) + " |")
    print("|---|---|" + "---|" * len(names))
    for key, label in METRICS:
        if key in col:
            vals = [r[col[key]] for r in data]
            print(f"| {label} (`{key}`) | {units[col[key]]} | " + " | ".join(vals) + " |")
    for s in STALLS:
        key = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
        if key in col:
            vals = [f"{float(r[col[key]]):.2f}" for r in data]
            print(f"| stall {s} | warps per issue | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()

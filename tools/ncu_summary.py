#!/usr/bin/env python
"""Print the handful of ncu raw-page metrics the design decisions rest on.
usage: tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            print("==", path, d.get("Kernel Name", "")[:90])
            for k in WANT:
                if k in d:
                    print("  %-68s %s %s" % (k, d[k], u[k]))
            stalls = sorted(((float(v), k[len(STALL):-len("_per_issue_active.ratio")])
                             for k, v in d.items()
                             if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v),
                            reverse=True)
            print("  stalls per issue: " + ", ".join("%s %.2f" % (k, v) for v, k in stalls[:8]))


if __name__ == "__main__":
    main()

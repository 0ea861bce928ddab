#!/usr/bin/env python
"""Condense an `ncu --set full` report into the handful of raw metrics DESIGN.md and bench.py's roofline quote.

usage: python profiles/summarize_ncu.py gpurun_out/k2_full_r1.ncu-rep > profiles/r1_k2_ncu_full.txt
(needs the `ncu` CLI only to read the report; no GPU)."""
import csv
import io
import subprocess
import sys

KEYS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            print(f"== {path}")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    print(f"{k}\t{r[i]}\t{units[i]}")


if __name__ == "__main__":
    main()

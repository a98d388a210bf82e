#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep [OUT.csv] : the key metrics of an `ncu --set full` capture as the
metric x kernel table kept under profiles/ (one column per captured launch).  Reads the report with
`ncu -i ... --page raw --csv`; runs where ncu is installed (no GPU needed)."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_write.sum.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
]


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    names = [r[ki] for r in data]
    out = io.StringIO()
    w = csv.writer(out)
    w.writerow(["metric", "unit"] + names)
    for key in KEYS:
        cols = [i for i, h in enumerate(hdr) if h == key or h.endswith("." + key)]
        if not cols:
            continue
        c = cols[0]
        w.writerow([key, units[c]] + [r[c] for r in data])
    s = out.getvalue()
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            f.write(s)
    sys.stdout.write(s)


if __name__ == "__main__":
    main()

"""Compacts an `ncu --set full --csv --page raw` log (one row per launch, ~700 columns) to the columns the round notes quote:

    python scripts/extract_ncu_full.py <raw.csv> <out.csv> [label-prefix]
"""
import csv
import sys

KEEP = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]
src, dst = sys.argv[1], sys.argv[2]
with open(src) as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = list(csv.reader(lines))
hdr, units, rows = rd[0], rd[1], rd[2:]
idx = [hdr.index(k) for k in KEEP if k in hdr]
with open(dst, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in rows:
        w.writerow([r[i].split("(")[0].replace("usb::", "") if hdr[i] == "Kernel Name" else r[i] for i in idx])
print(f"{dst}: {len(rows)} launches, {len(idx)} columns")

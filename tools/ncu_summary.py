"""Condense an `ncu -i X.ncu-rep --page raw --csv` dump into one row per kernel launch with the metrics DESIGN.md quotes.
usage: python tools/ncu_summary.py raw.csv > profiles/summary.csv"""
import csv, sys
COLS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__occupancy_limit_registers",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_membar", "smsp__pcsamp_warps_issue_stalled_selected"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
w = csv.writer(sys.stdout)
cols = [c for c in COLS if c in hdr]
w.writerow(["Kernel Name"] + cols)
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    name = name.split("(")[0] if not name.startswith("void") else name.split("(")[0] + (">" if "<" in name and ">" not in name.split("(")[0] else "")
    out = [name]
    for c in cols:
        i = hdr.index(c)
        out.append((r[i] + " " + units[i]).strip())
    w.writerow(out)

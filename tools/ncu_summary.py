"""Summarise an .ncu-rep (ncu --set full) per kernel: the metrics DESIGN.md / bench.py quote.  python tools/ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("== " + r[col["Kernel Name"]][:70])
    for k in KEYS:
        if k in col: print(f"   {k:85s} {r[col[k]]:>16s} {units[col[k]]}")
    stalls = [(float(r[i].replace(",", "")), h) for h, i in col.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
    for v, h in sorted(stalls, reverse=True)[:6]:
        print(f"   {h:85s} {v:16.6f} inst")
    print()

#!/usr/bin/env python3
"""ncu `--page raw --csv` -> the handful of per-launch numbers the docs quote (one block per profiled launch)."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "smsp__warps_active.avg.per_cycle_active",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("=" * 100)
    print("kernel:", r[ix.get("Kernel Name", 4)][:160])
    for k in KEYS:
        if k in ix:
            print(f"  {k:85s} {r[ix[k]]:>16s} {units[ix[k]]}")
    for h, i in ix.items():
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            print(f"  {h:85s} {r[i]:>16s}")

#!/usr/bin/env python3
"""Print the key metrics of every kernel in an .ncu-rep (run here, no GPU needed): python tools/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_sector_hit_rate.pct", "L2hit%"),
    ("l1tex__t_sector_hit_rate.pct", "L1hit%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_dispatch"),
    ("smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "st_imc"),
]
PIPES = "--pipes" in sys.argv  # also list every per-pipe utilisation column the report has
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("==", r[idx["Kernel Name"]][:100], "grid", r[idx.get("launch__grid_size", 0)])
    parts = []
    for key, short in WANT:
        if key in idx:
            v = r[idx[key]]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            parts.append(f"{short}={v}{units[idx[key]] if short in ('time', 'dram_rd', 'dram_wr') else ''}")
    print("   " + "  ".join(parts))
    if PIPES:
        pp = []
        for h, i in idx.items():
            if ("pipe_" in h) and ("pct" in h) and r[i] not in ("", "0", "n/a"):
                try:
                    if float(r[i]) >= 0.5:
                        pp.append(f"{h.replace('.avg.pct_of_peak_sustained_active', '').replace('sm__', '')}={float(r[i]):.3g}")
                except ValueError:
                    pass
        print("   pipes: " + "  ".join(pp))

#!/usr/bin/env python3
"""Condenses an `ncu --page raw --csv` export into the files kept under profiles/:

  python tools/ncu_summary.py gpurun_out/r02_ncu_raw.csv r02

writes profiles/<tag>_ncu_kernels_summary.csv (one line per launch, the columns the DESIGN/VERDICT discussion uses), and per-kernel
extracts in the raw-page format (header, units, rows) that bench.py reads its `roofline.traffic` from:
profiles/<tag>_ncu_k_step_full.csv, <tag>_ncu_gram_5a_terms3_full.csv, <tag>_ncu_gram_5a_terms1_full.csv, <tag>_ncu_gram_5b_block_full.csv."""
import csv
import os
import sys

src, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(src)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
KEEP = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.per_cycle_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]
keep = [k for k in KEEP if k in col]


def write(path, sel):
    with open(os.path.join(ROOT, "profiles", path), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(keep)
        w.writerow([units[col[k]] for k in keep])
        for r in sel:
            w.writerow([r[col[k]] for k in keep])
    print(path, len(sel), "launches")


name = col["Kernel Name"]
write("%s_ncu_kernels_summary.csv" % tag, data)
write("%s_ncu_k_step_full.csv" % tag, [r for r in data if "k_step<1, 1, 0>" in r[name]])
g2 = [r for r in data if "k_gram2<64, 3>" in r[name]]
write("%s_ncu_gram_5a_terms3_full.csv" % tag, g2[:1])
write("%s_ncu_gram_5b_block_full.csv" % tag, g2[1:])
write("%s_ncu_gram_5a_terms1_full.csv" % tag, [r for r in data if "k_gram<64, 1>" in r[name]])

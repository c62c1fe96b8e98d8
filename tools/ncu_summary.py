#!/usr/bin/env python3
"""Summarise an .ncu-rep (one row per captured launch) into the handful of numbers the roofline discussion needs."""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_sector_hit_rate.pct", "L2hit%"), ("l1tex__t_sector_hit_rate.pct", "L1hit%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__inst_executed.sum", "inst"), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_branch"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("lts__t_sectors_srcunit_tex_op_red.sum", "red_sectors")]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
seen = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0]
    seen[name] = seen.get(name, 0) + 1
    if seen[name] > int(sys.argv[2]) if len(sys.argv) > 2 else seen[name] > 1:
        continue
    print("==", name)
    line = []
    for key, short in WANT:
        if key in hdr:
            i = hdr.index(key)
            v = r[i]
            try:
                fv = float(v.replace(",", ""))
                v = "%.3g" % fv
            except ValueError:
                pass
            line.append("%s=%s%s" % (short, v, units[i] if units[i] in ("us", "ms", "Mbyte", "Gbyte", "Kbyte", "ns") else ""))
    print("   " + "  ".join(line))

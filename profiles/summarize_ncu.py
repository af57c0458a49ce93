#!/usr/bin/env python
"""Print the roofline-relevant counters of every kernel in an .ncu-rep (one block per launch).
usage: python profiles/summarize_ncu.py report.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_no_instructions",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_sample_buffer_full"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("== kernel:", name)
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print("   %-72s %-14s %s" % (w, units[i], r[i]))
    try:
        rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
        ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = rd*scale[ur] + wr*scale[uw]
        t = float(r[hdr.index("gpu__time_duration.sum")]); ut = units[hdr.index("gpu__time_duration.sum")]
        ts = t*{"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}[ut]
        print("   dram traffic per launch: %.4g bytes  (%.1f GB/s over the launch)" % (tot, tot/ts/1e9))
    except Exception as e:
        print("   (traffic n/a: %s)" % e)

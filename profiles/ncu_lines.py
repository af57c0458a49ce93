#!/usr/bin/env python
"""Aggregate warp-stall samples of an .ncu-rep by CUDA source line.
usage: python profiles/ncu_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None
agg = defaultdict(lambda: [0, 0, 0, 0, ""])   # samples, long_sb, no_inst, instr
cur_line = None; cur_src = ""
tot = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        iS = hdr.index("# Samples"); iL = hdr.index("stall_long_sb"); iN = hdr.index("stall_no_inst"); iI = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= iL:
        continue
    if r[0].strip():          # a CUDA source line row
        cur_line = (fname, r[0]); cur_src = r[1].strip()
        continue
    try:
        n = int(r[iS]); l = int(r[iL]); ni = int(r[iN]); ie = int(float(r[iI] or 0))
    except ValueError:
        continue
    a = agg[cur_line]; a[0] += n; a[1] += l; a[2] += ni; a[3] += ie; a[4] = cur_src
    tot += n
print("total samples", tot)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% long_sb %5.1f%% noinst %4.1f%% instr %9d  %s:%s  %s" % (100.0*a[0]/tot, 100.0*a[1]/tot, 100.0*a[2]/tot, a[3], k[0], k[1], a[4][:90]))

#!/usr/bin/env python
"""Top SASS instructions of an .ncu-rep by warp-stall samples, with the stall-reason split.
usage: python profiles/ncu_sass.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; recs = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    recs.append(r)
iS = hdr.index("# Samples"); iSrc = hdr.index("Source")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS]) for r in recs)
print("total samples", tot, " instructions", len(recs))
agg = {h: sum(int(r[hdr.index(h)]) for r in recs) for h in reasons}
print("by reason:", ", ".join("%s %.1f%%" % (h[6:], 100.0*v/tot) for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
# opcode histogram of executed instructions
iE = hdr.index("Instructions Executed")
ops = {}
for r in recs:
    op = r[iSrc].split()[0] if not r[iSrc].strip().startswith("@") else r[iSrc].split()[1]
    op = op.split(".")[0]
    ops[op] = ops.get(op, 0) + int(r[iE])
te = sum(ops.values())
print("executed warp-instructions by opcode:", ", ".join("%s %.1f%%" % (k, 100.0*v/te) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:18]))
for idx in sorted(range(len(recs)), key=lambda i: -int(recs[i][iS]))[:top]:
    r = recs[idx]
    rs = sorted(((int(r[hdr.index(h)]), h[6:]) for h in reasons), reverse=True)[:3]
    print("%5.2f%%  %-60s  %s   | prev: %s" % (100.0*int(r[iS])/tot, r[iSrc].strip()[:60], " ".join("%s=%d" % (n, v) for v, n in rs if v),
          recs[idx-1][iSrc].strip()[:40] if idx else ""))

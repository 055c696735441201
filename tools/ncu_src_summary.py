#!/usr/bin/env python3
"""Summarise an ncu report's source page: stall samples and executed instructions per source file and the hottest lines.
usage: tools/ncu_src_summary.py report.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None
hdr = None
agg = collections.defaultdict(lambda: [0, 0])
lines = []
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and cur and r[0] not in ("", "..."):
        try:
            si = hdr.index("# Samples")
            ii = hdr.index("Instructions Executed")
            s, ie = int(r[si]), int(r[ii])
        except Exception:
            continue
        agg[cur][0] += s
        agg[cur][1] += ie
        lines.append((s, ie, cur, r[0], r[1].strip()[:100]))
tot = max(1, sum(v[0] for v in agg.values()))
toti = max(1, sum(v[1] for v in agg.values()))
print("%-22s %10s %6s %14s %6s" % ("file", "samples", "%", "warp-inst", "%"))
for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("%-22s %10d %5.1f%% %14d %5.1f%%" % (k, v[0], 100.0 * v[0] / tot, v[1], 100.0 * v[1] / toti))
print("\nhottest lines (samples, warp-inst, file:line, source)")
for s, ie, f, ln, src in sorted(lines, reverse=True)[:top]:
    print("%8d %12d %s:%s  %s" % (s, ie, f, ln, src))

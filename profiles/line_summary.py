#!/usr/bin/env python3
"""Per-source-line view of an `ncu --page source --csv --print-source cuda,sass` export: stall samples (reliable: taken in an
un-instrumented pass), executed instructions and shared-memory wavefronts (taken in instrumented replay passes, which slow
the kernel down and therefore INFLATE every timing-dependent loop - here the drain warps' polling).
usage: line_summary.py export.csv [hw_inst_executed]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
sections, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = {"file": r[1], "rows": []}; sections.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
data, H = [], None
for sec in sections:
    for r in sec["rows"]:
        if r and r[0] == "Line No":
            H = r; ix = {h: i for i, h in enumerate(H)}; continue
        if H and len(r) >= len(H) and r[2] == "-":
            try:
                data.append((sec["file"].split("/")[-1], int(r[0]), r[1].strip(), float(r[6]), float(r[7]),
                             float(r[ix["L1 Wavefronts Shared"]]), float(r[ix["L1 Wavefronts Shared Ideal"]])))
            except ValueError:
                pass
ts = sum(d[3] for d in data); ti = sum(d[4] for d in data); tw = sum(d[5] for d in data)
print("lines %d  samples %d  instrumented warp instructions %.4g  instrumented shared wavefronts %.4g" % (len(data), ts, ti, tw))
if len(sys.argv) > 2:
    print("hardware counter smsp__inst_executed.sum %.4g -> %.0f %% of the instrumented count; the difference is extra polling"
          % (float(sys.argv[2]), 100 * float(sys.argv[2]) / ti))
print("-- top 40 lines by stall samples: samples | instrumented instructions | instrumented wavefronts (ideal)")
for d in sorted(data, key=lambda d: -d[3])[:40]:
    print("%-18s %5d  %5.2f%%  %5.2f%%  %5.2f%% (%5.2f%%)  %s" % (d[0][:18], d[1], 100 * d[3] / ts, 100 * d[4] / ti, 100 * d[5] / tw, 100 * d[6] / tw, d[2][:100]))
print("-- top 12 lines by shared-memory wavefronts")
for d in sorted(data, key=lambda d: -d[5])[:12]:
    print("%-18s %5d  wavefronts %5.2f%% (ideal %5.2f%%)  %s" % (d[0][:18], d[1], 100 * d[5] / tw, 100 * d[6] / tw, d[2][:100]))

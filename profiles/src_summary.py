#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv --print-source sass` export: top instructions by stall
samples, executed-instruction mix, shared-memory wavefronts.  usage: src_summary.py file.csv [kernel#]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
# split per kernel
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
b = blocks[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
H = b["rows"][0]; R = [r for r in b["rows"][1:] if len(r) == len(H)]
ix = {h: i for i, h in enumerate(H)}
def f(r, k):
    try: return float(r[ix[k]])
    except ValueError: return 0.0
tot_s = sum(f(r, "# Samples") for r in R); tot_i = sum(f(r, "Instructions Executed") for r in R)
tot_w = sum(f(r, "L1 Wavefronts Shared") for r in R)
print(b["name"]); print("instructions(warp) %.4g  samples %d  shared wavefronts %.4g" % (tot_i, tot_s, tot_w))
mix = collections.Counter(); smp = collections.Counter()
for r in R:
    op = r[ix["Source"]].split()
    op = [o for o in op if not o.startswith("@")][0].split(".")[0] if op else "?"
    mix[op] += f(r, "Instructions Executed"); smp[op] += f(r, "# Samples")
print("-- opcode mix (share of executed warp instructions | share of samples)")
for op, v in mix.most_common(22):
    print("  %-10s %6.2f%%  %6.2f%%" % (op, 100 * v / tot_i, 100 * smp[op] / tot_s))
stall_cols = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
print("-- stall reasons (all samples)")
agg = {h: sum(f(r, h) for r in R) for h in stall_cols}
for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
    print("  %-24s %6.2f%%" % (h, 100 * v / tot_s))
print("-- top 40 instructions by samples")
for r in sorted(R, key=lambda r: -f(r, "# Samples"))[:40]:
    top = max(stall_cols, key=lambda h: f(r, h))
    print("  %5.2f%%  exec %10.4g  wf %10.4g  %-18s %s" % (100 * f(r, "# Samples") / tot_s, f(r, "Instructions Executed"),
          f(r, "L1 Wavefronts Shared"), top, r[ix["Source"]].strip()[:90]))
print("-- top 12 by shared wavefronts")
for r in sorted(R, key=lambda r: -f(r, "L1 Wavefronts Shared"))[:12]:
    print("  wf %10.4g ideal %10.4g exec %10.4g  %s" % (f(r, "L1 Wavefronts Shared"), f(r, "L1 Wavefronts Shared Ideal"), f(r, "Instructions Executed"), r[ix["Source"]].strip()[:90]))

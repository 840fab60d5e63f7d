#!/usr/bin/env python3
"""Split an `ncu --page source --csv --print-source sass` export of pair_kernel into the filter role
(before USETMAXREG.TRY_ALLOC) and the drain role (after it): instruction counts, sample shares, stall mix."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
idx = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
blk = rows[idx[0] + 1: idx[1] if len(idx) > 1 else None]
H = blk[0]; R = [r for r in blk[1:] if len(r) == len(H)]
ix = {h: i for i, h in enumerate(H)}
def f(r, k):
    try: return float(r[ix[k]])
    except ValueError: return 0.0
tot = sum(f(r, '# Samples') for r in R); toti = sum(f(r, 'Instructions Executed') for r in R)
stall = [h for h in H if h.startswith('stall_') and 'Not Issued' not in h]
st = [n for n, r in enumerate(R) if 'USETMAXREG.TRY_ALLOC' in r[ix['Source']]][0]
spin = [n for n, r in enumerate(R) if 'SYNCS.PHASECHK' in r[ix['Source']] and f(r, 'Instructions Executed') > 50 * min([f(q, 'Instructions Executed') for q in R if 'SYNCS.PHASECHK' in q[ix['Source']]] + [1e30])]
print("total warp instructions %.4g, samples %d, static instructions %d" % (toti, tot, len(R)))
for name, seg in (("filter role + common", R[:st]), ("drain role + out-of-line blocks", R[st:])):
    ssum = sum(f(r, '# Samples') for r in seg); isum = sum(f(r, 'Instructions Executed') for r in seg)
    print("== %s: static %d, instr share %.3f (%.4g), sample share %.3f" % (name, len(seg), isum / toti, isum, ssum / tot))
    c = collections.Counter(); cs = collections.Counter()
    for r in seg:
        op = [o for o in r[ix['Source']].split() if not o.startswith('@')][0].split('.')[0]
        c[op] += f(r, 'Instructions Executed'); cs[op] += f(r, '# Samples')
    print("   opcode: share of role instructions | share of role samples")
    for k, v in c.most_common(14):
        print("     %-8s %5.1f%% %5.1f%%" % (k, 100 * v / isum, 100 * cs[k] / max(ssum, 1)))
    agg = {h: sum(f(r, h) for r in seg) for h in stall}
    print("   stalls: " + ", ".join("%s %.1f%%" % (h[6:], 100 * v / max(ssum, 1)) for h, v in sorted(agg.items(), key=lambda x: -x[1])[:7]))
    for r in sorted(seg, key=lambda r: -f(r, '# Samples'))[:8]:
        top = max(stall, key=lambda h: f(r, h))
        print("     %5.2f%% ex=%9.3g %-14s %s" % (100 * f(r, '# Samples') / tot, f(r, 'Instructions Executed'), top[6:], r[ix['Source']].strip()[:70]))

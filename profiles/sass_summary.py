#!/usr/bin/env python3
"""Opcode histogram and Blackwell-specific instruction evidence of the kernels in libgfn.so (sm_100a cubins):
    python profiles/sass_summary.py > profiles/r2/kernels.sass.txt
For every kernel of interest: registers / shared memory (cuobjdump -res-usage), the opcode histogram (mnemonic up to the
first '.'), and the lines that show TMA bulk copies (UBLKCP), mbarrier (SYNCS), register reallocation (USETMAXREG), packed
fp16 (HFMA2/HADD2/HMNMX2), warp match / redux, shared-memory atomics (ATOMS) and FP64 (DFMA/DADD/DMUL)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mdy-newanalysis-package_b200", "newanalysis", "libgfn.so")
WANT = [("pair_kernel<__half, 8, 8, true, 0, true, false>", r"pair_kernelI6__halfLi8ELi8ELb1ELi0ELb1ELb0EE"),
        ("count_kernel<true>", r"count_kernelILb1EE"),
        ("dense_kernel<__half, 12, 127>", r"dense_kernelI6__halfLi12ELi127EE"),
        ("dense_kernel<__half, 8, 31>", r"dense_kernelI6__halfLi8ELi31EE"),
        ("cells_list_kernel", r"cells_list_kernel"), ("prep_kernel", r"prep_kernel"), ("finalize_kernel", r"finalize_kernel")]
EVID = ["UBLKCP", "SYNCS", "USETMAXREG", "HFMA2", "HADD2", "HMNMX2", "MATCH", "REDUX", "ATOMS", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "FADD.RM"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
funcs = {}
cur = None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); funcs[cur] = []
        continue
    m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(.*?);", ln)
    if m and cur:
        funcs[cur].append(m.group(1).strip())
usage = {}
lines = res.splitlines()
for i, ln in enumerate(lines):
    m = re.search(r"Function (\S+):", ln)
    if m and i + 1 < len(lines):
        usage[m.group(1)] = lines[i + 1].strip()
print("# SASS summary of", os.path.relpath(LIB, ROOT), "(cuobjdump -sass, sm_100a)\n")
archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
print("cubin architectures:", ", ".join(archs), "\n")
for title, pat in WANT:
    hits = [f for f in funcs if re.search(pat, f)]
    if not hits:
        print("## %s: not found\n" % title)
        continue
    f = hits[0]
    ins = funcs[f]
    print("## %s\n   %s\n   %s\n   %d instructions" % (title, f, usage.get(f, ""), len(ins)))
    ops = collections.Counter()
    for t in ins:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        ops[t.split()[0].split(".")[0]] += 1
    print("   opcodes: " + ", ".join("%s %d" % kv for kv in ops.most_common(40)))
    for e in EVID:
        ex = [t for t in ins if re.search(r"(^|\s)" + re.escape(e), t)]
        if ex:
            print("   %-10s %4d   e.g. %s" % (e, len(ex), ex[0][:100]))
    print()

#!/usr/bin/env python3
"""Small golden cases through every kernel variant, for compute-sanitizer (profiles/sanitizer/run.sh).

    compute-sanitizer --tool memcheck|synccheck|initcheck|racecheck python profiles/sanitizer/run_cases.py [quick]

Each case is checked against its golden fixture as well (counts exact), so a sanitizer run that changes a result is seen.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "mdy-newanalysis-package_b200"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import cases as C                                   # noqa: E402
from newanalysis.gfunction import RDF, RDF_voronoi, calcHistoTS   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
names = ["rect_all", "self_odd", "hmin_pos", "tiny_1x5", "g000_nodip", "cfg2_shape"] if not quick else ["self_odd", "tiny_1x5"]
os.environ["GFN_CELLS_FORCE"] = "1"
variants = ["", "fp32", "fp64", "fp16,global", "per_particle", "fp16,dense", "fp16,nodense", "nodense,cells"] if not quick else ["", "fp64"]
done = 0
for name in names:
    case = dict(C.CASES[name], frames=1)
    z = np.load(os.path.join(GOLD, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    for flags in variants:
        n1, n2 = case["n1"], case["n2"]
        bx, by, bz = case["box"]
        try:
            r = RDF(case["modes"], case["hmin"], case["hmax"], case["dr"], n1, n2, n1, n2, bx, by, bz, gfn_flags=flags)
        except ValueError:
            continue                                # a flag this build does not know
        C.drive(r, case)
        h = r.raw_histogram()
        assert np.isfinite(h).all()
        if C.CASES[name]["frames"] == 1 and meta["mode_sel"] & 1:
            assert np.array_equal(h[:meta["histo_n"]], z["raw"][:meta["histo_n"]]), (name, flags)
        done += 1
if not quick:
    vc = dict(C.VORO_CASES["voro_rect"], frames=1)
    for flags in ("", "fp16,global"):
        C.drive_voro(C.make_voro(RDF_voronoi, vc, gfn_flags=flags), vc).raw_histogram()
        done += 1
    C.drive_ts(calcHistoTS, C.TS_CASES["ts_edge"])
    done += 1
libs = sorted(set(ln.split()[-1] for ln in open("/proc/self/maps") if "libgfn" in ln))
print("sanitizer cases done:", done, "with", ", ".join(os.path.relpath(p, ROOT) for p in libs))

#!/usr/bin/env python3
"""Throughput of the BASELINE.json configurations other than the bench workload (device-resident frames,
device-timed with gfn_mark): python profiles/run_configs.py [cfg1 cfg2 cfg3 cfg4 cfg5]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mdy-newanalysis-package_b200"))
import newanalysis.gfunction as G          # noqa: E402
from gfn_b200 import synth                 # noqa: E402


def run(name, n1, n2, L, modes, hmax, nframes, reps, self_pairs, flags="", env=None):
    os.environ.update(env or {})
    rng = np.random.default_rng(5)
    c1 = rng.random((nframes, n1, 3)) * L; d1 = rng.standard_normal((nframes, n1, 3))
    r = G.RDF(modes, 0.0, hmax, 0.05, n1, n2, n1, n2, L, gfn_flags=flags)
    dc1, dd1 = G.to_device(c1), G.to_device(d1)
    if self_pairs:
        dc2, dd2 = dc1, dd1
    else:
        c2 = rng.random((nframes, n2, 3)) * L; d2 = rng.standard_normal((nframes, n2, 3))
        dc2, dd2 = G.to_device(c2), G.to_device(d2)
    need_d = modes != ["000"]
    r.calcFramesDevice(nframes, dc1, dc2, dd1 if need_d else None, dd2 if need_d else None); r.sync()
    st0 = r.stats()
    r.mark(0)
    for _ in range(reps):
        r.calcFramesDevice(nframes, dc1, dc2, dd1 if need_d else None, dd2 if need_d else None)
    r.mark(1); r.sync()
    ms = r.mark_elapsed_ms(); st = r.stats()
    evals = st["pair_evals"] - st0["pair_evals"]
    f_in = (st["in_range"] - st0["in_range"]) / evals
    out = dict(config=name, n1=n1, n2=n2, modes=modes, frames=nframes * reps, gpair_per_s=evals / ms / 1e6,
               in_range_fraction=f_in, kernel_ms=st["kernel_ms"] - st0["kernel_ms"], wall_ms=ms, hist=st["hist_mode"],
               warps=st["warps_per_block"], replicas=st["hist_replicas"], batch_frames=st["batch_frames"], filter=st["filter"],
               kernel=st["kernel"], flags=flags, survivor_fraction=(st["survivors"] - st0["survivors"]) / evals,
               exact_fraction=(st["exact_evals"] - st0["exact_evals"]) / evals,
               tile_visit_fraction=((st["tile_visits"] - st0["tile_visits"]) / max(st["tile_visits_dense"] - st0["tile_visits_dense"], 1.0)) if st["cells"] else 1.0)
    for k in (env or {}):
        os.environ.pop(k, None)
    print(json.dumps(out), flush=True)


CONFIGS = {
    "cfg1": lambda: run("cfg1 SPC/E 1000 O-O g000", 1000, 1000, 31.04, ["000"], 15.0, 1398, 3, True),
    "cfg1_small_batches": lambda: run("cfg1, 100 frames per call", 1000, 1000, 31.04, ["000"], 15.0, 100, 20, True),
    "cfg2": lambda: run("cfg2 [C2mim][OTf] 1000x1000 5 modes 600 bins", 1000, 1000, 67.7, ["000", "110", "101", "011", "220"], 30.0, 699, 3, False),
    "cfg3_g000_count": lambda: run("cfg3 shape, g000 only (count kernel forced)", 32768, 32768, 99.33, ["000"], 15.0, 15, 2, True, "dense"),
    "cfg1_ring": lambda: run("cfg1 SPC/E 1000 O-O g000 (ring kernel)", 1000, 1000, 31.04, ["000"], 15.0, 100, 20, True, "nodense"),
    "cfg2_ring": lambda: run("cfg2 (ring kernel)", 1000, 1000, 67.7, ["000", "110", "101", "011", "220"], 30.0, 500, 4, False, "nodense"),
    "cfg1_all": lambda: run("cfg1 shape, all modes", 1000, 1000, 31.04, ["all"], 15.0, 100, 20, True),
    "cfg1_all_ring": lambda: run("cfg1 shape, all modes (ring kernel)", 1000, 1000, 31.04, ["all"], 15.0, 100, 20, True, "nodense"),
    "cfg3_dense": lambda: run("cfg3 water 32768 all modes (dense kernel forced)", 32768, 32768, 99.33, ["all"], 15.0, 15, 2, True, "dense"),
    "cfg3": lambda: run("cfg3 water 32768 all modes", 32768, 32768, 99.33, ["all"], 15.0, 15, 4, True),
    "cfg3_cells_ring": lambda: run("cfg3, tile skipping on pair_kernel", 32768, 32768, 99.33, ["all"], 15.0, 15, 4, True, "nodense,cells", {"GFN_CELLS_FORCE": "1"}),
    "cfg5_cells_ring": lambda: run("cfg5, tile skipping on pair_kernel", 1048576, 1048576, 315.4, ["all"], 15.0, 1, 2, True, "nodense,cells"),
    "cfg4a_cells": lambda: run("cfg4 solute COM 1024 x solvent 81920 all, tile skipping", 1024, 81920, 143.6, ["all"], 15.0, 16, 4, False, "cells", {"GFN_CELLS_FORCE": "1"}),
    "cfg5_g000": lambda: run("cfg5 shape, g000 only", 1048576, 1048576, 315.4, ["000"], 15.0, 1, 1, True),
    "cfg5_g000_cells": lambda: run("cfg5 shape, g000 only, tile skipping", 1048576, 1048576, 315.4, ["000"], 15.0, 1, 2, True, "cells"),
    "cfg3_g000_cells": lambda: run("cfg3 shape, g000 only, tile skipping", 32768, 32768, 99.33, ["000"], 15.0, 15, 4, True, "cells", {"GFN_CELLS_FORCE": "1"}),
    "cfg3_cells": lambda: run("cfg3 water 32768 all modes, tile skipping", 32768, 32768, 99.33, ["all"], 15.0, 15, 4, True, "cells", {"GFN_CELLS_FORCE": "1"}),
    "cfg5_cells": lambda: run("cfg5 1M sites self all, tile skipping", 1048576, 1048576, 315.4, ["all"], 15.0, 1, 2, True, "cells"),
    "cfg4b_cells": lambda: run("cfg4 solvent 81920 self all, tile skipping", 81920, 81920, 143.6, ["all"], 15.0, 2, 3, True, "cells", {"GFN_CELLS_FORCE": "1"}),
    "cfg3_g000": lambda: run("cfg3 shape, g000 only", 32768, 32768, 99.33, ["000"], 15.0, 15, 4, True),
    "cfg4a": lambda: run("cfg4 solute COM 1024 x solvent 81920 all", 1024, 81920, 143.6, ["all"], 15.0, 16, 4, False),
    "cfg4b": lambda: run("cfg4 solvent 81920 self all", 81920, 81920, 143.6, ["all"], 15.0, 2, 3, True),
    "cfg4c": lambda: run("cfg4 solute atoms 4096 x solvent O 81920 g000", 4096, 81920, 143.6, ["000"], 15.0, 8, 4, False),
    "cfg5": lambda: run("cfg5 1M sites self all", 1048576, 1048576, 315.4, ["all"], 15.0, 1, 1, True),
}

if __name__ == "__main__":
    for k in (sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg3_g000", "cfg4a", "cfg4b", "cfg4c"]):
        t0 = time.time()
        CONFIGS[k]()

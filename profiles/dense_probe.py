#!/usr/bin/env python3
"""Probe of the dense kernel on a cfg1-shaped box: python profiles/dense_probe.py <frames per call> <calls> [modes]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mdy-newanalysis-package_b200"))
import newanalysis.gfunction as G
nf, calls = int(sys.argv[1]), int(sys.argv[2])
modes = sys.argv[3].split("+") if len(sys.argv) > 3 else ["all"]
n, L = 1000, 31.04
rng = np.random.default_rng(5)
c = rng.random((nf, n, 3)) * L; d = rng.standard_normal((nf, n, 3))
r = G.RDF(modes, 0.0, 15.0, 0.05, n, n, n, n, L)
dc, dd = G.to_device(c), G.to_device(d)
t0 = time.time()
for k in range(calls):
    r.calcFramesDevice(nf, dc, dc, dd, dd)
    r.sync()
    print("call", k, "done", round(time.time() - t0, 3), "s", flush=True)
st = r.stats()
print(st["kernel"], st["warps_per_block"], st["hist_replicas"], st["batch_frames"], "in_range", st["in_range"], "kernel_ms", st["kernel_ms"], flush=True)

#!/usr/bin/env python3
"""bench.py - pair evaluations per second of the multipole g-function path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg4|cfg5]
    python bench.py --gpus N --single-process          (one process drives N devices; not the driver's launch)

Workload (config.workload), default cfg3 = BASELINE.json configs[2] - water, 32,768 molecules, self g-function, all
seven multipole modes, histogram 0-15 A / 0.05 A (300 bins), cubic box L = 99.33 A, synthetic uniform coordinates and
normal dipoles (gfn_b200/synth.py).  One STEP = one pass of the hot path over a batch of frames_per_step frames; one
pair evaluation = one iteration of the reference j loop (gfunction.pyx:130), n(n+1)/2 per frame.  cfg5 = configs[4]
(1,048,576 sites, one frame per step), cfg4 = configs[3] (four RDF objects per frame on one shared upload).

  value      device-timed (CUDA events on the library's compute stream, gfn_mark) throughput with the frames already
             resident in HBM; N ranks each process their own frames (weak scaling).
  e2e        the same metric through the C ABI with HOST buffers: RDF.calcFrames(pinned=True) -> gfn_enqueue_frames_pinned
             (strided DMA out of page-locked host arrays, no staging copy), kernels, and a histogram readout (D2H, plus
             the NCCL all-reduce over ranks when N > 1) every step; wall clock between two device syncs, max over ranks.
  e2e_calcframe  the reference's own call, RDF.calcFrame once per frame with pageable numpy arrays (staged through the
             pinned ring by the calling thread), same readout per step.
  strong     a FIXED number of frames (1,024 for cfg3) split over the N ranks from page-locked host arrays, wall clock
             from a barrier to the end of the final all-reduced readout, max over ranks.
  parity     (every N) exact check of the timed run (sum of g000 == 2 x in-range pairs summed over ranks) and a small
             seeded multi-rank case recomputed on rank 0 by the CPU oracle (checker only, outside every timed region).
  roofline   FP64 vector pipe: algorithmic flops/eval x evals / pair-kernel time (CUDA events around
             every pair-kernel launch, summed by the library) against the DFMA peak measured in this run.
  cpu_baseline  the UNMODIFIED reference module (oracle/_ref, kind "reference"; else the C port) timed on this box's
             host cores on a bounded sample of the same workload.
  other_configs  (N = 1) BASELINE configs[0], [1] and [4], device-resident, with flops/eval and roofline fraction.

--impl reference times the reference's own CPU implementation (oracle/_ref) on the same config with all host cores;
under torchrun rank 0 alone runs it (N GPUs are then compared with ONE host, not with N hosts).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

# stdout carries exactly one JSON line: NCCL's version banner (NCCL_DEBUG=VERSION, printed to stdout) goes to stderr instead
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "mdy-newanalysis-package_b200")
for p in (PKG, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

from gfn_b200 import sharding, synth  # noqa: E402

METRIC = "pair evals/sec (1e9), multipole g-function"
UNIT = "1e9 pair evals/s"
NOMINAL_FP64_TFLOPS = 37.2
HN = 300

# frames per step / frames in the host pool (the pool exceeds L2: 126 MB) / frames of the strong-scaling run
SHAPES = {"cfg1": (1398, 1398, 0), "cfg2": (699, 699, 0), "cfg3": (64, 128, 1024), "cfg5": (1, 2, 8)}


def cfg(name):
    """Workload description: synth.CONFIGS entry + frames_per_step / pool_frames / strong_frames."""
    c = dict(synth.CONFIGS[name])
    n = int(os.environ.get("GFN_BENCH_N", c["n1"]))
    if n != c["n1"]:                     # reduced-size dry runs only (never the reported number)
        c["L"] = c["L"] * (n / c["n1"]) ** (1.0 / 3.0)
        c["n1"] = c["n2"] = n
    c["name"] = name
    c["F"], c["pool"], c["strong"] = SHAPES[name]
    return c


def public_config(name, c):
    """The `config` object both arms print (identical keys and values)."""
    return {"workload": name, "n": c["n1"], "L": c["L"], "modes": "+".join(c["modes"]), "bins": HN,
            "l2": "inputs exceed L2: a pool of frames (> 126 MB) is cycled, site records are re-derived every step"}


def make_pool(c, nframes, rank, alloc=None):
    """(coordinates, dipoles) of selection 1 [and 2] for nframes frames; alloc(shape) -> array (page-locked or plain)."""
    alloc = alloc or (lambda shape: np.empty(shape))
    n1, n2 = c["n1"], c["n2"]
    co = alloc((nframes, n1, 3)); di = alloc((nframes, n1, 3))
    for f in range(nframes):
        co[f], di[f] = synth.frame(c["cfg"], f, n1, c["L"], stream=1000 * rank)
    if c["self_pairs"]:
        return co, di, co, di
    co2 = alloc((nframes, n2, 3)); di2 = alloc((nframes, n2, 3))
    for f in range(nframes):
        co2[f], di2[f] = synth.frame(c["cfg"], f, n2, c["L"], stream=1000 * rank + 1)
    return co, di, co2, di2


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons, sampled every 50 ms by one nvidia-smi process per rank that is started early
    (its start-up takes longer than a short timed region, above all on an 8-GPU box); window(t0, t1) reduces the samples
    whose time stamps fall into a timed region (wall-clock seconds, time.time())."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="gfn_clocks_", suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=fd, stderr=subprocess.DEVNULL)
            os.close(fd)
        except OSError:
            self.proc = None

    def window(self, t0, t1):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.12)           # let the sample that covers the end of the region land in the file
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        try:
            for ln in open(self.path):
                parts = [x.strip() for x in ln.split(",")]
                if len(parts) < 8:
                    continue
                try:
                    ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, float(parts[1]), float(parts[2]), [nm for nm, v in zip(names, parts[4:8]) if v.lower().startswith("active")]))
                except ValueError:
                    continue
        except OSError:
            return out
        inside = [r for r in rows if t0 <= r[0] <= t1 + 0.05]
        if not inside and rows:     # region shorter than the sampling period: the sample closest to it
            mid = 0.5 * (t0 + t1)
            near = min(rows, key=lambda r: abs(r[0] - mid))
            if abs(near[0] - mid) < 0.5:
                inside = [near]
                out["note"] = "nearest sample, %.0f ms from the region" % (1e3 * abs(near[0] - mid))
        if inside:
            out.update(sm_mhz=float(np.median([r[1] for r in inside])), sm_max_mhz=float(max(r[2] for r in inside)), samples=len(inside))
            out["reasons"] = sorted(set(nm for r in inside for nm in r[3]))
        return out

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        try:
            os.unlink(self.path)
        except OSError:
            pass


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the reference's prange must get every host core."""
    n = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


def reference_class():
    """(RDF-like class, kind): the unmodified reference class if oracle/_ref is present, else the port."""
    import oracle as O
    ref = O.load_ref()
    return (ref.RDF, "reference") if ref is not None else (O.OracleRDF, "port")


def time_reference(name, c, steps, warmup, rank=0):
    """The reference's own RDF.calcFrame on the host cores.  cfg1-cfg3: one full frame per step.  cfg5: a frame is 5.5e11
    pair evaluations (minutes on a CPU) - a step is a 1,024-row block of it against all sites, through the reference's
    rectangular path.  cfg4: the four objects of a frame, the solvent self object as a 2,048-row block."""
    use_all_host_threads()
    RDFc, kind = reference_class()
    hm = (c["hmin"], c["hmax"], c["dr"])
    jobs = []                                                  # (rdf, per-step pair evals, frame maker)
    if name == "cfg4":
        q = synth.CFG4
        ns, nu, na = q["n_solvent"], q["n_solute"], q["n_solute_atoms"]
        blk = 2048

        def frames4(f):
            return cfg4_frame(f, 0)
        jobs = [(RDFc(["all"], *hm, nu, ns, nu, ns, q["L"]), nu * ns, lambda a: (a["solu_c"], a["solv_c"], a["solu_d"], a["solv_d"])),
                (RDFc(["all"], *hm, blk, ns, blk, ns, q["L"]), blk * ns, lambda a: (a["solv_c"][:blk], a["solv_c"], a["solv_d"][:blk], a["solv_d"])),
                (RDFc(["000"], *hm, na, ns, nu, ns, q["L"]), na * ns, lambda a: (a["solu_a"], a["solv_o"])),
                (RDFc(["all"], *hm, nu, nu, nu, nu, q["L"]), synth.pair_evals(nu, nu), lambda a: (a["solu_c"], a["solu_c"], a["solu_d"], a["solu_d"]))]
        sample = "per step: the four RDF objects of one cfg4 frame, the solvent self object as a %d-row block (%d pair evals)" % (blk, sum(j[1] for j in jobs))
        frames = [frames4(f) for f in range(2)]
    else:
        n1, n2 = c["n1"], c["n2"]
        if name == "cfg5":
            blk = 1024
            rdf = RDFc(c["modes"], *hm, blk, n2, blk, n2, c["L"])
            per = blk * n2
            mk = lambda a: (np.ascontiguousarray(a[0][:blk]), a[0], np.ascontiguousarray(a[1][:blk]), a[1])      # noqa: E731
            sample = "per step: a %d-row block of one cfg5 frame against all %d sites (%d pair evals)" % (blk, n2, per)
        else:
            rdf = RDFc(c["modes"], *hm, n1, n2, n1, n2, c["L"])
            per = synth.pair_evals(n1, n2)
            mk = lambda a: (a[0], a[2], a[1], a[3])                                                            # noqa: E731
            sample = "per step: one full frame of %s (n=%d, %d pair evals)" % (name, n1, per)
        jobs = [(rdf, per, mk)]
        nfr = min(3, warmup + steps)
        pool = make_pool(c, nfr, rank)
        frames = [(pool[0][f], pool[1][f], pool[2][f], pool[3][f]) for f in range(nfr)]
    per_step = sum(j[1] for j in jobs)

    def step(k):
        fr = frames[k % len(frames)]
        for rdf, _, mk in jobs:
            rdf.calcFrame(*mk(fr))

    for k in range(warmup):
        step(k)
    t0 = time.perf_counter()
    for k in range(steps):
        step(warmup + k)
    dt = time.perf_counter() - t0
    return per_step * steps / dt / 1e9, dt, kind, sample + ", %d step(s), %d warm-up" % (steps, warmup)


def run_reference(args, rank, world):
    if rank != 0:
        return
    name = args.workload
    c = cfg(name) if name != "cfg4" else cfg4_config()
    val, dt, kind, sample = time_reference(name, c, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": public_config(name, c),
        "note": "reference Cython/OpenMP calcHisto on this box's host cores (rank 0 only: under torchrun N GPUs are compared with ONE host)",
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": host_threads(), "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# cfg4: BASELINE configs[3] - mixed solute / solvent box, four RDF objects per frame on one shared upload
# ---------------------------------------------------------------------------------------------------------------
def cfg4_config():
    q = dict(synth.CFG4)
    q.update(n1=q["n_solvent"], n2=q["n_solvent"], modes=["all"], name="cfg4")
    return q


def cfg4_frame(f, rank):
    q = synth.CFG4
    ns, nu, na = q["n_solvent"], q["n_solute"], q["n_solute_atoms"]
    rng = np.random.default_rng(4000 + f + 7919 * rank)
    solv_c = np.ascontiguousarray(rng.random((ns, 3)) * q["L"]); solv_d = rng.standard_normal((ns, 3))
    solv_o = np.ascontiguousarray(solv_c + 0.06 * rng.standard_normal((ns, 3)))
    solu_c = np.ascontiguousarray(rng.random((nu, 3)) * q["L"]); solu_d = rng.standard_normal((nu, 3))
    solu_a = np.ascontiguousarray(np.repeat(solu_c, na // nu, axis=0) + 1.5 * rng.standard_normal((na, 3)))
    return dict(solv_c=solv_c, solv_d=solv_d, solv_o=solv_o, solu_c=solu_c, solu_d=solu_d, solu_a=solu_a)


def cfg4_pairs():
    q = synth.CFG4
    ns, nu, na = q["n_solvent"], q["n_solute"], q["n_solute_atoms"]
    return nu * ns + synth.pair_evals(ns, ns) + na * ns + synth.pair_evals(nu, nu)


def run_cfg4(args, rank, world, local_rank, dist, G, barrier, allmax, sampler):
    """Per frame: six arrays are uploaded ONCE (DeviceArray.upload) and shared by four RDF objects through the reference's
    pre_*_gpu keywords of calcFrame (gfunction.pyx:390-391, 431-453).  value: the frames' arrays already resident;
    e2e: uploads inside the timed region, a readout of all four histograms per step."""
    q = synth.CFG4
    ns, nu, na, L = q["n_solvent"], q["n_solute"], q["n_solute_atoms"], q["L"]
    hm = (q["hmin"], q["hmax"], q["dr"])
    F, K, W = args.frames_per_step or 4, args.steps, args.warmup
    per = cfg4_pairs()
    kw = dict(devices=[local_rank], gfn_flags=args.flags or None)
    rdfs = [G.RDF(["all"], *hm, nu, ns, nu, ns, L, **kw), G.RDF(["all"], *hm, ns, ns, ns, ns, L, **kw),
            G.RDF(["000"], *hm, na, ns, nu, ns, L, **kw), G.RDF(["all"], *hm, nu, nu, nu, nu, L, **kw)]
    if world > 1:
        for r in rdfs:
            r.attach_comm(sharding.broadcast_unique_id(dist, G.comm_unique_id, device="cuda"), world, rank)
    pool = [cfg4_frame(f, rank) for f in range(F)]
    keys = ("solv_c", "solv_d", "solv_o", "solu_c", "solu_d", "solu_a")
    dev = [{k: G.to_device(fr[k], local_rank) for k in keys} for fr in pool]

    def frame(fr, g):
        rdfs[0].calcFrame(fr["solu_c"], fr["solv_c"], fr["solu_d"], fr["solv_d"], pre_coor_sel1_gpu=g["solu_c"], pre_coor_sel2_gpu=g["solv_c"],
                          pre_dip_sel1_gpu=g["solu_d"], pre_dip_sel2_gpu=g["solv_d"])
        rdfs[1].calcFrame(fr["solv_c"], fr["solv_c"], fr["solv_d"], fr["solv_d"], pre_coor_sel1_gpu=g["solv_c"], pre_coor_sel2_gpu=g["solv_c"],
                          pre_dip_sel1_gpu=g["solv_d"], pre_dip_sel2_gpu=g["solv_d"])
        rdfs[2].calcFrame(fr["solu_a"], fr["solv_o"], pre_coor_sel1_gpu=g["solu_a"], pre_coor_sel2_gpu=g["solv_o"])
        rdfs[3].calcFrame(fr["solu_c"], fr["solu_c"], fr["solu_d"], fr["solu_d"], pre_coor_sel1_gpu=g["solu_c"], pre_coor_sel2_gpu=g["solu_c"],
                          pre_dip_sel1_gpu=g["solu_d"], pre_dip_sel2_gpu=g["solu_d"])

    def sync_all():
        for r in rdfs:
            r.sync()

    def dev_step(s):
        for f in range(F):
            frame(pool[f], dev[f])

    up = {k: G.to_device(pool[0][k], local_rank) for k in keys}          # the e2e arm's upload slots, reused every frame

    def host_step(s):
        for f in range(F):
            for k in keys:
                up[k].upload(pool[f][k])
            frame(pool[f], up)
        return [r.raw_histogram() for r in rdfs]

    for s in range(W):
        dev_step(s)
    sync_all()
    st0 = [r.stats() for r in rdfs]
    barrier()
    tw0 = time.time(); t0 = time.perf_counter()
    for s in range(K):
        dev_step(W + s)
    sync_all()
    ms = allmax((time.perf_counter() - t0) * 1e3)
    tw1 = time.time()
    barrier()
    clocks = sampler.window(tw0, tw1)
    st1 = [r.stats() for r in rdfs]
    for s in range(W):
        host_step(s)
    sync_all()
    barrier()
    t0 = time.perf_counter()
    for s in range(K):
        hists = host_step(W + s)
    sync_all()
    dt = allmax(time.perf_counter() - t0)
    barrier()
    sampler.stop()
    # exact count check of every object: sum of g000 over ranks == weight x in-range pairs over ranks
    inr = np.array([a["in_range"] - a["overflow"] for a in [r.stats() for r in rdfs]], dtype=np.float64)
    if dist is not None:
        inr = sharding.allreduce_histogram(dist, inr, device="cuda")
    wts = [1.0, 2.0, 1.0, 2.0]
    exact = all(float(h[:HN].sum()) == w * v for h, w, v in zip(hists, wts, inr))
    if rank != 0:
        return None
    kernel_ms = sum(b["kernel_ms"] - a["kernel_ms"] for a, b in zip(st0, st1))
    launches = sum(b["kernel_launches"] - a["kernel_launches"] for a, b in zip(st0, st1))
    h2d = F * sum(pool[0][k].nbytes for k in keys)
    return {
        "metric": METRIC, "value": world * K * F * per / (ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": public_config("cfg4", cfg4_config()),
        "run": {"frames_per_step": F, "pair_evals_per_step": F * per, "objects": ["solute COM x solvent COM, all modes", "solvent self, all modes",
                "solute atoms x solvent O, g000", "solute self, all modes"], "kernels": [b["kernel"] for b in st1],
                "timing": "wall clock between device syncs (four handles, four compute streams), max over ranks"},
        "roofline": {"bound": "fp64", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None,
                     "note": "four handles with different kernels; per-kernel rooflines are reported for cfg3 (default workload)",
                     "kernel_ms_sum_over_handles_per_step": kernel_ms / K},
        "e2e": {"value": world * K * F * per / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * (8 * HN * 8 + 64),
                "ms_per_step": dt / K * 1e3, "api": "DeviceArray.upload x 6 + RDF.calcFrame(pre_*_gpu=...) x 4 per frame, four readouts per step"},
        "parity": {"g000_exact": bool(exact), "check": "per object: sum of g000 (all-reduced) == weight x in-range pairs summed over ranks"},
        "gpu_launches": int(launches),
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"], "samples": clocks["samples"]},
    }


# ---------------------------------------------------------------------------------------------------------------
# parity leg (checker; never timed)
# ---------------------------------------------------------------------------------------------------------------
def parity_case(G, dist, rank, world, local_rank, flags):
    """A small seeded case of the bench workload's shape and density (n = 8,192 water sites, all modes: the same kernel
    variant as cfg3) - two frames per rank through RDF.calcFrame, all-reduced by the library, recomputed on rank 0 by the
    CPU oracle for all ranks' frames."""
    n = 8192
    L = 99.33 * (n / 32768.0) ** (1.0 / 3.0)
    r = G.RDF(["all"], 0.0, 15.0, 0.05, n, n, n, n, L, devices=[local_rank], gfn_flags=flags)
    if world > 1:
        r.attach_comm(sharding.broadcast_unique_id(dist, G.comm_unique_id, device="cuda"), world, rank)
    nf = 2
    for f in range(nf):
        c, d = synth.frame(9, f, n, L, stream=rank)
        r.calcFrame(c, c, d, d)
    got = r.raw_histogram()
    kern = r.stats()["kernel"]
    if rank != 0:
        return None
    import oracle as O
    use_all_host_threads()
    o = O.OracleRDF(["all"], 0.0, 15.0, 0.05, n, n, n, n, L)
    for rk in range(world):
        for f in range(nf):
            c, d = synth.frame(9, f, n, L, stream=rk)
            o.calcFrame(c, c, d, d)
    ref = o.raw_sum()
    scale = np.maximum(np.abs(ref), np.tile(ref[:HN], 7))
    return {"g000_exact": bool(np.array_equal(got[:HN], ref[:HN])), "max_weighted_rel": float(np.max(np.abs(got - ref) / np.maximum(scale, 1.0))),
            "tolerance": 1e-12, "case": "n=%d, all modes, %d frames on each of %d rank(s), kernel %s, vs the CPU oracle on rank 0" % (n, nf, world, kern),
            "g000_pairs": float(ref[:HN].sum())}


# ---------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import newanalysis.gfunction as G          # fails loudly if the CUDA extension is missing

    def barrier():
        if dist is not None:
            dist.barrier()

    def allmax(x):
        return x if dist is None else sharding.allreduce_max(dist, x, device="cuda")

    def allsum(v):
        return np.asarray(v, dtype=np.float64) if dist is None else sharding.allreduce_histogram(dist, np.asarray(v, dtype=np.float64), device="cuda")

    sampler = ClockSampler(local_rank)
    sampler.start()
    name = args.workload
    if name == "cfg4":
        line = run_cfg4(args, rank, world, local_rank, dist, G, barrier, allmax, sampler)
        if line is not None:
            print(json.dumps(line), flush=True)
        if dist is not None:
            dist.destroy_process_group()
        return

    c = cfg(name)
    n1, n2, L = c["n1"], c["n2"], c["L"]
    per = synth.pair_evals(n1, n2)
    F, K, W = args.frames_per_step or c["F"], args.steps, args.warmup
    POOL = max(c["pool"], F)
    msel = 127 if c["modes"] == ["all"] else None
    flags = ("" if args.filter == "auto" else args.filter) + ("," + args.flags if args.flags else "")
    mk = lambda: G.RDF(c["modes"], c["hmin"], c["hmax"], c["dr"], n1, n2, n1, n2, L, devices=[local_rank], gfn_flags=flags)  # noqa: E731
    new_uid = lambda: sharding.broadcast_unique_id(dist, G.comm_unique_id, device="cuda")      # noqa: E731

    pc1, pd1, pc2, pd2 = make_pool(c, POOL, rank, alloc=G.pinned_empty)       # page-locked host pool
    need_d = c["modes"] != ["000"]
    fp64_peak, fp32_peak = G.measure_peaks(local_rank)

    # ---------------- arm 1: inputs resident in HBM, device-timed ------------------------------------
    r = mk()
    if world > 1:
        r.attach_comm(new_uid(), world, rank)
    dc1, dd1 = G.to_device(pc1, local_rank), (G.to_device(pd1, local_rank) if need_d else None)
    dc2, dd2 = (dc1, dd1) if c["self_pairs"] else (G.to_device(pc2, local_rank), G.to_device(pd2, local_rank) if need_d else None)

    def first_of(s):
        return (s * F) % (POOL - F + 1) if POOL > F else 0

    def dev_step(s):
        r.calcFramesDevice(F, dc1, dc2, dd1, dd2, first=first_of(s))

    for s in range(W):
        dev_step(s)
    r.sync()
    st0 = r.stats()
    barrier()
    tw0 = time.time()
    r.mark(0)
    for s in range(K):
        dev_step(W + s)
    r.mark(1)
    r.sync()
    tw1 = time.time()
    ms = allmax(r.mark_elapsed_ms())
    barrier()
    clocks = sampler.window(tw0, tw1)
    st1 = r.stats()
    hist = r.raw_histogram()                     # includes the NCCL all-reduce over ranks when world > 1
    # exact: every accepted pair of every rank sits in the all-reduced g000 row with its weight (self: 2, i == j never counts)
    tot = allsum([st1["in_range"] - st1["overflow"]])
    wt = 2.0 if n1 == n2 else 1.0
    counts_exact = bool(float(hist[:HN].sum()) == wt * float(tot[0]))
    assert counts_exact, "sum of g000 over ranks (%r) != %g x in-range pairs over ranks (%r)" % (hist[:HN].sum(), wt, tot[0])
    value = world * K * F * per / (ms * 1e-3) / 1e9
    kernel_ms = st1["kernel_ms"] - st0["kernel_ms"]
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    pair_launches = st1["pair_launches"] - st0["pair_launches"]
    f_in = (st1["in_range"] - st0["in_range"]) / (st1["pair_evals"] - st0["pair_evals"])
    f_surv = (st1["survivors"] - st0["survivors"]) / (st1["pair_evals"] - st0["pair_evals"])
    flops_eval = synth.flops_per_eval(mode_mask(c["modes"]), f_in)
    achieved = flops_eval * K * F * per / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else None
    del r, dc1, dd1, dc2, dd2

    # ---------------- arm 2: end to end, page-locked host arrays through gfn_enqueue_frames_pinned --------------------
    e = mk()
    if world > 1:           # an NCCL unique id makes exactly one communicator: a second handle needs its own
        e.attach_comm(new_uid(), world, rank)

    def block(lo, hi):
        c1v = pc1[lo:hi]; d1v = pd1[lo:hi] if need_d else None
        if c["self_pairs"]:
            return c1v, c1v, d1v, d1v
        return c1v, pc2[lo:hi], d1v, (pd2[lo:hi] if need_d else None)

    def host_step(s):
        lo = first_of(s)
        e.calcFrames(*block(lo, lo + F), pinned=True)
        return e.raw_histogram()

    for s in range(W):
        host_step(s)
    e.sync()
    se0 = e.stats()
    barrier()
    tw0 = time.time()
    t0 = time.perf_counter()
    for s in range(K):
        host_step(W + s)
    e.sync()
    dt = allmax(time.perf_counter() - t0)
    tw1 = time.time()
    barrier()
    clocks_e2e = sampler.window(tw0, tw1)
    se1 = e.stats()

    # ---------------- strong scaling: a fixed number of frames over all ranks, final readout included ----------------
    strong = None
    if c["strong"]:
        T = c["strong"]
        mine = len(sharding.shard_frames(T, rank, world))
        e.raw_histogram()
        barrier()
        t0 = time.perf_counter()
        done = 0
        while done < mine:
            k = min(F, mine - done)
            lo = done % (POOL - k + 1)
            e.calcFrames(*block(lo, lo + k), pinned=True)
            done += k
        e.raw_histogram()
        ds = allmax(time.perf_counter() - t0)
        barrier()
        strong = {"frames_total": T, "frames_this_rank": mine, "wall_ms": ds * 1e3, "value": T * per / ds / 1e9, "unit": UNIT,
                  "includes": "H2D from page-locked host arrays, kernels, one all-reduced readout at the end; barrier to barrier, max over ranks"}

    # ---------------- the reference's own call: RDF.calcFrame per frame, pageable arrays ---------------------------
    e2e_cf = None
    if n1 <= 131072:
        g = mk()
        if world > 1:
            g.attach_comm(new_uid(), world, rank)
        qc1, qd1 = np.array(pc1), np.array(pd1)
        qc2, qd2 = (qc1, qd1) if c["self_pairs"] else (np.array(pc2), np.array(pd2))

        def cf_step(s):
            lo = first_of(s)
            for f in range(lo, lo + F):
                if need_d:
                    g.calcFrame(qc1[f], qc2[f], qd1[f], qd2[f])
                else:
                    g.calcFrame(qc1[f], qc2[f])
            return g.raw_histogram()

        for s in range(W):
            cf_step(s)
        g.sync()
        sg0 = g.stats()
        barrier()
        t0 = time.perf_counter()
        for s in range(K):
            cf_step(W + s)
        g.sync()
        dtc = allmax(time.perf_counter() - t0)
        barrier()
        sg1 = g.stats()
        e2e_cf = {"value": world * K * F * per / dtc / 1e9, "unit": UNIT, "h2d_bytes_per_step": (sg1["h2d_bytes"] - sg0["h2d_bytes"]) / K,
                  "d2h_bytes_per_step": 8 * HN * 8 + 8 * 8, "ms_per_step": dtc / K * 1e3,
                  "api": "newanalysis.gfunction.RDF.calcFrame once per frame (pageable numpy arrays, staged through the pinned ring) + readout per step"}
        del g, qc1, qd1, qc2, qd2

    # ---------------- parity leg (checker) -------------------------------------------------------------------------
    parity = parity_case(G, dist, rank, world, local_rank, flags)

    # ---------------- N = 1: atoms / DCD arms (cfg3), the other BASELINE configurations ----------------------------
    atoms = None
    others = None
    if world == 1 and name == "cfg3":
        atoms = atom_arms(args, G, c, mk, np.array(pc1), F, K, W, per, POOL)
    if world == 1 and name == "cfg3" and not args.no_other_configs:
        del e
        others = [other_config(G, nm, local_rank, fp64_peak) for nm in ("cfg1", "cfg2", "cfg5")]
        others.append(other_config(G, "cfg5", local_rank, fp64_peak, flags="cells"))
        others.append(other_config(G, "cfg4_solvent_self", local_rank, fp64_peak))      # the dominant object of cfg4 (the whole
        # four-object workload with its shared uploads runs under --workload cfg4)
    sampler.stop()
    e2e_value = world * K * F * per / dt / 1e9
    h2d = (se1["h2d_bytes"] - se0["h2d_bytes"]) / K
    d2h = 8 * HN * 8 + 8 * 8

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        val, cdt, kind, sample = time_reference(name, c, 2, 1)
        cpu = {"value": val, "unit": UNIT, "cores": host_threads(), "kind": kind, "sample": sample}

    parity = dict(parity or {})
    parity["timed_run_counts_exact"] = counts_exact
    parity["timed_run_check"] = "sum of the all-reduced g000 row == %g x (in-range - overflow) pairs summed over ranks, exactly" % wt
    traffic, traffic_src = ncu_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": public_config(name, c),
        "run": {"frames_per_step": F, "pair_evals_per_step": F * per, "pool_frames": POOL,
                "parallelism": "frames sharded over %d GPU(s), one NCCL all-reduce at readout" % world,
                "filter": st1["filter"], "kernel": st1["kernel"], "hist": st1["hist_mode"], "batch_frames": st1["batch_frames"], "grid": st1["grid"],
                "warps_per_block": st1["warps_per_block"], "smem_bytes": st1["smem_bytes"]},
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": (achieved / fp64_peak) if achieved else None, "traffic": traffic,
                     "traffic_unit": "bytes of DRAM read+write per pair_kernel launch (ncu --set full capture, %s); algorithmic: %d"
                                     % (traffic_src, st1["batch_frames"] * ((n1 + 1023) // 1024 * 1024) * 80),
                     "peak_source": "measured in this run: dependent-free DFMA loop (gfn_measure_peaks); nominal %.1f" % NOMINAL_FP64_TFLOPS,
                     "frac_of_nominal": (achieved / NOMINAL_FP64_TFLOPS) if achieved else None,
                     "flops_per_eval": flops_eval, "in_range_fraction": f_in, "survivor_fraction": f_surv,
                     "kernel": "pair_kernel" if st1["kernel"] == "ring" else st1["kernel"] + "_kernel", "kernel_ms_per_launch": kernel_ms / max(pair_launches, 1),
                     "kernel_share_of_step": kernel_ms / ms, "fp32_peak_tflops": fp32_peak,
                     "pipes_from_ncu": ncu_pipes(),
                     "lever": {"fp32": "fp32 conservative pre-filter + exact fp64 survivors",
                               "fp16x2": "packed-fp16 conservative pre-filter (two core sites per instruction) + exact fp64 survivors",
                               "fp64": "fp64 pre-test + exact fp64 survivors"}[st1["filter"]]},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": dt / K * 1e3,
                "api": "newanalysis.gfunction.RDF.calcFrames(pinned=True) -> gfn_enqueue_frames_pinned (page-locked host arrays, no staging copy) + readout per step"},
        "parity": parity,
        "gpu_launches": int(launches),
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                   "samples": clocks["samples"], "e2e": clocks_e2e},
    }
    if e2e_cf is not None:
        line["e2e_calcframe"] = e2e_cf
    if strong is not None:
        line["strong"] = strong
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if atoms is not None:
        line["e2e_atoms"] = atoms
    if world == 1 and name == "cfg3" and not args.no_other_configs:
        line["time_series"] = time_series_arm(args)
    if others is not None:
        line["other_configs"] = others
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def time_series_arm(args):
    """calcHistoTS (gfunction.pyx:168-192), one call per frame with host arrays, result[t] written on
    the host: 1,024 core sites x 32,768 surround sites, 240 bins.  The reference function on all host threads runs the same
    frames (fewer of them); the two results must be equal."""
    from newanalysis.gfunction import calcHistoTS
    import oracle as O
    n1, n2, nb, T = 1024, 32768, 240, 12
    L = (n2 / 0.0334) ** (1.0 / 3.0)
    rng = np.random.default_rng(77)
    surr = [rng.random((n2, 3)) * L for _ in range(T)]
    core = [np.ascontiguousarray(x[:n1]) + 0.37 for x in surr]
    res = np.zeros((T, n1, nb))
    calcHistoTS(core[0], surr[0], res, L, 0, 0.0, 12.0, 20.0)               # builds the handle
    res[:] = 0.0
    t0 = time.perf_counter()
    for t in range(T):
        calcHistoTS(core[t], surr[t], res, L, t, 0.0, 12.0, 20.0)
    ms = (time.perf_counter() - t0) / T * 1e3
    out = {"what": "calcHistoTS, %d x %d sites, %d bins, host arrays in, result[t] out, per call" % (n1, n2, nb),
           "ms_per_frame": ms, "value": n1 * n2 / ms / 1e6, "unit": UNIT, "d2h_bytes_per_frame": 8 * (n1 + 1) * nb}
    if not args.no_cpu_baseline:
        ref = O.load_ref()
        use_all_host_threads()
        fn = ref.calcHistoTS if ref is not None else None
        TR = 2
        rr = np.zeros((TR, n1, nb))
        t0 = time.perf_counter()
        for t in range(TR):
            if fn is not None:
                fn(core[t], surr[t], rr, L, t, 0.0, 12.0, 20.0)
            else:
                O.calc_histo_ts_c(core[t], surr[t], rr[t], L, 0.0, 12.0, 20.0)
        rms = (time.perf_counter() - t0) / TR * 1e3
        out["reference_ms_per_frame"] = rms
        out["reference_kind"] = "reference" if fn is not None else "port"
        out["reference_cores"] = host_threads() if fn is not None else 1
        out["equal_to_reference"] = bool(np.array_equal(res[:TR], rr))
    return out


def counts_ok(g000_sum, accepted, n1, n2, self_pairs, frames):
    """Exact relation between the g000 row and the number of accepted pairs: every accepted pair weighs 2 in a triangular
    frame and 1 otherwise - except i == j between two DIFFERENT selections of equal size (quirk Q1), which weighs 1: then
    the row lies between 2 * accepted - n * frames and 2 * accepted."""
    if n1 != n2:
        return bool(float(g000_sum) == accepted)
    if self_pairs:
        return bool(float(g000_sum) == 2.0 * accepted)
    return bool(2.0 * accepted - n1 * frames <= float(g000_sum) <= 2.0 * accepted)


def mode_mask(modes):
    bits = {"000": 1, "110": 2, "101": 4, "011": 8, "220": 16, "202": 32, "022": 64}
    if modes == ["all"]:
        return 127
    m = 0
    for k in modes:
        m |= bits[k]
    return m


def other_config(G, name, dev, fp64_peak, flags=None):
    """One of the other BASELINE configurations, frames resident in HBM, device-timed: value, flops/eval, roofline fraction.
    flags="cells": tile skipping - `value` then counts REFERENCE-EQUIVALENT evaluations (the pairs the reference's loop visits),
    `visited_value` the pairs of the tile pairs the kernel actually went through; no roofline fraction (the flop count of
    SURVEY 8d assumes every pair is evaluated)."""
    if name == "cfg4_solvent_self":
        q = synth.CFG4
        c = dict(cfg=4, n1=q["n_solvent"], n2=q["n_solvent"], L=q["L"], modes=["all"], hmin=q["hmin"], hmax=q["hmax"], dr=q["dr"], self_pairs=True, F=2)
    else:
        c = cfg(name)
    n1, n2, L = c["n1"], c["n2"], c["L"]
    F = c["F"]
    reps = 3 if name not in ("cfg5",) else 2
    per = synth.pair_evals(n1, n2)
    need_d = c["modes"] != ["000"]
    rng = np.random.default_rng(50 + c["cfg"])
    c1 = rng.random((F, n1, 3)) * L
    r = G.RDF(c["modes"], c["hmin"], c["hmax"], c["dr"], n1, n2, n1, n2, L, devices=[dev], gfn_flags=flags)
    dc1 = G.to_device(c1, dev); dd1 = G.to_device(rng.standard_normal((F, n1, 3)), dev) if need_d else None
    if c["self_pairs"]:
        dc2, dd2 = dc1, dd1
    else:
        dc2 = G.to_device(rng.random((F, n2, 3)) * L, dev); dd2 = G.to_device(rng.standard_normal((F, n2, 3)), dev) if need_d else None
    del c1
    r.calcFramesDevice(F, dc1, dc2, dd1, dd2); r.sync()
    s0 = r.stats()
    r.mark(0)
    for _ in range(reps):
        r.calcFramesDevice(F, dc1, dc2, dd1, dd2)
    r.mark(1); r.sync()
    ms = r.mark_elapsed_ms(); s1 = r.stats()
    evals = s1["pair_evals"] - s0["pair_evals"]
    f_in = (s1["in_range"] - s0["in_range"]) / evals
    fl = synth.flops_per_eval(mode_mask(c["modes"]), f_in)
    kms = s1["kernel_ms"] - s0["kernel_ms"]
    ach = fl * evals / (kms * 1e-3) / 1e12
    hist = r.raw_histogram()
    wt = 2.0 if n1 == n2 else 1.0
    hn = int(r.histo_n)
    out = {"workload": name, "n1": n1, "n2": n2, "modes": "+".join(c["modes"]), "frames": F * reps, "value": evals / ms / 1e6, "unit": UNIT,
           "kernel": s1["kernel"], "filter": s1["filter"], "in_range_fraction": f_in, "flops_per_eval": fl,
           "achieved_tflops": ach, "frac": ach / fp64_peak, "kernel_share": kms / ms, "ms_per_frame": ms / (F * reps)}
    if s1["cells"]:
        share = (s1["tile_visits"] - s0["tile_visits"]) / max(s1["tile_visits_dense"] - s0["tile_visits_dense"], 1.0)
        out.update(flags=flags, tile_skipping=True, visited_share=share, visited_value=share * evals / ms / 1e6,
                   value_counts="reference-equivalent pair evaluations (all pairs the reference's loop visits)", achieved_tflops=None, frac=None)
    return dict(out,
                bins=hn, counts_exact=counts_ok(hist[:hn].sum(), s1["in_range"] - s1["overflow"], n1, n2, c["self_pairs"], F * (reps + 1)))


def atom_arms(args, G, c, mk, pool_c, F, K, W, per, POOL):
    """N = 1, cfg3: end to end from float32 ATOM positions (3-site water; COM + dipoles on the device), from host arrays and
    from a DCD file on disk (page cache) through the reader threads."""
    n = c["n1"]
    rng = np.random.default_rng(4242)
    off = (0.6 * rng.standard_normal((n, 3, 3))).astype(np.float32)
    pool_a = np.ascontiguousarray((pool_c.astype(np.float32)[:, :, None, :] + off[None]).reshape(POOL, 3 * n, 3))
    topo = (np.tile([15.999, 1.008, 1.008], n), np.full(n, 3, dtype=np.int32), (3 * np.arange(n)).astype(np.int32), np.tile([-0.8476, 0.4238, 0.4238], n))
    first_of = lambda s: (s * F) % (POOL - F + 1) if POOL > F else 0       # noqa: E731

    def timed(obj, step):
        for s in range(W):
            step(s)
        obj.sync()
        s0 = obj.stats()
        t0 = time.perf_counter()
        for s in range(K):
            step(W + s)
        obj.sync()
        return time.perf_counter() - t0, s0, obj.stats()

    a = mk()
    a.attach_topology(*topo)

    def atom_step(s):
        lo = first_of(s)
        for f in range(lo, lo + F):
            a.calcFrameAtoms(pool_a[f])
        return a.raw_histogram()

    dta, sa0, sa1 = timed(a, atom_step)
    out = {"value": K * F * per / dta / 1e9, "unit": UNIT, "h2d_bytes_per_step": (sa1["h2d_bytes"] - sa0["h2d_bytes"]) / K,
           "d2h_bytes_per_step": 8 * HN * 8 + 8 * 8, "ms_per_step": dta / K * 1e3,
           "gpu_launches": int(sa1["kernel_launches"] - sa0["kernel_launches"]),
           "api": "newanalysis.gfunction.RDF.calcFrameAtoms (host float32 positions, 3 atoms per site; COM + dipoles on the device) + readout per step"}
    del a
    tmpd = tempfile.mkdtemp(prefix="gfn_bench_")
    path = os.path.join(tmpd, "pool.dcd")
    synth.write_dcd(path, pool_a)
    traj = G.DCDFile(path)
    sel = np.arange(3 * n, dtype=np.int32)
    t4 = mk()
    t4.attach_topology(*topo)

    def traj_step(s):
        lo = first_of(s)
        t4.calcTrajectory(traj, sel, first=lo, last=lo + F, threads=args.reader_threads)
        return t4.raw_histogram()

    dtt, s40, s41 = timed(t4, traj_step)
    traj.close()
    os.unlink(path); os.rmdir(tmpd)
    out["from_dcd"] = {"value": K * F * per / dtt / 1e9, "unit": UNIT, "ms_per_step": dtt / K * 1e3,
                       "file_bytes_per_step": F * (3 * (8 + 4 * 3 * n)), "h2d_bytes_per_step": (s41["h2d_bytes"] - s40["h2d_bytes"]) / K,
                       "reader_threads": args.reader_threads,
                       "api": "newanalysis.gfunction.RDF.calcTrajectory(DCDFile) + readout per step (file in the page cache)"}
    return out


def run_single_process(args):
    """One process, N devices (RDF(devices=range(N)), NEWANALYSIS_GPUS=all): frames go round-robin over the devices' rings,
    histograms are combined at readout.  Host arrays only - page-locked (strided DMA, no host copy) and pageable (frames of
    a batch copied by the pool threads)."""
    import newanalysis.gfunction as G
    name = args.workload
    c = cfg(name)
    n1, n2, L = c["n1"], c["n2"], c["L"]
    per = synth.pair_evals(n1, n2)
    N = args.gpus
    F, K, W = (args.frames_per_step or c["F"]) * N, args.steps, args.warmup
    POOL = max(c["pool"], 2 * F if name == "cfg3" else F)
    pc1, pd1, pc2, pd2 = make_pool(c, POOL, 0, alloc=G.pinned_empty)
    qc1, qd1 = np.array(pc1), np.array(pd1)
    out = {"mode": "single_process", "metric": METRIC, "unit": UNIT, "n_gpus": N, "steps": K, "warmup": W, "config": public_config(name, c),
           "run": {"frames_per_step": F, "pool_frames": POOL}}
    for label, (a1, b1), pinned in (("e2e_pinned", (pc1, pd1), True), ("e2e_pageable", (qc1, qd1), False)):
        r = G.RDF(c["modes"], c["hmin"], c["hmax"], c["dr"], n1, n2, n1, n2, L, devices=list(range(N)))

        def step(s):
            lo = (s * F) % (POOL - F + 1) if POOL > F else 0
            r.calcFrames(a1[lo:lo + F], a1[lo:lo + F], b1[lo:lo + F], b1[lo:lo + F], pinned=pinned)
            return r.raw_histogram()

        for s in range(W):
            step(s)
        r.sync()
        s0 = r.stats()
        t0 = time.perf_counter()
        for s in range(K):
            hist = step(W + s)
        r.sync()
        dt = time.perf_counter() - t0
        s1 = r.stats()
        out[label] = {"value": K * F * per / dt / 1e9, "ms_per_step": dt / K * 1e3, "h2d_bytes_per_step": (s1["h2d_bytes"] - s0["h2d_bytes"]) / K,
                      "kernel_ms_sum_over_devices_per_step": (s1["kernel_ms"] - s0["kernel_ms"]) / K, "ndev": s1["ndev"],
                      "kernel_ms_busiest_device_per_step": (s1["kernel_ms_max_device"] - s0["kernel_ms_max_device"]) / K,
                      "frames_per_device_min_max": [s1["frames_min_device"] - s0["frames_min_device"], s1["frames_max_device"] - s0["frames_max_device"]],
                      "counts_exact": bool(float(hist[:HN].sum()) == 2.0 * (s1["in_range"] - s1["overflow"]))}
        del r
    print(json.dumps(out), flush=True)


def _latest_profile(fname):
    import glob
    import re
    hits = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*", fname)),
                  key=lambda f: [int(v) for v in re.findall(r"\d+", os.path.basename(os.path.dirname(f)))])
    return hits[-1] if hits else None


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one pair_kernel launch, from the committed summary of the
    latest `ncu --set full` capture of this very command (profiles/<round>/pair_kernel_raw_selected.txt); None if absent."""
    f = _latest_profile("pair_kernel_raw_selected.txt")
    if not f:
        return None, "no capture"
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for ln in open(f):
        t = ln.split()
        if len(t) == 3 and t[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(t[2]) * mult.get(t[1], 1.0)
    return tot, os.path.relpath(f, ROOT)


def ncu_pipes():
    """Pipe utilisations of the pair kernel from the same committed capture (per cent of peak): FP64 pipe, issue slots,
    shared-memory data pipe, fp16 pipe.  They explain `frac`: the kernel is reference-flops / time, not FP64-pipe busy."""
    f = _latest_profile("pair_kernel_raw_selected.txt")
    if not f:
        return None
    want = {"sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_util",
            "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_pipe",
            "sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active": "fp16_pipe_util"}
    out = {"source": os.path.relpath(f, ROOT)}
    for ln in open(f):
        t = ln.split()
        if len(t) >= 2 and t[0] in want:
            try:
                out.setdefault(want[t[0]], float(t[-1]))
            except ValueError:
                pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg4", "cfg5"],
                    help="BASELINE.json configuration: cfg3 (default, the metric's config), cfg4 (four objects per frame), cfg5 (1M sites)")
    ap.add_argument("--filter", default="auto", choices=["auto", "fp16", "fp32", "fp64"],
                    help="range pre-filter: auto = what the RDF class picks itself (fp16x2 for in-box coordinates)")
    ap.add_argument("--frames-per-step", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--single-process", action="store_true", help="one process drives --gpus devices (run without torchrun)")
    ap.add_argument("--reader-threads", type=int, default=2, help="decode threads of the DCD arm")
    ap.add_argument("--flags", default="", help="extra gfn flags (experiments)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = {"cfg3": 20, "cfg4": 6, "cfg5": 6}[args.workload] if args.impl == "ours" else {"cfg3": 20, "cfg4": 6, "cfg5": 10}[args.workload]
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.single_process:
        run_single_process(args)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

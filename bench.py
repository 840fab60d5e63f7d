#!/usr/bin/env python3
"""bench.py - pair evaluations per second of the multipole g-function path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[2] - water, 32,768 molecules, self g-function, all
seven multipole modes, histogram 0-15 A / 0.05 A (300 bins), cubic box L = 99.33 A, synthetic uniform
coordinates and normal dipoles (gfn_b200/synth.py).  One STEP = one pass of the hot path over a batch
of FRAMES_PER_STEP frames; one pair evaluation = one iteration of the reference j loop
(gfunction.pyx:130), n(n+1)/2 per frame.

  value      device-timed (CUDA events on the library's compute stream, gfn_mark) throughput with the
             frames already resident in HBM; N ranks each process their own frames (weak scaling).
  e2e        the same metric through the reference-facing call RDF.calcFrame with HOST numpy arrays:
             copy into the pinned ring, H2D, kernels, and a histogram readout (D2H, plus the NCCL
             all-reduce over ranks when N > 1) every step, wall clock between two device syncs.
  roofline   FP64 vector pipe: algorithmic flops/eval x evals / pair-kernel time
             (CUDA events around every pair-kernel launch, summed by the library) against the DFMA
             peak measured in this run (gfn_measure_peaks); nominal 37.2 TFLOP/s beside it.
  cpu_baseline  the UNMODIFIED reference module (oracle/_ref, kind "reference"; else the C port) timed on
             this box's host cores on a bounded sample of the same workload.

--impl reference times the reference's own CPU implementation (oracle/_ref) on the same config, one
frame per step.
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
WORKLOAD = "cfg3"
FRAMES_PER_STEP = 64
POOL_FRAMES = 128
NOMINAL_FP64_TFLOPS = 37.2


def cfg():
    c = dict(synth.CONFIGS[WORKLOAD])
    n = int(os.environ.get("GFN_BENCH_N", c["n1"]))
    if n != c["n1"]:                     # reduced-size dry runs only (never the reported number)
        c["L"] = c["L"] * (n / c["n1"]) ** (1.0 / 3.0)
        c["n1"] = c["n2"] = n
    return c


def make_pool(c, nframes, rank):
    n = c["n1"]
    co = np.empty((nframes, n, 3)); di = np.empty((nframes, n, 3))
    for f in range(nframes):
        co[f], di[f] = synth.frame(c["cfg"], f, n, c["L"], stream=1000 * rank)
    return co, di


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


def reference_rdf(c):
    """(RDF-like object, kind): the unmodified reference class if oracle/_ref is present, else the port."""
    import oracle as O
    ref = O.load_ref()
    n = c["n1"]
    args = (c["modes"], c["hmin"], c["hmax"], c["dr"], n, n, n, n, c["L"])
    if ref is not None:
        return ref.RDF(*args), "reference"
    return O.OracleRDF(*args), "port"


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


def time_reference(c, steps, warmup, rank=0):
    """One frame per step through the reference's own RDF.calcFrame on the host cores."""
    use_all_host_threads()
    rdf, kind = reference_rdf(c)
    n = c["n1"]
    per = synth.pair_evals(n, n)
    frames = [synth.frame(c["cfg"], f, n, c["L"], stream=1000 * rank) for f in range(min(4, warmup + steps))]
    for s in range(warmup):
        x, d = frames[s % len(frames)]
        rdf.calcFrame(x, x, d, d)
    t0 = time.perf_counter()
    for s in range(steps):
        x, d = frames[(warmup + s) % len(frames)]
        rdf.calcFrame(x, x, d, d)
    dt = time.perf_counter() - t0
    return per * steps / dt / 1e9, dt, kind, "%d full frames of %s (n=%d, %d pair evals each), %d warm-up" % (steps, WORKLOAD, n, per, warmup)


def run_reference(args, rank, world):
    if rank != 0:
        return
    c = cfg()
    val, dt, kind, sample = time_reference(c, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n": c["n1"], "modes": "all", "bins": 300, "frames_per_step": 1,
                   "note": "reference Cython/OpenMP calcHisto on host cores, one full frame per step"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": host_threads(), "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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

    c = cfg()
    n, L = c["n1"], c["L"]
    per = synth.pair_evals(n, n)
    F, K, W = args.frames_per_step, args.steps, args.warmup
    flags = ("" if args.filter == "auto" else args.filter) + ("," + args.flags if args.flags else "")
    mk = lambda: G.RDF(c["modes"], c["hmin"], c["hmax"], c["dr"], n, n, n, n, L, devices=[local_rank], gfn_flags=flags)  # noqa: E731

    uid = sharding.broadcast_unique_id(dist, G.comm_unique_id, device="cuda") if world > 1 else None

    sampler = ClockSampler(local_rank)
    sampler.start()
    pool_c, pool_d = make_pool(c, POOL_FRAMES, rank)
    fp64_peak, fp32_peak = G.measure_peaks(local_rank)

    # ---------------- arm 1: inputs resident in HBM, device-timed ------------------------------------
    r = mk()
    if uid is not None:
        r.attach_comm(uid, world, rank)
    dc, dd = G.to_device(pool_c, local_rank), G.to_device(pool_d, local_rank)

    def dev_step(s):
        first = (s * F) % (POOL_FRAMES - F + 1) if POOL_FRAMES > F else 0
        r.calcFramesDevice(F, dc, dc, dd, dd, first=first)

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
    hn = 300
    frames_done = (W + K) * F * world
    assert abs(hist[:hn].sum() - 2.0 * st1["in_range"] * (world if world > 1 else 1)) <= 0.02 * hist[:hn].sum() + 1, \
        "g000 checksum does not match the in-range pair count"
    value = world * K * F * per / (ms * 1e-3) / 1e9
    kernel_ms = st1["kernel_ms"] - st0["kernel_ms"]
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    pair_launches = st1["pair_launches"] - st0["pair_launches"]
    f_in = (st1["in_range"] - st0["in_range"]) / (st1["pair_evals"] - st0["pair_evals"])
    f_surv = (st1["survivors"] - st0["survivors"]) / (st1["pair_evals"] - st0["pair_evals"])
    flops_eval = synth.flops_per_eval(127, f_in)
    achieved = flops_eval * K * F * per / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else None

    # ---------------- arm 2: end to end through RDF.calcFrame with host arrays -----------------------
    e = mk()
    if world > 1:           # an NCCL unique id makes exactly one communicator: a second handle needs its own
        e.attach_comm(sharding.broadcast_unique_id(dist, G.comm_unique_id, device="cuda"), world, rank)

    def host_step(s):
        first = (s * F) % (POOL_FRAMES - F + 1) if POOL_FRAMES > F else 0
        for f in range(first, first + F):
            e.calcFrame(pool_c[f], pool_c[f], pool_d[f], pool_d[f])
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

    # ---------------- arm 3 (N=1 only): end to end from float32 ATOM positions (3-site water) ----------
    # RDF.calcFrameAtoms: the frame crosses PCIe as the trajectory's native float32 positions, centres of mass and
    # dipoles (helpers.comByResidue / dipByResidue of the reference workflow) are produced on the device
    atoms = None
    if world == 1:
        rng = np.random.default_rng(4242)
        off = (0.6 * rng.standard_normal((n, 3, 3))).astype(np.float32)
        pool_a = np.ascontiguousarray((pool_c.astype(np.float32)[:, :, None, :] + off[None]).reshape(POOL_FRAMES, 3 * n, 3))
        a = mk()
        a.attach_topology(np.tile([15.999, 1.008, 1.008], n), np.full(n, 3, dtype=np.int32), (3 * np.arange(n)).astype(np.int32),
                          np.tile([-0.8476, 0.4238, 0.4238], n))

        def atom_step(s):
            first = (s * F) % (POOL_FRAMES - F + 1) if POOL_FRAMES > F else 0
            for f in range(first, first + F):
                a.calcFrameAtoms(pool_a[f])
            return a.raw_histogram()

        for s in range(W):
            atom_step(s)
        a.sync()
        sa0 = a.stats()
        t0 = time.perf_counter()
        for s in range(K):
            atom_step(W + s)
        a.sync()
        dta = time.perf_counter() - t0
        sa1 = a.stats()
        atoms = {"value": K * F * per / dta / 1e9, "unit": UNIT, "h2d_bytes_per_step": (sa1["h2d_bytes"] - sa0["h2d_bytes"]) / K,
                 "d2h_bytes_per_step": 8 * hn * 8 + 8 * 8, "ms_per_step": dta / K * 1e3,
                 "gpu_launches": int(sa1["kernel_launches"] - sa0["kernel_launches"]),
                 "api": "newanalysis.gfunction.RDF.calcFrameAtoms (host float32 positions, 3 atoms per site; COM + dipoles on the device) + readout per step"}

        # ---------------- arm 4 (N=1 only): from a DCD file on disk (page cache) through reader threads -------------
        tmpd = tempfile.mkdtemp(prefix="gfn_bench_")
        path = os.path.join(tmpd, "pool.dcd")
        synth.write_dcd(path, pool_a)
        traj = G.DCDFile(path)
        sel = np.arange(3 * n, dtype=np.int32)
        t4 = mk()
        t4.attach_topology(np.tile([15.999, 1.008, 1.008], n), np.full(n, 3, dtype=np.int32), (3 * np.arange(n)).astype(np.int32),
                           np.tile([-0.8476, 0.4238, 0.4238], n))

        def traj_step(s):
            first = (s * F) % (POOL_FRAMES - F + 1) if POOL_FRAMES > F else 0
            t4.calcTrajectory(traj, sel, first=first, last=first + F, threads=args.reader_threads)
            return t4.raw_histogram()

        for s in range(W):
            traj_step(s)
        t4.sync()
        s40 = t4.stats()
        t0 = time.perf_counter()
        for s in range(K):
            traj_step(W + s)
        t4.sync()
        dtt = time.perf_counter() - t0
        s41 = t4.stats()
        traj.close()
        os.unlink(path); os.rmdir(tmpd)
        atoms["from_dcd"] = {"value": K * F * per / dtt / 1e9, "unit": UNIT, "ms_per_step": dtt / K * 1e3,
                             "file_bytes_per_step": F * (3 * (8 + 4 * 3 * n)), "h2d_bytes_per_step": (s41["h2d_bytes"] - s40["h2d_bytes"]) / K,
                             "reader_threads": args.reader_threads,
                             "api": "newanalysis.gfunction.RDF.calcTrajectory(DCDFile) + readout per step (file in the page cache)"}
    sampler.stop()
    e2e_value = world * K * F * per / dt / 1e9
    h2d = (se1["h2d_bytes"] - se0["h2d_bytes"]) / K
    d2h = 8 * hn * 8 + 8 * 8

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        val, cdt, kind, sample = time_reference(c, 2, 1)
        cpu = {"value": val, "unit": UNIT, "cores": host_threads(), "kind": kind, "sample": sample}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n": n, "L": L, "modes": "all", "bins": hn, "frames_per_step": F,
                   "pair_evals_per_step": F * per, "parallelism": "frames sharded over %d GPU(s), one NCCL all-reduce at readout" % world,
                   "l2": "inputs exceed L2: %d-frame pool (%.0f MB fp64) cycled, records re-derived every step" % (POOL_FRAMES, 2 * pool_c.nbytes / 1e6),
                   "filter": st1["filter"], "hist": st1["hist_mode"], "batch_frames": st1["batch_frames"], "grid": st1["grid"],
                   "warps_per_block": st1["warps_per_block"], "smem_bytes": st1["smem_bytes"]},
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": (achieved / fp64_peak) if achieved else None, "traffic": ncu_traffic()[0],
                     "traffic_unit": "bytes of DRAM read+write per pair_kernel launch (ncu --set full capture, %s); algorithmic: %d"
                                     % (ncu_traffic()[1], st1["batch_frames"] * ((n + 1023) // 1024 * 1024) * 80),
                     "peak_source": "measured in this run: dependent-free DFMA loop (gfn_measure_peaks); nominal %.1f" % NOMINAL_FP64_TFLOPS,
                     "frac_of_nominal": (achieved / NOMINAL_FP64_TFLOPS) if achieved else None,
                     "flops_per_eval": flops_eval, "in_range_fraction": f_in, "survivor_fraction": f_surv,
                     "kernel": "pair_kernel", "kernel_ms_per_launch": kernel_ms / max(pair_launches, 1),
                     "kernel_share_of_step": kernel_ms / ms, "fp32_peak_tflops": fp32_peak,
                     "lever": {"fp32": "fp32 conservative pre-filter + exact fp64 survivors",
                               "fp16x2": "packed-fp16 conservative pre-filter (two core sites per instruction) + exact fp64 survivors",
                               "fp64": "fp64 pre-test + exact fp64 survivors"}[st1["filter"]]},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": dt / K * 1e3, "api": "newanalysis.gfunction.RDF.calcFrame (host numpy arrays) + readout per step"},
        "gpu_launches": int(launches),
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                   "samples": clocks["samples"], "e2e": clocks_e2e},
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if atoms is not None:
        line["e2e_atoms"] = atoms
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one pair_kernel launch, from the committed summary of the
    latest `ncu --set full` capture of this very command (profiles/<round>/pair_kernel_raw_selected.txt); None if absent."""
    import glob
    import re
    hits = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r*", "pair_kernel_raw_selected.txt")),
                  key=lambda f: [int(v) for v in re.findall(r"\d+", os.path.basename(os.path.dirname(f)))])
    if not hits:
        return None, "no capture"
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for ln in open(hits[-1]):
        t = ln.split()
        if len(t) == 3 and t[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(t[2]) * mult.get(t[1], 1.0)
    return tot, os.path.relpath(hits[-1], os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--filter", default="auto", choices=["auto", "fp16", "fp32", "fp64"],
                    help="range pre-filter: auto = what the RDF class picks itself (fp16x2 for in-box coordinates)")
    ap.add_argument("--frames-per-step", type=int, default=FRAMES_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reader-threads", type=int, default=2, help="decode threads of the DCD arm")
    ap.add_argument("--flags", default="", help="extra gfn flags (experiments)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

"""Seeded synthetic periodic boxes of the BASELINE.json shapes.

Pure numpy, no GPU, no oracle: shared by tests/, bench.py and __graft_entry__.smoke().
seed = 1000*cfg + frame, rng = np.random.default_rng(seed); coordinates uniform in [0,L)^3 (float64),
dipoles standard normal (non-zero a.s.; the g-function formulas normalise them).
"""
import numpy as np

# name -> dict(n1, n2, L, modes, hmin, hmax, dr, frames, self)
CONFIGS = {
    # cfg1: SPC/E water, 1000 O sites, O-O g000 0-15 A dr 0.05, 100 frames (reference CPU-runnable case)
    "cfg1": dict(cfg=1, n1=1000, n2=1000, L=31.04, modes=["000"], hmin=0.0, hmax=15.0, dr=0.05,
                 frames=100, self_pairs=True),
    # cfg2: [C2mim][OTf] 1000 cation COM x 1000 anion COM, distinct selections of equal size (quirk Q1)
    "cfg2": dict(cfg=2, n1=1000, n2=1000, L=67.7, modes=["000", "110", "101", "011", "220"], hmin=0.0,
                 hmax=30.0, dr=0.05, frames=1000, self_pairs=False),
    # cfg3: water 32,768 molecules, self g-function, all dipole modes, 1 B200 (the headline metric config)
    "cfg3": dict(cfg=3, n1=32768, n2=32768, L=99.33, modes=["all"], hmin=0.0, hmax=15.0, dr=0.05,
                 frames=1000, self_pairs=True),
    # cfg5: 1M-site box, self, all modes
    "cfg5": dict(cfg=5, n1=1048576, n2=1048576, L=315.4, modes=["all"], hmin=0.0, hmax=15.0, dr=0.05,
                 frames=10000, self_pairs=True),
}

# cfg4: mixed solute/solvent box, 262,144 atoms; four RDF objects per frame sharing uploads
CFG4 = dict(cfg=4, L=143.6, n_solvent=81920, n_solute=1024, n_solute_atoms=4096, frames=2000,
            hmin=0.0, hmax=15.0, dr=0.05,
            pairs=[("solute_com", "solvent_com", ["all"]),
                   ("solvent_com", "solvent_com", ["all"]),
                   ("solute_atoms", "solvent_o", ["000"]),
                   ("solute_com", "solute_com", ["all"])])


def pair_evals(n1, n2):
    """One unit of work = one iteration of the reference j loop (gfunction.pyx:126-130):
    n(n+1)/2 when the selections have equal size (quirk Q1), else n1*n2."""
    return n1 * (n1 + 1) // 2 if n1 == n2 else n1 * n2


def frame(cfg, frame_idx, n, L, with_dipoles=True, stream=0, f32_roundtrip=False):
    """One selection of one frame: (coords (n,3) f64, dipoles (n,3) f64 or None)."""
    rng = np.random.default_rng(1000 * cfg + frame_idx + 7919 * stream)
    xyz = rng.random((n, 3)) * L
    if f32_roundtrip:
        xyz = xyz.astype(np.float32).astype(np.float64)
    dip = rng.standard_normal((n, 3)) if with_dipoles else None
    return np.ascontiguousarray(xyz), (np.ascontiguousarray(dip) if dip is not None else None)


def config_frame(name, frame_idx, n_override=None):
    """(c1, c2, d1, d2) for a named config; self configs return the same arrays twice."""
    c = CONFIGS[name]
    n1 = n_override or c["n1"]
    n2 = n_override or c["n2"]
    c1, d1 = frame(c["cfg"], frame_idx, n1, c["L"])
    if c["self_pairs"]:
        return c1, c1, d1, d1
    c2, d2 = frame(c["cfg"], frame_idx, n2, c["L"], stream=1)
    return c1, c2, d1, d2


def flops_per_eval(mode_sel, f_in):
    """Algorithmic FP64 flops per pair evaluation, SURVEY.md section 8d convention."""
    f_eval = 3 + 1.5 + 5 + 1 + 1
    f_inr = 2.0
    if mode_sel & 1:
        f_inr += 2
    for lo, hi in ((2, 16), (4, 32), (8, 64)):
        if mode_sel & (lo | hi):
            f_inr += 7 + (2 if mode_sel & lo else 0) + (5 if mode_sel & hi else 0)
    return f_eval + f_in * f_inr


def in_range_fraction(hmax, L):
    import math
    return min(1.0, 4.0 * math.pi * hmax ** 3 / (3.0 * L ** 3))


def write_dcd(path, positions, cells=None):
    """Synthetic trajectory file for bench.py / examples: positions (nframes, natoms, 3) float32 -> little-endian CHARMM
    DCD (optional per-frame cells A, B, C, alpha, beta, gamma in degrees).  Input generation only - the product reads
    trajectories with libgfn.so (newanalysis.gfunction.DCDFile); the checker's independent writer / reader with all the
    format variants lives in oracle/dcd.py and is not used outside tests/."""
    import struct
    pos = np.ascontiguousarray(positions, dtype="<f4")
    nf, na = pos.shape[0], pos.shape[1]
    icntrl = np.zeros(20, dtype="<i4")
    icntrl[0], icntrl[2], icntrl[3] = nf, 1, nf
    icntrl[10] = 1 if cells is not None else 0
    icntrl[19] = 24
    head = b"CORD" + icntrl.tobytes()
    head = head[:40] + struct.pack("<f", 0.002) + head[44:]          # icntrl[9]: time step, float32 in CHARMM files
    title = struct.pack("<i", 1) + b"synthetic trajectory (gfn_b200.synth.write_dcd)".ljust(80)
    mark = struct.pack("<i", 4 * na)
    with open(path, "wb") as f:
        for payload in (head, title, struct.pack("<i", na)):
            f.write(struct.pack("<i", len(payload)) + payload + struct.pack("<i", len(payload)))
        for i in range(nf):
            if cells is not None:
                A, B, C, al, be, ga = [float(v) for v in cells[i]]
                f.write(struct.pack("<i6di", 48, A, ga, B, be, al, C, 48))
            for k in range(3):
                f.write(mark); f.write(np.ascontiguousarray(pos[i, :, k]).tobytes()); f.write(mark)

"""Definitions of the golden parity cases (inputs are regenerated from seeds; outputs are committed).

Each case is a dict understood by `build_inputs`; the expected outputs in tests/golden/*.npz were
produced by the UNMODIFIED reference module (oracle/_ref) via tests/golden/make_golden.py.
"""
import hashlib

import numpy as np

CASES = {
    # name: n1, n2, box (x,y,z | y,z None = cubic), modes, hmin, hmax, dr, frames, same, extras
    "cfg1_shape":   dict(n1=1000, n2=1000, box=(31.04, None, None), modes=["000"], hmin=0.0, hmax=15.0, dr=0.05, frames=3, same=True, seed=1),
    "cfg2_shape":   dict(n1=1000, n2=1000, box=(67.7, None, None), modes=["000", "110", "101", "011", "220"], hmin=0.0, hmax=30.0, dr=0.05, frames=2, same=False, seed=2),
    "cfg3_density": dict(n1=2048, n2=2048, box=(39.42, None, None), modes=["all"], hmin=0.0, hmax=15.0, dr=0.05, frames=2, same=True, seed=3),
    "rect_all":     dict(n1=300, n2=400, box=(40.0, None, None), modes=["all"], hmin=0.0, hmax=15.0, dr=0.05, frames=2, same=False, seed=4),
    "rect_wide":    dict(n1=37, n2=1531, box=(52.0, None, None), modes=["all"], hmin=0.0, hmax=20.0, dr=0.1, frames=2, same=False, seed=5),
    "self_odd":     dict(n1=257, n2=257, box=(31.04, None, None), modes=["all"], hmin=0.0, hmax=15.0, dr=0.05, frames=2, same=True, seed=6),
    "ortho_npt":    dict(n1=333, n2=333, box=(30.0, 36.0, 41.0), modes=["all"], hmin=0.0, hmax=14.0, dr=0.05, frames=3, same=True, seed=7, npt=True),
    "g000_nodip":   dict(n1=200, n2=500, box=(35.0, None, None), modes=["000"], hmin=0.0, hmax=15.0, dr=0.05, frames=2, same=False, seed=8, nodip=True),
    "hmin_pos":     dict(n1=400, n2=400, box=(32.0, None, None), modes=["000", "110", "202"], hmin=2.0, hmax=12.0, dr=0.1, frames=2, same=True, seed=9),
    "f32_bounds":   dict(n1=256, n2=256, box=(40.0, None, None), modes=["all"], hmin=0.0, hmax=15.02, dr=0.05, frames=1, same=True, seed=10),
    "beyond_half":  dict(n1=128, n2=128, box=(20.0, None, None), modes=["all"], hmin=0.0, hmax=15.0, dr=0.05, frames=2, same=True, seed=11),
    "unwrapped":    dict(n1=150, n2=170, box=(25.0, None, None), modes=["all"], hmin=0.0, hmax=12.0, dr=0.05, frames=2, same=False, seed=12, unwrapped=True),
    "tiny_1x1":     dict(n1=1, n2=1, box=(40.0, None, None), modes=["all"], hmin=0.0, hmax=15.0, dr=0.05, frames=1, same=False, seed=13),
    "tiny_1x5":     dict(n1=1, n2=5, box=(12.0, None, None), modes=["all"], hmin=0.0, hmax=5.0, dr=0.05, frames=2, same=False, seed=14),
    "distinct_eq_subset": dict(n1=64, n2=64, box=(18.0, None, None), modes=["110", "011", "022"], hmin=0.0, hmax=8.0, dr=0.05, frames=2, same=False, seed=15),
}


def build_inputs(case, frame):
    """-> (c1, c2, d1, d2, box_xyz_for_this_frame or None)."""
    n1, n2 = case["n1"], case["n2"]
    bx, by, bz = case["box"]
    L = np.array([bx, by or bx, bz or bx], dtype=np.float64)
    rng = np.random.default_rng(100003 * case["seed"] + frame)
    newbox = None
    if case.get("npt"):
        s = 1.0 + 1e-3 * rng.standard_normal(3)
        L = L * s
        newbox = tuple(float(v) for v in L)
    lo, span = (-1.0, 3.0) if case.get("unwrapped") else (0.0, 1.0)
    c1 = np.ascontiguousarray((lo + span * rng.random((n1, 3))) * L)
    d1 = np.ascontiguousarray(rng.standard_normal((n1, 3)))
    if case["same"]:
        c2, d2 = c1, d1
    else:
        c2 = np.ascontiguousarray((lo + span * rng.random((n2, 3))) * L)
        d2 = np.ascontiguousarray(rng.standard_normal((n2, 3)))
    if case.get("nodip"):
        d1 = d2 = None
    return c1, c2, d1, d2, newbox


def input_digest(case):
    h = hashlib.sha256()
    for f in range(case["frames"]):
        for a in build_inputs(case, f)[:4]:
            if a is not None:
                h.update(a.tobytes())
    return h.hexdigest()


def make_rdf(cls, case):
    n1, n2 = case["n1"], case["n2"]
    bx, by, bz = case["box"]
    return cls(case["modes"], case["hmin"], case["hmax"], case["dr"], n1, n2, n1, n2, bx, by, bz)


def drive(rdf, case):
    """Run all frames of a case through an RDF-like object (reference, oracle or the GPU shim)."""
    for f in range(case["frames"]):
        c1, c2, d1, d2, newbox = build_inputs(case, f)
        if newbox is not None:
            rdf.update_box(*newbox)
        if d1 is None:
            rdf.calcFrame(c1, c2)
        else:
            rdf.calcFrame(c1, c2, d1, d2)
    return rdf


# hand-checkable micro cases: x=(0,0,0), 0-15/0.05, L=40 -> flat indices
MICRO = [((15.0, 0.0, 0.0), 300), ((14.999999999, 0.0, 0.0), 299), ((0.05, 0.0, 0.0), 1),
         ((0.049999, 0.0, 0.0), None)]


# N1 producers (helpers.comByResidue / dipByResidue): residues of ragged size, incl. single-atom and empty ones
RESIDUE_CASES = {
    "water":   dict(nres=500, apr=("const", 3), L=30.0, seed=21),
    "ragged":  dict(nres=333, apr=("range", 1, 40), L=45.0, seed=22),
    "ions":    dict(nres=64, apr=("range", 1, 2), L=20.0, seed=23),
    "gaps":    dict(nres=100, apr=("range", 0, 6), L=25.0, seed=24, gaps=True),
}


def build_residues(case):
    """-> (coor (natoms,3), masses, charges, apr int32, rfa int32).  'gaps' leaves unused atoms between
    residues and lets residues be empty (apr = 0: the reference then divides 0 by 0 -> nan, reproduced)."""
    rng = np.random.default_rng(900001 * case["seed"])
    nres = case["nres"]
    kind = case["apr"]
    apr = np.full(nres, kind[1], dtype=np.int32) if kind[0] == "const" else rng.integers(kind[1], kind[2] + 1, nres).astype(np.int32)
    gap = rng.integers(0, 3, nres) if case.get("gaps") else np.zeros(nres, dtype=np.int64)
    rfa = np.zeros(nres, dtype=np.int32)
    pos = 0
    for i in range(nres):
        pos += int(gap[i])
        rfa[i] = pos
        pos += int(apr[i])
    natoms = max(pos, 1)
    coor = np.ascontiguousarray(rng.random((natoms, 3)) * case["L"])
    masses = np.ascontiguousarray(1.0 + 15.0 * rng.random(natoms))
    charges = np.ascontiguousarray(rng.standard_normal(natoms))
    return coor, masses, charges, apr, rfa


# N2 shell-resolved g-functions (RDF_voronoi, gfunction.pyx:764-1148) and calcHistoTSVoro (:195-224).
# nmol*: molecules per selection (sites per molecule = n // nmol); the Delaunay matrix has shape (nmol1, n2) because
# the reference indexes it with row stride n2 (:602); entries are drawn from [-1, nshells + 2] so that the clamping
# of 0, negative and too large values is exercised.
VORO_CASES = {
    "voro_self":      dict(n1=600, n2=600, nmol1=200, nmol2=200, box=(28.0, None, None), modes=["all"], hmin=0.0, hmax=13.0, dr=0.1, nshells=3, frames=2, same=True, seed=31),
    "voro_self_excl": dict(n1=600, n2=600, nmol1=200, nmol2=200, box=(28.0, None, None), modes=["all"], hmin=0.0, hmax=13.0, dr=0.1, nshells=3, frames=2, same=True, seed=32, exclude_self=True),
    "voro_rect":      dict(n1=150, n2=420, nmol1=50, nmol2=210, box=(30.0, 33.0, 36.0), modes=["000", "110", "202", "022"], hmin=1.0, hmax=14.0, dr=0.05, nshells=4, frames=3, same=False, seed=33, npt=True),
    "voro_mol":       dict(n1=500, n2=500, nmol1=500, nmol2=500, box=(31.04, None, None), modes=["all"], hmin=0.0, hmax=15.0, dr=0.05, nshells=2, frames=2, same=False, seed=34, exclude_self=True),
    "voro_g000":      dict(n1=256, n2=128, nmol1=256, nmol2=64, box=(24.0, None, None), modes=["000"], hmin=0.0, hmax=12.0, dr=0.05, nshells=1, frames=2, same=False, seed=35, nodip=True),
    "voro_many":      dict(n1=2048, n2=2048, nmol1=2048, nmol2=2048, box=(39.42, None, None), modes=["all"], hmin=0.0, hmax=15.0, dr=0.05, nshells=5, frames=1, same=True, seed=36),
}


def build_voro_inputs(case, frame):
    """-> (c1, c2, d1, d2, newbox, delaunay int32 (nmol1, n2))."""
    c1, c2, d1, d2, newbox = build_inputs(case, frame)
    rng = np.random.default_rng(700001 * case["seed"] + frame)
    dl = rng.integers(-1, case["nshells"] + 3, size=(case["nmol1"], case["n2"])).astype(np.int32)
    return c1, c2, d1, d2, newbox, np.ascontiguousarray(dl)


def voro_digest(case):
    h = hashlib.sha256()
    for f in range(case["frames"]):
        for a in build_voro_inputs(case, f):
            if isinstance(a, np.ndarray):
                h.update(a.tobytes())
    return h.hexdigest()


def make_voro(cls, case, **kw):
    bx, by, bz = case["box"]
    return cls(case["modes"], case["hmin"], case["hmax"], case["dr"], case["n1"], case["n2"], case["nmol1"], case["nmol2"],
               bx, case["nshells"], by, bz, exclude_self_mol=bool(case.get("exclude_self")), **kw)


def drive_voro(rdf, case):
    for f in range(case["frames"]):
        c1, c2, d1, d2, newbox, dl = build_voro_inputs(case, f)
        if newbox is not None:
            rdf.update_box(*newbox)
        if d1 is None:
            rdf.calcFrame(c1, c2, dl)
        else:
            rdf.calcFrame(c1, c2, dl, d1, d2)
    return rdf


TSVORO_CASES = {
    "tsvoro_small": dict(n1=40, n2=90, box=22.0, hmin=0.0, hmax=10.0, dr=0.1, maxshell=3, frames=3, seed=41),
    "tsvoro_self":  dict(n1=300, n2=300, box=26.0, hmin=0.5, hmax=12.0, dr=0.05, maxshell=4, frames=2, seed=42, same=True),
}


def build_tsvoro_inputs(case, frame):
    """-> (core, surr, ds_t int8 (n1, n2)); shells drawn from [0, maxshell + 2] (0 wraps to the last shell)."""
    rng = np.random.default_rng(800001 * case["seed"] + frame)
    core = np.ascontiguousarray(rng.random((case["n1"], 3)) * case["box"])
    surr = core if case.get("same") else np.ascontiguousarray(rng.random((case["n2"], 3)) * case["box"])
    ds_t = rng.integers(0, case["maxshell"] + 3, size=(case["n1"], case["n2"])).astype(np.int8)
    return core, surr, ds_t


def drive_tsvoro(fn, case):
    """fn(core, surr, result, ds, maxshell, boxl, t, hmin, hmax, invdx) for every frame -> result (frames, n1, maxshell, nbins)."""
    nb = int(round((case["hmax"] - case["hmin"]) / case["dr"]))
    T = case["frames"]
    result = np.zeros((T, case["n1"], case["maxshell"], nb))
    ds = np.zeros((T, case["n1"], case["n2"]), dtype=np.int8)
    frames = [build_tsvoro_inputs(case, t) for t in range(T)]
    for t in range(T):
        ds[t] = frames[t][2]
    for t in range(T):
        fn(frames[t][0], frames[t][1], result, ds, case["maxshell"], case["box"], t, case["hmin"], case["hmax"], 1.0 / case["dr"])
    return result


# N3 RDF time series (calcHistoTS, gfunction.pyx:168-192): result[t, i, pos] += 1 for all n1 x n2 pairs, cubic box.
# The fixture ts.npz holds what the UNMODIFIED reference function wrote.  `edge` plants surround sites at exactly
# histo_max from core sites (quirk Q3: pos == nbins, which the reference's unchecked index turns into bin 0 of the next
# core row - or of the next frame's first row for the last core site); the result array therefore gets one spare frame.
TS_CASES = {
    "ts_rect":  dict(n1=150, n2=400, box=30.0, hmin=0.0, hmax=12.0, dr=0.05, frames=3, seed=61),
    "ts_self":  dict(n1=220, n2=220, box=27.0, hmin=0.0, hmax=12.0, dr=0.05, frames=2, seed=62, same=True),
    "ts_hmin":  dict(n1=64, n2=300, box=24.0, hmin=1.5, hmax=11.5, dr=0.1, frames=2, seed=63),
    "ts_edge":  dict(n1=5, n2=40, box=40.0, hmin=0.0, hmax=15.0, dr=0.05, frames=2, seed=64, edge=True),
}


def build_ts_inputs(case, frame):
    """-> (core (n1,3), surr (n2,3))."""
    rng = np.random.default_rng(600001 * case["seed"] + frame)
    core = np.ascontiguousarray(rng.random((case["n1"], 3)) * case["box"])
    surr = core if case.get("same") else np.ascontiguousarray(rng.random((case["n2"], 3)) * case["box"])
    if case.get("edge"):
        # exact distances histo_max (15.0) along x from core sites 1 and n1-1 (the last row), straight and across the boundary
        core[1] = (2.0, 3.0, 4.0); surr[0] = (17.0, 3.0, 4.0)
        core[case["n1"] - 1] = (30.0, 8.0, 9.0); surr[1] = (5.0, 8.0, 9.0)       # 5 - 30 = -25 -> +40 = 15
    return core, surr


def ts_nbins(case):
    return int(round((case["hmax"] - case["hmin"]) / case["dr"]))


def drive_ts(fn, case, spare_frames=1):
    """fn(core, surr, result, boxl, t, hmin, hmax, invdx) for every frame -> result (frames + spare, n1, nbins)."""
    T = case["frames"]
    result = np.zeros((T + spare_frames, case["n1"], ts_nbins(case)))
    for t in range(T):
        core, surr = build_ts_inputs(case, t)
        fn(core, surr, result, case["box"], t, case["hmin"], case["hmax"], 1.0 / case["dr"])
    return result


# N1 + N4 atom frames: float32 atom positions -> astype(float64) -> helpers.comByResidue / dipByResidue -> RDF.calcFrame,
# all three steps by the UNMODIFIED reference modules (tests/golden/make_golden.py); fixture atoms.npz holds the raw sums.
ATOM_CASES = {
    "atoms_water_self": dict(n1=900, apr1=("const", 3), same=True, L=33.0, modes=["all"], hmax=15.0, frames=3, seed=31),
    "atoms_ions_cross": dict(n1=160, apr1=("range", 2, 12), n2=210, apr2=("range", 2, 7), same=False, L=38.0,
                             modes=["000", "110", "101", "011", "220"], hmax=19.0, frames=2, seed=32, npt=True),
    "atoms_solute_water": dict(n1=1, apr1=("const", 24), n2=700, apr2=("const", 3), same=False, L=30.0, modes=["all"],
                               hmax=15.0, frames=4, seed=33),
}


def build_atoms(case):
    """-> dict(topo1=(masses, apr, rfa, charges), topo2=... or None, pos1 (frames, na1, 3) float32, pos2 or None,
    boxes (frames, 3) or None)."""
    rng = np.random.default_rng(700001 * case["seed"])
    nf, L = case["frames"], case["L"]

    def topo(n, kind):
        apr = np.full(n, kind[1], dtype=np.int32) if kind[0] == "const" else rng.integers(kind[1], kind[2] + 1, n).astype(np.int32)
        rfa = np.concatenate([[0], np.cumsum(apr)[:-1]]).astype(np.int32)
        na = int(apr.sum())
        return (np.ascontiguousarray(1.0 + 15.0 * rng.random(na)), apr, rfa, np.ascontiguousarray(rng.standard_normal(na)))

    def frames(tp):
        apr = tp[1]
        centre = np.repeat(rng.random((nf, apr.shape[0], 3)) * L, apr, axis=1)
        return np.ascontiguousarray((centre + 0.7 * rng.standard_normal(centre.shape)).astype(np.float32))

    t1 = topo(case["n1"], case["apr1"])
    p1 = frames(t1)
    t2 = p2 = None
    if not case["same"]:
        t2 = topo(case["n2"], case["apr2"])
        p2 = frames(t2)
    boxes = L * (1.0 + 1e-3 * rng.standard_normal((nf, 3))) if case.get("npt") else None
    return dict(topo1=t1, topo2=t2, pos1=p1, pos2=p2, boxes=boxes)


# Topology side (newanalysis.psf / newanalysis.functions): synthetic PSF files; fixture topology.npz holds what the
# UNMODIFIED reference functions atomsPerResidue / residueFirstAtom (py_functions.py:64-89) return for the selections.
PSF_CASES = {
    "solvated": dict(ext=False, seed=51, blocks=[("SOLU", "MQS0", 1, 24), ("WAT", "SWM4", 300, 5)]),
    "ionic":    dict(ext=True, seed=52, blocks=[("CAT", "C2MI", 40, 19), ("AN", "OTF", 40, 8), ("CAT", "C2MI", 3, 19)]),
    "ragged":   dict(ext=False, seed=53, blocks=[("A", "LIG", 7, (1, 9)), ("A", "LIG2", 5, (2, 4)), ("B", "LIG", 7, (1, 9))],
                     same_resid_across_segments=True),
}
PSF_SELECTIONS = {
    "solvated": [dict(resname="SWM4"), dict(resname="MQS0"), dict()],
    "ionic":    [dict(resname="C2MI"), dict(resname="OTF"), dict(segid="CAT"), dict(resname="C2MI OTF")],
    "ragged":   [dict(resname="LIG"), dict(segid="B"), dict(), dict(resname="LIG2")],
}


def build_psf(case):
    """-> (text of a PSF file, dict of per-atom arrays that were written)."""
    rng = np.random.default_rng(800001 * case["seed"])
    segs, resids, resnames, names, charges, masses = [], [], [], [], [], []
    for seg, resname, nres, na in case["blocks"]:
        first = 1 if case.get("same_resid_across_segments") else (int(resids[-1]) + 1 if resids else 1)
        for r in range(nres):
            n = na if isinstance(na, int) else int(rng.integers(na[0], na[1] + 1))
            for a in range(n):
                segs.append(seg); resids.append(str(first + r)); resnames.append(resname); names.append("A%d" % (a + 1))
                charges.append(round(float(rng.standard_normal()), 6)); masses.append(round(1.0 + 15.0 * float(rng.random()), 4))
    n = len(segs)
    ext = case["ext"]
    out = ["PSF EXT CMAP" if ext else "PSF CMAP", "", "%8d !NTITLE" % 2, "* synthetic topology (tests/golden/cases.py)", "* seed %d" % case["seed"], "",
           ("%10d !NATOM" if ext else "%8d !NATOM") % n]
    for k in range(n):
        if ext:
            out.append("%10d %-8s %-8s %-8s %-8s %-6s %14.6f%14.4f%8d" % (k + 1, segs[k], resids[k], resnames[k], names[k], "T%d" % (k % 7), charges[k], masses[k], 0))
        else:
            out.append("%8d %-4s %-4s %-4s %-4s %4d %14.6f%14.4f%8d" % (k + 1, segs[k], resids[k], resnames[k], names[k], k % 90, charges[k], masses[k], 0))
    out += ["", ("%10d !NBOND: bonds" if ext else "%8d !NBOND: bonds") % 0, "", ""]
    return "\n".join(out), dict(segids=np.array(segs), resids=np.array(resids), resnames=np.array(resnames), names=np.array(names),
                                 charges=np.array(charges), masses=np.array(masses))

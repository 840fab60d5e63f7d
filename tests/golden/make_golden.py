#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference module (oracle/_ref).

Run in the build container (needs oracle/_ref, i.e. /root/reference at build time):
    python oracle/build_ref.py && python tests/golden/make_golden.py
The GPU box never runs this; it only reads the committed fixtures.

Per case the fixture holds
  raw    sum_i histogram[k][i][b]  (un-normalised, what RDF._addHisto adds; gfunction.pyx:466-480)
  out    histogram_out after RDF.write()  (normalised; gfunction.pyx:483-550)
  dat    the text of every written <prefix>_gXYZ.dat, keyed by mode
  digest sha256 of the regenerated inputs (guards against generator drift)
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)

import cases as C          # noqa: E402
import oracle as O         # noqa: E402


def main():
    ref = O.load_ref()
    if ref is None:
        sys.exit("oracle/_ref is not built; run python oracle/build_ref.py first")
    for name, case in C.CASES.items():
        rdf = C.drive(C.make_rdf(ref.RDF, case), case)
        raw = np.zeros(7 * int(rdf.histo_n))
        rdf._addHisto(raw, rdf.histogram)
        with tempfile.TemporaryDirectory() as tmp:
            rdf.write(os.path.join(tmp, "g"))
            dat = {m: open(os.path.join(tmp, "g_g%s.dat" % m)).read()
                   for m in O.MODE_ORDER if int(rdf.mode_sel) & O.MODE_BITS[m]}
        meta = dict(histo_n=int(rdf.histo_n), mode_sel=int(rdf.mode_sel), ctr=int(rdf.ctr),
                    oneistwo=bool(rdf.oneistwo), histo_min=float(rdf.histo_min),
                    histo_max=float(rdf.histo_max), histo_dr=float(rdf.histo_dr),
                    histo_inv_dr=float(rdf.histo_inv_dr), volume=float(rdf.volume),
                    digest=C.input_digest(case), dat=dat)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), raw=raw, out=rdf.histogram_out,
                            meta=np.array(json.dumps(meta)))
        print("%-20s histo_n=%d in-range(g000 raw)=%.0f" % (name, meta["histo_n"], raw[:meta["histo_n"]].sum()))

    # micro cases: flat index hit by a single pair (quirks Q2/Q3)
    micro = []
    for y, _expect in C.MICRO:
        r = ref.RDF(["000"], 0.0, 15.0, 0.05, 1, 2, 1, 2, 40.0)
        c1 = np.zeros((1, 3))
        c2 = np.array([y, (30.0, 30.0, 30.0)], dtype=np.float64)
        # allocate one spare row so the Q3 spill (index 300) stays inside the array we can inspect
        r.histogram = np.zeros(2 * int(r.histo_n) * 7)
        r.calcFrame(c1, c2)
        nz = np.flatnonzero(r.histogram)
        micro.append(int(nz[0]) if nz.size else -1)
    np.savez_compressed(os.path.join(HERE, "micro.npz"), idx=np.array(micro))
    print("micro", micro)

    # N1 producers: comByResidue / dipByResidue / velcomByResidue of the unmodified reference helpers module
    H = O.load_ref_helpers()
    if H is None:
        sys.exit("oracle/_ref/helpers*.so is not built; run python oracle/build_ref.py first")
    out = {}
    for name in sorted(C.RESIDUE_CASES):
        coor, masses, charges, apr, rfa = C.build_residues(C.RESIDUE_CASES[name])
        nres = apr.shape[0]
        com = H.comByResidue(coor, masses, nres, apr, rfa)
        dip = H.dipByResidue(coor, charges, masses, nres, apr, rfa, com)
        vel = H.velcomByResidue(coor * 0.37, masses, nres, apr, rfa)
        out[name + "_com"] = com; out[name + "_dip"] = dip; out[name + "_vel"] = vel
        print("residues %-10s nres=%d natoms=%d" % (name, nres, coor.shape[0]))
    np.savez_compressed(os.path.join(HERE, "residues.npz"), **out)

    # N1 + N4 atom frames: the whole script-side chain with the unmodified reference modules
    out = {}
    for name in sorted(C.ATOM_CASES):
        case = C.ATOM_CASES[name]
        a = C.build_atoms(case)
        n1 = case["n1"]; n2 = n1 if case["same"] else case["n2"]
        rdf = ref.RDF(case["modes"], 0.0, case["hmax"], 0.05, n1, n2, n1, n2, case["L"])
        for f in range(case["frames"]):
            if a["boxes"] is not None:
                rdf.update_box(*[float(v) for v in a["boxes"][f]])
            sel = []
            for tp, pos in ((a["topo1"], a["pos1"]), (a["topo2"], a["pos2"])):
                if tp is None:
                    sel.append(sel[0]); continue
                masses, apr, rfa, charges = tp
                coor = np.ascontiguousarray(pos[f], dtype="double")          # the notebook's np.ascontiguousarray(sel.positions, dtype='double')
                com = H.comByResidue(coor, masses, apr.shape[0], apr, rfa)
                dip = H.dipByResidue(coor, charges, masses, apr.shape[0], apr, rfa, com)
                sel.append((com, dip))
            rdf.calcFrame(sel[0][0], sel[1][0], sel[0][1], sel[1][1])
        raw = np.zeros(7 * int(rdf.histo_n))
        rdf._addHisto(raw, rdf.histogram)
        out[name + "_raw"] = raw
        out[name + "_volumes"] = np.array(rdf.volume_list)
        out[name + "_oneistwo"] = np.array(bool(rdf.oneistwo))
        print("atoms %-20s histo_n=%d in-range(g000 raw)=%.0f oneistwo=%s" % (name, rdf.histo_n, raw[:int(rdf.histo_n)].sum(), rdf.oneistwo))
    np.savez_compressed(os.path.join(HERE, "atoms.npz"), **out)


def main_shells():
    """N2: RDF_voronoi and calcHistoTSVoro fixtures from the unmodified reference module."""
    ref = O.load_ref()
    if ref is None:
        sys.exit("oracle/_ref is not built; run python oracle/build_ref.py first")
    for name, case in C.VORO_CASES.items():
        rdf = C.drive_voro(C.make_voro(ref.RDF_voronoi, case), case)
        ns, hn = int(rdf.nshells), int(rdf.histo_n)
        raw = np.zeros(7 * hn * ns)
        rdf._addHisto(raw, rdf.histogram)
        with tempfile.TemporaryDirectory() as tmp:
            rdf.write(os.path.join(tmp, "g"))
            dat = {"%d_%s" % (sh + 1, m): open(os.path.join(tmp, "g_shell%d_g%s.dat" % (sh + 1, m))).read()
                   for sh in range(ns) for m in O.MODE_ORDER if int(rdf.mode_sel) & O.MODE_BITS[m]}
        meta = dict(histo_n=hn, nshells=ns, mode_sel=int(rdf.mode_sel), ctr=int(rdf.ctr), oneistwo=bool(rdf.oneistwo),
                    volume=float(rdf.volume), digest=C.voro_digest(case), dat=dat)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), raw=raw, out=rdf.histogram_out, meta=np.array(json.dumps(meta)))
        print("%-16s histo_n=%d nshells=%d g000 per shell=%s" % (name, hn, ns, raw[:hn * ns].reshape(hn, ns).sum(0)))
    out = {}
    for name, case in C.TSVORO_CASES.items():
        out[name] = C.drive_tsvoro(ref.calcHistoTSVoro, case)
        print("%-16s counts per shell=%s" % (name, out[name].sum(axis=(0, 1, 3))))
    np.savez_compressed(os.path.join(HERE, "tsvoro.npz"), **out)

    # N3: calcHistoTS (gfunction.pyx:168-192) by the unmodified reference function, incl. the Q3 spill of the last core row
    out = {}
    for name, case in C.TS_CASES.items():
        out[name] = C.drive_ts(ref.calcHistoTS, case)
        T = case["frames"]
        print("%-16s counts per frame=%s spare frame holds %.0f" % (name, out[name][:T].sum(axis=(1, 2)), out[name][T:].sum()))
    np.savez_compressed(os.path.join(HERE, "ts.npz"), **out)

    # topology helpers: the reference's own atomsPerResidue / residueFirstAtom (py_functions.py:64-89, pure Python) run on the
    # AtomGroup-like selections of newanalysis.psf
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_py_functions", "/root/reference/newanalysis/functions/py_functions.py")
    RF = importlib.util.module_from_spec(spec); spec.loader.exec_module(RF)
    sys.path.insert(0, os.path.join(ROOT, "mdy-newanalysis-package_b200", "newanalysis"))
    import psf as PSFMOD                      # the module file alone: the package __init__ would load the CUDA extension

    class RefView(object):                    # what the reference functions touch: n_atoms, n_residues, atoms[i].resid / .resname
        def __init__(self, sel):
            self.n_atoms, self.n_residues = sel.n_atoms, sel.n_residues
            self.atoms = [type("A", (), dict(resid=sel.resids[i], resname=sel.resnames[i])) for i in range(sel.n_atoms)]
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name in sorted(C.PSF_CASES):
            text, _ = C.build_psf(C.PSF_CASES[name])
            path = os.path.join(tmp, name + ".psf")
            open(path, "w").write(text)
            top = PSFMOD.PSF(path)
            for k, crit in enumerate(C.PSF_SELECTIONS[name]):
                sel = top.select(**crit)
                rv = RefView(sel)
                out["%s_%d_apr" % (name, k)] = RF.atomsPerResidue(rv)
                out["%s_%d_rfa" % (name, k)] = RF.residueFirstAtom(rv)
                out["%s_%d_n" % (name, k)] = np.array([sel.n_atoms, sel.n_residues])
                print("topology %-10s sel %d: %d atoms, %d residues" % (name, k, sel.n_atoms, sel.n_residues))
    np.savez_compressed(os.path.join(HERE, "topology.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "shells":
        main_shells()
    else:
        main()
        main_shells()

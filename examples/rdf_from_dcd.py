#!/usr/bin/env python3
"""The reference's RDF notebook (docs/notebooks/rdf.ipynb, cells 10-18) with the trajectory handed to the library:
one solute molecule x all waters, all seven g-functions, straight from a DCD file.

The notebook needs MDAnalysis for two things: the topology (masses, charges, residues -> atoms per residue / first atom)
and the frame loop.  The frame loop, the centres of mass, the dipoles and the pair loop all run behind libgfn.so here;
the topology arrays come from wherever you have them (MDAnalysis, a PSF parser, ...).  Without arguments the script
writes a small synthetic solvated system to a temporary DCD file first so that it runs anywhere with a B200.

    python examples/rdf_from_dcd.py [trajectory.dcd topology.psf SOLUTE_RESNAME SOLVENT_RESNAME]
e.g. the notebook's data:  python examples/rdf_from_dcd.py 1mq_swm4.dcd mqs0_swm4_1000.psf MQS0 SWM4
"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mdy-newanalysis-package_b200"))

from newanalysis.gfunction import RDF, DCDFile  # noqa: E402
from newanalysis.functions import atomsPerResidue, residueFirstAtom  # noqa: E402
from newanalysis.psf import PSF  # noqa: E402


def residues(resid):
    """atoms per residue / first atom of each residue from a per-atom residue id (functions.atomsPerResidue /
    residueFirstAtom of the reference, for contiguous residues)."""
    first = np.flatnonzero(np.r_[True, resid[1:] != resid[:-1]])
    apr = np.diff(np.r_[first, resid.shape[0]])
    return apr.astype(np.int32), first.astype(np.int32)


def synthetic(path, nwat=4000, nsol=24, L=49.3, nframes=50, seed=5):
    from gfn_b200 import synth                   # synthetic-input helpers: writes the demo file
    rng = np.random.default_rng(seed)
    wat = rng.random((nframes, nwat, 1, 3)) * L + 0.6 * rng.standard_normal((nframes, nwat, 3, 3))
    sol = L / 2 + 2.0 * rng.standard_normal((nframes, nsol, 3))
    pos = np.concatenate([sol, wat.reshape(nframes, 3 * nwat, 3)], axis=1).astype(np.float32)
    cells = np.tile([L, L, L, 90.0, 90.0, 90.0], (nframes, 1))
    synth.write_dcd(path, pos, cells)
    masses = np.r_[1.0 + 12.0 * rng.random(nsol), np.tile([15.999, 1.008, 1.008], nwat)]
    charges = np.r_[rng.standard_normal(nsol), np.tile([-0.8476, 0.4238, 0.4238], nwat)]
    resid = np.r_[np.zeros(nsol, dtype=np.int64), 1 + np.repeat(np.arange(nwat), 3)]
    return masses, charges, resid, 0


def main():
    tmp = None
    if len(sys.argv) == 5:
        path = sys.argv[1]
        top = PSF(sys.argv[2])
        s_sol, s_wat = top.select(resname=sys.argv[3]), top.select(resname=sys.argv[4])
        sel_solute, sel_wat = s_sol.indices, s_wat.indices
        apr_s, rfa_s, apr_w, rfa_w = atomsPerResidue(s_sol), residueFirstAtom(s_sol), atomsPerResidue(s_wat), residueFirstAtom(s_wat)
        m_s, q_s, m_w, q_w = s_sol.masses, s_sol.charges, s_wat.masses, s_wat.charges
    else:
        tmp = tempfile.mkdtemp(prefix="gfn_example_")
        path = os.path.join(tmp, "demo.dcd")
        masses, charges, resid, solute = synthetic(path)
        sel_solute = np.flatnonzero(resid == solute).astype(np.int32)
        sel_wat = np.flatnonzero(resid != solute).astype(np.int32)
        apr_s, rfa_s = residues(resid[sel_solute])
        apr_w, rfa_w = residues(resid[sel_wat])
        m_s, q_s, m_w, q_w = masses[sel_solute], charges[sel_solute], masses[sel_wat], charges[sel_wat]

    traj = DCDFile(path)
    _, cell = traj.read(0)
    boxl = np.float64(round(float(cell[0]), 4))                   # the notebook: round(u.coord.dimensions[0], 4)
    nsolute, nwat = apr_s.shape[0], apr_w.shape[0]

    sol_wat = RDF(["all"], 0.0, 20.0, 0.05, nsolute, nwat, nsolute, nwat, boxl)
    sol_wat.attach_topology(m_s, apr_s, rfa_s, q_s, m_w, apr_w, rfa_w, q_w)
    t0 = time.time()
    n = sol_wat.calcTrajectory(traj, sel_solute, sel_wat, step=1)
    out = os.path.join(tmp or ".", "rdf")
    sol_wat.write(out)
    print("%d frames of %d atoms in %.3f s -> %s_g000.dat ... (%d solute x %d water sites)"
          % (n, traj.n_atoms, time.time() - t0, out, nsolute, nwat))
    g000 = np.loadtxt(out + "_g000.dat")
    print("g000 at large r: %.3f (-> 1 for a uniform solvent)" % g000[-40:, 1].mean())


if __name__ == "__main__":
    main()

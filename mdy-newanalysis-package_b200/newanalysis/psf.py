"""newanalysis.psf (B200 backend): CHARMM / NAMD / X-PLOR PSF topologies without MDAnalysis.

The reference's scripts take masses, charges and the residue structure from `MDAnalysis.Universe(psf, dcd)`
(docs/notebooks/rdf.ipynb cells 10-16: `u.select_atoms("resname SWM4")`, `sel.charges`, `sel.masses`,
`atomsPerResidue(sel)`, `residueFirstAtom(sel)`).  Together with `newanalysis.gfunction.DCDFile` this module lets the
same analysis run from the two files alone:

    top = PSF("mqs0_swm4_1000.psf")
    sel = top.select(resname="SWM4");  sel_solute = top.select(resname="MQS0")
    rdf.attach_topology(sel_solute.masses, atomsPerResidue(sel_solute), residueFirstAtom(sel_solute), sel_solute.charges,
                        sel.masses, atomsPerResidue(sel), residueFirstAtom(sel), sel.charges)
    rdf.calcTrajectory(DCDFile("1mq_swm4.dcd"), sel_solute.indices, sel.indices, step=20)

What is read is the !NATOM section (standard and EXT column widths, CHARMM and X-PLOR atom-type columns, Drude/CHEQ
extra columns ignored): atom id, segment, residue id, residue name, atom name, atom type, charge, mass.  Bonds, angles
etc. are not needed by the path.  MDAnalysis is a third-party dependency of the reference that is not in the image: this
parser follows the published PSF layout and is pinned only by its own writer-side tests (tests/test_topology_cpu.py).
"""
import numpy


class Selection(object):
    """AtomGroup-like view of some atoms of a PSF: indices (into the trajectory's atoms), masses, charges, resids,
    resnames, segids, names, n_atoms, n_residues (runs of equal (segid, resid, resname)); `positions` is whatever you
    assign (e.g. DCDFile.read(frame, sel.indices)[0])."""

    def __init__(self, top, indices):
        self.top = top
        self.indices = numpy.ascontiguousarray(indices, dtype=numpy.int32)
        self.masses = numpy.ascontiguousarray(top.masses[self.indices])
        self.charges = numpy.ascontiguousarray(top.charges[self.indices])
        self.resids = top.resids[self.indices]
        self.resnames = top.resnames[self.indices]
        self.segids = top.segids[self.indices]
        self.names = top.names[self.indices]
        self.n_atoms = int(self.indices.shape[0])
        if self.n_atoms:
            new = numpy.ones(self.n_atoms, dtype=bool)
            new[1:] = ((self.resids[1:] != self.resids[:-1]) | (self.resnames[1:] != self.resnames[:-1])
                       | (self.segids[1:] != self.segids[:-1]))
            self.n_residues = int(new.sum())
        else:
            self.n_residues = 0
        self.positions = None

    def __len__(self):
        return self.n_atoms


class PSF(object):
    def __init__(self, path):
        with open(path, "r") as fh:
            lines = fh.read().splitlines()
        if not lines or not lines[0].startswith("PSF"):
            raise ValueError("%s is not a PSF file (first line does not start with PSF)" % path)
        self.flags = lines[0].split()[1:]
        start = None
        for k, ln in enumerate(lines):
            if "!NATOM" in ln:
                start = k
                break
        if start is None:
            raise ValueError("%s: no !NATOM section" % path)
        try:
            natom = int(lines[start].split()[0])
        except (ValueError, IndexError):
            raise ValueError("%s: cannot read the atom count from %r" % (path, lines[start]))
        body = lines[start + 1:start + 1 + natom]
        if len(body) < natom:
            raise ValueError("%s: !NATOM announces %d atoms, the file holds %d" % (path, natom, len(body)))
        ids = numpy.empty(natom, dtype=numpy.int64)
        segids, resids, resnames, names, types = [], [], [], [], []
        charges = numpy.empty(natom, dtype=numpy.float64)
        masses = numpy.empty(natom, dtype=numpy.float64)
        for k, ln in enumerate(body):
            t = ln.split()
            if len(t) < 8:
                raise ValueError("%s: atom line %d has %d fields, expected at least 8: %r" % (path, k + 1, len(t), ln))
            try:
                ids[k] = int(t[0])
                charges[k] = float(t[6])
                masses[k] = float(t[7])
            except ValueError:
                raise ValueError("%s: atom line %d: id / charge / mass unreadable: %r" % (path, k + 1, ln))
            segids.append(t[1]); resids.append(t[2]); resnames.append(t[3]); names.append(t[4]); types.append(t[5])
        self.path = str(path)
        self.n_atoms = natom
        self.ids = ids
        self.segids = numpy.array(segids)
        self.resids = numpy.array(resids)            # strings: residue ids may carry insertion codes
        self.resnames = numpy.array(resnames)
        self.names = numpy.array(names)
        self.types = numpy.array(types)
        self.charges = charges
        self.masses = masses

    def select(self, resname=None, segid=None, name=None, mask=None):
        """Atoms matching all given criteria (each a string or a collection of strings; mask: boolean array), in file
        order - the subset of MDAnalysis' `select_atoms("resname X and segid Y and name Z")` the reference's scripts use."""
        keep = numpy.ones(self.n_atoms, dtype=bool)
        for values, column in ((resname, self.resnames), (segid, self.segids), (name, self.names)):
            if values is not None:
                if isinstance(values, str):
                    values = values.split()
                keep &= numpy.isin(column, list(values))
        if mask is not None:
            keep &= numpy.asarray(mask, dtype=bool)
        return Selection(self, numpy.flatnonzero(keep))

    @property
    def atoms(self):
        return Selection(self, numpy.arange(self.n_atoms))

"""newanalysis.functions (B200 backend): the callers on the input side of the g-function path.

Only what sits between a selection and RDF.calcFrame is mirrored from the reference's functions module
(/root/reference/newanalysis/functions/py_functions.py; the notebook imports them in docs/notebooks/rdf.ipynb cell 7):

  atomsPerResidue(sel)            py_functions.py:64-79     same result, vectorised
  residueFirstAtom(sel)           py_functions.py:82-89
  centerOfMassByResidue(sel, **)  py_functions.py:91-114    -> helpers.comByResidue on the GPU (bit-identical arrays)
  dipoleMomentByResidue(sel, **)  py_functions.py:120-150   -> helpers.dipByResidue on the GPU

`sel` is anything AtomGroup-like: n_atoms, n_residues, per-atom resids / resnames (or an `atoms` sequence with
.resid / .resname like MDAnalysis), masses, charges, positions.  `newanalysis.psf.PSF(...).select(...)` gives such
objects without MDAnalysis.  The rest of the reference module (~40 unrelated functions) is out of scope.
"""
import numpy


def _resids_resnames(sel):
    if hasattr(sel, "resids") and hasattr(sel, "resnames") and len(sel.resids) == sel.n_atoms:
        return numpy.asarray(sel.resids), numpy.asarray(sel.resnames)
    atoms = sel.atoms
    return (numpy.array([atoms[i].resid for i in range(sel.n_atoms)]),
            numpy.array([atoms[i].resname for i in range(sel.n_atoms)]))


def atomsPerResidue(sel):
    """Atoms per residue: a new residue starts where resid or resname differs from the previous atom
    (py_functions.py:71-78).  int32 array of length sel.n_residues."""
    apr = numpy.zeros(sel.n_residues, dtype=numpy.int32)
    if sel.n_atoms == 0:
        return apr
    resid, resname = _resids_resnames(sel)
    new = numpy.ones(sel.n_atoms, dtype=bool)
    new[1:] = (resid[1:] != resid[:-1]) | (resname[1:] != resname[:-1])
    idx = numpy.cumsum(new) - 1
    if idx[-1] >= apr.shape[0]:
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (idx[-1], apr.shape[0]))   # as the reference's apr[idx]+=1
    numpy.add.at(apr, idx, 1)
    return apr


def residueFirstAtom(sel):
    """Index (within the selection) of the first atom of every residue (py_functions.py:82-89)."""
    apr = atomsPerResidue(sel)
    rfa = numpy.zeros(sel.n_residues, dtype=numpy.int32)
    if apr.shape[0] > 1:
        rfa[1:] = numpy.cumsum(apr[:-1])
    return rfa


def centerOfMassByResidue(sel, **kwargs):
    """Returns an array of the centers of mass of all residues in the current AtomGroup (py_functions.py:91-114)."""
    from newanalysis.helpers import comByResidue
    if "coor" not in kwargs:
        coor = sel.positions.astype('float64')
    else:
        coor = kwargs["coor"]
    if "masses" not in kwargs:
        masses = numpy.ascontiguousarray(sel.masses)
    else:
        masses = kwargs["masses"]
    if "apr" not in kwargs:
        print("Warning: apr is now calculated each timestep (very slow)! Give apr as an argument to speed up code")
        apr = atomsPerResidue(sel)
    else:
        apr = kwargs["apr"]
    if "rfa" not in kwargs:
        print("Warning: rfa is now calculated each timestep (very slow)! Give rfa as an argument to speed up code")
        rfa = residueFirstAtom(sel)
    else:
        rfa = kwargs["rfa"]
    numres = sel.n_residues
    return comByResidue(coor, masses, numres, apr, rfa)


def dipoleMomentByResidue(sel, **kwargs):
    """Returns an array of the dipole moments of all residues in the current AtomGroup (py_functions.py:120-150)."""
    from newanalysis.helpers import dipByResidue
    if "charges" not in kwargs:
        charges = numpy.ascontiguousarray(sel.charges)
    else:
        charges = kwargs["charges"]
    if "masses" not in kwargs:
        masses = numpy.ascontiguousarray(sel.masses)
    else:
        masses = kwargs["masses"]
    if "coor" not in kwargs:
        coor = sel.positions.astype('float64')
    else:
        coor = kwargs["coor"]
    if "apr" not in kwargs:
        print("Warning: apr is now calculated each timestep (very slow)! Give apr as an argument to speed up code")
        apr = atomsPerResidue(sel)
    else:
        apr = kwargs["apr"]
    if "rfa" not in kwargs:
        print("Warning: rfa is now calculated each timestep (very slow)! Give rfa as an argument to speed up code")
        rfa = residueFirstAtom(sel)
    else:
        rfa = kwargs["rfa"]
    if "com" not in kwargs:
        # the reference calls centerOfMassByResidue(sel, masses=masses, coor=coor) here, which recomputes apr / rfa
        # (and prints its warnings) even when they were passed; the arrays are the same
        com = centerOfMassByResidue(sel, masses=masses, coor=coor, apr=apr, rfa=rfa)
    else:
        com = kwargs["com"]
    numres = sel.n_residues
    return dipByResidue(coor, charges, masses, numres, apr, rfa, com)

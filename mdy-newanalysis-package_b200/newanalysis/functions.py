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


_SLOW = "Warning: %s is now calculated each timestep (very slow)! Give %s as an argument to speed up code"


def _inputs(sel, kwargs, want):
    """Keyword arguments of the reference wrappers with their defaults taken from the selection: coor (float64 copy of
    sel.positions), masses / charges (contiguous), apr / rfa (recomputed, with the reference's warning, when absent)."""
    out = {}
    for key in want:
        if key in kwargs:
            out[key] = kwargs[key]
        elif key == "coor":
            out[key] = sel.positions.astype('float64')
        elif key in ("masses", "charges"):
            out[key] = numpy.ascontiguousarray(getattr(sel, key))
        else:
            print(_SLOW % (key, key))
            out[key] = atomsPerResidue(sel) if key == "apr" else residueFirstAtom(sel)
    return out


def centerOfMassByResidue(sel, **kwargs):
    """Centres of mass of all residues of the AtomGroup-like `sel` (py_functions.py:91-114): optional keywords coor,
    masses, apr, rfa; computed by helpers.comByResidue on the GPU, bit-identical to the reference's arrays."""
    from newanalysis.helpers import comByResidue
    a = _inputs(sel, kwargs, ("coor", "masses", "apr", "rfa"))
    return comByResidue(a["coor"], a["masses"], sel.n_residues, a["apr"], a["rfa"])


def dipoleMomentByResidue(sel, **kwargs):
    """Dipole moments of all residues relative to `com` (default: their centres of mass) (py_functions.py:120-150):
    optional keywords charges, masses, coor, com, apr, rfa; helpers.dipByResidue on the GPU."""
    from newanalysis.helpers import dipByResidue
    a = _inputs(sel, kwargs, ("charges", "masses", "coor", "apr", "rfa"))
    # (the reference computes a missing com with centerOfMassByResidue(sel, masses=masses, coor=coor), which derives
    # apr / rfa once more and prints its warnings again; the arrays are the same)
    com = kwargs["com"] if "com" in kwargs else centerOfMassByResidue(sel, **{k: a[k] for k in ("masses", "coor", "apr", "rfa")})
    return dipByResidue(a["coor"], a["charges"], a["masses"], sel.n_residues, a["apr"], a["rfa"], com)

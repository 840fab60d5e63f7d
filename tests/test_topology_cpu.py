"""CPU tests of the topology side of the path's callers: newanalysis.psf (PSF files without MDAnalysis) and
newanalysis.functions.atomsPerResidue / residueFirstAtom / centerOfMassByResidue / dipoleMomentByResidue against
fixtures produced by the reference's own functions (tests/golden/make_golden.py, py_functions.py:64-150)."""
import os

import numpy as np
import pytest

import cases as C
import oracle as O


def _mods():
    from newanalysis import psf, functions
    return psf, functions


@pytest.mark.parametrize("name", sorted(C.PSF_CASES))
def test_psf_reads_what_was_written(tmp_path, name):
    psf, _ = _mods()
    text, cols = C.build_psf(C.PSF_CASES[name])
    path = tmp_path / (name + ".psf")
    path.write_text(text)
    top = psf.PSF(str(path))
    assert top.n_atoms == cols["masses"].shape[0]
    for key in ("segids", "resids", "resnames", "names"):
        assert np.array_equal(getattr(top, key), cols[key]), key
    assert np.array_equal(top.charges, cols["charges"]) and np.array_equal(top.masses, cols["masses"])
    assert np.array_equal(top.ids, np.arange(1, top.n_atoms + 1))
    assert ("EXT" in top.flags) == C.PSF_CASES[name]["ext"]


@pytest.mark.parametrize("name", sorted(C.PSF_CASES))
def test_apr_rfa_match_reference_functions(tmp_path, golden_dir, name):
    psf, F = _mods()
    z = np.load(os.path.join(golden_dir, "topology.npz"))
    text, cols = C.build_psf(C.PSF_CASES[name])
    path = tmp_path / (name + ".psf")
    path.write_text(text)
    top = psf.PSF(str(path))
    for k, crit in enumerate(C.PSF_SELECTIONS[name]):
        sel = top.select(**crit)
        assert [sel.n_atoms, sel.n_residues] == list(z["%s_%d_n" % (name, k)])
        apr, rfa = F.atomsPerResidue(sel), F.residueFirstAtom(sel)
        assert apr.dtype == np.int32 and rfa.dtype == np.int32
        assert np.array_equal(apr, z["%s_%d_apr" % (name, k)])
        assert np.array_equal(rfa, z["%s_%d_rfa" % (name, k)])
        assert np.array_equal(sel.masses, cols["masses"][sel.indices]) and np.array_equal(sel.charges, cols["charges"][sel.indices])


def test_mdanalysis_like_atom_objects_are_accepted():
    """The reference's functions walk sel.atoms[i].resid / .resname (MDAnalysis AtomGroup); so may ours."""
    _, F = _mods()

    class Atom(object):
        def __init__(self, resid, resname):
            self.resid, self.resname = resid, resname

    class Group(object):
        atoms = [Atom(1, "A"), Atom(1, "A"), Atom(1, "B"), Atom(2, "B"), Atom(2, "B"), Atom(2, "B")]
        n_atoms, n_residues = 6, 3
    assert list(F.atomsPerResidue(Group())) == [2, 1, 3]
    assert list(F.residueFirstAtom(Group())) == [0, 2, 3]
    Group.n_residues = 2
    with pytest.raises(IndexError):
        F.atomsPerResidue(Group())


def test_psf_errors(tmp_path):
    psf, _ = _mods()
    p = tmp_path / "x.psf"
    p.write_text("not a topology\n")
    with pytest.raises(ValueError, match="not a PSF"):
        psf.PSF(str(p))
    p.write_text("PSF\n\n       1 !NTITLE\n* t\n\n")
    with pytest.raises(ValueError, match="NATOM"):
        psf.PSF(str(p))
    p.write_text("PSF\n\n       1 !NTITLE\n* t\n\n       3 !NATOM\n       1 A    1    R    N    1   0.100000       14.0070           0\n")
    with pytest.raises(ValueError, match="announces 3 atoms"):
        psf.PSF(str(p))
    p.write_text("PSF\n\n       1 !NTITLE\n* t\n\n       1 !NATOM\n       1 A    1    R    N    1   ******       14.0070           0\n")
    with pytest.raises(ValueError, match="unreadable"):
        psf.PSF(str(p))


def test_wrappers_pass_the_reference_arguments(tmp_path, monkeypatch):
    """centerOfMassByResidue / dipoleMomentByResidue (py_functions.py:91-150): keyword handling and the order of the
    arguments they hand to helpers.comByResidue / dipByResidue.  The GPU producers are replaced by the oracle's C
    restatements here (no GPU in this test); their own parity is tests/test_gpu_parity.py::test_residue_producers_bit_exact."""
    psf, F = _mods()
    import newanalysis.helpers as H
    monkeypatch.setattr(H, "comByResidue", lambda coor, masses, nres, apr, rfa: O.com_by_residue_c(coor, masses, apr[:nres], rfa[:nres]))
    monkeypatch.setattr(H, "dipByResidue", lambda coor, charges, masses, nres, apr, rfa, com: O.dip_by_residue_c(coor, charges, apr[:nres], rfa[:nres], com))
    text, cols = C.build_psf(C.PSF_CASES["solvated"])
    path = tmp_path / "s.psf"
    path.write_text(text)
    sel = psf.PSF(str(path)).select(resname="SWM4")
    rng = np.random.default_rng(5)
    sel.positions = (rng.random((sel.n_atoms, 3)) * 30).astype(np.float32)
    apr, rfa = F.atomsPerResidue(sel), F.residueFirstAtom(sel)
    coor = sel.positions.astype("float64")
    com_ref = O.com_by_residue_c(coor, sel.masses, apr, rfa)
    dip_ref = O.dip_by_residue_c(coor, sel.charges, apr, rfa, com_ref)
    assert np.array_equal(F.centerOfMassByResidue(sel, apr=apr, rfa=rfa), com_ref)
    assert np.array_equal(F.centerOfMassByResidue(sel, coor=coor, masses=sel.masses, apr=apr, rfa=rfa), com_ref)
    assert np.array_equal(F.centerOfMassByResidue(sel), com_ref)                           # apr / rfa computed (with the warning)
    assert np.array_equal(F.dipoleMomentByResidue(sel, apr=apr, rfa=rfa), dip_ref)
    assert np.array_equal(F.dipoleMomentByResidue(sel, coor=coor, charges=sel.charges, masses=sel.masses, com=com_ref, apr=apr, rfa=rfa), dip_ref)
    shifted = com_ref + 1.0
    assert np.array_equal(F.dipoleMomentByResidue(sel, com=shifted, apr=apr, rfa=rfa), O.dip_by_residue_c(coor, sel.charges, apr, rfa, shifted))

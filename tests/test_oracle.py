"""CPU tests: pin the oracle (C + numpy restatements, host-logic port) to the reference.

Sources of truth, strongest first:
  1. committed golden fixtures generated from the unmodified reference module (tests/golden/*.npz);
  2. the reference module itself when oracle/_ref is present (build container and GPU box both have
     the prebuilt .so; it is skipped only if the file is missing).
"""
import json
import os

import numpy as np
import pytest

import cases as C
import oracle as O


def load_golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return z["raw"], z["out"], json.loads(str(z["meta"]))


@pytest.mark.parametrize("name", sorted(C.CASES))
def test_inputs_reproducible(golden_dir, name):
    _, _, meta = load_golden(golden_dir, name)
    assert C.input_digest(C.CASES[name]) == meta["digest"]


@pytest.mark.parametrize("name", sorted(C.CASES))
def test_c_oracle_matches_golden(golden_dir, name, tmp_path):
    raw, out, meta = load_golden(golden_dir, name)
    rdf = C.drive(C.make_rdf(O.OracleRDF, C.CASES[name]), C.CASES[name])
    assert int(rdf.histo_n) == meta["histo_n"] and int(rdf.mode_sel) == meta["mode_sel"]
    assert float(rdf.histo_max) == meta["histo_max"] and float(rdf.histo_inv_dr) == meta["histo_inv_dr"]
    assert bool(rdf.oneistwo) == meta["oneistwo"] and rdf.ctr == meta["ctr"]
    assert np.array_equal(rdf.raw_sum(), raw)            # bit-exact, all 7 modes, reference summation order
    rdf.write(str(tmp_path / "g"))
    assert np.array_equal(rdf.histogram_out, out)
    for mode, text in meta["dat"].items():
        assert (tmp_path / ("g_g%s.dat" % mode)).read_text() == text


@pytest.mark.parametrize("name", ["rect_all", "self_odd", "ortho_npt", "hmin_pos", "unwrapped", "tiny_1x5",
                                  "beyond_half", "f32_bounds"])
def test_numpy_restatement_matches_golden(golden_dir, name):
    raw, _, _ = load_golden(golden_dir, name)
    case = C.CASES[name]
    rdf = C.make_rdf(O.OracleRDF, case)
    for f in range(case["frames"]):
        c1, c2, d1, d2, newbox = C.build_inputs(case, f)
        if newbox:
            rdf.update_box(*newbox)
        O.calc_histo_numpy(c1, c2, d1, d2, rdf.histogram, rdf.box_dim, rdf.histo_min, rdf.histo_max,
                           int(rdf.histo_n), rdf.histo_inv_dr, int(rdf.mode_sel))
    assert np.array_equal(rdf.raw_sum(), raw)


def test_micro_cases(golden_dir):
    idx = np.load(os.path.join(golden_dir, "micro.npz"))["idx"]
    for (y, expect), got in zip(C.MICRO, idx):
        assert got == (-1 if expect is None else expect)          # what the reference did
        hist = np.zeros(2 * 300 * 7)
        box = np.array([40.0, 40, 40, 20, 20, 20])
        O.calc_histo_c(np.zeros((1, 3)), np.array([y, (30.0, 30, 30)]), None, None, hist, box, 0.0, 15.0, 300, 20.0, 1)
        nz = np.flatnonzero(hist)
        assert (int(nz[0]) if nz.size else -1) == got


def test_rows_variant_equals_full(golden_dir):
    case = C.CASES["self_odd"]
    raw, _, meta = load_golden(golden_dir, "self_odd")
    rdf = C.make_rdf(O.OracleRDF, case)
    out = np.zeros(7 * meta["histo_n"])
    for f in range(case["frames"]):
        c1, c2, d1, d2, _ = C.build_inputs(case, f)
        rc, oob = O.calc_histo_rows_c(c1, c2, d1, d2, out, rdf.box_dim, rdf.histo_min, rdf.histo_max,
                                      meta["histo_n"], rdf.histo_inv_dr, meta["mode_sel"])
        assert rc == 0 and oob == 0
    n = meta["histo_n"]
    assert np.array_equal(out[:n], raw[:n])                                  # counts exact
    np.testing.assert_allclose(out, raw, rtol=0, atol=1e-12 * max(1.0, raw[:n].max()))


def test_zero_norm_dipole_raises():
    rdf = O.OracleRDF(["all"], 0.0, 15.0, 0.05, 4, 5, 4, 5, 40.0)
    rng = np.random.default_rng(0)
    c1, c2 = rng.random((4, 3)) * 5, rng.random((5, 3)) * 5
    d1, d2 = rng.standard_normal((4, 3)), rng.standard_normal((5, 3))
    d2[2] = 0.0
    with pytest.raises(ZeroDivisionError):
        rdf.calcFrame(c1, c2, d1, d2)
    assert rdf.ctr == 1                      # Q5: counter already incremented


def test_missing_dipoles_typeerror():
    rdf = O.OracleRDF(["110"], 0.0, 15.0, 0.05, 4, 5, 4, 5, 40.0)
    with pytest.raises(TypeError, match="fist selection"):
        rdf.calcFrame(np.zeros((4, 3)), np.zeros((5, 3)))


def test_exact_thresholds_equivalent_to_reference_test():
    """T_lo < d2 <= T_hi  <=>  not (dist<hmin or dist>hmax or floor(dist*invdx)==0), probed adversarially."""
    rng = np.random.default_rng(5)
    for hmin, hmax, dr in [(0.0, 15.0, 0.05), (0.0, 15.02, 0.05), (2.0, 12.0, 0.1), (0.0, 30.0, 0.05), (0.3, 9.99, 0.03333)]:
        invdx = float(np.float64(1.0 / np.round(np.float64(dr), 5)))
        t_lo, t_hi = O.exact_thresholds(hmin, hmax, invdx)
        hm, hx = float(np.float32(hmin)), float(np.float32(hmax))
        pts = [t_lo, t_hi, hm * hm, hx * hx, (1.0 / invdx) ** 2]
        vals = list(rng.random(200000) * hx * hx * 1.2)
        for p in pts:   # walk +-64 ulps around every edge
            x = y = p
            vals.append(p)
            for _ in range(64):
                x = np.nextafter(x, -np.inf); y = np.nextafter(y, np.inf)
                vals += [x, y]
        d2 = np.abs(np.array(vals, dtype=np.float64))
        dist = np.sqrt(d2)
        ref_ok = ~((dist < hm) | (dist > hx) | (np.floor(dist * invdx) == 0))
        thr_ok = (d2 > t_lo) & (d2 <= t_hi)
        assert np.array_equal(ref_ok, thr_ok)


def test_reference_module_agrees_when_present():
    ref = O.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    case = C.CASES["rect_all"]
    a = C.drive(C.make_rdf(ref.RDF, case), case)
    b = C.drive(C.make_rdf(O.OracleRDF, case), case)
    assert np.array_equal(a.histogram, b.histogram)


# ---- N1 producers -------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(C.RESIDUE_CASES))
def test_residue_oracle_matches_golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "residues.npz"))
    coor, masses, charges, apr, rfa = C.build_residues(C.RESIDUE_CASES[name])
    com = O.com_by_residue_c(coor, masses, apr, rfa)
    assert np.array_equal(com, z[name + "_com"], equal_nan=True)
    assert np.array_equal(O.dip_by_residue_c(coor, charges, apr, rfa, com), z[name + "_dip"], equal_nan=True)
    assert np.array_equal(O.com_by_residue_c(coor * 0.37, masses, apr, rfa), z[name + "_vel"], equal_nan=True)


def test_residue_oracle_matches_reference_module_when_present():
    H = O.load_ref_helpers()
    if H is None:
        pytest.skip("oracle/_ref/helpers not built")
    coor, masses, charges, apr, rfa = C.build_residues(C.RESIDUE_CASES["ragged"])
    com = H.comByResidue(coor, masses, apr.shape[0], apr, rfa)
    assert np.array_equal(com, O.com_by_residue_c(coor, masses, apr, rfa))
    assert np.array_equal(H.dipByResidue(coor, charges, masses, apr.shape[0], apr, rfa, com),
                          O.dip_by_residue_c(coor, charges, apr, rfa, com))


# ---- N2: shell-resolved g-functions (RDF_voronoi gfunction.pyx:764-1148, calcHistoTSVoro :195-224) --------------
@pytest.mark.parametrize("name", sorted(C.VORO_CASES))
def test_voronoi_oracle_matches_golden(golden_dir, name, tmp_path):
    raw, out, meta = load_golden(golden_dir, name)
    case = C.VORO_CASES[name]
    assert C.voro_digest(case) == meta["digest"]
    rdf = C.drive_voro(C.make_voro(O.OracleRDFVoronoi, case), case)
    assert int(rdf.histo_n) == meta["histo_n"] and int(rdf.nshells) == meta["nshells"]
    assert bool(rdf.oneistwo) == meta["oneistwo"] and rdf.ctr == meta["ctr"]
    assert float(rdf.volume) == meta["volume"]            # update_box leaves the volume alone (reference quirk)
    assert np.array_equal(rdf.raw_sum(), raw)             # bit-exact, all modes and shells
    rdf.write(str(tmp_path / "g"))
    assert np.array_equal(rdf.histogram_out, out)
    for key, text in meta["dat"].items():
        sh, mode = key.split("_")
        assert (tmp_path / ("g_shell%s_g%s.dat" % (sh, mode))).read_text() == text


@pytest.mark.parametrize("name", sorted(C.TSVORO_CASES))
def test_tsvoro_oracle_matches_golden(golden_dir, name):
    want = np.load(os.path.join(golden_dir, "tsvoro.npz"))[name]

    def fn(core, surr, result, ds, maxshell, boxl, t, hmin, hmax, invdx):
        O.calc_histo_ts_voro_c(core, surr, result[t], ds[t], maxshell, boxl, hmin, hmax, invdx)
    assert np.array_equal(C.drive_tsvoro(fn, C.TSVORO_CASES[name]), want)


@pytest.mark.parametrize("name", sorted(C.TS_CASES))
def test_ts_oracle_matches_golden(golden_dir, name):
    """N3 pin: the C restatement of calcHistoTS (gfunction.pyx:168-192) equals what the unmodified reference function
    wrote, including the quirk-Q3 spill across rows and frames (ts_edge)."""
    want = np.load(os.path.join(golden_dir, "ts.npz"))[name]
    got = C.drive_ts(O.calc_histo_ts, C.TS_CASES[name])
    assert np.array_equal(got, want)
    if name == "ts_edge":
        T = C.TS_CASES[name]["frames"]
        assert want[0, 2, 0] == 1 and want[1, 0, 0] == 1 and want[T, 0, 0] == 1       # next row, next frame, spare frame
    # the slab-wise entry point (what the GPU test of larger cases uses) agrees wherever no pair leaves the slab
    if not C.TS_CASES[name].get("edge"):
        case = C.TS_CASES[name]
        for t in range(case["frames"]):
            slab = np.zeros_like(want[t])
            core, surr = C.build_ts_inputs(case, t)
            assert O.calc_histo_ts_c(core, surr, slab, case["box"], case["hmin"], case["hmax"], 1.0 / case["dr"]) == 0
            assert np.array_equal(slab, want[t])


def test_ts_oracle_matches_reference_module_when_present():
    ref = O.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    for name, case in C.TS_CASES.items():
        assert np.array_equal(C.drive_ts(ref.calcHistoTS, case), C.drive_ts(O.calc_histo_ts, case)), name


def test_voronoi_oracle_matches_reference_module():
    ref = O.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    case = dict(n1=90, n2=60, nmol1=30, nmol2=20, box=(18.0, None, None), modes=["all"], hmin=0.0, hmax=9.0, dr=0.1,
                nshells=3, frames=2, same=False, seed=77, exclude_self=True)
    a = C.drive_voro(C.make_voro(ref.RDF_voronoi, case), case)
    b = C.drive_voro(C.make_voro(O.OracleRDFVoronoi, case), case)
    assert np.array_equal(a.histogram, b.histogram)


# ---- N1 + N4 atom frames: the oracle's chain against the fixture made by the unmodified reference modules ----------
def atom_chain_oracle(case):
    a = C.build_atoms(case)
    n1 = case["n1"]; n2 = n1 if case["same"] else case["n2"]
    o = O.OracleRDF(case["modes"], 0.0, case["hmax"], 0.05, n1, n2, n1, n2, case["L"])
    for f in range(case["frames"]):
        if a["boxes"] is not None:
            o.update_box(*[float(v) for v in a["boxes"][f]])
        sel = []
        for tp, pos in ((a["topo1"], a["pos1"]), (a["topo2"], a["pos2"])):
            if tp is None:
                sel.append(sel[0]); continue
            masses, apr, rfa, charges = tp
            coor = np.ascontiguousarray(pos[f], dtype=np.float64)
            com = O.com_by_residue_c(coor, masses, apr, rfa)
            sel.append((com, O.dip_by_residue_c(coor, charges, apr, rfa, com)))
        o.calcFrame(sel[0][0], sel[1][0], sel[0][1], sel[1][1])
    return o


@pytest.mark.parametrize("name", sorted(C.ATOM_CASES))
def test_atom_chain_oracle_matches_golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "atoms.npz"))
    o = atom_chain_oracle(C.ATOM_CASES[name])
    assert np.array_equal(o.raw_sum(), z[name + "_raw"])
    assert np.array_equal(np.array(o.volume_list), z[name + "_volumes"])
    assert bool(o.oneistwo) == bool(z[name + "_oneistwo"])

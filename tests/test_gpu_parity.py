"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the
reference-facing RDF class -> C ABI (libgfn.so), against
  (1) the committed golden fixtures produced by the unmodified reference module, and
  (2) the CPU oracle on seeded inputs, up to BASELINE.json's full single-GPU size (cfg3, n=32768).
Bars: g000 / pair counts bit-exact; orientation-weighted rows within 1e-12 relative (fp64, summation
order differs), with the cancellation caveat of SURVEY.md 8c: atol = 1e-12 * pair weight of the bin.
"""
import json
import os

import numpy as np
import pytest

import cases as C
import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RTOL = 1e-12
# "" / fp16 / fp32: the library picks the kernel itself - the dense-neighbourhood kernel for most of the small golden boxes
# (expected in-range share > 10 %), the warp-specialised ring kernel otherwise; "dense" / "nodense" force one of them
VARIANTS = ["", "fp32", "fp64", "fp16", "fp32,global", "fp64,global", "fp16,global", "per_particle",
            "fp16,dense", "fp32,dense", "fp16,nodense", "fp32,nodense", "nodense,cells", "fp32,nodense,cells", "nodense,cells,global",
            "cells", "fp32,cells"]           # tile skipping on the dense / count kernels (forced: the golden boxes are small)


def RDF(*a, **k):
    from newanalysis.gfunction import RDF as R
    return R(*a, **k)


def load_golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return z["raw"], z["out"], json.loads(str(z["meta"]))


def assert_hist_close(got, ref, hn, pairw=None):
    got = np.asarray(got); ref = np.asarray(ref)
    if pairw is None:
        pairw = ref[:hn]
    # count-like rows: exact wherever the reference row is integer valued (g000)
    scale = np.maximum(np.abs(ref), np.tile(np.abs(pairw), 7))
    err = np.abs(got - ref)
    bad = err > RTOL * scale
    assert not bad.any(), "max err/scale = %g at %s" % (np.max(err / np.maximum(scale, 1e-300)), np.flatnonzero(bad)[:5])


@pytest.mark.parametrize("flags", VARIANTS)
@pytest.mark.parametrize("name", sorted(C.CASES))
def test_golden(golden_dir, name, flags, tmp_path, monkeypatch):
    if "cells" in flags:
        monkeypatch.setenv("GFN_CELLS_FORCE", "1")         # small boxes: the library would not bother to skip tiles
    raw, out, meta = load_golden(golden_dir, name)
    case = C.CASES[name]
    n1, n2 = case["n1"], case["n2"]
    bx, by, bz = case["box"]
    rdf = RDF(case["modes"], case["hmin"], case["hmax"], case["dr"], n1, n2, n1, n2, bx, by, bz, gfn_flags=flags)
    C.drive(rdf, case)
    hn = meta["histo_n"]
    assert int(rdf.histo_n) == hn and int(rdf.mode_sel) == meta["mode_sel"]
    assert float(rdf.histo_max) == meta["histo_max"] and float(rdf.histo_inv_dr) == meta["histo_inv_dr"]
    assert bool(rdf.oneistwo) == meta["oneistwo"] and rdf.ctr == meta["ctr"]
    got = rdf.raw_histogram()
    if meta["mode_sel"] & 1:
        assert np.array_equal(got[:hn], raw[:hn]), "g000 not bit-exact"
    # pair weight row is what g000 would hold (minus Q3 spills), whatever the modes
    assert_hist_close(got, raw, hn, pairw=np.maximum(rdf.pair_weight, raw[:hn]))
    rdf.write(str(tmp_path / "g"))
    if meta["mode_sel"] & 1:
        assert np.array_equal(rdf.histogram_out[:hn], out[:hn])
        assert (tmp_path / "g_g000.dat").read_text() == meta["dat"]["000"]
    for mode, text in meta["dat"].items():
        a = np.loadtxt(str(tmp_path / ("g_g%s.dat" % mode)))
        b = np.array([[float(v) for v in ln.split()] for ln in text.strip().splitlines()])
        assert a.shape == b.shape and np.allclose(a, b, rtol=0, atol=2e-5)


def test_checked_build_runs_the_small_cases_clean():
    """compute-sanitizer is closed on the GPU pool (profiles/sanitizer/README.md).  Its stand-in: libgfn compiled with
    -DGFN_CHECKED (every ring / stage / tile / histogram index and protocol invariant is tested on the device and raises a
    sticky error) runs the small golden cases of every kernel variant - ring, count, dense, tile skipping, shells, time
    series - in a child process that finds the checked library first."""
    import subprocess
    import sys
    chk = os.path.join(ROOT, "mdy-newanalysis-package_b200", "newanalysis", "checked")
    if not os.path.exists(os.path.join(chk, "libgfn.so")):
        pytest.skip("checked library not built (make -C mdy-newanalysis-package_b200/csrc checked)")
    env = dict(os.environ, LD_LIBRARY_PATH=chk + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "sanitizer", "run_cases.py")], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "sanitizer cases done" in out.stdout and "checked/libgfn.so" in out.stdout, out.stdout[-500:]


@pytest.mark.parametrize("modes,dr,stripes,kernel", [(["all"], 0.0125, 2, "dense"), (["all"], 0.005, 1, "dense"), (["000", "110", "220"], 0.02, 4, "dense"),
                                                    (["000"], 0.004, 1, "count"), (["000"], 0.0125, 16, "count")])
def test_dense_box_kernels_with_many_bins(modes, dr, stripes, kernel):
    """Histograms too large for the preferred shared-memory layout: the dense kernel falls back from 8 bank-group stripes to
    4, 2 or 1 (bin updates then replay, results do not change), the count kernel from one replica per warp to one per CTA."""
    rng = np.random.default_rng(int(dr * 1e5))
    n, L, nf, hmax = 900, 33.0, 3, 15.0
    g = RDF(modes, 0.0, hmax, dr, n, n, n, n, L)
    o = O.OracleRDF(modes, 0.0, hmax, dr, n, n, n, n, L)
    st = g.stats()
    assert st["kernel"] == kernel and st["hist_replicas"] == stripes, st
    for f in range(nf):
        c = rng.random((n, 3)) * L; d = rng.standard_normal((n, 3))
        if modes == ["000"]:
            g.calcFrame(c, c); o.calcFrame(c, c)
        else:
            g.calcFrame(c, c, d, d); o.calcFrame(c, c, d, d)
    hn = int(g.histo_n)
    got, ref = g.raw_histogram(), o.raw_sum()
    assert np.array_equal(got[:hn], ref[:hn]) and ref[:hn].sum() > 1000
    assert_hist_close(got, ref, hn)


@pytest.mark.parametrize("n1,n2", [(100, 6000), (64, 3000), (200, 200 * 13)])
def test_dense_kernel_few_units_per_tile(n1, n2):
    """A small core selection against a large surround one in a dense box: one to four 64-row units per surround tile, so the
    dense kernel's 16 warps run many tiles ahead of its 3-stage TMA ring.  (An mbarrier parity cannot tell a tile from the
    one two uses before it: the kernel checks the stage's generation word first.)"""
    rng = np.random.default_rng(n1 * 7 + n2)
    L, nf = 36.0, 6
    g = RDF(["all"], 0.0, 15.0, 0.05, n1, n2, n1, n2, L, gfn_flags="dense")
    assert g.stats()["kernel"] == "dense"
    o = O.OracleRDF(["all"], 0.0, 15.0, 0.05, n1, n2, n1, n2, L)
    for f in range(nf):
        c1 = rng.random((n1, 3)) * L; d1 = rng.standard_normal((n1, 3))
        c2 = rng.random((n2, 3)) * L; d2 = rng.standard_normal((n2, 3))
        g.calcFrame(c1, c2, d1, d2); o.calcFrame(c1, c2, d1, d2)
    got, ref = g.raw_histogram(), o.raw_sum()
    assert np.array_equal(got[:300], ref[:300]) and ref[:300].sum() > 1000
    assert_hist_close(got, ref, 300)


def test_tile_skipping_is_used_only_where_it_pays():
    """GFN_FLAG_CELLS is honoured when the expected share of visited tile pairs is small (a box many tile diameters wide)."""
    mk = lambda n, L, **kw: RDF(["all"], 0.0, 15.0, 0.05, n, n, n, n, L, **kw).stats()      # noqa: E731
    assert mk(32768, 99.33, gfn_flags="cells")["cells"] is False          # BASELINE configs[2]: most tile pairs stay
    st = mk(262144, 198.66, gfn_flags="cells")                            # the same density, eight times the volume
    assert st["cells"] is True
    assert mk(262144, 198.66)["cells"] is False
    assert mk(4096, 31.0, gfn_flags="cells")["cells"] is False            # a box of two tile diameters


def test_kernel_choice_follows_the_expected_in_range_share():
    """Dense boxes (BASELINE configs[0] / [1] shapes) get the dense kernel, sparse ones (cfg3 density) the ring kernel;
    flags override; handles the dense kernel does not support fall back to the ring kernel."""
    mk = lambda L, hmax, modes, n=256, **kw: RDF(modes, 0.0, hmax, 0.05, n, n, n, n, L, **kw)      # noqa: E731
    assert mk(31.04, 15.0, ["000"]).stats()["kernel"] == "count"            # g000 only: fp32 classification + exact pass
    assert mk(31.04, 15.0, ["000"], gfn_flags="fp16").stats()["kernel"] == "count"
    assert mk(31.04, 15.0, ["000"], gfn_flags="fp64").stats()["kernel"] == "ring"
    assert mk(99.33, 15.0, ["000"]).stats()["kernel"] == "ring"
    assert mk(67.7, 30.0, ["000", "110", "101", "011", "220"]).stats()["kernel"] == "dense"
    assert mk(99.33, 15.0, ["all"]).stats()["kernel"] == "ring"
    assert mk(99.33, 15.0, ["all"], gfn_flags="dense").stats()["kernel"] == "dense"
    assert mk(31.04, 15.0, ["all"], gfn_flags="nodense").stats()["kernel"] == "ring"
    assert mk(31.04, 15.0, ["all"], gfn_flags="fp64").stats()["kernel"] == "ring"
    assert mk(31.04, 15.0, ["all"], gfn_flags="global").stats()["kernel"] == "ring"
    st = mk(31.04, 15.0, ["000"]).stats()
    assert st["warps_per_block"] == 16 and st["hist_replicas"] == 16
    st = mk(31.04, 15.0, ["all"]).stats()
    assert st["warps_per_block"] == 16 and st["hist_replicas"] == 8          # two warps share a fixed-point replica
    st = RDF(["000", "110", "101", "011", "220"], 0.0, 30.0, 0.05, 256, 256, 256, 256, 67.7).stats()     # 600 bins, 4 weighted columns
    assert st["kernel"] == "dense" and st["warps_per_block"] == 16 and st["hist_replicas"] == 8


def test_micro_cases(golden_dir):
    idx = np.load(os.path.join(golden_dir, "micro.npz"))["idx"]
    for (y, _), want in zip(C.MICRO, idx):
        rdf = RDF(["000"], 0.0, 15.0, 0.05, 1, 2, 1, 2, 40.0, gfn_flags="per_particle")
        rdf.calcFrame(np.zeros((1, 3)), np.array([y, (30.0, 30.0, 30.0)]))
        h = rdf.histogram
        nz = np.flatnonzero(h)
        assert (int(nz[0]) if nz.size else -1) == want


def test_per_particle_matches_reference_layout():
    case = C.CASES["rect_all"]
    a = C.drive(C.make_rdf(O.OracleRDF, case), case)
    n1, n2 = case["n1"], case["n2"]
    b = RDF(case["modes"], case["hmin"], case["hmax"], case["dr"], n1, n2, n1, n2, case["box"][0], gfn_flags="per_particle")
    C.drive(b, case)
    hp = b.histogram
    hn = int(b.histo_n)
    assert hp.shape == a.histogram.shape
    assert np.array_equal(hp[:n1 * hn], a.histogram[:n1 * hn])
    scale = np.maximum(np.abs(a.histogram), np.tile(a.histogram[:n1 * hn], 7))
    assert np.all(np.abs(hp - a.histogram) <= RTOL * scale)
    out = np.zeros(7 * hn)
    b._addHisto(out, hp)
    assert_hist_close(out, a.raw_sum(), hn)


@pytest.mark.parametrize("filt", ["fp16", "fp32", "fp64"])
def test_exact_div_flag_gives_reference_increments_bit_for_bit(filt):
    """GFN_FLAG_EXACT_DIV: the cosines are the reference's literal divisions (gfunction.pyx:155,159,163), so a
    per-particle histogram entry that received exactly ONE pair holds exactly the reference's increment - all seven
    modes, bit for bit.  (Without the flag the reciprocal form may differ by a few ulp; checked to stay below 1e-14.)"""
    case = C.CASES["rect_all"]
    n1, n2 = case["n1"], case["n2"]
    a = C.drive(C.make_rdf(O.OracleRDF, dict(case, frames=1)), dict(case, frames=1))
    ref = a.histogram.reshape(7, n1, -1)
    single = ref[0] == 1.0                      # rectangular case: weight 1 per pair
    assert single.sum() > 5000
    for flags, exact in ((filt + ",per_particle,exact_div", True), (filt + ",per_particle", False)):
        b = RDF(case["modes"], case["hmin"], case["hmax"], case["dr"], n1, n2, n1, n2, case["box"][0], gfn_flags=flags)
        C.drive(b, dict(case, frames=1))
        got = b.histogram.reshape(7, n1, -1)
        assert np.array_equal(got[0], ref[0])
        for k in range(1, 7):
            if exact:
                assert np.array_equal(got[k][single], ref[k][single]), "mode row %d: increments differ from the reference's" % k
            else:
                assert np.max(np.abs(got[k][single] - ref[k][single])) < 1e-14


def test_exact_div_flag_on_shared_replicas():
    """The flag on the default (shared-memory replica) path: counts exact, sums within the usual bar, and closer to the
    reference's than 1e-13 of the pair weight (summation order only)."""
    case = C.CASES["cfg3_density"]
    a = C.drive(C.make_rdf(O.OracleRDF, case), case)
    b = C.drive(RDF(case["modes"], case["hmin"], case["hmax"], case["dr"], case["n1"], case["n2"], case["n1"], case["n2"],
                    case["box"][0], gfn_flags="exact_div"), case)
    got, ref = b.raw_histogram(), a.raw_sum()
    assert np.array_equal(got[:300], ref[:300])
    assert_hist_close(got, ref, 300)
    assert np.max(np.abs(got - ref) / np.maximum(np.tile(ref[:300], 7), 1.0)) < 1e-13


def _random_frame(rng, n, L, dip=True):
    return np.ascontiguousarray(rng.random((n, 3)) * L), (np.ascontiguousarray(rng.standard_normal((n, 3))) if dip else None)


@pytest.mark.parametrize("flags", ["fp32", "fp64", "fp16"])
@pytest.mark.parametrize("n,L", [(8192, 62.58), (32768, 99.33)])
def test_cfg3_shape_against_oracle(n, L, flags):
    """BASELINE configs[2] shape (water, self, all modes, 0-15/0.05); n=32768 is the full size."""
    rng = np.random.default_rng(3000 + n)
    c, d = _random_frame(rng, n, L)
    rdf = RDF(["all"], 0.0, 15.0, 0.05, n, n, n, n, L, gfn_flags=flags)
    rdf.calcFrame(c, c, d, d)
    got = rdf.raw_histogram()
    ref = np.zeros(7 * 300)
    box = np.array([L, L, L, L / 2, L / 2, L / 2])
    rc, oob = O.calc_histo_rows_c(c, c, d, d, ref, box, 0.0, 15.0, 300, 20.0, 127)
    assert rc == 0 and oob == 0
    assert np.array_equal(got[:300], ref[:300])
    # the rows oracle itself differs from the reference order by ~1e-13 relative: allow 2x
    scale = np.maximum(np.abs(ref), np.tile(ref[:300], 7))
    assert np.all(np.abs(got - ref) <= 2 * RTOL * scale)
    st = rdf.stats()
    assert st["pair_evals"] == n * (n + 1) / 2 and st["in_range"] * 2 - 0 >= ref[:300].sum() - 1


def test_cfg2_shape_many_frames_batched():
    """[C2mim][OTf] shape: 1000 x 1000 distinct selections of equal size (quirk Q1), 600 bins, 5 modes."""
    n, L, nf = 1000, 67.7, 12
    modes = ["000", "110", "101", "011", "220"]
    rng = np.random.default_rng(2000)
    c1 = rng.random((nf, n, 3)) * L; c2 = rng.random((nf, n, 3)) * L
    d1 = rng.standard_normal((nf, n, 3)); d2 = rng.standard_normal((nf, n, 3))
    a = RDF(modes, 0.0, 30.0, 0.05, n, n, n, n, L)
    a.calcFrames(c1, c2, d1, d2)
    b = RDF(modes, 0.0, 30.0, 0.05, n, n, n, n, L)
    o = O.OracleRDF(modes, 0.0, 30.0, 0.05, n, n, n, n, L)
    for f in range(nf):
        b.calcFrame(c1[f], c2[f], d1[f], d2[f])
        o.calcFrame(c1[f], c2[f], d1[f], d2[f])
    ref = o.raw_sum()
    ha, hb = a.raw_histogram(), b.raw_histogram()
    assert np.array_equal(ha[:600], hb[:600])          # batch call == per-frame calls
    if a.stats()["hist_replicas"] == 4:
        assert np.array_equal(ha, hb)                  # private replicas: fixed summation order, bit-identical
    else:
        assert_hist_close(ha, hb, 600)                 # replicas shared under a lock: order may differ
    assert np.array_equal(ha[:600], ref[:600])
    assert_hist_close(ha, ref, 600)
    assert a.ctr == nf and not a.oneistwo


def test_threshold_edges_exact():
    """Pairs placed within a few ulps of histo_max, of the first-bin edge 1/invdx and of bin edges,
    straight and across the periodic boundary: the pre-filter must never lose a pair the reference
    keeps, and bins must agree exactly."""
    L = 40.0
    rng = np.random.default_rng(11)
    pts = []
    for base in (15.0, 0.05, 7.35, 14.95, 0.1):
        for k in range(-6, 7):
            r = base
            for _ in range(abs(k)):
                r = np.nextafter(r, np.inf if k > 0 else -np.inf)
            for _ in range(4):
                u = rng.standard_normal(3); u /= np.linalg.norm(u)
                pts.append(u * r)
            pts.append(np.array([r, 0, 0])); pts.append(np.array([0, -r, 0]))
    pts = np.array(pts)
    origin = np.array([[39.7, 0.2, 20.0]])
    c2 = np.ascontiguousarray((origin + pts) % L)
    for flags in ("fp32", "fp64", "fp16"):
        g = RDF(["000"], 0.0, 15.0, 0.05, 1, len(c2), 1, len(c2), L, gfn_flags=flags + ",per_particle")
        o = O.OracleRDF(["000"], 0.0, 15.0, 0.05, 1, len(c2), 1, len(c2), L)
        o.histogram = np.zeros(2 * 300 * 7)          # spare row so the Q3 spill stays inspectable
        g.calcFrame(origin, c2); o.calcFrame(origin, c2)
        hp = g.histogram
        assert np.array_equal(hp[:300], o.histogram[:300])
        assert hp[:300].sum() > 50


@pytest.mark.parametrize("hmin,hmax,dr,L", [(0.0, 15.0, 0.05, 40.0), (2.0, 12.0, 0.1, 32.0), (0.0, 30.0, 0.05, 67.7), (0.0, 15.02, 0.05, 40.0)])
@pytest.mark.parametrize("filt,kernel", [("fp16", "count"), ("fp32", "count")])
def test_dense_threshold_table_at_every_bin_edge(hmin, hmax, dr, L, filt, kernel):
    """The exact pass of the g000 count kernel takes the bin from a table of exact d^2 thresholds instead of the square root:
    pairs placed within +-4 ulp of EVERY bin edge (and of histo_min, histo_max and the first-bin rule) along all three
    axes, straight and across the periodic boundary, must land in the reference's bins - also the quirk-Q3 spill count.
    (For the count kernel all of these pairs are ambiguous in fp32 and go through its exact pass.)"""
    nb = int(np.int32((hmax - hmin) / dr + 0.5))
    edges = [np.float64(np.float32(hmin)) + k * np.round(dr, 5) for k in range(nb + 2)] + [1.0 / (1.0 / np.round(dr, 5)), float(np.float32(hmax))]
    rs = []
    for e in edges:
        for k in range(-4, 5):
            r = e
            for _ in range(abs(k)):
                r = np.nextafter(r, np.inf if k > 0 else -np.inf)
            if 0 < r < L / 2 * 0.99:
                rs.append(r)
    rs = np.array(rs)
    rng = np.random.default_rng(12)
    origin = np.array([[L - 0.3, 0.2, L / 2]])
    pts = np.zeros((len(rs), 3))
    axis = rng.integers(0, 3, len(rs)); sign = rng.choice([-1.0, 1.0], len(rs))
    pts[np.arange(len(rs)), axis] = sign * rs
    c2 = np.ascontiguousarray((origin + pts) % L)
    g = RDF(["000"], hmin, hmax, dr, 1, len(c2), 1, len(c2), L, gfn_flags=filt + ",dense")
    assert g.stats()["kernel"] == kernel
    o = O.OracleRDF(["000"], hmin, hmax, dr, 1, len(c2), 1, len(c2), L)
    o.histogram = np.zeros(2 * nb * 7)               # spare row so the Q3 spill stays inspectable
    for _ in range(2):
        g.calcFrame(origin, c2); o.calcFrame(origin, c2)
    got = g.raw_histogram()
    assert np.array_equal(got[:nb], o.histogram[:nb]) and got[:nb].sum() > 100
    assert g.overflow == o.histogram[nb:2 * nb].sum()                     # pairs at exactly histo_max (pos == histo_n)
    assert np.array_equal(got[nb:2 * nb], o.histogram[nb:2 * nb])         # ... land in the next mode's row like the reference's


@pytest.mark.parametrize("case", ["self", "self_hmin", "rect", "rect_unwrapped", "self_far", "npt_fine", "two_selections_equal_size"])
def test_count_kernel_selfcheck_and_oracle(case, monkeypatch):
    """g000 count kernel (gfn_count.cu): GFN_DEBUG=8 makes it evaluate EVERY pair in fp64 as well and count the pairs whose
    fp32 decision (bin / rejection) differs from the exact one - there must be none - and the counts equal the oracle's.
    Cases: triangular and rectangular frames, histo_min > 0, coordinates far outside the box (large EPS: more pairs take
    the exact pass, none may be wrong), per-frame boxes with fine bins."""
    monkeypatch.setenv("GFN_DEBUG", "8")
    import zlib
    rng = np.random.default_rng(zlib.crc32(case.encode()))
    hmin, hmax, dr, L, nf = 0.0, 15.0, 0.05, 31.04, 3
    n1 = n2 = 1000
    shift = 0.0
    if case == "self_hmin":
        hmin, hmax, dr = 2.0, 14.0, 0.1
    if case.startswith("rect"):
        n1, n2 = 300, 1500
    if case == "rect_unwrapped":
        shift = 40.0 * L
    if case == "self_far":
        shift = 3000.0 * L
    if case == "npt_fine":
        dr, nf = 0.01, 4
    g = RDF(["000"], hmin, hmax, dr, n1, n2, n1, n2, L, gfn_flags="dense")
    assert g.stats()["kernel"] == "count"
    o = O.OracleRDF(["000"], hmin, hmax, dr, n1, n2, n1, n2, L)
    for f in range(nf):
        if case == "npt_fine":
            b = L * (1.0 + 0.01 * rng.standard_normal(3))
            g.update_box(*b); o.update_box(*b)
        c1 = rng.random((n1, 3)) * L + shift * rng.integers(-1, 2, (n1, 3))
        c2 = c1 if n1 == n2 else rng.random((n2, 3)) * L + shift * rng.integers(-1, 2, (n2, 3))
        if case == "two_selections_equal_size":
            # quirk Q1: the triangle j >= i is keyed on the SIZES; (i, i) is then a real pair of weight 1 - put every second
            # partner within range of its namesake
            c2 = np.ascontiguousarray(rng.random((n2, 3)) * L)
            c2[::2] = c1[::2] + 3.0 * rng.standard_normal(c1[::2].shape)
        g.calcFrame(c1, c2); o.calcFrame(c1, c2)
    hn = int(g.histo_n)
    got, ref = g.raw_histogram(), o.raw_sum()
    st = g.stats()
    assert st["selfcheck_bad"] == 0
    assert np.array_equal(got[:hn], ref[:hn]) and ref[:hn].sum() > 1000
    if case != "two_selections_equal_size":
        assert st["in_range"] * (2 if n1 == n2 else 1) == ref[:hn].sum() + (2 if n1 == n2 else 1) * st["overflow"]
    else:
        assert ref[:hn].sum() % 2 == 1 or ref[:hn].sum() < 2 * st["in_range"]         # pairs of weight 1 are in there
    if shift == 0.0:
        assert st["exact_evals"] < (0.01 if dr >= 0.05 else 0.06) * st["in_range"]      # the fp64 pass stays a small share


@pytest.mark.parametrize("case", ["self_all", "self_all_fp32", "rect", "unwrapped_npt", "per_particle", "g000_spill",
                                  "dense_self_all", "dense_rect_5modes", "dense_unwrapped_npt", "dense_two_selections",
                                  "count_self", "count_rect", "count_spill", "count_unwrapped_npt"])
def test_tile_skipping_equals_the_full_enumeration_and_the_oracle(case, monkeypatch):
    """GFN_FLAG_CELLS (gfn_cells.cu): frames sorted into Morton order on the device, only tile pairs within histo_max are
    visited.  Counts must equal the un-pruned run's and the oracle's bit for bit - in a triangular frame that needs the
    reference's core/surround roles restored from the ORIGINAL indices (modes 101 / 011 / 202 / 022 are not symmetric
    under the exchange) - weighted sums within the bar, and the visit count must actually drop."""
    monkeypatch.setenv("GFN_CELLS_FORCE", "1")
    rng = np.random.default_rng(zlib_seed(case))
    n1 = n2 = 20000; L, hmax, nf, modes, flags = 200.0, 10.0, 2, ["all"], ""
    if case == "self_all_fp32":
        flags = "fp32,"
    if case == "rect":
        n1, n2 = 3000, 24000
    if case == "per_particle":
        n1 = n2 = 3000; L = 100.0; flags = "per_particle,"
    if case == "g000_spill":
        modes, hmax = ["000"], 10.02
    # "dense_*": orientation modes with tile skipping run on the dense kernel (visit lists of 64-row blocks per surround tile)
    kern = "dense" if case.startswith("dense") else ("count" if case.startswith("count") else "ring")
    if kern == "count":                                   # g000 only: the count kernel walks lists of 128-row blocks
        modes = ["000"]
        if case == "count_rect":
            n1, n2 = 3000, 24000
        if case == "count_spill":
            hmax = 10.02
        case = {"count_self": "g000", "count_rect": "g000_rect", "count_spill": "g000_spill", "count_unwrapped_npt": "unwrapped_npt"}[case]
    if case == "dense_rect_5modes":
        n1, n2, modes = 3000, 24000, ["000", "110", "101", "011", "220"]
    if case == "dense_two_selections":
        n1 = n2 = 6000; L = 130.0
    case = case.replace("dense_", "") if case != "dense_two_selections" else case
    shift = 0.0
    a = RDF(modes, 0.0, hmax, 0.05, n1, n2, n1, n2, L, gfn_flags=flags + ("nodense,cells" if kern == "ring" else ("cells,dense" if kern == "count" else "cells")))
    b = RDF(modes, 0.0, hmax, 0.05, n1, n2, n1, n2, L, gfn_flags=flags + "nodense")
    o = O.OracleRDF(modes, 0.0, hmax, 0.05, n1, n2, n1, n2, L)
    hn = int(a.histo_n)
    for f in range(nf):
        if case == "unwrapped_npt":
            bx = L * (1.0 + 0.02 * rng.standard_normal(3)); shift = 2.0
            for r in (a, b, o):
                r.update_box(*bx)
        c1 = rng.random((n1, 3)) * L; d1 = rng.standard_normal((n1, 3))
        if shift:
            c1 = c1 + L * rng.integers(-2, 3, (n1, 3))        # single-shift minimum image (quirk Q8): far images do not count
        c2, d2 = (c1, d1) if n1 == n2 else (rng.random((n2, 3)) * L, rng.standard_normal((n2, 3)))
        if case == "dense_two_selections":          # equal sizes, different arrays (quirk Q1): such launches keep the full enumeration
            c2, d2 = rng.random((n2, 3)) * L, rng.standard_normal((n2, 3))
        for r in (a, b, o):
            if modes == ["000"]:
                r.calcFrame(c1, c2)
            else:
                r.calcFrame(c1, c2, d1, d2)
    ha, hb = a.raw_histogram(), b.raw_histogram()
    ref = o.raw_sum()
    sa, sb = a.stats(), b.stats()
    assert sa["cells"] and not sb["cells"] and sa["kernel"] == kern
    assert np.array_equal(ha[:hn], ref[:hn]) and np.array_equal(hb[:hn], ref[:hn]) and ref[:hn].sum() > 1000
    if case == "g000_spill":
        # pairs in (10.00, 10.02] have pos == histo_n (quirk Q3): they spill into the next row exactly where the un-pruned run puts them
        assert a.overflow == b.overflow and a.overflow > 0
        assert np.array_equal(ha, hb) and np.array_equal(ha, ref)
    else:
        assert_hist_close(ha, ref, hn); assert_hist_close(hb, ref, hn)
    assert sa["in_range"] == sb["in_range"]
    assert 0 <= sa["tile_visits"] <= max(sa["tile_visits_dense"], 1)
    if case == "dense_two_selections":
        assert sa["tile_visits"] == 0
    elif case not in ("per_particle", "rect", "g000_rect"):
        assert sa["tile_visits"] < 0.85 * sa["tile_visits_dense"]       # the box is a few tile diameters wide
    if case == "per_particle":
        pa, pb = a.histogram, b.histogram           # the reference's [mode][core i][bin] array: rows are ORIGINAL indices
        assert np.array_equal(pa[:n1 * hn], pb[:n1 * hn])
        assert np.allclose(pa, pb, rtol=0, atol=1e-9)


def zlib_seed(s):
    import zlib
    return zlib.crc32(s.encode())


def test_unwrapped_and_large_coordinates():
    """Coordinates far outside the box (single-shift minimum image, quirk Q8) and a large offset:
    the guard band scales with max|x|, results must still match the oracle exactly."""
    n, L = 600, 30.0
    rng = np.random.default_rng(5)
    c = (rng.random((n, 3)) * 3 - 1) * L + 1000.0
    d = rng.standard_normal((n, 3))
    g = RDF(["all"], 0.0, 14.0, 0.05, n, n, n, n, L)
    o = O.OracleRDF(["all"], 0.0, 14.0, 0.05, n, n, n, n, L)
    g.calcFrame(c, c, d, d); o.calcFrame(c, c, d, d)
    got, ref = g.raw_histogram(), o.raw_sum()
    assert np.array_equal(got[:280], ref[:280]) and ref[:280].sum() > 0
    assert_hist_close(got, ref, 280)


def test_zero_norm_dipole_raises_deferred():
    rdf = RDF(["all"], 0.0, 15.0, 0.05, 4, 5, 4, 5, 40.0)
    rng = np.random.default_rng(0)
    c1, c2 = rng.random((4, 3)) * 5, rng.random((5, 3)) * 5
    d1, d2 = rng.standard_normal((4, 3)), rng.standard_normal((5, 3))
    d2[2] = 0.0
    rdf.calcFrame(c1, c2, d1, d2)
    with pytest.raises(ZeroDivisionError, match="float division"):
        rdf.sync()
    assert rdf.ctr == 1


def test_argument_errors_match_reference():
    rdf = RDF(["110"], 0.0, 15.0, 0.05, 4, 5, 4, 5, 40.0)
    with pytest.raises(TypeError, match="fist selection"):
        rdf.calcFrame(np.zeros((4, 3)), np.zeros((5, 3)))
    with pytest.raises(TypeError, match="second selection"):
        rdf.calcFrame(np.zeros((4, 3)), np.zeros((5, 3)), np.ones((4, 3)))
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):
        rdf.calcFrame(np.zeros((4, 3), dtype=np.float32), np.zeros((5, 3)), np.ones((4, 3)), np.ones((5, 3)))
    with pytest.raises(ValueError, match="not C-contiguous"):
        rdf.calcFrame(np.zeros((3, 4)).T, np.zeros((5, 3)), np.ones((4, 3)), np.ones((5, 3)))


def test_g000_without_dipoles_and_unknown_modes():
    n, L = 500, 30.0
    rng = np.random.default_rng(8)
    c1, _ = _random_frame(rng, n, L, dip=False); c2, _ = _random_frame(rng, 300, L, dip=False)
    g = RDF(["000", "bogus"], 0.0, 15.0, 0.05, n, 300, n, 300, L)
    o = O.OracleRDF(["000", "bogus"], 0.0, 15.0, 0.05, n, 300, n, 300, L)
    g.calcFrame(c1, c2); o.calcFrame(c1, c2)
    assert np.array_equal(g.raw_histogram(), o.raw_sum())


def test_scale_reset_write_twice_semantics(tmp_path):
    """scale / resetArrays / non-idempotent write (quirk Q7) behave like the reference class."""
    case = C.CASES["self_odd"]
    g = C.drive(C.make_rdf(RDF, case), case)
    o = C.drive(C.make_rdf(O.OracleRDF, case), case)
    g.scale(0.5); o.scale(0.5)
    hn = int(g.histo_n)
    assert np.array_equal(g.raw_histogram()[:hn], o.raw_sum()[:hn])
    g.write(str(tmp_path / "a")); o.write(str(tmp_path / "b"))
    assert np.array_equal(g.histogram_out[:hn], o.histogram_out[:hn])
    assert (tmp_path / "a_g000.dat").read_text() == (tmp_path / "b_g000.dat").read_text()
    g.write(str(tmp_path / "a")); o.write(str(tmp_path / "b"))           # second write accumulates again
    # the reference adds the n1 rows one by one onto the already normalised (non-integer) values, this
    # backend adds their sum: same value up to rounding, not bit-identical (documented in INTEGRATION.md)
    assert np.allclose(g.histogram_out[:hn], o.histogram_out[:hn], rtol=1e-12, atol=0)
    g.resetArrays(); o.resetArrays()
    assert not g.raw_histogram().any() and not g.histogram_out.any() and g.ctr == o.ctr == case["frames"]
    c1, c2, d1, d2, _ = C.build_inputs(case, 0)
    g.calcFrame(c1, c2, d1, d2); o.calcFrame(c1, c2, d1, d2)
    assert np.array_equal(g.raw_histogram()[:hn], o.raw_sum()[:hn])


def test_linearity_and_determinism_at_full_size():
    """Size-independent properties at cfg3's full n: H(frame A) + H(frame B) == H(A then B) exactly for
    counts, and two runs of the same input are bit-identical in every row (fixed summation order)."""
    n, L = 32768, 99.33
    rng = np.random.default_rng(99)
    fa = _random_frame(rng, n, L); fb = _random_frame(rng, n, L)
    def run(frames):
        r = RDF(["all"], 0.0, 15.0, 0.05, n, n, n, n, L)
        for c, d in frames:
            r.calcFrame(c, c, d, d)
        return r.raw_histogram(), r
    ha, _ = run([fa]); hb, _ = run([fb]); hab, r = run([fa, fb]); hab2, _ = run([fa, fb])
    assert np.array_equal(hab, hab2)
    assert np.array_equal(ha[:300] + hb[:300], hab[:300])
    assert np.allclose(ha + hb, hab, rtol=1e-12, atol=1e-12 * hab[:300].max())
    # checksum of counts: in-range ordered pairs == sum of g000
    assert r.stats()["in_range"] * 2 == hab[:300].sum()            # self case: i==j never in range, every pair weight 2


def test_ring_recycling_and_npt(monkeypatch):
    """More batches than ring slots with a box that changes every frame."""
    monkeypatch.setenv("GFN_BATCH_FRAMES", "2")
    n, L, nf = 400, 30.0, 15
    g = RDF(["all"], 0.0, 14.0, 0.05, n, n, n, n, L)
    o = O.OracleRDF(["all"], 0.0, 14.0, 0.05, n, n, n, n, L)
    assert g.stats()["batch_frames"] == 2
    for f in range(nf):
        rng = np.random.default_rng(400 + f)
        s = L * (1 + 1e-3 * rng.standard_normal(3))
        c = np.ascontiguousarray(rng.random((n, 3)) * s); d = rng.standard_normal((n, 3))
        g.update_box(*s); o.update_box(*s)
        g.calcFrame(c, c, d, d); o.calcFrame(c, c, d, d)
        c[:] = 0; d[:] = 0          # caller reuses its arrays immediately
    got, ref = g.raw_histogram(), o.raw_sum()
    assert np.array_equal(got[:280], ref[:280])
    assert_hist_close(got, ref, 280)
    assert g.volume_list == o.volume_list


@pytest.mark.parametrize("case", ["self_all", "rect_5modes", "g000_npt"])
def test_pinned_frames_and_pool_copies_equal_staged_frames(case, monkeypatch):
    """Three host feeds of the same frames: per-frame calcFrame (staged by the caller thread), calcFrames blocks (frames of
    a batch copied by the pool threads) and calcFrames(pinned=True) (no staging at all: strided DMA from page-locked
    arrays, gfn_enqueue_frames_pinned).  Small batches so that every feed crosses several ring slots; all against the oracle."""
    from newanalysis.gfunction import pinned_empty
    monkeypatch.setenv("GFN_BATCH_FRAMES", "3")
    monkeypatch.setenv("GFN_POOL_THREADS", "3")
    rng = np.random.default_rng(515)
    nf, L = 11, 42.0
    if case == "self_all":
        n1 = n2 = 700; modes = ["all"]
    elif case == "rect_5modes":
        n1, n2 = 300, 900; modes = ["000", "110", "101", "011", "220"]
    else:
        n1 = n2 = 500; modes = ["000"]
    boxes = L * (1 + 2e-3 * rng.standard_normal((nf, 3))) if case == "g000_npt" else None
    same = n1 == n2
    pc1, pd1 = pinned_empty((nf, n1, 3)), pinned_empty((nf, n1, 3))
    pc1[:] = rng.random((nf, n1, 3)) * L; pd1[:] = rng.standard_normal((nf, n1, 3))
    if same:
        pc2, pd2 = pc1, pd1
    else:
        pc2, pd2 = pinned_empty((nf, n2, 3)), pinned_empty((nf, n2, 3))
        pc2[:] = rng.random((nf, n2, 3)) * L; pd2[:] = rng.standard_normal((nf, n2, 3))
    need_d = modes != ["000"]
    mk = lambda: RDF(modes, 0.0, 15.0, 0.05, n1, n2, n1, n2, L)      # noqa: E731
    a, b, c = mk(), mk(), mk()
    o = O.OracleRDF(modes, 0.0, 15.0, 0.05, n1, n2, n1, n2, L)
    for f in range(nf):
        if boxes is not None:
            a.update_box(*boxes[f]); o.update_box(*boxes[f])
        args = (pc1[f], pc2[f]) + ((pd1[f], pd2[f]) if need_d else ())
        a.calcFrame(*args); o.calcFrame(*args)
    dargs = (pd1, pd2) if need_d else (None, None)
    # two calls each, the second one starting in the middle of a batch
    for lo, hi in ((0, 4), (4, nf)):
        c1v = pc1[lo:hi].copy(); c2v = c1v if same else pc2[lo:hi].copy()
        d1v = pd1[lo:hi].copy() if need_d else None; d2v = (d1v if same else pd2[lo:hi].copy()) if need_d else None
        b.calcFrames(c1v, c2v, d1v, d2v, boxes=None if boxes is None else boxes[lo:hi])
    for lo, hi in ((0, 5), (5, nf)):
        c1v = pc1[lo:hi]; c2v = c1v if same else pc2[lo:hi]
        d1v = pd1[lo:hi] if need_d else None; d2v = (d1v if same else pd2[lo:hi]) if need_d else None
        c.calcFrames(c1v, c2v, d1v, d2v, boxes=None if boxes is None else boxes[lo:hi], pinned=True)
    ha, hb, hc, ref = a.raw_histogram(), b.raw_histogram(), c.raw_histogram(), o.raw_sum()
    assert np.array_equal(ha[:300], ref[:300]) and ref[:300].sum() > 1000
    assert np.array_equal(hb[:300], ref[:300]) and np.array_equal(hc[:300], ref[:300])
    for h in (ha, hb, hc):
        assert_hist_close(h, ref, 300)
    assert c.stats()["h2d_bytes"] == a.stats()["h2d_bytes"]         # the same bytes cross PCIe, without the host copy
    assert a.volume_list == o.volume_list and c.volume_list == o.volume_list
    with pytest.raises(ValueError, match="page-locked"):
        c.calcFrames(pc1[:2].copy(), pc2[:2].copy(), pd1[:2].copy() if need_d else None, pd2[:2].copy() if need_d else None, pinned=True)
    with pytest.raises(TypeError):
        c.calcFrames(pc1[:2].astype(np.float32), pc2[:2].astype(np.float32), pinned=True)


def test_device_resident_frames_equal_host_path():
    from newanalysis.gfunction import to_device
    n, L, nf = 3000, 45.0, 3
    rng = np.random.default_rng(17)
    c = rng.random((nf, n, 3)) * L; d = rng.standard_normal((nf, n, 3))
    a = RDF(["all"], 0.0, 15.0, 0.05, n, n, n, n, L); a.calcFrames(c, c, d, d)
    b = RDF(["all"], 0.0, 15.0, 0.05, n, n, n, n, L)
    dc, dd = to_device(c), to_device(d)
    b.calcFramesDevice(nf, dc, dc, dd, dd)
    ha, hb = a.raw_histogram(), b.raw_histogram()
    # same frames, same kernel: counts equal; the dense kernel (this box is dense) hands work units to warps dynamically,
    # so the order of the fp64 additions - not the set of addends - may differ between two runs
    assert np.array_equal(ha[:300], hb[:300]) and a.stats()["kernel"] == b.stats()["kernel"]
    assert_hist_close(hb, ha, 300)
    # legacy pre_*_gpu kwargs of calcFrame: one shared upload, two RDF objects (cfg4 pattern)
    e = RDF(["000"], 0.0, 15.0, 0.05, n, n, n, n, L); f = RDF(["all"], 0.0, 15.0, 0.05, n, n, n, n, L)
    pc, pd = to_device(c[0]), to_device(d[0])
    e.calcFrame(c[0], c[0], pre_coor_sel1_gpu=pc, pre_coor_sel2_gpu=pc)
    f.calcFrame(c[0], c[0], d[0], d[0], pre_coor_sel1_gpu=pc, pre_coor_sel2_gpu=pc, pre_dip_sel1_gpu=pd, pre_dip_sel2_gpu=pd)
    o = O.OracleRDF(["all"], 0.0, 15.0, 0.05, n, n, n, n, L); o.calcFrame(c[0], c[0], d[0], d[0])
    assert np.array_equal(e.raw_histogram()[:300], o.raw_sum()[:300])
    assert np.array_equal(f.raw_histogram()[:300], o.raw_sum()[:300])


def test_big_histogram_falls_back_to_global_atomics():
    n, L = 700, 60.0
    rng = np.random.default_rng(21)
    c, d = _random_frame(rng, n, L)
    g = RDF(["all"], 0.0, 29.0, 0.005, n, n, n, n, L)       # 5800 bins: replicas do not fit in shared memory
    assert g.stats()["hist_mode"] == "global_atomics"
    o = O.OracleRDF(["all"], 0.0, 29.0, 0.005, n, n, n, n, L)
    g.calcFrame(c, c, d, d); o.calcFrame(c, c, d, d)
    got, ref = g.raw_histogram(), o.raw_sum()
    assert np.array_equal(got[:5800], ref[:5800])
    assert_hist_close(got, ref, 5800)


def test_inrange_div_sqrt_sequences_equal_ieee_builtins():
    """The drain path's branch-free division / square-root sequences must return exactly what the
    correctly rounded built-ins return (bin assignment and per-pair increments depend on it)."""
    from newanalysis.gfunction import selftest_arith
    for seed in (1, 2, 3):
        assert selftest_arith(0, seed, 1 << 31) == (0, 0, 0)


def _ngpu():
    from newanalysis.gfunction import device_count
    return device_count()


@pytest.mark.skipif("_ngpu() < 2")
def test_single_process_two_devices_frame_sharding(monkeypatch):
    """NEWANALYSIS_GPUS=2: one handle deals frame batches round-robin to two GPUs and sums the per-device
    histograms on the host at readout (device order: deterministic)."""
    monkeypatch.setenv("GFN_BATCH_FRAMES", "2")
    n, L, nf = 900, 40.0, 9
    rng = np.random.default_rng(123)
    c = rng.random((nf, n, 3)) * L; d = rng.standard_normal((nf, n, 3))
    one = RDF(["all"], 0.0, 15.0, 0.05, n, n, n, n, L)
    two = RDF(["all"], 0.0, 15.0, 0.05, n, n, n, n, L, devices=[0, 1])
    assert two.stats()["ndev"] == 2
    for f in range(nf):
        one.calcFrame(c[f], c[f], d[f], d[f]); two.calcFrame(c[f], c[f], d[f], d[f])
    a, b = one.raw_histogram(), two.raw_histogram()
    assert np.array_equal(a[:300], b[:300])
    assert_hist_close(b, a, 300)
    two.calcFrame(c[0], c[0], d[0], d[0]); one.calcFrame(c[0], c[0], d[0], d[0])     # accumulation continues after a readout
    assert np.array_equal(one.raw_histogram()[:300], two.raw_histogram()[:300])


@pytest.mark.skipif("_ngpu() < 2")
@pytest.mark.parametrize("serial", [False, True])
def test_two_devices_block_and_page_locked_feeds(serial, monkeypatch):
    """One process, two devices: block calls staged by the copy pool and page-locked blocks whose batches are submitted by one
    worker per device (or, GFN_SERIAL_SUBMIT, by the calling thread); NpT boxes, several calls and readouts in between; the
    histograms of the devices are added on the host in device order."""
    from newanalysis.gfunction import pinned_empty
    monkeypatch.setenv("GFN_BATCH_FRAMES", "3")
    if serial:
        monkeypatch.setenv("GFN_SERIAL_SUBMIT", "1")
    n, L, nf = 1200, 60.0, 23
    rng = np.random.default_rng(321)
    c, d = pinned_empty((nf, n, 3)), pinned_empty((nf, n, 3))
    c[:] = rng.random((nf, n, 3)) * L; d[:] = rng.standard_normal((nf, n, 3))
    boxes = L * (1 + 2e-3 * rng.standard_normal((nf, 3)))
    a = RDF(["all"], 0.0, 14.0, 0.05, n, n, n, n, L, devices=[0, 1])
    b = RDF(["all"], 0.0, 14.0, 0.05, n, n, n, n, L, devices=[0, 1])
    o = O.OracleRDF(["all"], 0.0, 14.0, 0.05, n, n, n, n, L)
    for lo, hi in ((0, 9), (9, 10), (10, nf)):
        a.calcFrames(c[lo:hi], c[lo:hi], d[lo:hi], d[lo:hi], boxes=boxes[lo:hi], pinned=True)
        cc, dd = np.array(c[lo:hi]), np.array(d[lo:hi])
        b.calcFrames(cc, cc, dd, dd, boxes=boxes[lo:hi])
        for f in range(lo, hi):
            o.update_box(*boxes[f]); o.calcFrame(c[f], c[f], d[f], d[f])
        ha, hb, ref = a.raw_histogram(), b.raw_histogram(), o.raw_sum()
        assert np.array_equal(ha[:280], ref[:280]) and np.array_equal(hb[:280], ref[:280])
        assert_hist_close(ha, ref, 280); assert_hist_close(hb, ref, 280)
    assert a.stats()["frames"] == nf and a.stats()["ndev"] == 2 and a.volume_list == o.volume_list
    for r in (a, b):                                     # three calls, each dealt evenly (+- 1 frame)
        st = r.stats()
        assert st["frames_max_device"] - st["frames_min_device"] <= 3 and st["frames_min_device"] >= 9


@pytest.mark.skipif("_ngpu() < 2")
@pytest.mark.parametrize("pinned", [True, False])
def test_two_devices_block_call_is_dealt_evenly(pinned, monkeypatch):
    """A block call gives every device the same number of frames: dealing whole batches round-robin until the frames run
    out would leave device 0 a batch ahead (46 frames in batches of 5: 25 vs 21), and the call lasts as long as the
    fullest device."""
    from newanalysis.gfunction import pinned_empty
    monkeypatch.setenv("GFN_BATCH_FRAMES", "5")
    n, L, nf = 700, 50.0, 46
    rng = np.random.default_rng(99)
    c = pinned_empty((nf, n, 3)) if pinned else np.empty((nf, n, 3))
    c[:] = rng.random((nf, n, 3)) * L
    two = RDF(["000"], 0.0, 12.0, 0.05, n, n, n, n, L, devices=[0, 1])
    one = RDF(["000"], 0.0, 12.0, 0.05, n, n, n, n, L)
    two.calcFrames(c, c, None, None, pinned=pinned)
    cc = np.array(c)
    one.calcFrames(cc, cc, None, None)
    assert np.array_equal(two.raw_histogram()[:240], one.raw_histogram()[:240])
    st = two.stats()
    assert st["frames"] == nf and st["frames_max_device"] == 23 and st["frames_min_device"] == 23
    two.calcFrames(c[:7], c[:7], None, None, pinned=pinned)          # odd count: one device gets the extra frame
    two.sync()
    st = two.stats()
    assert st["frames"] == nf + 7 and st["frames_max_device"] - st["frames_min_device"] == 1


@pytest.mark.skipif("_ngpu() < 2")
def test_two_devices_device_resident_frames():
    """gfn_enqueue_device_frames runs the frames on the device that holds them: a two-device handle fed with frames
    uploaded to GPU 0 and GPU 1 equals the one-GPU run."""
    from newanalysis.gfunction import to_device
    n, L, nf = 3000, 45.0, 8
    rng = np.random.default_rng(92)
    c = rng.random((nf, n, 3)) * L; d = rng.standard_normal((nf, n, 3))
    one = RDF(["all"], 0.0, 15.0, 0.05, n, n, n, n, L)
    one.calcFrames(c, c, d, d)
    two = RDF(["all"], 0.0, 15.0, 0.05, n, n, n, n, L, devices=[0, 1])
    h = nf // 2
    c0, d0 = to_device(c[:h], 0), to_device(d[:h], 0)
    c1, d1 = to_device(c[h:], 1), to_device(d[h:], 1)
    two.calcFramesDevice(h, c0, c0, d0, d0)
    two.calcFramesDevice(nf - h, c1, c1, d1, d1)
    a, b = one.raw_histogram(), two.raw_histogram()
    assert np.array_equal(a[:300], b[:300]) and a[:300].sum() > 0
    assert_hist_close(b, a, 300)
    # frames on a device the handle does not own are refused
    with pytest.raises(ValueError, match="not one of this handle"):
        one.calcFramesDevice(1, c1, c1, d1, d1)


@pytest.mark.skipif("_ngpu() < 2")
def test_two_devices_atom_frames_and_trajectory(tmp_path, monkeypatch):
    """The topology is copied to every device of the handle: atom frames and a DCD trajectory sharded over two GPUs give
    the counts of the one-GPU run (weighted rows differ by the order of the final sum only)."""
    import dcd as D
    from newanalysis.gfunction import DCDFile
    monkeypatch.setenv("GFN_BATCH_FRAMES", "2")
    nres, L, nf = 1100, 36.0, 9
    pos, masses, charges, apr, rfa = _water_box(nres, L, nf, 83)
    path = str(tmp_path / "two.dcd")
    D.write_dcd(path, pos)
    sel = np.arange(3 * nres, dtype=np.int32)
    one = RDF(["all"], 0.0, 15.0, 0.05, nres, nres, nres, nres, L)
    one.attach_topology(masses, apr, rfa, charges)
    one.calcFramesAtoms(pos)
    ref = one.raw_histogram()
    for feed in ("atoms", "dcd"):
        two = RDF(["all"], 0.0, 15.0, 0.05, nres, nres, nres, nres, L, devices=[0, 1])
        two.attach_topology(masses, apr, rfa, charges)
        if feed == "atoms":
            two.calcFramesAtoms(pos)
        else:
            assert two.calcTrajectory(DCDFile(path), sel) == nf
        got = two.raw_histogram()
        assert two.stats()["ndev"] == 2
        assert np.array_equal(got[:300], ref[:300]) and ref[:300].sum() > 0
        assert_hist_close(got, ref, 300)


@pytest.mark.skipif("_ngpu() < 2")
def test_multi_process_write_normalises_over_all_ranks(tmp_path):
    """One process per GPU (torchrun, 2 ranks, NCCL): write() on every rank gives the g-functions of the whole run -
    histogram summed by ncclAllReduce, frame count and mean volume summed with gfn_allreduce_host."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    import mp_write_check as M
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                    "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(here, "mp_write_check.py"), str(tmp_path)],
                   check=True, timeout=180)
    n, L, nf = 1500, 40.0, 11
    c, d, boxes = M.frames(nf, n, L)
    o = O.OracleRDF(["all"], 0.0, 15.0, 0.05, n, n, n, n, L)
    for f in range(nf):
        o.update_box(*[float(v) for v in boxes[f]])
        o.calcFrame(c[f], c[f], d[f], d[f])
    o.write(str(tmp_path / "oracle"))
    for rank in (0, 1):
        got = np.load(str(tmp_path / ("rank%d_out.npy" % rank)))
        assert np.allclose(got[:300], o.histogram_out[:300], rtol=1e-13, atol=0)         # counts exact; mean volume summed in another order
        assert np.allclose(got, o.histogram_out, rtol=1e-9, atol=1e-10)
        assert (tmp_path / ("rank%d_g000.dat" % rank)).read_text() == (tmp_path / "oracle_g000.dat").read_text()


# ---- N1: centre of mass / dipole producers on the device ------------------------------------------------
@pytest.mark.parametrize("name", sorted(C.RESIDUE_CASES))
def test_residue_producers_bit_exact(golden_dir, name):
    from newanalysis.helpers import comByResidue, dipByResidue, velcomByResidue
    z = np.load(os.path.join(golden_dir, "residues.npz"))
    coor, masses, charges, apr, rfa = C.build_residues(C.RESIDUE_CASES[name])
    nres = apr.shape[0]
    com = comByResidue(coor, masses, nres, apr, rfa)
    assert np.array_equal(com, z[name + "_com"], equal_nan=True)
    assert np.array_equal(dipByResidue(coor, charges, masses, nres, apr, rfa, com), z[name + "_dip"], equal_nan=True)
    assert np.array_equal(velcomByResidue(coor * 0.37, masses, nres, apr, rfa), z[name + "_vel"], equal_nan=True)


def test_atoms_to_gfunction_on_device():
    """Raw atom coordinates -> (device) centres of mass + dipoles -> RDF, without a host round trip, equals the
    host pipeline of the reference (comByResidue, dipByResidue, calcFrame) bit for bit in the counts."""
    from newanalysis.gfunction import ResidueMap, to_device
    nres, L, nf = 2000, 40.0, 3
    rng = np.random.default_rng(31)
    apr = np.full(nres, 3, dtype=np.int32); rfa = (3 * np.arange(nres)).astype(np.int32)
    masses = np.tile([15.999, 1.008, 1.008], nres); charges = np.tile([-0.8476, 0.4238, 0.4238], nres)
    centre = rng.random((nf, nres, 1, 3)) * L
    coor = np.ascontiguousarray((centre + rng.standard_normal((nf, nres, 3, 3)) * 0.6).reshape(nf, 3 * nres, 3))
    rm = ResidueMap(masses, apr, rfa, charges)
    com, dip = rm.com_dip(coor)
    for f in range(nf):
        assert np.array_equal(com[f], O.com_by_residue_c(coor[f], masses, apr, rfa))
        assert np.array_equal(dip[f], O.dip_by_residue_c(coor[f], charges, apr, rfa, com[f]))
    dcom, ddip = rm.com_dip_device(to_device(coor))
    a = RDF(["all"], 0.0, 15.0, 0.05, nres, nres, nres, nres, L)
    a.calcFramesDevice(nf, dcom, dcom, ddip, ddip)
    o = O.OracleRDF(["all"], 0.0, 15.0, 0.05, nres, nres, nres, nres, L)
    for f in range(nf):
        o.calcFrame(com[f], com[f], dip[f], dip[f])
    got, ref = a.raw_histogram(), o.raw_sum()
    assert np.array_equal(got[:300], ref[:300]) and ref[:300].sum() > 0
    assert_hist_close(got, ref, 300)


def test_function_wrappers_from_psf_selection(tmp_path):
    """newanalysis.functions.centerOfMassByResidue / dipoleMomentByResidue (py_functions.py:91-150) on a selection of a
    PSF topology, producers on the GPU: bit-identical to the oracle's restatement of helpers.pyx:71-123."""
    from newanalysis import psf, functions as F
    text, _ = C.build_psf(C.PSF_CASES["ionic"])
    path = tmp_path / "i.psf"
    path.write_text(text)
    sel = psf.PSF(str(path)).select(resname="C2MI")
    rng = np.random.default_rng(77)
    sel.positions = (rng.random((sel.n_atoms, 3)) * 40).astype(np.float32)
    apr, rfa = F.atomsPerResidue(sel), F.residueFirstAtom(sel)
    coor = sel.positions.astype("float64")
    com = F.centerOfMassByResidue(sel, apr=apr, rfa=rfa)
    dip = F.dipoleMomentByResidue(sel, com=com, apr=apr, rfa=rfa)
    ref_com = O.com_by_residue_c(coor, sel.masses, apr, rfa)
    assert np.array_equal(com, ref_com)
    assert np.array_equal(dip, O.dip_by_residue_c(coor, sel.charges, apr, rfa, ref_com))
    assert np.array_equal(F.dipoleMomentByResidue(sel, apr=apr, rfa=rfa), dip)


# ---- N1 + N4: raw float32 atom positions through the upload ring --------------------------------------------
def _water_box(nres, L, nf, seed):
    rng = np.random.default_rng(seed)
    apr = np.full(nres, 3, dtype=np.int32); rfa = (3 * np.arange(nres)).astype(np.int32)
    masses = np.tile([15.999, 1.008, 1.008], nres); charges = np.tile([-0.8476, 0.4238, 0.4238], nres)
    centre = rng.random((nf, nres, 1, 3)) * L
    pos = (centre + rng.standard_normal((nf, nres, 3, 3)) * 0.6).reshape(nf, 3 * nres, 3).astype(np.float32)
    return np.ascontiguousarray(pos), masses, charges, apr, rfa


def _reference_chain(pos32, masses, charges, apr, rfa):
    """What an analysis script does per frame before RDF.calcFrame: positions.astype(float64) -> comByResidue ->
    dipByResidue (helpers.pyx:71-123), here through the oracle's C restatements."""
    c64 = np.ascontiguousarray(pos32.astype(np.float64))
    com = O.com_by_residue_c(c64, masses, apr, rfa)
    return com, O.dip_by_residue_c(c64, charges, apr, rfa, com)


@pytest.mark.parametrize("flags", ["", "fp32", "fp64"])
def test_atom_frames_self_equal_reference_chain(flags):
    nres, L, nf = 3000, 45.0, 5
    pos, masses, charges, apr, rfa = _water_box(nres, L, nf, 41)
    a = RDF(["all"], 0.0, 15.0, 0.05, nres, nres, nres, nres, L, gfn_flags=flags or None)
    a.attach_topology(masses, apr, rfa, charges)
    a.calcFrameAtoms(pos[0])                 # one frame per call ...
    a.calcFramesAtoms(pos[1:])               # ... and the rest in one call
    host = RDF(["all"], 0.0, 15.0, 0.05, nres, nres, nres, nres, L, gfn_flags=flags or None)
    o = O.OracleRDF(["all"], 0.0, 15.0, 0.05, nres, nres, nres, nres, L)
    for f in range(nf):
        com, dip = _reference_chain(pos[f], masses, charges, apr, rfa)
        host.calcFrame(com, com, dip, dip)
        o.calcFrame(com, com, dip, dip)
    got, ref = a.raw_histogram(), o.raw_sum()
    assert a.oneistwo is True and a.ctr == nf
    assert np.array_equal(got[:300], ref[:300]) and ref[:300].sum() > 0
    assert_hist_close(got, ref, 300)
    # the device-made centres of mass / dipoles are bit-identical to the host-made ones: counts are equal, and the fp64
    # sums hold the same addends (their order follows the launch's frame batch, which differs between the two feeds)
    hh = host.raw_histogram()
    assert np.array_equal(got[:300], hh[:300])
    assert_hist_close(got, hh, 300)
    st = a.stats()
    assert st["h2d_bytes"] == nf * (3 * nres * 3 * 4 + 48)      # float32 atoms once per frame + the box


def test_atom_frames_two_selections_ragged_and_npt():
    """Cations x anions with ragged residues (2..9 atoms; a single atom has a zero dipole, quirk Q5), a different topology per selection, per-frame boxes."""
    rng = np.random.default_rng(43)
    n1, n2, L, nf = 700, 500, 50.0, 4

    def topo(n):
        apr = rng.integers(2, 10, n).astype(np.int32)
        rfa = np.concatenate([[0], np.cumsum(apr)[:-1]]).astype(np.int32)
        na = int(apr.sum())
        return apr, rfa, 1.0 + 15.0 * rng.random(na), rng.standard_normal(na), na
    apr1, rfa1, m1, q1, na1 = topo(n1)
    apr2, rfa2, m2, q2, na2 = topo(n2)

    def frames(apr, na, n):
        centre = np.repeat(rng.random((nf, n, 3)) * L, apr, axis=1)
        return np.ascontiguousarray((centre + 0.8 * rng.standard_normal((nf, na, 3))).astype(np.float32))
    p1, p2 = frames(apr1, na1, n1), frames(apr2, na2, n2)
    boxes = L * (1.0 + 1e-3 * rng.standard_normal((nf, 3)))
    modes = ["000", "110", "101", "011", "220"]
    a = RDF(modes, 0.0, 20.0, 0.05, n1, n2, n1, n2, L)
    a.attach_topology(m1, apr1, rfa1, q1, m2, apr2, rfa2, q2)
    a.calcFramesAtoms(p1, p2, boxes=boxes)
    o = O.OracleRDF(modes, 0.0, 20.0, 0.05, n1, n2, n1, n2, L)
    for f in range(nf):
        o.update_box(*boxes[f])
        c1, d1 = _reference_chain(p1[f], m1, q1, apr1, rfa1)
        c2, d2 = _reference_chain(p2[f], m2, q2, apr2, rfa2)
        o.calcFrame(c1, c2, d1, d2)
    got, ref = a.raw_histogram(), o.raw_sum()
    assert a.oneistwo is False
    assert np.array_equal(got[:400], ref[:400]) and ref[:400].sum() > 0
    assert_hist_close(got, ref, 400)
    assert a.volume_list == o.volume_list


def test_atom_frames_mix_with_host_frames_and_small_batches(monkeypatch):
    """Atom frames and host frames may alternate on one handle (a layout change closes the batch); tiny batches make the
    ring recycle its atom buffers."""
    monkeypatch.setenv("GFN_BATCH_FRAMES", "2")
    nres, L, nf = 1200, 36.0, 9
    pos, masses, charges, apr, rfa = _water_box(nres, L, nf, 47)
    a = RDF(["all"], 0.0, 15.0, 0.05, nres, nres, nres, nres, L)
    a.attach_topology(masses, apr, rfa, charges)
    o = O.OracleRDF(["all"], 0.0, 15.0, 0.05, nres, nres, nres, nres, L)
    for f in range(nf):
        com, dip = _reference_chain(pos[f], masses, charges, apr, rfa)
        o.calcFrame(com, com, dip, dip)
        if f % 3 == 1:
            a.calcFrame(com, com, dip, dip)
        else:
            a.calcFrameAtoms(pos[f])
    got, ref = a.raw_histogram(), o.raw_sum()
    assert np.array_equal(got[:300], ref[:300])
    assert_hist_close(got, ref, 300)


def test_atom_frames_g011_only_and_g000_without_charges():
    """mode 011 needs only the dipoles of selection 2 (the frame is then not aliased); g000 needs no charges at all."""
    nres, L, nf = 900, 33.0, 2
    pos, masses, charges, apr, rfa = _water_box(nres, L, nf, 53)
    for modes, q in ((["011"], charges), (["000"], None)):
        a = RDF(modes, 0.0, 15.0, 0.05, nres, nres, nres, nres, L)
        a.attach_topology(masses, apr, rfa, q)
        a.calcFramesAtoms(pos)
        o = O.OracleRDF(modes, 0.0, 15.0, 0.05, nres, nres, nres, nres, L)
        for f in range(nf):
            com, dip = _reference_chain(pos[f], masses, charges, apr, rfa)
            o.calcFrame(com, com, dip, dip)
        got, ref = a.raw_histogram(), o.raw_sum()
        assert_hist_close(got, ref, 300, pairw=a.pair_weight)
        assert np.array_equal(a.pair_weight, O_pair_weight(o, nres, L, pos, masses, charges, apr, rfa))


def O_pair_weight(o, nres, L, pos, masses, charges, apr, rfa):
    g = O.OracleRDF(["000"], 0.0, 15.0, 0.05, nres, nres, nres, nres, L)
    for f in range(pos.shape[0]):
        com, dip = _reference_chain(pos[f], masses, charges, apr, rfa)
        g.calcFrame(com, com)
    return g.raw_sum()[:300]


def test_atom_frames_argument_errors():
    nres, L = 300, 30.0
    pos, masses, charges, apr, rfa = _water_box(nres, L, 1, 59)
    a = RDF(["all"], 0.0, 15.0, 0.05, nres, nres, nres, nres, L)
    with pytest.raises(RuntimeError, match="attach_topology"):
        a.calcFrameAtoms(pos[0])
    with pytest.raises(ValueError, match="residues"):
        a.attach_topology(masses[:-3], apr[:-1], rfa[:-1], charges[:-3])
    a.attach_topology(masses, apr, rfa)                       # no charges, but "all" needs dipoles
    with pytest.raises(ValueError, match="charges"):
        a.calcFrameAtoms(pos[0])
    a = RDF(["all"], 0.0, 15.0, 0.05, nres, nres, nres, nres, L)
    a.attach_topology(masses, apr, rfa, charges)
    with pytest.raises(ValueError, match="shapes"):
        a.calcFrameAtoms(pos[0, :-1])
    b = RDF(["000"], 0.0, 15.0, 0.05, nres, nres + 1, nres, nres + 1, L)
    b.attach_topology(masses, apr, rfa)
    with pytest.raises(ValueError, match="second selection"):
        b.calcFrameAtoms(pos[0])


# ---- N4: DCD trajectory -> reader threads -> upload ring ----------------------------------------------------
def _solvated_system(nwat, nsol_atoms, L, nf, seed):
    """A solute of nsol_atoms atoms followed by nwat 3-site waters; returns float32 positions (nf, natoms, 3)."""
    rng = np.random.default_rng(seed)
    wat, masses, charges, apr, rfa = _water_box(nwat, L, nf, seed + 1)
    sol = (L / 2 + 2.0 * rng.standard_normal((nf, nsol_atoms, 3))).astype(np.float32)
    pos = np.ascontiguousarray(np.concatenate([sol, wat], axis=1))
    return pos, masses, charges, apr, rfa


@pytest.mark.parametrize("big_endian,threads", [(False, 1), (True, 3)])
def test_trajectory_self_npt_equals_atom_frames_and_oracle(tmp_path, big_endian, threads):
    import dcd as D
    from newanalysis.gfunction import DCDFile
    nwat, nsol, L, nf = 2000, 17, 40.0, 23
    pos, masses, charges, apr, rfa = _solvated_system(nwat, nsol, L, nf, 61)
    rng = np.random.default_rng(62)
    cells = np.concatenate([L * (1 + 1e-3 * rng.standard_normal((nf, 3))), np.full((nf, 3), 90.0)], axis=1)
    path = str(tmp_path / "npt.dcd")
    D.write_dcd(path, pos, cells, big_endian=big_endian)
    sel = np.arange(nsol, nsol + 3 * nwat, dtype=np.int32)
    first, last, step = 2, 21, 3
    frames = list(range(first, last, step))

    a = RDF(["all"], 0.0, 15.0, 0.05, nwat, nwat, nwat, nwat, L)
    a.attach_topology(masses, apr, rfa, charges)
    traj = DCDFile(path)
    assert a.calcTrajectory(traj, sel, first=first, last=last, step=step, use_cell=True, threads=threads) == len(frames)

    b = RDF(["all"], 0.0, 15.0, 0.05, nwat, nwat, nwat, nwat, L)
    b.attach_topology(masses, apr, rfa, charges)
    b.calcFramesAtoms(pos[frames][:, sel], boxes=cells[frames, :3])
    o = O.OracleRDF(["all"], 0.0, 15.0, 0.05, nwat, nwat, nwat, nwat, L)
    ora = D.DCD(path)
    for f in frames:
        xyz, cell = ora.read(f)
        o.update_box(*cell[:3])
        com, dip = _reference_chain(xyz[sel], masses, charges, apr, rfa)
        o.calcFrame(com, com, dip, dip)
    got, ref = a.raw_histogram(), o.raw_sum()
    hb = b.raw_histogram()
    assert np.array_equal(got[:300], hb[:300])          # same frames: equal counts, sums in batch-dependent order
    assert_hist_close(got, hb, 300)
    assert np.array_equal(got[:300], ref[:300]) and ref[:300].sum() > 0
    assert_hist_close(got, ref, 300)
    assert a.ctr == len(frames) and a.oneistwo is True
    assert a.volume_list == o.volume_list and a.volume == o.volume
    assert np.array_equal(a.box_dim, o.box_dim)


def test_trajectory_solute_solvent_fixed_box_and_float32_cells(tmp_path):
    """rdf.ipynb's case: one solute molecule x all waters, box from the constructor; then the NpT variant with the
    cell lengths rounded through float32 (what a script reading ts.dimensions would pass to update_box)."""
    import dcd as D
    from newanalysis.gfunction import DCDFile
    nwat, nsol, L, nf = 1500, 24, 36.0, 12
    pos, masses, charges, apr, rfa = _solvated_system(nwat, nsol, L, nf, 67)
    rng = np.random.default_rng(68)
    msol = 1.0 + 12.0 * rng.random(nsol); qsol = rng.standard_normal(nsol)
    cells = np.concatenate([L * (1 + 1e-3 * rng.standard_normal((nf, 3))), np.full((nf, 3), 90.0)], axis=1)
    path = str(tmp_path / "sol.dcd")
    D.write_dcd(path, pos, cells)
    s_sol = np.arange(nsol, dtype=np.int32); s_wat = np.arange(nsol, nsol + 3 * nwat, dtype=np.int32)
    one = np.array([nsol], dtype=np.int32); zero = np.array([0], dtype=np.int32)
    for use_cell in (False, "float32"):
        a = RDF(["all"], 0.0, 15.0, 0.05, 1, nwat, 1, nwat, L)
        a.attach_topology(msol, one, zero, qsol, masses, apr, rfa, charges)
        a.calcTrajectory(DCDFile(path), s_sol, s_wat, use_cell=use_cell)
        o = O.OracleRDF(["all"], 0.0, 15.0, 0.05, 1, nwat, 1, nwat, L)
        for f in range(nf):
            if use_cell:
                o.update_box(*[float(np.float32(v)) for v in cells[f, :3]])
            c1, d1 = _reference_chain(pos[f][s_sol], msol, qsol, one, zero)
            c2, d2 = _reference_chain(pos[f][s_wat], masses, charges, apr, rfa)
            o.calcFrame(c1, c2, d1, d2)
        got, ref = a.raw_histogram(), o.raw_sum()
        assert np.array_equal(got[:300], ref[:300]) and ref[:300].sum() > 0
        assert_hist_close(got, ref, 300)
        assert a.volume_list == o.volume_list and a.oneistwo is False


def test_trajectory_argument_errors(tmp_path):
    import dcd as D
    from newanalysis.gfunction import DCDFile
    nwat, L, nf = 200, 30.0, 3
    pos, masses, charges, apr, rfa = _solvated_system(nwat, 5, L, nf, 71)
    p_nocell, p_tric = str(tmp_path / "a.dcd"), str(tmp_path / "b.dcd")
    D.write_dcd(p_nocell, pos)
    D.write_dcd(p_tric, pos, np.tile([L, L, L, 90.0, 90.0, 120.0], (nf, 1)))
    sel = np.arange(5, 5 + 3 * nwat, dtype=np.int32)
    a = RDF(["all"], 0.0, 15.0, 0.05, nwat, nwat, nwat, nwat, L)
    with pytest.raises(RuntimeError, match="attach_topology"):
        a.calcTrajectory(DCDFile(p_nocell), sel)
    a.attach_topology(masses, apr, rfa, charges)
    with pytest.raises(ValueError, match="selection 1 has"):
        a.calcTrajectory(DCDFile(p_nocell), sel[:-1])
    bad = sel.copy(); bad[7] = 5 + 3 * nwat
    with pytest.raises(ValueError, match="out of range"):
        a.calcTrajectory(DCDFile(p_nocell), bad)
    with pytest.raises(ValueError, match="no unit cell"):
        a.calcTrajectory(DCDFile(p_nocell), sel, use_cell=True)
    with pytest.raises(ValueError, match="orthorhombic"):
        a.calcTrajectory(DCDFile(p_tric), sel, use_cell=True)
    assert a.calcTrajectory(DCDFile(p_nocell), sel, first=5) == 0
    # nothing of the failed calls may have reached the histogram except frames before the failing one (none here)
    assert a.calcTrajectory(DCDFile(p_tric), sel) == nf            # cell ignored: fixed box
    b = RDF(["all"], 0.0, 15.0, 0.05, nwat, nwat, nwat, nwat, L)
    b.attach_topology(masses, apr, rfa, charges)
    b.calcFramesAtoms(pos[:, sel])
    assert np.array_equal(a.raw_histogram(), b.raw_histogram())


@pytest.mark.parametrize("feed", ["atoms", "dcd"])
@pytest.mark.parametrize("name", sorted(C.ATOM_CASES))
def test_atom_frames_golden(golden_dir, tmp_path, name, feed):
    """Fixture tests/golden/atoms.npz: positions.astype(float64) -> helpers.comByResidue -> helpers.dipByResidue ->
    RDF.calcFrame, all by the UNMODIFIED reference modules.  The GPU takes the float32 positions (or a DCD file
    holding them) and must reproduce the reference's un-normalised sums: counts bit-exact, weighted rows within 1e-12."""
    import dcd as D
    from newanalysis.gfunction import DCDFile
    case = C.ATOM_CASES[name]
    z = np.load(os.path.join(golden_dir, "atoms.npz"))
    a = C.build_atoms(case)
    n1 = case["n1"]; n2 = n1 if case["same"] else case["n2"]
    r = RDF(case["modes"], 0.0, case["hmax"], 0.05, n1, n2, n1, n2, case["L"])
    hn = int(r.histo_n)
    if a["topo2"] is None:
        r.attach_topology(*a["topo1"])
    else:
        r.attach_topology(*a["topo1"], *a["topo2"])
    if feed == "atoms":
        r.calcFramesAtoms(a["pos1"], a["pos2"], boxes=a["boxes"])
    else:
        na1 = a["pos1"].shape[1]
        pos = a["pos1"] if a["pos2"] is None else np.concatenate([a["pos1"], a["pos2"]], axis=1)
        cells = None if a["boxes"] is None else np.concatenate([a["boxes"], np.full_like(a["boxes"], 90.0)], axis=1)
        path = str(tmp_path / "g.dcd")
        D.write_dcd(path, pos, cells)
        s1 = np.arange(na1, dtype=np.int32)
        s2 = None if a["pos2"] is None else np.arange(na1, pos.shape[1], dtype=np.int32)
        r.calcTrajectory(DCDFile(path), s1, s2, use_cell=cells is not None)
    got, ref = r.raw_histogram(), z[name + "_raw"]
    assert got.shape == ref.shape
    assert np.array_equal(got[:hn], ref[:hn]) and ref[:hn].sum() > 0
    assert_hist_close(got, ref, hn, pairw=np.maximum(r.pair_weight, ref[:hn]))
    assert np.array_equal(np.array(r.volume_list), z[name + "_volumes"])
    assert bool(r.oneistwo) == bool(z[name + "_oneistwo"])


def _cfg4_box(rng, ns, nu, apu, L):
    """A cfg4-shaped frame: ns solvent molecules (COM + dipole, and an O site each), nu solute molecules
    (COM + dipole) of apu atoms each."""
    solv_c, solv_d = _random_frame(rng, ns, L)
    solv_o = np.ascontiguousarray(solv_c + 0.06 * rng.standard_normal((ns, 3)))
    solu_c, solu_d = _random_frame(rng, nu, L)
    solu_a = np.ascontiguousarray(np.repeat(solu_c, apu, axis=0) + 1.5 * rng.standard_normal((nu * apu, 3)))
    return solv_c, solv_d, solv_o, solu_c, solu_d, solu_a


def test_device_arrays_reused_frame_after_frame():
    """The upload slots of the pre_*_gpu path are overwritten every frame while earlier frames are still queued: upload()
    waits for the prep kernels that read the slot (not for the whole device), so no frame may see its successor's data."""
    from newanalysis.gfunction import to_device
    n, L, nf = 12000, 70.0, 7
    rng = np.random.default_rng(808)
    frames = [(rng.random((n, 3)) * L, rng.standard_normal((n, 3))) for _ in range(nf)]
    a = RDF(["all"], 0.0, 15.0, 0.05, n, n, n, n, L)
    b = RDF(["000"], 0.0, 15.0, 0.05, n, n, n, n, L)
    gc, gd = to_device(frames[0][0]), to_device(frames[0][1])
    o = O.OracleRDF(["all"], 0.0, 15.0, 0.05, n, n, n, n, L)
    for c, d in frames:
        gc.upload(c); gd.upload(d)
        a.calcFrame(c, c, d, d, pre_coor_sel1_gpu=gc, pre_coor_sel2_gpu=gc, pre_dip_sel1_gpu=gd, pre_dip_sel2_gpu=gd)
        b.calcFrame(c, c, pre_coor_sel1_gpu=gc, pre_coor_sel2_gpu=gc)
        o.calcFrame(c, c, d, d)
    ref = o.raw_sum()
    ha, hb = a.raw_histogram(), b.raw_histogram()
    assert np.array_equal(ha[:300], ref[:300]) and np.array_equal(hb[:300], ref[:300]) and ref[:300].sum() > 1e5
    assert_hist_close(ha, ref, 300)


def test_cfg4_four_objects_share_one_upload():
    """BASELINE configs[3] as north_star states it: four RDF objects per frame on ONE shared upload of the frame's arrays
    (the reference's pre_*_gpu kwargs, gfunction.pyx:390-391, 431-453): solute COM x solvent COM "all", solvent self
    "all", solute atoms x solvent O "000", solute self "all" - reduced sizes, three frames, against OracleRDF."""
    from newanalysis.gfunction import to_device
    ns, nu, apu, L = 3000, 96, 4, 48.0
    mk = lambda cls, **kw: [cls(["all"], 0.0, 15.0, 0.05, nu, ns, nu, ns, L, **kw),              # noqa: E731
                            cls(["all"], 0.0, 15.0, 0.05, ns, ns, ns, ns, L, **kw),
                            cls(["000"], 0.0, 15.0, 0.05, nu * apu, ns, nu, ns, L, **kw),
                            cls(["all"], 0.0, 15.0, 0.05, nu, nu, nu, nu, L, **kw)]
    gpu, cpu = mk(RDF), mk(O.OracleRDF)
    rng = np.random.default_rng(4004)
    for f in range(3):
        solv_c, solv_d, solv_o, solu_c, solu_d, solu_a = _cfg4_box(rng, ns, nu, apu, L)
        # one upload per array and frame, shared by the four objects
        g_solv_c, g_solv_d, g_solv_o = to_device(solv_c), to_device(solv_d), to_device(solv_o)
        g_solu_c, g_solu_d, g_solu_a = to_device(solu_c), to_device(solu_d), to_device(solu_a)
        gpu[0].calcFrame(solu_c, solv_c, solu_d, solv_d, pre_coor_sel1_gpu=g_solu_c, pre_coor_sel2_gpu=g_solv_c,
                         pre_dip_sel1_gpu=g_solu_d, pre_dip_sel2_gpu=g_solv_d)
        gpu[1].calcFrame(solv_c, solv_c, solv_d, solv_d, pre_coor_sel1_gpu=g_solv_c, pre_coor_sel2_gpu=g_solv_c,
                         pre_dip_sel1_gpu=g_solv_d, pre_dip_sel2_gpu=g_solv_d)
        gpu[2].calcFrame(solu_a, solv_o, pre_coor_sel1_gpu=g_solu_a, pre_coor_sel2_gpu=g_solv_o)
        gpu[3].calcFrame(solu_c, solu_c, solu_d, solu_d, pre_coor_sel1_gpu=g_solu_c, pre_coor_sel2_gpu=g_solu_c,
                         pre_dip_sel1_gpu=g_solu_d, pre_dip_sel2_gpu=g_solu_d)
        cpu[0].calcFrame(solu_c, solv_c, solu_d, solv_d)
        cpu[1].calcFrame(solv_c, solv_c, solv_d, solv_d)
        cpu[2].calcFrame(solu_a, solv_o)
        cpu[3].calcFrame(solu_c, solu_c, solu_d, solu_d)
    for k, (g, c) in enumerate(zip(gpu, cpu)):
        got, ref = g.raw_histogram(), c.raw_sum()
        assert bool(g.oneistwo) == bool(c.oneistwo) and g.ctr == c.ctr == 3, k
        assert np.array_equal(got[:300], ref[:300]) and ref[:300].sum() > 100, "object %d: g000 not bit-exact" % k
        assert_hist_close(got, ref, 300, pairw=np.maximum(g.pair_weight, ref[:300]))
        assert g.stats()["h2d_bytes"] < 1000, "object %d copied frame data although every array was shared" % k


def test_cfg4_solvent_self_full_size():
    """BASELINE configs[3], the dominant object at its real size: 81,920 solvent molecules, self, all modes, L = 143.6 A
    (3.36e9 pair evaluations) against the row-block CPU oracle."""
    n, L = 81920, 143.6
    rng = np.random.default_rng(4040)
    c, d = _random_frame(rng, n, L)
    rdf = RDF(["all"], 0.0, 15.0, 0.05, n, n, n, n, L)
    rdf.calcFrame(c, c, d, d)
    got = rdf.raw_histogram()
    ref = np.zeros(7 * 300)
    box = np.array([L, L, L, L / 2, L / 2, L / 2])
    rc, oob = O.calc_histo_rows_c(c, c, d, d, ref, box, 0.0, 15.0, 300, 20.0, 127)
    assert rc == 0 and oob == 0
    assert np.array_equal(got[:300], ref[:300]) and ref[:300].sum() > 1e7
    scale = np.maximum(np.abs(ref), np.tile(ref[:300], 7))
    assert np.all(np.abs(got - ref) <= 2 * RTOL * scale)


def test_cfg5_row_block_against_oracle():
    """BASELINE configs[4] shape (1M-site box, L = 315.4 A): the rectangular row block BASELINE.md section 3 names -
    1,024 core sites against ALL 1,048,576 surround sites of the box (gfunction.pyx:126-130, rectangular path) - vs
    the CPU oracle (1.07e9 pair evaluations)."""
    n1, n2, L = 1024, 1048576, 315.4
    rng = np.random.default_rng(5005)
    c1, d1 = _random_frame(rng, n1, L); c2, d2 = _random_frame(rng, n2, L)
    rdf = RDF(["all"], 0.0, 15.0, 0.05, n1, n2, n1, n2, L)
    rdf.calcFrame(c1, c2, d1, d2)
    got = rdf.raw_histogram()
    ref = np.zeros(7 * 300)
    box = np.array([L, L, L, L / 2, L / 2, L / 2])
    rc, oob = O.calc_histo_rows_c(c1, c2, d1, d2, ref, box, 0.0, 15.0, 300, 20.0, 127)
    assert rc == 0 and oob == 0
    assert np.array_equal(got[:300], ref[:300]) and ref[:300].sum() > 1000
    scale = np.maximum(np.abs(ref), np.tile(ref[:300], 7))
    assert np.all(np.abs(got - ref) <= 2 * RTOL * scale)


def test_filter_auto_selection_and_fp16_range_guard():
    """Default handles pick the packed-fp16 pre-filter for in-box coordinates and keep fp32 for far-away
    (unwrapped) ones; an explicit fp16 handle refuses coordinates it cannot represent instead of losing pairs."""
    n, L = 300, 30.0
    rng = np.random.default_rng(77)
    c = rng.random((n, 3)) * L; d = rng.standard_normal((n, 3))
    a = RDF(["all"], 0.0, 14.0, 0.05, n, n, n, n, L); a.calcFrame(c, c, d, d)
    assert a.stats()["filter"] == "fp16x2"
    far = c + 100 * L
    b = RDF(["all"], 0.0, 14.0, 0.05, n, n, n, n, L); b.calcFrame(far, far, d, d)
    assert b.stats()["filter"] == "fp32"
    o = O.OracleRDF(["all"], 0.0, 14.0, 0.05, n, n, n, n, L); o.calcFrame(far, far, d, d)
    assert np.array_equal(b.raw_histogram()[:280], o.raw_sum()[:280])
    # fp16 forced on data 100 box lengths away: still exact (the guard band widens with max|x|)
    e = RDF(["all"], 0.0, 14.0, 0.05, n, n, n, n, L, gfn_flags="fp16"); e.calcFrame(far, far, d, d)
    assert np.array_equal(e.raw_histogram()[:280], o.raw_sum()[:280])
    # ... and refused beyond what fp16 can hold
    g = RDF(["all"], 0.0, 14.0, 0.05, n, n, n, n, L, gfn_flags="fp16"); g.calcFrame(c + 1e6 * L, c + 1e6 * L, d, d)
    with pytest.raises(ValueError, match="fp16"):
        g.sync()


# ---- N3: RDF time series ------------------------------------------------------------------------------------
def test_calc_histo_ts_matches_oracle():
    """calcHistoTS (gfunction.pyx:168-192): per-frame, per-core-site g000 rows; counts are exact."""
    from newanalysis.gfunction import calcHistoTS
    n1, n2, nb, L, nf = 150, 400, 240, 30.0, 3
    rng = np.random.default_rng(41)
    res = np.zeros((nf, n1, nb)); ref = np.zeros((nf, n1, nb))
    for t in range(nf):
        c1 = rng.random((n1, 3)) * L; c2 = rng.random((n2, 3)) * L
        calcHistoTS(c1, c2, res, L, t, 0.0, 12.0, 20.0)
        O.calc_histo_ts_c(c1, c2, ref[t], L, 0.0, 12.0, 20.0)
    assert np.array_equal(res, ref) and ref.sum() > 1000
    # same selection on both sides: every ordered pair counts once, i == j is rejected by the first-bin rule
    c = rng.random((200, 3)) * L
    r2 = np.zeros((1, 200, nb)); o2 = np.zeros((1, 200, nb))
    calcHistoTS(c, c, r2, L, 0, 0.0, 12.0, 20.0); O.calc_histo_ts_c(c, c, o2[0], L, 0.0, 12.0, 20.0)
    assert np.array_equal(r2, o2)


def test_calc_histo_ts_adds_into_result_and_takes_strided_results():
    """The reference adds (`result[t,i,pos] += 1.0`, gfunction.pyx:192): what is already in `result` stays, a second call on
    the same frame doubles the counts, and a result array that is not C-contiguous gets the same rows."""
    from newanalysis.gfunction import calcHistoTS
    n1, n2, nb, L = 96, 700, 240, 28.0
    rng = np.random.default_rng(43)
    c1 = rng.random((n1, 3)) * L; c2 = rng.random((n2, 3)) * L
    want = np.zeros((n1, nb)); O.calc_histo_ts_c(c1, c2, want, L, 0.0, 12.0, 20.0)
    res = np.full((2, n1, nb), 3.0)
    calcHistoTS(c1, c2, res, L, 1, 0.0, 12.0, 20.0)
    calcHistoTS(c1, c2, res, L, 1, 0.0, 12.0, 20.0)
    assert np.array_equal(res[1], 3.0 + 2.0 * want) and np.array_equal(res[0], np.full((n1, nb), 3.0))
    wide = np.zeros((2, n1, 2 * nb))
    view = wide[:, :, ::2]
    assert not view.flags.c_contiguous
    calcHistoTS(c1, c2, view, L, 0, 0.0, 12.0, 20.0)
    assert np.array_equal(wide[0, :, ::2], want) and not wide[:, :, 1::2].any() and not wide[1].any()


def test_readout_per_particle_add_slices_and_errors():
    """gfn_readout_per_particle_add through raw ctypes: a slice of the [mode][core site][bin] array is ADDED to the
    caller's values; slices beyond the array and handles without GFN_FLAG_PER_PARTICLE are refused."""
    import ctypes
    import newanalysis
    lib = ctypes.CDLL(os.path.join(os.path.dirname(newanalysis.__file__), "libgfn.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    U = ctypes.c_ulonglong
    lib.gfn_create.restype = ctypes.c_int
    lib.gfn_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                               ctypes.c_float, ctypes.c_double, ctypes.c_int, dp, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                               ctypes.c_uint]
    lib.gfn_enqueue_frame.restype = ctypes.c_int
    lib.gfn_enqueue_frame.argtypes = [ctypes.c_void_p, dp, dp, dp, dp]
    lib.gfn_readout_per_particle.restype = ctypes.c_int
    lib.gfn_readout_per_particle.argtypes = [ctypes.c_void_p, dp]
    lib.gfn_readout_per_particle_add.restype = ctypes.c_int
    lib.gfn_readout_per_particle_add.argtypes = [ctypes.c_void_p, dp, U, U]
    lib.gfn_destroy.restype = None
    lib.gfn_destroy.argtypes = [ctypes.c_void_p]
    lib.gfn_last_error.restype = ctypes.c_char_p
    lib.gfn_last_error.argtypes = [ctypes.c_void_p]
    n, hn, L = 64, 100, 20.0
    rng = np.random.default_rng(44)
    c = np.ascontiguousarray(rng.random((n, 3)) * L)
    box = (ctypes.c_double * 6)(L, L, L, L / 2, L / 2, L / 2)
    ptr = lambda a: a.ctypes.data_as(dp)

    def handle(flags):
        h = ctypes.c_void_p()
        rc = lib.gfn_create(ctypes.byref(h), n, n, hn, 0.0, 10.0, 10.0, 1, box, None, 0, flags)
        assert rc == 0, lib.gfn_last_error(None)
        return h

    h = handle(1 | 32)                                            # per-particle, no triangle
    try:
        assert lib.gfn_enqueue_frame(h, ptr(c), ptr(c), None, None) == 0, lib.gfn_last_error(h)
        full = np.zeros(7 * n * hn)
        assert lib.gfn_readout_per_particle(h, ptr(full)) == 0 and full[:n * hn].sum() > 100
        part = np.full(3 * hn, 5.0)
        assert lib.gfn_readout_per_particle_add(h, ptr(part), 7 * hn, 3 * hn) == 0, lib.gfn_last_error(h)
        assert np.array_equal(part, 5.0 + full[7 * hn:10 * hn]) and full[7 * hn:10 * hn].sum() > 0
        assert lib.gfn_readout_per_particle_add(h, ptr(part), 0, 0) == 0
        assert lib.gfn_readout_per_particle_add(h, ptr(part), 7 * n * hn - 2, 3) != 0
        assert b"exceeds" in lib.gfn_last_error(h)
        assert lib.gfn_readout_per_particle_add(h, None, 0, 1) != 0
    finally:
        lib.gfn_destroy(h)
    g = handle(0)
    try:
        assert lib.gfn_readout_per_particle_add(g, ptr(part), 0, 1) != 0
        assert b"PER_PARTICLE" in lib.gfn_last_error(g)
    finally:
        lib.gfn_destroy(g)


@pytest.mark.parametrize("name", sorted(C.TS_CASES))
def test_calc_histo_ts_golden(golden_dir, name):
    """calcHistoTS against what the UNMODIFIED reference function wrote (tests/golden/ts.npz), including the quirk-Q3
    spill of a pair at exactly histo_max into the next core row / the next frame's first row (ts_edge)."""
    from newanalysis.gfunction import calcHistoTS
    want = np.load(os.path.join(golden_dir, "ts.npz"))[name]
    got = C.drive_ts(calcHistoTS, C.TS_CASES[name])
    assert got.shape == want.shape and want.sum() > 50
    assert np.array_equal(got, want)


# ---- N2: shell-resolved g-functions (RDF_voronoi gfunction.pyx:764-1148, calcHistoTSVoro :195-224) --------------
def assert_shell_hist_close(got, ref, rows, pairw):
    """rows = histo_n*nshells bins per mode; same bars as assert_hist_close."""
    scale = np.maximum(np.abs(ref), np.tile(np.abs(pairw), 7))
    err = np.abs(np.asarray(got) - np.asarray(ref))
    bad = err > RTOL * scale
    assert not bad.any(), "max err/scale = %g at %s" % (np.max(err / np.maximum(scale, 1e-300)), np.flatnonzero(bad)[:5])


@pytest.mark.parametrize("flags", ["", "fp32", "fp64", "fp16,global", "per_particle"])
@pytest.mark.parametrize("name", sorted(C.VORO_CASES))
def test_voronoi_golden(golden_dir, name, flags, tmp_path):
    from newanalysis.gfunction import RDF_voronoi
    raw, out, meta = load_golden(golden_dir, name)
    case = C.VORO_CASES[name]
    if flags == "per_particle" and case["n1"] > 1000:
        pytest.skip("per-particle layout only on the small cases")
    rdf = C.drive_voro(C.make_voro(RDF_voronoi, case, gfn_flags=flags), case)
    hn, ns = meta["histo_n"], meta["nshells"]
    rows = hn * ns
    assert int(rdf.histo_n) == hn and int(rdf.nshells) == ns and rdf.ctr == meta["ctr"]
    assert bool(rdf.oneistwo) == meta["oneistwo"] and float(rdf.volume) == meta["volume"]
    got = rdf.raw_histogram()
    assert got.shape == (7 * rows,)
    if meta["mode_sel"] & 1:
        assert np.array_equal(got[:rows], raw[:rows]), "g000 not bit-exact"
    assert_shell_hist_close(got, raw, rows, np.maximum(rdf.pair_weight, raw[:rows]))
    if flags == "per_particle":
        pp = rdf.histogram
        assert pp.shape == (7 * case["n1"] * rows,)
        assert_shell_hist_close(pp.reshape(7, case["n1"], rows).sum(1).reshape(-1), raw, rows, np.maximum(rdf.pair_weight, raw[:rows]))
    rdf.write(str(tmp_path / "g"))
    if meta["mode_sel"] & 1:
        assert np.array_equal(rdf.histogram_out[:rows], out[:rows])
    for key, text in meta["dat"].items():
        sh, mode = key.split("_")
        a = np.loadtxt(str(tmp_path / ("g_shell%s_g%s.dat" % (sh, mode))))
        b = np.array([[float(v) for v in ln.split()] for ln in text.strip().splitlines()])
        assert a.shape == b.shape and np.allclose(a, b, rtol=0, atol=2e-5)
        if mode == "000":
            assert (tmp_path / ("g_shell%s_g000.dat" % sh)).read_text() == text


def test_voronoi_against_oracle_larger():
    """2048 molecules of 2 sites, 3 shells, self: counts exact per shell, weights within 1e-12."""
    from newanalysis.gfunction import RDF_voronoi
    case = dict(n1=4096, n2=4096, nmol1=2048, nmol2=2048, box=(49.67, None, None), modes=["all"], hmin=0.0, hmax=15.0, dr=0.05,
                nshells=3, frames=2, same=True, seed=91, exclude_self=True)
    gpu = C.drive_voro(C.make_voro(RDF_voronoi, case), case)
    got = gpu.raw_histogram()
    rows = 300 * 3
    ref = np.zeros(7 * rows)
    o = C.make_voro(O.OracleRDFVoronoi, dict(case, n1=1, n2=1, nmol1=1, nmol2=1))     # host parameters only
    for f in range(case["frames"]):
        c1, c2, d1, d2, _, dl = C.build_voro_inputs(case, f)
        big = np.zeros(7 * case["n1"] * rows)
        rc, oob = O.calc_histo_shells_c(c1, c2, d1, d2, big, o.box_dim, o.histo_min, o.histo_max, 300, o.histo_inv_dr, 127,
                                        case["nmol1"], case["nmol2"], 3, dl, True)
        assert rc == 0 and oob == 0
        ref += big.reshape(7, case["n1"], rows).sum(1).reshape(-1)
    assert np.array_equal(got[:rows], ref[:rows])
    assert_shell_hist_close(got, ref, rows, np.maximum(gpu.pair_weight, ref[:rows]))


def test_voronoi_argument_errors():
    from newanalysis.gfunction import RDF_voronoi
    n = 64
    c = np.random.default_rng(1).random((n, 3)) * 20.0
    d = np.ones((n, 3))
    r = RDF_voronoi(["all"], 0.0, 8.0, 0.1, n, n, 16, 16, 20.0, 2)
    with pytest.raises(TypeError, match="first selection"):
        r.calcFrame(c, c, np.ones((16, n), dtype=np.int32))
    with pytest.raises(ValueError):                       # matrix smaller than the reference's indexing needs
        r.calcFrame(c, c, np.ones((16, 16), dtype=np.int32), d, d)
    r2 = RDF_voronoi(["000"], 0.0, 8.0, 0.1, n, n, 16, 16, 20.0, 2)
    r2.calcFrame(c, c, np.ones((16, n), dtype=np.int32))
    with pytest.raises(ValueError, match="changed shape"):
        r2.calcFrame(c, c, np.ones((17, n), dtype=np.int32))
    assert r2.raw_histogram().reshape(7, 80, 2)[0, :, 0].sum() > 0
    with pytest.raises(ValueError):
        RDF_voronoi(["000"], 0.0, 8.0, 0.1, n, n, 128, 16, 20.0, 2).calcFrame(c, c, np.ones((128, n), dtype=np.int32))


def test_voronoi_before_first_frame_and_batch_call():
    """RDF_voronoi creates its device handle with the first frame: the inherited helpers must not trip over that, and
    calcFrames (many frames per call) equals frame-by-frame calls."""
    from newanalysis.gfunction import RDF_voronoi
    case = C.VORO_CASES["voro_rect"]
    r = C.make_voro(RDF_voronoi, case)
    r.sync(); r.flush(); r.scale(2.0); r.resetArrays()
    assert not r.raw_histogram().any()
    with pytest.raises(RuntimeError, match="first calcFrame"):
        r.stats()
    with pytest.raises(TypeError):
        r.attach_topology(np.ones(3), [3], [0])
    a = C.drive_voro(C.make_voro(RDF_voronoi, case), case)
    fr = [C.build_voro_inputs(case, f) for f in range(case["frames"])]
    r.calcFrames(np.stack([x[0] for x in fr]), np.stack([x[1] for x in fr]), np.stack([x[5] for x in fr]),
                 np.stack([x[2] for x in fr]), np.stack([x[3] for x in fr]), boxes=np.array([x[4] for x in fr]))
    got, ref = r.raw_histogram(), a.raw_histogram()
    rows = int(r.histo_n) * int(r.nshells)
    assert np.array_equal(got[:rows], ref[:rows]) and r.ctr == a.ctr       # counts exact; weighted rows up to batch order
    assert_shell_hist_close(got, ref, rows, np.maximum(r.pair_weight, ref[:rows]))
    assert r.stats()["frames"] == case["frames"]


@pytest.mark.parametrize("name", sorted(C.TSVORO_CASES))
def test_calc_histo_ts_voro_golden(golden_dir, name):
    from newanalysis.gfunction import calcHistoTSVoro
    want = np.load(os.path.join(golden_dir, "tsvoro.npz"))[name]
    got = C.drive_tsvoro(calcHistoTSVoro, C.TSVORO_CASES[name])
    assert np.array_equal(got, want)


# ---- the C ABI itself, without the Cython shim ------------------------------------------------------------------
def test_c_abi_through_ctypes_matches_golden(golden_dir):
    """gfn_create -> gfn_enqueue_frame -> gfn_readout driven through raw ctypes (what a cgo / JNI / ctypes binding of
    include/gfn.h would do) on the golden case rect_all: g000 bit-exact, weighted rows 1e-12."""
    import ctypes
    import newanalysis
    lib = ctypes.CDLL(os.path.join(os.path.dirname(newanalysis.__file__), "libgfn.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    lib.gfn_create.restype = ctypes.c_int
    lib.gfn_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                               ctypes.c_float, ctypes.c_double, ctypes.c_int, dp, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                               ctypes.c_uint]
    lib.gfn_enqueue_frame.restype = ctypes.c_int
    lib.gfn_enqueue_frame.argtypes = [ctypes.c_void_p, dp, dp, dp, dp]
    lib.gfn_readout.restype = ctypes.c_int
    lib.gfn_readout.argtypes = [ctypes.c_void_p, dp, dp, ctypes.POINTER(ctypes.c_ulonglong)]
    lib.gfn_destroy.restype = None
    lib.gfn_destroy.argtypes = [ctypes.c_void_p]
    lib.gfn_last_error.restype = ctypes.c_char_p
    lib.gfn_last_error.argtypes = [ctypes.c_void_p]

    raw, _out, meta = load_golden(golden_dir, "rect_all")
    case = C.CASES["rect_all"]
    hn = meta["histo_n"]
    L = case["box"][0]
    box = (ctypes.c_double * 6)(L, L, L, L / 2, L / 2, L / 2)
    h = ctypes.c_void_p()
    # histo_min/max as C floats, inv_dr = 1/round(dr,5): what RDF.__init__ computes (gfunction.pyx:291-295)
    rc = lib.gfn_create(ctypes.byref(h), case["n1"], case["n2"], hn, ctypes.c_float(case["hmin"]), ctypes.c_float(case["hmax"]),
                        1.0 / np.round(case["dr"], 5), 127, box, None, 0, 0)
    assert rc == 0, lib.gfn_last_error(None)
    try:
        for f in range(case["frames"]):
            c1, c2, d1, d2, _ = C.build_inputs(case, f)
            p = [a.ctypes.data_as(dp) for a in (c1, c2, d1, d2)]
            assert lib.gfn_enqueue_frame(h, *p) == 0, lib.gfn_last_error(h)
        hist = np.zeros(7 * hn); pw = np.zeros(hn); ovf = ctypes.c_ulonglong(0)
        assert lib.gfn_readout(h, hist.ctypes.data_as(dp), pw.ctypes.data_as(dp), ctypes.byref(ovf)) == 0, lib.gfn_last_error(h)
    finally:
        lib.gfn_destroy(h)
    assert np.array_equal(hist[:hn], raw[:hn]) and raw[:hn].sum() > 0
    assert_hist_close(hist, raw, hn, pairw=np.maximum(pw, raw[:hn]))

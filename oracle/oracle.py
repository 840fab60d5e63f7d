"""oracle/oracle.py - Python side of the CPU checker for the g-function pair path.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product (libgfn.so + the newanalysis.gfunction
shim) never does, and has no CPU fallback.

Parity status: PINNED against the reference's own compiled module (oracle/_ref, see build_ref.py)
and the fixtures under tests/golden/ generated from it (tests/golden/make_golden.py).

Contents
  load_c()            ctypes handle on oracle/libgfn_oracle.so (gfn_oracle.c, built by oracle/Makefile)
  load_ref()          the unmodified reference module (oracle/_ref/gfunction*.so) or None
  calc_histo_numpy()  30-line numpy restatement of calcHisto (gfunction.pyx:110-165), small cases
  OracleRDF           restatement of the reference RDF class host logic (gfunction.pyx:227-566)
                      on top of the C restatement: __init__ (A1), calcFrame (A2), _addHisto (A4),
                      _normHisto (A5), write (A6), update_box/scale/resetArrays (A7)
  exact_thresholds()  the d^2 acceptance interval equivalent to gfunction.pyx:149 (used to test the
                      library's host-side threshold computation)
"""
import ctypes
import glob
import importlib.util
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MODE_BITS = {"000": 1, "110": 2, "101": 4, "011": 8, "220": 16, "202": 32, "022": 64}
MODE_ORDER = ["000", "110", "101", "011", "220", "202", "022"]   # histogram row k = index here

_c = None
_ref = False


def build_c(force=False):
    so = os.path.join(HERE, "libgfn_oracle.so")
    src = os.path.join(HERE, "gfn_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "CC=/usr/bin/gcc", "libgfn_oracle.so"])
    return so


def load_c():
    global _c
    if _c is not None:
        return _c
    lib = ctypes.CDLL(build_c())
    dp = ctypes.POINTER(ctypes.c_double)
    llp = ctypes.POINTER(ctypes.c_longlong)
    lib.gfn_oracle_calc_histo.restype = ctypes.c_int
    lib.gfn_oracle_calc_histo.argtypes = [dp, dp, dp, dp, dp, dp, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                          ctypes.c_double, ctypes.c_int, llp]
    lib.gfn_oracle_add_histo.restype = None
    lib.gfn_oracle_add_histo.argtypes = [dp, dp, ctypes.c_int, ctypes.c_int]
    lib.gfn_oracle_calc_histo_ts.restype = ctypes.c_longlong
    lib.gfn_oracle_calc_histo_ts.argtypes = [dp, dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_double, ctypes.c_float, ctypes.c_float,
                                             ctypes.c_double]
    lib.gfn_oracle_calc_histo_rows.restype = ctypes.c_int
    lib.gfn_oracle_calc_histo_rows.argtypes = [dp, dp, dp, dp, dp, dp, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                               ctypes.c_float, ctypes.c_int, ctypes.c_double,
                                               ctypes.c_int, llp]
    ip = ctypes.POINTER(ctypes.c_int)
    lib.gfn_oracle_calc_histo_shells.restype = ctypes.c_int
    lib.gfn_oracle_calc_histo_shells.argtypes = [dp, dp, dp, dp, dp, dp, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                                 ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int, ip, ctypes.c_longlong, ctypes.c_int, llp]
    lib.gfn_oracle_add_histo_shells.restype = None
    lib.gfn_oracle_add_histo_shells.argtypes = [dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.gfn_oracle_calc_histo_ts_voro.restype = ctypes.c_longlong
    lib.gfn_oracle_calc_histo_ts_voro.argtypes = [dp, dp, dp, ctypes.POINTER(ctypes.c_byte), ctypes.c_int,
                                                  ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                                  ctypes.c_float, ctypes.c_float, ctypes.c_double]
    _c = lib
    return lib


def load_ref():
    """Import the unmodified reference gfunction module from oracle/_ref (None if not built)."""
    global _ref
    if _ref is not False:
        return _ref
    hits = sorted(glob.glob(os.path.join(HERE, "_ref", "gfunction*.so")))
    if not hits:
        _ref = None
        return None
    spec = importlib.util.spec_from_file_location("gfunction", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _ref = mod
    return mod


def load_ref_helpers():
    """The unmodified reference helpers module (oracle/_ref/helpers*.so, newanalysis/helpers/helpers.pyx) or None.
    Its import line `from scipy.special import sph_harm` fails on scipy >= 1.17 (the name was removed); the
    loader provides a placeholder for that one unrelated name instead of touching the reference source."""
    hits = sorted(glob.glob(os.path.join(HERE, "_ref", "helpers*.so")))
    if not hits:
        return None
    import scipy.special
    if not hasattr(scipy.special, "sph_harm"):
        scipy.special.sph_harm = None
    spec = importlib.util.spec_from_file_location("helpers", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def com_by_residue_c(coor, masses, apr, rfa):
    """C restatement of helpers.comByResidue (helpers.pyx:71-98)."""
    lib = load_c()
    coor = _c64(coor); masses = _c64(masses)
    apr = np.ascontiguousarray(apr, dtype=np.int32); rfa = np.ascontiguousarray(rfa, dtype=np.int32)
    com = np.zeros((apr.shape[0], 3))
    ip = ctypes.POINTER(ctypes.c_int)
    lib.gfn_oracle_com_by_residue.restype = None
    lib.gfn_oracle_com_by_residue(_dptr(coor), _dptr(masses), ctypes.c_int(apr.shape[0]), apr.ctypes.data_as(ip),
                                  rfa.ctypes.data_as(ip), _dptr(com))
    return com


def dip_by_residue_c(coor, charges, apr, rfa, com):
    """C restatement of helpers.dipByResidue (helpers.pyx:101-123)."""
    lib = load_c()
    coor = _c64(coor); charges = _c64(charges); com = _c64(com)
    apr = np.ascontiguousarray(apr, dtype=np.int32); rfa = np.ascontiguousarray(rfa, dtype=np.int32)
    dip = np.zeros((apr.shape[0], 3))
    ip = ctypes.POINTER(ctypes.c_int)
    lib.gfn_oracle_dip_by_residue.restype = None
    lib.gfn_oracle_dip_by_residue(_dptr(coor), _dptr(charges), ctypes.c_int(apr.shape[0]), apr.ctypes.data_as(ip),
                                  rfa.ctypes.data_as(ip), _dptr(com), _dptr(dip))
    return dip


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if a is not None else None


def _c64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def calc_histo_c(c1, c2, d1, d2, dest, box_dim, hmin, hmax, histo_n, invdx, mode_sel):
    """C restatement of calcHisto; dest is the flat [7][n1][histo_n] float64 array (updated in place).
    Returns (zerodiv_flag, lost_writes)."""
    lib = load_c()
    c1 = _c64(c1); c2 = _c64(c2); d1 = _c64(d1); d2 = _c64(d2)
    box = np.ascontiguousarray(box_dim, dtype=np.float64)
    oob = ctypes.c_longlong(0)
    rc = lib.gfn_oracle_calc_histo(_dptr(c1), _dptr(c2), _dptr(d1), _dptr(d2), _dptr(dest), _dptr(box),
                                   c1.shape[0], c2.shape[0], np.float32(hmin), np.float32(hmax),
                                   int(histo_n), float(invdx), int(mode_sel), ctypes.byref(oob))
    return rc, oob.value


def calc_histo_rows_c(c1, c2, d1, d2, out7, box_dim, hmin, hmax, histo_n, invdx, mode_sel,
                      row_begin=0, row_end=None):
    """Reduced-memory C restatement: accumulates rows [row_begin,row_end) into out7 ([7][histo_n])."""
    lib = load_c()
    c1 = _c64(c1); c2 = _c64(c2); d1 = _c64(d1); d2 = _c64(d2)
    box = np.ascontiguousarray(box_dim, dtype=np.float64)
    if row_end is None:
        row_end = c1.shape[0]
    oob = ctypes.c_longlong(0)
    rc = lib.gfn_oracle_calc_histo_rows(_dptr(c1), _dptr(c2), _dptr(d1), _dptr(d2), _dptr(out7),
                                        _dptr(box), c1.shape[0], c2.shape[0], int(row_begin),
                                        int(row_end), np.float32(hmin), np.float32(hmax), int(histo_n),
                                        float(invdx), int(mode_sel), ctypes.byref(oob))
    return rc, oob.value


def calc_histo_ts_c(core, surr, result_t, boxl, hmin, hmax, invdx):
    """calcHistoTS on the slab result_t = result[t] ([n1][nbins]); a Q3 pair of the last core row is dropped and counted."""
    lib = load_c()
    core = _c64(core); surr = _c64(surr)
    return lib.gfn_oracle_calc_histo_ts(_dptr(core), _dptr(surr), _dptr(result_t), core.shape[0],
                                        surr.shape[0], result_t.shape[1], float(boxl),
                                        np.float32(hmin), np.float32(hmax), float(invdx))


def calc_histo_ts(core, surr, result, boxl, t, hmin, hmax, invdx):
    """Same signature as the reference's calcHistoTS (gfunction.pyx:168-192) on the whole C-contiguous result array:
    a Q3 pair (pos >= nbins) lands where the reference's unchecked flat index points, also across the frame boundary."""
    lib = load_c()
    core = _c64(core); surr = _c64(surr)
    assert result.flags.c_contiguous and result.dtype == np.float64
    slab = result[t]
    room = (result.shape[0] - t) * result.shape[1] * result.shape[2]
    lib.gfn_oracle_calc_histo_ts_room.restype = ctypes.c_longlong
    return lib.gfn_oracle_calc_histo_ts_room(_dptr(core), _dptr(surr), _dptr(slab), ctypes.c_int(core.shape[0]),
                                             ctypes.c_int(surr.shape[0]), ctypes.c_int(result.shape[2]), ctypes.c_double(boxl),
                                             ctypes.c_float(np.float32(hmin)), ctypes.c_float(np.float32(hmax)),
                                             ctypes.c_double(invdx), ctypes.c_longlong(room))


def calc_histo_shells_c(c1, c2, d1, d2, dest, box_dim, hmin, hmax, histo_n, invdx, mode_sel,
                        nm_core, nm_surround, nshells, delaunay, exclude_self=False):
    """calcHistoVoronoi / calcHistoVoronoiNonSelf (gfunction.pyx:570-761) on the C restatement.
    Returns (status, out_of_bounds)."""
    lib = load_c()
    c1 = _c64(c1); c2 = _c64(c2)
    d1 = _c64(d1) if d1 is not None else None
    d2 = _c64(d2) if d2 is not None else None
    dl = np.ascontiguousarray(delaunay, dtype=np.int32)
    oob = ctypes.c_longlong(0)
    rc = lib.gfn_oracle_calc_histo_shells(
        _dptr(c1), _dptr(c2), _dptr(d1) if d1 is not None else None, _dptr(d2) if d2 is not None else None,
        _dptr(dest), _dptr(box_dim), c1.shape[0], c2.shape[0], float(hmin), float(hmax), int(histo_n),
        float(invdx), int(mode_sel), int(nm_core), int(nm_surround), int(nshells),
        dl.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), dl.size, 1 if exclude_self else 0, ctypes.byref(oob))
    return rc, oob.value


def calc_histo_ts_voro_c(core, surr, result_t, ds_t, maxshell, boxl, hmin, hmax, invdx):
    """calcHistoTSVoro (gfunction.pyx:195-224) for one frame: result_t is result[t] ([ncore][maxshell][nbins]),
    ds_t is ds[t] ([ncore][nsurr] int8)."""
    core = _c64(core); surr = _c64(surr)
    ds_t = np.ascontiguousarray(ds_t, dtype=np.int8)
    assert result_t.flags.c_contiguous and result_t.dtype == np.float64
    return load_c().gfn_oracle_calc_histo_ts_voro(
        _dptr(core), _dptr(surr), _dptr(result_t), ds_t.ctypes.data_as(ctypes.POINTER(ctypes.c_byte)), int(maxshell),
        core.shape[0], surr.shape[0], result_t.shape[-1], float(boxl), float(hmin), float(hmax), float(invdx))


def calc_histo_numpy(c1, c2, d1, d2, dest, box_dim, hmin, hmax, histo_n, invdx, mode_sel):
    """numpy restatement of gfunction.pyx:110-165, one core row at a time (small cases only).

    Every arithmetic expression keeps the reference's association so per-pair values are bit-identical;
    np.add.at applies the adds of a row in ascending j, i.e. the reference order."""
    n1, n2 = c1.shape[0], c2.shape[0]
    hmin = np.float64(np.float32(hmin)); hmax = np.float64(np.float32(hmax))
    L, half = np.asarray(box_dim[:3], float), np.asarray(box_dim[3:6], float)
    ix1 = n1 * histo_n
    need_c = mode_sel & (2 | 4 | 16 | 32)
    need_s = mode_sel & (2 | 8 | 16 | 64)
    if need_s:
        slen_all = np.sqrt(d2[:, 0] * d2[:, 0] + d2[:, 1] * d2[:, 1] + d2[:, 2] * d2[:, 2])
    for i in range(n1):
        jb = i if n1 == n2 else 0
        d = c2[jb:] - c1[i]
        sgn = (0.0 < d).astype(float) - (d < 0.0).astype(float)
        d = np.where(np.abs(d) > half, d - sgn * L, d)
        dist = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
        ok = ~((dist < hmin) | (dist > hmax) | (np.floor(dist * invdx) == 0))
        pos = np.floor((dist - hmin) * invdx).astype(np.int64)
        w = np.ones(d.shape[0])
        if n1 == n2:
            w[:] = 2.0
            w[0] = 1.0          # j == i
        pos, w_ok, dd, dist_ok = pos[ok], w[ok], d[ok], dist[ok]
        base = i * histo_n + pos

        def put(k, v):
            ix = k * ix1 + base
            keep = ix < dest.shape[0]
            np.add.at(dest, ix[keep], v[keep])

        if mode_sel & 1:
            put(0, 1.0 * w_ok)
        if need_c:
            cp = d1[i]
            clen = math.sqrt(cp[0] * cp[0] + cp[1] * cp[1] + cp[2] * cp[2])
        if need_s:
            sp = d2[jb:][ok]
            slen = slen_all[jb:][ok]
        if mode_sel & (2 | 16):
            c = (cp[0] * sp[:, 0] + cp[1] * sp[:, 1] + cp[2] * sp[:, 2]) / (clen * slen)
            if mode_sel & 2: put(1, c * w_ok)
            if mode_sel & 16: put(4, (1.5 * c * c - 0.5) * w_ok)
        if mode_sel & (4 | 32):
            c = (cp[0] * dd[:, 0] + cp[1] * dd[:, 1] + cp[2] * dd[:, 2]) / (clen * dist_ok)
            if mode_sel & 4: put(2, c * w_ok)
            if mode_sel & 32: put(5, (1.5 * c * c - 0.5) * w_ok)
        if mode_sel & (8 | 64):
            c = (sp[:, 0] * dd[:, 0] + sp[:, 1] * dd[:, 1] + sp[:, 2] * dd[:, 2]) / (slen * dist_ok)
            if mode_sel & 8: put(3, c * w_ok)
            if mode_sel & 64: put(6, (1.5 * c * c - 0.5) * w_ok)


def modes_to_mask(modes):
    """gfunction.pyx:274-288 - unknown strings are silently ignored."""
    m = 0
    for mode in modes:
        if mode == "all":
            m |= 127
        m |= MODE_BITS.get(mode, 0)
    return m


class OracleRDF(object):
    """Host-logic restatement of the reference RDF class (gfunction.pyx:227-566), CPU path only."""

    def __init__(self, modes, histo_min, histo_max, histo_dr, sel1_n, sel2_n, sel1_nmol, sel2_nmol,
                 box_x, box_y=None, box_z=None, use_cuda=False, norm_volume=None):
        self.cuda = False                                                   # :260-266 (no pycuda)
        self.dtype = np.float64                                             # :268-271
        self.mode_sel = np.int32(modes_to_mask(modes))                      # :274-288
        self.histo_min = np.float32(histo_min)                              # :291
        self.histo_max = np.float32(histo_max)                              # :292
        self.histo_dr = np.round(self.dtype(histo_dr), 5)                   # :293
        self.histo_inv_dr = self.dtype(1.0 / self.histo_dr)                 # :294
        self.histo_n = np.int32((histo_max - histo_min) / histo_dr + 0.5)   # :295
        self.box_dim = np.zeros(6, dtype=self.dtype)                        # :297-311
        self.box_dim[0] = box_x
        self.box_dim[3] = box_x / 2.0
        self.box_dim[1], self.box_dim[4] = (box_y, box_y / 2.0) if box_y else (self.box_dim[0], self.box_dim[3])
        self.box_dim[2], self.box_dim[5] = (box_z, box_z / 2.0) if box_z else (self.box_dim[0], self.box_dim[3])
        self.volume = self.box_dim[0] * self.box_dim[1] * self.box_dim[2]   # :313
        self.volume_list = []
        self.norm_volume = norm_volume
        self.n1 = np.int32(sel1_n)
        self.n2 = np.int32(sel2_n)
        self.nmol1 = sel1_nmol
        self.nmol2 = sel2_nmol
        self.oneistwo = None
        self.ctr = 0
        self.histogram = np.zeros(int(self.n1) * int(self.histo_n) * 7, dtype=self.dtype)   # :332
        self.histogram_out = np.zeros(int(self.histo_n) * 7, dtype=np.float64)              # :333

    def update_box(self, box_x, box_y, box_z):                              # :367-383
        self.box_dim[0] = box_x; self.box_dim[3] = box_x / 2.0
        self.box_dim[1] = box_y; self.box_dim[4] = box_y / 2.0
        self.box_dim[2] = box_z; self.box_dim[5] = box_z / 2.0
        self.volume = self.box_dim[0] * self.box_dim[1] * self.box_dim[2]

    def calcFrame(self, coor_sel1, coor_sel2, dip_sel1=None, dip_sel2=None):    # :385-464
        m = int(self.mode_sel)
        if dip_sel1 is None and m & (2 | 4 | 16 | 32):
            raise TypeError("No dipole moment array for the fist selection was passed, although the requested gfunctions need it!")
        if dip_sel2 is None and m & (2 | 8 | 16 | 64):
            raise TypeError("No dipole moment array for the second selection was passed, although the requested gfunctions need it!")
        if self.oneistwo is None:                                           # :417-421
            self.oneistwo = bool((coor_sel1[0] == coor_sel2[0]).all() and self.nmol1 == self.nmol2)
        self.ctr += 1
        self.volume_list.append(self.volume)
        rc, _ = calc_histo_c(coor_sel1, coor_sel2, dip_sel1, dip_sel2, self.histogram, self.box_dim,
                             self.histo_min, self.histo_max, self.histo_n, self.histo_inv_dr, m)
        if rc:
            raise ZeroDivisionError("float division")

    def _addHisto(self, histo_out, histo):                                  # :466-480
        load_c().gfn_oracle_add_histo(_dptr(histo_out), _dptr(histo), int(self.n1), int(self.histo_n))

    def _normHisto(self):                                                   # :483-511
        if self.norm_volume:
            volume = self.norm_volume
        else:
            volume = 0.0
            for v in self.volume_list:
                volume += v
            volume /= len(self.volume_list)
        PI = 3.14159265358979323846
        if self.oneistwo:
            rho = float(self.nmol1 * (self.nmol1 - 1)) / volume
        else:
            rho = float(self.nmol1 * self.nmol2) / volume
        for i in range(self.histo_n):
            r = self.histo_min + float(i) * self.histo_dr
            r_out = r + self.histo_dr
            slice_vol = 4.0 / 3.0 * PI * (r_out ** 3 - r ** 3)
            norm = rho * slice_vol * float(self.ctr)
            for j in range(7):
                self.histogram_out[j * self.histo_n + i] /= norm

    def scale(self, value):                                                 # :513-515
        self.histogram_out[:] *= value
        self.histogram[:] *= value

    def write(self, filename="rdf"):                                        # :517-550
        self._addHisto(self.histogram_out, self.histogram.astype(np.float64))
        self._normHisto()
        for k, name in enumerate(MODE_ORDER):
            if int(self.mode_sel) & MODE_BITS[name]:
                with open(filename + "_g" + name + ".dat", "w") as fh:
                    for i in range(self.histo_n):
                        r = self.histo_min + i * self.histo_dr + self.histo_dr * 0.5
                        fh.write("%5.5f\t%5.5f\n" % (r, self.histogram_out[k * self.histo_n + i]))

    def resetArrays(self):                                                  # :552-557
        self.histogram_out[:] = 0.0
        self.histogram[:] = 0.0

    def raw_sum(self):
        """sum_i histogram[k][i][b] in the reference order (what _addHisto adds) - the quantity the
        GPU path is compared against, before normalisation."""
        out = np.zeros(7 * int(self.histo_n))
        self._addHisto(out, self.histogram)
        return out


class OracleRDFVoronoi(OracleRDF):
    """Host-logic restatement of the reference RDF_voronoi class (gfunction.pyx:764-1148), CPU path."""

    def __init__(self, modes, histo_min, histo_max, histo_dr, sel1_n, sel2_n, sel1_nmol, sel2_nmol,
                 box_x, nshells, box_y=None, box_z=None, use_cuda=False, norm_volume=None, exclude_self_mol=False):
        OracleRDF.__init__(self, modes, histo_min, histo_max, histo_dr, sel1_n, sel2_n, sel1_nmol, sel2_nmol,
                           box_x, box_y, box_z, use_cuda, norm_volume)
        self.nshells = np.int32(nshells)                                    # :842
        self.exclude_self_mol = exclude_self_mol                            # :843
        self.histogram = np.zeros(int(self.nshells) * int(self.n1) * int(self.histo_n) * 7, dtype=self.dtype)   # :881
        self.histogram_out = np.zeros(int(self.nshells) * int(self.histo_n) * 7, dtype=np.float64)              # :882

    def update_box(self, box_x, box_y, box_z):                              # :916-935: self.volume is NOT refreshed
        self.box_dim[0] = box_x; self.box_dim[3] = box_x / 2.0              #   (the assignment sits inside a string literal)
        self.box_dim[1] = box_y; self.box_dim[4] = box_y / 2.0
        self.box_dim[2] = box_z; self.box_dim[5] = box_z / 2.0

    def calcFrame(self, coor_sel1, coor_sel2, delaunay_matrix, dip_sel1=None, dip_sel2=None):   # :937-1021
        m = int(self.mode_sel)
        if dip_sel1 is None and m & (2 | 4 | 16 | 32):
            raise TypeError("No dipole moment array for the first selection was passed, although the requested gfunctions need it!")
        if dip_sel2 is None and m & (2 | 8 | 16 | 64):
            raise TypeError("No dipole moment array for the second selection was passed, although the requested gfunctions need it!")
        if self.oneistwo is None:
            self.oneistwo = bool((coor_sel1[0] == coor_sel2[0]).all() and self.nmol1 == self.nmol2)
        self.ctr += 1
        self.volume_list.append(self.volume)
        rc, _ = calc_histo_shells_c(coor_sel1, coor_sel2, dip_sel1, dip_sel2, self.histogram, self.box_dim,
                                    self.histo_min, self.histo_max, self.histo_n, self.histo_inv_dr, m,
                                    self.nmol1, self.nmol2, self.nshells, delaunay_matrix, self.exclude_self_mol)
        if rc == 1:
            raise ZeroDivisionError("float division")
        if rc:
            raise ValueError("delaunay matrix too small for the reference's indexing")

    def _addHisto(self, histo_out, histo):                                  # :1023-1047
        load_c().gfn_oracle_add_histo_shells(_dptr(histo_out), _dptr(histo), int(self.n1), int(self.histo_n), int(self.nshells))

    def _normHisto(self):                                                   # :1049-1089
        if self.norm_volume:
            volume = self.norm_volume
        else:
            volume = 0.0
            for v in self.volume_list:
                volume += v
            volume /= len(self.volume_list)
        PI = 3.14159265358979323846
        if self.oneistwo:
            rho = float(self.nmol1 * (self.nmol1 - 1)) / volume
        else:
            rho = float(self.nmol1 * self.nmol2) / volume
        ns = int(self.nshells)
        for shell in range(ns):
            for i in range(self.histo_n):
                r = self.histo_min + float(i) * self.histo_dr
                r_out = r + self.histo_dr
                slice_vol = 4.0 / 3.0 * PI * (r_out ** 3 - r ** 3)
                norm = rho * slice_vol * float(self.ctr)
                for j in range(7):
                    self.histogram_out[j * self.histo_n * ns + i * ns + shell] /= norm

    def write(self, filename="rdf"):                                        # :1095-1133
        self._addHisto(self.histogram_out, self.histogram.astype(np.float64))
        self._normHisto()
        ns = int(self.nshells)
        for shell in range(ns):
            for k, name in enumerate(MODE_ORDER):
                if int(self.mode_sel) & MODE_BITS[name]:
                    with open(filename + "_shell" + str(shell + 1) + "_g" + name + ".dat", "w") as fh:
                        for i in range(self.histo_n):
                            r = self.histo_min + i * self.histo_dr + self.histo_dr * 0.5
                            fh.write("%5.5f\t%5.5f\n" % (r, self.histogram_out[k * self.histo_n * ns + i * ns + shell]))

    def raw_sum(self):
        out = np.zeros(7 * int(self.histo_n) * int(self.nshells))
        self._addHisto(out, self.histogram)
        return out


def exact_thresholds(hmin_f32, hmax_f32, invdx):
    """(T_lo, T_hi): a pair with d2 = (dx*dx+dy*dy)+dz*dz passes gfunction.pyx:149 iff T_lo < d2 <= T_hi.

    All three rejection tests are monotone in d2 because sqrt, *invdx and floor are monotone:
      accept  <=>  dist >= hmin  and  dist <= hmax  and  fl(dist*invdx) >= 1.0
    T_hi = max{t : sqrt(t) <= hmax};  T_lo = max{t : sqrt(t) < d_acc} with d_acc the smallest accepted dist."""
    hmin = float(np.float32(hmin_f32)); hmax = float(np.float32(hmax_f32)); invdx = float(invdx)
    # smallest d with fl(d*invdx) >= 1.0
    d = 1.0 / invdx
    while d * invdx >= 1.0:
        d = math.nextafter(d, -math.inf)
    while d * invdx < 1.0:
        d = math.nextafter(d, math.inf)
    d_acc = max(d, hmin)

    def max_t_with_sqrt_le(x):          # max{t >= 0 : sqrt(t) <= x}
        t = x * x
        while math.sqrt(t) > x:
            t = math.nextafter(t, -math.inf)
        while math.sqrt(math.nextafter(t, math.inf)) <= x:
            t = math.nextafter(t, math.inf)
        return t

    t_hi = max_t_with_sqrt_le(hmax)
    t_lo = max_t_with_sqrt_le(math.nextafter(d_acc, -math.inf)) if d_acc > 0 else -1.0
    return t_lo, t_hi

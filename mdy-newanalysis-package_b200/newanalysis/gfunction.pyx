# cython: language_level=3, boundscheck=False, wraparound=False
"""newanalysis.gfunction - B200 backend behind the reference's RDF class.

Drop-in for the `RDF` class of /root/reference/newanalysis/gfunction/gfunction.pyx (:227-566): same
constructor signature, mode strings, methods (calcFrame, update_box, scale, write, resetArrays,
printInfo, _addHisto, _normHisto) and attribute names, so analysis scripts keep
`from newanalysis.gfunction import RDF` unchanged.  What changed is below the class: instead of the
OpenMP `calcHisto` loop (:110-165) or the PyCUDA kernel (:13-107), every frame goes through the C ABI
of libgfn.so (include/gfn.h), i.e. hand-written sm_100a kernels fed by a pinned upload ring.

There is NO CPU fallback: without a B200 the constructor raises RuntimeError (the reference printed a
message and continued on the CPU, :262-264).

Differences a user can observe (all documented in INTEGRATION.md):
  * `calcFrame` is asynchronous; a zero-norm dipole (ZeroDivisionError, reference :155-163) is raised
    at the next readout (`write`, `histogram`, `sync`) instead of inside `calcFrame`.
  * `histogram` is a read-only property.  By default it is the [mode][bin] sum over core particles
    (7*histo_n values, what `_addHisto` would produce); with NEWANALYSIS_GFN_FLAGS=per_particle (or
    gfn_flags="per_particle") it is the reference's 7*n1*histo_n per-particle array (:332).
  * extra, optional: `calcFrames` (many frames per call), `sync`, `stats`, `to_device`.
"""
import os

import numpy as np
cimport numpy as np

np.import_array()

cdef extern from "gfn.h":
    ctypedef struct gfn_t:
        pass
    ctypedef struct gfn_stats_t:
        double pair_evals
        double survivors
        double in_range
        double kernel_ms
        double kernel_launches
        double pair_launches
        double frames
        double h2d_bytes
        double overflow
        int filter_fp64
        int hist_mode
        int warps_per_block
        int blocks_per_sm
        int grid
        int smem_bytes
        int batch_frames
        int ndev
        int hist_replicas
        int dense
        double exact_evals
        double selfcheck_bad
        double tile_visits
        double tile_visits_dense
        int cells
        double frames_max_device
        double frames_min_device
        double kernel_ms_max_device
    int GFN_OK, GFN_EINVAL, GFN_ECUDA, GFN_ENCCL, GFN_EZERODIV, GFN_ENOMEM, GFN_ENODEV
    unsigned GFN_FLAG_PER_PARTICLE, GFN_FLAG_FILTER_FP64, GFN_FLAG_DISCARD_OVERFLOW, GFN_FLAG_HIST_GLOBAL
    int gfn_device_count() nogil
    int gfn_create(gfn_t **out, int n1, int n2, int histo_n, float histo_min, float histo_max,
                   double inv_dr, int mode_sel, const double *box_dim, const int *devices, int ndev,
                   unsigned flags) nogil
    int gfn_update_box(gfn_t *h, const double *box_dim) nogil
    int gfn_enqueue_frame(gfn_t *h, const double *c1, const double *c2, const double *d1, const double *d2) nogil
    int gfn_enqueue_frames_pinned(gfn_t *h, int nframes, const double *c1, long long s1, const double *c2, long long s2,
                                  const double *d1, long long s3, const double *d2, long long s4, const double *boxes) nogil
    int gfn_host_alloc(void **out, unsigned long long nbytes) nogil
    int gfn_host_free(void *p) nogil
    int gfn_enqueue_frames(gfn_t *h, int nframes, const double *c1, long long s1, const double *c2, long long s2,
                           const double *d1, long long s3, const double *d2, long long s4, const double *boxes) nogil
    int gfn_enqueue_device_frames(gfn_t *h, int nframes, const double *c1, long long s1, const double *c2, long long s2,
                                  const double *d1, long long s3, const double *d2, long long s4, const double *boxes) nogil
    ctypedef struct gfn_shell_cfg:
        int nshells
        int layout
        int apr_core
        int apr_surround
        int map_ld
        long long map_elems
        int exclude_self
    int gfn_create_shells(gfn_t **out, int n1, int n2, int histo_n, float histo_min, float histo_max,
                          double inv_dr, int mode_sel, const double *box_dim, const int *devices, int ndev,
                          unsigned flags, const gfn_shell_cfg *cfg) nogil
    int gfn_enqueue_frames_shells(gfn_t *h, int nframes, const double *c1, long long s1, const double *c2, long long s2,
                                  const double *d1, long long s3, const double *d2, long long s4, const double *boxes,
                                  const int *shell_map, long long s_map) nogil
    int gfn_flush(gfn_t *h) nogil
    int gfn_sync(gfn_t *h) nogil
    int gfn_readout(gfn_t *h, double *hist, double *pairw, unsigned long long *overflow) nogil
    int gfn_readout_per_particle(gfn_t *h, double *hist) nogil
    int gfn_readout_per_particle_add(gfn_t *h, double *dst, unsigned long long first, unsigned long long count) nogil
    int gfn_scale(gfn_t *h, double value) nogil
    int gfn_reset(gfn_t *h) nogil
    void gfn_destroy(gfn_t *h) nogil
    const char *gfn_last_error(const gfn_t *h) nogil
    int gfn_stats(gfn_t *h, gfn_stats_t *out) nogil
    int gfn_comm_unique_id(char *out) nogil
    int gfn_comm_attach(gfn_t *h, const char *uid, int nranks, int rank) nogil
    int gfn_allreduce_host(gfn_t *h, double *vals, int n) nogil
    int gfn_measure_peaks(int dev, double *fp64, double *fp32) nogil
    int gfn_mark(gfn_t *h, int which) nogil
    int gfn_selftest_arith(int dev, unsigned long long seed, long long n, unsigned long long *mism) nogil
    int gfn_mark_elapsed_ms(gfn_t *h, double *ms) nogil
    ctypedef struct gfn_residues_t:
        pass
    int gfn_residues_create(gfn_residues_t **out, int dev, int natoms, int nres, const int *apr, const int *rfa,
                            const double *masses, const double *charges) nogil
    void gfn_residues_destroy(gfn_residues_t *r) nogil
    int gfn_residues_com_dip(gfn_residues_t *r, int nframes, const double *coor, const double *com_in,
                             double *com_out, double *dip_out) nogil
    int gfn_residues_com_dip_device(gfn_residues_t *r, int nframes, const double *dcoor, double *dcom, double *ddip) nogil
    int gfn_attach_topology(gfn_t *h, int which, int natoms, int nres, const int *apr, const int *rfa,
                            const double *masses, const double *charges) nogil
    int gfn_enqueue_atom_frames(gfn_t *h, int nframes, const float *a1, long long s1, const float *a2, long long s2,
                                const double *boxes) nogil
    int gfn_topology_natoms(gfn_t *h, int *natoms1, int *natoms2) nogil
    ctypedef struct gfn_traj_t:
        pass
    int gfn_traj_open_dcd(gfn_traj_t **out, const char *path) nogil
    void gfn_traj_close(gfn_traj_t *t) nogil
    int gfn_traj_info(const gfn_traj_t *t, int *natoms, long long *nframes, int *has_cell) nogil
    int gfn_traj_read(gfn_traj_t *t, long long frame, const int *idx, int n, float *xyz, double *cell) nogil
    int gfn_enqueue_trajectory(gfn_t *h, gfn_traj_t *t, long long first, long long count, long long step,
                               const int *sel1, int nsel1, const int *sel2, int nsel2, int use_cell,
                               double *boxes_out, int nthreads) nogil
    const char *gfn_traj_last_error(const gfn_traj_t *t) nogil
    int gfn_dev_alloc(int dev, void **dptr, unsigned long long nbytes) nogil
    int gfn_dev_free(int dev, void *dptr) nogil
    int gfn_dev_upload(int dev, void *dptr, const void *host, unsigned long long nbytes) nogil


_FLAG_NAMES = {"cells": 512, "dense": 128, "nodense": 256, "exact_div": 64, "no_triangle": 32, "fp16": 16, "filter_fp16": 16, "per_particle": 1, "fp64": 2, "filter_fp64": 2, "discard_overflow": 4, "global": 8, "hist_global": 8}


def _parse_flags(spec):
    if spec is None:
        spec = os.environ.get("NEWANALYSIS_GFN_FLAGS", "")
    if isinstance(spec, int):
        return spec
    flags = 0
    for tok in str(spec).replace(" ", "").split(","):
        if tok:
            if tok not in _FLAG_NAMES:
                raise ValueError("unknown gfn flag %r (known: %s)" % (tok, ", ".join(sorted(_FLAG_NAMES))))
            flags |= _FLAG_NAMES[tok]
    return flags


def _parse_devices(spec):
    """NEWANALYSIS_GPUS: 'all', a count ('8'), or a comma list of ordinals ('0,2,3').  Default: device 0."""
    if spec is None:
        spec = os.environ.get("NEWANALYSIS_GPUS", "")
    if isinstance(spec, (list, tuple)):
        return [int(v) for v in spec]
    spec = str(spec).strip()
    if not spec:
        return [0]
    if spec == "all":
        return list(range(max(1, gfn_device_count())))
    if "," in spec:
        return [int(v) for v in spec.split(",") if v != ""]
    return list(range(int(spec)))


cdef object _raise(int rc, const gfn_t *h):
    cdef const char *msg = gfn_last_error(h)
    text = msg.decode("utf-8", "replace") if msg != NULL else ""
    if rc == GFN_EZERODIV:
        raise ZeroDivisionError("float division")
    if rc == GFN_EINVAL:
        raise ValueError(text)
    if rc == GFN_ENOMEM:
        raise MemoryError(text)
    raise RuntimeError("libgfn: " + text)


def device_count():
    return gfn_device_count()


def measure_peaks(int dev=0):
    """(fp64_tflops, fp32_tflops) of dependent-free FMA loops on device dev (roofline denominators)."""
    cdef double a = 0, b = 0
    cdef int rc
    with nogil:
        rc = gfn_measure_peaks(dev, &a, &b)
    if rc:
        _raise(rc, NULL)
    return a, b


def selftest_arith(int dev=0, unsigned long long seed=1, long long n_operands=1 << 30):
    """(division, sqrt, near-tie) mismatch counts of the kernel's in-range fp64 sequences vs the IEEE built-ins."""
    cdef unsigned long long m[3]
    cdef int rc
    with nogil:
        rc = gfn_selftest_arith(dev, seed, n_operands, m)
    if rc:
        _raise(rc, NULL)
    return int(m[0]), int(m[1]), int(m[2])


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it, every rank passes it to RDF.attach_comm)."""
    cdef char buf[128]
    cdef int rc = gfn_comm_unique_id(buf)
    if rc:
        _raise(rc, NULL)
    return bytes(buf[:128])


cdef class DeviceArray:
    """float64 array resident in HBM (the new library's analogue of the PyCUDA allocations the
    reference accepts as pre_*_gpu, gfunction.pyx:390-391)."""
    cdef void *ptr
    cdef readonly int device
    cdef readonly tuple shape
    cdef readonly unsigned long long nbytes
    cdef public object absmax      # max |value| of the last upload (None when produced on the device)

    def __cinit__(self):
        self.ptr = NULL
        self.absmax = None

    def __dealloc__(self):
        if self.ptr != NULL:
            gfn_dev_free(self.device, self.ptr)
            self.ptr = NULL

    def upload(self, arr):
        cdef np.ndarray[np.float64_t, ndim=1, mode="c"] flat = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
        if <unsigned long long> flat.nbytes != self.nbytes:
            raise ValueError("size mismatch")
        cdef int rc = gfn_dev_upload(self.device, self.ptr, <const void *> flat.data, self.nbytes)
        if rc:
            _raise(rc, NULL)
        self.absmax = float(np.abs(flat).max()) if flat.shape[0] else 0.0

    @property
    def address(self):
        return <unsigned long long> <size_t> self.ptr


class _PinnedOwner(object):
    """Frees a gfn_host_alloc block when the last numpy view of it is gone."""
    def __init__(self, size_t ptr):
        self.ptr = ptr

    def __del__(self):
        cdef size_t p = self.ptr
        self.ptr = 0
        if p:
            gfn_host_free(<void *> p)


def pinned_empty(shape, dtype=np.float64):
    """numpy array in page-locked host memory (cudaHostAlloc through libgfn): the source `RDF.calcFrames(..., pinned=True)`
    uploads from without a staging copy.  Freed with the last reference."""
    import ctypes
    dt = np.dtype(dtype)
    shape = tuple(int(v) for v in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    cdef unsigned long long nbytes = <unsigned long long> (int(np.prod(shape)) * dt.itemsize)
    cdef void *p = NULL
    cdef int rc = gfn_host_alloc(&p, nbytes)
    if rc != 0:
        raise MemoryError("libgfn: " + gfn_last_error(NULL).decode("utf-8", "replace"))
    buf = (ctypes.c_char * max(int(nbytes), 1)).from_address(<size_t> p)
    buf._gfn_owner = _PinnedOwner(<size_t> p)
    return np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)


def to_device(arr, int device=0):
    """Upload a float64 array ((n,3) or (nframes,n,3)) once; the result can be shared by several RDF objects."""
    a = np.ascontiguousarray(arr, dtype=np.float64)
    cdef DeviceArray d = DeviceArray()
    d.device = device
    d.shape = tuple(a.shape)
    d.nbytes = a.nbytes
    cdef int rc = gfn_dev_alloc(device, &d.ptr, d.nbytes)
    if rc:
        d.ptr = NULL
        _raise(rc, NULL)
    d.upload(a)
    return d


cdef DeviceArray _device_empty(tuple shape, int device):
    cdef DeviceArray d = DeviceArray()
    cdef unsigned long long n = 8
    for v in shape:
        n *= <unsigned long long> v
    d.device = device
    d.shape = shape
    d.nbytes = n
    cdef int rc = gfn_dev_alloc(device, &d.ptr, n)
    if rc:
        d.ptr = NULL
        _raise(rc, NULL)
    return d


cdef class ResidueMap:
    """Topology of a selection on the device (atoms per residue, first atom, masses, charges) for the GPU
    versions of helpers.comByResidue / dipByResidue / velcomByResidue
    (/root/reference/newanalysis/helpers/helpers.pyx:41-123).  Results are bit-identical to the reference."""
    cdef gfn_residues_t *r
    cdef readonly int natoms
    cdef readonly int nres
    cdef readonly int device
    cdef readonly bint has_charges

    def __cinit__(self):
        self.r = NULL

    def __init__(self, masses, apr, rfa, charges=None, int device=0):
        cdef np.ndarray[np.float64_t, ndim=1, mode="c"] m = np.ascontiguousarray(masses, dtype=np.float64)
        cdef np.ndarray[np.int32_t, ndim=1, mode="c"] a = np.ascontiguousarray(apr, dtype=np.int32)
        cdef np.ndarray[np.int32_t, ndim=1, mode="c"] f = np.ascontiguousarray(rfa, dtype=np.int32)
        cdef np.ndarray[np.float64_t, ndim=1, mode="c"] q = None
        if a.shape[0] != f.shape[0]:
            raise ValueError("apr and rfa must have one entry per residue")
        if charges is not None:
            q = np.ascontiguousarray(charges, dtype=np.float64)
            if q.shape[0] != m.shape[0]:
                raise ValueError("charges and masses must have one entry per atom")
        self.natoms = m.shape[0]
        self.nres = a.shape[0]
        self.device = device
        self.has_charges = q is not None
        cdef int rc = gfn_residues_create(&self.r, device, self.natoms, self.nres, <const int *> a.data, <const int *> f.data,
                                          <const double *> m.data, <const double *> q.data if q is not None else NULL)
        if rc:
            self.r = NULL
            _raise(rc, NULL)

    def __dealloc__(self):
        if self.r != NULL:
            gfn_residues_destroy(self.r)
            self.r = NULL

    def com(self, coor):
        """(natoms,3) or (nframes,natoms,3) host array -> centres of mass (…,nres,3)."""
        return self._run(coor, None, True, False)[0]

    def dip(self, coor, com=None):
        """Dipoles relative to `com` (default: the centres of mass)."""
        return self._run(coor, com, False, True)[1]

    def com_dip(self, coor):
        return self._run(coor, None, True, True)

    def _run(self, coor, com_in, bint want_com, bint want_dip):
        cdef np.ndarray[np.float64_t, ndim=1, mode="c"] c = np.ascontiguousarray(coor, dtype=np.float64).reshape(-1)
        sh = np.shape(coor)
        lead = tuple(sh[:len(sh) - 2])
        cdef int nframes = 1
        for v in lead:
            nframes *= v
        if c.shape[0] != <long long> nframes * self.natoms * 3:
            raise ValueError("coordinates must have shape (..., %d, 3)" % self.natoms)
        cdef np.ndarray[np.float64_t, ndim=1, mode="c"] ci = None
        cdef np.ndarray[np.float64_t, ndim=1, mode="c"] co = np.zeros(<long long> nframes * self.nres * 3)
        cdef np.ndarray[np.float64_t, ndim=1, mode="c"] do = None
        if com_in is not None:
            ci = np.ascontiguousarray(com_in, dtype=np.float64).reshape(-1)
            if ci.shape[0] != co.shape[0]:
                raise ValueError("com must have shape (..., %d, 3)" % self.nres)
        if want_dip:
            do = np.zeros(<long long> nframes * self.nres * 3)
        cdef int rc
        cdef const double *pc = <const double *> c.data
        cdef const double *pci = <const double *> ci.data if ci is not None else NULL
        cdef double *pco = <double *> co.data
        cdef double *pdo = <double *> do.data if do is not None else NULL
        with nogil:
            rc = gfn_residues_com_dip(self.r, nframes, pc, pci, pco, pdo)
        if rc:
            _raise(rc, NULL)
        shape = lead + (self.nres, 3)
        return (co.reshape(shape) if (want_com or com_in is None) else None), (do.reshape(shape) if want_dip else None)

    def com_dip_device(self, DeviceArray coor, bint want_dip=True):
        """Device-resident coordinates (nframes,natoms,3) -> (DeviceArray com, DeviceArray dip) of shape
        (nframes,nres,3), ready for RDF.calcFramesDevice / the pre_*_gpu arguments of RDF.calcFrame."""
        cdef long long per = 3 * <long long> self.natoms * 8
        if coor.nbytes % per:
            raise ValueError("device array does not hold whole frames of %d atoms" % self.natoms)
        cdef int nframes = <int> (coor.nbytes // per)
        cdef DeviceArray com = _device_empty((nframes, self.nres, 3), self.device)
        cdef DeviceArray dip = _device_empty((nframes, self.nres, 3), self.device) if want_dip else None
        cdef int rc
        cdef double *pd = <double *> dip.ptr if dip is not None else NULL
        with nogil:
            rc = gfn_residues_com_dip_device(self.r, nframes, <const double *> coor.ptr, <double *> com.ptr, pd)
        if rc:
            _raise(rc, NULL)
        return com, dip


def comByResidue(coor, masses, int nres, apr, rfa):
    """helpers.comByResidue(coor,masses,nres,apr,rfa) on the GPU (helpers.pyx:71-98), same result bit for bit.
    Builds a ResidueMap per call; keep a ResidueMap yourself when calling it every frame."""
    return ResidueMap(masses, np.asarray(apr)[:nres], np.asarray(rfa)[:nres]).com(np.asarray(coor, dtype=np.float64))


def velcomByResidue(vels, masses, int nres, apr, rfa):
    """helpers.velcomByResidue (helpers.pyx:41-68): the same arithmetic as comByResidue, applied to velocities."""
    return comByResidue(vels, masses, nres, apr, rfa)


def dipByResidue(coor, charges, masses, int nres, apr, rfa, com):
    """helpers.dipByResidue(coor,charges,masses,nres,apr,rfa,com) on the GPU (helpers.pyx:101-123)."""
    return ResidueMap(masses, np.asarray(apr)[:nres], np.asarray(rfa)[:nres], charges).dip(np.asarray(coor, dtype=np.float64), com)


cdef int _raise_traj(int rc, gfn_traj_t *t) except -1:
    msg = gfn_traj_last_error(t).decode("utf-8", "replace")
    if rc == GFN_EINVAL:
        raise ValueError(msg)
    raise RuntimeError(msg)


cdef class DCDFile:
    """CHARMM / NAMD DCD trajectory read by libgfn.so (no MDAnalysis needed, no GPU needed): the frame source of
    RDF.calcTrajectory.  n_atoms, n_frames, has_cell; read(frame, indices=None) -> (positions float32 (n,3) as stored,
    cell = [A, B, C, alpha, beta, gamma] or None)."""
    cdef gfn_traj_t *t
    cdef readonly int n_atoms
    cdef readonly long long n_frames
    cdef readonly bint has_cell
    cdef readonly str path

    def __cinit__(self):
        self.t = NULL

    def __init__(self, path):
        cdef bytes bp = os.fsencode(path)
        cdef int rc = gfn_traj_open_dcd(&self.t, bp)
        if rc:
            self.t = NULL
            _raise_traj(rc, NULL)
        cdef int na = 0, hc = 0
        cdef long long nf = 0
        gfn_traj_info(self.t, &na, &nf, &hc)
        self.n_atoms = na; self.n_frames = nf; self.has_cell = hc != 0
        self.path = str(path)

    def __dealloc__(self):
        if self.t != NULL:
            gfn_traj_close(self.t)
            self.t = NULL

    def close(self):
        if self.t != NULL:
            gfn_traj_close(self.t)
            self.t = NULL

    def __len__(self):
        return self.n_frames

    def read(self, long long frame, indices=None):
        if self.t == NULL:
            raise ValueError("trajectory is closed")
        cdef np.ndarray[np.int32_t, ndim=1, mode="c"] idx = None
        cdef int n = self.n_atoms
        if indices is not None:
            idx = np.ascontiguousarray(indices, dtype=np.int32)
            n = idx.shape[0]
        cdef np.ndarray[np.float32_t, ndim=2, mode="c"] xyz = np.empty((n, 3), dtype=np.float32)
        cdef np.ndarray[np.float64_t, ndim=1, mode="c"] cell = np.zeros(6)
        cdef int rc
        cdef const int *pi = <const int *> idx.data if idx is not None else NULL
        with nogil:
            rc = gfn_traj_read(self.t, frame, pi, n, <float *> xyz.data, <double *> cell.data)
        if rc:
            _raise_traj(rc, self.t)
        return xyz, (cell if self.has_cell else None)


cdef class _Backend:
    cdef gfn_t *h
    cdef int histo_n, n1, n2
    cdef public list keepalive

    def __cinit__(self):
        self.h = NULL
        self.keepalive = []

    def __dealloc__(self):
        if self.h != NULL:
            gfn_destroy(self.h)
            self.h = NULL

    cdef int check(self, int rc) except -1:
        if rc != GFN_OK:
            _raise(rc, self.h)
        return 0

    def open(self, int n1, int n2, int histo_n, float hmin, float hmax, double inv_dr, int mode_sel,
             np.ndarray[np.float64_t, ndim=1, mode="c"] box_dim, devices, unsigned flags):
        cdef np.ndarray[np.int32_t, ndim=1, mode="c"] devs = np.ascontiguousarray(devices, dtype=np.int32)
        cdef int rc
        cdef int ndev = devs.shape[0]
        rc = gfn_create(&self.h, n1, n2, histo_n, hmin, hmax, inv_dr, mode_sel, <const double *> box_dim.data,
                        <const int *> devs.data, ndev, flags)
        if rc != GFN_OK:
            self.h = NULL
            _raise(rc, NULL)
        self.histo_n = histo_n
        self.n1 = n1
        self.n2 = n2

    def open_shells(self, int n1, int n2, int histo_n, float hmin, float hmax, double inv_dr, int mode_sel,
                    np.ndarray[np.float64_t, ndim=1, mode="c"] box_dim, devices, unsigned flags,
                    int nshells, int layout, int apr_core, int apr_surround, int map_ld, long long map_elems, bint exclude_self):
        """Shell-resolved handle (gfn_create_shells); histogram rows hold histo_n*nshells bins."""
        cdef np.ndarray[np.int32_t, ndim=1, mode="c"] devs = np.ascontiguousarray(devices, dtype=np.int32)
        cdef gfn_shell_cfg cfg
        cfg.nshells = nshells; cfg.layout = layout; cfg.apr_core = apr_core; cfg.apr_surround = apr_surround
        cfg.map_ld = map_ld; cfg.map_elems = map_elems; cfg.exclude_self = 1 if exclude_self else 0
        cdef int ndev = devs.shape[0]
        cdef int rc = gfn_create_shells(&self.h, n1, n2, histo_n, hmin, hmax, inv_dr, mode_sel, <const double *> box_dim.data,
                                        <const int *> devs.data, ndev, flags, &cfg)
        if rc != GFN_OK:
            self.h = NULL
            _raise(rc, NULL)
        self.histo_n = histo_n * nshells
        self.n1 = n1
        self.n2 = n2

    cdef int enqueue_shells(self, const double *c1, const double *c2, const double *d1, const double *d2, const int *shell_map) except -1:
        cdef int rc
        with nogil:
            rc = gfn_enqueue_frames_shells(self.h, 1, c1, 0, c2, 0, d1, 0, d2, 0, NULL, shell_map, 0)
        return self.check(rc)

    def update_box(self, np.ndarray[np.float64_t, ndim=1, mode="c"] box_dim):
        self.check(gfn_update_box(self.h, <const double *> box_dim.data))

    cdef int enqueue(self, const double *c1, const double *c2, const double *d1, const double *d2) except -1:
        cdef int rc
        with nogil:
            rc = gfn_enqueue_frame(self.h, c1, c2, d1, d2)
        return self.check(rc)

    cdef int enqueue_many(self, int nframes, const double *c1, long long s1, const double *c2, long long s2,
                          const double *d1, long long s3, const double *d2, long long s4, const double *boxes, bint pinned=False) except -1:
        cdef int rc
        with nogil:
            if pinned:
                rc = gfn_enqueue_frames_pinned(self.h, nframes, c1, s1, c2, s2, d1, s3, d2, s4, boxes)
            else:
                rc = gfn_enqueue_frames(self.h, nframes, c1, s1, c2, s2, d1, s3, d2, s4, boxes)
        return self.check(rc)

    def enqueue_device(self, int nframes, DeviceArray c1, DeviceArray c2, DeviceArray d1, DeviceArray d2, boxes=None, long long first=0):
        cdef long long o1 = first * 3 * self.n1, o2 = first * 3 * self.n2
        if first < 0 or <unsigned long long> ((first + nframes) * 3 * self.n1 * 8) > c1.nbytes or <unsigned long long> ((first + nframes) * 3 * self.n2 * 8) > c2.nbytes:
            raise ValueError("frames [%d, %d) exceed the device array" % (first, first + nframes))
        cdef const double *p1 = <const double *> c1.ptr + o1
        cdef const double *p2 = <const double *> c2.ptr + o2
        cdef const double *q1 = (<const double *> d1.ptr + o1) if d1 is not None else NULL
        cdef const double *q2 = (<const double *> d2.ptr + o2) if d2 is not None else NULL
        cdef np.ndarray[np.float64_t, ndim=2, mode="c"] bx
        cdef const double *pb = NULL
        if boxes is not None:
            bx = np.ascontiguousarray(boxes, dtype=np.float64).reshape(nframes, 6)
            pb = <const double *> bx.data
        cdef long long s1 = 3 * self.n1, s2 = 3 * self.n2
        cdef int rc
        with nogil:
            rc = gfn_enqueue_device_frames(self.h, nframes, p1, s1, p2, s2, q1, s1, q2, s2, pb)
        # the arrays must outlive the kernels that read them.  A long run that never reads out would pile up one entry per
        # call: beyond a few dozen, wait for the device once and let them go.
        cdef Py_ssize_t nk = len(self.keepalive)
        if nk == 0 or self.keepalive[nk - 1] != (c1, c2, d1, d2):
            self.keepalive.append((c1, c2, d1, d2))
        self.check(rc)
        if len(self.keepalive) > 64:
            self.sync()

    def attach_topology(self, int which, masses, apr, rfa, charges=None):
        cdef np.ndarray[np.float64_t, ndim=1, mode="c"] m = np.ascontiguousarray(masses, dtype=np.float64)
        cdef np.ndarray[np.int32_t, ndim=1, mode="c"] a = np.ascontiguousarray(apr, dtype=np.int32)
        cdef np.ndarray[np.int32_t, ndim=1, mode="c"] f = np.ascontiguousarray(rfa, dtype=np.int32)
        cdef np.ndarray[np.float64_t, ndim=1, mode="c"] q = None
        if a.shape[0] != f.shape[0]:
            raise ValueError("apr and rfa must have one entry per residue")
        if charges is not None:
            q = np.ascontiguousarray(charges, dtype=np.float64)
            if q.shape[0] != m.shape[0]:
                raise ValueError("charges and masses must have one entry per atom")
        self.check(gfn_attach_topology(self.h, which, <int> m.shape[0], <int> a.shape[0], <const int *> a.data, <const int *> f.data,
                                       <const double *> m.data, <const double *> q.data if q is not None else NULL))

    cdef int enqueue_atoms(self, int nframes, const float *a1, long long s1, const float *a2, long long s2, const double *boxes) except -1:
        cdef int rc
        with nogil:
            rc = gfn_enqueue_atom_frames(self.h, nframes, a1, s1, a2, s2, boxes)
        return self.check(rc)

    def enqueue_trajectory(self, DCDFile traj, long long first, long long count, long long step, sel1, sel2, int use_cell, int nthreads):
        cdef np.ndarray[np.int32_t, ndim=1, mode="c"] s1 = np.ascontiguousarray(sel1, dtype=np.int32)
        cdef np.ndarray[np.int32_t, ndim=1, mode="c"] s2 = None
        cdef np.ndarray[np.float64_t, ndim=2, mode="c"] boxes = np.zeros((max(count, 1), 3))
        if sel2 is not None:
            s2 = np.ascontiguousarray(sel2, dtype=np.int32)
        if traj.t == NULL:
            raise ValueError("trajectory is closed")
        cdef const int *p2 = <const int *> s2.data if s2 is not None else NULL
        cdef int n1 = s1.shape[0]
        cdef int n2 = s2.shape[0] if s2 is not None else 0
        cdef int rc
        with nogil:
            rc = gfn_enqueue_trajectory(self.h, traj.t, first, count, step, <const int *> s1.data, n1, p2, n2,
                                        use_cell, <double *> boxes.data, nthreads)
        if rc:
            msg = gfn_traj_last_error(traj.t).decode("utf-8", "replace")
            if rc == GFN_EINVAL:
                raise ValueError(msg)
            raise RuntimeError(msg)
        return boxes[:count]

    def flush(self):
        cdef int rc
        with nogil:
            rc = gfn_flush(self.h)
        self.check(rc)

    def sync(self):
        cdef int rc
        with nogil:
            rc = gfn_sync(self.h)
        del self.keepalive[:]
        self.check(rc)

    def readout(self):
        """-> (hist[7*histo_n], pairw[histo_n], overflow)"""
        cdef np.ndarray[np.float64_t, ndim=1, mode="c"] hist = np.zeros(7 * self.histo_n, dtype=np.float64)
        cdef np.ndarray[np.float64_t, ndim=1, mode="c"] pw = np.zeros(self.histo_n, dtype=np.float64)
        cdef unsigned long long ovf = 0
        cdef int rc
        with nogil:
            rc = gfn_readout(self.h, <double *> hist.data, <double *> pw.data, &ovf)
        del self.keepalive[:]
        self.check(rc)
        return hist, pw, int(ovf)

    def readout_per_particle(self):
        cdef np.ndarray[np.float64_t, ndim=1, mode="c"] hist = np.zeros(7 * <long long> self.n1 * self.histo_n, dtype=np.float64)
        cdef int rc
        with nogil:
            rc = gfn_readout_per_particle(self.h, <double *> hist.data)
        self.check(rc)
        return hist

    cdef int add_per_particle(self, double *dst, unsigned long long first, unsigned long long count) except -1:
        cdef int rc
        with nogil:
            rc = gfn_readout_per_particle_add(self.h, dst, first, count)
        self.check(rc)
        return 0

    def scale(self, double v):
        self.check(gfn_scale(self.h, v))

    def reset(self):
        self.check(gfn_reset(self.h))

    def mark(self, int which):
        self.check(gfn_mark(self.h, which))

    def mark_elapsed_ms(self):
        cdef double ms = 0
        cdef int rc
        with nogil:
            rc = gfn_mark_elapsed_ms(self.h, &ms)
        self.check(rc)
        return ms

    def attach_comm(self, bytes uid, int nranks, int rank):
        if len(uid) != 128:
            raise ValueError("unique id must be 128 bytes")
        cdef const char *p = uid
        cdef int rc
        with nogil:
            rc = gfn_comm_attach(self.h, p, nranks, rank)
        self.check(rc)

    def allreduce_host(self, values):
        """Sum of a few host doubles over the ranks of the attached communicator (collective)."""
        cdef np.ndarray[np.float64_t, ndim=1, mode="c"] v = np.ascontiguousarray(values, dtype=np.float64).copy()
        cdef int n = v.shape[0]
        cdef int rc
        with nogil:
            rc = gfn_allreduce_host(self.h, <double *> v.data, n)
        self.check(rc)
        return v

    def stats(self):
        cdef gfn_stats_t s
        self.check(gfn_stats(self.h, &s))
        return dict(pair_evals=s.pair_evals, survivors=s.survivors, in_range=s.in_range, kernel_ms=s.kernel_ms,
                    kernel_launches=s.kernel_launches, pair_launches=s.pair_launches, frames=s.frames,
                    h2d_bytes=s.h2d_bytes, overflow=s.overflow, filter=("fp32", "fp64", "fp16x2")[s.filter_fp64],
                    hist_mode=("smem_warp_replicas", "global_atomics", "per_particle")[s.hist_mode],
                    warps_per_block=s.warps_per_block, blocks_per_sm=s.blocks_per_sm, grid=s.grid,
                    smem_bytes=s.smem_bytes, batch_frames=s.batch_frames, ndev=s.ndev, hist_replicas=s.hist_replicas,
                    kernel=("ring", "dense", "count")[s.dense], exact_evals=s.exact_evals, selfcheck_bad=s.selfcheck_bad,
                    cells=bool(s.cells), tile_visits=s.tile_visits, tile_visits_dense=s.tile_visits_dense,
                    frames_max_device=s.frames_max_device, frames_min_device=s.frames_min_device,
                    kernel_ms_max_device=s.kernel_ms_max_device)


_ts_cache = {}


def calcHistoTS(core_xyz, surr_xyz, np.ndarray[np.float64_t, ndim=3] result, double boxl, int t,
                float histo_min, float histo_max, double invdx):
    """RDF time series on the GPU: result[t, i, pos] += 1 for every surround site within range of core site i
    (reference: def calcHistoTS, gfunction.pyx:168-192; cubic box, all n1 x n2 pairs, no dipoles).

    Same signature and result as the reference (pinned by tests/golden/ts.npz, generated by the unmodified reference
    function).  A pair with pos >= nbins (quirk Q3) lands in result[t, i+1, pos-nbins] like the reference's unchecked
    flat index does; for the last core row that is the first row of frame t+1 when `result` is C-contiguous and has such
    a frame - beyond the array (where the reference writes out of bounds) the pair is dropped."""
    c1 = np.ascontiguousarray(core_xyz, dtype=np.float64)
    c2 = c1 if surr_xyz is core_xyz else np.ascontiguousarray(surr_xyz, dtype=np.float64)
    cdef int n1 = c1.shape[0], n2 = c2.shape[0], nb = result.shape[2]
    if result.shape[1] != n1:
        raise ValueError("result must have shape (nframes, %d, nbins)" % n1)
    if t < 0 or t >= result.shape[0]:
        raise IndexError("frame index out of range")
    key = (n1, n2, nb, float(np.float32(histo_min)), float(np.float32(histo_max)), float(invdx), float(boxl))
    be = _ts_cache.get(key)
    if be is None:
        if gfn_device_count() < 1:
            raise RuntimeError("newanalysis.gfunction (B200 backend): no CUDA device available and there is no CPU fallback")
        if len(_ts_cache) > 8:
            _ts_cache.clear()
        be = _Backend()
        box = np.array([boxl, boxl, boxl, boxl / 2.0, boxl / 2.0, boxl / 2.0], dtype=np.float64)
        be.open(n1, n2, nb, histo_min, histo_max, invdx, 1, box, [0], 1 | 32)      # per-particle, no triangle
        _ts_cache[key] = be
    cdef _Backend b = be
    b.reset()
    cdef np.ndarray[np.float64_t, ndim=2, mode="c"] a1 = c1
    cdef np.ndarray[np.float64_t, ndim=2, mode="c"] a2 = c2
    b.enqueue(<const double *> a1.data, <const double *> a2.data, NULL, NULL)
    # Q3 spill of the last core row: the device keeps it right behind the frame's slab (first row of the next mode), and
    # in a C-contiguous result that is the first row of frame t+1 - one add over (n1 + 1) rows covers both
    cdef unsigned long long slab = <unsigned long long> n1 * nb
    cdef bint spill = t + 1 < result.shape[0]
    cdef np.ndarray[np.float64_t, ndim=2] tmp
    if result.flags.c_contiguous:
        b.add_per_particle(<double *> result.data + <unsigned long long> t * slab, 0, slab + (nb if spill else 0))
    else:
        tmp = np.zeros((n1, nb))
        b.add_per_particle(<double *> tmp.data, 0, slab)
        result[t] += tmp


def calcHistoTSVoro(core_xyz, surr_xyz, np.ndarray[np.float64_t, ndim=4] result, ds, int maxshell,
                    double boxl, int t, float histo_min, float histo_max, double invdx):
    """Shell-resolved RDF time series on the GPU: result[t, i, shell-1, pos] += 1 with shell = min(ds[t, i, j], maxshell)
    (reference: def calcHistoTSVoro, gfunction.pyx:195-224; ds is the char Delaunay-distance array [frame][core][surround]).

    Same signature and result as the reference.  ds[t, i, j] == 0 addresses the last shell (index -1 wraps in the
    reference); a pair past the last radial bin lands where the reference's flat index points, except beyond the
    frame's slab, where the reference writes into the next frame (or past the array) and this version drops it."""
    c1 = np.ascontiguousarray(core_xyz, dtype=np.float64)
    c2 = c1 if surr_xyz is core_xyz else np.ascontiguousarray(surr_xyz, dtype=np.float64)
    cdef int n1 = c1.shape[0], n2 = c2.shape[0], nb = result.shape[3]
    if result.shape[1] != n1 or result.shape[2] != maxshell:
        raise ValueError("result must have shape (nframes, %d, %d, nbins)" % (n1, maxshell))
    if t < 0 or t >= result.shape[0]:
        raise IndexError("frame index out of range")
    cdef np.ndarray[np.int32_t, ndim=2, mode="c"] shells = np.ascontiguousarray(np.asarray(ds[t]), dtype=np.int32)
    if shells.shape[0] != n1 or shells.shape[1] != n2:
        raise ValueError("ds[t] must have shape (%d, %d)" % (n1, n2))
    key = ("voro", n1, n2, nb, maxshell, float(np.float32(histo_min)), float(np.float32(histo_max)), float(invdx), float(boxl))
    be = _ts_cache.get(key)
    if be is None:
        if gfn_device_count() < 1:
            raise RuntimeError("newanalysis.gfunction (B200 backend): no CUDA device available and there is no CPU fallback")
        if len(_ts_cache) > 8:
            _ts_cache.clear()
        be = _Backend()
        box = np.array([boxl, boxl, boxl, boxl / 2.0, boxl / 2.0, boxl / 2.0], dtype=np.float64)
        be.open_shells(n1, n2, nb, histo_min, histo_max, invdx, 1, box, [0], 1 | 32,       # per-particle, no triangle
                       maxshell, 2, 1, 1, n2, <long long> n1 * n2, False)
        _ts_cache[key] = be
    cdef _Backend b = be
    b.reset()
    cdef np.ndarray[np.float64_t, ndim=2, mode="c"] a1 = c1
    cdef np.ndarray[np.float64_t, ndim=2, mode="c"] a2 = c2
    b.enqueue_shells(<const double *> a1.data, <const double *> a2.data, NULL, NULL, <const int *> shells.data)
    cdef unsigned long long slab = <unsigned long long> n1 * maxshell * nb
    if result.flags.c_contiguous:
        b.add_per_particle(<double *> result.data + <unsigned long long> t * slab, 0, slab)
    else:
        result[t] += b.readout_per_particle()[:slab].reshape(n1, maxshell, nb)


class RDF(object):
    """
    class RDF

    A container object for calculating radial distribution functions (B200 backend).
    """

    def __init__(self, modes, histo_min, histo_max, histo_dr, sel1_n, sel2_n, sel1_nmol, sel2_nmol, box_x, box_y=None, box_z=None, use_cuda=False, norm_volume=None, gfn_flags=None, devices=None):
        """
        RDF.__init__(modes, histo_min, histo_max, histo_dr, sel1_n, sel2_n, sel1_nmol, sel2_nmol, box_x, box_y=None, box_z=None, use_cuda=False)

        Same arguments as the reference (gfunction.pyx:234).  use_cuda is accepted and ignored: this
        class always runs on the GPU.  gfn_flags / devices are optional extensions (see module docstring).
        """
        # the histogram always lives on the device; readout arrays are float64 like the reference CPU path (:268-271)
        self.cuda = True
        self.dtype = np.float64

        cdef int mode_sel = 0
        for mode in modes:                                   # :274-288, unknown strings are ignored
            if mode == "all": mode_sel = mode_sel | 127
            if mode == "000": mode_sel = mode_sel | 1
            if mode == "110": mode_sel = mode_sel | 2
            if mode == "101": mode_sel = mode_sel | 4
            if mode == "011": mode_sel = mode_sel | 8
            if mode == "220": mode_sel = mode_sel | 16
            if mode == "202": mode_sel = mode_sel | 32
            if mode == "022": mode_sel = mode_sel | 64
        self.mode_sel = np.int32(mode_sel)

        self.histo_min = np.float32(histo_min)               # :291-295
        self.histo_max = np.float32(histo_max)
        self.histo_dr = np.round(self.dtype(histo_dr), 5)
        self.histo_inv_dr = self.dtype(1.0 / self.histo_dr)
        self.histo_n = np.int32((histo_max - histo_min) / histo_dr + 0.5)

        self.box_dim = np.zeros(6, dtype=self.dtype)         # :297-311 (box_y / box_z falsy -> cubic)
        self.box_dim[0] = box_x
        self.box_dim[3] = box_x / 2.0
        if box_y:
            self.box_dim[1] = box_y
            self.box_dim[4] = box_y / 2.0
        else:
            self.box_dim[1] = self.box_dim[0]
            self.box_dim[4] = self.box_dim[3]
        if box_z:
            self.box_dim[2] = box_z
            self.box_dim[5] = box_z / 2.0
        else:
            self.box_dim[2] = self.box_dim[0]
            self.box_dim[5] = self.box_dim[3]
        self.volume = self.box_dim[0] * self.box_dim[1] * self.box_dim[2]
        self.volume_list = []
        self.norm_volume = norm_volume

        self.n1 = np.int32(sel1_n)
        self.n2 = np.int32(sel2_n)
        self.nmol1 = sel1_nmol
        self.nmol2 = sel2_nmol
        self.oneistwo = None
        self.ctr = 0

        self.histogram_out = np.zeros(self.histo_n * 7, dtype=np.float64)   # :333

        # filter choice: explicit flag, or "auto" (default): fp32 until the first host frame shows that the
        # coordinates stay within a few box lengths, then the cheaper packed-fp16 pre-filter (results identical)
        spec = gfn_flags if gfn_flags is not None else os.environ.get("NEWANALYSIS_GFN_FLAGS", "")
        self._auto_filter = not isinstance(spec, int) and not any(tok in str(spec) for tok in ("fp16", "fp32", "fp64"))
        if not isinstance(spec, int):
            spec = ",".join(tok for tok in str(spec).replace(" ", "").split(",") if tok and tok != "fp32")
        self._flags = _parse_flags(spec)
        self._devices = _parse_devices(devices)
        if gfn_device_count() < 1:
            raise RuntimeError("newanalysis.gfunction (B200 backend): no CUDA device available and there is no CPU fallback")
        self._open_backend()

    def _open_backend(self):
        self._be = _Backend()
        self._be.open(int(self.n1), int(self.n2), int(self.histo_n), self.histo_min, self.histo_max,
                      float(self.histo_inv_dr), int(self.mode_sel), self.box_dim, self._devices, self._flags)

    def _auto_pick_filter(self, c1, c2, absmax=None):
        """First frame of an "auto" handle: switch to the fp16 pre-filter when every coordinate lies within two
        box lengths of the origin (wrapped trajectories); far-away (unwrapped) coordinates keep fp32."""
        self._auto_filter = False
        if self._flags & (2 | 16):
            self._finish_comm()
            return
        lmax = float(max(self.box_dim[0], self.box_dim[1], self.box_dim[2]))
        if absmax is None:
            absmax = max(float(np.abs(c1).max()), float(np.abs(c2).max()))
        if absmax <= 2.0 * lmax and float(self.histo_max) >= 0.04 * lmax:
            self._flags |= 16
            self._open_backend()          # nothing has been enqueued yet
            topo = getattr(self, "_topology", None)
            if topo is not None:
                self._be.attach_topology(0, *topo[:4])
                if topo[4] is not None:
                    self._be.attach_topology(1, *topo[4:])
        self._finish_comm()

    # ------------------------------------------------------------------ reference API
    def update_box(self, box_x, box_y, box_z):
        """
        update_box(box_x,box_y,box_z)

        Update the box dimensions, when analysing NpT trajectories (:367-383).
        """
        self.box_dim[0] = box_x
        self.box_dim[3] = box_x / 2.0
        self.box_dim[1] = box_y
        self.box_dim[4] = box_y / 2.0
        self.box_dim[2] = box_z
        self.box_dim[5] = box_z / 2.0
        self._be.update_box(self.box_dim)
        self.volume = self.box_dim[0] * self.box_dim[1] * self.box_dim[2]

    def calcFrame(self,
                  np.ndarray[np.float64_t, ndim=2, mode="c"] coor_sel1,
                  np.ndarray[np.float64_t, ndim=2, mode="c"] coor_sel2,
                  np.ndarray[np.float64_t, ndim=2, mode="c"] dip_sel1 = None,
                  np.ndarray[np.float64_t, ndim=2, mode="c"] dip_sel2 = None,
                  pre_coor_sel1_gpu=None, pre_coor_sel2_gpu=None,
                  pre_dip_sel1_gpu=None, pre_dip_sel2_gpu=None):
        """
        calcFrame(coor_sel1, coor_sel2, dip_sel1=None, dip_sel2=None,
                  pre_coor_sel1_gpu=None, pre_coor_sel2_gpu=None,
                  pre_dip_sel1_gpu=None, pre_dip_sel2_gpu=None)

        Pass the data of the current frame (:385-464).  The arrays are copied before the call returns.
        pre_*_gpu take `DeviceArray` objects from `to_device` (shared uploads for several RDF objects).
        """
        cdef int m = <int> self.mode_sel
        if dip_sel1 is None and pre_dip_sel1_gpu is None and (m & 2 or m & 4 or m & 16 or m & 32):
            raise TypeError("No dipole moment array for the fist selection was passed, although the requested gfunctions need it!")
        if dip_sel2 is None and pre_dip_sel2_gpu is None and (m & 2 or m & 8 or m & 16 or m & 64):
            raise TypeError("No dipole moment array for the second selection was passed, although the requested gfunctions need it!")

        if coor_sel1.shape[0] != self.n1 or coor_sel1.shape[1] != 3 or coor_sel2.shape[0] != self.n2 or coor_sel2.shape[1] != 3:
            raise ValueError("coordinate arrays must have shapes (%d,3) and (%d,3)" % (self.n1, self.n2))
        if dip_sel1 is not None and (dip_sel1.shape[0] != self.n1 or dip_sel1.shape[1] != 3):
            raise ValueError("dip_sel1 must have shape (%d,3)" % self.n1)
        if dip_sel2 is not None and (dip_sel2.shape[0] != self.n2 or dip_sel2.shape[1] != 3):
            raise ValueError("dip_sel2 must have shape (%d,3)" % self.n2)

        if self.oneistwo is None:                            # :417-421
            if (coor_sel1[0] == coor_sel2[0]).all() and self.nmol1 == self.nmol2:
                self.oneistwo = True
            else:
                self.oneistwo = False

        if self._auto_filter and self.ctr == 0:
            self._auto_pick_filter(coor_sel1, coor_sel2)
        self.ctr += 1                                        # :424
        self.volume_list.append(self.volume)                 # :427

        cdef _Backend be = self._be
        if pre_coor_sel1_gpu is not None or pre_coor_sel2_gpu is not None or pre_dip_sel1_gpu is not None or pre_dip_sel2_gpu is not None:
            # arrays not shared by the caller are staged in per-object device buffers that are reused every frame
            # (DeviceArray.upload waits for the frames that still read them), so a long run allocates nothing per frame
            given = [a for a in (pre_coor_sel1_gpu, pre_coor_sel2_gpu, pre_dip_sel1_gpu, pre_dip_sel2_gpu) if a is not None]
            dev = given[0].device
            c1 = pre_coor_sel1_gpu if pre_coor_sel1_gpu is not None else self._stage(0, coor_sel1, dev)
            if pre_coor_sel2_gpu is not None:
                c2 = pre_coor_sel2_gpu
            elif coor_sel2 is coor_sel1:
                c2 = c1
            else:
                c2 = self._stage(1, coor_sel2, dev)
            d1 = pre_dip_sel1_gpu if pre_dip_sel1_gpu is not None else (self._stage(2, dip_sel1, dev) if dip_sel1 is not None else None)
            if pre_dip_sel2_gpu is not None:
                d2 = pre_dip_sel2_gpu
            elif dip_sel2 is dip_sel1:
                d2 = d1
            else:
                d2 = self._stage(3, dip_sel2, dev) if dip_sel2 is not None else None
            be.enqueue_device(1, c1, c2, d1, d2, self.box_dim.reshape(1, 6))
            return

        be.enqueue(<const double *> coor_sel1.data, <const double *> coor_sel2.data,
                   <const double *> dip_sel1.data if dip_sel1 is not None else NULL,
                   <const double *> dip_sel2.data if dip_sel2 is not None else NULL)

    def _stage(self, int slot, arr, int dev):
        """Per-object staging buffer `slot` (0 c1, 1 c2, 2 d1, 3 d2) on device `dev`, refilled with arr."""
        st = self.__dict__.setdefault("_staged", {})
        cdef DeviceArray d = st.get((slot, dev))
        if d is None or d.nbytes != <unsigned long long> arr.nbytes:
            d = to_device(arr, dev)
            st[(slot, dev)] = d
        else:
            d.upload(arr)
        return d

    def _addHisto(self, np.ndarray[np.float64_t, ndim=1] histo_out, np.ndarray[np.float64_t, ndim=1] histo):
        """
        Add the histograms of all core particles (:466-480).  `histo` may be the per-particle array
        (7*n1*histo_n) or the already summed [mode][bin] array (7*histo_n) this backend keeps.
        """
        cdef int hn = self.histo_n
        cdef long long n1 = self.n1
        if histo.shape[0] == 7 * hn:
            histo_out += histo
        elif histo.shape[0] == 7 * n1 * hn:
            histo_out += histo.reshape(7, n1, hn).sum(axis=1).reshape(-1)
        else:
            raise ValueError("histogram has %d elements, expected %d or %d" % (histo.shape[0], 7 * hn, 7 * n1 * hn))

    def _normHisto(self):
        """
        Normalize the histogram (:483-511); host code, same arithmetic as the reference.
        """
        ctr = self.ctr
        if getattr(self, "_nranks", 1) > 1:
            # frames are spread over processes: the histogram read out is the sum over all ranks, so the frame count and
            # the mean volume must be taken over all ranks as well (collective, like the readout)
            tot = self._be.allreduce_host([float(self.ctr), float(sum(self.volume_list))])
            ctr = int(tot[0])
            volume = self.norm_volume if self.norm_volume else tot[1] / tot[0]
        elif self.norm_volume:
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

        cdef int i, j
        for i in range(self.histo_n):
            r = self.histo_min + float(i) * self.histo_dr
            r_out = r + self.histo_dr
            slice_vol = 4.0 / 3.0 * PI * (r_out**3 - r**3)
            norm = rho * slice_vol * float(ctr)
            for j in range(7):
                self.histogram_out[j * self.histo_n + i] /= norm

    def scale(self, value):
        """:513-515"""
        self.histogram_out[:] *= value
        self._be.scale(float(value))

    def write(self, filename="rdf"):
        """
        RDF.write(filename="rdf")

        Norm the calculated histograms and write them to <filename>_g000.dat ... (:517-550).
        Like the reference, calling it twice accumulates into the already normalised histogram_out.
        """
        self._addHisto(self.histogram_out, self.raw_histogram())
        self._normHisto()

        names = ("000", "110", "101", "011", "220", "202", "022")
        files = {}
        for k in range(7):
            if self.mode_sel & (1 << k):
                files[k] = open(filename + "_g" + names[k] + ".dat", 'w')
        for i in range(self.histo_n):
            r = self.histo_min + i * self.histo_dr + self.histo_dr * 0.5
            for k in files:
                files[k].write("%5.5f\t%5.5f\n" % (r, self.histogram_out[k * self.histo_n + i]))
        for k in files:
            files[k].close()

    def resetArrays(self):
        """
        Reset histogram arrays to zero (:552-557; ctr and volume_list are kept, like the reference)
        """
        self.histogram_out[:] = 0.0
        self._be.reset()

    def printInfo(self):
        """
        Print some information about the RDF container (:560-566).
        """
        print("Modes: ", self.mode_sel)
        print("Frames read: ", self.ctr)

    # ------------------------------------------------------------------ extensions
    @property
    def histogram(self):
        if self._flags & 1:
            return self._be.readout_per_particle()
        return self.raw_histogram()

    def raw_histogram(self):
        """Un-normalised [mode][bin] sums accumulated since the last resetArrays (7*histo_n float64)."""
        if self._auto_filter and self.ctr == 0:
            self._auto_filter = False
        self._finish_comm()
        hist, pw, ovf = self._be.readout()
        self.pair_weight = pw
        self.overflow = ovf
        return hist

    def calcFrames(self, coor_sel1, coor_sel2, dip_sel1=None, dip_sel2=None, boxes=None, pinned=False):
        """Many frames per call: arrays of shape (nframes, n, 3); boxes (nframes, 3) for NpT or None.

        pinned=True: the arrays are page-locked (`pinned_empty`) float64 C arrays and stay unchanged until the next
        sync() / histogram readout: the GPU's copy engine reads them directly, nothing is staged on the host."""
        if pinned:
            for arr in (coor_sel1, coor_sel2, dip_sel1, dip_sel2):
                if arr is not None and not (isinstance(arr, np.ndarray) and arr.dtype == np.float64 and arr.flags.c_contiguous):
                    raise TypeError("pinned=True takes C-contiguous float64 arrays made by pinned_empty (no conversion copy is allowed)")
        cdef np.ndarray[np.float64_t, ndim=3, mode="c"] c1 = np.ascontiguousarray(coor_sel1, dtype=np.float64)
        cdef np.ndarray[np.float64_t, ndim=3, mode="c"] c2 = c1 if coor_sel2 is coor_sel1 else np.ascontiguousarray(coor_sel2, dtype=np.float64)
        cdef np.ndarray[np.float64_t, ndim=3, mode="c"] d1 = None
        cdef np.ndarray[np.float64_t, ndim=3, mode="c"] d2 = None
        cdef np.ndarray[np.float64_t, ndim=2, mode="c"] bx = None
        cdef int m = <int> self.mode_sel
        cdef int nf = c1.shape[0]
        if dip_sel1 is not None:
            d1 = np.ascontiguousarray(dip_sel1, dtype=np.float64)
        if dip_sel2 is not None:
            d2 = d1 if dip_sel2 is dip_sel1 else np.ascontiguousarray(dip_sel2, dtype=np.float64)
        if d1 is None and (m & 2 or m & 4 or m & 16 or m & 32):
            raise TypeError("No dipole moment array for the fist selection was passed, although the requested gfunctions need it!")
        if d2 is None and (m & 2 or m & 8 or m & 16 or m & 64):
            raise TypeError("No dipole moment array for the second selection was passed, although the requested gfunctions need it!")
        if c1.shape[1] != self.n1 or c1.shape[2] != 3 or c2.shape[1] != self.n2 or c2.shape[2] != 3 or c2.shape[0] != nf:
            raise ValueError("coordinate arrays must have shapes (nframes,%d,3) and (nframes,%d,3)" % (self.n1, self.n2))
        if nf == 0:
            return
        if self.oneistwo is None:
            self.oneistwo = bool((c1[0, 0] == c2[0, 0]).all() and self.nmol1 == self.nmol2)
        if boxes is not None:
            b3 = np.asarray(boxes, dtype=np.float64).reshape(nf, 3)
            bx = np.ascontiguousarray(np.concatenate([b3, b3 / 2.0], axis=1))
            vols = list(b3[:, 0] * b3[:, 1] * b3[:, 2])
        else:
            vols = [self.volume] * nf
        if self._auto_filter and self.ctr == 0:
            self._auto_pick_filter(c1[0], c2[0])
        cdef _Backend be = self._be
        be.enqueue_many(nf, <const double *> c1.data, 3 * <long long> self.n1, <const double *> c2.data, 3 * <long long> self.n2,
                        <const double *> d1.data if d1 is not None else NULL, 3 * <long long> self.n1,
                        <const double *> d2.data if d2 is not None else NULL, 3 * <long long> self.n2,
                        <const double *> bx.data if bx is not None else NULL, bool(pinned))
        self.ctr += nf
        self.volume_list.extend(vols)
        if pinned:
            be.keepalive.append(("pinned", c1, c2, d1, d2))          # until the next sync / readout
        if boxes is not None:
            self.update_box(*[float(v) for v in b3[nf - 1]])

    def calcFramesDevice(self, int nframes, c1, c2, d1=None, d2=None, boxes=None, first=0):
        """Frames already resident in HBM (`DeviceArray`s of shape (>=first+nframes,n,3)); no host->device copy."""
        if self.oneistwo is None:
            self.oneistwo = bool(c1 is c2 and self.nmol1 == self.nmol2)
        if self._auto_filter and self.ctr == 0:
            if c1.absmax is not None and c2.absmax is not None:
                self._auto_pick_filter(None, None, max(c1.absmax, c2.absmax))
            else:
                self._auto_filter = False
                self._finish_comm()
        self.ctr += nframes
        self.volume_list.extend([self.volume] * nframes)
        self._be.enqueue_device(nframes, c1, c2, d1, d2, boxes, first)

    def attach_topology(self, masses, apr, rfa, charges=None, masses2=None, apr2=None, rfa2=None, charges2=None):
        """Let the RDF object take raw atom positions (calcFrameAtoms / calcFramesAtoms): masses / charges per atom,
        apr / rfa per residue (atoms per residue, first atom - the arrays the reference's helpers.comByResidue and
        dipByResidue take, helpers.pyx:71-123).  The second set describes selection 2; without it selection 2 uses
        the topology of selection 1 (sel1_n == sel2_n)."""
        if self._auto_filter and self.ctr == 0:
            self._topology = (masses, apr, rfa, charges, masses2, apr2, rfa2, charges2)     # re-applied if the backend is re-opened
        self._be.attach_topology(0, masses, apr, rfa, charges)
        if masses2 is not None:
            self._be.attach_topology(1, masses2, apr2, rfa2, charges2)
        self._natoms = (len(masses), len(masses2) if masses2 is not None else len(masses))
        self._first_res = (int(np.asarray(apr)[0]), masses2 is None)

    def calcFrameAtoms(self, pos_sel1, pos_sel2=None):
        """One frame from atom positions (float32, (natoms,3), what MDAnalysis' AtomGroup.positions returns): the same
        result, bit for bit, as  c = comByResidue(pos.astype(float64), ...); d = dipByResidue(..., c);
        calcFrame(c1, c2, d1, d2)  of the reference workflow, with centres of mass and dipoles made on the device."""
        p1 = np.ascontiguousarray(pos_sel1, dtype=np.float32)
        p2 = p1 if (pos_sel2 is None or pos_sel2 is pos_sel1) else np.ascontiguousarray(pos_sel2, dtype=np.float32)
        self.calcFramesAtoms(p1[None], p2[None] if p2 is not p1 else None)

    def calcFramesAtoms(self, pos_sel1, pos_sel2=None, boxes=None):
        """Many frames of atom positions: float32 (nframes,natoms,3); boxes (nframes,3) for NpT or None."""
        if getattr(self, "_natoms", None) is None:
            raise RuntimeError("attach_topology() must be called before atom positions can be passed")
        cdef np.ndarray[np.float32_t, ndim=3, mode="c"] a1 = np.ascontiguousarray(pos_sel1, dtype=np.float32)
        cdef np.ndarray[np.float32_t, ndim=3, mode="c"] a2 = a1 if (pos_sel2 is None or pos_sel2 is pos_sel1) else np.ascontiguousarray(pos_sel2, dtype=np.float32)
        cdef np.ndarray[np.float64_t, ndim=2, mode="c"] bx = None
        cdef int nf = a1.shape[0]
        na1, na2 = self._natoms
        if a1.shape[1] != na1 or a1.shape[2] != 3 or a2.shape[0] != nf or a2.shape[1] != na2 or a2.shape[2] != 3:
            raise ValueError("position arrays must have shapes (nframes,%d,3) and (nframes,%d,3)" % (na1, na2))
        if nf == 0:
            return
        if self.oneistwo is None:                            # :417-421 compares the first sites of the two selections
            k, same_topology = self._first_res
            self.oneistwo = bool(same_topology and (a2 is a1 or (a1[0, :k] == a2[0, :k]).all()) and self.nmol1 == self.nmol2)
        if boxes is not None:
            b3 = np.asarray(boxes, dtype=np.float64).reshape(nf, 3)
            bx = np.ascontiguousarray(np.concatenate([b3, b3 / 2.0], axis=1))
            vols = list(b3[:, 0] * b3[:, 1] * b3[:, 2])
        else:
            vols = [self.volume] * nf
        if self._auto_filter and self.ctr == 0:
            # a centre of mass lies inside the hull of its atoms, so max|position| bounds max|site coordinate|
            self._auto_pick_filter(None, None, max(float(np.abs(a1[0]).max()), float(np.abs(a2[0]).max())))
        self.ctr += nf
        self.volume_list.extend(vols)
        cdef _Backend be = self._be
        be.enqueue_atoms(nf, <const float *> a1.data, 3 * <long long> a1.shape[1],
                         <const float *> a2.data if a2 is not a1 else NULL, 3 * <long long> a2.shape[1],
                         <const double *> bx.data if bx is not None else NULL)
        if boxes is not None:
            self.update_box(*[float(v) for v in b3[nf - 1]])

    def calcTrajectory(self, traj, sel1, sel2=None, first=0, last=None, step=1, use_cell=False, threads=2):
        """Frames first, first+step, ... (< last) of a `DCDFile` straight into the histogram: replaces the script loop
        `for ts in u.trajectory[first:last:step]: ...positions... comByResidue ... dipByResidue ... calcFrame`
        (docs/notebooks/rdf.ipynb cell 18).  sel1 / sel2: atom indices into the trajectory (e.g. AtomGroup.indices) in
        the order of the attached topology; sel2=None is the self g-function.  Reader threads decode ahead of the GPU.
        use_cell: False - the box of the constructor / update_box for every frame; True - each frame's own A, B, C;
        "float32" - A, B, C rounded through float32 (what MDAnalysis' ts.dimensions gives a script).
        Returns the number of frames processed."""
        if getattr(self, "_natoms", None) is None:
            raise RuntimeError("attach_topology() must be called before atom positions can be passed")
        cdef long long nfr = traj.n_frames
        if last is None or last > nfr:
            last = nfr
        first = int(first); step = int(step)
        if step < 1 or first < 0:
            raise ValueError("first must be >= 0 and step >= 1")
        cdef long long count = (last - first + step - 1) // step if last > first else 0
        if count == 0:
            return 0
        s1 = np.ascontiguousarray(sel1, dtype=np.int32)
        s2 = None if (sel2 is None or sel2 is sel1) else np.ascontiguousarray(sel2, dtype=np.int32)
        cdef int uc = 2 if use_cell == "float32" else (1 if use_cell else 0)
        if self.oneistwo is None:
            k, same_topology = self._first_res
            self.oneistwo = bool(same_topology and (s2 is None or np.array_equal(s1[:k], s2[:k])) and self.nmol1 == self.nmol2)
        if self._auto_filter and self.ctr == 0:
            p1, _ = traj.read(first, s1)
            amax = float(np.abs(p1).max())
            if s2 is not None:
                amax = max(amax, float(np.abs(traj.read(first, s2)[0]).max()))
            self._auto_pick_filter(None, None, amax)
        boxes = self._be.enqueue_trajectory(traj, first, count, step, s1, s2, uc, int(threads))
        self.ctr += count
        if uc:
            self.volume_list.extend(list(boxes[:, 0] * boxes[:, 1] * boxes[:, 2]))
            self.update_box(*[float(v) for v in boxes[count - 1]])
        else:
            self.volume_list.extend([self.volume] * count)
        return int(count)

    def flush(self):
        self._be.flush()

    def sync(self):
        """Wait for all enqueued frames; raises the deferred ZeroDivisionError if a zero-norm dipole was hit."""
        self._be.sync()

    def stats(self):
        return self._be.stats()

    def mark(self, which):
        """Record device-side stopwatch mark 0 or 1 on the compute stream (benchmarks)."""
        self._be.mark(which)

    def mark_elapsed_ms(self):
        return self._be.mark_elapsed_ms()

    def attach_comm(self, uid, nranks, rank):
        """One process per GPU: join an NCCL communicator; readout then sums the histograms of all ranks.
        (A unique id makes exactly one communicator, so while the pre-filter choice is still open - "auto", no
        frame seen yet - joining is postponed until the backend is final.)"""
        self._nranks = int(nranks)
        if self._auto_filter and self.ctr == 0:
            self._pending_comm = (uid, nranks, rank)
        else:
            self._be.attach_comm(uid, nranks, rank)

    def _finish_comm(self):
        pc = getattr(self, "_pending_comm", None)
        if pc is not None:
            self._pending_comm = None
            self._be.attach_comm(*pc)


class RDF_voronoi(RDF):
    """
    class RDF_voronoi

    Shell-resolved g-functions (B200 backend): drop-in for the reference's RDF_voronoi (gfunction.pyx:764-1148).
    Every in-range pair is filed under the Delaunay shell of its two molecules, read from the int32 matrix passed to
    calcFrame; histogram_out has the reference layout [mode][bin][shell].
    """

    def __init__(self, modes, histo_min, histo_max, histo_dr, sel1_n, sel2_n, sel1_nmol, sel2_nmol, box_x, nshells, box_y=None, box_z=None, use_cuda=False, norm_volume=None, exclude_self_mol=False, gfn_flags=None, devices=None):
        """
        Same arguments as the reference (gfunction.pyx:771).  use_cuda is accepted and ignored; gfn_flags / devices
        are optional extensions.  The device handle is created by the first calcFrame, when the size of the
        Delaunay matrix is known.
        """
        self.nshells = np.int32(nshells)                     # :842
        self.exclude_self_mol = exclude_self_mol             # :843
        self._map_shape = None
        RDF.__init__(self, modes, histo_min, histo_max, histo_dr, sel1_n, sel2_n, sel1_nmol, sel2_nmol, box_x,
                     box_y, box_z, use_cuda, norm_volume, gfn_flags, devices)
        if int(self.nshells) < 1:
            raise ValueError("nshells must be >= 1")
        self.histogram_out = np.zeros(self.nshells * self.histo_n * 7, dtype=np.float64)    # :882

    def _open_backend(self):
        if self._map_shape is None:          # first frame not seen yet
            self._be = None
            return
        cdef int n1 = self.n1, n2 = self.n2
        if int(self.nmol1) < 1 or int(self.nmol2) < 1 or n1 // int(self.nmol1) < 1 or n2 // int(self.nmol2) < 1:
            raise ValueError("sel*_nmol must be between 1 and the number of sites (apr = n // nmol, gfunction.pyx:582-583)")
        self._be = _Backend()
        self._be.open_shells(n1, n2, int(self.histo_n), self.histo_min, self.histo_max, float(self.histo_inv_dr),
                             int(self.mode_sel), self.box_dim, self._devices, self._flags,
                             int(self.nshells), 1, n1 // int(self.nmol1), n2 // int(self.nmol2), n2,
                             <long long> self._map_shape[0] * self._map_shape[1], bool(self.exclude_self_mol))

    def update_box(self, box_x, box_y, box_z):
        """
        update_box(box_x,box_y,box_z)

        Update the box dimensions (:916-935).  Like the reference, self.volume is NOT refreshed here (the
        assignment sits inside a string literal in the reference).
        """
        self.box_dim[0] = box_x
        self.box_dim[3] = box_x / 2.0
        self.box_dim[1] = box_y
        self.box_dim[4] = box_y / 2.0
        self.box_dim[2] = box_z
        self.box_dim[5] = box_z / 2.0
        if self._be is not None:
            self._be.update_box(self.box_dim)

    def calcFrame(self,
                  np.ndarray[np.float64_t, ndim=2, mode="c"] coor_sel1,
                  np.ndarray[np.float64_t, ndim=2, mode="c"] coor_sel2,
                  np.ndarray[np.int32_t, ndim=2, mode="c"] delaunay_matrix,
                  np.ndarray[np.float64_t, ndim=2, mode="c"] dip_sel1 = None,
                  np.ndarray[np.float64_t, ndim=2, mode="c"] dip_sel2 = None,
                  pre_coor_sel1_gpu=None, pre_coor_sel2_gpu=None,
                  pre_dip_sel1_gpu=None, pre_dip_sel2_gpu=None):
        """
        calcFrame(coor_sel1, coor_sel2, delaunay_matrix, dip_sel1=None, dip_sel2=None, ...)

        Pass the data of the current frame (:937-1021).  delaunay_matrix[a, b] is the Delaunay distance between
        molecule a of the first and molecule b of the second selection; the reference indexes it as
        flat[(i // apr_core) * sel2_n + j // apr_surround] (:602), so it must hold at least that many elements
        (ValueError otherwise; the reference would read out of bounds).  pre_*_gpu are ignored like in the reference.
        """
        cdef int m = <int> self.mode_sel
        if dip_sel1 is None and (m & 2 or m & 4 or m & 16 or m & 32):
            raise TypeError("No dipole moment array for the first selection was passed, although the requested gfunctions need it!")
        if dip_sel2 is None and (m & 2 or m & 8 or m & 16 or m & 64):
            raise TypeError("No dipole moment array for the second selection was passed, although the requested gfunctions need it!")
        if coor_sel1.shape[0] != self.n1 or coor_sel1.shape[1] != 3 or coor_sel2.shape[0] != self.n2 or coor_sel2.shape[1] != 3:
            raise ValueError("coordinate arrays must have shapes (%d,3) and (%d,3)" % (self.n1, self.n2))
        if dip_sel1 is not None and (dip_sel1.shape[0] != self.n1 or dip_sel1.shape[1] != 3):
            raise ValueError("dip_sel1 must have shape (%d,3)" % self.n1)
        if dip_sel2 is not None and (dip_sel2.shape[0] != self.n2 or dip_sel2.shape[1] != 3):
            raise ValueError("dip_sel2 must have shape (%d,3)" % self.n2)

        if self.oneistwo is None:                            # :972-976
            if (coor_sel1[0] == coor_sel2[0]).all() and self.nmol1 == self.nmol2:
                self.oneistwo = True
            else:
                self.oneistwo = False

        shape = (delaunay_matrix.shape[0], delaunay_matrix.shape[1])
        if self._map_shape is None:
            self._map_shape = shape
            if self._auto_filter:
                self._auto_filter = False
                lmax = float(max(self.box_dim[0], self.box_dim[1], self.box_dim[2]))
                absmax = max(float(np.abs(coor_sel1).max()), float(np.abs(coor_sel2).max()))
                if not (self._flags & (2 | 16)) and absmax <= 2.0 * lmax and float(self.histo_max) >= 0.04 * lmax:
                    self._flags |= 16
            self._open_backend()
            self._finish_comm()
        elif shape != self._map_shape:
            raise ValueError("delaunay_matrix changed shape from %r to %r" % (self._map_shape, shape))

        self.ctr += 1                                        # :979
        self.volume_list.append(self.volume)                 # :982
        cdef _Backend be = self._be
        be.enqueue_shells(<const double *> coor_sel1.data, <const double *> coor_sel2.data,
                          <const double *> dip_sel1.data if dip_sel1 is not None else NULL,
                          <const double *> dip_sel2.data if dip_sel2 is not None else NULL,
                          <const int *> delaunay_matrix.data)

    def _addHisto(self, np.ndarray[np.float64_t, ndim=1] histo_out, np.ndarray[np.float64_t, ndim=1] histo):
        """
        Add the histograms of all core particles (:1023-1047).  `histo` may be the per-particle array
        (7*n1*histo_n*nshells) or the already summed [mode][bin][shell] array this backend keeps.
        """
        cdef long long row = <long long> self.histo_n * self.nshells
        cdef long long n1 = self.n1
        if histo.shape[0] == 7 * row:
            histo_out += histo
        elif histo.shape[0] == 7 * n1 * row:
            histo_out += histo.reshape(7, n1, row).sum(axis=1).reshape(-1)
        else:
            raise ValueError("histogram has %d elements, expected %d or %d" % (histo.shape[0], 7 * row, 7 * n1 * row))

    def _normHisto(self):
        """
        Normalize the histogram (:1049-1089); host code, same arithmetic as the reference.
        """
        ctr = self.ctr
        if getattr(self, "_nranks", 1) > 1 and self._be is not None:
            # frames are spread over processes (attach_comm): frame count and mean volume over all ranks, like RDF._normHisto
            tot = self._be.allreduce_host([float(self.ctr), float(sum(self.volume_list))])
            ctr = int(tot[0])
            volume = self.norm_volume if self.norm_volume else tot[1] / tot[0]
        elif self.norm_volume:
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

        cdef int i, j, shell
        for shell in range(self.nshells):
            for i in range(self.histo_n):
                r = self.histo_min + float(i) * self.histo_dr
                r_out = r + self.histo_dr
                slice_vol = 4.0 / 3.0 * PI * (r_out**3 - r**3)
                norm = rho * slice_vol * float(ctr)
                for j in range(7):
                    self.histogram_out[j * self.histo_n * self.nshells + i * self.nshells + shell] /= norm

    def scale(self, value):
        """:1091-1093"""
        self.histogram_out[:] *= value
        if self._be is not None:
            self._be.scale(float(value))

    def write(self, filename="rdf"):
        """
        RDF_voronoi.write(filename="rdf")

        Norm the calculated histograms and write them to <filename>_shell<n>_g000.dat ... (:1095-1133).
        """
        self._addHisto(self.histogram_out, self.raw_histogram())
        self._normHisto()

        names = ("000", "110", "101", "011", "220", "202", "022")
        for shell in range(self.nshells):
            files = {}
            for k in range(7):
                if self.mode_sel & (1 << k):
                    files[k] = open(filename + "_shell" + str(shell + 1) + "_g" + names[k] + ".dat", 'w')
            for i in range(self.histo_n):
                r = self.histo_min + i * self.histo_dr + self.histo_dr * 0.5
                for k in files:
                    files[k].write("%5.5f\t%5.5f\n" % (r, self.histogram_out[k * self.histo_n * self.nshells + i * self.nshells + shell]))
            for k in files:
                files[k].close()

    def resetArrays(self):
        """
        Reset histogram arrays to zero (:1135-1140)
        """
        self.histogram_out[:] = 0.0
        if self._be is not None:
            self._be.reset()

    def raw_histogram(self):
        """Un-normalised [mode][bin][shell] sums accumulated since the last resetArrays (7*histo_n*nshells float64)."""
        if not self._ensure_backend():
            return np.zeros(7 * int(self.histo_n) * int(self.nshells), dtype=np.float64)
        hist, pw, ovf = self._be.readout()
        self.pair_weight = pw
        self.overflow = ovf
        return hist

    def calcFrames(self, coor_sel1, coor_sel2, delaunay_matrices, dip_sel1=None, dip_sel2=None, boxes=None):
        """Many frames per call: coordinates / dipoles (nframes, n, 3), delaunay_matrices (nframes, rows, cols) int32,
        boxes (nframes, 3) for NpT or None (the volume used by _normHisto stays the constructor's, like update_box)."""
        c1 = np.ascontiguousarray(coor_sel1, dtype=np.float64)
        c2 = c1 if coor_sel2 is coor_sel1 else np.ascontiguousarray(coor_sel2, dtype=np.float64)
        dm = np.ascontiguousarray(delaunay_matrices, dtype=np.int32)
        d1 = None if dip_sel1 is None else np.ascontiguousarray(dip_sel1, dtype=np.float64)
        d2 = None if dip_sel2 is None else (d1 if dip_sel2 is dip_sel1 else np.ascontiguousarray(dip_sel2, dtype=np.float64))
        if c1.ndim != 3 or c2.ndim != 3 or dm.ndim != 3 or c2.shape[0] != c1.shape[0] or dm.shape[0] != c1.shape[0]:
            raise ValueError("calcFrames takes arrays of shape (nframes, n, 3) and matrices of shape (nframes, rows, cols)")
        for f in range(c1.shape[0]):
            if boxes is not None:
                self.update_box(*[float(v) for v in np.asarray(boxes, dtype=np.float64).reshape(-1, 3)[f]])
            self.calcFrame(c1[f], c2[f], dm[f], None if d1 is None else d1[f], None if d2 is None else d2[f])

    def calcFramesDevice(self, *args, **kwargs):
        raise TypeError("RDF_voronoi takes host frames: the Delaunay matrix travels with every frame (calcFrame / calcFrames)")

    def attach_topology(self, *args, **kwargs):
        raise TypeError("RDF_voronoi takes host frames: the Delaunay matrix travels with every frame (calcFrame / calcFrames)")

    def calcFrameAtoms(self, *args, **kwargs):
        raise TypeError("RDF_voronoi takes host frames: the Delaunay matrix travels with every frame (calcFrame / calcFrames)")

    calcFramesAtoms = calcFrameAtoms
    calcTrajectory = calcFrameAtoms

    def attach_comm(self, uid, nranks, rank):
        """One process per GPU: join an NCCL communicator; readout then sums the shell-resolved histograms of all ranks
        (the device handle is created by the first calcFrame, so joining happens there or at the first readout)."""
        self._nranks = int(nranks)
        if self._be is None:
            self._pending_comm = (uid, nranks, rank)
        else:
            self._be.attach_comm(uid, nranks, rank)

    def _ensure_backend(self):
        """A rank that saw no frame still has to take part in the collective readout: open the handle with the smallest
        matrix the selections allow."""
        if self._be is None:
            if getattr(self, "_nranks", 1) <= 1:
                return False
            self._auto_filter = False
            self._map_shape = (max(1, int(self.nmol1)), int(self.n2))
            self._open_backend()
        self._finish_comm()
        return True

    def sync(self):
        if self._be is not None:
            self._be.sync()

    def flush(self):
        if self._be is not None:
            self._be.flush()

    def stats(self):
        if self._be is None:
            raise RuntimeError("RDF_voronoi: the device handle is created by the first calcFrame")
        return self._be.stats()

    def mark(self, which):
        if self._be is None:
            raise RuntimeError("RDF_voronoi: the device handle is created by the first calcFrame")
        self._be.mark(which)

    def mark_elapsed_ms(self):
        if self._be is None:
            raise RuntimeError("RDF_voronoi: the device handle is created by the first calcFrame")
        return self._be.mark_elapsed_ms()

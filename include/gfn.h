/*
 * gfn.h - C ABI of libgfn.so, the B200 (sm_100a) backend of newanalysis.gfunction.RDF.
 *
 * This is the drop-in boundary for ONE reference path: the per-frame pair loop of the g-function
 * RDF class.  Every entry point cites the reference interface it replaces; all paths are relative
 * to /root/reference/newanalysis/gfunction/gfunction.pyx.
 *
 *   reference native call being replaced (the Cython cdef the RDF class calls once per frame):
 *     cdef void calcHisto(double *core_xyz, double *surr_xyz, double *core_dip, double *surr_dip,
 *                         ndarray dest, ndarray box_dim, int ncore, int nsurround,
 *                         float histo_min, float histo_max, int histo_n, double invdx, int mode_sel)
 *                                                                        gfunction.pyx:110-112
 *     call site                                                         gfunction.pyx:460-464
 *     legacy GPU analogue  CUDA.calcHistoGPU(...)                        gfunction.pyx:104-105, 455-457
 *
 * Conventions: plain C, no torch / numpy types.  Every function returns an int status (GFN_OK = 0)
 * unless stated; gfn_last_error() gives the message of the last failure on that handle.  All host
 * pointers are caller-owned; input frames are COPIED into the library's pinned ring before
 * gfn_enqueue_frame returns, so the caller may overwrite them immediately (analysis scripts reuse
 * their arrays every frame).  There is NO CPU fallback: without a usable sm_100 device gfn_create
 * fails with GFN_ENODEV.
 *
 * Arithmetic contract (DESIGN.md sections 3 and 6):
 *   - Distances and bins: IEEE fp64 with exactly the reference's operation order and one rounding per operation (no
 *     fused multiply-add), so every pair lands in the reference's bin and the g000 row / pair counts are bit-exact.
 *     g000-only handles in dense boxes (count kernel, gfn_count.cu) decide a pair's bin in fp32 when the estimate is farther
 *     than a proven error bound from every bin edge and from both ends of the accepted range, and in fp64 otherwise: the
 *     same bins, bit for bit.
 *   - Orientation weights (gfunction.pyx:154-165): by default each cosine is formed as  dot * (1/|a| * 1/|b|)  with
 *     separately rounded reciprocals (1/norm per site, 1/dist from the square-root iteration) instead of the
 *     reference's  dot / (|a|*|b|):  a per-pair increment can differ from the reference's by up to 4 ulp (relative
 *     4.5e-16), far inside the 1e-12 bar of the weighted rows; sums additionally differ by summation order.  A pair
 *     whose needed norm is zero / outside [2^-300, 2^300] always takes the literal divisions (quirk Q5 semantics).
 *     The dense-box kernel (gfn_dense.cu) accumulates the weighted rows in 64-bit fixed point, 2^-43 per unit: one
 *     addend is off by at most 5.7e-14 absolute (bar: 1e-12 per unit of pair weight), sums are exact integers and
 *     therefore identical for any order of evaluation.
 *   - GFN_FLAG_EXACT_DIV makes every pair take the literal IEEE divisions: per-pair increments are then bit-identical to
 *     the reference's and the weighted rows differ from it by summation order only.
 */
#ifndef GFN_H
#define GFN_H

#ifdef __cplusplus
extern "C" {
#endif

#define GFN_OK        0
#define GFN_EINVAL    1   /* bad argument                                   -> ValueError      */
#define GFN_ECUDA     2   /* CUDA runtime failure                           -> RuntimeError    */
#define GFN_ENCCL     3   /* NCCL failure / NCCL library not loadable       -> RuntimeError    */
#define GFN_EZERODIV  4   /* in-range pair with a zero-norm dipole (Q5)     -> ZeroDivisionError("float division"),
                             gfunction.pyx:155,159,163 under cdivision=False; reported at the next
                             gfn_sync / gfn_readout after the offending frame (deferred) */
#define GFN_ENOMEM    5
#define GFN_ENODEV    6   /* no usable CUDA device: the product never falls back to the CPU */

/* gfn_create flags */
#define GFN_FLAG_PER_PARTICLE      1u  /* keep the reference's [mode][core i][bin] histogram on the device
                                          (7*n1*histo_n doubles, gfunction.pyx:332) instead of [mode][bin] */
#define GFN_FLAG_FILTER_FP64       2u  /* range pre-test in fp64 instead of the conservative fp32 pre-filter;
                                          results are identical, only the pipe that rejects far pairs changes */
#define GFN_FLAG_DISCARD_OVERFLOW  4u  /* pos >= histo_n (quirk Q3) is dropped instead of spilling into
                                          bin pos-histo_n of the next row like the reference does */
#define GFN_FLAG_HIST_GLOBAL       8u  /* force the global-atomic histogram path (default: per-warp shared-memory
                                          replicas whenever they fit) */
#define GFN_FLAG_FILTER_FP16       16u /* range pre-filter in packed fp16 (two core sites per instruction, coordinates in box
                                          units, wider guard band); results identical, see DESIGN.md section 4 */
#define GFN_FLAG_NO_TRIANGLE       32u /* evaluate all n1 x n2 pairs with weight 1 even when n1 == n2 (calcHistoTS,
                                          gfunction.pyx:168-192, has no j >= i shortcut) */
#define GFN_FLAG_DENSE             128u /* force the dense-box kernels (g000 only: count kernel, fp32 bin classification + fp64 for
                                          the ambiguous pairs; other mode sets: every warp filters AND evaluates, core sites in
                                          registers) where they support the handle; default: chosen when the share of in-range
                                          pairs expected from box and histo_max exceeds 10 % */
#define GFN_FLAG_NO_DENSE          256u /* never use them (the warp-specialised ring kernel serves every handle) */
#define GFN_FLAG_CELLS             512u /* tile skipping for sparse neighbourhoods: every frame is sorted into Morton order on the device
                                         * and only the tile pairs whose bounding boxes lie within histo_max of each other are
                                         * visited (gfn_cells.cu).  Same pairs accepted, same arithmetic: counts bit-exact.  Honoured
                                         * where the expected share of visited tile pairs is small enough to pay (a box many tile
                                         * diameters wide, e.g. 10^6 sites: 269 -> 20 ms per frame, g000 only 251 -> 7 ms); ignored by
                                         * shell-resolved handles.  Launches of two DIFFERENT selections of equal size keep the full
                                         * enumeration (quirk Q1: their triangle relates original indices). */
#define GFN_FLAG_EXACT_DIV         64u /* orientation cosines by the reference's literal divisions dot/(|a|*|b|)
                                          (gfunction.pyx:155,159,163) instead of reciprocal products: per-pair increments
                                          bit-identical to the reference, see "Arithmetic contract" above */

/* bits of mode_sel, gfunction.pyx:274-288 */
#define GFN_MODE_000 1
#define GFN_MODE_110 2
#define GFN_MODE_101 4
#define GFN_MODE_011 8
#define GFN_MODE_220 16
#define GFN_MODE_202 32
#define GFN_MODE_022 64

typedef struct gfn_handle gfn_t;

typedef struct gfn_stats {
    double pair_evals;        /* reference j-loop iterations submitted so far (n(n+1)/2 or n1*n2 per frame) */
    double survivors;         /* pairs that passed the range pre-filter and were evaluated exactly in fp64 */
    double in_range;          /* pairs that were binned (passed gfunction.pyx:149) */
    double kernel_ms;         /* device time of all pair-kernel launches so far (CUDA events, summed over devices) */
    double kernel_launches;   /* number of launches of library kernels so far (all kinds) */
    double pair_launches;     /* number of pair-kernel launches */
    double frames;            /* frames submitted */
    double h2d_bytes;         /* bytes copied host->device */
    double overflow;          /* pairs with pos >= histo_n (quirk Q3) */
    int    filter_fp64;       /* range pre-filter in use: 0 fp32, 1 fp64 pre-test, 2 packed fp16 */
    int    hist_mode;         /* 0 shared-memory warp replicas, 1 global atomics, 2 per-particle global */
    int    warps_per_block, blocks_per_sm, grid, smem_bytes, batch_frames, ndev, hist_replicas;
    int    dense;             /* 0 ring kernel, 1 dense-neighbourhood kernel, 2 g000 count kernel (fp32 classification) */
    double exact_evals;       /* count kernel: pairs the fp32 classification could not decide, evaluated in fp64 */
    double selfcheck_bad;     /* count kernel under GFN_DEBUG=8: pairs whose fp32 decision differs from the exact one (must be 0) */
    double tile_visits;       /* GFN_FLAG_CELLS: (core tile, surround tile) pairs the pair kernel visited so far ... */
    double tile_visits_dense; /* ... and how many the un-pruned enumeration has for the same frames */
    int    cells;             /* 1 if tile skipping is active */
    double frames_max_device; /* frames submitted to the device that got the most / the fewest of them (block calls deal */
    double frames_min_device; /* ... every device the same number +- 1) */
    double kernel_ms_max_device;  /* pair-kernel time of the busiest device */
} gfn_stats_t;

/* Number of CUDA devices visible (0 if none / no driver).  Never fails. */
int gfn_device_count(void);

/* Replaces the device-side half of RDF.__init__ (gfunction.pyx:234-364; legacy allocations :336-364).
 *   n1,n2        selection sizes (self.n1, self.n2)                       :321-322
 *   histo_n      number of bins  int32((max-min)/dr+0.5)                  :295
 *   histo_min/max  passed as C float exactly like calcHisto takes them    :112, :291-292 (quirk Q4)
 *   inv_dr       1/round(dr,5)                                            :293-294
 *   mode_sel     bitmask                                                  :274-288
 *   box_dim[6]   lengths then half lengths                                :297-311 (halves must equal length/2)
 *   devices/ndev CUDA ordinals to shard frames over (NULL/0 = device 0)
 */
int gfn_create(gfn_t **out, int n1, int n2, int histo_n, float histo_min, float histo_max,
               double inv_dr, int mode_sel, const double box_dim[6],
               const int *devices, int ndev, unsigned flags);

/* RDF.update_box, gfunction.pyx:367-383: applies to frames enqueued afterwards (the box travels
 * with each frame, so NpT trajectories are safe with asynchronous batches). */
int gfn_update_box(gfn_t *h, const double box_dim[6]);

/* One frame = one call of calcHisto (gfunction.pyx:460-464).  c1 (n1,3), c2 (n2,3), d1 (n1,3), d2 (n2,3)
 * C-contiguous float64; d1/d2 may be NULL when mode_sel does not need them (:411-415, quirk Q6).
 * c2 == c1 (same pointer) uploads the selection once.  Asynchronous with respect to the GPU. */
int gfn_enqueue_frame(gfn_t *h, const double *c1, const double *c2, const double *d1, const double *d2);

/* Batch extension: nframes frames, frame f of every array starts stride_k doubles after frame f-1
 * (stride 0 is not allowed).  box_dims: NULL = current box for all, else 6 doubles per frame.
 * On a handle with several devices a call of at least two frames per device gives every device the same number of
 * frames (+- 1): the call lasts as long as its fullest device (gfn_stats: frames_max_device / frames_min_device). */
int gfn_enqueue_frames(gfn_t *h, int nframes,
                       const double *c1, long long stride_c1, const double *c2, long long stride_c2,
                       const double *d1, long long stride_d1, const double *d2, long long stride_d2,
                       const double *box_dims);

/* The same frames WITHOUT the staging copy: every array must be page-locked host memory (gfn_host_alloc below, or
 * cudaHostRegister'ed by the caller; anything else is GFN_EINVAL) and must stay unchanged until the next gfn_sync /
 * gfn_readout returns - the DMA engine reads the caller's arrays directly (one strided copy per array and batch), so the
 * host does no per-frame work at all.  Meant for trajectory decoders that write into page-locked buffers and for one
 * process feeding several devices.  (The reference has no counterpart: calcFrame, gfunction.pyx:385-464, is synchronous.) */
int gfn_enqueue_frames_pinned(gfn_t *h, int nframes,
                              const double *c1, long long stride_c1, const double *c2, long long stride_c2,
                              const double *d1, long long stride_d1, const double *d2, long long stride_d2,
                              const double *box_dims);

/* Page-locked host memory for gfn_enqueue_frames_pinned (cudaHostAlloc, portable across devices). */
int gfn_host_alloc(void **out, unsigned long long bytes);
int gfn_host_free(void *p);

/* Device-resident frames (inputs already in HBM on one of h's devices; the legacy pre_*_gpu kwargs of
 * calcFrame, gfunction.pyx:390-391,431-453, map here).  Pointers are DEVICE pointers with the same
 * layout as gfn_enqueue_frames, all on the same device; the frames are processed on the device that
 * holds them (a multi-device handle takes frames on each of its GPUs).  Nothing is copied from the host
 * except the boxes. */
int gfn_enqueue_device_frames(gfn_t *h, int nframes,
                              const double *dc1, long long stride_c1, const double *dc2, long long stride_c2,
                              const double *dd1, long long stride_d1, const double *dd2, long long stride_d2,
                              const double *box_dims);

/* ---- N2: shell-resolved g-functions -------------------------------------------------
 * Replaces cdef void calcHistoVoronoi / calcHistoVoronoiNonSelf (gfunction.pyx:570-664, :667-761; call site
 * RDF_voronoi.calcFrame :1009-1021) and def calcHistoTSVoro (:195-224).  The pair loop is the one of calcHisto;
 * each in-range pair additionally reads its shell from an int32 matrix that travels with the frame, and the
 * histogram rows hold histo_n * nshells bins.
 *   layout 1 (RDF_voronoi):     bin = pos*nshells + shell, shell = map[(i/apr_core)*map_ld + j/apr_surround] - 1,
 *                               values outside [0, nshells) go to the last shell (:602-605); the reference passes
 *                               map_ld = nsurround (:602, "Correctly shaped?")
 *   layout 2 (calcHistoTSVoro): bin = shell*histo_n + pos, shell = min(map[i*map_ld + j], nshells) - 1, a negative
 *                               index wraps once like Cython's wraparound indexing (:219-224)
 * map_elems is the number of ints of one frame's matrix; it must cover the largest index the loop can form,
 * (ceil(n1/apr_core)-1)*map_ld + ceil(n2/apr_surround) (the reference would read out of bounds otherwise).
 * Handles created here take frames only through gfn_enqueue_frames_shells; readout sizes use
 * histo_n*nshells bins per row.  exclude_self != 0 skips pairs with i/apr_core == j/apr_surround (:700). */
typedef struct gfn_shell_cfg {
    int nshells, layout;
    int apr_core, apr_surround;
    int map_ld;
    long long map_elems;
    int exclude_self;
} gfn_shell_cfg;
int gfn_create_shells(gfn_t **out, int n1, int n2, int histo_n, float histo_min, float histo_max,
                      double inv_dr, int mode_sel, const double box_dim[6], const int *devices, int ndev,
                      unsigned flags, const gfn_shell_cfg *cfg);
int gfn_enqueue_frames_shells(gfn_t *h, int nframes,
                              const double *c1, long long s_c1, const double *c2, long long s_c2,
                              const double *d1, long long s_d1, const double *d2, long long s_d2,
                              const double *box_dims, const int *shell_map, long long s_map);

/* Submit a partially filled batch (asynchronous). */
int gfn_flush(gfn_t *h);
/* Flush and wait for every submitted frame; reports deferred GFN_EZERODIV. */
int gfn_sync(gfn_t *h);

/* Replaces memcpy_dtoh + RDF._addHisto (gfunction.pyx:529-531, 466-480): flush, wait, combine the
 * per-device histograms (one ncclAllReduce when more than one device/rank participates) and copy out.
 *   hist      7*histo_n doubles, layout [mode k][bin], k in order 000,110,101,011,220,202,022: the
 *             un-normalised sum over core particles of everything accumulated since the last gfn_reset
 *             (OVERWRITES hist; the caller adds it to histogram_out like _addHisto does)
 *   pairw     histo_n doubles or NULL: total pair weight per bin (what g000 would hold) regardless of mode_sel
 *   overflow  1 value or NULL: number of pairs whose pos >= histo_n (quirk Q3)
 * Everything is enqueued behind the last frame on the compute stream (status words, the all-reduce, the copy into
 * pinned memory) and the host waits once.  Deferred errors (GFN_EZERODIV, the fp16 range GFN_EINVAL) are summed over
 * the communicator together with the histogram: every rank returns the error if any rank hit it.
 */
int gfn_readout(gfn_t *h, double *hist, double *pairw, unsigned long long *overflow);

/* Per-particle histogram (GFN_FLAG_PER_PARTICLE only): 7*n1*histo_n doubles, reference layout :332. */
int gfn_readout_per_particle(gfn_t *h, double *hist);

/* A slice of the same array added to the caller's: dst[k] += per_particle[first + k], k < count.  This is the
 * reference's `result[t,i,pos] += 1.0` of calcHistoTS (gfunction.pyx:168-192) for a whole frame: the shim passes
 * &result[t,0,0] and only the g000 rows of the frame cross PCIe (through a pinned staging buffer the handle keeps). */
int gfn_readout_per_particle_add(gfn_t *h, double *dst, unsigned long long first, unsigned long long count);

/* RDF.scale (gfunction.pyx:513-515) and RDF.resetArrays (:552-557) on the device accumulators. */
int gfn_scale(gfn_t *h, double value);
int gfn_reset(gfn_t *h);

void gfn_destroy(gfn_t *h);
const char *gfn_last_error(const gfn_t *h);   /* h may be NULL: last gfn_create failure */
int gfn_stats(gfn_t *h, gfn_stats_t *out);

/* Multi-process frame sharding (one process per GPU): rank 0 calls gfn_comm_unique_id, ships the
 * 128 bytes to the other ranks by any means, every rank calls gfn_comm_attach; gfn_readout then
 * all-reduces over the ranks (ncclAllReduce, sum, fp64).  Single-process multi-device handles set
 * up their communicator internally (ncclCommInitAll). */
int gfn_comm_unique_id(char out[128]);
int gfn_comm_attach(gfn_t *h, const char unique_id[128], int nranks, int rank);
/* Sum of n host doubles over the ranks of the attached communicator, in place (collective; no-op without
 * gfn_comm_attach).  For what RDF._normHisto needs besides the histogram when frames are spread over processes: the
 * number of frames (self.ctr, gfunction.pyx:424) and the sum of their volumes (self.volume_list, :427, :487-492). */
int gfn_allreduce_host(gfn_t *h, double *vals, int n);

/* Device-side stopwatch for benchmarks: gfn_mark records a CUDA event on the compute stream of the
 * handle's first device (after flushing pending frames), gfn_mark_elapsed_ms waits for mark 1 and
 * returns the device time between mark 0 and mark 1. */
int gfn_mark(gfn_t *h, int which);
int gfn_mark_elapsed_ms(gfn_t *h, double *ms);

/* Measured pipe peaks for the roofline denominator: dependent-free DFMA / FFMA loops on every SM
 * of device dev.  Results in TFLOP/s (an FMA counts 2). */
int gfn_measure_peaks(int dev, double *fp64_tflops, double *fp32_tflops);

/* ---- N1: per-residue centre of mass and dipole on the device --------------------------------
 * Replaces helpers.comByResidue (/root/reference/newanalysis/helpers/helpers.pyx:71-98), dipByResidue (:101-123)
 * and, being the same arithmetic on velocities, velcomByResidue (:41-68).  Results are bit-identical to the
 * reference (same atom order, un-fused multiply/add, IEEE division).
 *   apr[i]  atoms of residue i, rfa[i] index of its first atom (the reference's atoms_per_residue / residue_first_atom)
 *   coor    nframes x natoms x 3 float64;  outputs nframes x nres x 3 float64 */
typedef struct gfn_residues gfn_residues_t;
int gfn_residues_create(gfn_residues_t **out, int dev, int natoms, int nres, const int *apr, const int *rfa,
                        const double *masses, const double *charges /* NULL: centres of mass only */);
void gfn_residues_destroy(gfn_residues_t *r);
/* host arrays: com_in == NULL computes the centres of mass (returned in com_out if not NULL); dip_out != NULL
 * computes dipoles relative to com_in (or to the computed centres of mass) */
int gfn_residues_com_dip(gfn_residues_t *r, int nframes, const double *coor, const double *com_in,
                         double *com_out, double *dip_out);
/* device arrays in and out (feed them to gfn_enqueue_device_frames); ddip may be NULL */
int gfn_residues_com_dip_device(gfn_residues_t *r, int nframes, const double *dcoor, double *dcom, double *ddip);

/* ---- N1 + N4: frames as raw float32 atom positions ------------------------------------------
 * Replaces, per frame, the script-side chain  positions.astype(float64) -> helpers.comByResidue (helpers.pyx:71-98)
 * -> helpers.dipByResidue (:101-123) -> RDF.calcFrame (gfunction.pyx:385-464): the ring carries the trajectory's
 * native float32 positions (12 B/atom), centres of mass and dipoles are produced on the device, bit-identical to
 * the reference chain (float32 widens exactly), and feed the pair kernel without touching the host again.
 *   gfn_attach_topology: which = 0 / 1 selects selection 1 / 2; nres must equal n1 / n2.  Without a topology for
 *     selection 2 (n1 == n2) the one of selection 1 serves both.  charges may be NULL when mode_sel does not need
 *     that selection's dipoles.  The arrays are copied (to every device of the handle).
 *   gfn_enqueue_atom_frames: a1 (nframes, natoms1, 3), a2 (nframes, natoms2, 3) float32, strides in floats;
 *     a2 == NULL or a2 == a1: the self g-function (positions uploaded once).  Like gfn_enqueue_frames the arrays
 *     are copied before the call returns; box_dims as there. */
int gfn_attach_topology(gfn_t *h, int which, int natoms, int nres, const int *apr, const int *rfa,
                        const double *masses, const double *charges);
int gfn_enqueue_atom_frames(gfn_t *h, int nframes, const float *a1, long long stride_a1,
                            const float *a2, long long stride_a2, const double *box_dims);
/* atoms of the attached topologies (natoms2 = natoms1 when selection 2 shares the topology; 0, 0 before attaching) */
int gfn_topology_natoms(gfn_t *h, int *natoms1, int *natoms2);

/* ---- N4: trajectory decode overlapped with compute ------------------------------------------
 * Replaces the script-side frame loop `for ts in u.trajectory[first:last:step]: ... sel.positions ...`
 * (/root/reference/docs/notebooks/rdf.ipynb cells 10-18; MDAnalysis' DCD reader, a third-party dependency of the
 * reference) for CHARMM / NAMD DCD files: both byte orders, optional unit-cell and fourth-dimension records; files
 * with fixed atoms or 64-bit record markers are refused.  Host code, usable without a GPU.
 *   gfn_traj_read: positions of the atoms idx[0..n) (idx == NULL: all atoms) of one frame as (n,3) float32 exactly as
 *     stored; cell = A, B, C, alpha, beta, gamma (degrees; zeros when the file has no cell).  xyz / cell may be NULL.
 *   gfn_enqueue_trajectory: frames first, first+step, ... (count of them) go through gfn_enqueue_atom_frames in
 *     order; nthreads reader threads read, de-interleave and gather ahead while the calling thread fills the pinned
 *     ring, so file decoding overlaps the uploads and kernels of earlier frames.  sel1 / sel2 are atom indices into
 *     the trajectory (sel2 == NULL: self g-function) matching the attached topologies.  use_cell 0: the handle's
 *     current box (gfn_update_box) for every frame; 1: each frame's own A, B, C as stored; 2: A, B, C rounded through
 *     float32, what MDAnalysis' ts.dimensions hands to a script.  boxes_out: count x 3 lengths used (use_cell != 0), or NULL.
 * Errors of these functions are reported by gfn_traj_last_error (t may be NULL: last gfn_traj_open_dcd failure). */
typedef struct gfn_traj gfn_traj_t;
int gfn_traj_open_dcd(gfn_traj_t **out, const char *path);
void gfn_traj_close(gfn_traj_t *t);
int gfn_traj_info(const gfn_traj_t *t, int *natoms, long long *nframes, int *has_cell);
int gfn_traj_read(gfn_traj_t *t, long long frame, const int *idx, int n, float *xyz, double cell[6]);
int gfn_enqueue_trajectory(gfn_t *h, gfn_traj_t *t, long long first, long long count, long long step,
                           const int *sel1, int nsel1, const int *sel2, int nsel2, int use_cell,
                           double *boxes_out, int nthreads);
const char *gfn_traj_last_error(const gfn_traj_t *t);

/* Self-test of the kernel's straight-line fp64 division / square-root sequences against the IEEE
 * built-ins (__ddiv_rn, __dsqrt_rn) on n_operands random operand sets with exponents in the range the
 * kernel uses them for.  mismatches = {division, square root, near-tie cases}; all must be 0. */
int gfn_selftest_arith(int dev, unsigned long long seed, long long n_operands, unsigned long long mismatches[3]);

/* Raw device memory helpers so a host framework without its own CUDA binding can stage
 * device-resident frames for gfn_enqueue_device_frames (bench: inputs resident in HBM). */
int gfn_dev_alloc(int dev, void **dptr, unsigned long long bytes);
int gfn_dev_free(int dev, void *dptr);
int gfn_dev_upload(int dev, void *dptr, const void *host, unsigned long long bytes);

#ifdef __cplusplus
}
#endif
#endif /* GFN_H */

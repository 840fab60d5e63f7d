// gfn_kernels.cuh - device-side data layout and kernel launch interface of libgfn.so (sm_100a only).
// See DESIGN.md for the layout rationale; reference citations are to
// /root/reference/newanalysis/gfunction/gfunction.pyx.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gfn {

// One site of a selection as the pair kernel consumes it: 80 bytes, so a tile of T consecutive
// sites is one contiguous 80*T-byte run that a single 1-D TMA bulk copy moves into shared memory.
// The 80-byte stride also spreads the 16-byte chunks of randomly indexed records over all eight
// 4-bank groups (80*i mod 128 takes 8 values), so survivor gathers are bank-conflict free.
//   x,y,z     coordinates exactly as passed to calcFrame (fp64)          gfunction.pyx:141-143
//   px,py,pz  dipole (0 when the modes do not need it)                   :122-124 / :136-138
//   norm      sqrt(px*px+py*py+pz*pz), un-fused like the reference       :125 / :139
//   fx,fy,fz  coordinates rounded to fp32 for the range pre-filter (1e30 for padding sites)
struct __align__(16) Site {
    double x, y, z, px, py, pz, norm, spare;
    float fx, fy, fz, fw;
};
static_assert(sizeof(Site) == 80, "Site must be 80 bytes");

constexpr int kTJ      = 128;    // surround sites per shared-memory stage
constexpr int kStages  = 3;      // surround stages in flight
constexpr int kIPT     = 2;      // core sites held in registers per thread
#ifdef GFN_EXP_12_6   // experiment: 12 filter + 6 drain warps (768-site core tiles)
constexpr int kPadTo   = 3072;   // a multiple of the 768-site core tile and of 1024
#else
constexpr int kPadTo   = 1024;   // selections are padded to a multiple of this many sites on the device
#endif
#ifndef GFN_DRAIN_ILP
#define GFN_DRAIN_ILP 2
#endif
constexpr int kDrainILP = GFN_DRAIN_ILP;     // survivors evaluated per drain lane per batch
#ifdef GFN_EXP_12_6
constexpr int kRingCap = 1024;   // 12 rings must fit next to 6 replicas (sparse tiles only: a fully dense mask push needs 2048)
#else
constexpr int kRingCap = 2048;   // survivor ring entries (16 bit) per filter warp: one push (<= 32*32) always fits behind < 64 unread ones
#endif
constexpr int kCtrlBytes = 640;  // shared-memory control block
constexpr int kCounters = 8;     // 0 overflow, 1 error flags (bit 0 zero-norm dipole, bit 1 fp16 filter range), 2 survivors, 3 in-range pairs,
                                 // 4 count kernel self-check mismatches (GFN_DEBUG=8), 5 count kernel: pairs evaluated in fp64,
                                 // 6 tile visits of cell-sorted launches (GFN_FLAG_CELLS)

struct PairParams {
    const Site* siteA;  const float* maxA;   // selection 1 (core)
    const Site* siteB;  const float* maxB;   // selection 2 (surround)
    const unsigned* nflagA; const unsigned* nflagB;   // [nframes] != 0: a dipole norm of the frame is zero or outside [2^-300, 2^300]
    const double* box;              // [nframes][6] lengths, half lengths
    long long strideA, strideB;     // padded sites per frame
    int n1, n2, nframes, tri;       // tri: n1 == n2 (reference keys the triangle on sizes, quirk Q1)
    int histo_n, mode_sel;
    double hmin, hmax, invdx;       // hmin/hmax = (double)(float)histo_min/max  (quirk Q4)
    int n_itiles, n_jtiles, tiles_per_frame;
    const int* tile_start;          // [n_itiles+1] prefix of surround-tile visits per core tile
    double* partials;               // [grid][histo_n*NSLOT] block partial histograms (shared-memory mode)
    double* gacc;                   // [8][histo_n] device accumulators: 7 modes + pair weight
    double* gpp;                    // [7][n1][histo_n] per-particle histogram or nullptr
    unsigned long long* counters;   // [kCounters]
    unsigned flags;                 // GFN_FLAG_DISCARD_OVERFLOW
    int nrep;                       // histogram replicas in shared memory
    unsigned poll_ns; int poll_reps; // drain-warp back-off while its ring holds less than a batch
    int debug;                      // GFN_DEBUG ablations (timing experiments only, results wrong): 1 drain discards, 2 filter does not publish
    // shell-resolved histograms (N2: RDF_voronoi gfunction.pyx:570-1148, calcHistoTSVoro :195-224); shell_mode 0 = off.
    // histo_n above counts ALL bins of a row (radial bins x shells); histo_nr is the number of radial bins.
    const int* shell_map;           // [nframes][...] int32 shell matrix of each frame (Delaunay distance)
    long long stride_map;           // ints between the matrices of consecutive frames
    int shell_mode;                 // 1: bin = pos*nshells + shell, shell = map-1 clamped to the last (:602-605)
                                    // 2: bin = shell*histo_nr + pos, shell = min(map, nshells)-1, negative wraps (:219-224)
    int nshells, histo_nr;
    int apr1, apr2;                 // sites per molecule: matrix row = i / apr1, column = j / apr2 (:582-583, :601)
    int map_ld;                     // row stride of the matrix as the reference indexes it (:602 uses nsurround)
    int excl_self;                  // skip pairs of the same molecule (calcHistoVoronoiNonSelf :700)
    // cell-sorted frames with tile skipping (gfn_cells.cu; nullptr = every tile pair is visited): siteA / siteB are then the
    // SORTED records (Site.fw holds the site's original index), the core tile of entry e = frame * n_itiles + ti visits
    // the surround tiles visit_list[e * n_jtiles + 0 .. visit_start[e+1] - visit_start[e])
    const long long* visit_start;   // [nframes * n_itiles + 1] exclusive prefix, last = total visits of the launch
    const int* visit_list;
    long long visit_cap;            // dense kernel: entries the compact list buffer holds (a launch that lists more does nothing and flags it)
    // dense kernel, g000-only handles: the bin of a pair from d^2 without a square root (gfn_dense.cu)
    const double* bin_thr;          // [histo_n + 4]  T[k] = min { t : sqrt(t) >= hmin and floor((sqrt(t) - hmin) * invdx) >= k }
    double acc_lo, acc_hi;          // a pair passes gfunction.pyx:149  <=>  acc_lo <= d^2 <= acc_hi
    int pos_max;                    // bin of d^2 = acc_hi (histo_n when dist == hmax lands on the edge, quirk Q3)
};

struct PairConfig {
    int filter_fp64;   // 0 fp32 pre-filter, 1 fp64 pre-test, 2 fp16x2 pre-filter
    int hist_mode;     // 0 shared-memory warp replicas, 1 global atomics, 2 per-particle global atomics
    int warps;         // warps per CTA: filter warps (12 or 8) + drain warps
    int drain_warps;   // 4, or 8 working in pairs on 4 replicas
    int nrep;          // shared-memory histogram replicas per CTA (drain warps sharing one take a lock)
    int nslot;         // 1 (g000 only) or 8
    int all_modes;     // mode_sel == 127: the kernel variant without mode masks
    int shells;        // shell-resolved rows (gfn_create_shells): the kernel variant with the shell lookup
    int blocks_per_sm;
    int grid;
    int smem_bytes;
    int dense;         // 1: the symmetric dense-neighbourhood kernel (gfn_dense.cu) serves this handle, 2: the g000 count kernel (gfn_count.cu)
    int dense_msel;    // its compile-time mode set (1, 31, 127) or 0 = run-time modes
};

// Chooses kernel variant, grid and shared-memory size for a histogram of histo_n bins; sets the
// function attributes.  Returns cudaSuccess or the failing error.
//   in_range_hint: expected share of pairs inside histo_max, (4/3) pi hmax^3 / V of the handle's box: above kDenseShare
//   the dense kernel is chosen when it supports the handle (GFN_FLAG_DENSE / GFN_FLAG_NO_DENSE override the hint).
constexpr double kDenseShare = 0.10;
struct HandleGeom { double hmax, invdx, box_max; int n1, n2, tri; double cells_share; };   // cells_share: expected share of visited units under tile skipping (0: off)    // what the kernel choice looks at besides the in-range share
cudaError_t pair_configure(int dev, int n1, int histo_n, int mode_sel, unsigned flags, int shells, double in_range_hint, const HandleGeom& geom, PairConfig* cfg);
bool count_configure(int dev, int histo_n, int mode_sel, double invdx, double hmax, double box_max, int max_smem, int sms, PairConfig* cfg, cudaError_t* err);
cudaError_t count_launch(const PairConfig& cfg, const PairParams& p, cudaStream_t s);
bool dense_configure(int dev, int histo_n, int mode_sel, int max_smem, int sms, const HandleGeom& geom, PairConfig* cfg, cudaError_t* err);
long long dense_units_per_frame(int n1, int n2, int tri);
cudaError_t dense_launch(const PairConfig& cfg, const PairParams& p, cudaStream_t s);
long long dense_max_units(const PairConfig& cfg);     // launch-size bound of the dense kernel's fixed-point sums
cudaError_t pair_launch(const PairConfig& cfg, const PairParams& p, cudaStream_t s);

// Converts nframes frames of one selection into Site records and the per-frame max |coordinate|.
cudaError_t prep_launch(const double* xyz, long long stride_xyz, const double* dip, long long stride_dip,
                        Site* site, float* maxabs, unsigned* nflag, int n, int n_pad, int nframes, const double* box, int half_filter,
                        unsigned long long* counters, cudaStream_t s);

// partials[grid][histo_n*nslot] -> gacc, blocks summed in ascending order (deterministic).
cudaError_t finalize_launch(const double* partials, int grid, int histo_n, int nslot, double* gacc, cudaStream_t s);

cudaError_t scale_launch(double* a, long long n, double v, cudaStream_t s);

// status[0] = 1 if the fp16 range flag is set in counters[1], status[1] = 1 if a zero-norm dipole was hit (quirk Q5):
// the two words ride behind the accumulators through the readout's all-reduce.
cudaError_t status_launch(const unsigned long long* counters, double* status, cudaStream_t s);

// N1 producers (helpers.pyx:71-123): nframes frames of natoms x 3 coordinates -> nres x 3 centres of mass / dipoles.
cudaError_t residue_com_launch(const double* coor, long long stride_coor, const double* masses, const int* apr, const int* rfa,
                               int nres, int nframes, double* com, cudaStream_t s);
cudaError_t residue_dip_launch(const double* coor, long long stride_coor, const double* charges, const int* apr, const int* rfa,
                               int nres, int nframes, const double* com, double* dip, cudaStream_t s);

cudaError_t residue_com_launch_f32(const float* coor, long long stride_coor, const double* masses, const int* apr, const int* rfa,
                                   int nres, int nframes, double* com, long long stride_out, cudaStream_t s);
cudaError_t residue_dip_launch_f32(const float* coor, long long stride_coor, const double* charges, const int* apr, const int* rfa,
                                   int nres, int nframes, const double* com, double* dip, long long stride_out, cudaStream_t s);

cudaError_t peaks_measure(int dev, double* fp64_tflops, double* fp32_tflops);

// Compares the kernel's in-range fp64 division / square-root sequences with the IEEE built-ins on
// n_operands random operand sets; out = {division mismatches, sqrt mismatches, near-tie mismatches}.
cudaError_t arith_selftest(int dev, unsigned long long seed, long long n_operands, unsigned long long out[3]);

int tile_i(const PairConfig& cfg);   // core sites per block tile

// tile skipping (gfn_cells.cu)
size_t cells_sort_temp_bytes(long long items);
cudaError_t cells_dense_lists_launch(const float* boxBlk, const float* boxTile, int n_blk, int n_jtiles, int tri, int blk_per_tile, int nframes,
                                     const double* box, double hmax, int* list, long long capacity, int* count, long long* start,
                                     unsigned long long* counters, cudaStream_t s);
cudaError_t cells_sort_launch(const Site* site, Site* sorted, int n, int n_pad, int nframes, const double* box,
                              unsigned long long* keys_in, unsigned long long* keys_out, unsigned* vals_in, unsigned* vals_out,
                              void* tmp, size_t tmp_bytes, cudaStream_t s);
cudaError_t cells_bbox_launch(const Site* sorted, int n, int n_pad, int nframes, const double* box, int tile, int ntiles, float* bbox, cudaStream_t s);
cudaError_t cells_list_launch(const float* boxA, const float* boxB, int n_itiles, int n_jtiles, int TI, int tri, int nframes,
                              const double* box, double hmax, int* list, int* count, long long* start, unsigned long long* visits, cudaStream_t s);

}  // namespace gfn

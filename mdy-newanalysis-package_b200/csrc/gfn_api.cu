// gfn_api.cu - host side of libgfn.so: the C ABI declared in include/gfn.h.
//
// What lives here (reference counterparts in /root/reference/newanalysis/gfunction/gfunction.pyx):
//   * handle = device half of RDF.__init__ (:234-364) - parameters, per-device accumulators
//   * pinned-host frame-batch upload ring on CUDA streams: gfn_enqueue_frame copies the caller's
//     arrays into a pinned batch slot and returns; full batches go H2D on a copy stream while the
//     compute stream works on the previous batch (replaces the synchronous memcpy_htod per frame of
//     the legacy branch, :429-453)
//   * frame-sharded multi-GPU driver: batches are dealt round-robin to the handle's devices; at
//     readout the small per-device histograms are summed with ONE ncclAllReduce (NCCL is loaded
//     with dlopen so the library has no hard dependency when a single GPU is used)
//   * readout / scale / reset (:529-531, :513-515, :552-557)
#include "gfn_kernels.cuh"
#include "../../include/gfn.h"

#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "copy_helper.h"

using namespace gfn;

namespace {

// NVTX range of one host-side stage (the reference has no tracing; Nsight shows these next to the
// kernels).  Header-only NVTX 3: without an attached tool a push/pop is one load and one branch.
struct Range {
    explicit Range(const char* name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
    Range(const Range&) = delete;
    Range& operator=(const Range&) = delete;
};

thread_local char g_create_err[512] = "";

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen (ABI of nccl.h 2.x; only what the histogram all-reduce needs)
// ------------------------------------------------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };   // ncclDataType_t: ncclDouble
enum { ncclSum = 0 };

struct Nccl {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string err;
};

Nccl& nccl()
{
    static Nccl n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (n.lib) break;
    }
    if (!n.lib) { n.err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return n; }
#define GFN_SYM(field, name) *(void**)(&n.field) = dlsym(n.lib, name); if (!n.field) { n.err = std::string("missing NCCL symbol ") + name; return n; }
    GFN_SYM(GetUniqueId, "ncclGetUniqueId")
    GFN_SYM(CommInitRank, "ncclCommInitRank")
    GFN_SYM(CommInitAll, "ncclCommInitAll")
    GFN_SYM(CommDestroy, "ncclCommDestroy")
    GFN_SYM(AllReduce, "ncclAllReduce")
    GFN_SYM(GroupStart, "ncclGroupStart")
    GFN_SYM(GroupEnd, "ncclGroupEnd")
    GFN_SYM(GetErrorString, "ncclGetErrorString")
#undef GFN_SYM
    n.ok = true;
    return n;
}

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
// Readers of caller-managed device arrays (gfn_enqueue_device_frames: the legacy pre_*_gpu path): the only kernel that reads
// the raw arrays of a frame is prep_kernel, and launch_frames records an event right behind it.  gfn_dev_upload waits for the
// events recorded on its device since the last upload - not for the whole device - before it overwrites a buffer, so the
// upload of the next frame overlaps the pair kernels of the previous one.
struct ReaderFence {
    std::mutex mu;
    std::vector<cudaEvent_t> ev[64];
    void add(int dev, cudaEvent_t e) { if (dev < 0 || dev >= 64) return; std::lock_guard<std::mutex> lk(mu); ev[dev].push_back(e); }
    void forget(int dev, cudaEvent_t e)
    {
        if (dev < 0 || dev >= 64) return;
        std::lock_guard<std::mutex> lk(mu);
        std::vector<cudaEvent_t>& v = ev[dev];
        for (size_t k = 0; k < v.size(); ) { if (v[k] == e) { v[k] = v.back(); v.pop_back(); } else ++k; }
    }
    // waits with the lock held: a handle cannot be destroyed (forget) under the events being waited for
    cudaError_t wait(int dev)
    {
        if (dev < 0 || dev >= 64) return cudaDeviceSynchronize();
        std::lock_guard<std::mutex> lk(mu);
        cudaError_t rc = cudaSuccess;
        for (cudaEvent_t e : ev[dev]) { const cudaError_t r = cudaEventSynchronize(e); if (r != cudaSuccess) rc = r; }
        ev[dev].clear();
        return rc;
    }
};
ReaderFence& reader_fence() { static ReaderFence f; return f; }

constexpr int kRing = 3;
constexpr int kXRing = 4;
constexpr int kStatusSlots = 2;   // behind the 8*histo_n accumulators: [0] ranks/devices that hit the fp16 range error, [1] ... a zero-norm dipole

// one source of frames for the ring: host selections (coordinates / dipoles [/ shell matrix]) or float32 atom positions
struct FrameSrc {
    const double* c1 = nullptr; const double* c2 = nullptr; const double* d1 = nullptr; const double* d2 = nullptr;
    long long s_c1 = 0, s_c2 = 0, s_d1 = 0, s_d2 = 0;
    const int* shell_map = nullptr; long long s_map = 0;
    const float* a1 = nullptr; const float* a2 = nullptr; long long s_a1 = 0, s_a2 = 0;
};

struct Batch {
    double* h_pin = nullptr;      // pinned staging: [frame][c1 | d1 | c2 | d2] (absent parts omitted)
    double* d_raw = nullptr;      // same layout on the device
    double* h_box = nullptr;      // pinned [frames][6]
    double* d_box = nullptr;
    Site* siteA = nullptr; Site* siteB = nullptr;
    float*  maxab = nullptr;      // [4][frames]: max |coordinate| of selection 1, 2; dipole-norm flag words of selection 1, 2
    cudaEvent_t uploaded = nullptr, done = nullptr, k0 = nullptr, k1 = nullptr;
    int nframes = 0;
    bool aliased = false;         // selection 2 is selection 1 (same pointer passed)
    bool inflight = false;
    bool timed = false;           // k0/k1 hold an un-harvested pair-kernel timing
    float* h_atoms = nullptr;     // pinned [frame][atoms of selection 1 | atoms of selection 2][3] float32 (atom batches)
    float* d_atoms = nullptr;
    bool atoms = false;           // this batch holds atom positions; COM / dipoles are produced on the device
    bool direct = false;          // the frames are NOT staged in h_pin: the DMA engine reads the caller's page-locked arrays
    FrameSrc dsrc;                // ... starting here (first frame of the batch)
};

struct Dev {
    int dev = 0;
    cudaStream_t copy = nullptr, comp = nullptr;
    Batch ring[kRing];
    int cur = 0;
    int ramp = 2;                         // page-locked feeds: size of the next batch (reset by every sync)
    // per-device share of the handle's statistics (a device is only ever driven by one thread at a time)
    struct Stats { double pairs = 0, frames = 0, h2d = 0, launches = 0, pair_launches = 0, kernel_ms = 0, visits_dense = 0; } st;
    double* gacc = nullptr;               // [8][histo_n] + kStatusSlots status words (filled at readout, summed over ranks)
    double* gsum = nullptr;               // all-reduce result, same layout
    double* h_out = nullptr;              // pinned: readout lands here ([8][histo_n] + status, then kCounters u64)
    double* gpp = nullptr;                // per-particle
    double* h_pp = nullptr;               // pinned staging of gfn_readout_per_particle_add
    size_t h_pp_cap = 0;                  // ... its capacity in doubles
    double* partials = nullptr;
    unsigned long long* counters = nullptr;
    int* item_start = nullptr;
    double* bin_thr = nullptr;            // dense kernel, g000-only handles: bin thresholds in d^2
    PairConfig cfg;
    ncclComm_t comm = nullptr;
    // device-frame scratch (gfn_enqueue_device_frames)
    Site* xsiteA = nullptr; Site* xsiteB = nullptr;
    float* xmax = nullptr; int xcap = 0;
    struct XSlot { double* h_box = nullptr; double* d_box = nullptr; cudaEvent_t k0 = nullptr, k1 = nullptr; bool timed = false; };
    XSlot xs[kXRing];
    int xcur = 0;
    // tile skipping (GFN_FLAG_CELLS, gfn_cells.cu): one workspace per device - every launch runs on the comp stream
    struct Cells {
        Site* sortA = nullptr; Site* sortB = nullptr;                  // [batch_frames][n_pad] records in Morton order
        unsigned long long* keys[2] = {nullptr, nullptr}; unsigned* vals[2] = {nullptr, nullptr};
        void* tmp = nullptr; size_t tmp_bytes = 0;
        float* boxA = nullptr; float* boxB = nullptr;                  // [batch_frames][tiles][6]
        int* list = nullptr; int* count = nullptr; long long* start = nullptr; long long list_cap = 0;
    } cells;
    cudaEvent_t mark[2] = {nullptr, nullptr};
    gfn_residues_t* res[2] = {nullptr, nullptr};   // topology of selection 1 / 2 on this device (gfn_attach_topology)
};

}  // namespace

struct gfn_residues {
    int dev, natoms, nres;
    int* apr = nullptr; int* rfa = nullptr;
    double* masses = nullptr; double* charges = nullptr;
    double* scratch = nullptr; size_t scratch_bytes = 0;     // staging for host-side calls
    char err[256];
};

struct gfn_handle {
    CopyHelper* helper = nullptr;
    int copy_threads = 1;
    CopyPool* pool = nullptr;     // block calls: frames of a batch copied in parallel
    int pool_threads = 0;
    int n1, n2, histo_n, mode_sel;   // histo_n: bins per histogram row (radial bins x shells)
    int histo_nr = 0;                // radial bins
    gfn_shell_cfg shells = {0, 0, 1, 1, 0, 0, 0};   // layout 0: plain g-functions
    long long off_map = 0;           // offset (doubles) of the shell matrix inside a frame
    // atoms through the ring (N1 + N4): selections produced on the device from float32 atom positions
    int na[2] = {0, 0};              // atoms of selection 1 / 2 (0: no topology attached)
    long long atom_cap = 0;          // floats per frame the rings' atom buffers were sized for
    float hmin_f, hmax_f;
    double inv_dr;
    double box[6];
    unsigned flags;
    bool need1, need2;
    int n1_pad, n2_pad;
    int batch_frames;
    int idle_frames;              // frames worth sending on their own when the device has nothing to do
    long long frame_doubles;      // worst case (not aliased)
    long long off_c1, off_d1, off_c2, off_d2;   // offsets (doubles) inside a non-aliased frame
    int n_itiles, n_jtiles, tiles_per_frame;
    double acc_lo = 0, acc_hi = 0; int pos_max = 0;      // dense kernel: acceptance interval of d^2, bin of its upper end
    bool cells = false;           // tile skipping active (GFN_FLAG_CELLS on a ring-kernel handle)
    double st_visits_dense = 0;   // tile visits the un-pruned enumeration has for the frames launched so far
    std::vector<Dev> devs;
    int next_dev = 0;
    // external communicator (one process per GPU)
    int nranks = 1, rank = 0;
    bool ext_comm = false;
    bool comm_ready = false;
    // stats
    double st_pairs = 0, st_frames = 0, st_h2d = 0, st_launches = 0, st_pair_launches = 0, st_kernel_ms = 0;
    char err[512];
};

namespace {

int fail(gfn_t* h, int code, const char* fmt, ...)
{
    char* dst = h ? h->err : g_create_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(h, GFN_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

int validate_box(gfn_t* h, const double* b)
{
    for (int k = 0; k < 3; ++k) {
        if (!(b[k] > 0.0) || !isfinite(b[k])) return fail(h, GFN_EINVAL, "box length %d must be positive and finite", k);
        if (b[3 + k] != b[k] / 2.0) return fail(h, GFN_EINVAL, "box_dim[%d] must equal box_dim[%d]/2 (gfunction.pyx:297-311)", 3 + k, k);
    }
    return GFN_OK;
}

void harvest_timing(Dev& d, Batch& b)
{
    if (b.timed) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, b.k0, b.k1) == cudaSuccess) d.st.kernel_ms += ms;
        b.timed = false;
    }
}

int wait_batch(gfn_t* h, Dev& d, Batch& b)
{
    if (b.inflight) {
        CU(cudaSetDevice(d.dev));
        CU(cudaEventSynchronize(b.done));
        b.inflight = false;
        harvest_timing(d, b);
    }
    return GFN_OK;
}

void fill_params(gfn_t* h, Dev& d, PairParams& p)
{
    memset(&p, 0, sizeof(p));
    p.n1 = h->n1; p.n2 = h->n2; p.tri = (h->n1 == h->n2) && !(h->flags & GFN_FLAG_NO_TRIANGLE);
    p.histo_n = h->histo_n; p.mode_sel = h->mode_sel;
    p.hmin = (double)h->hmin_f; p.hmax = (double)h->hmax_f; p.invdx = h->inv_dr;
    p.n_itiles = h->n_itiles; p.n_jtiles = h->n_jtiles; p.tiles_per_frame = h->tiles_per_frame;
    p.tile_start = d.item_start;
    p.partials = d.partials; p.gacc = d.gacc; p.gpp = d.gpp; p.counters = d.counters;
    p.flags = h->flags;
    p.nrep = d.cfg.nrep;
    p.poll_ns = 120; p.poll_reps = 1;
    if (const char* e = getenv("GFN_POLL_NS")) p.poll_ns = (unsigned)atoi(e);
    if (const char* e = getenv("GFN_POLL_REPS")) p.poll_reps = atoi(e);
    if (const char* e = getenv("GFN_DEBUG")) p.debug = atoi(e);
    p.strideA = h->n1_pad; p.strideB = h->n2_pad;
    p.shell_mode = h->shells.layout; p.nshells = h->shells.nshells; p.histo_nr = h->histo_nr;
    p.apr1 = h->shells.apr_core; p.apr2 = h->shells.apr_surround; p.map_ld = h->shells.map_ld; p.excl_self = h->shells.exclude_self;
    p.bin_thr = d.bin_thr; p.acc_lo = h->acc_lo; p.acc_hi = h->acc_hi; p.pos_max = h->pos_max;
}

// ---- exact bin thresholds in d^2 (dense kernel, g000-only handles) ------------------------------------------------
// The reference decides a pair's fate from dist = sqrt(d2) (gfunction.pyx:148-150):
//   reject if dist < hmin or dist > hmax or floor(dist*invdx) == 0;   pos = floor((dist - hmin) * invdx)
// sqrt, the subtraction, the multiplication (each correctly rounded) and floor are monotone non-decreasing, so both the
// acceptance and the bin are step functions of d2: the steps are found by bisection over the (ordered) bit patterns of the
// non-negative doubles, evaluating the reference's own expressions with the host's IEEE arithmetic (volatile: one rounding
// per operation, no contraction).
template <class Pred>
double min_true(Pred pred, double hi)
{
    if (!pred(hi)) return INFINITY;
    if (pred(0.0)) return 0.0;
    unsigned long long lo = 0, hb;
    memcpy(&hb, &hi, 8);
    while (hb - lo > 1) {          // invariant: !pred(lo), pred(hb)
        const unsigned long long mid = lo + (hb - lo) / 2;
        double t;
        memcpy(&t, &mid, 8);
        if (pred(t)) hb = mid; else lo = mid;
    }
    double t;
    memcpy(&t, &hb, 8);
    return t;
}

// Fills thr[0 .. histo_n+3]; returns false if the handle's parameters do not suit the table (the caller then keeps the
// ring kernel).
bool bin_thresholds(double hmin, double hmax, double invdx, int histo_n, std::vector<double>& thr, double* acc_lo, double* acc_hi, int* pos_max)
{
    if (!(hmax > 0.0) || !(hmax < 1e15) || !(invdx < 1e15) || !(invdx > 1e-15) || hmin < 0.0 || !(hmin <= hmax)) return false;
    auto dist_of = [](double t) { volatile double s = sqrt(t); return (double)s; };
    auto bin_of = [&](double t) { volatile double s = sqrt(t); volatile double d = s - hmin; volatile double q = d * invdx; return floor((double)q); };
    const double top = (hmax + 4.0 / invdx + 1.0) * (hmax + 4.0 / invdx + 1.0) * 4.0;
    const double lo = min_true([&](double t) { volatile double q = dist_of(t) * invdx; return dist_of(t) >= hmin && !((double)q < 1.0); }, top);
    const double above = min_true([&](double t) { return dist_of(t) > hmax; }, top);
    if (!isfinite(lo) || !isfinite(above) || !(above > 0.0)) return false;
    unsigned long long ab;
    memcpy(&ab, &above, 8);
    ab -= 1;
    double hi;
    memcpy(&hi, &ab, 8);
    if (!(lo <= hi)) return false;                      // nothing can be in range: leave such handles to the ring kernel
    const double pm = bin_of(hi);
    if (!(pm >= 0.0) || pm > (double)histo_n + 1.0) return false;
    *acc_lo = lo; *acc_hi = hi; *pos_max = (int)pm;
    thr.assign((size_t)histo_n + 4, INFINITY);
    for (int k = 0; k <= *pos_max + 1; ++k)
        thr[k] = min_true([&](double t) { return dist_of(t) >= hmin && bin_of(t) >= (double)k; }, top);
    return true;
}

// prep + pair + finalize for nframes frames whose raw arrays are on the device
int launch_frames(gfn_t* h, Dev& d, int nframes,
                  const double* c1, long long s_c1, const double* d1, long long s_d1,
                  const double* c2, long long s_c2, const double* d2, long long s_d2, bool aliased,
                  Site* siteA, Site* siteB, float* maxab, const double* d_box,
                  cudaEvent_t k0, cudaEvent_t k1, const int* d_map = nullptr, long long s_map = 0)
{
    Range nvtx_range("gfn: launch batch");
    cudaStream_t s = d.comp;
    CU(cudaMemsetAsync(maxab, 0, sizeof(float) * 4 * nframes, s));
    unsigned* nflag = reinterpret_cast<unsigned*>(maxab + 2 * nframes);
    const int hf = d.cfg.filter_fp64 == 2;
    CU(prep_launch(c1, s_c1, h->need1 ? d1 : nullptr, s_d1, siteA, maxab, nflag, h->n1, h->n1_pad, nframes, d_box, hf, d.counters, s));
    d.st.launches += 1;
    if (!aliased) {
        CU(prep_launch(c2, s_c2, h->need2 ? d2 : nullptr, s_d2, siteB, maxab + nframes, nflag + nframes, h->n2, h->n2_pad, nframes, d_box, hf, d.counters, s));
        d.st.launches += 1;
    }
    PairParams p;
    fill_params(h, d, p);
    // Two DIFFERENT selections of equal size: the reference still walks the triangle j >= i (quirk Q1), now a relation
    // between the original indices of two arrays - not a triangle of positions once each is sorted on its own.  Such
    // launches keep the full enumeration.
    const bool tri_launch = (h->n1 == h->n2) && !(h->flags & GFN_FLAG_NO_TRIANGLE);
    if (h->cells && d.cfg.dense == 1 && tri_launch && !aliased && (long long)h->tiles_per_frame * nframes > dense_max_units(d.cfg))
        return fail(h, GFN_EINVAL, "two different selections of equal size cannot use tile skipping (their triangle relates original indices, quirk Q1) and a full frame of this size does not fit the dense kernel: create the handle without GFN_FLAG_CELLS");
    if (h->cells && !(tri_launch && !aliased)) {
        // Morton order + visit lists: the pair kernel runs on the sorted copies
        Dev::Cells& c = d.cells;
        const int TI = tile_i(d.cfg);
        CU(cells_sort_launch(siteA, c.sortA, h->n1, h->n1_pad, nframes, d_box, c.keys[0], c.keys[1], c.vals[0], c.vals[1], c.tmp, c.tmp_bytes, s));
        CU(cells_bbox_launch(c.sortA, h->n1, h->n1_pad, nframes, d_box, TI, h->n_itiles, c.boxA, s));
        if (!aliased) CU(cells_sort_launch(siteB, c.sortB, h->n2, h->n2_pad, nframes, d_box, c.keys[0], c.keys[1], c.vals[0], c.vals[1], c.tmp, c.tmp_bytes, s));
        Site* sb = aliased ? c.sortA : c.sortB;
        CU(cells_bbox_launch(sb, h->n2, h->n2_pad, nframes, d_box, kTJ, h->n_jtiles, c.boxB, s));
        const int tri = (h->n1 == h->n2) && !(h->flags & GFN_FLAG_NO_TRIANGLE);
        double reach = (double)h->hmax_f;
        if (const char* e = getenv("GFN_CELLS_REACH")) reach = atof(e);          // experiments: a larger reach visits more tiles (never fewer pairs)
        if (d.cfg.dense) {
            CU(cells_dense_lists_launch(c.boxA, c.boxB, h->n_itiles, h->n_jtiles, tri, kTJ / TI, nframes, d_box, reach, c.list, c.list_cap, c.count, c.start, d.counters, s));
            p.visit_cap = c.list_cap;
        } else {
            CU(cells_list_launch(c.boxA, c.boxB, h->n_itiles, h->n_jtiles, TI, tri, nframes, d_box, reach, c.list, c.count, c.start, d.counters + 6, s));
        }
        d.st.launches += aliased ? 8 : 12;
        d.st.visits_dense += (double)h->tiles_per_frame * nframes;
        siteA = c.sortA; siteB = sb;
        p.visit_start = c.start; p.visit_list = c.list;
    }
    p.siteA = siteA; p.maxA = maxab;
    p.siteB = aliased ? siteA : siteB; p.maxB = aliased ? maxab : maxab + nframes;
    p.nflagA = nflag; p.nflagB = aliased ? nflag : nflag + nframes;
    p.box = d_box; p.nframes = nframes;
    p.shell_map = d_map; p.stride_map = s_map;
    CU(cudaEventRecord(k0, s));
    CU(pair_launch(d.cfg, p, s));
    CU(cudaEventRecord(k1, s));
    d.st.launches += 1; d.st.pair_launches += 1;
    if (d.cfg.hist_mode == 0) {
        CU(finalize_launch(d.partials, d.cfg.grid, h->histo_n, d.cfg.nslot, d.gacc, s));
        d.st.launches += 1;
    }
    const bool tri = (h->n1 == h->n2) && !(h->flags & GFN_FLAG_NO_TRIANGLE);
    const double per = tri ? 0.5 * (double)h->n1 * ((double)h->n1 + 1.0) : (double)h->n1 * (double)h->n2;
    d.st.pairs += per * nframes;
    d.st.frames += nframes;
    return GFN_OK;
}

int submit(gfn_t* h, Dev& d)
{
    Batch& b = d.ring[d.cur];
    if (b.nframes == 0) return GFN_OK;
    CU(cudaSetDevice(d.dev));
    const long long fd = b.aliased ? (h->off_c2) : h->frame_doubles;    // aliased frames hold only c1|d1
    size_t bytes = (size_t)fd * b.nframes * sizeof(double);
    if (b.atoms) {
        // the frame crosses PCIe as float32 atom positions; centres of mass and dipoles (helpers.pyx:71-123) are
        // written by the residue kernels into the same [c1|d1|c2|d2] layout the host path uploads
        const long long na1 = h->na[0], na2 = b.aliased ? 0 : (h->na[1] ? h->na[1] : h->na[0]);
        const long long fa = 3 * (na1 + na2);
        bytes = (size_t)fa * b.nframes * sizeof(float);
        CU(cudaMemcpyAsync(b.d_atoms, b.h_atoms, bytes, cudaMemcpyHostToDevice, d.copy));
    } else if (b.direct) {
        // page-locked caller arrays: one strided DMA per part, [frame] rows of the part's width into the frame layout
        const FrameSrc& q = b.dsrc;
        const size_t dp = (size_t)fd * sizeof(double);
        auto part = [&](long long off, const double* srcp, long long stride, size_t width) -> cudaError_t {
            const size_t sp = b.nframes > 1 ? (size_t)stride * sizeof(double) : width;
            return cudaMemcpy2DAsync(b.d_raw + off, dp, srcp, sp, width, (size_t)b.nframes, cudaMemcpyHostToDevice, d.copy);
        };
        bytes = 0;
        CU(part(h->off_c1, q.c1, q.s_c1, sizeof(double) * 3 * h->n1)); bytes += sizeof(double) * 3 * h->n1;
        if (h->need1) { CU(part(h->off_d1, q.d1, q.s_d1, sizeof(double) * 3 * h->n1)); bytes += sizeof(double) * 3 * h->n1; }
        if (!b.aliased) {
            CU(part(h->off_c2, q.c2, q.s_c2, sizeof(double) * 3 * h->n2)); bytes += sizeof(double) * 3 * h->n2;
            if (h->need2) { CU(part(h->off_d2, q.d2, q.s_d2, sizeof(double) * 3 * h->n2)); bytes += sizeof(double) * 3 * h->n2; }
        }
        bytes *= (size_t)b.nframes;
    } else {
        CU(cudaMemcpyAsync(b.d_raw, b.h_pin, bytes, cudaMemcpyHostToDevice, d.copy));
    }
    CU(cudaMemcpyAsync(b.d_box, b.h_box, sizeof(double) * 6 * b.nframes, cudaMemcpyHostToDevice, d.copy));
    CU(cudaEventRecord(b.uploaded, d.copy));
    CU(cudaStreamWaitEvent(d.comp, b.uploaded, 0));
    d.st.h2d += (double)bytes + 48.0 * b.nframes;
    if (b.atoms) {
        const long long na1 = h->na[0], na2 = b.aliased ? 0 : (h->na[1] ? h->na[1] : h->na[0]);
        const long long fa = 3 * (na1 + na2);
        for (int k = 0; k < (b.aliased ? 1 : 2); ++k) {
            const gfn_residues* r = (k == 1 && d.res[1]) ? d.res[1] : d.res[0];
            const float* at = b.d_atoms + (k == 0 ? 0 : 3 * na1);
            double* com = b.d_raw + (k == 0 ? h->off_c1 : h->off_c2);
            double* dip = b.d_raw + (k == 0 ? h->off_d1 : h->off_d2);
            const bool need = k == 0 ? h->need1 : h->need2;
            CU(residue_com_launch_f32(at, fa, r->masses, r->apr, r->rfa, r->nres, b.nframes, com, fd, d.comp));
            if (need) CU(residue_dip_launch_f32(at, fa, r->charges, r->apr, r->rfa, r->nres, b.nframes, com, dip, fd, d.comp));
            d.st.launches += need ? 2 : 1;
        }
    }
    int rc = launch_frames(h, d, b.nframes,
                           b.d_raw + h->off_c1, fd, b.d_raw + h->off_d1, fd,
                           b.d_raw + h->off_c2, fd, b.d_raw + h->off_d2, fd, b.aliased,
                           b.siteA, b.siteB, b.maxab, b.d_box, b.k0, b.k1,
                           h->shells.layout ? reinterpret_cast<const int*>(b.d_raw + h->off_map) : nullptr, 2 * fd);
    if (rc) return rc;
    b.timed = true;
    CU(cudaEventRecord(b.done, d.comp));
    b.inflight = true;
    d.cur = (d.cur + 1) % kRing;
    Batch& nb = d.ring[d.cur];
    nb.nframes = 0;
    return GFN_OK;
}

int check_zerodiv(gfn_t* h)
{
    for (Dev& d : h->devs) {
        unsigned long long c[kCounters];
        CU(cudaSetDevice(d.dev));
        CU(cudaMemcpy(c, d.counters, sizeof(c), cudaMemcpyDeviceToHost));
        if (c[1] & 2ull) return fail(h, GFN_EINVAL, "coordinates lie more than 30000 box lengths from the origin: the fp16 pre-filter cannot represent them, create the handle without GFN_FLAG_FILTER_FP16");
        if (c[1] & 16ull) return fail(h, GFN_ECUDA, "internal: the visit list of a cell-sorted launch overflowed its buffer");
        if (c[1] & 4ull) return fail(h, GFN_ECUDA, "a bin of one CTA took more than 2^20 pairs in one launch: beyond the dense kernel's fixed-point range (create the handle with GFN_FLAG_NO_DENSE)");
        if (c[1] & 8ull) return fail(h, GFN_ECUDA, "checked build: a device-side index / protocol check failed");
        if (c[1] & 1ull) return fail(h, GFN_EZERODIV, "float division");
    }
    return GFN_OK;
}

int setup_comm(gfn_t* h)
{
    if (h->comm_ready) return GFN_OK;
    const int nd = (int)h->devs.size();
    if (nd > 1) {
        Nccl& n = nccl();
        if (!n.ok) return fail(h, GFN_ENCCL, "%s", n.err.c_str());
        std::vector<ncclComm_t> comms(nd);
        std::vector<int> ids(nd);
        for (int i = 0; i < nd; ++i) ids[i] = h->devs[i].dev;
        int rc = n.CommInitAll(comms.data(), nd, ids.data());
        if (rc != ncclSuccess) return fail(h, GFN_ENCCL, "ncclCommInitAll: %s", n.GetErrorString(rc));
        for (int i = 0; i < nd; ++i) h->devs[i].comm = comms[i];
    }
    h->comm_ready = true;
    return GFN_OK;
}

void free_dev(Dev& d)
{
    cudaSetDevice(d.dev);
    if (d.comp) cudaStreamSynchronize(d.comp);
    if (d.copy) cudaStreamSynchronize(d.copy);
    for (Batch& b : d.ring) {
        if (b.h_pin) cudaFreeHost(b.h_pin);
        if (b.h_box) cudaFreeHost(b.h_box);
        if (b.h_atoms) cudaFreeHost(b.h_atoms);
        cudaFree(b.d_atoms);
        cudaFree(b.d_raw); cudaFree(b.d_box); cudaFree(b.siteA); cudaFree(b.siteB); cudaFree(b.maxab);
        if (b.uploaded) cudaEventDestroy(b.uploaded);
        if (b.done) cudaEventDestroy(b.done);
        if (b.k0) cudaEventDestroy(b.k0);
        if (b.k1) cudaEventDestroy(b.k1);
    }
    if (d.h_out) cudaFreeHost(d.h_out);
    if (d.h_pp) cudaFreeHost(d.h_pp);
    cudaFree(d.gacc); cudaFree(d.gsum); cudaFree(d.gpp); cudaFree(d.partials); cudaFree(d.counters); cudaFree(d.item_start); cudaFree(d.bin_thr);
    cudaFree(d.xsiteA); cudaFree(d.xsiteB); cudaFree(d.xmax);
    cudaFree(d.cells.sortA); cudaFree(d.cells.sortB); cudaFree(d.cells.keys[0]); cudaFree(d.cells.keys[1]); cudaFree(d.cells.vals[0]); cudaFree(d.cells.vals[1]);
    cudaFree(d.cells.tmp); cudaFree(d.cells.boxA); cudaFree(d.cells.boxB); cudaFree(d.cells.list); cudaFree(d.cells.count); cudaFree(d.cells.start);
    for (Dev::XSlot& x : d.xs) {
        if (x.k0) reader_fence().forget(d.dev, x.k0);
        if (x.h_box) cudaFreeHost(x.h_box);
        cudaFree(x.d_box);
        if (x.k0) cudaEventDestroy(x.k0);
        if (x.k1) cudaEventDestroy(x.k1);
    }
    for (cudaEvent_t e : d.mark) if (e) cudaEventDestroy(e);
    for (gfn_residues_t*& r : d.res) { gfn_residues_destroy(r); r = nullptr; }
    if (d.comm && nccl().ok) nccl().CommDestroy(d.comm);
    if (d.copy) cudaStreamDestroy(d.copy);
    if (d.comp) cudaStreamDestroy(d.comp);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int gfn_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* gfn_last_error(const gfn_t* h) { return h ? h->err : g_create_err; }

static int create_impl(gfn_t** out, int n1, int n2, int histo_nr, float histo_min, float histo_max,
                       double inv_dr, int mode_sel, const double box_dim[6], const int* devices, int ndev, unsigned flags,
                       const gfn_shell_cfg* sc)
{
    gfn_t* h = nullptr;
    int histo_n = histo_nr;
    if (sc) {
        if (sc->layout != 1 && sc->layout != 2) return fail(h, GFN_EINVAL, "shell layout must be 1 ([bin][shell]) or 2 ([shell][bin])");
        if (sc->nshells < 1) return fail(h, GFN_EINVAL, "nshells must be >= 1 (got %d)", sc->nshells);
        if (sc->apr_core < 1 || sc->apr_surround < 1)
            return fail(h, GFN_EINVAL, "sites per molecule must be >= 1 (got %d, %d): more molecules than sites", sc->apr_core, sc->apr_surround);
        if (histo_nr >= 1 && (long long)histo_nr * sc->nshells > 2000000000ll / 8) return fail(h, GFN_EINVAL, "histo_n * nshells too large");
        const long long need = (long long)((n1 + sc->apr_core - 1) / sc->apr_core - 1) * sc->map_ld + (n2 + sc->apr_surround - 1) / sc->apr_surround;
        if (sc->map_ld < 1 || sc->map_elems < need)
            return fail(h, GFN_EINVAL, "shell matrix too small: the pair loop indexes up to element %lld (row stride %d), the matrix has %lld", need - 1, sc->map_ld, sc->map_elems);
        histo_n = histo_nr * sc->nshells;
    }
    if (!out) return fail(h, GFN_EINVAL, "out is NULL");
    *out = nullptr;
    if (n1 < 1 || n2 < 1) return fail(h, GFN_EINVAL, "selection sizes must be >= 1 (got %d, %d)", n1, n2);
    if (histo_n < 1) return fail(h, GFN_EINVAL, "histo_n must be >= 1 (got %d)", histo_n);
    if (mode_sel < 0 || mode_sel > 127) return fail(h, GFN_EINVAL, "mode_sel out of range: %d", mode_sel);
    if (!(inv_dr > 0.0) || !isfinite(inv_dr)) return fail(h, GFN_EINVAL, "inv_dr must be positive and finite");
    if (!box_dim) return fail(h, GFN_EINVAL, "box_dim is NULL");
    int rc = validate_box(h, box_dim);
    if (rc) return rc;
    const int navail = gfn_device_count();
    if (navail < 1) return fail(h, GFN_ENODEV, "no CUDA device available; libgfn has no CPU fallback");
    std::vector<int> devs;
    if (!devices || ndev <= 0) devs.push_back(0);
    else for (int i = 0; i < ndev; ++i) {
        if (devices[i] < 0 || devices[i] >= navail) return fail(h, GFN_EINVAL, "device ordinal %d out of range (0..%d)", devices[i], navail - 1);
        devs.push_back(devices[i]);
    }
    for (int dv : devs) {
        cudaDeviceProp pr;
        if (cudaGetDeviceProperties(&pr, dv) != cudaSuccess) return fail(h, GFN_ECUDA, "cudaGetDeviceProperties(%d) failed", dv);
        if (pr.major != 10) return fail(h, GFN_ENODEV, "device %d is sm_%d%d; libgfn is built for sm_100a (B200) only", dv, pr.major, pr.minor);
    }
    if ((flags & GFN_FLAG_PER_PARTICLE) && devs.size() > 1)
        return fail(h, GFN_EINVAL, "GFN_FLAG_PER_PARTICLE supports a single device");

    h = new gfn_handle();
    h->err[0] = 0;
    h->n1 = n1; h->n2 = n2; h->histo_n = histo_n; h->mode_sel = mode_sel;
    h->hmin_f = histo_min; h->hmax_f = histo_max; h->inv_dr = inv_dr; h->flags = flags;
    memcpy(h->box, box_dim, sizeof(h->box));
    h->histo_nr = histo_nr;
    if (sc) h->shells = *sc;
    h->need1 = mode_sel & (2 | 4 | 16 | 32);       // gfunction.pyx:121 / :411
    h->need2 = mode_sel & (2 | 8 | 16 | 64);       // gfunction.pyx:135 / :414
    h->n1_pad = (n1 + kPadTo - 1) / kPadTo * kPadTo;
    h->n2_pad = (n2 + kPadTo - 1) / kPadTo * kPadTo;
    h->off_c1 = 0;
    h->off_d1 = 3ll * n1;
    h->off_c2 = h->off_d1 + (h->need1 ? 3ll * n1 : 0);
    h->off_d2 = h->off_c2 + 3ll * n2;
    h->frame_doubles = h->off_d2 + (h->need2 ? 3ll * n2 : 0);
    if (sc) { h->off_map = h->frame_doubles; h->frame_doubles += (sc->map_elems + 1) / 2; }

    // batch size: ~2^33 pair evaluations or 64 MB of frames per batch, whichever is smaller
    const bool tri_create = (n1 == n2) && !(flags & GFN_FLAG_NO_TRIANGLE);
    const double per = tri_create ? 0.5 * (double)n1 * ((double)n1 + 1.0) : (double)n1 * (double)n2;
    long long bf = (long long)(8589934592.0 / per);
    const long long by_bytes = (64ll << 20) / (h->frame_doubles * 8);
    if (bf > by_bytes) bf = by_bytes;
    if (bf > 4096) bf = 4096;
    if (bf < 1) bf = 1;
    if (const char* e = getenv("GFN_BATCH_FRAMES")) { long long v = atoll(e); if (v >= 1 && v <= 65535) bf = v; }
    h->batch_frames = (int)bf;
    // a helper thread for the ring copies when the box has cores to spare: at least 8 per local rank (LOCAL_WORLD_SIZE is
    // what torchrun exports; measured on the 16-core box: +2 % end to end at one rank, not measured beyond two)
    {
        int ranks = 1;
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) { const int v = atoi(e); if (v > 1) ranks = v; }
        h->copy_threads = (int)std::thread::hardware_concurrency() >= 8 * ranks ? 2 : 1;
        if (const char* e = getenv("GFN_COPY_THREADS")) { const int v = atoi(e); if (v >= 1 && v <= 2) h->copy_threads = v; }
        // block calls: the frames of a batch are copied by a pool - two workers per device this handle feeds, within a fair
        // share of the box's cores (the caller copies too)
        const int cores = (int)std::thread::hardware_concurrency() / ranks;
        int pt = 2 * (int)devs.size();
        if (pt > cores / 2) pt = cores / 2;
        if (pt > 16) pt = 16;
        if (h->copy_threads < 2 && devs.size() < 2) pt = 0;
        if (const char* e = getenv("GFN_POOL_THREADS")) { const int v = atoi(e); if (v >= 0 && v <= 64) pt = v; }
        h->pool_threads = pt < 0 ? 0 : pt;
    }
    // an idle device (first frames of a run, right after a readout) should not wait for a full batch: a quarter of
    // one, or ~2^28 pair evaluations (0.2 ms of kernel time) if that is less
    {
        long long idle = (long long)(268435456.0 / per) + 1;
        const long long quarter = (bf + 3) / 4;
        h->idle_frames = (int)(idle < quarter ? idle : quarter);
        if (h->idle_frames < 1) h->idle_frames = 1;
    }

#define CUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(nullptr, GFN_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); for (Dev& dd : h->devs) free_dev(dd); delete h; return GFN_ECUDA; } } while (0)
    double cells_share = 0.05;        // expected share of visited tile pairs (forced handles: assume few)
    if ((flags & GFN_FLAG_CELLS) && !getenv("GFN_CELLS_FORCE")) {
        // Tile skipping is worth it only when enough tile pairs can be skipped.  Expected share of visited tile pairs: a tile of
        // T sorted sites fills a blob of edge (V T / n)^(1/3); Morton blobs measure ~1.5 cubes across; a core blob sees the
        // surround blobs within histo_max of it.  Measured: orientation modes (dense kernel, 64-row core blocks) gain below ~15 %
        // (configs[4]: 1.3 % visited, 13 x faster; configs[3] solvent: 14 %, 1.09 x faster; configs[2]: 46 %, 1.9 x slower); g000-only handles (count kernel,
        // 128-row blocks, ~1000 x 10^9 visited pairs/s whatever the density: configs[2] shape 47 % visited, 1.17 x faster; configs[4]
        // shape 1.3 %, 36 x faster) gain below ~50 %; pair_kernel (GFN_FLAG_NO_DENSE,
        // 512-row core tiles), whose survivor path is built for evenly sparse tiles, below 15 %.
        const double vol = box_dim[0] * box_dim[1] * box_dim[2];
        const bool g000 = mode_sel == 1;
        const bool ring = (flags & GFN_FLAG_NO_DENSE) != 0;          // pair_kernel: 512-row core tiles; count / dense kernel: 128 / 64 rows
        const double ec = cbrt(vol * (ring ? 512.0 : (g000 ? 128.0 : 64.0)) / (double)(n1 > 0 ? n1 : 1)), es = cbrt(vol * kTJ / (double)(n2 > 0 ? n2 : 1));
        const double w = 1.5 * (ec + es) + 2.0 * (double)histo_max;
        double share = 1.0;
        for (int k = 0; k < 3; ++k) share *= fmin(1.0, w / box_dim[k]);
        if (share > (ring ? 0.15 : (g000 ? 0.5 : 0.15))) flags &= ~GFN_FLAG_CELLS;
        h->flags = flags;
        cells_share = share;
    }
    if (sc) { flags &= ~GFN_FLAG_CELLS; h->flags = flags; }
    h->devs.resize(devs.size());
    for (size_t i = 0; i < devs.size(); ++i) h->devs[i].dev = devs[i];
    for (Dev& d : h->devs) {
        CUC(cudaSetDevice(d.dev));
        // expected share of in-range pairs in a uniform box: decides between the ring kernel and the dense kernel
        const double vol = box_dim[0] * box_dim[1] * box_dim[2];
        const double share = fmin(1.0, 4.18879020478639098 * (double)histo_max * (double)histo_max * (double)histo_max / vol);
        const HandleGeom geom = {(double)(float)histo_max, inv_dr, fmax(box_dim[0], fmax(box_dim[1], box_dim[2])), n1, n2, tri_create ? 1 : 0, (flags & GFN_FLAG_CELLS) ? cells_share : 0.0};
        CUC(pair_configure(d.dev, n1, histo_n, mode_sel, flags, sc ? 1 : 0, share, geom, &d.cfg));
    }
    std::vector<double> thr;
    if (h->devs[0].cfg.dense &&
        !bin_thresholds((double)histo_min, (double)histo_max, inv_dr, histo_n, thr, &h->acc_lo, &h->acc_hi, &h->pos_max)) {
        // parameters the threshold table does not cover: the ring kernel serves the handle
        flags |= GFN_FLAG_NO_DENSE; flags &= ~GFN_FLAG_DENSE;
        h->flags = flags;
        for (Dev& d : h->devs) {
            CUC(cudaSetDevice(d.dev));
            CUC(pair_configure(d.dev, n1, histo_n, mode_sel, flags, sc ? 1 : 0, 0.0, HandleGeom{0.0, 0.0, 0.0, n1, n2, tri_create ? 1 : 0, 0.0}, &d.cfg));
        }
    }
    h->cells = (flags & GFN_FLAG_CELLS) && !sc;
    // work decomposition (identical on all devices)
    {
        const PairConfig& c = h->devs[0].cfg;
        const int TI = tile_i(c);
        h->n_itiles = (n1 + TI - 1) / TI;
        h->n_jtiles = (n2 + kTJ - 1) / kTJ;
        std::vector<int> start;
        if (c.dense) {
            // dense kernels: units of 64 (count kernel: 128) core rows per surround tile; prefix of units per surround tile (a triangular frame
            // needs only the row blocks that reach the tile's last column: rows <= 128 jt + 127)
            start.assign(h->n_jtiles + 1, 0);
            for (int jt = 0; jt < h->n_jtiles; ++jt) {
                long long cnt = h->n_itiles;
                if (tri_create && c.dense == 2 && jt + 1 < cnt) cnt = jt + 1;             // count kernel: 128-row blocks
                if (tri_create && c.dense == 1 && 2ll * jt + 2 < cnt) cnt = 2ll * jt + 2;
                start[jt + 1] = start[jt] + (int)cnt;
            }
            h->tiles_per_frame = start[h->n_jtiles];
            if (c.dense == 1) {
                // fixed-point sums of the dense kernel: a launch holds at most dense_max_units units
                double upf = (double)(h->tiles_per_frame > 0 ? h->tiles_per_frame : 1);
                if (flags & GFN_FLAG_CELLS) upf *= fmin(1.0, 2.0 * cells_share);
                long long cap = (long long)((double)dense_max_units(c) / fmax(upf, 1.0));
                if (cap < 1) cap = 1;             // (dense_configure has checked that one frame fits)
                if (cap < h->batch_frames) h->batch_frames = (int)cap;
                if (h->idle_frames > h->batch_frames) h->idle_frames = h->batch_frames;
            }
        } else {
            // prefix of surround-tile visits per core tile (the triangle starts at the tile holding the diagonal)
            start.assign(h->n_itiles + 1, 0);
            for (int t = 0; t < h->n_itiles; ++t) {
                const int first = tri_create ? (t * TI) / kTJ : 0;
                const int tiles = h->n_jtiles - first;
                start[t + 1] = start[t] + (tiles > 0 ? tiles : 0);
            }
            h->tiles_per_frame = start[h->n_itiles];
        }
        for (Dev& d : h->devs) {
            CUC(cudaSetDevice(d.dev));
            CUC(cudaMalloc(&d.item_start, sizeof(int) * start.size()));
            CUC(cudaMemcpy(d.item_start, start.data(), sizeof(int) * start.size(), cudaMemcpyHostToDevice));
            if (!thr.empty() && c.dense == 2) {
                CUC(cudaMalloc(&d.bin_thr, sizeof(double) * thr.size()));
                CUC(cudaMemcpy(d.bin_thr, thr.data(), sizeof(double) * thr.size(), cudaMemcpyHostToDevice));
            }
        }
    }
    for (Dev& d : h->devs) {
        CUC(cudaSetDevice(d.dev));
        CUC(cudaStreamCreateWithFlags(&d.copy, cudaStreamNonBlocking));
        CUC(cudaStreamCreateWithFlags(&d.comp, cudaStreamNonBlocking));
        CUC(cudaMalloc(&d.gacc, sizeof(double) * (8 * (size_t)histo_n + kStatusSlots)));
        CUC(cudaMalloc(&d.gsum, sizeof(double) * (8 * (size_t)histo_n + kStatusSlots)));
        CUC(cudaMemset(d.gacc, 0, sizeof(double) * (8 * (size_t)histo_n + kStatusSlots)));
        CUC(cudaHostAlloc(&d.h_out, sizeof(double) * (8 * (size_t)histo_n + kStatusSlots) + sizeof(unsigned long long) * kCounters, cudaHostAllocDefault));
        CUC(cudaMalloc(&d.counters, sizeof(unsigned long long) * kCounters));
        CUC(cudaMemset(d.counters, 0, sizeof(unsigned long long) * kCounters));
        if (d.cfg.hist_mode == 0) CUC(cudaMalloc(&d.partials, sizeof(double) * (size_t)d.cfg.grid * histo_n * d.cfg.nslot));
        if (h->cells) {
            Dev::Cells& c = d.cells;
            const size_t bf = (size_t)h->batch_frames;
            const size_t np = (size_t)(h->n1_pad > h->n2_pad ? h->n1_pad : h->n2_pad);
            CUC(cudaMalloc(&c.sortA, sizeof(Site) * (size_t)h->n1_pad * bf));
            CUC(cudaMalloc(&c.sortB, sizeof(Site) * (size_t)h->n2_pad * bf));
            for (int k = 0; k < 2; ++k) {
                CUC(cudaMalloc(&c.keys[k], sizeof(unsigned long long) * np * bf));
                CUC(cudaMalloc(&c.vals[k], sizeof(unsigned) * np * bf));
            }
            c.tmp_bytes = cells_sort_temp_bytes((long long)(np * bf));
            CUC(cudaMalloc(&c.tmp, c.tmp_bytes ? c.tmp_bytes : 16));
            CUC(cudaMalloc(&c.boxA, sizeof(float) * 6 * (size_t)h->n_itiles * bf));
            CUC(cudaMalloc(&c.boxB, sizeof(float) * 6 * (size_t)h->n_jtiles * bf));
            if (d.cfg.dense) {
                // dense / count kernel: entries are (frame, surround tile), lists are compact; room for the un-pruned unit count
                c.list_cap = (long long)h->tiles_per_frame * (long long)bf;
                CUC(cudaMalloc(&c.list, sizeof(int) * (size_t)c.list_cap));
                CUC(cudaMalloc(&c.count, sizeof(int) * (size_t)h->n_jtiles * bf));
                CUC(cudaMalloc(&c.start, sizeof(long long) * ((size_t)h->n_jtiles * bf + 1)));
            } else {
                CUC(cudaMalloc(&c.list, sizeof(int) * (size_t)h->n_itiles * (size_t)h->n_jtiles * bf));
                CUC(cudaMalloc(&c.count, sizeof(int) * (size_t)h->n_itiles * bf));
                CUC(cudaMalloc(&c.start, sizeof(long long) * ((size_t)h->n_itiles * bf + 1)));
            }
        }
        if (flags & GFN_FLAG_PER_PARTICLE) {
            const size_t bytes = sizeof(double) * 7ull * n1 * histo_n;
            CUC(cudaMalloc(&d.gpp, bytes));
            CUC(cudaMemset(d.gpp, 0, bytes));
        }
        for (Dev::XSlot& x : d.xs) {
            CUC(cudaEventCreateWithFlags(&x.k0, cudaEventDefault));
            CUC(cudaEventCreateWithFlags(&x.k1, cudaEventDefault));
            CUC(cudaHostAlloc(&x.h_box, sizeof(double) * 6 * h->batch_frames, cudaHostAllocDefault));
            CUC(cudaMalloc(&x.d_box, sizeof(double) * 6 * h->batch_frames));
        }
        CUC(cudaEventCreateWithFlags(&d.mark[0], cudaEventDefault));
        CUC(cudaEventCreateWithFlags(&d.mark[1], cudaEventDefault));
        for (Batch& b : d.ring) {
            const size_t fb = sizeof(double) * (size_t)h->frame_doubles * h->batch_frames;
            CUC(cudaHostAlloc(&b.h_pin, fb, cudaHostAllocDefault));
            CUC(cudaMalloc(&b.d_raw, fb));
            CUC(cudaHostAlloc(&b.h_box, sizeof(double) * 6 * h->batch_frames, cudaHostAllocDefault));
            CUC(cudaMalloc(&b.d_box, sizeof(double) * 6 * h->batch_frames));
            CUC(cudaMalloc(&b.siteA, sizeof(Site) * (size_t)h->n1_pad * h->batch_frames));
            CUC(cudaMalloc(&b.siteB, sizeof(Site) * (size_t)h->n2_pad * h->batch_frames));
            CUC(cudaMalloc(&b.maxab, sizeof(float) * 4 * h->batch_frames));
            CUC(cudaEventCreateWithFlags(&b.uploaded, cudaEventDisableTiming));
            CUC(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming));
            CUC(cudaEventCreateWithFlags(&b.k0, cudaEventDefault));
            CUC(cudaEventCreateWithFlags(&b.k1, cudaEventDefault));
        }
    }
#undef CUC
    *out = h;
    return GFN_OK;
}

int gfn_create(gfn_t** out, int n1, int n2, int histo_n, float histo_min, float histo_max,
               double inv_dr, int mode_sel, const double box_dim[6], const int* devices, int ndev, unsigned flags)
{
    return create_impl(out, n1, n2, histo_n, histo_min, histo_max, inv_dr, mode_sel, box_dim, devices, ndev, flags, nullptr);
}

int gfn_create_shells(gfn_t** out, int n1, int n2, int histo_n, float histo_min, float histo_max,
                      double inv_dr, int mode_sel, const double box_dim[6], const int* devices, int ndev, unsigned flags,
                      const gfn_shell_cfg* cfg)
{
    if (!cfg) return fail(nullptr, GFN_EINVAL, "cfg is NULL");
    return create_impl(out, n1, n2, histo_n, histo_min, histo_max, inv_dr, mode_sel, box_dim, devices, ndev, flags, cfg);
}

int gfn_update_box(gfn_t* h, const double box_dim[6])
{
    if (!h || !box_dim) return fail(h, GFN_EINVAL, "NULL argument");
    int rc = validate_box(h, box_dim);
    if (rc) return rc;
    memcpy(h->box, box_dim, sizeof(h->box));
    return GFN_OK;
}

// Copies the segments of one frame into the ring; large frames are split between the caller and the helper thread.
static void copy_frame(gfn_t* h, const CopyHelper::Seg* segs, int nseg)
{
    size_t total = 0;
    for (int k = 0; k < nseg; ++k) total += segs[k].n;
    if (h->copy_threads < 2 || total < CopyHelper::kMinSplitBytes) {
        for (int k = 0; k < nseg; ++k) copy_to_ring(segs[k].d, segs[k].s, segs[k].n);
        return;
    }
    if (!h->helper) h->helper = new CopyHelper();
    h->helper->copy(segs, nseg);
}

// Frames go into the ring batch by batch.  Per batch: the caller's frames are copied into their pinned slots (one frame:
// the caller, split with the helper thread when large; several frames: the copy pool, one frame per task) - or, for
// page-locked sources (direct), only referenced - then the batch is submitted when it is full or the device is idle.
// Page-locked frames on a multi-device handle: nothing is copied on the host, so all the calling thread would do is issue
// ~14 CUDA calls per batch and device, one device after the other - with 8 devices that serial submission, not the GPUs,
// bounds the rate.  The call's frames are dealt to the devices exactly as the serial loop would (round-robin by batch, ramp
// after a sync) and each device's batches are then submitted by its own worker.
static int enqueue_direct_parallel(gfn_t* h, int nframes, const FrameSrc& src, bool aliased, const double* box_dims)
{
    struct Job { int f, k; };
    const int nd = (int)h->devs.size();
    std::vector<std::vector<Job>> jobs(nd);
    // batches that hold staged frames of an earlier call go first (serial: rare)
    for (Dev& d : h->devs) {
        Batch& b = d.ring[d.cur];
        if (b.nframes > 0) { int rc = submit(h, d); if (rc) return rc; }
    }
    if (box_dims)
        for (int f = 0; f < nframes; ++f) { int rc = validate_box(h, box_dims + 6ll * f); if (rc) return rc; }
    // every device gets the same number of frames of this call (+- 1, the extra ones starting at the device whose turn it
    // is): dealing whole batches round-robin until the frames run out leaves the first devices up to a batch ahead of the
    // last ones, and the call takes as long as the fullest device (8 devices, 512 frames, batches of 15: 74 vs 59 frames)
    std::vector<int> quota(nd, nframes / nd);
    for (int r = 0; r < nframes % nd; ++r) quota[(h->next_dev + r) % nd] += 1;
    for (int f = 0; f < nframes; ) {
        if (quota[h->next_dev] == 0) { h->next_dev = (h->next_dev + 1) % nd; continue; }
        Dev& d = h->devs[h->next_dev];
        int k = h->batch_frames;
        if (d.ramp < h->batch_frames && k > d.ramp) k = d.ramp;
        if (k > quota[h->next_dev]) k = quota[h->next_dev];
        jobs[h->next_dev].push_back(Job{f, k});
        if (d.ramp < h->batch_frames) d.ramp *= 2;
        f += k;
        quota[h->next_dev] -= k;
        h->next_dev = (h->next_dev + 1) % nd;
    }
    std::vector<int> rcs(nd, GFN_OK);
    if (!h->pool) h->pool = new CopyPool(h->pool_threads);
    h->pool->run(nd, [&](int di) {
        Dev& d = h->devs[di];
        for (const Job& j : jobs[di]) {
            Batch* b = &d.ring[d.cur];
            int rc = wait_batch(h, d, *b);
            if (rc) { rcs[di] = rc; return; }
            b->aliased = aliased; b->atoms = false; b->direct = true;
            b->dsrc = src;
            b->dsrc.c1 = src.c1 + src.s_c1 * j.f; b->dsrc.c2 = src.c2 + src.s_c2 * j.f;
            if (src.d1) b->dsrc.d1 = src.d1 + src.s_d1 * j.f;
            if (src.d2) b->dsrc.d2 = src.d2 + src.s_d2 * j.f;
            for (int i = 0; i < j.k; ++i)
                memcpy(b->h_box + 6ll * i, box_dims ? box_dims + 6ll * (j.f + i) : h->box, sizeof(double) * 6);
            b->nframes = j.k;
            rc = submit(h, d);
            if (rc) { rcs[di] = rc; return; }
        }
    });
    for (int rc : rcs) if (rc) return rc;
    return GFN_OK;
}

static int enqueue_loop(gfn_t* h, int nframes, const FrameSrc& src, bool aliased, const double* box_dims, bool direct = false)
{
    const bool atoms = src.a1 != nullptr;
    if (direct && !atoms && h->devs.size() > 1 && h->pool_threads > 0 && !getenv("GFN_SERIAL_SUBMIT"))
        return enqueue_direct_parallel(h, nframes, src, aliased, box_dims);
    auto device_idle = [](Dev& d) {
        bool idle = true;
        for (Batch& o : d.ring)
            if (o.inflight && cudaEventQuery(o.done) == cudaErrorNotReady) { idle = false; break; }
        cudaGetLastError();
        return idle;
    };
    // a block call on several devices gives every device the same number of frames (see enqueue_direct_parallel); frame by
    // frame calls keep filling one device's batch at a time
    const int nd = (int)h->devs.size();
    std::vector<int> quota;
    if (nd > 1 && nframes >= 2 * nd) {
        quota.assign(nd, nframes / nd);
        for (int r = 0; r < nframes % nd; ++r) quota[(h->next_dev + r) % nd] += 1;
    }
    int f = 0;
    while (f < nframes) {
        if (!quota.empty() && quota[h->next_dev] == 0) { h->next_dev = (h->next_dev + 1) % nd; continue; }
        const int di = h->next_dev;
        Dev& d = h->devs[di];
        Batch* b = &d.ring[d.cur];
        if (b->nframes > 0 && (b->aliased != aliased || b->atoms != atoms || b->direct || direct)) {
            // layout change, or a batch that references caller memory (it never spans two calls): close the batch
            int rc = submit(h, d);
            if (rc) return rc;
            b = &d.ring[d.cur];
        }
        if (b->nframes == 0) {
            int rc = wait_batch(h, d, *b);
            if (rc) return rc;
            b->aliased = aliased;
            b->atoms = atoms;
            b->direct = direct;
            if (direct) {
                b->dsrc = src;
                b->dsrc.c1 = src.c1 + src.s_c1 * f; b->dsrc.c2 = src.c2 + src.s_c2 * f;
                if (src.d1) b->dsrc.d1 = src.d1 + src.s_d1 * f;
                if (src.d2) b->dsrc.d2 = src.d2 + src.s_d2 * f;
            }
        }
        // frames of this call that go into the batch now; an idle device (first frames of a run, right after a readout)
        // gets a short batch instead of waiting for a full one
        int k = h->batch_frames - b->nframes;
        if (k > nframes - f) k = nframes - f;
        if (!quota.empty() && k > quota[di]) k = quota[di];
        const bool idle = device_idle(d);
        if (direct) {
            // page-locked frames cost the host nothing, but a batch runs only after ALL its frames have crossed PCIe: an idle
            // device (after a sync / readout) starts with two frames and the batches double while it works (2, 4, 8, ... batch_frames)
            if (d.ramp < h->batch_frames && b->nframes == 0 && k > d.ramp) k = d.ramp;
        }
        if (idle && !direct) {
            // staged frames: what the copy threads move in one go costs the idle device no more than a single frame would
            int first = h->idle_frames;
            if (h->pool_threads > 0 && first < h->pool_threads + 1) first = h->pool_threads + 1;
            if (b->nframes < first && k > first - b->nframes) k = first - b->nframes;
        }
        if (box_dims)
            for (int i = 0; i < k; ++i) {
                int rc = validate_box(h, box_dims + 6ll * (f + i));
                if (rc) return rc;
            }
        auto frame_segs = [&](int slot, int fr, CopyHelper::Seg* segs) -> int {
            int nseg = 0;
            auto seg = [&](void* dstp, const void* srcp, size_t n) {
                segs[nseg++] = {static_cast<unsigned char*>(dstp), static_cast<const unsigned char*>(srcp), n};
            };
            if (atoms) {
                const long long na1 = h->na[0], na2 = aliased ? 0 : (h->na[1] ? h->na[1] : h->na[0]);
                float* dst = b->h_atoms + 3 * (na1 + na2) * slot;
                seg(dst, src.a1 + src.s_a1 * fr, sizeof(float) * 3 * na1);
                if (!aliased) seg(dst + 3 * na1, src.a2 + src.s_a2 * fr, sizeof(float) * 3 * na2);
            } else {
                const long long fd = aliased ? h->off_c2 : h->frame_doubles;
                double* dst = b->h_pin + fd * slot;
                seg(dst + h->off_c1, src.c1 + src.s_c1 * fr, sizeof(double) * 3 * h->n1);
                if (h->need1) seg(dst + h->off_d1, src.d1 + src.s_d1 * fr, sizeof(double) * 3 * h->n1);
                if (!aliased) {
                    seg(dst + h->off_c2, src.c2 + src.s_c2 * fr, sizeof(double) * 3 * h->n2);
                    if (h->need2) seg(dst + h->off_d2, src.d2 + src.s_d2 * fr, sizeof(double) * 3 * h->n2);
                }
                if (src.shell_map) seg(dst + h->off_map, src.shell_map + src.s_map * fr, sizeof(int) * (size_t)h->shells.map_elems);
            }
            return nseg;
        };
        if (!direct) {
            if (k >= 2 && h->pool_threads > 0) {
                if (!h->pool) h->pool = new CopyPool(h->pool_threads);
                const int base = b->nframes;
                h->pool->run(k, [&](int i) {
                    CopyHelper::Seg segs[8];
                    const int nseg = frame_segs(base + i, f + i, segs);
                    for (int q = 0; q < nseg; ++q) copy_to_ring(segs[q].d, segs[q].s, segs[q].n);
                });
            } else {
                for (int i = 0; i < k; ++i) {
                    CopyHelper::Seg segs[8];
                    const int nseg = frame_segs(b->nframes + i, f + i, segs);
                    copy_frame(h, segs, nseg);
                }
            }
        }
        for (int i = 0; i < k; ++i)
            memcpy(b->h_box + 6ll * (b->nframes + i), box_dims ? box_dims + 6ll * (f + i) : h->box, sizeof(double) * 6);
        b->nframes += k;
        f += k;
        bool go = b->nframes == h->batch_frames;
        if (!go && b->nframes >= h->idle_frames) go = idle || device_idle(d);
        if (!go && direct) go = true;                            // a direct batch is exactly what this pass referenced (it never spans calls)
        if (!quota.empty()) quota[di] -= k;
        if (go) {
            int rc = submit(h, d);
            if (rc) return rc;
            if (direct && d.ramp < h->batch_frames) d.ramp *= 2;
            h->next_dev = (h->next_dev + 1) % nd;
        } else if (!quota.empty() && quota[di] == 0) {
            h->next_dev = (h->next_dev + 1) % nd;                // this device has its share; the open batch goes out with the next call or flush
        }
    }
    return GFN_OK;
}

static int enqueue_impl(gfn_t* h, int nframes,
                       const double* c1, long long s_c1, const double* c2, long long s_c2,
                       const double* d1, long long s_d1, const double* d2, long long s_d2, const double* box_dims,
                       const int* shell_map, long long s_map, bool direct = false)
{
    if (!h) return GFN_EINVAL;
    if (h->shells.layout && !shell_map) return fail(h, GFN_EINVAL, "this handle resolves shells: frames need a shell matrix (gfn_enqueue_frames_shells)");
    if (!h->shells.layout && shell_map) return fail(h, GFN_EINVAL, "this handle was not created with gfn_create_shells");
    if (nframes < 0) return fail(h, GFN_EINVAL, "nframes < 0");
    if (!c1 || !c2) return fail(h, GFN_EINVAL, "coordinate array is NULL");
    if (h->need1 && !d1) return fail(h, GFN_EINVAL, "No dipole moment array for the first selection was passed, although the requested gfunctions need it");
    if (h->need2 && !d2) return fail(h, GFN_EINVAL, "No dipole moment array for the second selection was passed, although the requested gfunctions need it");
    const bool aliased = !h->shells.layout && (c2 == c1) && h->n1 == h->n2 && (!h->need2 || !h->need1 || d2 == d1) && (!h->need2 || h->need1);
    FrameSrc src;
    src.c1 = c1; src.c2 = c2; src.d1 = d1; src.d2 = d2; src.s_c1 = s_c1; src.s_c2 = s_c2; src.s_d1 = s_d1; src.s_d2 = s_d2;
    src.shell_map = shell_map; src.s_map = s_map;
    if (direct) {
        // the DMA engine reads the caller's arrays: every part must be page-locked, frames must not overlap
        if (h->shells.layout) return fail(h, GFN_EINVAL, "shell-resolved handles take staged frames (gfn_enqueue_frames_shells)");
        const void* parts[4] = {c1, h->need1 ? d1 : nullptr, aliased ? nullptr : c2, (!aliased && h->need2) ? d2 : nullptr};
        const long long strides[4] = {s_c1, s_d1, s_c2, s_d2};
        const long long widths[4] = {3ll * h->n1, 3ll * h->n1, 3ll * h->n2, 3ll * h->n2};
        for (int k = 0; k < 4; ++k) {
            if (!parts[k]) continue;
            if (nframes > 1 && strides[k] < widths[k]) return fail(h, GFN_EINVAL, "frame stride of part %d is smaller than a frame", k);
            cudaPointerAttributes at;
            const cudaError_t e = cudaPointerGetAttributes(&at, parts[k]);
            if (e != cudaSuccess || at.type != cudaMemoryTypeHost) {
                cudaGetLastError();
                return fail(h, GFN_EINVAL, "gfn_enqueue_frames_pinned: part %d is not page-locked host memory (gfn_host_alloc / cudaHostRegister)", k);
            }
        }
    }
    return enqueue_loop(h, nframes, src, aliased, box_dims, direct);
}

// ---- N1 + N4: atoms through the ring ---------------------------------------------------------------------
int gfn_attach_topology(gfn_t* h, int which, int natoms, int nres, const int* apr, const int* rfa,
                        const double* masses, const double* charges)
{
    if (!h) return GFN_EINVAL;
    if (which < 0 || which > 1) return fail(h, GFN_EINVAL, "which must be 0 (selection 1) or 1 (selection 2)");
    if (h->shells.layout) return fail(h, GFN_EINVAL, "shell-resolved handles take host frames (gfn_enqueue_frames_shells)");
    if (which == 1 && !h->na[0]) return fail(h, GFN_EINVAL, "attach the topology of selection 1 first");
    if (nres != (which == 0 ? h->n1 : h->n2))
        return fail(h, GFN_EINVAL, "topology has %d residues, selection %d of this handle has %d sites", nres, which + 1, which == 0 ? h->n1 : h->n2);
    int rc = gfn_flush(h);                       // batches filled so far were sized with the previous topology
    if (rc) return rc;
    for (Dev& d : h->devs) {
        gfn_residues_t* r = nullptr;
        rc = gfn_residues_create(&r, d.dev, natoms, nres, apr, rfa, masses, charges);
        if (rc) return fail(h, rc, "%s", gfn_last_error(nullptr));
        gfn_residues_destroy(d.res[which]);
        d.res[which] = r;
    }
    h->na[which] = natoms;
    const long long cap = 3ll * h->na[0] + 3ll * (h->na[1] ? h->na[1] : h->na[0]);
    if (cap > h->atom_cap) {
        for (Dev& d : h->devs) {
            CU(cudaSetDevice(d.dev));
            for (Batch& b : d.ring) {
                int wrc = wait_batch(h, d, b);
                if (wrc) return wrc;
                if (b.h_atoms) { CU(cudaFreeHost(b.h_atoms)); b.h_atoms = nullptr; }
                if (b.d_atoms) { CU(cudaFree(b.d_atoms)); b.d_atoms = nullptr; }
                const size_t bytes = sizeof(float) * (size_t)cap * h->batch_frames;
                CU(cudaHostAlloc(&b.h_atoms, bytes, cudaHostAllocDefault));
                CU(cudaMalloc(&b.d_atoms, bytes));
            }
        }
        h->atom_cap = cap;
    }
    return GFN_OK;
}

int gfn_topology_natoms(gfn_t* h, int* natoms1, int* natoms2)
{
    if (!h) return GFN_EINVAL;
    if (natoms1) *natoms1 = h->na[0];
    if (natoms2) *natoms2 = h->na[1] ? h->na[1] : h->na[0];
    return GFN_OK;
}

int gfn_enqueue_atom_frames(gfn_t* h, int nframes, const float* a1, long long stride_a1,
                            const float* a2, long long stride_a2, const double* box_dims)
{
    Range nvtx_range("gfn_enqueue_atom_frames");
    if (!h) return GFN_EINVAL;
    if (!h->na[0]) return fail(h, GFN_EINVAL, "no topology attached (gfn_attach_topology)");
    if (nframes < 0) return fail(h, GFN_EINVAL, "nframes < 0");
    if (!a1) return fail(h, GFN_EINVAL, "atom position array is NULL");
    if (!a2) {
        if (h->na[1] || h->n1 != h->n2) return fail(h, GFN_EINVAL, "atom positions of the second selection are NULL");
        a2 = a1; stride_a2 = stride_a1;
    }
    {
        const Dev& d0 = h->devs[0];
        if (h->need1 && !d0.res[0]->charges)
            return fail(h, GFN_EINVAL, "No charges for the first selection were attached, although the requested gfunctions need its dipole moments");
        if (h->need2 && !(d0.res[1] ? d0.res[1] : d0.res[0])->charges)
            return fail(h, GFN_EINVAL, "No charges for the second selection were attached, although the requested gfunctions need its dipole moments");
    }
    const bool aliased = (a2 == a1) && !h->na[1] && h->n1 == h->n2 && (!h->need2 || h->need1);
    FrameSrc src;
    src.a1 = a1; src.a2 = a2; src.s_a1 = stride_a1; src.s_a2 = stride_a2;
    return enqueue_loop(h, nframes, src, aliased, box_dims);
}

int gfn_enqueue_frames(gfn_t* h, int nframes,
                       const double* c1, long long s_c1, const double* c2, long long s_c2,
                       const double* d1, long long s_d1, const double* d2, long long s_d2, const double* box_dims)
{
    Range nvtx_range("gfn_enqueue_frames");
    return enqueue_impl(h, nframes, c1, s_c1, c2, s_c2, d1, s_d1, d2, s_d2, box_dims, nullptr, 0);
}

int gfn_enqueue_frames_pinned(gfn_t* h, int nframes,
                              const double* c1, long long s_c1, const double* c2, long long s_c2,
                              const double* d1, long long s_d1, const double* d2, long long s_d2, const double* box_dims)
{
    Range nvtx_range("gfn_enqueue_frames_pinned");
    return enqueue_impl(h, nframes, c1, s_c1, c2, s_c2, d1, s_d1, d2, s_d2, box_dims, nullptr, 0, true);
}

int gfn_host_alloc(void** out, unsigned long long bytes)
{
    if (!out) return fail(nullptr, GFN_EINVAL, "NULL argument");
    *out = nullptr;
    const cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, e == cudaErrorMemoryAllocation ? GFN_ENOMEM : GFN_ECUDA, "cudaHostAlloc(%llu bytes): %s", bytes, cudaGetErrorString(e)); }
    return GFN_OK;
}

int gfn_host_free(void* p)
{
    if (!p) return GFN_OK;
    const cudaError_t e = cudaFreeHost(p);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, GFN_ECUDA, "cudaFreeHost: %s", cudaGetErrorString(e)); }
    return GFN_OK;
}

int gfn_enqueue_frames_shells(gfn_t* h, int nframes,
                              const double* c1, long long s_c1, const double* c2, long long s_c2,
                              const double* d1, long long s_d1, const double* d2, long long s_d2,
                              const double* box_dims, const int* shell_map, long long s_map)
{
    Range nvtx_range("gfn_enqueue_frames_shells");
    if (h && !shell_map) return fail(h, GFN_EINVAL, "shell matrix is NULL");
    return enqueue_impl(h, nframes, c1, s_c1, c2, s_c2, d1, s_d1, d2, s_d2, box_dims, shell_map, s_map);
}

int gfn_enqueue_frame(gfn_t* h, const double* c1, const double* c2, const double* d1, const double* d2)
{
    return gfn_enqueue_frames(h, 1, c1, 0, c2, 0, d1, 0, d2, 0, nullptr);
}

int gfn_enqueue_device_frames(gfn_t* h, int nframes,
                              const double* dc1, long long s_c1, const double* dc2, long long s_c2,
                              const double* dd1, long long s_d1, const double* dd2, long long s_d2, const double* box_dims)
{
    Range nvtx_range("gfn_enqueue_device_frames");
    if (!h) return GFN_EINVAL;
    if (nframes < 0) return fail(h, GFN_EINVAL, "nframes < 0");
    if (!dc1 || !dc2) return fail(h, GFN_EINVAL, "coordinate array is NULL");
    if (h->need1 && !dd1) return fail(h, GFN_EINVAL, "dipoles of the first selection are required by mode_sel");
    if (h->need2 && !dd2) return fail(h, GFN_EINVAL, "dipoles of the second selection are required by mode_sel");
    if (h->shells.layout) return fail(h, GFN_EINVAL, "shell-resolved handles take host frames (gfn_enqueue_frames_shells)");
    // the frames run on the device that holds them: a multi-device handle takes device-resident frames on each of its GPUs
    Dev* dp = &h->devs[0];
    {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, dc1) != cudaSuccess || at.type != cudaMemoryTypeDevice) {
            cudaGetLastError();
            return fail(h, GFN_EINVAL, "dc1 is not a device pointer");
        }
        dp = nullptr;
        for (Dev& cand : h->devs) if (cand.dev == at.device) dp = &cand;
        if (!dp) return fail(h, GFN_EINVAL, "the frames live on device %d, which is not one of this handle's devices", at.device);
        const double* others[3] = {dc2, dd1, dd2};
        for (const double* q : others) {
            if (!q) continue;
            cudaPointerAttributes a2;
            if (cudaPointerGetAttributes(&a2, q) != cudaSuccess || a2.type != cudaMemoryTypeDevice || a2.device != at.device) {
                cudaGetLastError();
                return fail(h, GFN_EINVAL, "all arrays of a device-resident frame must live on the same device (%d)", at.device);
            }
        }
    }
    Dev& d = *dp;
    CU(cudaSetDevice(d.dev));
    const bool aliased = (dc2 == dc1) && h->n1 == h->n2 && (!h->need2 || !h->need1 || dd2 == dd1) && (!h->need2 || h->need1);
    const int chunk = h->batch_frames;
    if (d.xcap < chunk) {
        CU(cudaMalloc(&d.xsiteA, sizeof(Site) * (size_t)h->n1_pad * chunk));
        CU(cudaMalloc(&d.xsiteB, sizeof(Site) * (size_t)h->n2_pad * chunk));
        CU(cudaMalloc(&d.xmax, sizeof(float) * 4 * chunk));
        d.xcap = chunk;
    }
    for (int f0 = 0; f0 < nframes; f0 += chunk) {
        const int nf = nframes - f0 < chunk ? nframes - f0 : chunk;
        Dev::XSlot& x = d.xs[d.xcur];
        d.xcur = (d.xcur + 1) % kXRing;
        if (x.timed) {          // slot reuse: its previous launch is kXRing chunks old
            CU(cudaEventSynchronize(x.k1));
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, x.k0, x.k1) == cudaSuccess) d.st.kernel_ms += ms;
            x.timed = false;
        }
        for (int f = 0; f < nf; ++f) {
            const double* src = box_dims ? box_dims + 6ll * (f0 + f) : h->box;
            if (box_dims) { int rc = validate_box(h, src); if (rc) return rc; }
            memcpy(x.h_box + 6ll * f, src, sizeof(double) * 6);
        }
        // record scratch is reused chunk after chunk: stream order on d.comp serialises it
        CU(cudaMemcpyAsync(x.d_box, x.h_box, sizeof(double) * 6 * nf, cudaMemcpyHostToDevice, d.comp));
        int rc = launch_frames(h, d, nf,
                               dc1 + s_c1 * f0, s_c1, dd1 ? dd1 + s_d1 * f0 : nullptr, s_d1,
                               dc2 + s_c2 * f0, s_c2, dd2 ? dd2 + s_d2 * f0 : nullptr, s_d2, aliased,
                               d.xsiteA, d.xsiteB, d.xmax, x.d_box, x.k0, x.k1);
        if (rc) return rc;
        x.timed = true;
        reader_fence().add(d.dev, x.k0);          // recorded right behind prep_kernel, the only reader of the caller's arrays
    }
    return GFN_OK;
}

int gfn_flush(gfn_t* h)
{
    if (!h) return GFN_EINVAL;
    for (Dev& d : h->devs) {
        int rc = submit(h, d);
        if (rc) return rc;
    }
    return GFN_OK;
}

int gfn_sync(gfn_t* h)
{
    Range nvtx_range("gfn_sync");
    if (!h) return GFN_EINVAL;
    int rc = gfn_flush(h);
    if (rc) return rc;
    for (Dev& d : h->devs) {
        CU(cudaSetDevice(d.dev));
        CU(cudaStreamSynchronize(d.copy));
        CU(cudaStreamSynchronize(d.comp));
        for (Batch& b : d.ring) { b.inflight = false; harvest_timing(d, b); }
        d.ramp = 2;
        for (Dev::XSlot& x : d.xs) if (x.timed) { float ms = 0.f; if (cudaEventElapsedTime(&ms, x.k0, x.k1) == cudaSuccess) d.st.kernel_ms += ms; x.timed = false; }
    }
    return check_zerodiv(h);
}

// Everything of a readout is enqueued behind the last frame on each device's compute stream - status words, the
// all-reduce, the copy into pinned memory - and the host waits ONCE per device (the collective is fused
// into the tail of the last frame).  The two status words travel through the same all-reduce as the histogram, so every
// rank of a multi-process job learns about an error any rank hit and all of them return it: no rank is left waiting
// in a collective the failing rank never entered.
int gfn_readout(gfn_t* h, double* hist, double* pairw, unsigned long long* overflow)
{
    Range nvtx_range("gfn_readout");
    if (!h || !hist) return fail(h, GFN_EINVAL, "NULL argument");
    int rc = gfn_flush(h);
    if (rc) return rc;
    const size_t n8 = 8ull * h->histo_n, nall = n8 + kStatusSlots;
    // one process, several devices, no external communicator: every device copies its 19 KB into pinned memory and the host
    // adds them in device order (deterministic; no NCCL needed, and an in-process group all-reduce over 8 communicators
    // costs more launch latency than the eight copies)
    const bool host_sum = h->devs.size() > 1 && !h->ext_comm;
    const bool collective = (h->devs.size() > 1 || h->ext_comm) && !host_sum;
    if (collective && (rc = setup_comm(h))) return rc;
    for (Dev& d : h->devs) {
        CU(cudaSetDevice(d.dev));
        CU(status_launch(d.counters, d.gacc + n8, d.comp));
    }
    h->st_launches += (double)h->devs.size();
    if (collective) {
        Nccl& n = nccl();
        if (h->devs.size() > 1) n.GroupStart();
        for (Dev& d : h->devs) {
            CU(cudaSetDevice(d.dev));
            int nr = n.AllReduce(d.gacc, d.gsum, nall, ncclFloat64, ncclSum, d.comm, d.comp);
            if (nr != ncclSuccess) { if (h->devs.size() > 1) n.GroupEnd(); return fail(h, GFN_ENCCL, "ncclAllReduce: %s", n.GetErrorString(nr)); }
        }
        if (h->devs.size() > 1) { int nr = n.GroupEnd(); if (nr != ncclSuccess) return fail(h, GFN_ENCCL, "ncclGroupEnd: %s", n.GetErrorString(nr)); }
        h->st_launches += (double)h->devs.size();
    }
    for (size_t i = 0; i < h->devs.size(); ++i) {
        Dev& d = h->devs[i];
        CU(cudaSetDevice(d.dev));
        if (i == 0 || host_sum) CU(cudaMemcpyAsync(d.h_out, collective ? d.gsum : d.gacc, sizeof(double) * nall, cudaMemcpyDeviceToHost, d.comp));
        CU(cudaMemcpyAsync(d.h_out + nall, d.counters, sizeof(unsigned long long) * kCounters, cudaMemcpyDeviceToHost, d.comp));
    }
    unsigned long long ovf = 0;
    for (Dev& d : h->devs) {
        CU(cudaSetDevice(d.dev));
        CU(cudaStreamSynchronize(d.comp));       // the one wait: frames, all-reduce and copies are all behind it
        for (Batch& b : d.ring) { b.inflight = false; harvest_timing(d, b); }
        d.ramp = 2;
        for (Dev::XSlot& x : d.xs) if (x.timed) { float ms = 0.f; if (cudaEventElapsedTime(&ms, x.k0, x.k1) == cudaSuccess) d.st.kernel_ms += ms; x.timed = false; }
        ovf += reinterpret_cast<const unsigned long long*>(d.h_out + nall)[0];
        if (reinterpret_cast<const unsigned long long*>(d.h_out + nall)[1] & 8ull)
            return fail(h, GFN_ECUDA, "checked build: a device-side index / protocol check failed");
        if (reinterpret_cast<const unsigned long long*>(d.h_out + nall)[1] & 4ull)
            return fail(h, GFN_ECUDA, "a bin of one CTA took more than 2^20 pairs in one launch: beyond the dense kernel's fixed-point range (create the handle with GFN_FLAG_NO_DENSE)");
        if (reinterpret_cast<const unsigned long long*>(d.h_out + nall)[1] & 16ull)
            return fail(h, GFN_ECUDA, "internal: the visit list of a cell-sorted launch overflowed its buffer");
    }
    if (host_sum) {
        double* acc = h->devs[0].h_out;
        for (size_t i = 1; i < h->devs.size(); ++i) {
            const double* src = h->devs[i].h_out;
            for (size_t k = 0; k < nall; ++k) acc[k] += src[k];
        }
    }
    const double* out = h->devs[0].h_out;
    if (out[n8] != 0.0)
        return fail(h, GFN_EINVAL, "coordinates lie more than 30000 box lengths from the origin: the fp16 pre-filter cannot represent them, create the handle without GFN_FLAG_FILTER_FP16%s",
                    collective ? " (reported by at least one rank / device)" : "");
    memcpy(hist, out, sizeof(double) * 7 * h->histo_n);
    if (pairw) memcpy(pairw, out + 7ull * h->histo_n, sizeof(double) * h->histo_n);
    if (overflow) *overflow = ovf;      // local devices only; ranks of an external communicator report their own
    if (out[n8 + 1] != 0.0) return fail(h, GFN_EZERODIV, "float division");
    return GFN_OK;
}

int gfn_readout_per_particle(gfn_t* h, double* hist)
{
    Range nvtx_range("gfn_readout_per_particle");
    if (!h || !hist) return fail(h, GFN_EINVAL, "NULL argument");
    if (!(h->flags & GFN_FLAG_PER_PARTICLE)) return fail(h, GFN_EINVAL, "handle was not created with GFN_FLAG_PER_PARTICLE");
    int rc = gfn_sync(h);
    if (rc && rc != GFN_EZERODIV) return rc;
    Dev& d = h->devs[0];
    CU(cudaSetDevice(d.dev));
    CU(cudaMemcpy(hist, d.gpp, sizeof(double) * 7ull * h->n1 * h->histo_n, cudaMemcpyDeviceToHost));
    return rc;
}

int gfn_readout_per_particle_add(gfn_t* h, double* dst, unsigned long long first, unsigned long long count)
{
    Range nvtx_range("gfn_readout_per_particle_add");
    if (!h || !dst) return fail(h, GFN_EINVAL, "NULL argument");
    if (!(h->flags & GFN_FLAG_PER_PARTICLE)) return fail(h, GFN_EINVAL, "handle was not created with GFN_FLAG_PER_PARTICLE");
    const unsigned long long total = 7ull * h->n1 * h->histo_n;
    if (first > total || count > total - first) return fail(h, GFN_EINVAL, "slice [%llu, +%llu) exceeds the per-particle array of %llu values", first, count, total);
    int rc = gfn_flush(h);
    if (rc) return rc;
    Dev& d = h->devs[0];
    CU(cudaSetDevice(d.dev));
    if (count > d.h_pp_cap) {
        if (d.h_pp) { CU(cudaFreeHost(d.h_pp)); d.h_pp = nullptr; d.h_pp_cap = 0; }
        CU(cudaHostAlloc(&d.h_pp, sizeof(double) * count, cudaHostAllocDefault));
        d.h_pp_cap = count;
    }
    // enqueued behind the last frame; gfn_sync below is the one wait (and reports the deferred errors)
    if (count) CU(cudaMemcpyAsync(d.h_pp, d.gpp + first, sizeof(double) * count, cudaMemcpyDeviceToHost, d.comp));
    rc = gfn_sync(h);
    if (rc && rc != GFN_EZERODIV) return rc;
    const double* src = d.h_pp;
    for (unsigned long long k = 0; k < count; ++k) dst[k] += src[k];
    return rc;
}

int gfn_scale(gfn_t* h, double value)
{
    if (!h) return GFN_EINVAL;
    int rc = gfn_flush(h);
    if (rc) return rc;
    for (Dev& d : h->devs) {
        CU(cudaSetDevice(d.dev));
        CU(scale_launch(d.gacc, 8ll * h->histo_n, value, d.comp));
        if (d.gpp) CU(scale_launch(d.gpp, 7ll * h->n1 * h->histo_n, value, d.comp));
        h->st_launches += d.gpp ? 2 : 1;
    }
    return GFN_OK;
}

int gfn_reset(gfn_t* h)
{
    if (!h) return GFN_EINVAL;
    int rc = gfn_flush(h);
    if (rc) return rc;
    for (Dev& d : h->devs) {
        CU(cudaSetDevice(d.dev));
        CU(cudaMemsetAsync(d.gacc, 0, sizeof(double) * 8 * h->histo_n, d.comp));
        if (d.gpp) CU(cudaMemsetAsync(d.gpp, 0, sizeof(double) * 7ull * h->n1 * h->histo_n, d.comp));
        CU(cudaMemsetAsync(d.counters, 0, sizeof(unsigned long long) * kCounters, d.comp));
    }
    return GFN_OK;
}

void gfn_destroy(gfn_t* h)
{
    if (!h) return;
    delete h->helper;
    h->helper = nullptr;
    delete h->pool;
    h->pool = nullptr;
    for (Dev& d : h->devs) free_dev(d);
    delete h;
}

int gfn_stats(gfn_t* h, gfn_stats_t* out)
{
    if (!h || !out) return fail(h, GFN_EINVAL, "NULL argument");
    memset(out, 0, sizeof(*out));
    out->pair_evals = h->st_pairs; out->frames = h->st_frames; out->h2d_bytes = h->st_h2d;
    out->kernel_launches = h->st_launches; out->pair_launches = h->st_pair_launches; out->kernel_ms = h->st_kernel_ms;
    double visits_dense = h->st_visits_dense;
    for (Dev& d : h->devs) {
        out->pair_evals += d.st.pairs; out->frames += d.st.frames; out->h2d_bytes += d.st.h2d; out->kernel_launches += d.st.launches;
        out->pair_launches += d.st.pair_launches; out->kernel_ms += d.st.kernel_ms; visits_dense += d.st.visits_dense;
    }
    for (Dev& d : h->devs) {
        unsigned long long c[kCounters];
        CU(cudaSetDevice(d.dev));
        CU(cudaMemcpy(c, d.counters, sizeof(c), cudaMemcpyDeviceToHost));
        out->overflow += (double)c[0]; out->survivors += (double)c[2]; out->in_range += (double)c[3];
        out->selfcheck_bad += (double)c[4]; out->exact_evals += (double)c[5]; out->tile_visits += (double)c[6];
    }
    out->frames_min_device = h->devs.empty() ? 0.0 : h->devs[0].st.frames;
    for (Dev& d : h->devs) {
        if (d.st.frames > out->frames_max_device) out->frames_max_device = d.st.frames;
        if (d.st.frames < out->frames_min_device) out->frames_min_device = d.st.frames;
        if (d.st.kernel_ms > out->kernel_ms_max_device) out->kernel_ms_max_device = d.st.kernel_ms;
    }
    const PairConfig& c = h->devs[0].cfg;
    out->filter_fp64 = c.filter_fp64; out->hist_mode = c.hist_mode; out->warps_per_block = c.warps; out->hist_replicas = c.nrep;
    out->blocks_per_sm = c.blocks_per_sm; out->grid = c.grid; out->smem_bytes = c.smem_bytes;
    out->batch_frames = h->batch_frames; out->ndev = (int)h->devs.size(); out->dense = c.dense;
    out->cells = h->cells ? 1 : 0; out->tile_visits_dense = visits_dense;
    return GFN_OK;
}

int gfn_comm_unique_id(char out[128])
{
    Nccl& n = nccl();
    if (!n.ok) return fail(nullptr, GFN_ENCCL, "%s", n.err.c_str());
    ncclUniqueId id;
    int rc = n.GetUniqueId(&id);
    if (rc != ncclSuccess) return fail(nullptr, GFN_ENCCL, "ncclGetUniqueId: %s", n.GetErrorString(rc));
    memcpy(out, id.internal, 128);
    return GFN_OK;
}

int gfn_comm_attach(gfn_t* h, const char unique_id[128], int nranks, int rank)
{
    if (!h || !unique_id) return fail(h, GFN_EINVAL, "NULL argument");
    if (h->devs.size() != 1) return fail(h, GFN_EINVAL, "gfn_comm_attach needs a single-device handle (one process per GPU)");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(h, GFN_EINVAL, "bad rank %d of %d", rank, nranks);
    Nccl& n = nccl();
    if (!n.ok) return fail(h, GFN_ENCCL, "%s", n.err.c_str());
    ncclUniqueId id;
    memcpy(id.internal, unique_id, 128);
    Dev& d = h->devs[0];
    CU(cudaSetDevice(d.dev));
    int rc = n.CommInitRank(&d.comm, nranks, id, rank);
    if (rc != ncclSuccess) return fail(h, GFN_ENCCL, "ncclCommInitRank: %s", n.GetErrorString(rc));
    h->nranks = nranks; h->rank = rank; h->ext_comm = true; h->comm_ready = true;
    return GFN_OK;
}

int gfn_allreduce_host(gfn_t* h, double* vals, int n)
{
    if (!h || !vals || n < 0) return fail(h, GFN_EINVAL, "bad argument");
    if (!h->ext_comm || h->nranks == 1 || n == 0) return GFN_OK;      // the devices of one process share these host values
    if (n > 8 * h->histo_n) return fail(h, GFN_EINVAL, "at most %d values", 8 * h->histo_n);
    Nccl& nc = nccl();
    if (!nc.ok) return fail(h, GFN_ENCCL, "%s", nc.err.c_str());
    Dev& d = h->devs[0];
    CU(cudaSetDevice(d.dev));
    // gsum is free between readouts (gfn_readout waits for its all-reduce before it returns)
    CU(cudaMemcpyAsync(d.gsum, vals, sizeof(double) * n, cudaMemcpyHostToDevice, d.comp));
    const int nr = nc.AllReduce(d.gsum, d.gsum, (size_t)n, ncclFloat64, ncclSum, d.comm, d.comp);
    if (nr != ncclSuccess) return fail(h, GFN_ENCCL, "ncclAllReduce: %s", nc.GetErrorString(nr));
    CU(cudaMemcpyAsync(vals, d.gsum, sizeof(double) * n, cudaMemcpyDeviceToHost, d.comp));
    CU(cudaStreamSynchronize(d.comp));
    h->st_launches += 1;
    return GFN_OK;
}

int gfn_mark(gfn_t* h, int which)
{
    if (!h || which < 0 || which > 1) return fail(h, GFN_EINVAL, "gfn_mark: which must be 0 or 1");
    int rc = gfn_flush(h);
    if (rc) return rc;
    Dev& d = h->devs[0];
    CU(cudaSetDevice(d.dev));
    CU(cudaEventRecord(d.mark[which], d.comp));
    return GFN_OK;
}

int gfn_mark_elapsed_ms(gfn_t* h, double* ms)
{
    if (!h || !ms) return fail(h, GFN_EINVAL, "NULL argument");
    Dev& d = h->devs[0];
    CU(cudaSetDevice(d.dev));
    CU(cudaEventSynchronize(d.mark[1]));
    float f = 0.f;
    CU(cudaEventElapsedTime(&f, d.mark[0], d.mark[1]));
    *ms = f;
    return GFN_OK;
}

int gfn_measure_peaks(int dev, double* fp64_tflops, double* fp32_tflops)
{
    gfn_t* h = nullptr;
    CU(peaks_measure(dev, fp64_tflops, fp32_tflops));
    return GFN_OK;
}

// ---- N1: residue map (topology on the device) ----------------------------------------------------------
int gfn_residues_create(gfn_residues_t** out, int dev, int natoms, int nres, const int* apr, const int* rfa,
                        const double* masses, const double* charges)
{
    gfn_t* h = nullptr;
    if (!out || !apr || !rfa || !masses) return fail(h, GFN_EINVAL, "NULL argument");
    *out = nullptr;
    if (natoms < 1 || nres < 1) return fail(h, GFN_EINVAL, "natoms and nres must be >= 1");
    for (int i = 0; i < nres; ++i)
        if (apr[i] < 0 || rfa[i] < 0 || (long long)rfa[i] + apr[i] > natoms)
            return fail(h, GFN_EINVAL, "residue %d (first atom %d, %d atoms) exceeds natoms = %d", i, rfa[i], apr[i], natoms);
    if (gfn_device_count() < 1) return fail(h, GFN_ENODEV, "no CUDA device available; libgfn has no CPU fallback");
    CU(cudaSetDevice(dev));
    gfn_residues* r = new gfn_residues();
    r->dev = dev; r->natoms = natoms; r->nres = nres; r->err[0] = 0;
#define CUR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(nullptr, GFN_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); gfn_residues_destroy(r); return GFN_ECUDA; } } while (0)
    CUR(cudaMalloc(&r->apr, sizeof(int) * nres));
    CUR(cudaMalloc(&r->rfa, sizeof(int) * nres));
    CUR(cudaMalloc(&r->masses, sizeof(double) * natoms));
    CUR(cudaMemcpy(r->apr, apr, sizeof(int) * nres, cudaMemcpyHostToDevice));
    CUR(cudaMemcpy(r->rfa, rfa, sizeof(int) * nres, cudaMemcpyHostToDevice));
    CUR(cudaMemcpy(r->masses, masses, sizeof(double) * natoms, cudaMemcpyHostToDevice));
    if (charges) {
        CUR(cudaMalloc(&r->charges, sizeof(double) * natoms));
        CUR(cudaMemcpy(r->charges, charges, sizeof(double) * natoms, cudaMemcpyHostToDevice));
    }
#undef CUR
    *out = r;
    return GFN_OK;
}

void gfn_residues_destroy(gfn_residues_t* r)
{
    if (!r) return;
    cudaSetDevice(r->dev);
    cudaFree(r->apr); cudaFree(r->rfa); cudaFree(r->masses); cudaFree(r->charges); cudaFree(r->scratch);
    delete r;
}

// device pointers in, device pointers out (chained into gfn_enqueue_device_frames without touching the host)
int gfn_residues_com_dip_device(gfn_residues_t* r, int nframes, const double* dcoor, double* dcom, double* ddip)
{
    gfn_t* h = nullptr;
    if (!r || !dcoor || !dcom) return fail(h, GFN_EINVAL, "NULL argument");
    if (ddip && !r->charges) return fail(h, GFN_EINVAL, "dipoles requested but the residue map was created without charges");
    CU(cudaSetDevice(r->dev));
    CU(residue_com_launch(dcoor, 3ll * r->natoms, r->masses, r->apr, r->rfa, r->nres, nframes, dcom, 0));
    if (ddip) CU(residue_dip_launch(dcoor, 3ll * r->natoms, r->charges, r->apr, r->rfa, r->nres, nframes, dcom, ddip, 0));
    CU(cudaStreamSynchronize(0));     // the RDF handles consume these arrays on their own (non-blocking) streams
    return GFN_OK;
}

// host arrays in, host arrays out: comByResidue (com_in == NULL: com is computed and returned in com_out) and
// dipByResidue (dip_out != NULL; com_in, if given, is the reference point array the caller passes, helpers.pyx:101)
int gfn_residues_com_dip(gfn_residues_t* r, int nframes, const double* coor, const double* com_in, double* com_out, double* dip_out)
{
    gfn_t* h = nullptr;
    if (!r || !coor || nframes < 1) return fail(h, GFN_EINVAL, "bad argument");
    if (dip_out && !r->charges) return fail(h, GFN_EINVAL, "dipoles requested but the residue map was created without charges");
    CU(cudaSetDevice(r->dev));
    const size_t cb = sizeof(double) * 3 * (size_t)r->natoms * nframes, ob = sizeof(double) * 3 * (size_t)r->nres * nframes;
    if (r->scratch_bytes < cb + 2 * ob) {
        cudaFree(r->scratch); r->scratch = nullptr; r->scratch_bytes = 0;
        CU(cudaMalloc(&r->scratch, cb + 2 * ob));
        r->scratch_bytes = cb + 2 * ob;
    }
    double* dcoor = r->scratch; double* dcom = dcoor + 3ll * r->natoms * nframes; double* ddip = dcom + 3ll * r->nres * nframes;
    CU(cudaMemcpy(dcoor, coor, cb, cudaMemcpyHostToDevice));
    if (com_in) CU(cudaMemcpy(dcom, com_in, ob, cudaMemcpyHostToDevice));
    else CU(residue_com_launch(dcoor, 3ll * r->natoms, r->masses, r->apr, r->rfa, r->nres, nframes, dcom, 0));
    if (dip_out) CU(residue_dip_launch(dcoor, 3ll * r->natoms, r->charges, r->apr, r->rfa, r->nres, nframes, dcom, ddip, 0));
    if (com_out) CU(cudaMemcpy(com_out, dcom, ob, cudaMemcpyDeviceToHost));
    if (dip_out) CU(cudaMemcpy(dip_out, ddip, ob, cudaMemcpyDeviceToHost));
    return GFN_OK;
}

int gfn_selftest_arith(int dev, unsigned long long seed, long long n_operands, unsigned long long mismatches[3])
{
    gfn_t* h = nullptr;
    if (!mismatches) return fail(h, GFN_EINVAL, "NULL argument");
    CU(arith_selftest(dev, seed, n_operands, mismatches));
    return GFN_OK;
}

int gfn_dev_alloc(int dev, void** dptr, unsigned long long bytes)
{
    gfn_t* h = nullptr;
    CU(cudaSetDevice(dev));
    CU(cudaMalloc(dptr, bytes));
    return GFN_OK;
}

int gfn_dev_free(int dev, void* dptr)
{
    gfn_t* h = nullptr;
    CU(cudaSetDevice(dev));
    CU(cudaFree(dptr));
    return GFN_OK;
}

int gfn_dev_upload(int dev, void* dptr, const void* host, unsigned long long bytes)
{
    gfn_t* h = nullptr;
    CU(cudaSetDevice(dev));
    CU(reader_fence().wait(dev));     // frames enqueued earlier may still be reading this buffer: wait for their prep kernels only
    CU(cudaMemcpy(dptr, host, bytes, cudaMemcpyHostToDevice));
    return GFN_OK;
}

}  // extern "C"

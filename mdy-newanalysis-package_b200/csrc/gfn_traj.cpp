// gfn_traj.cpp - N4: the trajectory side of the upload ring.
//
// The reference's analysis scripts read CHARMM/NAMD DCD trajectories through MDAnalysis and, per frame, hand
// float32 positions to helpers.comByResidue / dipByResidue and then to RDF.calcFrame
// (/root/reference/docs/notebooks/rdf.ipynb cells 10-18: `u = MDAnalysis.Universe(psf, "...dcd")`,
// `for ts in u.trajectory[::skip]: ... sel.positions ... sol_wat.calcFrame(...)`).  MDAnalysis is a third-party
// dependency that is not part of the reference tree; what is restated here is the published DCD record layout
// (the one MDAnalysis' libdcd and VMD's dcdplugin read):
//
//   record 1   int32 84 | "CORD" | icntrl[20] | int32 84      icntrl[0] frames, [1] first step, [2] step interval,
//                                                             [8] fixed atoms, [9] time step, [10] unit cell present,
//                                                             [11] fourth dimension present, [19] CHARMM version
//   record 2   int32 n  | int32 ntitle | ntitle*80 chars | int32 n
//   record 3   int32 4  | int32 natoms | int32 4
//   per frame  [int32 48 | A gamma B beta alpha C (float64) | int32 48]    when icntrl[10] != 0
//              int32 4N | X (float32 x N) | int32 4N ;  the same for Y and Z ; [and W when icntrl[11] != 0]
//
// Both byte orders are read (the first record length tells).  Files with fixed atoms (icntrl[8] != 0: later frames
// hold only the free atoms) and 64-bit record markers are rejected with a message.
//
// gfn_enqueue_trajectory is the decode/compute overlap: reader threads pread + de-interleave + gather the two
// selections of frames k, k+T, ... into staging slots while the calling thread feeds finished slots, in frame order,
// to gfn_enqueue_atom_frames (pinned ring -> H2D on the copy stream -> residue kernels -> pair kernel).  Host code only;
// everything that touches the device goes through the public C ABI.
#include "../../include/gfn.h"

#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

struct gfn_traj {
    int fd = -1;
    bool swap = false;           // file byte order differs from the host's
    int natoms = 0;
    long long nframes = 0;       // frames actually present in the file (the header count is checked against the size)
    bool has_cell = false, has_4d = false;
    long long first_off = 0;     // offset of frame 0
    long long frame_bytes = 0;
    int istart = 0, nsavc = 0;
    double delta = 0.0;
    // staging of gfn_enqueue_trajectory, kept between calls (fresh 1 MB buffers cost a page fault per 4 KB)
    std::vector<std::vector<float>> slot_atoms;
    std::vector<std::vector<unsigned char>> raws;
    std::vector<unsigned char> read_raw;     // frame buffer of gfn_traj_read (kept: a fresh 1 MB vector costs more than the read)
    std::mutex read_mu;
    char err[256];
};

namespace {

thread_local char g_traj_err[256] = "";

int tfail(gfn_traj* t, int code, const char* fmt, ...)
{
    char* dst = t ? t->err : g_traj_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 256, fmt, ap);
    va_end(ap);
    return code;
}

inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }
inline uint64_t bswap64(uint64_t v) { return __builtin_bswap64(v); }

inline int32_t rd_i32(const unsigned char* p, bool swap)
{
    uint32_t v;
    memcpy(&v, p, 4);
    if (swap) v = bswap32(v);
    int32_t r;
    memcpy(&r, &v, 4);
    return r;
}

inline float rd_f32(const unsigned char* p, bool swap)
{
    uint32_t v;
    memcpy(&v, p, 4);
    if (swap) v = bswap32(v);
    float r;
    memcpy(&r, &v, 4);
    return r;
}

inline double rd_f64(const unsigned char* p, bool swap)
{
    uint64_t v;
    memcpy(&v, p, 8);
    if (swap) v = bswap64(v);
    double r;
    memcpy(&r, &v, 8);
    return r;
}

bool pread_all(int fd, void* buf, size_t n, long long off)
{
    unsigned char* p = static_cast<unsigned char*>(buf);
    while (n > 0) {
        ssize_t got = pread(fd, p, n, (off_t)off);
        if (got < 0) { if (errno == EINTR) continue; return false; }
        if (got == 0) return false;
        p += got; n -= (size_t)got; off += got;
    }
    return true;
}

// Checks the record markers of one raw frame and returns pointers to its X, Y, Z blocks and the cell record.
int locate(gfn_traj* t, const unsigned char* raw, long long frame, const unsigned char** cell, const unsigned char* xyz[3])
{
    const unsigned char* p = raw;
    *cell = nullptr;
    if (t->has_cell) {
        if (rd_i32(p, t->swap) != 48 || rd_i32(p + 52, t->swap) != 48)
            return tfail(t, GFN_EINVAL, "frame %lld: unit cell record is damaged", frame);
        *cell = p + 4;
        p += 56;
    }
    const long long nb = 4ll * t->natoms;        // fits an int32 marker: gfn_traj_open_dcd refuses natoms beyond that
    for (int k = 0; k < 3; ++k) {
        if ((long long)rd_i32(p, t->swap) != nb || (long long)rd_i32(p + 4 + nb, t->swap) != nb)
            return tfail(t, GFN_EINVAL, "frame %lld: coordinate record %d is damaged", frame, k);
        xyz[k] = p + 4;
        p += 8 + nb;
    }
    return GFN_OK;
}

// cell[6] = A, B, C, alpha, beta, gamma (degrees) from the record's A gamma B beta alpha C; CHARMM stores the
// cosines of the angles when all three lie in [-1, 1] (the convention MDAnalysis and VMD decode the same way)
void decode_cell(const gfn_traj* t, const unsigned char* rec, double cell[6])
{
    double v[6];
    for (int k = 0; k < 6; ++k) v[k] = rd_f64(rec + 8 * k, t->swap);
    double alpha = v[4], beta = v[3], gamma = v[1];
    if (fabs(alpha) <= 1.0 && fabs(beta) <= 1.0 && fabs(gamma) <= 1.0) {
        const double r2d = 180.0 / 3.14159265358979323846;
        alpha = 90.0 - asin(alpha) * r2d; beta = 90.0 - asin(beta) * r2d; gamma = 90.0 - asin(gamma) * r2d;
    }
    cell[0] = v[0]; cell[1] = v[2]; cell[2] = v[5]; cell[3] = alpha; cell[4] = beta; cell[5] = gamma;
}

void gather(const gfn_traj* t, const unsigned char* const xyz[3], const int* idx, int n, float* out)
{
    if (!t->swap) {
        const float* X = reinterpret_cast<const float*>(xyz[0]);     // blocks start 4 bytes after a marker: 4-byte aligned
        const float* Y = reinterpret_cast<const float*>(xyz[1]);
        const float* Z = reinterpret_cast<const float*>(xyz[2]);
        if (idx) for (int k = 0; k < n; ++k) { const int a = idx[k]; out[3 * k] = X[a]; out[3 * k + 1] = Y[a]; out[3 * k + 2] = Z[a]; }
        else     for (int k = 0; k < n; ++k) { out[3 * k] = X[k]; out[3 * k + 1] = Y[k]; out[3 * k + 2] = Z[k]; }
    } else {
        for (int k = 0; k < n; ++k) {
            const long long a = idx ? idx[k] : k;
            out[3 * k] = rd_f32(xyz[0] + 4 * a, true); out[3 * k + 1] = rd_f32(xyz[1] + 4 * a, true); out[3 * k + 2] = rd_f32(xyz[2] + 4 * a, true);
        }
    }
}

}  // namespace

extern "C" {

const char* gfn_traj_last_error(const gfn_traj_t* t) { return t ? t->err : g_traj_err; }

int gfn_traj_open_dcd(gfn_traj_t** out, const char* path)
{
    gfn_traj* t = nullptr;
    if (!out || !path) return tfail(t, GFN_EINVAL, "NULL argument");
    *out = nullptr;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return tfail(t, GFN_EINVAL, "cannot open %s: %s", path, strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return tfail(t, GFN_EINVAL, "cannot stat %s", path); }
    unsigned char hd[92];
    if (st.st_size < 92 + 4 || !pread_all(fd, hd, 92, 0)) { close(fd); return tfail(t, GFN_EINVAL, "%s is too short to be a DCD file", path); }
    bool swap;
    if (rd_i32(hd, false) == 84) swap = false;
    else if (rd_i32(hd, true) == 84) swap = true;
    else {
        close(fd);
        if (rd_i32(hd, false) == 0 || rd_i32(hd, true) == 0)
            return tfail(t, GFN_EINVAL, "%s: 64-bit record markers are not supported", path);
        return tfail(t, GFN_EINVAL, "%s is not a DCD file (first record length is not 84)", path);
    }
    if (memcmp(hd + 4, "CORD", 4) != 0) { close(fd); return tfail(t, GFN_EINVAL, "%s: header magic is not CORD (velocity files are not read)", path); }
    if (rd_i32(hd + 88, swap) != 84) { close(fd); return tfail(t, GFN_EINVAL, "%s: header record is damaged", path); }
    int32_t ic[20];
    for (int k = 0; k < 20; ++k) ic[k] = rd_i32(hd + 8 + 4 * k, swap);
    if (ic[8] != 0) { close(fd); return tfail(t, GFN_EINVAL, "%s: %d fixed atoms - trajectories with fixed atoms are not supported", path, ic[8]); }
    const bool charmm = ic[19] != 0;
    // title record
    unsigned char w[8];
    long long off = 92;
    if (!pread_all(fd, w, 8, off)) { close(fd); return tfail(t, GFN_EINVAL, "%s: truncated title record", path); }
    const int32_t tl = rd_i32(w, swap);
    if (tl < 4 || (tl - 4) % 80 != 0 || off + 8 + tl > st.st_size) { close(fd); return tfail(t, GFN_EINVAL, "%s: bad title record length %d", path, tl); }
    off += 4 + (long long)tl;
    if (!pread_all(fd, w, 4, off) || rd_i32(w, swap) != tl) { close(fd); return tfail(t, GFN_EINVAL, "%s: title record is damaged", path); }
    off += 4;
    unsigned char na[12];
    if (!pread_all(fd, na, 12, off) || rd_i32(na, swap) != 4 || rd_i32(na + 8, swap) != 4) { close(fd); return tfail(t, GFN_EINVAL, "%s: atom count record is damaged", path); }
    const int32_t natoms = rd_i32(na + 4, swap);
    if (natoms < 1) { close(fd); return tfail(t, GFN_EINVAL, "%s: %d atoms", path, natoms); }
    off += 12;
    // natoms comes from the file: a coordinate record (4*natoms bytes behind an int32 marker) must fit the marker, and one
    // whole frame must fit into what is left of the file - otherwise the header is damaged (or hostile)
    {
        const bool cell_ = ic[19] != 0 && ic[10] != 0, fourd_ = ic[19] != 0 && ic[11] != 0;
        const long long fb = (cell_ ? 56 : 0) + (3 + (fourd_ ? 1 : 0)) * (8 + 4ll * natoms);
        if (4ll * natoms > 2147483647ll || fb > (long long)st.st_size - off) {
            close(fd);
            return tfail(t, GFN_EINVAL, "%s: the header announces %d atoms, but the file holds less than one frame of that size", path, natoms);
        }
    }

    t = new gfn_traj();
    t->err[0] = 0;
    t->fd = fd; t->swap = swap; t->natoms = natoms;
    t->has_cell = charmm && ic[10] != 0;
    t->has_4d = charmm && ic[11] != 0;
    t->istart = ic[1]; t->nsavc = ic[2];
    t->delta = charmm ? (double)rd_f32(hd + 8 + 36, swap) : rd_f64(hd + 8 + 36, swap);
    t->first_off = off;
    t->frame_bytes = (t->has_cell ? 56 : 0) + (3 + (t->has_4d ? 1 : 0)) * (8 + 4ll * natoms);
    const long long present = (st.st_size - off) / t->frame_bytes;
    // the frame count of the header is what the writing program intended; a run that was cut short leaves fewer
    t->nframes = (ic[0] > 0 && ic[0] < present) ? ic[0] : present;
    *out = t;
    return GFN_OK;
}

void gfn_traj_close(gfn_traj_t* t)
{
    if (!t) return;
    if (t->fd >= 0) close(t->fd);
    delete t;
}

int gfn_traj_info(const gfn_traj_t* t, int* natoms, long long* nframes, int* has_cell)
{
    if (!t) return GFN_EINVAL;
    if (natoms) *natoms = t->natoms;
    if (nframes) *nframes = t->nframes;
    if (has_cell) *has_cell = t->has_cell ? 1 : 0;
    return GFN_OK;
}

int gfn_traj_read(gfn_traj_t* t, long long frame, const int* idx, int n, float* xyz, double cell[6])
{
    if (!t) return GFN_EINVAL;
    if (frame < 0 || frame >= t->nframes) return tfail(t, GFN_EINVAL, "frame %lld out of range (0..%lld)", frame, t->nframes - 1);
    if (idx) for (int k = 0; k < n; ++k) if (idx[k] < 0 || idx[k] >= t->natoms) return tfail(t, GFN_EINVAL, "atom index %d out of range (0..%d)", idx[k], t->natoms - 1);
    if (!idx) n = t->natoms;
    std::lock_guard<std::mutex> lk(t->read_mu);
    std::vector<unsigned char>& raw = t->read_raw;
    try {
        if (raw.size() < (size_t)t->frame_bytes) raw.resize((size_t)t->frame_bytes);
    } catch (const std::bad_alloc&) {
        return tfail(t, GFN_ENOMEM, "out of memory for a frame buffer of %lld bytes", t->frame_bytes);
    }
    if (!pread_all(t->fd, raw.data(), (size_t)t->frame_bytes, t->first_off + frame * t->frame_bytes)) return tfail(t, GFN_EINVAL, "frame %lld: read failed", frame);
    const unsigned char* cr; const unsigned char* b[3];
    int rc = locate(t, raw.data(), frame, &cr, b);
    if (rc) return rc;
    if (xyz) gather(t, b, idx, n, xyz);
    if (cell) {
        if (cr) decode_cell(t, cr, cell);
        else for (int k = 0; k < 6; ++k) cell[k] = 0.0;
    }
    return GFN_OK;
}

int gfn_enqueue_trajectory(gfn_t* h, gfn_traj_t* t, long long first, long long count, long long step,
                           const int* sel1, int nsel1, const int* sel2, int nsel2, int use_cell,
                           double* boxes_out, int nthreads)
{
    if (!h || !t) return GFN_EINVAL;
    if (count < 0 || step < 1 || first < 0) return tfail(t, GFN_EINVAL, "bad frame range (first %lld, count %lld, step %lld)", first, count, step);
    if (count == 0) return GFN_OK;
    if (first + (count - 1) * step >= t->nframes)
        return tfail(t, GFN_EINVAL, "frames %lld..%lld exceed the trajectory (%lld frames)", first, first + (count - 1) * step, t->nframes);
    int na1 = 0, na2 = 0;
    int rc = gfn_topology_natoms(h, &na1, &na2);
    if (rc) return tfail(t, rc, "%s", gfn_last_error(h));
    if (!sel1 || nsel1 != na1) return tfail(t, GFN_EINVAL, "selection 1 has %d atoms, the attached topology %d", nsel1, na1);
    const bool self = sel2 == nullptr;
    if (!self && nsel2 != na2) return tfail(t, GFN_EINVAL, "selection 2 has %d atoms, the attached topology %d", nsel2, na2);
    for (int k = 0; k < nsel1; ++k) if (sel1[k] < 0 || sel1[k] >= t->natoms) return tfail(t, GFN_EINVAL, "atom index %d out of range (0..%d)", sel1[k], t->natoms - 1);
    if (!self) for (int k = 0; k < nsel2; ++k) if (sel2[k] < 0 || sel2[k] >= t->natoms) return tfail(t, GFN_EINVAL, "atom index %d out of range (0..%d)", sel2[k], t->natoms - 1);
    if (use_cell && !t->has_cell) return tfail(t, GFN_EINVAL, "the trajectory stores no unit cell");
    if (use_cell < 0 || use_cell > 2) return tfail(t, GFN_EINVAL, "use_cell must be 0, 1 or 2");
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 16) nthreads = 16;
    if (nthreads > count) nthreads = (int)count;

    struct Slot { float* atoms = nullptr; double box[6]; int state = 0; /* 0 free, 1 ready */ int rc = 0; };
    const int nslots = 2 * nthreads + 2;
    std::vector<Slot> slots(nslots);
    const size_t per = 3 * ((size_t)nsel1 + (self ? 0 : (size_t)nsel2));
    try {       // no C++ exception may cross the C ABI
        if ((int)t->slot_atoms.size() < nslots) t->slot_atoms.resize(nslots);
        if ((int)t->raws.size() < nthreads) t->raws.resize(nthreads);
        for (int k = 0; k < nslots; ++k) {
            if (t->slot_atoms[k].size() < per) t->slot_atoms[k].resize(per);
            slots[k].atoms = t->slot_atoms[k].data();
        }
        for (int k = 0; k < nthreads; ++k) if (t->raws[k].size() < (size_t)t->frame_bytes) t->raws[k].resize((size_t)t->frame_bytes);
    } catch (const std::bad_alloc&) {
        return tfail(t, GFN_ENOMEM, "out of memory for the staging buffers (%d slots of %zu floats, %d frame buffers of %lld bytes)",
                     nslots, per, nthreads, t->frame_bytes);
    }
    std::mutex mu;
    std::condition_variable cv;
    long long consumed = 0;          // frames handed to the ring so far
    bool stop = false;

    auto worker = [&](int w) {
        std::vector<unsigned char>& raw = t->raws[w];
        for (long long k = w; k < count; k += nthreads) {
            Slot& s = slots[k % nslots];
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || (consumed > k - nslots && s.state == 0); });
                if (stop) return;
            }
            const long long frame = first + k * step;
            int wrc = GFN_OK;
            const unsigned char* cr = nullptr; const unsigned char* b[3];
            if (!pread_all(t->fd, raw.data(), (size_t)t->frame_bytes, t->first_off + frame * t->frame_bytes)) {
                std::lock_guard<std::mutex> lk(mu);
                wrc = tfail(t, GFN_EINVAL, "frame %lld: read failed", frame);
            } else {
                std::lock_guard<std::mutex> lk(mu);      // error text is shared
                wrc = locate(t, raw.data(), frame, &cr, b);
            }
            if (!wrc) {
                gather(t, b, sel1, nsel1, s.atoms);
                if (!self) gather(t, b, sel2, nsel2, s.atoms + 3 * (size_t)nsel1);
                if (use_cell) {
                    double c[6];
                    decode_cell(t, cr, c);
                    if (fabs(c[3] - 90.0) > 1e-6 || fabs(c[4] - 90.0) > 1e-6 || fabs(c[5] - 90.0) > 1e-6) {
                        std::lock_guard<std::mutex> lk(mu);
                        wrc = tfail(t, GFN_EINVAL, "frame %lld: the unit cell is not orthorhombic (%g %g %g degrees); the minimum image of gfunction.pyx:145-147 needs a rectangular box", frame, c[3], c[4], c[5]);
                    }
                    for (int a = 0; a < 3; ++a) {
                        const double len = use_cell == 2 ? (double)(float)c[a] : c[a];   // 2: through float32 like MDAnalysis' ts.dimensions
                        s.box[a] = len; s.box[3 + a] = len / 2.0;
                    }
                }
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                s.rc = wrc; s.state = 1;
            }
            cv.notify_all();
        }
    };

    std::vector<std::thread> pool;
    for (int w = 0; w < nthreads; ++w) pool.emplace_back(worker, w);
    int out_rc = GFN_OK;
    for (long long k = 0; k < count; ++k) {
        Slot& s = slots[k % nslots];
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return s.state == 1; });
        }
        out_rc = s.rc;
        if (!out_rc) {
            const float* a1 = s.atoms;
            const float* a2 = self ? nullptr : a1 + 3 * (size_t)nsel1;
            out_rc = gfn_enqueue_atom_frames(h, 1, a1, 0, a2, 0, use_cell ? s.box : nullptr);
            if (out_rc) { std::lock_guard<std::mutex> lk(mu); tfail(t, out_rc, "%s", gfn_last_error(h)); }    // err text is shared with the workers
            if (!out_rc && boxes_out && use_cell) memcpy(boxes_out + 3 * k, s.box, sizeof(double) * 3);
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            s.state = 0;
            consumed = k + 1;
            if (out_rc) stop = true;
        }
        cv.notify_all();
        if (out_rc) break;
    }
    for (std::thread& th : pool) th.join();
    return out_rc;
}

}  // extern "C"

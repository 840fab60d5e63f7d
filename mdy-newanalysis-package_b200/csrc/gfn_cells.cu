// gfn_cells.cu - tile skipping for sparse neighbourhoods (GFN_FLAG_CELLS): frames are sorted into space-filling-curve
// order on the device, and the pair kernel visits only the (core tile, surround tile) pairs that can hold a pair within
// histo_max.  The reference (cdef void calcHisto, gfunction.pyx:110-165) evaluates ALL n1 x n2 pairs and rejects almost
// all of them at :149 (98.6 % at BASELINE configs[2], 99.95 % at configs[4]); a pair it accepts has a single-shift
// minimum-image distance <= histo_max, and that distance is never smaller than the true periodic distance - so a tile
// pair whose bounding boxes are farther apart than histo_max (periodic, per component) contributes nothing and is not
// visited.  Every visited pair goes through the unchanged pair kernel, so counts stay bit-exact and weighted sums keep
// their bound.
//
//   K_key     site -> 64-bit key  frame << 32 | Morton code of its cell (1024^3 cells over the wrapped box), padding last
//   sort      cub::DeviceRadixSort::SortPairs (stable: the permutation is deterministic), one call per selection and launch
//   K_gather  Site records into sorted order, scrambled again inside each full tile; Site.fw := original index (the pair kernel needs it to restore the
//             reference's core/surround roles in triangular frames and for the quirk-Q3 spill position)
//   K_bbox    per tile of consecutive sorted sites: bounding box of the wrapped coordinates (rounded outwards)
//   K_list    one warp per (frame, core tile): surround tiles whose box lies within histo_max of the core tile's box,
//             in ascending order (triangular frames: from the tile that holds the core tile's first site)
//   K_scan    exclusive prefix of the list lengths = the flat work sequence the pair kernel's CTAs slice
#include "gfn_device.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <stdlib.h>

namespace gfn {

__device__ __forceinline__ unsigned spread10(unsigned v)
{
    v &= 1023u;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// coordinate wrapped into [0, L): what the bounding boxes are built from (unwrapped input allowed, quirk Q8)
__device__ __forceinline__ double wrap01(double x, double L)
{
    double t = x / L;
    t -= floor(t);
    if (!(t >= 0.0) || t >= 1.0) t = 0.0;      // NaN / rounding to 1
    return t;
}

__global__ void __launch_bounds__(256)
cells_key_kernel(const Site* __restrict__ site, int n, int n_pad, const double* __restrict__ box,
                 unsigned long long* __restrict__ keys, unsigned* __restrict__ vals)
{
    const int f = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_pad) return;
    unsigned code = 0xFFFFFFFFu;
    if (s < n) {
        const Site& r = site[(long long)f * n_pad + s];
        const double* b = box + 6ll * f;
        const unsigned ix = min(1023u, (unsigned)(wrap01(r.x, b[0]) * 1024.0));
        const unsigned iy = min(1023u, (unsigned)(wrap01(r.y, b[1]) * 1024.0));
        const unsigned iz = min(1023u, (unsigned)(wrap01(r.z, b[2]) * 1024.0));
        code = spread10(ix) | (spread10(iy) << 1) | (spread10(iz) << 2);
    }
    keys[(long long)f * n_pad + s] = ((unsigned long long)f << 32) | code;
    vals[(long long)f * n_pad + s] = (unsigned)s;
}

__global__ void __launch_bounds__(256)
cells_gather_kernel(const Site* __restrict__ in, Site* __restrict__ out, const unsigned* __restrict__ perm, int n, int n_pad, int scramble)
{
    const int f = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pad) return;
    // Inside a full tile of kTJ sites the Morton order is scrambled again (x -> 37 x + 11 mod 128): the tile keeps its sites
    // and its bounding box, but neighbouring slots no longer hold neighbouring sites.  Otherwise the survivors a drain warp
    // evaluates together (consecutive ring entries: adjacent core rows against consecutive surround sites) would have nearly
    // equal distances, hit the same few bins and serialise in the histogram update (measured: 4x slower than unsorted frames).
    int kk = k;
    if (scramble > 1 && (k | (scramble - 1)) < n) kk = (k & ~(scramble - 1)) | ((293 * (k & (scramble - 1)) + 11) & (scramble - 1));
    const unsigned s = perm[(long long)f * n_pad + kk];
    const uint4* src = reinterpret_cast<const uint4*>(in + (long long)f * n_pad + s);
    uint4* dst = reinterpret_cast<uint4*>(out + (long long)f * n_pad + k);
    uint4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2), d = __ldg(src + 3), e = __ldg(src + 4);
    e.w = s;                                    // Site.fw: the original index
    dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = d; dst[4] = e;
}

// bbox[(f * ntiles + t) * 6 + {0,1,2}] = lower corner, {3,4,5} = upper corner (wrapped coordinates, rounded outwards);
// an empty tile gets lower > upper and is never within range of anything
__global__ void __launch_bounds__(256)
cells_bbox_kernel(const Site* __restrict__ sorted, int n, int n_pad, int tile, int ntiles, const double* __restrict__ box, float* __restrict__ bbox)
{
    const int f = blockIdx.y;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= ntiles) return;
    const double* b = box + 6ll * f;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    const int end = min((t + 1) * tile, n);
    for (int k = t * tile + lane; k < end; k += 32) {
        const Site& r = sorted[(long long)f * n_pad + k];
        const double w[3] = {wrap01(r.x, b[0]) * b[0], wrap01(r.y, b[1]) * b[1], wrap01(r.z, b[2]) * b[2]};
#pragma unroll
        for (int c = 0; c < 3; ++c) { lo[c] = fminf(lo[c], __double2float_rd(w[c])); hi[c] = fmaxf(hi[c], __double2float_ru(w[c])); }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
        for (int d = 16; d; d >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], d));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], d));
        }
    if (lane == 0) {
        float* o = bbox + ((long long)f * ntiles + t) * 6;
        o[0] = lo[0]; o[1] = lo[1]; o[2] = lo[2]; o[3] = hi[0]; o[4] = hi[1]; o[5] = hi[2];
    }
}

// smallest periodic distance between the intervals [alo, ahi] and [blo, bhi] (both inside [0, L)), never overestimated
__device__ __forceinline__ float interval_gap(float alo, float ahi, float blo, float bhi, float L)
{
    const float g0 = fmaxf(fmaxf(blo - ahi, alo - bhi), 0.f);
    const float g1 = fmaxf(fmaxf(blo + L - ahi, alo - (bhi + L)), 0.f);
    const float g2 = fmaxf(fmaxf(blo - L - ahi, alo - (bhi - L)), 0.f);
    return fminf(g0, fminf(g1, g2));
}

__global__ void __launch_bounds__(256)
cells_list_kernel(const float* __restrict__ boxA, const float* __restrict__ boxB, int n_itiles, int n_jtiles, int TI, int tri,
                  const double* __restrict__ box, float reach, int* __restrict__ list, int* __restrict__ count)
{
    const int f = blockIdx.y;
    const int ti = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (ti >= n_itiles) return;
    const float* a = boxA + ((long long)f * n_itiles + ti) * 6;
    const double* b = box + 6ll * f;
    // box lengths rounded down and a relative slack on the reach: both make the test more permissive
    const float Lx = __double2float_rd(b[0]), Ly = __double2float_rd(b[1]), Lz = __double2float_rd(b[2]);
    const float r = reach * (1.0f + 4e-6f) + 4e-6f * fmaxf(Lx, fmaxf(Ly, Lz));
    const float r2 = r * r * (1.0f + 1e-6f);
    int* out = list + ((long long)f * n_itiles + ti) * n_jtiles;
    int nout = 0;
    const int jt_first = tri ? (ti * TI) / kTJ : 0;
    for (int base = jt_first; base < n_jtiles; base += 32) {
        const int jt = base + lane;
        bool hit = false;
        if (jt < n_jtiles) {
            const float* q = boxB + ((long long)f * n_jtiles + jt) * 6;
            const float gx = interval_gap(a[0], a[3], q[0], q[3], Lx);
            const float gy = interval_gap(a[1], a[4], q[1], q[4], Ly);
            const float gz = interval_gap(a[2], a[5], q[2], q[5], Lz);
            const float d2 = gx * gx + gy * gy + gz * gz;
            hit = (a[0] <= a[3]) && (q[0] <= q[3]) && d2 <= r2;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) out[nout + __popc(m & ((1u << lane) - 1u))] = jt;
        nout += __popc(m);
    }
    if (lane == 0) count[(long long)f * n_itiles + ti] = nout;
}

// Dense-kernel units of cell-sorted frames: one warp per (frame, surround tile jt) lists the 64-row core BLOCKS whose box
// lies within reach of the tile's box, in ascending order - a triangular frame only the blocks that reach the tile's last
// column (64-row blocks of the dense kernel: ib <= 2 jt + 1; 128-row blocks of the count kernel: ib <= jt - as in the un-pruned
// enumerations).  An entry is never empty (block 0 stands in: the kernel's tile
// ring counts the units of every tile).  Two passes: fill == 0 counts, fill == 1 writes the lists at their prefix offsets
// (compact storage, capacity checked: an overflowing launch raises the error flag and lists nothing beyond).
__global__ void __launch_bounds__(256)
cells_dense_list_kernel(const float* __restrict__ boxBlk, const float* __restrict__ boxTile, int n_blk, int n_jtiles, int tri, int blk_per_tile,
                        const double* __restrict__ box, float reach, int fill, int* __restrict__ count, const long long* __restrict__ start,
                        int* __restrict__ list, long long capacity, unsigned long long* __restrict__ counters)
{
    const int f = blockIdx.y;
    const int jt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (jt >= n_jtiles) return;
    const long long e = (long long)f * n_jtiles + jt;
    const float* a = boxTile + e * 6;
    const double* b = box + 6ll * f;
    const float Lx = __double2float_rd(b[0]), Ly = __double2float_rd(b[1]), Lz = __double2float_rd(b[2]);
    const float r = reach * (1.0f + 4e-6f) + 4e-6f * fmaxf(Lx, fmaxf(Ly, Lz));
    const float r2 = r * r * (1.0f + 1e-6f);
    const long long base = fill ? start[e] : 0;
    const bool room = !fill || start[e + 1] <= capacity;
    if (fill && !room) { if (lane == 0) atomicOr(counters + 1, 16ull); return; }
    int nout = 0;
    const int ib_end = tri ? min(blk_per_tile * jt + blk_per_tile, n_blk) : n_blk;
    for (int b0 = 0; b0 < ib_end; b0 += 32) {
        const int ib = b0 + lane;
        bool hit = false;
        if (ib < ib_end) {
            const float* q = boxBlk + ((long long)f * n_blk + ib) * 6;
            const float gx = interval_gap(a[0], a[3], q[0], q[3], Lx);
            const float gy = interval_gap(a[1], a[4], q[1], q[4], Ly);
            const float gz = interval_gap(a[2], a[5], q[2], q[5], Lz);
            hit = (a[0] <= a[3]) && (q[0] <= q[3]) && gx * gx + gy * gy + gz * gz <= r2;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (fill && hit) list[base + nout + __popc(m & ((1u << lane) - 1u))] = ib;
        nout += __popc(m);
    }
    if (nout == 0) { if (fill && lane == 0) list[base] = 0; nout = 1; }
    if (!fill && lane == 0) count[e] = nout;
}

// exclusive prefix of count[0 .. nent) -> start[0 .. nent]; one block
__global__ void __launch_bounds__(1024)
cells_scan_kernel(const int* __restrict__ count, long long* __restrict__ start, long long nent, unsigned long long* __restrict__ visits)
{
    __shared__ long long part[1024];
    const int tid = threadIdx.x;
    const long long per = (nent + 1023) / 1024;
    const long long lo = (long long)tid * per, hi = lo + per < nent ? lo + per : nent;
    long long s = 0;
    for (long long k = lo; k < hi; ++k) s += count[k];
    part[tid] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const long long v = tid >= d ? part[tid - d] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    long long run = tid ? part[tid - 1] : 0;
    for (long long k = lo; k < hi; ++k) { start[k] = run; run += count[k]; }
    if (tid == 1023) { start[nent] = part[1023]; if (visits) atomicAdd(visits, (unsigned long long)part[1023]); }
}

size_t cells_sort_temp_bytes(long long items)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const unsigned*)nullptr, (unsigned*)nullptr, (int)items, 0, 64);
    return bytes;
}

// Sorts one selection of nframes frames: site -> sorted (Site.fw = original index).
cudaError_t cells_sort_launch(const Site* site, Site* sorted, int n, int n_pad, int nframes, const double* box,
                              unsigned long long* keys_in, unsigned long long* keys_out, unsigned* vals_in, unsigned* vals_out,
                              void* tmp, size_t tmp_bytes, cudaStream_t s)
{
    if (nframes > 65535) return cudaErrorInvalidValue;
    const dim3 grid((n_pad + 255) / 256, nframes);
    cells_key_kernel<<<grid, 256, 0, s>>>(site, n, n_pad, box, keys_in, vals_in);
    int top = 32;
    while ((1ll << (top - 32)) < nframes) ++top;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, vals_out, n_pad * nframes, 0, top, s);
    if (e != cudaSuccess) return e;
    int scramble = kTJ;                 // power of two: block of sorted slots whose order is scrambled again
    if (const char* e = getenv("GFN_CELLS_SCRAMBLE")) scramble = atoi(e);
    cells_gather_kernel<<<grid, 256, 0, s>>>(site, sorted, vals_out, n, n_pad, scramble);
    return cudaGetLastError();
}

// Bounding boxes of the tiles of `tile` consecutive sorted sites.
cudaError_t cells_bbox_launch(const Site* sorted, int n, int n_pad, int nframes, const double* box, int tile, int ntiles, float* bbox, cudaStream_t s)
{
    cells_bbox_kernel<<<dim3((ntiles + 7) / 8, nframes), 256, 0, s>>>(sorted, n, n_pad, tile, ntiles, box, bbox);
    return cudaGetLastError();
}

cudaError_t cells_list_launch(const float* boxA, const float* boxB, int n_itiles, int n_jtiles, int TI, int tri, int nframes,
                              const double* box, double hmax, int* list, int* count, long long* start, unsigned long long* visits, cudaStream_t s)
{
    cells_list_kernel<<<dim3((n_itiles + 7) / 8, nframes), 256, 0, s>>>(boxA, boxB, n_itiles, n_jtiles, TI, tri, box, (float)hmax * (1.0f + 1e-6f), list, count);
    cells_scan_kernel<<<1, 1024, 0, s>>>(count, start, (long long)nframes * n_itiles, visits);
    return cudaGetLastError();
}

// Visit lists of the dense kernel's units (count, prefix, fill).  capacity: ints the list buffer holds.
cudaError_t cells_dense_lists_launch(const float* boxBlk, const float* boxTile, int n_blk, int n_jtiles, int tri, int blk_per_tile, int nframes,
                                     const double* box, double hmax, int* list, long long capacity, int* count, long long* start,
                                     unsigned long long* counters, cudaStream_t s)
{
    const dim3 grid((n_jtiles + 7) / 8, nframes);
    const float reach = (float)hmax * (1.0f + 1e-6f);
    cells_dense_list_kernel<<<grid, 256, 0, s>>>(boxBlk, boxTile, n_blk, n_jtiles, tri, blk_per_tile, box, reach, 0, count, nullptr, nullptr, capacity, counters);
    cells_scan_kernel<<<1, 1024, 0, s>>>(count, start, (long long)nframes * n_jtiles, counters + 6);
    cells_dense_list_kernel<<<grid, 256, 0, s>>>(boxBlk, boxTile, n_blk, n_jtiles, tri, blk_per_tile, box, reach, 1, count, start, list, capacity, counters);
    return cudaGetLastError();
}

}  // namespace gfn

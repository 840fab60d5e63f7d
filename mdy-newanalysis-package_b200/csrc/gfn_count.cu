// gfn_count.cu - the pair kernel for g000-ONLY handles with dense neighbourhoods (BASELINE configs[0]: SPC/E water, O-O
// g(r) up to ~L/2, 47 % of all pairs in range).
//
// It replaces the same reference loop as the other two pair kernels - cdef void calcHisto, gfunction.pyx:110-165 with
// mode_sel == 1: minimum image :141-147, dist :148, acceptance :149, bin :150, histogram[pos] += off_diag :152 - and
// produces the same integers.  With nothing but a pair COUNT to produce, the fp64 work of a pair is its bin and nothing
// else, and almost every pair's bin can be decided in fp32:
//
//   * CLASSIFY (all pairs, fp32, FMA/ALU pipes, 27.7 SASS instructions per pair, no branch): minimum-image distance from the fp32-rounded
//     coordinates, q = (dist - hmin) * invdx, floor(q) through a round-down add of 2^23.  The estimate q differs from the
//     reference's exact value by less than EPS (bound below, ~1.5e-3 of a bin for a wrapped water box):
//         certain   : q is farther than EPS from every bin edge and from both ends of the accepted range
//                     -> the reference's bin is floor(q): one native shared-memory integer atomic (red.shared.add.u32)
//         rejected  : q is farther than EPS outside the accepted range -> nothing
//         ambiguous : everything else (~0.3 % of the in-range pairs) -> one bit in a per-lane mask
//   * EXACT (ambiguous pairs only, fp64): after each chunk of 32 surround sites every lane walks its mask and evaluates
//     those pairs with the reference's own fp64 operations (the g000 path of gfn_dense.cu: |d| minimum image, d^2, bin
//     from the table of exact d^2 thresholds).  A frame whose coordinates are far outside the box has a large EPS, hence
//     more ambiguous pairs - slower, never wrong.
//
//   * a lane owns 4 core rows (fp32 coordinates in registers), the surround sites of a 128-site tile are broadcast from
//     shared memory (one LDS.128 per 4 pairs): WORK UNIT = 128 core rows x 128 surround sites, one unit per warp at a time,
//     units enumerated (frame, surround tile, row block) - a triangular (n1 == n2) frame has only the row blocks up to the
//     diagonal, and the diagonal unit skips / masks its lower half in 32 x 32 sub-blocks.
//   * warps are independent: each stages the fp32 part of its surround tile itself (2 KB), no CTA-wide barrier inside
//     the loop.  Counts are integers, so the result does not depend on the order of anything.
//   * its cost is proportional to the pairs it visits whatever their density, which also makes it the kernel of g000-only
//     handles with tile skipping (GFN_FLAG_CELLS, gfn_cells.cu): the units then come from per-tile visit lists of 128-row
//     blocks of the Morton-sorted frame (1 M sites: 251 -> 7 ms per frame).
//
// EPS (in bins), u = 2^-24, M = max |coordinate| of the frame, h = largest half box length:
//   component:  the fp32 minimum image  w = h - ||s - c| - h|  equals the reference's |component| in exact arithmetic
//               (single shift included, quirk Q8) and is 1-Lipschitz in s - c, 2-Lipschitz in h:
//               |w - ref| <= e_w = u (10 M + 7 h) + |L - 2h|    (input roundings 4uM + 2uh, three fp32 operations)
//   distance:   | ||w|| - dist | <= sqrt(3) e_w; the fma chain, sqrt.approx (2^-21 budgeted) and the two products add
//               at most 15 u relative
//   q:          |q - q_ref| <= invdx (sqrt(3) e_w + 32 u hmax) + 2 u (hmin invdx + histo_n + 4) + 1e-9,  times 1.25
// GFN_DEBUG bit 3 (value 8) makes the kernel evaluate EVERY pair exactly as well and count the certain pairs whose
// fp32 bin differs (counters[4]); tests/test_gpu_parity.py requires zero.
#include "gfn_device.cuh"

#include <stdio.h>
#include <stdlib.h>

namespace gfn {

constexpr int kCW = 16;             // warps per CTA
constexpr int kCR = 4;              // core rows per lane
constexpr int kCRows = 32 * kCR;    // core rows per unit

__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ void red_shared_u32(unsigned addr, unsigned v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// One ambiguous (or, in self-check mode, any) pair with the reference's fp64 operations.  Returns the bin (>= 0) of an
// accepted pair, -1 otherwise.
__device__ __forceinline__ int exact_bin(const PairParams& P, const Box& bx, const Site* __restrict__ sa, const Site* __restrict__ sb, int row, int j)
{
    if (row >= P.n1 || j >= P.n2 || (P.tri && j < row)) return -1;
    const double2 cxy = __ldg(reinterpret_cast<const double2*>(&sa[row].x));
    const double cz = __ldg(&sa[row].z);
    const double2 sxy = __ldg(reinterpret_cast<const double2*>(&sb[j].x));
    const double sz = __ldg(&sb[j].z);
    const double ax = wrap_abs(sxy.x, cxy.x, bx.hx, bx.Lx);                                         // :141-147
    const double ay = wrap_abs(sxy.y, cxy.y, bx.hy, bx.Ly);
    const double az = wrap_abs(sz, cz, bx.hz, bx.Lz);
    const double d2 = DADD(DADD(DMUL(ax, ax), DMUL(ay, ay)), DMUL(az, az));                         // :148
    if (!(d2 >= P.acc_lo && d2 <= P.acc_hi)) return -1;                                             // :149
    // the bin is floor(e) or floor(e) + 1 for an estimate e a quarter bin low and good to +-0.2 (gfn_dense.cu)
    const float fd = __double2float_rn(d2);
    const float e = fmaf(fd * rsqrt_approx(fd), (float)P.invdx, -(float)(P.hmin * P.invdx + 0.25));
    const unsigned b = min((unsigned)max(__float2int_rz(e), 0), (unsigned)P.pos_max);
    return (int)b + (d2 >= __ldg(P.bin_thr + b + 1) ? 1 : 0);                                       // :150
}

template <bool TRI>
__global__ void __launch_bounds__(kCW * 32, 1)
count_kernel(const PairParams P)
{
    extern __shared__ __align__(128) unsigned char smem[];
    float4* tiles = reinterpret_cast<float4*>(smem);                               // [kCW][kTJ] fp32 surround coordinates
    unsigned* hist = reinterpret_cast<unsigned*>(tiles + kCW * kTJ);               // [nrep][hn]
    const int hn = P.histo_n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int k = tid; k < P.nrep * hn; k += kCW * 32) hist[k] = 0u;
    __syncthreads();

    float4* tile = tiles + warp * kTJ;
    unsigned* hw_ = hist + (P.nrep == kCW ? warp : 0) * hn;
    const unsigned hraw = smem_u32(hw_);
    const unsigned hbase = hraw - (0x4B000000u << 2);      // + (bits of (floor(q) + 2^23) << 2) = address of the bin
    const unsigned wt = TRI ? 2u : 1u;                     // :131-134 (i == j has dist 0 and is never accepted, :149)
    const bool selfcheck = P.debug & 8;

    // cell-sorted frames (gfn_cells.cu): entry gt = frame * n_jtiles + jt lists the 128-row core blocks within histo_max of
    // surround tile jt; visit_start is the exclusive prefix of the list lengths = the flat unit sequence
    const bool listed = P.visit_start != nullptr;
    const long long nent = (long long)P.nframes * P.n_jtiles;
    const int upf = P.tiles_per_frame;
    long long total_units = listed ? P.visit_start[nent] : (long long)upf * P.nframes;
    if (listed && total_units > P.visit_cap) total_units = 0;       // the list buffer overflowed (flagged by gfn_cells.cu)
    const long long ubeg = total_units * blockIdx.x / gridDim.x;
    const long long uend = total_units * (blockIdx.x + 1) / gridDim.x;

    // handle constants in bins: accepted  <=>  qlo <= q_ref <= qhi, counted in the row itself while q_ref < hn (quirk Q3)
    const double qlo = fmax(0.0, 1.0 - P.hmin * P.invdx);
    const double qhi = (P.hmax - P.hmin) * P.invdx;
    const double qtop = fmin(qhi, (double)hn);
    const double qmid_d = 0.5 * (qlo + qtop);
    const float qmid = (float)qmid_d;
    const float invdx_f = (float)P.invdx, nhminv_f = -(float)(P.hmin * P.invdx);

    unsigned long long n_amb = 0, n_amb_acc = 0, n_ovf = 0, n_bad = 0, n_w1 = 0;
    int f_cur = -1;
    Box bx; bx.Lx = bx.Ly = bx.Lz = bx.hx = bx.hy = bx.hz = 0.0;
    float hxf = 0.f, hyf = 0.f, hzf = 0.f, e1 = 0.f, hc = 0.f, hwid = 0.f;

    for (int iter = 0;; ++iter) {
        const long long u = ubeg + warp + (long long)kCW * iter;
        if (u >= uend) break;
        int f, jt, ib;
        if (listed) {
            long long lo = 0, hi = nent;                       // the last entry with start <= u (entries are never empty)
            while (hi - lo > 1) { const long long mid = (lo + hi) >> 1; if (P.visit_start[mid] <= u) lo = mid; else hi = mid; }
            f = (int)(lo / P.n_jtiles); jt = (int)(lo - (long long)f * P.n_jtiles);
            ib = __ldg(P.visit_list + u);
        } else {
            f = (int)(u / upf);
            const int ru = (int)(u - (long long)f * upf);
            int lo = 0, hi = P.n_jtiles;                       // surround tile jt with tile_start[jt] <= ru < tile_start[jt+1]
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (P.tile_start[mid] <= ru) lo = mid; else hi = mid; }
            jt = lo; ib = ru - P.tile_start[lo];
        }
        const int j0 = jt * kTJ, i0 = ib * kCRows;
        const bool diag = TRI && ib == jt;
        const Site* sa = P.siteA + (long long)f * P.strideA;
        const Site* sb = P.siteB + (long long)f * P.strideB;

        if (f != f_cur) {
            f_cur = f;
            const double* b = P.box + 6ll * f;
            bx.Lx = b[0]; bx.Ly = b[1]; bx.Lz = b[2]; bx.hx = b[3]; bx.hy = b[4]; bx.hz = b[5];
            const double uu = 5.9604644775390625e-08;
            const double M = (double)fmaxf(P.maxA[f], P.maxB[f]) * (1.0 + 1.2e-7);
            const double hb = fmax(b[3], fmax(b[4], b[5]));
            const double inc = fmax(fabs(b[0] - 2.0 * b[3]), fmax(fabs(b[1] - 2.0 * b[4]), fabs(b[2] - 2.0 * b[5])));
            const double ew = uu * (10.0 * M + 7.0 * hb) + inc;
            const double eps = 1.25 * (P.invdx * (1.7320508075688774 * ew + 32.0 * uu * P.hmax) + 2.0 * uu * (P.hmin * P.invdx + hn + 4.0) + 1e-9);
            const double slack = 4.0 * uu * ((double)hn + 4.0 + fabs(qmid_d));      // roundings of qmid and of q - qmid
            e1 = __double2float_rd(0.5 - eps);
            hc = __double2float_rd(0.5 * (qtop - qlo) - eps - slack);
            hwid = __double2float_ru(fmax(qhi + eps - qmid_d, qmid_d - qlo + eps) + slack);
            if (!(eps < 0.5)) { e1 = -1.f; hc = -1.f; }        // nothing is certain (also eps = NaN / inf)
            if (!(hwid == hwid)) hwid = 3.0e38f;
            hxf = (float)b[3]; hyf = (float)b[4]; hzf = (float)b[5];
        }

        // ---- stage the unit: fp32 surround coordinates -> this warp's tile, 4 core rows -> registers ------------------
        __syncwarp();
#pragma unroll
        for (int k = 0; k < kTJ / 32; ++k)
            tile[lane + 32 * k] = __ldg(reinterpret_cast<const float4*>(&sb[j0 + lane + 32 * k].fx));
        float cx[kCR], cy[kCR], cz[kCR];
#pragma unroll
        for (int r = 0; r < kCR; ++r) {
            const float4 c4 = __ldg(reinterpret_cast<const float4*>(&sa[i0 + 32 * r + lane].fx));
            cx[r] = c4.x; cy[r] = c4.y; cz[r] = c4.z;
        }
        __syncwarp();

        // One pair: classification and, if certain, the count.  The tail is predicated PTX - no branch per pair, so the
        // four rows' chains interleave: p = certain, w = in the widened window and not certain (ambiguous).
#define GFN_COUNT_HEAD(R)                                                                                             \
            const float dx = s4.x - cx[R], dy = s4.y - cy[R], dz = s4.z - cz[R];                                      \
            const float vx = fabsf(dx) - hxf, vy = fabsf(dy) - hyf, vz = fabsf(dz) - hzf;                             \
            const float wx = hxf - fabsf(vx), wy = hyf - fabsf(vy), wz = hzf - fabsf(vz);                             \
            float t = wx * wx;                                                                                        \
            t = fmaf(wy, wy, t);                                                                                      \
            t = fmaf(wz, wz, t);                                                                                      \
            const float q = fmaf(sqrt_approx(t), invdx_f, nhminv_f);                                                  \
            const float bf = __fadd_rd(q, 8388608.0f);                 /* floor(q) + 2^23 for 0 <= q < 2^23 */        \
            const float g = q - (bf - 8388607.5f);                     /* q - (floor(q) + 0.5) */                    \
            const float m = q - qmid;                                                                                 \
            const unsigned addr = hbase + (__float_as_uint(bf) << 2);                                                 \
            GFN_CHECK(P, !((fabsf(g) < e1) && (fabsf(m) <= hc)) || (addr - hraw) < 4u * (unsigned)hn);
#define GFN_COUNT_PAIR(R, BIT, LOC)                                                                                   \
        do {                                                                                                          \
            GFN_COUNT_HEAD(R)                                                                                         \
            asm volatile("{\n\t.reg .pred p, w;\n\t"                                                                 \
                         "setp.lt.f32 p, %3, %5;\n\t"                                                                 \
                         "setp.le.and.f32 p, %4, %6, p;\n\t"                                                          \
                         "setp.le.and.f32 w, %4, %7, !p;\n\t"                                                         \
                         "@p red.shared.add.u32 [%1], %2;\n\t"                                                        \
                         "@w or.b32 %0, %0, %8;\n\t}"                                                                 \
                         : "+r"(LOC) : "r"(addr), "r"(wt), "f"(fabsf(g)), "f"(fabsf(m)), "f"(e1), "f"(hc), "f"(hwid), "r"(BIT));  \
        } while (0)
        // ... with the triangle predicate  JI > TR  (surround site right of the lane's row)
#define GFN_COUNT_PAIR_TRI(R, BIT, LOC, JI, TR)                                                                       \
        do {                                                                                                          \
            GFN_COUNT_HEAD(R)                                                                                         \
            asm volatile("{\n\t.reg .pred p, w, c;\n\t"                                                              \
                         "setp.gt.s32 c, %9, %10;\n\t"                                                                \
                         "setp.lt.and.f32 p, %3, %5, c;\n\t"                                                          \
                         "setp.le.and.f32 p, %4, %6, p;\n\t"                                                          \
                         "setp.le.and.f32 w, %4, %7, !p;\n\t"                                                         \
                         "and.pred w, w, c;\n\t"                                                                      \
                         "@p red.shared.add.u32 [%1], %2;\n\t"                                                        \
                         "@w or.b32 %0, %0, %8;\n\t}"                                                                 \
                         : "+r"(LOC) : "r"(addr), "r"(wt), "f"(fabsf(g)), "f"(fabsf(m)), "f"(e1), "f"(hc), "f"(hwid), "r"(BIT), "r"(JI), "r"(TR));  \
        } while (0)

#pragma unroll 1
        for (int jc = 0; jc < kTJ / 32; ++jc) {
            const float4* tj = tile + jc * 32;
            unsigned amb[kCR];
#pragma unroll
            for (int r = 0; r < kCR; ++r) amb[r] = 0u;
            if (!diag) {
#pragma unroll 1
                for (int jo = 0; jo < 4; ++jo) {
                    unsigned loc[kCR];
#pragma unroll
                    for (int r = 0; r < kCR; ++r) loc[r] = 0u;
#pragma unroll
                    for (int ji = 0; ji < 8; ++ji) {
                        const float4 s4 = tj[jo * 8 + ji];
#pragma unroll
                        for (int r = 0; r < kCR; ++r) GFN_COUNT_PAIR(r, 1u << ji, loc[r]);
                    }
#pragma unroll
                    for (int r = 0; r < kCR; ++r) amb[r] |= loc[r] << (jo * 8);
                }
            } else {
                // diagonal unit, 32 x 32 sub-blocks (row group r, chunk jc): above the diagonal (r < jc) everything, on it
                // (r == jc) surround sites right of the lane's row (j > i; j == i has dist 0), below it nothing (:126-129)
#pragma unroll
                for (int r = 0; r < kCR; ++r) {
                    if (r > jc) continue;
                    const int lt = r < jc ? -1 : lane;
#pragma unroll 1
                    for (int jo = 0; jo < 4; ++jo) {
                        unsigned loc = 0u;
                        const int tr = lt - jo * 8;
#pragma unroll
                        for (int ji = 0; ji < 8; ++ji) {
                            const float4 s4 = tj[jo * 8 + ji];
                            GFN_COUNT_PAIR_TRI(r, 1u << ji, loc, ji, tr);
                        }
                        amb[r] |= loc << (jo * 8);
                    }
                    // j == i itself: distance 0 between identical selections (never accepted), but a real pair with
                    // off_diag = 1 between two DIFFERENT selections of equal size (quirk Q1) - always through the exact pass
                    if (r == jc) amb[r] |= 1u << lane;
                }
            }
            if (selfcheck) {
                // every pair of the chunk exactly as well: a certain pair must have landed in the reference's bin.  The
                // fp32 classification is repeated here (same instructions, same result) without the count.
                for (int r = 0; r < kCR; ++r)
                    for (int jj = 0; jj < 32; ++jj) {
                        const float4 s4 = tj[jj];
                        const float ccx = r == 0 ? cx[0] : r == 1 ? cx[1] : r == 2 ? cx[2] : cx[3];
                        const float ccy = r == 0 ? cy[0] : r == 1 ? cy[1] : r == 2 ? cy[2] : cy[3];
                        const float ccz = r == 0 ? cz[0] : r == 1 ? cz[1] : r == 2 ? cz[2] : cz[3];
                        const float dx = s4.x - ccx, dy = s4.y - ccy, dz = s4.z - ccz;
                        const float vx = fabsf(dx) - hxf, vy = fabsf(dy) - hyf, vz = fabsf(dz) - hzf;
                        const float wx = hxf - fabsf(vx), wy = hyf - fabsf(vy), wz = hzf - fabsf(vz);
                        float t = wx * wx;
                        t = fmaf(wy, wy, t);
                        t = fmaf(wz, wz, t);
                        const float q = fmaf(sqrt_approx(t), invdx_f, nhminv_f);
                        const float bf = __fadd_rd(q, 8388608.0f);
                        const float g = q - (bf - 8388607.5f);
                        const float m = q - qmid;
                        const int row = i0 + 32 * r + lane, j = j0 + jc * 32 + jj;
                        if (diag && j == row) continue;              // always taken by the exact pass
                        const bool cond = !diag || j > row;
                        const bool cert = (fabsf(g) < e1) && (fabsf(m) <= hc) && cond;
                        const bool wide = (fabsf(m) <= hwid) && cond;
                        const int pos = exact_bin(P, bx, sa, sb, row, j);
                        const int fb = (int)(__float_as_uint(bf) - 0x4B000000u);
                        if (cert && pos != fb) ++n_bad;             // counted in a wrong bin (or counted although rejected)
                        if (!wide && pos >= 0) ++n_bad;             // dropped although the reference accepts it
                    }
            }
            // ---- exact pass over the chunk's ambiguous pairs: every lane walks its own four masks -----------------------
            unsigned m0 = amb[0], m1 = amb[1], m2 = amb[2], m3 = amb[3];
            while (__any_sync(0xffffffffu, (m0 | m1 | m2 | m3) != 0u)) {
                const int r = m0 ? 0 : (m1 ? 1 : (m2 ? 2 : 3));
                const unsigned m = m0 ? m0 : (m1 ? m1 : (m2 ? m2 : m3));
                if (m != 0u) {
                    const int jj = __ffs((int)m) - 1;
                    if (r == 0) m0 &= m0 - 1u; else if (r == 1) m1 &= m1 - 1u; else if (r == 2) m2 &= m2 - 1u; else m3 &= m3 - 1u;
                    const int row = i0 + 32 * r + lane, j = j0 + jc * 32 + jj;
                    const int pos = exact_bin(P, bx, sa, sb, row, j);
                    ++n_amb;
                    if (pos >= 0) {
                        ++n_amb_acc;
                        const unsigned w = (P.tri && row != j) ? 2u : 1u;                                    // :131-134
                        GFN_CHECK(P, jj >= 0 && jj < 32 && row < P.strideA && j < P.strideB);
                        if (pos < hn) { red_shared_u32(hraw + ((unsigned)pos << 2), w); if (TRI && w == 1u) ++n_w1; }   // :152
                        else {                                  // quirk Q3: the flat index runs into the next row / mode
                            ++n_ovf;
                            if (!(P.flags & GFN_FLAG_DISCARD_OVERFLOW)) global_add(P, 1, 0, listed ? __float_as_int(sa[row].fw) : row, pos, (double)w);
                        }
                    }
                }
            }
        }
#undef GFN_COUNT_PAIR
#undef GFN_COUNT_PAIR_TRI
#undef GFN_COUNT_HEAD
    }

    // ---- flush: replicas summed -> this CTA's partial histogram; the counters follow from the counts -----------------
    __syncthreads();
    unsigned long long sum = 0;
    double* out = P.partials + (size_t)blockIdx.x * hn;
    for (int k = tid; k < hn; k += kCW * 32) {
        unsigned long long s = 0;
        for (int r = 0; r < P.nrep; ++r) s += hist[r * hn + k];
        out[k] = (double)s;
        sum += s;
    }
    // Counters.  The histogram holds S = wt * (pairs accepted in the rows with weight wt) + (pairs accepted with weight 1:
    // i == j of two different selections in a triangular frame), so accepted-in-rows = (S + W1) / 2 for wt = 2, S for wt = 1 -
    // exact per CTA (its replicas hold exactly its own pairs).
    __shared__ unsigned long long red[5];
    if (tid < 5) red[tid] = 0ull;
    __syncthreads();
    unsigned long long rej = n_amb - n_amb_acc;
    for (int d = 16; d; d >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
        n_w1 += __shfl_xor_sync(0xffffffffu, n_w1, d);
        n_ovf += __shfl_xor_sync(0xffffffffu, n_ovf, d);
        rej += __shfl_xor_sync(0xffffffffu, rej, d);
        n_bad += __shfl_xor_sync(0xffffffffu, n_bad, d);
        n_amb += __shfl_xor_sync(0xffffffffu, n_amb, d);
    }
    if (lane == 0) {
        atomicAdd(&red[0], sum); atomicAdd(&red[1], n_w1); atomicAdd(&red[2], n_ovf); atomicAdd(&red[3], rej);
        if (n_bad) atomicAdd(P.counters + 4, n_bad);
        if (n_amb) atomicAdd(P.counters + 5, n_amb);
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned long long inr = (TRI ? (red[0] + red[1]) / 2ull : red[0]) + red[2];
        if (red[2]) atomicAdd(P.counters + 0, red[2]);
        if (inr + red[3]) atomicAdd(P.counters + 2, inr + red[3]);
        if (inr) atomicAdd(P.counters + 3, inr);
    }
}

// ------------------------------------------------------------------------------------------------
static size_t count_smem(int nrep, int hn) { return (size_t)kCW * kTJ * sizeof(float4) + (size_t)nrep * hn * sizeof(unsigned); }

// Serves g000-only handles with shared-memory histograms and an fp32-representable classification; fills cfg
// (dense = 2) and returns true, or leaves cfg untouched.
bool count_configure(int dev, int histo_n, int mode_sel, double invdx, double hmax, double box_max, int max_smem, int sms, PairConfig* cfg, cudaError_t* err)
{
    *err = cudaSuccess;
    if (mode_sel != 1 || cfg->filter_fp64 == 1 || cfg->shells || cfg->hist_mode != 0) return false;
    if (const char* env = getenv("GFN_NO_COUNT")) if (atoi(env)) return false;
    // expected EPS of a wrapped trajectory (M ~ box length): beyond a few per cent of a bin the exact pass dominates
    const double uu = 5.9604644775390625e-08;
    const double eps = 1.25 * (invdx * (1.7320508075688774 * uu * 13.5 * box_max + 32.0 * uu * hmax) + 2.0 * uu * (histo_n + 4.0));
    if (!(eps < 0.03) || histo_n > (1 << 22)) return false;
    int nrep = 0;
    if (count_smem(kCW, histo_n) <= (size_t)max_smem) nrep = kCW;
    else if (count_smem(1, histo_n) <= (size_t)max_smem) nrep = 1;
    if (!nrep) return false;
    PairConfig c = *cfg;
    c.dense = 2; c.dense_msel = 1;
    c.filter_fp64 = 0;                      // Site.fx/fy/fz hold the fp32-rounded coordinates
    c.warps = kCW; c.drain_warps = 0; c.nrep = nrep; c.nslot = 1;
    c.smem_bytes = (int)count_smem(nrep, histo_n);
    c.blocks_per_sm = 1; c.grid = sms;
    if ((*err = cudaFuncSetAttribute(count_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, c.smem_bytes)) != cudaSuccess) return false;
    if ((*err = cudaFuncSetAttribute(count_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c.smem_bytes)) != cudaSuccess) return false;
    *cfg = c;
    return true;
}

cudaError_t count_launch(const PairConfig& cfg, const PairParams& p, cudaStream_t s)
{
    if (p.tri) count_kernel<true><<<cfg.grid, kCW * 32, cfg.smem_bytes, s>>>(p);
    else       count_kernel<false><<<cfg.grid, kCW * 32, cfg.smem_bytes, s>>>(p);
    return cudaGetLastError();
}

}  // namespace gfn

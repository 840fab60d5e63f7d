// gfn_dense.cu - the pair kernel for DENSE neighbourhoods (a large share of all pairs lies inside histo_max:
// BASELINE configs[0] SPC/E water at hmax ~ L/2, 47 % in range; configs[1] ionic liquid, 36 %).
//
// Same arithmetic as pair_kernel (gfn_kernels.cu) - it replaces the same reference loop, cdef void calcHisto,
// gfunction.pyx:110-165 - but a different division of labour.  pair_kernel is built for sparse neighbourhoods (1.4 % in
// range at cfg3): half of its warps run the cheap range pre-filter, the other half evaluates the few survivors handed
// over through shared-memory rings.  When every second or third pair survives, the hand-off (per-survivor bit extraction
// and ring stores in the filter warps, two record gathers per survivor in the drain warps) costs more than the filter
// saves and the eight drain warps carry all the work.  Here every warp is the same:
//
//   * WORK UNIT = 64 core rows x one surround tile of 128 sites.  Units are enumerated (frame, surround tile, row
//     block) - for a triangular (n1 == n2) frame only the blocks that reach the diagonal - and a CTA owns a contiguous
//     range of them; its warps PULL units from a shared-memory counter, so warps never wait for each other and the
//     triangle balances itself.  Surround tiles stream through a 3-stage TMA ring (cp.async.bulk + mbarrier); a stage is
//     refilled by the warp that finishes the last unit of its tile.
//   * a lane OWNS two adjacent core rows of the unit and keeps their fp64 records (coordinates, dipole, 1/norm) in
//     REGISTERS, so the exact path never gathers a core record.
//   * per unit the warp runs the packed-fp16 (or fp32) range pre-filter of pair_kernel - two core rows per instruction -
//     and keeps the survivor masks in registers: 4 chunks x 2 rows x 32 bits.
//   * then every lane walks the survivors of its OWN two rows, one survivor of each row per step (two independent fp64
//     chains): the surround record is gathered from the shared-memory stage (four 128-bit loads per survivor; the
//     80-byte record stride spreads random rows over all bank groups), distance, bin and weights in the reference's
//     arithmetic.  These gathers make the shared-memory pipe the busiest unit of the kernel (profiles/r2/); the
//     alternative - going through the surround sites one by one with the record read as a broadcast and every lane
//     testing it against its two rows - has no gathers but idles the lanes whose rows did not survive: measured equal in
//     dense boxes and four times slower in the sparse units of a cell-sorted frame.
//   * ONE histogram per CTA: pair counts with native shared-memory integer atomics; orientation sums as 64-bit FIXED POINT
//     under a per-bin lock word (native exchange), stored in 8 bank-group stripes so that the read-modify-write of random
//     bins is conflict-free.  Integer sums are associative: any order gives the same bits, so the warps pull units
//     dynamically and share the histogram, and the result is reproducible.
//   * the acceptance test :149 is taken on d^2: acc_lo <= d^2 <= acc_hi with the exact thresholds the host finds by
//     bisection over the reference's own expressions (gfn_api.cu bin_thresholds), so rejected pairs never reach the
//     square root.  g000-only handles are not served here: gfn_count.cu decides their bins in fp32.
//
// Results: bins bit-exact, weighted sums within the same bound as pair_kernel's.
#include "gfn_device.cuh"

#include <stdio.h>
#include <stdlib.h>

namespace gfn {

constexpr int kDStages = 3;

struct DCtrl {
    uint64_t full[kDStages];     // TMA landed in stage s
    int done[4];                 // units finished on the tile held by stage s (the last one refills the stage)
    int next;                    // next unit of this CTA's range to hand out
    volatile int gen[3];         // ordinal (inside the CTA's range) of the tile stage s holds or is being loaded with
};
constexpr int kDCtrlBytes = 64;
static_assert(sizeof(DCtrl) <= kDCtrlBytes, "control block");

// weighted columns of a bin record for a compile-time mode set (MSEL = 0: chosen at run time, all six columns kept)
__host__ __device__ constexpr int dense_ncol(int msel)
{
    return msel == 0 ? 6 : ((msel >> 1) & 1) + ((msel >> 2) & 1) + ((msel >> 3) & 1) + ((msel >> 4) & 1) + ((msel >> 5) & 1) + ((msel >> 6) & 1);
}
// column of mode bit k (1..6) inside the record, or -1
__host__ __device__ constexpr int dense_col(int msel, int k)
{
    if (msel == 0) return k - 1;
    if (!((msel >> k) & 1)) return -1;
    int c = 0;
    for (int b = 1; b < k; ++b) c += (msel >> b) & 1;
    return c;
}

__host__ __device__ inline size_t pad16(size_t b) { return (b + 15) & ~(size_t)15; }
// ONE histogram per CTA, shared by all its warps:
//   sums   [bin][chunk][Q] 16-byte words: two int64 fixed-point columns (2^-43) per word, Q copies ("stripes") of every
//          word, copy q at the 16-byte slot q of its group.  A lane only ever touches stripe lane % Q: the 8 lanes of a
//          quarter warp - what one shared-memory wavefront of a 128-bit access serves - hit 8 different bank groups
//          whatever their bins are, so the read-modify-write of random bins is free of bank conflicts (with [bin][column]
//          records it replayed 2.2 times, and the shared-memory pipe was the kernel's limiter at 74 % of its peak)
//   count / lock words [bin][Q] u32: bit 31 = the stripe of the bin is being updated, bits 0-30 = pairs it has taken
__host__ __device__ inline size_t dense_hist_bytes(int hn, int msel, int Q)
{
    return (size_t)hn * ((size_t)(dense_ncol(msel) / 2) * Q * 16 + 4 * (size_t)Q);
}

// Orientation sums are kept as 64-bit fixed point, 2^-43 per unit: integer addition is associative, so a bin's sum does
// not depend on which warp or lane got there first - all warps of a CTA can share one histogram and pull their work
// dynamically, and the result is still reproducible bit for bit.  One addend is off by at most
// 2^-44 = 5.7e-14 (the bar is 1e-12 per unit of pair weight); |sum| < 2^20 per bin and stripe holds as long as no bin of a CTA
// takes 2^20 pairs in one launch: checked at the flush (error, not a wrap-around); launches are sized far below (dense_max_units).
constexpr double kFixScale = 8796093022208.0;              // 2^43
constexpr long long kFixPairs = (1ll << 20) - 1;           // accepted pairs one stripe of a bin may take per launch
__device__ __forceinline__ long long to_fixed(double v)
{
    // v * 2^43 + 1.5 * 2^52 has its integer part in the mantissa: |v| <= 2^7 keeps the exponent fixed
    return __double_as_longlong(fma(v, kFixScale, 6755399441055744.0)) - 0x4338000000000000ll;
}

template <typename F, int NW, int MSEL>
__global__ void __launch_bounds__(NW * 32, 1)
dense_kernel(const PairParams P)
{
    constexpr bool kH16 = std::is_same<F, __half>::value;
    static_assert(MSEL != 1, "g000-only handles are served by gfn_count.cu");
    constexpr int TJ = kTJ, S = kDStages;
    constexpr int NCOL = dense_ncol(MSEL);
    static_assert(TJ / 32 == 4, "four chunks per surround tile");
    static_assert(std::is_same<F, __half>::value || std::is_same<F, float>::value, "fp16x2 or fp32 pre-filter");

    extern __shared__ __align__(128) unsigned char smem[];
    Site* jsite = reinterpret_cast<Site*>(smem);                                           // [S][TJ]
    DCtrl* ctrl = reinterpret_cast<DCtrl*>(jsite + S * TJ);
    const int hn = P.histo_n;
    unsigned char* hist = reinterpret_cast<unsigned char*>(ctrl) + kDCtrlBytes;                      // [NW] replicas

    const int tid = threadIdx.x, lane = tid & 31;
    const int mode_sel = MSEL == 0 ? P.mode_sel : MSEL;
    const bool m_cs = mode_sel & (2 | 16), m_cd = mode_sel & (4 | 32), m_sd = mode_sel & (8 | 64);
    const double wt_d = P.tri ? 2.0 : 1.0;                   // off_diag of every accepted pair of this launch (:131-134)

    // this CTA's range of units and the surround tiles it touches (tile index gt = frame * n_jtiles + jt)
    const int upf = P.tiles_per_frame;                        // units per frame
    // cell-sorted frames (gfn_cells.cu): entry gt = frame * n_jtiles + jt lists the 64-row core blocks within histo_max of
    // surround tile jt (never empty), visit_start is the exclusive prefix of the list lengths = the flat unit sequence
    const bool listed = P.visit_start != nullptr;
    const long long nent = (long long)P.nframes * P.n_jtiles;
    long long total_units = listed ? P.visit_start[nent] : (long long)upf * P.nframes;
    if (listed && total_units > P.visit_cap) total_units = 0;       // the list buffer overflowed (flagged by gfn_cells.cu): only the flush runs
    const long long ubeg = total_units * blockIdx.x / gridDim.x;
    const long long uend = total_units * (blockIdx.x + 1) / gridDim.x;
    auto tile_of = [&](long long u, int& iu) -> long long {
        if (listed) {
            long long lo = 0, hi = nent;           // the last entry with start <= u (entries are never empty)
            while (hi - lo > 1) { const long long mid = (lo + hi) >> 1; if (P.visit_start[mid] <= u) lo = mid; else hi = mid; }
            iu = __ldg(P.visit_list + u);
            return lo;
        }
        const int f = (int)(u / upf);
        const int r = (int)(u - (long long)f * upf);
        int lo = 0, hi = P.n_jtiles;       // jt with tile_start[jt] <= r < tile_start[jt+1]
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (P.tile_start[mid] <= r) lo = mid; else hi = mid; }
        iu = r - P.tile_start[lo];
        return (long long)f * P.n_jtiles + lo;
    };
    auto units_in_range = [&](long long gt) -> int {          // units of tile gt that belong to this CTA
        long long b, e;
        if (listed) { b = P.visit_start[gt]; e = P.visit_start[gt + 1]; }
        else {
            const int f = (int)(gt / P.n_jtiles), jt = (int)(gt - (long long)f * P.n_jtiles);
            b = (long long)f * upf + P.tile_start[jt]; e = (long long)f * upf + P.tile_start[jt + 1];
        }
        if (b < ubeg) b = ubeg;
        if (e > uend) e = uend;
        return (int)(e - b);
    };
    int dummy_iu;
    const long long gt_first = ubeg < uend ? tile_of(ubeg, dummy_iu) : 0;
    const long long gt_last = ubeg < uend ? tile_of(uend - 1, dummy_iu) : -1;

    auto load_tile = [&](long long gt, int st) {
        const int f = (int)(gt / P.n_jtiles), jt = (int)(gt - (long long)f * P.n_jtiles);
        ctrl->gen[st] = (int)(gt - gt_first);          // a waiter first sees WHICH tile the stage is for, then waits for its bytes
        __threadfence_block();
        mbar_expect_tx(&ctrl->full[st], TJ * (uint32_t)sizeof(Site));
        bulk_g2s(jsite + st * TJ, P.siteB + (long long)f * P.strideB + (long long)jt * TJ, TJ * (uint32_t)sizeof(Site), &ctrl->full[st]);
    };
    const int Q = P.nrep;                                    // stripes (8, 4, 2 or 1)
    const size_t hbytes = dense_hist_bytes(hn, MSEL, Q);
    {
        unsigned* z = reinterpret_cast<unsigned*>(hist);
        for (size_t k = tid; k < hbytes / 4; k += NW * 32) z[k] = 0u;
    }
    if (tid < 4) ctrl->done[tid] = 0;
    if (tid < 3) ctrl->gen[tid] = -1;
    if (tid == 0) {
        ctrl->next = 0;
        for (int k = 0; k < S; ++k) mbar_init(&ctrl->full[k], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) for (int k = 0; k < S && gt_first + k <= gt_last; ++k) load_tile(gt_first + k, k);

    constexpr int NCH = NCOL / 2;                            // 16-byte words per bin
    const int q = lane & (Q - 1);
    longlong2* wsum = reinterpret_cast<longlong2*>(hist) + q;                                        // word (bin, c): [(bin * NCH + c) * Q]
    unsigned* lock = reinterpret_cast<unsigned*>(hist + (size_t)hn * NCH * Q * 16) + q;              // count / lock word (bin): [bin * Q]
    unsigned n_surv = 0, n_inr = 0, n_ovf = 0;
    bool zerodiv = false;
    const double acc_lo = P.acc_lo, acc_hi = P.acc_hi;

    // frame constants, refreshed when a unit belongs to another frame
    int f_cur = -1;
    bool exact_div = false;                                  // this frame: literal divisions (a norm is zero / out of range, or asked for)
    Box bx; bx.Lx = bx.Ly = bx.Lz = bx.hx = bx.hy = bx.hz = 0.0;
    float Tg = 0.f, hxf = 0.f, hyf = 0.f, hzf = 0.f;
    __half2 hTg, hLx, hLy, hLz;
    hTg = hLx = hLy = hLz = __halves2half2(__ushort_as_half(0), __ushort_as_half(0));

    for (;;) {
        // the warps PULL the units of the CTA's range from a shared-memory counter: nobody waits for a slower warp, the
        // triangle balances itself, and - the sums being integers - the result does not depend on who got which unit
        int un = 0;
        if (lane == 0) un = atomicAdd(&ctrl->next, 1);
        un = __shfl_sync(0xffffffffu, un, 0);
        const long long u = ubeg + un;
        if (u >= uend) break;
        int iu;
        const long long gt = tile_of(u, iu);
        const int f = (int)(gt / P.n_jtiles), jt = (int)(gt - (long long)f * P.n_jtiles);
        const long long k = gt - gt_first;                    // ordinal of the tile inside this CTA's range
        const int st = (int)(k % S);
        const int j0 = jt * TJ;
        const Site* siteA = P.siteA + (long long)f * P.strideA;

        if (f != f_cur) {
            f_cur = f;
            const double* b = P.box + 6ll * f;
            bx.Lx = b[0]; bx.Ly = b[1]; bx.Lz = b[2]; bx.hx = b[3]; bx.hy = b[4]; bx.hz = b[5];
            const double M = (double)fmaxf(P.maxA[f], P.maxB[f]) * (1.0 + 1.2e-7);
            // a dipole norm of this frame that is zero or outside [2^-300, 2^300] (prep_kernel flags it): the whole frame
            // takes the reference's literal divisions with the zero-denominator test of quirk Q5
            exact_div = (P.flags & GFN_FLAG_EXACT_DIV) || ((m_cs || m_cd) && P.nflagA[f]) || ((m_cs || m_sd) && P.nflagB[f]);
            // filter constants (same guard bands as pair_kernel)
            if (kH16) {
                const double Sb = fmax(b[0], fmax(b[1], b[2])), invS = 1.0 / Sb;
                const double uu = FOps<__half>::kU;
                const double hm = P.hmax * invS;
                const double ew = uu * (2.1 * M * invS + 4.2 * fmax(b[3], fmax(b[4], b[5])) * invS + 2.2 * hm);
                const double T = (hm * hm * (1.0 + 1e-12) + 3.5 * ew * hm + 3.0 * ew * ew + 4.0 * 5.9604644775390625e-08) * (1.0 + 8.0 * uu);
                __half th = __double2half(T);
                if (__half2float(th) <= (float)T) th = __ushort_as_half((unsigned short)(__half_as_ushort(th) + 1));   // round up
                hTg = __halves2half2(__hneg(th), __hneg(th));
                const __half a = __double2half(b[0] * invS), bq = __double2half(b[1] * invS), cq = __double2half(b[2] * invS);
                hLx = __halves2half2(a, a); hLy = __halves2half2(bq, bq); hLz = __halves2half2(cq, cq);
            } else {
                Tg = guard_threshold<float>(P.hmax, M, fmax(b[3], fmax(b[4], b[5])));
                hxf = (float)b[3]; hyf = (float)b[4]; hzf = (float)b[5];
            }
        }
        // ---- this lane's two core rows -----------------------------------------------------------------------------
        const int row0 = iu * 64 + 2 * lane;
        double cx[2], cy[2], cz[2], cpx[2], cpy[2], cpz[2], cn[2], icn[2];
        unsigned fxb[2], fyb[2], fzb[2]; int corig[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const double2* rp = reinterpret_cast<const double2*>(siteA + row0 + r);
            const double2 a0 = __ldg(rp), a1 = __ldg(rp + 1);
            cx[r] = a0.x; cy[r] = a0.y; cz[r] = a1.x; cpx[r] = a1.y;
            const double2 a2 = __ldg(rp + 2), a3 = __ldg(rp + 3);
            cpy[r] = a2.x; cpz[r] = a2.y; cn[r] = a3.x; icn[r] = a3.y;
            const uint4 f4 = __ldg(reinterpret_cast<const uint4*>(&siteA[row0 + r].fx));
            fxb[r] = f4.x; fyb[r] = f4.y; fzb[r] = f4.z; corig[r] = (int)f4.w;      // Site.fw: original index in cell-sorted frames
        }
        __half2 hcx, hcy, hcz;
        float fcx[2], fcy[2], fcz[2];
        {
            const unsigned px = __byte_perm(fxb[0], fxb[1], 0x5410), py = __byte_perm(fyb[0], fyb[1], 0x5410), pz = __byte_perm(fzb[0], fzb[1], 0x5410);
            hcx = *reinterpret_cast<const __half2*>(&px); hcy = *reinterpret_cast<const __half2*>(&py); hcz = *reinterpret_cast<const __half2*>(&pz);
#pragma unroll
            for (int r = 0; r < 2; ++r) { fcx[r] = __uint_as_float(fxb[r]); fcy[r] = __uint_as_float(fyb[r]); fcz[r] = __uint_as_float(fzb[r]); }
        }

        // the stage may still hold (or be waiting for) an EARLIER tile: with few units per tile the 16 warps run several tiles
        // ahead of the ring, and an mbarrier parity alone cannot tell a tile from the one two uses before it
        while (ctrl->gen[st] != (int)k) __nanosleep(40);
        mbar_wait(&ctrl->full[st], (unsigned)((k / S) & 1));
        const Site* js = jsite + st * TJ;

        // ---- range pre-filter: survivor masks of the unit, [chunk][row], bit 31-jj = surround site jj ---------------
        unsigned mk[4][2];
#pragma unroll
        for (int jc = 0; jc < 4; ++jc) {
            mk[jc][0] = 0u; mk[jc][1] = 0u;
            const int jbase = jc * 32;
            if (j0 + jbase >= P.n2) continue;                               // padding only
            if (P.tri && j0 + jbase + 31 < iu * 64) continue;              // chunk entirely left of the diagonal
            unsigned m0 = 0u, m1 = 0u;
            if (kH16) {
                // 9 HADD2 (with |.| modifiers) + 3 HFMA2 + SHF + LOP3 per TWO pairs, see pair_kernel
                unsigned acc[2] = {0u, 0u};
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                    const uint4 s4 = *reinterpret_cast<const uint4*>(&js[jbase + jj].fx);
                    const __half2 sx = *reinterpret_cast<const __half2*>(&s4.x), sy = *reinterpret_cast<const __half2*>(&s4.y), sz = *reinterpret_cast<const __half2*>(&s4.z);
                    const __half2 dx = __hsub2(sx, hcx), dy = __hsub2(sy, hcy), dz = __hsub2(sz, hcz);
                    const __half2 ax = __habs2(dx), ay = __habs2(dy), az = __habs2(dz);
                    const __half2 wx = __hmin2(ax, __hsub2(hLx, ax));
                    const __half2 wy = __hmin2(ay, __hsub2(hLy, ay));
                    const __half2 wz = __hmin2(az, __hsub2(hLz, az));
                    const __half2 tt = __hfma2(wz, wz, __hfma2(wy, wy, __hfma2(wx, wx, hTg)));
                    const unsigned tb = *reinterpret_cast<const unsigned*>(&tt);
                    acc[jj >> 4] |= (tb >> (jj & 15)) & (0x80008000u >> (jj & 15));
                }
                m0 = __byte_perm(acc[1], acc[0], 0x5410);
                m1 = __byte_perm(acc[1], acc[0], 0x7632);
            } else {
#pragma unroll 16
                for (int jj = 0; jj < 32; ++jj) {
                    const float4 s4 = *reinterpret_cast<const float4*>(&js[jbase + jj].fx);
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const float dx = s4.x - fcx[r], dy = s4.y - fcy[r], dz = s4.z - fcz[r];
                        const float vx = fabsf(dx) - hxf, vy = fabsf(dy) - hyf, vz = fabsf(dz) - hzf;
                        const float wx = hxf - fabsf(vx), wy = hyf - fabsf(vy), wz = hzf - fabsf(vz);
                        float tt = fmaf(wx, wx, -Tg);
                        tt = fmaf(wy, wy, tt);
                        tt = fmaf(wz, wz, tt);
                        if (r == 0) m0 = __funnelshift_l(__float_as_uint(tt), m0, 1);
                        else        m1 = __funnelshift_l(__float_as_uint(tt), m1, 1);
                    }
                }
            }
            // rows beyond the selection; for a triangular frame everything left of the diagonal (j < i, :126-129)
            if (row0 >= P.n1) m0 = 0u;
            if (row0 + 1 >= P.n1) m1 = 0u;
            if (P.tri) {
                const int k0 = row0 - (j0 + jbase);                         // first allowed jj of row 0 (row 1: one more)
                if (k0 > 0) m0 = k0 >= 32 ? 0u : (m0 & (0xffffffffu >> k0));
                if (k0 + 1 > 0) m1 = k0 + 1 >= 32 ? 0u : (m1 & (0xffffffffu >> (k0 + 1)));
            }
            mk[jc][0] = m0; mk[jc][1] = m1;
            n_surv += __popc(m0) + __popc(m1);
        }

        // ---- walk: every lane pops one survivor of EACH of its two rows per step (two independent fp64 chains).  The four
        //      chunk masks of a row form one stream (a shift register of words), so lanes wait for each other once per unit;
        //      a step costs the same whether one or both chains are live, and taking one survivor from each row needs
        //      max(c0, c1) steps where two per row would need ceil(c0/2) + ceil(c1/2) - the same in a dense unit, half in the
        //      sparse units a cell-sorted frame is full of.
        //      (Going through the surround sites one by one instead - the record read as a broadcast, every lane evaluating it
        //      against its two rows under the rows' filter bits - removes the gathers' bank conflicts but idles the lanes whose
        //      rows did not survive: measured equal in dense boxes, four times slower in sparse ones.)
        {
            unsigned ms[2][4];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int jc = 0; jc < 4; ++jc) ms[u][jc] = mk[jc][u];
            int jb[2] = {0, 0};
            while (__any_sync(0xffffffffu, (ms[0][0] | ms[0][1] | ms[0][2] | ms[0][3] | ms[1][0] | ms[1][1] | ms[1][2] | ms[1][3]) != 0u)) {
                int jg[2], sorig[2];
                double bpx[2], bpy[2], bpz[2], bn[2], ibn[2];
                double bx0[2], by0[2], bz0[2];
                bool have[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (ms[u][0] == 0u) { ms[u][0] = ms[u][1]; ms[u][1] = ms[u][2]; ms[u][2] = ms[u][3]; ms[u][3] = 0u; jb[u] += 32; }   // next chunk of the row's stream
                    const int fl = 31 - __clz(ms[u][0]);           // highest set bit = first surround site; -1: none left
                    have[u] = fl >= 0;
                    ms[u][0] &= ~(1u << (fl & 31));
                    const int jl = (jb[u] + 31 - fl) & (TJ - 1);
                    jg[u] = j0 + jl;
                    sorig[u] = listed ? __float_as_int(js[jl].fw) : 0;
                    const double2* bp = reinterpret_cast<const double2*>(js + jl);
                    const double2 b0 = bp[0], b1 = bp[1], b2 = bp[2], b3 = bp[3];   // x y | z px | py pz | norm 1/norm
                    bx0[u] = b0.x; by0[u] = b0.y; bz0[u] = b1.x;
                    bpx[u] = b1.y; bpy[u] = b2.x; bpz[u] = b2.y; bn[u] = b3.x; ibn[u] = b3.y;
                }
                bool ok[2]; int pos[2];
                double v[2][6];
                double dxa[2], dya[2], dza[2], d2[2], dist[2], idist[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    dxa[u] = wrap_signed(bx0[u], cx[u], bx.hx, bx.Lx);                       // :141-147
                    dya[u] = wrap_signed(by0[u], cy[u], bx.hy, bx.Ly);
                    dza[u] = wrap_signed(bz0[u], cz[u], bx.hz, bx.Lz);
                    d2[u] = DOT3(dxa[u], dya[u], dza[u], dxa[u], dya[u], dza[u]);
                    // :149 as an interval of d^2 (exact thresholds from the host; false for NaN).  Accepted values lie
                    // far inside the range where the straight-line square root below is exact (gfn_create checks
                    // histo_max and inv_dr), the others are replaced by 1.
                    ok[u] = have[u] && d2[u] >= acc_lo && d2[u] <= acc_hi;
                    if (!ok[u]) d2[u] = 1.0;
                }
                // --- distance :148, bin :150
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    dist[u] = sqrt_rsqrt_inrange(d2[u], idist[u]);
                    pos[u] = __double2int_rd(DMUL(DSUB(dist[u], P.hmin), P.invdx));
                }
                n_inr += (ok[0] ? 1u : 0u) + (ok[1] ? 1u : 0u);
                // --- orientation weights :154-165: reciprocal form (see pair_kernel) unless this frame holds a norm that
                //     is zero / out of range or the caller asked for the literal divisions
                double c[2][3];
                if (!exact_div) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        c[u][0] = DMUL(DOT3(cpx[u], cpy[u], cpz[u], bpx[u], bpy[u], bpz[u]), DMUL(icn[u], ibn[u]));
                        c[u][1] = DMUL(DOT3(cpx[u], cpy[u], cpz[u], dxa[u], dya[u], dza[u]), DMUL(icn[u], idist[u]));
                        c[u][2] = DMUL(DOT3(bpx[u], bpy[u], bpz[u], dxa[u], dya[u], dza[u]), DMUL(ibn[u], idist[u]));
                    }
                } else {
                    double den[2][3], num[2][3];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        den[u][0] = DMUL(cn[u], bn[u]);
                        den[u][1] = DMUL(cn[u], dist[u]);
                        den[u][2] = DMUL(bn[u], dist[u]);
                        num[u][0] = DOT3(cpx[u], cpy[u], cpz[u], bpx[u], bpy[u], bpz[u]);
                        num[u][1] = DOT3(cpx[u], cpy[u], cpz[u], dxa[u], dya[u], dza[u]);
                        num[u][2] = DOT3(bpx[u], bpy[u], bpz[u], dxa[u], dya[u], dza[u]);
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const bool sel = k == 0 ? m_cs : (k == 1 ? m_cd : m_sd);
                            if (ok[u] && sel && den[u][k] == 0.0) zerodiv = true;          // quirk Q5
                            if (!ok[u] || !sel || den[u][k] == 0.0) { den[u][k] = 1.0; num[u][k] = 0.0; }
                            c[u][k] = DDIV(num[u][k], den[u][k]);
                        }
                    }
                }
                // cell-sorted triangular frame: the pair was found by POSITION; the reference's core is the site with the smaller
                // ORIGINAL index (:126-130).  Exchanging the roles exchanges cos(core dipole, d) and cos(surround dipole, d) and
                // flips d: the same products, negated - exact.
                int rowo[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    rowo[u] = listed ? corig[u] : row0 + u;
                    if (listed && P.tri && sorig[u] < corig[u]) {
                        const double t = c[u][1];
                        c[u][1] = -c[u][2]; c[u][2] = -t;
                        rowo[u] = sorig[u];
                    }
                }
                // every accepted pair of a launch carries the same weight off_diag (:131-134: 2 in a triangular frame, else 1 -
                // but see i == j below): the sums are kept unweighted and multiplied at the flush (exact, a power of two)
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        v[u][k] = c[u][k];                                                       // incr
                        v[u][3 + k] = DSUB(DMUL(DMUL(1.5, c[u][k]), c[u][k]), 0.5);              // 1.5*incr*incr-0.5
                    }
                // ---- accumulate -------------------------------------------------------------------------------
                bool pend[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (ok[u] && P.tri && row0 + u == jg[u]) {
                        // i == j of two DIFFERENT selections of equal size (quirk Q1; identical selections have d = 0 here
                        // and were rejected): the one pair of the triangle with off_diag = 1 - straight to the global rows
                        if (pos[u] < hn) {
                            atomicAdd(P.gacc + 7ll * hn + pos[u], 1.0);
                            if (mode_sel & 1) atomicAdd(P.gacc + pos[u], 1.0);
#pragma unroll
                            for (int k = 1; k <= 6; ++k)
                                if (mode_sel & (1 << k)) atomicAdd(P.gacc + (long long)k * hn + pos[u], v[u][k - 1]);
                        } else {
                            ++n_ovf;
                            if (!(P.flags & GFN_FLAG_DISCARD_OVERFLOW)) {
                                if (mode_sel & 1) global_add(P, 1, 0, rowo[u], pos[u], 1.0);
#pragma unroll
                                for (int k = 1; k <= 6; ++k)
                                    if (mode_sel & (1 << k)) global_add(P, 1, k, rowo[u], pos[u], v[u][k - 1]);
                            }
                        }
                        ok[u] = false;
                    }
                    pend[u] = ok[u] && pos[u] < hn;
                    GFN_CHECK(P, !pend[u] || (pos[u] >= 0 && (size_t)(((size_t)pos[u] * NCH + NCH - 1) * Q + q) * 16 + 16 <= (size_t)hn * NCH * Q * 16));
                    GFN_CHECK(P, !have[u] || (jg[u] - j0 >= 0 && jg[u] - j0 < TJ && st < S));
                    if (ok[u] && pos[u] >= hn) {
                        // quirk Q3: the reference's flat index runs into the next row / mode; global accumulators
                        ++n_ovf;
                        if (!(P.flags & GFN_FLAG_DISCARD_OVERFLOW)) {
                            const double w = wt_d;
                            if (mode_sel & 1)  global_add(P, 1, 0, rowo[u], pos[u], w);
#pragma unroll
                            for (int k = 1; k <= 6; ++k)
                                if (mode_sel & (1 << k)) global_add(P, 1, k, rowo[u], pos[u], DMUL(v[u][k - 1], w));
                        }
                    }
                }
                // sums: the lane claims its stripe of the bin by setting bit 31 of the stripe's count word (native shared-memory
                // atomic OR: one winner among all lanes of the CTA that use this stripe), updates the sums with plain 128-bit
                // loads / stores and releases by storing the count + 1
                while (__any_sync(0xffffffffu, pend[0] || pend[1])) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        unsigned held = 0x80000000u;
                        if (pend[u]) held = atomicOr(lock + (size_t)pos[u] * Q, 0x80000000u);       // old value: the lock bit and the count
                        if (pend[u] && !(held & 0x80000000u)) {
                            // column k of the bin holds mode bit b with dense_col(MSEL, b) == k (a compile-time map;
                            // MSEL == 0: all six columns, unselected ones are never added to)
                            long long addv[NCOL];
#pragma unroll
                            for (int b = 1; b <= 6; ++b) {
                                const int col = dense_col(MSEL, b);
                                if (col >= 0) addv[col] = (MSEL != 0 || (mode_sel & (1 << b))) ? to_fixed(v[u][b - 1]) : 0ll;
                            }
                            longlong2* h2 = wsum + (size_t)pos[u] * NCH * Q;
                            longlong2 hv[NCH];
#pragma unroll
                            for (int c2 = 0; c2 < NCH; ++c2) hv[c2] = h2[c2 * Q];
#pragma unroll
                            for (int c2 = 0; c2 < NCH; ++c2) { hv[c2].x += addv[2 * c2]; hv[c2].y += addv[2 * c2 + 1]; }
#pragma unroll
                            for (int c2 = 0; c2 < NCH; ++c2) h2[c2 * Q] = hv[c2];
                            __threadfence_block();
                            *reinterpret_cast<volatile unsigned*>(lock + (size_t)pos[u] * Q) = held + 1u;      // :152 and the weight row; releases
                            pend[u] = false;
                        }
                    }
                }
            }
        }
        // ---- unit finished; the warp that finishes the last unit of a tile refills its stage with the tile S ahead ---
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            const int d = atomicAdd(&ctrl->done[st], 1) + 1;
            if (d == units_in_range(gt)) {
                atomicExch(&ctrl->done[st], 0);
                if (gt + S <= gt_last) load_tile(gt + S, st);
            }
        }
    }

    // counters: warp-reduce then one atomic per warp
    for (int d = 16; d; d >>= 1) {
        n_surv += __shfl_xor_sync(0xffffffffu, n_surv, d);
        n_inr  += __shfl_xor_sync(0xffffffffu, n_inr, d);
        n_ovf  += __shfl_xor_sync(0xffffffffu, n_ovf, d);
    }
    const bool anyz = __any_sync(0xffffffffu, zerodiv);
    if (lane == 0) {
        if (n_surv) atomicAdd(P.counters + 2, (unsigned long long)n_surv);
        if (n_inr)  atomicAdd(P.counters + 3, (unsigned long long)n_inr);
        if (n_ovf)  atomicAdd(P.counters + 0, (unsigned long long)n_ovf);
        if (anyz)   atomicOr(P.counters + 1, 1ull);
    }

    // ---- flush: stripes summed -> this CTA's partial histogram (layout of pair_kernel) -----------------------------
    __syncthreads();
    {
        constexpr int PS = 8;
        double* out = P.partials + (size_t)blockIdx.x * hn * PS;
        const long long* sums = reinterpret_cast<const long long*>(hist);
        const unsigned* counts = reinterpret_cast<const unsigned*>(hist + (size_t)hn * NCH * Q * 16);      // [bin][Q]
        for (int k = tid; k < hn * PS; k += NW * 32) {
            const int bin = k / PS, sl = k - bin * PS;
            double s = 0.0;
            int col = -1;
            if (sl >= 1 && sl <= 6) col = MSEL == 0 ? sl - 1 : dense_col(MSEL, sl);
            if (sl == 0 || sl == 7) {
                if (sl == 7 || (mode_sel & 1)) {
                    unsigned long long c = 0;
                    for (int z = 0; z < Q; ++z) c += counts[(size_t)bin * Q + z];
                    s = (double)c;
                }
            } else if (col >= 0) {
                double acc = 0.0;                                  // each stripe < 2^63: add as doubles, in stripe order
                for (int z = 0; z < Q; ++z) acc += (double)sums[(((size_t)bin * NCH + (col >> 1)) * Q + z) * 2 + (col & 1)];
                s = acc * (1.0 / kFixScale);
            }
            out[k] = s * wt_d;
        }
        // the fixed-point range: a stripe of a bin holds exactly as many addends as its count word says
        for (int k = tid; k < hn * Q; k += NW * 32)
            if (counts[k] > (unsigned)kFixPairs) atomicOr(P.counters + 1, 4ull);
    }
}

// ------------------------------------------------------------------------------------------------
// variant selection
// ------------------------------------------------------------------------------------------------
typedef void (*pair_fn_t)(const PairParams);

constexpr int kDW = 16;             // warps per CTA
long long dense_max_units(const PairConfig& cfg);

static size_t dense_smem(int Q, int hn, int msel)
{
    return (size_t)kDStages * kTJ * sizeof(Site) + kDCtrlBytes + dense_hist_bytes(hn, msel, Q);
}

template <typename F>
static pair_fn_t dense_pick_t(int msel)
{
    switch (msel) {
        case 127: return dense_kernel<F, kDW, 127>;
        case 31:  return dense_kernel<F, kDW, 31>;
        default:  return dense_kernel<F, kDW, 0>;
    }
}

static pair_fn_t dense_pick(const PairConfig& c)
{
    return c.filter_fp64 == 2 ? dense_pick_t<__half>(c.dense_msel) : dense_pick_t<float>(c.dense_msel);
}

// Fills the dense fields of cfg (dense, dense_msel, warps, nrep, smem_bytes, nslot, grid) when the dense kernel can serve
// this handle; returns false (cfg untouched) otherwise.
long long dense_units_per_frame(int n1, int n2, int tri)
{
    const long long ni = (n1 + 63) / 64, nj = (n2 + kTJ - 1) / kTJ;
    if (!tri) return ni * nj;
    long long u = 0;
    for (long long jt = 0; jt < nj; ++jt) u += (2 * jt + 2 < ni) ? 2 * jt + 2 : ni;     // row blocks that reach the tile's last column
    return u;
}

bool dense_configure(int dev, int histo_n, int mode_sel, int max_smem, int sms, const HandleGeom& geom, PairConfig* cfg, cudaError_t* err)
{
    *err = cudaSuccess;
    if (cfg->filter_fp64 == 1 || cfg->shells || cfg->hist_mode != 0) return false;
    if (mode_sel == 0 || mode_sel == 1) return false;     // g000 only: gfn_count.cu (or the ring kernel)
    const int msel = mode_sel == 127 ? 127 : (mode_sel == 31 ? 31 : 0);
    // stripes of the CTA's histogram: 8 make its updates conflict-free, fewer when 8 do not fit
    int nrep = 0;
    for (int cand = 8; cand >= 1 && !nrep; cand >>= 1)
        if (dense_smem(cand, histo_n, msel) <= (size_t)max_smem) nrep = cand;
    if (const char* env = getenv("GFN_DENSE_STRIPES")) {
        const int v = atoi(env);
        if ((v == 1 || v == 2 || v == 4 || v == 8) && dense_smem(v, histo_n, msel) <= (size_t)max_smem) nrep = v;
    }
    if (!nrep) return false;                          // the histogram does not fit: the ring kernel takes the handle
    PairConfig c = *cfg;
    c.dense = 1; c.dense_msel = msel;
    c.warps = kDW; c.drain_warps = 0; c.nrep = nrep;
    c.nslot = 8;
    c.smem_bytes = (int)dense_smem(nrep, histo_n, msel);
    c.blocks_per_sm = 1; c.grid = sms;
    {
        // a single frame must fit a launch; under tile skipping only the listed units count (twice the expected share)
        double units = (double)dense_units_per_frame(geom.n1, geom.n2, geom.tri);
        if (geom.cells_share > 0.0) units *= fmin(1.0, 2.0 * geom.cells_share);
        if ((double)dense_max_units(c) < units) return false;
    }
    pair_fn_t fn = dense_pick(c);
    if ((*err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, c.smem_bytes)) != cudaSuccess) return false;
    *cfg = c;
    return true;
}

// Units (64 core rows x 128 surround sites) per launch: small enough that a CTA's 2^20-pairs-per-bin limit is out of reach
// for any but pathological frames (a CTA then sees at most 2^27 pairs; a bin of a uniform box takes < 1 % of them).
long long dense_max_units(const PairConfig& cfg)
{
    return (1ll << 14) * cfg.grid;
}

cudaError_t dense_launch(const PairConfig& cfg, const PairParams& p, cudaStream_t s)
{
    pair_fn_t fn = dense_pick(cfg);
    fn<<<cfg.grid, cfg.warps * 32, cfg.smem_bytes, s>>>(p);
    return cudaGetLastError();
}

}  // namespace gfn

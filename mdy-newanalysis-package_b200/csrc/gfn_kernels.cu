// gfn_kernels.cu - hand-written sm_100a kernels of the g-function pair path.
//
// Replaces cdef void calcHisto (gfunction.pyx:110-165) and RDF._addHisto (:466-480) of the
// reference.  Compiled with -fmad=false: every fp64 expression that decides a bin or a weight is
// evaluated with the reference's own operation order and one rounding per operation; fused
// multiply-adds appear only where written explicitly (the range pre-filter, whose result never
// decides a bin).
//
// Kernel K1 (pair_kernel), one persistent CTA per SM slot:
//   work item = (frame, core tile of TI sites, segment of surround tiles).  The CTA keeps the core
//   tile's fp64 records in shared memory and each thread's IPT core positions in registers; surround
//   tiles of TJ sites stream through a 2-stage shared-memory ring filled by 1-D TMA bulk copies
//   (cp.async.bulk + mbarrier complete_tx).  Every thread range-tests its IPT x TJ pairs with a
//   branch-free pre-filter  w = h - | |dx| - h |  (equal in magnitude to the reference's
//   single-shift minimum image, :145-147),  t = fma(wz,wz,fma(wy,wy,fma(wx,wx,-Tg)));  the sign bit
//   of t is funnel-shifted into a 32-pair survivor mask.  Tg carries a proven guard band, so the
//   filter never rejects a pair the reference accepts.  Survivors are compacted into a per-warp
//   queue and evaluated 32 at a time, one lane each, with the exact reference arithmetic in fp64
//   (distance, the literal rejection test :149, bin :150, weights :152-165), then added into the
//   warp's private shared-memory histogram replica [bin][8 slots] - lanes that hit the same bin are
//   ordered with match.any, so no atomics are needed and the result is deterministic.
// Kernel K2 (finalize_kernel): block partials -> device accumulators in a fixed order.
// Kernel K0 (prep_kernel): user layout (n,3) fp64 -> Rec64 + float4 records, dipole norms, max |x|.
#include "gfn_kernels.cuh"
#include "../../include/gfn.h"

#include <stdio.h>
#include <stdlib.h>

namespace gfn {

// ------------------------------------------------------------------------------------------------
// PTX helpers (TMA 1-D bulk copy + mbarrier)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// K0: prep
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prep_kernel(const double* __restrict__ xyz, long long stride_xyz, const double* __restrict__ dip, long long stride_dip,
            Site* __restrict__ site, float* __restrict__ maxabs, int n, int n_pad)
{
    const int f = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_pad) return;
    Site r;
    float m = 0.f;
    if (s < n) {
        const double* c = xyz + (long long)f * stride_xyz + 3ll * s;
        r.x = c[0]; r.y = c[1]; r.z = c[2];
        if (dip) {
            const double* d = dip + (long long)f * stride_dip + 3ll * s;
            r.px = d[0]; r.py = d[1]; r.pz = d[2];
            // gfunction.pyx:125/:139  sqrt(px*px+py*py+pz*pz), left-to-right, un-fused
            r.norm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(r.px, r.px), __dmul_rn(r.py, r.py)), __dmul_rn(r.pz, r.pz)));
        } else {
            r.px = r.py = r.pz = 0.0; r.norm = 0.0;
        }
        r.spare = 0.0;
        r.fx = __double2float_rn(r.x); r.fy = __double2float_rn(r.y); r.fz = __double2float_rn(r.z); r.fw = 0.f;
        m = fmaxf(fmaxf(__double2float_ru(fabs(r.x)), __double2float_ru(fabs(r.y))), __double2float_ru(fabs(r.z)));
    } else {
        // padding: never passes the fp32 pre-filter against a finite site; the fp64 pre-test masks it by index
        r.x = r.y = r.z = r.px = r.py = r.pz = r.norm = r.spare = 0.0;
        r.fx = r.fy = r.fz = 1e30f; r.fw = 0.f;
    }
    site[(long long)f * n_pad + s] = r;
    unsigned mb = __reduce_max_sync(0xffffffffu, __float_as_uint(m));   // non-negative floats order like their bits
    if ((threadIdx.x & 31) == 0 && mb) atomicMax(reinterpret_cast<unsigned*>(maxabs) + f, mb);
}

cudaError_t prep_launch(const double* xyz, long long stride_xyz, const double* dip, long long stride_dip,
                        Site* site, float* maxabs, int n, int n_pad, int nframes, cudaStream_t s)
{
    for (int f0 = 0; f0 < nframes; f0 += 65535) {
        int nf = nframes - f0 < 65535 ? nframes - f0 : 65535;
        dim3 grid((n_pad + 255) / 256, nf);
        prep_kernel<<<grid, 256, 0, s>>>(xyz + (long long)f0 * stride_xyz, stride_xyz,
                                         dip ? dip + (long long)f0 * stride_dip : nullptr, stride_dip,
                                         site + (long long)f0 * n_pad, maxabs + f0, n, n_pad);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K1: pair kernel
// ------------------------------------------------------------------------------------------------
template <typename F> struct FOps;
template <> struct FOps<float> {
    static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
    static __device__ __forceinline__ unsigned signword(float a) { return __float_as_uint(a); }
    static constexpr double kU = 5.9604644775390625e-08;   // 2^-24
    static __device__ __forceinline__ float up(double t) { return __uint_as_float(__float_as_uint(__double2float_ru(t)) + 1u); }
};
template <> struct FOps<double> {
    static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
    static __device__ __forceinline__ double abs(double a) { return fabs(a); }
    static __device__ __forceinline__ unsigned signword(double a) { return (unsigned)__double2hiint(a); }
    static constexpr double kU = 1.1102230246251565e-16;   // 2^-53
    static __device__ __forceinline__ double up(double t) { return t * (1.0 + 1e-15); }
};

// Guard-banded squared cut-off of the pre-filter (see DESIGN.md "filter guard band" for the proof).
//   M        bound on |coordinate| of both selections in this frame, halfmax = max half box length
//   e_w      bound on the error of one filtered component  w  against the exact minimum-image component
//   accept   <=  fma-chain(w) - Tg < 0   for every pair with  d2_ref <= hmax^2 * (1 + 1e-12)
template <typename F>
__device__ __forceinline__ F guard_threshold(double hmax, double M, double halfmax)
{
    const double u = FOps<F>::kU;
    const double ew = u * (10.0 * M + 7.0 * halfmax);
    const double T = (hmax * hmax * (1.0 + 1e-12) + 3.5 * ew * hmax + 3.0 * ew * ew) * (1.0 + 16.0 * u);
    return FOps<F>::up(T);
}

struct Box { double Lx, Ly, Lz, hx, hy, hz; };

#define DMUL(a, b) __dmul_rn((a), (b))
#define DADD(a, b) __dadd_rn((a), (b))
#define DSUB(a, b) __dsub_rn((a), (b))
#define DDIV(a, b) __ddiv_rn((a), (b))
#define DOT3(ax, ay, az, bx, by, bz) DADD(DADD(DMUL(ax, bx), DMUL(ay, by)), DMUL(az, bz))

// Adds one value into the global accumulators, reproducing where the reference's flat index
// ix_mode + i*histo_n + pos lands (quirk Q3: pos may reach past the row).
__device__ __forceinline__ void global_add(const PairParams& P, int hist_mode, int k, int i, int pos, double v)
{
    const long long hn = P.histo_n;
    if (hist_mode == 2) {   // per-particle: literal reference index, bounds-checked
        const long long ix = (long long)k * P.n1 * hn + (long long)i * hn + pos;
        if (ix < 7ll * P.n1 * hn) atomicAdd(P.gpp + ix, v);
    }
    // compact [mode][bin]: the sum over rows of wherever the reference wrote
    const long long flat = (long long)i * hn + pos;
    const int kk = flat >= (long long)P.n1 * hn ? k + 1 : k;
    const int bin = pos >= hn ? (int)(pos - hn) : pos;
    if (kk < 7 && bin < hn) atomicAdd(P.gacc + (long long)kk * hn + bin, v);
}

// reference minimum image, gfunction.pyx:145-147:  if fabs(d) > half: d = d - sign(d)*L
// (|d| > half > 0 implies d != 0, so sign(d)*L == copysign(L, d) exactly; one rounding in the subtraction)
__device__ __forceinline__ double min_image(double d, double L, double half)
{
    const double t = DSUB(d, copysign(L, d));
    return fabs(d) > half ? t : d;
}

// 16-byte chunk c (0..3) of bin pos inside a replica with 8 slots per bin: chunks are XOR-swizzled
// with bits 1..2 of pos so that lanes holding different bins hit different 4-bank groups.
__device__ __forceinline__ int hist_chunk(int pos, int c) { return pos * 4 + (c ^ ((pos >> 1) & 3)); }

template <typename F, int NW, int NSLOT, int HIST>
__global__ void __launch_bounds__(NW * 32, 1)
pair_kernel(const PairParams P)
{
    constexpr int TI = NW * 32 * kIPT;
    constexpr int TJ = kTJ;
    constexpr int S = kStages;
    constexpr bool kF32 = sizeof(F) == 4;
    constexpr uint32_t kTileBytes = TJ * (uint32_t)sizeof(Site);

    extern __shared__ __align__(128) unsigned char smem[];
    Site* isite = reinterpret_cast<Site*>(smem);                                  // [TI]
    Site* jsite = isite + TI;                                                     // [S][TJ]
    unsigned short* queue = reinterpret_cast<unsigned short*>(jsite + S * TJ);    // [NW][kQCap]
    uint64_t* bars = reinterpret_cast<uint64_t*>(queue + NW * kQCap);             // full[S], ifull
    int* done_cnt = reinterpret_cast<int*>(bars + 4);                             // [S] warps finished with stage
    int* locks = done_cnt + 4;                                                    // [NW] replica locks
    double* hist = reinterpret_cast<double*>(locks + 20);                         // [nrep][histo_n*NSLOT] (HIST==0)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hn = P.histo_n;
    const int hsz = hn * NSLOT;
    const int nrep = P.nrep;
    const int rep = warp % nrep;
    const bool use_lock = nrep < NW;
    double* hist_w = hist + (size_t)rep * hsz;
    unsigned short* q = queue + warp * kQCap;

    if (HIST == 0) {
        for (int k = tid; k < nrep * hsz; k += NW * 32) hist[k] = 0.0;
    }
    if (tid < 4) done_cnt[tid] = 0;
    if (tid < NW) locks[tid] = 0;
    if (tid == 0) {
        for (int k = 0; k < 4; ++k) mbar_init(&bars[k], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int mode_sel = P.mode_sel;
    const bool m_cs = mode_sel & (2 | 16), m_cd = mode_sel & (4 | 32), m_sd = mode_sel & (8 | 64);
    unsigned tcount = 0;        // surround tiles consumed so far by this CTA (stage = tcount % S, parity = (tcount / S) & 1)
    unsigned icount = 0;        // items consumed (parity of the core-tile barrier)
    unsigned n_surv = 0, n_inr = 0, n_ovf = 0;
    bool zerodiv = false;

    const long long total_items = (long long)P.items_per_frame * P.nframes;
    for (long long item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int f = (int)(item / P.items_per_frame);
        const int r = (int)(item - (long long)f * P.items_per_frame);
        // core tile ti with item_start[ti] <= r < item_start[ti+1]
        int lo = 0, hi = P.n_itiles;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (P.item_start[mid] <= r) lo = mid; else hi = mid; }
        const int ti = lo;
        const int seg = r - P.item_start[ti];
        const int i0 = ti * TI;
        const int jt_first = P.tri ? (i0 / TJ) : 0;
        const int jt_begin = jt_first + seg * P.seg_tiles;
        const int jt_end = min(jt_begin + P.seg_tiles, P.n_jtiles);
        const int ntiles = jt_end - jt_begin;

        const Site* siteA = P.siteA + (long long)f * P.strideA;
        const Site* siteB = P.siteB + (long long)f * P.strideB;

        __syncthreads();     // previous item fully consumed: the core tile and every stage are free
        if (tid == 0) {
            mbar_expect_tx(&bars[3], TI * (uint32_t)sizeof(Site));
            bulk_g2s(isite, siteA + i0, TI * (uint32_t)sizeof(Site), &bars[3]);
            for (int pre = 0; pre < S && pre < ntiles; ++pre) {
                const int st = (tcount + pre) % S;
                mbar_expect_tx(&bars[st], kTileBytes);
                bulk_g2s(jsite + st * TJ, siteB + (long long)(jt_begin + pre) * TJ, kTileBytes, &bars[st]);
            }
        }

        // frame constants (uniform)
        Box bx;
        {
            const double* b = P.box + 6ll * f;
            bx.Lx = b[0]; bx.Ly = b[1]; bx.Lz = b[2]; bx.hx = b[3]; bx.hy = b[4]; bx.hz = b[5];
        }
        const double M = (double)fmaxf(P.maxA[f], P.maxB[f]) * (1.0 + 1.2e-7);
        const F Tg = guard_threshold<F>(P.hmax, M, fmax(bx.hx, fmax(bx.hy, bx.hz)));
        const F hxf = (F)bx.hx, hyf = (F)bx.hy, hzf = (F)bx.hz;

        // this thread's core positions: local index il(k) = warp*32*IPT + k*32 + lane
        F cx[kIPT], cy[kIPT], cz[kIPT];
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
            const int il = warp * (32 * kIPT) + k * 32 + lane;
            const Site* c = siteA + i0 + il;
            if (kF32) { cx[k] = (F)c->fx; cy[k] = (F)c->fy; cz[k] = (F)c->fz; }
            else if (i0 + il < P.n1) { cx[k] = (F)c->x; cy[k] = (F)c->y; cz[k] = (F)c->z; }
            else { cx[k] = cy[k] = cz[k] = (F)1e30; }
        }
        const int my_i_min = i0 + warp * (32 * kIPT);     // smallest core index of this warp

        mbar_wait(&bars[3], icount & 1);
        ++icount;

        int qcount = 0;   // warp-uniform
        for (int t = 0; t < ntiles; ++t) {
            const unsigned g = tcount + t;
            const int st = g % S;
            const int j0 = (jt_begin + t) * TJ;
            mbar_wait(&bars[st], (g / S) & 1);
            const Site* js = jsite + st * TJ;

            // ---- exact evaluation of up to 32 queued survivors (one per lane) -------------------------
            auto drain = [&](int cnt) {
                const bool have = lane < cnt;
                const unsigned e = have ? q[qcount - cnt + lane] : 0u;
                __syncwarp();          // all lanes have read their entry before anyone pushes again
                const int il = warp * (32 * kIPT) + (int)(e >> 7), jl = e & 127;
                const int i = i0 + il, j = j0 + jl;
                bool ok = have && i < P.n1 && j < P.n2 && !(P.tri && j < i);
                int pos = 0;
                double w = 1.0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0, v6 = 0;
                if (ok) {
                    const double2* ap = reinterpret_cast<const double2*>(isite + il);
                    const double2* bp = reinterpret_cast<const double2*>(js + jl);
                    const double2 a0 = ap[0], a1 = ap[1], a2 = ap[2], a3 = ap[3];   // x y | z px | py pz | norm -
                    const double2 b0 = bp[0], b1 = bp[1], b2 = bp[2], b3 = bp[3];
                    const double dx = min_image(DSUB(b0.x, a0.x), bx.Lx, bx.hx);                     // :141-147
                    const double dy = min_image(DSUB(b0.y, a0.y), bx.Ly, bx.hy);
                    const double dz = min_image(DSUB(b1.x, a1.x), bx.Lz, bx.hz);
                    const double dist = __dsqrt_rn(DOT3(dx, dy, dz, dx, dy, dz));                      // :148
                    ++n_surv;
                    // :149  dist<hmin | dist>hmax | floor(dist*invdx)==0  (dist >= 0, so floor(x)==0 <=> x<1); NaN rejected
                    if (!(dist >= P.hmin) || dist > P.hmax || DMUL(dist, P.invdx) < 1.0) {
                        ok = false;
                    } else {
                        pos = __double2int_rd(DMUL(DSUB(dist, P.hmin), P.invdx));                     // :150
                        w = (P.tri && i != j) ? 2.0 : 1.0;                                            // :131-134
                        ++n_inr;
                        const double apx = a1.y, apy = a2.x, apz = a2.y, an = a3.x;
                        const double bpx = b1.y, bpy = b2.x, bpz = b2.y, bn = b3.x;
                        if (m_cs) {                                                                   // :154-157
                            const double den = DMUL(an, bn);
                            if (den == 0.0) zerodiv = true;
                            else {
                                const double c = DDIV(DOT3(apx, apy, apz, bpx, bpy, bpz), den);
                                v1 = DMUL(c, w);
                                v4 = DMUL(DSUB(DMUL(DMUL(1.5, c), c), 0.5), w);
                            }
                        }
                        if (m_cd) {                                                                   // :158-161
                            const double den = DMUL(an, dist);
                            if (den == 0.0) zerodiv = true;
                            else {
                                const double c = DDIV(DOT3(apx, apy, apz, dx, dy, dz), den);
                                v2 = DMUL(c, w);
                                v5 = DMUL(DSUB(DMUL(DMUL(1.5, c), c), 0.5), w);
                            }
                        }
                        if (m_sd) {                                                                   // :162-165
                            const double den = DMUL(bn, dist);
                            if (den == 0.0) zerodiv = true;
                            else {
                                const double c = DDIV(DOT3(bpx, bpy, bpz, dx, dy, dz), den);
                                v3 = DMUL(c, w);
                                v6 = DMUL(DSUB(DMUL(DMUL(1.5, c), c), 0.5), w);
                            }
                        }
                    }
                }
                const double v0 = (mode_sel & 1) ? w : 0.0;                                           // :152
                if (!(mode_sel & 2)) v1 = 0.0;
                if (!(mode_sel & 4)) v2 = 0.0;
                if (!(mode_sel & 8)) v3 = 0.0;
                if (!(mode_sel & 16)) v4 = 0.0;
                if (!(mode_sel & 32)) v5 = 0.0;
                if (!(mode_sel & 64)) v6 = 0.0;

                const bool spill = ok && pos >= hn;          // quirk Q3
                if ((HIST != 0 && ok) || spill) {
                    if (spill) ++n_ovf;
                    if (!(spill && (P.flags & GFN_FLAG_DISCARD_OVERFLOW))) {
                        const int hm = HIST == 0 ? 1 : HIST;
                        if (mode_sel & 1)  global_add(P, hm, 0, i, pos, v0);
                        if (mode_sel & 2)  global_add(P, hm, 1, i, pos, v1);
                        if (mode_sel & 4)  global_add(P, hm, 2, i, pos, v2);
                        if (mode_sel & 8)  global_add(P, hm, 3, i, pos, v3);
                        if (mode_sel & 16) global_add(P, hm, 4, i, pos, v4);
                        if (mode_sel & 32) global_add(P, hm, 5, i, pos, v5);
                        if (mode_sel & 64) global_add(P, hm, 6, i, pos, v6);
                        if (!spill) atomicAdd(P.gacc + 7ll * hn + pos, w);      // pair weight row
                    }
                }
                if (HIST == 0) {
                    const bool add = ok && !spill;
                    const unsigned key = add ? (unsigned)pos : (0x40000000u | lane);
                    const unsigned grp = __match_any_sync(0xffffffffu, key);
                    const int rank = __popc(grp & ((1u << lane) - 1u));
                    const int maxr = __reduce_max_sync(0xffffffffu, add ? (unsigned)rank : 0u);
                    if (use_lock) {
                        if (lane == 0) { while (atomicCAS(&locks[rep], 0, 1) != 0) __nanosleep(40); }
                        __syncwarp();
                        __threadfence_block();
                    }
                    for (int rr = 0; rr <= maxr; ++rr) {
                        if (add && rank == rr) {
                            if (NSLOT == 1) {
                                hist_w[pos] += w;
                            } else {
                                double2* h2 = reinterpret_cast<double2*>(hist_w);
                                const int c0 = hist_chunk(pos, 0), c1 = hist_chunk(pos, 1), c2 = hist_chunk(pos, 2), c3 = hist_chunk(pos, 3);
                                double2 h0 = h2[c0], h1 = h2[c1], hh2 = h2[c2], h3 = h2[c3];
                                h0.x += v0; h0.y += v1; h1.x += v2; h1.y += v3;
                                hh2.x += v4; hh2.y += v5; h3.x += v6; h3.y += w;
                                h2[c0] = h0; h2[c1] = h1; h2[c2] = hh2; h2[c3] = h3;
                            }
                        }
                        __syncwarp();
                    }
                    if (use_lock) {
                        __threadfence_block();
                        __syncwarp();
                        if (lane == 0) atomicExch(&locks[rep], 0);
                    }
                }
                qcount -= cnt;
            };

            // ---- pre-filter over the TJ surround sites of this stage --------------------------------
            for (int jc = 0; jc < TJ / 32; ++jc) {
                const int jbase = jc * 32;
                if (j0 + jbase >= P.n2) break;                              // padding only
                if (P.tri && j0 + jbase + 31 < my_i_min) continue;         // chunk entirely below this warp's rows
                unsigned msk[kIPT];
#pragma unroll
                for (int k = 0; k < kIPT; ++k) msk[k] = 0u;
#pragma unroll 16
                for (int jj = 0; jj < 32; ++jj) {
                    F sx, sy, sz;
                    if (kF32) {
                        const float4 s = *reinterpret_cast<const float4*>(&js[jbase + jj].fx);
                        sx = (F)s.x; sy = (F)s.y; sz = (F)s.z;
                    } else {
                        const double2 sxy = *reinterpret_cast<const double2*>(&js[jbase + jj].x);
                        sx = (F)sxy.x; sy = (F)sxy.y; sz = (F)js[jbase + jj].z;
                        if (j0 + jbase + jj >= P.n2) { sx = sy = sz = (F)-1e30; }
                    }
#pragma unroll
                    for (int k = 0; k < kIPT; ++k) {
                        const F dx = sx - cx[k], dy = sy - cy[k], dz = sz - cz[k];
                        const F vx = FOps<F>::abs(dx) - hxf, vy = FOps<F>::abs(dy) - hyf, vz = FOps<F>::abs(dz) - hzf;
                        const F wx = hxf - FOps<F>::abs(vx), wy = hyf - FOps<F>::abs(vy), wz = hzf - FOps<F>::abs(vz);
                        F tt = FOps<F>::fma(wx, wx, -Tg);
                        tt = FOps<F>::fma(wy, wy, tt);
                        tt = FOps<F>::fma(wz, wz, tt);
                        msk[k] = __funnelshift_l(FOps<F>::signword(tt), msk[k], 1);    // first jj ends at bit 31
                    }
                }
                // ---- compact survivors into the warp queue (one packed scan for both core slots) ---------
                static_assert(kIPT == 2, "the packed scan below assumes two core sites per thread");
                if (!__any_sync(0xffffffffu, (msk[0] | msk[1]) != 0u)) continue;
                const unsigned cnt2 = (unsigned)__popc(msk[0]) | ((unsigned)__popc(msk[1]) << 16);
                unsigned incl = cnt2;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned up = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += up;
                }
                const unsigned tot2 = __shfl_sync(0xffffffffu, incl, 31);
                const unsigned excl = incl - cnt2;
#pragma unroll
                for (int k = 0; k < kIPT; ++k) {
                    const int total = (tot2 >> (16 * k)) & 0xffff;
                    if (total == 0) continue;
                    unsigned m = msk[k];
                    int at = qcount + (int)((excl >> (16 * k)) & 0xffff);
                    const unsigned tag = (unsigned)(k * 32 + lane) << 7;
                    while (m) {
                        const int jj = __clz(m);
                        m &= ~(0x80000000u >> jj);
                        q[at++] = (unsigned short)(tag | (unsigned)(jbase + jj));
                    }
                    qcount += total;
                    __syncwarp();
                    while (qcount >= 32) drain(32);
                }
            }
            if (qcount > 0) drain(qcount);   // the stage buffer is recycled next: empty the queue

            // ---- release the stage: the last warp to finish refills it -----------------------------------
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                const int old = atomicAdd(&done_cnt[st], 1);
                if (old == NW - 1) {
                    atomicExch(&done_cnt[st], 0);
                    if (t + S < ntiles) {
                        mbar_expect_tx(&bars[st], kTileBytes);
                        bulk_g2s(jsite + st * TJ, siteB + (long long)(jt_begin + t + S) * TJ, kTileBytes, &bars[st]);
                    }
                }
            }
        }
        tcount += ntiles;
    }

    // ---- flush -----------------------------------------------------------------------------------
    __syncthreads();
    if (HIST == 0) {
        double* out = P.partials + (size_t)blockIdx.x * hsz;
        for (int k = tid; k < hsz; k += NW * 32) {
            // k = bin*NSLOT + slot in the un-swizzled layout the finalize kernel expects
            int src = k;
            if (NSLOT == 8) { const int bin = k >> 3, sl = k & 7; src = hist_chunk(bin, sl >> 1) * 2 + (sl & 1); }
            double s = 0.0;
            for (int wv = 0; wv < nrep; ++wv) s += hist[(size_t)wv * hsz + src];
            out[k] = s;
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
}

// ------------------------------------------------------------------------------------------------
// K2: finalize
// ------------------------------------------------------------------------------------------------
__global__ void finalize_kernel(const double* __restrict__ partials, int grid, int hn, int nslot, double* __restrict__ gacc)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;    // over hn*nslot, layout [bin][slot]
    if (idx >= hn * nslot) return;
    const int bin = idx / nslot, slot = idx - bin * nslot;
    double s = 0.0;
    for (int b = 0; b < grid; ++b) s += partials[(size_t)b * hn * nslot + idx];
    if (nslot == 1) {                       // g000-only kernels: slot 0 is both g000 and the pair weight
        gacc[bin] += s;
        gacc[7ll * hn + bin] += s;
    } else {
        gacc[(long long)slot * hn + bin] += s;
    }
}

cudaError_t finalize_launch(const double* partials, int grid, int histo_n, int nslot, double* gacc, cudaStream_t s)
{
    const int n = histo_n * nslot;
    finalize_kernel<<<(n + 127) / 128, 128, 0, s>>>(partials, grid, histo_n, nslot, gacc);
    return cudaGetLastError();
}

__global__ void scale_kernel(double* a, long long n, double v)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += step) a[i] *= v;
}

cudaError_t scale_launch(double* a, long long n, double v, cudaStream_t s)
{
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    scale_kernel<<<(int)blocks, 256, 0, s>>>(a, n, v);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// variant selection
// ------------------------------------------------------------------------------------------------
static size_t smem_fixed_bytes(int nw)
{
    const size_t ti = (size_t)nw * 32 * kIPT;
    return ti * sizeof(Site) + (size_t)kStages * kTJ * sizeof(Site) + (size_t)nw * kQCap * sizeof(unsigned short)
         + 4 * sizeof(uint64_t) + 4 * sizeof(int) + 20 * sizeof(int);
}

typedef void (*pair_fn_t)(const PairParams);

template <typename F>
static pair_fn_t pick_fn(int hist_mode, int warps, int nslot)
{
    if (hist_mode == 0) {
        if (warps == 16) return nslot == 1 ? pair_kernel<F, 16, 1, 0> : pair_kernel<F, 16, 8, 0>;
        return nslot == 1 ? pair_kernel<F, 8, 1, 0> : pair_kernel<F, 8, 8, 0>;
    }
    if (hist_mode == 1) return pair_kernel<F, 16, 8, 1>;
    return pair_kernel<F, 16, 8, 2>;
}

static pair_fn_t pick(const PairConfig& c)
{
    return c.filter_fp64 ? pick_fn<double>(c.hist_mode, c.warps, c.nslot) : pick_fn<float>(c.hist_mode, c.warps, c.nslot);
}

int tile_i(const PairConfig& cfg) { return cfg.warps * 32 * kIPT; }

cudaError_t pair_configure(int dev, int histo_n, int mode_sel, unsigned flags, PairConfig* cfg)
{
    cudaError_t e;
    int max_smem = 0, sms = 0;
    if ((e = cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    cfg->filter_fp64 = (flags & GFN_FLAG_FILTER_FP64) ? 1 : 0;
    cfg->nslot = (mode_sel == 1) ? 1 : 8;
    cfg->nrep = 1;
    if (flags & GFN_FLAG_PER_PARTICLE) { cfg->hist_mode = 2; cfg->warps = 16; cfg->nslot = 8; }
    else if (flags & GFN_FLAG_HIST_GLOBAL) { cfg->hist_mode = 1; cfg->warps = 16; cfg->nslot = 8; }
    else {
        // per-replica bytes; replicas = as many as fit (<= warps); warps sharing a replica serialise on a lock
        const long long rb = (long long)histo_n * cfg->nslot * (long long)sizeof(double);
        const long long r16 = ((long long)max_smem - (long long)smem_fixed_bytes(16)) / rb;
        const long long r8  = ((long long)max_smem - (long long)smem_fixed_bytes(8)) / rb;
        cfg->hist_mode = 0;
        if (r16 >= 4)      { cfg->warps = 16; cfg->nrep = (int)(r16 > 16 ? 16 : r16); }
        else if (r8 >= 2)  { cfg->warps = 8;  cfg->nrep = (int)(r8 > 8 ? 8 : r8); }
        else if (r16 >= 1) { cfg->warps = 16; cfg->nrep = (int)r16; }
        else { cfg->hist_mode = 1; cfg->warps = 16; cfg->nslot = 8; }
        if (const char* env = getenv("GFN_WARPS")) {
            const int w = atoi(env);
            if (cfg->hist_mode == 0 && (w == 16 || w == 8)) {
                const long long rr = w == 16 ? r16 : r8;
                if (rr >= 1) { cfg->warps = w; cfg->nrep = (int)(rr > w ? w : rr); }
            }
        }
        if (const char* env = getenv("GFN_NREP")) {
            const int r = atoi(env);
            if (cfg->hist_mode == 0 && r >= 1 && r <= cfg->nrep) cfg->nrep = r;
        }
    }
    const bool sh = cfg->hist_mode == 0;
    cfg->smem_bytes = (int)(smem_fixed_bytes(cfg->warps) + (sh ? (size_t)cfg->nrep * histo_n * cfg->nslot * sizeof(double) : 0));
    pair_fn_t fn = pick(*cfg);
    if ((e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg->smem_bytes)) != cudaSuccess) return e;
    int occ = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, cfg->warps * 32, cfg->smem_bytes)) != cudaSuccess) return e;
    if (occ < 1) return cudaErrorInvalidConfiguration;
    cfg->blocks_per_sm = occ;
    cfg->grid = sms * occ;
    return cudaSuccess;
}

cudaError_t pair_launch(const PairConfig& cfg, const PairParams& p, cudaStream_t s)
{
    pair_fn_t fn = pick(cfg);
    long long items = (long long)p.items_per_frame * p.nframes;
    int grid = cfg.grid;
    if (cfg.hist_mode != 0 && items < grid) grid = (int)(items > 0 ? items : 1);   // partials need the full grid only in smem mode
    fn<<<grid, cfg.warps * 32, cfg.smem_bytes, s>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// pipe peak micro-benchmarks (roofline denominators)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters, T a, T b)
{
    T x0 = (T)threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = FOps<T>::fma(x0, a, b); x1 = FOps<T>::fma(x1, a, b); x2 = FOps<T>::fma(x2, a, b); x3 = FOps<T>::fma(x3, a, b);
            x4 = FOps<T>::fma(x4, a, b); x5 = FOps<T>::fma(x5, a, b); x6 = FOps<T>::fma(x6, a, b); x7 = FOps<T>::fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

template <typename T>
static cudaError_t measure_one(int sms, int iters, double* tflops)
{
    cudaError_t e;
    const int blocks = sms * 8, threads = 256;
    T* out = nullptr;
    if ((e = cudaMalloc(&out, sizeof(T) * blocks * threads)) != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fma_peak_kernel<T><<<blocks, threads>>>(out, iters / 8, (T)0.999999, (T)1e-6);     // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fma_peak_kernel<T><<<blocks, threads>>>(out, iters, (T)0.999999, (T)1e-6);
        cudaEventRecord(e1);
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) break;
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    if (e != cudaSuccess) return e;
    const double flops = 2.0 * 64.0 * (double)iters * blocks * threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    return cudaGetLastError();
}

cudaError_t peaks_measure(int dev, double* fp64_tflops, double* fp32_tflops)
{
    cudaError_t e;
    int sms = 0;
    if ((e = cudaSetDevice(dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if (fp64_tflops && (e = measure_one<double>(sms, 4096, fp64_tflops)) != cudaSuccess) return e;
    if (fp32_tflops && (e = measure_one<float>(sms, 8192, fp32_tflops)) != cudaSuccess) return e;
    return cudaSuccess;
}

}  // namespace gfn

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
#include "gfn_device.cuh"

#include <stdio.h>
#include <stdlib.h>

namespace gfn {

// ------------------------------------------------------------------------------------------------
// K0: prep
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prep_kernel(const double* __restrict__ xyz, long long stride_xyz, const double* __restrict__ dip, long long stride_dip,
            Site* __restrict__ site, float* __restrict__ maxabs, unsigned* __restrict__ nflag, int n, int n_pad, const double* __restrict__ box, int half_filter,
            unsigned long long* __restrict__ counters)
{
    const int f = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_pad) return;
    Site r;
    float m = 0.f;
    bool odd_norm = false;      // a dipole norm the reciprocal form of the cosines cannot take (zero: quirk Q5; outside [2^-300, 2^300])
    if (s < n) {
        const double* c = xyz + (long long)f * stride_xyz + 3ll * s;
        r.x = c[0]; r.y = c[1]; r.z = c[2];
        if (dip) {
            const double* d = dip + (long long)f * stride_dip + 3ll * s;
            r.px = d[0]; r.py = d[1]; r.pz = d[2];
            // gfunction.pyx:125/:139  sqrt(px*px+py*py+pz*pz), left-to-right, un-fused
            r.norm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(r.px, r.px), __dmul_rn(r.py, r.py)), __dmul_rn(r.pz, r.pz)));
            odd_norm = !mid_range(r.norm);
        } else {
            r.px = r.py = r.pz = 0.0; r.norm = 0.0;
        }
        r.spare = __drcp_rn(r.norm);       // 1/norm for the cosines (inf for a zero dipole: such sites take the exact division path)
        if (half_filter) {
            // fp16 filter: coordinates in units of the longest box edge, each duplicated into both halves of a
            // half2 so that one broadcast LDS.128 feeds the two core sites a filter thread keeps in one register
            const double* bb = box + 6ll * f;
            const double invS = 1.0 / fmax(bb[0], fmax(bb[1], bb[2]));
            // beyond ~3e4 box lengths fp16 overflows and the filter would no longer be conservative: refuse loudly
            if (fmax(fabs(r.x), fmax(fabs(r.y), fabs(r.z))) * invS > 30000.0) atomicOr(counters + 1, 2ull);
            const __half hx = __double2half(r.x * invS), hy = __double2half(r.y * invS), hz = __double2half(r.z * invS);
            const __half2 xx = __halves2half2(hx, hx), yy = __halves2half2(hy, hy), zz = __halves2half2(hz, hz);
            r.fx = __uint_as_float(*reinterpret_cast<const unsigned*>(&xx));
            r.fy = __uint_as_float(*reinterpret_cast<const unsigned*>(&yy));
            r.fz = __uint_as_float(*reinterpret_cast<const unsigned*>(&zz));
            r.fw = 0.f;
        } else {
            r.fx = __double2float_rn(r.x); r.fy = __double2float_rn(r.y); r.fz = __double2float_rn(r.z); r.fw = 0.f;
        }
        m = fmaxf(fmaxf(__double2float_ru(fabs(r.x)), __double2float_ru(fabs(r.y))), __double2float_ru(fabs(r.z)));
    } else {
        // padding: never passes the fp32 pre-filter against a finite site; the fp64 pre-test masks it by index
        r.x = r.y = r.z = r.px = r.py = r.pz = r.norm = r.spare = 0.0;
        r.fx = r.fy = r.fz = 1e30f; r.fw = 0.f;
        if (half_filter) r.fx = r.fy = r.fz = __uint_as_float(0x73537353u);     // half2(15000, 15000): squares overflow to +inf
    }
    site[(long long)f * n_pad + s] = r;
    unsigned mb = __reduce_max_sync(0xffffffffu, __float_as_uint(m));   // non-negative floats order like their bits
    if ((threadIdx.x & 31) == 0 && mb) atomicMax(reinterpret_cast<unsigned*>(maxabs) + f, mb);
    if (__any_sync(0xffffffffu, odd_norm) && (threadIdx.x & 31) == 0) atomicOr(nflag + f, 1u);
}

cudaError_t prep_launch(const double* xyz, long long stride_xyz, const double* dip, long long stride_dip,
                        Site* site, float* maxabs, unsigned* nflag, int n, int n_pad, int nframes, const double* box, int half_filter,
                        unsigned long long* counters, cudaStream_t s)
{
    for (int f0 = 0; f0 < nframes; f0 += 65535) {
        int nf = nframes - f0 < 65535 ? nframes - f0 : 65535;
        dim3 grid((n_pad + 255) / 256, nf);
        prep_kernel<<<grid, 256, 0, s>>>(xyz + (long long)f0 * stride_xyz, stride_xyz,
                                         dip ? dip + (long long)f0 * stride_dip : nullptr, stride_dip,
                                         site + (long long)f0 * n_pad, maxabs + f0, nflag + f0, n, n_pad, box + 6ll * f0, half_filter, counters);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K1: pair kernel
// ------------------------------------------------------------------------------------------------
// Shared-memory control block of one CTA (after the data arrays).
struct Ctrl {
    uint64_t full[kStages];      // TMA landed in stage s
    uint64_t ifull;              // TMA landed in the core tile
    int done_cnt[4];             // warps finished with stage s (the last one refills it)
    int locks[8];                // replica locks (only when fewer replicas than drain warps)
    volatile unsigned tail[16];  // ring w: entries published by filter warp w
    volatile unsigned head[16];  // ring w: entries consumed by its drain warp
    volatile unsigned tdone[16]; // ring w: number of surround tiles completely filtered
    volatile unsigned tend[16][4];   // ring w: tail at the end of tile g (slot g & 3)
};

// Shared-memory histogram replica of one drain warp (or of several under a lock):
//   wsum[bin][6]  fp64 sums of g110,g101,g011,g220,g202,g022 (48-byte bins: 48*pos mod 128 takes 8 values,
//                 so lanes holding different bins spread over all 4-bank groups)        -- only when WEIGHTS
//   cnt[bin]      u32 pair weight (1 or 2 per pair, native shared-memory integer atomics; it is g000 and the
//                 weight row at once, exact)
__host__ __device__ inline size_t replica_bytes(int hn, bool weights)
{
    size_t b = (size_t)hn * (weights ? 48 + 4 : 4);
    return (b + 15) & ~(size_t)15;
}

// The last warp (filter or drain) to finish with a stage refills it with the tile S steps ahead.
__device__ __forceinline__ void release_stage(Ctrl* c, Site* jsite, const Site* siteB, int st, int nwarps,
                                              int t, const Item& it, int lane)
{
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();
        const int old = atomicAdd(&c->done_cnt[st], 1);
        if (old == nwarps - 1) {
            atomicExch(&c->done_cnt[st], 0);
            if (t + kStages < it.ntiles) {
                mbar_expect_tx(&c->full[st], kTJ * (uint32_t)sizeof(Site));
                bulk_g2s(jsite + st * kTJ, siteB + (long long)run_tile(it, t + kStages) * kTJ, kTJ * (uint32_t)sizeof(Site), &c->full[st]);
            }
        }
    }
}

// Warp-specialised pair kernel.  Warps [0,NF) are FILTER warps: fp32 (or fp64) range pre-filter over
// 64 core sites x the surround stage, survivors published into a per-warp single-producer ring.
// Warps [NF,NF+4) are DRAIN warps (one per SM sub-partition): they consume the rings of filter warps
// w = d, d+4, d+8 in that fixed order, evaluate two survivors per lane exactly in fp64 and add them
// into their private histogram replica.  Registers are re-balanced with setmaxnreg.
template <typename F, int NF, int ND, bool WEIGHTS, int HIST, bool ALLM, bool SHELLS>
__global__ void __launch_bounds__((NF + ND) * 32, 1)
pair_kernel(const PairParams P)
{
    constexpr int NW = NF + ND;
    static_assert(ND == 4 || (ND == 8 && NF == 8) || (ND == 6 && NF == 12), "4 drain warps (rings d, d+4, d+8), 6 behind 12 (rings d, d+6), or 8 behind 8 filter warps (ring d)");
    constexpr int TI = NF * 32 * kIPT;
    constexpr int TJ = kTJ;
    constexpr int S = kStages;
    constexpr bool kF32 = std::is_same<F, float>::value;
    constexpr bool kH16 = std::is_same<F, __half>::value;
    constexpr unsigned RMASK = kRingCap - 1;
    static_assert(kIPT == 2, "two core sites per filter thread");
    constexpr int kU = kDrainILP;

    extern __shared__ __align__(128) unsigned char smem[];
    Site* isite = reinterpret_cast<Site*>(smem);                                  // [TI]
    Site* jsite = isite + TI;                                                     // [S][TJ]
    unsigned short* rings = reinterpret_cast<unsigned short*>(jsite + S * TJ);    // [NF][kRingCap]
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(rings + NF * kRingCap);
    unsigned char* hist = reinterpret_cast<unsigned char*>(ctrl) + kCtrlBytes;                         // [nrep] replicas

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hn = P.histo_n;
    const size_t rbytes = replica_bytes(hn, WEIGHTS);
    const int nrep = P.nrep;

    if (HIST == 0) {
        unsigned* z = reinterpret_cast<unsigned*>(hist);
        for (size_t k = tid; k < (size_t)nrep * rbytes / 4; k += NW * 32) z[k] = 0u;
    }
    if (tid < 16) { ctrl->tail[tid] = 0; ctrl->head[tid] = 0; ctrl->tdone[tid] = 0; }
    if (tid < 4) ctrl->done_cnt[tid] = 0;
    if (tid < 8) ctrl->locks[tid] = 0;
    if (tid == 0) {
        for (int k = 0; k < S; ++k) mbar_init(&ctrl->full[k], 1);
        mbar_init(&ctrl->ifull, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const long long total_tiles = total_visits(P);
    const long long slice_begin = total_tiles * blockIdx.x / gridDim.x;
    const long long slice_end = total_tiles * (blockIdx.x + 1) / gridDim.x;
    unsigned tcount = 0;        // surround tiles consumed so far by this CTA (stage = g % S, parity = (g / S) & 1)

    if (warp < NF) {
        // =========================================== FILTER ===========================================
#ifndef GFN_REGS_F
#define GFN_REGS_F "104"
#define GFN_REGS_D "152"
#endif
#ifndef GFN_REGS_F6
#define GFN_REGS_F6 "72"
#define GFN_REGS_D6 "128"
#endif
        if (ND == 8)   asm volatile("setmaxnreg.dec.sync.aligned.u32 " GFN_REGS_F ";");
        else if (ND == 6) asm volatile("setmaxnreg.dec.sync.aligned.u32 " GFN_REGS_F6 ";");    // experiment (profiles/exp_build.sh): 576 threads launch with 96 registers; 384*F + 192*D must stay BELOW 576*96
        else if (kF32 || kH16) asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
        else           asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        unsigned short* q = rings + warp * kRingCap;
        unsigned tail = 0, head_seen = 0;

        for (long long cur = slice_begin; cur < slice_end; ) {
            const Item it = decode_run(P, cur, slice_end, TI);
            cur += it.ntiles;
            const Site* siteA = P.siteA + (long long)it.f * P.strideA;
            const Site* siteB = P.siteB + (long long)it.f * P.strideB;

            __syncthreads();     // previous item fully consumed: the core tile and every stage are free
            if (tid == 0) {
                mbar_expect_tx(&ctrl->ifull, TI * (uint32_t)sizeof(Site));
                bulk_g2s(isite, siteA + it.i0, TI * (uint32_t)sizeof(Site), &ctrl->ifull);
                for (int pre = 0; pre < S && pre < it.ntiles; ++pre) {
                    const int st = (tcount + pre) % S;
                    mbar_expect_tx(&ctrl->full[st], TJ * (uint32_t)sizeof(Site));
                    bulk_g2s(jsite + st * TJ, siteB + (long long)run_tile(it, pre) * TJ, TJ * (uint32_t)sizeof(Site), &ctrl->full[st]);
                }
            }
            // frame constants (uniform)
            const double* b = P.box + 6ll * it.f;
            const double M = (double)fmaxf(P.maxA[it.f], P.maxB[it.f]) * (1.0 + 1.2e-7);
            // ---- filter constants and this thread's two core positions ------------------------------------
            // scalar filters (fp32 / fp64): Tg, half box lengths and positions in the filter type
            // fp16x2 filter: everything in units of the longest box edge S; the two core sites share a half2
            typedef typename std::conditional<kH16, float, F>::type FS;          // scalar type of the non-half paths
            FS Tg = 0, hxf = 0, hyf = 0, hzf = 0;
            FS cx[kIPT], cy[kIPT], cz[kIPT];
            __half2 hTg, hLx, hLy, hLz, hcx, hcy, hcz;
            if (kH16) {
                const double S = fmax(b[0], fmax(b[1], b[2])), invS = 1.0 / S;
                const double u = FOps<__half>::kU;
                // component error of the fp16 fold min(|d|, L - |d|) for pairs near the cut-off:
                //   e_w <= u (2 M' + 4 h' + 2 hmax') + O(u^2)   (primes: in units of S; 2M': rounding of the two
                //   coordinates; 2h' + hmax' >= |d'| of an in-range pair: rounding of d; 2h': rounding of L';
                //   hmax': rounding of L - |d| when it is not exact)
                const double hm = P.hmax * invS;
                const double ew = u * (2.1 * M * invS + 4.2 * fmax(b[3], fmax(b[4], b[5])) * invS + 2.2 * hm);
                const double T = (hm * hm * (1.0 + 1e-12) + 3.5 * ew * hm + 3.0 * ew * ew + 4.0 * 5.9604644775390625e-08) * (1.0 + 8.0 * u);
                __half th = __double2half(T);
                if (__half2float(th) <= (float)T) th = __ushort_as_half((unsigned short)(__half_as_ushort(th) + 1));   // round up
                hTg = __halves2half2(__hneg(th), __hneg(th));
                const __half a = __double2half(b[0] * invS), bq = __double2half(b[1] * invS), cq = __double2half(b[2] * invS);
                hLx = __halves2half2(a, a); hLy = __halves2half2(bq, bq); hLz = __halves2half2(cq, cq);
#ifdef GFN_EXP_SLOT_OLD
                const Site* c0 = siteA + it.i0 + warp * (32 * kIPT) + lane;
                const Site* c1 = c0 + 32;
#else
                const Site* c0 = siteA + it.i0 + warp * (32 * kIPT) + 2 * lane;     // core slots 2*lane, 2*lane + 1: adjacent records
                const Site* c1 = c0 + 1;
#endif
                const unsigned x0 = __float_as_uint(c0->fx), y0 = __float_as_uint(c0->fy), z0 = __float_as_uint(c0->fz);
                const unsigned x1 = __float_as_uint(c1->fx), y1 = __float_as_uint(c1->fy), z1 = __float_as_uint(c1->fz);
                const unsigned px = __byte_perm(x0, x1, 0x5410), py = __byte_perm(y0, y1, 0x5410), pz = __byte_perm(z0, z1, 0x5410);
                hcx = *reinterpret_cast<const __half2*>(&px); hcy = *reinterpret_cast<const __half2*>(&py); hcz = *reinterpret_cast<const __half2*>(&pz);
            } else {
                Tg = (FS)guard_threshold<FS>(P.hmax, M, fmax(b[3], fmax(b[4], b[5])));
                hxf = (FS)b[3]; hyf = (FS)b[4]; hzf = (FS)b[5];
#pragma unroll
                for (int k = 0; k < kIPT; ++k) {
#ifdef GFN_EXP_SLOT_OLD
                    const int il = warp * (32 * kIPT) + k * 32 + lane;
#else
                    const int il = warp * (32 * kIPT) + 2 * lane + k;
#endif
                    const Site* c = siteA + it.i0 + il;
                    if (kF32) { cx[k] = (FS)c->fx; cy[k] = (FS)c->fy; cz[k] = (FS)c->fz; }
                    else if (it.i0 + il < P.n1) { cx[k] = (FS)c->x; cy[k] = (FS)c->y; cz[k] = (FS)c->z; }
                    else { cx[k] = cy[k] = cz[k] = (FS)1e30; }
                }
            }
            const int my_i_min = it.i0 + warp * (32 * kIPT);     // smallest core index of this warp

            for (int t = 0; t < it.ntiles; ++t) {
                const unsigned g = tcount + t;
                const int st = g % S;
                const int j0 = run_tile(it, t) * TJ;
                mbar_wait(&ctrl->full[st], (g / S) & 1);
                const Site* js = jsite + st * TJ;

                unsigned mk0 = 0u, mk1 = 0u, mk2 = 0u, mk3 = 0u, mk4 = 0u, mk5 = 0u, mk6 = 0u, mk7 = 0u;   // masks of the tile: [chunk][core slot]
                for (int jc = 0; jc < TJ / 32; ++jc) {
                    const int jbase = jc * 32;
                    if (j0 + jbase >= P.n2) break;                              // padding only
                    if (P.tri && j0 + jbase + 31 < my_i_min) continue;         // chunk entirely below this warp's rows
                    unsigned m0 = 0u, m1 = 0u;
                    if (kH16) {
                        // 9 HADD2 (with |.| modifiers) + 3 HFMA2 + SHF + LOP3 per TWO pairs.  The two sign bits of a
                        // result (bit 15: core site 0, bit 31: core site 1) are shifted right by jj and OR-ed into an
                        // accumulator that covers 16 surround sites; the two accumulators of a chunk are re-packed
                        // into the usual masks (bit 31-jj = surround site jj) with two PRMTs.
                        unsigned acc[2] = {0u, 0u};
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) {
                            const uint4 s4 = *reinterpret_cast<const uint4*>(&js[jbase + jj].fx);
                            const __half2 sx = *reinterpret_cast<const __half2*>(&s4.x), sy = *reinterpret_cast<const __half2*>(&s4.y), sz = *reinterpret_cast<const __half2*>(&s4.z);
                            const __half2 dx = __hsub2(sx, hcx), dy = __hsub2(sy, hcy), dz = __hsub2(sz, hcz);
                            // |w| = min(|d|, L - |d|) squared equals the single-shift minimum image squared for every d
                            // (L - |d| < 0 beyond one box length, where the reference leaves |d| - L); L - |d| is exact
                            // whenever it is small (Sterbenz).  The min runs on the ALU pipe, off the fp16 pipe.
                            const __half2 ax = __habs2(dx), ay = __habs2(dy), az = __habs2(dz);
                            const __half2 wx = __hmin2(ax, __hsub2(hLx, ax));
                            const __half2 wy = __hmin2(ay, __hsub2(hLy, ay));
                            const __half2 wz = __hmin2(az, __hsub2(hLz, az));
                            const __half2 tt = __hfma2(wz, wz, __hfma2(wy, wy, __hfma2(wx, wx, hTg)));
                            const unsigned tb = *reinterpret_cast<const unsigned*>(&tt);
                            acc[jj >> 4] |= (tb >> (jj & 15)) & (0x80008000u >> (jj & 15));
                        }
                        m0 = __byte_perm(acc[1], acc[0], 0x5410);      // low halves: core site k = 0, sites 0-15 on top
                        m1 = __byte_perm(acc[1], acc[0], 0x7632);      // high halves: core site k = 1
                    } else {
#pragma unroll 16
                        for (int jj = 0; jj < 32; ++jj) {
                            FS sx, sy, sz;
                            if (kF32) {
                                const float4 s4 = *reinterpret_cast<const float4*>(&js[jbase + jj].fx);
                                sx = (FS)s4.x; sy = (FS)s4.y; sz = (FS)s4.z;
                            } else {
                                const double2 sxy = *reinterpret_cast<const double2*>(&js[jbase + jj].x);
                                sx = (FS)sxy.x; sy = (FS)sxy.y; sz = (FS)js[jbase + jj].z;
                                if (j0 + jbase + jj >= P.n2) { sx = sy = sz = (FS)-1e30; }
                            }
#pragma unroll
                            for (int k = 0; k < kIPT; ++k) {
                                const FS dx = sx - cx[k], dy = sy - cy[k], dz = sz - cz[k];
                                const FS vx = FOps<FS>::abs(dx) - hxf, vy = FOps<FS>::abs(dy) - hyf, vz = FOps<FS>::abs(dz) - hzf;
                                const FS wx = hxf - FOps<FS>::abs(vx), wy = hyf - FOps<FS>::abs(vy), wz = hzf - FOps<FS>::abs(vz);
                                FS tt = FOps<FS>::fma(wx, wx, -Tg);
                                tt = FOps<FS>::fma(wy, wy, tt);
                                tt = FOps<FS>::fma(wz, wz, tt);
                                if (k == 0) m0 = __funnelshift_l(FOps<FS>::signword(tt), m0, 1);    // first jj ends at bit 31
                                else        m1 = __funnelshift_l(FOps<FS>::signword(tt), m1, 1);
                            }
                        }
                    }
                    if (jc == 0) { mk0 = m0; mk1 = m1; } else if (jc == 1) { mk2 = m0; mk3 = m1; }
                    else if (jc == 2) { mk4 = m0; mk5 = m1; } else { mk6 = m0; mk7 = m1; }
                }
                // ---- publish the tile's survivors into this warp's ring, once per tile: one scan over the lanes' counts,
                //      then every lane writes its own entries (lane-major, then chunk, core slot, site: a fixed order, so
                //      the summation order stays deterministic).  Two branch-free extraction steps per mask cover almost
                //      every lane (0.5 survivors per mask and lane at cfg3); the rest loops. ----
                if (!(P.debug & 2)) {
                    static_assert(TJ / 32 == 4, "eight masks per tile");
                    const unsigned c0 = __popc(mk0), c1 = __popc(mk1), c2 = __popc(mk2), c3 = __popc(mk3);
                    const unsigned c4 = __popc(mk4), c5 = __popc(mk5), c6 = __popc(mk6), c7 = __popc(mk7);
                    const unsigned cnt = ((c0 + c1) + (c2 + c3)) + ((c4 + c5) + (c6 + c7));
                    unsigned incl = cnt;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const unsigned up = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += up;
                    }
                    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
#ifdef GFN_EXP_SLOT_OLD
                    constexpr unsigned kSlot1 = 32u << 7;
                    const unsigned tagl = ((unsigned)st << 13) | ((unsigned)lane << 7);     // stage | core slot | (site below)
#else
                    // core slot of mask k of this lane: 2*lane + k.  Consecutive ring entries then address adjacent core
                    // records, which the 80-byte stride spreads over distinct bank groups
                    constexpr unsigned kSlot1 = 1u << 7;
                    const unsigned tagl = ((unsigned)st << 13) | ((unsigned)(2 * lane) << 7);     // stage | core slot | (site below)
#endif
                    // one mask: entries from ring position `at` on
                    // (the first two survivors of a mask go out branch-free; `rest` collects what is left for emit_rest)
                    unsigned rest = 0u;
                    auto emit = [&](unsigned& m, unsigned tag, unsigned& at) {
#pragma unroll
                        for (int s2 = 0; s2 < 2; ++s2) {
                            const bool hv = m != 0u;
                            const int jj = __clz(m);
                            GFN_CHECK(P, !hv || (at - ctrl->head[warp]) < (unsigned)kRingCap);        // never over an unread entry
                            if (hv) q[at & RMASK] = (unsigned short)(tag + (unsigned)jj);
                            m = hv ? (m & ~(0x80000000u >> (jj & 31))) : 0u;
                            at += hv ? 1u : 0u;
                        }
                        rest |= m;
                    };
                    auto emit_rest = [&](unsigned m, unsigned tag, unsigned at) {
                        while (m) {
                            const int jj = __clz(m);
                            GFN_CHECK(P, (at - ctrl->head[warp]) < (unsigned)kRingCap);
                            q[at & RMASK] = (unsigned short)(tag + (unsigned)jj);
                            m &= ~(0x80000000u >> jj);
                            ++at;
                        }
                    };
                    if (total != 0u && total <= (unsigned)kRingCap / 2u) {
                        // a push must fit behind a partial batch the drain warp is still waiting to fill: at most half the ring
                        while (tail + total - head_seen > (unsigned)kRingCap) {      // ring full: wait for the drain warp
                            head_seen = ctrl->head[warp];
                            if (tail + total - head_seen > (unsigned)kRingCap) __nanosleep(64);
                        }
                        unsigned a0 = tail + (incl - cnt);
                        unsigned a1 = a0 + c0, a2 = a1 + c1, a3 = a2 + c2, a4 = a3 + c3, a5 = a4 + c4, a6 = a5 + c5, a7 = a6 + c6;
                        emit(mk0, tagl, a0);               emit(mk1, tagl + kSlot1, a1);
                        emit(mk2, tagl + 32u, a2);         emit(mk3, tagl + kSlot1 + 32u, a3);
                        emit(mk4, tagl + 64u, a4);         emit(mk5, tagl + kSlot1 + 64u, a5);
                        emit(mk6, tagl + 96u, a6);         emit(mk7, tagl + kSlot1 + 96u, a7);
                        if (__any_sync(0xffffffffu, rest != 0u)) {       // some lane holds three or more survivors of one mask
                            emit_rest(mk0, tagl, a0);               emit_rest(mk1, tagl + kSlot1, a1);
                            emit_rest(mk2, tagl + 32u, a2);         emit_rest(mk3, tagl + kSlot1 + 32u, a3);
                            emit_rest(mk4, tagl + 64u, a4);         emit_rest(mk5, tagl + kSlot1 + 64u, a5);
                            emit_rest(mk6, tagl + 96u, a6);         emit_rest(mk7, tagl + kSlot1 + 96u, a7);
                        }
                        tail += total;
                        __syncwarp();
                        if (lane == 0) { __threadfence_block(); ctrl->tail[warp] = tail; }
                    } else if (total != 0u) {
                        // dense tile: chunk by chunk (two masks, one scan); a chunk that would not fit behind a partial
                        // batch the drain warp may still be waiting to fill goes mask by mask (<= 32*32 entries each)
#pragma unroll 1
                        for (int c = 0; c < 4; ++c) {
                            const unsigned ma = c == 0 ? mk0 : c == 1 ? mk2 : c == 2 ? mk4 : mk6;
                            const unsigned mb = c == 0 ? mk1 : c == 1 ? mk3 : c == 2 ? mk5 : mk7;
                            const unsigned ca = __popc(ma), cb = __popc(mb);
                            const unsigned both = __reduce_add_sync(0xffffffffu, ca + cb);
                            if (both == 0u) continue;
                            const int npush = both <= (unsigned)kRingCap - 32u * kU ? 1 : 2;
                            for (int ps = 0; ps < npush; ++ps) {
                                const unsigned ci = npush == 1 ? ca + cb : (ps == 0 ? ca : cb);
                                unsigned inc = ci;
#pragma unroll
                                for (int d = 1; d < 32; d <<= 1) {
                                    const unsigned up = __shfl_up_sync(0xffffffffu, inc, d);
                                    if (lane >= d) inc += up;
                                }
                                const unsigned tot = __shfl_sync(0xffffffffu, inc, 31);
                                if (tot == 0u) continue;
                                while (tail + tot - head_seen > (unsigned)kRingCap) {
                                    head_seen = ctrl->head[warp];
                                    if (tail + tot - head_seen > (unsigned)kRingCap) __nanosleep(64);
                                }
                                const unsigned at = tail + (inc - ci);
                                const unsigned tg = tagl + (unsigned)c * 32u;
                                if (npush == 1) { emit_rest(ma, tg, at); emit_rest(mb, tg + kSlot1, at + ca); }
                                else if (ps == 0) emit_rest(ma, tg, at);
                                else emit_rest(mb, tg + kSlot1, at);
                                tail += tot;
                                __syncwarp();
                                if (lane == 0) { __threadfence_block(); ctrl->tail[warp] = tail; }
                            }
                        }
                    }
                }
                // ---- tile finished: tell the drain warp where it ends, release the stage ----------------
                __syncwarp();
                if (lane == 0) {
                    ctrl->tend[warp][g & 3] = tail;
                    __threadfence_block();
                    ctrl->tdone[warp] = g + 1;
                }
                release_stage(ctrl, jsite, siteB, st, NW, t, it, lane);
            }
            tcount += it.ntiles;
        }
    } else {
        // =========================================== DRAIN ============================================
        if (ND == 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 " GFN_REGS_D ";");
        else if (ND == 6) asm volatile("setmaxnreg.inc.sync.aligned.u32 " GFN_REGS_D6 ";");
        else         asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int dw = warp - NF;                   // drain warp: serves the rings of filter warps dw, dw+ND, ...
        const int rep = dw % nrep;
        const bool use_lock = nrep < ND;
        double* wsum = reinterpret_cast<double*>(hist + (size_t)rep * rbytes);
        unsigned* cnt = reinterpret_cast<unsigned*>(hist + (size_t)rep * rbytes + (WEIGHTS ? (size_t)hn * 48 : 0));
        const int mode_sel = ALLM ? 127 : P.mode_sel;       // ALLM: all seven modes selected, masks fold away
        const bool m_cs = mode_sel & (2 | 16), m_cd = mode_sel & (4 | 32), m_sd = mode_sel & (8 | 64);
        unsigned n_surv = 0, n_inr = 0, n_ovf = 0;
        bool zerodiv = false;
        const double t_small = 0.25 / (P.invdx * P.invdx);       // (0.5/invdx)^2

        unsigned icount = 0;
        for (long long cur = slice_begin; cur < slice_end; ) {
            const Item it = decode_run(P, cur, slice_end, TI);
            cur += it.ntiles;
            const Site* siteB = P.siteB + (long long)it.f * P.strideB;
            __syncthreads();
            Box bx;
            {
                const double* b = P.box + 6ll * it.f;
                bx.Lx = b[0]; bx.Ly = b[1]; bx.Lz = b[2]; bx.hx = b[3]; bx.hy = b[4]; bx.hz = b[5];
            }
            mbar_wait(&ctrl->ifull, icount & 1);
            ++icount;

            int j0s0 = 0, j0s1 = 0, j0s2 = 0;          // first surround index of the tile currently held by stage 0 / 1 / 2

            // ---- one batch: n <= 32*kU entries of ring fw starting at ring position head ------------------------
            auto process_batch = [&](const unsigned short* q, const int fw, const unsigned head, const int n) {
                        // ---- exact evaluation: kU survivors per lane, branch-free so that ptxas interleaves
                        //      the dependent fp64 chains of all of them ---------------------------------------
                        if (P.debug & 1) { __syncwarp(); if (lane == 0) ctrl->head[fw] = head + (unsigned)n; return; }
                        bool ok[kU]; int pos[kU], gi[kU], gj[kU]; double w[kU], v[kU][6];
                        double dxa[kU], dya[kU], dza[kU], d2[kU], dist[kU];
                        double apx[kU], apy[kU], apz[kU], an[kU], bpx[kU], bpy[kU], bpz[kU], bn[kU], ian[kU], ibn[kU], idist[kU];
                        bool self_pair[kU];
#pragma unroll
                        for (int u = 0; u < kU; ++u) {
                            const int idx = lane + 32 * u;
                            const bool have = idx < n;
                            const unsigned e = have ? q[(head + idx) & RMASK] : 0u;
                            const int il = fw * (32 * kIPT) + (int)((e >> 7) & 63u), jl = e & 127;
                            const int stg = (int)(e >> 13);
                            GFN_CHECK(P, stg < S && il < TI && (int)(head + idx - ctrl->tail[fw]) < 0 || !have);   // a published entry of a loaded stage
                            const Site* js = jsite + stg * TJ;
                            const int i = it.i0 + il, j = (stg == 0 ? j0s0 : (stg == 1 ? j0s1 : j0s2)) + jl;
                            gi[u] = i; if (SHELLS) gj[u] = j;
                            self_pair[u] = (i == j);
                            ok[u] = have && i < P.n1 && j < P.n2 && !(P.tri && j < i);
                            const double2* ap = reinterpret_cast<const double2*>(isite + il);
                            const double2* bp = reinterpret_cast<const double2*>(js + jl);
                            if (P.visit_start) {
                                // cell-sorted frame: i, j are positions in the sorted order, Site.fw keeps the original index.
                                // A triangular frame visits every unordered pair once BY POSITION; the reference's roles
                                // (core = the smaller original index, gfunction.pyx:126-130) are restored by exchanging the
                                // two records - same operands, same operations, same result.
                                const int oi = __float_as_int(isite[il].fw), oj = __float_as_int(js[jl].fw);
                                gi[u] = oi;
                                if (P.tri && oj < oi) { const double2* sw = ap; ap = bp; bp = sw; gi[u] = oj; }
                            }
                            const double2 a0 = ap[0], a1 = ap[1], a2 = ap[2], a3 = ap[3];   // x y | z px | py pz | norm -
                            const double2 b0 = bp[0], b1 = bp[1], b2 = bp[2], b3 = bp[3];
                            apx[u] = a1.y; apy[u] = a2.x; apz[u] = a2.y; an[u] = a3.x; ian[u] = a3.y;
                            bpx[u] = b1.y; bpy[u] = b2.x; bpz[u] = b2.y; bn[u] = b3.x; ibn[u] = b3.y;
                            dxa[u] = min_image(DSUB(b0.x, a0.x), bx.Lx, bx.hx);                     // :141-147
                            dya[u] = min_image(DSUB(b0.y, a0.y), bx.Ly, bx.hy);
                            dza[u] = min_image(DSUB(b1.x, a1.x), bx.Lz, bx.hz);
                            d2[u] = DOT3(dxa[u], dya[u], dza[u], dxa[u], dya[u], dza[u]);
                            // pairs closer than half the first bin edge are rejected by :149 whatever the rounding
                            // (floor(dist*invdx)==0); dropping them here keeps d2 = 0 (i == j) away from the sqrt
                            ok[u] = ok[u] && d2[u] >= t_small;           // false for NaN too
                        }
                        unsigned ns = 0;
#pragma unroll
                        for (int u = 0; u < kU; ++u) ns += ok[u] ? 1u : 0u;
                        n_surv += ns;
                        // --- distance :148
                        bool rng = true;
#pragma unroll
                        for (int u = 0; u < kU; ++u) { if (!ok[u]) d2[u] = 1.0; rng = rng && mid_range(d2[u]); }
                        if (__all_sync(0xffffffffu, rng)) {
#pragma unroll
                            for (int u = 0; u < kU; ++u) dist[u] = sqrt_rsqrt_inrange(d2[u], idist[u]);
                        } else {
#pragma unroll
                            for (int u = 0; u < kU; ++u) { dist[u] = __dsqrt_rn(d2[u]); idist[u] = __drcp_rn(dist[u]); }
                        }
                        // --- range test :149, bin :150, weight :131-134
#pragma unroll
                        for (int u = 0; u < kU; ++u) {
                            // dist<hmin | dist>hmax | floor(dist*invdx)==0  (dist >= 0: floor(x)==0 <=> x<1)
                            ok[u] = ok[u] && dist[u] >= P.hmin && !(dist[u] > P.hmax) && !(DMUL(dist[u], P.invdx) < 1.0);
                            pos[u] = __double2int_rd(DMUL(DSUB(dist[u], P.hmin), P.invdx));
                            w[u] = (P.tri && !self_pair[u]) ? 2.0 : 1.0;
                        }
                        if (SHELLS) {
                            // shell-resolved rows: the bin index gains the shell read from this frame's matrix
                            const int* map = P.shell_map + (long long)it.f * P.stride_map;
#pragma unroll
                            for (int u = 0; u < kU; ++u) {
                                const int r = gi[u] / P.apr1, c = gj[u] / P.apr2;
                                if (P.excl_self && r == c) ok[u] = false;                                  // :700
                                int sh = ok[u] ? __ldg(map + (long long)r * P.map_ld + c) : 1;             // :602
                                if (P.shell_mode == 1) {
                                    sh -= 1;                                                               // :603
                                    if (sh < 0 || sh >= P.nshells) sh = P.nshells - 1;                     // :604-605
                                    pos[u] = pos[u] * P.nshells + sh;                                      // :648-651
                                } else {
                                    sh = (sh > P.nshells ? P.nshells : sh) - 1;                            // :220-222
                                    if (sh < 0) sh += P.nshells;                                           // negative index wraps (Cython wraparound)
                                    if (sh < 0) { ok[u] = false; sh = 0; }
                                    pos[u] = sh * P.histo_nr + pos[u];                                     // :224
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < kU; ++u) n_inr += ok[u] ? 1u : 0u;
                        // shared-memory replicas: lanes that hit the same bin will take turns in lane order (deterministic,
                        // no fp64 atomics).  The ranking is started here so that MATCH/REDUX overlap the weight arithmetic.
                        bool addf[kU]; int rnk[kU], mxr[kU];
#pragma unroll
                        for (int u = 0; u < kU; ++u) {
                            addf[u] = ok[u] && pos[u] < hn;            // pos >= hn: quirk Q3, handled below
                            rnk[u] = 0; mxr[u] = 0;
                            if (HIST == 0 && WEIGHTS) {
                                const unsigned key = addf[u] ? (unsigned)pos[u] : (0x40000000u | lane);
                                const unsigned grp = __match_any_sync(0xffffffffu, key);
                                rnk[u] = __popc(grp & ((1u << lane) - 1u));
                                mxr[u] = (int)__reduce_max_sync(0xffffffffu, addf[u] ? (unsigned)rnk[u] : 0u);
                            }
                        }
                        if (WEIGHTS) {
                            // --- orientation weights :154-165, all three cosines (unselected ones are never added)
                            double c[kU][3];
                            // Usual case: every needed norm of the warp's in-range pairs lies in [2^-300, 2^300].  Then no
                            // denominator clen*slen, clen*dist, slen*dist can be zero or leave the normal range (quirk Q5
                            // cannot fire), and the cosines are formed with the reciprocals 1/norm (prep_kernel) and 1/dist
                            // (by-product of the root): 2 multiplications each, relative error < 4 ulp - far inside the
                            // 1e-12 bar for the weighted rows; bins never depend on them.
                            bool nrm = true;
#pragma unroll
                            for (int u = 0; u < kU; ++u)
                                nrm = nrm && (!ok[u] || ((!(m_cs || m_cd) || mid_range(an[u])) && (!(m_cs || m_sd) || mid_range(bn[u]))));
                            if (P.flags & GFN_FLAG_EXACT_DIV) nrm = false;      // the caller asked for the literal divisions
                            if (__all_sync(0xffffffffu, nrm)) {
#pragma unroll
                                for (int u = 0; u < kU; ++u) {
                                    c[u][0] = DMUL(DOT3(apx[u], apy[u], apz[u], bpx[u], bpy[u], bpz[u]), DMUL(ian[u], ibn[u]));
                                    c[u][1] = DMUL(DOT3(apx[u], apy[u], apz[u], dxa[u], dya[u], dza[u]), DMUL(ian[u], idist[u]));
                                    c[u][2] = DMUL(DOT3(bpx[u], bpy[u], bpz[u], dxa[u], dya[u], dza[u]), DMUL(ibn[u], idist[u]));
                                }
                            } else {
                            // rare: IEEE divisions with the reference's zero-denominator semantics
                            double den[kU][3], num[kU][3];
#pragma unroll
                            for (int u = 0; u < kU; ++u) {
                                den[u][0] = DMUL(an[u], bn[u]);
                                den[u][1] = DMUL(an[u], dist[u]);
                                den[u][2] = DMUL(bn[u], dist[u]);
                                num[u][0] = DOT3(apx[u], apy[u], apz[u], bpx[u], bpy[u], bpz[u]);
                                num[u][1] = DOT3(apx[u], apy[u], apz[u], dxa[u], dya[u], dza[u]);
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
#pragma unroll
                            for (int u = 0; u < kU; ++u)
#pragma unroll
                                for (int k = 0; k < 3; ++k) {
                                    v[u][k] = DMUL(c[u][k], w[u]);                                          // incr*off_diag
                                    v[u][3 + k] = DMUL(DSUB(DMUL(DMUL(1.5, c[u][k]), c[u][k]), 0.5), w[u]);   // (1.5*incr*incr-0.5)*off_diag
                                }
                        } else {
#pragma unroll
                            for (int u = 0; u < kU; ++u)
#pragma unroll
                                for (int k = 0; k < 6; ++k) v[u][k] = 0.0;
                        }
                        __syncwarp();
                        if (lane == 0) ctrl->head[fw] = head + (unsigned)n;     // entries are in registers: the producer may reuse them

                        // ---- accumulate ----------------------------------------------------------------------
                        bool slow = false;
#pragma unroll
                        for (int u = 0; u < kU; ++u) slow = slow || (ok[u] && (HIST != 0 || pos[u] >= hn));
                        if (HIST != 0 || __any_sync(0xffffffffu, slow)) {
                            // global accumulators: the per-particle / global-atomic variants, and quirk Q3 spills
#pragma unroll
                            for (int u = 0; u < kU; ++u) {
                                const bool spill = ok[u] && pos[u] >= hn;
                                if ((HIST != 0 && ok[u]) || spill) {
                                    const double v0 = (mode_sel & 1) ? w[u] : 0.0;                                // :152
                                    if (spill) ++n_ovf;
                                    if (!(spill && (P.flags & GFN_FLAG_DISCARD_OVERFLOW))) {
                                        const int hm = HIST == 0 ? 1 : HIST;
                                        if (mode_sel & 1)  global_add(P, hm, 0, gi[u], pos[u], v0);
                                        if (mode_sel & 2)  global_add(P, hm, 1, gi[u], pos[u], v[u][0]);
                                        if (mode_sel & 4)  global_add(P, hm, 2, gi[u], pos[u], v[u][1]);
                                        if (mode_sel & 8)  global_add(P, hm, 3, gi[u], pos[u], v[u][2]);
                                        if (mode_sel & 16) global_add(P, hm, 4, gi[u], pos[u], v[u][3]);
                                        if (mode_sel & 32) global_add(P, hm, 5, gi[u], pos[u], v[u][4]);
                                        if (mode_sel & 64) global_add(P, hm, 6, gi[u], pos[u], v[u][5]);
                                        if (!spill) atomicAdd(P.gacc + 7ll * hn + pos[u], w[u]);      // pair weight row
                                    }
                                }
                            }
                        }
                        if (HIST == 0) {
#pragma unroll
                            for (int u = 0; u < kU; ++u) {
                                GFN_CHECK(P, !addf[u] || (pos[u] >= 0 && pos[u] < hn));
                                if (addf[u]) atomicAdd(cnt + pos[u], (P.tri && !self_pair[u]) ? 2u : 1u);   // :152 and the weight row
                            }
                            if (WEIGHTS) {
                                if (use_lock) {
                                    if (lane == 0) { while (atomicCAS(&ctrl->locks[rep], 0, 1) != 0) __nanosleep(40); }
                                    __syncwarp();
                                    __threadfence_block();
                                }
#pragma unroll
                                for (int u = 0; u < kU; ++u) {
                                    double2* hp = reinterpret_cast<double2*>(wsum + (size_t)pos[u] * 6);
                                    for (int rr = 0; rr <= mxr[u]; ++rr) {
                                        if (addf[u] && rnk[u] == rr) {
                                            double2 h0 = hp[0], h1 = hp[1], h2 = hp[2];
                                            if (ALLM || (mode_sel & 2))  h0.x += v[u][0];       // g110
                                            if (ALLM || (mode_sel & 4))  h0.y += v[u][1];       // g101
                                            if (ALLM || (mode_sel & 8))  h1.x += v[u][2];       // g011
                                            if (ALLM || (mode_sel & 16)) h1.y += v[u][3];       // g220
                                            if (ALLM || (mode_sel & 32)) h2.x += v[u][4];       // g202
                                            if (ALLM || (mode_sel & 64)) h2.y += v[u][5];       // g022
                                            hp[0] = h0; hp[1] = h1; hp[2] = h2;
                                        }
                                        __syncwarp();
                                    }
                                }
                                if (use_lock) {
                                    __threadfence_block();
                                    __syncwarp();
                                    if (lane == 0) atomicExch(&ctrl->locks[rep], 0);
                                }
                            }
                        }
            };

            if (ND == 8) {
                // ---- one ring per drain warp: batches may span tile boundaries (entries carry their stage), so the
                //      only partial batches are the ones needed to let a stage go.  Tile t_rel is released once all
                //      its entries are evaluated; it must be flushed when the filter has finished the tile after it
                //      (it will soon need the stage) or the run is over. ------------------------------------------------
                const int fw = dw;
                const unsigned short* q = rings + fw * kRingCap;
                unsigned head = ctrl->head[fw];
                int t_wait = 0, t_rel = 0;
                unsigned seen_td = 0xffffffffu;     // tdone at which the full check below last found nothing to do
                for (;;) {
                    const unsigned td_raw = ctrl->tdone[fw];
                    if (td_raw == seen_td) {
                        // fast poll: with tdone and head unchanged, only a full batch can give this warp work.
                        // The back-off follows the deficit (survivors arrive every few tens of ns).
                        const unsigned a = ctrl->tail[fw] - head;
                        if (a < 32u * kU) { __nanosleep(min((32u * kU - a) * P.poll_ns, 4000u)); continue; }
                    }
                    seen_td = 0xffffffffu;
                    int td = (int)(td_raw - tcount);                    // tiles of this run the filter has finished
                    td = td < 0 ? 0 : (td > it.ntiles ? it.ntiles : td);
                    __threadfence_block();
                    const unsigned tail_now = ctrl->tail[fw];
                    const unsigned avail = tail_now - head;
                    // stage data of every tile that may have entries in the ring: finished ones and the one in progress
                    // (never beyond t_rel + S - 1: a later tile is only loaded after this warp has released t_rel)
                    while (t_wait < it.ntiles && t_wait <= td && t_wait < t_rel + S) {
                        const unsigned g = tcount + t_wait;
                        mbar_wait(&ctrl->full[g % S], (g / S) & 1);
                        const int j0 = run_tile(it, t_wait) * TJ;
                        if (g % S == 0) j0s0 = j0; else if (g % S == 1) j0s1 = j0; else j0s2 = j0;
                        ++t_wait;
                    }
                    bool flush = false;
                    if (t_rel < it.ntiles && td > t_rel && (td >= t_rel + 2 || td == it.ntiles))
                        flush = (int)(ctrl->tend[fw][(tcount + t_rel) & 3] - head) > 0;
                    if (avail >= 32u * kU || (flush && avail > 0u)) {
                        __threadfence_block();
                        const int n = (int)min(avail, 32u * kU);
                        process_batch(q, fw, head, n);
                        head += (unsigned)n;
                    } else if (!(t_rel < it.ntiles && td > t_rel && (int)(head - ctrl->tend[fw][(tcount + t_rel) & 3]) >= 0)) {
                        if (t_rel == it.ntiles) break;
                        seen_td = td_raw;
                        __nanosleep(min((32u * kU - min(avail, 32u * kU - 1u)) * P.poll_ns, 4000u));
                        continue;
                    }
                    while (t_rel < it.ntiles && td > t_rel && (int)(head - ctrl->tend[fw][(tcount + t_rel) & 3]) >= 0) {
                        release_stage(ctrl, jsite, siteB, (int)((tcount + t_rel) % S), NW, t_rel, it, lane);
                        ++t_rel;
                    }
                    if (t_rel == it.ntiles) break;
                }
            } else {
                // ---- several rings per drain warp: tile by tile, ring after ring, the rest of a ring's tile as a
                //      partial batch (fixed grouping, no carry-over) ------------------------------------------------
                for (int t = 0; t < it.ntiles; ++t) {
                    const unsigned g = tcount + t;
                    const int st = g % S;
                    mbar_wait(&ctrl->full[st], (g / S) & 1);
                    const int j0 = run_tile(it, t) * TJ;
                    if (st == 0) j0s0 = j0; else if (st == 1) j0s1 = j0; else j0s2 = j0;
#pragma unroll 1
                    for (int fw = dw; fw < NF; fw += ND) {      // this warp's rings, in this fixed order
                        const unsigned short* q = rings + fw * kRingCap;
                        unsigned head = ctrl->head[fw];         // this warp is its only writer
                        for (;;) {
                            const bool fin = ctrl->tdone[fw] > g;
                            __threadfence_block();
                            const unsigned limit = fin ? ctrl->tend[fw][g & 3] : ctrl->tail[fw];
                            const unsigned avail = limit - head;
                            if (avail < 32u * kU && !(fin && avail > 0u)) {
                                if (fin) break;
#pragma unroll 1
                                for (int z = P.poll_reps; z > 0; --z) __nanosleep(P.poll_ns * 10u);
                                continue;
                            }
                            __threadfence_block();
                            const int n = (int)min(avail, 32u * kU);
                            process_batch(q, fw, head, n);
                            head += (unsigned)n;
                        }
                    }
                    release_stage(ctrl, jsite, siteB, st, NW, t, it, lane);
                }
            }
            tcount += it.ntiles;
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

    // ---- flush: replicas summed in a fixed order -> this CTA's partial histogram ------------------------
    // partial layout: WEIGHTS ? [bin][8] = {g000, g110, g101, g011, g220, g202, g022, pair weight} : [bin] = pair weight
    __syncthreads();
    if (HIST == 0) {
        constexpr int PS = WEIGHTS ? 8 : 1;
        double* out = P.partials + (size_t)blockIdx.x * hn * PS;
        for (int k = tid; k < hn * PS; k += NW * 32) {
            const int bin = k / PS, sl = k - bin * PS;
            double s = 0.0;
            for (int r = 0; r < nrep; ++r) {
                const unsigned char* base = hist + (size_t)r * rbytes;
                if (!WEIGHTS || sl == 0 || sl == 7) {
                    const unsigned c = reinterpret_cast<const unsigned*>(base + (WEIGHTS ? (size_t)hn * 48 : 0))[bin];
                    if (!WEIGHTS || sl == 7 || (P.mode_sel & 1)) s += (double)c;
                } else {
                    s += reinterpret_cast<const double*>(base)[(size_t)bin * 6 + (sl - 1)];
                }
            }
            out[k] = s;
        }
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
    // blocks are added in ascending order (fixed summation order); loads are issued eight at a time
    double s = 0.0;
    const size_t stride = (size_t)hn * nslot;
    int b = 0;
    for (; b + 8 <= grid; b += 8) {
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = partials[(size_t)(b + k) * stride + idx];
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[k];
    }
    for (; b < grid; ++b) s += partials[(size_t)b * stride + idx];
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

__global__ void status_kernel(const unsigned long long* __restrict__ counters, double* __restrict__ status)
{
    const unsigned long long e = counters[1];
    status[0] = (e & 2ull) ? 1.0 : 0.0;
    status[1] = (e & 1ull) ? 1.0 : 0.0;
}

cudaError_t status_launch(const unsigned long long* counters, double* status, cudaStream_t s)
{
    status_kernel<<<1, 1, 0, s>>>(counters, status);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// variant selection
// ------------------------------------------------------------------------------------------------
static size_t smem_fixed_bytes(int nf)
{
    const size_t ti = (size_t)nf * 32 * kIPT;
    return ti * sizeof(Site) + (size_t)kStages * kTJ * sizeof(Site) + (size_t)nf * kRingCap * sizeof(unsigned short) + kCtrlBytes;
}

typedef void (*pair_fn_t)(const PairParams);

template <typename F, int NF, int ND, bool SHELLS>
static pair_fn_t pick_shape(int hist_mode, bool weights, bool allm)
{
    if (hist_mode == 0) {
        if (!weights) return pair_kernel<F, NF, ND, false, 0, false, SHELLS>;
        return allm ? pair_kernel<F, NF, ND, true, 0, true, SHELLS> : pair_kernel<F, NF, ND, true, 0, false, SHELLS>;
    }
    return hist_mode == 1 ? pair_kernel<F, NF, ND, true, 1, false, SHELLS> : pair_kernel<F, NF, ND, true, 2, false, SHELLS>;
}

template <typename F>
static pair_fn_t pick_fn(const PairConfig& c)
{
    const bool weights = c.nslot == 8, allm = c.all_modes != 0;
    if (c.shells) return c.drain_warps == 8 ? pick_shape<F, 8, 8, true>(c.hist_mode, weights, allm) : pick_shape<F, 8, 4, true>(c.hist_mode, weights, allm);
    return c.drain_warps == 8 ? pick_shape<F, 8, 8, false>(c.hist_mode, weights, allm) : pick_shape<F, 8, 4, false>(c.hist_mode, weights, allm);
}

#ifdef GFN_EXP_12_6
static pair_fn_t pick(const PairConfig& c) { return pair_kernel<__half, 12, 6, true, 0, true, false>; }
#elif defined(GFN_EXP_ONLY_F32)
static pair_fn_t pick(const PairConfig& c) { return pair_kernel<float, 8, 8, true, 0, true, false>; }
#elif defined(GFN_EXP_ONLY_H16)      // experiment builds (profiles/exp.sh): one instantiation, seconds to compile; cfg3 fp16 shape only
static pair_fn_t pick(const PairConfig& c) { return pair_kernel<__half, 8, 8, true, 0, true, false>; }
#else
static pair_fn_t pick(const PairConfig& c) { return c.filter_fp64 == 1 ? pick_fn<double>(c) : (c.filter_fp64 == 2 ? pick_fn<__half>(c) : pick_fn<float>(c)); }
#endif

int tile_i(const PairConfig& cfg) { return cfg.dense == 2 ? 128 : (cfg.dense ? 64 : (cfg.warps - cfg.drain_warps) * 32 * kIPT); }

cudaError_t pair_configure(int dev, int n1, int histo_n, int mode_sel, unsigned flags, int shells, double in_range_hint, const HandleGeom& geom, PairConfig* cfg)
{
    cfg->shells = shells ? 1 : 0;
    cfg->dense = 0; cfg->dense_msel = 0;
    static_assert(sizeof(Ctrl) <= kCtrlBytes, "control block does not fit");
    cudaError_t e;
    int max_smem = 0, sms = 0;
    if ((e = cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    cfg->filter_fp64 = (flags & GFN_FLAG_FILTER_FP64) ? 1 : ((flags & GFN_FLAG_FILTER_FP16) ? 2 : 0);
    cfg->nslot = (mode_sel == 1) ? 1 : 8;
    cfg->all_modes = (mode_sel == 127) ? 1 : 0;
    cfg->nrep = 1;
    // CTA shapes in order of preference: 8 filter + 8 drain warps, 8 + 4 (fewer histogram replicas to fit)
#ifdef GFN_EXP_12_6
    const int shapes[2][2] = {{12, 6}, {12, 6}};
#else
    const int shapes[2][2] = {{8, 8}, {8, 4}};
#endif
    int force_nf = 0, force_nd = 0;
    if (const char* env = getenv("GFN_FILTER_WARPS")) force_nf = atoi(env);
    if (const char* env = getenv("GFN_DRAIN_WARPS")) force_nd = atoi(env);
    int nf = 8, nd = 8;
    if (flags & GFN_FLAG_PER_PARTICLE) { cfg->hist_mode = 2; cfg->nslot = 8; }
    else if (flags & GFN_FLAG_HIST_GLOBAL) { cfg->hist_mode = 1; cfg->nslot = 8; }
    else {
        const long long rb = (long long)replica_bytes(histo_n, cfg->nslot == 8);
        cfg->hist_mode = 0;
        int best = -1; long long best_fit = 0;
        for (int k = 0; k < 2; ++k) {
            const int f = shapes[k][0], d = shapes[k][1];
            if ((force_nf && force_nf != f) || (force_nd && force_nd != d)) continue;
            const long long fit = ((long long)max_smem - (long long)smem_fixed_bytes(f)) / rb;
            if (fit >= d) { best = k; best_fit = d; break; }              // a private replica per drain warp
            if (best < 0 && fit >= 1) { best = k; best_fit = fit >= 4 ? 4 : (fit >= 2 ? 2 : 1); }   // shared under a lock
        }
        if (best < 0) { cfg->hist_mode = 1; cfg->nslot = 8; }
        else { nf = shapes[best][0]; nd = shapes[best][1]; cfg->nrep = (int)best_fit; }
        if (const char* env = getenv("GFN_NREP")) {
            const int r = atoi(env);
            if (cfg->hist_mode == 0 && r >= 1 && r <= cfg->nrep) cfg->nrep = r;
        }
    }
    if (cfg->hist_mode != 0 && force_nd == 4) { nf = 8; nd = 4; }
    cfg->drain_warps = nd;
    cfg->warps = nf + nd;
    const bool sh = cfg->hist_mode == 0;
    cfg->smem_bytes = (int)(smem_fixed_bytes(nf) + (sh ? (size_t)cfg->nrep * replica_bytes(histo_n, cfg->nslot == 8) : 0));
    pair_fn_t fn = pick(*cfg);
    if ((e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg->smem_bytes)) != cudaSuccess) return e;
    cfg->blocks_per_sm = 1;
    cfg->grid = sms;
    // dense neighbourhoods: the symmetric kernel, when it supports the handle
    // tile skipping with orientation modes runs on the dense kernel (its warps all filter AND evaluate, so the survivors a
    // cell-sorted frame concentrates in its near tiles do not stall anybody)
    // (g000-only handles: the count kernel, whose warps are independent anyway)
    const bool cells_dense = (flags & GFN_FLAG_CELLS) && !(flags & GFN_FLAG_NO_DENSE);
    const bool want_dense = (flags & GFN_FLAG_DENSE) || cells_dense || (!(flags & GFN_FLAG_NO_DENSE) && in_range_hint > kDenseShare);
    if (want_dense) {
        cudaError_t de = cudaSuccess;
        // (a histogram too large for pair_kernel's per-warp replicas sent it to global atomics above; the dense-box kernels
        // have their own, smaller shared layouts and refuse only what the CALLER forced to global / per-particle)
        const int ring_hist = cfg->hist_mode;
        if (!(flags & (GFN_FLAG_PER_PARTICLE | GFN_FLAG_HIST_GLOBAL))) cfg->hist_mode = 0;
        // g000 only: the count kernel (fp32 classification, fp64 for the ambiguous pairs), else the dense walk kernel
        if (count_configure(dev, histo_n, mode_sel, geom.invdx, geom.hmax, geom.box_max, max_smem, sms, cfg, &de)) return cudaSuccess;
        if (de != cudaSuccess) return de;
        if (dense_configure(dev, histo_n, mode_sel, max_smem, sms, geom, cfg, &de)) return cudaSuccess;
        if (de != cudaSuccess) return de;
        cfg->hist_mode = ring_hist;
    }
    return cudaSuccess;
}

cudaError_t pair_launch(const PairConfig& cfg, const PairParams& p, cudaStream_t s)
{
    if (cfg.dense == 2) return count_launch(cfg, p, s);
    if (cfg.dense) return dense_launch(cfg, p, s);
    pair_fn_t fn = pick(cfg);
    long long tiles = (long long)p.tiles_per_frame * p.nframes;
    int grid = cfg.grid;
    if (cfg.hist_mode != 0 && tiles < grid) grid = (int)(tiles > 0 ? tiles : 1);   // partials need the full grid only in smem mode
    fn<<<grid, cfg.warps * 32, cfg.smem_bytes, s>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// N1: per-residue centre of mass and dipole (the producers of the pair path's inputs)
//   helpers.pyx:71-98   comByResidue   com[i] = (sum_j coor[a]*m[a]) / (sum_j m[a]),  a = rfa[i] + j, j ascending
//   helpers.pyx:101-123 dipByResidue   dip[i] = sum_j (coor[a] - com[i]) * q[a]
// One thread per residue walks its atoms in the reference's order with un-fused multiply/add, so the
// results are bit-identical to the reference.  velcomByResidue (helpers.pyx:41-68) is the same arithmetic
// as comByResidue applied to velocities.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
residue_com_kernel(const T* __restrict__ coor, long long stride_coor, const double* __restrict__ masses,
                   const int* __restrict__ apr, const int* __restrict__ rfa, int nres,
                   double* __restrict__ com, long long stride_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nres) return;
    const T* c = coor + (long long)blockIdx.y * stride_coor;      // float positions widen exactly, like astype(float64)
    double sx = 0.0, sy = 0.0, sz = 0.0, tot = 0.0;
    const int first = rfa[i], n = apr[i];
    for (int j = 0; j < n; ++j) {
        const int a = first + j;
        const double m = masses[a];
        sx = __dadd_rn(sx, __dmul_rn((double)c[3ll * a], m));
        sy = __dadd_rn(sy, __dmul_rn((double)c[3ll * a + 1], m));
        sz = __dadd_rn(sz, __dmul_rn((double)c[3ll * a + 2], m));
        tot = __dadd_rn(tot, m);
    }
    double* o = com + (long long)blockIdx.y * stride_out + 3ll * i;
    o[0] = __ddiv_rn(sx, tot); o[1] = __ddiv_rn(sy, tot); o[2] = __ddiv_rn(sz, tot);
}

template <typename T>
__global__ void __launch_bounds__(128)
residue_dip_kernel(const T* __restrict__ coor, long long stride_coor, const double* __restrict__ charges,
                   const int* __restrict__ apr, const int* __restrict__ rfa, int nres,
                   const double* __restrict__ com, double* __restrict__ dip, long long stride_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nres) return;
    const T* c = coor + (long long)blockIdx.y * stride_coor;
    const double* cm = com + (long long)blockIdx.y * stride_out + 3ll * i;
    const double cx = cm[0], cy = cm[1], cz = cm[2];
    double dx = 0.0, dy = 0.0, dz = 0.0;
    const int first = rfa[i], n = apr[i];
    for (int j = 0; j < n; ++j) {
        const int a = first + j;
        const double q = charges[a];
        dx = __dadd_rn(dx, __dmul_rn(__dsub_rn((double)c[3ll * a], cx), q));
        dy = __dadd_rn(dy, __dmul_rn(__dsub_rn((double)c[3ll * a + 1], cy), q));
        dz = __dadd_rn(dz, __dmul_rn(__dsub_rn((double)c[3ll * a + 2], cz), q));
    }
    double* o = dip + (long long)blockIdx.y * stride_out + 3ll * i;
    o[0] = dx; o[1] = dy; o[2] = dz;
}

// stride_out: doubles between the outputs of consecutive frames (3*nres when the output is a dense (nframes,nres,3) array)
template <typename T>
static cudaError_t residue_com_launch_t(const T* coor, long long stride_coor, const double* masses, const int* apr, const int* rfa,
                                        int nres, int nframes, double* com, long long stride_out, cudaStream_t s)
{
    for (int f0 = 0; f0 < nframes; f0 += 65535) {
        const int nf = nframes - f0 < 65535 ? nframes - f0 : 65535;
        dim3 grid((nres + 127) / 128, nf);
        residue_com_kernel<T><<<grid, 128, 0, s>>>(coor + (long long)f0 * stride_coor, stride_coor, masses, apr, rfa, nres,
                                                   com + (long long)f0 * stride_out, stride_out);
    }
    return cudaGetLastError();
}

template <typename T>
static cudaError_t residue_dip_launch_t(const T* coor, long long stride_coor, const double* charges, const int* apr, const int* rfa,
                                        int nres, int nframes, const double* com, double* dip, long long stride_out, cudaStream_t s)
{
    for (int f0 = 0; f0 < nframes; f0 += 65535) {
        const int nf = nframes - f0 < 65535 ? nframes - f0 : 65535;
        dim3 grid((nres + 127) / 128, nf);
        residue_dip_kernel<T><<<grid, 128, 0, s>>>(coor + (long long)f0 * stride_coor, stride_coor, charges, apr, rfa, nres,
                                                   com + (long long)f0 * stride_out, dip + (long long)f0 * stride_out, stride_out);
    }
    return cudaGetLastError();
}

cudaError_t residue_com_launch(const double* coor, long long stride_coor, const double* masses, const int* apr, const int* rfa,
                               int nres, int nframes, double* com, cudaStream_t s)
{
    return residue_com_launch_t<double>(coor, stride_coor, masses, apr, rfa, nres, nframes, com, 3ll * nres, s);
}

cudaError_t residue_dip_launch(const double* coor, long long stride_coor, const double* charges, const int* apr, const int* rfa,
                               int nres, int nframes, const double* com, double* dip, cudaStream_t s)
{
    return residue_dip_launch_t<double>(coor, stride_coor, charges, apr, rfa, nres, nframes, com, dip, 3ll * nres, s);
}

// float32 atom positions (trajectory-native) -> centres of mass / dipoles written straight into a frame batch
cudaError_t residue_com_launch_f32(const float* coor, long long stride_coor, const double* masses, const int* apr, const int* rfa,
                                   int nres, int nframes, double* com, long long stride_out, cudaStream_t s)
{
    return residue_com_launch_t<float>(coor, stride_coor, masses, apr, rfa, nres, nframes, com, stride_out, s);
}

cudaError_t residue_dip_launch_f32(const float* coor, long long stride_coor, const double* charges, const int* apr, const int* rfa,
                                   int nres, int nframes, const double* com, double* dip, long long stride_out, cudaStream_t s)
{
    return residue_dip_launch_t<float>(coor, stride_coor, charges, apr, rfa, nres, nframes, com, dip, stride_out, s);
}

// ------------------------------------------------------------------------------------------------
// arithmetic self-test: in-range division / square-root sequences against the IEEE built-ins
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long splitmix(unsigned long long& x)
{
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// random double with a uniformly random mantissa and sign, biased exponent uniform in [723, 1323] (mid_range)
__device__ __forceinline__ double rand_mid(unsigned long long& st, bool positive)
{
    const unsigned long long r = splitmix(st);
    const unsigned long long man = r & 0x000fffffffffffffull;
    const unsigned long long ex = 723ull + (splitmix(st) % 601ull);
    const unsigned long long sg = positive ? 0ull : (r >> 63);
    return __longlong_as_double((long long)((sg << 63) | (ex << 52) | man));
}

__global__ void arith_selftest_kernel(unsigned long long seed, int per_thread, unsigned long long* mism)
{
    unsigned long long st = seed + 0x1234567ull * (blockIdx.x * blockDim.x + threadIdx.x);
    unsigned long long bad_div = 0, bad_sqrt = 0, bad_near = 0;
    for (int k = 0; k < per_thread; ++k) {
        const double a = rand_mid(st, false), b = rand_mid(st, false), x = rand_mid(st, true);
        if (__double_as_longlong(div_inrange(a, b)) != __double_as_longlong(__ddiv_rn(a, b))) ++bad_div;
        if (__double_as_longlong(sqrt_inrange(x)) != __double_as_longlong(__dsqrt_rn(x))) ++bad_sqrt;
        // operands close to each other / to perfect squares stress the final rounding
        const double b2 = __longlong_as_double(__double_as_longlong(a) + (long long)(splitmix(st) % 1024ull) - 512ll);
        if (mid_range(b2) && __double_as_longlong(div_inrange(a, b2)) != __double_as_longlong(__ddiv_rn(a, b2))) ++bad_near;
        const double sq = __dmul_rn(x, x);
        const double sq2 = __longlong_as_double(__double_as_longlong(sq) + (long long)(splitmix(st) % 8ull) - 4ll);
        if (mid_range(sq2) && __double_as_longlong(sqrt_inrange(sq2)) != __double_as_longlong(__dsqrt_rn(sq2))) ++bad_near;
    }
    if (bad_div) atomicAdd(mism + 0, bad_div);
    if (bad_sqrt) atomicAdd(mism + 1, bad_sqrt);
    if (bad_near) atomicAdd(mism + 2, bad_near);
}

cudaError_t arith_selftest(int dev, unsigned long long seed, long long n_operands, unsigned long long out[3])
{
    cudaError_t e;
    if ((e = cudaSetDevice(dev)) != cudaSuccess) return e;
    unsigned long long* d = nullptr;
    if ((e = cudaMalloc(&d, 3 * sizeof(unsigned long long))) != cudaSuccess) return e;
    cudaMemset(d, 0, 3 * sizeof(unsigned long long));
    const int blocks = 148 * 8, threads = 256;
    long long per = n_operands / ((long long)blocks * threads);
    if (per < 1) per = 1;
    if (per > 1000000) per = 1000000;
    arith_selftest_kernel<<<blocks, threads>>>(seed, (int)per, d);
    e = cudaMemcpy(out, d, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e != cudaSuccess ? e : cudaGetLastError();
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

// Micro-benchmark: issue/pipe cost of candidate range pre-filters (pairs per clock per SM).
//   float fold : 9 FADD + 3 FFMA + SHF         (shipping filter)
//   fixed point: 3 IADD + 3 IMAD.HI + ISETP + IADD3.X  (32-bit fractional coordinates, wrap is free)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o filter_loops filter_loops.cu && ./filter_loops
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int TJ = 128;

__global__ void __launch_bounds__(256) k_float(const float4* __restrict__ js, const float4* __restrict__ is, unsigned* out, int iters, float h, float Tg)
{
    __shared__ float4 sj[TJ];
    if (threadIdx.x < TJ) sj[threadIdx.x] = js[threadIdx.x];
    __syncthreads();
    const float4 c0 = is[threadIdx.x], c1 = is[threadIdx.x + 256];
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
        unsigned m0 = 0, m1 = 0;
#pragma unroll 16
        for (int j = 0; j < TJ; ++j) {
            const float4 s = sj[j];
            {
                float dx = s.x - c0.x, dy = s.y - c0.y, dz = s.z - c0.z;
                float wx = h - fabsf(fabsf(dx) - h), wy = h - fabsf(fabsf(dy) - h), wz = h - fabsf(fabsf(dz) - h);
                float t = fmaf(wz, wz, fmaf(wy, wy, fmaf(wx, wx, -Tg)));
                m0 = __funnelshift_l(__float_as_uint(t), m0, 1);
            }
            {
                float dx = s.x - c1.x, dy = s.y - c1.y, dz = s.z - c1.z;
                float wx = h - fabsf(fabsf(dx) - h), wy = h - fabsf(fabsf(dy) - h), wz = h - fabsf(fabsf(dz) - h);
                float t = fmaf(wz, wz, fmaf(wy, wy, fmaf(wx, wx, -Tg)));
                m1 = __funnelshift_l(__float_as_uint(t), m1, 1);
            }
        }
        acc += __popc(m0) + __popc(m1);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(256) k_fixed(const int4* __restrict__ js, const int4* __restrict__ is, unsigned* out, int iters, unsigned Tg)
{
    __shared__ int4 sj[TJ];
    if (threadIdx.x < TJ) sj[threadIdx.x] = js[threadIdx.x];
    __syncthreads();
    const int4 c0 = is[threadIdx.x], c1 = is[threadIdx.x + 256];
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
        unsigned m0 = 0, m1 = 0;
#pragma unroll 16
        for (int j = 0; j < TJ; ++j) {
            const int4 s = sj[j];
            {
                int dx = s.x - c0.x, dy = s.y - c0.y, dz = s.z - c0.z;
                unsigned t = (unsigned)__mulhi(dx, dx) + (unsigned)__mulhi(dy, dy) + (unsigned)__mulhi(dz, dz);
                m0 = m0 + m0 + (t < Tg ? 1u : 0u);
            }
            {
                int dx = s.x - c1.x, dy = s.y - c1.y, dz = s.z - c1.z;
                unsigned t = (unsigned)__mulhi(dx, dx) + (unsigned)__mulhi(dy, dy) + (unsigned)__mulhi(dz, dz);
                m1 = m1 + m1 + (t < Tg ? 1u : 0u);
            }
        }
        acc += __popc(m0) + __popc(m1);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// fixed point with 16-bit squares: dX>>16 squared fits 32 bits -> IMAD (low) accumulate
__global__ void __launch_bounds__(256) k_fixed16(const int4* __restrict__ js, const int4* __restrict__ is, unsigned* out, int iters, int Tg)
{
    __shared__ int4 sj[TJ];
    if (threadIdx.x < TJ) sj[threadIdx.x] = js[threadIdx.x];
    __syncthreads();
    const int4 c0 = is[threadIdx.x], c1 = is[threadIdx.x + 256];
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
        unsigned m0 = 0, m1 = 0;
#pragma unroll 16
        for (int j = 0; j < TJ; ++j) {
            const int4 s = sj[j];      // 15-bit signed fixed point: differences wrap in 16 bits via the shift pair
            {
                int dx = (short)(s.x - c0.x), dy = (short)(s.y - c0.y), dz = (short)(s.z - c0.z);
                int t = dx * dx + dy * dy + dz * dz - Tg;
                m0 = __funnelshift_l((unsigned)t, m0, 1);
            }
            {
                int dx = (short)(s.x - c1.x), dy = (short)(s.y - c1.y), dz = (short)(s.z - c1.z);
                int t = dx * dx + dy * dy + dz * dz - Tg;
                m1 = __funnelshift_l((unsigned)t, m1, 1);
            }
        }
        acc += __popc(m0) + __popc(m1);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}


// packed f32x2 (Blackwell FADD2/FFMA2): two surround sites per instruction, minimum image by magic-number rounding
__device__ __forceinline__ unsigned long long pk(float a, float b) { return ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(a); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) { unsigned long long d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

__global__ void __launch_bounds__(256) k_packed(const float4* __restrict__ js, const float4* __restrict__ is, unsigned* out, int iters, float L, float invL, float Tg)
{
    // surround sites stored pairwise: sx2[j/2] = (x_j, x_j+1) etc.
    __shared__ unsigned long long sx2[TJ / 2], sy2[TJ / 2], sz2[TJ / 2];
    if (threadIdx.x < TJ / 2) {
        const float4 a = js[2 * threadIdx.x], b = js[2 * threadIdx.x + 1];
        sx2[threadIdx.x] = pk(a.x, b.x); sy2[threadIdx.x] = pk(a.y, b.y); sz2[threadIdx.x] = pk(a.z, b.z);
    }
    __syncthreads();
    const float4 c0 = is[threadIdx.x], c1 = is[threadIdx.x + 256];
    const unsigned long long c0x = pk(c0.x, c0.x), c0y = pk(c0.y, c0.y), c0z = pk(c0.z, c0.z);
    const unsigned long long c1x = pk(c1.x, c1.x), c1y = pk(c1.y, c1.y), c1z = pk(c1.z, c1.z);
    const float MAGIC = 12582912.f;      // 1.5 * 2^23
    const unsigned long long iL = pk(invL, invL), mg = pk(MAGIC, MAGIC), nmg = pk(-MAGIC, -MAGIC), nL = pk(-L, -L), nT = pk(-Tg, -Tg);
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
        unsigned m0 = 0, m1 = 0;
#pragma unroll 16
        for (int j = 0; j < TJ / 2; ++j) {
            const unsigned long long X = sx2[j], Y = sy2[j], Z = sz2[j];
            {
                unsigned long long dx = sub2(X, c0x), dy = sub2(Y, c0y), dz = sub2(Z, c0z);
                unsigned long long qx = add2(fma2(dx, iL, mg), nmg), qy = add2(fma2(dy, iL, mg), nmg), qz = add2(fma2(dz, iL, mg), nmg);
                unsigned long long wx = fma2(qx, nL, dx), wy = fma2(qy, nL, dy), wz = fma2(qz, nL, dz);
                unsigned long long t = fma2(wz, wz, fma2(wy, wy, fma2(wx, wx, nT)));
                m0 = __funnelshift_l((unsigned)t, m0, 1); m0 = __funnelshift_l((unsigned)(t >> 32), m0, 1);
            }
            {
                unsigned long long dx = sub2(X, c1x), dy = sub2(Y, c1y), dz = sub2(Z, c1z);
                unsigned long long qx = add2(fma2(dx, iL, mg), nmg), qy = add2(fma2(dy, iL, mg), nmg), qz = add2(fma2(dz, iL, mg), nmg);
                unsigned long long wx = fma2(qx, nL, dx), wy = fma2(qy, nL, dy), wz = fma2(qz, nL, dz);
                unsigned long long t = fma2(wz, wz, fma2(wy, wy, fma2(wx, wx, nT)));
                m1 = __funnelshift_l((unsigned)t, m1, 1); m1 = __funnelshift_l((unsigned)(t >> 32), m1, 1);
            }
        }
        acc += __popc(m0) + __popc(m1);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// hybrid: the thread's two core sites share packed registers; FADD2 for s - c (surround value broadcast),
// scalar FADDs with |.| modifiers for the two folds, FFMA2 for the accumulation.  Same arithmetic as
// "float fold" (every operation is the same IEEE single-precision op), 10 issue slots per pair instead of 13.
__device__ __forceinline__ float lo32(unsigned long long v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi32(unsigned long long v) { return __uint_as_float((unsigned)(v >> 32)); }
__global__ void __launch_bounds__(256) k_hybrid(const float4* __restrict__ js, const float4* __restrict__ is, unsigned* out, int iters, float h, float Tg)
{
    __shared__ float4 sj[TJ];
    if (threadIdx.x < TJ) sj[threadIdx.x] = js[threadIdx.x];
    __syncthreads();
    const float4 c0 = is[threadIdx.x], c1 = is[threadIdx.x + 256];
    const unsigned long long cx = pk(c0.x, c1.x), cy = pk(c0.y, c1.y), cz = pk(c0.z, c1.z), nT = pk(-Tg, -Tg);
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
        unsigned m0 = 0, m1 = 0;
#pragma unroll 16
        for (int j = 0; j < TJ; ++j) {
            const float4 s = sj[j];
            const unsigned long long dx = sub2(pk(s.x, s.x), cx), dy = sub2(pk(s.y, s.y), cy), dz = sub2(pk(s.z, s.z), cz);
            const unsigned long long wx = pk(h - fabsf(fabsf(lo32(dx)) - h), h - fabsf(fabsf(hi32(dx)) - h));
            const unsigned long long wy = pk(h - fabsf(fabsf(lo32(dy)) - h), h - fabsf(fabsf(hi32(dy)) - h));
            const unsigned long long wz = pk(h - fabsf(fabsf(lo32(dz)) - h), h - fabsf(fabsf(hi32(dz)) - h));
            const unsigned long long t = fma2(wz, wz, fma2(wy, wy, fma2(wx, wx, nT)));
            m0 = __funnelshift_l((unsigned)t, m0, 1);
            m1 = __funnelshift_l((unsigned)(t >> 32), m1, 1);
        }
        acc += __popc(m0) + __popc(m1);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// fp16x2: the thread's two core sites in one half2; surround values pre-duplicated (x,x),(y,y),(z,z) in shared
// memory.  Coordinates in box units.  HADD2 (|.| modifiers) + HFMA2: ~6.5 issue slots per pair.
#include <cuda_fp16.h>
__global__ void __launch_bounds__(256) k_half(const float4* __restrict__ js, const float4* __restrict__ is, unsigned* out, int iters, float invL, float hh, float Tg)
{
    __shared__ uint4 sj[TJ];
    if (threadIdx.x < TJ) {
        const float4 a = js[threadIdx.x];
        __half2 xx = __floats2half2_rn(a.x * invL, a.x * invL), yy = __floats2half2_rn(a.y * invL, a.y * invL), zz = __floats2half2_rn(a.z * invL, a.z * invL);
        sj[threadIdx.x] = make_uint4(*reinterpret_cast<unsigned*>(&xx), *reinterpret_cast<unsigned*>(&yy), *reinterpret_cast<unsigned*>(&zz), 0u);
    }
    __syncthreads();
    const float4 c0 = is[threadIdx.x], c1 = is[threadIdx.x + 256];
    const __half2 cx = __floats2half2_rn(c0.x * invL, c1.x * invL), cy = __floats2half2_rn(c0.y * invL, c1.y * invL), cz = __floats2half2_rn(c0.z * invL, c1.z * invL);
    const __half2 h = __floats2half2_rn(hh, hh), nT = __floats2half2_rn(-Tg, -Tg);
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
        unsigned m0 = 0, m1 = 0;
#pragma unroll 16
        for (int j = 0; j < TJ; ++j) {
            const uint4 s = sj[j];
            const __half2 sx = *reinterpret_cast<const __half2*>(&s.x), sy = *reinterpret_cast<const __half2*>(&s.y), sz = *reinterpret_cast<const __half2*>(&s.z);
            const __half2 dx = __hsub2(sx, cx), dy = __hsub2(sy, cy), dz = __hsub2(sz, cz);
            const __half2 wx = __hsub2(h, __habs2(__hsub2(__habs2(dx), h)));
            const __half2 wy = __hsub2(h, __habs2(__hsub2(__habs2(dy), h)));
            const __half2 wz = __hsub2(h, __habs2(__hsub2(__habs2(dz), h)));
            const __half2 t = __hfma2(wz, wz, __hfma2(wy, wy, __hfma2(wx, wx, nT)));
            const unsigned tb = *reinterpret_cast<const unsigned*>(&t);
            m0 = __funnelshift_l(tb << 16, m0, 1);      // sign of the low half (core site 0)
            m1 = __funnelshift_l(tb, m1, 1);            // sign of the high half (core site 1)
        }
        acc += __popc(m0) + __popc(m1);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main(int argc, char** argv)
{
    const int sms = 148, bps = argc > 1 ? atoi(argv[1]) : 2, blocks = sms * bps, threads = 256, iters = 2000; printf("blocks per SM: %d (%d warps/SM)\n", bps, bps * 8);
    float4* dj; float4* di; unsigned* out;
    cudaMalloc(&dj, TJ * 16); cudaMalloc(&di, 512 * 16); cudaMalloc(&out, blocks * threads * 4);
    float4 hj[TJ], hi[512];
    for (int k = 0; k < TJ; ++k) hj[k] = make_float4(k * 0.7f, k * 0.3f, k * 0.11f, 0);
    for (int k = 0; k < 512; ++k) hi[k] = make_float4(k * 0.19f, k * 0.23f, k * 0.05f, 0);
    cudaMemcpy(dj, hj, sizeof(hj), cudaMemcpyHostToDevice); cudaMemcpy(di, hi, sizeof(hi), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double pairs = (double)blocks * threads * 2.0 * TJ * iters;
    for (int which = 0; which < 6; ++which) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (which == 0) k_float<<<blocks, threads>>>(dj, di, out, iters, 49.6f, 225.01f);
            if (which == 1) k_fixed<<<blocks, threads>>>((int4*)dj, (int4*)di, out, iters, 98000000u);
            if (which == 2) k_fixed16<<<blocks, threads>>>((int4*)dj, (int4*)di, out, iters, 98000000);
            if (which == 4) k_hybrid<<<blocks, threads>>>(dj, di, out, iters, 49.6f, 225.01f);
            if (which == 5) k_half<<<blocks, threads>>>(dj, di, out, iters, 1.0f / 99.33f, 0.5f, 0.0228f);
            if (which == 3) k_packed<<<blocks, threads>>>(dj, di, out, iters, 99.33f, 1.0f / 99.33f, 225.01f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%s: %.3f ms  %.1f Gpair/s  (%.2f pairs/clk/SM at 1.965 GHz)\n", which == 0 ? "float fold " : which == 1 ? "fixed mulhi" : which == 2 ? "fixed 16bit" : which == 3 ? "packed x2  " : which == 4 ? "hybrid x2  " : "half2 fold ",
               best, pairs / best / 1e6, pairs / (best * 1e-3) / 148 / 1.965e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// Micro-benchmark: issue/pipe cost of candidate range pre-filters (pairs per clock per SM).
//   float fold : 9 FADD + 3 FFMA + SHF         (shipping filter)
//   fixed point: 3 IADD + 3 IMAD.HI + ISETP + IADD3.X  (32-bit fractional coordinates, wrap is free)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o filter_loops filter_loops.cu && ./filter_loops
#include <cstdio>
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

int main()
{
    const int sms = 148, blocks = sms * 2, threads = 256, iters = 2000;
    float4* dj; float4* di; unsigned* out;
    cudaMalloc(&dj, TJ * 16); cudaMalloc(&di, 512 * 16); cudaMalloc(&out, blocks * threads * 4);
    float4 hj[TJ], hi[512];
    for (int k = 0; k < TJ; ++k) hj[k] = make_float4(k * 0.7f, k * 0.3f, k * 0.11f, 0);
    for (int k = 0; k < 512; ++k) hi[k] = make_float4(k * 0.19f, k * 0.23f, k * 0.05f, 0);
    cudaMemcpy(dj, hj, sizeof(hj), cudaMemcpyHostToDevice); cudaMemcpy(di, hi, sizeof(hi), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double pairs = (double)blocks * threads * 2.0 * TJ * iters;
    for (int which = 0; which < 3; ++which) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (which == 0) k_float<<<blocks, threads>>>(dj, di, out, iters, 49.6f, 225.01f);
            if (which == 1) k_fixed<<<blocks, threads>>>((int4*)dj, (int4*)di, out, iters, 98000000u);
            if (which == 2) k_fixed16<<<blocks, threads>>>((int4*)dj, (int4*)di, out, iters, 98000000);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%s: %.3f ms  %.1f Gpair/s  (%.2f pairs/clk/SM at 1.965 GHz)\n", which == 0 ? "float fold " : which == 1 ? "fixed mulhi" : "fixed 16bit",
               best, pairs / best / 1e6, pairs / (best * 1e-3) / 148 / 1.965e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

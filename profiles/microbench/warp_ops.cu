// warp_ops.cu - issue cost of the warp-wide primitives the histogram update can be built from (B200, sm_100a):
// match.any with many / few distinct keys, redux, vote/ballot, the 10-ballot software match, shared-memory atomics.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_ops warp_ops.cu && ./warp_ops
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(512) k(unsigned* out, int iters, int distinct)
{
    __shared__ unsigned sh[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    unsigned key = (lane % distinct) * 37u + 5u, acc = 0;
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) { acc += __match_any_sync(0xffffffffu, key); }
        if (OP == 1) { acc += __reduce_max_sync(0xffffffffu, key); }
        if (OP == 2) { acc += __ballot_sync(0xffffffffu, key & 1u); }
        if (OP == 3) {           // software match over 10 key bits
            unsigned eq = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 10; ++b) { const unsigned v = __ballot_sync(0xffffffffu, (key >> b) & 1u); eq &= ((key >> b) & 1u) ? v : ~v; }
            acc += eq;
        }
        if (OP == 4) { acc += atomicAdd(&sh[(key * 9u + (threadIdx.x >> 5) * 256u) & 4095u], 1u); }       // with return
        if (OP == 5) { atomicAdd(&sh[(key * 9u + (threadIdx.x >> 5) * 256u) & 4095u], 1u); }               // no return
        if (OP == 6) { unsigned* p = &sh[(key * 9u + (threadIdx.x >> 5) * 256u) & 4095u]; *p = *p + 1u; __syncwarp(); }   // plain RMW
        key = key * 1664525u + 1013904223u + acc;          // keys change every iteration (data-dependent)
        if (distinct < 32) key = (key % distinct) * 37u + 5u;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + sh[threadIdx.x];
}

template <int OP>
static void run(const char* name, int distinct)
{
    unsigned* out; cudaMalloc(&out, 148 * 512 * 4);
    const int iters = 4096;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<OP><<<148, 512>>>(out, 64, distinct);
    cudaEventRecord(a); k<OP><<<148, 512>>>(out, iters, distinct); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    // 16 warps per SM, iters ops each
    const double clk = ms * 1e-3 * 1.965e9;
    printf("%-44s distinct=%2d  %.1f clk per warp-op per SM (16 warps in flight)  [%s]\n", name, distinct, clk / (16.0 * iters), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main()
{
    for (int d : {32, 8, 1}) run<0>("match.any", d);
    run<1>("redux.max", 32);
    run<2>("ballot", 32);
    run<3>("software match (10 ballots + 10 lop3)", 32);
    for (int d : {32, 4, 1}) run<4>("shared atomicAdd u32 with return", d);
    for (int d : {32, 4, 1}) run<5>("shared atomicAdd u32 no return", d);
    for (int d : {32, 4}) run<6>("plain shared RMW (ld, add, st, syncwarp)", d);
    return 0;
}

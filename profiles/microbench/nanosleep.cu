// How long does __nanosleep(ns) really sleep on sm_100a?  One warp per SM, clock64 around 200 sleeps.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned ns, long long* out)
{
    long long t0 = clock64();
    for (int i = 0; i < 200; ++i) __nanosleep(ns);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = (t1 - t0) / 200;
}
int main()
{
    long long* d; cudaMalloc(&d, 8);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const unsigned v[] = {20, 50, 100, 200, 400, 800, 1000, 1500, 2000, 3000, 4000, 8000, 16000, 100000};
    for (unsigned ns : v) {
        k<<<148, 32>>>(ns, d);
        long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("nanosleep(%6u): %8lld cycles = %8.0f ns per call\n", ns, c, c * 1e6 / khz);
    }
    return 0;
}

// start-up reference: how long does a process need for cuInit + primary context on this box?
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
__global__ void k(int *p) { if (p) *p = 1; }
int main()
{
    auto t0 = std::chrono::steady_clock::now();
    cudaFree(0);
    auto t1 = std::chrono::steady_clock::now();
    int *d; cudaMalloc(&d, 4); k<<<1, 1>>>(d); cudaDeviceSynchronize();
    auto t2 = std::chrono::steady_clock::now();
    printf("context %.1f ms, first launch %.1f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count());
    return 0;
}

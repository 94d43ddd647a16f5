// where do the warps of small CTAs land?  prints histogram of (%warpid % 4) for 96-thread CTAs
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* out, int spin) {
  unsigned smid, wid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  long long t0 = clock64(); while (clock64() - t0 < spin) {}
  if ((threadIdx.x & 31) == 0) { out[(blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32) * 2] = smid; out[(blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32) * 2 + 1] = wid; }
}
int main() {
  int* d; cudaMalloc(&d, 1 << 20);
  for (int nthr : {96, 192, 128}) {
    int nblk = 400, nw = nthr / 32;
    k<<<nblk, nthr>>>(d, 200000); cudaDeviceSynchronize();
    static int h[1 << 18]; cudaMemcpy(h, d, nblk * nw * 8, cudaMemcpyDeviceToHost);
    int hist[4] = {0, 0, 0, 0}; int persm[148][4] = {};
    for (int i = 0; i < nblk * nw; i++) { hist[h[2 * i + 1] % 4]++; persm[h[2 * i]][h[2 * i + 1] % 4]++; }
    printf("threads=%d: warpid%%4 histogram: %d %d %d %d ; SM0: %d %d %d %d ; SM77: %d %d %d %d\n", nthr, hist[0], hist[1], hist[2], hist[3],
           persm[0][0], persm[0][1], persm[0][2], persm[0][3], persm[77][0], persm[77][1], persm[77][2], persm[77][3]);
    printf("  first CTAs: "); for (int i = 0; i < 4 * nw; i++) printf("(sm%d,w%d) ", h[2 * i], h[2 * i + 1]); printf("\n");
  }
  return 0;
}

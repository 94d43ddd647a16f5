// pipe probe 3: does HMNMX2 (fp16x2 max on bit patterns) issue on a different pipe than the integer ALU?
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
constexpr int NCH = 8; constexpr int ITER = 2048;
__device__ __forceinline__ unsigned hmax2u(unsigned a, unsigned b) {
  __half2 x = *reinterpret_cast<__half2*>(&a), y = *reinterpret_cast<__half2*>(&b);
  __half2 r = __hmax2(x, y); return *reinterpret_cast<unsigned*>(&r);
}
template<int OP> __global__ void k(unsigned* out, unsigned seed, long long* cyc) {
  unsigned r[NCH]; unsigned b = (seed * 3 + threadIdx.x) & 0x3fff3fff, c = 0x10011001;
#pragma unroll
  for (int j = 0; j < NCH; j++) r[j] = ((threadIdx.x * 17 + j * 1315423911u + seed) & 0x3fff3fffu) | 0x04000400u;
  long long t0 = clock64();
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int j = 0; j < NCH; j++) {
      unsigned a = r[j], n = r[(j + 1) % NCH];
      if (OP == 0) r[j] = hmax2u(a, n);                                            // HMNMX2
      if (OP == 1) r[j] = __vmaxs2(a, n);                                          // VIMNMX.S16x2
      if (OP == 2) { r[j] = hmax2u(a, n); r[j] = __vmaxs2(r[j], b); }              // HMNMX2 + VIMNMX
      if (OP == 3) { r[j] = hmax2u(a, n); r[j] = r[j] + c; }                       // HMNMX2 + IADD
      if (OP == 4) { r[j] = hmax2u(a, n); r[j] = r[j] * 3u + c; }                  // HMNMX2 + IMAD
      if (OP == 5) { r[j] = hmax2u(a, n); r[j] = __vmaxs2(r[j], b); r[j] = r[j] * 3u + c; }   // all three
      if (OP == 6) { r[j] = hmax2u(a, n); r[j] = (r[j] | 0x30003u) ^ b; }          // HMNMX2 + LOP3
    }
  }
  long long t1 = clock64();
  unsigned acc = 0;
#pragma unroll
  for (int j = 0; j < NCH; j++) acc ^= r[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template<int OP> void run(const char* name, unsigned* dout, long long* dcyc) {
  int nblk = 148, nthr = 512;
  k<OP><<<nblk, nthr>>>(dout, 1234u, dcyc); cudaDeviceSynchronize();
  k<OP><<<nblk, nthr>>>(dout, 1234u, dcyc); cudaDeviceSynchronize();
  static long long hc[256]; cudaMemcpy(hc, dcyc, 8 * nblk, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < nblk; i++) avg += hc[i]; avg /= nblk;
  printf("%-28s cycles/warp-iter/SMSP=%.2f\n", name, avg / ((double)ITER * NCH) / 4.0);
}
int main() {
  unsigned* dout; long long* dcyc; cudaMalloc(&dout, 148 * 512 * 4); cudaMalloc(&dcyc, 4096);
  run<0>("HMNMX2", dout, dcyc); run<1>("VIMNMX.S16x2", dout, dcyc); run<2>("HMNMX2+VIMNMX", dout, dcyc);
  run<3>("HMNMX2+IADD", dout, dcyc); run<4>("HMNMX2+IMAD", dout, dcyc); run<5>("HMNMX2+VIMNMX+IMAD", dout, dcyc); run<6>("HMNMX2+LOP3", dout, dcyc);
  return 0;
}

// second-pass pipe probe: non-foldable chains + pairwise mixes (exploration tool)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int NCH = 8; constexpr int ITER = 2048;
#define N(j) r[((j)+1)%NCH]
template<int OP> __device__ __forceinline__ void body(unsigned (&r)[NCH], unsigned b, unsigned c) {
#pragma unroll
  for (int j = 0; j < NCH; j++) {
    unsigned a = r[j], n = N(j);
    if (OP == 0) r[j] = a + n;                                   // IADD3
    if (OP == 1) r[j] = (a | 0x30003u) ^ n;                      // LOP3
    if (OP == 2) r[j] = (unsigned)max((int)a, (int)n) ;          // VIMNMX s32
    if (OP == 3) r[j] = __vmaxs2(a, n);                          // VIMNMX.S16x2
    if (OP == 4) r[j] = __viaddmax_s16x2(a, b, n);               // VIADDMNMX.S16x2
    if (OP == 5) r[j] = a * b + n;                               // IMAD
    if (OP == 6) r[j] = __vadd2(a, n);                           // VIADD.16x2
    if (OP == 7) r[j] = __funnelshift_r(a, n, 2);                // SHF
    if (OP == 8) r[j] = (a << 2) + n;                            // LEA
    if (OP == 9) r[j] = __vimax3_s16x2(a, b, n);                 // VIMNMX3
    if (OP == 10) { r[j] = __viaddmax_s16x2(a, b, n); r[j] = r[j] * b + n; }       // DPX + IMAD (2 ops)
    if (OP == 11) { r[j] = __viaddmax_s16x2(a, b, n); r[j] = r[j] + n; }           // DPX + IADD (may become IMAD.IADD)
    if (OP == 12) { r[j] = __vmaxs2(a, n); r[j] = r[j] * b + n; }                  // VIMNMX + IMAD
    if (OP == 13) { r[j] = __vmaxs2(a, n); r[j] = (r[j] | 0x30003u) ^ b; }         // VIMNMX + LOP3
    if (OP == 14) { r[j] = __viaddmax_s16x2(a, b, n); r[j] = __vmaxs2(r[j], c); }  // DPX + VIMNMX
    if (OP == 15) { r[j] = __viaddmax_s16x2(a, b, n); r[j] = (r[j] | 0x30003u) ^ c; } // DPX + LOP3
    if (OP == 16) { r[j] = __vadd2(a, b); r[j] = __vmaxs2(r[j], n); }              // VIADD.16x2 + VIMNMX (unfused add-max)
    if (OP == 17) { unsigned t = a + b; r[j] = __vmaxs2(t, n); }                   // IADD + VIMNMX
    if (OP == 18) { r[j] = __viaddmax_s32((int)a, (int)b, (int)n); }               // VIADDMNMX s32
    if (OP == 19) { r[j] = __vibmax_s16x2(a, n, nullptr, nullptr); }
  }
}
template<int OP> __global__ void k(unsigned* out, unsigned seed, long long* cyc) {
  unsigned r[NCH]; unsigned b = seed * 3 + threadIdx.x, c = seed + 7;
#pragma unroll
  for (int j = 0; j < NCH; j++) r[j] = threadIdx.x * 17 + j * 1315423911u + seed;
  long long t0 = clock64();
  for (int it = 0; it < ITER; it++) body<OP>(r, b, c);
  long long t1 = clock64();
  unsigned acc = 0;
#pragma unroll
  for (int j = 0; j < NCH; j++) acc ^= r[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template<int OP> void run(const char* name, int nops, unsigned* dout, long long* dcyc) {
  int nblk = 148, nthr = 512;
  k<OP><<<nblk, nthr>>>(dout, 1234u, dcyc); cudaDeviceSynchronize();
  k<OP><<<nblk, nthr>>>(dout, 1234u, dcyc); cudaDeviceSynchronize();
  static long long hc[256]; cudaMemcpy(hc, dcyc, 8 * nblk, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < nblk; i++) avg += hc[i]; avg /= nblk;
  double per = avg / ((double)ITER * NCH) / 4.0;   // cycles per warp-"body" per SMSP (4 warps/SMSP)
  printf("%-28s cycles/warp-iter/SMSP=%.2f  (%d src ops)\n", name, per, nops);
}
int main() {
  unsigned* dout; long long* dcyc; cudaMalloc(&dout, 148 * 512 * 4); cudaMalloc(&dcyc, 4096);
  run<0>("IADD3", 1, dout, dcyc); run<1>("LOP3", 1, dout, dcyc); run<2>("VIMNMX.s32", 1, dout, dcyc);
  run<3>("VIMNMX.S16x2", 1, dout, dcyc); run<4>("VIADDMNMX.S16x2", 1, dout, dcyc); run<5>("IMAD", 1, dout, dcyc);
  run<6>("VIADD.16x2", 1, dout, dcyc); run<7>("SHF", 1, dout, dcyc); run<8>("LEA", 1, dout, dcyc); run<9>("VIMNMX3.S16x2", 1, dout, dcyc);
  run<18>("VIADDMNMX.s32", 1, dout, dcyc); run<19>("vibmax_s16x2(no pred)", 1, dout, dcyc);
  run<10>("DPX+IMAD", 2, dout, dcyc); run<11>("DPX+IADD", 2, dout, dcyc); run<12>("VIMNMX2+IMAD", 2, dout, dcyc);
  run<13>("VIMNMX2+LOP3", 2, dout, dcyc); run<14>("DPX+VIMNMX2", 2, dout, dcyc); run<15>("DPX+LOP3", 2, dout, dcyc);
  run<16>("VIADD16x2+VIMNMX2", 2, dout, dcyc); run<17>("IADD+VIMNMX2", 2, dout, dcyc);
  return 0;
}

// INT/DPX issue-rate microbenchmark for sm_100a (exploration tool; not on the product path).
// Each kernel runs ITER x UNROLL independent-chain ops per thread; reports lane-ops/clk/SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

constexpr int NCH = 8;      // independent chains per thread
constexpr int ITER = 4096;

template<int OP> __device__ __forceinline__ unsigned op(unsigned a, unsigned b, unsigned c) {
  if (OP == 0) return a + b;                                   // IADD3
  if (OP == 1) return (unsigned)__viaddmax_s32((int)a, (int)b, (int)c);   // VIADDMNMX
  if (OP == 2) return __viaddmax_s16x2(a, b, c);               // VIADDMNMX.S16x2
  if (OP == 3) return __vmaxs2(a, b);                          // VIMNMX.S16x2
  if (OP == 4) return __vimax3_s16x2(a, b, c);                 // VIMNMX3.S16x2
  if (OP == 5) return (a | 0x00030003u) ^ b;                   // LOP3
  if (OP == 6) return a * 5u + b;                              // IMAD
  if (OP == 7) return __vadd2(a, b);                           // VIADD.16x2
  if (OP == 8) return __funnelshift_r(a, b, 2);                // SHF
  if (OP == 9) return (unsigned)max((int)a, (int)b);           // IMNMX/VIMNMX s32
  if (OP == 10) return (unsigned)__vimax3_s32((int)a,(int)b,(int)c);
  if (OP == 11) return (a << 2) + b;                           // LEA
  return a;
}

template<int OP> __global__ void k_alu(unsigned* out, unsigned seed, long long* cyc) {
  unsigned r[NCH];
  unsigned b = seed * 3 + threadIdx.x, c = seed + 7;
#pragma unroll
  for (int j = 0; j < NCH; j++) r[j] = threadIdx.x * 17 + j + seed;
  long long t0 = clock64();
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int j = 0; j < NCH; j++) r[j] = op<OP>(r[j], b, c);
  }
  long long t1 = clock64();
  unsigned acc = 0;
#pragma unroll
  for (int j = 0; j < NCH; j++) acc ^= r[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// mixed: per "cell": 4 ALU-only DPX ops + 3 flexible ops (add, sub, mul-add), mimicking the sweep inner loop
__global__ void k_mix(unsigned* out, unsigned seed, long long* cyc) {
  unsigned r[NCH], w[NCH];
  unsigned b = seed * 3 + threadIdx.x, c = seed + 7;
#pragma unroll
  for (int j = 0; j < NCH; j++) { r[j] = threadIdx.x * 17 + j + seed; w[j] = j; }
  long long t0 = clock64();
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int j = 0; j < NCH; j++) {
      unsigned m1 = __viaddmax_s16x2(r[j], b, 0u);
      unsigned m1s = m1 + c;
      unsigned m2 = __viaddmax_s16x2(r[(j + 1) % NCH], c, m1s);
      unsigned h = __vmaxs2(w[(j + 3) % NCH], m2);
      unsigned z = h | 0x00030003u;
      unsigned d = z - h;
      w[j] = w[j] * 4u + d;
      r[j] = z;
    }
  }
  long long t1 = clock64();
  unsigned acc = 0;
#pragma unroll
  for (int j = 0; j < NCH; j++) acc ^= r[j] ^ w[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_lds128(unsigned* out, unsigned seed, long long* cyc) {
  extern __shared__ uint4 sm[];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_uint4(i, seed, i * 3, 1);
  __syncthreads();
  uint4 acc = make_uint4(0, 0, 0, 0);
  int idx = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int j = 0; j < NCH; j++) {
      uint4 v = sm[(idx + j * 256) & 2047];
      acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w;
    }
    idx = (idx + acc.w) & 2047 & ~0 ;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x ^ acc.y ^ acc.z ^ acc.w;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_shfl(unsigned* out, unsigned seed, long long* cyc) {
  unsigned r[NCH];
#pragma unroll
  for (int j = 0; j < NCH; j++) r[j] = threadIdx.x * 17 + j + seed;
  long long t0 = clock64();
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int j = 0; j < NCH; j++) r[j] = __shfl_up_sync(0xffffffffu, r[j], 1) + 1;
  }
  long long t1 = clock64();
  unsigned acc = 0;
#pragma unroll
  for (int j = 0; j < NCH; j++) acc ^= r[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// latency probes: dependent chain, 1 warp
__global__ void k_lat(unsigned* out, unsigned seed, long long* cyc) {
  __shared__ int smk[4];
  unsigned r = threadIdx.x + seed, b = seed | 1;
  long long t0 = clock64();
  for (int it = 0; it < ITER; it++) r = __viaddmax_s16x2(r, b, r ^ 1);   // dependent DPX
  long long t1 = clock64();
  for (int it = 0; it < ITER; it++) r = __shfl_up_sync(0xffffffffu, r, 1) ^ b;     // dependent SHFL+LOP
  long long t2 = clock64();
  for (int it = 0; it < ITER; it++) r = (unsigned)__reduce_max_sync(0xffffffffu, (int)(r + threadIdx.x));   // REDUX
  long long t3 = clock64();
  if (threadIdx.x < 4) smk[threadIdx.x] = 0;
  __syncthreads();
  long long t4 = clock64();
  for (int it = 0; it < ITER; it++) { atomicMax(&smk[it & 3], (int)(r & 0xffff) + it); __syncthreads(); r += smk[(it + 1) & 3]; }
  long long t5 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t5 - t4; }
}

template<class F> int run(const char* name, F launch, int nblk, int nthr, double ops_per_thread, unsigned* dout, long long* dcyc) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); CK(cudaDeviceSynchronize());
  cudaEventRecord(e0); launch(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  static long long hc[4096]; CK(cudaMemcpy(hc, dcyc, sizeof(long long) * nblk, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < nblk; i++) avg += hc[i]; avg /= nblk;
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double blk_per_sm = (double)nblk / p.multiProcessorCount;
  double laneops_per_clk_sm = ops_per_thread * nthr * blk_per_sm / avg;
  double tops = ops_per_thread * nthr * (double)nblk / (ms * 1e-3) / 1e12;
  printf("%-22s blocks=%d thr=%d  cyc/blk=%.0f  lane-ops/clk/SM=%.1f  wall=%.3f ms  Tlane-ops/s=%.2f  eff_clk=%.0f MHz\n", name, nblk, nthr, avg,
         laneops_per_clk_sm, ms, tops, avg / (ms * 1e-3) / 1e6);
  return 0;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  unsigned* dout; long long* dcyc;
  CK(cudaMalloc(&dout, 4096 * 1024 * 4)); CK(cudaMalloc(&dcyc, 4096 * 8));
  const char* names[] = {"IADD", "VIADDMNMX.s32", "VIADDMNMX.s16x2", "VIMNMX.s16x2", "VIMNMX3.s16x2", "LOP3x2", "IMAD", "VIADD.16x2", "SHF.funnel", "IMNMX.s32", "VIMNMX3.s32", "LEA"};
  int nsm = p.multiProcessorCount;
  for (int cfg = 0; cfg < 2; cfg++) {
    int nblk = nsm * (cfg == 0 ? 1 : 2), nthr = cfg == 0 ? 512 : 1024;
    double opt = (double)ITER * NCH;
#define RUN(OP) run(names[OP], [&]{ k_alu<OP><<<nblk, nthr>>>(dout, 1234u, dcyc); }, nblk, nthr, opt, dout, dcyc);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11)
    run("MIX(8 ops/cell)", [&]{ k_mix<<<nblk, nthr>>>(dout, 1234u, dcyc); }, nblk, nthr, opt * 8, dout, dcyc);
    run("LDS.128", [&]{ k_lds128<<<nblk, nthr, 32768>>>(dout, 0u, dcyc); }, nblk, nthr, opt, dout, dcyc);
    run("SHFL.up", [&]{ k_shfl<<<nblk, nthr>>>(dout, 1u, dcyc); }, nblk, nthr, opt, dout, dcyc);
  }
  k_lat<<<1, 32>>>(dout, 5u, dcyc); CK(cudaDeviceSynchronize());
  long long hc[4]; CK(cudaMemcpy(hc, dcyc, 32, cudaMemcpyDeviceToHost));
  printf("latency (cycles/op, 1 warp): DPX dep=%.1f  SHFL+LOP dep=%.1f  REDUX dep=%.1f\n", hc[0] / (double)ITER, hc[1] / (double)ITER, hc[2] / (double)ITER);
  k_lat<<<1, 96>>>(dout, 5u, dcyc); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(hc, dcyc, 32, cudaMemcpyDeviceToHost));
  printf("3-warp CTA: ATOMS.max+BAR+LDS round = %.1f cycles\n", hc[3] / (double)ITER);
  return 0;
}

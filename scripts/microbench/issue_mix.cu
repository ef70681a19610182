// Issue-slot microbenchmark for the fp64 roofline discussion (profiles/): on B200 an fp64 warp instruction occupies the
// fp64 pipe of an SM sub-partition for 2 cycles. Questions this answers, at 16 warps / SM (4 per sub-partition, like
// k_eval_staged) and 256-thread CTAs:
//   (a) do integer / shared-memory-load instructions issue in the shadow of the fp64 pipe (second cycle), or do they
//       cost extra cycles? -> DFMA chains with N independent integer ops (or LDS.128) per DFMA
//   (b) dependent-issue latency of DFMA / DADD / DMUL (1 chain, 1 warp per sub-partition)
//   (c) MUFU.RCP64H cost next to DFMA
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o issue_mix issue_mix.cu ; run: ./issue_mix
#include <cstdio>
#include <cuda_runtime.h>

#define FMA(s, x, y) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(s) : "d"(x), "d"(y))
#define ADD(s, y) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(s) : "d"(y))
#define MUL(s, y) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(s) : "d"(y))
#define IADD(a, b) asm volatile("add.s32 %0, %0, %1;" : "+r"(a) : "r"(b))
#define IMAD(a, b, c) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c))
#define LOP(a, b) asm volatile("xor.b32 %0, %0, %1;" : "+r"(a) : "r"(b))

// MODE 0: CH DFMA chains + NI integer adds per iteration; MODE 1: + NI LDS.128 per iteration; MODE 2: + NI MUFU.RCP64H
// MODE 3: CH DFMA chains + NI IMAD; MODE 4: CH DADD; MODE 5: CH DMUL
template <int MODE, int CH, int NI>
__global__ void __launch_bounds__(256, 2) k(double *out, int iters, double x, double y, int ia, int ib) {
  __shared__ double2 sm[512];
  double s[CH];
  int a[NI > 0 ? NI : 1];
  double r[NI > 0 ? NI : 1];
#pragma unroll
  for (int i = 0; i < CH; i++) s[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < (NI > 0 ? NI : 1); i++) { a[i] = ia + i + threadIdx.x; r[i] = 0.0; }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) sm[i] = make_double2(x, y);
  __syncthreads();
  const double2 *sp = sm + (threadIdx.x & 1);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
      if (MODE == 4) ADD(s[i], y);
      else if (MODE == 5) MUL(s[i], x);
      else FMA(s[i], x, y);
      // spread the NI extra instructions evenly between the fp64 instructions
      if (NI > 0 && (i * NI) / CH != ((i + 1) * NI) / CH) {
        const int e = (i * NI) / CH;
        if (MODE == 0) IADD(a[e], ib);
        else if (MODE == 3) IMAD(a[e], ib, ia);
        else if (MODE == 1) {
          double2 v;
          asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(sp + 2 * e)));
          r[e] += v.x;   // one extra DADD per load (counted below)
        } else if (MODE == 2) {
          double v;
          asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(v) : "d"(s[i]));
          a[e] ^= __double2hiint(v);
        }
      }
    }
    if (NI > CH) {   // more extra instructions than fp64 ones: the remainder as a block
#pragma unroll
      for (int e = CH; e < NI; e++) { if (MODE == 0) IADD(a[e], ib); else if (MODE == 3) IMAD(a[e], ib, ia); }
    }
  }
  double t = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) t += s[i];
#pragma unroll
  for (int i = 0; i < (NI > 0 ? NI : 1); i++) t += a[i] + r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE, int CH, int NI> void run(const char *name, int ctasPerSm, int threads) {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int iters = 20000, grid = sms * ctasPerSm;
  double *out;
  cudaMalloc(&out, sizeof(double) * grid * threads);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE, CH, NI><<<grid, threads>>>(out, 100, 1.0000001, 1e-9, 3, 5);
  cudaEventRecord(a);
  k<MODE, CH, NI><<<grid, threads>>>(out, iters, 1.0000001, 1e-9, 3, 5);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  const double warpsPerSmsp = ctasPerSm * (threads / 32) / 4.0;
  const int fp64PerIter = CH + (MODE == 1 ? NI : 0);
  const double cycles = ms * 1e-3 * khz * 1e3;
  const double perSmspFp64 = warpsPerSmsp * iters * (double)fp64PerIter;
  printf("%-58s %7.3f ms  %6.3f cycles per fp64 warp-instr per sub-partition (%.2f warps/sub-partition, %d fp64 + %d other per iter)\n",
         name, ms, cycles / perSmspFp64, warpsPerSmsp, fp64PerIter, NI);
  cudaFree(out);
}

int main() {
  // (a) shadow issue
  run<0, 10, 0>("DFMA x10", 2, 256);
  run<0, 10, 5>("DFMA x10 + 5 IADD", 2, 256);
  run<0, 10, 10>("DFMA x10 + 10 IADD", 2, 256);
  run<0, 10, 15>("DFMA x10 + 15 IADD", 2, 256);
  run<0, 10, 20>("DFMA x10 + 20 IADD", 2, 256);
  run<3, 10, 10>("DFMA x10 + 10 IMAD", 2, 256);
  run<1, 10, 2>("DFMA x10 + 2 (LDS.128 + DADD)", 2, 256);
  run<1, 10, 5>("DFMA x10 + 5 (LDS.128 + DADD)", 2, 256);
  run<2, 10, 2>("DFMA x10 + 2 MUFU.RCP64H", 2, 256);
  run<2, 10, 5>("DFMA x10 + 5 MUFU.RCP64H", 2, 256);
  run<2, 10, 10>("DFMA x10 + 10 MUFU.RCP64H", 2, 256);
  // (b) latency: one dependent chain, one warp per sub-partition (1 CTA of 128 threads per SM)
  run<0, 1, 0>("DFMA latency (1 chain, 1 warp/sub-partition)", 1, 128);
  run<4, 1, 0>("DADD latency", 1, 128);
  run<5, 1, 0>("DMUL latency", 1, 128);
  run<0, 2, 0>("DFMA 2 chains, 1 warp", 1, 128);
  run<0, 4, 0>("DFMA 4 chains, 1 warp", 1, 128);
  run<0, 8, 0>("DFMA 8 chains, 1 warp", 1, 128);
  // (c) occupancy sensitivity: 2 and 8 warps per sub-partition with 10 chains
  run<0, 10, 0>("DFMA x10, 2 warps/sub-partition", 1, 256);
  run<0, 10, 10>("DFMA x10 + 10 IADD, 2 warps/sub-partition", 1, 256);
  run<0, 10, 10>("DFMA x10 + 10 IADD, 8 warps/sub-partition", 4, 256);
  return 0;
}

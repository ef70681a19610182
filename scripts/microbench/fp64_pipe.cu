// fp64 issue-rate microbenchmark for the roofline discussion in profiles/: how many fp64 warp instructions per cycle and
// SM sub-partition does B200 sustain for (a) independent DFMA chains, (b) independent DADD chains, (c) Kahan steps
// (4 dependent DADD + 1 DMUL, 10 chains per thread -- the shape of pass 1 of k_eval_staged), at 16 warps per SM?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_pipe fp64_pipe.cu ; run: ./fp64_pipe
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int CH>
__global__ void __launch_bounds__(256, 2) k(double *out, int iters, double x, double y) {
  double s[CH], c[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) { s[i] = threadIdx.x * 1e-3 + i; c[i] = 0.0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
      if (MODE == 0) s[i] = __fma_rn(s[i], x, y);
      else if (MODE == 1) s[i] = __dadd_rn(s[i], y);
      else {
        const double v = __dmul_rn(x, y + i);          // stands for r * prior
        const double yy = __dsub_rn(v, c[i]);
        const double t = __dadd_rn(s[i], yy);
        c[i] = __dsub_rn(__dsub_rn(t, s[i]), yy);
        s[i] = t;
      }
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) r += s[i] + c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE, int CH> void run(const char *name, int perIter) {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int iters = 20000, grid = sms * 2;
  double *out;
  cudaMalloc(&out, sizeof(double) * grid * 256);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE, CH><<<grid, 256>>>(out, 100, 1.0000001, 1e-9);
  cudaEventRecord(a);
  k<MODE, CH><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  const double warpInstr = (double)grid * 8 * iters * CH * perIter;      // fp64 warp instructions
  const double cycles = ms * 1e-3 * khz * 1e3;                           // SM cycles (nominal max clock)
  printf("%-34s %8.3f ms  %.3f fp64 warp-instr / cycle / sub-partition (1/%.2f cycles)\n", name, ms,
         warpInstr / (sms * 4.0) / cycles, (sms * 4.0) * cycles / warpInstr);
  cudaFree(out);
}

int main() {
  run<0, 10>("DFMA, 10 independent chains", 1);
  run<1, 10>("DADD, 10 independent chains", 1);
  run<2, 10>("Kahan step, 10 chains (pass 1)", 5);
  run<2, 5>("Kahan step, 5 chains", 5);
  run<2, 20>("Kahan step, 20 chains", 5);
  return 0;
}

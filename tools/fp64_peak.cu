// FP64 pipe microbenchmark for the roofline of the FP64-bound kernels (MEASURED_PEAKS.json has
// no FP64 entry).  Independent chains of DFMA (2 flop) and of DMUL/DADD pairs (what -fmad=false code issues) per
// thread, enough warps to saturate every SM.  Prints one JSON line.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CHAINS = 8;
constexpr int ITERS = 4096;

template <bool FMA> __global__ void __launch_bounds__(256) k(double* out, double a, double b)
{
  double x[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) x[i] = threadIdx.x * 1e-3 + i;
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (FMA) x[i] = fma(x[i], a, b);
      else x[i] = __dadd_rn(__dmul_rn(x[i], a), b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

template <bool FMA> double run(int blocks)
{
  double* d;
  cudaMalloc(&d, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) k<FMA><<<blocks, 256>>>(d, 0.999999, 1e-9);
  float best = 1e30f;
  for (int r = 0; r < 10; ++r) {
    cudaEventRecord(e0);
    k<FMA><<<blocks, 256>>>(d, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaFree(d);
  const double flop = 2.0 * CHAINS * ITERS * 256.0 * blocks;
  return flop / (best * 1e-3) / 1e12;
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 8 * 4;
  const double fma = run<true>(blocks), nofma = run<false>(blocks);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmul_dadd_tflops\": %.2f, \"sm_clock_mhz\": %.0f, "
         "\"dfma_per_clk_per_sm\": %.1f, \"how\": \"8 independent chains x 4096 iterations per thread, 256-thread blocks, "
         "32 blocks per SM, best of 10, CUDA events\"}\n",
         p.name, p.multiProcessorCount, fma, nofma, clk / 1e3, fma * 1e12 / 2.0 / p.multiProcessorCount / (clk * 1e3));
  return 0;
}

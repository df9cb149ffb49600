// FP64 pipe micro-benchmark (B200): dependent DFMA latency and independent DFMA / DADD+DMUL throughput.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = __fma_rn(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmuladd_kernel(double* out, int iters, double a, double b) {   // -fmad=false style: separate mul and add
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = __dadd_rn(__dmul_rn(x[i], a), b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
float time_kernel(K k, int blocks, int threads, double* out, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 64 * 1024);
  const int iters = 20000;
  // latency: one warp, one dependent chain
  float ms = time_kernel(dfma_kernel<1>, 1, 32, out, iters);
  printf("{\"sms\": %d, \"clock_mhz\": %.0f,\n", sms, khz / 1e3);
  printf(" \"dfma_dependent_latency_cycles\": %.2f,\n", ms * 1e-3 * khz * 1e3 / iters);
  ms = time_kernel(dmuladd_kernel<1>, 1, 32, out, iters);
  printf(" \"dmul_dadd_dependent_latency_cycles\": %.2f,\n", ms * 1e-3 * khz * 1e3 / iters);
  // throughput: full chip, 8 independent chains per thread, 1024 threads per SM x 2 blocks
  ms = time_kernel(dfma_kernel<8>, sms * 2, 1024, out, iters);
  double flops = 2.0 * 8 * iters * (double)sms * 2 * 1024 / (ms * 1e-3);
  printf(" \"dfma_tflops\": %.2f, \"dfma_per_clk_per_sm\": %.1f,\n", flops / 1e12, flops / 2 / sms / (khz * 1e3));
  ms = time_kernel(dmuladd_kernel<8>, sms * 2, 1024, out, iters);
  flops = 2.0 * 8 * iters * (double)sms * 2 * 1024 / (ms * 1e-3);
  printf(" \"dmul_dadd_tflops\": %.2f, \"dmul_dadd_ops_per_clk_per_sm\": %.1f}\n", flops / 1e12, flops / sms / (khz * 1e3));
  return 0;
}

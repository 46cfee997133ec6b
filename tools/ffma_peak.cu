// FP32 FFMA peak of the device (the roofline denominator of the CUDA-core `fp32` mode):
// every thread runs 16 independent FMA chains; 2 resident CTAs of 1024 threads per SM.   nvcc -arch=sm_100a -O3 tools/ffma_peak.cu -o tools/bin/ffma_peak
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(1024, 2) ffma(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-6f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 12345.678f) out[0] = s;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float* d; cudaMalloc(&d, 4);
  const int blocks = p.multiProcessorCount * 2, iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0, sustained = 0;
  for (int rep = 0; rep < 12; ++rep) {
    cudaEventRecord(e0);
    ffma<<<blocks, 1024>>>(d, iters, 1.0000001f, 1e-7f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 16 * 8 * (double)iters * 1024.0 * blocks / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
    if (rep >= 6) sustained += tf / 6;
  }
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"fp32_ffma_tflops_burst\": %.1f, \"fp32_ffma_tflops_sustained\": %.1f, \"sms\": %d, \"max_sm_mhz\": %d, \"nominal\": \"%d SMs x 128 FMA/clk x 2 x %.3f GHz = %.1f TFLOP/s\"}\n",
         best, sustained, p.multiProcessorCount, clk / 1000, p.multiProcessorCount, clk / 1e6, p.multiProcessorCount * 256.0 * clk / 1e9);
  return 0;
}

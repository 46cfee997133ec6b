// TMEM -> register bandwidth probe (tcgen05.ld.32x32b.x32): how many bytes per clock can the epilogue warps pull?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I deepphysinet_b200/csrc -I tools tools/tmem_probe.cu -o tools/bin/tmem_probe
#include <cstdio>
#include <cstdlib>
#include "dpn_umma.cuh"
#include "probe_extra.cuh"
using namespace dpn::umma;

__global__ void __launch_bounds__(512, 1) probe(long long* out, float* sink, int nwarps, int reps, int batch) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tbase, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t t = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < reps; ++i) {
      uint32_t ra[32], rb[32];
      if (batch == 1) {
        float v[32];
        tmem_ld32(t + ((i * 32 + (warp >> 2) * 128) & 511), v);
        acc += v[0] + v[31];
      } else {
        tmem_ld32_issue(t + ((i * 64) & 511), ra);
        tmem_ld32_issue(t + ((i * 64 + 32) & 511), rb);
        tmem_ld_wait(ra); tmem_ld_wait(rb);
        acc += __uint_as_float(ra[0]) + __uint_as_float(rb[31]);
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink[threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
  long long* d; float* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4096);
  for (int batch = 1; batch <= 2; ++batch)
    for (int nw : {1, 4, 8, 16}) {
      int reps = 512;
      probe<<<1, 512>>>(d, s, nw, reps, batch);
      cudaError_t e = cudaDeviceSynchronize();
      long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      double bytes = (double)nw * reps * (batch == 1 ? 1 : 2) * 32 * 32 * 4;
      printf("batch %d warps %2d: %lld cycles, %.1f B/clk per SM, %.1f cycles per x32 load per warp (%s)\n", batch, nw, c, bytes / c,
             (double)c / (reps * (batch == 1 ? 1 : 2)), cudaGetErrorString(e));
    }
  return 0;
}

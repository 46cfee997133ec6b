// Hardware probe: does a swizzled K-major operand layout feed tcgen05.mma (SS form) faster than the no-swizzle core-matrix
// layout dpn_tc.cu uses?  Answer (profiles/r01g_umma_sw_probe.txt): the no-swizzle layout already runs at 128.6 cycles per
// M=128 N=256 K=16 MMA = the tensor-pipe floor when the issue loop is lean (the 152.4 of tools/umma_probe.cu mode 5 was that
// probe's branchy single-lane loop).  The swizzled modes here compute correct results but their timings are issue-bound by the
// runtime divisions in sw_desc() - do not read them as hardware rates.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I deepphysinet_b200/csrc -I tools tools/umma_sw_probe.cu -o tools/bin/umma_sw_probe
//   ./umma_sw_probe <swizzle bytes: 0 | 32 | 64 | 128> <reps>
// Layout of a [rows x 64] 16-bit tile with swizzle width Wb: K is cut into slabs of Wb/2 elements; inside a slab row r starts at
// r * Wb and the 16-byte chunk index is XORed with the address bits [7, 7 + log2(Wb/16)) (the Swizzle<B,4,3> pattern), slabs are
// rows * Wb bytes apart.  Descriptor: start address = slab + byte offset of the first k of the MMA, SBO = 8 * Wb, layout type
// 2 / 4 / 6 for 128 / 64 / 32 bytes.  Wb = 0: the no-swizzle layout of dpn_umma.cuh:tile_off.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>

#include "dpn_umma.cuh"
#include "probe_extra.cuh"

using namespace dpn::umma;

static inline uint16_t f2bf(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  uint32_t r = u + 0x7FFF + ((u >> 16) & 1);
  return (uint16_t)(r >> 16);
}
static inline float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

constexpr int M = 128, N = 256, K = 64;

__host__ __device__ inline uint32_t sw_off(int rows, int r, int k, int Wb) {
  if (Wb == 0) return tile_off(rows, r, k);
  const int Ke = Wb / 2, kb = k / Ke, kin = k % Ke;
  uint32_t raw = (uint32_t)r * Wb + kin * 2;
  raw ^= ((raw >> 7) & (uint32_t)(Wb / 16 - 1)) << 4;
  return (uint32_t)kb * rows * Wb + raw;
}
__device__ inline uint64_t sw_desc(uint32_t tile, int rows, int ks, int Wb) {
  if (Wb == 0) return smem_desc(tile + ks * 2 * rows * 16, rows * 16, 128);
  const int Ke = Wb / 2, kb = (ks * 16) / Ke, kin = (ks * 16) % Ke;
  const uint32_t addr = tile + kb * rows * Wb + kin * 2;
  const uint64_t type = Wb == 128 ? 2 : Wb == 64 ? 4 : 6;
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)((8 * Wb) >> 4) << 32) | (1ull << 46) | (type << 61);
}

struct Params { const uint16_t *At, *Bt; float* D; long long* cycles; int Wb, reps; };

__global__ void __launch_bounds__(160, 1) probe(Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sA = smem;
  uint8_t* sB = smem + M * K * 2;
  if (tid == 0) { mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 4) tmem_alloc(&tmem_base, 256);
  for (uint32_t i = tid; i < M * K * 2 / 16; i += blockDim.x) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(p.At)[i];
  for (uint32_t i = tid; i < N * K * 2 / 16; i += blockDim.x) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(p.Bt)[i];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base;
  if (tid == 128) {
    const uint32_t idesc = idesc_bf16(N, 0, 0);
    const long long t0 = clock64();
    for (int rep = 0; rep < p.reps; ++rep)
      for (int ks = 0; ks < K / 16; ++ks)
        mma_bf16(tbase, sw_desc(smem_u32(sA), M, ks, p.Wb), sw_desc(smem_u32(sB), N, ks, p.Wb), idesc, (rep | ks) ? 1u : 0u);
    mma_commit(&bar_mma);
    mbar_wait(&bar_mma, 0);
    *p.cycles = clock64() - t0;
  }
  if (warp < 4) {
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int j = 0; j < 32; ++j) p.D[(size_t)tid * N + c0 + j] = v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tbase, 256);
}

int main(int argc, char** argv) {
  const int Wb = argc > 1 ? atoi(argv[1]) : 128, reps = argc > 2 ? atoi(argv[2]) : 1;
  std::vector<float> A(M * K), B((size_t)N * K);
  srand(321);
  for (auto& v : A) v = bf2f(f2bf((rand() % 2001 - 1000) / 1000.f));
  for (auto& v : B) v = bf2f(f2bf((rand() % 2001 - 1000) / 1000.f));
  std::vector<uint16_t> At(M * K), Bt((size_t)N * K);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) At[sw_off(M, m, k, Wb) / 2] = f2bf(A[m * K + k]);
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) Bt[sw_off(N, n, k, Wb) / 2] = f2bf(B[(size_t)n * K + k]);
  uint16_t *dA, *dB; float* dD; long long* dC;
  cudaMalloc(&dA, At.size() * 2); cudaMalloc(&dB, Bt.size() * 2); cudaMalloc(&dD, (size_t)M * N * 4); cudaMalloc(&dC, 8);
  cudaMemcpy(dA, At.data(), At.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bt.data(), Bt.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, (size_t)M * N * 4);
  Params p{dA, dB, dD, dC, Wb, reps};
  const size_t smem = (size_t)(M + N) * K * 2 + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 160, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("swizzle %d: CUDA error %s\n", Wb, cudaGetErrorString(e)); return 2; }
  std::vector<float> D((size_t)M * N);
  long long cyc = 0;
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[(size_t)n * K + k];
      s *= reps;
      maxerr = fmax(maxerr, fabs(s - D[(size_t)m * N + n]));
      maxref = fmax(maxref, fabs(s));
    }
  printf("swizzle %3d B: N=%d K=%d reps=%d  max|err|=%.4g (max|ref|=%.4g)  %s  %.1f cycles per MMA (%d MMAs)\n", Wb, N, K, reps, maxerr, maxref,
         maxerr < 1e-3 * maxref * (reps > 1 ? 10 : 1) ? "PASS" : "FAIL", (double)cyc / (reps * (K / 16)), reps * (K / 16));
  return 0;
}

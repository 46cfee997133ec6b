// Hardware probe for the hand-rolled tcgen05 plumbing in deepphysinet_b200/csrc/dpn_umma.cuh.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I deepphysinet_b200/csrc -I tools tools/umma_probe.cu -o gpurun_out/umma_probe
//   ./umma_probe <mode>
// mode 0  K-major A,B   (LBO = k-core stride, SBO = 8-row-group stride)    <- what the kernels assume
// mode 1  K-major A,B   with LBO/SBO swapped                                (must FAIL if 0 is right)
// mode 2  MN-major A,B  (SBO = 8-element MN-group stride, LBO = 8-k-group stride)
// mode 3  MN-major A,B  with LBO/SBO swapped
// mode 4  like 0, B tile fetched with cp.async.bulk from a pre-tiled global image
// mode 5  throughput: cycles per tcgen05.mma (M=128,N=256,K=16) issued back to back, no-swizzle operands
// mode 6  like 0 with N=192
// mode 7  mixed: A K-major, B MN-major
#include <cstdio>
#include <cstdlib>
#include <cmath>
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

constexpr int M = 128;

struct Params {
  const uint16_t* Atile;   // pre-tiled images (layout (*)), A: rows x cols as used by the mode
  const uint16_t* Btile;
  float* D;                // [128][N]
  long long* cycles;
  int N, K, mode, reps;
};

__global__ void __launch_bounds__(160, 1) probe(Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mma, bar_load;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = p.N, K = p.K;
  const bool mn = (p.mode == 2 || p.mode == 3);
  const bool b_mn = mn || p.mode == 7;
  uint8_t* sA = smem;
  uint8_t* sB = smem + M * K * 2;
  const uint32_t bytesA = M * K * 2, bytesB = N * K * 2;
  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_load, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(&tmem_base, 256);
  // operands -> smem
  for (uint32_t i = tid; i < bytesA / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(p.Atile)[i];
  if (p.mode != 4)
    for (uint32_t i = tid; i < bytesB / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(p.Btile)[i];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base;
  if (tid == 128) {   // warp 4 lane 0: producer + MMA issuer
    if (p.mode == 4) {
      mbar_arrive_expect_tx(&bar_load, bytesB);
      bulk_g2s(sB, p.Btile, bytesB, &bar_load);
      mbar_wait(&bar_load, 0);
      tc_fence_after();
    }
    const uint32_t idesc = idesc_bf16(N, mn ? 1 : 0, b_mn ? 1 : 0);
    long long t0 = clock64();
    for (int rep = 0; rep < p.reps; ++rep) {
      for (int ks = 0; ks < K / 16; ++ks) {
        uint64_t ad, bd;
        if (!mn) {
          uint32_t lboA = M * 16, sbo = 128;
          if (p.mode == 1) { ad = smem_desc(smem_u32(sA) + ks * 2 * M * 16, sbo, lboA); }
          else ad = smem_desc(smem_u32(sA) + ks * 2 * M * 16, lboA, sbo);
        } else {
          // A tile is [K rows][M cols]: MN-group stride = K*16, k-group stride = 128
          if (p.mode == 3) ad = smem_desc(smem_u32(sA) + ks * 2 * 128, (uint32_t)K * 16, 128);
          else ad = smem_desc(smem_u32(sA) + ks * 2 * 128, 128, (uint32_t)K * 16);
        }
        if (!b_mn) {
          uint32_t lboB = N * 16, sbo = 128;
          if (p.mode == 1) bd = smem_desc(smem_u32(sB) + ks * 2 * N * 16, sbo, lboB);
          else bd = smem_desc(smem_u32(sB) + ks * 2 * N * 16, lboB, sbo);
        } else {
          if (p.mode == 3) bd = smem_desc(smem_u32(sB) + ks * 2 * 128, (uint32_t)K * 16, 128);
          else bd = smem_desc(smem_u32(sB) + ks * 2 * 128, 128, (uint32_t)K * 16);
        }
        mma_bf16(tbase, ad, bd, idesc, (rep | ks) ? 1u : 0u);
      }
    }
    mma_commit(&bar_mma);
    mbar_wait(&bar_mma, 0);
    long long t1 = clock64();
    if (p.cycles) *p.cycles = t1 - t0;
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
  int mode = argc > 1 ? atoi(argv[1]) : 0;
  int N = (mode == 6) ? 192 : 256, K = 64, reps = 1;
  if (mode == 5) { reps = 256; }
  const bool mn = (mode == 2 || mode == 3), b_mn = mn || mode == 7;
  std::vector<float> A(M * K), B((size_t)N * K);
  srand(123 + mode);
  for (auto& v : A) v = bf2f(f2bf((rand() % 2001 - 1000) / 1000.f));
  for (auto& v : B) v = bf2f(f2bf((rand() % 2001 - 1000) / 1000.f));
  // tiled images per layout (*): K-major: rows = m (or n), k = contraction; MN-major: rows = contraction, k = m (or n)
  std::vector<uint16_t> At(M * K), Bt((size_t)N * K);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      uint32_t off = mn ? tile_off(K, k, m) : tile_off(M, m, k);
      At[off / 2] = f2bf(A[m * K + k]);
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      uint32_t off = b_mn ? tile_off(K, k, n) : tile_off(N, n, k);
      Bt[off / 2] = f2bf(B[(size_t)n * K + k]);
    }
  uint16_t *dA, *dB; float* dD; long long* dC;
  cudaMalloc(&dA, At.size() * 2); cudaMalloc(&dB, Bt.size() * 2); cudaMalloc(&dD, (size_t)M * N * 4); cudaMalloc(&dC, 8);
  cudaMemcpy(dA, At.data(), At.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bt.data(), Bt.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, (size_t)M * N * 4);
  Params p{dA, dB, dD, dC, N, K, mode, reps};
  size_t smem = (size_t)(M + N) * K * 2 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 160, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 2; }
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
  printf("mode %d: N=%d K=%d reps=%d  max|err|=%.4g (max|ref|=%.4g)  %s", mode, N, K, reps, maxerr, maxref,
         maxerr < 1e-3 * maxref * (reps > 1 ? 10 : 1) ? "PASS" : "FAIL");
  if (mode == 5) printf("  cycles=%lld -> %.1f cycles per MMA (%d MMAs)", cyc, (double)cyc / (reps * (K / 16)), reps * (K / 16));
  printf("\n");
  return 0;
}

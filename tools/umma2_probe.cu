// Hardware probe for the 2-CTA (cta_group::2) tcgen05 plumbing: cluster of two CTAs, M = 256 (128 rows per CTA),
// each CTA holds its A tile [128 x K] and HALF of B ([N/2 x K], rows r*N/2 ...), the leader issues the MMAs, the commit is
// multicast to both CTAs, each CTA reads its own D [128 x N] from its TMEM.  Also exercises a remote mbarrier arrive
// (peer -> leader) before the leader may issue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I deepphysinet_b200/csrc -I tools tools/umma2_probe.cu -o tools/bin/umma2_probe
//   ./umma2_probe <mode>     mode 0: both CTAs alloc/dealloc with cta_group::2 ; mode 1: N = 192 ; mode 2: fp16 operands ; mode 3: throughput
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "dpn_umma.cuh"
#include "probe_extra.cuh"

using namespace dpn::umma;

static inline uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); uint32_t r = u + 0x7FFF + ((u >> 16) & 1); return (uint16_t)(r >> 16); }
static inline float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

constexpr int M = 128;
struct Params { const uint16_t* Atile; const uint16_t* Btile; float* D; int N, K, mode, reps; long long* cycles; };

__device__ __forceinline__ void tmem_alloc2(uint32_t* dst, uint32_t n) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(n) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t a, uint32_t n) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(a), "r"(n) : "memory");
}
__device__ __forceinline__ void mma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b),
               "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void remote_arrive(uint64_t* bar, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(160, 1) probe(Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mma, bar_peer;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int N = p.N, K = p.K, Nh = N / 2;
  const bool f16 = p.mode == 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + M * K * 2;
  const uint32_t bytesA = M * K * 2, bytesB = Nh * K * 2;
  if (tid == 0) { mbar_init(&bar_mma, 1); mbar_init(&bar_peer, 1); fence_barrier_init(); }
  if (warp == 4) tmem_alloc2(&tmem_base, 256);
  const uint16_t* Ag = p.Atile + (size_t)rank * M * K;            // this CTA's 128 rows
  const uint16_t* Bg = p.Btile + (size_t)rank * Nh * K;           // this CTA's half of the N rows (pre-tiled per half)
  for (uint32_t i = tid; i < bytesA / 16; i += blockDim.x) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(Ag)[i];
  for (uint32_t i = tid; i < bytesB / 16; i += blockDim.x) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(Bg)[i];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tbase = tmem_base;
  if (tid == 128) {
    if (rank == 1) {
      remote_arrive(&bar_peer, 0);                               // "my operands are in place" -> leader
    } else {
      wait_cluster(&bar_peer, 0);
      tc_fence_after();
      const uint32_t idesc = idesc_16(f16, N, 0, 0, 256);
      const long long t0 = clock64();
      for (int rep = 0; rep < p.reps; ++rep)
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t ad = smem_desc(smem_u32(sA) + ks * 2 * M * 16, M * 16, 128);
        const uint64_t bd = smem_desc(smem_u32(sB) + ks * 2 * Nh * 16, Nh * 16, 128);
        mma2(tbase, ad, bd, idesc, (rep | ks) ? 1u : 0u);
      }
      commit2(&bar_mma, 3);
      mbar_wait(&bar_mma, 0);
      if (p.cycles) *p.cycles = clock64() - t0;
    }
  }
  if (warp < 4) {
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int j = 0; j < 32; ++j) p.D[((size_t)rank * M + tid) * N + c0 + j] = v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 4) tmem_dealloc2(tbase, 256);
}

int main(int argc, char** argv) {
  int mode = argc > 1 ? atoi(argv[1]) : 0;
  int N = (mode == 1) ? 192 : 256, K = 64, Nh = N / 2;
  std::vector<float> A(2 * M * K), B((size_t)N * K);
  srand(321 + mode);
  for (auto& v : A) v = bf2f(f2bf((rand() % 2001 - 1000) / 1000.f));
  for (auto& v : B) v = bf2f(f2bf((rand() % 2001 - 1000) / 1000.f));
  if (mode == 2) {                                               // values exactly representable in fp16 AND bf16: multiples of 1/8
    for (auto& v : A) v = (float)(rand() % 17 - 8) / 8.f;
    for (auto& v : B) v = (float)(rand() % 17 - 8) / 8.f;
  }
  auto enc = [&](float f) -> uint16_t {
    if (mode != 2) return f2bf(f);
    // fp16 encode of a small exactly representable value
    if (f == 0.f) return 0;
    uint16_t s = f < 0 ? 0x8000 : 0; float a = fabsf(f); int e = 0; float m = frexpf(a, &e);   // a = m * 2^e, m in [0.5,1)
    int E = e - 1 + 15; uint16_t frac = (uint16_t)((m * 2.f - 1.f) * 1024.f);
    return s | (uint16_t)(E << 10) | frac;
  };
  std::vector<uint16_t> At(2 * M * K), Bt((size_t)N * K);
  for (int r = 0; r < 2; ++r)
    for (int m = 0; m < M; ++m)
      for (int k = 0; k < K; ++k) At[(size_t)r * M * K + tile_off(M, m, k) / 2] = enc(A[((size_t)r * M + m) * K + k]);
  for (int r = 0; r < 2; ++r)
    for (int n = 0; n < Nh; ++n)
      for (int k = 0; k < K; ++k) Bt[(size_t)r * Nh * K + tile_off(Nh, n, k) / 2] = enc(B[((size_t)r * Nh + n) * K + k]);
  uint16_t *dA, *dB; float* dD;
  cudaMalloc(&dA, At.size() * 2); cudaMalloc(&dB, Bt.size() * 2); cudaMalloc(&dD, (size_t)2 * M * N * 4);
  cudaMemcpy(dA, At.data(), At.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bt.data(), Bt.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, (size_t)2 * M * N * 4);
  int reps = mode == 3 ? 256 : 1;
  long long* dC; cudaMalloc(&dC, 8); cudaMemset(dC, 0, 8);
  Params p{dA, dB, dD, N, K, mode, reps, dC};
  size_t smem = (size_t)(M + Nh) * K * 2 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<2, 160, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 2; }
  std::vector<float> D((size_t)2 * M * N);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 2 * M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
      s *= reps;
      maxerr = fmax(maxerr, fabs(s - D[(size_t)m * N + n]));
      maxref = fmax(maxref, fabs(s));
    }
  long long cyc = 0; cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
  printf("umma2 mode %d: M=256 (2 CTAs) N=%d K=%d reps=%d  max|err|=%.4g (max|ref|=%.4g)  %s  %.1f cycles per pair-MMA\n", mode, N, K, reps, maxerr, maxref,
         maxerr < 1e-3 * maxref * (reps > 1 ? 10 : 1) ? "PASS" : "FAIL", (double)cyc / (reps * (K / 16)));
  return 0;
}

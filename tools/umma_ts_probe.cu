// Hardware probe for the TS form of tcgen05.mma: A operand in TENSOR MEMORY (written by the CTA's own threads with tcgen05.st),
// B operand in shared memory (no-swizzle K-major tile as everywhere in dpn_tc.cu), D in TMEM.
// A [128 x K] bf16 is laid out lane = row, 32-bit column j holds elements (2j, 2j+1) -> K/2 columns; D uses columns 256...
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I deepphysinet_b200/csrc -I tools tools/umma_ts_probe.cu -o tools/bin/umma_ts_probe
//   ./umma_ts_probe <mode>     mode 0: correctness (K = 64) ; mode 1: throughput (256 x K = 64, back to back)
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
struct Params { const uint16_t* A; const uint16_t* Btile; float* D; long long* cycles; int N, K, reps; };

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void probe_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem),
               "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(160, 1) probe(Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = p.N, K = p.K;
  uint8_t* sB = smem;
  if (tid == 0) { mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 4) tmem_alloc(&tmem_base, 512);
  for (uint32_t i = tid; i < (uint32_t)N * K * 2 / 16; i += blockDim.x) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(p.Btile)[i];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base;
  if (warp < 4) {                                   // thread = row: pack my K elements two per column and store them to TMEM columns 0..K/2
    uint32_t r[32];
    for (int j = 0; j < K / 2; ++j) r[j] = (uint32_t)p.A[(size_t)tid * K + 2 * j] | ((uint32_t)p.A[(size_t)tid * K + 2 * j + 1] << 16);
    for (int j = K / 2; j < 32; ++j) r[j] = 0;
    tmem_st32(tbase + ((uint32_t)(warp * 32) << 16), r);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 128) {
    const uint32_t idesc = idesc_bf16(N, 0, 0);
    const long long t0 = clock64();
    for (int rep = 0; rep < p.reps; ++rep)
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t bd = smem_desc(smem_u32(sB) + ks * 2 * N * 16, N * 16, 128);
        probe_mma_ts(tbase + 256, tbase + ks * 8, bd, idesc, (rep | ks) ? 1u : 0u);       // 16 bf16 = 8 columns per K step
      }
    mma_commit(&bar_mma);
    mbar_wait(&bar_mma, 0);
    if (p.cycles) *p.cycles = clock64() - t0;
  }
  if (warp < 4) {
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + 256 + c0, v);
      for (int j = 0; j < 32; ++j) p.D[(size_t)tid * N + c0 + j] = v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tbase, 512);
}

int main(int argc, char** argv) {
  int mode = argc > 1 ? atoi(argv[1]) : 0;
  int N = 256, K = 64, reps = mode == 1 ? 256 : 1;
  std::vector<float> A(M * K), B((size_t)N * K);
  srand(77 + mode);
  for (auto& v : A) v = bf2f(f2bf((rand() % 2001 - 1000) / 1000.f));
  for (auto& v : B) v = bf2f(f2bf((rand() % 2001 - 1000) / 1000.f));
  std::vector<uint16_t> Ar(M * K), Bt((size_t)N * K);
  for (int i = 0; i < M * K; ++i) Ar[i] = f2bf(A[i]);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) Bt[tile_off(N, n, k) / 2] = f2bf(B[(size_t)n * K + k]);
  uint16_t *dA, *dB; float* dD; long long* dC;
  cudaMalloc(&dA, Ar.size() * 2); cudaMalloc(&dB, Bt.size() * 2); cudaMalloc(&dD, (size_t)M * N * 4); cudaMalloc(&dC, 8);
  cudaMemcpy(dA, Ar.data(), Ar.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bt.data(), Bt.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, (size_t)M * N * 4);
  Params p{dA, dB, dD, dC, N, K, reps};
  size_t smem = (size_t)N * K * 2 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 160, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("ts mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 2; }
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
  printf("ts mode %d: A in TMEM, N=%d K=%d reps=%d  max|err|=%.4g (max|ref|=%.4g)  %s  %.1f cycles per MMA\n", mode, N, K, reps, maxerr, maxref,
         maxerr < 1e-3 * maxref * (reps > 1 ? 10 : 1) ? "PASS" : "FAIL", (double)cyc / (reps * (K / 16)));
  return 0;
}

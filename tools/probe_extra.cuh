// Helpers that only the probe programs under tools/ use (they were part of deepphysinet_b200/csrc/dpn_umma.cuh while the round-1
// kernels needed them): software-pipelined TMEM loads, the bf16-only instruction descriptor, the cta_group::2 pair primitives and the
// element offset of layout (*).  Include after dpn_umma.cuh.
#pragma once
#include "dpn_umma.cuh"

namespace dpn {
namespace umma {

// Split form for software pipelining: issue the load of the NEXT 32 columns, work on the current ones, then wait.
// The wait takes the destination registers as in/out operands so that no consumer can be scheduled above it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
// Visits NB consecutive 32-column blocks of this thread's TMEM lane, double-buffered: while f works on block i the
// load of block i+1 is in flight.  f(block_index, float (&v)[32]).
template <int NB, class F>
__device__ __forceinline__ void tmem_for_each_block(uint32_t taddr, F&& f) {
  uint32_t ra[32], rb[32];
  tmem_ld32_issue(taddr, ra);
#pragma unroll 1
  for (int cb = 0; cb < NB; cb += 2) {
    tmem_ld_wait(ra);
    if (cb + 1 < NB) tmem_ld32_issue(taddr + (cb + 1) * 32, rb);
    f(cb, reinterpret_cast<float(&)[32]>(ra));
    if (cb + 1 < NB) {
      tmem_ld_wait(rb);
      if (cb + 2 < NB) tmem_ld32_issue(taddr + (cb + 2) * 32, ra);
      f(cb + 1, reinterpret_cast<float(&)[32]>(rb));
    }
  }
}


// Instruction descriptor for kind::f16: D=f32, A=B=bf16, M=128; major bits: 0 = K-major, 1 = MN-major.
__host__ __device__ constexpr uint32_t idesc_bf16(int n, int a_mn_major, int b_mn_major, int m = 128) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}


// ---- CTA pair (cta_group::2): two CTAs of a cluster on the two SMs of a TPC run ONE M = 256 MMA -----------------------
// Each CTA holds its 128 rows of A and HALF of B (N/2 rows) at the same shared-memory offsets; the leader (cluster rank 0)
// issues, D rows of a CTA land in its own TMEM.  Verified on hardware by tools/umma2_probe.cu.
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {   // one full warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs once all previously issued pair-MMAs have completed
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// arrive on the barrier at the same offset in another CTA of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta_rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(cta_rank));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");   // default .release.cta: a cluster-scope release costs ~700 cycles per call
}
// wait on a local barrier whose arrivals come from another CTA
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ uint32_t ld_remote_u32(const uint32_t* p, uint32_t cta_rank) {
  uint32_t ra, v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(p)), "r"(cta_rank));
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(ra) : "memory");
  return v;
}


__host__ __device__ constexpr uint32_t tile_off(int rows, int r, int k) {   // byte offset of element (r,k)
  return (uint32_t)((k >> 3) * (rows * 16) + r * 16 + (k & 7) * 2);
}


}  // namespace umma
}  // namespace dpn

// DPN_MODE_BF16: tcgen05 / TMEM / bulk-copy path (see dpn_tc.cu).
#pragma once
#include "dpn_fp32.cuh"

namespace dpn {
namespace tc {

constexpr int DEFAULT_CHUNK = 131072;   // points per sample per pass (multiple of 128)

size_t workspace_bytes(int P, int Kn, int B);
int run(const Job& job, cudaStream_t st);

}  // namespace tc
}  // namespace dpn

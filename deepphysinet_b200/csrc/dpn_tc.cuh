// DPN_MODE_BF16: tcgen05 / TMEM / bulk-copy path (see dpn_tc.cu).
#pragma once
#include "dpn_fp32.cuh"

namespace dpn {
namespace tc {

constexpr int DEFAULT_POINTS_IN_FLIGHT = 262144;   // B * chunk: bounds the workspace (~34 KB per point)

int default_chunk(int B);                          // points per sample per pass (multiple of 128)
size_t workspace_bytes(int chunk, int Kn, int B);
int run(const Job& job, cudaStream_t st);

}  // namespace tc
}  // namespace dpn

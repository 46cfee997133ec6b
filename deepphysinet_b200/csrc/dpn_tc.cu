#include "dpn_tc.cuh"
namespace dpn { namespace tc {
size_t workspace_bytes(int P, int Kn, int B) { return f32::workspace_bytes(P < 16384 ? P : 16384, Kn, B); }
int run(const Job& job, cudaStream_t st) { set_error("bf16 mode not built yet"); return DPN_E_UNSUPPORTED; }
}}

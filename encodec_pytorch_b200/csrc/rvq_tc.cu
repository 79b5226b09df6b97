// tcgen05 / TMEM fused multi-stage search (placeholder until the kernel lands in this file).
#include "rvq_common.cuh"
namespace rvq {
int tc_encode(const EncodeArgs& a, cudaStream_t st) { return simt_encode(a, st); }
}  // namespace rvq

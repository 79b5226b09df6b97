// tcgen05 / TMEM fused multi-stage search (placeholder until the kernel lands in this file).
#include "rvq_common.cuh"
namespace rvq {
int tc_encode(const void* pack, int K, int D, const float* x, int64_t sxb, int64_t sxd, int64_t sxt,
              int B, int T, int stage0, int n_q, int64_t* codes, float* quantized, double* sqerr,
              int flags, cudaStream_t st) {
  return simt_encode(pack, K, D, x, sxb, sxd, sxt, B, T, stage0, n_q, codes, quantized, sqerr, flags, st);
}
}  // namespace rvq

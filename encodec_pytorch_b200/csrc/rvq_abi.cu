// C ABI glue: error plumbing, device gate, launch counter, pack / encode dispatch.
#include "rvq_common.cuh"

#include <atomic>
#include <cstdarg>

namespace rvq {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return RVQ_ECUDA;
}

static thread_local unsigned long long* g_counters = nullptr;
unsigned long long* search_counters() { return g_counters; }

static std::atomic<int> g_bound_mode{0};
int pack_bound_mode() { return g_bound_mode.load(std::memory_order_relaxed); }

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// The library is built for sm_100a only: refuse anything else loudly (no fallback of any kind).
int check_device() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_rc = RVQ_ENODEV;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
    return RVQ_ENODEV;
  }
  if (dev == cached_dev) {
    if (cached_rc != RVQ_OK) set_error("current CUDA device is not sm_100 (B200); no fallback path exists");
    return cached_rc;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  cached_dev = dev;
  cached_rc = (major == 10) ? RVQ_OK : RVQ_ENODEV;
  if (cached_rc != RVQ_OK)
    set_error("current CUDA device is sm_%d%d, this library needs sm_100 (B200); no fallback path exists", major, minor);
  return cached_rc;
}

}  // namespace rvq

using namespace rvq;

extern "C" {

int rvq_version(void) { return RVQ_ABI_VERSION; }
int rvq_pack_bound_mode(int mode) { return g_bound_mode.exchange(mode < 0 || mode > 2 ? 0 : mode, std::memory_order_relaxed); }
const char* rvq_last_error(void) { return g_err; }
int rvq_device_ok(void) { return check_device(); }
uint64_t rvq_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

size_t rvq_pack_bytes(int n_q, int K, int D) {
  if (n_q < 0 || K <= 0 || D <= 0) return 0;
  return size_t(kHeaderBytes) + size_t(n_q) * stage_layout(K, D).stride;
}

int rvq_pack(const float* const* embed_ptrs_host, int n_q, int K, int D, void* pack, size_t pack_bytes, void* stream) {
  if (int e = check_device()) return e;
  RVQ_REQUIRE(embed_ptrs_host && pack, "rvq_pack: null pointer");
  RVQ_REQUIRE(n_q >= 0 && K > 0 && D > 0, "rvq_pack: bad shape n_q=%d K=%d D=%d", n_q, K, D);
  if (pack_bytes < rvq_pack_bytes(n_q, K, D)) {
    set_error("rvq_pack: buffer of %zu bytes, need %zu", pack_bytes, rvq_pack_bytes(n_q, K, D));
    return RVQ_ESIZE;
  }
  RVQ_REQUIRE((reinterpret_cast<uintptr_t>(pack) & 255) == 0, "rvq_pack: pack must be 256-byte aligned");
  for (int i = 0; i < n_q; ++i) RVQ_REQUIRE(embed_ptrs_host[i] != nullptr, "rvq_pack: embed pointer %d is null", i);
  return simt_pack(embed_ptrs_host, n_q, K, D, pack, (cudaStream_t)stream);
}

int rvq_encode(const void* pack, int K, int D, const float* x, int64_t sxb, int64_t sxd, int64_t sxt,
               int B, int T, int stage0, int n_q, int64_t* codes, float* quantized, float* residual_out,
               double* stage_sqerr, int flags, void* stream) {
  if (int e = check_device()) return e;
  RVQ_REQUIRE(pack, "rvq_encode: null pack");
  RVQ_REQUIRE(B >= 0 && T >= 0 && n_q >= 0 && stage0 >= 0 && K > 0 && D > 0, "rvq_encode: bad shape");
  if (int64_t(B) * T == 0 || n_q == 0) return RVQ_OK;
  RVQ_REQUIRE(x && codes, "rvq_encode: null pointer");
  RVQ_REQUIRE(int64_t(B) * T < (int64_t(1) << 31), "rvq_encode: more than 2^31 frames in one call");
  cudaStream_t st = (cudaStream_t)stream;
  EncodeArgs a{pack, K, D, x, sxb, sxd, sxt, B, T, stage0, n_q, codes, quantized, residual_out, stage_sqerr, flags};
  const bool want_tc = tc_shape(K, D) && !(flags & RVQ_FLAG_FORCE_EXACT);
  return want_tc ? tc_encode(a, st) : simt_encode(a, st);
}

int rvq_encode_train(const void* pack, int K, int D, const float* x, int64_t sxb, int64_t sxd, int64_t sxt,
                     int B, int T, int stage0, int n_q, int64_t* codes, float* quantized, float* residual_out,
                     double* stage_sqerr, float* counts, float* embed_sum, int flags, void* stream) {
  if (int e = check_device()) return e;
  RVQ_REQUIRE(pack && counts && embed_sum, "rvq_encode_train: null pointer");
  RVQ_REQUIRE(B >= 0 && T >= 0 && n_q >= 0 && stage0 >= 0 && K > 0 && D > 0, "rvq_encode_train: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const bool want_tc = tc_shape(K, D) && !(flags & RVQ_FLAG_FORCE_EXACT);
  if (!want_tc) {      // other shapes: the fp32 search, then the statistics pass over its codes
    if (int e = rvq_encode(pack, K, D, x, sxb, sxd, sxt, B, T, stage0, n_q, codes, quantized, residual_out, stage_sqerr, flags, stream)) return e;
    return rvq_ema_stats(pack, K, D, x, sxb, sxd, sxt, B, T, stage0, n_q, codes, counts, embed_sum, flags, stream);
  }
  RVQ_CUDA(cudaMemsetAsync(counts, 0, size_t(n_q) * K * 4, st));
  RVQ_CUDA(cudaMemsetAsync(embed_sum, 0, size_t(n_q) * K * D * 4, st));
  if (int64_t(B) * T == 0 || n_q == 0) return RVQ_OK;
  RVQ_REQUIRE(x && codes, "rvq_encode_train: null pointer");
  RVQ_REQUIRE(int64_t(B) * T < (int64_t(1) << 31), "rvq_encode_train: more than 2^31 frames in one call");
  EncodeArgs a{pack, K, D, x, sxb, sxd, sxt, B, T, stage0, n_q, codes, quantized, residual_out, stage_sqerr, flags};
  a.ema_counts = counts; a.ema_sum = embed_sum;
  return tc_encode(a, st);
}

int rvq_kmeans_assign(const void* pack, int K, int D, const float* samples, int64_t N, int64_t* buckets, void* stream) {
  if (int e = check_device()) return e;
  RVQ_REQUIRE(pack && (N == 0 || (samples && buckets)), "rvq_kmeans_assign: null pointer");
  RVQ_REQUIRE(N >= 0 && N < (int64_t(1) << 31), "rvq_kmeans_assign: N out of range");
  if (N == 0) return RVQ_OK;
  // samples [N, D] viewed as x[B=1, D, T=N] with strides (0, 1, D)
  // the assignment is the search's contraction: on the tensor cores where the shape allows, exact fp32 re-scores with the
  // direct distance of core_vq.py:86-88 (ties -> lowest index); other shapes take the fp32 SIMT kernel
  EncodeArgs a{pack, K, D, samples, 0, 1, D, 1, (int)N, 0, 1, buckets, nullptr, nullptr, nullptr, RVQ_FLAG_DIRECT_DIST};
  return tc_shape(K, D) ? tc_encode(a, (cudaStream_t)stream) : simt_encode(a, (cudaStream_t)stream);
}

/* debug only (not part of the public header): timeline of CTA 0 of the last tcgen05 encode (RVQ_TC_TRACE builds) */
int rvq_debug_trace(long long* out_host, int n) { cudaDeviceSynchronize(); return rvq::tc_debug_trace(out_host, n); }

int rvq_search_counters(uint64_t* counters_dev) {
  g_counters = reinterpret_cast<unsigned long long*>(counters_dev);
  return RVQ_OK;
}

}  // extern "C"

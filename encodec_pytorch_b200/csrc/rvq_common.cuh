// Shared definitions for the B200 RVQ library: pack layout, error plumbing, launch counter.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rvq_b200.h"

namespace rvq {

// ----------------------------------------------------------------------------------------------
// Pack layout (device memory, caller-allocated, rvq_pack_bytes()):
//   [0, 256)                 header (reserved; the pack is read-only after rvq_pack)
//   per stage s (stride stage_bytes(K, D), 256-B aligned sections):
//     tab32   fp32 [K][D]    row-major copy of embed             (gathers, exact re-score)
//     tab32T  fp32 [D][K]    transposed copy                     (exact SIMT search tiles)
//     cnorm   fp32 [K]       |c_k|^2
//     tc      fp16 image     (only D==128 && K%128==0) K/128 chunks of 36864 B, each the
//                            shared-memory image of a 128-code x 144-K UMMA B operand
//     meta    StageMeta
//     outl    u8 [K]         1 = code excluded from the fp16 image (outlier norm / fp16 range)
//     gab     float2 [K]     per-code score-error coefficients {g16 + a_k, b_k} (StageMeta::percode), then
//             fp16 [K]       g16_k: the coefficient carried by the fp16 image (augmented columns 2 and 4)
// ----------------------------------------------------------------------------------------------
constexpr int kHeaderBytes  = 256;
constexpr int kTcChunkCodes = 128;                 // codes per UMMA N tile
constexpr int kTcKPad       = 144;                 // 128 dims + 16 augmented K columns
constexpr int kTcChunkBytes = kTcChunkCodes * kTcKPad * 2;  // 36864
// image of chunk c: 18 K-groups (8 halves = 16 B each) x 128 code rows; element (row r, group g) at
//   c*36864 + g*2048 + r*16  -> K-major, no swizzle: core matrix = 8 rows x 16 B contiguous,
//   LBO (K direction) = 2048 B, SBO (row direction) = 128 B
constexpr int kTcLBO = kTcChunkCodes * 16;         // 2048
constexpr int kTcSBO = 128;

struct StageMeta {
  // Two-sided bound on the fp16 tensor-core score error of a frame against any live code:
  //   |S_k - s_k| <= delta/2,  delta = margin_coef * |r| + margin_dr * |r - fp16(r)| + margin_abs
  // (|r| may be an upper bound; |r - fp16(r)| is the exact rounding residue of the frame's fp16 operand;
  //  margin_coef covers the rounding of the codes: 2 max_k |(-2c_k) - fp16(-2c_k)|, margin_dr = 2 (2 cref + that)).
  float margin_coef;
  float margin_abs;
  float xlimit;        // frames with |x| >= xlimit take the exact path (outlier codes / fp16 range)
  float cref;          // largest norm among non-outlier codes
  float cmin;          // smallest code norm
  int   n_outliers;    // codes excluded from the fp16 image (provably non-winning under xlimit)
  float cmax_all;      // largest norm among ALL codes (bounds the residual growth of exact-path frames)
  float margin_dr;
  // Per-code bound (tables whose code norms are heterogeneous, e.g. fitted ones: the codes that compete for a frame are
  // the small ones, and the bound above is set by the largest).  With a_k = |fp16(-2c_k) + 2c_k| (+ accumulation) and
  // b_k = 2|c_k| + that:   |S_k - s_k| <= a_k |r| + b_k |r - fp16(r)|,   |r - fp16(r)| <= 2^-11 |r| (+ subnormals).
  // The image carries g16_k >= a_k + 2^-11 b_k in its augmented columns and the frame's operand an upper bound R of
  // |r|, so the tensor core delivers the LOWER bounds T_k = S_k - g16_k R <= s_k directly; for ANY code k',
  // s_winner <= s_k' <= T_k' + (g16_k' + a_k') R + b_k' |r - fp16(r)| + abs_pc: the threshold the candidates' T are held against.
  int   percode;       // 1: the search uses the per-code bound for this stage
  float abs_pc;        // absolute term of the per-code threshold
  float g16max;        // largest g16_k (fallback threshold when the minimum of T is not unique)
};

__host__ __device__ inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }
__host__ __device__ inline bool tc_shape(int K, int D) { return D == 128 && K >= kTcChunkCodes && K <= 1024 && (K % kTcChunkCodes) == 0; }

struct StageLayout {
  size_t off_tab32, off_tab32T, off_cnorm, off_tc, off_meta, off_outl, off_gab, stride;
};
__host__ __device__ inline StageLayout stage_layout(int K, int D) {
  StageLayout L;
  size_t o = 0;
  L.off_tab32 = o;  o += align256(size_t(K) * D * 4);
  L.off_tab32T = o; o += align256(size_t(K) * D * 4);
  L.off_cnorm = o;  o += align256(size_t(K) * 4);
  L.off_tc = o;     o += tc_shape(K, D) ? size_t(K / kTcChunkCodes) * kTcChunkBytes : 0;
  L.off_meta = o;   o += 256;
  L.off_outl = o;   o += tc_shape(K, D) ? align256(size_t(K)) : 0;      // u8 [K]: codes excluded from the fp16 image
  L.off_gab = o;    o += tc_shape(K, D) ? align256(size_t(K) * 10) : 0; // float2 [K] {g16 + a, b}, then fp16 [K] g16
  L.stride = o;
  return L;
}

struct PackView {
  const unsigned char* base;
  StageLayout L;
  __host__ __device__ PackView(const void* p, int K, int D) : base((const unsigned char*)p), L(stage_layout(K, D)) {}
  __host__ __device__ const unsigned char* stage(int s) const { return base + kHeaderBytes + size_t(s) * L.stride; }
  __host__ __device__ const float* tab32(int s)  const { return (const float*)(stage(s) + L.off_tab32); }
  __host__ __device__ const float* tab32T(int s) const { return (const float*)(stage(s) + L.off_tab32T); }
  __host__ __device__ const float* cnorm(int s)  const { return (const float*)(stage(s) + L.off_cnorm); }
  __host__ __device__ const unsigned char* tc(int s) const { return stage(s) + L.off_tc; }
  __host__ __device__ const StageMeta* meta(int s) const { return (const StageMeta*)(stage(s) + L.off_meta); }
  __host__ __device__ const unsigned char* outl(int s) const { return stage(s) + L.off_outl; }
  __host__ __device__ const float2* gab(int s) const { return (const float2*)(stage(s) + L.off_gab); }
  __host__ __device__ const __half* g16(int s, int K) const { return (const __half*)(stage(s) + L.off_gab + size_t(K) * 8); }
};

// ----------------------------------------------------------------------------------------------
// error plumbing
// ----------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);   // records + returns RVQ_ECUDA
int  check_device();                               // RVQ_OK or RVQ_ENODEV (cached per device)
void count_launch(int n = 1);

#define RVQ_CUDA(expr)                                             \
  do {                                                             \
    cudaError_t _e = (expr);                                       \
    if (_e != cudaSuccess) return rvq::cuda_fail(_e, #expr);       \
  } while (0)

#define RVQ_LAUNCH_CHECK(name)                                     \
  do {                                                             \
    rvq::count_launch();                                           \
    cudaError_t _e = cudaGetLastError();                           \
    if (_e != cudaSuccess) return rvq::cuda_fail(_e, name);        \
  } while (0)

#define RVQ_REQUIRE(cond, ...)                                     \
  do {                                                             \
    if (!(cond)) { rvq::set_error(__VA_ARGS__); return RVQ_EINVAL; } \
  } while (0)

// frame n = b*T + t  ->  element offset of (b, d=0, t) in a strided [B, D, T] tensor
struct FrameAddr {
  int64_t sxb, sxd, sxt; int T;
  __device__ inline int64_t base(int64_t n) const { int64_t b = n / T; int64_t t = n - b * T; return b * sxb + t * sxt; }
};

// torch's CPU argmax (core_vq.py:188) propagates NaN: the first NaN distance wins; otherwise the smallest
// distance, lowest index on ties.  (best, bcode) starts as (+inf, 0x7fffffff).
__device__ __forceinline__ bool nan_aware_better(float dist, int code, float best, int bcode) {
  if (dist != dist) return best == best || code < bcode;
  return best == best && (dist < best || (dist == best && code < bcode));
}

// element index of the code of (stage s of the call, frame n) in the codes output: [n_q, B, T] by default,
// [B, n_q, T] with RVQ_FLAG_CODES_BKT (n < 2^31)
__device__ __forceinline__ int64_t code_index(int bkt, int n_q, int T, int64_t N, int s, int64_t n) {
  if (!bkt) return int64_t(s) * N + n;
  const uint32_t b = uint32_t(n) / uint32_t(T), t = uint32_t(n) - b * uint32_t(T);
  return (int64_t(b) * n_q + s) * T + t;
}

// optional search counters of the calling thread (rvq_search_counters); nullptr = off
unsigned long long* search_counters();
// rvq_pack_bound_mode: 0 = per stage (heterogeneous norms -> per-code bound), 1 = per-code everywhere, 2 = per-stage everywhere
int pack_bound_mode();

// ---- entry points implemented per translation unit -------------------------------------------
struct EncodeArgs {
  const void* pack; int K, D;
  const float* x; int64_t sxb, sxd, sxt; int B, T;
  int stage0, n_q;
  int64_t* codes; float* quantized; float* residual_out; double* sqerr;
  int flags;
  float* ema_counts = nullptr; float* ema_sum = nullptr;   // rvq_encode_train: EMA statistics accumulated by the search itself
};
int simt_pack(const float* const* embed_ptrs_host, int n_q, int K, int D, void* pack, cudaStream_t st);
int simt_encode(const EncodeArgs& a, cudaStream_t st);
int tc_encode(const EncodeArgs& a, cudaStream_t st);    // tcgen05 search, two tiles in flight per CTA (rvq_tc.cu)
int tc_debug_trace(long long* out_host, int n);         // RVQ_TC_TRACE builds: timeline of CTA 0

}  // namespace rvq

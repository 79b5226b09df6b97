/*
 * rvq_b200.h -- C ABI of the B200-native EnCodec residual vector quantizer.
 *
 * Drop-in boundary for the hot path of Madhudorai/encodec-pytorch
 * (quantization/core_vq.py + quantization/vq.py).  The Python host mirror in
 * encodec_pytorch_b200/quantization/ binds these with ctypes; INTEGRATION.md
 * shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer
 *     unless its name ends in _host;
 *   - the library never allocates or frees caller-visible memory: outputs and
 *     scratch are caller-allocated (the host mirror uses torch tensors);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *     every call is asynchronous on that stream and does no host sync;
 *   - return 0 on success, a negative RVQ_E* code on failure; the message is
 *     available from rvq_last_error() (thread-local);
 *   - frames are the B*T latent vectors of a [B, D, T] tensor, flattened
 *     b-major / t-minor exactly like core_vq.py:290 + :178;
 *   - there is NO CPU fallback: on a machine without an sm_100 device every
 *     compute entry point fails with RVQ_ENODEV.
 */
#ifndef RVQ_B200_H
#define RVQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVQ_ABI_VERSION 1

#define RVQ_OK        0
#define RVQ_EINVAL   -1   /* bad argument / unsupported shape              */
#define RVQ_ENODEV   -2   /* no sm_100 CUDA device                         */
#define RVQ_ECUDA    -3   /* CUDA runtime error (message has the detail)   */
#define RVQ_ESIZE    -4   /* caller buffer too small                       */

/* rvq_encode / rvq_ema_stats / rvq_residual_combine / rvq_expire_codes flags */
#define RVQ_FLAG_STE          1  /* training arithmetic of core_vq.py:309/:348: q <- r + (q - r) */
#define RVQ_FLAG_FORCE_EXACT  2  /* use the fp32 SIMT search even where the tcgen05 path applies   */
#define RVQ_FLAG_DIRECT_DIST  4  /* k-means distance sum((x-c)^2) of core_vq.py:86-91 (exact path)  */
#define RVQ_FLAG_ACCUM_Q      8  /* `quantized` holds a running sum on entry (stage segments)       */
#define RVQ_FLAG_CODES_BKT   16  /* rvq_encode: codes are written as [B, n_q, T] (what model.py:166 makes of the search's
                                     [n_q, B, T] with a transpose) instead of [n_q, B, T]                             */
#define RVQ_FLAG_OUT_BDT     32  /* rvq_encode / rvq_decode_ex: the fp32 frame output (`quantized` / `out`) is written as a
                                     contiguous [B, D, T] tensor -- the layout SEANet's decoder consumes
                                     (modules/seanet.py:193-195) -- instead of [B, T, D]                              */

int         rvq_version(void);                 /* RVQ_ABI_VERSION                                    */
const char* rvq_last_error(void);              /* last error message of the calling thread           */
int         rvq_device_ok(void);               /* 0 if the current device is sm_100, else RVQ_ENODEV */
/* number of kernels launched by this library in this process (bench.py's gpu_launches) */
uint64_t    rvq_launch_count(void);

/* ---- codebook pack ---------------------------------------------------------------------------
 * The search image of n_q codebooks [K, D] fp32 (EuclideanCodebook.embed, core_vq.py:143):
 * fp32 copy [n_q,K,D], transposed fp32 copy [n_q,D,K], |c|^2 [n_q,K], and (D==128, K%128==0)
 * the fp16 UMMA operand image + per-stage margin metadata used by the tcgen05 search.
 * embed_ptrs_host: HOST array of n_q DEVICE pointers (the per-stage `embed` buffers).          */
size_t rvq_pack_bytes(int n_q, int K, int D);
int    rvq_pack(const float* const* embed_ptrs_host, int n_q, int K, int D,
                void* pack, size_t pack_bytes, void* stream);

/* ---- encode: replaces ResidualVectorQuantization.encode (core_vq.py:357-367) and the search /
 * gather / residual part of .forward (core_vq.py:337-355 with :212-221, :301-324).
 *   x            fp32 [B, D, T] with ELEMENT strides (sxb, sxd, sxt)
 *   stage0,n_q   stages [stage0, stage0+n_q) of the pack are applied in order
 *   codes        int64 [n_q, B, T] contiguous (out); [B, n_q, T] with RVQ_FLAG_CODES_BKT
 *   quantized    fp32 [B, T, D] contiguous (out; [B, D, T] with RVQ_FLAG_OUT_BDT) or NULL: sum over stages, in stage order, of the
 *                gathered codewords (of the straight-through values with RVQ_FLAG_STE)
 *                (RVQ_FLAG_ACCUM_Q: added onto the values already in the buffer)
 *   residual_out fp32 [B, T, D] contiguous (out) or NULL: the residual after the last stage
 *   stage_sqerr  double [n_q] (out, accumulated by atomics; caller zeroes) or NULL:
 *                sum over frames and dims of (q_i - r_i)^2 -> commitment loss numerator (:319)   */
int rvq_encode(const void* pack, int K, int D,
               const float* x, int64_t sxb, int64_t sxd, int64_t sxt, int B, int T,
               int stage0, int n_q,
               int64_t* codes, float* quantized, float* residual_out, double* stage_sqerr,
               int flags, void* stream);

/* ---- training forward in one launch: rvq_encode plus the EMA statistics of core_vq.py:227-228 -- what rvq_ema_stats
 * computes from the codes afterwards, accumulated by the search itself while each stage's input residual is on chip
 * (counts fp32 [n_q, K] = embed_onehot.sum(0), embed_sum fp32 [n_q, K, D] = (x^T @ onehot)^T; both zeroed by the call;
 * float atomics: sums agree with rvq_ema_stats to summation order).  Other arguments as rvq_encode.           */
int rvq_encode_train(const void* pack, int K, int D,
                     const float* x, int64_t sxb, int64_t sxd, int64_t sxt, int B, int T,
                     int stage0, int n_q,
                     int64_t* codes, float* quantized, float* residual_out, double* stage_sqerr,
                     float* counts, float* embed_sum, int flags, void* stream);

/* ---- decode: replaces ResidualVectorQuantization.decode (core_vq.py:369-375).
 *   codes  int64 [n_q, B, T] with ELEMENT strides (scq, scb, sct) (model.py:188 passes a transposed view)
 *   out    fp32 [B, T, D] contiguous: ((0 + e_0[c_0]) + e_1[c_1]) + ...                          */
int rvq_decode(const void* pack, int K, int D,
               const int64_t* codes, int64_t scq, int64_t scb, int64_t sct,
               int n_q, int B, int T, float* out, void* stream);
/* same with flags: RVQ_FLAG_OUT_BDT writes out as contiguous fp32 [B, D, T]                          */
int rvq_decode_ex(const void* pack, int K, int D,
                  const int64_t* codes, int64_t scq, int64_t scb, int64_t sct,
                  int n_q, int B, int T, float* out, int flags, void* stream);

/* ---- EMA statistics: bincount + scatter-add of core_vq.py:227-228 for all stages at once.
 * Re-derives each stage's input residual from x and codes with the encode arithmetic.
 *   counts     fp32 [n_q, K]     (out; zeroed by the call)  = embed_onehot.sum(0)
 *   embed_sum  fp32 [n_q, K, D]  (out; zeroed by the call)  = (x^T @ onehot)^T
 * The caller all-reduces both across ranks (distrib.py:32-34) before rvq_ema_apply.              */
int rvq_ema_stats(const void* pack, int K, int D,
                  const float* x, int64_t sxb, int64_t sxd, int64_t sxt, int B, int T,
                  int stage0, int n_q, const int64_t* codes,
                  float* counts, float* embed_sum, int flags, void* stream);

/* ---- EMA apply: core_vq.py:227-235 + :49-60 for n_q stages (in-place on the module buffers).
 * *_ptrs_host: HOST arrays of n_q DEVICE pointers (cluster_size [K], embed_avg [K,D], embed [K,D]). */
int rvq_ema_apply(float* const* cluster_size_ptrs_host, float* const* embed_avg_ptrs_host,
                  float* const* embed_ptrs_host, int n_q, int K, int D,
                  const float* counts, const float* embed_sum,
                  double decay, double epsilon, void* stream);

/* ---- dead-code expiry: core_vq.py:159-163 (replace_): embed[k] <- samples[k] where
 * cluster_size[k] < threshold.  samples fp32 [K, D] (rows already drawn by the host RNG).        */
int rvq_expire_replace(float* embed, const float* cluster_size, const float* samples,
                       int K, int D, float threshold, void* stream);

/* ---- fused expiry for stage `stage` (relative to stage0) of a residual stack (core_vq.py:165-175):
 * embed[k] <- r_stage[sel[k]] where cluster_size[k] < threshold; r_stage is the input residual of
 * that stage, recomputed from x and codes for the K selected frames only.
 *   sel  int64 [K] DEVICE: frame numbers in [0, B*T) drawn by the host RNG (sample_vectors, :69-77) */
int rvq_expire_codes(const void* pack, int K, int D,
                     const float* x, int64_t sxb, int64_t sxd, int64_t sxt, int B, int T,
                     int stage0, int stage, const int64_t* codes, const int64_t* sel,
                     const float* cluster_size, float threshold, float* embed,
                     int flags, void* stream);

/* ---- expiry of a whole residual stack with no host round trip (core_vq.py:165-175 applied to n_q stages; replaces
 * the per-stage `torch.any` host sync + `randperm` + rvq_expire_codes sequence).  Per stage i:
 *   fired[i] = any(cluster_size_i < threshold)                       (int32 [n_q] DEVICE, out)
 *   sel[i,:] = K distinct frame numbers in [0, B*T), uniformly random and in random order -- the law of
 *              randperm(N)[:K] (sample_vectors, :69-77); K draws with replacement when B*T < K (:75).  Drawn on the
 *              device from the counter-based stream (seed, offset); the same pair reproduces the same indices.
 *              (int64 [n_q, K] DEVICE, out; rows of stages that did not fire are left untouched)
 *   embed_i[k] <- r_i[sel[i,k]] where cluster_size_i[k] < threshold  (r_i: input residual of stage i, as in
 *              rvq_expire_codes).  Stages index relative to stage0; *_ptrs_host are HOST arrays of DEVICE pointers.
 * K <= 2048.                                                                                        */
int rvq_expire_stack(const void* pack, int K, int D,
                     const float* x, int64_t sxb, int64_t sxd, int64_t sxt, int B, int T,
                     int stage0, int n_q, const int64_t* codes,
                     const float* const* cluster_size_ptrs_host, float* const* embed_ptrs_host, float threshold,
                     uint64_t seed, uint64_t offset, int64_t* sel, int* fired, int flags, void* stream);

/* ---- code bit-packing (the callers' side of the path: binary.BitPacker driven by compress.compress_to_file,
 * binary.py:69-87 + compress.py:70-92, and binary.BitUnpacker, binary.py:104-121), one byte stream per batch item.
 * Value number i = t*n_q + k (time-major, codebook-minor) occupies stream bits [i*bits, (i+1)*bits), little-endian in
 * bits; a trailing partial byte is zero-padded.  codes: int64, element strides (sq, sb, st) for index [k, b, t] (the search's
 * [n_q, B, T] output or the model's [B, K, T] view alike); values are masked to `bits` bits.  out / in: uint8, stream b at
 * byte offset b * stride, rvq_bitpack_bytes(n_q, T, bits) bytes each.  1 <= bits <= 16, n_q <= 64 (pack).             */
size_t rvq_bitpack_bytes(int n_q, int T, int bits);
int rvq_bitpack(const int64_t* codes, int64_t sq, int64_t sb, int64_t st, int n_q, int B, int T, int bits,
                unsigned char* out, int64_t out_stride, void* stream);
int rvq_bitunpack(const unsigned char* in, int64_t in_stride, int n_q, int B, int T, int bits,
                  int64_t* codes, int64_t sq, int64_t sb, int64_t st, void* stream);

/* ---- k-means (core_vq.py:80-102) on flat fp32 samples [N, D] (contiguous).
 * assign: buckets[n] = argmin_k sum_d (x[n,d]-means[k,d])^2, lowest index on ties (:86-91);
 *         `pack` is an rvq_pack() image (n_q = 1) of the current means.
 * update: bins = bincount(buckets); means[k] <- mean of assigned rows, empty clusters keep the
 *         old mean (:92-100).  sums fp32 [K, D] and bins int64 [K] are caller scratch/outputs.    */
int rvq_kmeans_assign(const void* pack, int K, int D, const float* samples, int64_t N,
                      int64_t* buckets, void* stream);
int rvq_kmeans_update(const float* samples, int64_t N, int D, const int64_t* buckets, int K,
                      float* means, int64_t* bins, float* sums, void* stream);

/* ---- backward helper: out[B,T,D] = sum_i w[i] * r_{i+1}, r_{i+1} the residual after stage i,
 * recomputed from x and codes (gradient of the commitment losses).
 *   w  fp32 [n_q] DEVICE                                                                         */
int rvq_residual_combine(const void* pack, int K, int D,
                         const float* x, int64_t sxb, int64_t sxd, int64_t sxt, int B, int T,
                         int stage0, int n_q, const int64_t* codes, const float* w,
                         float* out, int flags, void* stream);

/* ---- debug / evidence: optional counters of the tcgen05 search.  Registers, for the CALLING THREAD, a device buffer of
 * 32 x uint64 that every later rvq_encode of this thread accumulates into (atomics on the encode's stream):
 *   [0] frame-stages searched, [1] certified unique, [2] re-scored (candidate list / wide set), [3] exact scans,
 *   [11] re-scored frame-stages with more than 4 candidates, [12] the total of their candidates.
 * NULL (the default) turns the counters off; the codebook pack is never written after rvq_pack.  The caller zeroes
 * and reads the buffer.                                                                               */
int rvq_search_counters(uint64_t* counters_dev);

/* ---- which score-error bound the tcgen05 search certifies winners with (StageMeta in the pack; core_vq.py:181-189 is the
 * arithmetic being certified).  Both bounds are rigorous -- the choice changes how many frames take the exact fp32
 * re-score, never a result:
 *   0 (default) per stage at rvq_pack time: the per-code bound where the code norms are heterogeneous (fitted tables),
 *               the per-stage bound where they are uniform (freshly initialised tables);
 *   1 per-code bound on every stage;   2 per-stage bound on every stage.
 * Process-wide; read by rvq_pack.  Returns the previous mode.                                             */
int rvq_pack_bound_mode(int mode);

#ifdef __cplusplus
}
#endif
#endif /* RVQ_B200_H */

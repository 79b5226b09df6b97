// fp32 SIMT kernels of the B200 RVQ library: codebook pack, exact (fp32) fused multi-stage
// search, gather/decode, EMA statistics + update, dead-code replacement, k-means update and the
// residual-combine backward helper.  The tcgen05 search lives in rvq_tc.cu.
//
// Reference behaviour restated here (paths relative to the reference repository):
//   quantization/core_vq.py:181-189  distance / argmax (ties -> lowest index)
//   quantization/core_vq.py:357-367  residual encode loop
//   quantization/core_vq.py:369-375  decode sum order
//   quantization/core_vq.py:227-235  EMA statistics, Laplace smoothing, table overwrite
//   quantization/core_vq.py:80-102   k-means
#include "rvq_common.cuh"
#include <stdlib.h>

namespace rvq {

// ================================================================================================
// pack
// ================================================================================================
struct PtrTable32 { const float* p[32]; };
struct MutPtrTable32 { float* p[32]; };

// copy embed -> tab32 and tab32T through a 32x32 shared tile (coalesced both ways)
__global__ void pack_copy_kernel(PtrTable32 src, unsigned char* pack, int stage_base, int K, int D) {
  __shared__ float tile[32][33];
  const int s = blockIdx.z;
  PackView pv(pack, K, D);
  const float* in = src.p[s];
  float* t32 = const_cast<float*>(pv.tab32(stage_base + s));
  float* t32T = const_cast<float*>(pv.tab32T(stage_base + s));
  const int k0 = blockIdx.y * 32, d0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int k = k0 + i, d = d0 + threadIdx.x;
    float v = 0.f;
    if (k < K && d < D) { v = in[size_t(k) * D + d]; t32[size_t(k) * D + d] = v; }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int d = d0 + i, k = k0 + threadIdx.x;
    if (k < K && d < D) t32T[size_t(d) * K + k] = tile[threadIdx.x][i];
  }
}

// one block per stage: |c|^2, norm statistics, margin metadata and the fp16 UMMA image
// |score error| <= beta*|x|*|c|: fp16 rounding of x and of -2c (2 * 2^-11 * 2|x||c| by Cauchy-Schwarz)
// plus slack (x1.125) for the tensor core's fp32 accumulation of 144 products
constexpr float kMarginSlack = 1.0625f;              // fp32 accumulation of the 144 products, fp32 norm arithmetic
constexpr float kHalfUlp     = 4.8828125e-4f;        // 2^-11: relative rounding error bound of fp16
constexpr float kOutlierMul = 64.f;                  // codes with |c| > 64 * (15/16-quantile of the norms) are outliers
constexpr float kBigScore   = 60000.f;               // fp16-representable score of an outlier code
constexpr float kHalfSafe   = 3.0e4f;                // |2c| elements and |c|^2 must stay below fp16 max

__global__ void __launch_bounds__(1024) pack_meta_kernel(unsigned char* pack, int stage_base, int K, int D, float margin_scale, int bound_mode) {
  extern __shared__ float sh[];          // [Kpow2] sorted norms, then [K] flags
  const int s = stage_base + blockIdx.x;
  PackView pv(pack, K, D);
  const float* t32 = pv.tab32(s);
  const float* tT = pv.tab32T(s);
  float* cn = const_cast<float*>(pv.cnorm(s));
  const bool tc = tc_shape(K, D);
  int Kp = 1; while (Kp < K) Kp <<= 1;
  float* norms = sh;                      // Kp
  // (the three arrays below exist only for shapes the tensor-core search serves: simt_pack sizes the block's memory)
  unsigned long long* hashes = (unsigned long long*)(sh + Kp);   // K: 64-bit hash of every row (duplicate detection)
  unsigned long long* tab_h = hashes + K;                        // 2 Kp: open-addressing table of the hashes ...
  int* tab_i = (int*)(tab_h + 2 * Kp);                           // 2 Kp: ... and the lowest row index that carries each
  float* e2s = (float*)(tab_i + 2 * Kp);                         // K: |(-2c) - fp16(-2c)|^2 of every row
  unsigned char* outl = tc ? (unsigned char*)(e2s + K) : (unsigned char*)(sh + Kp);   // K

  for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
    float nv = __int_as_float(0x7f800000);
    if (k < K) {
      // element (k, d) read from the transposed copy: consecutive threads touch consecutive addresses; the sum runs over d
      // in the same order as before
      const float* col = tT + k;
      float acc = 0.f, amax = 0.f, e2 = 0.f;
      unsigned h1 = 0x811c9dc5u, h2 = 0x9747b28cu;      // two 32-bit multiplicative hashes of the row's bit patterns
      #pragma unroll 8
      for (int d = 0; d < D; ++d) {
        float v = col[size_t(d) * K]; acc = fmaf(v, v, acc); amax = fmaxf(amax, fabsf(v));
        const unsigned bits = v == 0.f ? 0u : __float_as_uint(v);                   // (-0 and +0 are the same row element)
        h1 = h1 * 31u + bits;
        h2 = (h2 ^ bits) * 0x9e3779b1u;
        // exact rounding residue of this code's fp16 operand row (-2c): |b - fp16(b)|^2
        const float bb = -2.f * v; const float e = bb - __half2float(__float2half_rn(bb)); e2 = fmaf(e, e, e2);
      }
      if (tc) { hashes[k] = ((static_cast<unsigned long long>(h1) << 32) | h2) | 1ull; e2s[k] = e2; }      // (hash 0 marks an empty table slot)
      cn[k] = acc;
      nv = sqrtf(acc);
      // range flags for the fp16 image: B holds -2c, the augmented column holds |c|^2
      outl[k] = (!(2.f * amax < kHalfSafe) || !(acc < kHalfSafe)) ? 1 : 0;
    }
    norms[k] = nv;
  }
  if (tc)
    for (int i = threadIdx.x; i < 2 * Kp; i += blockDim.x) { tab_h[i] = 0ull; tab_i[i] = 0x7fffffff; }
  __syncthreads();
  if (!tc) return;
  // lowest row index per distinct hash: insert with linear probing (at most K of the 2 Kp slots fill)
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const unsigned long long hk = hashes[k];
    unsigned slot = unsigned(hk >> 20) & unsigned(2 * Kp - 1);
    while (true) {
      const unsigned long long prev = atomicCAS(&tab_h[slot], 0ull, hk);
      if (prev == 0ull || prev == hk) { atomicMin(&tab_i[slot], k); break; }
      slot = (slot + 1) & unsigned(2 * Kp - 1);
    }
  }
  __syncthreads();

  // bitonic sort of the norms (ascending); +inf padding sinks to the end
  for (int size = 2; size <= Kp; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < Kp; i += blockDim.x) {
        int j = i ^ stride;
        if (j > i) {
          bool up = ((i & size) == 0);
          float a = norms[i], b = norms[j];
          if ((a > b) == up) { norms[i] = b; norms[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  __shared__ float s_thr, s_cref, s_cmin, s_outmin, s_cmax, s_db2, s_gmax, s_bmax, s_nlow;
  __shared__ int s_nout, s_nalias;
  if (threadIdx.x == 0) {
    // reference norm for the outlier test: a high quantile (15/16), not the median -- a codebook whose norms are bimodal
    // with the large group in the minority (e.g. the first EMA steps after a k-means init, where most rows have shrunk) must
    // not have its large, winning codes classified as outliers (every frame would then take the exact scan)
    const float ref = norms[K - 1 - K / 16];
    s_thr = kOutlierMul * ref;
    s_cmin = norms[0];
    s_cmax = norms[K - 1];
    s_cref = 0.f; s_outmin = __int_as_float(0x7f800000); s_nout = 0; s_nalias = 0; s_db2 = 0.f; s_gmax = 0.f; s_bmax = 0.f;
    s_nlow = norms[K / 64];                // a low quantile of the norms (the K/64 smallest codes are at or below it)
  }
  __syncthreads();
  // classify (a code is an outlier by norm ratio or by fp16 range); reduce cref / min outlier norm
  {
    float lcref = 0.f, loutmin = __int_as_float(0x7f800000), ldb2 = 0.f, lgmax = 0.f, lbmax = 0.f; int lnout = 0;
    float2* gab = const_cast<float2*>(pv.gab(s));
    __half* g16 = const_cast<__half*>(pv.g16(s, K));
    int lnalias = 0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      float nv = sqrtf(cn[k]);
      bool o = outl[k] || !(nv <= s_thr);
      // An exact duplicate of a row with a lower index can never win (equal distances, ties go to the lowest index,
      // core_vq.py:188): it leaves the image like an outlier does, so that a crowd of identical rows (zero residuals drawn as
      // k-means means, silence) does not put the whole crowd into every candidate set.  (The exact scans still see it.)
      bool alias = false;
      if (!o) {
        // first row with the same hash (table look-up), then one element-wise check
        const unsigned long long hk = hashes[k];
        unsigned slot = unsigned(hk >> 20) & unsigned(2 * Kp - 1);
        while (tab_h[slot] != hk) slot = (slot + 1) & unsigned(2 * Kp - 1);
        const int first = tab_i[slot];
        if (first < k) {
          bool same = true;
          for (int d = 0; d < D && same; ++d) same = tT[size_t(d) * K + first] == tT[size_t(d) * K + k];
          alias = same;
        }
      }
      outl[k] = (o || alias) ? 1 : 0;
      float2 ab = make_float2(0.f, 0.f);
      __half gh = __float2half_rn(0.f);
      if (o) { loutmin = fminf(loutmin, nv); ++lnout; }
      else if (alias) ++lnalias;
      else {
        lcref = fmaxf(lcref, nv);
        const float e2 = e2s[k];      // rounding residue of the fp16 operand row, from the first pass
        ldb2 = fmaxf(ldb2, e2);
        // per-code coefficients (StageMeta): |S_k - s_k| <= a_k |r| + b_k |r - fp16(r)|; the image carries
        // g16_k >= a_k + 2^-11 b_k (rounded UP to fp16, so that S_k - g16_k R stays a lower bound of s_k)
        const float nvu = nv * 1.0001f;
        const float dbk = sqrtf(e2) * 1.0001f;
        const float ak = margin_scale * (kMarginSlack * dbk + 144.f * 1.1920929e-7f * 2.f * nvu);
        const float bk = margin_scale * kMarginSlack * (2.f * nvu + dbk);
        // (at least the smallest normal fp16: nothing then depends on how the tensor core treats subnormal operands)
        gh = __float2half_ru(fmaxf((ak + bk * kHalfUlp * 1.001f) * 1.00001f, 6.2e-5f));
        const float g = __half2float(gh);
        ab = make_float2((g + ak) * 1.00001f, bk);
        lgmax = fmaxf(lgmax, g); lbmax = fmaxf(lbmax, bk);
      }
      gab[k] = ab; g16[k] = gh;
    }
    atomicMax((int*)&s_cref, __float_as_int(lcref));          // non-negative floats order as ints
    atomicMax((int*)&s_db2, __float_as_int(ldb2));
    atomicMax((int*)&s_gmax, __float_as_int(lgmax));
    atomicMax((int*)&s_bmax, __float_as_int(lbmax));
    atomicMin((int*)&s_outmin, __float_as_int(loutmin));
    atomicAdd(&s_nout, lnout);
    atomicAdd(&s_nalias, lnalias);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    StageMeta m;
    m.cref = s_cref; m.cmin = s_cmin; m.n_outliers = s_nout;
    m.cmax_all = s_cmax;
    // score error of code k for a frame r with fp16 image r~ = r - dr, operand row b_k = fp16(-2 c_k) = -2 c_k - db_k:
    //   sum_d (r~_d b_kd + 2 r_d c_kd) = -r.db_k + 2 c_k.dr + dr.db_k   =>   |.| <= |r| dbmax + (2 cref + dbmax) |dr|
    const float dbmax = sqrtf(s_db2) * 1.0001f;
    // + the tensor core's fp32 accumulation of the 144 products of a score (|sum| <= |r~| |b_k| <= 2 |r| cref): an absolute
    // term, so that operands which happen to be exact in fp16 (dbmax = 0, |dr| = 0) are still covered
    m.margin_coef = margin_scale * (2.f * kMarginSlack * dbmax + 144.f * 1.1920929e-7f * 2.f * s_cref);
    m.margin_dr = margin_scale * 2.f * kMarginSlack * (2.f * s_cref + dbmax);
    // |c|^2 is carried as fp16 hi + fp16 lo: error <= 2^-22 |c|^2 (+ 2^-24 when lo is subnormal)
    m.margin_abs = 2.f * (2.4e-7f * s_cref * s_cref + 6e-8f);
    // per-code bound: worth its extra look-up when a group of codes is well below the largest one, which sets the per-stage
    // bound (fitted tables: the codes that compete for typical frames are the small, populous ones -- measured on B200 at
    // cfg2: 2.27 -> 0.53 ms on a fitted stack, 0.52 -> 0.56 ms on uniform norms).  + the part of |r - fp16(r)| that fp16
    // subnormals add to the relative bound 2^-11 |r| (sqrt(128) 2^-25) times the largest b_k.
    m.percode = bound_mode == 1 ? 1 : bound_mode == 2 ? 0 : (s_cref > 1.5f * s_nlow ? 1 : 0);
    m.abs_pc = m.margin_abs + s_bmax * 3.5e-7f + s_gmax * 6.2e-5f;      // (last term: a frame bound R below the normal fp16 range)
    m.g16max = s_gmax;
    // |x| bound under which (a) outlier codes provably lose to the smallest-norm code and
    // (b) every live score + margin stays below the outlier score, (c) x fits fp16.
    float xl = 6.0e4f;
    if (s_nout + s_nalias > 0) {     // (no outlier by norm: s_outmin = +inf and the first clause is void)
      xl = fminf(xl, 0.5f * (s_outmin - s_cmin));
      float denom = 2.f * s_cmin + (m.margin_coef + m.margin_dr * kHalfUlp * 1.01f) + 1e-30f;   // |dr| <= 2^-11 |r| (+ subnormals)
      xl = fminf(xl, (0.9f * kBigScore - s_cmin * s_cmin) / denom);
    }
    if (s_nout >= K) xl = 0.f;
    m.xlimit = xl > 0.f ? xl : 0.f;
    *const_cast<StageMeta*>(pv.meta(s)) = m;
  }
  // the flags go to the pack: pack_image_kernel (many blocks per stage) builds the fp16 image from them
  unsigned char* og = const_cast<unsigned char*>(pv.outl(s));
  for (int k = threadIdx.x; k < K; k += blockDim.x) og[k] = outl[k];
}

// fp16 UMMA image of a stage (layout in rvq_common.cuh): chunk c = 128 codes x 18 K-groups of 8 halves.  grid (stages, 18):
// a block writes one K-group of every code (consecutive threads -> consecutive codes -> contiguous 16 B).
__global__ void __launch_bounds__(256) pack_image_kernel(unsigned char* pack, int stage_base, int K, int D) {
  const int s = stage_base + blockIdx.x, g = blockIdx.y;
  PackView pv(pack, K, D);
  const float* t32 = pv.tab32(s);
  const float* cn = pv.cnorm(s);
  const unsigned char* outl = pv.outl(s);
  unsigned char* img = const_cast<unsigned char*>(pv.tc(s));
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const int c = k / kTcChunkCodes, r = k % kTcChunkCodes;
    const bool o = outl[k] != 0;
    __align__(16) __half h[8];
    if (g < 16) {
      const float4* row = reinterpret_cast<const float4*>(t32 + size_t(k) * D + g * 8);
      const float4 a = row[0], b = row[1];
      const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      #pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = __float2half_rn(o ? 0.f : -2.f * v[j]);
    } else if (g == 16) {
      float ee = o ? kBigScore : cn[k];
      __half hi = __float2half_rn(ee);
      __half lo = __float2half_rn(o ? 0.f : ee - __half2float(hi));
      h[0] = hi; h[1] = lo;
      #pragma unroll
      for (int j = 2; j < 8; ++j) h[j] = __float2half_rn(0.f);
      // columns 2 and 4: -g16_k, met by the frame's bound(s) of |r| in the operand's augmented block (rvq_tc.cu)
      h[2] = h[4] = __hneg(pv.g16(s, K)[k]);
    } else {
      #pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = __float2half_rn(0.f);
    }
    *reinterpret_cast<uint4*>(img + size_t(c) * kTcChunkBytes + size_t(g) * kTcLBO + size_t(r) * 16) =
        *reinterpret_cast<const uint4*>(h);
  }
}

int simt_pack(const float* const* embed_ptrs_host, int n_q, int K, int D, void* pack, cudaStream_t st) {
  int Kp = 1; while (Kp < K) Kp <<= 1;
  size_t meta_smem = size_t(Kp) * 4 + size_t(K) + (tc_shape(K, D) ? size_t(K) * 12 + size_t(Kp) * 24 : 0);
  RVQ_REQUIRE(meta_smem <= 200 * 1024, "rvq_pack: codebook_size %d too large", K);
  RVQ_CUDA(cudaFuncSetAttribute(pack_meta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)meta_smem));
  RVQ_CUDA(cudaMemsetAsync(pack, 0, kHeaderBytes, st));
  const float margin_scale = 1.f;
  for (int s0 = 0; s0 < n_q; s0 += 32) {
    int ns = n_q - s0 < 32 ? n_q - s0 : 32;
    PtrTable32 tab;
    for (int i = 0; i < 32; ++i) tab.p[i] = i < ns ? embed_ptrs_host[s0 + i] : nullptr;
    dim3 grid((D + 31) / 32, (K + 31) / 32, ns), block(32, 8);
    pack_copy_kernel<<<grid, block, 0, st>>>(tab, (unsigned char*)pack, s0, K, D);
    RVQ_LAUNCH_CHECK("pack_copy_kernel");
    pack_meta_kernel<<<ns, 1024, meta_smem, st>>>((unsigned char*)pack, s0, K, D, margin_scale, pack_bound_mode());
    RVQ_LAUNCH_CHECK("pack_meta_kernel");
    if (tc_shape(K, D)) {
      pack_image_kernel<<<dim3(ns, kTcKPad / 8), 256, 0, st>>>((unsigned char*)pack, s0, K, D);
      RVQ_LAUNCH_CHECK("pack_image_kernel");
    }
  }
  return RVQ_OK;
}

// ================================================================================================
// exact fused multi-stage search (fp32 SIMT).  Block = 256 threads, tile = 64 frames.
// ================================================================================================
constexpr int kXF = 64;    // frames per block
constexpr int kXC = 64;    // codes per chunk

template <bool DIRECT>
__global__ void __launch_bounds__(256, 2)
exact_encode_kernel(const unsigned char* pack, int K, int D,
                    const float* __restrict__ x, FrameAddr fa, int64_t N,
                    int stage0, int n_q, int64_t* __restrict__ codes,
                    float* __restrict__ quantized, float* __restrict__ residual_out,
                    double* __restrict__ sqerr, int ste, int accum_q, int bkt, int bdt) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                    // [D][64] residual, frame-contiguous
  float* cs = xs + size_t(D) * kXF;    // [D][64] codebook chunk, code-contiguous
  float* xx = cs + size_t(D) * kXC;    // [64]
  float* xpart = xx + kXF;             // [4][64]
  int*   win = (int*)(xpart + 4 * kXF);// [64]
  float* red = (float*)(win + kXF);    // [8]
  float* qs = red + 8;                 // [D][64] running quantized sum (only when requested)

  PackView pv(pack, K, D);
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t n0 = int64_t(blockIdx.x) * kXF;
  const int uf = tid & 63, uq = tid >> 6;          // update-phase mapping: frame, dim quarter
  const int dq = D >> 2;                            // dims per quarter (D % 4 == 0)
  const int64_t un = n0 + uf;
  const bool uvalid = un < N;

  // load the latent tile (coalesced along t when sxt == 1)
  {
    const int64_t xb = uvalid ? fa.base(un) : 0;
    float part = 0.f;
    for (int d = uq * dq; d < (uq + 1) * dq; ++d) {
      float v = uvalid ? x[xb + int64_t(d) * fa.sxd] : 0.f;
      xs[d * kXF + uf] = v;
      part = fmaf(v, v, part);
    }
    xpart[uq * kXF + uf] = part;
  }
  // element of frame n, dim d of the fp32 frame output: [B, T, D], or [B, D, T] with RVQ_FLAG_OUT_BDT
  auto qidx = [&](int64_t n, int d) -> int64_t {
    if (!bdt) return n * D + d;
    const int64_t b = n / fa.T, t = n - b * fa.T;
    return (b * D + d) * fa.T + t;
  };
  if (quantized != nullptr) {
    if (accum_q) {
      for (int i = tid; i < kXF * D; i += 256) {
        int f, d;
        if (bdt) { d = i / kXF; f = i - d * kXF; } else { f = i / D; d = i - f * D; }
        qs[d * kXF + f] = (n0 + f < N) ? quantized[qidx(n0 + f, d)] : 0.f;
      }
    } else {
      for (int d = uq * dq; d < (uq + 1) * dq; ++d) qs[d * kXF + uf] = 0.f;
    }
  }
  __syncthreads();
  if (tid < kXF) xx[tid] = ((xpart[tid] + xpart[kXF + tid]) + xpart[2 * kXF + tid]) + xpart[3 * kXF + tid];

  for (int si = 0; si < n_q; ++si) {
    const int s = stage0 + si;
    const float* t32T = pv.tab32T(s);
    const float* t32 = pv.tab32(s);
    const float* cn = pv.cnorm(s);
    float bd[4]; int bi[4];
    #pragma unroll
    for (int i = 0; i < 4; ++i) { bd[i] = __int_as_float(0x7f800000); bi[i] = 0x7fffffff; }

    for (int c0 = 0; c0 < K; c0 += kXC) {
      __syncthreads();   // previous chunk consumed / residual + xx updated
      for (int i = tid; i < D * (kXC / 4); i += 256) {
        int d = i / (kXC / 4), j4 = (i - d * (kXC / 4)) * 4;
        float4 v;
        const float* src = t32T + size_t(d) * K + c0 + j4;
        if (c0 + j4 + 3 < K && ((K & 3) == 0)) v = *reinterpret_cast<const float4*>(src);
        else {
          v.x = c0 + j4 + 0 < K ? src[0] : 0.f; v.y = c0 + j4 + 1 < K ? src[1] : 0.f;
          v.z = c0 + j4 + 2 < K ? src[2] : 0.f; v.w = c0 + j4 + 3 < K ? src[3] : 0.f;
        }
        *reinterpret_cast<float4*>(cs + d * kXC + j4) = v;
      }
      __syncthreads();
      float acc[4][4];
      #pragma unroll
      for (int i = 0; i < 4; ++i)
        #pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      #pragma unroll 4
      for (int d = 0; d < D; ++d) {
        float4 xv = *reinterpret_cast<const float4*>(xs + d * kXF + ty * 4);
        float4 cv = *reinterpret_cast<const float4*>(cs + d * kXC + tx * 4);
        float xa[4] = {xv.x, xv.y, xv.z, xv.w}, ca[4] = {cv.x, cv.y, cv.z, cv.w};
        #pragma unroll
        for (int i = 0; i < 4; ++i)
          #pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (DIRECT) { float t = xa[i] - ca[j]; acc[i][j] = fmaf(t, t, acc[i][j]); }
            else acc[i][j] = fmaf(xa[i], ca[j], acc[i][j]);
          }
      }
      #pragma unroll
      for (int j = 0; j < 4; ++j) {
        int code = c0 + tx * 4 + j;
        if (code < K) {
          float cnj = DIRECT ? 0.f : cn[code];
          #pragma unroll
          for (int i = 0; i < 4; ++i) {
            // core_vq.py:183-187: (|x|^2 - 2 x.e) + |e|^2, maximised after negation
            float dist = DIRECT ? acc[i][j] : (xx[ty * 4 + i] - 2.f * acc[i][j]) + cnj;
            if (nan_aware_better(dist, code, bd[i], bi[i])) { bd[i] = dist; bi[i] = code; }
          }
        }
      }
    }
    // reduce over the 16 code lanes of each frame row; ties -> lowest index
    #pragma unroll
    for (int i = 0; i < 4; ++i) {
      #pragma unroll
      for (int off = 8; off > 0; off >>= 1) {
        float od = __shfl_xor_sync(0xffffffffu, bd[i], off);
        int oi = __shfl_xor_sync(0xffffffffu, bi[i], off);
        if (oi != 0x7fffffff && nan_aware_better(od, oi, bd[i], bi[i])) { bd[i] = od; bi[i] = oi; }
      }
      // no code at all (K == 0 cannot happen): code 0
      if (tx == 0) win[ty * 4 + i] = bi[i] == 0x7fffffff ? 0 : bi[i];
    }
    __syncthreads();
    // gather + residual update (+ straight-through arithmetic), exact fp32
    {
      const int idx = win[uf];
      const float* row = t32 + size_t(idx) * D;
      float part = 0.f;
      for (int d = uq * dq; d < (uq + 1) * dq; ++d) {
        float r = xs[d * kXF + uf];
        float q = row[d];
        if (ste) q = r + (q - r);
        float rn = r - q;
        xs[d * kXF + uf] = rn;
        part = fmaf(rn, rn, part);
        if (quantized != nullptr) qs[d * kXF + uf] += q;
      }
      xpart[uq * kXF + uf] = part;
      if (uvalid && uq == 0) codes[code_index(bkt, n_q, fa.T, N, si, un)] = idx;
      if (sqerr != nullptr) {
        float v = uvalid ? part : 0.f;
        #pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((tid & 31) == 0) red[tid >> 5] = v;
      }
    }
    __syncthreads();
    if (tid < kXF) xx[tid] = ((xpart[tid] + xpart[kXF + tid]) + xpart[2 * kXF + tid]) + xpart[3 * kXF + tid];
    if (sqerr != nullptr && tid == 0) {
      double t = 0.0;
      for (int i = 0; i < 8; ++i) t += (double)red[i];
      atomicAdd(&sqerr[si], t);
    }
  }
  __syncthreads();
  // frame-major stores: consecutive threads write consecutive dims of one frame
  if (quantized != nullptr) {
    for (int i = tid; i < kXF * D; i += 256) {
      int f, d;
      if (bdt) { d = i / kXF; f = i - d * kXF; } else { f = i / D; d = i - f * D; }      // consecutive threads -> consecutive addresses
      if (n0 + f < N) quantized[qidx(n0 + f, d)] = qs[d * kXF + f];
    }
  }
  if (residual_out != nullptr) {
    for (int i = tid; i < kXF * D; i += 256) {
      int f = i / D, d = i - f * D;
      if (n0 + f < N) residual_out[(n0 + f) * D + d] = xs[d * kXF + f];
    }
  }
}

int simt_encode(const EncodeArgs& a, cudaStream_t st) {
  const int K = a.K, D = a.D;
  RVQ_REQUIRE(D % 4 == 0 && D >= 4 && D <= 256, "rvq_encode: dimension %d unsupported (multiple of 4, <= 256)", D);
  RVQ_REQUIRE(K >= 1, "rvq_encode: codebook_size %d", K);
  const int64_t N = int64_t(a.B) * a.T;
  if (N == 0 || a.n_q == 0) return RVQ_OK;
  size_t smem = (size_t(D) * (kXF + kXC + (a.quantized ? kXF : 0)) + kXF + 4 * kXF + kXF + 8) * 4;
  FrameAddr fa{a.sxb, a.sxd, a.sxt, a.T};
  unsigned grid = unsigned((N + kXF - 1) / kXF);
  const bool direct = (a.flags & RVQ_FLAG_DIRECT_DIST) != 0;
  auto kern = direct ? exact_encode_kernel<true> : exact_encode_kernel<false>;
  RVQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, 256, smem, st>>>((const unsigned char*)a.pack, K, D, a.x, fa, N, a.stage0, a.n_q, a.codes,
                                a.quantized, a.residual_out, a.sqerr, (a.flags & RVQ_FLAG_STE) ? 1 : 0,
                                (a.flags & RVQ_FLAG_ACCUM_Q) ? 1 : 0, (a.flags & RVQ_FLAG_CODES_BKT) ? 1 : 0,
                                (a.flags & RVQ_FLAG_OUT_BDT) ? 1 : 0);
  RVQ_LAUNCH_CHECK("exact_encode_kernel");
  return RVQ_OK;
}

// ================================================================================================
// chain kernels: one warp per frame walks the stages (decode / EMA statistics / residual combine)
// ================================================================================================
enum ChainMode { kDecode = 0, kStats = 1, kCombine = 2, kQuantSum = 3 };

__device__ __forceinline__ void red_add_f4(float* addr, float4 v) {
#if __CUDA_ARCH__ >= 900
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
#else
  atomicAdd(addr, v.x); atomicAdd(addr + 1, v.y); atomicAdd(addr + 2, v.z); atomicAdd(addr + 3, v.w);
#endif
}

template <int MODE>
__global__ void __launch_bounds__(256)
chain_kernel(const unsigned char* pack, int K, int D,
             const float* __restrict__ x, FrameAddr fa, int64_t N, int stage0, int n_q,
             const int64_t* __restrict__ codes, int64_t scq, int64_t scb, int64_t sct, int T,
             const float* __restrict__ w, float* __restrict__ out,
             float* __restrict__ counts, float* __restrict__ embed_sum, int ste, int accum = 0) {
  PackView pv(pack, K, D);
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int nd4 = D >> 2;     // float4 groups per frame (D % 4 == 0)
  for (int64_t n = warp; n < N; n += nwarps) {
    const int64_t b = n / T, t = n - b * T;
    for (int g0 = 0; g0 < nd4; g0 += 32) {
      const int g = g0 + lane;
      const bool act = g < nd4;
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f), acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == kQuantSum && accum && act) acc = *reinterpret_cast<const float4*>(out + n * D + g * 4);
      if (MODE != kDecode && (MODE != kQuantSum || ste) && act) {
        const int64_t xb = fa.base(n) + int64_t(g) * 4 * fa.sxd;
        r.x = x[xb]; r.y = x[xb + fa.sxd]; r.z = x[xb + 2 * fa.sxd]; r.w = x[xb + 3 * fa.sxd];
      }
      for (int s0 = 0; s0 < n_q; s0 += 32) {
        // lane i fetches the code of stage s0+i once; broadcast per stage below
        int mycode = 0;
        if (s0 + lane < n_q) {
          int64_t c = codes[int64_t(s0 + lane) * scq + b * scb + t * sct];
          mycode = c < 0 ? 0 : (c >= K ? K - 1 : int(c));
        }
        const int ns = n_q - s0 < 32 ? n_q - s0 : 32;
        for (int i = 0; i < ns; ++i) {
          const int s = stage0 + s0 + i;
          const int idx = __shfl_sync(0xffffffffu, mycode, i);
          if (MODE == kStats && g0 == 0 && lane == 0) atomicAdd(&counts[size_t(s0 + i) * K + idx], 1.0f);
          if (!act) continue;
          float4 q = *reinterpret_cast<const float4*>(pv.tab32(s) + size_t(idx) * D + g * 4);
          if (MODE == kDecode) {
            acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
          } else if (MODE == kQuantSum) {
            // running sum of the per-stage outputs (core_vq.py:349), straight-through values in training
            if (ste) {
              q.x = r.x + (q.x - r.x); q.y = r.y + (q.y - r.y); q.z = r.z + (q.z - r.z); q.w = r.w + (q.w - r.w);
              r.x -= q.x; r.y -= q.y; r.z -= q.z; r.w -= q.w;
            }
            acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
          } else {
            if (MODE == kStats) red_add_f4(embed_sum + (size_t(s0 + i) * K + idx) * D + g * 4, r);
            if (ste) { q.x = r.x + (q.x - r.x); q.y = r.y + (q.y - r.y); q.z = r.z + (q.z - r.z); q.w = r.w + (q.w - r.w); }
            r.x -= q.x; r.y -= q.y; r.z -= q.z; r.w -= q.w;
            if (MODE == kCombine) {
              const float ws = w[s0 + i];
              acc.x = fmaf(ws, r.x, acc.x); acc.y = fmaf(ws, r.y, acc.y);
              acc.z = fmaf(ws, r.z, acc.z); acc.w = fmaf(ws, r.w, acc.w);
            }
          }
        }
      }
      if (MODE != kStats && act) *reinterpret_cast<float4*>(out + n * D + g * 4) = acc;
    }
  }
}

// decode / quantized sum with the output written as contiguous [B, D, T] (RVQ_FLAG_OUT_BDT): a block takes 32 consecutive
// frames (a warp walks the stages of four of them exactly like chain_kernel), parks the sums in a transposed shared-memory
// tile and writes 128-byte runs along t.
template <int MODE>
__global__ void __launch_bounds__(256)
chain_bdt_kernel(const unsigned char* pack, int K, int D,
                 const float* __restrict__ x, FrameAddr fa, int64_t N, int stage0, int n_q,
                 const int64_t* __restrict__ codes, int64_t scq, int64_t scb, int64_t sct, int T,
                 float* __restrict__ out, int ste, int accum) {
  extern __shared__ float tile[];            // [D][33]
  PackView pv(pack, K, D);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nd4 = D >> 2;
  for (int64_t n0 = int64_t(blockIdx.x) * 32; n0 < N; n0 += int64_t(gridDim.x) * 32) {
    for (int fl = warp; fl < 32; fl += 8) {
      const int64_t n = n0 + fl;
      if (n >= N) break;
      const int64_t b = n / T, t = n - b * T;
      for (int g0 = 0; g0 < nd4; g0 += 32) {
        const int g = g0 + lane;
        const bool act = g < nd4;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f), acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == kQuantSum && accum && act) {      // the running sum continues in stage order (same association as one call)
          const float* o = out + (b * D + g * 4) * T + t;
          acc = make_float4(o[0], o[T], o[2 * int64_t(T)], o[3 * int64_t(T)]);
        }
        if (MODE == kQuantSum && ste && act) {
          const int64_t xb = fa.base(n) + int64_t(g) * 4 * fa.sxd;
          r.x = x[xb]; r.y = x[xb + fa.sxd]; r.z = x[xb + 2 * fa.sxd]; r.w = x[xb + 3 * fa.sxd];
        }
        for (int s0 = 0; s0 < n_q; s0 += 32) {
          int mycode = 0;
          if (s0 + lane < n_q) {
            int64_t c = codes[int64_t(s0 + lane) * scq + b * scb + t * sct];
            mycode = c < 0 ? 0 : (c >= K ? K - 1 : int(c));
          }
          const int ns = n_q - s0 < 32 ? n_q - s0 : 32;
          for (int i = 0; i < ns; ++i) {
            const int idx = __shfl_sync(0xffffffffu, mycode, i);
            if (!act) continue;
            float4 q = *reinterpret_cast<const float4*>(pv.tab32(stage0 + s0 + i) + size_t(idx) * D + g * 4);
            if (MODE == kQuantSum && ste) {
              q.x = r.x + (q.x - r.x); q.y = r.y + (q.y - r.y); q.z = r.z + (q.z - r.z); q.w = r.w + (q.w - r.w);
              r.x -= q.x; r.y -= q.y; r.z -= q.z; r.w -= q.w;
            }
            acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
          }
        }
        if (act) {
          float* tp = tile + (g * 4) * 33 + fl;
          tp[0] = acc.x; tp[33] = acc.y; tp[66] = acc.z; tp[99] = acc.w;
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D * 32; i += 256) {
      const int d = i >> 5, fl = i & 31;
      const int64_t n = n0 + fl;
      if (n < N) {
        const int64_t b = n / T, t = n - b * T;
        out[(b * D + d) * T + t] = tile[d * 33 + fl];
      }
    }
    __syncthreads();
  }
}

static unsigned chain_grid(int64_t N) {
  int64_t blocks = (N + 7) / 8;              // 8 warps per block, one frame per warp per pass
  const int64_t cap = 148 * 16;              // persistent-ish: a few waves over 148 SMs
  if (blocks > cap) blocks = cap;
  return unsigned(blocks < 1 ? 1 : blocks);
}

// quantized [N, D] = (accum ? quantized : 0) + sum over stages of the gathered rows (or of the
// straight-through values), in stage order; companion launch of the tensor-core search
static unsigned chain_bdt_grid(int64_t N) {
  int64_t blocks = (N + 31) / 32;
  const int64_t cap = 148 * 8;
  if (blocks > cap) blocks = cap;
  return unsigned(blocks < 1 ? 1 : blocks);
}

int simt_quant_sum(const void* pack, int K, int D, const float* x, FrameAddr fa, int64_t N, int T, int stage0, int n_q,
                   const int64_t* codes, float* out, int flags, cudaStream_t st) {
  const int ste = (flags & RVQ_FLAG_STE) ? 1 : 0, accum = (flags & RVQ_FLAG_ACCUM_Q) ? 1 : 0;
  // codes as the search wrote them: [n_q, B, T] or, with RVQ_FLAG_CODES_BKT, [B, n_q, T]
  const bool bkt = (flags & RVQ_FLAG_CODES_BKT) != 0;
  const int64_t scq = bkt ? int64_t(T) : N, scb = bkt ? int64_t(n_q) * T : int64_t(T);
  if (flags & RVQ_FLAG_OUT_BDT) {
    chain_bdt_kernel<kQuantSum><<<chain_bdt_grid(N), 256, size_t(D) * 33 * 4, st>>>((const unsigned char*)pack, K, D, x, fa, N, stage0,
                                                                                 n_q, codes, scq, scb, 1, T, out, ste, accum);
    RVQ_LAUNCH_CHECK("chain_bdt_kernel<quant_sum>");
    return RVQ_OK;
  }
  chain_kernel<kQuantSum><<<chain_grid(N), 256, 0, st>>>((const unsigned char*)pack, K, D, x, fa, N, stage0, n_q, codes,
                                                        scq, scb, 1, T, nullptr, out, nullptr, nullptr, ste, accum);
  RVQ_LAUNCH_CHECK("chain_kernel<quant_sum>");
  return RVQ_OK;
}

}  // namespace rvq

using namespace rvq;

// ------------------------------------------------------------------------------------------------
// EMA apply / expiry / k-means update kernels
// ------------------------------------------------------------------------------------------------
// cluster_size <- decay*cluster_size + (1-decay)*bincount   (core_vq.py:227, :49-56); one block per stage
__global__ void __launch_bounds__(1024)
ema_cluster_kernel(MutPtrTable32 cs_tab, int stage_base, int K, const float* __restrict__ counts, float decay, float alpha) {
  float* cs = cs_tab.p[blockIdx.x];
  const float* cnt = counts + size_t(stage_base + blockIdx.x) * K;
  for (int k = threadIdx.x; k < K; k += blockDim.x) cs[k] = cs[k] * decay + alpha * cnt[k];
}

// embed_avg <- EMA; embed <- embed_avg / (laplace(cluster_size) * total)   (core_vq.py:229-235).  grid (stages, kEmaSplit):
// every block re-derives the stage's total = sum(cluster_size) with the same 1024-thread reduction (so all blocks of a stage,
// and the former one-block-per-stage kernel, agree bit for bit) and then takes its slice of the K*D elements.
constexpr int kEmaSplit = 16;
__global__ void __launch_bounds__(1024)
ema_apply_kernel(MutPtrTable32 cs_tab, MutPtrTable32 ea_tab, MutPtrTable32 em_tab, int stage_base, int K, int D,
                 const float* __restrict__ embed_sum, float decay, float alpha, float eps, float keps) {
  __shared__ float red[32];
  __shared__ float s_total;
  const int s = blockIdx.x;
  const float* cs = cs_tab.p[s];
  float* ea = ea_tab.p[s];
  float* em = em_tab.p[s];
  const float* es = embed_sum + size_t(stage_base + s) * K * D;
  float part = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) part += cs[k];
  #pragma unroll
  for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (threadIdx.x == 0) s_total = v;
  }
  __syncthreads();
  const float total = s_total;
  const int n = K * D;
  const int per = (n + gridDim.y - 1) / gridDim.y;
  const int i0 = blockIdx.y * per, i1 = min(n, i0 + per);
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    int k = i / D;
    float a = ea[i] * decay + alpha * es[i];
    ea[i] = a;
    float sm = (cs[k] + eps) / (total + keps) * total;
    em[i] = a / sm;
  }
}

__global__ void expire_replace_kernel(float* embed, const float* __restrict__ cluster_size,
                                      const float* __restrict__ samples, int K, int D, float thr) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K * D && cluster_size[i / D] < thr) embed[i] = samples[i];
}

// one warp per code k: if the code is dead, walk the selected frame's residual chain up to `stage`
__global__ void __launch_bounds__(256)
expire_codes_kernel(const unsigned char* pack, int K, int D, const float* __restrict__ x, rvq::FrameAddr fa,
                    int64_t N, int stage0, int stage, const int64_t* __restrict__ codes,
                    const int64_t* __restrict__ sel, const float* __restrict__ cluster_size, float thr,
                    float* __restrict__ embed, int ste) {
  rvq::PackView pv(pack, K, D);
  const int lane = threadIdx.x & 31;
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (k >= K || !(cluster_size[k] < thr)) return;
  int64_t n = sel[k];
  n = n < 0 ? 0 : (n >= N ? N - 1 : n);
  for (int d = lane; d < D; d += 32) {
    float r = x[fa.base(n) + int64_t(d) * fa.sxd];
    for (int i = 0; i < stage; ++i) {
      int64_t c = codes[int64_t(i) * N + n];
      int idx = c < 0 ? 0 : (c >= K ? K - 1 : int(c));
      float q = pv.tab32(stage0 + i)[size_t(idx) * D + d];
      if (ste) q = r + (q - r);
      r -= q;
    }
    embed[size_t(k) * D + d] = r;
  }
}

// ---- expiry of a whole residual stack without a host round trip (core_vq.py:165-175 for n stages) ----------------
// counter-based generator for the index draw: 64-bit finalizer of (seed, offset, stage, slot, round)
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ int64_t draw_index(uint64_t seed, uint64_t offset, int stage, int slot, int round, int64_t N) {
  const uint64_t r = mix64(mix64(seed ^ mix64(offset + uint64_t(stage))) ^ (uint64_t(slot) << 24) ^ uint64_t(round));
  return int64_t(__umul64hi(r, uint64_t(N)));             // uniform on [0, N) up to 2^-64 N
}
constexpr int kSelTable = 4096;                           // hash set of the draws of one stage (K <= kSelTable / 2)

// One block per stage.  fired[stage] = any(cluster_size < thr) (the reference's host-side torch.any, :168-170); for a
// firing stage, sel[stage, 0..K) = K distinct frame numbers, uniformly random and in random order -- the distribution
// of randperm(N)[:K] (sample_vectors, :69-77) -- or K draws with replacement when N < K (:75).  Duplicate draws are
// settled deterministically (lowest (round, slot) keeps the value, the others draw again), so a (seed, offset) pair
// reproduces the same indices.
__global__ void __launch_bounds__(1024)
expire_sample_kernel(PtrTable32 cs_tab, int stage_base, int K, int64_t N, float thr, uint64_t seed, uint64_t offset,
                     int64_t* __restrict__ sel, int* __restrict__ fired) {
  __shared__ unsigned keys[kSelTable];
  __shared__ int owner[kSelTable];
  const int tid = threadIdx.x, stage = stage_base + blockIdx.x;
  const float* cs = cs_tab.p[blockIdx.x];
  int dead = 0;
  for (int k = tid; k < K; k += blockDim.x) dead |= (cs[k] < thr) ? 1 : 0;
  dead = __syncthreads_or(dead);
  if (tid == 0) fired[stage] = dead;
  if (!dead) return;
  int64_t* out = sel + int64_t(stage) * K;
  if (N < K) {
    for (int k = tid; k < K; k += blockDim.x) out[k] = draw_index(seed, offset, stage, k, 0, N);
    return;
  }
  int M = 64;
  while (M < 2 * K) M <<= 1;                              // <= kSelTable (checked by the host entry)
  for (int i = tid; i < M; i += blockDim.x) { keys[i] = 0xffffffffu; owner[i] = 0x7fffffff; }
  __syncthreads();
  bool done[2] = {false, false};                          // slots tid, tid + 1024
  for (int round = 0;; ++round) {
    unsigned v[2] = {0u, 0u}; int h[2] = {0, 0};
    #pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = tid + j * 1024;
      if (k >= K || done[j]) continue;
      v[j] = unsigned(draw_index(seed, offset, stage, k, round, N));
      int hh = int((v[j] * 2654435761u) >> 7) & (M - 1);
      for (;;) {
        const unsigned prev = atomicCAS(&keys[hh], 0xffffffffu, v[j]);
        if (prev == 0xffffffffu || prev == v[j]) break;
        hh = (hh + 1) & (M - 1);
      }
      atomicMin(&owner[hh], (round << 12) | k);
      h[j] = hh;
    }
    __syncthreads();
    int pending = 0;
    #pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = tid + j * 1024;
      if (k >= K || done[j]) continue;
      if (owner[h[j]] == ((round << 12) | k)) { done[j] = true; out[k] = int64_t(v[j]); }
      else pending = 1;
    }
    if (!__syncthreads_or(pending)) break;
  }
}

// one warp per (stage, code k): a dead code takes the stage's input residual of frame sel[stage, k], recomputed from x and
// the codes with the encode arithmetic
__global__ void __launch_bounds__(256)
expire_stack_kernel(const unsigned char* pack, int K, int D, const float* __restrict__ x, rvq::FrameAddr fa, int64_t N,
                    int stage0, int stage_base, const int64_t* __restrict__ codes, const int64_t* __restrict__ sel,
                    const int* __restrict__ fired, PtrTable32 cs_tab, MutPtrTable32 em_tab, float thr, int ste) {
  const int rel = blockIdx.y, stage = stage_base + rel;   // stage: relative to stage0
  if (!fired[stage]) return;
  rvq::PackView pv(pack, K, D);
  const int lane = threadIdx.x & 31;
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (k >= K || !(cs_tab.p[rel][k] < thr)) return;
  int64_t n = sel[int64_t(stage) * K + k];
  n = n < 0 ? 0 : (n >= N ? N - 1 : n);
  float* embed = em_tab.p[rel];
  for (int d = lane; d < D; d += 32) {
    float r = x[fa.base(n) + int64_t(d) * fa.sxd];
    for (int i = 0; i < stage; ++i) {
      int64_t c = codes[int64_t(i) * N + n];
      int idx = c < 0 ? 0 : (c >= K ? K - 1 : int(c));
      float q = pv.tab32(stage0 + i)[size_t(idx) * D + d];
      if (ste) q = r + (q - r);
      r -= q;
    }
    embed[size_t(k) * D + d] = r;
  }
}

__global__ void kmeans_scatter_kernel(const float* __restrict__ samples, int64_t N, int D,
                                      const int64_t* __restrict__ buckets, int K,
                                      float* sums, unsigned long long* bins) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t n = warp; n < N; n += nwarps) {
    int64_t k = buckets[n];
    if (k < 0 || k >= K) continue;
    if (lane == 0) atomicAdd(&bins[k], 1ull);
    for (int d = lane; d < D; d += 32) atomicAdd(&sums[k * D + d], samples[n * D + d]);
  }
}

// D == 128: a warp walks a run of consecutive samples, lane = 16-byte chunk of the row, and keeps the running sum of the
// current bucket in registers; it is flushed (one red.global.add.v4.f32 per lane + one count) when the bucket changes.
// Lloyd iterations on residuals routinely produce a giant cluster (tens of thousands of samples on one mean): one scalar
// atomic per element then serialises 40 000 adds on each of its 128 addresses (0.5-1 ms per iteration, measured); runs of
// equal buckets collapse here, and every flush is a quarter of the instructions.
constexpr int kKmRun = 32;
__global__ void __launch_bounds__(256) kmeans_scatter128_kernel(const float* __restrict__ samples, int64_t N,
                                                                const int64_t* __restrict__ buckets, int K,
                                                                float* sums, unsigned long long* bins) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t n0 = warp * kKmRun;
  if (n0 >= N) return;
  const int64_t n1 = n0 + kKmRun < N ? n0 + kKmRun : N;
  // the run's buckets: one per lane, handed round with shuffles
  const int64_t kb = n0 + lane < n1 ? buckets[n0 + lane] : -1;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int cur = -1, cnt = 0;
  auto flush = [&]() {
    if (cur >= 0 && cnt > 0) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(sums + size_t(cur) * 128 + 4 * lane), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
      if (lane == 0) atomicAdd(&bins[cur], (unsigned long long)cnt);
    }
  };
  for (int i = 0; i < int(n1 - n0); ++i) {
    const int64_t k64 = __shfl_sync(0xffffffffu, kb, i);
    const int k = (k64 < 0 || k64 >= K) ? -1 : int(k64);
    if (k != cur) { flush(); cur = k; cnt = 0; acc = make_float4(0.f, 0.f, 0.f, 0.f); }
    if (k >= 0) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(samples + (n0 + i) * 128) + lane);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; ++cnt;
    }
  }
  flush();
}

__global__ void kmeans_finalize_kernel(float* means, const float* __restrict__ sums,
                                       const unsigned long long* __restrict__ bins, int K, int D) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K * D) {
    unsigned long long b = bins[i / D];
    if (b != 0) means[i] = sums[i] / float((long long)b);   // empty clusters keep the old mean (:100)
  }
}

// ------------------------------------------------------------------------------------------------
// C ABI entry points served by this translation unit
// ------------------------------------------------------------------------------------------------
extern "C" {

int rvq_decode(const void* pack, int K, int D, const int64_t* codes, int64_t scq, int64_t scb, int64_t sct,
               int n_q, int B, int T, float* out, void* stream) {
  return rvq_decode_ex(pack, K, D, codes, scq, scb, sct, n_q, B, T, out, 0, stream);
}

int rvq_decode_ex(const void* pack, int K, int D, const int64_t* codes, int64_t scq, int64_t scb, int64_t sct,
                  int n_q, int B, int T, float* out, int flags, void* stream) {
  if (int e = check_device()) return e;
  RVQ_REQUIRE(D % 4 == 0 && D > 0 && K > 0 && n_q >= 0 && B >= 0 && T >= 0, "rvq_decode: bad shape K=%d D=%d n_q=%d", K, D, n_q);
  const int64_t N = int64_t(B) * T;
  if (N == 0) return RVQ_OK;
  RVQ_REQUIRE(pack && out && (codes || n_q == 0), "rvq_decode: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  FrameAddr fa{0, 0, 0, T};
  if (flags & RVQ_FLAG_OUT_BDT) {
    RVQ_REQUIRE(D <= 256, "rvq_decode: dimension %d too large for the [B, D, T] output path", D);
    chain_bdt_kernel<kDecode><<<chain_bdt_grid(N), 256, size_t(D) * 33 * 4, st>>>((const unsigned char*)pack, K, D, nullptr, fa, N, 0, n_q,
                                                                               codes, scq, scb, sct, T, out, 0, 0);
    RVQ_LAUNCH_CHECK("chain_bdt_kernel<decode>");
    return RVQ_OK;
  }
  chain_kernel<kDecode><<<chain_grid(N), 256, 0, st>>>((const unsigned char*)pack, K, D, nullptr, fa, N, 0, n_q,
                                                      codes, scq, scb, sct, T, nullptr, out, nullptr, nullptr, 0);
  RVQ_LAUNCH_CHECK("chain_kernel<decode>");
  return RVQ_OK;
}

int rvq_ema_stats(const void* pack, int K, int D, const float* x, int64_t sxb, int64_t sxd, int64_t sxt,
                  int B, int T, int stage0, int n_q, const int64_t* codes, float* counts, float* embed_sum,
                  int flags, void* stream) {
  if (int e = check_device()) return e;
  RVQ_REQUIRE(pack && counts && embed_sum, "rvq_ema_stats: null pointer");
  RVQ_REQUIRE(D % 4 == 0 && D > 0 && K > 0 && n_q >= 0, "rvq_ema_stats: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  RVQ_CUDA(cudaMemsetAsync(counts, 0, size_t(n_q) * K * 4, st));
  RVQ_CUDA(cudaMemsetAsync(embed_sum, 0, size_t(n_q) * K * D * 4, st));
  const int64_t N = int64_t(B) * T;
  if (N == 0 || n_q == 0) return RVQ_OK;
  RVQ_REQUIRE(x && codes, "rvq_ema_stats: null pointer");
  FrameAddr fa{sxb, sxd, sxt, T};
  chain_kernel<kStats><<<chain_grid(N), 256, 0, st>>>((const unsigned char*)pack, K, D, x, fa, N, stage0, n_q, codes,
                                                     int64_t(B) * T, T, 1, T, nullptr, nullptr, counts, embed_sum,
                                                     (flags & RVQ_FLAG_STE) ? 1 : 0);
  RVQ_LAUNCH_CHECK("chain_kernel<stats>");
  return RVQ_OK;
}

int rvq_residual_combine(const void* pack, int K, int D, const float* x, int64_t sxb, int64_t sxd, int64_t sxt,
                         int B, int T, int stage0, int n_q, const int64_t* codes, const float* w, float* out,
                         int flags, void* stream) {
  if (int e = check_device()) return e;
  RVQ_REQUIRE(pack && out, "rvq_residual_combine: null pointer");
  RVQ_REQUIRE(D % 4 == 0 && D > 0 && K > 0 && n_q >= 0, "rvq_residual_combine: bad shape");
  const int64_t N = int64_t(B) * T;
  if (N == 0) return RVQ_OK;
  RVQ_REQUIRE(x && (n_q == 0 || (codes && w)), "rvq_residual_combine: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  FrameAddr fa{sxb, sxd, sxt, T};
  chain_kernel<kCombine><<<chain_grid(N), 256, 0, st>>>((const unsigned char*)pack, K, D, x, fa, N, stage0, n_q, codes,
                                                       int64_t(B) * T, T, 1, T, w, out, nullptr, nullptr,
                                                       (flags & RVQ_FLAG_STE) ? 1 : 0);
  RVQ_LAUNCH_CHECK("chain_kernel<combine>");
  return RVQ_OK;
}

int rvq_ema_apply(float* const* cluster_size_ptrs_host, float* const* embed_avg_ptrs_host,
                  float* const* embed_ptrs_host, int n_q, int K, int D, const float* counts,
                  const float* embed_sum, double decay, double epsilon, void* stream) {
  if (int e = check_device()) return e;
  RVQ_REQUIRE(cluster_size_ptrs_host && embed_avg_ptrs_host && embed_ptrs_host && counts && embed_sum,
              "rvq_ema_apply: null pointer");
  RVQ_REQUIRE(K > 0 && D > 0 && n_q >= 0, "rvq_ema_apply: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const float alpha = float(1.0 - decay);             // python: (1 - decay) in double, then fp32
  const float keps = float(double(K) * epsilon);
  for (int s0 = 0; s0 < n_q; s0 += 32) {
    int ns = n_q - s0 < 32 ? n_q - s0 : 32;
    MutPtrTable32 a, b, c;
    for (int i = 0; i < 32; ++i) {
      a.p[i] = i < ns ? cluster_size_ptrs_host[s0 + i] : nullptr;
      b.p[i] = i < ns ? embed_avg_ptrs_host[s0 + i] : nullptr;
      c.p[i] = i < ns ? embed_ptrs_host[s0 + i] : nullptr;
    }
    ema_cluster_kernel<<<ns, 1024, 0, st>>>(a, s0, K, counts, float(decay), alpha);
    RVQ_LAUNCH_CHECK("ema_cluster_kernel");
    ema_apply_kernel<<<dim3(ns, kEmaSplit), 1024, 0, st>>>(a, b, c, s0, K, D, embed_sum, float(decay), alpha, float(epsilon), keps);
    RVQ_LAUNCH_CHECK("ema_apply_kernel");
  }
  return RVQ_OK;
}

int rvq_expire_replace(float* embed, const float* cluster_size, const float* samples, int K, int D,
                       float threshold, void* stream) {
  if (int e = check_device()) return e;
  RVQ_REQUIRE(embed && cluster_size && samples && K > 0 && D > 0, "rvq_expire_replace: bad argument");
  int n = K * D;
  expire_replace_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(embed, cluster_size, samples, K, D, threshold);
  RVQ_LAUNCH_CHECK("expire_replace_kernel");
  return RVQ_OK;
}

int rvq_expire_codes(const void* pack, int K, int D, const float* x, int64_t sxb, int64_t sxd, int64_t sxt,
                     int B, int T, int stage0, int stage, const int64_t* codes, const int64_t* sel,
                     const float* cluster_size, float threshold, float* embed, int flags, void* stream) {
  if (int e = check_device()) return e;
  RVQ_REQUIRE(pack && x && sel && cluster_size && embed && (codes || stage == 0), "rvq_expire_codes: null pointer");
  RVQ_REQUIRE(K > 0 && D > 0 && stage >= 0 && stage0 >= 0, "rvq_expire_codes: bad shape");
  const int64_t N = int64_t(B) * T;
  RVQ_REQUIRE(N > 0, "rvq_expire_codes: empty batch");
  FrameAddr fa{sxb, sxd, sxt, T};
  expire_codes_kernel<<<(K + 7) / 8, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)pack, K, D, x, fa, N, stage0,
                                                                     stage, codes, sel, cluster_size, threshold, embed,
                                                                     (flags & RVQ_FLAG_STE) ? 1 : 0);
  RVQ_LAUNCH_CHECK("expire_codes_kernel");
  return RVQ_OK;
}

int rvq_expire_stack(const void* pack, int K, int D, const float* x, int64_t sxb, int64_t sxd, int64_t sxt,
                     int B, int T, int stage0, int n_q, const int64_t* codes,
                     const float* const* cluster_size_ptrs_host, float* const* embed_ptrs_host, float threshold,
                     uint64_t seed, uint64_t offset, int64_t* sel, int* fired, int flags, void* stream) {
  if (int e = check_device()) return e;
  RVQ_REQUIRE(pack && x && sel && fired && cluster_size_ptrs_host && embed_ptrs_host && (codes || n_q <= 1),
              "rvq_expire_stack: null pointer");
  RVQ_REQUIRE(K > 0 && 2 * K <= kSelTable && D > 0 && n_q >= 0 && stage0 >= 0, "rvq_expire_stack: bad shape (K=%d)", K);
  const int64_t N = int64_t(B) * T;
  RVQ_REQUIRE(N > 0 && N < (int64_t(1) << 31), "rvq_expire_stack: frame count out of range");
  cudaStream_t st = (cudaStream_t)stream;
  FrameAddr fa{sxb, sxd, sxt, T};
  for (int s0 = 0; s0 < n_q; s0 += 32) {
    const int ns = n_q - s0 < 32 ? n_q - s0 : 32;
    PtrTable32 cs;
    MutPtrTable32 em;
    for (int i = 0; i < 32; ++i) {
      cs.p[i] = i < ns ? cluster_size_ptrs_host[s0 + i] : nullptr;
      em.p[i] = i < ns ? embed_ptrs_host[s0 + i] : nullptr;
    }
    expire_sample_kernel<<<ns, 1024, 0, st>>>(cs, s0, K, N, threshold, seed, offset, sel, fired);
    RVQ_LAUNCH_CHECK("expire_sample_kernel");
    expire_stack_kernel<<<dim3((K + 7) / 8, ns), 256, 0, st>>>((const unsigned char*)pack, K, D, x, fa, N, stage0, s0, codes,
                                                             sel, fired, cs, em, threshold, (flags & RVQ_FLAG_STE) ? 1 : 0);
    RVQ_LAUNCH_CHECK("expire_stack_kernel");
  }
  return RVQ_OK;
}

int rvq_kmeans_update(const float* samples, int64_t N, int D, const int64_t* buckets, int K, float* means,
                      int64_t* bins, float* sums, void* stream) {
  if (int e = check_device()) return e;
  RVQ_REQUIRE(samples && buckets && means && bins && sums && K > 0 && D > 0 && N >= 0, "rvq_kmeans_update: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  RVQ_CUDA(cudaMemsetAsync(bins, 0, size_t(K) * 8, st));
  RVQ_CUDA(cudaMemsetAsync(sums, 0, size_t(K) * D * 4, st));
  if (N > 0) {
    if (D == 128 && (reinterpret_cast<uintptr_t>(samples) & 15) == 0 && (reinterpret_cast<uintptr_t>(sums) & 15) == 0) {
      const int64_t warps = (N + kKmRun - 1) / kKmRun;
      kmeans_scatter128_kernel<<<unsigned((warps + 7) / 8), 256, 0, st>>>(samples, N, buckets, K, sums, (unsigned long long*)bins);
    } else {
      kmeans_scatter_kernel<<<chain_grid(N), 256, 0, st>>>(samples, N, D, buckets, K, sums, (unsigned long long*)bins);
    }
    RVQ_LAUNCH_CHECK("kmeans_scatter_kernel");
  }
  int n = K * D;
  kmeans_finalize_kernel<<<(n + 255) / 256, 256, 0, st>>>(means, sums, (const unsigned long long*)bins, K, D);
  RVQ_LAUNCH_CHECK("kmeans_finalize_kernel");
  return RVQ_OK;
}

}  // extern "C"

// Fused multi-stage nearest-code search on the 5th-generation tensor cores (tcgen05 / TMEM).
//
// One persistent CTA per SM walks 128-frame tiles.  Per tile the fp32 residual of every frame
// stays in shared memory across all n_q stages (core_vq.py:357-367 without the per-stage round
// trips); per stage the scores  S[f,k] = -2 r_f . c_k + |c_k|^2  of all K codes are produced by
// tcgen05.mma (fp16 operands, fp32 accumulate in TMEM) from
//   A = fp16(r)  [128 frames x 144]  written by the frame threads (cols 128,129 = 1.0),
//   B = fp16 image of the stage's codebook [128 codes x 144] per chunk (cols 0..127 = -2c,
//       cols 128,129 = hi/lo halves of |c|^2), streamed by the TMA engine (cp.async.bulk) from the
//       pre-arranged pack into a 3-slot ring.
// The 128 frame threads (thread = TMEM lane = frame) read the scores back with tcgen05.ld and
// keep, per frame, the minimum over each 32-code batch and over each residue class (code mod 32).
// A code is within `delta` of the minimum iff its batch AND its class are; delta bounds the fp16
// score error two-sidedly, so the exact fp32 winner is certified when one batch and one class
// qualify, and otherwise the (at most four) candidates are re-scored with the reference's fp32
// formula (core_vq.py:181-189, ties -> lowest index).  Anything else (fp16-range outliers, wide
// ties) falls back to an exact fp32 scan of the whole table for that frame.  The winner's fp32
// row is gathered, the residual updated exactly (core_vq.py:364 / :348 with the straight-through
// arithmetic of :309 in training), and the new fp16 A operand written for the next stage.
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc), 2..5 = frame threads.
#include "rvq_common.cuh"
#include "rvq_ptx.cuh"

namespace rvq {

namespace {

constexpr int kM = 128;                 // frames per tile (UMMA M, TMEM lanes)
constexpr int kNChunk = kTcChunkCodes;  // 128 codes per MMA group (UMMA N)
constexpr int kRing = 3;                // B ring slots
constexpr int kAcc = 4;                 // TMEM accumulator buffers of 128 columns
constexpr int kKSteps = kTcKPad / 16;   // 9 UMMA K steps of 16
constexpr int kThreadsTc = 192;
constexpr int kMaxBatches = 32;         // K <= 1024 on this path
constexpr uint32_t kABytes = kM * kTcKPad * 2;   // 36864
constexpr uint32_t kLBO = 2048, kSBO = 128;       // see rvq_common.cuh (pack image layout)

struct SmemLayout {
  static constexpr uint32_t a = 0;
  static constexpr uint32_t b = a + kABytes;
  static constexpr uint32_t rs = b + kRing * kTcChunkBytes;          // fp32 residual [128 d][128 f]
  static constexpr uint32_t bmin = rs + 128 * kM * 4;                // fp32 batch minima [32][128 f]
  static constexpr uint32_t bars = bmin + kMaxBatches * kM * 4;
  static constexpr uint32_t total = bars + 256;
};
struct Bars {
  uint64_t full[kRing], empty[kRing], acc_full[kAcc], acc_empty[kAcc], a_ready;
  uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 256, "barrier block");
static_assert(SmemLayout::total <= 227 * 1024, "shared memory budget");

struct TcParams {
  const unsigned char* pack; int K;
  const float* x; FrameAddr fa; int64_t N;
  int stage0, n_q;
  int64_t* codes; float* residual_out; double* sqerr;
  int ste;
  unsigned long long* counters;
};

__device__ __forceinline__ float inf_f() { return __int_as_float(0x7f800000); }

}  // namespace

__global__ void __launch_bounds__(kThreadsTc, 1) tc_encode_kernel(const TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  Bars* bars = reinterpret_cast<Bars*>(smem + SmemLayout::bars);
  float* rs = reinterpret_cast<float*>(smem + SmemLayout::rs);
  float* sbmin = reinterpret_cast<float*>(smem + SmemLayout::bmin);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = 128;
  PackView pv(p.pack, p.K, D);
  const int nchunks = p.K / kNChunk;
  const int64_t ntiles = (p.N + kM - 1) / kM;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) { ptx::mbar_init(ptx::smem_u32(&bars->full[i]), 1); ptx::mbar_init(ptx::smem_u32(&bars->empty[i]), 1); }
    for (int i = 0; i < kAcc; ++i) { ptx::mbar_init(ptx::smem_u32(&bars->acc_full[i]), 1); ptx::mbar_init(ptx::smem_u32(&bars->acc_empty[i]), 4); }
    ptx::mbar_init(ptx::smem_u32(&bars->a_ready), kM);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    // ===== TMA producer: the same chunk sequence (stage-major) for every tile of this CTA =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int s = 0; s < p.n_q; ++s) {
          const unsigned char* img = pv.tc(p.stage0 + s);
          for (int c = 0; c < nchunks; ++c, ++it) {
            const uint32_t slot = it % kRing, ph = (it / kRing) & 1;
            ptx::mbar_wait(ptx::smem_u32(&bars->empty[slot]), ph ^ 1);
            const uint32_t fb = ptx::smem_u32(&bars->full[slot]);
            ptx::mbar_expect_tx(fb, kTcChunkBytes);
            ptx::bulk_g2s(sbase + SmemLayout::b + slot * kTcChunkBytes, img + size_t(c) * kTcChunkBytes, kTcChunkBytes, fb);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_f16_f32(kM, kNChunk);
      uint32_t it = 0, ait = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int s = 0; s < p.n_q; ++s, ++ait) {
          ptx::mbar_wait(ptx::smem_u32(&bars->a_ready), ait & 1);          // fp16 residual operand written
          ptx::tc_fence_after();
          for (int c = 0; c < nchunks; ++c, ++it) {
            const uint32_t slot = it % kRing, ph = (it / kRing) & 1;
            const uint32_t buf = it % kAcc, aph = (it / kAcc) & 1;
            ptx::mbar_wait(ptx::smem_u32(&bars->acc_empty[buf]), aph ^ 1);   // epilogue drained this accumulator
            ptx::mbar_wait(ptx::smem_u32(&bars->full[slot]), ph);            // codebook chunk landed
            ptx::tc_fence_after();
            const uint32_t a_addr = sbase + SmemLayout::a;
            const uint32_t b_addr = sbase + SmemLayout::b + slot * kTcChunkBytes;
            #pragma unroll
            for (int k = 0; k < kKSteps; ++k) {
              const uint64_t ad = ptx::umma_desc_kmajor_noswz(a_addr + k * 2 * kLBO, kLBO, kSBO);
              const uint64_t bd = ptx::umma_desc_kmajor_noswz(b_addr + k * 2 * kLBO, kLBO, kSBO);
              ptx::umma_f16_ss(tmem + buf * kNChunk, ad, bd, idesc, k > 0 ? 1u : 0u);
            }
            ptx::umma_commit(ptx::smem_u32(&bars->empty[slot]));       // ring slot reusable once the MMAs have read it
            ptx::umma_commit(ptx::smem_u32(&bars->acc_full[buf]));     // scores ready for the frame threads
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ===== frame threads: thread <-> TMEM lane <-> frame of the tile =====
    const int q = warp & 3;                    // TMEM lane quadrant this warp may access
    const int f = q * 32 + lane;               // frame row within the tile
    const uint32_t tlane = tmem + (uint32_t(q * 32) << 16);
    unsigned char* arow = smem + SmemLayout::a + f * 16;
    // augmented K columns never change: col 128,129 = 1 (pick up hi/lo of |c|^2), rest 0
    {
      __align__(16) __half h[8];
      #pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = __float2half_rn(j < 2 ? 1.f : 0.f);
      *reinterpret_cast<uint4*>(arow + 16 * kLBO) = *reinterpret_cast<const uint4*>(h);
      *reinterpret_cast<uint4*>(arow + 17 * kLBO) = make_uint4(0u, 0u, 0u, 0u);
    }
    unsigned long long n_cert = 0, n_resc = 0, n_full = 0, n_all = 0;
    long long t_wait = 0, t_epi = 0, t_win = 0, t_upd = 0, t_load = 0;   // phase cycles (lane 0 of each warp)
    const long long t_begin = clock64();
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t n = tile * kM + f;
      const bool valid = n < p.N;
      long long tc0 = clock64();
      // ---- load the latent, |x|^2 with the exact path's summation order, fp16 operand row ----
      float xx;
      {
        const int64_t xb = valid ? p.fa.base(n) : 0;
        float part[4] = {0.f, 0.f, 0.f, 0.f};
        #pragma unroll 4
        for (int g = 0; g < 16; ++g) {
          __align__(16) __half h[8];
          #pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int d = g * 8 + j;
            const float v = valid ? __ldg(p.x + xb + int64_t(d) * p.fa.sxd) : 0.f;
            rs[d * kM + f] = v;
            part[g >> 2] = fmaf(v, v, part[g >> 2]);
            h[j] = __float2half_rn(v);
          }
          *reinterpret_cast<uint4*>(arow + g * kLBO) = *reinterpret_cast<const uint4*>(h);
        }
        xx = ((part[0] + part[1]) + part[2]) + part[3];
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(ptx::smem_u32(&bars->a_ready));
      { const long long t = clock64(); t_load += t - tc0; tc0 = t; }

      for (int s = 0; s < p.n_q; ++s) {
        const int st = p.stage0 + s;
        const StageMeta* meta = pv.meta(st);
        const float xnorm = sqrtf(xx);
        const float delta = meta->margin_coef * (xnorm + 1e-3f) + meta->margin_abs;
        const bool outl = !(xnorm < meta->xlimit);      // also true for NaN
        float cm[32];
        #pragma unroll
        for (int j = 0; j < 32; ++j) cm[j] = inf_f();

        for (int c = 0; c < nchunks; ++c, ++it) {
          const uint32_t buf = it % kAcc, aph = (it / kAcc) & 1;
          ptx::mbar_wait(ptx::smem_u32(&bars->acc_full[buf]), aph);
          ptx::tc_fence_after();
          { const long long t = clock64(); t_wait += t - tc0; tc0 = t; }
          #pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t v0[32], v1[32];
            ptx::tmem_ld32(tlane + buf * kNChunk + half * 64, v0);
            ptx::tmem_ld32(tlane + buf * kNChunk + half * 64 + 32, v1);
            ptx::tmem_ld_wait();
            #pragma unroll
            for (int j = 0; j < 32; ++j) cm[j] = ptx::fmin3(cm[j], __uint_as_float(v0[j]), __uint_as_float(v1[j]));
            float t0[11], t1[11];
            #pragma unroll
            for (int j = 0; j < 10; ++j) {
              t0[j] = ptx::fmin3(__uint_as_float(v0[3 * j]), __uint_as_float(v0[3 * j + 1]), __uint_as_float(v0[3 * j + 2]));
              t1[j] = ptx::fmin3(__uint_as_float(v1[3 * j]), __uint_as_float(v1[3 * j + 1]), __uint_as_float(v1[3 * j + 2]));
            }
            t0[10] = fminf(__uint_as_float(v0[30]), __uint_as_float(v0[31]));
            t1[10] = fminf(__uint_as_float(v1[30]), __uint_as_float(v1[31]));
            float b0 = ptx::fmin3(ptx::fmin3(t0[0], t0[1], t0[2]), ptx::fmin3(t0[3], t0[4], t0[5]), ptx::fmin3(t0[6], t0[7], t0[8]));
            float b1 = ptx::fmin3(ptx::fmin3(t1[0], t1[1], t1[2]), ptx::fmin3(t1[3], t1[4], t1[5]), ptx::fmin3(t1[6], t1[7], t1[8]));
            b0 = ptx::fmin3(b0, t0[9], t0[10]);
            b1 = ptx::fmin3(b1, t1[9], t1[10]);
            sbmin[(c * 4 + half * 2) * kM + f] = b0;
            sbmin[(c * 4 + half * 2 + 1) * kM + f] = b1;
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->acc_empty[buf]));
          { const long long t = clock64(); t_epi += t - tc0; tc0 = t; }
        }

        // ---- winner: certified / re-scored / exact scan ----
        float m = inf_f();
        #pragma unroll
        for (int j = 0; j < 32; ++j) m = fminf(m, cm[j]);
        const float thr = m + delta;
        uint32_t cmask = 0, bmask = 0;
        #pragma unroll
        for (int j = 0; j < 32; ++j) cmask |= (cm[j] <= thr) ? (1u << j) : 0u;
        const int nb = nchunks * 4;
        for (int a = 0; a < nb; ++a) bmask |= (sbmin[a * kM + f] <= thr) ? (1u << a) : 0u;
        const int ncl = __popc(cmask), nba = __popc(bmask);
        const float* t32 = pv.tab32(st);
        const float* cn = pv.cnorm(st);
        int idx = 0;
        const bool need_full = outl || cmask == 0u || bmask == 0u;       // masks are empty only for NaN scores
        if (!need_full) {
          if (ncl == 1 && nba == 1) {
            idx = (__ffs(bmask) - 1) * 32 + (__ffs(cmask) - 1);
            ++n_cert;
          } else {
            // candidates = flagged batches x flagged classes, visited in ascending code order so that
            // the strict '<' keeps the lowest index among exact ties (core_vq.py:188)
            float best = inf_f(); int bi = 0x7fffffff;
            uint32_t bm2 = bmask;
            while (bm2) {
              const int a = __ffs(bm2) - 1; bm2 &= bm2 - 1;
              uint32_t cm2 = cmask;
              while (cm2) {
                const int j = __ffs(cm2) - 1; cm2 &= cm2 - 1;
                const int code = a * 32 + j;
                const float4* row = reinterpret_cast<const float4*>(t32 + size_t(code) * D);
                float acc = 0.f;
                #pragma unroll 4
                for (int d4 = 0; d4 < 32; ++d4) {
                  const float4 cv = __ldg(row + d4);
                  acc = fmaf(rs[(d4 * 4 + 0) * kM + f], cv.x, acc);
                  acc = fmaf(rs[(d4 * 4 + 1) * kM + f], cv.y, acc);
                  acc = fmaf(rs[(d4 * 4 + 2) * kM + f], cv.z, acc);
                  acc = fmaf(rs[(d4 * 4 + 3) * kM + f], cv.w, acc);
                }
                const float dist = (xx - 2.f * acc) + __ldg(cn + code);     // core_vq.py:183-187
                if (dist < best) { best = dist; bi = code; }
              }
            }
            idx = bi == 0x7fffffff ? 0 : bi;
            ++n_resc;
          }
        }
        if (__any_sync(0xffffffffu, need_full)) {
          // frames outside the fp16 image's validity range (or NaN): exact fp32 scan of the whole
          // table, all lanes in lockstep (code rows are warp-uniform loads)
          float best = inf_f(); int bi = 0x7fffffff;
          for (int k = 0; k < p.K; ++k) {
            const float4* row = reinterpret_cast<const float4*>(t32 + size_t(k) * D);
            float acc = 0.f;
            #pragma unroll 4
            for (int d4 = 0; d4 < 32; ++d4) {
              const float4 cv = __ldg(row + d4);
              acc = fmaf(rs[(d4 * 4 + 0) * kM + f], cv.x, acc);
              acc = fmaf(rs[(d4 * 4 + 1) * kM + f], cv.y, acc);
              acc = fmaf(rs[(d4 * 4 + 2) * kM + f], cv.z, acc);
              acc = fmaf(rs[(d4 * 4 + 3) * kM + f], cv.w, acc);
            }
            const float dist = (xx - 2.f * acc) + __ldg(cn + k);
            if (dist < best) { best = dist; bi = k; }
          }
          if (need_full) { idx = bi == 0x7fffffff ? 0 : bi; ++n_full; }
        }
        ++n_all;
        { const long long t = clock64(); t_win += t - tc0; tc0 = t; }

        // ---- gather the fp32 row, exact residual update, next stage's fp16 operand ----
        {
          const float4* row = reinterpret_cast<const float4*>(t32 + size_t(idx) * D);
          float part[4] = {0.f, 0.f, 0.f, 0.f};
          #pragma unroll 4
          for (int g = 0; g < 16; ++g) {
            const float4 c0 = __ldg(row + 2 * g), c1 = __ldg(row + 2 * g + 1);
            const float cq[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
            __align__(16) __half h[8];
            #pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int d = g * 8 + j;
              const float r = rs[d * kM + f];
              float qv = cq[j];
              if (p.ste) qv = r + (qv - r);                 // core_vq.py:309
              const float rn = r - qv;                       // core_vq.py:364 / :348
              rs[d * kM + f] = rn;
              part[g >> 2] = fmaf(rn, rn, part[g >> 2]);
              h[j] = __float2half_rn(rn);
            }
            *reinterpret_cast<uint4*>(arow + g * kLBO) = *reinterpret_cast<const uint4*>(h);
          }
          xx = ((part[0] + part[1]) + part[2]) + part[3];
        }
        ptx::fence_proxy_async_smem();
        if (s + 1 < p.n_q) ptx::mbar_arrive(ptx::smem_u32(&bars->a_ready));
        if (valid) p.codes[int64_t(s) * p.N + n] = idx;
        if (p.sqerr != nullptr) {
          float v = valid ? xx : 0.f;
          #pragma unroll
          for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
          if (lane == 0) atomicAdd(&p.sqerr[s], (double)v);
        }
        { const long long t = clock64(); t_upd += t - tc0; tc0 = t; }
      }
      if (p.residual_out != nullptr && valid) {
        float* out = p.residual_out + n * D;
        #pragma unroll 4
        for (int d4 = 0; d4 < 32; ++d4)
          *reinterpret_cast<float4*>(out + d4 * 4) = make_float4(rs[(d4 * 4) * kM + f], rs[(d4 * 4 + 1) * kM + f],
                                                                 rs[(d4 * 4 + 2) * kM + f], rs[(d4 * 4 + 3) * kM + f]);
      }
    }
    // search statistics (evidence for the certified / re-scored / exact-scan split)
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      n_all += __shfl_xor_sync(0xffffffffu, n_all, off);
      n_cert += __shfl_xor_sync(0xffffffffu, n_cert, off);
      n_resc += __shfl_xor_sync(0xffffffffu, n_resc, off);
      n_full += __shfl_xor_sync(0xffffffffu, n_full, off);
    }
    if (lane == 0 && p.counters != nullptr) {
      atomicAdd(&p.counters[0], n_all); atomicAdd(&p.counters[1], n_cert);
      atomicAdd(&p.counters[2], n_resc); atomicAdd(&p.counters[3], n_full);
      atomicAdd(&p.counters[4], (unsigned long long)t_wait); atomicAdd(&p.counters[5], (unsigned long long)t_epi);
      atomicAdd(&p.counters[6], (unsigned long long)t_win);  atomicAdd(&p.counters[7], (unsigned long long)t_upd);
      atomicAdd(&p.counters[8], (unsigned long long)t_load); atomicAdd(&p.counters[9], (unsigned long long)(clock64() - t_begin));
      atomicAdd(&p.counters[10], 1ull);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, 512);
}

int simt_quant_sum(const void* pack, int K, int D, const float* x, FrameAddr fa, int64_t N, int T, int stage0, int n_q,
                   const int64_t* codes, float* out, int ste, int accum, cudaStream_t st);

int tc_encode(const EncodeArgs& a, cudaStream_t st) {
  const int64_t N = int64_t(a.B) * a.T;
  if (N == 0 || a.n_q == 0) return RVQ_OK;
  RVQ_REQUIRE(a.D == 128 && a.K % kNChunk == 0 && a.K <= kMaxBatches * 32, "tc_encode: shape D=%d K=%d", a.D, a.K);
  static thread_local int sm_count = 0, sm_dev = -1;
  int dev = 0;
  RVQ_CUDA(cudaGetDevice(&dev));
  if (dev != sm_dev) {
    RVQ_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    RVQ_CUDA(cudaFuncSetAttribute(tc_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemLayout::total));
    sm_dev = dev;
  }
  PackView pv(a.pack, a.K, a.D);
  RVQ_CUDA(cudaMemsetAsync(pv.counters(), 0, 16 * sizeof(unsigned long long), st));
  TcParams p;
  p.pack = (const unsigned char*)a.pack; p.K = a.K;
  p.x = a.x; p.fa = FrameAddr{a.sxb, a.sxd, a.sxt, a.T}; p.N = N;
  p.stage0 = a.stage0; p.n_q = a.n_q;
  p.codes = a.codes; p.residual_out = a.residual_out; p.sqerr = a.sqerr;
  p.ste = (a.flags & RVQ_FLAG_STE) ? 1 : 0;
  p.counters = pv.counters();
  const int64_t ntiles = (N + kM - 1) / kM;
  const unsigned grid = unsigned(ntiles < sm_count ? ntiles : sm_count);
  tc_encode_kernel<<<grid, kThreadsTc, SmemLayout::total, st>>>(p);
  RVQ_LAUNCH_CHECK("tc_encode_kernel");
  if (a.quantized != nullptr)
    return simt_quant_sum(a.pack, a.K, a.D, a.x, p.fa, N, a.T, a.stage0, a.n_q, a.codes, a.quantized, p.ste,
                          (a.flags & RVQ_FLAG_ACCUM_Q) ? 1 : 0, st);
  return RVQ_OK;
}

}  // namespace rvq

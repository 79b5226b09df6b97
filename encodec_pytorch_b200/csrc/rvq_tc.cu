// Fused multi-stage nearest-code search on the 5th-generation tensor cores (tcgen05 / TMEM).
//
// One persistent CTA per SM owns a contiguous range of 128-frame tiles and walks it two tiles at
// a time ("round"); both tiles of a round consume the SAME codebook chunk stream, so each chunk
// is fetched from L2 once per 256 frames.  Per tile the fp32 residual of every frame stays in
// shared memory across all n_q stages (core_vq.py:357-367 without the per-stage round trips);
// per stage the scores  S[f,k] = -2 r_f . c_k + |c_k|^2  of all K codes come from tcgen05.mma
// (fp16 operands, fp32 accumulation in TMEM):
//   A = fp16(r) [128 frames x 144], kept in TENSOR MEMORY (written with tcgen05.st by the frame
//       threads; columns 128,129 = 1.0 pick up the two halves of |c|^2),
//   B = fp16 image of the codebook, 64 codes x 144 per chunk (cols 0..127 = -2c, cols 128,129 =
//       hi/lo halves of |c|^2), streamed by the TMA engine (cp.async.bulk) from the pre-arranged
//       pack into a shared-memory ring.
// Each tile has 128 frame threads (thread = TMEM lane = frame).  They read the scores back with
// tcgen05.ld and keep, per frame, the minimum over every 32-code batch and over every residue
// class (code mod 32).  A code is within `delta` of the minimum iff its batch AND its class are;
// delta bounds the fp16 score error two-sidedly (rvq_common.cuh, StageMeta), so the exact fp32
// winner is certified when exactly one batch and one class qualify.  Otherwise the candidates
// (flagged batches x flagged classes, usually 2..4 codes) are re-scored in fp32 with the
// reference's formula (core_vq.py:181-189, ties -> lowest index), warp-cooperatively.  Frames
// outside the fp16 image's validity range fall back to an exact fp32 scan of the table.  Then the
// winner's fp32 row is gathered, the residual updated exactly (core_vq.py:364 / :348, with the
// straight-through arithmetic of :309 in training) and the fp16 operand of the next stage written.
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc), 2..5 = frames of tile slot 0,
// 6..9 = frames of tile slot 1.
#include "rvq_common.cuh"
#include "rvq_ptx.cuh"

namespace rvq {

namespace {

constexpr int kM = 128;                 // frames per tile (UMMA M, TMEM lanes)
constexpr int kN = kTcChunkCodes;       // 64 codes per MMA group (UMMA N)
constexpr int kRing = 3;                // B ring slots
constexpr int kKSteps = kTcKPad / 16;   // 9 UMMA K steps of 16
constexpr int kGroups = 2;              // tile slots per CTA
constexpr int kThreadsTc = 64 + kGroups * 128;
constexpr int kMaxBatches = 32;         // K <= 1024 on this path
constexpr int kListMax = 64;            // re-score entries per warp and stage handled cooperatively
// TMEM columns of tile slot g: [g*256, +64) and [+64, +128) score buffers, [+128, +200) fp16 A operand
constexpr uint32_t kTmemSlot = 256, kTmemA = 128;

struct SmemLayout {
  static constexpr uint32_t b = 0;                                         // ring of codebook chunks
  static constexpr uint32_t rs = b + kRing * kTcChunkBytes;                // fp32 residual [g][128 d][128 f] (swizzled)
  static constexpr uint32_t bmin = rs + kGroups * 128 * kM * 4;            // fp32 batch minima [g][32][128 f]
  static constexpr uint32_t list = bmin + kGroups * kMaxBatches * kM * 4;  // re-score lists [8 warps][64]
  static constexpr uint32_t bars = list + kGroups * 4 * kListMax * 4;
  static constexpr uint32_t total = bars + 256;
};
struct Bars {
  uint64_t full[kRing], empty[kRing], acc_full[kGroups][2], acc_empty[kGroups][2], a_ready[kGroups];
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
// residual element (dim d, frame f) with an XOR swizzle: conflict-free both for "thread = frame,
// fixed d" and for "fixed frame, lane = dim/4" (the cooperative re-score)
__device__ __forceinline__ int rs_idx(int d, int f) { return d * kM + (f ^ ((d >> 2) & 31)); }
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float warp_sum(float v) {
  #pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// tiles [start, start+cnt) of this CTA
__device__ __forceinline__ void cta_range(int64_t ntiles, int64_t& start, int64_t& cnt) {
  const int64_t base = ntiles / gridDim.x, rem = ntiles % gridDim.x;
  const int64_t b = blockIdx.x;
  start = b * base + (b < rem ? b : rem);
  cnt = base + (b < rem ? 1 : 0);
}

}  // namespace

__global__ void __launch_bounds__(kThreadsTc, 1) tc_encode_kernel(const TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  Bars* bars = reinterpret_cast<Bars*>(smem + SmemLayout::bars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = 128;
  PackView pv(p.pack, p.K, D);
  const int nchunks = p.K / kN;
  const int64_t ntiles = (p.N + kM - 1) / kM;
  int64_t tile0, tcnt;
  cta_range(ntiles, tile0, tcnt);
  const int rounds = int((tcnt + kGroups - 1) / kGroups);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) { ptx::mbar_init(ptx::smem_u32(&bars->full[i]), 1); ptx::mbar_init(ptx::smem_u32(&bars->empty[i]), 1); }
    for (int g = 0; g < kGroups; ++g) {
      for (int i = 0; i < 2; ++i) { ptx::mbar_init(ptx::smem_u32(&bars->acc_full[g][i]), 1); ptx::mbar_init(ptx::smem_u32(&bars->acc_empty[g][i]), 4); }
      ptx::mbar_init(ptx::smem_u32(&bars->a_ready[g]), kM);
    }
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
    // ===== TMA producer: one chunk stream (stage-major) per round, shared by both tile slots =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int r = 0; r < rounds; ++r) {
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
      const uint32_t idesc = ptx::umma_idesc_f16_f32(kM, kN);
      uint32_t it = 0, ait = 0, acc_it[kGroups] = {0, 0};
      for (int r = 0; r < rounds; ++r) {
        const bool act1 = int64_t(r) * kGroups + 1 < tcnt;       // slot 1 idle in an odd last round
        for (int s = 0; s < p.n_q; ++s, ++ait) {
          for (int c = 0; c < nchunks; ++c, ++it) {
            const uint32_t slot = it % kRing, ph = (it / kRing) & 1;
            ptx::mbar_wait(ptx::smem_u32(&bars->full[slot]), ph);                 // codebook chunk landed
            const uint32_t b_addr = sbase + SmemLayout::b + slot * kTcChunkBytes;
            #pragma unroll
            for (int g = 0; g < kGroups; ++g) {
              if (g == 1 && !act1) continue;
              if (c == 0) ptx::mbar_wait(ptx::smem_u32(&bars->a_ready[g]), ait & 1);   // fp16 residual operand in TMEM
              const uint32_t buf = acc_it[g] & 1, aph = (acc_it[g] >> 1) & 1;
              ++acc_it[g];
              ptx::mbar_wait(ptx::smem_u32(&bars->acc_empty[g][buf]), aph ^ 1);   // frame threads drained this buffer
              ptx::tc_fence_after();
              const uint32_t d_tmem = tmem + g * kTmemSlot + buf * kN;
              const uint32_t a_tmem = tmem + g * kTmemSlot + kTmemA;
              #pragma unroll
              for (int k = 0; k < kKSteps; ++k) {
                const uint64_t bd = ptx::umma_desc_kmajor_noswz(b_addr + k * 2 * kTcLBO, kTcLBO, kTcSBO);
                ptx::umma_f16_ts(d_tmem, a_tmem + k * 8, bd, idesc, k > 0 ? 1u : 0u);
              }
              ptx::umma_commit(ptx::smem_u32(&bars->acc_full[g][buf]));   // scores ready for the frame threads
            }
            ptx::umma_commit(ptx::smem_u32(&bars->empty[slot]));          // ring slot reusable once read
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ===== frame threads: thread <-> TMEM lane <-> frame of the slot's tile =====
    const int g = (warp - 2) >> 2;             // tile slot
    const int wq = warp & 3;                   // TMEM lane quadrant this warp may access
    const int f = wq * 32 + lane;              // frame row within the tile
    const uint32_t tlane = tmem + g * kTmemSlot + (uint32_t(wq * 32) << 16);
    float* rs = reinterpret_cast<float*>(smem + SmemLayout::rs) + g * 128 * kM;
    float* sbmin = reinterpret_cast<float*>(smem + SmemLayout::bmin) + g * kMaxBatches * kM;
    uint32_t* wlist = reinterpret_cast<uint32_t*>(smem + SmemLayout::list) + (warp - 2) * kListMax;
    const uint32_t bar_a = ptx::smem_u32(&bars->a_ready[g]);
    // augmented K columns never change: cols 128,129 = 1 (pick up hi/lo of |c|^2), rest 0
    {
      uint32_t w8[8] = {pack_half2(1.f, 1.f), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      ptx::tmem_st8(tlane + kTmemA + 64, w8);
    }
    unsigned long long n_cert = 0, n_resc = 0, n_full = 0, n_all = 0;
    long long t_wait = 0, t_epi = 0, t_win = 0, t_upd = 0, t_load = 0;   // phase cycles (lane 0 of each warp)
    const long long t_begin = clock64();
    uint32_t acc_it = 0;
    for (int r = 0; r < rounds; ++r) {
      if (int64_t(r) * kGroups + g >= tcnt) break;     // idle slot in the last round
      const int64_t tile = tile0 + int64_t(r) * kGroups + g;
      const int64_t n = tile * kM + f;
      const bool valid = n < p.N;
      long long tc0 = clock64();
      // ---- load the latent, |x|^2 with the exact path's summation order, fp16 operand row ----
      float xx;
      {
        const int64_t xb = valid ? p.fa.base(n) : 0;
        float part[4] = {0.f, 0.f, 0.f, 0.f};
        #pragma unroll
        for (int h = 0; h < 4; ++h) {
          float v[32];
          #pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = valid ? __ldg(p.x + xb + int64_t(h * 32 + j) * p.fa.sxd) : 0.f;
          #pragma unroll
          for (int j = 0; j < 32; ++j) {
            rs[rs_idx(h * 32 + j, f)] = v[j];
            part[h] = fmaf(v[j], v[j], part[h]);
          }
          uint32_t w0[8], w1[8];
          #pragma unroll
          for (int j = 0; j < 8; ++j) { w0[j] = pack_half2(v[2 * j], v[2 * j + 1]); w1[j] = pack_half2(v[16 + 2 * j], v[17 + 2 * j]); }
          ptx::tmem_st8(tlane + kTmemA + h * 16, w0);
          ptx::tmem_st8(tlane + kTmemA + h * 16 + 8, w1);
        }
        xx = ((part[0] + part[1]) + part[2]) + part[3];
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_a);
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

        for (int c = 0; c < nchunks; ++c, ++acc_it) {
          const uint32_t buf = acc_it & 1, aph = (acc_it >> 1) & 1;
          ptx::mbar_wait(ptx::smem_u32(&bars->acc_full[g][buf]), aph);
          ptx::tc_fence_after();
          { const long long t = clock64(); t_wait += t - tc0; tc0 = t; }
          {
            uint32_t v0[32], v1[32];
            ptx::tmem_ld32(tlane + buf * kN, v0);
            ptx::tmem_ld32(tlane + buf * kN + 32, v1);
            ptx::tmem_ld_wait();
            // scores are in registers: hand the accumulator back before reducing them
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->acc_empty[g][buf]));
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
            sbmin[(c * 2) * kM + f] = b0;
            sbmin[(c * 2 + 1) * kM + f] = b1;
          }
          { const long long t = clock64(); t_epi += t - tc0; tc0 = t; }
        }

        // ---- winner: certified / re-scored / exact scan ----
        float m = inf_f();
        #pragma unroll
        for (int j = 0; j < 32; j += 2) m = ptx::fmin3(m, cm[j], cm[j + 1]);
        const float thr = m + delta;
        uint32_t cmask = 0, bmask = 0;
        #pragma unroll
        for (int j = 0; j < 32; ++j) cmask |= (cm[j] <= thr) ? (1u << j) : 0u;
        const int nb = nchunks * 2;
        for (int a = 0; a < nb; ++a) bmask |= (sbmin[a * kM + f] <= thr) ? (1u << a) : 0u;
        const int ncl = __popc(cmask), nba = __popc(bmask);
        const float* t32 = pv.tab32(st);
        const float* cn = pv.cnorm(st);
        int idx = 0;
        bool need_full = outl || cmask == 0u || bmask == 0u;       // masks are empty only for NaN scores
        const bool certified = !need_full && ncl == 1 && nba == 1;
        if (certified) { idx = (__ffs(bmask) - 1) * 32 + (__ffs(cmask) - 1); ++n_cert; }
        const bool need_resc = !need_full && !certified;
        const uint32_t resc_mask = __ballot_sync(0xffffffffu, need_resc);
        if (resc_mask != 0u) {
          // candidates = flagged batches x flagged classes.  All lanes' candidates go to one list
          // (ascending code order per frame) that the warp then scores cooperatively: lane l holds
          // dims 4l..4l+3 of the code row (coalesced 512-B read) and of the frame's residual.
          const int mine = need_resc ? ncl * nba : 0;
          int pre = mine;
          #pragma unroll
          for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, pre, off); if (lane >= off) pre += o; }
          const int total = __shfl_sync(0xffffffffu, pre, 31);
          if (total > kListMax) {
            if (need_resc) need_full = true;           // pathological tie width: exact scan instead
          } else {
            int pos = pre - mine;
            if (need_resc) {
              uint32_t bm2 = bmask;
              while (bm2) {
                const int a = __ffs(bm2) - 1; bm2 &= bm2 - 1;
                uint32_t cm2 = cmask;
                while (cm2) { const int j = __ffs(cm2) - 1; cm2 &= cm2 - 1; wlist[pos++] = (uint32_t(lane) << 16) | uint32_t(a * 32 + j); }
              }
            }
            __syncwarp();
            float best = inf_f(); int bi = 0x7fffffff;
            for (int e0 = 0; e0 < total; e0 += 4) {
              float part[4], xxo[4]; int code[4], own[4];
              #pragma unroll
              for (int u = 0; u < 4; ++u) {
                const uint32_t ent = wlist[min(e0 + u, total - 1)];
                own[u] = int(ent >> 16); code[u] = int(ent & 0xffffu);
                const float4 cv = __ldg(reinterpret_cast<const float4*>(t32 + size_t(code[u]) * D) + lane);
                const int fo = wq * 32 + own[u];
                float acc = rs[rs_idx(4 * lane + 0, fo)] * cv.x;
                acc = fmaf(rs[rs_idx(4 * lane + 1, fo)], cv.y, acc);
                acc = fmaf(rs[rs_idx(4 * lane + 2, fo)], cv.z, acc);
                acc = fmaf(rs[rs_idx(4 * lane + 3, fo)], cv.w, acc);
                part[u] = acc;
                xxo[u] = __shfl_sync(0xffffffffu, xx, own[u]);
              }
              #pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float dot = warp_sum(part[u]);
                const float dist = (xxo[u] - 2.f * dot) + __ldg(cn + code[u]);      // core_vq.py:183-187
                if (e0 + u < total && lane == own[u] && dist < best) { best = dist; bi = code[u]; }
              }
            }
            if (need_resc) { idx = bi == 0x7fffffff ? 0 : bi; ++n_resc; }
            __syncwarp();
          }
        }
        if (__any_sync(0xffffffffu, need_full)) {
          // frames outside the fp16 image's validity range, NaN, or absurd tie widths: exact fp32
          // scan of the whole table, all lanes in lockstep (code rows are warp-uniform loads)
          float best = inf_f(); int bi = 0x7fffffff;
          for (int k = 0; k < p.K; ++k) {
            const float4* row = reinterpret_cast<const float4*>(t32 + size_t(k) * D);
            float acc = 0.f;
            #pragma unroll 4
            for (int d4 = 0; d4 < 32; ++d4) {
              const float4 cv = __ldg(row + d4);
              acc = fmaf(rs[rs_idx(d4 * 4 + 0, f)], cv.x, acc);
              acc = fmaf(rs[rs_idx(d4 * 4 + 1, f)], cv.y, acc);
              acc = fmaf(rs[rs_idx(d4 * 4 + 2, f)], cv.z, acc);
              acc = fmaf(rs[rs_idx(d4 * 4 + 3, f)], cv.w, acc);
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
          #pragma unroll
          for (int h = 0; h < 4; ++h) {
            float4 cq[8];
            #pragma unroll
            for (int j = 0; j < 8; ++j) cq[j] = __ldg(row + h * 8 + j);
            uint32_t w0[8], w1[8];
            #pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float cv[4] = {cq[j].x, cq[j].y, cq[j].z, cq[j].w};
              float rn[4];
              #pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int d = h * 32 + j * 4 + i;
                const float rv = rs[rs_idx(d, f)];
                float qv = cv[i];
                if (p.ste) qv = rv + (qv - rv);               // core_vq.py:309
                rn[i] = rv - qv;                               // core_vq.py:364 / :348
                rs[rs_idx(d, f)] = rn[i];
                part[h] = fmaf(rn[i], rn[i], part[h]);
              }
              if (j < 4) { w0[2 * j] = pack_half2(rn[0], rn[1]); w0[2 * j + 1] = pack_half2(rn[2], rn[3]); }
              else       { w1[2 * (j - 4)] = pack_half2(rn[0], rn[1]); w1[2 * (j - 4) + 1] = pack_half2(rn[2], rn[3]); }
            }
            ptx::tmem_st8(tlane + kTmemA + h * 16, w0);
            ptx::tmem_st8(tlane + kTmemA + h * 16 + 8, w1);
          }
          xx = ((part[0] + part[1]) + part[2]) + part[3];
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        if (s + 1 < p.n_q) ptx::mbar_arrive(bar_a);
        if (valid) p.codes[int64_t(s) * p.N + n] = idx;
        if (p.sqerr != nullptr) {
          const float v = warp_sum(valid ? xx : 0.f);
          if (lane == 0) atomicAdd(&p.sqerr[s], (double)v);
        }
        { const long long t = clock64(); t_upd += t - tc0; tc0 = t; }
      }
      if (p.residual_out != nullptr && valid) {
        float* out = p.residual_out + n * D;
        #pragma unroll 4
        for (int d4 = 0; d4 < 32; ++d4)
          *reinterpret_cast<float4*>(out + d4 * 4) = make_float4(rs[rs_idx(d4 * 4, f)], rs[rs_idx(d4 * 4 + 1, f)],
                                                                 rs[rs_idx(d4 * 4 + 2, f)], rs[rs_idx(d4 * 4 + 3, f)]);
      }
    }
    // search statistics and phase cycles (evidence; see rvq_search_stats)
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
  RVQ_REQUIRE(tc_shape(a.K, a.D), "tc_encode: shape D=%d K=%d", a.D, a.K);
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
  const int64_t pairs = (ntiles + kGroups - 1) / kGroups;
  const unsigned grid = unsigned(pairs < sm_count ? pairs : sm_count);
  tc_encode_kernel<<<grid, kThreadsTc, SmemLayout::total, st>>>(p);
  RVQ_LAUNCH_CHECK("tc_encode_kernel");
  if (a.quantized != nullptr)
    return simt_quant_sum(a.pack, a.K, a.D, a.x, p.fa, N, a.T, a.stage0, a.n_q, a.codes, a.quantized, p.ste,
                          (a.flags & RVQ_FLAG_ACCUM_Q) ? 1 : 0, st);
  return RVQ_OK;
}

}  // namespace rvq

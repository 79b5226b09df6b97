// Fused multi-stage nearest-code search on the 5th-generation tensor cores (tcgen05 / TMEM).
//
// One persistent CTA per SM walks a contiguous range of 128-frame tiles.  The fp32 residual of every
// frame of the tile stays in shared memory across all n_q stages (core_vq.py:357-367 without the
// per-stage round trips through HBM).  Per stage the scores  S[f,k] = -2 r_f . c_k + |c_k|^2  of all K
// codes come from tcgen05.mma (fp16 operands, fp32 accumulation in tensor memory), 128 codes per MMA
// group (M=128, N=128, 9 K-steps of 16 -- the issue-rate floor measured by scripts/ubench_tc.cu):
//   A = fp16(r) [128 frames x 144] in shared memory: two 128B-swizzled K-major blocks of 64 dims that
//       the frame warps rewrite every stage, plus a constant block whose columns 128,129 = 1 pick up
//       the hi/lo halves of |c|^2;
//   B = fp16 image of the codebook, 128 codes x 144 per chunk (cols 0..127 = -2c, cols 128,129 = hi/lo
//       of |c|^2), streamed by the TMA engine (cp.async.bulk) from the pre-arranged pack into a ring.
// Warp roles (320 threads): 0..3 = score warps (thread = TMEM lane = frame), 4..7 = helper warps, 8 = TMA
// producer, 9..11 = MMA issuers (chunks round-robin; warp 9 also allocates TMEM).  A score warp reads the accumulators with
// tcgen05.ld and keeps, per frame, the minimum over every 32-code batch and over every residue class
// (code mod 32).  A code is within `delta` of the minimum iff its batch AND its class are; delta bounds
// the fp16 score error two-sidedly (rvq_common.cuh, StageMeta), so the exact fp32 winner is certified
// when exactly one batch and one class qualify.  Otherwise the candidates (flagged batches x flagged
// classes, usually 2..4 codes) are re-scored in fp32 with the reference's formula (core_vq.py:181-189,
// ties -> lowest index).  Frames outside the fp16 image's validity range take an exact fp32 scan.
// The update pass runs "lane = dimension": for each frame a quarter-warp reads the winner's (or the
// candidates') rows of the fp32 table (128-byte segments, all certified rows of a warp in flight at once),
// re-scores if needed, updates the residual exactly (core_vq.py:364 / :348, with the straight-through arithmetic of :309 in training)
// and writes the fp16 operand of the next stage.  Score warp q and helper warp q split the 32 frames
// of TMEM lane quadrant q for that pass.
#include "rvq_common.cuh"
#include "rvq_ptx.cuh"

namespace rvq {

namespace {

constexpr int kM = 128;                 // frames per tile (UMMA M, TMEM lanes)
constexpr int kN = kTcChunkCodes;       // 128 codes per MMA group (UMMA N)
constexpr int kRing = 3;                // B ring slots
constexpr int kKSteps = kTcKPad / 16;   // 9 UMMA K steps of 16
constexpr int kAccBufs = 4;             // accumulator buffers of kN TMEM columns
constexpr int kMaxChunks = 8;           // K <= 1024 on this path
constexpr int kThreadsTc = 12 * 32;        // 8 frame warps + warpgroup {TMA producer, 2 MMA issuers, idle}
constexpr int kBig = 5;                 // ncnt marker: more than 4 candidates (enumerate the masks)
constexpr int kFull = 6;                // ncnt marker: exact scan of the whole table

struct SmemLayout {
  static constexpr uint32_t a_sw = 0;                              // 2 x [128 rows][128 B], 128B swizzle
  static constexpr uint32_t a_aug = a_sw + 2 * 16384;              // [2 k-groups][128 rows][16 B], no swizzle
  static constexpr uint32_t b = a_aug + 4096;                      // ring of codebook chunks
  static constexpr uint32_t rs = b + kRing * kTcChunkBytes;        // fp32 residual [128 f][128 d], chunk-swizzled
  static constexpr uint32_t cand = rs + kM * 128 * 4;              // int4 [128]: candidate codes (-1 = none)
  static constexpr uint32_t ncnt = cand + kM * 16;                 // int [128]
  static constexpr uint32_t cmask = ncnt + kM * 4;                 // u32 [128] flagged classes
  static constexpr uint32_t bmask = cmask + kM * 4;                // u32 [128] flagged batches
  static constexpr uint32_t xpart = bmask + kM * 4;                // float [4][128] partial |x|^2
  static constexpr uint32_t dr2 = xpart + 4 * kM * 4;              // float [128]: |r - fp16(r)|^2 of every frame's current operand
  static constexpr uint32_t slowq = dr2 + kM * 4;            // u8 [2][128]: frames with 2..4 listed candidates (per stage parity)
  static constexpr uint32_t wideq = slowq + 2 * kM;                // u8 [2][128]: frames with a wide candidate set
  static constexpr uint32_t qcnt = wideq + 2 * kM;                 // int [2][2]: queue lengths {slow, wide} per stage parity
  static constexpr uint32_t bars = qcnt + 16;
  static constexpr uint32_t total = bars + 256;
};
struct Bars {
  uint64_t full[kRing], empty[kRing], acc_full[kAccBufs], acc_empty[kAccBufs], a_ready;
  long long t0;          // kernel start (debug trace)
  uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 256, "barrier block");
static_assert(SmemLayout::total <= 227 * 1024, "shared memory budget");
static_assert(kKSteps == 9 && kN == 128, "operand geometry");

// debug timeline of CTA 0 (first tile, first kTraceStages stages): g_trace[stage][warp][event] = cycles since kernel start
constexpr int kTraceStages = 4, kTraceEv = 16;
__device__ long long g_trace[kTraceStages * 11 * kTraceEv];
#ifdef RVQ_TC_TRACE
#define RVQ_TRACE(stage, ev) do { if (blockIdx.x == 0 && t_tile == 0 && (stage) < kTraceStages && lane == 0) \
    g_trace[((stage) * 11 + warp) * kTraceEv + (ev)] = clock64() - t_kernel0; } while (0)
#else
#define RVQ_TRACE(stage, ev) do { (void)t_tile; } while (0)
#endif
// phase timers of the score warps / MMA issuer (rvq_search_stats); they live in registers only when enabled
#ifdef RVQ_TC_TIMERS
#define RVQ_TICK(acc) do { const unsigned tt_ = (unsigned)clock(); acc += tt_ - tc0; tc0 = tt_; } while (0)
#else
#define RVQ_TICK(acc) do { } while (0)
#endif

struct TcParams {
  const unsigned char* pack; int K;
  const float* x; FrameAddr fa; int64_t N;
  int stage0, n_q;
  int64_t* codes; float* residual_out; double* sqerr;
  int ste;
  unsigned long long* counters;
};

__device__ __forceinline__ float inf_f() { return __int_as_float(0x7f800000); }
// residual element group: 16-byte chunk ch (dims 4ch..4ch+3) of frame f, XOR-swizzled so that both
// "lanes = consecutive chunks of one frame" and "lanes = consecutive frames, one chunk" spread over banks
__device__ __forceinline__ int rs_off(int f, int ch) { return f * 128 + ((ch ^ (f & 31)) << 2); }
// byte offset inside a 128B-swizzled K block of dims 4g..4g+3 (g = 0..15) of row f
__device__ __forceinline__ uint32_t asw_off(int f, int g) { return uint32_t(f * 128 + ((((g >> 1) ^ (f & 7))) << 4) + ((g & 1) << 3)); }
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float warp_sum(float v) {
  #pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
__device__ __forceinline__ float half_warp_sum(float v) {
  #pragma unroll
  for (int off = 8; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); return fmaf(a.w, b.w, acc);
}
__device__ __forceinline__ float min32(const uint32_t (&v)[32]) {
  float t[11];
  #pragma unroll
  for (int j = 0; j < 10; ++j) t[j] = ptx::fmin3(__uint_as_float(v[3 * j]), __uint_as_float(v[3 * j + 1]), __uint_as_float(v[3 * j + 2]));
  t[10] = fminf(__uint_as_float(v[30]), __uint_as_float(v[31]));
  const float a = ptx::fmin3(t[0], t[1], t[2]), b = ptx::fmin3(t[3], t[4], t[5]), c = ptx::fmin3(t[6], t[7], t[8]);
  return ptx::fmin3(ptx::fmin3(a, b, c), t[9], t[10]);
}

// tiles [start, start+cnt) of this CTA
__device__ __forceinline__ void cta_range(int ntiles, int& start, int& cnt) {
  const int base = ntiles / int(gridDim.x), rem = ntiles % int(gridDim.x);
  const int b = blockIdx.x;
  start = b * base + (b < rem ? b : rem);
  cnt = base + (b < rem ? 1 : 0);
}

// ---- update pass -------------------------------------------------------------------------------------
// torch's CPU argmax (core_vq.py:188) propagates NaN: the first NaN distance wins; otherwise the smallest
// distance, lowest index on ties.  (best, bcode) starts as (+inf, 0x7fffffff).
__device__ __forceinline__ bool nan_aware_better(float dist, int code, float best, int bcode) {
  if (dist != dist) return best == best || code < bcode;
  return best == best && (dist < best || (dist == best && code < bcode));
}

// Frames whose candidate set does not fit an Item (more than 4 candidates, or an exact scan): the whole
// warp scores the set, one candidate per lane, and rewrites the frame's entry as a certified winner.
__device__ __noinline__ void resolve_big(unsigned char* smem, int f, int lane, int K, int rot, const float* __restrict__ t32,
                                         const float* __restrict__ cn) {
  const float* rs = reinterpret_cast<const float*>(smem + SmemLayout::rs);
  int* ncnt = reinterpret_cast<int*>(smem + SmemLayout::ncnt);
  const bool full = ncnt[f] == kFull;
  const uint32_t cm = full ? 0xffffffffu : *reinterpret_cast<const uint32_t*>(smem + SmemLayout::cmask + f * 4);
  const uint32_t bm = full ? 0xffffffffu : *reinterpret_cast<const uint32_t*>(smem + SmemLayout::bmask + f * 4);
  const int nc = __popc(cm);
  const int total = full ? K : nc * __popc(bm);
  // |r|^2: lane l owns chunk l
  const float4 rl = *reinterpret_cast<const float4*>(rs + rs_off(f, lane));
  const float rr = warp_sum(dot4(rl, rl, 0.f));
  float best = inf_f(); int bcode = 0x7fffffff;
  for (int t = lane; t < total; t += 32) {
    int code;
    if (full) code = t;
    else {
      const int a = int(__fns(bm, 0, t / nc + 1));       // batch in processing order -> actual batch
      int pc = (a >> 2) + rot; pc = pc < (K >> 7) ? pc : pc - (K >> 7);
      code = pc * 128 + (a & 3) * 32 + int(__fns(cm, 0, t % nc + 1));
    }
    if (code >= K) continue;
    const float4* rp = reinterpret_cast<const float4*>(t32 + size_t(code) * 128);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    #pragma unroll 2
    for (int ch = 0; ch < 32; ch += 4) {
      a0 = dot4(*reinterpret_cast<const float4*>(rs + rs_off(f, ch + 0)), __ldg(rp + ch + 0), a0);
      a1 = dot4(*reinterpret_cast<const float4*>(rs + rs_off(f, ch + 1)), __ldg(rp + ch + 1), a1);
      a2 = dot4(*reinterpret_cast<const float4*>(rs + rs_off(f, ch + 2)), __ldg(rp + ch + 2), a2);
      a3 = dot4(*reinterpret_cast<const float4*>(rs + rs_off(f, ch + 3)), __ldg(rp + ch + 3), a3);
    }
    const float dot = (a0 + a1) + (a2 + a3);
    const float dist = (rr - 2.f * dot) + __ldg(cn + code);          // core_vq.py:183-187
    if (nan_aware_better(dist, code, best, bcode)) { best = dist; bcode = code; }
  }
  #pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, off);
    const int oc = __shfl_xor_sync(0xffffffffu, bcode, off);
    if (oc != 0x7fffffff && nan_aware_better(ob, oc, best, bcode)) { best = ob; bcode = oc; }
  }
  __syncwarp();
  if (lane == 0) {
    *reinterpret_cast<int4*>(smem + SmemLayout::cand + f * 16) = make_int4(bcode == 0x7fffffff ? 0 : bcode, -1, -1, -1);
    ncnt[f] = 1;
  }
  __syncwarp();
}

// ---- update pass: one frame per QUARTER-warp; lane j (0..7) of the quarter owns the 16-byte chunks
// j, 8+j, 16+j, 24+j of the frame's 512-byte row (dims 4c..4c+3 of chunk c), so every row access of the
// quarter is one contiguous 128-byte segment ------------------------------------------------------------
struct Row4 { float4 v[4]; };

__device__ __forceinline__ float quarter_sum(float v) {
  #pragma unroll
  for (int off = 4; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
__device__ __forceinline__ Row4 load_row(const float* __restrict__ t32, int code, int j) {
  const float4* rp = reinterpret_cast<const float4*>(t32 + size_t(code) * 128);
  Row4 r;
  #pragma unroll
  for (int i = 0; i < 4; ++i) r.v[i] = __ldg(rp + 8 * i + j);
  return r;
}
__device__ __forceinline__ Row4 load_res(const float* rs, int f, int j) {
  Row4 r;
  #pragma unroll
  for (int i = 0; i < 4; ++i) r.v[i] = *reinterpret_cast<const float4*>(rs + rs_off(f, 8 * i + j));
  return r;
}
__device__ __forceinline__ float dot_row(const Row4& a, const Row4& b) {
  return (dot4(a.v[0], b.v[0], 0.f) + dot4(a.v[1], b.v[1], 0.f)) + (dot4(a.v[2], b.v[2], 0.f) + dot4(a.v[3], b.v[3], 0.f));
}
// fp16 operand of the next stage: chunk c = dims 4c..4c+3 -> K block c/16, 16-byte group (c%16)/2, half c%2.
// Also records the exact squared rounding residue |r - fp16(r)|^2 of the frame (it enters the score-error
// margin of the next stage).  Called by whole quarter-warps (8 converged lanes).
__device__ __forceinline__ void store_operand(unsigned char* smem, int f, int j, int qq, const Row4& n) {
  float e2 = 0.f;
  #pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = 8 * i + j;
    const __half2 h0 = __floats2half2_rn(n.v[i].x, n.v[i].y), h1 = __floats2half2_rn(n.v[i].z, n.v[i].w);
    *reinterpret_cast<uint2*>(smem + SmemLayout::a_sw + (c >> 4) * 16384 + asw_off(f, c & 15)) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    const float2 b0 = __half22float2(h0), b1 = __half22float2(h1);
    const float ex = n.v[i].x - b0.x, ey = n.v[i].y - b0.y, ez = n.v[i].z - b1.x, ew = n.v[i].w - b1.y;
    e2 = fmaf(ex, ex, e2); e2 = fmaf(ey, ey, e2); e2 = fmaf(ez, ez, e2); e2 = fmaf(ew, ew, e2);
  }
  const uint32_t qmask = 0xffu << (8 * qq);
  #pragma unroll
  for (int off = 4; off > 0; off >>= 1) e2 += __shfl_xor_sync(qmask, e2, off);
  if (j == 0) reinterpret_cast<float*>(smem + SmemLayout::dr2)[f] = e2;
}
// exact fp32 r <- r - q (core_vq.py:364 / :348; straight-through arithmetic of :309 in training), fp16
// operand of the next stage, code store, squared-error partial
template <bool TRAIN>
__device__ __forceinline__ void apply_row(const TcParams& p, unsigned char* smem, int f, int j, int qq, const Row4& r, const Row4& qrow,
                                          int code, int s, int64_t tile_n0, float& sq_acc) {
  float* rs = reinterpret_cast<float*>(smem + SmemLayout::rs);
  Row4 n;
  #pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 q = qrow.v[i];
    const float4 rv = r.v[i];
    if (TRAIN && p.ste) { q.x = rv.x + (q.x - rv.x); q.y = rv.y + (q.y - rv.y); q.z = rv.z + (q.z - rv.z); q.w = rv.w + (q.w - rv.w); }
    n.v[i] = make_float4(rv.x - q.x, rv.y - q.y, rv.z - q.z, rv.w - q.w);
    *reinterpret_cast<float4*>(rs + rs_off(f, 8 * i + j)) = n.v[i];
  }
  const int64_t nfr = tile_n0 + f;
  if (nfr < p.N) {
    if (j == 0) p.codes[int64_t(s) * p.N + nfr] = code;
    if (TRAIN && p.sqerr != nullptr) sq_acc += dot_row(n, n);
  }
  store_operand(smem, f, j, qq, n);
}

// Exact fp32 re-score of up to 4 candidate codes (-1 = none) of the frame whose residual this quarter-warp
// holds: load4 puts the four rows in flight, score4 computes the distances with the reference's formula
// (core_vq.py:183-187) and keeps the best (lowest code on ties) and its row.  Every lane of the warp must
// call score4 (shuffles).
struct Cand4 { int c[4]; Row4 w[4]; float nrm[4]; };
__device__ __forceinline__ void load4(Cand4& k, int j, const float* __restrict__ t32, const float* __restrict__ cn) {
  #pragma unroll
  for (int u = 0; u < 4; ++u) {
    k.nrm[u] = 0.f;
    if (k.c[u] >= 0) { k.w[u] = load_row(t32, k.c[u], j); k.nrm[u] = __ldg(cn + k.c[u]); }
  }
}
__device__ __forceinline__ void score4(const Cand4& k, const Row4& r, float rr, float& best, int& bcode, Row4& brow) {
  float d[4];
  #pragma unroll
  for (int u = 0; u < 4; ++u) d[u] = dot_row(r, k.w[u]);
  #pragma unroll
  for (int off = 4; off > 0; off >>= 1) {
    #pragma unroll
    for (int u = 0; u < 4; ++u) d[u] += __shfl_xor_sync(0xffffffffu, d[u], off);
  }
  #pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float e = (rr - 2.f * d[u]) + k.nrm[u];
    if (k.c[u] >= 0 && (e < best || (e == best && k.c[u] < bcode))) { best = e; bcode = k.c[u]; brow = k.w[u]; }
  }
}

// A frame whose flagged batches x flagged classes give more than 4 candidates: the whole warp works on it,
// 16 candidates per step (4 per quarter-warp); the quarter that holds the winner's row updates the frame.
template <bool TRAIN>
__device__ __noinline__ float resolve_wide(const TcParams& p, unsigned char* smem, int f, int lane, int s, int rot, int nchunks,
                                              int64_t tile_n0, const float* __restrict__ t32, const float* __restrict__ cn) {
  float sq_acc = 0.f;
  const float* rs = reinterpret_cast<const float*>(smem + SmemLayout::rs);
  const int qq = lane >> 3, j = lane & 7;
  const uint32_t cm = *reinterpret_cast<const uint32_t*>(smem + SmemLayout::cmask + f * 4);
  const uint32_t bm = *reinterpret_cast<const uint32_t*>(smem + SmemLayout::bmask + f * 4);
  const int nc = __popc(cm);
  const Row4 r = load_res(rs, f, j);
  const float rr = quarter_sum(dot_row(r, r));
  float best = inf_f(); int bcode = 0x7fffffff; Row4 brow = r;
  // quarter qq takes the flagged batches number qq, qq + 4, ...; within a batch the flagged classes four at a time
  uint32_t bmq = bm;
  for (int i = 0; i < qq; ++i) bmq &= bmq - 1;
  const int nb = __popc(bm);
  #pragma unroll 1
  for (int ob = 0; ob < nb; ob += 4) {
    const int a = bmq ? __ffs(bmq) - 1 : -1;
    #pragma unroll
    for (int i = 0; i < 4; ++i) bmq &= bmq - 1;
    int base = 0;
    if (a >= 0) { int pc = (a >> 2) + rot; pc = pc < nchunks ? pc : pc - nchunks; base = pc * 128 + (a & 3) * 32; }   // processing order -> code
    uint32_t cmq = cm;
    #pragma unroll 1
    for (int oc = 0; oc < nc; oc += 4) {
      Cand4 k;
      #pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int jj = cmq ? __ffs(cmq) - 1 : -1;
        cmq &= cmq - 1;
        k.c[u] = (a >= 0 && jj >= 0) ? base + jj : -1;
      }
      load4(k, j, t32, cn);
      score4(k, r, rr, best, bcode, brow);
    }
  }
  // best over the four quarters (candidate codes are distinct, so the winner's quarter is unique)
  float wb = best; int wc = bcode;
  #pragma unroll
  for (int off = 8; off <= 16; off <<= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, wb, off);
    const int oc = __shfl_xor_sync(0xffffffffu, wc, off);
    if (ob < wb || (ob == wb && oc < wc)) { wb = ob; wc = oc; }
  }
  bool mine = bcode == wc;
  if (wc == 0x7fffffff) {                                  // NaN distances only: lowest candidate, like an exact scan would
    mine = qq == 0;
    if (mine) { const int c0 = (__ffs(bm) - 1); int pc = (c0 >> 2) + rot; pc = pc < nchunks ? pc : pc - nchunks;
                bcode = pc * 128 + (c0 & 3) * 32 + (__ffs(cm) - 1); brow = load_row(t32, bcode, j); }
  }
  if (mine) apply_row<TRAIN>(p, smem, f, j, qq, r, brow, bcode, s, tile_n0, sq_acc);
  return sq_acc;
}

// FIRST: the residual rows were just loaded from x; only the fp16 operand is produced.
// Otherwise every quarter-warp walks a short list of items: the certified frames among the 4 it owns (frame
// f16 + 4 qq + k: the four frames of a step then sit in different bank groups of the operand tile), then its
// share of the stage's re-score queue (items gq, gq + 32, ... -- the frames that need a re-score are spread over
// all 32 quarter-warps of the CTA, so no warp is left with several of them).  ONE rolled loop body serves both
// kinds; the first candidate row of the next item is fetched while the current one is processed.  Keeping this
// code small matters more than hiding every latency: a stage's hot code has to fit the SM's instruction cache
// (ncu showed a 67 % icc hit rate and a saturated GPC instruction cache with the unrolled version).
template <bool FIRST, bool TRAIN>
__device__ __forceinline__ void update_pass(const TcParams& p, unsigned char* smem, int q, int h, int lane, int s, int rot,
                                            int nchunks, int64_t tile_n0, const float* __restrict__ t32,
                                            const float* __restrict__ cn, float& sq_acc, int qpar) {
  float* rs = reinterpret_cast<float*>(smem + SmemLayout::rs);
  const int qq = lane >> 3, j = lane & 7;
  const int f16 = q * 32 + h * 16;
  if (FIRST) {
    #pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      const int f = f16 + 4 * qq + k;
      store_operand(smem, f, j, qq, load_res(rs, f, j));
    }
    return;
  }
  const int* qc = reinterpret_cast<const int*>(smem + SmemLayout::qcnt) + qpar * 2;
  const int nslow = qc[0], nwide = qc[1];
  const unsigned char* slowq = smem + SmemLayout::slowq + qpar * kM;
  const unsigned char* wideq = smem + SmemLayout::wideq + qpar * kM;
  if (threadIdx.x == 0) { int* nx = reinterpret_cast<int*>(smem + SmemLayout::qcnt) + (qpar ^ 1) * 2; nx[0] = 0; nx[1] = 0; }
  const int nv = lane < 16 ? *reinterpret_cast<const int*>(smem + SmemLayout::ncnt + (f16 + lane) * 4) : 1;
  const uint32_t slow = __ballot_sync(0xffffffffu, nv > 1);     // bit i = frame f16+i is in one of the queues
  const int gq = (h * 4 + q) * 4 + qq;                          // quarter-warp number in the CTA, 0..31
  const int nitems = 4 + ((nslow + 31) >> 5);
  // item it -> frame (or -1) and its candidate list
  auto fetch = [&](int it, int& f, int4& cd) {
    if (it < 4) { const int fi = 4 * qq + it; f = ((slow >> fi) & 1u) ? -1 : f16 + fi; }
    else { const int qi = (it - 4) * 32 + gq; f = qi < nslow ? int(slowq[qi]) : -1; }
    cd = make_int4(-1, -1, -1, -1);
    if (f >= 0) cd = *reinterpret_cast<const int4*>(smem + SmemLayout::cand + f * 16);
  };
  int fn; int4 cdn; Row4 rown;
  fetch(0, fn, cdn);
  if (fn >= 0) rown = load_row(t32, cdn.x, j);
  // wide candidate sets (their latency overlaps the first row in flight): one frame per warp at a time
  #pragma unroll 1
  for (int i = h * 4 + q; i < nwide; i += 8) sq_acc += resolve_wide<TRAIN>(p, smem, wideq[i], lane, s, rot, nchunks, tile_n0, t32, cn);
  #pragma unroll 1
  for (int it = 0; it < nitems; ++it) {
    const int f = fn; const int4 cd = cdn;
    Row4 row = rown;
    if (it + 1 < nitems) {
      fetch(it + 1, fn, cdn);
      if (fn >= 0) rown = load_row(t32, cdn.x, j);
    }
    Row4 r;
    #pragma unroll
    for (int i = 0; i < 4; ++i) r.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f >= 0) r = load_res(rs, f, j);
    int code = cd.x;
    if (__any_sync(0xffffffffu, cd.y >= 0)) {                    // some quarter has 2..4 candidates: exact fp32 re-score
      const bool multi = cd.y >= 0;
      Row4 w1 = row, w2 = row, w3 = row;
      float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
      if (multi) n0 = __ldg(cn + cd.x);
      if (cd.y >= 0) { w1 = load_row(t32, cd.y, j); n1 = __ldg(cn + cd.y); }
      if (cd.z >= 0) { w2 = load_row(t32, cd.z, j); n2 = __ldg(cn + cd.z); }
      if (cd.w >= 0) { w3 = load_row(t32, cd.w, j); n3 = __ldg(cn + cd.w); }
      float d0 = dot_row(r, row), d1 = dot_row(r, w1), d2 = dot_row(r, w2), d3 = dot_row(r, w3), rr = dot_row(r, r);
      #pragma unroll
      for (int off = 4; off > 0; off >>= 1) {
        d0 += __shfl_xor_sync(0xffffffffu, d0, off); d1 += __shfl_xor_sync(0xffffffffu, d1, off);
        d2 += __shfl_xor_sync(0xffffffffu, d2, off); d3 += __shfl_xor_sync(0xffffffffu, d3, off);
        rr += __shfl_xor_sync(0xffffffffu, rr, off);
      }
      if (multi) {                                               // core_vq.py:183-187, lowest index on ties
        float best = (rr - 2.f * d0) + n0;
        if (!(best == best)) best = inf_f();                     // NaN distances: keep the first candidate unless a finite one exists
        const float e1 = (rr - 2.f * d1) + n1, e2 = (rr - 2.f * d2) + n2, e3 = (rr - 2.f * d3) + n3;
        if (e1 < best || (e1 == best && cd.y < code)) { best = e1; code = cd.y; row = w1; }
        if (cd.z >= 0 && (e2 < best || (e2 == best && cd.z < code))) { best = e2; code = cd.z; row = w2; }
        if (cd.w >= 0 && (e3 < best || (e3 == best && cd.w < code))) { best = e3; code = cd.w; row = w3; }
      }
    }
    if (f >= 0) apply_row<TRAIN>(p, smem, f, j, qq, r, row, code, s, tile_n0, sq_acc);
  }
}

}  // namespace

template <bool TRAIN>
__global__ void __launch_bounds__(kThreadsTc, 1) tc_encode_kernel(const TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  Bars* bars = reinterpret_cast<Bars*>(smem + SmemLayout::bars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PackView pv(p.pack, p.K, 128);
  const int nchunks = p.K / kN;
  // every CTA walks the chunks of a stage in its own rotation, so that the 148 SMs (which run the same stage
  // at about the same time) do not all pull the same lines out of the same L2 slices at once
  const int rot = int(blockIdx.x % unsigned(nchunks));
  const int ntiles = int((p.N + kM - 1) / kM);       // N < 2^31 frames per call (checked by rvq_encode)
  int tile0, tcnt;
  cta_range(ntiles, tile0, tcnt);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) { ptx::mbar_init(ptx::smem_u32(&bars->full[i]), 1); ptx::mbar_init(ptx::smem_u32(&bars->empty[i]), 1); }
    for (int i = 0; i < kAccBufs; ++i) { ptx::mbar_init(ptx::smem_u32(&bars->acc_full[i]), 1); ptx::mbar_init(ptx::smem_u32(&bars->acc_empty[i]), 4); }
    ptx::mbar_init(ptx::smem_u32(&bars->a_ready), 8);
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < 4) reinterpret_cast<int*>(smem + SmemLayout::qcnt)[threadIdx.x] = 0;
  // constant augmented K block of A: k-group 0 = (1, 1, 0, ...) picks up hi/lo of |c|^2, k-group 1 = 0
  for (int i = threadIdx.x; i < 4096 / 16; i += blockDim.x)
    *reinterpret_cast<uint4*>(smem + SmemLayout::a_aug + i * 16) = make_uint4(i < 128 ? pack_half2(1.f, 1.f) : 0u, 0u, 0u, 0u);
  if (warp == 9) {
    ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  if (threadIdx.x == 0) bars->t0 = clock64();
  __syncthreads();
  const long long t_kernel0 = bars->t0;

  // register budget: the two frame warpgroups take what the producer / issuer warpgroup gives up
  if (warp >= 8) {
  ptx::reg_dec<64>();
  if (warp == 8) {
    // ===== TMA producer: the chunk stream (stage-major) of every tile =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = 0; t < tcnt; ++t) {
        for (int s = 0; s < p.n_q; ++s) {
          const unsigned char* img = pv.tc(p.stage0 + s);
#ifdef RVQ_TC_TMA_AFTER_A
          ptx::mbar_wait(ptx::smem_u32(&bars->a_ready), uint32_t(t * p.n_q + s) & 1);   // experiment: no prefetch during the update pass
#endif
          for (int c = 0; c < nchunks; ++c, ++it) {
            const uint32_t slot = it % kRing, ph = (it / kRing) & 1;
            ptx::mbar_wait(ptx::smem_u32(&bars->empty[slot]), ph ^ 1);
            const uint32_t fb = ptx::smem_u32(&bars->full[slot]);
            ptx::mbar_expect_tx(fb, kTcChunkBytes);
            const int pc = c + rot < nchunks ? c + rot : c + rot - nchunks;
            ptx::bulk_g2s(sbase + SmemLayout::b + slot * kTcChunkBytes, img + size_t(pc) * kTcChunkBytes, kTcChunkBytes, fb);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ===== MMA issuers: descriptors hoisted, 9 MMAs + 2 commits per 128-code chunk.  The issue path of one
    // thread (two barrier waits of ~100 cycles each + the scalar code around every tcgen05.mma) is longer than
    // the 576 tensor cycles of a chunk, so three warps take the chunks round-robin. =====
    const uint32_t who = warp - 9;
    const uint32_t stride = nchunks >= 3 ? 3u : (nchunks >= 2 ? 2u : 1u);
    if (lane == 0 && who < stride) {
      constexpr uint32_t idesc = ptx::umma_idesc_f16_f32(kM, kN);
      uint64_t ad[kKSteps];
      #pragma unroll
      for (int k = 0; k < 8; ++k) ad[k] = ptx::umma_desc_kmajor_sw128(sbase + SmemLayout::a_sw + (k >> 2) * 16384 + (k & 3) * 32);
      ad[8] = ptx::umma_desc_kmajor_noswz(sbase + SmemLayout::a_aug, 2048, 128);
      const uint64_t bd0 = ptx::umma_desc_kmajor_noswz(sbase + SmemLayout::b, kTcLBO, kTcSBO);
      const uint32_t total = uint32_t(tcnt) * uint32_t(p.n_q) * uint32_t(nchunks);
      uint32_t seen = 0xffffffffu;          // last (tile, stage) index whose operand this thread waited for
      for (uint32_t it = who; it < total; it += stride) {
        const uint32_t ar = it / uint32_t(nchunks);
        if (ar != seen) {
          ptx::mbar_wait(ptx::smem_u32(&bars->a_ready), ar & 1);                 // fp16 residual operand written
          seen = ar;
          { const int t_tile = int(ar / uint32_t(p.n_q)); RVQ_TRACE(int(ar % uint32_t(p.n_q)), 0); }
        }
        const uint32_t slot = it % kRing, buf = it % kAccBufs;
        ptx::mbar_wait(ptx::smem_u32(&bars->full[slot]), (it / kRing) & 1);              // codebook chunk landed
        ptx::mbar_wait(ptx::smem_u32(&bars->acc_empty[buf]), ((it / kAccBufs) & 1) ^ 1); // accumulator drained
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem + buf * kN;
        const uint64_t b0 = bd0 + uint64_t((slot * kTcChunkBytes) >> 4);
        #pragma unroll
        for (int k = 0; k < kKSteps; ++k)
          ptx::umma_f16_ss(d_tmem, ad[k], b0 + uint64_t((k * 2 * kTcLBO) >> 4), idesc, k > 0 ? 1u : 0u);
        ptx::umma_commit(ptx::smem_u32(&bars->acc_full[buf]));     // scores ready for the score warps
        ptx::umma_commit(ptx::smem_u32(&bars->empty[slot]));       // ring slot reusable once read
        { const int t_tile = int(ar / uint32_t(p.n_q)); RVQ_TRACE(int(ar % uint32_t(p.n_q)), 1 + int(it % uint32_t(nchunks))); }
      }
    }
    __syncwarp();
  }
  } else {
    ptx::reg_inc<216>();
    // ===== frame warps =====
    const int q = warp & 3;                    // TMEM lane quadrant = frames 32q..32q+31 of the tile
    const int h = warp >= 4 ? 1 : 0;           // 0 = score warp, 1 = helper warp
    const int f = q * 32 + lane;               // tile load / score mapping: thread <-> frame
    const uint32_t pair_bar = 1 + q;
    const uint32_t tlane = tmem + (uint32_t(q * 32) << 16);
    float* rs = reinterpret_cast<float*>(smem + SmemLayout::rs);
    float* xpart = reinterpret_cast<float*>(smem + SmemLayout::xpart);
    const uint32_t bar_a = ptx::smem_u32(&bars->a_ready);
    uint32_t n_cert = 0, n_resc = 0, n_full = 0;                        // search statistics (rvq_search_stats)
#ifdef RVQ_TC_TIMERS
    uint32_t t_wait = 0, t_epi = 0, t_win = 0, t_upd = 0, t_load = 0;  // phase cycles (lane 0 of each score warp)
    const long long t_begin = clock64();
#endif
    uint32_t acc_it = 0;
    for (int t = 0; t < tcnt; ++t) {
      const int64_t tile_n0 = int64_t(tile0 + t) * kM;
      const int t_tile = int(t);
#ifdef RVQ_TC_TIMERS
      unsigned tc0 = (unsigned)clock();
#endif
      // ---- load the latent tile: this warp takes dims 64h..64h+63 of its quadrant's 32 frames ----
      {
        const int64_t n = tile_n0 + f;
        const bool valid = n < p.N;
        const int64_t xb = valid ? p.fa.base(n) : 0;
        #pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          float v[32];
          const int d0 = h * 64 + hb * 32;
          #pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = valid ? __ldg(p.x + xb + int64_t(d0 + j) * p.fa.sxd) : 0.f;
          float part = 0.f;
          #pragma unroll
          for (int j = 0; j < 32; j += 4) {
            *reinterpret_cast<float4*>(rs + rs_off(f, (d0 + j) >> 2)) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            part = fmaf(v[j], v[j], part); part = fmaf(v[j + 1], v[j + 1], part);
            part = fmaf(v[j + 2], v[j + 2], part); part = fmaf(v[j + 3], v[j + 3], part);
          }
          xpart[(h * 2 + hb) * kM + f] = part;
        }
      }
      ptx::named_bar_sync(pair_bar, 64);
      float xx = ((xpart[f] + xpart[kM + f]) + xpart[2 * kM + f]) + xpart[3 * kM + f];   // exact path's order
      float sq_dummy = 0.f;
      update_pass<true, TRAIN>(p, smem, q, h, lane, 0, rot, nchunks, tile_n0, nullptr, nullptr, sq_dummy, 0);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_a);
      RVQ_TICK(t_load);

      for (int s = 0; s < p.n_q; ++s) {
        const int st = p.stage0 + s;
        const float* t32 = pv.tab32(st);
        const float* cn = pv.cnorm(st);
        const int qpar = int((t * p.n_q + s) & 1);          // parity of the re-score queues of this (tile, stage)
        RVQ_TRACE(s, 0);
        if (h == 0) {
          // ---- scores: per-class and per-batch minima of the K approximate scores of this frame ----
          const StageMeta* meta = pv.meta(st);
          const float xnorm = sqrtf(xx);
          const bool outl = !(xnorm < meta->xlimit);      // also true for NaN
          float cm[32], bmin[32];
          #pragma unroll
          for (int j = 0; j < 32; ++j) { cm[j] = inf_f(); bmin[j] = inf_f(); }
          // one rolled iteration per 128-code chunk (the hot loops of a stage must stay inside the instruction cache);
          // bmin is a shift register: after the loop the a-th batch in processing order sits at 32 - 4*nchunks + a
          #pragma unroll 1
          for (int c = 0; c < nchunks; ++c) {
            const uint32_t buf = acc_it % kAccBufs, aph = (acc_it / kAccBufs) & 1;
            ++acc_it;
            ptx::mbar_wait(ptx::smem_u32(&bars->acc_full[buf]), aph);
            ptx::tc_fence_after();
            RVQ_TICK(t_wait);
            RVQ_TRACE(s, 1 + c);
            uint32_t v0[32], v1[32];
            ptx::tmem_ld32(tlane + buf * kN, v0);
            ptx::tmem_ld32(tlane + buf * kN + 32, v1);
            #pragma unroll
            for (int j = 0; j < 28; ++j) bmin[j] = bmin[j + 4];
            ptx::tmem_ld_wait();
            #pragma unroll
            for (int j = 0; j < 32; ++j) cm[j] = ptx::fmin3(cm[j], __uint_as_float(v0[j]), __uint_as_float(v1[j]));
            bmin[28] = min32(v0);
            bmin[29] = min32(v1);
            ptx::tmem_ld32(tlane + buf * kN + 64, v0);
            ptx::tmem_ld32(tlane + buf * kN + 96, v1);
            ptx::tmem_ld_wait();
            // scores are in registers: hand the accumulator back before reducing them
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->acc_empty[buf]));
            #pragma unroll
            for (int j = 0; j < 32; ++j) cm[j] = ptx::fmin3(cm[j], __uint_as_float(v0[j]), __uint_as_float(v1[j]));
            bmin[30] = min32(v0);
            bmin[31] = min32(v1);
            RVQ_TICK(t_epi);
          }
          RVQ_TRACE(s, 9);
          // the rounding residue of this frame's operand was written by whichever warp updated the frame; the scores
          // above could only exist after every warp had finished that update
          const float drn = sqrtf(reinterpret_cast<const float*>(smem + SmemLayout::dr2)[f]) * 1.001f;
          const float delta = meta->margin_coef * xnorm + meta->margin_dr * drn + meta->margin_abs;
          // ---- candidates: certified winner / up to 4 codes to re-score / mask enumeration / exact scan ----
          float m4[4];
          #pragma unroll
          for (int j = 0; j < 4; ++j) {
            m4[j] = ptx::fmin3(cm[8 * j], cm[8 * j + 1], cm[8 * j + 2]);
            m4[j] = ptx::fmin3(m4[j], cm[8 * j + 3], cm[8 * j + 4]);
            m4[j] = ptx::fmin3(m4[j], cm[8 * j + 5], cm[8 * j + 6]);
            m4[j] = fminf(m4[j], cm[8 * j + 7]);
          }
          const float m = fminf(ptx::fmin3(m4[0], m4[1], m4[2]), m4[3]);
          const float thr = m + delta;
          uint32_t cm4[4] = {0u, 0u, 0u, 0u}, bm4[4] = {0u, 0u, 0u, 0u};
          #pragma unroll
          for (int j = 0; j < 32; ++j) {
            cm4[j & 3] |= (cm[j] <= thr) ? (1u << j) : 0u;
            bm4[j & 3] |= (bmin[j] <= thr) ? (1u << j) : 0u;
          }
          const uint32_t cmask = (cm4[0] | cm4[1]) | (cm4[2] | cm4[3]);
          const uint32_t bmask = ((bm4[0] | bm4[1]) | (bm4[2] | bm4[3])) >> (32 - 4 * nchunks);   // bit a = a-th batch processed
          const int nc = __popc(cmask), nb = __popc(bmask);
          const bool full = outl || cmask == 0u || bmask == 0u;     // masks are empty only for NaN scores
          const int ncand = nc * nb;
          // bmask bit a = a-th batch in this CTA's processing order; its codes start at batch_base(a)
          auto batch_base = [&](int a) { int pc = (a >> 2) + rot; pc = pc < nchunks ? pc : pc - nchunks; return pc * 128 + (a & 3) * 32; };
          int4 cd = make_int4(batch_base(__ffs(bmask) - 1) + (__ffs(cmask) - 1), -1, -1, -1);
          if (!full && ncand > 1 && ncand <= 4) {
            int cc[4] = {-1, -1, -1, -1};
            int w = 0;
            uint32_t bm2 = bmask;
            while (bm2) {
              const int a = __ffs(bm2) - 1; bm2 &= bm2 - 1;
              uint32_t cm2 = cmask;
              while (cm2) {
                const int j = __ffs(cm2) - 1; cm2 &= cm2 - 1;
                const int code = batch_base(a) + j;
                if (w == 0) cc[0] = code; else if (w == 1) cc[1] = code; else if (w == 2) cc[2] = code; else cc[3] = code;
                ++w;
              }
            }
            cd = make_int4(cc[0], cc[1], cc[2], cc[3]);
          }
          *reinterpret_cast<int4*>(smem + SmemLayout::cand + f * 16) = cd;
          *reinterpret_cast<int*>(smem + SmemLayout::ncnt + f * 4) = full ? kFull : (ncand > 4 ? kBig : ncand);
          *reinterpret_cast<uint32_t*>(smem + SmemLayout::cmask + f * 4) = cmask;
          *reinterpret_cast<uint32_t*>(smem + SmemLayout::bmask + f * 4) = bmask;
          {
            int* qc = reinterpret_cast<int*>(smem + SmemLayout::qcnt) + qpar * 2;
            if (!full && ncand > 1) {
              if (ncand <= 4) smem[SmemLayout::slowq + qpar * kM + atomicAdd(&qc[0], 1)] = (unsigned char)f;
              else            smem[SmemLayout::wideq + qpar * kM + atomicAdd(&qc[1], 1)] = (unsigned char)f;
            }
            // frames outside the fp16 image's validity range (or NaN): exact scan right here, then they are certified
            uint32_t fm = __ballot_sync(0xffffffffu, full);
            if (fm) __syncwarp();                     // the lanes' ncnt / mask entries are read by the whole warp below
            while (fm) {
              const int i = __ffs(fm) - 1; fm &= fm - 1;
              resolve_big(smem, q * 32 + i, lane, p.K, rot, t32, cn);
            }
          }
          n_full += full ? 1u : 0u; n_cert += (!full && ncand == 1) ? 1u : 0u; n_resc += (!full && ncand > 1) ? 1u : 0u;
          // upper bound of the next residual's |r|^2 (only the margin and the validity test use it):
          // the winner's approximate score is <= m + delta and off by <= delta/2
          if (full) { const float g2 = xnorm + meta->cmax_all; xx = g2 * g2; }
          else xx = fmaxf(xx + m + 1.5f * delta, 0.f) * 1.00001f + 1e-30f;
        }
        RVQ_TICK(t_win);
        RVQ_TRACE(s, 10);
        ptx::named_bar_sync(5, 256);             // candidate lists and re-score queues visible to all 8 frame warps
        RVQ_TRACE(s, 11);
#ifdef RVQ_TC_TRACE
        if (blockIdx.x == 0 && t_tile == 0 && s < kTraceStages && lane == 0) {     // probe: latency of one dependent table load
          const long long ta = clock64();
          const float pv0 = __ldg(t32 + size_t((s * 37 + warp * 5 + 3) % p.K) * 128 + 64);
          const long long tb = clock64() + (pv0 == 123.456f ? 1 : 0);
          g_trace[(s * 11 + warp) * kTraceEv + 15] = tb - ta;
        }
#endif
        float sq = 0.f;
        update_pass<false, TRAIN>(p, smem, q, h, lane, s, rot, nchunks, tile_n0, t32, cn, sq, qpar);
        RVQ_TRACE(s, 13);
        if (s + 1 < p.n_q) {
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_a);
        }
        if (TRAIN && p.sqerr != nullptr) {
          sq = warp_sum(sq);
          if (lane == 0) atomicAdd(&p.sqerr[s], (double)sq);
        }
        RVQ_TICK(t_upd);
      }
      ptx::named_bar_sync(5, 256);               // every frame of the tile has its final residual (re-scores run on any warp)
      if (TRAIN && p.residual_out != nullptr) {
        // each warp writes the 16 frames it owns, 512 contiguous bytes per frame
        for (int i = 0; i < 16; ++i) {
          const int fo = q * 32 + h * 16 + i;
          const int64_t n = tile_n0 + fo;
          if (n < p.N) *reinterpret_cast<float4*>(p.residual_out + n * 128 + lane * 4) = *reinterpret_cast<const float4*>(rs + rs_off(fo, lane));
        }
      }
      ptx::named_bar_sync(pair_bar, 64);         // both warps are done with this tile's rows
    }
    // search statistics (evidence; see rvq_search_stats)
    if (h == 0 && p.counters != nullptr) {
      #pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        n_cert += __shfl_xor_sync(0xffffffffu, n_cert, off);
        n_resc += __shfl_xor_sync(0xffffffffu, n_resc, off);
        n_full += __shfl_xor_sync(0xffffffffu, n_full, off);
      }
      if (lane == 0) {
        atomicAdd(&p.counters[0], (unsigned long long)(n_cert + n_resc + n_full)); atomicAdd(&p.counters[1], (unsigned long long)n_cert);
        atomicAdd(&p.counters[2], (unsigned long long)n_resc); atomicAdd(&p.counters[3], (unsigned long long)n_full);
#ifdef RVQ_TC_TIMERS
        atomicAdd(&p.counters[4], (unsigned long long)t_wait); atomicAdd(&p.counters[5], (unsigned long long)t_epi);
        atomicAdd(&p.counters[6], (unsigned long long)t_win);  atomicAdd(&p.counters[7], (unsigned long long)t_upd);
        atomicAdd(&p.counters[8], (unsigned long long)t_load); atomicAdd(&p.counters[9], (unsigned long long)(clock64() - t_begin));
        atomicAdd(&p.counters[10], 1ull);
#endif
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) ptx::tmem_dealloc(tmem, 512);
}

int tc_debug_trace(long long* out_host, int n) {
  const int m = n < kTraceStages * 11 * kTraceEv ? n : kTraceStages * 11 * kTraceEv;
  RVQ_CUDA(cudaMemcpyFromSymbol(out_host, g_trace, size_t(m) * sizeof(long long)));
  return m;
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
    RVQ_CUDA(cudaFuncSetAttribute(tc_encode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemLayout::total));
    RVQ_CUDA(cudaFuncSetAttribute(tc_encode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemLayout::total));
    sm_dev = dev;
  }
  PackView pv(a.pack, a.K, a.D);
  RVQ_CUDA(cudaMemsetAsync(pv.counters(), 0, 32 * sizeof(unsigned long long), st));
  TcParams p;
  p.pack = (const unsigned char*)a.pack; p.K = a.K;
  p.x = a.x; p.fa = FrameAddr{a.sxb, a.sxd, a.sxt, a.T}; p.N = N;
  p.stage0 = a.stage0; p.n_q = a.n_q;
  p.codes = a.codes; p.residual_out = a.residual_out; p.sqerr = a.sqerr;
  p.ste = (a.flags & RVQ_FLAG_STE) ? 1 : 0;
  p.counters = pv.counters();
  const int64_t ntiles = (N + kM - 1) / kM;
  const unsigned grid = unsigned(ntiles < sm_count ? ntiles : sm_count);
  // the lean variant serves plain encodes; straight-through arithmetic, loss numerators and the residual output
  // live in the other one (a stage's hot code has to fit the instruction cache)
  if (p.ste || p.sqerr != nullptr || p.residual_out != nullptr) tc_encode_kernel<true><<<grid, kThreadsTc, SmemLayout::total, st>>>(p);
  else tc_encode_kernel<false><<<grid, kThreadsTc, SmemLayout::total, st>>>(p);
  RVQ_LAUNCH_CHECK("tc_encode_kernel");
  if (a.quantized != nullptr)
    return simt_quant_sum(a.pack, a.K, a.D, a.x, p.fa, N, a.T, a.stage0, a.n_q, a.codes, a.quantized, p.ste,
                          (a.flags & RVQ_FLAG_ACCUM_Q) ? 1 : 0, st);
  return RVQ_OK;
}

}  // namespace rvq

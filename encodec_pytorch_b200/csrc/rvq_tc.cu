// Fused multi-stage nearest-code search on the 5th-generation tensor cores (tcgen05 / TMEM), two tiles in flight.
//
// One persistent CTA per SM walks a contiguous range of 128-frame tiles, TWO AT A TIME (slot 0 takes the even
// tiles of the range, slot 1 the odd ones).  A tile's stage is a serial chain
//     MMA (scores of the 128 frames against all K codes) -> score reduction -> winner -> residual update -> next MMA,
// so the two slots run half a stage apart: while the tensor pipe computes the scores of one slot, the SIMT warps
// choose the winners of / update the other one.  The fp32 residual of every frame stays in shared memory across
// all n_q stages (core_vq.py:357-367 without the per-stage round trips through HBM).
//
//   scores   S[f,k] = -2 r_f . c_k + |c_k|^2 from tcgen05.mma kind::f16 (fp32 accumulation in tensor memory),
//            M=128 frames, N=128 codes per chunk, 9 K-steps of 16:
//            A = fp16(r), held IN TENSOR MEMORY (64 columns per slot; K-steps 0..7) + one constant shared-memory
//                block whose columns 128,129 = 1 (K-step 8: picks up the hi/lo halves of |c|^2);
//            B = fp16 image of the codebook (cols 0..127 = -2c, cols 128,129 = hi/lo of |c|^2), streamed by the
//                TMA engine (cp.async.bulk) from the L2-resident pack into a ring of 7 third-of-a-chunk
//                slots (6 K-groups = 12 KB each; small slots keep more bytes in flight than whole chunks would);
//            D = 3 accumulator buffers of 128 TMEM columns shared by both slots.
//   warps    0..3   score warps  (thread = TMEM lane = frame): tcgen05.ld, per-class / per-batch minima, certified
//                   winner or candidate list (see below), codes of certified frames;
//            4..11  update warps: gather of the winners' fp32 rows ("lane = dimension", a quarter-warp per frame, four
//                   rows in flight per quarter-warp), exact fp32 re-score of candidate lists, r <- r - q; then, thread =
//                   frame, the fp16 operand of the next stage goes to tensor memory (tcgen05.st) together with its
//                   exact rounding residue; tile loads;
//            12  TMA producer;  13  MMA issuer (owns the TMEM allocation).
//            setmaxnreg: 160 registers for the score warps, 152 for the update warps, 48 for the last warpgroup.
//   sync     mbarriers only between roles: a_ready[slot] (update -> MMA), acc_full/acc_empty (MMA <-> score),
//            cand_ready[slot] (score -> update), full/empty (TMA <-> MMA).
//
// Certified argmin: a score warp keeps per frame the minimum over every 32-code batch and over every residue class
// (code mod 32).  A code is within `delta` of the minimum iff its batch AND its class are; delta bounds the fp16
// score error two-sidedly (rvq_common.cuh, StageMeta), so the exact fp32 winner is certified when exactly one batch
// and one class qualify.  Otherwise the candidates (flagged batches x flagged classes) are re-scored in fp32 with
// the reference's formula (core_vq.py:181-189, ties -> lowest index).  Frames outside the fp16 image's validity
// range take an exact fp32 scan.
#include "rvq_common.cuh"
#include "rvq_ptx.cuh"

#include <cstdlib>

namespace rvq {

namespace {

constexpr int kM = 128;                 // frames per tile (UMMA M, TMEM lanes)
constexpr int kN = kTcChunkCodes;       // 128 codes per MMA group (UMMA N)
constexpr int kRing = 7;                // B ring slots; each holds one K-third of a chunk (6 K-groups = 3 K-steps)
constexpr int kSlotBytes = 6 * kTcLBO;  // 12288 B
constexpr int kAccBufs = 3;             // accumulator buffers of kN TMEM columns
constexpr int kTmemA = kAccBufs * kN;   // first TMEM column of the fp16 operands (64 columns per slot)
constexpr int kThreadsTc = 16 * 32;
constexpr int kUpdWarps = 8;
#ifndef RVQ_TC_WIN
#define RVQ_TC_WIN 3
#endif
constexpr int kWin = RVQ_TC_WIN;          // winner rows in flight per quarter-warp
constexpr int kBig = 5;                 // ncnt marker: more than 4 candidates (enumerate the masks)
constexpr int kFull = 6;                // ncnt marker: exact scan of the whole table
constexpr int kRsBytes = kM * 128 * 4;  // fp32 residual of one tile

struct Sm {
  static constexpr uint32_t aug = 0;                               // [2 k-groups][128 rows][16 B], no swizzle
  static constexpr uint32_t ring = aug + 4096;
  static constexpr uint32_t rs = ring + kRing * kSlotBytes;        // 2 x fp32 [128 f][128 d], chunk-swizzled
  static constexpr uint32_t misc = rs + 2 * kRsBytes;              // 2 x per-slot block (offsets m_*)
  static constexpr uint32_t m_cand = 0;                            // int4 [128]: candidate codes (-1 = none)
  static constexpr uint32_t m_ncnt = m_cand + kM * 16;             // int [128]
  static constexpr uint32_t m_cmask = m_ncnt + kM * 4;             // u32 [128] flagged classes   (a fresh tile: |x|^2 of dims 0..63)
  static constexpr uint32_t m_bmask = m_cmask + kM * 4;            // u32 [128] flagged batches   (a fresh tile: |x|^2 of dims 64..127)
  static constexpr uint32_t m_dr2 = m_bmask + kM * 4;              // float [2][128]: |r - fp16(r)|^2 of the current operand, per half of the dims
  static constexpr uint32_t m_slowq = m_dr2 + 2 * kM * 4;          // u8 [128]: frames with 2..4 listed candidates
  static constexpr uint32_t m_wideq = m_slowq + kM;                // u8 [128]: frames with a wide candidate set
  static constexpr uint32_t m_qcnt = m_wideq + kM;                 // int [2]: queue lengths {slow, wide}
  static constexpr uint32_t m_size = m_qcnt + 16;
  static constexpr uint32_t bars = misc + 2 * m_size;
  static constexpr uint32_t total = bars + 256;
};
struct Bars {
  uint64_t full[kRing], empty[kRing], acc_full[kAccBufs], acc_empty[kAccBufs], a_ready[2], cand_ready[2];
  uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 256, "barrier block");
static_assert(Sm::total <= 227 * 1024, "shared memory budget");
static_assert(kTcKPad / 16 == 9 && kN == 128 && kTmemA + 2 * 64 == 512, "operand geometry");
static_assert((Sm::m_size % 16) == 0 && (Sm::misc % 16) == 0, "alignment");

#ifdef RVQ_TC_TIMERS
#define RVQ_TICK(acc) do { const unsigned tt_ = (unsigned)clock(); acc += tt_ - tc0; tc0 = tt_; } while (0)
#define RVQ_TICK0() unsigned tc0 = (unsigned)clock()
#else
#define RVQ_TICK(acc) do { } while (0)
#define RVQ_TICK0() do { } while (0)
#endif

struct TcParams {
  const unsigned char* pack; int K;
  const float* x; FrameAddr fa; int64_t N;
  int stage0, n_q;
  int64_t* codes; float* residual_out; double* sqerr;
  int ste;
  int tf;                  // frames per tile (<= 128): chosen by the host so that every CTA gets an even number of tiles
  unsigned long long* counters;
};

__device__ __forceinline__ float inf_f() { return __int_as_float(0x7f800000); }
// residual element group: 16-byte chunk ch (dims 4ch..4ch+3) of frame f, XOR-swizzled so that both
// "lanes = consecutive chunks of one frame" and "lanes = consecutive frames, one chunk" spread over banks
__device__ __forceinline__ int rs_off(int f, int ch) { return f * 128 + ((ch ^ (f & 31)) << 2); }
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float warp_sum(float v) {
  #pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
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
// squared rounding residue of two floats against their fp16 pair
__device__ __forceinline__ float residue2(float a, float b, uint32_t& word, float e2) {
  const __half2 h = __floats2half2_rn(a, b);
  word = *reinterpret_cast<const uint32_t*>(&h);
  const float2 bk = __half22float2(h);
  const float ea = a - bk.x, eb = b - bk.y;
  return fmaf(eb, eb, fmaf(ea, ea, e2));
}

// tiles [start, start+cnt) of this CTA
__device__ __forceinline__ void cta_range(int ntiles, int& start, int& cnt) {
  const int base = ntiles / int(gridDim.x), rem = ntiles % int(gridDim.x);
  const int b = blockIdx.x;
  start = b * base + (b < rem ? b : rem);
  cnt = base + (b < rem ? 1 : 0);
}

// torch's CPU argmax (core_vq.py:188) propagates NaN: the first NaN distance wins; otherwise the smallest
// distance, lowest index on ties.  (best, bcode) starts as (+inf, 0x7fffffff).
__device__ __forceinline__ bool nan_aware_better(float dist, int code, float best, int bcode) {
  if (dist != dist) return best == best || code < bcode;
  return best == best && (dist < best || (dist == best && code < bcode));
}

// Score-warp side.  Frames whose candidate set is the whole table (outside the fp16 image's validity range, NaN):
// the whole warp scores the table, one code per lane, stores the code and rewrites the frame's entry as a
// certified winner for the update warps.
__device__ __forceinline__ void resolve_full(const float* rs, unsigned char* ms, int f, int lane, int K, const float* __restrict__ t32,
                                          const float* __restrict__ cn, int64_t* code_out) {
  const float4 rl = *reinterpret_cast<const float4*>(rs + rs_off(f, lane));
  const float rr = warp_sum(dot4(rl, rl, 0.f));
  float best = inf_f(); int bcode = 0x7fffffff;
  for (int code = lane; code < K; code += 32) {
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
    const int code = bcode == 0x7fffffff ? 0 : bcode;
    *reinterpret_cast<int4*>(ms + Sm::m_cand + f * 16) = make_int4(code, -1, -1, -1);
    *reinterpret_cast<int*>(ms + Sm::m_ncnt + f * 4) = 1;
    if (code_out != nullptr) *code_out = code;
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
// exact fp32 r <- r - q (core_vq.py:364 / :348; straight-through arithmetic of :309 in training), the exact squared
// rounding residue |r - fp16(r)|^2 of the new residual (it enters the score-error margin of the next stage) and
// the squared-error partial.  Called by whole quarter-warps (8 converged lanes; all lanes of the warp shuffle).
// new residual n = r - q (core_vq.py:364 / :348; straight-through arithmetic of :309 in training)
template <bool TRAIN>
__device__ __forceinline__ float4 sub_row(const TcParams& p, const float4& rv, float4 q) {
  if (TRAIN && p.ste) { q.x = rv.x + (q.x - rv.x); q.y = rv.y + (q.y - rv.y); q.z = rv.z + (q.z - rv.z); q.w = rv.w + (q.w - rv.w); }
  return make_float4(rv.x - q.x, rv.y - q.y, rv.z - q.z, rv.w - q.w);
}
// exact fp32 r <- r - q for a frame at an arbitrary position (candidate-list / wide paths)
template <bool TRAIN>
__device__ __forceinline__ void apply_row(const TcParams& p, float* rs, int f, int j, const Row4& r, const Row4& qrow) {
  #pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(rs + rs_off(f, 8 * i + j)) = sub_row<TRAIN>(p, r.v[i], qrow.v[i]);
}

// A frame whose flagged batches x flagged classes give more than 4 candidates: the whole warp works on it,
// 8 candidates per step (2 per quarter-warp); the quarter that found the winner updates the frame.
template <int NC> struct Cand { int c[NC]; Row4 w[NC]; float nrm[NC]; };
template <int NC>
__device__ __forceinline__ void load_cand(Cand<NC>& k, int j, const float* __restrict__ t32, const float* __restrict__ cn) {
  #pragma unroll
  for (int u = 0; u < NC; ++u) {
    k.nrm[u] = 0.f;
    if (k.c[u] >= 0) { k.w[u] = load_row(t32, k.c[u], j); k.nrm[u] = __ldg(cn + k.c[u]); }
  }
}
// exact distances of up to NC candidates (core_vq.py:183-187); keeps the best (lowest code on ties) and its slot u
template <int NC>
__device__ __forceinline__ void score_cand(const Cand<NC>& k, const Row4& r, float rr, float& best, int& bcode, int& bidx) {
  float d[NC];
  #pragma unroll
  for (int u = 0; u < NC; ++u) d[u] = dot_row(r, k.w[u]);
  #pragma unroll
  for (int off = 4; off > 0; off >>= 1) {
    #pragma unroll
    for (int u = 0; u < NC; ++u) d[u] += __shfl_xor_sync(0xffffffffu, d[u], off);
  }
  #pragma unroll
  for (int u = 0; u < NC; ++u) {
    const float e = (rr - 2.f * d[u]) + k.nrm[u];
    if (k.c[u] >= 0 && (e < best || (e == best && k.c[u] < bcode))) { best = e; bcode = k.c[u]; bidx = u; }
  }
}
template <bool TRAIN>
__device__ __forceinline__ void resolve_wide(const TcParams& p, float* rs, unsigned char* ms, int f, int lane, int s, int rot, int nchunks,
                                             int64_t tile_n0, const float* __restrict__ t32, const float* __restrict__ cn) {
  const int qq = lane >> 3, j = lane & 7;
  const uint32_t cm = *reinterpret_cast<const uint32_t*>(ms + Sm::m_cmask + f * 4);
  const uint32_t bm = *reinterpret_cast<const uint32_t*>(ms + Sm::m_bmask + f * 4);
  const int nc = __popc(cm);
  const Row4 r = load_res(rs, f, j);
  const float rr = quarter_sum(dot_row(r, r));
  float best = inf_f(); int bcode = 0x7fffffff, bidx = 0;
  // quarter qq takes the flagged batches number qq, qq + 4, ...; within a batch the flagged classes two at a time
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
    for (int oc = 0; oc < nc; oc += 2) {
      Cand<2> k;
      #pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int jj = cmq ? __ffs(cmq) - 1 : -1;
        cmq &= cmq - 1;
        k.c[u] = (a >= 0 && jj >= 0) ? base + jj : -1;
      }
      load_cand<2>(k, j, t32, cn);
      score_cand<2>(k, r, rr, best, bcode, bidx);
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
                bcode = pc * 128 + (c0 & 3) * 32 + (__ffs(cm) - 1); }
  }
  const int64_t nfr = tile_n0 + f;
  if (mine) {
    apply_row<TRAIN>(p, rs, f, j, r, load_row(t32, bcode, j));     // the winner's row again (rare path; an L1/L2 hit)
    if (j == 0 && f < p.tf && nfr < p.N) p.codes[int64_t(s) * p.N + nfr] = bcode;
  }
}

// The residual update of one (slot, stage), executed by the eight update warps (u = 0..7: TMEM lane quadrant q = u & 3,
// half h = u >> 2 of its 32 frames).
//   1. frames with wide candidate sets: one frame per warp at a time (resolve_wide);
//   2. certified frames, branch-free: quarter-warp qq owns frames f0 .. f0+3 (f0 = 32q + 16h + 4qq).  Lane j works on the
//      16-byte slots 8(i ^ ix) + j (i = 0..3) of each residual row; with the row swizzle of rs_off that slot holds the
//      logical chunk 8i + (j ^ jx) of the frame, so the winner's row is fetched with its 16-byte pieces permuted by
//      jx (the 8 lanes of a quarter still cover one contiguous 128-byte segment per i; all addresses are a base +
//      an immediate).  All
//      four rows are in flight at once.  A queued frame's slot fetches row 0 and its stores are predicated off;
//   3. frames with a candidate list (2..4 codes) come from the slot's queue, spread over the 32 quarter-warps:
//      the four candidate rows are loaded together and re-scored in exact fp32 (core_vq.py:183-187).
template <bool TRAIN>
__device__ __forceinline__ void update_pass(const TcParams& p, float* rs, unsigned char* ms, int u, int lane, int s, int rot,
                                            int nchunks, int64_t tile_n0, const float* __restrict__ t32,
                                            const float* __restrict__ cn, uint32_t (&tsub)[4]) {
  const int qq = lane >> 3, j = lane & 7;
  const int q = u & 3, h = u >> 2;
  const int fbase = q * 32 + h * 16;
  RVQ_TICK0();
  const int* qc = reinterpret_cast<const int*>(ms + Sm::m_qcnt);
  const int nslow = qc[0], nwide = qc[1];
  const unsigned char* slowq = ms + Sm::m_slowq;
  const unsigned char* wideq = ms + Sm::m_wideq;
  const int nv = lane < 16 ? *reinterpret_cast<const int*>(ms + Sm::m_ncnt + (fbase + lane) * 4) : 1;
  const uint32_t slow = __ballot_sync(0xffffffffu, nv > 1);     // bit i = frame fbase+i is in one of the queues
  const int4* cand = reinterpret_cast<const int4*>(ms + Sm::m_cand);
  {
    const int f0 = fbase + 4 * qq;
    const int ix = 2 * h + (qq >> 1), jx0 = 4 * (qq & 1);
    // physical slot 8(i ^ ix) + j of frame f0 + k holds the logical chunk 8i + (j ^ (jx0 + k)): four shared-memory bases
    // (one per i) + immediates on the residual side, one row pointer per frame + immediates on the table side
    float* rb[4];
    #pragma unroll
    for (int i = 0; i < 4; ++i) rb[i] = rs + f0 * 128 + 32 * (i ^ ix) + 4 * j;
    int oc[4];
    #pragma unroll
    for (int k = 0; k < 4; ++k) oc[k] = ((slow >> (4 * qq + k)) & 1u) ? -1 : cand[f0 + k].x;
    // a rolling window of kWin rows in flight: the buffer of frame k is refilled with the row of frame k + kWin
    Row4 buf[kWin];
    auto fetch = [&](int k) {
#ifdef RVQ_EXP_ROW0       // experiment (wrong results): every gather hits the same L1-resident row
      const float4* rp = reinterpret_cast<const float4*>(t32 + size_t(oc[k] >= 0 ? 0 : 0) * 128 + 4 * (j ^ (jx0 + k)));
#else
      const float4* rp = reinterpret_cast<const float4*>(t32 + size_t(oc[k] >= 0 ? oc[k] : 0) * 128 + 4 * (j ^ (jx0 + k)));
#endif
      #pragma unroll
      for (int i = 0; i < 4; ++i) buf[k % kWin].v[i] = __ldg(rp + 8 * i);
    };
    #pragma unroll
    for (int k = 0; k < kWin; ++k) fetch(k);
    RVQ_TICK(tsub[0]);
    #pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 r[4];
      #pragma unroll
      for (int i = 0; i < 4; ++i) r[i] = *reinterpret_cast<const float4*>(rb[i] + k * 128);
      #pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 n = sub_row<TRAIN>(p, r[i], buf[k % kWin].v[i]);
        if (oc[k] >= 0) *reinterpret_cast<float4*>(rb[i] + k * 128) = n;
      }
      if (k + kWin < 4) fetch(k + kWin);
    }
  }
  RVQ_TICK(tsub[1]);
  // candidate lists: item qi goes to quarter qi / 8 of warp qi % 8 (the first eight items land on eight different
  // warps), then qi + 32, ...; a warp none of whose quarters has an item skips the body
  const int gq = qq * kUpdWarps + u;
  #pragma unroll 1
  for (int qi = gq; qi < ((nslow + 31) & ~31); qi += 32) {
    const int f = qi < nslow ? int(slowq[qi]) : -1;
    if (!__any_sync(0xffffffffu, f >= 0)) continue;
    int4 cd = make_int4(-1, -1, -1, -1);
    if (f >= 0) cd = cand[f];
    Cand<4> k;
    k.c[0] = cd.x; k.c[1] = cd.y; k.c[2] = cd.z; k.c[3] = cd.w;
    load_cand<4>(k, j, t32, cn);
    Row4 r;
    #pragma unroll
    for (int i = 0; i < 4; ++i) r.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f >= 0) r = load_res(rs, f, j);
    const float rr = quarter_sum(dot_row(r, r));
    // core_vq.py:183-187, lowest index on ties; NaN distances: keep the first candidate unless a finite one exists
    float best = inf_f(); int bcode = 0x7fffffff, bidx = 0;
    score_cand<4>(k, r, rr, best, bcode, bidx);
    if (bcode == 0x7fffffff) bcode = cd.x;
    const int64_t nfr = tile_n0 + f;
    if (f >= 0) {
      #pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a01 = bidx == 1 ? k.w[1].v[i] : k.w[0].v[i], a23 = bidx == 3 ? k.w[3].v[i] : k.w[2].v[i];
        *reinterpret_cast<float4*>(rs + rs_off(f, 8 * i + j)) = sub_row<TRAIN>(p, r.v[i], bidx >= 2 ? a23 : a01);
      }
      if (j == 0 && f < p.tf && nfr < p.N) p.codes[int64_t(s) * p.N + nfr] = bcode;
    }
  }
  RVQ_TICK(tsub[2]);
  // frames with wide candidate sets: one frame per warp at a time, handed out from the last warp down (the candidate
  // lists start at the first)
  #pragma unroll 1
  for (int i = kUpdWarps - 1 - u; i < nwide; i += kUpdWarps) resolve_wide<TRAIN>(p, rs, ms, wideq[i], lane, s, rot, nchunks, tile_n0, t32, cn);
  RVQ_TICK(tsub[3]);
}

// Thread = frame = TMEM lane (f = 32q + lane), dims 64h .. 64h+63 of the new residual: fp16 operand of the next stage
// to tensor memory (16 dims per tcgen05.st), its exact squared rounding residue |r - fp16(r)|^2 (it enters the score
// margin of the next stage) and, in training, the squared-error partial sum((q - r)^2) = |new residual|^2 (core_vq.py:319).
template <bool TRAIN>
__device__ __forceinline__ void operand_pass(const float* rs, unsigned char* ms, int f, int h, uint32_t taddr, bool store, float& sq) {
  float e[4] = {0.f, 0.f, 0.f, 0.f};
  #pragma unroll 2
  for (int g = 0; g < 4; ++g) {
    uint32_t w[8];
    #pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(rs + rs_off(f, 16 * h + 4 * g + c));
      e[c] = residue2(v.x, v.y, w[2 * c], e[c]);
      e[c] = residue2(v.z, v.w, w[2 * c + 1], e[c]);
      if (TRAIN) sq = dot4(v, v, sq);
    }
    if (store) ptx::tmem_st8(taddr + 32 * h + 8 * g, w);
  }
  reinterpret_cast<float*>(ms + Sm::m_dr2)[h * kM + f] = (e[0] + e[1]) + (e[2] + e[3]);
}

}  // namespace

template <bool TRAIN>
__global__ void __launch_bounds__(kThreadsTc, 1) tc_encode_kernel(const TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  Bars* bars = reinterpret_cast<Bars*>(smem + Sm::bars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PackView pv(p.pack, p.K, 128);
  const int nchunks = p.K / kN;
  // every CTA walks the chunks of a stage in its own rotation, so that the 148 SMs (which run the same stage
  // at about the same time) do not all pull the same lines out of the same L2 slices at once
  const int rot = int(blockIdx.x % unsigned(nchunks));
  const int ntiles = int((p.N + p.tf - 1) / p.tf);   // N < 2^31 frames per call (checked by rvq_encode)
  int tile0, tcnt;
  cta_range(ntiles, tile0, tcnt);
  // slot 0 takes tiles tile0, tile0+2, ...; slot 1 takes tile0+1, tile0+3, ...; a slot's step n = (tile-in-slot) * n_q + stage.
  // Every role walks the same global order: (slot 0, n), (slot 1, n) for n = 0, 1, ...; slot 1 may run out one tile earlier.
  const int steps0 = ((tcnt + 1) >> 1) * p.n_q, steps1 = (tcnt >> 1) * p.n_q;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) { ptx::mbar_init(ptx::smem_u32(&bars->full[i]), 1); ptx::mbar_init(ptx::smem_u32(&bars->empty[i]), 1); }
    for (int i = 0; i < kAccBufs; ++i) { ptx::mbar_init(ptx::smem_u32(&bars->acc_full[i]), 1); ptx::mbar_init(ptx::smem_u32(&bars->acc_empty[i]), 4); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(ptx::smem_u32(&bars->a_ready[i]), kUpdWarps); ptx::mbar_init(ptx::smem_u32(&bars->cand_ready[i]), 4); }
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < 2) {
    int* qc = reinterpret_cast<int*>(smem + Sm::misc + threadIdx.x * Sm::m_size + Sm::m_qcnt);
    qc[0] = 0; qc[1] = 0;
  }
  // constant augmented K block of A: k-group 0 = (1, 1, 0, ...) picks up hi/lo of |c|^2, k-group 1 = 0
  for (int i = threadIdx.x; i < 4096 / 16; i += blockDim.x)
    *reinterpret_cast<uint4*>(smem + Sm::aug + i * 16) = make_uint4(i < 128 ? pack_half2(1.f, 1.f) : 0u, 0u, 0u, 0u);
  if (warp == 13) {
    ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp >= 12) {
    ptx::reg_dec<48>();
    if (warp == 12) {
      // ===== TMA producer: the three K-thirds of every 128-code chunk, in the global step order.  One thread: the slots
      // come free in the order they were filled, and a spinning warp costs the working warps of its scheduler issue slots =====
      if (lane == 0) {
        uint32_t slot = 0, ph = 0;             // ring position of the next K-third
        for (int n = 0; n < steps0; ++n) {
          for (int X = 0; X < 2; ++X) {
            if (X == 1 && n >= steps1) break;
            const unsigned char* img = pv.tc(p.stage0 + n % p.n_q);
            for (int c = 0; c < nchunks; ++c) {
              const int pc = c + rot < nchunks ? c + rot : c + rot - nchunks;
              #pragma unroll 1
              for (int third = 0; third < 3; ++third) {
                ptx::mbar_wait(ptx::smem_u32(&bars->empty[slot]), ph ^ 1);
                const uint32_t fb = ptx::smem_u32(&bars->full[slot]);
                ptx::mbar_expect_tx(fb, kSlotBytes);
                ptx::bulk_g2s(sbase + Sm::ring + slot * kSlotBytes, img + size_t(pc) * kTcChunkBytes + third * kSlotBytes, kSlotBytes, fb);
                if (++slot == kRing) { slot = 0; ph ^= 1; }
              }
            }
          }
        }
      }
      __syncwarp();
    } else if (warp == 13) {
      // ===== MMA issuer: per chunk 8 MMAs with A from tensor memory + 1 with the constant shared-memory block =====
      if (lane == 0) {
        constexpr uint32_t idesc = ptx::umma_idesc_f16_f32(kM, kN);
        const uint64_t ad_aug = ptx::umma_desc_kmajor_noswz(sbase + Sm::aug, 2048, 128);
        const uint64_t bd0 = ptx::umma_desc_kmajor_noswz(sbase + Sm::ring, kTcLBO, kTcSBO);
#ifdef RVQ_TC_TIMERS
        uint32_t m_wa = 0, m_wf = 0, m_wc = 0, m_is = 0; const long long m_t0 = clock64();
#endif
        uint32_t slot = 0, ph = 0, buf = 0, bph = 0;      // ring slot / phase of the next K-third, accumulator buffer / phase
        for (int n = 0; n < steps0; ++n) {
          for (int X = 0; X < 2; ++X) {
            if (X == 1 && n >= steps1) break;
            RVQ_TICK0();
            ptx::mbar_wait(ptx::smem_u32(&bars->a_ready[X]), uint32_t(n) & 1);       // fp16 operand of this step is in TMEM
            ptx::tc_fence_after();
            RVQ_TICK(m_wa);
            const uint32_t a_tmem = tmem + kTmemA + 64 * X;
            for (int c = 0; c < nchunks; ++c) {
              ptx::mbar_wait(ptx::smem_u32(&bars->acc_empty[buf]), bph ^ 1);                     // accumulator drained
              RVQ_TICK(m_wc);
              const uint32_t d_tmem = tmem + buf * kN;
              #pragma unroll
              for (int h = 0; h < 3; ++h) {
                ptx::mbar_wait(ptx::smem_u32(&bars->full[slot]), ph);                            // K-third landed
                ptx::tc_fence_after();
                RVQ_TICK(m_wf);
                const uint64_t bs = bd0 + uint64_t((slot * kSlotBytes) >> 4);
                #pragma unroll
                for (int k = 0; k < 3; ++k) {
                  const uint64_t bk = bs + uint64_t((k * 2 * kTcLBO) >> 4);
                  if (h == 2 && k == 2) ptx::umma_f16_ss(d_tmem, ad_aug, bk, idesc, 1u);
                  else ptx::umma_f16_ts(d_tmem, a_tmem + 8 * (3 * h + k), bk, idesc, (h | k) ? 1u : 0u);
                }
                ptx::umma_commit(ptx::smem_u32(&bars->empty[slot]));   // ring slot reusable once read
                if (++slot == kRing) { slot = 0; ph ^= 1; }
                RVQ_TICK(m_is);
              }
              ptx::umma_commit(ptx::smem_u32(&bars->acc_full[buf]));   // scores ready for the score warps
              if (++buf == kAccBufs) { buf = 0; bph ^= 1; }
            }
          }
        }
#ifdef RVQ_TC_TIMERS
        if (p.counters != nullptr) {
          atomicAdd(&p.counters[11], (unsigned long long)m_wa); atomicAdd(&p.counters[12], (unsigned long long)m_wf);
          atomicAdd(&p.counters[13], (unsigned long long)m_wc); atomicAdd(&p.counters[14], (unsigned long long)(clock64() - m_t0));
          atomicAdd(&p.counters[19], (unsigned long long)m_is);
        }
#endif
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    ptx::reg_inc<152>();
    // ===== update warps =====
    const int u = warp - 4;
    const int q = u & 3, h = u >> 2;           // TMEM lane quadrant (frames 32q..32q+31 of a tile), half of the dims / of the frames
    const int f = q * 32 + lane;               // thread <-> frame mapping of tile loads and operand stores
    const uint32_t tq = tmem + (uint32_t(q * 32) << 16) + kTmemA;
#ifdef RVQ_TC_TIMERS
    uint32_t t_wait = 0, t_upd = 0, t_tr = 0, t_load = 0, t_stw = 0, t_bar = 0;
#endif
    uint32_t tsub[4] = {0u, 0u, 0u, 0u};     // (timers) wide sets / own frames / candidate lists / barrier
    // load dims 64h..64h+63 of the latent tile `tile` into slot X: fp32 residual rows, |x|^2, fp16 operand in tensor
    // memory, rounding residue
    auto load_tile = [&](int X, int tile) {
      float* rs = reinterpret_cast<float*>(smem + Sm::rs + X * kRsBytes);
      unsigned char* ms = smem + Sm::misc + X * Sm::m_size;
      const int64_t n = int64_t(tile) * p.tf + f;
      const bool valid = f < p.tf && n < p.N;     // lanes beyond the tile's frames carry zeros and write nothing
      const int64_t xb = valid ? p.fa.base(n) : 0;
      float xsum = 0.f, e2 = 0.f;
      // the warp's 64 lines (one per dim: 32 consecutive frames x 4 B) are asked for at once, then read 16 dims at a time
      if (valid) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x + p.fa.base(n - lane) + int64_t(h * 64 + lane) * p.fa.sxd));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x + p.fa.base(n - lane) + int64_t(h * 64 + 32 + lane) * p.fa.sxd));
      }
      #pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        float v[16];
        const int d0 = h * 64 + g * 16;
        const float* xp = p.x + xb + int64_t(d0) * p.fa.sxd;
        #pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = valid ? __ldg(xp + int64_t(j) * p.fa.sxd) : 0.f;
        uint32_t w[8];
        #pragma unroll
        for (int j = 0; j < 16; j += 4) {
          *reinterpret_cast<float4*>(rs + rs_off(f, (d0 + j) >> 2)) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          xsum = fmaf(v[j], v[j], xsum); xsum = fmaf(v[j + 1], v[j + 1], xsum);
          xsum = fmaf(v[j + 2], v[j + 2], xsum); xsum = fmaf(v[j + 3], v[j + 3], xsum);
          e2 = residue2(v[j], v[j + 1], w[j / 2], e2);
          e2 = residue2(v[j + 2], v[j + 3], w[j / 2 + 1], e2);
        }
        ptx::tmem_st8(tq + 64 * X + d0 / 2, w);
      }
      // |x|^2 of this half of the dims goes where the (not yet written) class / batch masks of the tile's first stage live
      reinterpret_cast<float*>(ms + (h ? Sm::m_bmask : Sm::m_cmask))[f] = xsum;
      reinterpret_cast<float*>(ms + Sm::m_dr2)[h * kM + f] = e2;
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->a_ready[X]));
    };
    // n = -1 is the prologue: it only loads the first tile of each slot (one call site for the tile load)
    for (int n = -1; n < steps0; ++n) {
      for (int X = 0; X < 2; ++X) {
        if (X == 1 && (n < 0 ? steps1 == 0 : n >= steps1)) break;      // slot 1: prologue only if it has a tile, then steps while n < steps1
        RVQ_TICK0();
        int next_tile = -1;
        if (n < 0) next_tile = tile0 + X;
        else {
          const int jt = n / p.n_q, s = n - jt * p.n_q;
          const int st = p.stage0 + s;
          const int64_t tile_n0 = int64_t(tile0 + X + 2 * jt) * p.tf;
          float* rs = reinterpret_cast<float*>(smem + Sm::rs + X * kRsBytes);
          unsigned char* ms = smem + Sm::misc + X * Sm::m_size;
          // winners / candidate lists of this step: one warp polls the mbarrier, the others block on a hardware barrier
          // (a blocked warp costs no issue slots, a polling one does)
          if (u == 0) ptx::mbar_wait(ptx::smem_u32(&bars->cand_ready[X]), uint32_t(n) & 1);
          ptx::named_bar_sync(8, kUpdWarps * 32);
          RVQ_TICK(t_wait);
          update_pass<TRAIN>(p, rs, ms, u, lane, s, rot, nchunks, tile_n0, pv.tab32(st), pv.cnorm(st), tsub);
          RVQ_TICK(t_upd);
          ptx::named_bar_sync(6, kUpdWarps * 32);  // every frame of the tile has its new residual (re-scores run on any warp)
          RVQ_TICK(t_bar);
          if (threadIdx.x == 128) { int* qc = reinterpret_cast<int*>(ms + Sm::m_qcnt); qc[0] = 0; qc[1] = 0; }
          const bool last = s + 1 == p.n_q;
          float sq = 0.f;
          if (!last || (TRAIN && p.sqerr != nullptr)) operand_pass<TRAIN>(rs, ms, f, h, tq + 64 * X, !last, sq);
          RVQ_TICK(t_tr);
          if (!last) {
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->a_ready[X]));
            RVQ_TICK(t_stw);
          } else {
            if (TRAIN && p.residual_out != nullptr) {
              // each warp writes 16 frames, 512 contiguous bytes per frame
              for (int i = 0; i < 16; ++i) {
                const int fo = q * 32 + h * 16 + i;
                const int64_t nn = tile_n0 + fo;
                if (fo < p.tf && nn < p.N) *reinterpret_cast<float4*>(p.residual_out + nn * 128 + lane * 4) = *reinterpret_cast<const float4*>(rs + rs_off(fo, lane));
              }
            }
            if ((jt + 1) * p.n_q < (X ? steps1 : steps0)) {
              next_tile = tile0 + X + 2 * (jt + 1);
              ptx::named_bar_sync(7, kUpdWarps * 32);   // the other warp of this quadrant may still read these rows
            }
          }
          if (TRAIN && p.sqerr != nullptr) {
            if (!(f < p.tf && tile_n0 + f < p.N)) sq = 0.f;
            sq = warp_sum(sq);
            if (lane == 0) atomicAdd(&p.sqerr[s], (double)sq);
          }
        }
        if (next_tile >= 0) { load_tile(X, next_tile); RVQ_TICK(t_load); }
      }
    }
#ifdef RVQ_TC_TIMERS
    if (lane == 0 && p.counters != nullptr) {
      atomicAdd(&p.counters[15], (unsigned long long)t_wait); atomicAdd(&p.counters[7], (unsigned long long)t_upd);
      atomicAdd(&p.counters[16], (unsigned long long)t_tr);   atomicAdd(&p.counters[8], (unsigned long long)t_load);
      atomicAdd(&p.counters[20], (unsigned long long)tsub[0]); atomicAdd(&p.counters[21], (unsigned long long)tsub[1]);
      atomicAdd(&p.counters[22], (unsigned long long)tsub[2]); atomicAdd(&p.counters[23], (unsigned long long)tsub[3]);
      atomicAdd(&p.counters[17], (unsigned long long)t_stw); atomicAdd(&p.counters[18], (unsigned long long)t_bar);
    }
#endif
  } else {
    ptx::reg_inc<160>();
    // ===== score warps =====
    const int q = warp;                        // TMEM lane quadrant = frames 32q..32q+31 of a tile
    const int f = q * 32 + lane;
    const uint32_t tlane = tmem + (uint32_t(q * 32) << 16);
    uint32_t n_cert = 0, n_resc = 0, n_full = 0;                        // search statistics (rvq_search_stats)
#ifdef RVQ_TC_TIMERS
    uint32_t t_wait = 0, t_epi = 0, t_win = 0;
    const long long t_begin = clock64();
#endif
    float xx_0 = 0.f, xx_1 = 0.f;              // upper bound of |r|^2 of this thread's frame in slot 0 / 1
    uint32_t acc_it = 0;
    for (int n = 0; n < steps0; ++n) {
      for (int X = 0; X < 2; ++X) {
        if (X == 1 && n >= steps1) break;
        const int jt = n / p.n_q, s = n - jt * p.n_q;
        const int st = p.stage0 + s;
        const int64_t nfr = int64_t(tile0 + X + 2 * jt) * p.tf + f;
        const float* rs = reinterpret_cast<const float*>(smem + Sm::rs + X * kRsBytes);
        unsigned char* ms = smem + Sm::misc + X * Sm::m_size;
        const float* t32 = pv.tab32(st);
        const float* cn = pv.cnorm(st);
        const StageMeta* meta = pv.meta(st);
        RVQ_TICK0();
        // ---- scores: per-class and per-batch minima of the K approximate scores of this frame ----
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
        // |x|^2 of a new tile and the rounding residue of this frame's operand were written by the update warps; the
        // scores above could only exist after they had finished
        float xx = X ? xx_1 : xx_0;
        if (s == 0) xx = reinterpret_cast<const float*>(ms + Sm::m_cmask)[f] + reinterpret_cast<const float*>(ms + Sm::m_bmask)[f];
        const float xnorm = sqrtf(xx);
        const bool outl = !(xnorm < meta->xlimit);      // also true for NaN
        const float drn = sqrtf(reinterpret_cast<const float*>(ms + Sm::m_dr2)[f] + reinterpret_cast<const float*>(ms + Sm::m_dr2)[kM + f]) * 1.001f;
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
        *reinterpret_cast<int4*>(ms + Sm::m_cand + f * 16) = cd;
        *reinterpret_cast<int*>(ms + Sm::m_ncnt + f * 4) = full ? kFull : (ncand > 4 ? kBig : ncand);
        *reinterpret_cast<uint32_t*>(ms + Sm::m_cmask + f * 4) = cmask;
        *reinterpret_cast<uint32_t*>(ms + Sm::m_bmask + f * 4) = bmask;
        int64_t* code_out = (f < p.tf && nfr < p.N) ? p.codes + int64_t(s) * p.N + nfr : nullptr;
        if (!full && ncand == 1 && code_out != nullptr) *code_out = cd.x;      // certified: the warp's codes are one 256-byte run
        {
          int* qc = reinterpret_cast<int*>(ms + Sm::m_qcnt);
          if (!full && ncand > 1) {
            if (ncand <= 4) ms[Sm::m_slowq + atomicAdd(&qc[0], 1)] = (unsigned char)f;
            else            ms[Sm::m_wideq + atomicAdd(&qc[1], 1)] = (unsigned char)f;
          }
          // frames outside the fp16 image's validity range (or NaN): exact scan right here, then they are certified
          uint32_t fm = __ballot_sync(0xffffffffu, full);
          while (fm) {
            const int i = __ffs(fm) - 1; fm &= fm - 1;
            int64_t* co = reinterpret_cast<int64_t*>(__shfl_sync(0xffffffffu, (unsigned long long)code_out, i));
            resolve_full(rs, ms, q * 32 + i, lane, p.K, t32, cn, co);
          }
        }
        n_full += full ? 1u : 0u; n_cert += (!full && ncand == 1) ? 1u : 0u; n_resc += (!full && ncand > 1) ? 1u : 0u;
        // upper bound of the next residual's |r|^2 (only the margin and the validity test use it):
        // the winner's approximate score is <= m + delta and off by <= delta/2
        if (full) { const float g2 = xnorm + meta->cmax_all; xx = g2 * g2; }
        else xx = fmaxf(xx + m + 1.5f * delta, 0.f) * 1.00001f + 1e-30f;
        if (X) xx_1 = xx; else xx_0 = xx;
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->cand_ready[X]));    // winners and queues visible to the update warps
        RVQ_TICK(t_win);
      }
    }
    // search statistics (evidence; see rvq_search_stats)
    if (p.counters != nullptr) {
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
        atomicAdd(&p.counters[6], (unsigned long long)t_win);
        atomicAdd(&p.counters[9], (unsigned long long)(clock64() - t_begin));
        atomicAdd(&p.counters[10], 1ull);
#endif
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 13) ptx::tmem_dealloc(tmem, 512);
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
    RVQ_CUDA(cudaFuncSetAttribute(tc_encode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Sm::total));
    RVQ_CUDA(cudaFuncSetAttribute(tc_encode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Sm::total));
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
  // Two tiles are in flight per CTA, so every CTA should own an even number of tiles of equal size: with `passes`
  // rounds of 2 tiles per CTA, a tile gets ceil(N / (2 * SMs * passes)) <= 128 frames (cfg2: 48 000 frames -> 2 rounds
  // of 82-frame tiles instead of 2.5 tiles of 128 with one of them running alone).
  const int64_t per_round = 2ll * sm_count * kM;
  const int64_t passes = (N + per_round - 1) / per_round;
  int64_t tf = (N + 2ll * sm_count * passes - 1) / (2ll * sm_count * passes);
  tf = tf < 16 ? 16 : (tf > kM ? kM : tf);
  if (const char* e = getenv("RVQ_TC_TILE_FRAMES")) { const int v = atoi(e); if (v >= 1 && v <= kM) tf = v; }   // tuning knob
  p.tf = int(tf);
  const int64_t ntiles = (N + tf - 1) / tf;
  const int64_t want = (ntiles + 1) / 2;
  const unsigned grid = unsigned(want < sm_count ? want : sm_count);
  // the lean variant serves plain encodes; straight-through arithmetic, loss numerators and the residual output
  // live in the other one (a stage's hot code has to fit the instruction cache)
  if (p.ste || p.sqerr != nullptr || p.residual_out != nullptr) tc_encode_kernel<true><<<grid, kThreadsTc, Sm::total, st>>>(p);
  else tc_encode_kernel<false><<<grid, kThreadsTc, Sm::total, st>>>(p);
  RVQ_LAUNCH_CHECK("tc_encode_kernel");
  if (a.quantized != nullptr)
    return simt_quant_sum(a.pack, a.K, a.D, a.x, p.fa, N, a.T, a.stage0, a.n_q, a.codes, a.quantized, p.ste,
                          (a.flags & RVQ_FLAG_ACCUM_Q) ? 1 : 0, st);
  return RVQ_OK;
}

}  // namespace rvq

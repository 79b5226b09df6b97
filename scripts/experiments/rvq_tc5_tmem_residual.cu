// Fused multi-stage nearest-code search on the 5th-generation tensor cores (tcgen05 / TMEM), two tiles in flight,
// residuals resident in tensor memory.
//
// One persistent CTA per SM walks a contiguous range of 128-frame tiles, TWO AT A TIME (slot 0 takes the even
// tiles of the range, slot 1 the odd ones).  A tile's stage is a serial chain
//     MMA (scores of the 128 frames against all K codes) -> score reduction -> winner -> residual update -> next MMA,
// so the two slots run half a stage apart: while the tensor pipe computes the scores of one slot, the SIMT warps
// choose the winners of / update the other one.  The fp32 residual of every frame stays ON CHIP across all n_q
// stages (core_vq.py:357-367 without the per-stage round trips through HBM): in tensor memory, next to its fp16 copy.
//
//   tensor memory (512 columns x 128 lanes, lane = frame of a tile):
//            [0,128)    two accumulator buffers of 64 columns (one half of a 128-code chunk each)
//            [128,256)  fp16 operand A of slot 0 / slot 1 (64 columns each)
//            [256,512)  fp32 residual R of slot 0 / slot 1 (128 columns each; dims permuted inside blocks of 16 so that
//                       the 16x256b access shape hands a group of 4 lanes the same dims as its fp16 operand cells)
//   scores   S[f,k] = -2 r_f . c_k + |c_k|^2 from tcgen05.mma kind::f16 (fp32 accumulation in tensor memory),
//            M=128 frames, N=64 codes per accumulator, 9 K-steps of 16:
//            A = fp16(r) in tensor memory (K-steps 0..7) + one constant shared-memory block whose columns 128,129 = 1
//                (K-step 8: picks up the hi/lo halves of |c|^2);
//            B = fp16 image of the codebook (cols 0..127 = -2c, cols 128,129 = hi/lo of |c|^2), streamed by the
//                TMA engine (cp.async.bulk) from the L2-resident pack into a ring of third-of-a-chunk slots
//                (6 K-groups x 128 codes = 12 KB each); both halves of a chunk read the same three slots.
//   warps    0..3   score warps  (thread = TMEM lane = frame): tcgen05.ld of 16-column pairs, per-class minima in
//                   registers, per-batch minima to shared memory, then the certified winner or the class/batch masks
//                   of the candidates (see below), codes of certified frames;
//            4..11  update warps, each owns 16 frames: (1) winner rows of its certified frames and the candidate rows of
//                   its listed frames are requested with async copies (cp.async, lane = 16-byte chunk of a row: whole
//                   512-byte rows per instruction, no registers held in flight) into a shared-memory staging tile /
//                   pool; (2) the residual rows come from tensor memory (tcgen05.ld.16x256b); (3) listed frames are
//                   settled by the warp from shared memory (exact fp32 re-score), the winning row is moved into the
//                   frame's staging row; (4) r - q in registers -> fp16 operand of the next stage to tensor memory
//                   -> a_ready; (5) off the chain: residual back to tensor memory, exact rounding residue of the
//                   operand, squared error -> dr_ready; tile loads;
//            12  TMA producer;  13  MMA issuer (owns the TMEM allocation);  14, 15 idle.
//            setmaxnreg: 152 registers for the score warps, 160 for the update warps, 40 for the last warpgroup.
//   sync     mbarriers only between roles: a_ready[slot] (update -> MMA), acc_full/acc_empty (MMA <-> score),
//            cand_ready[slot] (score -> update), dr_ready[slot] (update -> score), full/empty (TMA <-> MMA).
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
constexpr int kN = kTcChunkCodes;       // 128 codes per chunk of the image
constexpr int kNh = 64;                 // codes per accumulator (UMMA N): half a chunk
constexpr int kRing = 7;                // B ring slots; each holds one K-third of a chunk (6 K-groups = 3 K-steps)
constexpr int kSlotBytes = 6 * kTcLBO;  // 12288 B
constexpr int kTmemA = 128;             // first TMEM column of the fp16 operands (64 columns per slot)
constexpr int kTmemR = 256;             // first TMEM column of the fp32 residuals (128 columns per slot)
constexpr int kThreadsTc = 16 * 32;
constexpr int kUpdWarps = 8;
constexpr int kBig = 5;                 // ncnt marker: more than 4 candidates (enumerate the masks)
constexpr int kFull = 6;                // ncnt marker: exact scan of the whole table
constexpr int kPoolRows = 48;           // staged candidate rows of listed frames (512 B each)
constexpr int kNoPool = 255;            // pool exhausted: the frame's candidate rows come through registers

struct Sm {
  static constexpr uint32_t aug = 0;                               // [2 k-groups][128 rows][16 B], no swizzle
  static constexpr uint32_t ring = aug + 4096;
  static constexpr uint32_t stage = ring + kRing * kSlotBytes;     // fp32 [128 f][128 d]: winner rows of the slot being updated, chunk-swizzled
  static constexpr uint32_t pool = stage + kM * 512;               // fp32 [kPoolRows][128 d]: candidate rows of listed frames
  static constexpr uint32_t pnorm = pool + kPoolRows * 512;        // float [kPoolRows]: their |c|^2
  static constexpr uint32_t scratch = pnorm + 256;                 // fp32 [12 warps][128 d]: one residual row per warp
  static constexpr uint32_t bmn = scratch + 12 * 512;              // float2 [16 half-chunks][128 f]: per-batch minima of the stage being scored
  static constexpr uint32_t misc = bmn + 16 * kM * 8;              // 2 x per-slot block (offsets m_*)
  static constexpr uint32_t m_cand = 0;                            // int4 [128]: winner (x) / candidate codes (-1 = none)
  static constexpr uint32_t m_ncnt = m_cand + kM * 16;             // int [128]
  static constexpr uint32_t m_cmask = m_ncnt + kM * 4;             // u32 [128] flagged classes   (a fresh tile: |x|^2 of dims 0..63)
  static constexpr uint32_t m_bmask = m_cmask + kM * 4;            // u32 [128] flagged batches   (a fresh tile: |x|^2 of dims 64..127)
  static constexpr uint32_t m_dr2 = m_bmask + kM * 4;              // float [2][128]: |r - fp16(r)|^2 of the current operand, per half of the dims
  static constexpr uint32_t m_pbase = m_dr2 + 2 * kM * 4;          // u8 [128]: first pool row of a listed frame (kNoPool = none)
  static constexpr uint32_t m_size = m_pbase + kM;
  static constexpr uint32_t bars = misc + 2 * m_size;
  static constexpr uint32_t total = bars + 256;
};
struct Bars {
  uint64_t full[kRing], empty[kRing], acc_full[2], acc_empty[2], a_ready[2], cand_ready[2], dr_ready[2];
  uint32_t tmem_base;
  int pool_cnt;
};
static_assert(sizeof(Bars) <= 256, "barrier block");
// shared-window address of a barrier, from the CTA's window base (no generic->shared conversion inside the hot loops)
#define RVQ_BAR(field, i) (sbase + Sm::bars + uint32_t(offsetof(Bars, field)) + 8u * uint32_t(i))
static_assert(Sm::total <= 227 * 1024, "shared memory budget");
static_assert(kTcKPad / 16 == 9 && kN == 128, "operand geometry");
static_assert((Sm::m_size % 16) == 0 && (Sm::misc % 16) == 0 && (Sm::stage % 1024) == 0, "alignment");

// debug timeline of CTA 0 (RVQ_TC_TRACE builds; slots 0/1, steps kTraceN0 .. kTraceN0 + kTraceSteps - 1; -DRVQ_TRACE_N0=36
// looks at the CTA's third tile, which runs alone): g_trace[X][step][event] = cycles since kernel start
#ifndef RVQ_TRACE_N0
#define RVQ_TRACE_N0 2
#endif
constexpr int kTraceSteps = 6, kTraceEv = 16;
[[maybe_unused]] constexpr int kTraceN0 = RVQ_TRACE_N0;
__device__ long long g_trace[2 * kTraceSteps * kTraceEv + 128 + 2 * 64];   // + update-pass detail for step N0+2
#ifdef RVQ_TC_TRACE
#define RVQ_TRACE(X, n, ev, cond) do { if (blockIdx.x == 0 && (cond) && (n) >= kTraceN0 && (n) < kTraceN0 + kTraceSteps) { \
    asm volatile("" ::: "memory"); g_trace[(((X) * kTraceSteps) + (n) - kTraceN0) * kTraceEv + (ev)] = clock64() - t_kernel0; asm volatile("" ::: "memory"); } } while (0)
// update-pass detail (slot X, step N0+2): g_trace[base + 128 + 64 X + 8 u + e], lane 0 of update warp u
#define RVQ_TRACE3(X, n, u, e) do { if (blockIdx.x == 0 && (n) == kTraceN0 + 2 && (threadIdx.x & 31) == 0) { \
    asm volatile("" ::: "memory"); g_trace[2 * kTraceSteps * kTraceEv + 128 + 64 * (X) + 8 * (u) + (e)] = clock64() - t_kernel0; asm volatile("" ::: "memory"); } } while (0)
#else
#define RVQ_TRACE(X, n, ev, cond) do { } while (0)
#define RVQ_TRACE3(X, n, u, e) do { } while (0)
#endif

struct TcParams {
  const unsigned char* pack; int K;
  const float* x; FrameAddr fa; int64_t N;
  int stage0, n_q;
  int64_t* codes; float* residual_out; double* sqerr;
  int ste;
  int tf;                  // frames per tile (<= 128)
  unsigned long long* counters;
};

__device__ __forceinline__ float inf_f() { return __int_as_float(0x7f800000); }
// staging rows: 16-byte chunk ch (dims 4ch..4ch+3) of frame f is XOR-swizzled with the frame number (its low three bits
// reversed) so that both access patterns of the kernel spread over the banks: 32 lanes = the 32 chunks of one frame;
// 8 groups of 4 lanes = 8 consecutive frames x 4 consecutive chunks
__device__ __forceinline__ int rs_swz(int f) { return (f & 24) | ((f & 1) << 2) | (f & 2) | ((f >> 2) & 1); }
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
// code of (batch a in this CTA's processing order, class j): batch = 32 consecutive codes, class = code mod 32
__device__ __forceinline__ int code_of(int a, int j, int rot, int nchunks) {
  int pc = (a >> 2) + rot; pc = pc < nchunks ? pc : pc - nchunks;
  return pc * 128 + (a & 3) * 32 + j;
}
// squared rounding residue of two floats against their fp16 pair
__device__ __forceinline__ float residue2(float a, float b, uint32_t& word, float e2) {
  const __half2 h = __floats2half2_rn(a, b);
  word = *reinterpret_cast<const uint32_t*>(&h);
  const float2 bk = __half22float2(h);
  const float ea = a - bk.x, eb = b - bk.y;
  return fmaf(eb, eb, fmaf(ea, ea, e2));
}
// squared rounding residue of two floats against their (already packed) fp16 pair
__device__ __forceinline__ float residue_of(float a, float b, uint32_t word, float e2) {
  const float2 bk = __half22float2(*reinterpret_cast<const __half2*>(&word));
  const float ea = a - bk.x, eb = b - bk.y;
  return fmaf(eb, eb, fmaf(ea, ea, e2));
}
// column of dim 16 i + j inside the residual block i of tensor memory: j = 4m + e sits at 8 (e >> 1) + 2m + (e & 1), so that
// the 16x256b access shape gives thread (g, m) of a warp the dims 16i + 4m .. 16i + 4m + 3 of frames g and g + 8
__host__ __device__ constexpr int perm16(int j) { return 8 * ((j & 3) >> 1) + 2 * (j >> 2) + (j & 1); }

// tiles [start, start+cnt) of this CTA
__device__ __forceinline__ void cta_range(int ntiles, int& start, int& cnt) {
  const int base = ntiles / int(gridDim.x), rem = ntiles % int(gridDim.x);
  const int b = blockIdx.x;
  start = b * base + (b < rem ? b : rem);
  cnt = base + (b < rem ? 1 : 0);
}

// minima of one pair of 16-column reads (the same 16 classes of two batches): running class minima, and the minimum of
// each half-batch
__device__ __forceinline__ float min16(const uint32_t (&v)[16]) {
  float t[5];
  #pragma unroll
  for (int j = 0; j < 5; ++j) t[j] = ptx::fmin3(__uint_as_float(v[3 * j]), __uint_as_float(v[3 * j + 1]), __uint_as_float(v[3 * j + 2]));
  return fminf(ptx::fmin3(t[0], t[1], t[2]), ptx::fmin3(t[3], t[4], __uint_as_float(v[15])));
}
__device__ __forceinline__ void pair_min(const uint32_t (&u)[16], const uint32_t (&v)[16], float* cm16, float& bu, float& bv) {
  #pragma unroll
  for (int j = 0; j < 16; ++j) cm16[j] = ptx::fmin3(cm16[j], __uint_as_float(u[j]), __uint_as_float(v[j]));
  bu = min16(u); bv = min16(v);
}

// torch's CPU argmax (core_vq.py:188) propagates NaN: the first NaN distance wins; otherwise the smallest
// distance, lowest index on ties.  (best, bcode) starts as (+inf, 0x7fffffff).
__device__ __forceinline__ bool nan_aware_better(float dist, int code, float best, int bcode) {
  if (dist != dist) return best == best || code < bcode;
  return best == best && (dist < best || (dist == best && code < bcode));
}

// Score-warp side.  Frames whose candidate set is the whole table (outside the fp16 image's validity range, NaN):
// the whole warp scores the table, one code per lane, against the frame's residual row `row` (plain fp32 [128] in
// shared memory), stores the code and rewrites the frame's entry as a certified winner for the update warps.
__device__ __forceinline__ void resolve_full(const float* row, unsigned char* ms, int f, int lane, int K, const float* __restrict__ t32,
                                             const float* __restrict__ cn, int64_t* code_out) {
  const float4 rl = *reinterpret_cast<const float4*>(row + 4 * lane);
  const float rr = warp_sum(dot4(rl, rl, 0.f));
  float best = inf_f(); int bcode = 0x7fffffff;
  for (int code = lane; code < K; code += 32) {
    const float4* rp = reinterpret_cast<const float4*>(t32 + size_t(code) * 128);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    #pragma unroll 2
    for (int ch = 0; ch < 32; ch += 4) {
      a0 = dot4(*reinterpret_cast<const float4*>(row + 4 * (ch + 0)), __ldg(rp + ch + 0), a0);
      a1 = dot4(*reinterpret_cast<const float4*>(row + 4 * (ch + 1)), __ldg(rp + ch + 1), a1);
      a2 = dot4(*reinterpret_cast<const float4*>(row + 4 * (ch + 2)), __ldg(rp + ch + 2), a2);
      a3 = dot4(*reinterpret_cast<const float4*>(row + 4 * (ch + 3)), __ldg(rp + ch + 3), a3);
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

// new residual n = r - q (core_vq.py:364 / :348; straight-through arithmetic of :309 in training)
template <bool TRAIN>
__device__ __forceinline__ float4 sub_row(const TcParams& p, const float4& rv, float4 q) {
  if (TRAIN && p.ste) { q.x = rv.x + (q.x - rv.x); q.y = rv.y + (q.y - rv.y); q.z = rv.z + (q.z - rv.z); q.w = rv.w + (q.w - rv.w); }
  return make_float4(rv.x - q.x, rv.y - q.y, rv.z - q.z, rv.w - q.w);
}

// A frame whose flagged batches x flagged classes give more than 4 candidates: the whole warp works on it, one candidate per
// quarter-warp and step (8 lanes x 4 chunks of 16 B per row), against the residual row `row` (plain fp32 [128] in shared
// memory).  Returns the winning code (exact fp32 distances, core_vq.py:183-187, lowest code on ties).
__device__ __forceinline__ int resolve_wide(const float* row, const unsigned char* ms, int f, int lane, int rot, int nchunks,
                                            const float* __restrict__ t32, const float* __restrict__ cn) {
  const int qq = lane >> 3, j = lane & 7;
  const uint32_t cm = *reinterpret_cast<const uint32_t*>(ms + Sm::m_cmask + f * 4);
  const uint32_t bm = *reinterpret_cast<const uint32_t*>(ms + Sm::m_bmask + f * 4);
  float4 r[4];
  #pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = *reinterpret_cast<const float4*>(row + 4 * (8 * i + j));
  float rr = (dot4(r[0], r[0], 0.f) + dot4(r[1], r[1], 0.f)) + (dot4(r[2], r[2], 0.f) + dot4(r[3], r[3], 0.f));
  #pragma unroll
  for (int off = 4; off > 0; off >>= 1) rr += __shfl_xor_sync(0xffffffffu, rr, off);
  float best = inf_f(); int bcode = 0x7fffffff;
  // candidates in enumeration order (batches outer, classes inner); quarter qq takes candidates qq, qq + 4, ...
  const int nc = __popc(cm), ntot = nc * __popc(bm);
  #pragma unroll 1
  for (int base = 0; base < ntot; base += 4) {
    const int idx = base + qq;
    int code = -1;
    if (idx < ntot) {
      const int ib = idx / nc, ic = idx - ib * nc;
      uint32_t bmq = bm, cmq = cm;
      for (int i = 0; i < ib; ++i) bmq &= bmq - 1;
      for (int i = 0; i < ic; ++i) cmq &= cmq - 1;
      code = code_of(__ffs(bmq) - 1, __ffs(cmq) - 1, rot, nchunks);
    }
    float d = 0.f;
    if (code >= 0) {
      const float4* rp = reinterpret_cast<const float4*>(t32 + size_t(code) * 128);
      d = (dot4(r[0], __ldg(rp + j), 0.f) + dot4(r[1], __ldg(rp + 8 + j), 0.f)) + (dot4(r[2], __ldg(rp + 16 + j), 0.f) + dot4(r[3], __ldg(rp + 24 + j), 0.f));
    }
    #pragma unroll
    for (int off = 4; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
    if (code >= 0) {
      const float e = (rr - 2.f * d) + __ldg(cn + code);
      if (e < best || (e == best && code < bcode)) { best = e; bcode = code; }
    }
  }
  #pragma unroll
  for (int off = 8; off <= 16; off <<= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, off);
    const int oc = __shfl_xor_sync(0xffffffffu, bcode, off);
    if (ob < best || (ob == best && oc < bcode)) { best = ob; bcode = oc; }
  }
  if (bcode == 0x7fffffff) bcode = code_of(__ffs(bm) - 1, __ffs(cm) - 1, rot, nchunks);   // NaN distances only: the first candidate
  return bcode;
}

// The residual update of one (slot, stage), executed by the eight update warps; warp u owns the 16 frames f0 .. f0 + 15
// (TMEM lanes 32 (u & 3) + 16 (u >> 2) ..).  A group of 4 lanes (g = lane / 4, m = lane % 4) holds frames g and g + 8 of them:
// lane m the 16-byte chunks 4i + m (i = 0..7) of each row, which is both what the 16x256b tensor-memory shape delivers of the
// residual (perm16) and what it takes for the fp16 operand.
template <bool TRAIN>
__device__ __forceinline__ void update_pass(const TcParams& p, unsigned char* smem, uint32_t sbase, unsigned char* ms, int u, int lane, int s,
                                            int rot, int nchunks, int64_t tile_n0, const float* __restrict__ t32,
                                            const float* __restrict__ cn, uint32_t taddr_r, uint32_t taddr_a, bool store, bool last,
                                            uint32_t bar_a, float& sq, int* pool_cnt, int trX, int trn, long long t_kernel0) {
  RVQ_TRACE3(trX, trn, u, 0);
  const int4* cand = reinterpret_cast<const int4*>(ms + Sm::m_cand);
  const int* ncnt = reinterpret_cast<const int*>(ms + Sm::m_ncnt);
  unsigned char* pbase = ms + Sm::m_pbase;
  const int g = lane >> 2, m = lane & 3;
  const int f0 = (u & 3) * 32 + (u >> 2) * 16;
  const int fA = f0 + g, fB = fA + 8;
  const uint32_t stage_s = sbase + Sm::stage, pool_s = sbase + Sm::pool;
  float* scr = reinterpret_cast<float*>(smem + Sm::scratch + (4 + u) * 512);
  // ---- 1. requests.  ncnt == 1: certified (or settled by an exact scan) -> the winner's row goes to the frame's staging row;
  //         ncnt 2..4: candidate list -> rows to the pool; ncnt kBig: wide set -> settled below through registers ----
  const int fw = f0 + (lane & 15);
  const int nw = ncnt[fw];
  const uint32_t listm = __ballot_sync(0xffffffffu, lane < 16 && nw >= 2 && nw <= 4);
  const uint32_t widem = __ballot_sync(0xffffffffu, lane < 16 && nw == kBig);
  {
    const int codew = cand[fw].x;
    #pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int nk = __shfl_sync(0xffffffffu, nw, k), code = __shfl_sync(0xffffffffu, codew, k);
      const int f = f0 + k;
      if (nk == 1) ptx::cp_async16(stage_s + uint32_t(f) * 512u + (uint32_t(lane ^ rs_swz(f)) << 4), reinterpret_cast<const float4*>(t32 + size_t(code) * 128) + lane);
    }
  }
  #pragma unroll 1
  for (uint32_t lm = listm; lm; lm &= lm - 1) {
    const int f = f0 + __ffs(lm) - 1;
    // the 2..4 candidates = flagged batches x flagged classes (warp-uniform enumeration; bit a of the batch mask is the
    // a-th batch in this CTA's processing order)
    int4 cd = make_int4(-1, -1, -1, -1);
    const uint32_t cmk = *reinterpret_cast<const uint32_t*>(ms + Sm::m_cmask + f * 4);
    uint32_t bm2 = *reinterpret_cast<const uint32_t*>(ms + Sm::m_bmask + f * 4);
    int w = 0;
    while (bm2) {
      const int a = __ffs(bm2) - 1; bm2 &= bm2 - 1;
      uint32_t cm2 = cmk;
      while (cm2) {
        const int code = code_of(a, __ffs(cm2) - 1, rot, nchunks); cm2 &= cm2 - 1;
        if (w == 0) cd.x = code; else if (w == 1) cd.y = code; else if (w == 2) cd.z = code; else cd.w = code;
        ++w;
      }
    }
    int base = 0;
    if (lane == 0) base = atomicAdd(pool_cnt, w);
    base = __shfl_sync(0xffffffffu, base, 0);
    const bool fits = base + w <= kPoolRows;
    if (fits) {
      const uint32_t dst = pool_s + uint32_t(base) * 512u + uint32_t(lane) * 16u;
      ptx::cp_async16(dst, reinterpret_cast<const float4*>(t32 + size_t(cd.x) * 128) + lane);
      ptx::cp_async16(dst + 512, reinterpret_cast<const float4*>(t32 + size_t(cd.y) * 128) + lane);
      if (cd.z >= 0) ptx::cp_async16(dst + 1024, reinterpret_cast<const float4*>(t32 + size_t(cd.z) * 128) + lane);
      if (cd.w >= 0) ptx::cp_async16(dst + 1536, reinterpret_cast<const float4*>(t32 + size_t(cd.w) * 128) + lane);
      const int ck = lane == 0 ? cd.x : lane == 1 ? cd.y : lane == 2 ? cd.z : cd.w;
      if (lane < 4 && ck >= 0) ptx::cp_async4(sbase + Sm::pnorm + uint32_t(base + lane) * 4u, cn + ck);
    }
    if (lane == 0) {
      *reinterpret_cast<int4*>(ms + Sm::m_cand + f * 16) = cd;
      pbase[f] = (unsigned char)(fits ? base : kNoPool);
    }
  }
  ptx::cp_async_commit();
  // ---- 2. residual rows of the 16 frames from tensor memory ----
  uint32_t r[64];
  ptx::tmem_ld_16x256b_x16(taddr_r, r);
  ptx::cp_async_wait_all();
  ptx::tmem_ld_wait();
  __syncwarp();
  RVQ_TRACE3(trX, trn, u, 1);
  // residual value (frame half hb, chunk i, element e) of this lane
#define RVQ_R(hb, i, e) r[4 * (2 * (i) + ((e) >> 1)) + 2 * (hb) + ((e) & 1)]
  // ---- 3. listed / wide frames: the frame's residual row goes to the warp's scratch row, the warp settles it (lane = 16-byte
  //         chunk) with exact fp32 distances (core_vq.py:183-187, lowest code on ties), the winner's row goes to the staging row ----
  if (listm | widem) {
    #pragma unroll 1
    for (uint32_t lm = listm | widem; lm; lm &= lm - 1) {
      const int L = __ffs(lm) - 1, f = f0 + L;
      if (g == (L & 7)) {
        if (L < 8) {
          #pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<uint4*>(scr + 4 * (4 * i + m)) = make_uint4(RVQ_R(0, i, 0), RVQ_R(0, i, 1), RVQ_R(0, i, 2), RVQ_R(0, i, 3));
        } else {
          #pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<uint4*>(scr + 4 * (4 * i + m)) = make_uint4(RVQ_R(1, i, 0), RVQ_R(1, i, 1), RVQ_R(1, i, 2), RVQ_R(1, i, 3));
        }
      }
      __syncwarp();
      int bcode;
      float4 wsel;
      if ((widem >> L) & 1) {
        bcode = resolve_wide(scr, ms, f, lane, rot, nchunks, t32, cn);
        wsel = __ldg(reinterpret_cast<const float4*>(t32 + size_t(bcode) * 128) + lane);
      } else {
        const int4 cd = cand[f];
        const int base = pbase[f];
        const float4 rl = *reinterpret_cast<const float4*>(scr + 4 * lane);
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 w0, w1, w2, w3;
        float n0, n1, n2, n3;
        if (base != kNoPool) {
          const float4* rows = reinterpret_cast<const float4*>(smem + Sm::pool + base * 512) + lane;
          const float* pn = reinterpret_cast<const float*>(smem + Sm::pnorm) + base;
          w0 = rows[0]; w1 = rows[32]; w2 = cd.z >= 0 ? rows[64] : z4; w3 = cd.w >= 0 ? rows[96] : z4;
          n0 = pn[0]; n1 = pn[1]; n2 = cd.z >= 0 ? pn[2] : 0.f; n3 = cd.w >= 0 ? pn[3] : 0.f;
        } else {
          w0 = __ldg(reinterpret_cast<const float4*>(t32 + size_t(cd.x) * 128) + lane);
          w1 = __ldg(reinterpret_cast<const float4*>(t32 + size_t(cd.y) * 128) + lane);
          w2 = cd.z >= 0 ? __ldg(reinterpret_cast<const float4*>(t32 + size_t(cd.z) * 128) + lane) : z4;
          w3 = cd.w >= 0 ? __ldg(reinterpret_cast<const float4*>(t32 + size_t(cd.w) * 128) + lane) : z4;
          n0 = __ldg(cn + cd.x); n1 = __ldg(cn + cd.y);
          n2 = cd.z >= 0 ? __ldg(cn + cd.z) : 0.f; n3 = cd.w >= 0 ? __ldg(cn + cd.w) : 0.f;
        }
        float rr = dot4(rl, rl, 0.f), d0 = dot4(rl, w0, 0.f), d1 = dot4(rl, w1, 0.f), d2 = dot4(rl, w2, 0.f), d3 = dot4(rl, w3, 0.f);
        #pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          rr += __shfl_xor_sync(0xffffffffu, rr, off);
          d0 += __shfl_xor_sync(0xffffffffu, d0, off); d1 += __shfl_xor_sync(0xffffffffu, d1, off);
          d2 += __shfl_xor_sync(0xffffffffu, d2, off); d3 += __shfl_xor_sync(0xffffffffu, d3, off);
        }
        float best = inf_f(); bcode = 0x7fffffff; wsel = w0;       // NaN distances only: the first candidate
        auto consider = [&](float d, float nrm, int code, const float4& w) {
          const float e = (rr - 2.f * d) + nrm;
          if (code >= 0 && (e < best || (e == best && code < bcode))) { best = e; bcode = code; wsel = w; }
        };
        consider(d0, n0, cd.x, w0); consider(d1, n1, cd.y, w1); consider(d2, n2, cd.z, w2); consider(d3, n3, cd.w, w3);
        if (bcode == 0x7fffffff) bcode = cd.x;
      }
      *reinterpret_cast<float4*>(smem + Sm::stage + f * 512 + ((lane ^ rs_swz(f)) << 4)) = wsel;
      const int64_t nfr = tile_n0 + f;
      if (lane == 0 && f < p.tf && nfr < p.N) p.codes[int64_t(s) * p.N + nfr] = bcode;
      __syncwarp();
    }
  }
  RVQ_TRACE3(trX, trn, u, 2);
  // ---- 4. phase 1 (on the chain to the next stage's MMA): n = r - q in registers, fp16 operand to tensor memory, a_ready ----
  uint32_t w[32];
  {
    #pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
      const int f = hb ? fB : fA;
      const int sw = rs_swz(f);
      const unsigned char* qbase = smem + Sm::stage + f * 512 + ((m ^ (sw & 3)) << 4);
      #pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 qv = *reinterpret_cast<const float4*>(qbase + ((i ^ (sw >> 2)) << 6));
        const float4 rv = make_float4(__uint_as_float(RVQ_R(hb, i, 0)), __uint_as_float(RVQ_R(hb, i, 1)), __uint_as_float(RVQ_R(hb, i, 2)),
                                      __uint_as_float(RVQ_R(hb, i, 3)));
        const float4 n = sub_row<TRAIN>(p, rv, qv);
        RVQ_R(hb, i, 0) = __float_as_uint(n.x); RVQ_R(hb, i, 1) = __float_as_uint(n.y);
        RVQ_R(hb, i, 2) = __float_as_uint(n.z); RVQ_R(hb, i, 3) = __float_as_uint(n.w);
        w[4 * i + 2 * hb] = pack_half2(n.x, n.y);
        w[4 * i + 2 * hb + 1] = pack_half2(n.z, n.w);
      }
    }
  }
  RVQ_TRACE3(trX, trn, u, 3);
  if (store) {
    ptx::tmem_st_16x256b_x8(taddr_a, w);
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(bar_a);
  }
  RVQ_TRACE3(trX, trn, u, 4);
  // ---- 5. phase 2 (off that chain): residual back to tensor memory, exact rounding residue of the operand, squared error;
  //         the caller publishes it on dr_ready, which the score warps await before their winner phase ----
  if (!last) ptx::tmem_st_16x256b_x16(taddr_r, r);
  float e2[2];
  #pragma unroll
  for (int hb = 0; hb < 2; ++hb) {
    const int f = hb ? fB : fA;
    float e[4] = {0.f, 0.f, 0.f, 0.f};
    float sqf = 0.f;
    #pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 n = make_float4(__uint_as_float(RVQ_R(hb, i, 0)), __uint_as_float(RVQ_R(hb, i, 1)), __uint_as_float(RVQ_R(hb, i, 2)),
                                   __uint_as_float(RVQ_R(hb, i, 3)));
      e[i & 3] = residue_of(n.x, n.y, w[4 * i + 2 * hb], e[i & 3]);
      e[i & 3] = residue_of(n.z, n.w, w[4 * i + 2 * hb + 1], e[i & 3]);
      if (TRAIN) sqf = dot4(n, n, sqf);
      if (TRAIN && last && p.residual_out != nullptr && f < p.tf && tile_n0 + f < p.N)
        *reinterpret_cast<float4*>(p.residual_out + (tile_n0 + f) * 128 + 4 * (4 * i + m)) = n;
    }
    e2[hb] = (e[0] + e[1]) + (e[2] + e[3]);
    if (TRAIN && f < p.tf && tile_n0 + f < p.N) sq += sqf;      // sum((q - r)^2) of core_vq.py:319 = |new residual|^2
  }
  // exact rounding residue of the new operand rows: sum over the 4 lanes of the group
  e2[0] += __shfl_xor_sync(0xffffffffu, e2[0], 1); e2[1] += __shfl_xor_sync(0xffffffffu, e2[1], 1);
  e2[0] += __shfl_xor_sync(0xffffffffu, e2[0], 2); e2[1] += __shfl_xor_sync(0xffffffffu, e2[1], 2);
  if (m == 0) {
    float* dr2 = reinterpret_cast<float*>(ms + Sm::m_dr2);
    dr2[fA] = e2[0]; dr2[kM + fA] = 0.f;
    dr2[fB] = e2[1]; dr2[kM + fB] = 0.f;
  }
  if (!last) ptx::tmem_st_wait();
#undef RVQ_R
  RVQ_TRACE3(trX, trn, u, 5);
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
    for (int i = 0; i < kRing; ++i) { ptx::mbar_init(RVQ_BAR(full, i), 1); ptx::mbar_init(RVQ_BAR(empty, i), 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(RVQ_BAR(acc_full, i), 1); ptx::mbar_init(RVQ_BAR(acc_empty, i), 4); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(RVQ_BAR(a_ready, i), kUpdWarps); ptx::mbar_init(RVQ_BAR(cand_ready, i), 4);
                                  ptx::mbar_init(RVQ_BAR(dr_ready, i), kUpdWarps); }
    bars->pool_cnt = 0;
    ptx::fence_mbar_init();
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
#ifdef RVQ_TC_TRACE
  __shared__ long long s_t0;
  if (threadIdx.x == 0) s_t0 = clock64();
  __syncthreads();
  const long long t_kernel0 = s_t0;
#endif

  if (warp >= 12) {
    ptx::reg_dec<40>();
    // opaque copy of the shared window base: barrier addresses are formed from a register, not re-derived per use
    uint32_t sb = sbase;
    asm volatile("" : "+r"(sb));
    const size_t stage_stride = pv.L.stride;
    if (warp == 12) {
      // ===== TMA producer: the three K-thirds of every 128-code chunk, in the global step order (warp-uniform loop, elected
      // issue: the copy's operands stay in uniform registers).  The slots come free in the order they were filled. =====
      uint32_t slot = 0, ph = 1;               // ring position of the next K-third, parity its `empty` barrier must have passed
      const unsigned char* img0 = pv.tc(p.stage0);
      for (int n = 0; n < steps0; ++n) {
        const int st = n % p.n_q;
        for (int X = 0; X < 2; ++X) {
          if (X == 1 && n >= steps1) break;
          const unsigned char* img = img0 + size_t(st) * stage_stride;
          int pc = rot;
          for (int c = 0; c < nchunks; ++c) {
            const unsigned char* src = img + size_t(pc) * kTcChunkBytes;
            if (++pc == nchunks) pc = 0;
            #pragma unroll 1
            for (int third = 0; third < 3; ++third) {
              ptx::mbar_wait(sb + Sm::bars + uint32_t(offsetof(Bars, empty)) + 8u * slot, ph);
              ptx::bulk_g2s_expect_w(sb + Sm::ring + slot * kSlotBytes, src + third * kSlotBytes, kSlotBytes,
                                     sb + Sm::bars + uint32_t(offsetof(Bars, full)) + 8u * slot);
              if (++slot == kRing) { slot = 0; ph ^= 1; }
            }
          }
        }
      }
      __syncwarp();
    } else if (warp == 13) {
      // ===== MMA issuer: a chunk is two accumulators of 64 codes; each takes 8 MMAs with A from tensor memory + 1 with the
      // constant shared-memory block, three per K-third.  The first half waits for the K-thirds as they land, the second half
      // frees them.  The loop is warp-uniform (every lane waits and forms the same descriptors, one elected lane issues), so the
      // operands of tcgen05.mma stay in uniform registers. =====
      constexpr uint32_t idesc = ptx::umma_idesc_f16_f32(kM, kNh);
      const uint32_t tmem_u = __reduce_max_sync(0xffffffffu, tmem);      // a provably warp-uniform copy (uniform register)
      const uint64_t ad_aug = ptx::umma_desc_kmajor_noswz(sb + Sm::aug, 2048, 128);
      const uint64_t bd0 = ptx::umma_desc_kmajor_noswz(sb + Sm::ring, kTcLBO, kTcSBO);
      const uint32_t bar_full0 = sb + Sm::bars + uint32_t(offsetof(Bars, full));
      const uint32_t bar_empty0 = sb + Sm::bars + uint32_t(offsetof(Bars, empty));
      const uint32_t bar_accf0 = sb + Sm::bars + uint32_t(offsetof(Bars, acc_full));
      const uint32_t bar_acce0 = sb + Sm::bars + uint32_t(offsetof(Bars, acc_empty));
      uint32_t slot = 0, sph = 0;              // ring position of the chunk's first K-third and the parity of its `full` barrier
      uint32_t aph = 1;                        // parity the accumulators' `empty` barriers must have passed
      static_assert(((2 * kTcLBO) >> 4) == 256, "B descriptor step of one K=16 MMA");
      constexpr uint32_t kHalfRows = (kNh * 16) >> 4;                  // descriptor offset of code row 64 inside a K-group
      for (int n = 0; n < steps0; ++n) {
        for (int X = 0; X < 2; ++X) {
          if (X == 1 && n >= steps1) break;
          ptx::mbar_wait(sb + Sm::bars + uint32_t(offsetof(Bars, a_ready)) + 8u * X, uint32_t(n) & 1);   // fp16 operand of this step is in TMEM
          ptx::tc_fence_after();
          RVQ_TRACE(X, n, 0, lane == 0);
          const uint32_t a_tmem = tmem_u + kTmemA + 64 * X;
          #pragma unroll 1
          for (int c = 0; c < nchunks; ++c) {
            uint32_t s3[3], p3[3];
            #pragma unroll
            for (int h = 0; h < 3; ++h) { s3[h] = slot; p3[h] = sph; if (++slot == kRing) { slot = 0; sph ^= 1; } }
            // first half: codes 0..63 of the chunk -> accumulator 0
            ptx::mbar_wait(bar_acce0, aph);                                           // accumulator 0 drained
            ptx::tc_fence_after();
            #pragma unroll
            for (int h = 0; h < 3; ++h) {
              ptx::mbar_wait(bar_full0 + 8u * s3[h], p3[h]);                          // K-third landed
              const uint64_t bs = bd0 + uint64_t(s3[h] * (kSlotBytes >> 4));
              if (h < 2) ptx::umma3_ts_w(tmem_u, a_tmem + 24 * h, bs, idesc, h == 0 ? 0u : 1u);
              else ptx::umma3_last_w(tmem_u, a_tmem + 48, ad_aug, bs, idesc);
            }
            ptx::umma_commit_w(bar_accf0);
            if (c == 0) RVQ_TRACE(X, n, 1, lane == 0);
            // second half: codes 64..127 -> accumulator 1; each K-third is released as its MMAs retire
            ptx::mbar_wait(bar_acce0 + 8u, aph);                                      // accumulator 1 drained
            ptx::tc_fence_after();
            #pragma unroll
            for (int h = 0; h < 3; ++h) {
              const uint64_t bs = bd0 + uint64_t(s3[h] * (kSlotBytes >> 4) + kHalfRows);
              if (h < 2) ptx::umma3_ts_w(tmem_u + kNh, a_tmem + 24 * h, bs, idesc, h == 0 ? 0u : 1u);
              else ptx::umma3_last_w(tmem_u + kNh, a_tmem + 48, ad_aug, bs, idesc);
              ptx::umma_commit_w(bar_empty0 + 8u * s3[h]);
            }
            ptx::umma_commit_w(bar_accf0 + 8u);
            aph ^= 1;
          }
          RVQ_TRACE(X, n, 2, lane == 0);
        }
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    ptx::reg_inc<160>();
    // ===== update warps =====
    const int u = warp - 4;
    const int q = u & 3, h = u >> 2;           // TMEM lane quadrant (frames 32q..32q+31 of a tile), half of the dims / of the frames
    const int f = q * 32 + lane;               // thread <-> frame mapping of tile loads
    const uint32_t tq = tmem + (uint32_t(q * 32) << 16);
    // pull this warp's share of a tile (64 lines: one per dim, 32 consecutive frames x 4 B) into L2
    auto prefetch_tile = [&](int tile) {
      const int64_t n = int64_t(tile) * p.tf + f;
      if (f < p.tf && n < p.N) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x + p.fa.base(n - lane) + int64_t(h * 64 + lane) * p.fa.sxd));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x + p.fa.base(n - lane) + int64_t(h * 64 + 32 + lane) * p.fa.sxd));
      }
    };
    // load dims 64h..64h+63 of the latent tile `tile` into slot X: fp32 residual and fp16 operand in tensor memory, |x|^2,
    // rounding residue
    auto load_tile = [&](int X, int tile) {
      unsigned char* ms = smem + Sm::misc + X * Sm::m_size;
      const int64_t n = int64_t(tile) * p.tf + f;
      const bool valid = f < p.tf && n < p.N;     // lanes beyond the tile's frames carry zeros and write nothing
      const int64_t xb = valid ? p.fa.base(n) : 0;
      float xsum = 0.f, e2 = 0.f;
      // the warp's 64 lines are asked for at once (a slot's later tiles were already requested a few stages before the end of
      // the previous tile), then read 16 dims at a time
      prefetch_tile(tile);
      #pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        float v[16];
        const int d0 = h * 64 + g * 16;
        const float* xp = p.x + xb + int64_t(d0) * p.fa.sxd;
        #pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = valid ? __ldg(xp + int64_t(j) * p.fa.sxd) : 0.f;
        uint32_t w[8], rr[16];
        #pragma unroll
        for (int j = 0; j < 16; ++j) rr[perm16(j)] = __float_as_uint(v[j]);
        #pragma unroll
        for (int j = 0; j < 16; j += 4) {
          xsum = fmaf(v[j], v[j], xsum); xsum = fmaf(v[j + 1], v[j + 1], xsum);
          xsum = fmaf(v[j + 2], v[j + 2], xsum); xsum = fmaf(v[j + 3], v[j + 3], xsum);
          e2 = residue2(v[j], v[j + 1], w[j / 2], e2);
          e2 = residue2(v[j + 2], v[j + 3], w[j / 2 + 1], e2);
        }
        ptx::tmem_st16(tq + kTmemR + 128 * X + d0, rr);
        ptx::tmem_st8(tq + kTmemA + 64 * X + d0 / 2, w);
      }
      // |x|^2 of this half of the dims goes where the (not yet written) class / batch masks of the tile's first stage live
      reinterpret_cast<float*>(ms + (h ? Sm::m_bmask : Sm::m_cmask))[f] = xsum;
      reinterpret_cast<float*>(ms + Sm::m_dr2)[h * kM + f] = e2;
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) { ptx::mbar_arrive(RVQ_BAR(a_ready, X)); ptx::mbar_arrive(RVQ_BAR(dr_ready, X)); }
    };
    // n = -1 is the prologue: it only loads the first tile of each slot (one call site for the tile load)
    for (int n = -1; n < steps0; ++n) {
      for (int X = 0; X < 2; ++X) {
        if (X == 1 && (n < 0 ? steps1 == 0 : n >= steps1)) break;      // slot 1: prologue only if it has a tile, then steps while n < steps1
        int next_tile = -1;
        if (n < 0) next_tile = tile0 + X;
        else {
          const int jt = n / p.n_q, s = n - jt * p.n_q;
          const int st = p.stage0 + s;
          const int64_t tile_n0 = int64_t(tile0 + X + 2 * jt) * p.tf;
          unsigned char* ms = smem + Sm::misc + X * Sm::m_size;
          // winners / candidate lists of this step: one warp polls the mbarrier, the others block on a hardware barrier
          // (a blocked warp costs no issue slots, a polling one does); the barrier also separates the passes' use of the
          // staging rows, the pool and its counter
          if (u == 0) { ptx::mbar_wait(RVQ_BAR(cand_ready, X), uint32_t(n) & 1); if (lane == 0) bars->pool_cnt = 0; }
          ptx::named_bar_sync(8, kUpdWarps * 32);
          RVQ_TRACE(X, n, 6, u == 0 && lane == 0);
          const bool last = s + 1 == p.n_q;
          // a few stages before the tile ends: the slot's next tile starts its way from HBM to L2
          if ((s + 4 == p.n_q || (p.n_q < 4 && s == 0)) && (jt + 1) * p.n_q < (X ? steps1 : steps0)) prefetch_tile(tile0 + X + 2 * (jt + 1));
          float sq = 0.f;
          const uint32_t lane16 = uint32_t(q * 32 + h * 16) << 16;      // TMEM lanes of this warp's 16 frames
          update_pass<TRAIN>(p, smem, sbase, ms, u, lane, s, rot, nchunks, tile_n0, pv.tab32(st), pv.cnorm(st),
                             tmem + lane16 + kTmemR + 128 * X, tmem + lane16 + kTmemA + 64 * X, !last, last, RVQ_BAR(a_ready, X), sq,
                             &bars->pool_cnt, X, n,
#ifdef RVQ_TC_TRACE
                             t_kernel0
#else
                             0
#endif
                             );
          RVQ_TRACE(X, n, 7, u == 0 && lane == 0);
          if (!last) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(RVQ_BAR(dr_ready, X));      // residual, rounding residues are in place
            RVQ_TRACE(X, n, 8, u == 0 && lane == 0);
          } else if ((jt + 1) * p.n_q < (X ? steps1 : steps0)) {
            next_tile = tile0 + X + 2 * (jt + 1);
            ptx::named_bar_sync(7, kUpdWarps * 32);     // the quadrant's other warp may not have read its residual rows yet
          }
          if (TRAIN && p.sqerr != nullptr) {
            sq = warp_sum(sq);
            if (lane == 0) atomicAdd(&p.sqerr[s], (double)sq);
          }
        }
        if (next_tile >= 0) load_tile(X, next_tile);
      }
    }
  } else {
    ptx::reg_inc<152>();
    // ===== score warps =====
    const int q = warp;                        // TMEM lane quadrant = frames 32q..32q+31 of a tile
    const int f = q * 32 + lane;
    uint32_t n_cert = 0, n_resc = 0, n_full = 0;                        // search statistics (rvq_search_stats)
    float xx_0 = 0.f, xx_1 = 0.f;              // upper bound of |r|^2 of this thread's frame in slot 0 / 1
    // opaque copies of loop invariants: kept in registers instead of being re-derived from special registers per chunk
    uint32_t tl = tmem + (uint32_t(q * 32) << 16), bar_full = RVQ_BAR(acc_full, 0), bar_empty = RVQ_BAR(acc_empty, 0);
    asm volatile("" : "+r"(tl), "+r"(bar_full), "+r"(bar_empty));
    float2* bmn = reinterpret_cast<float2*>(smem + Sm::bmn) + f;
    float* scr = reinterpret_cast<float*>(smem + Sm::scratch + warp * 512);
    uint32_t acc_ph = 0;                       // phase of the accumulators' `full` barriers
    const int nh = 2 * nchunks;                // half-chunks (accumulators) per stage
    for (int n = 0; n < steps0; ++n) {
      for (int X = 0; X < 2; ++X) {
        if (X == 1 && n >= steps1) break;
        const int jt = n / p.n_q, s = n - jt * p.n_q;
        const int st = p.stage0 + s;
        const int64_t nfr = int64_t(tile0 + X + 2 * jt) * p.tf + f;
        unsigned char* ms = smem + Sm::misc + X * Sm::m_size;
        const float* t32 = pv.tab32(st);
        const float* cn = pv.cnorm(st);
        const StageMeta* meta = pv.meta(st);
        // margin coefficients of this stage: requested before the chunk loop, used after it
        const float mt_coef = __ldg(&meta->margin_coef), mt_abs = __ldg(&meta->margin_abs), mt_xlimit = __ldg(&meta->xlimit);
        const float mt_cmax = __ldg(&meta->cmax_all), mt_dr = __ldg(&meta->margin_dr);
        // ---- scores: per-class minima (registers) and per-batch minima (shared memory) of the K approximate scores ----
        float cm[32];
        #pragma unroll
        for (int j = 0; j < 32; ++j) cm[j] = inf_f();
        // One rolled iteration per 64-code accumulator (the hot loops of a stage must stay inside the instruction cache): it
        // is read as two pairs of 16-column loads (columns c and c + 32: the same 16 classes of two batches).  The second pair
        // is in flight while the minima of the first are taken, the first pair of the NEXT accumulator while those of the
        // second are: only the first read of a stage waits for tensor memory.
        uint32_t x0[16], x1[16], y0[16], y1[16];
        ptx::mbar_wait(bar_full, acc_ph);
        ptx::tc_fence_after();
        RVQ_TRACE(X, n, 3, warp == 0 && lane == 0);
        ptx::tmem_ld16(tl, x0);
        ptx::tmem_ld16(tl + 32, x1);
        ptx::tmem_ld_wait();
        #pragma unroll 1
        for (int hs = 0; hs < nh; ++hs) {
          const uint32_t b = uint32_t(hs) & 1u;
          ptx::tmem_ld16(tl + b * kNh + 16, y0);
          ptx::tmem_ld16(tl + b * kNh + 48, y1);
          float ba, bb, ca, cb;
          pair_min(x0, x1, cm, ba, bb);
          ptx::tmem_ld_wait();
          // scores are in registers: hand the accumulator back before reducing the second pair
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_empty + 8u * b);
          if (b) acc_ph ^= 1;
          if (hs + 1 < nh) {
            ptx::mbar_wait(bar_full + 8u * (b ^ 1u), acc_ph);
            ptx::tc_fence_after();
            ptx::tmem_ld16(tl + (b ^ 1u) * kNh, x0);
            ptx::tmem_ld16(tl + (b ^ 1u) * kNh + 32, x1);
          }
          pair_min(y0, y1, cm + 16, ca, cb);
          bmn[hs * kM] = make_float2(fminf(ba, ca), fminf(bb, cb));
          ptx::tmem_ld_wait();
        }
        RVQ_TRACE(X, n, 4, warp == 0 && lane == 0);
        // |x|^2 of a new tile and the rounding residue of this frame's operand were written by the update warps; the
        // scores above could only exist after they had finished
        ptx::mbar_wait(RVQ_BAR(dr_ready, X), uint32_t(n) & 1);          // (completed long ago: the update warps finish phase 2 during the MMAs)
        ptx::tc_fence_after();
        float xx = X ? xx_1 : xx_0;
        if (s == 0) xx = reinterpret_cast<const float*>(ms + Sm::m_cmask)[f] + reinterpret_cast<const float*>(ms + Sm::m_bmask)[f];
        const float xnorm = sqrtf(xx);
        const bool outl = !(xnorm < mt_xlimit);      // also true for NaN
        const float drn = sqrtf(reinterpret_cast<const float*>(ms + Sm::m_dr2)[f] + reinterpret_cast<const float*>(ms + Sm::m_dr2)[kM + f]) * 1.001f;
        const float delta = mt_coef * xnorm + mt_dr * drn + mt_abs;
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
        for (int j = 0; j < 32; ++j)                  // two instructions per value: compare, predicated OR with an immediate
          asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(cm4[j & 3]) : "f"(cm[j]), "f"(thr), "r"(1u << j));
        #pragma unroll
        for (int hs = 0; hs < 16; ++hs) {
          if (hs < nh) {                              // bit a = a-th batch processed (a = 2 hs, 2 hs + 1)
            const float2 b2 = bmn[hs * kM];
            asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(bm4[hs & 3]) : "f"(b2.x), "f"(thr), "r"(1u << (2 * hs)));
            asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(bm4[(hs + 2) & 3]) : "f"(b2.y), "f"(thr), "r"(1u << (2 * hs + 1)));
          }
        }
        const uint32_t cmask = (cm4[0] | cm4[1]) | (cm4[2] | cm4[3]);
        const uint32_t bmask = (bm4[0] | bm4[1]) | (bm4[2] | bm4[3]);
        const int nc = __popc(cmask), nb = __popc(bmask);
        const bool full = outl || cmask == 0u || bmask == 0u;     // masks are empty only for NaN scores
        const int ncand = nc * nb;

        // first candidate (= the winner when certified); the update warps enumerate the other codes of a short list
        // from the two masks
        const int4 cd = make_int4(code_of(__ffs(bmask) - 1, __ffs(cmask) - 1, rot, nchunks), -1, -1, -1);
        *reinterpret_cast<int4*>(ms + Sm::m_cand + f * 16) = cd;
        *reinterpret_cast<int*>(ms + Sm::m_ncnt + f * 4) = full ? kFull : (ncand > 4 ? kBig : ncand);
        *reinterpret_cast<uint32_t*>(ms + Sm::m_cmask + f * 4) = cmask;
        *reinterpret_cast<uint32_t*>(ms + Sm::m_bmask + f * 4) = bmask;
        int64_t* code_out = (f < p.tf && nfr < p.N) ? p.codes + int64_t(s) * p.N + nfr : nullptr;
        if (!full && ncand == 1 && code_out != nullptr) *code_out = cd.x;      // certified: the warp's codes are one 256-byte run
        {
          // frames outside the fp16 image's validity range (or NaN): exact scan right here, then they are certified.  The
          // frame's residual row comes out of tensor memory through its own lane into the warp's scratch row.
          uint32_t fm = __ballot_sync(0xffffffffu, full);
          while (fm) {
            const int i = __ffs(fm) - 1; fm &= fm - 1;
            #pragma unroll 1
            for (int blk = 0; blk < 8; ++blk) {
              uint32_t v[16];
              ptx::tmem_ld16(tl + kTmemR + 128 * X + 16 * blk, v);
              ptx::tmem_ld_wait();
              if (lane == i) {
                #pragma unroll
                for (int j = 0; j < 16; j += 4)
                  *reinterpret_cast<uint4*>(scr + 16 * blk + j) = make_uint4(v[perm16(j)], v[perm16(j + 1)], v[perm16(j + 2)], v[perm16(j + 3)]);
              }
            }
            __syncwarp();
            int64_t* co = reinterpret_cast<int64_t*>(__shfl_sync(0xffffffffu, (unsigned long long)code_out, i));
            resolve_full(scr, ms, q * 32 + i, lane, p.K, t32, cn, co);
          }
        }
        n_full += full ? 1u : 0u; n_cert += (!full && ncand == 1) ? 1u : 0u; n_resc += (!full && ncand > 1) ? 1u : 0u;
        // upper bound of the next residual's |r|^2 (only the margin and the validity test use it):
        // the winner's approximate score is <= m + delta and off by <= delta/2
        if (full) { const float g2 = xnorm + mt_cmax; xx = g2 * g2; }
        else xx = fmaxf(xx + m + 1.5f * delta, 0.f) * 1.00001f + 1e-30f;
        if (X) xx_1 = xx; else xx_0 = xx;
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(RVQ_BAR(cand_ready, X));    // winners and lists visible to the update warps
        RVQ_TRACE(X, n, 5, warp == 0 && lane == 0);
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
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 13) ptx::tmem_dealloc(tmem, 512);
}

int tc_debug_trace(long long* out_host, int n) {
  const int cap = 2 * kTraceSteps * kTraceEv + 128 + 2 * 64;
  const int m = n < cap ? n : cap;
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
  // full 128-frame tiles: a CTA's odd last tile runs alone (measured on B200: this beats smaller balanced tiles)
  p.tf = kM;
  const int64_t ntiles = (N + p.tf - 1) / p.tf;
  // one tile per CTA while there are SMs to spare (a lone tile's stage is shorter than a pair's: small calls are latency-bound),
  // two or more tiles per CTA beyond that
  const unsigned grid = unsigned(ntiles < sm_count ? ntiles : sm_count);
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

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
//            A = fp16(r), held IN TENSOR MEMORY (64 columns per slot; K-steps 0..7) + one shared-memory block per slot
//                whose columns 128,129 = 1 (K-step 8: picks up the hi/lo halves of |c|^2) and whose columns 130 / 132
//                carry the frame's bound R >= |r| when the stage certifies with the per-code bound (else 0);
//            B = fp16 image of the codebook (cols 0..127 = -2c, cols 128,129 = hi/lo of |c|^2, cols 130,132 = -g16_k,
//                the code's own error coefficient: the accumulator then holds LOWER BOUNDS of the true scores), streamed by the
//                TMA engine (cp.async.bulk) from the L2-resident pack into a ring of 7 third-of-a-chunk
//                slots (6 K-groups = 12 KB each; small slots keep more bytes in flight than whole chunks would);
//            D = 3 accumulator buffers of 128 TMEM columns shared by both slots.
//   warps    0..3   score warps  (thread = TMEM lane = frame): tcgen05.ld, per-class / per-batch minima, certified
//                   winner or class/batch masks of the candidates (see below), codes of certified frames;
//            4..11  update warps, in the order of what sits on the chain to the next stage's MMA: (1) winner rows of the
//                   certified frames requested from the fp32 table (4 lanes own 2 frames, lane m = 16-byte chunks 4i+m);
//                   (2) while they fly, frames with a candidate list are settled by a warp each (exact fp32 re-score, the
//                   warp applies r <- r - q itself); (3) r - q in registers -> fp16 operand of the next stage to tensor
//                   memory (one tcgen05.st.16x256b.x8) -> a_ready; (4) off the chain: residual rows back to shared
//                   memory, exact rounding residue of the operand, squared error -> dr_ready; tile loads;
//            12  TMA producer;  13  MMA issuer (owns the TMEM allocation);  14, 15 idle.
//            setmaxnreg: 152 registers for the score warps, 160 for the update warps, 40 for the last warpgroup.
//   sync     mbarriers only between roles: a_ready[slot] (update -> MMA), acc_full/acc_empty (MMA <-> score),
//            cand_ready[slot] (score -> update), dr_ready[slot] (update -> score), full/empty (TMA <-> MMA).
//
// Certified argmin: a score warp keeps per frame the minimum over every 32-code batch and over every residue class
// (code mod 32).  A code is within `delta` of the minimum iff its batch AND its class are; delta bounds the fp16
// score error two-sidedly (rvq_common.cuh, StageMeta), so the exact fp32 winner is certified when exactly one batch
// and one class qualify.  Otherwise the candidates (flagged batches x flagged classes) are re-scored in fp32 with
// the reference's formula (core_vq.py:181-189, ties -> lowest index): 2..4 candidates by one warp per frame, wider sets by
// all update warps together.  Frames outside the fp16 image's validity range take an exact fp32 scan.  Two bounds, both
// rigorous (rvq_common.cuh, StageMeta): per stage (set by the largest code) and per code (the threshold comes from the
// coefficients of the code that attains the minimum; tables with heterogeneous norms, i.e. fitted ones); stages pick at pack
// time, and the kernel runs the specialisation of its body that matches the call (tc_encode_body<TRAIN, PC, STE>).
// The training variant also accumulates the EMA statistics of core_vq.py:227-228 (rvq_encode_train).
#include "rvq_common.cuh"
#include "rvq_ptx.cuh"

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
constexpr int kBig = 5;                 // ncnt marker: more than 4 candidates (enumerate the masks)
constexpr int kFull = 6;                // ncnt marker: exact scan of the whole table
constexpr int kRsBytes = kM * 128 * 4;  // fp32 residual of one tile

struct Sm {
  // augmented K block of the A operand (K-step 8), no swizzle: k-group 0 of slot 0, k-group 0 of slot 1 (per frame:
  // 1, 1, R0, 0, R1, 0, 0, 0 -- the ones meet the hi/lo halves of |c|^2, R0 + R1 >= |r| meets -g16_k when the stage uses
  // the per-code bound), then the all-zero k-group 1 shared by both slots
  static constexpr uint32_t aug = 0;
  static constexpr uint32_t ring = aug + 6144;
  static constexpr uint32_t rs = ring + kRing * kSlotBytes;        // 2 x fp32 [128 f][128 d], chunk-swizzled
  static constexpr uint32_t misc = rs + 2 * kRsBytes;              // 2 x per-slot block (offsets m_*)
  static constexpr uint32_t m_cand = 0;                            // int4 [128]: candidate codes (-1 = none)
  static constexpr uint32_t m_ncnt = m_cand + kM * 16;             // u8 [128]
  static constexpr uint32_t m_cmask = m_ncnt + kM;                 // u32 [128] flagged classes   (a fresh tile: |x|^2 of dims 0..63)
  static constexpr uint32_t m_bmask = m_cmask + kM * 4;            // u32 [128] flagged batches   (a fresh tile: |x|^2 of dims 64..127)
  static constexpr uint32_t m_dr2 = m_bmask + kM * 4;              // float [2][128]: |r - fp16(r)|^2 of the current operand, per half of the dims
  static constexpr uint32_t m_slowq = m_dr2 + 2 * kM * 4;          // u8 [128]: frames with 2..4 listed candidates
  static constexpr uint32_t m_wideq = m_slowq + kM;                // u8 [128]: frames with a wide candidate set
  static constexpr uint32_t m_qcnt = m_wideq + kM;                 // int [2]: queue lengths {slow, wide}
  static constexpr uint32_t m_size = m_qcnt + 16;
  static constexpr uint32_t bars = misc + 2 * m_size;
  static constexpr uint32_t total = bars + 224;
};
struct Bars {
  uint64_t full[kRing], empty[kRing], acc_full[kAccBufs], acc_empty[kAccBufs], a_ready[2], cand_ready[2], dr_ready[2];
  uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 224, "barrier block");
// shared-window address of a barrier, from the CTA's window base (no generic->shared conversion inside the hot loops)
#define RVQ_BAR(field, i) (sbase + Sm::bars + uint32_t(offsetof(Bars, field)) + 8u * uint32_t(i))
static_assert(Sm::total <= 227 * 1024, "shared memory budget");
static_assert(kTcKPad / 16 == 9 && kN == 128 && kTmemA + 2 * 64 == 512, "operand geometry");
static_assert((Sm::m_size % 16) == 0 && (Sm::misc % 16) == 0, "alignment");

// debug timeline of CTA 0 (RVQ_TC_TRACE builds; slots 0/1, steps kTraceN0 .. kTraceN0 + kTraceSteps - 1; -DRVQ_TRACE_N0=36
// looks at the CTA's third tile, which runs alone): g_trace[X][step][event] = cycles since kernel start
#ifndef RVQ_TRACE_N0
#define RVQ_TRACE_N0 2
#endif
constexpr int kTraceSteps = 6, kTraceEv = 16;
[[maybe_unused]] constexpr int kTraceN0 = RVQ_TRACE_N0;
__device__ long long g_trace[2 * kTraceSteps * kTraceEv + 128 + 2 * 64];   // + per-chunk detail of the MMA thread / the producer for (slot 0, step 4)
#ifdef RVQ_TC_TRACE
#define RVQ_TRACE(X, n, ev, cond) do { if (blockIdx.x == 0 && (cond) && (n) >= kTraceN0 && (n) < kTraceN0 + kTraceSteps) { \
    asm volatile("" ::: "memory"); g_trace[(((X) * kTraceSteps) + (n) - kTraceN0) * kTraceEv + (ev)] = clock64() - t_kernel0; asm volatile("" ::: "memory"); } } while (0)
#define RVQ_TRACE2(X, n, idx) do { if (blockIdx.x == 0 && (X) == 0 && (n) == kTraceN0 + 2) { \
    asm volatile("" ::: "memory"); g_trace[2 * kTraceSteps * kTraceEv + (idx)] = clock64() - t_kernel0; asm volatile("" ::: "memory"); } } while (0)
// update-pass detail (slot X, step 4): g_trace[base + 128 + 64 X + 8 u + e], lane 0 of update warp u
#define RVQ_TRACE3(X, n, u, e) do { if (blockIdx.x == 0 && (n) == kTraceN0 + 2 && (threadIdx.x & 31) == 0) { \
    asm volatile("" ::: "memory"); g_trace[2 * kTraceSteps * kTraceEv + 128 + 64 * (X) + 8 * (u) + (e)] = clock64() - t_kernel0; asm volatile("" ::: "memory"); } } while (0)
#else
#define RVQ_TRACE(X, n, ev, cond) do { } while (0)
#define RVQ_TRACE2(X, n, idx) do { } while (0)
#define RVQ_TRACE3(X, n, u, e) do { } while (0)
#endif
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
  float* ema_counts; float* ema_sum;   // EMA statistics of core_vq.py:227-228 ([n_q, K] / [n_q, K, D], accumulated) or nullptr
  int ste;
  int bkt;                 // codes as [B, n_q, T] (RVQ_FLAG_CODES_BKT)
  int direct;              // exact re-scores use the k-means distance sum((x - c)^2) of core_vq.py:86-91 (RVQ_FLAG_DIRECT_DIST)
  int tf;                  // frames per tile (<= 128): chosen by the host so that every CTA gets an even number of tiles
  unsigned long long* counters;
};

__device__ __forceinline__ float inf_f() { return __int_as_float(0x7f800000); }
// vector reduction into global memory (no return value: the adds are settled by the L2)
__device__ __forceinline__ void red_add_f4(float* addr, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// residual element group: 16-byte chunk ch (dims 4ch..4ch+3) of frame f, XOR-swizzled with the frame number (its low
// three bits reversed) so that all three access patterns of the kernel spread over the banks: 8 lanes = 8 consecutive
// chunks of one frame; 8 lanes = 8 consecutive frames, one chunk; 8 lanes = 2 consecutive frames x 4 consecutive chunks
__device__ __forceinline__ int rs_swz(int f) { return (f & 24) | ((f & 1) << 2) | (f & 2) | ((f >> 2) & 1); }
__device__ __forceinline__ int rs_off(int f, int ch) { return f * 128 + ((ch ^ rs_swz(f)) << 2); }
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
// sum of squared differences of two float4 (the k-means distance of core_vq.py:86-88)
__device__ __forceinline__ float sqd4(const float4& a, const float4& b, float acc) {
  const float x = a.x - b.x, y = a.y - b.y, z = a.z - b.z, w = a.w - b.w;
  acc = fmaf(x, x, acc); acc = fmaf(y, y, acc); acc = fmaf(z, z, acc); return fmaf(w, w, acc);
}
__device__ __forceinline__ float min32(const uint32_t (&v)[32]) {
  float t[11];
  #pragma unroll
  for (int j = 0; j < 10; ++j) t[j] = ptx::fmin3(__uint_as_float(v[3 * j]), __uint_as_float(v[3 * j + 1]), __uint_as_float(v[3 * j + 2]));
  t[10] = fminf(__uint_as_float(v[30]), __uint_as_float(v[31]));
  const float a = ptx::fmin3(t[0], t[1], t[2]), b = ptx::fmin3(t[3], t[4], t[5]), c = ptx::fmin3(t[6], t[7], t[8]);
  return ptx::fmin3(ptx::fmin3(a, b, c), t[9], t[10]);
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

// tiles [start, start+cnt) of this CTA (contiguous ranges: measured 1 % faster at cfg2 than the interleaved order
// blockIdx.x + j * gridDim.x, whose only merit is to spread a stretch of expensive frames over more CTAs)
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

// Score-warp side.  Frames whose candidate set is the whole table (outside the fp16 image's validity range, NaN):
// the whole warp scores the table, one code per lane, stores the code and rewrites the frame's entry as a
// certified winner for the update warps.
__device__ __forceinline__ void resolve_full(const float* rs, unsigned char* ms, int f, int lane, int K, const float* __restrict__ t32,
                                          const float* __restrict__ cn, int64_t* code_out, bool direct) {
  const float4 rl = *reinterpret_cast<const float4*>(rs + rs_off(f, lane));
  const float rr = warp_sum(dot4(rl, rl, 0.f));
  float best = inf_f(); int bcode = 0x7fffffff;
  for (int code = lane; code < K; code += 32) {
    const float4* rp = reinterpret_cast<const float4*>(t32 + size_t(code) * 128);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    float dist;
    if (direct) {
      #pragma unroll 2
      for (int ch = 0; ch < 32; ch += 4) {
        a0 = sqd4(*reinterpret_cast<const float4*>(rs + rs_off(f, ch + 0)), __ldg(rp + ch + 0), a0);
        a1 = sqd4(*reinterpret_cast<const float4*>(rs + rs_off(f, ch + 1)), __ldg(rp + ch + 1), a1);
        a2 = sqd4(*reinterpret_cast<const float4*>(rs + rs_off(f, ch + 2)), __ldg(rp + ch + 2), a2);
        a3 = sqd4(*reinterpret_cast<const float4*>(rs + rs_off(f, ch + 3)), __ldg(rp + ch + 3), a3);
      }
      dist = (a0 + a1) + (a2 + a3);                                    // core_vq.py:86-88
    } else {
      #pragma unroll 2
      for (int ch = 0; ch < 32; ch += 4) {
        a0 = dot4(*reinterpret_cast<const float4*>(rs + rs_off(f, ch + 0)), __ldg(rp + ch + 0), a0);
        a1 = dot4(*reinterpret_cast<const float4*>(rs + rs_off(f, ch + 1)), __ldg(rp + ch + 1), a1);
        a2 = dot4(*reinterpret_cast<const float4*>(rs + rs_off(f, ch + 2)), __ldg(rp + ch + 2), a2);
        a3 = dot4(*reinterpret_cast<const float4*>(rs + rs_off(f, ch + 3)), __ldg(rp + ch + 3), a3);
      }
      const float dot = (a0 + a1) + (a2 + a3);
      dist = (rr - 2.f * dot) + __ldg(cn + code);                      // core_vq.py:183-187
    }
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
    ms[Sm::m_ncnt + f] = 1;
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
template <bool STE>
__device__ __forceinline__ float4 sub_row(const float4& rv, float4 q) {
  if (STE) { q.x = rv.x + (q.x - rv.x); q.y = rv.y + (q.y - rv.y); q.z = rv.z + (q.z - rv.z); q.w = rv.w + (q.w - rv.w); }
  return make_float4(rv.x - q.x, rv.y - q.y, rv.z - q.z, rv.w - q.w);
}
// A frame whose flagged batches x flagged classes give more than 4 candidates: the whole warp works on it,
// 8 candidates per step (2 per quarter-warp); the quarter that found the winner updates the frame.
template <int NC> struct Cand { int c[NC]; Row4 w[NC]; float nrm[NC]; };
template <int NC>
__device__ __forceinline__ void load_cand(Cand<NC>& k, int j, const float* __restrict__ t32, const float* __restrict__ cn) {
  #pragma unroll
  for (int u = 0; u < NC; ++u) {      // (unconditional: a conditionally initialised struct is kept in local memory)
    const int c = k.c[u] < 0 ? 0 : k.c[u];
    k.w[u] = load_row(t32, c, j); k.nrm[u] = __ldg(cn + c);
  }
}
// exact distances of up to NC candidates (core_vq.py:183-187); keeps the best (lowest code on ties) and its slot u
template <int NC>
__device__ __forceinline__ void score_cand(const Cand<NC>& k, const Row4& r, float rr, float& best, int& bcode, int& bidx, bool direct) {
  float d[NC];
  #pragma unroll
  for (int u = 0; u < NC; ++u)
    d[u] = direct ? (sqd4(r.v[0], k.w[u].v[0], 0.f) + sqd4(r.v[1], k.w[u].v[1], 0.f)) + (sqd4(r.v[2], k.w[u].v[2], 0.f) + sqd4(r.v[3], k.w[u].v[3], 0.f))
                  : dot_row(r, k.w[u]);
  #pragma unroll
  for (int off = 4; off > 0; off >>= 1) {
    #pragma unroll
    for (int u = 0; u < NC; ++u) d[u] += __shfl_xor_sync(0xffffffffu, d[u], off);
  }
  #pragma unroll
  for (int u = 0; u < NC; ++u) {
    const float e = direct ? d[u] : (rr - 2.f * d[u]) + k.nrm[u];
    if (k.c[u] >= 0 && (e < best || (e == best && k.c[u] < bcode))) { best = e; bcode = k.c[u]; bidx = u; }
  }
}
// A frame whose flagged batches x flagged classes give more than 4 candidates: ALL update warps work on it, two candidates per
// quarter-warp and step (64 per step), exact fp32 distances (core_vq.py:183-187).  Every quarter folds its best into the
// frame's 64-bit key {orderable distance, code} with a shared-memory atomicMin -- smallest distance, lowest code on ties; the
// key lives in the upper half of the frame's candidate entry, which the score warp left at all ones -- and the winner is read
// after the update warps' barrier.  (Degenerate tables -- hundreds of near-identical codes -- give candidate sets of several
// hundred codes: one warp per frame, 8 candidates per step, took 100 k cycles for such a frame.)
template <bool TRAIN>
__device__ __forceinline__ void resolve_wide(const TcParams& p, const float* rs, unsigned char* ms, int f, int u, int lane, int rot, int nchunks,
                                             const float* __restrict__ t32, const float* __restrict__ cn) {
  const int qq = lane >> 3, j = lane & 7;
  const uint32_t cm = *reinterpret_cast<const uint32_t*>(ms + Sm::m_cmask + f * 4);
  const uint32_t bm = *reinterpret_cast<const uint32_t*>(ms + Sm::m_bmask + f * 4);
  const int nc = __popc(cm), total = nc * __popc(bm);
  constexpr int kPer = 4;                                  // candidates per quarter-warp and step
  constexpr int kCoop = kUpdWarps;                         // warps that share a frame's candidates
  if (4 * kPer * u >= total) return;                       // (warp-uniform) no candidate left for this warp
  const int w = 4 * u + qq;                                // this quarter's number among the 32 of the update warps
  // lane L: position of the L-th flagged batch / class (candidate t = flagged batch t / nc, flagged class t % nc)
  uint32_t mb = bm, mc = cm;
  const int npos = min(lane, max(nc, total / nc));         // (positions beyond the populations are never asked for)
  for (int i = 0; i < npos; ++i) { mb &= mb - 1; mc &= mc - 1; }
  const int posb = mb ? __ffs(mb) - 1 : 0, posc = mc ? __ffs(mc) - 1 : 0;
  const float inv_nc = 1.f / float(nc);
  const Row4 r = load_res(rs, f, j);
  const float rr = quarter_sum(dot_row(r, r));
  float best = inf_f(); int bcode = 0x7fffffff, bidx = 0;
  #pragma unroll 1
  for (int t0 = 0; t0 < total; t0 += 4 * kPer * kCoop) {
    Cand<kPer> k;
    #pragma unroll
    for (int v = 0; v < kPer; ++v) {
      const int t = t0 + kPer * w + v;
      const int ia = int((float(t) + 0.5f) * inv_nc);      // exact for t < 1024, nc <= 32
      const int ic = t - ia * nc;
      const int a = __shfl_sync(0xffffffffu, posb, ia & 31), jj = __shfl_sync(0xffffffffu, posc, ic & 31);
      k.c[v] = t < total ? code_of(a, jj, rot, nchunks) : -1;
    }
    load_cand<kPer>(k, j, t32, cn);
    score_cand<kPer>(k, r, rr, best, bcode, bidx, TRAIN && p.direct);
  }
  if (j == 0 && bcode != 0x7fffffff) {                     // (NaN distances never enter: the key then stays at all ones)
    uint32_t b = __float_as_uint(best + 0.f);              // -0 -> +0; then the usual order-preserving map to unsigned
    b ^= (b >> 31) ? 0xffffffffu : 0x80000000u;
    atomicMin(reinterpret_cast<unsigned long long*>(ms + Sm::m_cand + f * 16 + 8), (static_cast<unsigned long long>(b) << 32) | uint32_t(bcode));
  }
}

// The residual update of one (slot, stage), executed by the eight update warps (u = 0..7: TMEM lane quadrant q = u & 3,
// half h = u >> 2 of its 32 frames).
//   1. RESOLVE.  Frames with a candidate list (2..4 codes) come from the slot's queue, spread over the 32 quarter-warps
//      (8 lanes per frame): the candidate rows are loaded together and scored in exact fp32 (core_vq.py:183-187); frames
//      with wide candidate sets take a whole warp.  Only the winning code is written back (candidate entry + output).
//   2. UPDATE + OPERAND.  A group of 4 lanes owns two frames (TMEM lanes g and g + 8 of the warp's 16): lane m holds the
//      16-byte chunks 4i + m (i = 0..7) of each.  Both winner rows are fetched (64 contiguous bytes per group and i), the
//      residual row is updated in shared memory (r <- r - q, exact fp32), converted to fp16 with its exact rounding
//      residue, and the 16 frames x 64 columns of the next stage's operand go to tensor memory with ONE
//      tcgen05.st.16x256b.x8 -- the register layout of that shape is exactly this lane/chunk assignment, so no
//      transposition pass and no second barrier are needed.
template <bool TRAIN, bool STE>
__device__ __forceinline__ void update_pass(const TcParams& p, float* rs, unsigned char* ms, int u, int lane, int s, int rot,
                                            int nchunks, int64_t tile_n0, const float* __restrict__ t32,
                                            const float* __restrict__ cn, uint32_t taddr, bool store, uint32_t bar_a, float& sq,
                                            int trX, int trn, long long t_kernel0) {
  const int q = u & 3, h = u >> 2;
  RVQ_TRACE3(trX, trn, u, 0);
  const int* qc = reinterpret_cast<const int*>(ms + Sm::m_qcnt);
  const int nslow = qc[0], nwide = qc[1];
  const int4* cand = reinterpret_cast<const int4*>(ms + Sm::m_cand);
  const unsigned char* ncnt = ms + Sm::m_ncnt;
  const int g = lane >> 2, m = lane & 3;
  const int fA = q * 32 + h * 16 + g, fB = fA + 8;
  // The winner rows of the certified frames are requested FIRST: they are in flight while the listed frames are resolved.
  //   ncnt == 1: certified (or settled by an exact scan)   -> row requested now, subtracted below
  //   ncnt 2..4: candidate list -> the resolving quarter-warp holds the winner's row and applies r <- r - q itself
  //   ncnt kBig: wide set       -> only the code is resolved; the row is requested after the barrier
  const int nA = ncnt[fA], nB = ncnt[fB];
  const unsigned char* slowq = ms + Sm::m_slowq;
  const unsigned char* wideq = ms + Sm::m_wideq;
  // candidate lists (2..4 codes): one frame per warp at a time, lane = 16-byte chunk of the row, so a frame costs a
  // handful of registers next to the winner rows in flight.  Exact fp32 distances (core_vq.py:183-187), lowest code on
  // ties; the warp holds the winner's row and applies r <- r - q itself.
  struct Item { int f, ck; int4 cd; float4 rl[4], w[4]; float nrm; };
  auto item_cands = [&](int i, Item& it) {
    it.f = slowq[i];
    // the 2..4 candidates = flagged batches x flagged classes (bit a of the batch mask is the a-th batch in this CTA's
    // processing order).  Straight-line for the usual case of at most two flagged batches and classes (the first candidate
    // is always (lowest batch, lowest class), the order of the others does not matter: ties go to the lowest CODE);
    // three or four flagged classes of one batch (or vice versa) take the generic walk.
    it.cd = cand[it.f];                       // the score warp left all codes of a short list
    if (it.cd.y != -2) return;
    const uint32_t cmk = *reinterpret_cast<const uint32_t*>(ms + Sm::m_cmask + it.f * 4);
    const uint32_t bmk = *reinterpret_cast<const uint32_t*>(ms + Sm::m_bmask + it.f * 4);
    it.cd = make_int4(-1, -1, -1, -1);
    uint32_t bm2 = bmk;
    int w = 0;
    while (bm2) {
      const int a = __ffs(bm2) - 1; bm2 &= bm2 - 1;
      uint32_t cm2 = cmk;
      while (cm2) {
        const int code = code_of(a, __ffs(cm2) - 1, rot, nchunks); cm2 &= cm2 - 1;
        if (w == 0) it.cd.x = code; else if (w == 1) it.cd.y = code; else if (w == 2) it.cd.z = code; else it.cd.w = code;
        ++w;
      }
    }
  };
  // One candidate per quarter-warp: lane (qq, j) = (lane / 8, lane % 8) holds the chunks 8k + j (k = 0..3) of candidate qq's
  // row and of the frame's residual row (whole 128-byte lines per load instruction), so a distance needs 3 shuffle levels
  // and the four candidates are compared with 2 more.
  auto item_load = [&](int i, Item& it) {
    item_cands(i, it);
    const int qq = lane >> 3, j = lane & 7;
    it.ck = qq == 0 ? it.cd.x : qq == 1 ? it.cd.y : qq == 2 ? it.cd.z : it.cd.w;
    const float4* rp = reinterpret_cast<const float4*>(t32 + size_t(it.ck < 0 ? 0 : it.ck) * 128) + j;
    #pragma unroll
    for (int k = 0; k < 4; ++k) it.w[k] = __ldg(rp + 8 * k);
    it.nrm = __ldg(cn + (it.ck < 0 ? 0 : it.ck));
    #pragma unroll
    for (int k = 0; k < 4; ++k) it.rl[k] = *reinterpret_cast<const float4*>(rs + rs_off(it.f, 8 * k + j));
  };
  auto item_finish = [&](const Item& it) {
    const bool direct = TRAIN && p.direct;
    const int j = lane & 7;
    float rr = (dot4(it.rl[0], it.rl[0], 0.f) + dot4(it.rl[1], it.rl[1], 0.f)) + (dot4(it.rl[2], it.rl[2], 0.f) + dot4(it.rl[3], it.rl[3], 0.f));
    float d = direct ? (sqd4(it.rl[0], it.w[0], 0.f) + sqd4(it.rl[1], it.w[1], 0.f)) + (sqd4(it.rl[2], it.w[2], 0.f) + sqd4(it.rl[3], it.w[3], 0.f))
                     : (dot4(it.rl[0], it.w[0], 0.f) + dot4(it.rl[1], it.w[1], 0.f)) + (dot4(it.rl[2], it.w[2], 0.f) + dot4(it.rl[3], it.w[3], 0.f));
    #pragma unroll
    for (int off = 4; off > 0; off >>= 1) {
      rr += __shfl_xor_sync(0xffffffffu, rr, off);
      d += __shfl_xor_sync(0xffffffffu, d, off);
    }
    // exact fp32 distance of this quarter's candidate (core_vq.py:183-187; :86-88 for k-means), then the best of the four,
    // lowest code on ties
    float best = direct ? d : (rr - 2.f * d) + it.nrm;
    int bcode = it.ck;
    if (it.ck < 0 || !(best == best)) { best = inf_f(); bcode = 0x7fffffff; }
    #pragma unroll
    for (int off = 8; off <= 16; off <<= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int oc = __shfl_xor_sync(0xffffffffu, bcode, off);
      if (ob < best || (ob == best && oc < bcode)) { best = ob; bcode = oc; }
    }
    if (bcode == 0x7fffffff) bcode = it.cd.x;                   // NaN distances only: the first candidate
    const int64_t nfr = tile_n0 + it.f;
    if (it.ck == bcode) {                                        // the winner's quarter holds its row: r <- r - q
      #pragma unroll
      for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(rs + rs_off(it.f, 8 * k + j)) = sub_row<STE>(it.rl[k], it.w[k]);
      if (TRAIN && p.ema_sum != nullptr && it.f < p.tf && nfr < p.N) {      // EMA statistics: embed_sum[code] += r, counts[code] += 1
        float* row = p.ema_sum + (size_t(s) * p.K + bcode) * 128 + 4 * j;
        #pragma unroll
        for (int k = 0; k < 4; ++k) red_add_f4(row + 32 * k, it.rl[k]);
        if (j == 0) atomicAdd(p.ema_counts + size_t(s) * p.K + bcode, 1.f);
      }
    }
    if (lane == 0) {
      *reinterpret_cast<int*>(ms + Sm::m_cand + it.f * 16) = bcode;
      if (it.f < p.tf && nfr < p.N) p.codes[code_index(p.bkt, p.n_q, p.fa.T, p.N, s, nfr)] = bcode;
    }
  };
  // Listed frames FIRST, up to two per warp in flight at once (their candidate rows are the only loads of the warp at
  // that point: one L2 round trip for all of them instead of one per frame queued behind the 64 winner rows), then the
  // winner rows of the certified frames, which fly while the warps meet at the barrier.
  if (nslow + nwide > 0) {
    #pragma unroll 1
    for (int i = u; i < nslow; i += kUpdWarps) {
      Item it0;
      item_load(i, it0);
      item_finish(it0);
    }
    // wide candidate sets: one frame at a time, all warps together
    #pragma unroll 1
    // (the first candidates of frame i go to warp 7 - i, 6 - i, ...: the warps at the low end hold most of the listed frames)
    for (int i = 0; i < nwide; ++i) resolve_wide<TRAIN>(p, rs, ms, wideq[i], (kUpdWarps - 1 - u + i) & (kUpdWarps - 1), lane, rot, nchunks, t32, cn);
    RVQ_TRACE3(trX, trn, u, 1);
  }
  float4 qa[8], qb[8];
  int codeA = cand[fA].x, codeB = cand[fB].x;
  {
    const float4* ra = reinterpret_cast<const float4*>(t32 + size_t(codeA) * 128) + m;
    const float4* rb = reinterpret_cast<const float4*>(t32 + size_t(codeB) * 128) + m;
    #pragma unroll
    for (int i = 0; i < 8; ++i) qa[i] = nA == 1 ? __ldg(ra + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    #pragma unroll
    for (int i = 0; i < 8; ++i) qb[i] = nB == 1 ? __ldg(rb + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (nslow + nwide > 0) {
    ptx::named_bar_sync(6, kUpdWarps * 32);      // listed frames are updated, wide frames have their winner
    RVQ_TRACE3(trX, trn, u, 2);
    if (nwide > 0 && __any_sync(0xffffffffu, nA == kBig || nB == kBig)) {
      // winners of the wide sets: low word of the frame's key (all ones = NaN distances only: the first candidate, like an
      // exact scan would answer); the group's first lane writes the code out
      auto wide_winner = [&](int f) {
        const int4 cd = cand[f];
        const int code = cd.z == -1 ? cd.x : cd.z;
        const int64_t nfr = tile_n0 + f;
        if (m == 0 && f < p.tf && nfr < p.N) p.codes[code_index(p.bkt, p.n_q, p.fa.T, p.N, s, nfr)] = code;
        return code;
      };
      if (nA == kBig) {
        codeA = wide_winner(fA);
        const float4* ra = reinterpret_cast<const float4*>(t32 + size_t(codeA) * 128) + m;
        #pragma unroll
        for (int i = 0; i < 8; ++i) qa[i] = __ldg(ra + 4 * i);
      }
      if (nB == kBig) {
        codeB = wide_winner(fB);
        const float4* rb = reinterpret_cast<const float4*>(t32 + size_t(codeB) * 128) + m;
        #pragma unroll
        for (int i = 0; i < 8; ++i) qb[i] = __ldg(rb + 4 * i);
      }
    }
  }
  const bool doneA = nA >= 2 && nA <= 4, doneB = nB >= 2 && nB <= 4;     // already updated by the resolving quarter-warp
  RVQ_TRACE3(trX, trn, u, 3);
  // Phase 1 (on the chain to the next stage's MMA): n = r - q in registers (the winner rows are overwritten), fp16 operand
  // to tensor memory, a_ready.  Phase 2 (off that chain): residual rows back to shared memory, exact rounding residue of
  // the operand, squared error; the caller publishes it on dr_ready, which the score warps await before their winner phase.
  uint32_t w[32];
  auto sub_pack = [&](int f, float4 (&qr)[8], int half, bool done, int code) {
    const int sw = rs_swz(f);
    const float* rbase = rs + f * 128 + ((m ^ (sw & 3)) << 2);
    // EMA statistics of the training forward (core_vq.py:227-228): the stage's input residual is added to its code's row
    // of embed_sum (frames with a candidate list were counted by the warp that settled them)
    const bool stats = TRAIN && p.ema_sum != nullptr && !done && f < p.tf && tile_n0 + f < p.N;
    float* srow = stats ? p.ema_sum + (size_t(s) * p.K + code) * 128 + 4 * m : nullptr;
    if (stats && m == 0) atomicAdd(p.ema_counts + size_t(s) * p.K + code, 1.f);
    #pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 rv = *reinterpret_cast<const float4*>(rbase + ((i ^ (sw >> 2)) << 4));
      if (TRAIN && stats) red_add_f4(srow + 16 * i, rv);
      const float4 n = done ? rv : sub_row<STE>(rv, qr[i]);
      qr[i] = n;
      w[4 * i + 2 * half] = pack_half2(n.x, n.y);
      w[4 * i + 2 * half + 1] = pack_half2(n.z, n.w);
    }
  };
  sub_pack(fA, qa, 0, doneA, codeA);
  RVQ_TRACE3(trX, trn, u, 4);
  sub_pack(fB, qb, 1, doneB, codeB);
  RVQ_TRACE3(trX, trn, u, 5);
  if (store) {
    ptx::tmem_st_16x256b_x8(taddr, w);
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(bar_a);
  }
  RVQ_TRACE3(trX, trn, u, 6);
  float e2a, e2b;
  auto finish = [&](int f, const float4 (&nr)[8], int half, float& e2) {
    const int sw = rs_swz(f);
    float* rbase = rs + f * 128 + ((m ^ (sw & 3)) << 2);
    float e[4] = {0.f, 0.f, 0.f, 0.f};
    float sqf = 0.f;
    #pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 n = nr[i];
      *reinterpret_cast<float4*>(rbase + ((i ^ (sw >> 2)) << 4)) = n;
      e[i & 3] = residue_of(n.x, n.y, w[4 * i + 2 * half], e[i & 3]);
      e[i & 3] = residue_of(n.z, n.w, w[4 * i + 2 * half + 1], e[i & 3]);
      if (TRAIN) sqf = dot4(n, n, sqf);
    }
    e2 = (e[0] + e[1]) + (e[2] + e[3]);
    if (TRAIN && f < p.tf && tile_n0 + f < p.N) sq += sqf;      // sum((q - r)^2) of core_vq.py:319 = |new residual|^2
  };
  finish(fA, qa, 0, e2a);
  finish(fB, qb, 1, e2b);
  // exact rounding residue of the new operand rows: sum over the 4 lanes of the group
  e2a += __shfl_xor_sync(0xffffffffu, e2a, 1); e2b += __shfl_xor_sync(0xffffffffu, e2b, 1);
  e2a += __shfl_xor_sync(0xffffffffu, e2a, 2); e2b += __shfl_xor_sync(0xffffffffu, e2b, 2);
  if (m == 0) {
    float* dr2 = reinterpret_cast<float*>(ms + Sm::m_dr2);
    dr2[fA] = e2a; dr2[kM + fA] = 0.f;
    dr2[fB] = e2b; dr2[kM + fB] = 0.f;
  }
  RVQ_TRACE3(trX, trn, u, 6);
}

}  // namespace

// PC: some stage of the call certifies with the per-code bound (StageMeta::percode; bit s of pcmask = stage stage0 + s).  The
// kernel picks the specialisation at its top, so a stack on the per-stage bound runs code that does not contain the per-code
// branch at all (its mere presence in the winner phase cost 1.8 % at cfg2).
template <bool TRAIN, bool PC, bool STE>
__device__ __forceinline__ void tc_encode_body(const TcParams& p, const unsigned long long pcmask) {
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
    for (int i = 0; i < kAccBufs; ++i) { ptx::mbar_init(RVQ_BAR(acc_full, i), 1); ptx::mbar_init(RVQ_BAR(acc_empty, i), 4); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(RVQ_BAR(a_ready, i), kUpdWarps); ptx::mbar_init(RVQ_BAR(cand_ready, i), 4);
                                  ptx::mbar_init(RVQ_BAR(dr_ready, i), kUpdWarps); }
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < 2) {
    int* qc = reinterpret_cast<int*>(smem + Sm::misc + threadIdx.x * Sm::m_size + Sm::m_qcnt);
    qc[0] = 0; qc[1] = 0;
  }
  // augmented K block of A: k-group 0 of each slot = (1, 1, 0, ...) picks up hi/lo of |c|^2 (the per-frame bounds of |r|
  // are written per stage), k-group 1 = 0
  for (int i = threadIdx.x; i < 6144 / 16; i += blockDim.x)
    *reinterpret_cast<uint4*>(smem + Sm::aug + i * 16) = make_uint4(i < 256 ? pack_half2(1.f, 1.f) : 0u, 0u, 0u, 0u);
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
  long long* s_t0 = reinterpret_cast<long long*>(smem + Sm::bars + 216);      // padding of the barrier block
  if (threadIdx.x == 0) *s_t0 = clock64();
  __syncthreads();
  const long long t_kernel0 = *s_t0;
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
        int st = n % p.n_q;
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
      // ===== MMA issuer: per chunk 8 MMAs with A from tensor memory + 1 with the constant shared-memory block, three at a
      // time as their K-third lands.  The loop is warp-uniform (every lane waits and forms the same descriptors, one elected
      // lane issues), so the operands of tcgen05.mma stay in uniform registers and a chunk costs a few dozen instructions. =====
      constexpr uint32_t idesc = ptx::umma_idesc_f16_f32(kM, kN);
      const uint32_t tmem_u = __reduce_max_sync(0xffffffffu, tmem);      // a provably warp-uniform copy (uniform register)
      // slot 0: k-group 0 at +0, slot 1: at +2048; the shared zero k-group 1 at +4096 (LBO 4096 / 2048)
      const uint64_t ad_aug0 = ptx::umma_desc_kmajor_noswz(sb + Sm::aug, 4096, 128);
      const uint64_t ad_aug1 = ptx::umma_desc_kmajor_noswz(sb + Sm::aug + 2048, 2048, 128);
      const uint64_t bd0 = ptx::umma_desc_kmajor_noswz(sb + Sm::ring, kTcLBO, kTcSBO);
      const uint32_t bar_full0 = sb + Sm::bars + uint32_t(offsetof(Bars, full));
      const uint32_t bar_empty0 = sb + Sm::bars + uint32_t(offsetof(Bars, empty));
      const uint32_t bar_accf0 = sb + Sm::bars + uint32_t(offsetof(Bars, acc_full));
      const uint32_t bar_acce0 = sb + Sm::bars + uint32_t(offsetof(Bars, acc_empty));
      uint32_t slot = 0, sph = 0;              // ring position of the next K-third and the parity of its `full` barrier
      uint32_t ab = 0, aph = 1;                // accumulator buffer of the next chunk, parity its `empty` barrier must have passed
      static_assert(((2 * kTcLBO) >> 4) == 256, "B descriptor step of one K=16 MMA");
      for (int n = 0; n < steps0; ++n) {
        for (int X = 0; X < 2; ++X) {
          if (X == 1 && n >= steps1) break;
          ptx::mbar_wait(sb + Sm::bars + uint32_t(offsetof(Bars, a_ready)) + 8u * X, uint32_t(n) & 1);   // fp16 operand of this step is in TMEM
          ptx::tc_fence_after();
          RVQ_TRACE(X, n, 0, lane == 0);
          const uint32_t a_tmem = tmem_u + kTmemA + 64 * X;
          const uint64_t ad_aug = X ? ad_aug1 : ad_aug0;
          #pragma unroll 1
          for (int c = 0; c < nchunks; ++c) {
            const uint32_t d_tmem = tmem_u + ab * kN;
            ptx::mbar_wait(bar_acce0 + 8u * ab, aph);                               // accumulator drained
            ptx::tc_fence_after();
            if (lane == 0) RVQ_TRACE2(X, n, 8 * c + 0);
            #pragma unroll
            for (int h = 0; h < 3; ++h) {
              ptx::mbar_wait(bar_full0 + 8u * slot, sph);                           // K-third landed
              const uint64_t bs = bd0 + uint64_t(slot * (kSlotBytes >> 4));
              if (h < 2) ptx::umma_third_ts_w(d_tmem, a_tmem + 24 * h, bs, idesc, h == 0 ? 0u : 1u, bar_empty0 + 8u * slot);
              else ptx::umma_third_last_w(d_tmem, a_tmem + 48, ad_aug, bs, idesc, bar_empty0 + 8u * slot, bar_accf0 + 8u * ab);
              if (++slot == kRing) { slot = 0; sph ^= 1; }
            }
            if (lane == 0) RVQ_TRACE2(X, n, 8 * c + 4);
            if (c == 0) RVQ_TRACE(X, n, 1, lane == 0);
            if (++ab == kAccBufs) { ab = 0; aph ^= 1; }
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
    const int f = q * 32 + lane;               // thread <-> frame mapping of tile loads and operand stores
    const uint32_t tq = tmem + (uint32_t(q * 32) << 16) + kTmemA;
#ifdef RVQ_TC_TIMERS
    uint32_t t_wait = 0, t_upd = 0, t_tr = 0, t_load = 0, t_stw = 0, t_bar = 0;
#endif
    // load dims 64h..64h+63 of the latent tile `tile` into slot X: fp32 residual rows, |x|^2, fp16 operand in tensor
    // memory, rounding residue
    // pull this warp's share of a tile (64 lines: one per dim, 32 consecutive frames x 4 B) into L2
    auto prefetch_tile = [&](int tile) {
      const int64_t n = int64_t(tile) * p.tf + f;
      if (f < p.tf && n < p.N) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x + p.fa.base(n - lane) + int64_t(h * 64 + lane) * p.fa.sxd));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x + p.fa.base(n - lane) + int64_t(h * 64 + 32 + lane) * p.fa.sxd));
      }
    };
    const int pc_first = PC ? int(pcmask & 1ull) : 0;
    auto load_tile = [&](int X, int tile) {
      float* rs = reinterpret_cast<float*>(smem + Sm::rs + X * kRsBytes);
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
      // per-code bound: this half's |x| (rounded up to fp16) goes to column 2 + 2h of the frame's augmented operand row;
      // the two halves meet -g16_k in columns 2 and 4 of the image, and sqrt(a) + sqrt(b) >= sqrt(a + b) = |x|
      *reinterpret_cast<uint32_t*>(smem + Sm::aug + X * 2048 + f * 16 + 4 + 4 * h) =
          pc_first ? uint32_t(__half_as_ushort(__float2half_ru(sqrtf(xsum) * 1.0001f))) : 0u;
      ptx::fence_proxy_async_smem();
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) { ptx::mbar_arrive(RVQ_BAR(a_ready, X)); ptx::mbar_arrive(RVQ_BAR(dr_ready, X)); }
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
          if (u == 0) ptx::mbar_wait(RVQ_BAR(cand_ready, X), uint32_t(n) & 1);
          ptx::named_bar_sync(8, kUpdWarps * 32);
          RVQ_TICK(t_wait);
          RVQ_TRACE(X, n, 6, u == 0 && lane == 0);
          const bool last = s + 1 == p.n_q;
          // a few stages before the tile ends: the slot's next tile starts its way from HBM to L2
          if ((s + 4 == p.n_q || (p.n_q < 4 && s == 0)) && (jt + 1) * p.n_q < (X ? steps1 : steps0)) prefetch_tile(tile0 + X + 2 * (jt + 1));
          float sq = 0.f;
          // operand rows of this warp: TMEM lanes 32q + 16h .. +15, columns of slot X
          update_pass<TRAIN, STE>(p, rs, ms, u, lane, s, rot, nchunks, tile_n0, pv.tab32(st), pv.cnorm(st),
                             tmem + (uint32_t(q * 32 + h * 16) << 16) + kTmemA + 64 * X, !last, RVQ_BAR(a_ready, X), sq, X, n,
#ifdef RVQ_TC_TRACE
                             t_kernel0
#else
                             0
#endif
                             );
          RVQ_TICK(t_upd);
          RVQ_TRACE(X, n, 7, u == 0 && lane == 0);
          if (threadIdx.x == 128) { int* qc = reinterpret_cast<int*>(ms + Sm::m_qcnt); qc[0] = 0; qc[1] = 0; }
          if (!last) {
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(RVQ_BAR(dr_ready, X));      // residual rows, rounding residues and queue reset are in place
            RVQ_TICK(t_stw);
            RVQ_TRACE(X, n, 8, u == 0 && lane == 0);
          } else {
            if (TRAIN && p.residual_out != nullptr) {
              // each warp writes the 16 frames it has just updated, 512 contiguous bytes per frame
              __syncwarp();
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
      atomicAdd(&p.counters[17], (unsigned long long)t_stw); atomicAdd(&p.counters[18], (unsigned long long)t_bar);
    }
#endif
  } else {
    ptx::reg_inc<152>();
    // ===== score warps =====
    const int q = warp;                        // TMEM lane quadrant = frames 32q..32q+31 of a tile
    const int f = q * 32 + lane;
    const uint32_t tlane = tmem + (uint32_t(q * 32) << 16);
    uint32_t n_cert = 0, n_resc = 0, n_full = 0, n_wide = 0, n_widec = 0;   // search statistics (rvq_search_counters)
#ifdef RVQ_TC_TIMERS
    uint32_t t_wait = 0, t_epi = 0, t_win = 0;
    const long long t_begin = clock64();
#endif
    float xx_0 = 0.f, xx_1 = 0.f;              // upper bound of |r|^2 of this thread's frame in slot 0 / 1
    // opaque copies of loop invariants: kept in registers instead of being re-derived from special registers per chunk
    uint32_t tl = tlane, bar_full = RVQ_BAR(acc_full, 0), bar_empty = RVQ_BAR(acc_empty, 0);
    asm volatile("" : "+r"(tl), "+r"(bar_full), "+r"(bar_empty));
    uint32_t acc_buf = 0, acc_ph = 0;          // accumulator buffer of the next chunk and its phase
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
        // margin coefficients of this stage: requested before the chunk loop, used after it
        const float mt_coef = __ldg(&meta->margin_coef), mt_abs = __ldg(&meta->margin_abs), mt_xlimit = __ldg(&meta->xlimit);
        const float mt_cmax = __ldg(&meta->cmax_all), mt_dr = __ldg(&meta->margin_dr);
        // per-code bound (StageMeta): this stage's switch and the next stage's (its operand's bound of |r| is written at the
        // end of this one), from the mask gathered at kernel start
        const int pc = PC && s < 64 ? int(pcmask >> s) & 1 : 0, pc_next = PC && s < 63 ? int(pcmask >> (s + 1)) & 1 : 0;
        RVQ_TICK0();
        // ---- scores: per-class and per-batch minima of the K approximate scores of this frame ----
        float cm[32], bmin[32];
        #pragma unroll
        for (int j = 0; j < 32; ++j) { cm[j] = inf_f(); bmin[j] = inf_f(); }
        // one rolled iteration per 128-code chunk (the hot loops of a stage must stay inside the instruction cache);
        // bmin is a shift register: after the loop the a-th batch in processing order sits at 32 - 4*nchunks + a.
        // A chunk is read as four pairs of 16-column loads (columns c and c + 32: the same 16 classes of two batches);
        // the next pair is in flight while the minima of the current one are taken.
        #pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
          ptx::mbar_wait(bar_full + 8u * acc_buf, acc_ph);
          ptx::tc_fence_after();
          RVQ_TICK(t_wait);
          if (c == 0) RVQ_TRACE(X, n, 3, warp == 0 && lane == 0);
          if (warp == 0 && lane == 0) RVQ_TRACE2(X, n, 64 + 3 * c);
          const uint32_t ta = tl + acc_buf * kN;
          uint32_t x0[16], x1[16], y0[16], y1[16];
          ptx::tmem_ld16(ta, x0);
          ptx::tmem_ld16(ta + 32, x1);
          #pragma unroll
          for (int j = 0; j < 28; ++j) bmin[j] = bmin[j + 4];
          ptx::tmem_ld_wait();
          ptx::tmem_ld16(ta + 16, y0);
          ptx::tmem_ld16(ta + 48, y1);
          float ba, bb;
          pair_min(x0, x1, cm, ba, bb);
          ptx::tmem_ld_wait();
          ptx::tmem_ld16(ta + 64, x0);
          ptx::tmem_ld16(ta + 96, x1);
          {
            float ca, cb;
            pair_min(y0, y1, cm + 16, ca, cb);
            bmin[28] = fminf(ba, ca); bmin[29] = fminf(bb, cb);
          }
          ptx::tmem_ld_wait();
          ptx::tmem_ld16(ta + 80, y0);
          ptx::tmem_ld16(ta + 112, y1);
          pair_min(x0, x1, cm, ba, bb);
          ptx::tmem_ld_wait();
          // scores are in registers: hand the accumulator back before reducing the last pair
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_empty + 8u * acc_buf);
          if (warp == 0 && lane == 0) RVQ_TRACE2(X, n, 64 + 3 * c + 1);
          {
            float ca, cb;
            pair_min(y0, y1, cm + 16, ca, cb);
            bmin[30] = fminf(ba, ca); bmin[31] = fminf(bb, cb);
          }
          if (++acc_buf == kAccBufs) { acc_buf = 0; acc_ph ^= 1; }
          RVQ_TICK(t_epi);
          if (warp == 0 && lane == 0) RVQ_TRACE2(X, n, 64 + 3 * c + 2);
        }
        RVQ_TRACE(X, n, 4, warp == 0 && lane == 0);
        // |x|^2 of a new tile and the rounding residue of this frame's operand were written by the update warps; the
        // scores above could only exist after they had finished
        ptx::mbar_wait(RVQ_BAR(dr_ready, X), uint32_t(n) & 1);          // (completed long ago: the update warps finish phase 2 during the MMAs)
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
        float thr = m + delta;
        if (pc) {
          // Per-code bound: the accumulator holds the lower bounds T_k = S_k - g16_k R of the true scores.  For the code k'
          // that attains the minimum of T (it must be the only one, else the per-stage bound serves):
          //   s_winner <= s_k' <= T_k' + (g16_k' + a_k') R + b_k' |r - fp16(r)| + abs =: thr,
          // and every code whose T exceeds thr is beaten by k'.
          uint32_t ce4[4] = {0u, 0u, 0u, 0u}, be4[4] = {0u, 0u, 0u, 0u};
          #pragma unroll
          for (int j = 0; j < 32; ++j) {
            // (volatile: the compiler must not hoist these out of the branch -- stages on the per-stage bound skip them)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(ce4[j & 3]) : "f"(cm[j]), "f"(m), "r"(1u << j));
            asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(be4[j & 3]) : "f"(bmin[j]), "f"(m), "r"(1u << j));
          }
          const uint32_t ceq = (ce4[0] | ce4[1]) | (ce4[2] | ce4[3]);
          const uint32_t beq = ((be4[0] | be4[1]) | (be4[2] | be4[3])) >> (32 - 4 * nchunks);
          const bool uniq = __popc(ceq) == 1 && __popc(beq) == 1;
          const float mt_abs_pc = __ldg(&meta->abs_pc), mt_g16max = __ldg(&meta->g16max);
          const float2 ab = __ldg(pv.gab(st) + (uniq ? code_of(__ffs(beq) - 1, __ffs(ceq) - 1, rot, nchunks) : 0));
          const uint4 arow = *reinterpret_cast<const uint4*>(smem + Sm::aug + X * 2048 + f * 16);
          const float Rm = __half2float(__ushort_as_half((unsigned short)(arow.y & 0xffffu))) +
                           __half2float(__ushort_as_half((unsigned short)(arow.z & 0xffffu)));      // the R the tensor core multiplied with
          thr = uniq ? fmaf(ab.x, Rm, fmaf(ab.y, drn, m + mt_abs_pc)) : m + fmaf(mt_g16max, Rm, delta);
        }
        uint32_t cm4[4] = {0u, 0u, 0u, 0u}, bm4[4] = {0u, 0u, 0u, 0u};
        #pragma unroll
        for (int j = 0; j < 32; ++j) {                // two instructions per value: compare, predicated OR with an immediate
          asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(cm4[j & 3]) : "f"(cm[j]), "f"(thr), "r"(1u << j));
          asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(bm4[j & 3]) : "f"(bmin[j]), "f"(thr), "r"(1u << j));
        }
        const uint32_t cmask = (cm4[0] | cm4[1]) | (cm4[2] | cm4[3]);
        const uint32_t bmask = ((bm4[0] | bm4[1]) | (bm4[2] | bm4[3])) >> (32 - 4 * nchunks);   // bit a = a-th batch processed
        const int nc = __popc(cmask), nb = __popc(bmask);
        const bool full = outl || cmask == 0u || bmask == 0u;     // masks are empty only for NaN scores
        const int ncand = nc * nb;
        // bmask bit a = a-th batch in this CTA's processing order (code_of)

        // first candidate (= the winner when certified); the update warps enumerate the other codes of a short list
        // from the two masks
        // ... for the usual short list (at most two flagged batches and classes) all of its codes: the update warps then
        // start from addresses instead of walking masks (y = -2: more than two of a kind, enumerate the masks)
        int4 cd;
        {
          const int b0 = __ffs(bmask) - 1, b1 = 31 - __clz(bmask | 1u), c0 = __ffs(cmask) - 1, c1 = 31 - __clz(cmask | 1u);
          const bool two = nc == 2 && nb == 2;
          cd.x = code_of(b0, c0, rot, nchunks);
          cd.y = (nc > 2 || nb > 2) ? -2 : (nc == 2 ? code_of(b0, c1, rot, nchunks) : code_of(b1, c0, rot, nchunks));
          cd.z = two ? code_of(b1, c0, rot, nchunks) : -1;
          cd.w = two ? code_of(b1, c1, rot, nchunks) : -1;
        }
        *reinterpret_cast<int4*>(ms + Sm::m_cand + f * 16) = cd;
        ms[Sm::m_ncnt + f] = (unsigned char)(full ? kFull : (ncand > 4 ? kBig : ncand));
        *reinterpret_cast<uint32_t*>(ms + Sm::m_cmask + f * 4) = cmask;
        *reinterpret_cast<uint32_t*>(ms + Sm::m_bmask + f * 4) = bmask;
        int64_t* code_out = (f < p.tf && nfr < p.N) ? p.codes + code_index(p.bkt, p.n_q, p.fa.T, p.N, s, nfr) : nullptr;
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
            resolve_full(rs, ms, q * 32 + i, lane, p.K, t32, cn, co, TRAIN && p.direct);
          }
        }
        n_full += full ? 1u : 0u; n_cert += (!full && ncand == 1) ? 1u : 0u; n_resc += (!full && ncand > 1) ? 1u : 0u;
        if (!full && ncand > 4) { n_wide += 1u; n_widec += uint32_t(ncand); }
        // upper bound of the next residual's |r|^2 (only the margin and the validity test use it):
        // the winner's approximate score is <= m + delta and off by <= delta/2
        if (full) { const float g2 = xnorm + mt_cmax; xx = g2 * g2; }
        else if (pc) xx = fmaxf(xx + thr, 0.f) * 1.00001f + 1e-30f;      // thr bounds the winner's true score
        else xx = fmaxf(xx + m + 1.5f * delta, 0.f) * 1.00001f + 1e-30f;
        if (X) xx_1 = xx; else xx_0 = xx;
        if (s + 1 < p.n_q && (pc | pc_next)) {
          // the next stage's augmented operand row: (1, 1, R, 0, 0, 0, 0, 0), R >= |r| rounded up to fp16 (0 switches the
          // per-code term off; a row that is 0 and stays 0 is left alone: stacks with uniform norms never pay for the
          // store and its proxy fence).  All MMAs of this stage have completed (their scores were consumed above).
          const uint32_t rw = pc_next ? uint32_t(__half_as_ushort(__float2half_ru(sqrtf(xx) * 1.0001f))) : 0u;
          *reinterpret_cast<uint4*>(smem + Sm::aug + X * 2048 + f * 16) = make_uint4(pack_half2(1.f, 1.f), rw, 0u, 0u);
          ptx::fence_proxy_async_smem();
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(RVQ_BAR(cand_ready, X));    // winners and queues visible to the update warps
        RVQ_TRACE(X, n, 5, warp == 0 && lane == 0);
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
        n_wide += __shfl_xor_sync(0xffffffffu, n_wide, off);
        n_widec += __shfl_xor_sync(0xffffffffu, n_widec, off);
      }
      if (lane == 0) {
        atomicAdd(&p.counters[0], (unsigned long long)(n_cert + n_resc + n_full)); atomicAdd(&p.counters[1], (unsigned long long)n_cert);
        atomicAdd(&p.counters[2], (unsigned long long)n_resc); atomicAdd(&p.counters[3], (unsigned long long)n_full);
        atomicAdd(&p.counters[11], (unsigned long long)n_wide); atomicAdd(&p.counters[12], (unsigned long long)n_widec);
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

template <bool TRAIN>
__global__ void __launch_bounds__(kThreadsTc, 1) tc_encode_kernel(const TcParams p) {
  // which stages of this call use the per-code bound: gathered by every warp for itself (two L2 reads per lane at most).
  // Stages beyond 63 use the per-stage bound, which is valid for every stage (their operand rows then carry R = 0).
  const int lane = threadIdx.x & 31;
  PackView pv(p.pack, p.K, 128);
  const unsigned lo = __ballot_sync(0xffffffffu, lane < p.n_q && __ldg(&pv.meta(p.stage0 + lane)->percode) != 0);
  const unsigned hi = __ballot_sync(0xffffffffu, lane + 32 < p.n_q && __ldg(&pv.meta(p.stage0 + (lane + 32 < p.n_q ? lane + 32 : 0))->percode) != 0);
  const unsigned long long pcmask = (static_cast<unsigned long long>(hi) << 32) | lo;
  // (the straight-through arithmetic of core_vq.py:309 is a compile-time property of the body too: as a run-time switch
  // inside the residual update it cost the training call 3 %)
  if (TRAIN && p.ste) {
    if (pcmask != 0ull) tc_encode_body<TRAIN, true, TRAIN>(p, pcmask);
    else tc_encode_body<TRAIN, false, TRAIN>(p, 0ull);
  } else {
    if (pcmask != 0ull) tc_encode_body<TRAIN, true, false>(p, pcmask);
    else tc_encode_body<TRAIN, false, false>(p, 0ull);
  }
}

int tc_debug_trace(long long* out_host, int n) {
  const int cap = 2 * kTraceSteps * kTraceEv + 128 + 2 * 64;
  const int m = n < cap ? n : cap;
  RVQ_CUDA(cudaMemcpyFromSymbol(out_host, g_trace, size_t(m) * sizeof(long long)));
  return m;
}

int simt_quant_sum(const void* pack, int K, int D, const float* x, FrameAddr fa, int64_t N, int T, int stage0, int n_q,
                   const int64_t* codes, float* out, int flags, cudaStream_t st);

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
  TcParams p;
  p.pack = (const unsigned char*)a.pack; p.K = a.K;
  p.x = a.x; p.fa = FrameAddr{a.sxb, a.sxd, a.sxt, a.T}; p.N = N;
  p.stage0 = a.stage0; p.n_q = a.n_q;
  p.codes = a.codes; p.residual_out = a.residual_out; p.sqerr = a.sqerr;
  p.ema_counts = a.ema_counts; p.ema_sum = a.ema_sum;
  p.ste = (a.flags & RVQ_FLAG_STE) ? 1 : 0;
  p.bkt = (a.flags & RVQ_FLAG_CODES_BKT) ? 1 : 0;
  p.direct = (a.flags & RVQ_FLAG_DIRECT_DIST) ? 1 : 0;
  p.counters = search_counters();     // nullptr unless the caller registered a buffer (rvq_search_counters)
  // full 128-frame tiles: a CTA's odd last tile runs alone (measured on B200 at cfg2: 0.62 ms against 0.67 ms for balanced
  // 82-frame tiles, whose MMAs cost the same as full ones)
  const int64_t tf = kM;
  p.tf = int(tf);
  const int64_t ntiles = (N + tf - 1) / tf;
  // one tile per CTA while there are SMs to spare (a lone tile's stage is shorter than a pair's: small calls are latency-bound),
  // two or more tiles per CTA beyond that
  const unsigned grid = unsigned(ntiles < sm_count ? ntiles : sm_count);
  // the lean variant serves plain encodes; straight-through arithmetic, loss numerators and the residual output
  // live in the other one (a stage's hot code has to fit the instruction cache)
  if (p.ste || p.direct || p.sqerr != nullptr || p.residual_out != nullptr || p.ema_sum != nullptr) tc_encode_kernel<true><<<grid, kThreadsTc, Sm::total, st>>>(p);
  else tc_encode_kernel<false><<<grid, kThreadsTc, Sm::total, st>>>(p);
  RVQ_LAUNCH_CHECK("tc_encode_kernel");
  if (a.quantized != nullptr)
    return simt_quant_sum(a.pack, a.K, a.D, a.x, p.fa, N, a.T, a.stage0, a.n_q, a.codes, a.quantized, a.flags, st);
  return RVQ_OK;
}

}  // namespace rvq

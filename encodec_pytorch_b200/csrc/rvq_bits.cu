// Code bit-packing on the device: the byte stream binary.BitPacker produces when driven by the loop
// of compress.compress_to_file (binary.py:69-87, compress.py:70-92) and its inverse (binary.py:104-121), one stream per
// batch item.  The stream is little-endian in bits: value number i = t*K + k (time-major, codebook-minor) occupies stream
// bits [i*bits, (i+1)*bits); a trailing partial byte is zero-padded.  HBM-bound byte work: 8 B of int64 code in, bits/8 B out
// per value (10.25 B at bits = 10).
#include "rvq_common.cuh"

namespace rvq {
namespace {

constexpr int kBT = 128;            // time steps per block: a multiple of 8, so every block starts on a byte of the stream
constexpr int kBitsThreads = 256;

// grid (ceil(T / kBT), B).  Phase 1 reads the block's codes coalesced along t (the layout the search writes: [K, B, T]) into
// shared memory in stream order; phase 2 assembles 32-bit words of the stream.
__global__ void __launch_bounds__(kBitsThreads)
bitpack_kernel(const int64_t* __restrict__ codes, int64_t sq, int64_t sb, int64_t st, int K, int T, int bits,
               unsigned char* __restrict__ out, int64_t out_stride) {
  extern __shared__ unsigned short vals[];                 // [kBT][K + 1]
  const int b = blockIdx.y, t0 = blockIdx.x * kBT, nt = min(kBT, T - t0), ld = K + 1;
  const unsigned mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
  {
    // thread = (time step, every other codebook): 128 consecutive time steps per codebook row are one coalesced kilobyte;
    // eight independent loads in flight per thread
    const int tt = threadIdx.x & (kBT - 1);
    if (tt < nt) {
      const int64_t* src = codes + int64_t(b) * sb + int64_t(t0 + tt) * st;
      #pragma unroll 8
      for (int k = threadIdx.x >> 7; k < K; k += kBitsThreads / kBT)
        vals[tt * ld + k] = (unsigned short)(unsigned(__ldg(src + int64_t(k) * sq)) & mask);
    }
  }
  __syncthreads();
  const int64_t nvals = int64_t(nt) * K, nbits = nvals * bits, nbytes = (nbits + 7) >> 3;
  const int nwords = int((nbits + 31) >> 5);
  unsigned char* o = out + int64_t(b) * out_stride + ((int64_t(t0) * K * bits) >> 3);
  const bool aligned = (reinterpret_cast<uintptr_t>(o) & 3) == 0;
  for (int w = threadIdx.x; w < nwords; w += blockDim.x) {
    const int64_t bit0 = int64_t(w) << 5;
    int i = int(bit0 / bits);
    const int i1 = int(min((bit0 + 31) / bits, nvals - 1));
    int tt = i / K, k = i - tt * K;
    unsigned long long acc = 0ull;
    for (; i <= i1; ++i) {
      const unsigned long long v = vals[tt * ld + k];
      const int sh = int(int64_t(i) * bits - bit0);          // negative only for the value that straddles the word's start
      acc |= sh >= 0 ? (v << sh) : (v >> (-sh));
      if (++k == K) { k = 0; ++tt; }
    }
    const unsigned word = unsigned(acc);
    const int64_t byte0 = int64_t(w) << 2;
    if (aligned && byte0 + 4 <= nbytes) *reinterpret_cast<unsigned*>(o + byte0) = word;
    else {
      #pragma unroll
      for (int j = 0; j < 4; ++j) if (byte0 + j < nbytes) o[byte0 + j] = (unsigned char)(word >> (8 * j));
    }
  }
}

// grid (ceil(T / kBT), B): thread = one value, t fastest so the int64 stores are coalesced along t
__global__ void __launch_bounds__(kBitsThreads)
bitunpack_kernel(const unsigned char* __restrict__ in, int64_t in_stride, int K, int T, int bits,
                 int64_t* __restrict__ codes, int64_t sq, int64_t sb, int64_t st) {
  const int b = blockIdx.y, t0 = blockIdx.x * kBT, nt = min(kBT, T - t0);
  const unsigned mask = (1u << bits) - 1u;
  const unsigned char* s = in + int64_t(b) * in_stride;
  const int64_t nbytes = (int64_t(T) * K * bits + 7) >> 3;
  for (int idx = threadIdx.x; idx < K * kBT; idx += blockDim.x) {
    const int k = idx / kBT, tt = idx - k * kBT;
    if (tt >= nt) continue;
    const int64_t bit0 = (int64_t(t0 + tt) * K + k) * bits, byte0 = bit0 >> 3;
    unsigned v = s[byte0];
    if (byte0 + 1 < nbytes) v |= unsigned(s[byte0 + 1]) << 8;
    if (byte0 + 2 < nbytes) v |= unsigned(s[byte0 + 2]) << 16;
    codes[int64_t(k) * sq + int64_t(b) * sb + int64_t(t0 + tt) * st] = int64_t((v >> int(bit0 & 7)) & mask);
  }
}

}  // namespace
}  // namespace rvq

extern "C" {

size_t rvq_bitpack_bytes(int n_q, int T, int bits) {
  if (n_q <= 0 || T <= 0 || bits <= 0) return 0;
  return size_t((int64_t(n_q) * T * bits + 7) >> 3);
}

int rvq_bitpack(const int64_t* codes, int64_t sq, int64_t sb, int64_t st, int n_q, int B, int T, int bits,
                unsigned char* out, int64_t out_stride, void* stream) {
  using namespace rvq;
  if (int e = check_device()) return e;
  RVQ_REQUIRE(n_q >= 1 && n_q <= 64 && bits >= 1 && bits <= 16 && B >= 0 && T >= 0, "rvq_bitpack: bad shape (n_q=%d bits=%d)", n_q, bits);
  if (B == 0 || T == 0) return RVQ_OK;
  RVQ_REQUIRE(codes && out && out_stride >= (int64_t)rvq_bitpack_bytes(n_q, T, bits), "rvq_bitpack: bad buffer");
  RVQ_REQUIRE(B <= 65535, "rvq_bitpack: more than 65535 streams per call");
  dim3 grid((T + kBT - 1) / kBT, B);
  bitpack_kernel<<<grid, kBitsThreads, size_t(kBT) * (n_q + 1) * 2, (cudaStream_t)stream>>>(codes, sq, sb, st, n_q, T, bits, out, out_stride);
  RVQ_LAUNCH_CHECK("bitpack_kernel");
  return RVQ_OK;
}

int rvq_bitunpack(const unsigned char* in, int64_t in_stride, int n_q, int B, int T, int bits,
                  int64_t* codes, int64_t sq, int64_t sb, int64_t st, void* stream) {
  using namespace rvq;
  if (int e = check_device()) return e;
  RVQ_REQUIRE(n_q >= 1 && bits >= 1 && bits <= 16 && B >= 0 && T >= 0, "rvq_bitunpack: bad shape (n_q=%d bits=%d)", n_q, bits);
  if (B == 0 || T == 0) return RVQ_OK;
  RVQ_REQUIRE(codes && in && in_stride >= (int64_t)rvq_bitpack_bytes(n_q, T, bits), "rvq_bitunpack: bad buffer");
  RVQ_REQUIRE(B <= 65535, "rvq_bitunpack: more than 65535 streams per call");
  dim3 grid((T + kBT - 1) / kBT, B);
  bitunpack_kernel<<<grid, kBitsThreads, 0, (cudaStream_t)stream>>>(in, in_stride, n_q, T, bits, codes, sq, sb, st);
  RVQ_LAUNCH_CHECK("bitunpack_kernel");
  return RVQ_OK;
}

}  // extern "C"

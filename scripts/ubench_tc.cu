// Micro-benchmarks that ground the design of the tcgen05 search kernel (DESIGN.md cites the numbers):
//   1. tcgen05.ld throughput (32x32b.x32) with 4 / 8 warps per SM
//   2. tcgen05.mma issue rate: M=128, N in {64,128,256}, K=16 steps; A from smem (no-swizzle / 128B
//      swizzle descriptors) or from TMEM; optionally with concurrent tcgen05.ld + FMNMX epilogue warps
//   3. mbarrier try_wait latency on an already completed phase
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tc ubench_tc.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../encodec_pytorch_b200/csrc/rvq_ptx.cuh"

using namespace rvq;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr >> 4) & 0x3FFF);
  d |= uint64_t(1) << 16;              // LBO (ignored for swizzled K-major)
  d |= uint64_t(1024 >> 4) << 32;      // SBO = 8 rows x 128 B
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;              // SWIZZLE_128B
  return d;
}

// MODE: 0 = A smem no-swizzle, 1 = A and B 128B swizzle, 2 = A in TMEM (B no-swizzle), 3 = A sw128 / B noswz,
// 4 = A noswz / B sw128, 5 = A noswz LBO=144 SBO=2304, 6 = A noswz LBO=128 SBO=2048
// epi_warps: number of extra warps doing tcgen05.ld + min-reduce concurrently (0, 4, 8)
template <int MODE, int N>
__global__ void __launch_bounds__(320, 1) mma_bench(int iters, int epi_warps, int ld_only, long long* out, const unsigned char* img = nullptr, volatile int* stop = nullptr) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t tfull[3];
  __shared__ int done_flag;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = ptx::smem_u32(smem);
  if (threadIdx.x == 0) { done_flag = 0; for (int i = 0; i < 3; ++i) ptx::mbar_init(ptx::smem_u32(&tfull[i]), 1); }
  for (int i = threadIdx.x; i < (200 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&tmem_base), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base;
  long long cyc = 0;
  if (warp == 0) {
    if (lane == 0 && !ld_only) {
      constexpr uint32_t idesc = ptx::umma_idesc_f16_f32(128, N);
      const uint32_t a_addr = sbase, b_addr = sbase + 64 * 1024;
      uint64_t ad[8], bd[8];
      #pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (MODE == 0 || MODE == 4) ad[k] = ptx::umma_desc_kmajor_noswz(a_addr + k * 2 * (128 * 16), 128 * 16, 128);
        else if (MODE == 1 || MODE == 3) ad[k] = desc_sw128(a_addr + (k & 3) * 32 + (k >> 2) * 16384);
        else if (MODE == 5) ad[k] = ptx::umma_desc_kmajor_noswz(a_addr + k * 2 * 144, 144, 2304);
        else if (MODE == 6) ad[k] = ptx::umma_desc_kmajor_noswz(a_addr + k * 2 * 128, 128, 2048);
        else ad[k] = 0;
        if (MODE == 1 || MODE == 4) bd[k] = desc_sw128(b_addr + (k & 3) * 32 + (k >> 2) * 32768);
        else bd[k] = ptx::umma_desc_kmajor_noswz(b_addr + k * 2 * (N * 16), N * 16, 128);
      }
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tmem + (it & 1) * (MODE == 2 ? 128 : 256);
        #pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (MODE == 2) ptx::umma_f16_ts(d, tmem + 384 + k * 8, bd[k], idesc, k > 0);
          else ptx::umma_f16_ss(d, ad[k], bd[k], idesc, k > 0);
        }
      }
      ptx::umma_commit(ptx::smem_u32(&bar));
      ptx::mbar_wait(ptx::smem_u32(&bar), 0);
      cyc = clock64() - t0;
      out[blockIdx.x * 4 + 0] = cyc;
      *reinterpret_cast<volatile int*>(&done_flag) = 1;
    }
  } else if (warp == 1) {
    // optional TMA stream: 36 KB chunks into a 3-slot ring at smem + 90 KB, re-issued as soon as they land
    if (lane == 0 && img != nullptr) {
      uint32_t it = 0; long long bytes = 0;
      const long long t0 = clock64();
      for (int i = 0; i < 3; ++i) { ptx::mbar_expect_tx(ptx::smem_u32(&tfull[i]), 36864); ptx::bulk_g2s(sbase + 90 * 1024 + i * 36864, img + size_t(i) * 36864, 36864, ptx::smem_u32(&tfull[i])); }
      while (!*reinterpret_cast<volatile int*>(&done_flag)) {
        const uint32_t slot = it % 3, ph = (it / 3) & 1;
        ptx::mbar_wait(ptx::smem_u32(&tfull[slot]), ph);
        bytes += 36864;
        ptx::mbar_expect_tx(ptx::smem_u32(&tfull[slot]), 36864);
        ptx::bulk_g2s(sbase + 90 * 1024 + slot * 36864, img + size_t((it + 3) % 200) * 36864, 36864, ptx::smem_u32(&tfull[slot]));
        ++it;
      }
      const long long t1 = clock64();
      for (int i = 0; i < 3; ++i) { ptx::mbar_wait(ptx::smem_u32(&tfull[(it + i) % 3]), ((it + i) / 3) & 1); }
      out[blockIdx.x * 4 + 3] = bytes * 100 / (t1 - t0);
    }
  } else if (warp >= 2 && warp < 2 + epi_warps) {
    const uint32_t tl = tmem + (uint32_t((warp & 3) * 32) << 16);
    float cm[32];
    #pragma unroll
    for (int j = 0; j < 32; ++j) cm[j] = 1e30f;
    const int n = ld_only ? iters : iters * 2;
    const long long t0 = clock64();
    for (int it = 0; it < n; ++it) {
      uint32_t v0[32], v1[32];
      ptx::tmem_ld32(tl + ((it * 64) & 255), v0);
      ptx::tmem_ld32(tl + ((it * 64 + 32) & 255), v1);
      ptx::tmem_ld_wait();
      #pragma unroll
      for (int j = 0; j < 32; ++j) cm[j] = ptx::fmin3(cm[j], __uint_as_float(v0[j]), __uint_as_float(v1[j]));
    }
    const long long t1 = clock64();
    float m = 1e30f;
    #pragma unroll
    for (int j = 0; j < 32; ++j) m = fminf(m, cm[j]);
    if (m == 123.f) out[1000000] = 1;
    if (lane == 0 && warp == 2) { out[blockIdx.x * 4 + 1] = t1 - t0; out[blockIdx.x * 4 + 2] = n; }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}


// Kernel-like issue loop: per chunk 9 MMAs (8 from a 128B-swizzled A + 1 from a no-swizzle block) into one of 4
// accumulator buffers, then two tcgen05.commit (like rvq_tc.cu).  fill != 0: pseudo-random fp16 operand data.
__global__ void __launch_bounds__(320, 1) chunk_bench(int chunks, int ncommit, int fill, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar[8];
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = ptx::smem_u32(smem);
  for (int i = threadIdx.x; i < (200 * 1024) / 4; i += blockDim.x) {
    uint32_t v = 0;
    if (fill) { uint32_t h = (i * 2654435761u) ^ (i >> 7); v = ((h & 0x03ff03ffu) | 0x30003000u) ^ ((h >> 3) & 0x80008000u); }  // halves in +-[0.125, 0.25)
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) ptx::mbar_init(ptx::smem_u32(&bar[i]), 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&tmem_base), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (warp == 0 && lane == 0) {
    constexpr uint32_t idesc = ptx::umma_idesc_f16_f32(128, 128);
    uint64_t ad[9];
    #pragma unroll
    for (int k = 0; k < 8; ++k) ad[k] = desc_sw128(sbase + (k >> 2) * 16384 + (k & 3) * 32);
    ad[8] = ptx::umma_desc_kmajor_noswz(sbase + 32768, 2048, 128);
    const uint64_t bd = ptx::umma_desc_kmajor_noswz(sbase + 36864, 2048, 128);
    const long long t0 = clock64();
    for (int it = 0; it < chunks; ++it) {
      const uint32_t d = tmem + (it & 3) * 128;
      const uint64_t b0 = bd + uint64_t(((it % 3) * 36864) >> 4);
      #pragma unroll
      for (int k = 0; k < 9; ++k) ptx::umma_f16_ss(d, ad[k], b0 + uint64_t((k * 4096) >> 4), idesc, k > 0);
      if (ncommit > 0) ptx::umma_commit(ptx::smem_u32(&bar[it & 3]));
      if (ncommit > 1) ptx::umma_commit(ptx::smem_u32(&bar[4 + it % 3]));
    }
    const long long t1 = clock64();
    ptx::umma_commit(ptx::smem_u32(&bar[7]));
    // bar[7] may have received earlier commits: just spin on time
    while (clock64() - t1 < 200000) {}
    out[blockIdx.x * 4 + 0] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

__global__ void mbar_bench(long long* out) {
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
    ptx::fence_mbar_init();
    ptx::mbar_arrive(ptx::smem_u32(&bar));
    const long long t0 = clock64();
    int ok = 0;
    for (int i = 0; i < 1000; ++i) ok += ptx::mbar_try_wait(ptx::smem_u32(&bar), 0) ? 1 : 0;
    const long long t1 = clock64();
    out[0] = t1 - t0; out[1] = ok;
  }
}

static unsigned char* g_img = nullptr;
template <int MODE, int N>
void run(long long* d, int grid, int epi, const char* name, bool tma = false) {
  long long h[148 * 4];
  const int iters = 2000;
  CK(cudaFuncSetAttribute(mma_bench<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaMemset(d, 0, sizeof(h)));
  mma_bench<MODE, N><<<grid, 320, 200 * 1024>>>(iters, epi, 0, d, tma ? g_img : nullptr, nullptr);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  double mx = 0, sum = 0;
  for (int b = 0; b < grid; ++b) { sum += h[b * 4]; if (h[b * 4] > mx) mx = h[b * 4]; }
  printf("grid %3d  %-13s N=%3d epi_warps=%d : %.1f cyc per K=16 MMA (max CTA %.1f; floor %d)", grid, name, N, epi,
         sum / grid / (iters * 8.0), mx / (iters * 8.0), N / 2);
  if (epi) printf("  | epi warp: %.1f cyc per 64-col ld+min", double(h[1]) / double(h[2]));
  if (tma) printf("  | TMA stream %.1f B/cyc", h[3] / 100.0);
  printf("\n");
}
template <int MODE>
void run_mode(long long* d, const char* name) {
  for (int epi : {0, 8}) { run<MODE, 64>(d, 148, epi, name); run<MODE, 128>(d, 148, epi, name); run<MODE, 256>(d, 148, epi, name); }
}

int main() {
  long long* d; CK(cudaMalloc(&d, 8 * 1000001 + 64)); CK(cudaMemset(d, 0, 8 * 1000001 + 64));
  CK(cudaMalloc(&g_img, size_t(200) * 36864)); CK(cudaMemset(g_img, 0, size_t(200) * 36864));
  run<3, 128>(d, 148, 8, "Asw128 Bnoswz", true);
  CK(cudaFuncSetAttribute(chunk_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int fill : {0, 1}) for (int nc : {0, 1, 2}) {
    long long hh[148 * 4];
    chunk_bench<<<148, 320, 200 * 1024>>>(2000, nc, fill, d);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hh, d, sizeof(hh), cudaMemcpyDeviceToHost));
    double sum = 0; for (int b = 0; b < 148; ++b) sum += hh[b * 4];
    printf("kernel-like chunk loop: fill=%d commits/chunk=%d : %.1f cyc per 9-MMA chunk (issue side; floor 576)\n", fill, nc, sum / 148 / 2000.0);
  }
  long long h[4];
  mbar_bench<<<1, 32>>>(d);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
  printf("mbarrier.try_wait (completed phase): %.1f cyc each (%lld ok)\n", h[0] / 1000.0, h[1]);
  return 0;
}

// Micro-benchmark: how fast can all SMs stream the SAME L2-resident codebook image into shared memory
// with cp.async.bulk (1-D TMA), with and without cluster multicast?  Grounds the "codebook chunks are
// shared by several tiles" decision of the tcgen05 search.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_l2 ubench_l2.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../encodec_pytorch_b200/csrc/rvq_ptx.cuh"

using namespace rvq;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kRing = 3;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
// remote arrive on the same-offset barrier of CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(bar), "r"(cta));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}

// Every CTA streams `nchunks` chunks of `chunk` bytes (cycling over `image_bytes` of the image).
// CS = cluster size (1 = plain unicast).  With CS > 1 each CTA loads 1/CS of every chunk and multicasts
// it to all CTAs of the cluster; a ring slot is refilled only after every CTA of the cluster released it.
template <int CS>
__global__ void __launch_bounds__(64, 1) stream_kernel(const unsigned char* image, size_t image_bytes, int chunk, int nchunks, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[kRing], empty[kRing];
  const uint32_t sbase = ptx::smem_u32(smem);
  const uint32_t rank = CS > 1 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) { ptx::mbar_init(ptx::smem_u32(&full[i]), 1); ptx::mbar_init(ptx::smem_u32(&empty[i]), CS); }
    ptx::fence_mbar_init();
  }
  if (CS > 1) cluster_sync(); else __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    // producer
    size_t off = 0;
    const uint32_t part = chunk / CS;
    for (int it = 0; it < nchunks; ++it) {
      const uint32_t slot = it % kRing, ph = (it / kRing) & 1;
      ptx::mbar_wait(ptx::smem_u32(&empty[slot]), ph ^ 1);
      const uint32_t fb = ptx::smem_u32(&full[slot]);
      ptx::mbar_expect_tx(fb, chunk);
      if (CS == 1) ptx::bulk_g2s(sbase + slot * chunk, image + off, chunk, fb);
      else bulk_g2s_mc(sbase + slot * chunk + rank * part, image + off + rank * part, part, fb, uint16_t((1u << CS) - 1));
      off += chunk; if (off + chunk > image_bytes) off = 0;
    }
  } else if (threadIdx.x == 32) {
    // consumer: wait for the data, release the slot to every CTA of the cluster
    for (int it = 0; it < nchunks; ++it) {
      const uint32_t slot = it % kRing, ph = (it / kRing) & 1;
      ptx::mbar_wait(ptx::smem_u32(&full[slot]), ph);
      if (CS == 1) ptx::mbar_arrive(ptx::smem_u32(&empty[slot]));
      else for (uint32_t c = 0; c < CS; ++c) mbar_arrive_remote(ptx::smem_u32(&empty[slot]), c);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
  if (CS > 1) cluster_sync();
}

template <int CS>
void run(const unsigned char* img, size_t bytes, int chunk, int nchunks, long long* d, int grid) {
  CK(cudaFuncSetAttribute(stream_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRing * chunk));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = kRing * chunk; cfg.stream = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, stream_kernel<CS>, img, bytes, chunk, nchunks, d));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
  }
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long h[148]; CK(cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
  const double total = double(grid) * nchunks * chunk;
  printf("cluster %d grid %3d chunk %6d B: %.3f ms, smem fill %.2f TB/s (%.1f B/cyc/SM), L2 reads %.2f TB/s\n", CS, grid, chunk, ms,
         total / ms * 1e-9, double(nchunks) * chunk / avg, total / CS / ms * 1e-9);
}

int main() {
  const size_t bytes = size_t(32) * 1024 * 288;   // 32 stages x 1024 codes x 144 halves
  unsigned char* img; CK(cudaMalloc(&img, bytes)); CK(cudaMemset(img, 1, bytes));
  long long* d; CK(cudaMalloc(&d, 8 * 256));
  for (int chunk : {18432, 36864, 73728}) {
    const int nchunks = int(4 * bytes / chunk);
    run<1>(img, bytes, chunk, nchunks, d, 148);
    run<2>(img, bytes, chunk, nchunks, d, 148);
    run<4>(img, bytes, chunk, nchunks, d, 148);
  }
  return 0;
}

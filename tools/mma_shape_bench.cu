// Microbenchmark: sustained tcgen05.mma rate per instruction shape, operands in shared memory (SS),
// back-to-back issue from one thread.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// tools/mma_shape_bench.cu -o /tmp/mma_shape_bench ; run under gpurun.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../mae_clip_b200/csrc/tc_ptx.cuh"
using namespace mc::ptx;

__device__ __forceinline__ void mma_f16_1cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_1cta(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// mode 0: cta_group::2, mode 1: cta_group::1.  nb = number of distinct B tiles cycled (reuse pattern)
__global__ void __launch_bounds__(128, 1) bench(int mode, int M, int N, int iters, int pattern, int nacc, int accstride, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint64_t barmem;
  __shared__ uint32_t slot;
  const uint32_t bar = smem_u32(&barmem);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) smem_raw[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 1) {
    if (mode == 0) { tmem_alloc_pair(smem_u32(&slot), 512); tmem_relinquish_pair(); }
    else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = slot;
  const bool leader = cluster_ctarank() == 0;
  if (warp == 1 && (leader || mode == 1) && elect_one()) {
    const uint32_t idesc = idesc_f16(M, N);
    const uint64_t a0 = smem_desc_sw128(base), b0 = smem_desc_sw128(base + 16384);
    long long t0 = clock64();
    // 8 MMAs per iteration, descriptor offsets are immediates: 4 k-steps x 2 A chunks; accumulators
    // rotate over `nacc` TMEM regions (nacc in {1,2,4})
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        // pattern bit0: A fixed, bit1: B fixed, bit2: operands change only every other MMA
        const int uu = (pattern & 4) ? (u >> 1) : u;
        const uint64_t a = (pattern & 1) ? a0 : a0 + (uint64_t)(((uu >> 2) & 1) * (8192 >> 4)) + 2 * (uu & 3);
        const uint64_t b = (pattern & 2) ? b0 : b0 + (uint64_t)(((uu >> 2) & 1) * (8192 >> 4)) + 2 * (uu & 3);
        const uint32_t td = tm + (uint32_t)((u & (nacc - 1)) * accstride);
        if (mode == 0) mma_f16_pair(td, a, b, idesc, (i | u) >= nacc); else mma_f16_1cta(td, a, b, idesc, (i | u) >= nacc);
      }
    }
    if (mode == 0) mma_commit_pair(bar, 1); else commit_1cta(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    if (mode == 0) tmem_dealloc_pair(tm, 512);
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
  }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  struct C { int mode, M, N, nacc, pattern; } cfgs[] = {
      {0,128,128,2,0},{0,128,128,2,1},{0,128,128,2,2},{0,128,128,2,3},{0,128,128,2,4},
      {0,128,256,2,0},{0,128,256,2,1},{0,128,256,2,2},{0,128,256,2,3},
      {0,256,128,2,0},{0,256,128,2,1},{0,256,128,2,2},{0,256,128,2,3},
      {0,256,256,2,0},{0,256,256,2,3},
      {1,128,128,2,0},{1,128,128,2,1},{1,128,128,2,2},{1,128,128,2,3},{1,128,256,2,0},{1,128,256,2,3}};
  for (auto c : cfgs) {
    const int iters = 4096;
    const int accstride = (c.mode == 0 && c.M == 128) ? c.N / 2 : c.N;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 64 * 1024;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {2,1,1};
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, bench, c.mode, c.M, c.N, iters, c.pattern, c.nacc, accstride, d);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    double per = (double)cyc / iters;
    double macs = (double)c.M * c.N * 16 / per / (c.mode == 0 ? 2 : 1);
    printf("cta_group::%d M=%3d N=%3d pattern=%d (%s%s%s): %7.2f cyc/MMA -> %7.1f MAC/cycle/SM (%s %s)\n", c.mode == 0 ? 2 : 1, c.M, c.N,
           c.pattern, (c.pattern & 1) ? "A fixed " : "A varies ", (c.pattern & 2) ? "B fixed" : "B varies", (c.pattern & 4) ? " pairs" : "", per, macs,
           cudaGetErrorString(e), cudaGetErrorString(e2));
  }
  return 0;
}

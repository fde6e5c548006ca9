// Microbenchmark: sustained tcgen05.mma issue rate of ONE elected thread per CTA pair (cta_group::2), all 74 pairs of the
// chip busy, as a function of the instruction shape and of the ORDER in which accumulators are addressed (the sweeps of
// clip_loss_tc.cu issue S,S,S,Z / S,S,S,Z,Z,Z / ... sequences).  Operands live in shared memory (SS form).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/mma_issue_bench.cu -o tools/mma_issue_bench.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../mae_clip_b200/csrc/tc_ptx.cuh"
using namespace mc::ptx;

// SEQ: 12 accumulator indices (one period of the pattern)
template <int S0, int S1, int S2, int S3, int S4, int S5, int S6, int S7, int S8, int S9, int S10, int S11>
__device__ __forceinline__ void issue_period(uint32_t tm, uint32_t accstride, uint64_t a0, uint64_t b0, uint32_t idesc, bool first) {
  constexpr int seq[12] = {S0, S1, S2, S3, S4, S5, S6, S7, S8, S9, S10, S11};
#pragma unroll
  for (int u = 0; u < 12; ++u) {
    const uint64_t a = a0 + (uint64_t)(((u >> 2) & 1) * (8192 >> 4)) + 2 * (u & 3);
    const uint64_t b = b0 + (uint64_t)((((u + 1) >> 2) & 1) * (8192 >> 4)) + 2 * ((u + 1) & 3);
    mma_f16_pair(tm + seq[u] * accstride, a, b, idesc, first ? 0u : 1u);
  }
}

__global__ void __launch_bounds__(384, 1) bench(int M, int N, int periods, int pattern, int traffic /* 0 none, 1 st.shared, 2 ld.shared */, int traffic_warps, int extras, long long* out) {
  __shared__ uint64_t dummy_bar[2];
  if (threadIdx.x == 0) { mbar_init(smem_u32(&dummy_bar[0]), 1); mbar_init(smem_u32(&dummy_bar[1]), 1); }
  __shared__ volatile int stop_flag;
  if (threadIdx.x == 0) stop_flag = 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint64_t barmem;
  __shared__ uint32_t slot;
  const uint32_t bar = smem_u32(&barmem);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024; i += blockDim.x) smem_raw[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc_pair(smem_u32(&slot), 512); tmem_relinquish_pair(); }
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = slot;
  const bool leader = cluster_ctarank() == 0;
  if (warp == 1 && leader && elect_one()) {
    const uint32_t idesc = idesc_f16(M, N);
    const uint32_t accstride = (M == 128) ? N / 2 : N;   // TMEM columns per accumulator
    const uint64_t a0 = smem_desc_sw128(base), b0 = smem_desc_sw128(base + 32768);
    const long long t0 = clock64();
    const uint32_t db = smem_u32(&dummy_bar[0]), db1 = smem_u32(&dummy_bar[1]);
    mbar_arrive_local(db1);   // phase 0 of dummy_bar[1] is complete: waiting on parity 0 succeeds at once
    for (int i = 0; i < periods; ++i) {
      const bool f = i == 0;
      if (extras & 4) { mbar_wait(db1, 0); tc_fence_after(); }
      if (extras & 1) mma_commit_pair(db, 1);
      if (extras & 8) { mma_commit_pair(db, 3); }
      switch (pattern) {
        case 0: issue_period<0,0,0,0,0,0,0,0,0,0,0,0>(tm, accstride, a0, b0, idesc, f); break;   // one accumulator
        case 1: issue_period<0,1,0,1,0,1,0,1,0,1,0,1>(tm, accstride, a0, b0, idesc, f); break;   // alternate 2
        case 2: issue_period<0,0,0,1,1,1,0,0,0,1,1,1>(tm, accstride, a0, b0, idesc, f); break;   // S,S,S,Z,Z,Z
        case 3: issue_period<0,0,0,1,0,0,0,1,0,0,0,1>(tm, accstride, a0, b0, idesc, f); break;   // S,S,S,Z
        case 4: issue_period<0,1,2,0,1,2,0,1,2,0,1,2>(tm, accstride, a0, b0, idesc, f); break;   // rotate 3
        default: issue_period<0,0,0,0,1,1,1,1,0,0,0,0>(tm, accstride, a0, b0, idesc, f); break;  // runs of 4
      }
    }
    mma_commit_pair(bar, 1);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
    stop_flag = 1;
  } else if (warp >= 4 && warp < 4 + traffic_warps && traffic) {
    // competing shared-memory traffic (the role TMA fills and the epilogue play in the real sweeps): 16-byte accesses,
    // conflict-free, on a region the MMAs do not read
    uint4* reg = reinterpret_cast<uint4*>(smem_raw + (base - raw) + 65536) + threadIdx.x;
    uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
    unsigned long long n = 0;
    while (!stop_flag && cluster_ctarank() == 0 || (cluster_ctarank() != 0 && n < (unsigned long long)periods * 12 * 2)) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (traffic == 1) reg[(u & 1) * 384] = v;
        else { uint4 w = reg[(u & 1) * 384]; v.x ^= w.x; }
      }
      n += 8;
    }
    if (v.x == 0x12345) out[1] = (long long)v.x;
    if (threadIdx.x == 128 && blockIdx.x == 0) out[2] = (long long)n;
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 32);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char* names[] = {"one accumulator", "alternate 2", "S,S,S,Z,Z,Z", "S,S,S,Z", "rotate 3", "runs of 4"};
  int shapes[][2] = {{128, 128}, {128, 256}, {256, 128}, {256, 256}};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int extras : {0, 1, 4, 5, 8})
  for (int traffic = 0; traffic < 1; ++traffic)
  for (int tw = (traffic ? 2 : 8); tw <= 8; tw *= 2)
  for (auto& sh : shapes) {
    if (sh[0] != 128 || sh[1] != 128) continue;
    for (int pat = 3; pat < 4; ++pat) {
      const int M = sh[0], N = sh[1];
      const int accs = (M == 128 ? N / 2 : N);
      if (accs * 3 > 512 && pat == 4) continue;
      if (accs * 2 > 512 && pat != 0) continue;
      const int periods = 2048;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(148); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = 100 * 1024;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {2,1,1};
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, bench, M, N, 64, pat, traffic, tw, extras, d);   // warm-up
      cudaEventRecord(e0);
      cudaError_t e = cudaLaunchKernelEx(&cfg, bench, M, N, periods, pat, traffic, tw, extras, d);
      cudaEventRecord(e1);
      cudaError_t e2 = cudaDeviceSynchronize();
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      long long res[4] = {}; cudaMemcpy(res, d, 32, cudaMemcpyDeviceToHost);
      long long cyc = res[0];
      const double n = periods * 12.0, per = (double)cyc / n;
      const double tflops = 74.0 * n * 2.0 * M * N * 16 / (ms * 1e-3) / 1e12;
      const double tbytes = traffic ? (double)res[2] * 16.0 * 32 * tw / cyc : 0.0;   // competing bytes per cycle per SM
      printf("extras=%d M=%3d N=%3d %-10s traffic %s x%d warps (%5.1f B/clk) %7.2f cyc/MMA  %7.1f MAC/cyc/SM  %7.1f TFLOP/s chip (%.3f ms) %s %s\n", extras, M, N,
             names[pat], traffic == 0 ? "none" : (traffic == 1 ? "st.shared" : "ld.shared"), traffic ? tw : 0, tbytes, per,
             (double)M * N * 16 / per / 2, tflops, ms, cudaGetErrorString(e), cudaGetErrorString(e2));
    }
  }
  return 0;
}

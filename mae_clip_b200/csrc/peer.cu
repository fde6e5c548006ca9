// Peer-memory plumbing for the row-sharded contrastive loss (SURVEY.md section 8 e): one process
// per GPU, every rank maps every other rank's exchange region over NVLink / NVSwitch (CUDA IPC)
// and our own kernels load / store it directly.  The reference is single-process (no
// counterpart under /root/reference); the semantics are those of the all-gathers in
// mae_clip_b200/dist.py, which these kernels replace:
//   * peer_barrier_kernel  - all-to-all flag exchange with system-scope release / acquire
//   * peer_publish_kernel  - push small vectors (row statistics, loss partials, amax) into every
//                            peer's region (stores are fire-and-forget over NVLink)
// The embedding "all-gather" itself is fused into the operand staging: see
// stage_planes_peers_kernel in clip_loss_tc.cu, which pulls fp32 rows from the owning rank while
// it writes the local fp16 planes.
#include <string.h>

#include "common.cuh"

namespace mc {

constexpr int kMaxPeers = 16;
struct PeerWords {
  uint32_t* p[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// flags.p[q] = rank q's flag array (kMaxPeers words, zero at creation).  Thread q tells rank q
// "rank `rank` reached epoch e" and waits until rank q said the same here.  The epoch lives in
// device memory (one private word per rank, bumped by the kernel itself: every rank runs the same
// sequence of barriers), so a launch carries no step-dependent argument and the whole step can be
// captured in a CUDA graph.  Epochs only grow, so the arrays never need a reset.  A peer that never
// arrives trips the timeout: the kernel records 1 + (the rank it was waiting for) in epoch_counter[1]
// and RETURNS - no __trap, which would poison the CUDA context with a sticky error; the host reads the
// word (PeerExchange.check() / the NaN loss of PeerStep) and can tear the exchange down and go on over NCCL.
__global__ void __launch_bounds__(32) peer_barrier_kernel(PeerWords flags, int rank, int world,
                                                          uint32_t* __restrict__ epoch_counter,
                                                          unsigned long long timeout_ns) {
  const int q = threadIdx.x;
  uint32_t epoch = 0;
  if (q == 0) {
    epoch = *epoch_counter + 1;
    *epoch_counter = epoch;
  }
  epoch = __shfl_sync(0xffffffffu, epoch, 0);
  if (q < world) {
    __threadfence_system();  // order this GPU's earlier peer stores before the flag
    st_release_sys(flags.p[q] + rank, epoch);
    const uint32_t* mine = flags.p[rank] + q;
    const unsigned long long t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      if (globaltimer_ns() - t0 > timeout_ns) {
        printf("mae_clip_b200: peer barrier timed out (rank %d waiting for rank %d, epoch %u)\n", rank, q, epoch);
        atomicMax(epoch_counter + 1, (uint32_t)(q + 1));
        break;
      }
      __nanosleep(64);
    }
  }
  __syncwarp();
  __threadfence_system();
}

// dst.p[q][dst_offset + kk * dst_stride + i] = src[kk * src_stride + i]   for every peer q
__global__ void __launch_bounds__(256) peer_publish_kernel(const uint32_t* __restrict__ src, int k, int n,
                                                           int64_t src_stride, PeerWords dst, int64_t dst_stride,
                                                           int64_t dst_offset, int world) {
  const int q = blockIdx.y;
  if (q >= world) return;
  uint32_t* d = dst.p[q] + dst_offset;
  const int64_t total = (int64_t)k * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t kk = idx / n, i = idx - kk * n;
    d[kk * dst_stride + i] = src[kk * src_stride + i];
  }
}

// out[i] = sum_q src.p[q][i]: the reduce-scatter step of the stored-weights gradient (every rank pulls ITS rows of the
// peers' partial dI straight out of their memory: 16-byte loads, all `world` of them in flight before the first add;
// a fixed summation order, so every run gives the same bits)
__global__ void __launch_bounds__(256) peer_reduce_kernel(PeerWords src, int world, size_t n4, float4* __restrict__ out,
                                                          const int* __restrict__ gate) {
  if (gate != nullptr && *gate != 1) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v[kMaxPeers];
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q)
      if (q < world) v[q] = reinterpret_cast<const float4*>(src.p[q])[i];
    float4 a = v[0];
#pragma unroll
    for (int q = 1; q < kMaxPeers; ++q)
      if (q < world) { a.x += v[q].x; a.y += v[q].y; a.z += v[q].z; a.w += v[q].w; }
    out[i] = a;
  }
}

static int pack(PeerWords& w, void* const* host_ptrs, int world, const char* who) {
  MC_REQUIRE(host_ptrs != nullptr && world >= 1 && world <= kMaxPeers, MC_ERR_BAD_ARG,
             "%s: world %d outside [1, %d] or null pointer table", who, world, kMaxPeers);
  for (int q = 0; q < kMaxPeers; ++q) w.p[q] = nullptr;
  for (int q = 0; q < world; ++q) {
    MC_REQUIRE(host_ptrs[q] != nullptr && aligned(host_ptrs[q], 4), MC_ERR_BAD_ARG, "%s: peer pointer %d is null/misaligned",
               who, q);
    w.p[q] = static_cast<uint32_t*>(host_ptrs[q]);
  }
  return MC_OK;
}

}  // namespace mc

using namespace mc;

extern "C" {

int mc_peer_alloc(size_t bytes, void** dev_ptr_out, void* ipc_handle_out) {
  MC_ARCH_GUARD();
  MC_REQUIRE(bytes > 0 && dev_ptr_out && ipc_handle_out, MC_ERR_BAD_ARG, "peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == MC_PEER_HANDLE_BYTES, "handle size");
  void* p = nullptr;
  MC_CUDA(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("peer_alloc: %s", cudaGetErrorString(e));
    return MC_ERR_CUDA;
  }
  memcpy(ipc_handle_out, &h, sizeof(h));
  *dev_ptr_out = p;
  return MC_OK;
}

int mc_peer_open(const void* ipc_handle, void** dev_ptr_out) {
  MC_ARCH_GUARD();
  MC_REQUIRE(ipc_handle && dev_ptr_out, MC_ERR_BAD_ARG, "peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  void* p = nullptr;
  MC_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr_out = p;
  return MC_OK;
}

int mc_peer_close(void* dev_ptr) {
  if (!dev_ptr) return MC_OK;
  MC_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return MC_OK;
}

int mc_peer_free(void* dev_ptr) {
  if (!dev_ptr) return MC_OK;
  MC_CUDA(cudaFree(dev_ptr));
  return MC_OK;
}

int mc_peer_barrier(void* const* flag_ptrs_host, int rank, int world, unsigned int* epoch_counter,
                    double timeout_s, void* stream) {
  MC_ARCH_GUARD();
  PeerWords w;
  int rc = pack(w, flag_ptrs_host, world, "peer_barrier");
  if (rc) return rc;
  MC_REQUIRE(rank >= 0 && rank < world && epoch_counter, MC_ERR_BAD_ARG,
             "peer_barrier: rank %d outside world %d or null epoch counter", rank, world);
  if (timeout_s <= 0.0) timeout_s = 600.0;  // the order of a process-group timeout, not of a step
  peer_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(w, rank, world, epoch_counter,
                                                                       (unsigned long long)(timeout_s * 1e9));
  MC_LAUNCH_CHECK();
  return MC_OK;
}

int mc_peer_publish(const void* src, int k, int n, int64_t src_stride, void* const* dst_ptrs_host,
                    int64_t dst_stride, int64_t dst_offset, int world, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(src && aligned(src, 4) && k > 0 && n > 0 && dst_offset >= 0, MC_ERR_BAD_ARG, "peer_publish: bad argument");
  PeerWords w;
  int rc = pack(w, dst_ptrs_host, world, "peer_publish");
  if (rc) return rc;
  const int64_t total = (int64_t)k * n;
  int bx = (int)((total + 255) / 256);
  if (bx > 64) bx = 64;
  dim3 grid(bx, world);
  peer_publish_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint32_t*>(src), k, n,
                                                                          src_stride, w, dst_stride, dst_offset, world);
  MC_LAUNCH_CHECK();
  return MC_OK;
}

int mc_peer_reduce(void* const* src_ptrs_host, int world, size_t n_floats, float* out, const int* gate, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(out && aligned(out, 16) && n_floats > 0 && n_floats % 4 == 0, MC_ERR_BAD_ARG, "peer_reduce: bad argument");
  PeerWords w;
  int rc = pack(w, src_ptrs_host, world, "peer_reduce");
  if (rc) return rc;
  for (int q = 0; q < world; ++q) MC_REQUIRE(aligned(src_ptrs_host[q], 16), MC_ERR_ALIGN, "peer_reduce: source %d not 16-byte aligned", q);
  const size_t n4 = n_floats / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > num_sms() * 4) blocks = num_sms() * 4;
  peer_reduce_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, world, n4, reinterpret_cast<float4*>(out), gate);
  MC_LAUNCH_CHECK();
  return MC_OK;
}

}  // extern "C"

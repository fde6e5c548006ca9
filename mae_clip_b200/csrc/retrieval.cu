// "Next" row 3 (SURVEY.md section 8 f): inference-side similarity + top-k retrieval.
// Replaces /root/reference inference.py:42-46 (find_matches) and the same pattern in
// CIFAR.ipynb / classifier.ipynb:
//     image_n = F.normalize(image_embeddings, p=2, dim=-1); text_n = F.normalize(text_embeddings, p=2, dim=-1)
//     dot_similarity = text_n @ image_n.T;  values, indices = torch.topk(dot_similarity, k)
// Eager PyTorch writes a normalised copy of the whole image bank, then a GEMM, then a sort-based
// top-k.  Here the bank is streamed ONCE (HBM-bound: N*D*4 bytes): one warp per image row computes
// the row norm and its dot product with every (pre-normalised, shared-memory resident) query; the
// (Q, N) scores then go through an exact radix select per query.  Keys are 64-bit
// (order-preserving score bits << 32 | ~index), hence unique: the k-th largest key is an exact
// threshold, ties resolve to the LOWER index, and the result is deterministic.
#include "common.cuh"

namespace mc {

constexpr int kMaxQ = 32;      // queries per streaming pass (host loops over larger Q)
constexpr int kMaxK = 1024;    // top-k kept in shared memory
constexpr float kNormEps = 1e-12f;  // F.normalize default eps

__device__ __forceinline__ uint32_t desc_bits(float f) {  // larger float -> larger unsigned; NaN largest (torch.topk)
  if (f != f) return 0xFFFFFFFFu;
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// text (Q, D) -> normalised queries (Q, D): one warp per query
__global__ void __launch_bounds__(256) normalize_rows_kernel(const float* __restrict__ x, int rows, int D,
                                                             float* __restrict__ out) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= rows) return;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) { const float v = x[(size_t)r * D + k]; s = fmaf(v, v, s); }
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(s)), kNormEps);
  for (int k = lane; k < D; k += 32) out[(size_t)r * D + k] = x[(size_t)r * D + k] * inv;
}

// scores[q][n] = <t_n[q], x[n]> / max(||x[n]||, eps).  kVec float4 per lane cover a row (D = 128 * kVec).
template <int kVec>
__global__ void __launch_bounds__(256) sim_scores_kernel(const float* __restrict__ tn, int Q, const float* __restrict__ x,
                                                         long long N, float* __restrict__ scores) {
  constexpr int D = 128 * kVec;
  extern __shared__ __align__(16) float sq[];  // [Q][D] normalised queries
  for (int i = threadIdx.x; i < Q * D / 4; i += blockDim.x)
    reinterpret_cast<float4*>(sq)[i] = reinterpret_cast<const float4*>(tn)[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long n = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps) {
    float4 v[kVec];
#pragma unroll
    for (int u = 0; u < kVec; ++u) v[u] = ld_stream(reinterpret_cast<const float4*>(x + n * D) + lane + 32 * u);
    float nn = 0.f;
#pragma unroll
    for (int u = 0; u < kVec; ++u) nn += (v[u].x * v[u].x + v[u].y * v[u].y) + (v[u].z * v[u].z + v[u].w * v[u].w);
    const float inv = 1.f / fmaxf(sqrtf(warp_sum(nn)), kNormEps);
    for (int q = 0; q < Q; ++q) {
      float d = 0.f;
#pragma unroll
      for (int u = 0; u < kVec; ++u) {
        const float4 t = reinterpret_cast<const float4*>(sq + q * D)[lane + 32 * u];
        d += (v[u].x * t.x + v[u].y * t.y) + (v[u].z * t.z + v[u].w * t.w);
      }
      d = warp_sum(d);
      if (lane == 0) scores[(size_t)q * N + n] = d * inv;
    }
  }
}

// generic D (any multiple of 4 is not required): scalar loads
__global__ void __launch_bounds__(256) sim_scores_generic_kernel(const float* __restrict__ tn, int Q, int D,
                                                                 const float* __restrict__ x, long long N,
                                                                 float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long n = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps) {
    float nn = 0.f;
    for (int k = lane; k < D; k += 32) { const float v = x[n * D + k]; nn = fmaf(v, v, nn); }
    const float inv = 1.f / fmaxf(sqrtf(warp_sum(nn)), kNormEps);
    for (int q = 0; q < Q; ++q) {
      float d = 0.f;
      for (int k = lane; k < D; k += 32) d = fmaf(x[n * D + k], tn[(size_t)q * D + k], d);
      d = warp_sum(d);
      if (lane == 0) scores[(size_t)q * N + n] = d * inv;
    }
  }
}

__device__ __forceinline__ unsigned long long topk_key(float s, long long n) {
  return ((unsigned long long)desc_bits(s) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)n);
}

// one block per query: exact k-th largest 64-bit key by 8 radix passes, then gather + bitonic sort of the k winners
__global__ void __launch_bounds__(1024) topk_kernel(const float* __restrict__ scores, long long N, int k, int kpad,
                                                    float* __restrict__ out_vals, int64_t* __restrict__ out_idx) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long prefix_s;
  __shared__ unsigned int remaining_s, count_s;
  extern __shared__ unsigned long long winners[];  // [kpad]
  const float* row = scores + (size_t)blockIdx.x * N;
  if (threadIdx.x == 0) { prefix_s = 0ull; remaining_s = (unsigned)k; count_s = 0u; }
  __syncthreads();
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const unsigned long long prefix = prefix_s;
    const unsigned long long himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
    for (long long n = threadIdx.x; n < N; n += blockDim.x) {
      const unsigned long long key = topk_key(row[n], n);
      if ((key & himask) == prefix) atomicAdd(&hist[(unsigned)(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int rem = remaining_s;
      int d = 255;
      for (; d > 0; --d) {
        if (hist[d] >= rem) break;
        rem -= hist[d];
      }
      remaining_s = rem;  // the k-th largest is the rem-th largest inside digit d
      prefix_s = prefix | ((unsigned long long)d << shift);
    }
    __syncthreads();
  }
  const unsigned long long thr = prefix_s;  // exactly k keys are >= thr
  for (int i = threadIdx.x; i < kpad; i += blockDim.x) winners[i] = 0ull;
  __syncthreads();
  for (long long n = threadIdx.x; n < N; n += blockDim.x) {
    const unsigned long long key = topk_key(row[n], n);
    if (key >= thr) {
      const unsigned int slot = atomicAdd(&count_s, 1u);
      if (slot < (unsigned)kpad) winners[slot] = key;
    }
  }
  __syncthreads();
  for (int kk = 2; kk <= kpad; kk <<= 1) {  // bitonic, descending
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (kpad >> 1); t += blockDim.x) {
        const int i = ((t / j) * 2 * j) + (t % j), l = i + j;
        const bool desc = ((i & kk) == 0);
        const unsigned long long a = winners[i], b = winners[l];
        if ((a < b) == desc) { winners[i] = b; winners[l] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const long long n = (long long)(0xFFFFFFFFu - (uint32_t)(winners[i] & 0xFFFFFFFFull));
    out_idx[(size_t)blockIdx.x * k + i] = n;
    out_vals[(size_t)blockIdx.x * k + i] = row[n];
  }
}

}  // namespace mc

using namespace mc;

extern "C" {

size_t mc_similarity_topk_workspace_bytes(int Q, long long N, int D) {
  if (Q <= 0 || N <= 0 || D <= 0) return 0;
  return round_up((size_t)Q * D * sizeof(float), 256) + (size_t)Q * (size_t)N * sizeof(float);
}

int mc_similarity_topk(const float* text, int Q, const float* image, long long N, int D, int k, float* out_vals,
                       int64_t* out_idx, float* scores_out, void* ws, size_t ws_bytes, void* stream) {
  MC_ARCH_GUARD();
  MC_REQUIRE(text && image && ws && (k == 0 || (out_vals && out_idx)), MC_ERR_BAD_ARG, "similarity_topk: null pointer");
  MC_REQUIRE(Q > 0 && N > 0 && D > 0 && k >= 0, MC_ERR_BAD_ARG, "similarity_topk: bad sizes Q=%d N=%lld D=%d k=%d", Q, N, D, k);
  MC_REQUIRE(k <= N, MC_ERR_BAD_ARG, "similarity_topk: k=%d exceeds the %lld candidates (torch.topk raises too)", k, N);
  MC_REQUIRE(k <= kMaxK, MC_ERR_UNSUPPORTED, "similarity_topk: k=%d > %d", k, kMaxK);
  MC_REQUIRE(N < 0xFFFFFFFFll, MC_ERR_UNSUPPORTED, "similarity_topk: more than 2^32 - 1 candidates");
  MC_REQUIRE(ws_bytes >= mc_similarity_topk_workspace_bytes(Q, N, D), MC_ERR_WORKSPACE, "similarity_topk: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* tn = static_cast<float*>(ws);
  float* scores = scores_out ? scores_out : reinterpret_cast<float*>(static_cast<char*>(ws) + round_up((size_t)Q * D * 4, 256));
  normalize_rows_kernel<<<(Q + 7) / 8, 256, 0, st>>>(text, Q, D, tn);
  MC_LAUNCH_CHECK();
  long long nb = (N + 7) / 8;
  const long long cap = (long long)num_sms() * 8;
  if (nb > cap) nb = cap;
  const bool vec = (D == 128 || D == 256 || D == 512) && aligned(image, 16);
  for (int q0 = 0; q0 < Q; q0 += kMaxQ) {
    const int qn = Q - q0 < kMaxQ ? Q - q0 : kMaxQ;
    const float* tq = tn + (size_t)q0 * D;
    float* sc = scores + (size_t)q0 * N;
    const size_t smem = (size_t)qn * D * sizeof(float);
    if (vec && D == 128) sim_scores_kernel<1><<<(int)nb, 256, smem, st>>>(tq, qn, image, N, sc);
    else if (vec && D == 256) sim_scores_kernel<2><<<(int)nb, 256, smem, st>>>(tq, qn, image, N, sc);
    else if (vec && D == 512) {
      static std::atomic<unsigned long long> done{0};
      MC_CUDA(ensure_dynamic_smem(sim_scores_kernel<4>, kMaxQ * 512 * 4, done));
      sim_scores_kernel<4><<<(int)nb, 256, smem, st>>>(tq, qn, image, N, sc);
    } else sim_scores_generic_kernel<<<(int)nb, 256, 0, st>>>(tq, qn, D, image, N, sc);
    MC_LAUNCH_CHECK();
  }
  if (k > 0) {
    int kpad = 2;
    while (kpad < k) kpad <<= 1;
    topk_kernel<<<Q, 1024, (size_t)kpad * 8, st>>>(scores, N, k, kpad, out_vals, out_idx);
    MC_LAUNCH_CHECK();
  }
  return MC_OK;
}

}  // extern "C"

// Generic strided fp32 GEMM on CUDA cores: the bring-up / cross-check engine
// (mode MC_GEMM_SIMT_FP32).  64x64x16 tiles, 256 threads, 4x4 register tile.
// True fp32 FMA arithmetic - the same precision class as the reference's
// CPU/cuBLAS sgemm (CLIP.py:34-36, modules.py:70,72) - so it doubles as the
// on-device check for the tcgen05 paths at sizes the CPU oracle is slow at.
#include "common.cuh"

namespace mc {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256) sgemm_kernel(SgemmArgs a) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
  float acc[4][4] = {};

  // pick the load mapping so that the unit-stride dimension runs across threads
  const bool a_k_fast = (a.sak == 1);
  const bool b_j_fast = (a.sbj == 1);

  for (int k0 = 0; k0 < a.K; k0 += TK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = tid + e * 256;  // 1024 elements of the 64x16 A tile
      int ii, kk;
      if (a_k_fast) { kk = idx % TK; ii = idx / TK; } else { ii = idx % TM; kk = idx / TM; }
      int gi = i0 + ii, gk = k0 + kk;
      As[kk][ii] = (gi < a.M && gk < a.K) ? a.A[gi * a.sai + gk * a.sak] : 0.f;
      int jj, kb;
      if (b_j_fast) { jj = idx % TN; kb = idx / TN; } else { kb = idx % TK; jj = idx / TK; }
      int gj = j0 + jj, gkb = k0 + kb;
      Bs[kb][jj] = (gj < a.N && gkb < a.K) ? a.B[gkb * a.sbk + gj * a.sbj] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) av[r] = As[kk][ty * 4 + r];
#pragma unroll
      for (int c = 0; c < 4; ++c) bv[c] = Bs[kk][tx * 4 + c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int gi = i0 + ty * 4 + r;
    if (gi >= a.M) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int gj = j0 + tx * 4 + c;
      if (gj >= a.N) continue;
      float v = a.alpha * acc[r][c];
      if (a.bias) v += a.bias[gj];
      float* dst = a.C + (int64_t)gi * a.ldc + gj;
      if (a.accumulate) v += *dst;
      *dst = v;
      if (a.gelu_out) a.gelu_out[(int64_t)gi * a.ldc + gj] = gelu_erf(v);
    }
  }
}

int sgemm(const SgemmArgs& a, cudaStream_t stream) {
  MC_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, MC_ERR_BAD_ARG, "sgemm: empty shape %d %d %d", a.M, a.N,
             a.K);
  dim3 grid((a.N + TN - 1) / TN, (a.M + TM - 1) / TM);
  sgemm_kernel<<<grid, 256, 0, stream>>>(a);
  MC_LAUNCH_CHECK();
  return MC_OK;
}

}  // namespace mc

// fp32-class GEMM on the 5th-gen tensor cores for the ProjectionHead (modules.py:63-72 and their
// autograd): C[M,N] = A[M,K] . B[N,K]^T with both operands staged as K-major fp16 hi/lo planes
// (x = (hi + lo) / s, s a per-matrix power of two) and three tcgen05 passes hi*hi + hi*lo + lo*hi.
#pragma once
#include "common.cuh"

#include <cuda_fp16.h>

namespace mc {
namespace tcg {

struct Planes {          // a staged operand: rows x cols fp16, row pitch `pitch` elements (multiple of 8)
  __half* hi;
  __half* lo;
  float* scale;          // device: {s, 1/s}
  int rows, cols, pitch;
};

size_t planes_bytes(int rows, int cols);  // hi + lo + scale slot, 256-byte aligned pieces

// carve `mem` (>= planes_bytes) into a Planes descriptor
Planes carve_planes(void* mem, int rows, int cols);

// dst = split(src * s): src is (R, C) fp32 with row stride lds; transpose = 1 stages src^T (C rows, R cols).
// known_amax: device word with the bit pattern of max |src| when something already reduced it (the kernel that
// produced src, or an earlier staging of the same tensor); null = reduce it here (one more pass over src).
int stage(const float* src, int R, int C, int64_t lds, int transpose, const Planes& dst, cudaStream_t st,
          const unsigned int* known_amax = nullptr);
// bit pattern of max |src| over an (R, C) fp32 matrix with row stride lds -> *slot (zeroed first)
int amax(const float* src, int R, int C, int64_t lds, unsigned int* slot, cudaStream_t st);
inline const unsigned int* amax_slot(const Planes& p) { return reinterpret_cast<const unsigned int*>(p.scale) + 2; }

// Two small matrices (the weights of a head) staged in ONE launch: max |.| of each, a grid-wide barrier, then the
// hi / lo split (transposed or not).  `sync` = 4 device words, zeroed by the call.  Replaces 2 x (memset + amax + stage).
struct StageJob {
  const float* src;   // (R, C) dense fp32
  int R, C, transpose;
  Planes dst;         // (R, C) planes, or (C, R) when transpose
};
// known: two device words with the maxima of a / b when an earlier call already reduced them (then no barrier, no
// memset); amax_out: optional two words that receive them.
int stage_pair(const StageJob& a, const StageJob& b, unsigned int* sync, cudaStream_t st,
               const unsigned int* known = nullptr, unsigned int* amax_out = nullptr);

enum Epilogue { kEpiPlain = 0, kEpiGelu = 1 };

struct GemmOut {
  float* C;              // (M, N) fp32, row stride ldc
  int64_t ldc;
  const float* bias;     // length N or null
  float* gelu_out;       // kEpiGelu: gelu(C) with the same layout, or null
};

// ksplit > 1: partial products go to `ws` (ksplit x M x N fp32) and a second kernel reduces them
size_t gemm_workspace_bytes(int M, int N, int K);
int gemm(const Planes& A, const Planes& B, int M, int N, int K, const GemmOut& out, int epilogue, void* ws,
         size_t ws_bytes, cudaStream_t st);

}  // namespace tcg
}  // namespace mc
